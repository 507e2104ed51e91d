"""CPU oracle (test infrastructure).  See reverso_oracle.py — only tests/, smoke() and bench.py's CPU legs may import this."""
