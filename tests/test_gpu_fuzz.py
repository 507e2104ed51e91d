"""Differential fuzz of K2 on the GPU (scripts/dev/fuzz_search.py): random shapes, thresholds, tie structures, adjacent near-copies
and forced paths (fp32 scan, tensor, dense) against an fp32 restatement, scores to 1e-5 and id sets equal except among such ties."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [3, 11])
def test_fuzz_search_against_fp32_restatement(seed):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev", "fuzz_search.py"), "150", str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "150 cases, 0 mismatches" in r.stdout


@pytest.mark.parametrize("seed", [5])
def test_fuzz_mask_pool_against_fp32_restatement(seed):
    """K1 with and without the fused ingest (scripts/dev/fuzz_maskpool.py): widths on and off the tensor path, bf16 / fp16, class
    token, region caps, empty regions and images, any non-zero mask byte — counts, source indices and rows to 2e-5."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev", "fuzz_maskpool.py"), "150", str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "150 cases, 0 mismatches" in r.stdout


@pytest.mark.parametrize("seed,shards", [(1, 1), (2, 3)])
def test_fuzz_vector_db_against_numpy_model(seed, shards):
    """Model-based fuzz of the drop-in boundary (scripts/dev/fuzz_vector_db.py): random upserts (new ids, overwrites, duplicates
    inside a batch, int / uuid ids and the switch between them), bulk ingest, searches, collection re-creation and close + reopen
    from disk, checked against a numpy model after every step; `shards` = 3 runs the row-sharded collection code (on one GPU the
    three shards share the device)."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev", "fuzz_vector_db.py"), "300", str(seed), str(shards)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "300 steps, 0 mismatches" in r.stdout


def test_fuzz_selfjoin_against_bruteforce():
    """configs[4] path (scripts/dev/fuzz_selfjoin.py): sizes around the 4096-row query block, thresholds, planted near-copies,
    static-scene cliques, row ranges, id offsets, and small candidate / output capacities that force the exact fallbacks —
    every pair with cos >= threshold (fp32, 1e-4 band at the threshold) exactly once, i < j, scores to 1e-4."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev", "fuzz_selfjoin.py"), "60", "2"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "60 cases, 0 mismatches" in r.stdout


def test_fuzz_merge_against_numpy():
    """K3 (scripts/dev/fuzz_merge.py): up to 64 lists, short / empty lists, cross-shard score ties (lower id first), G * k at
    the 4096 limit (48 KB of dynamic shared memory), a shard's overflow flag must survive the merge — bit-exact."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev", "fuzz_merge.py"), "100", "4"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "100 cases, 0 mismatches" in r.stdout
