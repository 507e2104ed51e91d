"""Boundary parity of the error/status-string convention: every user-facing status line the reference's
`SimpleReverso` produces in the methods this repo mirrors must exist, character for character (f-string holes aside), in
revers_o_b200/core_system.py.  Reads the reference source, so it only runs where /root/reference is mounted (this
container); checkpoint messages are excluded (dead code in the reference)."""
import ast
import os

import pytest

REF = "/root/reference/core_system.py"
OURS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "revers_o_b200", "core_system.py")
METHODS = ["load_database", "delete_database", "unlock_database", "create_database", "search_similar"]
PREFIXES = ("❌", "✅", "⚠️", "ℹ️", "🎯", "📦", "💾", "🔄", "📊", "📁", "📂", "🔧", "🛑", "🔍", "\n📊", "\n\n⏸️")
SKIP = ("checkpoint", "Resuming", "already processed", "Error loading/processing image", "Image not found")


def _templates(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    out = {}
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SimpleReverso"]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]:
            found = set()
            for node in ast.walk(fn):
                if isinstance(node, ast.JoinedStr):
                    text = "".join(v.value if isinstance(v, ast.Constant) else "{}" for v in node.values)
                elif isinstance(node, ast.Constant) and isinstance(node.value, str):
                    text = node.value
                else:
                    continue
                if text.startswith(PREFIXES) and not any(w in text for w in SKIP):
                    found.add(text)
            out[fn.name] = found
    return out


@pytest.mark.skipif(not os.path.exists(REF), reason="reference source not mounted")
def test_every_reference_status_line_exists_in_the_drop_in():
    ref, ours = _templates(REF), _templates(OURS)
    missing = {m: sorted(ref[m] - ours.get(m, set())) for m in METHODS if ref.get(m, set()) - ours.get(m, set())}
    assert not missing, missing


@pytest.mark.skipif(not os.path.exists(REF), reason="reference source not mounted")
def test_mirrored_method_signatures_match_the_reference():
    def sigs(path):
        tree = ast.parse(open(path, encoding="utf-8").read())
        cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SimpleReverso"][0]
        out = {}
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            a = fn.args
            defaults = [ast.unparse(d).replace("'", '"') for d in a.defaults]
            out[fn.name] = ([x.arg for x in a.args], defaults)
        return out
    ref, ours = sigs(REF), sigs(OURS)
    for m in ["load_database", "delete_database", "unlock_database", "list_databases", "detect_regions", "extract_embeddings",
              "process_image_direct_pe", "request_stop", "create_database", "search_similar", "visualize_detections"]:
        assert m in ours, m
        assert ours[m] == ref[m], (m, ours[m], ref[m])
