"""Generates tests/golden/reference_trace.json + reference_trace.npz: what the reference's OWN code does on the path.

Run in the authoring container (needs /root/reference):   python tests/golden/make_reference_trace.py

The UNMODIFIED /root/reference/core_system.py and ui.py are loaded with stand-ins for the third-party packages that are not
installed (tests/reference_harness.py); `qdrant_client` is the oracle's qdrant-local restatement wrapped in a recorder.  The
script drives the reference's UI callbacks (build_database_ui -> load_selected_database_ui -> detect_and_extract_ui ->
search_database_ui, ui.py:29-159,207-214) and records
  * every call the reference made to the vector DB (recreate_collection / upsert with ids, vectors, payloads / search with the
    query, limit, threshold) and what came back,
  * every text / hit list the UI callbacks returned.
The fixtures pin everything on the path that is NOT qdrant's arithmetic to the real reference: the embeddings it computes
(e / e.norm(), core_system.py:407,447), the empty-mask skip, payload keys, upsert batching, status strings, hit formatting.
tests/test_reference_source.py replays them against the CUDA library on the GPU box.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def generate():
    import reference_harness as H
    kind, where = H.reference_location()
    assert kind == "source", "needs /root/reference"
    rec = H.Recorder()
    saved = H.install_stubs(H.oracle_backend_factory(), rec)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="rvo_trace_"))
    try:
        ui = H.load_reference(kind, where)
        out = H.run_scenario(ui)
    finally:
        os.chdir(cwd)
        H.restore_modules(saved)
    arrays, calls = {}, []
    for i, c in enumerate(rec.calls):
        c = dict(c)
        for key in ("vectors", "query"):
            if key in c:
                arrays[f"{key}_{i}"] = c.pop(key)
                c[key] = f"{key}_{i}"
        calls.append(c)
    return {"ui": out, "calls": calls}, arrays


if __name__ == "__main__":
    trace, arrays = generate()
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "reference_trace.json"), "w") as f:
        json.dump(trace, f, indent=1, ensure_ascii=False, default=float)
    np.savez_compressed(os.path.join(here, "reference_trace.npz"), **arrays)
    n_up = sum(len(c["ids"]) for c in trace["calls"] if c["op"] == "upsert")
    print(f"{len(trace['calls'])} vector-DB calls ({n_up} points upserted, "
          f"{sum(c['op'] == 'search' for c in trace['calls'])} searches), {len(trace['ui']['searches'])} UI searches")
