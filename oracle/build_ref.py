"""Build `oracle/_ref/`: the REAL reference, compiled, for use as a checker on the GPU box.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python, so "compiling it from the sources where they lie" means byte-compiling
`/root/reference/{core_system,ui,video_processing}.py` with this image's interpreter into `oracle/_ref/*.rvoc`.  Nothing is
copied into the repository: `oracle/_ref/` is git-ignored (it is NOT gpurun-ignored, so the bytecode travels to the GPU box like
the built .so).  tests/test_reference_source.py loads these modules UNMODIFIED — with stand-ins only for the third-party packages
that are not installed (qdrant_client -> B200VectorDB, gradio, supervision, autodistill*, perception_models) — and drives the
reference's own UI callbacks against the CUDA library.

Run by `__graft_entry__.build()` whenever /root/reference is present (the authoring container); a no-op elsewhere.
"""
from __future__ import annotations

import os
import py_compile

REFERENCE_DIR = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ("core_system", "ui", "video_processing")


def build() -> list[str]:
    if not os.path.exists(os.path.join(REFERENCE_DIR, "core_system.py")):
        return []
    os.makedirs(OUT, exist_ok=True)
    made = []
    for m in MODULES:
        src, dst = os.path.join(REFERENCE_DIR, m + ".py"), os.path.join(OUT, m + ".rvoc")
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        made.append(dst)
    return made


if __name__ == "__main__":
    print("\n".join(build()) or "no /root/reference here: nothing to build")
