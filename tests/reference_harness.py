"""Harness that runs the reference's OWN, UNMODIFIED code (core_system.py + ui.py) through the drop-in boundary.

The reference cannot be imported as is in this image: qdrant_client, gradio, supervision, autodistill*, perception_models,
yt_dlp and scenedetect are not installed.  This module installs small stand-ins for exactly those
third-party modules in `sys.modules` and then loads the reference modules without touching a byte of them:

  qdrant_client / qdrant_client.http.models   -> a recording wrapper around a backend of the caller's choice:
                                                 `oracle.QdrantLocalOracle` (CPU; generates tests/golden/reference_trace.*)
                                                 or `revers_o_b200.vector_db.B200VectorDB` (GPU; the product)
  core.vision_encoder.pe / .transforms        -> a deterministic fake PE encoder ([1,577,1024] tokens) and transform
  autodistill_grounded_sam / autodistill.detection / supervision.detection.core
                                              -> a deterministic fake detector (4 masks, one empty)
  gradio, yt_dlp, scenedetect                 -> inert shims (ui.py only needs gr.update / gr.Progress at import time)

Where the reference comes from:
  * `/root/reference/*.py`                    in the authoring container (source, loaded in place, never copied);
  * `oracle/_ref/*.rvoc`                       on the GPU box: bytecode compiled from those sources by `oracle/build_ref.py`
                                              (`__graft_entry__.build()`), git-ignored, travels with the gpurun snapshot.
TEST INFRASTRUCTURE ONLY: nothing under revers_o_b200/ imports this.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types
from types import SimpleNamespace as NS

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_DIR = "/root/reference"
REF_PYC_DIR = os.path.join(ROOT, "oracle", "_ref")
STUBBED = ["qdrant_client", "qdrant_client.http", "qdrant_client.http.models", "core", "core.vision_encoder",
           "core.vision_encoder.pe", "core.vision_encoder.transforms", "autodistill_grounded_sam", "autodistill",
           "autodistill.detection", "supervision", "supervision.detection", "supervision.detection.core", "gradio", "yt_dlp",
           "scenedetect", "core_system", "ui", "video_processing"]


def reference_location():
    """('source', dir) | ('pyc', dir) | (None, None)."""
    if os.path.exists(os.path.join(REFERENCE_DIR, "core_system.py")):
        return "source", REFERENCE_DIR
    if os.path.exists(os.path.join(REF_PYC_DIR, "core_system.rvoc")):
        return "pyc", REF_PYC_DIR
    return None, None


# ---- deterministic stand-ins for the third-party models ------------------------------------------------------------
def make_image(seed, size=(96, 80)):
    from PIL import Image
    rs = np.random.RandomState(seed)
    coarse = Image.fromarray((rs.rand(5, 6, 3) * 255).astype(np.uint8))       # low-frequency content: images differ globally
    img = np.asarray(coarse.resize(size, Image.BILINEAR), dtype=np.float32)
    img += rs.rand(size[1], size[0], 3) * 20
    return Image.fromarray(np.clip(img, 0, 255).astype(np.uint8))


class FakePE:
    """Stand-in for pe.CLIP: encode_image([1,3,336,336]) -> [1, 1 + 24*24, 1024] tokens, deterministic, any device."""
    created = []

    def __init__(self, name):
        import torch
        self.name, self.device, self.is_half = name, torch.device("cpu"), False
        g = torch.Generator().manual_seed(0)
        self.w = torch.randn((3, 1024), generator=g)
        FakePE.created.append(self)

    @staticmethod
    def available_configs():
        return ["PE-Core-B16-224", "PE-Core-L14-336", "PE-Core-G14-448"]

    @classmethod
    def from_config(cls, name, pretrained=True):
        return cls(name)

    def to(self, device):
        import torch
        self.device = torch.device(device)
        self.w = self.w.to(self.device)
        return self

    def half(self):
        self.is_half = True           # core_system.py:195-196: fp16 tokens out on CUDA
        return self

    def encode_image(self, x):
        import torch
        p = torch.nn.functional.adaptive_avg_pool2d(x.float(), 24).flatten(2).transpose(1, 2)   # [1,576,3]
        t = torch.tanh(((p - 0.5) * 4.0) @ self.w + 0.1 * torch.linspace(-1, 1, 1024, device=x.device))
        out = torch.cat([t.mean(1, keepdim=True), t], 1)
        return out.half() if self.is_half else out


def fake_transform(size):
    def preprocess(pil):
        import torch
        a = np.asarray(pil.resize((size, size)), dtype=np.float32) / 255.0
        return torch.from_numpy(a).permute(2, 0, 1)
    return preprocess


class FakeDetections:
    """supervision.detection.core.Detections: the four attributes core_system.py reads, and len()."""
    def __init__(self, xyxy, mask=None, confidence=None, class_id=None):
        self.xyxy, self.mask, self.confidence, self.class_id = xyxy, mask, confidence, class_id

    def __len__(self):
        return len(self.xyxy)


class FakeOntology:
    def __init__(self, mapping):
        self.mapping = dict(mapping)

    def classes(self):
        return list(self.mapping.values())


class FakeGroundedSAM:
    """autodistill_grounded_sam.GroundedSAM: predict(path) -> detections with 4 masks (one empty, core_system.py:402-404)."""
    instances = 0

    def __init__(self, ontology, box_threshold=0.35, text_threshold=0.25):
        self.ontology, self.box_threshold, self.text_threshold = ontology, box_threshold, text_threshold
        FakeGroundedSAM.instances += 1

    def predict(self, path):
        from PIL import Image
        with Image.open(path) as im:
            W, H = im.size
        m = np.zeros((4, H, W), bool)
        m[0, : H // 2, : W // 2] = True
        m[1, H // 4:, W // 3:] = True
        m[3, H // 2:, :] = True
        n_cls = max(1, len(self.ontology.classes()))
        return FakeDetections(xyxy=np.array([[0, 0, W // 2, H // 2], [W // 3, H // 4, W, H], [0, 0, 1, 1], [0, H // 2, W, H]], float),
                              mask=m, confidence=np.array([0.9, 0.8, 0.7, 0.6]), class_id=np.arange(4) % n_cls)


# ---- qdrant_client stand-in: records every call the reference makes -------------------------------------------------
class Recorder:
    """What the unmodified reference sent to / got from the vector DB, in call order."""
    def __init__(self):
        self.calls = []

    def reset(self):
        self.calls = []


def make_qdrant_module(backend_factory, recorder: Recorder):
    """`from qdrant_client import QdrantClient` / `from qdrant_client.http import models` for core_system.py:20-21."""
    models = NS(
        Distance=NS(COSINE="Cosine"),
        VectorParams=lambda size, distance="Cosine": NS(size=size, distance=distance),
        PointStruct=lambda id, vector, payload=None: NS(id=id, vector=vector, payload=payload),
    )

    class QdrantClient:
        def __init__(self, path=None, **kw):
            self.path = path
            self._b = backend_factory(path)
            recorder.calls.append({"op": "open", "path": path})

        def get_collections(self):
            r = self._b.get_collections()
            recorder.calls.append({"op": "get_collections", "names": [c.name for c in r.collections]})
            return r

        def recreate_collection(self, collection_name, vectors_config=None, **kw):
            recorder.calls.append({"op": "recreate_collection", "name": collection_name, "size": int(vectors_config.size),
                                   "distance": str(vectors_config.distance)})
            return self._b.recreate_collection(collection_name=collection_name, vectors_config=vectors_config, **kw)

        def upsert(self, collection_name, points):
            points = list(points)
            recorder.calls.append({"op": "upsert", "name": collection_name, "ids": [p.id for p in points],
                                   "vectors": np.asarray([p.vector for p in points], dtype=np.float32),
                                   "payloads": [p.payload for p in points],
                                   "vector_type": type(points[0].vector).__name__ if points else None})
            return self._b.upsert(collection_name=collection_name, points=points)

        def search(self, collection_name, query_vector, limit=10, score_threshold=None, **kw):
            hits = self._b.search(collection_name=collection_name, query_vector=query_vector, limit=limit,
                                  score_threshold=score_threshold, **kw)
            recorder.calls.append({"op": "search", "name": collection_name, "query": np.asarray(query_vector, dtype=np.float32),
                                   "query_type": type(query_vector).__name__, "limit": limit, "score_threshold": score_threshold,
                                   "hits": [{"id": h.id, "score": float(h.score), "payload": h.payload} for h in hits]})
            return hits

    qc = types.ModuleType("qdrant_client")
    qc.QdrantClient = QdrantClient
    http = types.ModuleType("qdrant_client.http")
    mm = types.ModuleType("qdrant_client.http.models")
    for k, v in vars(models).items():
        setattr(mm, k, v)
    http.models = mm
    qc.http = http
    return {"qdrant_client": qc, "qdrant_client.http": http, "qdrant_client.http.models": mm}


def oracle_backend_factory():
    """QdrantLocalOracle with qdrant-local's persistence emulated per path (a second client on the same directory sees the
    collections, core_system.py:100)."""
    from oracle import reverso_oracle as O
    store = {}

    def factory(path):
        key = os.path.abspath(path) if path else None
        if key not in store:
            store[key] = O.QdrantLocalOracle(path)
            if path:
                os.makedirs(path, exist_ok=True)       # qdrant-local creates the directory (list_databases sees it)
        return store[key]
    return factory


def b200_backend_factory(**kw):
    def factory(path):
        from revers_o_b200.vector_db import B200VectorDB
        return B200VectorDB(path=path, **kw)
    return factory


def _gradio_module():
    gr = types.ModuleType("gradio")

    class Progress:
        def __init__(self, *a, **k):
            self.events = []

        def __call__(self, value, desc=None, **k):
            self.events.append((value, desc))

        def tqdm(self, it, *a, **k):
            return it

    class _Ctx:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def __getattr__(self, name):
            return lambda *a, **k: None

    gr.Progress = Progress
    gr.update = lambda **kw: dict(kw, __type__="update")
    for name in ("Blocks", "Tab", "Tabs", "Row", "Column", "Group", "Accordion"):
        setattr(gr, name, _Ctx)
    for name in ("State", "Markdown", "Image", "Textbox", "Checkbox", "Button", "Dropdown", "Slider", "Gallery", "Radio", "File",
                 "Number", "HTML"):
        setattr(gr, name, _Ctx)
    return gr


def install_stubs(backend_factory, recorder: Recorder):
    saved = {k: sys.modules.get(k) for k in STUBBED}
    mods = make_qdrant_module(backend_factory, recorder)
    core = types.ModuleType("core")
    ve = types.ModuleType("core.vision_encoder")
    pe = types.ModuleType("core.vision_encoder.pe")
    pe.CLIP = FakePE
    tr = types.ModuleType("core.vision_encoder.transforms")
    tr.get_image_transform = fake_transform
    core.vision_encoder, ve.pe, ve.transforms = ve, pe, tr
    ags = types.ModuleType("autodistill_grounded_sam")
    ags.GroundedSAM = FakeGroundedSAM
    ad = types.ModuleType("autodistill")
    add = types.ModuleType("autodistill.detection")
    add.CaptionOntology = FakeOntology
    ad.detection = add
    sv = types.ModuleType("supervision")
    svd = types.ModuleType("supervision.detection")
    svc = types.ModuleType("supervision.detection.core")
    svc.Detections = FakeDetections
    sv.detection, svd.core = svd, svc
    ytd = types.ModuleType("yt_dlp")
    sd = types.ModuleType("scenedetect")
    sd.open_video = sd.SceneManager = sd.ContentDetector = lambda *a, **k: None
    mods.update({"core": core, "core.vision_encoder": ve, "core.vision_encoder.pe": pe, "core.vision_encoder.transforms": tr,
                 "autodistill_grounded_sam": ags, "autodistill": ad, "autodistill.detection": add, "supervision": sv,
                 "supervision.detection": svd, "supervision.detection.core": svc, "gradio": _gradio_module(), "yt_dlp": ytd,
                 "scenedetect": sd})
    for k in ("core_system", "ui", "video_processing"):
        sys.modules.pop(k, None)
    sys.modules.update(mods)
    return saved


def restore_modules(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _load(name, kind, where):
    if kind == "source":
        spec = importlib.util.spec_from_file_location(name, os.path.join(where, name + ".py"))
    else:
        path = os.path.join(where, name + ".rvoc")
        spec = importlib.util.spec_from_loader(name, importlib.machinery.SourcelessFileLoader(name, path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference(kind, where):
    """Import the unmodified reference modules (ui.py instantiates `reverso = SimpleReverso()` at import, ui.py:20)."""
    _load("core_system", kind, where)
    _load("video_processing", kind, where)
    return _load("ui", kind, where)


# ---- the scenario: the UI callbacks a user triggers, in order (ui.py:29-159, 200-214) ---------------------------------------
N_IMAGES = 8


def write_images(folder):
    os.makedirs(folder, exist_ok=True)
    for i in range(N_IMAGES):
        make_image(i).save(os.path.join(folder, f"img{i}.png"))


def run_scenario(ui):
    """Drives the reference's UI callbacks; returns their outputs.  Relative paths only, so that the texts do not depend on
    where the scenario ran (cwd is a scratch directory: core_system.py uses ./simple_reverso_db)."""
    out = {}
    write_images("imgs")
    out["build_regions"] = ui.build_database_ui("imgs", "trace", "object .", False, False, False)
    out["build_direct"] = ui.build_database_ui("imgs", "direct", "", True, False, False)
    out["list"] = ui.list_available_databases_ui()
    out["load_missing"] = ui.load_selected_database_ui("nope")
    out["load"] = ui.load_selected_database_ui("trace")
    q = make_image(2)
    viz, text, upd, meta = ui.detect_and_extract_ui(q, "object .", False)
    out["detect_text"], out["detect_choices"] = text, upd.get("choices")
    out["detect_meta"] = [{k: v for k, v in m.items() if k != "region_id"} for m in meta]
    searches = []
    for thr, k, sel in ((0.5, 5, None), (0.5, 5, "Region 2: object (Conf: 0.80)"), (0.0, 20, "Region 3: object (Conf: 0.60)"),
                        (0.9999, 3, None), (1.01, 5, None)):
        text, gallery, items = ui.search_database_ui(thr, k, [], sel)
        searches.append({"threshold": thr, "max_results": k, "selected": sel, "text": text,
                         "items": [{"score": it["score"], "filename": it["filename"], "bbox": it["bbox"]} for it in (items or [])]})
    out["load_direct"] = ui.load_selected_database_ui("direct")
    viz, text, upd, meta = ui.detect_and_extract_ui(make_image(5), "", True)
    out["direct_text"] = text
    text, gallery, items = ui.search_database_ui(0.3, 10, [], None)
    searches.append({"threshold": 0.3, "max_results": 10, "selected": None, "text": text,
                     "items": [{"score": it["score"], "filename": it["filename"], "bbox": it["bbox"]} for it in (items or [])]})
    out["searches"] = searches
    out["unlock"] = ui.unlock_selected_database_ui("trace")
    out["delete"] = ui.delete_selected_database_ui("direct")
    out["list_after"] = ui.list_available_databases_ui()
    return out
