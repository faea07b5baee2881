"""TEST INFRASTRUCTURE ONLY -- import the *real* reference hot-path modules (dev container only).

``/root/reference`` is a read-only checkout of Kami-code/HandsOnVLM-release.  It cannot be
imported normally on this image: ``llava/__init__`` drags in a vendored MPT
that needs symbols removed from transformers 5.x, and ``handsonvlm.model`` imports
``deepspeed``/``wandb``.  The shim registers empty package stubs so the ``__init__`` files never
run, stubs the absent third-party modules, and then imports exactly the modules on the path.

Nothing here ships to the GPU box: ``make_golden.py`` uses it to freeze fixtures; the tests
that call it are skipped when ``/root/reference`` does not exist.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REF = os.environ.get("HVLM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "handsonvlm"))


_loaded = None


def load():
    """Returns a namespace with the reference classes on the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF}")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def stub(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, rel)]
        sys.modules[name] = m

    for n, p in [("llava", "llava"), ("llava.model", "llava/model"), ("lita", "lita"),
                 ("lita.model", "lita/model"), ("handsonvlm", "handsonvlm"),
                 ("handsonvlm.model", "handsonvlm/model")]:
        if n not in sys.modules:
            stub(n, p)

    if "deepspeed" not in sys.modules:
        ds = types.ModuleType("deepspeed")
        ds.comm = types.SimpleNamespace(barrier=lambda: None)
        ds.zero = types.SimpleNamespace(Init=lambda *a, **k: contextlib.nullcontext())
        sys.modules["deepspeed"] = ds
    if "wandb" not in sys.modules:
        wb = types.ModuleType("wandb")
        wb.run = None
        wb.log = lambda *a, **k: None
        sys.modules["wandb"] = wb

    import transformers.generation as tg
    import transformers.generation.utils as gu
    for n in ("SampleOutput", "SampleEncoderDecoderOutput"):
        if not hasattr(gu, n):
            setattr(gu, n, object)
    if not hasattr(tg, "validate_stopping_criteria"):
        tg.validate_stopping_criteria = lambda sc, ml: sc

    ns = types.SimpleNamespace()
    ns.clip_encoder = importlib.import_module("llava.model.multimodal_encoder.clip_encoder")
    ns.llava_arch = importlib.import_module("llava.model.llava_arch")
    ns.lita_arch = importlib.import_module("lita.model.lita_arch")
    ns.v2t = importlib.import_module("hoi_forecast.model.visual_to_tokens")
    try:
        ns.handsonvlm = importlib.import_module("handsonvlm.model.language_model.handsonvlm")
    except Exception as e:  # pragma: no cover - depends on installed transformers
        ns.handsonvlm = None
        ns.handsonvlm_error = e
    _loaded = ns
    return ns


def build_tower(ns, hf_model, select_layer=-2, select_feature="patch"):
    """A reference ``CLIPVisionTower`` around an already-built HF CLIPVisionModel (the ctor calls
    ``from_pretrained`` which needs the network)."""
    import torch.nn as nn
    T = ns.clip_encoder.CLIPVisionTower
    tower = T.__new__(T)
    nn.Module.__init__(tower)
    tower.is_loaded = True
    tower.select_layer = select_layer
    tower.select_feature = select_feature
    tower.vision_tower = hf_model
    tower.vision_tower.requires_grad_(False)
    return tower


class FakeHost:
    """Light host object exposing what the mixin methods reach through ``self``:
    get_model()/get_vision_tower()/config/token_dim/B/device."""

    def __init__(self, tower, projector, embed, config, B):
        import torch
        self._inner = types.SimpleNamespace(vision_tower=tower, mm_projector=projector, embed_tokens=embed,
                                            get_vision_tower=lambda: tower)
        self.config = config
        self.token_dim = projector.out_features
        self.B = B
        self.device = torch.device("cpu")

    def get_model(self):
        return self._inner

    def get_vision_tower(self):
        return self._inner.vision_tower
