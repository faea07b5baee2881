"""One call that switches an imported reference checkout to the B200 path.

``patch_reference()`` swaps exactly the symbols INTEGRATION.md section 3 lists, in the reference modules that are already
imported (``sys.modules``): the vision-tower class and builder, the three mixins' hot-path methods, ``VisualToTokenHelper``
and ``HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal``.  State dicts, configs, trainers, the LLM and every other
reference file stay untouched; ``unpatch_reference()`` restores the originals.  (The inline ``<hand_traj>`` gather loop of
``HandsOnVLMForCausalLM.forward``, handsonvlm.py:146-187, is code inside a method and cannot be swapped by assignment: the
model class gets ``gather_hand_traj_states`` added, and INTEGRATION.md shows the three-line edit that calls it.)
"""
from __future__ import annotations

import sys

from . import arch
from .builder import build_vision_tower
from .tower import CLIPVisionTower

_saved: list = []


def _swap(obj, name, new):
    _saved.append((obj, name, getattr(obj, name, None), hasattr(obj, name)))
    setattr(obj, name, new)


def patch_reference() -> list:
    """-> list of 'module.symbol' strings that were swapped (only modules that are imported are touched)."""
    done = []
    m = sys.modules.get

    def swap(modname, path, new):
        mod = m(modname)
        if mod is None:
            return
        obj = mod
        parts = path.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p, None)
            if obj is None:
                return
        if len(parts) > 1 or hasattr(mod, parts[0]):
            _swap(obj, parts[-1], new)
            done.append(f"{modname}.{path}")

    # llava/model/multimodal_encoder/{clip_encoder,builder}.py and the names llava_arch imported from them
    swap("llava.model.multimodal_encoder.clip_encoder", "CLIPVisionTower", CLIPVisionTower)
    swap("llava.model.multimodal_encoder.builder", "CLIPVisionTower", CLIPVisionTower)
    swap("llava.model.multimodal_encoder.builder", "build_vision_tower", build_vision_tower)
    swap("llava.model.llava_arch", "build_vision_tower", build_vision_tower)
    # llava/model/llava_arch.py:73-234
    for meth in ("encode_images", "images_to_tokens", "visual_to_tokens", "prepare_inputs_labels_for_multimodal"):
        swap("llava.model.llava_arch", f"LlavaMetaForCausalLM.{meth}", getattr(arch.LlavaMetaForCausalLM, meth))
    # lita/model/lita_arch.py:17-85
    for meth in ("videos_to_tokens", "visual_to_tokens"):
        swap("lita.model.lita_arch", f"LitaMetaForCausalLM.{meth}", getattr(arch.LitaMetaForCausalLM, meth))
    # hoi_forecast/model/visual_to_tokens.py and the name handsonvlm.py:18 imported from it
    swap("hoi_forecast.model.visual_to_tokens", "VisualToTokenHelper", arch.VisualToTokenHelper)
    hv = "handsonvlm.model.language_model.handsonvlm"
    swap(hv, "VisualToTokenHelper", arch.VisualToTokenHelper)
    # handsonvlm/model/language_model/handsonvlm.py:212-451 (+ the gather as a method for the edit of :146-187)
    swap(hv, "HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal",
         arch.HandsOnVLMMetaForCausalLM.prepare_inputs_labels_for_multimodal)
    swap(hv, "HandsOnVLMForCausalLM.gather_hand_traj_states", arch.HandsOnVLMMetaForCausalLM.gather_hand_traj_states)
    swap(hv, "HandsOnVLMForCausalLM.clear_visual_token_cache", arch.HandsOnVLMMetaForCausalLM.clear_visual_token_cache)
    return done


def unpatch_reference() -> None:
    while _saved:
        obj, name, old, had = _saved.pop()
        if had:
            setattr(obj, name, old)
        else:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
