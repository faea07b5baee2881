"""handsonvlm-release_b200 -- B200-native (sm_100a) visual-token path for HandsOnVLM.

Only what the hot path needs: ``csrc/`` (CUDA kernels + the C ABI of
include/hvlm_b200.h), ``ops`` (torch.library custom ops over that ABI) and the host-side mirror of the
reference interface (``tower``, ``arch``).  The directory name is not a Python identifier; import it as
``import hvlm_b200`` (alias module at the repo root) or ``importlib.import_module("handsonvlm-release_b200")``.
"""
from . import _lib
from .constants import HAND_TRAJ_TOKEN_ID, IGNORE_INDEX, IMAGE_TOKEN_INDEX  # noqa: F401


def build(verbose: bool = False) -> str:
    return _lib.build(verbose)


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `build()` works before torch is paged in
    if name in ("ops", "tower", "arch", "weights", "dist", "traj_decoder", "builder", "integrate"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("CLIPVisionTower",):
        from .tower import CLIPVisionTower
        return CLIPVisionTower
    if name in ("VisualToTokenHelper", "LlavaMetaForCausalLM", "LitaMetaForCausalLM", "HandsOnVLMMetaForCausalLM",
                "gather_hand_traj_states"):
        from . import arch
        return getattr(arch, name)
    if name in ("patch_reference", "unpatch_reference"):
        from . import integrate
        return getattr(integrate, name)
    if name == "build_vision_tower":
        from .builder import build_vision_tower
        return build_vision_tower
    if name in ("CVAETrajDecoder", "MLPTrajDecoder"):
        from . import traj_decoder
        return getattr(traj_decoder, name)
    raise AttributeError(name)
