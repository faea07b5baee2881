"""ctypes binding of libhvlm_b200.so -- the C ABI declared in include/hvlm_b200.h.

The product path has NO fallback: if the shared library is missing or the device is not an sm_100 GPU the
import of the compute ops fails loudly (RuntimeError), it never routes to torch eager or to any CPU restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhvlm_b200.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "hvlm_b200.h")

# enums (include/hvlm_b200.h)
F32, BF16, F16 = 0, 1, 2
POOL_MODES = {"temporal_spatial_pool": 0, "spatial_pool": 1, "temporal": 2, "spatial": 3, "temporal_spatial": 4}
SPLICE_LLAVA, SPLICE_HANDSONVLM = 0, 1
SPLICE_FLAG_IM_START_END = 0x100   # OR-able into the splice variant (include/hvlm_b200.h)
EPI_BIAS, EPI_BIAS_QUICKGELU, EPI_BIAS_RESIDUAL = 0, 1, 2
PLAN_ERR_LEN_OVERFLOW, PLAN_ERR_IMG_OVERFLOW, PLAN_ERR_HAND_COUNT, PLAN_ERR_BAD_ID, PLAN_NOT_UNIFORM = 1, 2, 4, 8, 16
VIT_MAX_LAYERS = 24
ABI_VERSION = 3
STAGES = ["im2col", "patch_gemm", "layernorm", "qkv_gemm", "attention", "outproj_gemm", "fc1_gemm", "fc2_gemm", "pool",
          "gemm", "splice", "gather", "other"]

p = C.c_void_p
i32, i64, f32, sz = C.c_int, C.c_int64, C.c_float, C.c_size_t


class _Layer(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b",
                                          "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


class _FoldLayer(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("w_qkv_f", "c_qkv", "b_qkv_f", "w_fc1_f", "c_fc1", "b_fc1_f")]


class VitLayout(C.Structure):
    _fields_ = [("patch_w", C.c_uint64), ("cls", C.c_uint64), ("pos", C.c_uint64), ("pre_ln_g", C.c_uint64),
                ("pre_ln_b", C.c_uint64), ("layer", _Layer * VIT_MAX_LAYERS), ("total_bytes", C.c_uint64),
                ("n_layers", C.c_int32), ("_pad", C.c_int32), ("fold", _FoldLayer * VIT_MAX_LAYERS)]


class ResizePlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_h", "in_w", "new_h", "new_w", "top", "left", "out_h", "out_w", "xk", "yk", "x_lo",
                                         "x_cols", "rows_cap", "table_ints")]


# symbol -> (restype, argtypes); must list every HVLM_API symbol of the header (tests/test_abi.py checks it)
SIGNATURES = {
    "hvlm_strerror": (C.c_char_p, [i32]),
    "hvlm_abi_version": (i32, []),
    "hvlm_device_check": (i32, [i32]),
    "hvlm_gemm_bf16": (i32, [p, p, p, p, p, i32, i32, i32, i32, i32, p]),
    "hvlm_vit_l14_layout": (i32, [i32, C.POINTER(VitLayout)]),
    "hvlm_vit_l14_workspace_bytes": (sz, [i32]),
    "hvlm_vit_l14_fwd": (i32, [p, i32, p, i32, i32, p, p, sz, p]),
    "hvlm_vit_l14_fwd_u8": (i32, [p, i32, p, C.POINTER(f32), C.POINTER(f32), i32, p, p, sz, p]),
    "hvlm_vit_l14_fwd_open_mlp": (i32, [p, i32, p, i32, p, p, i32, p, p, sz, C.POINTER(C.c_uint64), p]),
    "hvlm_feature_select": (i32, [p, p, i32, i32, i32, p]),
    "hvlm_layernorm_1024": (i32, [p, p, p, p, i32, i32, f32, p]),
    "hvlm_layernorm_1024_stats": (i32, [p, p, p, p, p, p, p, i32, f32, p]),
    "hvlm_gemm_ln_fold_bf16": (i32, [p, p, p, p, p, p, i32, i32, i32, i32, f32, p, p]),
    "hvlm_gemm_resid_stats": (i32, [p, p, p, p, p, p, p, i32, i32, p]),
    "hvlm_vit_set_ln_fold": (i32, [i32]),
    "hvlm_vit_qkv_gemm": (i32, [p, p, p, p, i32, p]),
    "hvlm_vit_attention": (i32, [p, p, i32, p]),
    "hvlm_pool_out_tokens": (i32, [i32, i32]),
    "hvlm_pool_slowfast_fwd": (i32, [p, i32, i64, p, i32, i32, i32, i32, i32, p]),
    "hvlm_pool_slowfast_fwd_mapped": (i32, [p, i32, i64, p, p, i32, i32, i32, i32, i32, p]),
    "hvlm_pool_slowfast_bwd": (i32, [p, i32, p, i32, i32, i32, i32, i32, p]),
    "hvlm_splice_count": (i32, [p, i32, i32, p, p]),
    "hvlm_splice_info": (i32, [p, i32, i32, i32, p, p]),
    "hvlm_splice_plan": (i32, [p, p, i32, i32, i32, i32, i32, i32, i32, i32, i32, p, p, p, p, p, p, p]),
    "hvlm_splice_plan_ragged": (i32, [p, p, p, i32, i32, i32, i32, i32, i32, i32, i32, p, p, p, p, p, p, p]),
    "hvlm_splice_fwd": (i32, [p, p, p, p, p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, i32, p, p, p, p]),
    "hvlm_splice_bwd": (i32, [p, i32, p, p, i32, i32, i32, i32, i32, p, p, p]),
    "hvlm_hand_gather_fwd": (i32, [p, i32, p, i64, i32, i32, i32, p, p, p, p, p]),
    "hvlm_hand_gather_bwd": (i32, [p, i32, p, i32, i32, i32, p, p]),
    "hvlm_hand_gather_step": (i32, [p, i32, i32, i32, p, p]),
    "hvlm_traj_decode_workspace_bytes": (sz, [i32, i32]),
    "hvlm_traj_decode": (i32, [p, i64, i32, p, p, p, p, p, i32, i32, i32, i32, i32, p, p, sz, p]),
    "hvlm_skinny_linear": (i32, [p, i64, p, p, i32, i32, i32, i32, i32, p, p]),
    "hvlm_frame_dedup_workspace_bytes": (sz, [i32]),
    "hvlm_frame_dedup": (i32, [p, sz, i32, i32, p, p, p, p, sz, p]),
    "hvlm_gather_rows": (i32, [p, sz, i32, p, i32, p, p]),
    "hvlm_resize_table_host": (i32, [i32, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hvlm_resize_plan_host": (i32, [i32, i32, i32, i32, p]),
    "hvlm_resize_tables_host": (i32, [p, C.POINTER(C.c_int32)]),
    "hvlm_resize_crop_u8": (i32, [p, i32, p, p, p, p]),
    "hvlm_pad_square_u8": (i32, [p, i32, i32, i32, p, i32, i32, i32, p]),
    "hvlm_transpose_to_bf16": (i32, [p, i32, p, i32, i32, i32, p]),
    "hvlm_colsum": (i32, [p, i32, p, i32, i32, p]),
    "hvlm_launch_count": (C.c_uint64, []),
    "hvlm_profile_enable": (i32, [i32]),
    "hvlm_profile_collect": (i32, [C.POINTER(f32), C.POINTER(C.c_int32)]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libhvlm_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once) and declare every signature.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no non-CUDA fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.hvlm_abi_version() != ABI_VERSION:
        raise RuntimeError("libhvlm_b200.so ABI version mismatch")
    _lib = L
    return L


def strerror(rc: int) -> str:
    return lib().hvlm_strerror(rc).decode()


class HvlmError(RuntimeError):
    def __init__(self, what: str, rc: int):
        super().__init__(f"{what} failed: {strerror(rc)} (status {rc})")
        self.status = rc


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise HvlmError(what, rc)
