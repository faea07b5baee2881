"""PyTorch custom ops over the C ABI (include/hvlm_b200.h).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; every op hands raw
pointers to libhvlm_b200.so.  Ops are registered with ``torch.library`` (namespace ``hvlm``) with fake
(meta) kernels for shape propagation and autograd formulas for the training-shaped variant
(pooling, projector, splice, gather).  There is no CPU implementation: calling an op with a
non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float16: L.F16}


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hvlm ops run on sm_100a CUDA devices only (no CPU fallback); got a "
                               f"{t.device} tensor")


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


_checked = set()


def ensure_device(index: Optional[int] = None) -> None:
    """Fail loudly unless the current device is an sm_100 GPU and the extension is loaded."""
    if not torch.cuda.is_available():
        raise RuntimeError("hvlm_b200 needs an sm_100a (B200) GPU; no CUDA device is visible")
    index = torch.cuda.current_device() if index is None else index
    if index not in _checked:
        L.check(L.lib().hvlm_device_check(index), "hvlm_device_check")
        _checked.add(index)


# ------------------------------------------------------------------------------------------------
# GEMM (nn.Linear):  y = x @ w.T (+ b) with fused epilogue
# ------------------------------------------------------------------------------------------------
def gemm(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: str = "bias",
         resid: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [M,K] bf16, w [N,K] bf16, bias [N] f32, resid [M,N] f32 -> [M,N]."""
    _need_cuda(x, w, bias, resid)
    ensure_device()
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16, "operands must be bf16"
    assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1]
    x, w = x.contiguous(), w.contiguous()
    M, K = x.shape
    N = w.shape[0]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        bias = bias.contiguous()
    epi = {"bias": L.EPI_BIAS, "quick_gelu": L.EPI_BIAS_QUICKGELU, "residual": L.EPI_BIAS_RESIDUAL}[epilogue]
    if epi == L.EPI_BIAS_RESIDUAL:
        assert resid is not None and resid.dtype == torch.float32 and tuple(resid.shape) == (M, N)
        resid = resid.contiguous()
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=x.device)
    L.check(L.lib().hvlm_gemm_bf16(_p(x), _p(w), _p(bias), _p(resid), _p(out), M, N, K, epi, _DT[out.dtype], _stream()),
            "hvlm_gemm_bf16")
    return out


def transpose_to_bf16(x: torch.Tensor, r_pad: Optional[int] = None) -> torch.Tensor:
    """x [R,C] -> bf16 [C, R_pad] (zero padded; R_pad multiple of 8)."""
    _need_cuda(x)
    x = x.contiguous()
    R, Cc = x.shape
    r_pad = (R + 7) // 8 * 8 if r_pad is None else r_pad
    out = torch.empty(Cc, r_pad, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().hvlm_transpose_to_bf16(_p(x), _dt(x), _p(out), R, Cc, r_pad, _stream()), "hvlm_transpose_to_bf16")
    return out


def colsum(dy: torch.Tensor) -> torch.Tensor:
    _need_cuda(dy)
    dy = dy.contiguous()
    M, N = dy.shape
    out = torch.empty(N, dtype=torch.float32, device=dy.device)
    L.check(L.lib().hvlm_colsum(_p(dy), _dt(dy), _p(out), M, N, _stream()), "hvlm_colsum")
    return out


@torch.library.custom_op("hvlm::linear", mutates_args=())
def linear(x: torch.Tensor, w_bf16: torch.Tensor, bias_f32: torch.Tensor, out_f32: bool) -> torch.Tensor:
    """mm_projector forward: x [M,K] (any float dtype, rounded to bf16), w [N,K] bf16, bias [N] f32."""
    xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
    return gemm(xb, w_bf16, bias_f32, epilogue="bias", out_dtype=torch.float32 if out_f32 else torch.bfloat16)


@linear.register_fake
def _(x, w_bf16, bias_f32, out_f32):
    return x.new_empty(x.shape[0], w_bf16.shape[0], dtype=torch.float32 if out_f32 else torch.bfloat16)


# (weight shape, device) -> (dW fp32 [N,K], db fp32 [N] | None): where the projector wgrad / bias grad are written when a
# dist.ProjectorGradReducer owns a flat all-reduce bucket (the collective then reads what the kernels wrote, no packing)
_wgrad_sinks: dict = {}


def register_wgrad_sink(weight_shape, device, dW: torch.Tensor, db: Optional[torch.Tensor]) -> None:
    """One sink per (projector weight shape, device): the backward of ``hvlm::linear`` finds it by that key (the weight it
    sees may be a bf16 copy of an fp32 parameter, so there is no stabler identity).  Two projectors of the same shape on
    one device cannot both own a sink -- the second registration raises; use ``ProjectorGradReducer(use_sink=False)``."""
    assert dW.dtype == torch.float32 and tuple(dW.shape) == tuple(weight_shape) and dW.is_contiguous()
    key = (tuple(weight_shape), torch.device(device))
    if key in _wgrad_sinks and _wgrad_sinks[key][0].data_ptr() != dW.data_ptr():
        raise RuntimeError(f"a gradient bucket is already registered for projector weights of shape {key[0]} on {key[1]}")
    _wgrad_sinks[key] = (dW, db)


def unregister_wgrad_sink(weight_shape, device) -> None:
    _wgrad_sinks.pop((tuple(weight_shape), torch.device(device)), None)


@torch.library.custom_op("hvlm::linear_bwd", mutates_args=())
def linear_bwd(dy: torch.Tensor, x: torch.Tensor, w_bf16: torch.Tensor, need_dx: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """dW [N,K] f32 = dY^T X, db [N] f32 = colsum(dY), dX [M,K] f32 = dY W (only if need_dx)."""
    M, N = dy.shape
    K = x.shape[1]
    dyT = transpose_to_bf16(dy)                  # [N, Mp]   A operand, K-major over tokens
    xT = transpose_to_bf16(x)                    # [K, Mp]   B operand
    sink = _wgrad_sinks.get(((N, K), dy.device))
    if sink is not None:
        # results land in the reducer's flat bucket; autograd gets the bf16 cast it would make anyway (custom-op outputs
        # must not alias the sink)
        gemm(dyT, xT, None, out_dtype=torch.float32, out=sink[0])
        dW = sink[0].to(torch.bfloat16)
        if sink[1] is not None:
            L.check(L.lib().hvlm_colsum(_p(dy.contiguous()), _dt(dy), _p(sink[1]), M, N, _stream()), "hvlm_colsum")
            db = sink[1].clone()
        else:
            db = colsum(dy)
    else:
        dW = gemm(dyT, xT, None, out_dtype=torch.float32)
        db = colsum(dy)
    if need_dx:
        wT = transpose_to_bf16(w_bf16)           # [K, N]
        dyb = dy if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16)
        dx = gemm(dyb.contiguous(), wT, None, out_dtype=torch.float32)
    else:
        dx = torch.empty(0, dtype=torch.float32, device=dy.device)
    return dW, db, dx


@linear_bwd.register_fake
def _(dy, x, w_bf16, need_dx):
    N, K = w_bf16.shape
    return (dy.new_empty(N, K, dtype=torch.float32), dy.new_empty(N, dtype=torch.float32),
            dy.new_empty((x.shape[0], K) if need_dx else (0,), dtype=torch.float32))


def _linear_setup(ctx, inputs, output):
    x, w, b, _ = inputs
    ctx.save_for_backward(x, w)
    ctx.need_dx = x.requires_grad


def _linear_backward(ctx, dy):
    x, w = ctx.saved_tensors
    dW, db, dx = linear_bwd(dy.contiguous(), x, w, ctx.need_dx)
    return (dx.to(x.dtype) if ctx.need_dx else None), dW.to(w.dtype), db, None


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------------
# ViT-L/14 tower
# ------------------------------------------------------------------------------------------------
def _vit_workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    """Scratch for hvlm_vit_l14_fwd: the C side wants a 1024-byte aligned base (SWIZZLE_128B TMA tiles); the caching
    allocator only promises 512, so over-allocate and slice."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


@torch.library.custom_op("hvlm::vit_l14_hidden", mutates_args=())
def vit_l14_hidden(weight_blob: torch.Tensor, pixels: torch.Tensor, n_layers_run: int) -> torch.Tensor:
    """pixels [N,3,224,224] -> residual stream f32 [N,257,1024] after n_layers_run layers."""
    _need_cuda(weight_blob, pixels)
    ensure_device()
    if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, 224, 224):
        raise ValueError(f"Input image size ({tuple(pixels.shape[-2:])}) doesn't match model (224*224).")
    pixels = pixels.contiguous()
    N = pixels.shape[0]
    lib = L.lib()
    ws_bytes = lib.hvlm_vit_l14_workspace_bytes(N)
    ws = _vit_workspace(ws_bytes, pixels.device)
    hidden = torch.empty(N, 257, 1024, dtype=torch.float32, device=pixels.device)
    L.check(lib.hvlm_vit_l14_fwd(_p(weight_blob), n_layers_run, _p(pixels), _dt(pixels), N, _p(hidden), _p(ws),
                                 ws_bytes, _stream()), "hvlm_vit_l14_fwd")
    return hidden


def vit_l14_hidden_open(weight_blob: torch.Tensor, pixels: torch.Tensor, n_layers_run: int):
    """Tower forward with the last layer's second MLP matmul left open (hvlm_vit_l14_fwd_open_mlp):
    -> (hidden f32 [N,257,1024] after the last attention block, f1 bf16 [N,257,4096] = gelu(fc1(LN2(hidden)))).
    ``pixels``: float [N,3,224,224] or raw uint8 [N,224,224,3].  Inference-only helper of the pooled token path; ``f1`` is
    a view into the call's workspace."""
    _need_cuda(weight_blob, pixels)
    ensure_device()
    u8 = pixels.dtype == torch.uint8
    if u8:
        if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (224, 224, 3):
            raise ValueError(f"expected uint8 frames [N,224,224,3], got {tuple(pixels.shape)}")
    elif pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, 224, 224):
        raise ValueError(f"Input image size ({tuple(pixels.shape[-2:])}) doesn't match model (224*224).")
    pixels = pixels.contiguous()
    N = pixels.shape[0]
    lib = L.lib()
    ws_bytes = lib.hvlm_vit_l14_workspace_bytes(N)
    ws = _vit_workspace(ws_bytes, pixels.device)
    hidden = torch.empty(N, 257, 1024, dtype=torch.float32, device=pixels.device)
    f1_off = C.c_uint64(0)
    mean = (C.c_float * 3)(*CLIP_MEAN)
    std = (C.c_float * 3)(*CLIP_STD)
    L.check(lib.hvlm_vit_l14_fwd_open_mlp(_p(weight_blob), n_layers_run, _p(pixels), -1 if u8 else _dt(pixels), mean, std, N,
                                          _p(hidden), _p(ws), ws_bytes, C.byref(f1_off), _stream()),
            "hvlm_vit_l14_fwd_open_mlp")
    o = int(f1_off.value)
    f1 = ws[o:o + N * 257 * 4096 * 2].view(torch.bfloat16).view(N, 257, 4096)
    return hidden, f1


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)    # OPENAI_CLIP_MEAN / STD (transformers CLIPImageProcessor defaults)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@torch.library.custom_op("hvlm::vit_l14_hidden_u8", mutates_args=())
def vit_l14_hidden_u8(weight_blob: torch.Tensor, frames: torch.Tensor, n_layers_run: int) -> torch.Tensor:
    """raw uint8 frames [N,224,224,3] (NHWC) -> residual stream f32 [N,257,1024]; rescale + CLIP normalisation fused
    into the patch extraction."""
    _need_cuda(weight_blob, frames)
    ensure_device()
    if frames.dtype != torch.uint8 or frames.dim() != 4 or tuple(frames.shape[1:]) != (224, 224, 3):
        raise ValueError(f"expected uint8 frames [N,224,224,3], got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    N = frames.shape[0]
    lib = L.lib()
    ws_bytes = lib.hvlm_vit_l14_workspace_bytes(N)
    ws = _vit_workspace(ws_bytes, frames.device)
    hidden = torch.empty(N, 257, 1024, dtype=torch.float32, device=frames.device)
    mean = (C.c_float * 3)(*CLIP_MEAN)
    std = (C.c_float * 3)(*CLIP_STD)
    L.check(lib.hvlm_vit_l14_fwd_u8(_p(weight_blob), n_layers_run, _p(frames), mean, std, N, _p(hidden), _p(ws), ws_bytes,
                                    _stream()), "hvlm_vit_l14_fwd_u8")
    return hidden


@vit_l14_hidden_u8.register_fake
def _(weight_blob, frames, n_layers_run):
    return frames.new_empty(frames.shape[0], 257, 1024, dtype=torch.float32)


@vit_l14_hidden.register_fake
def _(weight_blob, pixels, n_layers_run):
    return pixels.new_empty(pixels.shape[0], 257, 1024, dtype=torch.float32)


def feature_select(hidden: torch.Tensor, out_dtype: torch.dtype, keep_cls: bool = False) -> torch.Tensor:
    _need_cuda(hidden)
    N = hidden.shape[0]
    out = torch.empty(N, 257 if keep_cls else 256, 1024, dtype=out_dtype, device=hidden.device)
    L.check(L.lib().hvlm_feature_select(_p(hidden), _p(out), N, _DT[out_dtype], int(keep_cls), _stream()),
            "hvlm_feature_select")
    return out


def layernorm_1024(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, out_dtype=torch.bfloat16, eps=1e-5):
    _need_cuda(x, g, b)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    L.check(L.lib().hvlm_layernorm_1024(_p(x), _p(g), _p(b), _p(out), x.numel() // 1024, _DT[out_dtype], eps, _stream()),
            "hvlm_layernorm_1024")
    return out


def layernorm_1024_stats(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps=1e-5):
    """LayerNorm with fp32 output plus what a folded GEMM consumes next:
    (out f32, bf16(out - shift), stats f32 [rows,8,2] of (out - shift), shift f32 [rows] = row mean of out)."""
    _need_cuda(x, g, b)
    ensure_device()
    x = x.contiguous()
    rows = x.numel() // 1024
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    xb = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    stats = torch.empty(rows, 8, 2, dtype=torch.float32, device=x.device)
    shift = torch.empty(rows, dtype=torch.float32, device=x.device)
    L.check(L.lib().hvlm_layernorm_1024_stats(_p(x), _p(g), _p(b), _p(out), _p(xb), _p(stats), _p(shift), rows, eps, _stream()),
            "hvlm_layernorm_1024_stats")
    return out, xb, stats, shift


def gemm_ln_fold(xb: torch.Tensor, stats: torch.Tensor, w_f: torch.Tensor, c: torch.Tensor, b_f: torch.Tensor, *,
                 epilogue: str = "bias", qkv_hm: bool = False, eps: float = 1e-5,
                 shift_io: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm folded into its consumer GEMM (hvlm_gemm_ln_fold_bf16): xb bf16 [M,1024] un-normalised (centred) rows, stats
    [M,8,2], w_f bf16 [N,1024] = gamma*W, c [N] its row sums, b_f [N] = b + W beta -> bf16 [M,N] (or [48,M,64]).
    ``shift_io`` f32 [M] (optional, updated IN PLACE): += the mean of each row of xb (the next producer's centre)."""
    _need_cuda(xb, stats, w_f, c, b_f)
    ensure_device()
    M, N = xb.shape[0], w_f.shape[0]
    assert xb.dtype == torch.bfloat16 and w_f.dtype == torch.bfloat16 and xb.shape[1] == 1024 and w_f.shape[1] == 1024
    assert stats.dtype == torch.float32 and stats.numel() == M * 16 and stats.is_contiguous()
    out = torch.empty((48, M, 64) if qkv_hm else (M, N), dtype=torch.bfloat16, device=xb.device)
    L.check(L.lib().hvlm_gemm_ln_fold_bf16(_p(xb.contiguous()), _p(stats), _p(w_f.contiguous()), _p(c), _p(b_f), _p(out), M, N,
                                           {"bias": L.EPI_BIAS, "quick_gelu": L.EPI_BIAS_QUICKGELU}[epilogue], int(qkv_hm), eps,
                                           _p(shift_io), _stream()), "hvlm_gemm_ln_fold_bf16")
    return out


def gemm_resid_stats(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], hidden: torch.Tensor,
                     shift: Optional[torch.Tensor] = None):
    """hidden f32 [M,1024] += a w^T + bias IN PLACE; returns (bf16(hidden - shift), stats [M,8,2] of (hidden - shift)) from
    the same epilogue (shift f32 [M], default 0)."""
    _need_cuda(a, w, hidden)
    ensure_device()
    M, K = a.shape
    assert hidden.dtype == torch.float32 and tuple(hidden.shape) == (M, 1024) and hidden.is_contiguous()
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and tuple(w.shape) == (1024, K)
    xb = torch.empty(M, 1024, dtype=torch.bfloat16, device=a.device)
    stats = torch.empty(M, 8, 2, dtype=torch.float32, device=a.device)
    L.check(L.lib().hvlm_gemm_resid_stats(_p(a.contiguous()), _p(w.contiguous()), _p(bias), _p(hidden), _p(shift), _p(xb),
                                          _p(stats), M, K, _stream()), "hvlm_gemm_resid_stats")
    return xb, stats


def vit_set_ln_fold(on: int) -> int:
    """Process-wide switch of the tower's folded LayerNorms (hvlm_vit_set_ln_fold); returns the previous setting."""
    return int(L.lib().hvlm_vit_set_ln_fold(int(on)))


def vit_qkv(y: torch.Tensor, w_qkv: torch.Tensor, b_qkv: torch.Tensor, n_frames: int) -> torch.Tensor:
    """y bf16 [M=n_frames*257,1024] -> qkv column-block-major bf16 [48, M, 64] (q0..15|k0..15|v0..15)."""
    _need_cuda(y, w_qkv, b_qkv)
    ensure_device()
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (n_frames * 257, 1024)
    out = torch.empty(48, n_frames * 257, 64, dtype=torch.bfloat16, device=y.device)
    L.check(L.lib().hvlm_vit_qkv_gemm(_p(y.contiguous()), _p(w_qkv.contiguous()), _p(b_qkv), _p(out), n_frames, _stream()),
            "hvlm_vit_qkv_gemm")
    return out


def vit_attention(qkv_hm: torch.Tensor) -> torch.Tensor:
    """qkv column-block-major bf16 [48, n_frames*257, 64] (q pre-scaled) -> bf16 [n_frames*257, 1024]."""
    _need_cuda(qkv_hm)
    ensure_device()
    assert qkv_hm.dtype == torch.bfloat16 and qkv_hm.dim() == 3 and qkv_hm.shape[0] == 48 and qkv_hm.shape[2] == 64
    assert qkv_hm.shape[1] % 257 == 0 and qkv_hm.is_contiguous()
    n_frames = qkv_hm.shape[1] // 257
    out = torch.empty(n_frames * 257, 1024, dtype=torch.bfloat16, device=qkv_hm.device)
    L.check(L.lib().hvlm_vit_attention(_p(qkv_hm), _p(out), n_frames, _stream()), "hvlm_vit_attention")
    return out


def launch_count() -> int:
    return int(L.lib().hvlm_launch_count())


def profile_enable(on: bool) -> None:
    L.lib().hvlm_profile_enable(int(on))


def profile_collect() -> dict:
    """-> {stage: (device_ms, launches)} summed over everything recorded since profile_enable(True)."""
    n = len(L.STAGES)
    ms = (C.c_float * n)()
    cnt = (C.c_int32 * n)()
    L.check(L.lib().hvlm_profile_collect(ms, cnt), "hvlm_profile_collect")
    return {L.STAGES[i]: (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}


# ------------------------------------------------------------------------------------------------
# slow-fast pooling
# ------------------------------------------------------------------------------------------------
def pool_out_tokens(t: int, mode: str) -> int:
    return L.lib().hvlm_pool_out_tokens(t, L.POOL_MODES[mode])


@torch.library.custom_op("hvlm::pool_slowfast", mutates_args=())
def pool_slowfast(tok: torch.Tensor, B: int, t: int, frame_stride: int, row_offset: int, mode: int,
                  out_bf16: bool) -> torch.Tensor:
    """tok: flat buffer viewed as [B*t, frame_stride, C] (rows row_offset..row_offset+255 of each frame are the
    256 tokens) -> [B, n_out, C].  frame_stride=257,row_offset=1 reads the tower's hidden state in place."""
    _need_cuda(tok)
    ensure_device()
    assert tok.is_contiguous() and tok.dim() == 3 and tok.shape[0] == B * t and tok.shape[1] == frame_stride
    assert row_offset + 256 <= frame_stride
    Cc = tok.shape[2]
    lib = L.lib()
    n_out = lib.hvlm_pool_out_tokens(t, mode)
    out = torch.empty(B, n_out, Cc, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=tok.device)
    base = C.c_void_p(tok.data_ptr() + row_offset * Cc * tok.element_size())
    L.check(lib.hvlm_pool_slowfast_fwd(base, _dt(tok), frame_stride, _p(out), _DT[out.dtype], B, t, Cc, mode, _stream()),
            "hvlm_pool_slowfast_fwd")
    return out


@pool_slowfast.register_fake
def _(tok, B, t, frame_stride, row_offset, mode, out_bf16):
    n_out = {0: t + 256, 1: 256, 2: t, 3: 256, 4: t + 256}[mode]
    return tok.new_empty(B, n_out, tok.shape[2], dtype=torch.bfloat16 if out_bf16 else torch.float32)


@torch.library.custom_op("hvlm::pool_slowfast_bwd", mutates_args=())
def pool_slowfast_bwd(dout: torch.Tensor, t: int, mode: int, out_bf16: bool) -> torch.Tensor:
    _need_cuda(dout)
    dout = dout.contiguous()
    B, _, Cc = dout.shape
    dtok = torch.empty(B, t, 256, Cc, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dout.device)
    L.check(L.lib().hvlm_pool_slowfast_bwd(_p(dout), _dt(dout), _p(dtok), _DT[dtok.dtype], B, t, Cc, mode, _stream()),
            "hvlm_pool_slowfast_bwd")
    return dtok


@pool_slowfast_bwd.register_fake
def _(dout, t, mode, out_bf16):
    return dout.new_empty(dout.shape[0], t, 256, dout.shape[2], dtype=torch.bfloat16 if out_bf16 else torch.float32)


def _pool_setup(ctx, inputs, output):
    tok, B, t, frame_stride, row_offset, mode, _ = inputs
    ctx.t, ctx.mode, ctx.frame_stride, ctx.row_offset, ctx.in_dtype, ctx.B = t, mode, frame_stride, row_offset, tok.dtype, B


def _pool_backward(ctx, dout):
    d = pool_slowfast_bwd(dout.contiguous(), ctx.t, ctx.mode, ctx.in_dtype == torch.bfloat16)
    d = d.reshape(ctx.B * ctx.t, 256, -1)
    if ctx.frame_stride != 256:
        full = d.new_zeros(ctx.B * ctx.t, ctx.frame_stride, d.shape[-1])
        full[:, ctx.row_offset:ctx.row_offset + 256] = d
        d = full
    return d.to(ctx.in_dtype), None, None, None, None, None, None


pool_slowfast.register_autograd(_pool_backward, setup_context=_pool_setup)


def pool_slowfast_mapped(tok: torch.Tensor, frame_map: torch.Tensor, B: int, t: int, frame_stride: int, row_offset: int,
                         mode: int, out_bf16: bool) -> torch.Tensor:
    """pool_slowfast with a frame indirection (frame de-duplication): tok [n_unique, frame_stride, C], frame_map int32
    [B*t] -> [B, n_out, C].  Forward only (the tower output never requires grad)."""
    _need_cuda(tok, frame_map)
    ensure_device()
    assert tok.is_contiguous() and tok.dim() == 3 and tok.shape[1] == frame_stride and row_offset + 256 <= frame_stride
    assert frame_map.dtype == torch.int32 and frame_map.numel() == B * t and frame_map.is_contiguous()
    Cc = tok.shape[2]
    lib = L.lib()
    n_out = lib.hvlm_pool_out_tokens(t, mode)
    out = torch.empty(B, n_out, Cc, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=tok.device)
    base = C.c_void_p(tok.data_ptr() + row_offset * Cc * tok.element_size())
    L.check(lib.hvlm_pool_slowfast_fwd_mapped(base, _dt(tok), frame_stride, _p(frame_map), _p(out), _DT[out.dtype], B, t, Cc,
                                              mode, _stream()), "hvlm_pool_slowfast_fwd_mapped")
    return out


def pool_tokens(tokens: torch.Tensor, mode: str, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """tokens [b,t,256,C] -> [b,n_out,C]  (compress_tokens / videos_to_tokens pooling)."""
    b, t, s, c = tokens.shape
    assert s == 256, f"tokens.shape = {tuple(tokens.shape)}"
    out_dtype = tokens.dtype if out_dtype is None else out_dtype
    return pool_slowfast(tokens.contiguous().reshape(b * t, 256, c), b, t, 256, 0, L.POOL_MODES[mode],
                         out_dtype == torch.bfloat16)


# ------------------------------------------------------------------------------------------------
# splice
# ------------------------------------------------------------------------------------------------
def splice_count(ids: torch.Tensor) -> torch.Tensor:
    _need_cuda(ids)
    ensure_device()
    ids = ids.contiguous()
    B, T = ids.shape
    counts = torch.empty(B, dtype=torch.int32, device=ids.device)
    L.check(L.lib().hvlm_splice_count(_p(ids), B, T, _p(counts), _stream()), "hvlm_splice_count")
    return counts


def splice_info(ids: torch.Tensor, vocab: int) -> torch.Tensor:
    """-> int32 [4, B]: image-token count | last image-token position | <hand_traj> tokens after it | bad-id flag."""
    _need_cuda(ids)
    ensure_device()
    ids = ids.contiguous()
    B, T = ids.shape
    info = torch.empty(4, B, dtype=torch.int32, device=ids.device)
    L.check(L.lib().hvlm_splice_info(_p(ids), B, T, int(vocab), _p(info), _stream()), "hvlm_splice_info")
    return info


def splice_plan(ids: torch.Tensor, counts: torch.Tensor, Nv: int, n_img: int, Lout: int, vocab: int, variant: int,
                hand_mode: int, n_hand: int, slot_offsets: Optional[torch.Tensor] = None,
                last_visual_end: Optional[torch.Tensor] = None):
    """``slot_offsets`` (int32 [n_img+1], device): per-slot row ranges of a row-concatenated visual tensor, for visual
    token blocks of different lengths (``Nv`` is ignored then).  ``last_visual_end``: optional 0-d int64 device tensor
    that receives the reference's ``last_visual_token_index`` (handsonvlm.py:288)."""
    if last_visual_end is not None:
        assert last_visual_end.dtype == torch.int64 and last_visual_end.numel() == 1 and last_visual_end.is_cuda
    ids = ids.contiguous()
    B, T = ids.shape
    dev = ids.device
    src_index = torch.empty(B, Lout, dtype=torch.int32, device=dev)
    hand_code = torch.empty(B, Lout, dtype=torch.int8, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    hand_scale = torch.empty(B, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if slot_offsets is not None:
        assert slot_offsets.dtype == torch.int32 and slot_offsets.numel() == n_img + 1 and slot_offsets.is_cuda
        L.check(L.lib().hvlm_splice_plan_ragged(_p(ids), _p(counts), _p(slot_offsets.contiguous()), B, T, n_img, Lout,
                                                vocab, variant, hand_mode, n_hand, _p(src_index), _p(hand_code),
                                                _p(lens), _p(hand_scale), _p(status), _p(last_visual_end), _stream()),
                "hvlm_splice_plan_ragged")
        return src_index, hand_code, lens, hand_scale, status
    L.check(L.lib().hvlm_splice_plan(_p(ids), _p(counts), B, T, Nv, n_img, Lout, vocab, variant, hand_mode, n_hand,
                                     _p(src_index), _p(hand_code), _p(lens), _p(hand_scale), _p(status),
                                     _p(last_visual_end), _stream()),
            "hvlm_splice_plan")
    return src_index, hand_code, lens, hand_scale, status


@torch.library.custom_op("hvlm::splice_gather", mutates_args=())
def splice_gather(src_index: torch.Tensor, hand_code: torch.Tensor, lens: torch.Tensor, hand_scale: torch.Tensor,
                  ids: torch.Tensor, labels: Optional[torch.Tensor], mask: Optional[torch.Tensor],
                  table: torch.Tensor, visual: torch.Tensor, visual_mask: Optional[torch.Tensor],
                  future_hands: Optional[torch.Tensor], variant: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Index-driven gather -> (embeds [B,L,D], labels [B,L] i64, mask [B,L] bool)."""
    _need_cuda(src_index, ids, table, visual)
    B, Lout = src_index.shape
    T = ids.shape[1]
    n_img, Nv, D = visual.shape
    assert table.dtype == visual.dtype, (table.dtype, visual.dtype)
    dev = ids.device
    embeds = torch.empty(B, Lout, D, dtype=table.dtype, device=dev)
    out_labels = torch.empty(B, Lout, dtype=torch.int64, device=dev)
    out_mask = torch.empty(B, Lout, dtype=torch.bool, device=dev)
    n_hand = 0
    if future_hands is not None:
        future_hands = future_hands.to(torch.float32).contiguous()
        n_hand = future_hands.shape[2]
    mask_u8 = None if mask is None else mask.contiguous().view(torch.uint8)
    vmask_u8 = None if visual_mask is None else visual_mask.contiguous().view(torch.uint8)
    L.check(L.lib().hvlm_splice_fwd(_p(src_index), _p(hand_code), _p(lens), _p(hand_scale), _p(ids.contiguous()),
                                    _p(None if labels is None else labels.contiguous()), _p(mask_u8),
                                    _p(table.contiguous()), _p(visual.contiguous()), _p(vmask_u8), _p(future_hands),
                                    n_hand, B, T, Lout, Nv, D, _dt(table), variant, _p(embeds), _p(out_labels),
                                    _p(out_mask.view(torch.uint8)), _stream()), "hvlm_splice_fwd")
    return embeds, out_labels, out_mask


@splice_gather.register_fake
def _(src_index, hand_code, lens, hand_scale, ids, labels, mask, table, visual, visual_mask, future_hands, variant):
    B, Lout = src_index.shape
    return (table.new_empty(B, Lout, table.shape[1]), ids.new_empty(B, Lout), ids.new_empty(B, Lout, dtype=torch.bool))


@torch.library.custom_op("hvlm::splice_bwd", mutates_args=())
def splice_bwd(d_embeds: torch.Tensor, src_index: torch.Tensor, ids: torch.Tensor, n_visual_rows: int, vocab: int,
               need_table: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(d_embeds)
    d_embeds = d_embeds.contiguous()
    B, Lout, D = d_embeds.shape
    dev = d_embeds.device
    # visual rows that no sample references keep a zero gradient
    d_visual = torch.zeros(n_visual_rows, D, dtype=torch.float32, device=dev)
    d_table = torch.zeros(vocab, D, dtype=torch.float32, device=dev) if need_table else torch.empty(0, device=dev)
    L.check(L.lib().hvlm_splice_bwd(_p(d_embeds), _dt(d_embeds), _p(src_index), _p(ids.contiguous()), B, ids.shape[1],
                                    Lout, n_visual_rows, D, _p(d_visual), _p(d_table if need_table else None),
                                    _stream()), "hvlm_splice_bwd")
    return d_visual, d_table


@splice_bwd.register_fake
def _(d_embeds, src_index, ids, n_visual_rows, vocab, need_table):
    D = d_embeds.shape[2]
    return (d_embeds.new_empty(n_visual_rows, D, dtype=torch.float32),
            d_embeds.new_empty((vocab, D) if need_table else (0,), dtype=torch.float32))


def _splice_setup(ctx, inputs, output):
    src_index, _, _, _, ids, _, _, table, visual, _, _, variant = inputs
    ctx.save_for_backward(src_index, ids)
    ctx.im_start_end = bool(variant & L.SPLICE_FLAG_IM_START_END)
    ctx.vshape = visual.shape
    ctx.vocab = table.shape[0]
    ctx.need_table = table.requires_grad
    ctx.need_visual = visual.requires_grad
    ctx.tdtype, ctx.vdtype = table.dtype, visual.dtype
    ctx.set_materialize_grads(False)


def _splice_backward(ctx, d_embeds, d_labels, d_mask):
    src_index, ids = ctx.saved_tensors
    gv = gt = None
    if d_embeds is not None and (ctx.need_table or ctx.need_visual):
        n_img, Nv, D = ctx.vshape
        if ctx.im_start_end and ctx.need_table:
            # llava_arch.py:150,181: every text segment is .detach()ed except the token right before / after an image
            # token -> drop the other text rows from the plan the backward kernel walks (INT32_MIN = skip)
            img = ids == -200
            near = torch.zeros_like(img)
            near[:, :-1] |= img[:, 1:]
            near[:, 1:] |= img[:, :-1]
            is_text = src_index >= 0
            keep = near.gather(1, src_index.clamp_min(0).long())
            src_index = torch.where(is_text & ~keep, torch.full_like(src_index, -2 ** 31), src_index)
        d_visual, d_table = splice_bwd(d_embeds, src_index, ids, n_img * Nv, ctx.vocab, ctx.need_table)
        if ctx.need_visual:
            gv = d_visual.reshape(n_img, Nv, D).to(ctx.vdtype)
        if ctx.need_table:
            gt = d_table.to(ctx.tdtype)
    return None, None, None, None, None, None, None, gt, gv, None, None, None


splice_gather.register_autograd(_splice_backward, setup_context=_splice_setup)


# ------------------------------------------------------------------------------------------------
# <hand_traj> gather
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("hvlm::hand_gather", mutates_args=())
def hand_gather(hidden: torch.Tensor, labels: torch.Tensor, hand_id: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (out [B,2,4,D/2], valid [B] bool, rows [B,4] i32, counts [B] i32)."""
    _need_cuda(hidden, labels)
    ensure_device()
    hidden, labels = hidden.contiguous(), labels.contiguous()
    B, Lh, D = hidden.shape
    dev = hidden.device
    out = torch.empty(B, 2, 4, D // 2, dtype=hidden.dtype, device=dev)
    valid = torch.empty(B, dtype=torch.bool, device=dev)
    rows = torch.empty(B, 4, dtype=torch.int32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(L.lib().hvlm_hand_gather_fwd(_p(hidden), _dt(hidden), _p(labels), hand_id, B, Lh, D, _p(out),
                                         _p(valid.view(torch.uint8)), _p(rows), _p(counts), _stream()),
            "hvlm_hand_gather_fwd")
    return out, valid, rows, counts


@hand_gather.register_fake
def _(hidden, labels, hand_id):
    B, Lh, D = hidden.shape
    return (hidden.new_empty(B, 2, 4, D // 2), hidden.new_empty(B, dtype=torch.bool),
            hidden.new_empty(B, 4, dtype=torch.int32), hidden.new_empty(B, dtype=torch.int32))


@torch.library.custom_op("hvlm::hand_gather_bwd", mutates_args=())
def hand_gather_bwd(dout: torch.Tensor, rows: torch.Tensor, Lh: int) -> torch.Tensor:
    _need_cuda(dout, rows)
    dout = dout.contiguous()
    B, _, _, half = dout.shape
    dh = torch.zeros(B, Lh, 2 * half, dtype=torch.float32, device=dout.device)
    L.check(L.lib().hvlm_hand_gather_bwd(_p(dout), _dt(dout), _p(rows), B, Lh, 2 * half, _p(dh), _stream()),
            "hvlm_hand_gather_bwd")
    return dh


@hand_gather_bwd.register_fake
def _(dout, rows, Lh):
    return dout.new_empty(dout.shape[0], Lh, 2 * dout.shape[3], dtype=torch.float32)


def _gather_setup(ctx, inputs, output):
    hidden, _, _ = inputs
    ctx.save_for_backward(output[2])
    ctx.Lh, ctx.dtype = hidden.shape[1], hidden.dtype
    ctx.set_materialize_grads(False)


def _gather_backward(ctx, dout, dvalid, drows, dcounts):
    if dout is None:
        return None, None, None
    (rows,) = ctx.saved_tensors
    return hand_gather_bwd(dout, rows, ctx.Lh).to(ctx.dtype), None, None


hand_gather.register_autograd(_gather_backward, setup_context=_gather_setup)


_traj_ws: dict = {}


def _traj_workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch for hvlm_traj_decode, one per (device, stream); the kernel leaves it zeroed."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _traj_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
        _traj_ws[key] = ws
    return ws


def traj_decode(cond: torch.Tensor, z: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
                b2: torch.Tensor, interleaved: bool = False) -> torch.Tensor:
    """CVAE trajectory decoder, generation side (TrajCVAE.inference, hoi_forecast/architecture/traj_decoder.py:75-91):
    ``ELU(cat(z, cond) W1^T + b1) W2^T + b2`` in one launch -> [R, 2] fp32.

    cond [R, Dc] (``interleaved=False``) or the raw last hidden rows [R/2, 2*Dc] (``interleaved=True``: the even/odd
    de-interleave of handsonvlm.py:613-616 happens inside the kernel).  z [R, L] is the already scaled noise."""
    _need_cuda(cond)
    ensure_device()
    if cond.dim() != 2 or z.dim() != 2:
        raise AssertionError((cond.shape, z.shape))
    if cond.stride(1) != 1:
        cond = cond.contiguous()
    R, Lz = z.shape
    Dc = cond.shape[1] // 2 if interleaved else cond.shape[1]
    H = w1.shape[0]
    if interleaved:
        assert cond.shape[0] * 2 == R and cond.shape[1] == 2 * Dc, (cond.shape, z.shape)
    else:
        assert cond.shape[0] == R, (cond.shape, z.shape)
    assert w1.shape == (H, Lz + Dc) and b1.shape == (H,) and w2.shape == (2, H) and b2.shape == (2,), \
        (w1.shape, b1.shape, w2.shape, b2.shape)
    dt = cond.dtype
    z, w1, b1, w2, b2 = (t.to(dt).contiguous() for t in (z, w1, b1, w2, b2))
    out = torch.empty(R, 2, dtype=torch.float32, device=cond.device)
    need = int(L.lib().hvlm_traj_decode_workspace_bytes(R, H))
    ws = _traj_workspace(cond.device, need)
    L.check(L.lib().hvlm_traj_decode(_p(cond), cond.stride(0), int(interleaved), _p(z), _p(w1), _p(b1), _p(w2), _p(b2),
                                     _dt(cond), R, Dc, Lz, H, _p(out), _p(ws), ws.numel(), _stream()),
            "hvlm_traj_decode")
    return out


def skinny_linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act: str = "none") -> torch.Tensor:
    """``act(x W^T + b)`` for a handful of rows (one warp per output unit, fp32 accumulate): the layers of the MLP trajectory
    decoder in the generation loop.  x [R,K], w [H,K] -> [R,H] in x.dtype."""
    _need_cuda(x, w)
    ensure_device()
    assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1], (x.shape, w.shape)
    if x.stride(1) != 1:
        x = x.contiguous()
    w = w.to(x.dtype).contiguous()
    b = None if b is None else b.to(x.dtype).contiguous()
    R, K = x.shape
    H = w.shape[0]
    out = torch.empty(R, H, dtype=x.dtype, device=x.device)
    L.check(L.lib().hvlm_skinny_linear(_p(x), x.stride(0), _p(w), _p(b), {"none": 0, "relu": 1, "elu": 2}[act], _dt(x),
                                       R, K, H, _p(out), _stream()), "hvlm_skinny_linear")
    return out


def hand_gather_step(hidden_last: torch.Tensor) -> torch.Tensor:
    """Generation-time gather (handsonvlm.py:613-616): [B,D] -> [B,2,1,D/2]."""
    _need_cuda(hidden_last)
    ensure_device()
    hidden_last = hidden_last.contiguous()
    B, D = hidden_last.shape
    out = torch.empty(B, 2, 1, D // 2, dtype=hidden_last.dtype, device=hidden_last.device)
    L.check(L.lib().hvlm_hand_gather_step(_p(hidden_last), _dt(hidden_last), B, D, _p(out), _stream()),
            "hvlm_hand_gather_step")
    return out


# ------------------------------------------------------------------------------------------------
# frame de-duplication and the uint8 resize / centre crop in front of the tower
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("hvlm::frame_dedup", mutates_args=())
def frame_dedup(frames: torch.Tensor, capacity: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """frames [N, ...] (any dtype, contiguous; bytes per frame % 16 == 0) -> (frame_map int32 [N], rep int32 [N],
    n_unique int32 [1]), all on the device, no host sync.  ``rep[:n_unique]`` lists the first occurrence of every distinct
    frame (padded with 0), ``frame_map[i]`` is frame i's index in that list."""
    _need_cuda(frames)
    ensure_device()
    frames = frames.contiguous()
    N = frames.shape[0]
    fb = frames[0].numel() * frames.element_size()
    dev = frames.device
    frame_map = torch.empty(N, dtype=torch.int32, device=dev)
    rep = torch.empty(N, dtype=torch.int32, device=dev)
    n_unique = torch.empty(1, dtype=torch.int32, device=dev)
    lib = L.lib()
    need = int(lib.hvlm_frame_dedup_workspace_bytes(N))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    L.check(lib.hvlm_frame_dedup(_p(frames), fb, N, int(capacity), _p(frame_map), _p(rep), _p(n_unique), _p(ws), need, _stream()),
            "hvlm_frame_dedup")
    return frame_map, rep, n_unique


@frame_dedup.register_fake
def _(frames, capacity=0):
    n = frames.shape[0]
    return (frames.new_empty(n, dtype=torch.int32), frames.new_empty(n, dtype=torch.int32),
            frames.new_empty(1, dtype=torch.int32))


@torch.library.custom_op("hvlm::gather_rows", mutates_args=())
def gather_rows(src: torch.Tensor, idx: torch.Tensor, n_out: Optional[int] = None) -> torch.Tensor:
    """out[i] = src[idx[i]] for i < n_out (idx int32 on the device; n_out is a HOST number, default len(idx))."""
    _need_cuda(src, idx)
    ensure_device()
    src = src.contiguous()
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    n_out = idx.numel() if n_out is None else n_out
    assert 0 < n_out <= idx.numel()
    rb = src[0].numel() * src.element_size()
    out = torch.empty((n_out,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    L.check(L.lib().hvlm_gather_rows(_p(src), rb, src.shape[0], _p(idx), n_out, _p(out), _stream()), "hvlm_gather_rows")
    return out


@gather_rows.register_fake
def _(src, idx, n_out=None):
    return src.new_empty((idx.numel() if n_out is None else n_out,) + tuple(src.shape[1:]))


_resize_plans: dict = {}


def resize_plan(H: int, W: int, size: int, device):
    """(hvlm_resize_plan, device coefficient table) for decoded frames of H x W, cached per geometry and device: geometry
    as transformers derives it (shortest edge -> size, centre crop size x size), Pillow's fixed-point bicubic taps."""
    key = (H, W, size, str(device))
    hit = _resize_plans.get(key)
    if hit is None:
        lib = L.lib()
        plan = L.ResizePlan()
        L.check(lib.hvlm_resize_plan_host(H, W, size, size, C.byref(plan)), "hvlm_resize_plan_host")
        tab = (C.c_int32 * plan.table_ints)()
        L.check(lib.hvlm_resize_tables_host(C.byref(plan), tab), "hvlm_resize_tables_host")
        hit = (plan, torch.frombuffer(tab, dtype=torch.int32).clone().to(device))
        _resize_plans[key] = hit
    return hit


@torch.library.custom_op("hvlm::resize_center_crop_u8", mutates_args=())
def resize_center_crop_u8(frames: torch.Tensor, size: int = 224) -> torch.Tensor:
    """Decoded frames uint8 [N,H,W,3] -> uint8 [N,size,size,3]: CLIPImageProcessor's resize (shortest edge -> size, PIL
    BICUBIC, bit-identical to Pillow) + centre crop, one launch; feed the result to the tower's uint8 path."""
    _need_cuda(frames)
    ensure_device()
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError(f"expected uint8 frames [N,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    N, H, W, _ = frames.shape
    plan, table = resize_plan(H, W, size, frames.device)
    out = torch.empty(N, size, size, 3, dtype=torch.uint8, device=frames.device)
    L.check(L.lib().hvlm_resize_crop_u8(_p(frames), N, C.byref(plan), _p(table), _p(out), _stream()), "hvlm_resize_crop_u8")
    return out


@resize_center_crop_u8.register_fake
def _(frames, size=224):
    return frames.new_empty(frames.shape[0], size, size, 3)


def pad_square_u8(frames: torch.Tensor, background=(122, 116, 104)) -> torch.Tensor:
    """expand2square (hoi_forecast/dataset/video_utils.py:13-25): uint8 [N,H,W,3] -> [N,S,S,3], S = max(H,W), the frame
    centred on a canvas of ``background`` (default: int(255 * OPENAI_CLIP_MEAN), what load_image passes for 'pad')."""
    _need_cuda(frames)
    ensure_device()
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError(f"expected uint8 frames [N,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    N, H, W, _ = frames.shape
    if H == W:
        return frames
    S = max(H, W)
    out = torch.empty(N, S, S, 3, dtype=torch.uint8, device=frames.device)
    r, g, b = (int(v) for v in background)
    L.check(L.lib().hvlm_pad_square_u8(_p(frames), N, H, W, _p(out), r, g, b, _stream()), "hvlm_pad_square_u8")
    return out
