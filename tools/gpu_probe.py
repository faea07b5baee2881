"""GPU bring-up probe: runs one stage of the CUDA path against the oracle and prints error statistics.
Usage (on the GPU box):  timeout 300 python tools/gpu_probe.py <stage> ;  stages: simple gemm attn vit bench
Each stage is a separate process so that a hang in one kernel cannot take the others down."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hvlm_b200  # noqa: E402
from hvlm_b200 import _lib as _L  # noqa: E402
if os.environ.get("HVLM_PROBE_LIB"):          # a library variant built by tools/build_probe_libs.sh
    _L.LIB_PATH = os.path.abspath(os.environ["HVLM_PROBE_LIB"])
from hvlm_b200 import ops  # noqa: E402
from oracle import restate, synth  # noqa: E402

dev = torch.device("cuda:0")
RES = {}


def stats(name, a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    d = (a - b).abs()
    r = dict(max_abs=float(d.max()), ref_max=float(b.abs().max()), rel=float(d.max() / b.abs().max().clamp_min(1e-30)),
             mean_rel=float(d.mean() / b.abs().mean().clamp_min(1e-30)), nan=int(torch.isnan(a).sum()))
    RES[name] = r
    print(f"{name:40s} rel={r['rel']:.3e} mean_rel={r['mean_rel']:.3e} max_abs={r['max_abs']:.3e} nan={r['nan']}", flush=True)
    return r


def stage_simple():
    # layernorm
    x = synth.gen("ln.x", (777, 1024), 2.0, 1, 0.3)
    g = synth.gen("ln.g", (1024,), 0.2, 1, 1.0)
    b = synth.gen("ln.b", (1024,), 0.1, 1)
    ref = torch.nn.functional.layer_norm(x, (1024,), g, b, 1e-5)
    stats("layernorm_f32", ops.layernorm_1024(x.to(dev), g.to(dev), b.to(dev), torch.float32), ref)
    stats("layernorm_bf16", ops.layernorm_1024(x.to(dev), g.to(dev), b.to(dev), torch.bfloat16), ref)
    # pooling
    for t, C, dt in ((100, 1024, torch.float32), (10, 4096, torch.float32), (2, 64, torch.float32), (100, 1024, torch.bfloat16)):
        tok = synth.gen(f"pool{t}", (2, t, 256, C), 1.0, 2)
        for mode in ("temporal_spatial_pool", "spatial_pool", "temporal", "spatial", "temporal_spatial"):
            ref = restate.pool_tokens(tok.to(dt).float(), mode)
            out = ops.pool_tokens(tok.to(dev).to(dt), mode, torch.float32)
            stats(f"pool_{mode}_t{t}_C{C}_{str(dt)[6:]}", out, ref)
    dout = synth.gen("pooldout", (2, 266, 512), 1.0, 3)
    ref = restate.pool_tokens_backward(dout, 10)
    stats("pool_bwd", ops.pool_slowfast_bwd(dout.to(dev), 10, 0, False), ref)
    # in-place read of the tower layout: frame_stride 257, row offset 1
    hid = synth.gen("hid", (6, 257, 1024), 1.0, 4)
    ref = restate.pool_tokens(hid[:, 1:].reshape(2, 3, 256, 1024), "temporal_spatial_pool")
    stats("pool_strided", ops.pool_slowfast(hid.to(dev), 2, 3, 257, 1, 0, False), ref)
    # gather
    hidden = synth.gen("gh", (4, 40, 64), 1.0, 41)
    labels = torch.full((4, 40), -100, dtype=torch.int64)
    labels[0, 30:34] = 32100
    labels[1, [3, 9, 17, 39]] = 32100
    labels[3, 1:5] = 32100
    ro, rv, rr = restate.gather_hand_traj(hidden, labels)
    o, v, r, c = ops.hand_gather(hidden.to(dev), labels.to(dev), 32100)
    stats("gather", o, ro)
    print("gather valid/rows/counts eq:", torch.equal(v.cpu(), rv), torch.equal(r.cpu(), rr), c.tolist())
    RES["gather_exact"] = bool(torch.equal(o.cpu(), ro) and torch.equal(v.cpu(), rv) and torch.equal(r.cpu(), rr))
    # splice
    from hvlm_b200 import _lib as L
    D = 256
    table = synth.embed_table(D)
    vis = synth.gen("vis", (3, 356, D), 1.0, 5)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=3, seed=22, ragged=True)
    rm, re_, rl = restate.splice(ids, mask, labels, vis, table, "handsonvlm", future_hands=fh)
    idsd = ids.to(dev)
    counts = ops.splice_count(idsd)
    Lout = ids.shape[1] - 1 + 356
    plan = ops.splice_plan(idsd, counts, 356, 3, Lout, table.shape[0], L.SPLICE_HANDSONVLM, 1, 4)
    e, l2, m2 = ops.splice_gather(plan[0], plan[1], plan[2], plan[3], idsd, labels.to(dev), mask.to(dev), table.to(dev),
                                  vis.to(dev), None, fh.to(dev), L.SPLICE_HANDSONVLM)
    stats("splice_embeds", e, re_)
    RES["splice_exact"] = bool(torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm))
    print("splice labels/mask eq:", torch.equal(l2.cpu(), rl), torch.equal(m2.cpu(), rm), "status", int(plan[4].item()),
          "counts", counts.tolist())


def stage_gemm():
    torch.manual_seed(0)
    cases = [(128, 256, 64), (128, 256, 1024), (256, 512, 128), (300, 1024, 1024), (25700, 3072, 1024),
             (1000, 4096, 1024), (1000, 1024, 4096), (356, 4096, 1024), (356, 5120, 1024), (200, 128, 64),
             (4096, 1024, 1424)]
    for (M, N, K) in cases:
        a = synth.gen(f"A{M}", (M, K), 1.0, 1).to(torch.bfloat16)
        w = synth.gen(f"W{N}", (N, K), K ** -0.5, 1).to(torch.bfloat16)
        bias = synth.gen(f"b{N}", (N,), 0.5, 1)
        ad, wd, bd = a.to(dev), w.to(dev), bias.to(dev)
        ref = (ad.float() @ wd.float().t() + bd).cpu()
        stats(f"gemm_f32_{M}x{N}x{K}", ops.gemm(ad, wd, bd, out_dtype=torch.float32), ref)
        if M <= 1000:
            stats(f"gemm_bf16_{M}x{N}x{K}", ops.gemm(ad, wd, bd, out_dtype=torch.bfloat16), ref)
            stats(f"gemm_gelu_{M}x{N}x{K}", ops.gemm(ad, wd, bd, epilogue="quick_gelu", out_dtype=torch.float32),
                  restate.quick_gelu(ref))
            res = synth.gen(f"r{M}", (M, N), 1.0, 2).to(dev)
            stats(f"gemm_resid_{M}x{N}x{K}", ops.gemm(ad, wd, bd, epilogue="residual", resid=res, out_dtype=torch.float32),
                  ref + res.cpu())
    # timing of the big ones
    for (M, N, K) in ((25700, 3072, 1024), (25700, 4096, 1024), (25700, 1024, 4096), (25700, 1024, 1024)):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.gemm(a, w, bias, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, bias, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2 * M * N * K / ms / 1e9
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(a, w)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        RES[f"time_gemm_{M}x{N}x{K}"] = dict(ms=ms, tflops=tf, cublas_ms=ms2, cublas_tflops=2 * M * N * K / ms2 / 1e9)
        print(f"gemm {M}x{N}x{K}: {ms:.3f} ms {tf:.0f} TF/s | cublas {ms2:.3f} ms {2*M*N*K/ms2/1e9:.0f} TF/s", flush=True)


def _attn_ref(q, k, v):
    # q,k,v [F,16,257,64] fp32 (q pre-scaled)
    s = q @ k.transpose(-1, -2)
    return torch.softmax(s, -1) @ v


def _attn_ref_from_qkv(qkv_hm):
    """qkv [48, F*257, 64] -> [F*257,1024] fp32 reference."""
    F = qkv_hm.shape[1] // 257
    t = qkv_hm.float().cpu().reshape(3, 16, F, 257, 64)
    s = t[0] @ t[1].transpose(-1, -2)                                   # [16,F,257,257]
    return (torch.softmax(s, -1) @ t[2]).permute(1, 2, 0, 3).reshape(F * 257, 1024)


def _time(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stage_attn():
    F = 3
    y = synth.gen("attn.y", (F * 257, 1024), 1.0, 1).to(torch.bfloat16)
    w = synth.gen("attn.w", (3072, 1024), 1024 ** -0.5, 1).to(torch.bfloat16)
    w[:1024] *= 0.125
    b = synth.gen("attn.b", (3072,), 0.1, 1)
    qkv = ops.vit_qkv(y.to(dev), w.to(dev), b.to(dev), F)
    ref = (y.float() @ w.float().t() + b).reshape(F * 257, 48, 64).permute(1, 0, 2)
    stats("qkv_gemm_cb_major", qkv, ref)
    stats("attention", ops.vit_attention(qkv), _attn_ref_from_qkv(qkv))
    qkv2 = qkv.clone()
    qkv2[:16] *= 6
    stats("attention_sharp", ops.vit_attention(qkv2), _attn_ref_from_qkv(qkv2))
    Fb = 100
    qb = torch.randn(48, Fb * 257, 64, device=dev).to(torch.bfloat16)
    qb[:16] *= 0.3
    ob = ops.vit_attention(qb)
    ms = _time(lambda: ops.vit_attention(qb))
    RES["time_attn_100f"] = dict(ms=ms, tflops=Fb * 16 * 4 * 257 * 257 * 64 / ms / 1e9)
    print(f"attention 100 frames: {ms * 1e3:.1f} us", flush=True)
    t = qb.reshape(3, 16, Fb, 257, 64).float()
    stats("attention_big_vs_sdpa", ob.reshape(Fb, 257, 16, 64).permute(2, 0, 1, 3),
          torch.nn.functional.scaled_dot_product_attention(t[0], t[1], t[2], scale=1.0))
    yb = torch.randn(Fb * 257, 1024, device=dev).to(torch.bfloat16)
    ms = _time(lambda: ops.vit_qkv(yb, w.to(dev), b.to(dev), Fb))
    print(f"qkv gemm head-major 100 frames: {ms * 1e3:.1f} us  {2 * 25700 * 3072 * 1024 / ms / 1e9:.0f} TF/s", flush=True)


def stage_vit():
    from hvlm_b200.tower import CLIPVisionTower
    for profile, nl, nf in (("strong", 2, 2), ("strong", 23, 2), ("hf", 23, 3)):
        sd = synth.clip_state_dict(synth.VIT_L14, 0, profile, n_layers=nl)
        px = synth.pixels((nf, 3, 224, 224), seed=3)
        tower = CLIPVisionTower("synthetic", None, delay_load=True)
        tower.select_layer = nl
        tower.load_model(sd)
        tower = tower.to(dev)
        t0 = time.time()
        hid = tower.forward_hidden(px.to(dev))
        torch.cuda.synchronize()
        ref = restate.vit_hidden(px, sd, nl)
        stats(f"vit_{profile}_{nl}L_hidden", hid, ref)
        emu = restate.vit_hidden(px, sd, nl, emulate="bf16")
        stats(f"vit_{profile}_{nl}L_vs_bf16emu", hid, emu)
        stats(f"vit_{profile}_{nl}L_emu_vs_fp32", emu, ref)
        feats = tower(px.to(dev))
        stats(f"vit_{profile}_{nl}L_feats", feats, ref[:, 1:])


def stage_bench():
    from hvlm_b200.tower import CLIPVisionTower
    NF = int(os.environ.get("PROBE_FRAMES", "100"))
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
    tower = CLIPVisionTower("synthetic", None, delay_load=True)
    tower.load_model(sd)
    tower = tower.to(dev)
    px = torch.randn(NF, 3, 224, 224, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        tower.forward_hidden(px)
    torch.cuda.synchronize()
    reps = []
    for _ in range(5):                      # 5 repeats of 10 forwards: report the median repeat (run-to-run noise ~2 %)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            tower.forward_hidden(px)
        e1.record()
        torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1) / 10)
    ms = sorted(reps)[len(reps) // 2]
    print("  repeats (ms):", " ".join(f"{r:.3f}" for r in reps), flush=True)
    RES["time_vit_%df" % NF] = dict(ms=ms, frames_per_s=NF / ms * 1e3, tflops=NF * 155.29 / ms)
    print(f"ViT {NF} frames: {ms:.3f} ms  {NF/ms*1e3:.0f} frames/s  {NF*155.29/ms:.0f} TF/s", flush=True)
    ops.profile_enable(True)
    for _ in range(3):
        tower.forward_hidden(px)
    prof = ops.profile_collect()
    ops.profile_enable(False)
    flops = dict(qkv_gemm=2 * NF * 257 * 3072 * 1024, outproj_gemm=2 * NF * 257 * 1024 * 1024, fc1_gemm=2 * NF * 257 * 4096 * 1024,
                 fc2_gemm=2 * NF * 257 * 4096 * 1024, attention=NF * 16 * 4 * 257 * 257 * 64)
    for k, (t, n) in prof.items():
        extra = f"  {flops[k] / (t / n) / 1e9:7.0f} TF/s" if k in flops else ""
        print(f"  {k:14s} {t / 3:8.3f} ms/fwd  n={n // 3:3d}  avg {t / n * 1e3:7.1f} us{extra}", flush=True)
    RES["profile"] = {k: dict(ms_per_fwd=t / 3, launches=n // 3) for k, (t, n) in prof.items()}


def stage_attn_trace():
    import ctypes as C
    from hvlm_b200 import _lib as L
    lib = L.lib()
    Fb = 100
    qb = torch.randn(48, Fb * 257, 64, device=dev).to(torch.bfloat16)
    qb[:16] *= 0.3
    out = torch.empty(Fb * 257, 1024, dtype=torch.bfloat16, device=dev)
    grid = min(Fb * 16, 2 * 148)
    trace = torch.zeros(grid, 5, 4, 16, dtype=torch.int64, device=dev)
    fn = lib.hvlm_debug_attention_trace
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int]
    for _ in range(2):
        ops.vit_attention(qb)
    dbg = int(os.environ.get("ATTN_DBG", "0"))
    rc = fn(qb.data_ptr(), out.data_ptr(), Fb, trace.data_ptr(), torch.cuda.current_stream().cuda_stream, dbg)
    torch.cuda.synchronize()
    t = trace.cpu()
    names_m = ["item start", "qk_full", "T0 o_read ok", "T0 S issued", "T0 p_full", "T0 PV issued", "", "",
               "T1 o_read ok", "T1 S issued(+qk next)", "T1 p_full", "T1 PV issued", "", "", "v free"]
    names_s = ["item start", "qk_full", "T0 wait s", "T0 s_full", "T0 pass1 done", "T0 p arrived", "T0 o_full", "T0 o read",
               "T0 epi done", "T1 wait s", "T1 s_full", "T1 pass1 done", "T1 p arrived", "T1 o_full", "T1 o read", "T1 epi done"]
    # absolute time of each softmax warp's p arrival (relative to warp 1's item start): shows stragglers
    for cta in (0, 1, 150):
        for it in (1, 2):
            b0 = int(t[cta, 1, it][0])
            print(f"cta{cta} item{it} p-arrive T0/T1 per warp:",
                  [(int(t[cta, w, it][5]) - b0, int(t[cta, w, it][12]) - b0) for w in (1, 2, 3, 4)],
                  "epi done:", [(int(t[cta, w, it][8]) - b0, int(t[cta, w, it][15]) - b0) for w in (1, 2, 3, 4)], flush=True)
    for cta in (0, 150):
        for role, names in ((0, names_m), (1, names_s), (2, names_s), (3, names_s), (4, names_s)):
            for it in (1,):
                row = t[cta, role, it]
                base = int(t[cta, role, it][0])
                prev = base
                line = []
                for i, nm in enumerate(names):
                    v = int(row[i])
                    if nm and v:
                        line.append(f"{nm}:+{v - prev}")
                        prev = v
                print(f"cta{cta} {'MMA' if role == 0 else 'SMX%d' % role} item{it} total={prev - base}: " + " | ".join(line), flush=True)


def stage_gemm_epi():
    """epilogue cost by variant, ViT shapes"""
    for (M, N, K) in ((25700, 1024, 1024), (25700, 1024, 4096), (25700, 4096, 1024), (25700, 3072, 1024)):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        o16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        o32 = torch.zeros(M, N, dtype=torch.float32, device=dev)
        fl = 2 * M * N * K
        for name, fn in (("bf16 store", lambda: ops.gemm(a, w, bias, out=o16)),
                         ("bf16 gelu", lambda: ops.gemm(a, w, bias, epilogue="quick_gelu", out=o16)),
                         ("f32 store", lambda: ops.gemm(a, w, bias, out=o32)),
                         ("f32 reduce-add", lambda: ops.gemm(a, w, bias, epilogue="residual", resid=o32, out=o32))):
            ms = _time(fn, 20)
            print(f"gemm {M}x{N}x{K} {name:16s}: {ms * 1e3:7.1f} us  {fl / ms / 1e9:6.0f} TF/s", flush=True)


def stage_hbm():
    """HBM-bound kernels at sizes larger than L2: achieved GB/s (algorithmic bytes / CUDA-event time)."""
    from hvlm_b200 import _lib as L
    # pooling, reference order: fp32 [B,100,256,4096] -> [B,356,4096]
    for B, C, dt in ((1, 4096, torch.float32), (4, 4096, torch.float32), (4, 4096, torch.bfloat16), (4, 1024, torch.float32)):
        tok = torch.randn(B, 100, 256, C, device=dev, dtype=dt)
        ms = _time(lambda: ops.pool_tokens(tok, "temporal_spatial_pool"), 20)
        by = tok.numel() * tok.element_size() + B * 356 * C * tok.element_size()
        print(f"pool fwd B={B} C={C} {str(dt)[6:]:8s}: {ms * 1e3:7.1f} us  {by / ms / 1e6:7.0f} GB/s", flush=True)
    hid = torch.randn(400, 257, 1024, device=dev)
    ms = _time(lambda: ops.pool_slowfast(hid, 4, 100, 257, 1, 0, True), 20)
    by = 400 * 256 * 1024 * 4 + 4 * 356 * 1024 * 2
    print(f"pool fwd in-place hidden layout B=4: {ms * 1e3:7.1f} us  {by / ms / 1e6:7.0f} GB/s", flush=True)
    dout = torch.randn(4, 356, 4096, device=dev)
    ms = _time(lambda: ops.pool_slowfast_bwd(dout, 100, 0, False), 20)
    print(f"pool bwd B=4 C=4096 f32: {ms * 1e3:7.1f} us  {(4 * 100 * 256 * 4096 * 4) / ms / 1e6:7.0f} GB/s (write)", flush=True)
    # splice
    for B, D in ((16, 4096), (64, 4096), (16, 5120)):
        table = torch.randn(32101, D, device=dev).to(torch.bfloat16)
        vis = torch.randn(B, 356, D, device=dev).to(torch.bfloat16)
        ids = torch.randint(1, 32000, (B, 64), device=dev)
        ids[:, 35] = -200
        ids[:, 56:60] = 32100
        labels = ids.clone()
        mask = torch.ones_like(ids, dtype=torch.bool)
        fh = torch.rand(B, 2, 4, 2, device=dev)
        counts = ops.splice_count(ids)
        Lout = 64 - 1 + 356
        plan = ops.splice_plan(ids, counts, 356, B, Lout, 32101, L.SPLICE_HANDSONVLM, 1, 4)
        fn = lambda: ops.splice_gather(plan[0], plan[1], plan[2], plan[3], ids, labels, mask, table, vis, None, fh,
                                       L.SPLICE_HANDSONVLM)
        hidden = torch.randn(B, Lout, D, device=dev).to(torch.bfloat16)
        lab2 = fn()[1]
        big = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ops.profile_enable(True)
        for _ in range(10):
            big.zero_()                      # flush L2 between iterations (512 MB write)
            torch.cuda.synchronize()         # keep host launch latency out of the event pairs
            fn()
            ops.hand_gather(hidden, lab2, 32100)
        prof = ops.profile_collect()
        ops.profile_enable(False)
        by = B * ((63 + 356) * D * 2 + Lout * (D * 2 + 9))
        t, n = prof["splice"]
        print(f"splice fwd B={B} D={D}: {t / n * 1e3:7.1f} us (device, L2 flushed)  {by / (t / n) / 1e6:7.0f} GB/s", flush=True)
        t, n = prof["gather"]
        print(f"hand gather B={B} D={D}: {t / n * 1e3:7.1f} us (device)", flush=True)
    x = torch.randn(25700, 1024, device=dev)
    g = torch.ones(1024, device=dev)
    b = torch.zeros(1024, device=dev)
    ms = _time(lambda: ops.layernorm_1024(x, g, b), 20)
    print(f"layernorm 25700 rows: {ms * 1e3:7.1f} us  {25700 * 1024 * 6 / ms / 1e6:7.0f} GB/s", flush=True)


def stage_traj():
    """Fused gather + CVAE decoder step vs the same math as eager torch ops (what the reference's loop launches)."""
    from hvlm_b200.traj_decoder import CVAETrajDecoder
    for D in (4096, 5120):
        dec = CVAETrajDecoder(token_dim=D // 2).to(dev).to(torch.bfloat16)
        hl = torch.randn(1, D, device=dev).to(torch.bfloat16)
        z = (2.0 * torch.randn(2, 256, device=dev)).to(torch.bfloat16)
        d = dec.hand_traj_decoder.cvae.dec_MLP

        def eager():
            e = hl.reshape(1, D // 2, 2).permute(0, 2, 1).unsqueeze(2).reshape(-1, D // 2)
            return d(torch.cat((z, e), dim=-1))

        def fused():
            return ops.traj_decode(hl, z, d[0].weight, d[0].bias, d[2].weight, d[2].bias, interleaved=True)

        a, b = eager().float(), fused()
        stats(f"traj step D={D} fused vs eager bf16", b, a)
        for name, fn in (("eager", eager), ("fused", fused)):
            g = torch.cuda.CUDAGraph()
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for _ in range(20):
                    fn()
            ms = _time(g.replay, 10) / 20
            ms_host = _time(fn, 50)
            print(f"traj step D={D} {name}: {ms * 1e3:6.2f} us/step device (graph replay), {ms_host * 1e3:6.1f} us/step "
                  f"with host launch", flush=True)


def stage_twostream():
    """Two half-batches on two streams: does one half's LayerNorm / kernel tails hide under the other half's GEMMs?"""
    import types as _t
    from hvlm_b200.tower import CLIPVisionTower
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
    tw = CLIPVisionTower("synthetic", _t.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    tw.load_model(sd)
    tw = tw.to(dev)
    px = torch.randn(100, 3, 224, 224, device=dev, dtype=torch.bfloat16)
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]

    def run(parts):
        if parts == 1:
            return [tw.forward_hidden(px)]
        cur = torch.cuda.current_stream()
        outs = []
        chunks = px.chunk(parts)
        for st, c in zip(streams, chunks):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(tw.forward_hidden(c))
        for st in streams[:parts]:
            cur.wait_stream(st)
        return outs

    ref = run(1)[0]
    for parts in (1, 2, 1, 2, 4):
        for _ in range(3):
            o = run(parts)
        torch.cuda.synchronize()
        reps = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                o = run(parts)
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1) / 10)
        same = torch.equal(torch.cat(o), ref)
        print(f"parts={parts}: median {sorted(reps)[2]:.3f} ms  ({' '.join(f'{r:.2f}' for r in reps)})  bit-identical={same}", flush=True)


def stage_hf_eager():
    """Context number: the reference's own GPU path for the tower -- HF CLIPVisionModel (24 layers, fp16/bf16 eager, SDPA,
    output_hidden_states=True as clip_encoder.py:44 calls it) on the same 100-frame clip."""
    import transformers
    cfg = transformers.CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                        num_attention_heads=16, image_size=224, patch_size=14, projection_dim=768)
    for dt in (torch.float16, torch.bfloat16):
        m = transformers.CLIPVisionModel(cfg).to(dev).to(dt).eval()
        px = torch.randn(100, 3, 224, 224, device=dev, dtype=dt)
        with torch.no_grad():
            fn = lambda: m(px, output_hidden_states=True).hidden_states[-2][:, 1:]
            ms = _time(fn, 10)
        print(f"HF CLIPVisionModel eager {str(dt)[6:]} 100 frames: {ms:.2f} ms  {100 / ms * 1e3:.0f} frames/s", flush=True)
        RES[f"hf_eager_{str(dt)[6:]}"] = ms
        del m


def stage_stress():
    """Race / nondeterminism hunt: many forwards of the same clip must be bit-identical (tower, pooled tokens, splice)."""
    import types as _t
    from hvlm_b200.tower import CLIPVisionTower
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
    tw = CLIPVisionTower("synthetic", _t.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    tw.load_model(sd)
    tw = tw.to(dev)
    bad = 0
    for n in (100, 37, 3):
        px = torch.randn(n, 3, 224, 224, device=dev, dtype=torch.bfloat16)
        ref = tw.forward_hidden(px).clone()
        assert torch.isfinite(ref).all()
        for i in range(40 if n == 100 else 100):
            out = tw.forward_hidden(px)
            if not torch.equal(out, ref):
                bad += 1
                print(f"MISMATCH n={n} iter={i} max|d|={(out - ref).abs().max().item():.3e}", flush=True)
        print(f"n={n}: done, mismatches so far {bad}", flush=True)
    RES["stress_mismatches"] = bad
    print("stress mismatches:", bad, flush=True)


def stage_latency():
    """Small-batch tower latency (configs[0]: one image): stream launches vs one CUDA-graph replay."""
    import types as _t
    from hvlm_b200.tower import CLIPVisionTower
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
    tw = CLIPVisionTower("synthetic", _t.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    tw.load_model(sd)
    tw = tw.to(dev)
    for n in (1, 2, 4, 10, 20):
        px = torch.randn(n, 3, 224, 224, device=dev)
        fn = lambda: tw(px)
        ms = _time(fn, 20)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            out = fn()
        ms_g = _time(g.replay, 20)
        print(f"tower N={n:3d}: {ms:7.3f} ms stream launches, {ms_g:7.3f} ms graph replay", flush=True)
        RES[f"latency_{n}"] = {"stream_ms": ms, "graph_ms": ms_g}


if __name__ == "__main__":
  for stage in sys.argv[1:]:
    t0 = time.time()
    RES.pop("exception", None)
    try:
        globals()["stage_" + stage]()
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        import traceback
        traceback.print_exc()
        RES["exception"] = repr(e)
    RES["seconds"] = time.time() - t0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"probe_{stage}.json"), "w") as f:
        json.dump(RES, f, indent=1)
