"""Stand-alone launches of the tower GEMMs at the 100-frame M, plain vs folded-LayerNorm variants (for ncu and for timing).
usage: python tools/fold_probe.py [time|once]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hvlm_b200
from hvlm_b200 import _lib as _L
if os.environ.get("HVLM_PROBE_LIB"):          # a library variant built with -DHVLM_P_* (see tools/build_probe_libs.sh)
    _L.LIB_PATH = os.path.abspath(os.environ["HVLM_PROBE_LIB"])
from hvlm_b200 import ops
from hvlm_b200.weights import fold_layernorm
from oracle import synth

DEV = "cuda:0"
M = 25700
mode = sys.argv[1] if len(sys.argv) > 1 else "time"


def block_stats(x):
    xb = x.float().reshape(x.shape[0], 8, 128)
    return torch.stack([xb.sum(-1), (xb * xb).sum(-1)], -1).contiguous()


x = (torch.randn(M, 1024, device=DEV) * 1.5 + 0.5)
xb = x.to(torch.bfloat16)
stats = block_stats(x)
g = synth.gen("g", (1024,), 0.2, 5, mean=1.0)
beta = synth.gen("beta", (1024,), 0.1, 6)
cases = {}
for name, N in (("qkv", 3072), ("fc1", 4096)):
    W = synth.gen(name + ".W", (N, 1024), 1024 ** -0.5, 5)
    b = synth.gen(name + ".b", (N,), 0.3, 5)
    w_f, c, b_f = fold_layernorm(W, b, g, beta)
    cases[name] = (W.to(torch.bfloat16).to(DEV), b.to(DEV), w_f.to(DEV), c.to(DEV), b_f.to(DEV))
a1 = torch.randn(M, 1024, device=DEV).to(torch.bfloat16)
a4 = torch.randn(M, 4096, device=DEV).to(torch.bfloat16)
w1 = (torch.randn(1024, 1024, device=DEV) / 32).to(torch.bfloat16)
w4 = (torch.randn(1024, 4096, device=DEV) / 64).to(torch.bfloat16)
bo = torch.randn(1024, device=DEV)
h = torch.randn(M, 1024, device=DEV)
f1out = torch.empty(M, 4096, dtype=torch.bfloat16, device=DEV)

runs = [
    ("qkv plain", lambda: ops.vit_qkv(xb, cases["qkv"][0], cases["qkv"][1], 100)),
    ("qkv fold", lambda: ops.gemm_ln_fold(xb, stats, *cases["qkv"][2:], qkv_hm=True)),
    ("fc1 plain", lambda: ops.gemm(xb, cases["fc1"][0], cases["fc1"][1], epilogue="quick_gelu", out=f1out)),
    ("fc1 fold", lambda: ops.gemm_ln_fold(xb, stats, *cases["fc1"][2:], epilogue="quick_gelu")),
    ("out_proj plain", lambda: ops.gemm(a1, w1, bo, epilogue="residual", resid=h, out=h)),
    ("out_proj fold", lambda: ops.gemm_resid_stats(a1, w1, bo, h)),
    ("fc2 plain", lambda: ops.gemm(a4, w4, bo, epilogue="residual", resid=h, out=h)),
    ("fc2 fold", lambda: ops.gemm_resid_stats(a4, w4, bo, h)),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, fn in runs:
    if mode == "once":
        fn()
        torch.cuda.synchronize()
        continue
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:16s} median {ts[len(ts) // 2]:7.1f} us  min {ts[0]:7.1f}")
