"""Per-stage device times of the tower for a small batch (latency regime).  usage: python tools/stage_probe.py [frames ...]"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import hvlm_b200  # noqa: E402
from hvlm_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    tower = bench.build_tower(bench.clip_state_dict(23)).to(dev)
    for n in [int(a) for a in sys.argv[1:]] or [10]:
        px = torch.randn(n, 3, 224, 224, device=dev).to(torch.bfloat16)
        for _ in range(5):
            tower.forward_hidden(px)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            tower.forward_hidden(px)
        e1.record()
        torch.cuda.synchronize()
        total = e0.elapsed_time(e1) / 20
        ops.profile_enable(True)
        for _ in range(3):
            tower.forward_hidden(px)
        prof = ops.profile_collect()
        ops.profile_enable(False)
        print(f"frames={n} tower forward {total:.3f} ms; per-launch avg us (with event bracketing): " +
              ", ".join(f"{k} {t / c * 1e3:.1f}x{c // 3}" for k, (t, c) in prof.items()))


if __name__ == "__main__":
    main()
