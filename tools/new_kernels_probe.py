"""Exercise the round-2 kernels at bench sizes (for `ncu --set full -k regex:...`): frame de-duplication on a tiled
100-frame bf16 clip, the row gather, the resize / crop kernel on 100 decoded 256x456 frames, the splice at 64 clips."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hvlm_b200  # noqa: E402
from hvlm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
clip = torch.randn(10, 3, 224, 224, device=dev).to(torch.bfloat16).repeat(10, 1, 1, 1)
dec = torch.randint(0, 256, (100, 256, 456, 3), device=dev, dtype=torch.uint8)
for _ in range(3):
    fmap, rep, n = ops.frame_dedup(clip)
    ops.gather_rows(clip, rep, 10)
    ops.resize_center_crop_u8(dec)
torch.cuda.synchronize()
print("ok", int(n.item()))
