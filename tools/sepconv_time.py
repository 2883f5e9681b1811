"""Times the fused separable convolution (premvos_sepconv2d_forward) on the entry-flow shapes (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from premvos_b200 import _lib, ops

for (N, C, H, W, Cout) in ((40, 128, 193, 193, 128), (40, 64, 193, 193, 128), (40, 128, 97, 97, 128)):
    x = torch.randn(N, C, H, W, device="cuda")
    dw, db = torch.randn(C, 3, 3) / 3, torch.randn(C) * 0.1
    pw, pb = torch.randn(Cout, C) / C ** 0.5, torch.randn(Cout) * 0.1
    os.environ["PREMVOS_CONV_REPEAT"] = "4"
    ops.sepconv2d(x, dw, db, pw, pb, False, False, 0.0)
    _lib.profile_begin()
    ops.sepconv2d(x, dw, db, pw, pb, False, False, 0.0)
    p = _lib.profile_end()["sepconv_fused_kernel"]
    us = p["ms"] * 1e3 / p["launches"]
    print("n%d c%d %dx%d -> %d: %.1f us per launch, %.0f GB/s of algorithmic bytes (in + out), %.1f TF/s" % (
        N, C, H, W, Cout, us, p["bytes"] / p["launches"] / us / 1e3, p["flops"] / p["launches"] / us / 1e6))
    del x
