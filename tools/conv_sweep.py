"""Tuning sweep of the tensor-core convolution on representative layer shapes (GPU box only).
Times conv_umma_kernel through the library's per-launch CUDA-event profiler for several (KC, TPS, MT)."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from premvos_b200 import _lib, ops  # noqa: E402

SHAPES = {
    # name: (N, Cin, H, W, Cout, k, stride, dil)
    "dc_conv1": (4, 565, 112, 256, 128, 3, 1, 1),
    "conv2_0": (4, 117, 112, 256, 128, 3, 1, 1),
    "conv2_4": (4, 533, 112, 256, 32, 3, 1, 1),
    "head2": (4, 565, 112, 256, 10, 3, 1, 1),
    "dc_conv3": (4, 128, 112, 256, 128, 3, 1, 4),
    "dc_conv5": (4, 96, 112, 256, 64, 3, 1, 16),
    "conv1aa": (8, 16, 224, 512, 16, 3, 1, 1),
    "conv3a": (8, 32, 112, 256, 64, 3, 2, 1),
    "conv6_4": (4, 529, 7, 16, 32, 3, 1, 1),
    "conv4_2": (4, 437, 28, 64, 96, 3, 1, 1),
    "conv2_4h": (16, 533, 112, 256, 48, 3, 1, 1),
    "dc_conv4": (16, 128, 112, 256, 96, 3, 1, 8),
    "dc_conv3b": (16, 128, 112, 256, 128, 3, 1, 4),
    "conv2_3": (16, 469, 112, 256, 64, 3, 1, 1),
    "xc_mid": (20, 728, 25, 25, 728, 1, 1, 1),
    "xc_entry": (20, 128, 193, 193, 128, 1, 1, 1),
    "xc_exit": (20, 1536, 25, 25, 2048, 1, 1, 1),
    "head_c3": (100, 512, 7, 7, 2048, 1, 1, 1),
    "head_c2": (100, 512, 7, 7, 512, 3, 1, 1),
    "res1x1": (1, 1024, 46, 83, 256, 1, 1, 1),
    "res1x1b": (1, 256, 46, 83, 1024, 1, 1, 1),
}


def run(name, env):
    N, Cin, H, W, Cout, k, stride, dil = SHAPES[name]
    for key in ("PREMVOS_KC", "PREMVOS_TPS", "PREMVOS_MT", "PREMVOS_NACC", "PREMVOS_DBG", "PREMVOS_BUDGET_KB", "PREMVOS_HALO"):
        os.environ.pop(key, None)
    os.environ.update({k2: str(v) for k2, v in env.items() if v is not None})
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, k, k) * 0.05
    pad = dil * (k // 2)
    try:
        ops.conv2d(x, w, None, stride, dil, (pad, pad, pad, pad), 0.1)
        _lib.profile_begin()
        for _ in range(3):
            ops.conv2d(x, w, None, stride, dil, (pad, pad, pad, pad), 0.1)
        prof = _lib.profile_end()
    except _lib.PremvosError as e:
        return None, str(e)[:60]
    r = prof["conv_umma_kernel"]
    us = r["ms"] * 1e3 / r["launches"]
    return us, "%.0f TF/s" % (r["flops"] / r["launches"] / us / 1e6)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    for name in names:
        print("==", name, SHAPES[name])
        k = SHAPES[name][5]
        for kc, tps in ([(2, 3), (2, 9), (4, 3), (8, 1)] if k == 3 else [(2, None), (4, None), (8, None)]):
            for mt in (1, 2):
                us, note = run(name, {"PREMVOS_KC": kc, "PREMVOS_TPS": tps, "PREMVOS_MT": mt})
                print("  KC=%s TPS=%s MT=%s NACC=0: %s %s" % (kc, tps, mt, "%.1f us" % us if us else "--", note), flush=True)
