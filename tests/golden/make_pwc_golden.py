#!/usr/bin/env python
"""Generate tests/golden/pwc_golden_*.npz by running the REFERENCE's own models/PWCNet.py.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_pwc_golden.py

What is and is not the reference here:
  * ``models/PWCNet.py`` is imported unmodified from /root/reference (network definition,
    forward wiring, concat order, flow scaling, deconvs, context net: PWCNet.py:43-272).
  * ``correlation_package.modules.corr.Correlation`` is a shim around
    ``oracle.pwc_oracle.correlation_forward``: the reference's correlation is a CUDA-only cffi
    extension for torch 0.2 (``torch.utils.ffi``/THC are gone) and its CPU file is a stub
    (src/corr.c:3-16).  The shim's arithmetic is pinned separately by the reference's own KAT.
  * ``PWCDCNet.warp`` is replaced by a copy that differs in exactly two tokens: no ``.cuda()``
    (PWCNet.py:166 hard-codes it) and ``align_corners=True`` passed explicitly (torch 0.2's
    grid_sample implemented that convention; modern torch changed the default).
Weights are the seeded synthetic state_dict of premvos_b200.synth (no checkpoint is available
offline); they are loaded with the reference's own load_state_dict, which also pins the key names.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/code/optical_flow_net-PWC-Net"

from oracle import pwc_oracle as O  # noqa: E402
from premvos_b200 import synth  # noqa: E402


def install_correlation_shim():
    class Correlation(torch.nn.Module):
        def __init__(self, pad_size=None, kernel_size=None, max_displacement=None, stride1=None,
                     stride2=None, corr_multiply=None):
            super().__init__()
            self.a = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)

        def forward(self, x1, x2):
            out = O.correlation_forward(x1.detach().numpy(), x2.detach().numpy(), *self.a)
            return torch.from_numpy(out)

    pkg = types.ModuleType("correlation_package")
    mods = types.ModuleType("correlation_package.modules")
    corr = types.ModuleType("correlation_package.modules.corr")
    corr.Correlation = Correlation
    pkg.modules = mods
    mods.corr = corr
    sys.modules["correlation_package"] = pkg
    sys.modules["correlation_package.modules"] = mods
    sys.modules["correlation_package.modules.corr"] = corr


def patched_warp(self, x, flo):
    # PWCNet.py:140-176 with `.cuda()` removed and align_corners=True made explicit
    B, C, H, W = x.size()
    xx = torch.arange(0, W).view(1, -1).repeat(H, 1)
    yy = torch.arange(0, H).view(-1, 1).repeat(1, W)
    xx = xx.view(1, 1, H, W).repeat(B, 1, 1, 1)
    yy = yy.view(1, 1, H, W).repeat(B, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float()
    vgrid = grid + flo
    vgrid[:, 0, :, :] = 2.0 * vgrid[:, 0, :, :] / max(W - 1, 1) - 1.0
    vgrid[:, 1, :, :] = 2.0 * vgrid[:, 1, :, :] / max(H - 1, 1) - 1.0
    vgrid = vgrid.permute(0, 2, 3, 1)
    output = torch.nn.functional.grid_sample(x, vgrid, align_corners=True)
    mask = torch.ones(x.size())
    mask = torch.nn.functional.grid_sample(mask, vgrid, align_corners=True)
    mask[mask < 0.9999] = 0
    mask[mask > 0] = 1
    return output * mask


def main():
    install_correlation_shim()
    sys.path.insert(0, REF)
    import models  # the reference package

    outdir = os.path.dirname(os.path.abspath(__file__))
    for tag, (h, w, batch, wseed, iseed) in {"a": (64, 128, 1, 0, 1), "b": (128, 192, 2, 3, 7)}.items():
        net = models.pwc_dc_net(None)
        net.warp = types.MethodType(patched_warp, net)
        sd = synth.pwc_synthetic_state_dict(wseed)
        ref_sd = net.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), "state_dict key order differs from reference"
        for k in sd:
            assert tuple(ref_sd[k].shape) == sd[k].shape, k
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        net.eval()
        x = synth.synthetic_pwc_input(batch, h, w, seed=iseed)
        with torch.no_grad():
            flow2 = net(torch.from_numpy(x)).numpy()
        # restatement vs reference, right here
        mine = O.pwc_forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x)).numpy()
        err = np.abs(mine - flow2).max() / np.abs(flow2).max()
        print("golden %s: flow2 %s |max| %.4f  oracle-vs-reference rel err %.3e" % (tag, flow2.shape, np.abs(flow2).max(), err))
        assert err < 1e-5
        np.savez_compressed(os.path.join(outdir, "pwc_golden_%s.npz" % tag), flow2=flow2.astype(np.float32),
                            h=h, w=w, batch=batch, weight_seed=wseed, input_seed=iseed,
                            x_checksum=np.float64(x.astype(np.float64).sum()),
                            w_checksum=np.float64(sum(float(v.astype(np.float64).sum()) for v in sd.values())))


if __name__ == "__main__":
    main()
