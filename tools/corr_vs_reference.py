"""BASELINE config C1 on the GPU box: the product's correlation (premvos_corr_forward, fp32 NCHW in / out) timed beside
the REFERENCE's own CUDA kernels (oracle/_ref/libcorr_reference.so = corr_cuda_kernel.cu compiled unchanged + the
corr_cuda.c:52-78 call sequence: 3 memsets, 2 blob_rearrange, CorrelateData, scratch alloc/free), same inputs, CUDA
events, inputs rotated over > L2 worth of buffers.  Prints one JSON line per shape.  A measurement tool (it imports
oracle/ only for the reference arm); not part of the product."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_corr
from premvos_b200 import pwc

PEAK = 6548.8
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]

SHAPES = [(1, 196, 7, 16), (1, 128, 14, 32), (1, 96, 28, 64), (1, 64, 56, 128), (1, 32, 112, 256), (1, 32, 256, 256),
          (16, 32, 112, 256), (16, 64, 56, 128)]


def time_fn(fn, bufs, iters):
    for i in range(3):
        fn(*bufs[i % len(bufs)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*bufs[i % len(bufs)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


for shape in SHAPES:
    B, C, H, W = shape
    nbytes = 4 * B * H * W * (2 * C + 81)
    nb = max(2, min(64, int(300e6 // nbytes) + 1))
    g = torch.Generator(device="cuda").manual_seed(0)
    bufs = [(torch.randn(*shape, device="cuda", generator=g), torch.randn(*shape, device="cuda", generator=g)) for _ in range(nb)]
    iters = 200 if nbytes < 50e6 else 50
    t_ours = time_fn(lambda a, b: pwc.correlation_forward(a, b), bufs, iters)
    line = {"shape": list(shape), "algorithmic_MB": nbytes / 1e6, "ours_us": t_ours * 1e6, "ours_GBps": nbytes / t_ours / 1e9,
            "ours_frac_of_hbm_peak": nbytes / t_ours / 1e9 / PEAK}
    if ref_corr.available():
        t_ref = time_fn(lambda a, b: ref_corr.corr_cuda_forward(a, b), bufs, iters)
        err = float((pwc.correlation_forward(*bufs[0]) - ref_corr.corr_cuda_forward(*bufs[0])).abs().max())
        line.update({"reference_kernel_us": t_ref * 1e6, "reference_GBps": nbytes / t_ref / 1e9, "speedup": t_ref / t_ours,
                     "max_abs_diff": err})
    print(json.dumps(line), flush=True)
