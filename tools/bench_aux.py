"""HBM roofline of the byte-work kernels around the networks (GPU box only): device cv2.resize, mask warp + boxes, flow
post-processing.  Inputs are sized above the 126 MB L2 and rotated; CUDA events on the launching stream; prints one JSON line per
kernel with algorithmic bytes / time against MEASURED_PEAKS.json's HBM figure.
    python tools/bench_aux.py            (under ncu: add `ncu` as argv[1] -> one launch of each between cudaProfilerStart/Stop)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from premvos_b200 import _lib, mergetrack, ops

under_ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm_peak = float(peaks.get("hbm_gbs", 6555.5))


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, alg_bytes, note):
    print(json.dumps({"kernel": name, "ms": ms, "algorithmic_MB": alg_bytes / 1e6, "achieved_GBps": alg_bytes / ms / 1e6,
                      "peak_GBps": hbm_peak, "frac": alg_bytes / ms / 1e6 / hbm_peak, "note": note}), flush=True)


g = torch.Generator(device="cuda").manual_seed(0)
# --- cv2.resize on the device: 128 frames 436x1024x3 -> 448x1024 (flow input) and -> 568x1333 BGR (proposal input) ---
B = 8 if under_ncu else 128
src = torch.randint(0, 256, (2, B, 436, 1024, 3), dtype=torch.uint8, device="cuda", generator=g)
dst_a = torch.empty((B, 448, 1024, 3), dtype=torch.uint8, device="cuda")
dst_b = torch.empty((B, 568, 1333, 3), dtype=torch.uint8, device="cuda")
it = [0]
def rz_a():
    it[0] ^= 1
    ops.resize_linear_u8(src[it[0]], 448, 1024, out=dst_a)
def rz_b():
    it[0] ^= 1
    ops.resize_linear_u8(src[it[0]], 568, 1333, reverse_channels=True, out=dst_b)
# --- mask warp: 40 masks of a 1080x1920 frame (83 MB in, 83 MB out, 16.6 MB flow) ---
n, H, W = (4, 480, 854) if under_ncu else (40, 1080, 1920)
# object-shaped masks (unions of ellipses) and a SMOOTH flow field (translation + slow spatial variation), as a video has them:
# a white-noise flow scatters the 4 taps of neighbouring pixels over separate cache lines and measures the L1 gather rate instead
from premvos_b200 import synth
m_np = synth.synthetic_masks(n, H, W, seed=5)
masks = torch.from_numpy(np.stack([m_np, m_np[::-1].copy()])).cuda()
yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
flow = torch.stack([3.7 + 2.0 * torch.sin(yy / 97.0) + 0.01 * xx, -2.2 + 1.5 * torch.cos(xx / 131.0)], dim=-1).contiguous()
wout = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
bbox = torch.empty((n, 4), device="cuda")
def wm():
    it[0] ^= 1
    mergetrack.warp_masks_device(masks[it[0]], flow, out=wout, bbox=bbox)
# --- flow post-processing: 64 pairs 448x1024 -> 436x1024 ---
FB = 4 if under_ncu else 64
flow2 = torch.randn((2, FB, 2, 112, 256), device="cuda", generator=g)
fout = torch.empty((FB, 436, 1024, 2), device="cuda")
def fp():
    it[0] ^= 1
    mergetrack.flow_postprocess_device(flow2[it[0]], 436, 1024, out=fout)

if under_ncu:
    for f in (rz_a, rz_b, wm, fp):
        f()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for f in (rz_a, rz_b, wm, fp):
        f()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    px = 436 * 1024 * 3
    report("resize_linear_u8_kernel 436x1024 -> 448x1024 x%d" % B, timed(rz_a), B * (px + 448 * 1024 * 3), "source read once + destination written once")
    report("resize_linear_u8_kernel 436x1024 -> 568x1333 BGR x%d" % B, timed(rz_b), B * (px + 568 * 1333 * 3), "source read once + destination written once")
    report("warp_masks_kernel %d masks %dx%d (+ boxes)" % (n, H, W), timed(wm), H * W * (8 + 2 * n), "flow once, every mask read + written once; 3 launches timed")
    report("flow_postprocess_kernel %d x 112x256 -> 436x1024" % FB, timed(fp), FB * (2 * 112 * 256 * 4 + 436 * 1024 * 8), "quarter-resolution flow read once, frame-resolution flow written once")
