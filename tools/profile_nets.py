"""Per-kernel device-time breakdown of the proposal and refinement networks at BASELINE sizes (GPU box only), using the
library's CUDA-event profiler (premvos_profile_begin/end)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import torch
from premvos_b200 import _lib, propnet, refnet, synth

which = sys.argv[1:] or ["propnet", "refnet"]
if "propnet" in which:
    H, W = propnet.custom_resize_shape(480, 854)
    net = propnet.ProposalNet().load_params(synth.propnet_synthetic_params(1))
    img = cv2.resize(synth.synthetic_bgr_frame(480, 854, seed=2), (W, H)).astype(np.float32)
    net(img); net(img)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): net(img)
    print("propnet wall ms/frame", (time.perf_counter() - t0) / 5 * 1e3)
    _lib.profile_begin(); net(img); prof = _lib.profile_end()
    tot = sum(v["ms"] for v in prof.values())
    print("propnet device ms (sum of kernels)", tot)
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("  %-28s n=%4d  %8.3f ms  %5.1f%%  %7.1f TF/s  %7.1f GB/s" % (k, v["launches"], v["ms"], 100 * v["ms"] / tot,
              v["flops"] / max(v["ms"], 1e-9) / 1e9, v["bytes"] / max(v["ms"], 1e-9) / 1e6))
    del net
if "refnet" in which:
    rn = refnet.RefinementNet(max_batch=20).load_params(synth.refnet_synthetic_params(2))
    frame = synth.synthetic_bgr_frame(480, 854, seed=3)
    boxes = synth.synthetic_boxes(100, 480, 854, seed=3)
    rn.refine(frame, boxes[:20])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rn.refine(frame, boxes)
    print("refnet wall ms/100 crops", (time.perf_counter() - t0) * 1e3)
    _lib.profile_begin(); rn.refine(frame, boxes[:20]); prof = _lib.profile_end()
    tot = sum(v["ms"] for v in prof.values())
    print("refnet device ms per 20 crops (sum of kernels)", tot)
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("  %-28s n=%4d  %8.3f ms  %5.1f%%  %7.1f TF/s  %7.1f GB/s" % (k, v["launches"], v["ms"], 100 * v["ms"] / tot,
              v["flops"] / max(v["ms"], 1e-9) / 1e9, v["bytes"] / max(v["ms"], 1e-9) / 1e6))
