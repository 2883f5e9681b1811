"""Per-layer device-time breakdown of the proposal and refinement networks at the bench sizes (GPU box only): the library's
CUDA-event profiler with PREMVOS_PROFILE_LAYERS=1 labels every conv_umma launch with its geometry and plan."""
import os, sys
os.environ["PREMVOS_PROFILE_LAYERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
import torch
from premvos_b200 import _lib, propnet, refnet, synth

H0, W0 = 436, 1024


def report(title, prof, top=40):
    tot = sum(v["ms"] for v in prof.values())
    print("== %s: %.3f ms (sum of kernels)" % (title, tot))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:top]:
        print("  %7.3f ms %5.1f%% n=%3d %7.1f us/launch %6.1f TF/s %7.1f GB/s  %s" % (
            v["ms"], 100 * v["ms"] / tot, v["launches"], 1e3 * v["ms"] / v["launches"], v["flops"] / max(v["ms"], 1e-9) / 1e9,
            v["bytes"] / max(v["ms"], 1e-9) / 1e6, k))


which = sys.argv[1:] or ["propnet", "refnet"]
if "propnet" in which:
    H, W = propnet.custom_resize_shape(H0, W0)
    net = propnet.ProposalNet().load_params(synth.propnet_synthetic_params(1))
    img = cv2.resize(synth.synthetic_bgr_frame(H0, W0, seed=2), (W, H)).astype(np.float32)
    net(img); net(img)
    _lib.profile_begin(); net(img); prof = _lib.profile_end()
    report("proposal net %dx%d" % (H, W), prof)
    del net
if "refnet" in which:
    NB = int(os.environ.get("REFNET_BOXES", "40"))
    rn = refnet.RefinementNet(max_batch=NB).load_params(synth.refnet_synthetic_params(2))
    frame = synth.synthetic_bgr_frame(H0, W0, seed=3)
    boxes = synth.synthetic_boxes(NB, H0, W0, seed=3)
    rn.refine(frame, boxes)
    _lib.profile_begin(); rn.refine(frame, boxes); prof = _lib.profile_end()
    report("refinement net, %d crops" % NB, prof)
if "reid" in which:
    from premvos_b200 import reid
    NB = int(os.environ.get("REID_BOXES", "40"))
    net = reid.ReIDNet(max_batch=NB).load_params(synth.reid_synthetic_params(0))
    frame = torch.from_numpy(synth.synthetic_bgr_frame(H0, W0, seed=3)).cuda()
    boxes = torch.from_numpy(synth.synthetic_boxes(NB, H0, W0, seed=3)).cuda()
    net.embed_device(frame, boxes); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net.embed_device(frame, boxes)
    e1.record(); torch.cuda.synchronize()
    print("ReID net, %d crops: %.3f ms per frame (device, 5 runs back to back), %d launches" % (NB, e0.elapsed_time(e1) / 5, net.launches_per_forward()))
    _lib.profile_begin(); net.embed_device(frame, boxes); torch.cuda.synchronize(); prof = _lib.profile_end()
    report("ReID net, %d crops" % NB, prof, top=60)
