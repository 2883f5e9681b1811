"""Target of the ncu captures (GPU box only): builds the resident pipeline (or one network), warms up, then runs ONE unit of
work between cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly that work.
    python tools/ncu_step.py pipeline      one bench step (4 frame pairs: flow + 2 proposal passes + 40 refine boxes each)
    python tools/ncu_step.py refnet        one refinement launch group (40 crops, the benchmarked group)
    python tools/ncu_step.py flow          one PWC forward (4 pairs)
    python tools/ncu_step.py propnet       one batched proposal forward (4 frames; ncu profiles the kernel nodes of the CUDA graph one by one)
    python tools/ncu_step.py reid          one ReID forward (40 crops)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from premvos_b200 import pipeline, refnet, synth
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "pipeline"
H, W = bench.H_IN, bench.W_IN
if what == "reid":
    from premvos_b200 import reid
    net = reid.ReIDNet(max_batch=40).load_params(synth.reid_synthetic_params(0))
    frame = torch.from_numpy(synth.synthetic_bgr_frame(H, W, seed=3)).cuda()
    boxes = torch.from_numpy(synth.synthetic_boxes(40, H, W, seed=3)).cuda()
    run = lambda: net.embed_device(frame, boxes)
elif what == "propnet":
    from premvos_b200 import ops, propnet
    Hp, Wp = propnet.custom_resize_shape(H, W)
    net = propnet.ProposalNet().load_params(synth.propnet_synthetic_params(8))
    frames = np.stack([synth.synthetic_bgr_frame(H, W, seed=2 + i) for i in range(4)])
    imgs = ops.resize_linear_u8(torch.from_numpy(frames).cuda(), Hp, Wp)
    run = lambda: net.forward_device(imgs)
elif what == "refnet":
    rn = refnet.RefinementNet(max_batch=40).load_params(synth.refnet_synthetic_params(2))
    frame = torch.from_numpy(synth.synthetic_bgr_frame(H, W, seed=3)).cuda()
    boxes = torch.from_numpy(synth.synthetic_boxes(40, H, W, seed=3)).cuda()
    run = lambda: rn.refine_device(frame, boxes)
else:
    sd = {k: torch.from_numpy(v) for k, v in synth.pwc_synthetic_state_dict(0).items()}
    pipe = pipeline.FramePipeline(sd, synth.propnet_synthetic_params(8), synth.propnet_synthetic_params(3),
                                  synth.refnet_synthetic_params(2), (H, W), pairs_per_step=4, boxes_per_frame=40)
    units = bench.make_units(4, 40)
    dev = [torch.from_numpy(np.stack([u[i] for u in units])).cuda() for i in range(3)]
    if what == "flow":
        ff, _ = pipe.prepare_device(dev[0], dev[1])
        run = lambda: pipe.flow_net.forward_u8(ff, out=pipe.out["flow"])
    else:
        run = lambda: pipe.run_frames_device(*dev, concurrent=False)
for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one", what)
