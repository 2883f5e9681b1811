"""Per-layer device-time breakdown of one BATCHED proposal-network forward at the bench size (GPU box only):
    PREMVOS_PROFILE_LAYERS=1 python tools/profile_propnet_batch.py [batch]"""
import os, sys
os.environ.setdefault("PREMVOS_PROFILE_LAYERS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from premvos_b200 import _lib, ops, propnet, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = propnet.custom_resize_shape(436, 1024)
net = propnet.ProposalNet().load_params(synth.propnet_synthetic_params(1))
frames = np.stack([synth.synthetic_bgr_frame(436, 1024, seed=2 + i) for i in range(B)])
imgs = ops.resize_linear_u8(torch.from_numpy(frames).cuda(), H, W)
x = imgs if B > 1 else imgs[0]
for _ in range(3):
    net.forward_device(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net.forward_device(x)
e1.record()
torch.cuda.synchronize()
print("batch %d: %.3f ms per forward, %.3f ms per frame (CUDA graph)" % (B, e0.elapsed_time(e1) / 5, e0.elapsed_time(e1) / 5 / B))
_lib.profile_begin()
net.forward_device(x)
prof = _lib.profile_end()
tot = sum(v["ms"] for v in prof.values())
print("sum of kernels %.3f ms" % tot)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:45]:
    print("  %7.3f ms %5.1f%% n=%3d %7.1f us/launch %6.1f TF/s %7.1f GB/s  %s" % (
        v["ms"], 100 * v["ms"] / tot, v["launches"], 1e3 * v["ms"] / v["launches"], v["flops"] / max(v["ms"], 1e-9) / 1e9,
        v["bytes"] / max(v["ms"], 1e-9) / 1e6, k))
