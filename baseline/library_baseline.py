"""GPU *library* baseline of the same unit of work (context for bench.py, never the product path).

What the reference really executes in production is cuDNN: torch 0.2 modules for PWC-Net
(/root/reference/code/optical_flow_net-PWC-Net/models/PWCNet.py:24-34) and TensorFlow 1.8 convolutions for the proposal and
refinement networks (proposal_net/basemodel.py:74-99, refinement_net/network/deeplab/core/xception.py:152-190).  Neither stack is
installable here, so this module runs the ORACLE restatements of the three forwards (oracle/*.py, plain torch.nn.functional
graphs) on the GPU through torch 2.11 / cuDNN 9, in fp32 with TF32 disabled ("fp32") and allowed ("tf32"):

  flow      oracle.pwc_oracle.pwc_forward on a [B,6,448,1024] batch, correlation as 81 shifted products in torch
  proposals ResNet-101 C4 backbone + RPN head + conv5 head on 100 RoIs (oracle.propnet_oracle layer functions; top-k / NMS / RoIAlign
            replaced by random RoI features: they are not convolution time), once per weight set
  refine    Xception-65 + ASPP + decoder + logits on [K,4,385,385] crops (oracle.refnet_oracle layer functions)

BatchNorm is executed unfolded, as the reference graphs do.  The numbers say how fast a plain library implementation of the
same arithmetic runs on this GPU; they are reported beside the product's, not compared as a ratio by the driver.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F


def _corr81(a, b):
    """cost volume of PWCNet.py:69 (md = 4) as 81 shifted channel means: [B,C,H,W] x2 -> [B,81,H,W]"""
    B, C, H, W = a.shape
    bp = F.pad(b, (4, 4, 4, 4))
    out = [(a * bp[:, :, dy:dy + H, dx:dx + W]).mean(1, keepdim=True) for dy in range(9) for dx in range(9)]
    return torch.cat(out, 1)


def _time(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def run(pairs=4, boxes=40, iters=3, hn=448, wn=1024, hp=568, wp=1333):
    """-> {"fp32": {...}, "tf32": {...}}: seconds per unit (flow + 2 proposal passes + `boxes` crops) and pairs/s, batched as the
    product runs them (`pairs` frames per forward, `boxes` crops per forward) and one at a time as the reference iterates."""
    from oracle import propnet_oracle as PO, pwc_oracle as O, refnet_oracle as RO
    from premvos_b200 import synth
    dev = torch.device("cuda")
    sd = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in synth.pwc_synthetic_state_dict(0).items()}
    PP = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in synth.propnet_synthetic_params(1).items()}
    RP = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in synth.refnet_synthetic_params(2).items()}
    g = torch.Generator(device="cuda").manual_seed(0)

    def flow(b):
        x = torch.rand(b, 6, hn, wn, device=dev, generator=g)
        return lambda: O.pwc_forward(sd, x, corr_fn=_corr81)

    def proposals(b):
        img = torch.randn(b, 3, hp, wp, device=dev, generator=g)
        rois = torch.randn(100 * b, 1024, 14, 14, device=dev, generator=g)

        def f():
            fm = PO.pretrained_resnet_conv4(PP, img, PO.RESNET_NUM_BLOCK)
            hidden = F.relu(F.conv2d(fm, PO._w(PP, "rpn/conv0/W"), PP["rpn/conv0/b"], padding=1))     # model.py:31-51
            F.conv2d(hidden, PO._w(PP, "rpn/class/W"), PP["rpn/class/b"])
            F.conv2d(hidden, PO._w(PP, "rpn/box/W"), PP["rpn/box/b"])
            return PO.resnet_conv5(PP, rois, PO.RESNET_NUM_BLOCK[-1]).mean(dim=(2, 3))
        return f

    def refine(k):
        x = torch.randn(k, 4, 385, 385, device=dev, generator=g)

        def f():
            ep = {}
            feat = RO.xception_65(RP, x, ep)
            low = ep["xception_65/" + RO.LOW_LEVEL_FEATURE]
            h, w = feat.shape[2:]
            br = [RO._conv_bn(RP, "image_pooling", feat.mean(dim=(2, 3), keepdim=True), RO.ASPP_BN_EPS).expand(-1, -1, h, w),
                  RO._conv_bn(RP, "aspp0", feat, RO.ASPP_BN_EPS)]
            for i, r in enumerate((6, 12, 18), 1):
                br.append(RO._separable(RP, "aspp%d" % i, feat, 256, RO.ASPP_BN_EPS, 1, r, True))
            aspp = RO._conv_bn(RP, "concat_projection", torch.cat(br, 1), RO.ASPP_BN_EPS)
            dh = int((385 - 1.0) * 0.25 + 1.0)
            up = F.interpolate(aspp, (dh, dh), mode="bilinear", align_corners=True)
            lo = F.interpolate(RO._conv_bn(RP, "decoder/feature_projection0", low, RO.ASPP_BN_EPS), (dh, dh), mode="bilinear",
                               align_corners=True)
            dec = RO._separable(RP, "decoder/decoder_conv0", torch.cat([up, lo], 1), 256, RO.ASPP_BN_EPS, 1, 1, True)
            dec = RO._separable(RP, "decoder/decoder_conv1", dec, 256, RO.ASPP_BN_EPS, 1, 1, True)
            wl = RP["logits/features/weights"].permute(3, 2, 0, 1).contiguous()
            return F.conv2d(dec, wl, RP["logits/features/biases"])
        return f

    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        with torch.no_grad(), torch.device("cuda"):
            torch.backends.cudnn.benchmark = True
            for name, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                t_flow_b, t_prop_b = _time(flow(pairs), iters) / pairs, _time(proposals(pairs), iters) / pairs
                chunk = min(boxes, 20)   # 40 crops of Xception-65 activations in one torch graph do not pay; 20 per forward
                t_ref_b = _time(refine(chunk), iters) / chunk
                t_flow_1, t_prop_1, t_ref_1 = _time(flow(1), iters), _time(proposals(1), iters), _time(refine(1), iters)
                unit_b = t_flow_b + 2 * t_prop_b + boxes * t_ref_b
                unit_1 = t_flow_1 + 2 * t_prop_1 + boxes * t_ref_1
                out[name] = {"batched_pairs_per_s": 1.0 / unit_b, "batch1_pairs_per_s": 1.0 / unit_1,
                             "batched_ms": {"flow_per_pair": t_flow_b * 1e3, "proposal_pass_per_frame": t_prop_b * 1e3,
                                            "refine_per_crop": t_ref_b * 1e3},
                             "batch1_ms": {"flow_per_pair": t_flow_1 * 1e3, "proposal_pass_per_frame": t_prop_1 * 1e3,
                                           "refine_per_crop": t_ref_1 * 1e3}}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["note"] = ("oracle restatements of the three forwards on the GPU through torch %s / cuDNN (the reference's production path is "
                   "cuDNN under torch 0.2 / TF 1.8, not installable here); unit = flow + 2 proposal passes + %d crops; batched = %d "
                   "frames / up to 20 crops per forward, batch1 = one at a time as the reference iterates; unfolded BatchNorm; context "
                   "only" % (torch.__version__, boxes, pairs))
    return out


if __name__ == "__main__":
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    print(json.dumps(run(), indent=1))
