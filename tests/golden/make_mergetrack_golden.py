"""Generates tests/golden/mergetrack_golden.npz with the reference's own warp_flow (MergeTrack/merge_functions.py:209-217,
source exec'd unmodified) and with OpenCV itself for the flow post-processing (cv2.resize as
optical_flow_net-PWC-Net/script_pwc_multi.py:59-68 calls it), run in the build container.
IPP is switched off for the float resize so that the vectors are those of OpenCV's own published code path.

    python tests/golden/make_mergetrack_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from premvos_b200 import synth  # noqa: E402


def _reference_warp_flow():
    """MergeTrack/merge_functions.py:209-217 itself: the function's unmodified source, exec'd from where it lies."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from make_reference_function_goldens import extract
    return extract("/root/reference/code/MergeTrack/merge_functions.py", ["warp_flow"])["warp_flow"]


def main():
    reference_warp_flow = _reference_warp_flow()
    H, W = 60, 84
    masks = synth.synthetic_masks(3, H, W, seed=11)
    rng = np.random.default_rng(12)
    flow = (rng.standard_normal((H, W, 2)) * 6).astype(np.float32)
    flow[:8] = np.round(flow[:8] * 64) / 64          # ties of the 1/32-pixel quantisation
    flow[8, :4] = [[1e6, -1e6], [3e9, 0], [np.nan, 0], [np.inf, -np.inf]]
    warped = np.stack([reference_warp_flow(m, flow.copy()) for m in masks])
    gray = rng.integers(0, 256, (H, W), dtype=np.uint8)
    remapped = reference_warp_flow(gray, flow.copy(), binarize=False)
    # script_pwc_multi.py:59-68 for a 64 x 128 network input and a 60 x 84 frame
    flow2 = (rng.standard_normal((2, 16, 32)) * 0.3).astype(np.float32)
    cv2.ipp.setUseIPP(False)
    flo = np.swapaxes(np.swapaxes(flow2 * 20.0, 0, 1), 1, 2)
    u_ = cv2.resize(flo[:, :, 0], (W, H))
    v_ = cv2.resize(flo[:, :, 1], (W, H))
    u_ *= W / float(128)
    v_ *= H / float(64)
    post = np.dstack((u_, v_)).astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mergetrack_golden.npz"), masks=masks, flow=flow, warped=warped,
                        gray=gray, remapped=remapped, flow2=flow2, post=post, cv2_version=cv2.__version__)
    print("wrote mergetrack_golden.npz", warped.sum(axis=(1, 2)), cv2.__version__)


if __name__ == "__main__":
    main()
