"""CPU restatement of MergeTrack's live mask propagation -- TEST INFRASTRUCTURE ONLY.

Follows, under /root/reference/code/MergeTrack/:
  * merge_functions.py:209-217  warp_flow:      map = grid - flow, cv2.remap(mask, map, None, INTER_LINEAR), res == 1
  * merge_functions.py:219-243  warp_proposals: warped mask -> RLE, toBbox, score = 0.5 * (final_score + 1)
  * merge.py:95-100             the per-frame loop: warp_proposals -> do_refinement on the next frame

OpenCV (cv2.remap) and pycocotools (toBbox) are third-party dependencies that are not vendored under /root/reference and not
pinned by a lock file (README.md:18-19 names no versions; this container has opencv-python 4.13.0).  Their published
algorithms are restated here:

cv2.remap(src uint8, map CV_32FC2, None, INTER_LINEAR, BORDER_CONSTANT 0)  (modules/imgproc/src/imgwarp.cpp, remap /
RemapInvoker / remapBilinear with FixedPtCast<int, uchar, 15>):
  * the float map is quantised to 1/32 pixel:  s = cvRound(coord * 32) (float32 product, round half to even, saturated to
    int32);  integer part i = saturate_cast<short>(s >> 5), fraction a = s & 31
  * weights from the 32 x 32 bilinear table in 15-bit fixed point:  w00 = (32-ay)(32-ax)*32, w01 = (32-ay)ax*32,
    w10 = ay(32-ax)*32, w11 = ay*ax*32; the (0,0) entry saturates to 32767 (short)
  * taps outside the image read the border value 0;  out = (sum(w * tap) + 16384) >> 15
For a 0/1 mask `res == 1` therefore means: the fixed-point weights on the taps that are 1 sum to >= 16384.

pycocotools toBbox (common/maskApi.c rleToBbox): the tight box [x, y, w, h] of the mask, [0, 0, 0, 0] when it is empty.

Pinned (tests/test_oracle_mergetrack.py) against cv2.remap itself run in this container and against the golden vectors in
tests/golden/mergetrack_golden.npz made by tests/golden/make_mergetrack_golden.py with cv2: bit-exact.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS           # 32
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS           # 32768


def _quantise(coord):
    """float32 map coordinate -> (integer part as int16-saturated, 5-bit fraction), as RemapInvoker does."""
    prod = (coord.astype(np.float32) * np.float32(INTER_TAB)).astype(np.float32)
    # cvRound = round half to even; out-of-range / NaN -> INT_MIN as cvtss2si does
    with np.errstate(invalid="ignore"):
        r = np.rint(prod.astype(np.float64))
    bad = ~np.isfinite(r) | (r >= 2147483648.0) | (r < -2147483648.0)
    s = np.where(bad, -2147483648.0, r).astype(np.int64)
    i = np.clip(s >> INTER_BITS, -32768, 32767)
    a = s & (INTER_TAB - 1)
    return i, a


def remap_linear_u8(img, mapxy):
    """cv2.remap(img, mapxy, None, cv2.INTER_LINEAR) for a single-channel uint8 image and a float32 [h, w, 2] map."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2
    H, W = img.shape
    ix, ax = _quantise(mapxy[:, :, 0])
    iy, ay = _quantise(mapxy[:, :, 1])
    w00 = (INTER_TAB - ay) * (INTER_TAB - ax) * 32
    w01 = (INTER_TAB - ay) * ax * 32
    w10 = ay * (INTER_TAB - ax) * 32
    w11 = ay * ax * 32
    w00 = np.minimum(w00, 32767)       # saturate_cast<short>(1.0 * 32768)
    S = img.astype(np.int64)

    def tap(y, x):
        ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
        return np.where(ok, S[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)], 0)

    acc = w00 * tap(iy, ix) + w01 * tap(iy, ix + 1) + w10 * tap(iy + 1, ix) + w11 * tap(iy + 1, ix + 1)
    return np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)


def flow_to_map(flow):
    """merge_functions.py:210-214: map = -flow + (x, y), in float32 like the numpy in-place adds of the reference."""
    flow = np.asarray(flow, dtype=np.float32)
    h, w = flow.shape[:2]
    m = -flow
    m[:, :, 0] += np.arange(w)            # float32 += int64 stays float32 (in place)
    m[:, :, 1] += np.arange(h)[:, np.newaxis]
    return m


def warp_flow(img, flow, binarize=True):
    """merge_functions.py:209-217."""
    res = remap_linear_u8(img, flow_to_map(flow))
    if binarize:
        res = np.equal(res, 1).astype(np.uint8)
    return res


def to_bbox(mask):
    """pycocotools.mask.toBbox(encode(mask)) -> float64 [x, y, w, h]; zeros for an empty mask (maskApi.c rleToBbox)."""
    m = np.asarray(mask) != 0
    if not m.any():
        return np.zeros(4, np.float64)
    ys = np.flatnonzero(m.any(axis=1))
    xs = np.flatnonzero(m.any(axis=0))
    return np.array([xs[0], ys[0], xs[-1] - xs[0] + 1, ys[-1] - ys[0] + 1], np.float64)


def warp_proposals(proposals, flow):
    """merge_functions.py:219-243 with the flow array instead of the .flo file name; 'segmentation' is left to the RLE codec
    (oracle/refnet_oracle.py rle_encode), everything else as the reference builds it."""
    out = []
    for prop in proposals:
        f_mask = warp_flow(prop["mask"], flow)
        out.append({"bbox": to_bbox(f_mask), "score": 0.5 * (prop["final_score"] + 1), "final_score": prop["final_score"],
                    "object_score": prop["object_score"], "mask": f_mask, "id": prop["id"]})
    return out
