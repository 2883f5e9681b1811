"""CPU restatement of `cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)` for uint8 images -- TEST INFRASTRUCTURE ONLY.

The reference's stage drivers resize every frame on the host with OpenCV before a network sees it:
  * optical_flow_net-PWC-Net/script_pwc_multi.py:38-45   both frames -> multiples of 64 (cv2.resize, default INTER_LINEAR)
  * proposal_net/eval.py:75-78 + common.py:35-62          CustomResize -> tensorpack ResizeTransform -> cv2.resize INTER_LINEAR
OpenCV is a third-party dependency that is not vendored under /root/reference (README.md names no version; this container has
opencv-python 4.13.0).  Its published algorithm for 8-bit linear resize (modules/imgproc/src/resize.cpp: resizeGeneric_ with
HResizeLinear / VResizeLinear, INTER_RESIZE_COEF_BITS = 11) is restated here:
  * source coordinate of destination index d:  f = (float)((d + 0.5) * (1 / (dst / src)) - 0.5),  s = floor(f),  f -= s
  * horizontal: s < 0 -> (s, f) = (0, 0);  s >= src - 1 -> (s, f) = (src - 1, 0);  taps s and min(s + 1, src - 1)
  * vertical:   f is NOT clamped, the two row indices are clipped to [0, src - 1] instead
  * coefficients  c0 = cvRound((1 - f) * 2048), c1 = cvRound(f * 2048)   (float products, round half to even)
  * horizontal pass in int32:  r = S[s] * a0 + S[s + 1] * a1
  * vertical pass:             out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
Pinned (tests/test_oracle_resize.py) against outputs of cv2 itself run in this container: bit-exact on every size the path uses.
"""
import numpy as np

COEF_SCALE = 2048


def linear_coefficients(dst_n, src_n, clamp_fraction):
    """-> (index int32[dst_n], c0 int32[dst_n], c1 int32[dst_n]) as OpenCV's resize() builds xofs/ialpha (clamp_fraction=True)
    and yofs/ibeta (clamp_fraction=False)."""
    scale = 1.0 / (float(dst_n) / src_n)
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src_n - 1
        f[hi] = 0
        s[hi] = src_n - 1
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    c1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    return s.astype(np.int32), c0, c1


def resize_linear_u8(src, dst_h, dst_w):
    """src uint8 [H, W] or [H, W, C] -> uint8 [dst_h, dst_w(, C)], bit-exact with cv2.resize(..., INTER_LINEAR)."""
    src = np.asarray(src)
    assert src.dtype == np.uint8
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    sh, sw = src.shape[:2]
    sx, a0, a1 = linear_coefficients(dst_w, sw, True)
    sy, b0, b1 = linear_coefficients(dst_h, sh, False)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy0, sy1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    S = src.astype(np.int64)
    rows = S[:, sx] * a0[None, :, None].astype(np.int64) + S[:, sx1] * a1[None, :, None].astype(np.int64)
    r0, r1 = rows[sy0], rows[sy1]
    out = (((b0[:, None, None].astype(np.int64) * (r0 >> 4)) >> 16) + ((b1[:, None, None].astype(np.int64) * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_linear_f32(src, dst_h, dst_w):
    """cv2.resize(src float32 [H, W], (dst_w, dst_h), INTER_LINEAR): same coordinates as the 8-bit path, float coefficients
    (1 - f, f), horizontal pass then vertical pass, every product / sum rounded to float32 (OpenCV's SIMD rows may fuse a
    multiply-add: pinned against cv2 to 1e-6 relative, not bit-exact).  Used by script_pwc_multi.py:64-65 on the flow."""
    src = np.asarray(src, dtype=np.float32)
    sh, sw = src.shape
    sx, _, _ = linear_coefficients(dst_w, sw, True)
    sy, _, _ = linear_coefficients(dst_h, sh, False)

    def frac(dst_n, src_n, clamp):
        scale = 1.0 / (float(dst_n) / src_n)
        f = ((np.arange(dst_n, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if clamp:
            f[(s < 0) | (s >= src_n - 1)] = 0
        return f

    fx, fy = frac(dst_w, sw, True), frac(dst_h, sh, False)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy0, sy1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    one = np.float32(1.0)
    rows = (src[:, sx] * (one - fx)[None, :]).astype(np.float32) + (src[:, sx1] * fx[None, :]).astype(np.float32)
    rows = rows.astype(np.float32)
    out = (rows[sy0] * (one - fy)[:, None]).astype(np.float32) + (rows[sy1] * fy[:, None]).astype(np.float32)
    return out.astype(np.float32)
