"""Pins the cv2.resize restatement (oracle/cv_resize_oracle.py) against OpenCV itself, run here: bit-exact on the sizes the
hot path uses (script_pwc_multi.py:38-45, proposal_net common.py:49-62) and on ragged / tiny / down-scaling cases."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import cv_resize_oracle as R

CASES = [(436, 1024, 448, 1024), (436, 1024, 568, 1333), (480, 854, 512, 896), (480, 854, 749, 1333), (100, 140, 128, 192),
         (100, 140, 800, 1120), (37, 53, 64, 64), (64, 64, 37, 53), (5, 7, 64, 128), (480, 854, 480, 854), (301, 203, 300, 202)]


@pytest.mark.parametrize("sh,sw,dh,dw", CASES)
def test_resize_restatement_is_bit_exact_with_cv2(sh, sw, dh, dw):
    rng = np.random.default_rng(sh * 1000 + dw)
    for ch in (3, 1):
        img = rng.integers(0, 256, (sh, sw, ch), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR).reshape(dh, dw, ch)
        got = R.resize_linear_u8(img, dh, dw)
        np.testing.assert_array_equal(got, ref)


def test_channel_order_commutes_with_resize():
    # the proposal stage resizes the BGR frame, the flow stage the RGB frames: resize is per channel
    img = np.random.default_rng(1).integers(0, 256, (50, 70, 3), dtype=np.uint8)
    a = R.resize_linear_u8(np.ascontiguousarray(img[:, :, ::-1]), 64, 128)
    b = R.resize_linear_u8(img, 64, 128)[:, :, ::-1]
    np.testing.assert_array_equal(a, b)
