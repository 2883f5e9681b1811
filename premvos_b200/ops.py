"""Single-op entry points of the CUDA library (bring-up / parity hooks).  No CPU fallback."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def conv2d(x: torch.Tensor, weight, bias=None, stride=1, dilation=1, padding=(0, 0, 0, 0), slope=1.0,
           residual: torch.Tensor = None) -> torch.Tensor:
    """act(conv2d(x, weight) + bias + residual) on the tcgen05 tensor-core path (premvos_conv2d_forward).

    x: CUDA float32 [N,Cin,H,W]; weight: [Cout,Cin,kh,kw] (host numpy / CPU tensor); padding =
    (top, left, bottom, right) zeros; act = LeakyReLU(slope) (1 = identity, 0 = ReLU)."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise TypeError("x must be a CUDA float32 tensor (premvos_b200 has no CPU path)")
    x = x.contiguous()
    w = np.ascontiguousarray(weight.detach().cpu().numpy() if isinstance(weight, torch.Tensor) else weight, dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(
        bias.detach().cpu().numpy() if isinstance(bias, torch.Tensor) else bias, dtype=np.float32)
    N, Cin, H, W = x.shape
    Cout, Cin2, kh, kw = w.shape
    if Cin2 != Cin:
        raise ValueError("weight has %d input channels, x has %d" % (Cin2, Cin))
    pt, pl, pb, pr = padding
    Ho = (H + pt + pb - dilation * (kh - 1) - 1) // stride + 1
    Wo = (W + pl + pr - dilation * (kw - 1) - 1) // stride + 1
    out = torch.empty((N, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    res_ptr = None
    if residual is not None:
        residual = residual.contiguous()
        if tuple(residual.shape) != tuple(out.shape):
            raise ValueError("residual shape %s != output shape %s" % (tuple(residual.shape), tuple(out.shape)))
        res_ptr = residual.data_ptr()
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_conv2d_forward(
            x.data_ptr(), w.ctypes.data_as(ctypes.c_void_p), None if b is None else b.ctypes.data_as(ctypes.c_void_p),
            res_ptr, out.data_ptr(), N, Cin, H, W, Cout, kh, kw, stride, dilation, pt, pl, pb, pr, float(slope), st))
    return out


def sepconv2d(x: torch.Tensor, dw_weight, dw_bias, pw_weight, pw_bias, relu_in=False, relu_mid=False, slope=1.0) -> torch.Tensor:
    """act(pointwise(relu_mid?(depthwise3x3(relu_in?(x)) + dw_bias)) + pw_bias) on the FUSED tcgen05 path
    (premvos_sepconv2d_forward): the depthwise tile is computed into the pointwise GEMM's A operand.  x CUDA float32 [N,C,H,W];
    dw_weight [C,3,3] (or [C,1,3,3]), pw_weight [Cout,C] (or [Cout,C,1,1]), Cout <= 128; depthwise SAME padding, stride 1."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise TypeError("x must be a CUDA float32 tensor (premvos_b200 has no CPU path)")
    x = x.contiguous()
    as_np = lambda t: None if t is None else np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)
    dw, db, pw, pb = as_np(dw_weight), as_np(dw_bias), as_np(pw_weight), as_np(pw_bias)
    N, C, H, W = x.shape
    dw = dw.reshape(C, 3, 3)
    pw = pw.reshape(-1, C)
    Cout = pw.shape[0]
    out = torch.empty((N, Cout, H, W), dtype=torch.float32, device=x.device)
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_sepconv2d_forward(x.data_ptr(), vp(dw), vp(db), vp(pw), vp(pb), out.data_ptr(), N, C, H, W, Cout,
                                                        1 if relu_in else 0, 1 if relu_mid else 0, float(slope), st))
    return out


def top_k(scores: np.ndarray, k: int) -> np.ndarray:
    """tf.nn.top_k indices (proposal_net/model.py:189-190), ordered by (score desc, index asc); k <= 1024."""
    s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    out = np.zeros(1024, np.int32)
    n = ctypes.c_int()
    _lib.check(_lib.lib().premvos_topk_host(s.ctypes.data_as(ctypes.c_void_p), s.size, int(k),
                                            out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
    return out[:n.value].copy()


def non_max_suppression(boxes: np.ndarray, scores: np.ndarray, max_output_size: int, iou_threshold: float) -> np.ndarray:
    """tf.image.non_max_suppression (proposal_net/model.py:205-209, 466-467) for <= 1024 boxes -> int32 indices."""
    b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    if b.shape[0] != s.size:
        raise ValueError("boxes and scores disagree: %d vs %d" % (b.shape[0], s.size))
    out = np.zeros(1024, np.int32)
    n = ctypes.c_int()
    _lib.check(_lib.lib().premvos_nms_host(b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), s.size,
                                           float(iou_threshold), int(max_output_size), out.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.byref(n)))
    return out[:n.value].copy()


def resize_linear_u8(img: torch.Tensor, dst_h: int, dst_w: int, reverse_channels: bool = False, out: torch.Tensor = None) -> torch.Tensor:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR), bit-exact, on the device (premvos_resize_linear_u8).
    img: CUDA uint8 [H,W,C] or [B,H,W,C] with C in (1, 3).  Enqueues on the current stream."""
    if not isinstance(img, torch.Tensor) or not img.is_cuda or img.dtype != torch.uint8 or not img.is_contiguous():
        raise TypeError("img must be a contiguous CUDA uint8 tensor (premvos_b200 has no CPU path)")
    if img.dim() not in (3, 4):
        raise ValueError("expected [H,W,C] or [B,H,W,C], got %s" % (tuple(img.shape),))
    B = 1 if img.dim() == 3 else int(img.shape[0])
    H, W, C = (int(v) for v in img.shape[-3:])
    shape = (dst_h, dst_w, C) if img.dim() == 3 else (B, dst_h, dst_w, C)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=img.device)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("out must be a contiguous CUDA uint8 tensor of shape %s" % (shape,))
    with torch.cuda.device(img.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_resize_linear_u8(img.data_ptr(), B, H, W, out.data_ptr(), int(dst_h), int(dst_w), C,
                                                       1 if reverse_channels else 0, st))
    return out


def packed_mask_bytes(hw: int) -> int:
    """Bytes of one bit-packed mask of `hw` pixels (whole 64-pixel words)."""
    return 8 * ((int(hw) + 63) // 64)


def pack_mask_bits(masks: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """uint8 CUDA masks [..., H, W] (non-zero = 1) -> uint8 [..., packed_mask_bytes(H*W)], 8 pixels per byte, pixel i of a mask
    at bit i % 8 of byte i / 8 (premvos_pack_mask_bits).  Enqueues on the current stream."""
    if not isinstance(masks, torch.Tensor) or not masks.is_cuda or masks.dtype != torch.uint8 or not masks.is_contiguous() or masks.dim() < 2:
        raise TypeError("masks must be a contiguous CUDA uint8 tensor [..., H, W] (premvos_b200 has no CPU path)")
    hw = int(masks.shape[-2]) * int(masks.shape[-1])
    n = masks.numel() // hw if hw else 0
    shape = tuple(masks.shape[:-2]) + (packed_mask_bytes(hw),)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=masks.device)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("out must be a contiguous CUDA uint8 tensor of shape %s" % (shape,))
    with torch.cuda.device(masks.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_pack_mask_bits(masks.data_ptr(), n, hw, out.data_ptr(), st))
    return out


def unpack_mask_bits(packed, height: int, width: int) -> np.ndarray:
    """Host inverse of pack_mask_bits: uint8 [..., packed_mask_bytes(H*W)] (numpy or CPU tensor) -> uint8 0/1 [..., H, W]."""
    a = packed.numpy() if isinstance(packed, torch.Tensor) else np.asarray(packed)
    bits = np.unpackbits(np.ascontiguousarray(a, dtype=np.uint8), axis=-1, bitorder="little")
    return bits[..., :height * width].reshape(a.shape[:-1] + (height, width))
