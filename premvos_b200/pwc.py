"""Host-side mirror of the reference's optical-flow call surface, backed by the CUDA library.

Reference surface reproduced here (paths relative to code/optical_flow_net-PWC-Net):
  models.pwc_dc_net(path=None) -> module            models/PWCNet.py:496-505
  net.cuda(); net.eval(); net(x) -> flow2           script_pwc_multi.py:88-90, PWCNet.py:179-272
  Correlation(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)(a, b)
                                                    correlation_package/modules/corr.py:4-20
  calculate_flow(net, im1_fn, im2_fn) -> [H,W,2]    script_pwc_multi.py:33-70
  writeFlowFile(filename, uv)                       script_pwc_multi.py:16-31

Tensors stay torch CUDA tensors at the boundary (plumbing only); every FLOP runs in
libpremvos_b200.so.  Nothing here falls back to PyTorch ops or to the CPU oracle.
"""
from __future__ import annotations

import ctypes
import math
import sys
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .synth import pwc_param_shapes

__all__ = ["pwc_dc_net", "PWCDCNet", "Correlation", "correlation_forward", "calculate_flow", "writeFlowFile",
           "readFlowFile"]


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (premvos_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def correlation_forward(input1: torch.Tensor, input2: torch.Tensor, pad_size=4, kernel_size=1,
                        max_displacement=4, stride1=1, stride2=1, corr_multiply=1) -> torch.Tensor:
    """Drop-in for `corr.corr_cuda_forward(input1, input2, rbot1, rbot2, output, ...)`
    (functions/corr.py:24-32): returns the freshly allocated output instead of resizing a passed one."""
    a = _require_cuda_f32(input1, "input1")
    b = _require_cuda_f32(input2, "input2")
    if a.dim() != 4 or a.shape != b.shape:
        raise ValueError("inputs must both be [B,C,H,W] with equal shapes, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    L = _lib.lib()
    B, C, H, W = a.shape
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.premvos_corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                           ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)))
    out = torch.empty((B, oc.value, oh.value, ow.value), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.premvos_corr_forward(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, C, H, W, pad_size,
                                          kernel_size, max_displacement, stride1, stride2, corr_multiply, st))
    return out


class Correlation(torch.nn.Module):
    """Same constructor and call as correlation_package.modules.corr.Correlation (forward only)."""

    def __init__(self, pad_size=None, kernel_size=None, max_displacement=None, stride1=None, stride2=None,
                 corr_multiply=None):
        super().__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def reset_params(self):
        return

    def forward(self, input1, input2):
        return correlation_forward(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                   self.stride1, self.stride2, self.corr_multiply)

    def __repr__(self):
        return self.__class__.__name__


class PWCDCNet:
    """PWC-DC-Net with the reference module's call surface: state_dict()/load_state_dict(), cuda(),
    eval(), __call__(x[B,6,H,W]) -> flow2[B,2,H/4,W/4].  Device handles are created lazily per
    (batch, H, W) because the library pre-allocates every buffer and captures a CUDA graph."""

    def __init__(self, md=4, tensor_cores=None, cuda_graph=True):
        if md != 4:
            raise NotImplementedError("the reference pipeline only instantiates md=4 (PWCNet.py:496-498)")
        self._shapes = pwc_param_shapes()
        self._state = OrderedDict()
        self._handles = {}
        self._tensor_cores = tensor_cores
        self._cuda_graph = cuda_graph
        self.training = False
        self._device = None
        self._init_weights()

    # -- nn.Module-like surface -------------------------------------------------------------------
    def _init_weights(self):
        # kaiming_normal(fan_in) weights, zero bias like PWCNet.py:133-137 (overwritten by load_state_dict)
        g = torch.Generator().manual_seed(0)
        for k, shp in self._shapes.items():
            if k.endswith(".weight"):
                fan_in = shp[1] * shp[2] * shp[3]
                self._state[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in)
            else:
                self._state[k] = torch.zeros(shp)

    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._state.items())

    def load_state_dict(self, sd, strict=True):
        keys = set(sd.keys())
        missing = [k for k in self._shapes if k not in keys]
        unexpected = [k for k in keys if k not in self._shapes]
        if strict and (missing or unexpected):
            raise RuntimeError("Error(s) in loading state_dict for PWCDCNet: missing keys %s, unexpected keys %s"
                               % (missing, unexpected))
        for k, shp in self._shapes.items():
            if k not in sd:
                continue
            v = sd[k]
            v = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v.detach().cpu()
            if tuple(v.shape) != tuple(shp):
                raise RuntimeError("size mismatch for %s: got %s, expected %s" % (k, tuple(v.shape), tuple(shp)))
            self._state[k] = v.to(torch.float32).contiguous().clone()
        self._drop_handles()
        return self

    def cuda(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("premvos_b200 needs a CUDA device (sm_100a); none is visible")
        self._device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                    (device if isinstance(device, int) else torch.device(device).index or 0))
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("premvos_b200 implements the inference path only")
        return self

    def parameters(self):
        return iter(self._state.values())

    # -- device handles -----------------------------------------------------------------------------
    def _drop_handles(self):
        for h in self._handles.values():
            _lib.lib().premvos_pwc_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self._drop_handles()
        except Exception:
            pass

    def _handle(self, B, H, W):
        # one handle (buffers, captured graph) per (device, batch, size): a handle built on one GPU must never serve tensors of
        # another; callers enter torch.cuda.device(tensor.device) first.  A handle supports ONE stream at a time.
        key = (torch.cuda.current_device(), B, H, W)
        if key in self._handles:
            return self._handles[key]
        L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(L.premvos_pwc_create(ctypes.byref(h), B, H, W))
        try:
            if self._tensor_cores is not None:
                _lib.check(L.premvos_pwc_set_option(h, b"tensor_cores", int(bool(self._tensor_cores))))
            _lib.check(L.premvos_pwc_set_option(h, b"cuda_graph", int(bool(self._cuda_graph))))
            for k, v in self._state.items():
                a = np.ascontiguousarray(v.numpy(), dtype=np.float32)
                _lib.check(L.premvos_pwc_set_param(h, k.encode(), a.ctypes.data_as(ctypes.c_void_p), a.size))
            _lib.check(L.premvos_pwc_finalize(h))
        except Exception:
            L.premvos_pwc_destroy(h)
            raise
        self._handles[key] = h
        return h

    def launches_per_forward(self, B, H, W) -> int:
        return int(_lib.lib().premvos_pwc_launches_per_forward(self._handle(B, H, W)))

    def tensor_core_layers(self, B, H, W) -> int:
        return int(_lib.lib().premvos_pwc_tensor_core_layers(self._handle(B, H, W)))

    # -- forward ------------------------------------------------------------------------------------
    def __call__(self, x):
        return self.forward(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _require_cuda_f32(x, "x")
        if x.dim() != 4 or x.shape[1] != 6:
            raise ValueError("expected x of shape [B,6,H,W], got %s" % (tuple(x.shape),))
        B, _, H, W = x.shape
        if H % 64 or W % 64:
            raise ValueError("H and W must be multiples of 64 (script_pwc_multi.py:38-45), got %dx%d" % (H, W))
        with torch.cuda.device(x.device):
            h = self._handle(B, H, W)
            out = torch.empty((B, 2, H // 4, W // 4), dtype=torch.float32, device=x.device)
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().premvos_pwc_forward(h, x.data_ptr(), out.data_ptr(), st))
        return out

    def forward_host(self, x: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        """End-to-end call on HOST buffers: H2D copy, forward, D2H copy, synchronise (all inside the
        library).  `x` float32 [B,6,H,W] C-contiguous (pinned memory makes the copies asynchronous)."""
        if isinstance(x, torch.Tensor):
            xa = x.numpy()
        else:
            xa = x
        if xa.dtype != np.float32 or not xa.flags["C_CONTIGUOUS"] or xa.ndim != 4 or xa.shape[1] != 6:
            raise ValueError("x must be C-contiguous float32 [B,6,H,W]")
        B, _, H, W = xa.shape
        if out is None:
            out = np.empty((B, 2, H // 4, W // 4), dtype=np.float32)
        oa = out.numpy() if isinstance(out, torch.Tensor) else out
        h = self._handle(B, H, W)
        _lib.check(_lib.lib().premvos_pwc_forward_host(h, xa.ctypes.data_as(ctypes.c_void_p),
                                                       oa.ctypes.data_as(ctypes.c_void_p)))
        return out

    def forward_host_u8(self, frames: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        """End-to-end call on decoded frames: uint8 RGB [B,2,H,W,3] (H, W multiples of 64) -> flow2 [B,2,H/4,W/4].
        BGR / 255 / planar layout (script_pwc_multi.py:47-56) run on the device; identical bits to forward_host."""
        fa = frames.numpy() if isinstance(frames, torch.Tensor) else frames
        if fa.dtype != np.uint8 or not fa.flags["C_CONTIGUOUS"] or fa.ndim != 5 or fa.shape[1] != 2 or fa.shape[4] != 3:
            raise ValueError("frames must be C-contiguous uint8 [B,2,H,W,3]")
        B, _, H, W, _ = fa.shape
        if H % 64 or W % 64:
            raise ValueError("H and W must be multiples of 64 (script_pwc_multi.py:38-45), got %dx%d" % (H, W))
        if out is None:
            out = np.empty((B, 2, H // 4, W // 4), dtype=np.float32)
        oa = out.numpy() if isinstance(out, torch.Tensor) else out
        h = self._handle(B, H, W)
        _lib.check(_lib.lib().premvos_pwc_forward_host_u8(h, fa.ctypes.data_as(ctypes.c_void_p), oa.ctypes.data_as(ctypes.c_void_p)))
        return out

    def forward_u8(self, frames: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """Device-resident variant of forward_host_u8: `frames` is a CUDA uint8 tensor [B,2,H,W,3] (RGB); enqueues on the
        current stream and returns the CUDA flow tensor without synchronising."""
        if not isinstance(frames, torch.Tensor) or not frames.is_cuda or frames.dtype != torch.uint8:
            raise TypeError("frames must be a CUDA uint8 tensor (this build has no CPU path)")
        if frames.dim() != 5 or frames.shape[1] != 2 or frames.shape[4] != 3 or not frames.is_contiguous():
            raise ValueError("frames must be contiguous [B,2,H,W,3], got %s" % (tuple(frames.shape),))
        B, _, H, W, _ = frames.shape
        if H % 64 or W % 64:
            raise ValueError("H and W must be multiples of 64 (script_pwc_multi.py:38-45), got %dx%d" % (H, W))
        with torch.cuda.device(frames.device):
            h = self._handle(B, H, W)
            if out is None:
                out = torch.empty((B, 2, H // 4, W // 4), dtype=torch.float32, device=frames.device)
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().premvos_pwc_forward_u8(h, frames.data_ptr(), out.data_ptr(), st))
        return out

    def get_tensor(self, name: str, B, H, W) -> np.ndarray:
        """Test hook: intermediate of the last forward, NCHW (see include/premvos_b200.h)."""
        L = _lib.lib()
        h = self._handle(B, H, W)
        n = ctypes.c_int64()
        _lib.check(L.premvos_pwc_get_tensor(h, name.encode(), None, ctypes.byref(n)))
        buf = np.empty(n.value, dtype=np.float32)
        _lib.check(L.premvos_pwc_get_tensor(h, name.encode(), buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
        return buf


def pwc_dc_net(path=None, **kw) -> PWCDCNet:
    """models.pwc_dc_net (PWCNet.py:496-505): accepts a raw state_dict file or {'state_dict': ...}."""
    model = PWCDCNet(**kw)
    if path is not None:
        data = torch.load(path, map_location="cpu")
        if "state_dict" in data.keys():
            model.load_state_dict(data["state_dict"])
        else:
            model.load_state_dict(data)
    return model


# ---- stage-1 driver pieces (script_pwc_multi.py) ----------------------------------------------------
def writeFlowFile(filename, uv):
    """Middlebury .flo: magic 202021.25 (f32), W (i32), H (i32), H*W*2 f32 row-major."""
    TAG_STRING = np.array(202021.25, dtype=np.float32)
    if uv.shape[2] != 2:
        sys.exit("writeFlowFile: flow must have two bands!")
    H = np.array(uv.shape[0], dtype=np.int32)
    W = np.array(uv.shape[1], dtype=np.int32)
    with open(filename, "wb") as f:
        f.write(TAG_STRING.tobytes())
        f.write(W.tobytes())
        f.write(H.tobytes())
        f.write(np.ascontiguousarray(uv, dtype=np.float32).tobytes())


def readFlowFile(filename):
    """Inverse of writeFlowFile; what MergeTrack's get_flow does (MergeTrack/merge_functions.py:197-207)."""
    with open(filename, "rb") as f:
        magic = np.frombuffer(f.read(4), dtype=np.float32)[0]
        if magic != np.float32(202021.25):
            raise ValueError("%s: bad .flo magic %r" % (filename, magic))
        w = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
        h = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
        data = np.frombuffer(f.read(4 * 2 * w * h), dtype=np.float32)
    return data.reshape(h, w, 2).copy()


def preprocess_frames(im1: np.ndarray, im2: np.ndarray):
    """script_pwc_multi.py:34-56 on already decoded RGB uint8 frames -> float32 [1,6,H_,W_]."""
    import cv2
    H, W = im1.shape[:2]
    H_ = int(math.ceil(H / 64.0) * 64)
    W_ = int(math.ceil(W / 64.0) * 64)
    chans = []
    for im in (im1, im2):
        im = cv2.resize(im[:, :, :3], (W_, H_))
        im = im[:, :, ::-1]
        im = 1.0 * im / 255.0
        chans.append(np.transpose(im, (2, 0, 1)))
    x = np.ascontiguousarray(np.concatenate(chans, 0)[None], dtype=np.float32)
    return x, (H, W, H_, W_)


def postprocess_flow(flow2: np.ndarray, H, W, H_, W_):
    """script_pwc_multi.py:59-68: x20, HWC, resize u/v back to (W,H), rescale by W/W_, H/H_."""
    import cv2
    flo = flow2 * 20.0
    flo = np.swapaxes(np.swapaxes(flo, 0, 1), 1, 2)
    u_ = cv2.resize(flo[:, :, 0], (W, H))
    v_ = cv2.resize(flo[:, :, 1], (W, H))
    u_ = u_ * (W / float(W_))
    v_ = v_ * (H / float(H_))
    return np.dstack((u_, v_)).astype(np.float32)


def calculate_flow(net: PWCDCNet, im1_fn, im2_fn):
    """script_pwc_multi.py:33-70.  `im*_fn` may be file names or already decoded RGB arrays."""
    import cv2

    def _read(fn):
        if isinstance(fn, np.ndarray):
            return fn
        im = cv2.imread(fn, cv2.IMREAD_COLOR)
        if im is None:
            raise FileNotFoundError(fn)
        return im[:, :, ::-1]  # scipy.ndimage.imread returned RGB

    im1, im2 = _read(im1_fn), _read(im2_fn)
    H, W = im1.shape[:2]
    H_ = int(math.ceil(H / 64.0) * 64)
    W_ = int(math.ceil(W / 64.0) * 64)
    if im1.dtype == np.uint8 and im2.dtype == np.uint8:
        # decoded frames stay uint8 until they are on the device (a quarter of the upload of the float tensor)
        frames = np.ascontiguousarray(np.stack([cv2.resize(im[:, :, :3], (W_, H_)) for im in (im1, im2)])[None])
        flo = net.forward_host_u8(frames)[0]
    else:
        x, _ = preprocess_frames(im1, im2)
        flo = net.forward_host(x)[0]
    return postprocess_flow(flo, H, W, H_, W_)
