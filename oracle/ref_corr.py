"""Loader of oracle/_ref/libcorr_reference.so -- the REFERENCE's own correlation CUDA kernels
(correlation_package/src/corr_cuda_kernel.cu, compiled unchanged for sm_100a from where it lies under
/root/reference by `make -C oracle ref`) behind the thin non-THC host wrapper oracle/ref_corr_wrapper.cu,
which reproduces corr_cuda.c:23-78.  TEST INFRASTRUCTURE ONLY: the second oracle of the correlation op
and the "reference GPU" kernel the product's correlation is timed against.  Nothing under premvos_b200/
loads it.  The .so is git-ignored but travels to the GPU box with the snapshot; when it is absent
`available()` is False and the tests that need it skip."""
import ctypes
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libcorr_reference.so")
_LIB = None


def available() -> bool:
    return os.path.exists(_PATH)


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_PATH)
        L.ref_corr_cuda_forward.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 10 + [ctypes.c_void_p]
        L.ref_corr_cuda_forward.restype = ctypes.c_int
        _LIB = L
    return _LIB


def output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    """corr_cuda.c:23-45 (float ceil included)."""
    import math
    kr = (kernel_size - 1) // 2
    border = max_displacement + kr
    ow = math.ceil(float(W + 2 * pad_size - 2 * border) / float(stride1))
    oh = math.ceil(float(H + 2 * pad_size - 2 * border) / float(stride1))
    gw = 2 * (max_displacement // stride2) + 1
    return gw * gw, oh, ow


def corr_cuda_forward(input1, input2, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                      corr_multiply=1):
    """Runs the reference kernels on two contiguous fp32 NCHW CUDA tensors; returns the output tensor."""
    import torch
    assert input1.is_cuda and input1.dtype == torch.float32 and input1.is_contiguous()
    assert input2.is_cuda and input2.dtype == torch.float32 and input2.is_contiguous()
    B, C, H, W = input1.shape
    oc, oh, ow = output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    out = torch.empty((B, oc, oh, ow), dtype=torch.float32, device=input1.device)
    st = torch.cuda.current_stream().cuda_stream
    rc = _lib().ref_corr_cuda_forward(input1.data_ptr(), input2.data_ptr(), out.data_ptr(), B, C, H, W, pad_size,
                                      kernel_size, max_displacement, stride1, stride2, corr_multiply, st)
    if rc != 0:
        raise RuntimeError("reference correlation kernel failed: CUDA error %d" % rc)
    return out
