"""ctypes binding of include/premvos_b200.h.  Fails loudly: there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpremvos_b200.so")

# every symbol include/premvos_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "premvos_version", "premvos_last_error", "premvos_kernel_launch_count", "premvos_profile_begin",
    "premvos_profile_end",
    "premvos_corr_output_shape", "premvos_corr_forward", "premvos_conv2d_forward", "premvos_resize_linear_u8",
    "premvos_warp_masks_u8", "premvos_flow_postprocess", "premvos_pack_mask_bits",
    "premvos_resize_linear_u8_prepare", "premvos_resize_linear_u8_release", "premvos_flow_postprocess_prepare", "premvos_flow_postprocess_release",
    "premvos_pwc_create", "premvos_pwc_set_param", "premvos_pwc_finalize", "premvos_pwc_forward",
    "premvos_pwc_forward_host", "premvos_pwc_forward_host_u8", "premvos_pwc_forward_u8",
    "premvos_pwc_launches_per_forward", "premvos_pwc_set_option",
    "premvos_pwc_get_tensor", "premvos_pwc_destroy", "premvos_pwc_tensor_core_layers",
    "premvos_propnet_create", "premvos_propnet_set_option", "premvos_propnet_set_param", "premvos_propnet_finalize",
    "premvos_propnet_forward", "premvos_propnet_forward_u8", "premvos_propnet_read_results", "premvos_propnet_read_results_image", "premvos_propnet_copy_results", "premvos_propnet_forward_host",
    "premvos_propnet_read_masks", "premvos_fill_full_masks_host", "premvos_propnet_launches_per_forward", "premvos_propnet_get_tensor", "premvos_propnet_destroy",
    "premvos_topk_host", "premvos_nms_host",
    "premvos_refnet_create", "premvos_refnet_set_param", "premvos_refnet_finalize", "premvos_refnet_forward",
    "premvos_refnet_forward_host",
    "premvos_refnet_launches_per_forward", "premvos_refnet_get_tensor", "premvos_refnet_destroy",
    "premvos_sepconv2d_forward",
    "premvos_reidnet_create", "premvos_reidnet_set_param", "premvos_reidnet_finalize", "premvos_reidnet_forward",
    "premvos_reidnet_forward_host", "premvos_reidnet_launches_per_forward", "premvos_reidnet_get_tensor",
    "premvos_reidnet_destroy",
]

_lib = None


class PremvosError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("premvos_b200 error %d: %s" % (code, msg))
        self.code = code


def lib() -> ctypes.CDLL:
    """Load (once) and return the CUDA library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "premvos_b200: %s is missing. Build it with `python -m premvos_b200.build` (needs nvcc). "
            "There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_void_p, c_char_p, c_i64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64
    P = ctypes.POINTER
    L.premvos_version.restype = c_char_p
    L.premvos_last_error.restype = c_char_p
    L.premvos_kernel_launch_count.restype = c_i64
    L.premvos_profile_end.argtypes = [c_char_p, c_int]
    L.premvos_corr_output_shape.argtypes = [c_int] * 7 + [P(c_int)] * 3
    L.premvos_corr_forward.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]
    L.premvos_conv2d_forward.argtypes = [c_void_p] * 5 + [c_int] * 13 + [ctypes.c_float, c_void_p]
    L.premvos_resize_linear_u8.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.premvos_warp_masks_u8.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    L.premvos_flow_postprocess.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]
    L.premvos_pack_mask_bits.argtypes = [c_void_p, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p]
    L.premvos_pwc_create.argtypes = [P(c_void_p), c_int, c_int, c_int]
    L.premvos_pwc_set_param.argtypes = [c_void_p, c_char_p, c_void_p, c_i64]
    L.premvos_pwc_finalize.argtypes = [c_void_p]
    L.premvos_pwc_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    L.premvos_pwc_forward_host.argtypes = [c_void_p, c_void_p, c_void_p]
    L.premvos_pwc_forward_host_u8.argtypes = [c_void_p, c_void_p, c_void_p]
    L.premvos_pwc_forward_u8.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    L.premvos_pwc_launches_per_forward.argtypes = [c_void_p]
    L.premvos_pwc_tensor_core_layers.argtypes = [c_void_p]
    L.premvos_pwc_set_option.argtypes = [c_void_p, c_char_p, c_int]
    L.premvos_pwc_get_tensor.argtypes = [c_void_p, c_char_p, c_void_p, P(c_i64)]
    L.premvos_pwc_destroy.argtypes = [c_void_p]
    L.premvos_pwc_destroy.restype = None
    L.premvos_refnet_create.argtypes = [P(c_void_p), c_int, c_int, c_int]
    L.premvos_refnet_set_param.argtypes = [c_void_p, c_char_p, c_void_p, c_i64]
    L.premvos_refnet_finalize.argtypes = [c_void_p]
    L.premvos_refnet_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.premvos_refnet_forward_host.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.premvos_refnet_launches_per_forward.argtypes = [c_void_p]
    L.premvos_refnet_get_tensor.argtypes = [c_void_p, c_char_p, c_void_p, P(c_i64)]
    L.premvos_refnet_destroy.argtypes = [c_void_p]
    L.premvos_refnet_destroy.restype = None
    L.premvos_sepconv2d_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            ctypes.c_float, c_void_p]
    L.premvos_reidnet_create.argtypes = [P(c_void_p), c_int]
    L.premvos_reidnet_set_param.argtypes = [c_void_p, c_char_p, c_void_p, c_i64]
    L.premvos_reidnet_finalize.argtypes = [c_void_p]
    L.premvos_reidnet_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]
    L.premvos_reidnet_forward_host.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]
    L.premvos_reidnet_launches_per_forward.argtypes = [c_void_p]
    L.premvos_reidnet_get_tensor.argtypes = [c_void_p, c_char_p, c_void_p, P(c_i64)]
    L.premvos_reidnet_destroy.argtypes = [c_void_p]
    L.premvos_reidnet_destroy.restype = None
    L.premvos_topk_host.argtypes = [c_void_p, c_int, c_int, c_void_p, P(c_int)]
    L.premvos_nms_host.argtypes = [c_void_p, c_void_p, c_int, ctypes.c_float, c_int, c_void_p, P(c_int)]
    L.premvos_propnet_create.argtypes = [P(c_void_p), c_int, c_int, c_int, c_int]
    L.premvos_propnet_set_option.argtypes = [c_void_p, c_char_p, c_int]
    L.premvos_propnet_set_param.argtypes = [c_void_p, c_char_p, c_void_p, c_i64]
    L.premvos_propnet_finalize.argtypes = [c_void_p]
    L.premvos_propnet_forward.argtypes = [c_void_p, c_void_p, c_void_p]
    L.premvos_propnet_forward_u8.argtypes = [c_void_p, c_void_p, c_void_p]
    L.premvos_propnet_read_results.argtypes = [c_void_p, c_void_p, P(c_int)] + [c_void_p] * 6
    L.premvos_propnet_read_results_image.argtypes = [c_void_p, c_void_p, c_int, P(c_int)] + [c_void_p] * 6
    L.premvos_propnet_copy_results.argtypes = [c_void_p] * 6
    L.premvos_propnet_forward_host.argtypes = [c_void_p, c_void_p, P(c_int)] + [c_void_p] * 6
    L.premvos_propnet_read_masks.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int]
    L.premvos_fill_full_masks_host.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.premvos_propnet_launches_per_forward.argtypes = [c_void_p]
    L.premvos_propnet_get_tensor.argtypes = [c_void_p, c_char_p, c_void_p, P(c_i64)]
    L.premvos_propnet_destroy.argtypes = [c_void_p]
    L.premvos_propnet_destroy.restype = None
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        raise PremvosError(code, lib().premvos_last_error().decode("utf-8", "replace"))


def kernel_launch_count() -> int:
    return int(lib().premvos_kernel_launch_count())


def profile_begin() -> None:
    check(lib().premvos_profile_begin())


def profile_end() -> dict:
    """-> {kernel name: {"launches", "ms", "flops", "bytes"}} for everything launched since profile_begin()."""
    buf = ctypes.create_string_buffer(1 << 16)
    need = lib().premvos_profile_end(buf, len(buf))
    if need > 0:        # the report (per-layer labels) did not fit: the library kept it, ask again with the size it named
        buf = ctypes.create_string_buffer(need + 16)
        need = lib().premvos_profile_end(buf, len(buf))
    check(need)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms, fl, by = line.rsplit(" ", 4)
        out[name] = {"launches": int(cnt), "ms": float(ms), "flops": float(fl), "bytes": float(by)}
    return out
