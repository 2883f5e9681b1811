"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, and rejects
bad arguments with the documented codes (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def L():
    from premvos_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_exports_every_symbol_the_header_declares(L):
    from premvos_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "premvos_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(premvos_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(L, sym), "libpremvos_b200.so does not export %s" % sym
    assert sorted(_lib.EXPORTS) == declared
    assert b"sm_100a" in L.premvos_version()


def test_library_is_blackwell_only():
    import subprocess
    from premvos_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_errors(L):
    from premvos_b200 import _lib
    h = ctypes.c_void_p()
    assert L.premvos_pwc_create(ctypes.byref(h), 1, 100, 128) == -1      # not a multiple of 64
    assert b"multiples of 64" in L.premvos_last_error()
    assert L.premvos_pwc_create(ctypes.byref(h), 0, 64, 64) == -1
    assert L.premvos_pwc_create(None, 1, 64, 64) == -1
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.premvos_corr_output_shape(436, 1024, 4, 1, 4, 1, 1, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (81, 436, 1024)
    assert L.premvos_corr_output_shape(2, 2, 0, 1, 4, 1, 1, None, None, None) == -1   # empty output
    assert L.premvos_corr_output_shape(8, 8, 4, 2, 4, 1, 1, None, None, None) == -1   # even kernel
    assert L.premvos_corr_forward(None, None, None, 1, 1, 2, 2, 0, 1, 0, 1, 1, 1, None) == -1
    # the fused separable-convolution hook: null pointers, then (with dummy non-null pointers) more than one output-channel tile
    assert L.premvos_sepconv2d_forward(None, None, None, None, None, None, 1, 32, 8, 8, 32, 0, 0, 1.0, None) == -1
    buf = (ctypes.c_float * 4)()
    assert L.premvos_sepconv2d_forward(buf, buf, None, buf, None, buf, 1, 32, 8, 8, 256, 0, 0, 1.0, None) == -2
    assert b"cout <= 128" in L.premvos_last_error()
    h2 = ctypes.c_void_p()
    assert L.premvos_reidnet_create(ctypes.byref(h2), 0) == -1 and L.premvos_reidnet_create(None, 4) == -1
    with pytest.raises(_lib.PremvosError):
        _lib.check(L.premvos_pwc_forward(None, None, None, None))


def test_argument_errors_of_the_byte_work_entry_points(L):
    # argument checks come before any CUDA call: they must answer on a machine without a GPU
    assert L.premvos_resize_linear_u8(None, 1, 8, 8, None, 16, 16, 3, 0, None) == -1
    assert L.premvos_warp_masks_u8(None, 2, 8, 8, None, None, None, 1, None) == -1
    assert L.premvos_warp_masks_u8(None, 0, 8, 8, None, None, None, 1, None) == 0        # no masks: a no-op
    assert L.premvos_warp_masks_u8(None, -1, 8, 8, None, None, None, 1, None) == -1
    assert L.premvos_flow_postprocess(None, 1, 64, 64, None, 60, 60, None) == -1
    assert L.premvos_fill_full_masks_host(None, None, 0, 14, 32, 32, None) == 0           # no boxes: a no-op
    assert L.premvos_fill_full_masks_host(None, None, 2, 14, 32, 32, None) == -1
    assert L.premvos_fill_full_masks_host(None, None, 0, 13, 32, 32, None) == -1          # odd mask size
    h = ctypes.c_void_p()
    assert L.premvos_propnet_create(ctypes.byref(h), 128, 160, 2, 81) in (0, -6)          # -6: no device visible here
    if h.value:
        assert L.premvos_propnet_set_option(h, b"batch", 0) == -1 and L.premvos_propnet_set_option(h, b"batch", 17) == -1
        assert L.premvos_propnet_set_option(h, b"batch", 4) == 0 and L.premvos_propnet_set_option(h, b"mode_mask", 1) == 0
        assert L.premvos_propnet_set_option(h, b"no_such_option", 1) == -1
        L.premvos_propnet_destroy(h)


def test_corr_shapes_match_oracle_shape_math(L):
    from oracle import pwc_oracle as O
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for (h, w, pad, k, md, s1, s2) in [(7, 16, 4, 1, 4, 1, 1), (100, 100, 40, 1, 40, 1, 1), (48, 64, 20, 3, 20, 2, 2),
                                       (2, 2, 1, 1, 1, 1, 1), (9, 11, 3, 1, 3, 1, 1)]:
        assert L.premvos_corr_output_shape(h, w, pad, k, md, s1, s2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
        assert (1, oc.value, oh.value, ow.value) == O.correlation_output_shape(1, 1, h, w, pad, k, md, s1, s2)


def test_host_mirror_surface_without_gpu():
    import torch
    from premvos_b200 import pwc, synth
    net = pwc.pwc_dc_net(None)
    sd = net.state_dict()
    assert list(sd.keys()) == list(synth.pwc_param_shapes().keys())
    net.load_state_dict(synth.pwc_synthetic_state_dict(0))
    with pytest.raises(RuntimeError):
        net.load_state_dict({"bogus": torch.zeros(1)})
    bad = dict(sd)
    bad["conv1a.0.weight"] = torch.zeros(1, 3, 3, 3)
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad)
    with pytest.raises(RuntimeError):           # no CPU fallback
        net(torch.zeros(1, 6, 64, 64))
    with pytest.raises(RuntimeError):
        pwc.Correlation(4, 1, 4, 1, 1, 1)(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    assert net.eval() is net


def test_product_package_never_imports_oracle():
    import subprocess
    import sys
    code = "import sys; import premvos_b200, premvos_b200.pwc, premvos_b200.shard; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "premvos_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
