"""GPU parity tests (run on a B200: `pytest -m gpu`).  Every call goes through the C ABI
(premvos_b200/_lib.py -> libpremvos_b200.so); the CPU oracle is only the checker.

Tolerance (BASELINE.json north_star): <= 1e-3 relative, measured as ||d||_inf / ||ref||_inf per tensor.
The fp32 SIMT mode is held to 1e-4, the tensor-core mode (split-bf16 x3, fp32 accumulate) to 1e-3."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import pwc_oracle as O
from premvos_b200 import _lib, pwc, synth

pytestmark = pytest.mark.gpu

TOL = 1e-3
TOL_FP32 = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert torch.cuda.get_device_capability(0)[0] == 10, "sm_100a library needs a Blackwell GPU"
    _lib.lib()   # raises if the CUDA library is missing -- no fallback


# ---------------------------------------------------------------------------------------------
# correlation through premvos_corr_forward
# ---------------------------------------------------------------------------------------------
def _corr_case(shape, cfg, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    ref = O.correlation_forward(a, b, *cfg, 1)
    got = pwc.correlation_forward(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), *cfg, 1).cpu().numpy()
    return got, ref


def test_corr_kat_from_reference_test():
    A = torch.tensor([[1., 2.], [3., 4.]]).view(1, 1, 2, 2).cuda()
    B = torch.tensor([[5., 6.], [7., 8.]]).view(1, 1, 2, 2).cuda()
    y = pwc.Correlation(0, 1, 0, 1, 1, 1)(A, B).cpu().numpy()
    np.testing.assert_array_equal(y[0, 0], np.array([[5, 12], [21, 32]], dtype=np.float32))
    y2 = pwc.Correlation(1, 1, 1, 1, 1, 1)(A, B)
    assert tuple(y2.shape) == (1, 9, 2, 2)
    np.testing.assert_array_equal(y2.cpu().numpy()[0, 4], y[0, 0])


# BASELINE config C1: the five PWC levels of a 256x256 pair + the stress shape, + ragged edges
@pytest.mark.parametrize("shape", [(1, 196, 4, 4), (1, 128, 8, 8), (1, 96, 16, 16), (1, 64, 32, 32), (1, 32, 64, 64),
                                   (1, 32, 256, 256), (2, 33, 5, 35), (1, 1, 1, 1), (3, 7, 13, 70)])
def test_corr_fast_path_matches_oracle(shape):
    got, ref = _corr_case(shape, (4, 1, 4, 1, 1))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 2e-6


@pytest.mark.parametrize("shape,cfg", [((1, 8, 12, 10), (3, 1, 3, 1, 1)), ((2, 5, 9, 9), (2, 3, 2, 1, 1)),
                                       ((1, 4, 16, 12), (4, 1, 4, 2, 2)), ((1, 6, 10, 14), (6, 1, 4, 1, 2)),
                                       ((1, 16, 20, 20), (20, 1, 20, 1, 2))])
def test_corr_generic_path_matches_oracle(shape, cfg):
    got, ref = _corr_case(shape, cfg, seed=3)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 2e-6


REF_CASES = [((1, 196, 4, 4), (4, 1, 4, 1, 1)), ((1, 128, 8, 8), (4, 1, 4, 1, 1)), ((1, 96, 16, 16), (4, 1, 4, 1, 1)),
             ((1, 64, 32, 32), (4, 1, 4, 1, 1)), ((1, 32, 64, 64), (4, 1, 4, 1, 1)), ((1, 32, 256, 256), (4, 1, 4, 1, 1)),
             ((2, 33, 5, 35), (4, 1, 4, 1, 1)), ((1, 32, 112, 256), (4, 1, 4, 1, 1)),
             ((1, 8, 12, 10), (3, 1, 3, 1, 1)), ((2, 5, 9, 9), (2, 3, 2, 1, 1)), ((1, 4, 16, 12), (4, 1, 4, 2, 2)),
             ((1, 16, 20, 20), (20, 1, 20, 1, 2))]


@pytest.mark.parametrize("shape,cfg", REF_CASES)
def test_corr_matches_reference_cuda_kernel(shape, cfg):
    """Second oracle: the reference's own corr_cuda_kernel.cu, compiled unchanged for sm_100a into oracle/_ref
    (blob_rearrange x2 + CorrelateData + memsets as corr_cuda.c:52-78 drives them), on the same inputs -- pins
    both the product kernel and the CPU restatement to outputs of the reference itself."""
    from oracle import ref_corr
    if not ref_corr.available():
        pytest.skip("oracle/_ref/libcorr_reference.so not built (needs /root/reference at build time)")
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(*shape, device="cuda", generator=g)
    b = torch.randn(*shape, device="cuda", generator=g)
    ref = ref_corr.corr_cuda_forward(a, b, *cfg, 1).cpu().numpy()
    got = pwc.correlation_forward(a, b, *cfg, 1).cpu().numpy()
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 2e-6
    if a.numel() <= 1 << 18:
        cpu = O.correlation_forward(a.cpu().numpy(), b.cpu().numpy(), *cfg, 1)
        assert rel_err(cpu, ref) < 2e-6


def test_corr_properties_full_size():
    # size-independent properties at the bench resolution (level 2 of a 448x1024 pair)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(1, 32, 112, 256, device="cuda", generator=g)
    b = torch.randn(1, 32, 112, 256, device="cuda", generator=g)
    c = pwc.correlation_forward(a, b)
    # centre displacement is the per-pixel channel mean of the product
    assert torch.allclose(c[:, 40], (a * b).mean(1), rtol=1e-5, atol=1e-6)
    # linearity in the second argument
    c2 = pwc.correlation_forward(a, 2.5 * b)
    assert torch.allclose(c2, 2.5 * c, rtol=1e-5, atol=1e-6)
    # shift property: displacing b by (dy,dx) moves the centre plane to channel (dy+4)*9+(dx+4)
    bs = torch.zeros_like(b)
    bs[:, :, 2:, :-3] = b[:, :, :-2, 3:]          # bs(y,x) = b(y-2, x+3)
    cs = pwc.correlation_forward(a, bs)
    ch = (2 + 4) * 9 + (-3 + 4)
    assert torch.allclose(cs[:, ch, 4:-4, 4:-4], c[:, 40, 4:-4, 4:-4], rtol=1e-5, atol=1e-6)
    with pytest.raises(_lib.PremvosError):
        pwc.correlation_forward(a, b, corr_multiply=0)


# ---------------------------------------------------------------------------------------------
# PWC-DC-Net forward through premvos_pwc_*
# ---------------------------------------------------------------------------------------------
def _net(seed, **kw):
    net = pwc.pwc_dc_net(None, **kw)
    net.load_state_dict(synth.pwc_synthetic_state_dict(seed))
    return net.cuda().eval()


MODES = [pytest.param(False, id="fp32-simt"), pytest.param(True, id="tensor-core")]


@pytest.mark.parametrize("tc", MODES)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_pwc_matches_reference_goldens(golden_dir, tag, tc):
    g = np.load(os.path.join(golden_dir, "pwc_golden_%s.npz" % tag))
    net = _net(int(g["weight_seed"]), tensor_cores=tc)
    x = synth.synthetic_pwc_input(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["input_seed"]))
    flow = net(torch.from_numpy(x).cuda()).cpu().numpy()
    assert flow.shape == g["flow2"].shape
    assert rel_err(flow, g["flow2"]) < (TOL if tc else TOL_FP32)


@pytest.mark.parametrize("tc", MODES)
def test_pwc_intermediates_match_oracle(tc):
    B, H, W = 2, 128, 192
    sd = synth.pwc_synthetic_state_dict(11)
    net = pwc.pwc_dc_net(None, tensor_cores=tc)
    net.load_state_dict(sd)
    net.cuda()
    x = synth.synthetic_pwc_input(B, H, W, seed=5)
    flow = net(torch.from_numpy(x).cuda()).cpu().numpy()
    ref, inter = O.pwc_forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x), True)
    report = []
    for name, t in inter.items():
        if name in ("flow2",):
            continue
        key = "flow2" if name == "flow2_pre" else name
        got = net.get_tensor(key, B, H, W).reshape(t.shape)
        report.append((name, rel_err(got, t.numpy())))
    report.append(("flow2_final", rel_err(flow, ref.numpy())))
    print("\n".join("%-12s %.3e" % r for r in report))
    tol = TOL if tc else TOL_FP32
    bad = [r for r in report if not (r[1] < tol)]
    assert not bad, bad


@pytest.mark.parametrize("tc", MODES)
def test_pwc_full_size_against_oracle_and_properties(tc):
    # BASELINE config C2: 1024x436 -> 448x1024
    B, H, W = 1, 448, 1024
    sd = synth.pwc_synthetic_state_dict(0)
    net = pwc.pwc_dc_net(None, tensor_cores=tc)
    net.load_state_dict(sd)
    net.cuda()
    x = synth.synthetic_pwc_input(2, H, W, seed=1)
    xd = torch.from_numpy(x).cuda()
    f0 = net(xd[:1])
    f0b = net(xd[:1])
    assert torch.equal(f0, f0b), "forward is not deterministic"
    # batch independence: a batch-2 handle gives the same per-pair results.  In tensor-core mode the planner may pick a
    # different (deterministic) split-K factor for the low pyramid levels at another batch size, i.e. another fp32
    # summation order -- equal within rounding, not bit for bit.
    btol = 1e-4 if tc else 1e-6
    f01 = net(xd)
    assert rel_err(f01[:1].cpu().numpy(), f0.cpu().numpy()) < btol
    f1 = net(xd[1:])
    assert rel_err(f01[1:].cpu().numpy(), f1.cpu().numpy()) < btol
    # host entry point == device entry point
    fh = net.forward_host(x[:1].copy())
    assert np.array_equal(fh, f0.cpu().numpy())
    # CUDA graph on/off gives identical bits
    net2 = pwc.pwc_dc_net(None, tensor_cores=tc, cuda_graph=False)
    net2.load_state_dict(sd)
    net2.cuda()
    assert torch.equal(net2(xd[:1]), f0)
    # and the oracle at full size
    torch.set_num_threads(os.cpu_count() or 1)
    ref = O.pwc_forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x[:1])).numpy()
    assert rel_err(f0.cpu().numpy(), ref) < (TOL if tc else TOL_FP32)


def test_pwc_benchmarked_plan_batch4_full_size():
    # the launch plan bench.py times: tensor-core mode, batch 4 at 448x1024 (the planner picks tile shapes / split-K by batch
    # size, so this is another set of kernels than the batch-1 / batch-2 handles above); every pair against the oracle
    B, H, W = 4, 448, 1024
    sd = synth.pwc_synthetic_state_dict(0)
    net = pwc.pwc_dc_net(None, tensor_cores=True)
    net.load_state_dict(sd)
    net.cuda()
    x = synth.synthetic_pwc_input(B, H, W, seed=11)
    got = net(torch.from_numpy(x).cuda()).cpu().numpy()
    torch.set_num_threads(os.cpu_count() or 1)
    P = {k: torch.from_numpy(v) for k, v in sd.items()}
    for i in range(B):
        ref = O.pwc_forward(P, torch.from_numpy(x[i:i + 1])).numpy()
        assert rel_err(got[i:i + 1], ref) < TOL, i


def test_tma_staged_cp8_correlation_gives_the_same_flow(monkeypatch):
    # the opt-in TMA-staged correlation of the CP8 path (csrc/corr_tma.cu, PREMVOS_CORR_CP8_TMA=1): same products, same channel
    # order per accumulator -> the same bits as the default kernel, on a size with ragged tiles and five pyramid levels
    B, H, W = 2, 192, 320
    sd = synth.pwc_synthetic_state_dict(4)
    x = torch.from_numpy(synth.synthetic_pwc_input(B, H, W, seed=12)).cuda()
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("PREMVOS_CORR_CP8_TMA", flag)
        net = pwc.pwc_dc_net(None, tensor_cores=True)
        net.load_state_dict(sd)
        net.cuda()
        outs.append(net(x).cpu().numpy())
        del net
    assert np.array_equal(outs[0], outs[1])


def test_calculate_flow_end_to_end():
    f1, f2 = synth.synthetic_frame_pair(100, 150, seed=3)
    sd = synth.pwc_synthetic_state_dict(2)
    net = _net(2)
    got = pwc.calculate_flow(net, f1, f2)
    assert got.shape == (100, 150, 2) and got.dtype == np.float32
    x, meta = O.preprocess_pair(f1, f2)
    ref = O.postprocess_flow(O.pwc_forward({k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x)).numpy()[0], *meta)
    assert rel_err(got, ref) < TOL


def test_forward_host_u8_equals_float_path():
    f1, f2 = synth.synthetic_frame_pair(128, 192, seed=4)
    net = _net(3)
    x, _ = pwc.preprocess_frames(f1, f2)                       # the float tensor the reference uploads
    frames = np.ascontiguousarray(np.stack([f1, f2])[None])
    a = net.forward_host(x)
    b = net.forward_host_u8(frames)
    np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        net.forward_host_u8(frames.astype(np.float32))


def test_errors_are_loud():
    net = _net(0)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 6, 100, 128, device="cuda"))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 5, 64, 64, device="cuda"))
    with pytest.raises(TypeError):
        net(torch.zeros(1, 6, 64, 64, device="cuda", dtype=torch.float16))
    assert _lib.kernel_launch_count() > 0
