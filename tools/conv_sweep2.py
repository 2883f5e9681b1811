"""Planner sweep (GPU box only): times conv_umma_kernel on given layer shapes for combinations of the planner overrides
PREMVOS_KC / PREMVOS_MT / PREMVOS_BUDGET_KB / PREMVOS_KSPLIT / PREMVOS_TPS (see conv_umma.cu:plan_conv_umma)."""
import itertools, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from premvos_b200 import _lib, ops  # noqa: E402

SHAPES = {
    "p_c2": (1, 256, 35, 83, 256, 3, 1, 1),
    "p_c3": (1, 256, 35, 83, 1024, 1, 1, 1),
    "p_c1": (1, 1024, 35, 83, 256, 1, 1, 1),
    "h_c2": (100, 512, 7, 7, 512, 3, 1, 1),
    "h_c3": (100, 512, 7, 7, 2048, 1, 1, 1),
    "h_c1": (100, 2048, 7, 7, 512, 1, 1, 1),
    "g1_c2": (1, 128, 71, 166, 128, 3, 1, 1),
    "g0_c3": (1, 64, 142, 333, 256, 1, 1, 1),
    "xc_mid": (20, 728, 25, 25, 728, 1, 1, 1),
    "xc_e2": (20, 256, 97, 97, 256, 1, 1, 1),
    "xc_e1": (20, 128, 193, 193, 128, 1, 1, 1),
    # the benchmarked launch group: 40 crops
    "xc40_mid": (40, 728, 25, 25, 728, 1, 1, 1),
    "xc40_x1": (40, 728, 25, 25, 1024, 1, 1, 1),
    "xc40_x2": (40, 1024, 25, 25, 1536, 1, 1, 1),
    "xc40_x3": (40, 1536, 25, 25, 2048, 1, 1, 1),
    "xc40_aspp": (40, 2048, 25, 25, 256, 1, 1, 1),
    "xc40_e2": (40, 256, 97, 97, 256, 1, 1, 1),
    "xc40_dec": (40, 304, 97, 97, 256, 1, 1, 1),
    "xc40_e3": (40, 728, 49, 49, 728, 1, 1, 1),
    # proposal network, batch 4 (568 x 1333 -> C4 35 x 83)
    "p4_c2": (4, 256, 35, 83, 256, 3, 1, 1),
    "p4_rpn": (4, 1024, 35, 83, 1024, 3, 1, 1),
    "p4_g1c2": (4, 128, 71, 166, 128, 3, 1, 1),
    "p4_c3": (4, 256, 35, 83, 1024, 1, 1, 1),
    "p4_c1": (4, 1024, 35, 83, 256, 1, 1, 1),
    "p4_g1c3": (4, 128, 71, 166, 512, 1, 1, 1),
    "p4_g0c3": (4, 64, 142, 333, 256, 1, 1, 1),
    "h4_c3": (400, 512, 7, 7, 2048, 1, 1, 1),
    "h4_c1": (400, 2048, 7, 7, 512, 1, 1, 1),
}
KEYS = ("PREMVOS_KC", "PREMVOS_TPS", "PREMVOS_MT", "PREMVOS_BUDGET_KB", "PREMVOS_KSPLIT", "PREMVOS_BN", "PREMVOS_DBG", "PREMVOS_NBUF", "PREMVOS_FLAT", "PREMVOS_CONV_REPEAT", "PREMVOS_LOCKSTEP", "PREMVOS_EPI8", "PREMVOS_TAIL", "PREMVOS_TAIL_SPLIT", "PREMVOS_PAIR", "PREMVOS_PAIR_STAGES", "PREMVOS_PAIR_MIN_ITEMS", "PREMVOS_PAIR_EPI16", "PREMVOS_STREAMK_ALL", "PREMVOS_PAIR_MIN_UNITS")


def run(name, env):
    N, Cin, H, W, Cout, k, stride, dil = SHAPES[name]
    for key in KEYS:
        os.environ.pop(key, None)
    os.environ.update({k2: str(v) for k2, v in env.items() if v is not None})
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, k, k) * 0.05
    pad = dil * (k // 2)
    try:
        ops.conv2d(x, w, None, stride, dil, (pad, pad, pad, pad), 0.1)
        _lib.profile_begin()
        for _ in range(5):
            ops.conv2d(x, w, None, stride, dil, (pad, pad, pad, pad), 0.1)
        prof = _lib.profile_end()
    except _lib.PremvosError as e:
        return None, str(e)[:70]
    r = prof["conv_umma_kernel"]
    us = r["ms"] * 1e3 / r["launches"]
    fin = prof.get("conv_finish_kernel")
    fus = fin["ms"] * 1e3 / fin["launches"] if fin else 0.0
    return us + fus, "%.0f TF/s (finish %.1f us)" % (r["flops"] / r["launches"] / (us + fus) / 1e6, fus)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in SHAPES] or list(SHAPES)
    grid = {"PREMVOS_KC": [None, 2, 4, 8], "PREMVOS_BUDGET_KB": [None, 220], "PREMVOS_KSPLIT": [None, 1, 2, 4], "PREMVOS_BN": [None]}
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("=")
            grid["PREMVOS_" + k] = [None if x == "-" else int(x) for x in v.split(",")]
    for name in names:
        print("==", name, SHAPES[name])
        res = []
        for combo in itertools.product(*grid.values()):
            env = dict(zip(grid.keys(), combo))
            us, note = run(name, env)
            res.append((us if us else 1e9, env, note))
        show = sorted(res, key=lambda r: r[0]) if len(res) <= 12 else sorted(res, key=lambda r: r[0])[:6] + [r for r in res if all(v is None for v in r[1].values())]
        for us, env, note in show:
            print("  %8.1f us  %s  %s" % (us, " ".join("%s=%s" % (k[8:], v) for k, v in env.items() if v is not None) or "(default)", note), flush=True)
