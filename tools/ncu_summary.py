"""Summarises ncu captures (run here, no GPU): per launch duration, DRAM bytes, tensor-pipe activity, SM activity.

    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] [--traffic profiles/r02_ncu_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor_el": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm_active": "sm__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "smem": "launch__shared_mem_per_block_dynamic", "grid": "launch__grid_size",
        "block": "launch__block_size", "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    args = sys.argv[1:]
    traffic_path = None
    if "--traffic" in args:
        i = args.index("--traffic")
        traffic_path = args[i + 1]
        args = args[:i] + args[i + 2:]
    traffic = {}
    for rep in args:
        hdr, units, rows = rows_of(rep)
        col = {k: (hdr.index(v) if v in hdr else None) for k, v in KEYS.items()}
        name_i = hdr.index("Kernel Name")
        print("== %s" % rep)
        for r in rows:
            def g(k, scale=1.0):
                try:
                    return float(r[col[k]].replace(",", "")) * scale
                except Exception:
                    return float("nan")
            def to_bytes(k):
                v, u = g(k), units[col[k]] if col[k] is not None else ""
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            name = r[name_i].split("(")[0].replace("premvos::<unnamed>::", "").replace("void ", "")
            du = g("dur") * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[col["dur"]], 1)
            rd, wr = to_bytes("rd"), to_bytes("wr")
            print("  %-34s grid %5d x %3d  %8.1f us  dram read %7.1f MB write %7.1f MB (%4.1f%%)  L2 %4.1f%%  tensor pipe %5.1f%% of active "
                  "(%5.1f%% of elapsed)  SM active %5.1f%%  issue %4.1f%%  regs %3d  smem %5.1f KB"
                  % (name[:34], g("grid"), g("block"), du, rd / 1e6, wr / 1e6, g("dram"), g("l2"), g("tensor"), g("tensor_el"), g("sm_active"),
                     g("issue"), g("regs"), to_bytes("smem") / 1e3))
            key = "conv_pair_kernel" if "conv_pair" in name else ("conv_umma_kernel" if "conv_umma" in name else
                                                                  ("depthwise3x3_kernel" if "depthwise" in name else name))
            traffic.setdefault(key, []).append(rd + wr)
    if traffic_path:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(traffic_path, "w"), indent=1)
        print("mean DRAM bytes per launch ->", traffic_path, {k: "%.1f MB over %d launches" % (sum(v) / len(v) / 1e6, len(v)) for k, v in traffic.items()})


if __name__ == "__main__":
    main()
