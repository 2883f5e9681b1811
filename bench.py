#!/usr/bin/env python
"""Benchmark of the PReMVOS per-frame hot path on B200 (contract: task statement section 4 / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pairs B] [--boxes K]

Workload = BASELINE.json's metric, "frame-pairs/sec (flow+proposal+refine) 1024x436": for every synthetic Sintel-shaped
frame pair (t, t+1) of 1024x436 pixels
    flow       PWC-Net full forward on the pair (448x1024 network input)                       BASELINE configs[1]
    proposals  proposal network (ResNet-101 C4, 100 RoIs) on frame t+1 with BOTH weight sets    simple_run.sh:27-42
               (568x1333 after CustomResize)
    refine     refinement network (DeepLabv3+ / Xception-65, 385x385 crops) on 40 boxes of frame t+1 = the most the two
               proposal passes can hand over (2 x RESULTS_PER_IM); the boxes are fixed synthetic ones because the detections
               of random-init heads are data dependent -- the work per step stays constant
random-init weights of the reference architectures.  A step = `--pairs` frame pairs per GPU through
premvos_b200.pipeline.FramePipeline; the metric is whole-job frame pairs per second.

  value    : device-resident ORIGINAL frames (the stage drivers' cv2.resize calls run on the device), K steps timed with CUDA events on the launching stream, barrier + synchronize on
             both sides, max over ranks.
  e2e      : the same through FramePipeline.run_frames_host: pinned HOST uint8 frames t, t+1 in, flow + detections + masks +
             conf_scores out to pinned host memory, one synchronisation per step; copies inside the timed region.
  stages   : each network alone, device resident (flow alone is BASELINE configs[1], proposals configs[2], refine configs[3]).
  roofline : dominant kernel of the step (by device time) from a per-launch CUDA-event pass over one step run serially
             on one stream (graphs bypassed so each launch can be bracketed): algorithmic FLOPs / measured time vs
             MEASURED_PEAKS.json.
  cpu_baseline / --impl reference : the CPU oracle restatement of the reference forwards on all host threads.  A step of the
             reference arm is ONE full unit actually executed (1 flow pair + 2 proposal passes + all `--boxes` refinement crops,
             one at a time as the reference iterates) -- the bounded sample of the product arm's 4-unit step; --steps / --warmup
             are honoured as given.  (The reference's own CPU path cannot run: its CPU correlation is a stub, warp() hard-codes
             .cuda(), TensorFlow 1.8 / tensorpack are not installable.)
  N > 1    : every rank packs the results of its step (flow, detections, conf_scores, bit-packed masks: pipeline.pack_step_results)
             into one flat buffer and gathers it to rank 0 with ONE NCCL collective per step (shard.gather_tensor, asynchronous, two
             buffers in flight) INSIDE the timed region: `value` is gather-inclusive.
  e2e_detected : the product mode of the pipeline, run_frames_host(boxes=None): every unit refines the boxes its own two proposal
             passes detect (a device->host read of the counts mid-step; data-dependent work, mean boxes per frame reported).
  --c5     : BASELINE configs[4]: a 90-frame 854x480 synthetic video, units sharded round-robin over the ranks, detected boxes,
             per-step results gathered to rank 0; value = frame pairs of the video / wall time of the slowest rank.
  library_baseline : the oracle graphs on the GPU through torch / cuDNN in fp32 and TF32 (baseline/library_baseline.py) -- what a
             plain library implementation of the same arithmetic does on this GPU; context, not the product path.
  c1_correlation : BASELINE configs[0] on the device: the cost-volume kernel alone, GB/s of algorithmic bytes vs the HBM peak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IN, W_IN = 436, 1024
METRIC = "frame-pairs/sec (flow+proposal+refine) 1024x436"
UNIT = "frame-pairs/s"
WORKLOAD = ("per frame pair 1024x436: PWC-Net full forward (448x1024 input) + proposal net x2 weight sets (ResNet-101 C4, "
            "568x1333 input, 100 RoIs) + refinement net on 40 boxes (DeepLabv3+ Xception-65, 385x385 crops)")
# algorithmic GFLOP per unit (2*MAC of every convolution of the reference graphs)
GFLOP_FLOW, GFLOP_CROP = 168.2, 61.8


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"],
                "tensor_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "source": "fallback"}


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of the benchmarked step
    (profiles/r02_ncu_traffic.json, else round 1's), or None."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p)).get(kernel)
    return None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 100 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_units(n_units, boxes_per_frame):
    """`n_units` distinct synthetic units: (frame t, frame t+1) uint8 RGB [436,1024,3] as a decoder hands them over, and the
    boxes of frame t+1 (float32 [K,4] xywh)."""
    from premvos_b200 import synth
    f1, f2 = synth.synthetic_frame_pair(H_IN, W_IN, seed=1)
    units = []
    for k in range(n_units):
        a = np.ascontiguousarray(np.roll(f1, shift=(7 * k, 13 * k), axis=(0, 1)))
        b = np.ascontiguousarray(np.roll(f2, shift=(7 * k, 13 * k), axis=(0, 1)))
        units.append((a, b, synth.synthetic_boxes(boxes_per_frame, H_IN, W_IN, seed=100 + k)))
    return units


def oracle_unit_times(repeats, warmup, boxes_per_frame):
    """`repeats` timed FULL units with the CPU oracle on all host threads, one at a time as the reference iterates: 1 PWC forward +
    2 proposal-net forwards (general, specific) + `boxes_per_frame` refinement crops.  -> (seconds of every timed unit, cores, parts)"""
    import cv2
    import torch
    from oracle import propnet_oracle as PO, pwc_oracle as O, refnet_oracle as RO
    from premvos_b200 import pipeline, propnet, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Hn, Wn = pipeline.flow_input_shape(H_IN, W_IN)
    Hp, Wp = propnet.custom_resize_shape(H_IN, W_IN)
    sd = {k: torch.from_numpy(v) for k, v in synth.pwc_synthetic_state_dict(0).items()}
    x = torch.from_numpy(synth.synthetic_pwc_input(1, Hn, Wn, seed=1))
    PPs = [synth.propnet_synthetic_params(1), synth.propnet_synthetic_params(4)]
    img = cv2.resize(synth.synthetic_bgr_frame(H_IN, W_IN, seed=2), (Wp, Hp)).astype(np.float32)
    RP = synth.refnet_synthetic_params(2)
    frame = synth.synthetic_bgr_frame(H_IN, W_IN, seed=3)
    boxes = synth.synthetic_boxes(boxes_per_frame, H_IN, W_IN, seed=3)
    image = (frame / 255).astype(np.float32)

    def crops():
        for b in boxes:
            inputs, crop = RO.make_network_input(image, b, 385)
            logits = RO.deeplab_logits(RP, inputs[None])[0]
            RO.segmentation_output(logits, crop, H_IN, W_IN, 385)

    units, parts = [], {"flow": [], "proposal_pass": [], "refine_crop": []}
    for it in range(warmup + repeats):
        t0 = time.perf_counter(); O.pwc_forward(sd, x)
        t1 = time.perf_counter(); PO.propnet_forward(PPs[0], img); PO.propnet_forward(PPs[1], img)
        t2 = time.perf_counter(); crops()
        t3 = time.perf_counter()
        if it >= warmup:
            units.append(t3 - t0)
            parts["flow"].append(t1 - t0); parts["proposal_pass"].append((t2 - t1) / 2); parts["refine_crop"].append((t3 - t2) / len(boxes))
    return units, cores, {k: float(np.mean(v)) for k, v in parts.items()}


def sample_text(repeats, parts, boxes_per_frame):
    return ("oracle (torch CPU, all host threads), %d full unit(s) executed: 1 PWC forward 448x1024 (%.2f s) + 2 proposal-net forwards "
            "568x1333 (%.2f s each) + %d refinement crops 385x385 one at a time (%.3f s each); nothing extrapolated"
            % (repeats, parts["flow"], parts["proposal_pass"], boxes_per_frame, parts["refine_crop"]))


def bench_config(B, K, world, refine_batch, l2_note):
    return {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "global_pairs_per_step": B * world, "boxes_per_frame": K,
            "refine_batch": refine_batch or K, "parallelism": "dp%d" % world,
            "weights": "seeded random init of the reference architectures (PWC-DC-Net 9.4M, 2 x ResNet-101 C4 51.9M, Xception-65 "
                       "DeepLabv3+ 40.8M params)",
            "l2": l2_note}


def l2_note(B, K):
    per_set = B * (2 * H_IN * W_IN * 3 + K * 16)
    sets = max(2, -(-(2 * 126 << 20) // per_set))
    return sets, per_set, ("inputs rotate over %d device input sets (%d MB > 126 MB L2); per-step activations are > 10 GB"
                           % (sets, sets * per_set >> 20))


def run_reference(args, rank):
    """The reference arm: K timed steps after W warm-up steps, a step = one full unit on the host cores (see the module docstring)."""
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    units, cores, parts = oracle_unit_times(steps, warm, args.boxes)
    sec = float(np.mean(units))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.pairs, args.boxes, max(1, args.gpus), args.refine_batch, l2_note(args.pairs, args.boxes)[2]),
            "reference_note": "CPU oracle port of the reference forwards on the host cores; a step of this arm is one unit (one frame "
                              "pair) of the product arm's %d-pair step; the reference's own CPU path cannot run (corr.c is a stub, "
                              "warp() hard-codes .cuda(), TF 1.8 / tensorpack not installable)" % args.pairs,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_text(steps, parts, args.boxes)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def c1_correlation(peak_gbs):
    """BASELINE configs[0] on the device: the cost-volume operator alone (pad 4, md 4) on the stress shape [1,32,256,256], on PWC
    level 2 of the benchmarked step [4,32,112,256] and on the 256x256 pyramid; CUDA events, inputs rotated over > L2 of buffers."""
    import torch
    from premvos_b200 import pwc
    corr = pwc.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    out = {}
    for name, shapes in (("stress_1x32x256x256", [(1, 32, 256, 256)]), ("pwc_level2_4x32x112x256", [(4, 32, 112, 256)]),
                         ("pyramid_256x256", [(1, 196, 4, 4), (1, 128, 8, 8), (1, 96, 16, 16), (1, 64, 32, 32), (1, 32, 64, 64)])):
        nbytes = sum(4 * (2 * c + 81) * h * w * b for b, c, h, w in shapes)
        sets = max(2, min(64, (2 * 126 << 20) // max(nbytes, 1) + 1))
        bufs = [[(torch.randn(s, device="cuda"), torch.randn(s, device="cuda")) for s in shapes] for _ in range(sets)]
        def run(i):
            for a, b in bufs[i % sets]:
                corr(a, b)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        iters = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        out[name] = {"us": us, "algorithmic_bytes": nbytes, "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak_gbs,
                     "launches": len(shapes)}
    out["note"] = ("premvos_corr_forward through pwc.Correlation (includes the operator's output allocation); bytes = 4*(2C+81)*H*W per "
                   "image (SURVEY 8d C1); the pyramid is launch-latency bound (5 launches of <= 0.7 MB)")
    return out


def run_c5(args, rank, local_rank, world):
    """BASELINE configs[4] / SURVEY 8(d) C5: a 90-frame 854x480 synthetic video through stages 1, 2, 3, 5 (simple_run.sh:19-61), units
    (frame pairs) sharded round-robin over the ranks (shard.shard_units), every unit refining its own detections, the per-step
    results gathered to rank 0 (bit-packed masks, one NCCL gather per step).  Timed: whole video, wall clock of the slowest rank,
    decode-time host preparation excluded (frames are resident in pinned host memory, as a decoder would hand them over)."""
    import torch
    import torch.distributed as dist
    from premvos_b200 import pipeline, shard, synth
    Hc, Wc = 480, 854
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    def weights(make):
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in make().items()} if rank == 0 else {}
        return shard.broadcast_state_dict(sd, src=0)
    sd_flow = weights(lambda: synth.pwc_synthetic_state_dict(0))
    P_gen = {k: v.numpy() for k, v in weights(lambda: synth.propnet_synthetic_params(8)).items()}
    P_spec = {k: v.numpy() for k, v in weights(lambda: synth.propnet_synthetic_params(3)).items()}
    P_ref = {k: v.numpy() for k, v in weights(lambda: synth.refnet_synthetic_params(2)).items()}
    B, K = args.pairs, args.boxes
    pipe = pipeline.FramePipeline(sd_flow, P_gen, P_spec, P_ref, (Hc, Wc), pairs_per_step=B, boxes_per_frame=K)
    f1, f2 = synth.synthetic_frame_pair(Hc, Wc, seed=1)
    frames = [np.ascontiguousarray(np.roll(f1 if t % 2 == 0 else f2, shift=(3 * t, 5 * t), axis=(0, 1))) for t in range(args.frames)]
    units = shard.shard_units(len(frames) - 1, rank, world)
    steps = [units[i:i + B] for i in range(0, len(units), B)]
    steps = [c + [c[-1]] * (B - len(c)) for c in steps]          # the last step of a shard is padded by repeating its last unit
    n_steps_all = -(-(-(-(len(frames) - 1) // world)) // B)      # every rank runs the same number of collectives
    prev = [torch.from_numpy(np.stack([frames[t] for t in c])).pin_memory() for c in steps]
    cur = [torch.from_numpy(np.stack([frames[t + 1] for t in c])).pin_memory() for c in steps]
    flats = [pipe.pack_step_results() for _ in range(2)]

    def run_video():
        pending = [None, None]
        nb = 0
        for i in range(n_steps_all):
            j = i & 1
            if i < len(steps):
                r = pipe.run_frames_host(prev[i], cur[i])
                nb += int(np.sum(r["num_boxes"][:len(set(steps[i]))]))
            if pending[j] is not None and pending[j][1] is not None:
                pending[j][1].wait()
            pipe.pack_step_results(out=flats[j])
            pending[j] = shard.gather_tensor(flats[j], dst=0, async_op=True)
        for p in pending:
            if p is not None and p[1] is not None:
                p[1].wait()
        torch.cuda.synchronize()
        return nb

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()

    run_video()                                                   # warm-up: one whole pass
    reps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    nb = 0
    for _ in range(reps):
        nb += run_video()
    barrier()
    sec = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([sec, float(nb)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        sec, nb = float(tmax[0]), float(t[1])
    pairs = (len(frames) - 1) * reps
    if rank == 0:
        print(json.dumps({"metric": "frame-pairs/sec, end-to-end 90-frame 854x480 video (flow + 2 x proposals + refinement of the detected "
                                    "boxes), sharded over the ranks, results gathered to rank 0",
                          "value": pairs / sec, "unit": UNIT, "n_gpus": world, "steps": reps, "warmup": 1, "ms_per_step": sec / reps * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "bf16x3 split-fp32 (fp32 accumulate)", "data": "synthetic",
                          "config": {"workload": "BASELINE configs[4] (SURVEY 8d C5): %d frames 854x480, %d frame pairs, pairs_per_step %d, "
                                                 "up to %d boxes per frame from the unit's own detections" % (len(frames), len(frames) - 1, B, K),
                                     "mean_boxes_refined_per_frame": nb / pairs, "parallelism": "dp%d" % world,
                                     "gather_bytes_per_rank_per_step": int(flats[0].numel()),
                                     "timing": "wall clock of the slowest rank over the whole video incl. host<->device copies and the "
                                               "NCCL gathers; a step of this line = one pass over the video"}}), flush=True)
    if distributed:
        dist.destroy_process_group()


class stdout_to_stderr:
    """NCCL prints its version banner on the process's stdout when the communicator is created (first collective); the contract is ONE
    JSON line on stdout, so file descriptor 1 points at stderr while the process group comes up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def time_stage(fn, iters, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=4, help="frame pairs per GPU per step")
    ap.add_argument("--boxes", type=int, default=40, help="refinement boxes per frame")
    ap.add_argument("--refine-batch", type=int, default=0, help="refinement crops per launch group (0 = all boxes of a frame)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-network context timings")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the torch / cuDNN context timings")
    ap.add_argument("--no-reid", action="store_true", help="skip the ReID network's context timing (stages.reid)")
    ap.add_argument("--c5", action="store_true", help="BASELINE configs[4]: 90-frame 854x480 video sharded over the ranks")
    ap.add_argument("--frames", type=int, default=90, help="frames of the --c5 video")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.c5:
        return run_c5(args, rank, local_rank, world)

    import torch
    import torch.distributed as dist
    from premvos_b200 import _lib, pipeline, shard, synth

    assert torch.cuda.is_available(), "bench.py needs a GPU; the product has no CPU path"
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()                                        # creates the communicator (and prints NCCL's banner) now
    args.warmup = max(args.warmup, 3)
    B, K = args.pairs, args.boxes

    # weights: generated on rank 0, broadcast over NCCL (the path's only start-up collective)
    def weights(make):
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in make().items()} if rank == 0 else {}
        return shard.broadcast_state_dict(sd, src=0)
    sd_flow = weights(lambda: synth.pwc_synthetic_state_dict(0))
    P_gen = {k: v.numpy() for k, v in weights(lambda: synth.propnet_synthetic_params(8)).items()}
    P_spec = {k: v.numpy() for k, v in weights(lambda: synth.propnet_synthetic_params(3)).items()}
    P_ref = {k: v.numpy() for k, v in weights(lambda: synth.refnet_synthetic_params(2)).items()}
    pipe = pipeline.FramePipeline(sd_flow, P_gen, P_spec, P_ref, (H_IN, W_IN), pairs_per_step=B, boxes_per_frame=K, refine_batch=args.refine_batch or None)

    # `sets` different input batches, rotated so that consecutive steps never read the same input (> L2 in total)
    sets, per_set, l2_text = l2_note(B, K)
    units = make_units(sets * B, K)
    host_sets, dev_sets = [], []
    for s in range(sets):
        u = units[s * B:(s + 1) * B]
        hs = [torch.from_numpy(np.stack([x[i] for x in u])).pin_memory() for i in range(3)]
        host_sets.append(hs)
        dev_sets.append([t.cuda() for t in hs])

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not distributed:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # N > 1: the per-step results of every rank are gathered to rank 0 inside the timed region (one NCCL gather of the flat,
    # bit-packed result buffer per step; asynchronous: the gather of step i overlaps the compute of step i + 1)
    flats = [pipe.pack_step_results() for _ in range(2)] if distributed else None
    pending = [None, None]

    def step_device(i):
        pipe.run_frames_device(*dev_sets[i % sets])
        if distributed:
            j = i & 1
            if pending[j] is not None:
                pending[j][1].wait()            # the buffer's previous gather is done before it is overwritten
            pipe.pack_step_results(out=flats[j])
            pending[j] = shard.gather_tensor(flats[j], dst=0, async_op=True)

    def drain():
        for j in range(2):
            if pending[j] is not None:
                pending[j][1].wait()
                pending[j] = None

    # ---- device-resident arm ----
    for i in range(args.warmup):
        step_device(i)
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    drain()
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.kernel_launch_count() - launches0
    value = world * B * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end arm: pinned host buffers in, pinned host results out, every step ----
    for i in range(2):
        pipe.run_frames_host(*host_sets[i % sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipe.run_frames_host(*host_sets[i % sets])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e = world * B * args.steps / e2e_s

    # ---- the product mode: boxes from the unit's own proposal passes (data-dependent work) ----
    det_steps = max(1, min(args.steps, 10))
    pipe.run_frames_host(host_sets[0][0], host_sets[0][1])
    barrier()
    t0 = time.perf_counter()
    nboxes = 0
    for i in range(det_steps):
        r = pipe.run_frames_host(host_sets[i % sets][0], host_sets[i % sets][1])
        nboxes += int(np.sum(r["num_boxes"]))
    torch.cuda.synchronize()
    det_s = max_over_ranks(time.perf_counter() - t0)
    e2e_detected = {"value": world * B * det_steps / det_s, "unit": UNIT, "steps": det_steps,
                    "mean_boxes_refined_per_frame": nboxes / (det_steps * B),
                    "call": "FramePipeline.run_frames_host(boxes=None): each unit refines the boxes its two proposal passes detect "
                            "(random-init heads: the count is data dependent), one device->host read of the counts mid-step"}

    # ---- per-launch profile of one step (rank 0), stages serial on one stream -> roofline of the dominant kernel ----
    roofline, kernels = None, None
    if rank == 0:
        peaks = load_peaks()
        _lib.profile_begin()
        pipe.run_frames_device(*dev_sets[0], concurrent=False)
        prof = _lib.profile_end()
        total_ms = sum(v["ms"] for v in prof.values()) or 1.0
        # every kernel of the step with its share and its fraction of the roofline that bounds it: algorithmic bytes (or FLOPs, x3
        # issued bf16 FLOPs for the split-bf16 convolutions) of its launches / their summed device time / the measured peak
        kernels = {}
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            e = {"launches_per_step": v["launches"], "ms_per_step": v["ms"], "share": v["ms"] / total_ms}
            sec = max(v["ms"], 1e-9) * 1e-3
            if k in ("conv_umma_kernel", "conv_pair_kernel"):
                e.update({"bound": "tensor", "algorithmic_tflops": v["flops"] / sec / 1e12,
                          "frac_of_bf16_peak": v["flops"] / sec / 1e12 / peaks["tensor_sustained"],
                          "tensor_pipe_frac": 3 * v["flops"] / sec / 1e12 / peaks["tensor_sustained"]})
            elif v["bytes"] > 0:
                e.update({"bound": "hbm", "algorithmic_gbs": v["bytes"] / sec / 1e9, "frac_of_hbm_peak": v["bytes"] / sec / 1e9 / peaks["hbm"]})
            kernels[k] = e
        # The tensor-core convolution (launch_conv_umma in csrc/conv_umma.cu -- the launch set round 1 reported as "conv_umma_kernel")
        # runs as two kernels since round 2: conv_umma_kernel<4|8> (one CTA per tile) and conv_pair_kernel (cta_group::2 CTA pairs,
        # stream-K, for the wide 1x1 layers).  The roofline entry covers ALL of its launches; each kernel alone is listed beside it.
        CONV = ["conv_umma_kernel", "conv_pair_kernel"]
        def tensor_entry(names):
            ms = sum(prof[n]["ms"] for n in names if n in prof)
            fl = sum(prof[n]["flops"] for n in names if n in prof)
            ln = sum(prof[n]["launches"] for n in names if n in prof)
            if not ln:
                return None
            ach = fl / (ms * 1e-3) / 1e12
            return {"kernel": "+".join(n for n in names if n in prof), "achieved": ach, "frac": ach / peaks["tensor_sustained"],
                    "tensor_pipe_frac": 3 * ach / peaks["tensor_sustained"], "ms_per_step": ms, "launches_per_step": ln,
                    "avg_launch_us": ms * 1e3 / ln, "algorithmic_gflop_per_step": fl / 1e9, "share_of_step": ms / total_ms}
        name, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        sec = top["ms"] * 1e-3
        traffic = load_traffic(name)
        if name in CONV:
            both = tensor_entry(CONV)
            peak = peaks["tensor_sustained"]
            roofline = {"kernel": both["kernel"], "bound": "tensor", "achieved": both["achieved"], "peak": peak, "unit": "TFLOP/s",
                        "frac": both["frac"], "traffic": traffic, "peak_source": peaks["source"] + " bf16 sustained",
                        "avg_launch_us": both["avg_launch_us"], "launches_per_step": both["launches_per_step"],
                        "algorithmic_gflop_per_step": both["algorithmic_gflop_per_step"], "share_of_step": both["share_of_step"],
                        "serial_step_ms": total_ms,
                        "note": "the tcgen05 convolution (csrc/conv_umma.cu, one launcher, two kernels: per_kernel); achieved = algorithmic "
                                "fp32 FLOPs of all its launches in one step / their summed device time; every algorithmic FLOP is issued "
                                "as 3 bf16 tensor-core FLOPs (split-bf16 x3 for 1e-3 fp32 parity), so tensor-pipe occupancy is 3x frac; "
                                "traffic = mean DRAM bytes per launch (ncu)",
                        "bf16_tflops_issued": 3 * both["achieved"], "tensor_pipe_frac": both["tensor_pipe_frac"],
                        "per_kernel": {n: tensor_entry([n]) for n in CONV},
                        "traffic_per_kernel": {n: load_traffic(n) for n in CONV},
                        "traffic_note": "ncu --set full capture of the refinement network's launch group (profiles/r02_ncu_full_refnet."
                                        "summary.txt): mean DRAM bytes per launch; `traffic` is the value of " + name}
        else:
            ach = top["bytes"] / sec / 1e9
            peak = peaks["hbm"]
            roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": traffic, "peak_source": peaks["source"] + " copy",
                        "avg_launch_us": top["ms"] * 1e3 / top["launches"], "launches_per_step": top["launches"]}

    # ---- each network alone, device resident (BASELINE configs[1..3]) ----
    stages = None
    if rank == 0 and not args.no_stages:
        fr, bx = dev_sets[0][1], dev_sets[0][2]
        ff, pi = pipe.prepare_device(dev_sets[0][0], fr)
        o = pipe.out
        t_flow = time_stage(lambda: pipe.flow_net.forward_u8(ff, out=o["flow"]), 20)
        t_prop = time_stage(lambda: pipe.general.forward_device(pi if B > 1 else pi[0]), 10) / B     # as the pipeline runs it: B frames per forward
        t_prop1 = time_stage(lambda: pipe.general.forward_device(pi[0]), 10)                        # BASELINE configs[2]: one frame per forward
        t_ref = time_stage(lambda: pipe.refine.refine_device(fr[0], bx[0], masks=o["masks"][0], conf=o["conf"][0]), 5)
        stages = {"flow": {"ms_per_pair": t_flow / B, "pairs_per_s": 1e3 * B / t_flow, "batch": B,
                           "algorithmic_tflop_per_s": GFLOP_FLOW * B / t_flow},
                  "proposal_pass": {"ms_per_frame": t_prop, "frames_per_s": 1e3 / t_prop, "input": [pipe.Hp, pipe.Wp], "batch": B,
                                    "ms_per_frame_batch1": t_prop1},
                  "refine": {"ms_per_frame_of_%d_boxes" % K: t_ref, "crops_per_s": 1e3 * K / t_ref,
                             "algorithmic_tflop_per_s": GFLOP_CROP * K / t_ref},
                  "sum_serial_ms_per_pair": t_flow / B + 2 * t_prop + t_ref,
                  "note": "each network alone on one stream, CUDA events; a unit = flow + 2 proposal passes + refine"}
        # BASELINE configs[2] / [3] at their own sizes (context): one 854x480 DAVIS-shaped frame -> 749x1333 after CustomResize, one
        # frame per forward; 100 crops through the refinement network (launch groups of K)
        from premvos_b200 import ops as _ops, propnet as _propnet
        Hd, Wd = _propnet.custom_resize_shape(480, 854)
        davis = _ops.resize_linear_u8(torch.from_numpy(synth.synthetic_bgr_frame(480, 854, seed=21)).cuda(), Hd, Wd)
        t_c3 = time_stage(lambda: pipe.general.forward_device(davis), 10)
        bx100 = torch.from_numpy(synth.synthetic_boxes(100, H_IN, W_IN, seed=22)).cuda()
        t_c4 = time_stage(lambda: pipe.refine.refine_device(fr[0], bx100), 3)
        stages["c3_proposal_854x480"] = {"ms_per_frame": t_c3, "input": [Hd, Wd], "batch": 1}
        stages["c4_refine_100_crops"] = {"ms": t_c4, "crops_per_s": 1e5 / t_c4, "launch_groups_of": K}
        if not args.no_reid:
            # SURVEY 8(f) N2, not part of the headline unit: the ReID embeddings MergeTrack adds to the proposals of a frame
            from premvos_b200 import reid
            rnet = reid.ReIDNet(max_batch=K).load_params(synth.reid_synthetic_params(0))
            t_reid = time_stage(lambda: rnet.embed_device(fr[0], bx[0]), 5)
            stages["reid"] = {"ms_per_frame_of_%d_boxes" % K: t_reid, "crops_per_s": 1e3 * K / t_reid,
                              "note": "ReID network (128 x 128 crops, 17 residual units, 124.9M parameters), not in the headline unit"}
            rnet.close()
            del rnet

    h2d_bytes, d2h_bytes = pipe.h2d_bytes_per_step(original_frames=True), pipe.d2h_bytes_per_step()
    launches_per_step = pipe.launches_per_step_from_frames()

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        units_s, cores, parts = oracle_unit_times(1, 0, K)
        cpu_baseline = {"value": 1.0 / units_s[0], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_text(1, parts, K)}

    # ---- context: the same arithmetic through torch / cuDNN on this GPU, and BASELINE configs[0] (rank 0, N=1 only) ----
    library_baseline, c1 = None, None
    if rank == 0 and world == 1:
        c1 = c1_correlation(load_peaks()["hbm"])
        if not args.no_library_baseline:
            del pipe
            torch.cuda.empty_cache()
            try:
                from baseline import library_baseline as LB
                library_baseline = LB.run(pairs=B, boxes=K)
            except Exception as e:   # context only: never fail the bench line over it
                library_baseline = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3 split-fp32 (fp32 accumulate)", "data": "synthetic",
                "config": bench_config(B, K, world, args.refine_batch, l2_text),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "call": "FramePipeline.run_frames_host: pinned uint8 frames t, t+1 + boxes in (as a decoder hands them over; the stage "
                                "drivers' cv2.resize calls run on the device, bit-exact), flow + detections + per-box masks + conf_scores "
                                "out to pinned host memory, synchronous per step"},
                "e2e_detected": e2e_detected,
                "gather": ({"bytes_per_rank_per_step": int(flats[0].numel()), "collective": "NCCL gather to rank 0, one per step, inside the "
                            "timed region, asynchronous (two buffers in flight)"} if distributed else None),
                "gpu_launches": int(launches), "launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "library_baseline": library_baseline,
                "c1_correlation": c1, "kernels": kernels, "stages": stages}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
