"""Resident per-frame pipeline: stages 1, 2, 3 and 5 of the reference's simple_run.sh for one GPU's share of the frames.

    stage 1  flow                 script_pwc_multi.py:33-70,100            unit = frame pair (t, t+1)
    stage 2  general proposals    proposal_net/train.py --forward          unit = frame (simple_run.sh:29-34)
    stage 3  specific proposals   same graph, second weight set            (simple_run.sh:36-42)
    stage 4  general + specific   combine_general_and_specific.py:23-44    list concatenation (<= 40 boxes / frame)
    stage 5  refinement           refinement_net main.py configs/run       unit = (frame, proposal)

The reference runs these as five processes over the whole dataset with JSON / .flo files in between and one
`session.run` per frame / per proposal.  Here one process per GPU keeps the three networks resident and runs a step of
`pairs_per_step` units with the stages on four CUDA streams (flow | general proposals | specific proposals | refinement):
the streams only join at the end of the step, so the small-grid layers of one network fill the SMs the others leave
idle.  Frames shard across ranks by unit (premvos_b200/shard.py) with no data-path collective.

Everything numeric runs in libpremvos_b200.so through the device entry points (premvos_pwc_forward_u8,
premvos_propnet_forward_u8 + premvos_propnet_copy_results, premvos_refnet_forward); there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops as _ops
from . import propnet as _propnet
from . import pwc as _pwc
from . import refnet as _refnet

RESULTS_PER_IM = _propnet.RESULTS_PER_IM


def flow_input_shape(h, w):
    """script_pwc_multi.py:38-45: network input size = frame size rounded up to multiples of 64."""
    return -(-h // 64) * 64, -(-w // 64) * 64


class FramePipeline:
    """Holds PWC-DC-Net, the proposal network with its two weight sets and the refinement network on the current device.

    A unit is the frame pair (t, t+1): flow t -> t+1, both proposal passes on frame t+1 and the refinement of frame
    t+1's boxes.  `run_device` takes device tensors and never synchronises; `run_host` is the end-to-end call on pinned
    host buffers (uploads, runs, downloads, one synchronisation per step).
    """

    def __init__(self, pwc_state_dict, general_params, specific_params, refine_params, frame_hw, pairs_per_step=4,
                 boxes_per_frame=2 * RESULTS_PER_IM, refine_batch=None, num_blocks=(3, 4, 23, 3), middle_units=16,
                 refine_input_size=385, tensor_cores=True):
        self.H, self.W = int(frame_hw[0]), int(frame_hw[1])
        self.B, self.K = int(pairs_per_step), int(boxes_per_frame)
        if refine_batch is None:     # all boxes of a frame in one launch group (measured: 40 per group beats 2 x 20 by 6 %)
            refine_batch = max(1, self.K)
        self.Hn, self.Wn = flow_input_shape(self.H, self.W)
        self.Hp, self.Wp = _propnet.custom_resize_shape(self.H, self.W)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.flow_net = _pwc.pwc_dc_net(None, tensor_cores=tensor_cores)
        self.flow_net.load_state_dict(pwc_state_dict)
        self.flow_net.cuda(self.dev.index).eval()
        self.general = _propnet.ProposalNet(num_blocks).load_params(general_params)
        self.specific = _propnet.ProposalNet(num_blocks).load_params(specific_params)
        self.refine = _refnet.RefinementNet(max_batch=refine_batch, input_size=refine_input_size,
                                            middle_units=middle_units).load_params(refine_params)
        # build every handle now (weight packing, buffer allocation, graph capture)
        self.flow_net._handle(self.B, self.Hn, self.Wn)
        # the B frames of a step go through each proposal network as ONE batched forward (every conv launch covers all of them)
        self.general._handle(self.Hp, self.Wp, self.B)
        self.specific._handle(self.Hp, self.Wp, self.B)
        self.refine._h()
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(4)]
        B, K, d = self.B, self.K, self.dev
        self.out = {
            "flow": torch.empty((B, 2, self.Hn // 4, self.Wn // 4), dtype=torch.float32, device=d),
            "det_count": torch.zeros((2, B), dtype=torch.int32, device=d),
            "det_boxes": torch.zeros((2, B, RESULTS_PER_IM, 4), dtype=torch.float32, device=d),
            "det_probs": torch.zeros((2, B, RESULTS_PER_IM), dtype=torch.float32, device=d),
            "masks": torch.zeros((B, K, self.H, self.W), dtype=torch.uint8, device=d),
            "conf": torch.zeros((B, K), dtype=torch.float32, device=d),
        }
        # what leaves the device (host buffers, rank 0): the masks bit-packed, 8 pixels per byte (ops.pack_mask_bits)
        self.packed = torch.zeros((B, K, _ops.packed_mask_bytes(self.H * self.W)), dtype=torch.uint8, device=d)
        self._stage = None   # device + pinned staging of run_host, allocated on first use
        self._resized = None
        self._frames_prev_dev = None

    # ---- sizes ---------------------------------------------------------------------------------------------
    def input_shapes(self):
        B = self.B
        return {"flow_frames": (B, 2, self.Hn, self.Wn, 3), "prop_images": (B, self.Hp, self.Wp, 3),
                "frames": (B, self.H, self.W, 3), "boxes": (B, self.K, 4)}

    def launches_per_step(self, num_boxes=None):
        k = self.K if num_boxes is None else num_boxes
        groups = -(-k // self.refine.max_batch)
        from . import _lib
        per_group = int(_lib.lib().premvos_refnet_launches_per_forward(self.refine._h()))
        return (self.flow_net.launches_per_forward(self.B, self.Hn, self.Wn) + 1
                + 2 * (self.general.launches_per_forward(self.Hp, self.Wp, self.B) + 1) + self.B * groups * per_group)

    def launches_per_step_from_frames(self, num_boxes=None):
        resize = 1 + (0 if (self.Hn, self.Wn) == (self.H, self.W) else 2 * self.B)
        return self.launches_per_step(num_boxes) + resize

    # ---- device-resident step ------------------------------------------------------------------------------
    def run_device(self, flow_frames, prop_images, frames, boxes=None, concurrent=True):
        """flow_frames uint8 RGB [B,2,Hn,Wn,3], prop_images uint8 BGR [B,Hp,Wp,3], frames uint8 RGB [B,H,W,3], all CUDA.
        boxes: CUDA float32 [B,K,4] xywh in frame coordinates -> every unit refines exactly K given boxes (fixed work:
        benchmarking, or boxes that come from a tracker); None -> each unit refines the boxes its two proposal passes
        detect, as stage 4 combines them (the first K of them; K = 40 holds every possible detection; costs one small
        device->host read per step).
        concurrent=False runs the stages one after the other on the current stream (per-kernel profiling).
        Returns the dict of device outputs (overwritten by the next step); with boxes=None also 'num_boxes' [B]."""
        B = self.B
        if tuple(flow_frames.shape) != (B, 2, self.Hn, self.Wn, 3) or tuple(prop_images.shape) != (B, self.Hp, self.Wp, 3) \
                or tuple(frames.shape) != (B, self.H, self.W, 3):
            raise ValueError("inputs must have the shapes of input_shapes(): %s" % (self.input_shapes(),))
        cur = torch.cuda.current_stream(self.dev)
        streams = self.streams if concurrent else [cur] * 4
        s_flow, s_gen, s_spec, s_ref = streams
        if concurrent:
            for s in streams:
                s.wait_stream(cur)
        o = self.out
        with torch.cuda.stream(s_flow):
            self.flow_net.forward_u8(flow_frames, out=o["flow"])
        for which, (net, s) in enumerate(((self.general, s_gen), (self.specific, s_spec))):
            with torch.cuda.stream(s):
                net.forward_device(prop_images if B > 1 else prop_images[0])
                net.copy_results_device(self.Hp, self.Wp, o["det_count"][which], o["det_boxes"][which], o["det_probs"][which], batch=B)
        result = dict(o)
        if boxes is not None:
            if tuple(boxes.shape) != (B, self.K, 4):
                raise ValueError("boxes must be [B,K,4] = %s" % ((B, self.K, 4),))
            with torch.cuda.stream(s_ref):
                for b in range(B):
                    self.refine.refine_device(frames[b], boxes[b], masks=o["masks"][b], conf=o["conf"][b])
        else:
            if concurrent:
                s_ref.wait_stream(s_gen)
                s_ref.wait_stream(s_spec)
            with torch.cuda.stream(s_ref):
                counts = o["det_count"].cpu().numpy()           # synchronises s_ref (hence both proposal streams)
                det = o["det_boxes"].cpu().numpy()
                nb = np.zeros((B,), np.int64)
                for b in range(B):
                    bx = self.combined_boxes_xywh(det[0, b, :counts[0, b]], det[1, b, :counts[1, b]])[:self.K]
                    nb[b] = len(bx)
                    if nb[b]:
                        bd = torch.from_numpy(bx).to(self.dev)
                        self.refine.refine_device(frames[b], bd, masks=o["masks"][b, :nb[b]], conf=o["conf"][b, :nb[b]])
                result["num_boxes"] = nb
        if concurrent:
            for s in streams:
                cur.wait_stream(s)
        return result

    # ---- the same step from the ORIGINAL frames: the stage drivers' cv2.resize calls run on the device -------------
    def prepare_device(self, frames_prev, frames_cur):
        """frames_prev, frames_cur: CUDA uint8 RGB [B,H,W,3] (frames t and t+1 of every unit).  Builds the two resized
        inputs on the device with the bit-exact cv2.resize kernel (premvos_resize_linear_u8): both frames at multiples of 64
        for the flow network (script_pwc_multi.py:38-45) and the CustomResize'd BGR frame t+1 for the proposal passes
        (eval.py:75-78).  -> (flow_frames [B,2,Hn,Wn,3], prop_images [B,Hp,Wp,3]), overwritten by the next call."""
        B = self.B
        if tuple(frames_prev.shape) != (B, self.H, self.W, 3) or tuple(frames_cur.shape) != (B, self.H, self.W, 3):
            raise ValueError("frames must be [B,H,W,3] = %s" % ((B, self.H, self.W, 3),))
        if self._resized is None:
            self._resized = (torch.empty((B, 2, self.Hn, self.Wn, 3), dtype=torch.uint8, device=self.dev),
                             torch.empty((B, self.Hp, self.Wp, 3), dtype=torch.uint8, device=self.dev))
        ff, pi = self._resized
        for b in range(B):
            if (self.Hn, self.Wn) == (self.H, self.W):
                ff[b, 0].copy_(frames_prev[b]); ff[b, 1].copy_(frames_cur[b])
            else:
                _ops.resize_linear_u8(frames_prev[b], self.Hn, self.Wn, out=ff[b, 0])
                _ops.resize_linear_u8(frames_cur[b], self.Hn, self.Wn, out=ff[b, 1])
        _ops.resize_linear_u8(frames_cur, self.Hp, self.Wp, reverse_channels=True, out=pi)
        return ff, pi

    def run_frames_device(self, frames_prev, frames_cur, boxes=None, concurrent=True):
        """run_device on the original frames (CUDA uint8 RGB [B,H,W,3] each): resizes on the device, then the step."""
        ff, pi = self.prepare_device(frames_prev, frames_cur)
        return self.run_device(ff, pi, frames_cur, boxes, concurrent=concurrent)

    def run_frames_host(self, frames_prev, frames_cur, boxes=None):
        """End-to-end step from pinned HOST frames (uint8 RGB [B,H,W,3] each, what a decoder hands over): one upload per
        frame, everything else on the device; results in pinned host buffers, one synchronisation."""
        dev, host = self._staging()
        if self._frames_prev_dev is None:
            self._frames_prev_dev = torch.empty((self.B, self.H, self.W, 3), dtype=torch.uint8, device=self.dev)
        cur = torch.cuda.current_stream(self.dev)
        self._frames_prev_dev.copy_(frames_prev, non_blocking=True)
        dev["frames"].copy_(frames_cur, non_blocking=True)
        if boxes is not None:
            dev["boxes"].copy_(boxes, non_blocking=True)
        res = self.run_frames_device(self._frames_prev_dev, dev["frames"], dev["boxes"] if boxes is not None else None)
        return self._download(res, host, cur)

    def _download(self, res, host, cur):
        """Device results -> pinned host buffers, one synchronisation.  The masks travel bit-packed ('masks_packed',
        ops.unpack_mask_bits restores uint8 [B,K,H,W]); `LazyMasks` unpacks them on first access of 'masks'."""
        _ops.pack_mask_bits(self.out["masks"], out=self.packed)
        for k, v in self.out.items():
            if k != "masks":
                host[k].copy_(v, non_blocking=True)
        host["masks_packed"].copy_(self.packed, non_blocking=True)
        cur.synchronize()
        out = HostResults(host, self.H, self.W)
        if "num_boxes" in res:
            out["num_boxes"] = res["num_boxes"]
        return out

    def combined_boxes_xywh(self, general_x1y1x2y2, specific_x1y1x2y2):
        return combine_proposals(general_x1y1x2y2, specific_x1y1x2y2, (self.H, self.W), (self.Hp, self.Wp))

    # ---- end-to-end step on host buffers -------------------------------------------------------------------
    def _staging(self):
        if self._stage is None:
            shp = self.input_shapes()
            dev = {k: torch.empty(v, dtype=torch.float32 if k == "boxes" else torch.uint8, device=self.dev) for k, v in shp.items()}
            host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in self.out.items() if k != "masks"}
            host["masks_packed"] = torch.empty(self.packed.shape, dtype=torch.uint8).pin_memory()
            self._stage = (dev, host)
        return self._stage

    def run_host(self, flow_frames, prop_images, frames, boxes=None):
        """Same step on HOST tensors (pinned for asynchronous copies): uploads the inputs, runs run_device, downloads every
        output into pinned host buffers and synchronises once.  Returns a dict of host tensors (reused by the next call)."""
        dev, host = self._staging()
        cur = torch.cuda.current_stream(self.dev)
        dev["flow_frames"].copy_(flow_frames, non_blocking=True)
        dev["prop_images"].copy_(prop_images, non_blocking=True)
        dev["frames"].copy_(frames, non_blocking=True)
        if boxes is not None:
            dev["boxes"].copy_(boxes, non_blocking=True)
        res = self.run_device(dev["flow_frames"], dev["prop_images"], dev["frames"], dev["boxes"] if boxes is not None else None)
        return self._download(res, host, cur)

    def h2d_bytes_per_step(self, with_boxes=True, original_frames=False):
        shp = self.input_shapes()
        if original_frames:   # run_frames_host: frames t and t+1 of every unit
            n = 2 * int(np.prod(shp["frames"]))
        else:                 # run_host: the two resized inputs + frame t+1
            n = sum(int(np.prod(shp[k])) for k in ("flow_frames", "prop_images", "frames"))
        return n + (int(np.prod(shp["boxes"])) * 4 if with_boxes else 0)

    def d2h_bytes_per_step(self):
        """Bytes run_host / run_frames_host bring back per step: flow, detections, conf_scores and the bit-packed masks."""
        return sum(v.numel() * v.element_size() for k, v in self.out.items() if k != "masks") + self.packed.numel()

    # ---- the results of a step as ONE flat device buffer (what a rank sends to rank 0) --------------------------------------
    def _flat_layout(self):
        keys = [k for k in self.out if k != "masks"]
        sizes = [self.out[k].numel() * self.out[k].element_size() for k in keys] + [self.packed.numel()]
        offs = np.concatenate([[0], np.cumsum([-(-n // 16) * 16 for n in sizes])])
        return keys + ["masks_packed"], sizes, offs

    def pack_step_results(self, out: torch.Tensor = None) -> torch.Tensor:
        """flow | det_count | det_boxes | det_probs | conf | bit-packed masks of the last step -> one contiguous uint8 CUDA
        tensor (16-byte aligned pieces; ~2.5 MB per unit instead of 18 MB with byte masks).  Enqueues on the current stream."""
        keys, sizes, offs = self._flat_layout()
        if out is None:
            out = torch.empty((int(offs[-1]),), dtype=torch.uint8, device=self.dev)
        _ops.pack_mask_bits(self.out["masks"], out=self.packed)
        for k, n, o in zip(keys, sizes, offs[:-1]):
            src = self.packed if k == "masks_packed" else self.out[k]
            out[int(o):int(o) + n].copy_(src.reshape(-1).view(torch.uint8), non_blocking=True)
        return out

    def unpack_step_results(self, flat) -> dict:
        """Host inverse of pack_step_results: uint8 [bytes] (CPU tensor / numpy) -> {'flow', 'det_*', 'conf', 'masks'} numpy."""
        a = flat.cpu().numpy() if isinstance(flat, torch.Tensor) else np.asarray(flat)
        keys, sizes, offs = self._flat_layout()
        res = {}
        for k, n, o in zip(keys, sizes, offs[:-1]):
            piece = a[int(o):int(o) + n]
            if k == "masks_packed":
                res["masks"] = _ops.unpack_mask_bits(piece.reshape(self.B, self.K, -1), self.H, self.W)
            else:
                t = self.out[k]
                res[k] = piece.view(np.dtype(str(t.dtype).replace("torch.", ""))).reshape(tuple(t.shape))
        return res

    def result_bytes_per_unit(self):
        """Bytes of one unit's results as they are gathered to rank 0 (pack_unit_results)."""
        return self.d2h_bytes_per_step() // self.B


class HostResults(dict):
    """The host result dict of a step.  'masks' (uint8 0/1 [B,K,H,W]) is unpacked from 'masks_packed' on first access."""

    def __init__(self, host, H, W):
        super().__init__(host)
        self._hw = (H, W)

    def __missing__(self, key):
        if key == "masks":
            m = torch.from_numpy(_ops.unpack_mask_bits(self["masks_packed"], *self._hw))
            self[key] = m
            return m
        raise KeyError(key)

    def items_unpacked(self):
        self["masks"]
        return [(k, v) for k, v in self.items() if k != "masks_packed"]


def combine_proposals(general_x1y1x2y2, specific_x1y1x2y2, frame_hw, resized_hw):
    """Stage-2/3 post-processing + stage 4 on the boxes of one frame: boxes / scale, clip to the frame (eval.py:93-96),
    x1y1x2y2 -> xywh rounded to one decimal as the proposal JSON stores them (train.py:399-408), general then specific
    (combine_general_and_specific.py:37).  -> float32 [n,4] as the refinement stage reads them (proposal['bbox'])."""
    H, W = frame_hw
    scale = (resized_hw[0] * 1.0 / H + resized_hw[1] * 1.0 / W) / 2
    out = []
    for boxes in (general_x1y1x2y2, specific_x1y1x2y2):
        b = np.array(boxes, dtype=np.float32).reshape(-1, 4) / scale
        b = _propnet.clip_boxes(b, (H, W))
        for box in b:
            box = np.array(box, dtype=np.float32)     # float32 subtraction and np.float32 rounding, as train.py:396-401 does
            box[2] -= box[0]
            box[3] -= box[1]
            out.append([float(round(x, 1)) for x in box])
    return np.asarray(out, dtype=np.float32).reshape(-1, 4)


def prepare_unit(frame_t, frame_t1):
    """Host-side decode-time preparation of one unit from two RGB uint8 frames [H,W,3], exactly what the reference's
    stage drivers do with cv2 before their networks see a tensor: script_pwc_multi.py:38-45 (resize both frames to
    multiples of 64) and proposal_net eval.py:75-78 + common.py:49-62 (CustomResize of the BGR frame)."""
    import cv2
    H, W = frame_t1.shape[:2]
    Hn, Wn = flow_input_shape(H, W)
    Hp, Wp = _propnet.custom_resize_shape(H, W)
    pair = np.stack([cv2.resize(frame_t, (Wn, Hn)), cv2.resize(frame_t1, (Wn, Hn))])
    bgr = np.ascontiguousarray(frame_t1[:, :, ::-1])
    prop = cv2.resize(bgr, (Wp, Hp), interpolation=cv2.INTER_LINEAR)
    return pair, prop, np.ascontiguousarray(frame_t1)


def run_video(pipe: FramePipeline, frames, rank=0, world=1, boxes_of_frame=None):
    """BASELINE config C5: a video (list of RGB uint8 frames) sharded by unit across `world` ranks; rank r runs units
    r, r+world, ... `pairs_per_step` at a time (the last step of a shard is padded by repeating its last unit) and returns
    {unit index t: {'flow', 'det_count', 'det_boxes', 'det_probs', 'masks', 'conf'[, 'num_boxes']}} for its own units;
    merge across ranks with shard.gather_results."""
    from . import shard
    units = shard.shard_units(len(frames) - 1, rank, world)
    results = {}
    B = pipe.B
    for i in range(0, len(units), B):
        chunk = units[i:i + B]
        padded = chunk + [chunk[-1]] * (B - len(chunk))
        prep = [prepare_unit(frames[t], frames[t + 1]) for t in padded]
        ff = torch.from_numpy(np.stack([p[0] for p in prep])).pin_memory()
        pi = torch.from_numpy(np.stack([p[1] for p in prep])).pin_memory()
        fr = torch.from_numpy(np.stack([p[2] for p in prep])).pin_memory()
        bx = None
        if boxes_of_frame is not None:
            bx = torch.from_numpy(np.stack([np.asarray(boxes_of_frame(t + 1), np.float32).reshape(pipe.K, 4) for t in padded])).pin_memory()
        out = pipe.run_host(ff, pi, fr, bx)
        for j, t in enumerate(chunk):
            results[t] = {k: (v[:, j] if k.startswith("det_") else v[j]).numpy().copy() for k, v in out.items_unpacked() if k != "num_boxes"}
            if "num_boxes" in out:
                results[t]["num_boxes"] = int(out["num_boxes"][j])
    return results
