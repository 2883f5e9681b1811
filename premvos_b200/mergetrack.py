"""Host-side mirror of MergeTrack's live mask propagation, backed by the CUDA library.

Reference surface reproduced here (paths under /root/reference/code/MergeTrack/):
  get_flow(filename) -> float32 [h, w, 2]                          merge_functions.py:197-207
  warp_flow(img, flow, binarize=True) -> uint8                     merge_functions.py:209-217
  warp_proposals(proposals, optflow_fn) -> warped proposals        merge_functions.py:219-243
  the per-frame step `warp_proposals -> do_refinement`             merge.py:95-100

`warp_flow` / `warp_proposals` keep the reference's signatures (host arrays in, host arrays / dicts out; all masks of a frame
in one launch instead of one cv2.remap per mask).  `LivePropagator` is the resident form of the same step: the flow field
leaves the flow network, is brought to frame resolution (script_pwc_multi.py:59-68), warps the masks of frame t, and the
boxes of the warped masks feed the refinement network on frame t+1 -- no .flo file, no RLE, no host round trip in between
(the reference: .flo written by stage 1, read back by get_flow; masks RLE-encoded and decoded around every call).
The tracking logic itself (scores, template updates, ReID) is out of scope.  No CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .pwc import readFlowFile
from .refnet import rle_encode


def get_flow(filename):
    """merge_functions.py:197-207, error behaviour included: a file with a wrong magic number prints the reference's message and
    yields None (readFlowFile raises instead)."""
    try:
        return readFlowFile(filename)
    except ValueError:
        print("Magic number incorrect. Invalid .flo file")
        return None


def _cuda_u8(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA uint8 tensor (premvos_b200 has no CPU path)" % name)
    return t


def _cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA float32 tensor (premvos_b200 has no CPU path)" % name)
    return t


def warp_masks_device(masks, flow, binarize=True, out=None, bbox=None, want_bbox=True):
    """masks CUDA uint8 [n,H,W], flow CUDA float32 [H,W,2] -> (warped CUDA uint8 [n,H,W], bbox CUDA float32 [n,4] xywh or
    None).  premvos_warp_masks_u8; enqueues on the current torch stream, never synchronises."""
    _cuda_u8(masks, "masks")
    _cuda_f32(flow, "flow")
    if masks.dim() != 3:
        raise ValueError("masks must be [n,H,W], got %s" % (tuple(masks.shape),))
    n, H, W = (int(v) for v in masks.shape)
    if tuple(flow.shape) != (H, W, 2):
        raise ValueError("flow must be [H,W,2] = %s, got %s" % ((H, W, 2), tuple(flow.shape)))
    if out is None:
        out = torch.empty_like(masks)
    elif _cuda_u8(out, "out").shape != masks.shape:
        raise ValueError("out must have the shape of masks")
    if want_bbox and bbox is None:
        bbox = torch.empty((n, 4), dtype=torch.float32, device=masks.device)
    elif bbox is not None and tuple(_cuda_f32(bbox, "bbox").shape) != (n, 4):
        raise ValueError("bbox must be [n,4]")
    with torch.cuda.device(masks.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_warp_masks_u8(masks.data_ptr(), n, H, W, flow.data_ptr(), out.data_ptr(),
                                                    bbox.data_ptr() if bbox is not None else None, 1 if binarize else 0, st))
    return out, bbox


def flow_postprocess_device(flow2, H, W, out=None):
    """flow2 CUDA float32 [B,2,H_/4,W_/4] (the flow network's output) -> CUDA float32 [B,H,W,2], the flow field at frame
    resolution in frame pixels (script_pwc_multi.py:59-68; premvos_flow_postprocess)."""
    _cuda_f32(flow2, "flow2")
    if flow2.dim() != 4 or flow2.shape[1] != 2:
        raise ValueError("flow2 must be [B,2,h,w], got %s" % (tuple(flow2.shape),))
    B, _, h, w = (int(v) for v in flow2.shape)
    if out is None:
        out = torch.empty((B, H, W, 2), dtype=torch.float32, device=flow2.device)
    elif tuple(_cuda_f32(out, "out").shape) != (B, H, W, 2):
        raise ValueError("out must be [B,H,W,2]")
    with torch.cuda.device(flow2.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_flow_postprocess(flow2.data_ptr(), B, 4 * h, 4 * w, out.data_ptr(), int(H), int(W), st))
    return out


def warp_flow(img, flow, binarize=True):
    """merge_functions.py:209-217 on host arrays: img uint8 [h,w] (or [n,h,w]), flow float32 [h,w,2] -> uint8 like img.
    Unlike the reference this does not modify `flow` in place."""
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim not in (2, 3):
        raise ValueError("expected a uint8 mask [h,w] or [n,h,w], got %s %s" % (img.dtype, img.shape))
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    m = torch.from_numpy(img.reshape((-1,) + img.shape[-2:])).cuda()
    out, _ = warp_masks_device(m, torch.from_numpy(flow).cuda(), binarize=binarize, want_bbox=False)
    return out.cpu().numpy().reshape(img.shape)


def warp_proposals(proposals, optflow_fn):
    """merge_functions.py:219-243.  `optflow_fn`: .flo file name or the flow array.  One launch for all proposals."""
    flow = optflow_fn if isinstance(optflow_fn, np.ndarray) else get_flow(optflow_fn)
    if not proposals:
        return []
    masks = np.ascontiguousarray(np.stack([np.asarray(prop["mask"], dtype=np.uint8) for prop in proposals]))
    out, bbox = warp_masks_device(torch.from_numpy(masks).cuda(), torch.from_numpy(np.ascontiguousarray(flow, dtype=np.float32)).cuda())
    warped_masks, boxes = out.cpu().numpy(), bbox.cpu().numpy().astype(np.float64)
    warped_props = []
    for prop, f_mask, box in zip(proposals, warped_masks, boxes):
        warped_props.append({"segmentation": rle_encode(f_mask), "bbox": box, "score": 0.5 * (prop["final_score"] + 1),
                             "final_score": prop["final_score"], "object_score": prop["object_score"], "mask": f_mask,
                             "id": prop["id"]})
    return warped_props


class LivePropagator:
    """The resident form of merge.py:95-101 for one video: masks of frame t + the pair (t, t+1) -> flow -> frame-resolution
    flow -> warped masks + their boxes -> refinement on frame t+1 (+ the ReID embeddings of the same boxes, merge.py:101, when a
    ReID network is given).  Everything stays on the device and on one stream.

    flow_net: premvos_b200.pwc.PWCDCNet (cuda, eval), refine_net: premvos_b200.refnet.RefinementNet (params loaded),
    reid_net: premvos_b200.reid.ReIDNet (params loaded) or None."""

    def __init__(self, flow_net, refine_net, frame_hw, max_objects=None, reid_net=None):
        from .pipeline import flow_input_shape
        from . import ops
        self._ops = ops
        self.flow_net, self.refine_net, self.reid_net = flow_net, refine_net, reid_net
        self.H, self.W = int(frame_hw[0]), int(frame_hw[1])
        self.Hn, self.Wn = flow_input_shape(self.H, self.W)
        self.max_objects = int(max_objects or refine_net.max_batch)
        d = torch.device("cuda", torch.cuda.current_device())
        self._pair = torch.empty((1, 2, self.Hn, self.Wn, 3), dtype=torch.uint8, device=d)
        self._flow2 = torch.empty((1, 2, self.Hn // 4, self.Wn // 4), dtype=torch.float32, device=d)
        self._flow = torch.empty((1, self.H, self.W, 2), dtype=torch.float32, device=d)

    def flow(self, frame_t, frame_t1):
        """CUDA uint8 RGB frames [H,W,3] -> flow field CUDA float32 [H,W,2] (what stage 1 writes as <frame_t>.flo)."""
        for k, f in enumerate((frame_t, frame_t1)):
            if (self.Hn, self.Wn) == (self.H, self.W):
                self._pair[0, k].copy_(f)
            else:
                self._ops.resize_linear_u8(f, self.Hn, self.Wn, out=self._pair[0, k])
        self.flow_net.forward_u8(self._pair, out=self._flow2)
        return flow_postprocess_device(self._flow2, self.H, self.W, out=self._flow)[0]

    def step(self, masks_t, frame_t, frame_t1):
        """masks_t CUDA uint8 [n,H,W] (0/1, the selected proposals of frame t) -> dict of CUDA tensors:
        warped [n,H,W], bbox [n,4] (xywh of the warped masks), masks [n,H,W] (refined on frame t+1), conf [n], flow [H,W,2],
        and with a ReID network 'reid' [n,128] (add_ReID embeds prop['bbox'], which do_refinement leaves untouched).
        An object whose warped mask is empty has bbox 0,0,0,0 (the reference refines / embeds that box too)."""
        flow = self.flow(frame_t, frame_t1)
        warped, bbox = warp_masks_device(masks_t, flow)
        masks, conf = self.refine_net.refine_device(frame_t1, bbox)
        out = {"flow": flow, "warped": warped, "bbox": bbox, "masks": masks, "conf": conf}
        if self.reid_net is not None:
            out["reid"] = self.reid_net.embed_device(frame_t1, bbox)
        return out


# ---- on-disk formats of stage 7 (host side) -------------------------------------------------------------
def pascal_colormap():
    """The 256-entry PASCAL VOC palette as uint8 [256,3] -- what `(np.array(pascal_colormap) * 255).round()` gives for the
    table at merge_functions.py:250-506 (bit-interleaved class index: bit 3j+c of the index is bit 7-j of channel c)."""
    cm = np.zeros((256, 3), np.uint8)
    for i in range(256):
        c = i
        for j in range(8):
            for ch in range(3):
                cm[i, ch] |= ((c >> ch) & 1) << (7 - j)
            c >>= 3
    return cm


def save_with_pascal_colormap(filename, arr):
    """merge_functions.py:508-514: an indexed PNG whose pixel values are the object ids and whose palette is the VOC colormap
    (PIL's `quantize(palette=...)` of a single-band image copies the data as is and attaches the palette)."""
    from PIL import Image
    im = Image.fromarray(np.squeeze(np.asarray(arr).astype("uint8")))     # single band "L"; putpalette turns it into "P"
    im.putpalette(pascal_colormap().reshape(-1).tolist())
    im.save(filename)


def save_pngs(proposals, output_fn, empty=False):
    """merge_functions.py:516-525: one indexed PNG per frame, pixel = id of the proposal that covers it (later ones win)."""
    import os
    png = np.zeros_like(np.asarray(proposals[0]["mask"]))
    if not empty:
        for prop in proposals:
            png[np.asarray(prop["mask"]).astype("bool")] = prop["id"]
    output_fol = os.path.dirname(output_fn)
    if output_fol and not os.path.exists(output_fol):
        os.makedirs(output_fol)
    save_with_pascal_colormap(output_fn, png)


def to_bbox_host(mask):
    """pycocotools toBbox of a binary mask on the host: [x, y, w, h] float64, zeros when empty."""
    m = np.asarray(mask) != 0
    if not m.any():
        return np.zeros(4, np.float64)
    ys, xs = np.flatnonzero(m.any(axis=1)), np.flatnonzero(m.any(axis=0))
    return np.array([xs[0], ys[0], xs[-1] - xs[0] + 1, ys[-1] - ys[0] + 1], np.float64)


def read_ann(ann_fn):
    """merge_functions.py:14-25: first-frame annotation PNG (or an id array) -> one template proposal per object id."""
    if isinstance(ann_fn, np.ndarray):
        ann = ann_fn
    else:
        from PIL import Image
        ann = np.array(Image.open(ann_fn))
    new_proposals = []
    for id_ in [i for i in np.unique(ann) if i != 0]:
        ann_mask = (ann == id_).astype(np.uint8)
        new_proposals.append({"id": id_, "bbox": to_bbox_host(ann_mask), "segmentation": rle_encode(ann_mask), "conf_score": "1.0",
                              "score": 1.0})
    return new_proposals


def read_props(prop_fn):
    """merge_functions.py:27-36: proposals JSON of a frame; a missing / unreadable file is an empty list; proposals without a
    ReID vector get an all-inf one."""
    import json
    try:
        with open(prop_fn, "r") as f:
            proposals = json.load(f)
        for prop in proposals:
            if "ReID" not in prop.keys():
                prop["ReID"] = np.inf * np.ones((128))
    except Exception:
        proposals = []
    return proposals
