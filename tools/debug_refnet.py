import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refnet_oracle as O
from premvos_b200 import refnet, synth
S, mu = 385, 16
P = synth.refnet_synthetic_params(2, mu)
net = refnet.RefinementNet(max_batch=4, input_size=S, middle_units=mu).load_params(P)
frame = synth.synthetic_bgr_frame(480, 854, seed=3)
boxes = synth.synthetic_boxes(2, 480, 854, seed=3)
masks, conf, post = net.refine(frame, boxes, want_posteriors=True)
lg_gpu = net.get_tensor("logits").reshape(4, 97, 97, 16)[..., :2]
crops = net.get_tensor("crops").reshape(4, 4).astype(int)
for i in range(2):
    crop = tuple(crops[i])
    m_ref, p_ref = O.segmentation_output(torch.from_numpy(lg_gpu[i].copy()), crop, 480, 854, S)
    d = np.abs(post[i] - p_ref)
    y, x = np.unravel_index(d.argmax(), d.shape)
    print("box", i, "crop", crop, "max|lg|", np.abs(lg_gpu[i]).max(), "post diff (GPU logits -> oracle output fn)", d.max(), "at", (y, x),
          "gpu", post[i][y, x], "ref", p_ref[y, x], "mask diff", int((masks[i] != m_ref).sum()))
    cy0, cx0, cy1, cx1 = crop
    ch, cw = cy1 - cy0, cx1 - cx0
    yy, xx = y - cy0, x - cx0
    print("   yy,xx", yy, xx, "src", np.float32(yy) * (np.float32(S) / np.float32(ch)), np.float32(xx) * (np.float32(S) / np.float32(cw)))
    lgS = O.tf_resize_bilinear(torch.from_numpy(lg_gpu[i].copy()).permute(2, 0, 1), S, S)
    pr = torch.softmax(lgS, 0)[1].numpy()
    sy, sx = int(np.floor(np.float32(yy) * (np.float32(S) / np.float32(ch)))), int(np.floor(np.float32(xx) * (np.float32(S) / np.float32(cw))))
    print("   neighbours prob", pr[sy:sy + 2, sx:sx + 2], "logits", lgS[:, sy:sy + 2, sx:sx + 2].numpy().tolist())
