"""Weight-file readers for the three networks (the part that needs no TensorFlow).

The reference restores its weights with three mechanisms:
  * PWC-Net:        torch.load of `pwc_net.pth.tar` (models/PWCNet.py:496-505)                    -> pwc.pwc_dc_net(path)
  * proposal net:   tensorpack `get_model_loader(path)` (proposal_net/train.py:653-657): a `.npz` / `.npy` dict of
                    variables (DictRestore) or a TensorFlow checkpoint (SaverRestore)
  * refinement net: `tf.train.Saver.restore` of a TensorFlow checkpoint (refinement_net/core/Engine.py)
This module reads `.npz` / `.npy` variable dictionaries (tensorpack's own exchange format,
`tensorpack/scripts/dump-model-params.py`) and, through premvos_b200/tf_checkpoint.py, TensorFlow checkpoints themselves
(`<prefix>.index` + `<prefix>.data-*`; that reader restates the bundle format without TensorFlow and is unpinned -- when in
doubt convert once where TensorFlow is installed: `np.savez(out, **{v.name: sess.run(v) for v in tf.global_variables()})`).

What a dictionary from either tool looks like and what is done with it:
  * names may carry the `:0` tensor suffix and a tower / scope prefix (`tower0/`, `tower-pred-0/`)       -> stripped
  * optimizer slots and bookkeeping variables (`global_step`, `learning_rate`, `*/Momentum`, `*/Adam*`,
    `EMA/*`, `*/ExponentialMovingAverage`, `beta1_power` ...)                                            -> dropped
  * every variable the network needs must be present with the expected shape                            -> checked, loudly
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

from .synth import propnet_param_shapes, refnet_param_shapes, reid_param_shapes

_DROP_SUFFIXES = ("/Momentum", "/Adam", "/Adam_1", "/ExponentialMovingAverage", "/AccumGrad", "/RMSProp", "/RMSProp_1")
_DROP_NAMES = {"global_step", "learning_rate", "beta1_power", "beta2_power"}
_PREFIXES = ("tower-pred-0/", "tower0/", "tower_0/", "InferenceTower/")


def _read_dict(path):
    path = str(path)
    if os.path.exists(path + ".index"):   # a TensorFlow checkpoint prefix
        from . import tf_checkpoint
        return tf_checkpoint.read_checkpoint(path)
    if path.endswith(".npy"):
        d = np.load(path, allow_pickle=True, encoding="latin1").item()
    else:
        with np.load(path) as z:
            d = {k: z[k] for k in z.files}
    if not isinstance(d, dict):
        raise ValueError("%s does not hold a {variable name: array} dictionary" % path)
    return d


def normalise_variable_names(d):
    """-> OrderedDict with `:0` / tower prefixes stripped and optimizer / bookkeeping variables removed."""
    out = OrderedDict()
    for k, v in d.items():
        name = k[:-2] if k.endswith(":0") else k
        for p in _PREFIXES:
            if name.startswith(p):
                name = name[len(p):]
        if name in _DROP_NAMES or name.startswith("EMA/") or name.endswith(_DROP_SUFFIXES) or "/Adam" in name:
            continue
        if name in out:
            raise ValueError("variable %r appears twice after name normalisation" % name)
        out[name] = np.asarray(v)
    return out


def _select(d, shapes, what):
    d = normalise_variable_names(d)
    # an unknown enclosing scope (e.g. the network's layer name in refinement_net's graph): accept `<scope>/<name>` when it is
    # the only variable that ends with the wanted name
    tails = None
    for k in shapes:
        if k in d:
            continue
        if tails is None:
            tails = {}
            for full in d:
                parts = full.split("/")
                for i in range(1, len(parts)):
                    tails.setdefault("/".join(parts[i:]), []).append(full)
        cands = tails.get(k, [])
        if len(cands) == 1:
            d[k] = d[cands[0]]
    missing = [k for k in shapes if k not in d]
    if missing:
        raise KeyError("%s: %d variable(s) missing, e.g. %s" % (what, len(missing), missing[:5]))
    out = OrderedDict()
    for k, shp in shapes.items():
        v = np.ascontiguousarray(d[k], dtype=np.float32)
        if tuple(v.shape) != tuple(shp):
            raise ValueError("%s: %s has shape %s, expected %s" % (what, k, tuple(v.shape), tuple(shp)))
        out[k] = v
    return out


def select_proposal_net_variables(d, num_blocks=(3, 4, 23, 3), num_class=2, second_num_class=81, mode_mask=False):
    """{name: array} as read from any source -> the dict ProposalNet.load_params takes."""
    return _select(d, propnet_param_shapes(num_blocks, num_class, second_num_class, mode_mask), "proposal_net")


def select_refinement_net_variables(d, middle_units=16, n_classes=2):
    return _select(d, refnet_param_shapes(middle_units, n_classes), "refinement_net")


def load_proposal_net_variables(path, num_blocks=(3, 4, 23, 3), num_class=2, second_num_class=81, mode_mask=False):
    """`.npz` / `.npy` dictionary of tensorpack variables, or a TensorFlow checkpoint prefix -> the dict
    ProposalNet.load_params takes (float32, exact shapes; unrelated variables ignored)."""
    return _select(_read_dict(path), propnet_param_shapes(num_blocks, num_class, second_num_class, mode_mask), "proposal_net")


def load_refinement_net_variables(path, middle_units=16, n_classes=2):
    """`.npz` / `.npy` dictionary of the slim variables (xception_65/..., aspp*, decoder/..., logits/...) -> the dict
    RefinementNet.load_params takes."""
    return _select(_read_dict(path), refnet_param_shapes(middle_units, n_classes), "refinement_net")


def load_reid_net_variables(path):
    """`.npz` / `.npy` dictionary or TensorFlow checkpoint prefix of the ReID network's variables (conv0/W, res<k>/..., conv1/...,
    fc1/..., fc2/..., outputTriplet/...) -> the dict ReIDNet.load_params takes."""
    return _select(_read_dict(path), reid_param_shapes(), "ReID_net")


def save_variables(path, params):
    """The inverse (tests, weight conversion scripts): np.savez of a {name: array} dict."""
    np.savez(path, **{k: np.asarray(v) for k, v in params.items()})
