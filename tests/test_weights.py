"""Weight dictionary readers (premvos_b200/weights.py): tensorpack / slim `.npz` dumps with tensor suffixes, tower prefixes
and optimizer slots are normalised to exactly the variables the networks take; missing or mis-shaped variables raise."""
import numpy as np
import pytest

from premvos_b200 import synth, weights


def test_proposal_net_npz_roundtrip_with_noise(tmp_path):
    nb = (1, 1, 1, 1)
    P = synth.propnet_synthetic_params(3, nb)
    dump = {"tower0/" + k + ":0": v for k, v in P.items()}
    dump["global_step:0"] = np.array(7)
    dump["learning_rate:0"] = np.array(0.003, np.float32)
    dump["tower0/conv0/W/Momentum:0"] = np.zeros_like(P["conv0/W"])
    dump["EMA/cost:0"] = np.array(1.0)
    weights.save_variables(tmp_path / "pn.npz", dump)
    got = weights.load_proposal_net_variables(tmp_path / "pn.npz", nb)
    assert list(got) == list(synth.propnet_param_shapes(nb))
    for k in P:
        np.testing.assert_array_equal(got[k], P[k])
        assert got[k].dtype == np.float32 and got[k].flags["C_CONTIGUOUS"]
    # the mask-head variables are demanded only with mode_mask
    with pytest.raises(KeyError):
        weights.load_proposal_net_variables(tmp_path / "pn.npz", nb, mode_mask=True)
    dump.update({k + ":0": v for k, v in synth.maskrcnn_synthetic_params(3).items()})
    weights.save_variables(tmp_path / "pnm.npz", dump)
    assert "maskrcnn/deconv/W" in weights.load_proposal_net_variables(tmp_path / "pnm.npz", nb, mode_mask=True)


def test_errors_and_npy(tmp_path):
    R = synth.refnet_synthetic_params(0, 0)
    np.save(tmp_path / "rn.npy", dict(R), allow_pickle=True)
    got = weights.load_refinement_net_variables(tmp_path / "rn.npy", middle_units=0)
    assert list(got) == list(synth.refnet_param_shapes(0)) and all(np.array_equal(got[k], R[k]) for k in R)
    bad = dict(R)
    first = next(iter(bad))
    bad[first] = bad[first][..., :-1]
    weights.save_variables(tmp_path / "bad.npz", bad)
    with pytest.raises(ValueError):
        weights.load_refinement_net_variables(tmp_path / "bad.npz", middle_units=0)
    del bad[first]
    weights.save_variables(tmp_path / "missing.npz", bad)
    with pytest.raises(KeyError):
        weights.load_refinement_net_variables(tmp_path / "missing.npz", middle_units=0)
    with pytest.raises(ValueError):
        weights.normalise_variable_names({"a:0": np.zeros(1), "tower0/a": np.zeros(1)})


def test_unknown_enclosing_scope_is_stripped(tmp_path):
    R = synth.refnet_synthetic_params(1, 0)
    weights.save_variables(tmp_path / "scoped.npz", {"deeplab_layer/" + k + ":0": v for k, v in R.items()})
    got = weights.load_refinement_net_variables(tmp_path / "scoped.npz", middle_units=0)
    assert all(np.array_equal(got[k], R[k]) for k in R)
    # ambiguous tails are not guessed
    amb = {"a/" + k: v for k, v in R.items()}
    amb.update({"b/" + k: v for k, v in R.items()})
    weights.save_variables(tmp_path / "amb.npz", amb)
    with pytest.raises(KeyError):
        weights.load_refinement_net_variables(tmp_path / "amb.npz", middle_units=0)
