"""The numpy restatement of the reference's returns / GAE / advantage normalisation against the
golden vectors the reference itself produced (tests/golden/make_returns_golden.py)."""
import os

import numpy as np
import pytest

from oracle import returns_oracle as ro

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "returns.npz"))
CASES = {"gae_vn": (True, True), "gae_plain": (True, False), "disc_vn": (False, True), "disc_plain": (False, False),
         "gae_vn_ptl": (True, True)}


def oracle_case(name):
    use_gae, vn = CASES[name]
    mean, std = ro.valuenorm_mean_std(*G[name + "_vn_state"]) if vn else (0.0, 1.0)
    return ro.compute_returns(G["value_preds"], G["rewards"], G["done"], 0.99, 0.95, use_gae, mean, std)


@pytest.mark.parametrize("name", sorted(CASES))
def test_returns_oracle_is_bit_identical_to_reference(name):
    ret, adv = oracle_case(name)
    T = G["rewards"].shape[0]
    assert np.array_equal(ret[:T], G[name + "_returns"][:T])
    if not CASES[name][0]:
        assert np.array_equal(ret[T], G[name + "_returns"][T])
    assert np.array_equal(adv, G[name + "_adv"])
    np.testing.assert_allclose(ro.normalize_advantages(adv), G[name + "_adv_norm"], rtol=2e-6, atol=2e-6)


def test_masks_cut_the_bootstrap_at_episode_ends():
    ret, _ = oracle_case("gae_plain")
    v, r, d = G["value_preds"], G["rewards"], G["done"]
    t, n = np.argwhere(d)[0]
    # at a done step the return is r + (gae of nothing): delta = r - v, returns = r
    assert np.allclose(ret[t, :, n], r[t, :, n].astype(np.float32), atol=1e-5)
