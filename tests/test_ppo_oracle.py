"""The PPO-minibatch oracle (oracle/ppo_oracle.py) against goldens produced by the reference's own
feed_forward_generator / evaluate_actions / R_MAPPO.ppo_update (tests/golden/make_ppo_golden.py)."""
import os

import numpy as np
import pytest

from oracle import ppo_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ppo.npz"))
CASES = ["default", "mse_unclipped", "no_valuenorm", "inactive", "inactive_nomask", "small_delta"]
T, N, P, NMB = int(G["T"]), int(G["N"]), int(G["P"]), int(G["num_mini_batch"])


def case_cfg(name):
    c = G[name + "_cfg"]
    return dict(clip=c[0], delta=c[1], use_clipped_value_loss=bool(c[2]), use_huber_loss=bool(c[3]),
                use_value_active_masks=bool(c[5]), use_policy_active_masks=bool(c[6])), bool(c[4])


def test_flat_index_mapping_reproduces_the_generators_minibatch():
    rows = po.flat_to_rows(G["mb0_flat_index"], N, P)
    obs = G["obs_seat_major"].reshape(T * P * N, *G["obs_seat_major"].shape[3:])
    assert np.array_equal(obs[rows], G["mb0_obs_batch"])
    for key, src in (("mb0_actions_batch", "default_actions"), ("mb0_value_preds_batch", "default_value_preds"),
                     ("mb0_return_batch", "default_returns"), ("mb0_old_action_log_probs_batch", "default_old_logp"),
                     ("mb0_adv_targ", "default_adv")):
        assert np.array_equal(G[src].reshape(-1)[rows], G[key].reshape(-1).astype(G[src].dtype)), key
    # the three index lists partition a permutation of all samples
    allrows = np.concatenate([po.flat_to_rows(G["mb%d_flat_index" % i], N, P) for i in range(NMB)])
    assert np.array_equal(np.sort(allrows), np.arange(T * N * P))


@pytest.mark.parametrize("name", CASES)
def test_head_and_loss_match_the_reference(name):
    kw, use_vn = case_cfg(name)
    state = G[name + "_vn_state0"] if use_vn else None
    inactive = name.startswith("inactive")
    for i in range(NMB):
        k = "%s_mb%d_" % (name, i)
        rows = po.flat_to_rows(G["mb%d_flat_index" % i], N, P)
        pick = lambda a: G[name + "_" + a].reshape(-1)[rows]
        logp, ent = po.evaluate_head(G[k + "logits"], pick("actions"))
        np.testing.assert_allclose(logp, G[k + "logp"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(ent, G[k + "entropy"], rtol=0, atol=2e-6)
        out = po.ppo_loss(G[k + "logp"], G[k + "entropy"], G[k + "values"], pick("old_logp"), pick("adv"), pick("value_preds"),
                          pick("returns"), pick("active") if inactive else None, state, **kw)
        pl, vl, de, rm = G[k + "losses"]
        np.testing.assert_allclose(out["policy_loss"], pl, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["value_loss"], vl, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["dist_entropy"], de, rtol=2e-6)
        np.testing.assert_allclose(out["ratio_mean"], rm, rtol=2e-6)
        np.testing.assert_allclose(out["imp_weights"], G[k + "imp_weights"], rtol=1e-6)
        np.testing.assert_allclose(out["dlogp"], G[k + "dlogp"], rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(out["dvalues"], G[k + "dvalues"], rtol=2e-5, atol=1e-9)
        if use_vn:
            np.testing.assert_allclose(out["vn_state"], G[k + "vn_state"], rtol=1e-6)
            state = out["vn_state"]
