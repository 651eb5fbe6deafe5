"""The drop-in claim, checked with the UNMODIFIED reference trainers (marker ``reference``; CPU twin of
tests/test_gpu_reference_trainers.py).

The reference's MainPlayer / CentralizedAgent / run_sim / XDPlayer / CentralizedMultiAgent / MixedAgent are run
twice with the same seeds: on the reference's own ``SyncVectorEnv([SimplifiedOvercooked] * N)`` and on this
package's ``VectorMultiAgentEnv`` (vector_api.py) with the adapter's dtype conventions (int8 observations, int32
rewards / dones, static views) — here backed by the CPU oracle (tests/oracle_env.py), on the GPU box by
``B200Overcooked``.  Everything the trainers leave behind must be identical: every SharedReplayBuffer field after
``chooseinsert`` / ``diaginsert`` / ``partinsert``, the episode scores, the returns, the PPO update's statistics and
the updated actor weights.
"""
import pytest
import torch

import trainer_harness as th
from oracle_env import OracleOvercooked

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ns():
    return th.load()


@pytest.mark.parametrize("layout,N,T,horizon", [("simple", 6, 12, 5), ("random1", 4, 9, 20)])
def test_main_player_with_centralized_partner(ns, tmp_path, layout, N, T, horizon):
    args = th.make_args(ns, layout, N, T)
    ref = th.run_main_player(ns, th.reference_env(ns, layout, N, horizon, "cpu"), "cpu", args, tmp_path / "ref")
    new = th.run_main_player(ns, OracleOvercooked(layout, N, horizon), "cpu", args, tmp_path / "new")
    for e, (a, b) in enumerate(zip(ref["episodes"], new["episodes"])):
        th.assert_buffers_equal(a["buffer"], b["buffer"], "episode %d" % e)
        assert a["scores"] == b["scores"]
        assert torch.equal(a["returns"], b["returns"])
        assert a["train_infos"] == b["train_infos"]
    assert all(torch.equal(x, y) for x, y in zip(ref["actor"], new["actor"]))
    buf = ref["episodes"][0]["buffer"]
    if horizon < T:  # episodes ended inside the rollout: masks carry the dones, scores were recorded
        assert (buf["masks"][1:] == 0).any() and len(ref["episodes"][0]["scores"]) > 0


def test_run_sim(ns, tmp_path):
    args = th.make_args(ns, "simple", 5, 8)
    ref = th.run_sim_text(ns, th.reference_env(ns, "simple", 5, 200, "cpu"), "cpu", args, tmp_path / "ref")
    new = th.run_sim_text(ns, OracleOvercooked("simple", 5, 200), "cpu", args, tmp_path / "new")
    assert ref == new and "STDEV" in ref


def test_xd_player_slices_and_mixed_play(ns, tmp_path):
    L, threads, horizon = 6, 3, 4
    args = th.make_args(ns, "simple", threads, L, extra=["--mp_weight", "0.5"])
    ref = th.run_xd_player(ns, lambda n: th.reference_env(ns, "simple", n, horizon, "cpu"), "cpu", args, tmp_path / "r", threads)
    new = th.run_xd_player(ns, lambda n: OracleOvercooked("simple", n, horizon), "cpu", args, tmp_path / "n", threads)
    for k in ("sp", "xp0", "xp1", "mp"):
        th.assert_buffers_equal(ref[k], new[k], k)
    assert ref["scores"] == new["scores"] and ref["mp_scores"] == new["mp_scores"] and ref["best_i"] == new["best_i"]
    assert any(len(s) for s in ref["scores"])
