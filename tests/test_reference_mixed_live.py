"""Live check of the mixed-play restatement (oracle/mixed_oracle.py) against the UNMODIFIED reference for several
episode lengths / horizons / layouts (build container only, marker `reference`): the committed golden
(tests/golden/mixed.npz) pins one configuration (L = 7); here the reference's XDPlayer.collect_mp_episode + MixedAgent +
SharedReplayBuffer.diaginsert / partinsert are run again (tests/golden/make_mixed_golden.py: collect) and the forced-main
schedule, the record placement and the env stream are compared."""
import os
import sys

import numpy as np
import pytest

from diverse_conventions_b200 import layouts
from oracle import mixed_oracle as mo
from oracle.c_oracle import COracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("L,horizon,layout,seed", [(3, 2, "simple", 1), (4, 3, "random1", 9), (10, 6, "simple", 3),
                                                   (13, 4, "unident_s", 4)])
def test_schedule_and_placement_match_the_live_reference(L, horizon, layout, seed):
    import make_mixed_golden as mg
    g = mg.collect(L=L, HORIZON=horizon, LAYOUT=layout, SEED=seed)
    G = L - 1
    forced, slot = mo.schedule(L)
    used = g["turn_values"] >= 1000  # [2L, N, 2]: the stub partner policy answers values >= 1000
    for seat in range(2):
        assert not used[:, :, seat][forced].any()          # forced worlds always act with the main policy
    if L >= 4:
        assert used[~forced].any() and not used[~forced].all()
    for j in range(G):                                     # every world is recorded exactly once per buffer slot
        assert sorted(slot[:, j][slot[:, j] >= 0].tolist()) == list(range(L))
    for name, fill in (("obs", 0), ("actions", 0), ("values", 0), ("logp", 0), ("rewards", 0), ("masks", 1), ("active", 1)):
        want = g["buf_" + name]
        assert np.array_equal(mo.place(L, g["turn_" + name], world_axis=0, fill=fill), want[:L]), name
        if want.shape[0] == L + 1:
            assert (want[L] == fill).all(), name           # slot L is never written by the collection
    assert (g["buf_values"][:L] < 1000).all()              # only main-policy records reach the buffer

    lp = layouts.load_layout(layout, horizon)              # the env stream of the run replays through the C oracle
    orc = COracle(lp, G)
    obs = orc.observe()
    for s in range(2 * L):
        assert np.array_equal(obs, np.moveaxis(g["turn_obs"][s], 1, 0)), s
        obs, rew, done = orc.step(np.ascontiguousarray(g["turn_actions"][s].T))
        assert np.array_equal(rew.T.astype(np.float32), g["turn_rewards"][s]), s
        assert np.array_equal(1.0 - done, g["turn_masks"][s][:, 0]), s
