"""Trajectory dump for the reference's browser replay (diverse_conventions_b200/trajectory.py): the state dicts built
from the packed states of the golden trajectories (produced by the reference env itself, tests/golden/make_golden.py)
carry exactly what the observation planes of the same steps encode, in the schema dictToState reads
(overcooked_flask/static/js/demo/replay.js:3792-3828)."""
import json
import os

import numpy as np
import pytest

from diverse_conventions_b200 import layouts, trajectory


@pytest.mark.parametrize("layout", ["simple", "random1", "unident_s", "simple_tomato"])
def test_state_dicts_agree_with_the_observation_planes(golden_dir, layout):
    g = np.load(os.path.join(golden_dir, "overcooked_%s.npz" % layout))
    lp = layouts.load_layout(layout, int(g["horizon"]))
    P, W, H = lp.num_players, lp.width, lp.height
    states, obs = g["states"], g["obs_head"]          # post-step states / observations of the first 64 steps
    traj = trajectory.build_trajectory(lp, states[:64], g["actions"][:64], g["rewards"][:64])
    json.loads(json.dumps(traj))                       # serialisable
    assert len(traj["ep_states"][0]) == 64 and len(traj["ep_actions"][0]) == 64 and len(traj["ep_rewards"][0]) == 64
    shift = 5 * P
    seen_held, seen_soup = False, False
    for t in range(64):
        d = traj["ep_states"][0][t]
        assert set(d) == {"players", "objects", "order_list"} and len(d["players"]) == P
        o = obs[t][0]                                  # player 0's view, [W, H, C]
        for i, pl in enumerate(d["players"]):
            x, y = pl["position"]
            assert o[x, y, i] == 1                     # position plane of player i (viewer 0: own index order)
            k = trajectory.DIRECTIONS.index(pl["orientation"])
            assert o[x, y, P + 4 * i + k] == 1         # orientation plane
            if pl["held_object"] is not None:
                seen_held = True
                assert pl["held_object"]["position"] == [x, y]
                name = pl["held_object"]["name"]
                if name == "onion":
                    assert o[x, y, shift + 9] == 1
                if name == "dish":
                    assert o[x, y, shift + 8] == 1
        for ob in d["objects"]:
            x, y = ob["position"]
            assert lp.terrain[y * W + x] in (1, 2)     # objects lie on pots / counters only
            if ob["name"] == "soup" and lp.terrain[y * W + x] == 1:
                seen_soup = True
                kind, n, cook = ob["state"]
                assert o[x, y, shift + 6] == cook and 1 <= n <= 3 and kind in ("onion", "tomato")
            if ob["name"] == "onion":
                assert o[x, y, shift + 9] == 1
            if ob["name"] == "dish":
                assert o[x, y, shift + 8] == 1
    assert seen_held
    if layout == "simple":
        assert seen_soup


def test_actions_and_grid_follow_the_replay_conventions():
    assert [trajectory.action_to_js(a) for a in range(6)] == [[0, -1], [0, 1], [1, 0], [-1, 0], [0, 0], "interact"]
    lp = layouts.load_layout("simple", 400)
    assert trajectory.terrain_rows(lp) == ["XXPXX", "O  2O", "X1  X", "XDXSX"]   # envs/layouts/simple.layout
    with pytest.raises(ValueError):
        trajectory.state_to_dict(lp, np.zeros(5, dtype=np.int32))
