"""Layout front-end vs the reference parser output (tests/golden/layouts.json,
generated from envs/overcooked2_env.py:171-291 by tests/golden/make_golden.py)."""
import ctypes
import json
import os

import pytest

from diverse_conventions_b200 import layouts


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "layouts.json")) as f:
        return json.load(f)


def test_all_layouts_match_reference_parser(golden_dir):
    gold = _golden(golden_dir)
    assert len(gold) == 63
    for key, ref in gold.items():
        name, mp = key.split("|")
        mp = None if mp == "None" else int(mp)
        assert layouts.get_base_layout_params(name, 400, mp) == ref, key


def test_classic_aliases():
    for nice, short in layouts.CLASSIC_LAYOUTS.items():
        assert layouts.load_layout(nice, 400).as_dict() == layouts.load_layout(short, 400).as_dict()


def test_layout_file_path(tmp_path):
    p = tmp_path / "mine.layout"
    p.write_text('{"grid": """XXPXX\n O  2O\n X1  X\n XDXSX""", "cook_time": 7, "delivery_reward": 11,'
                 ' "rew_shaping_params": {"PLACEMENT_IN_POT_REW": 1, "DISH_PICKUP_REWARD": 2, "SOUP_PICKUP_REWARD": 4}}')
    lp = layouts.load_layout(str(p), 50)
    assert (lp.width, lp.height, lp.num_players, lp.horizon) == (5, 4, 2, 50)
    assert lp.recipe_times == [7] * 16 and lp.recipe_values == [11] * 16
    assert (lp.placement_in_pot_rew, lp.dish_pickup_rew, lp.soup_pickup_rew) == (1, 2, 4)


def test_onion_tomato_time_tables():
    d = layouts.builtin_layout_dict("simple")
    del d["cook_time"], d["delivery_reward"]
    d.update(onion_time=5, tomato_time=3, onion_value=7, tomato_value=2)
    lp = layouts.parse_layout_dict(d, 400)
    assert lp.recipe_times[4 * 2 + 1] == 13 and lp.recipe_values[4 * 3 + 0] == 21


def test_config_struct_roundtrip():
    lp = layouts.load_layout("unident_s", 400)
    c = lp.to_config()
    assert c.struct_size == ctypes.sizeof(layouts.ocb_config)
    assert (c.width, c.height, c.num_players) == (9, 5, 2)
    assert list(c.terrain)[: lp.size] == lp.terrain
    assert layouts.io_bytes_per_world_step(lp) == 1820
    assert layouts.io_bytes_per_world_step(layouts.load_layout("simple", 400)) == 820


def test_validation_rejects_open_border_and_long_cook():
    d = layouts.builtin_layout_dict("simple")
    d["grid"] = "XXPXX\nO  2 \nX1  X\nXDXSX"
    with pytest.raises(ValueError):
        layouts.parse_layout_dict(d, 400).to_config()
    d = layouts.builtin_layout_dict("simple")
    d["cook_time"] = 500
    with pytest.raises(ValueError):
        layouts.parse_layout_dict(d, 400).to_config()
    with pytest.raises(FileNotFoundError):
        layouts.load_layout("no_such_layout", 400)


def test_validation_rejects_two_players_on_one_start_cell():
    lp = layouts.load_layout("simple", 400)
    lp.start_player_x[1], lp.start_player_y[1] = lp.start_player_x[0], lp.start_player_y[0]
    with pytest.raises(ValueError):
        lp.to_config()
