"""Layout front-end: `.layout` description -> flat simulator parameters.

Re-states the *semantics* of the reference's ``get_base_layout_params``
(envs/overcooked2_env.py:171-291) and ``read_layout_dict`` (:18-24): terrain codes
``[' ','P','X','O','D','S','T'] -> 0..6`` (:152,202), player start cells taken from the
digit characters listed in ``PLAYER_NUMS`` (:153-156,189-206), shaping rewards
(:136-140,208-213) and the 16-entry recipe time / value tables indexed by
``4*onions + tomatoes`` (:231-284).

The grids of the layouts shipped with the reference (envs/layouts/*.layout) are kept here
as plain data so that the GPU box (which has no reference checkout) can run them; a path
ending in ``.layout`` is read from disk instead, as in the reference.
"""
from __future__ import annotations

import ast
import ctypes
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

MAX_PLAYERS = 4
MAX_CELLS = 256
NUM_RECIPES = 16
MAX_COOK_TIME = 120
MAX_INGREDIENTS = 3

TERRAIN_CHARS = " PXODST"  # index == terrain code
PLAYER_CHARS = list("1234567890") + list("!@#$%^&*()") + list("abcdefghij") + list("klmnopqrst")

DEFAULT_SHAPING = {"PLACEMENT_IN_POT_REW": 3, "DISH_PICKUP_REWARD": 3, "SOUP_PICKUP_REWARD": 5}

# name used in the paper / BASELINE.json -> layout file name (train/test_vs_bc.py:39-49)
CLASSIC_LAYOUTS = {
    "cramped_room": "simple",
    "asymmetric_advantages": "unident_s",
    "coordination_ring": "random1",
    "forced_coordination": "random0",
    "counter_circuit": "random3",
}

# Grids of the shipped layouts; every one uses cook_time 20, delivery_reward 20 and
# the default shaping rewards.
_GRIDS: Dict[str, Sequence[str]] = {
    "simple": ("XXPXX", "O  2O", "X1  X", "XDXSX"),
    "simple_single": ("XXPXX", "O   O", "X1  X", "XDXSX"),
    "simple_tomato": ("XXPXX", "T  2T", "X1  O", "XXDSX"),
    "mdp_test": ("XXPXX", "O  2O", "T1  T", "XDPSX"),
    "five_by_five": ("XDPXX", "X   S", "O 2 X", "X1  D", "XOXPX"),
    "random0": ("XXXPX", "O X1P", "O2X X", "D X X", "XXXSX"),
    "random1": ("XXXPX", "X 1 P", "D2X X", "O   X", "XOSXX"),
    "random2": ("XXXPX", "O X1P", "O2X X", "D X X", "XXXSX"),
    "random3": ("XXXPPXXX", "X  2   X", "D XXXX S", "X  1   X", "XXXOOXXX"),
    "unident_s": ("XXXXXXXXX", "O XSXOX S", "X   P 1 X", "X2  P   X", "XXXDXDXXX"),
    "unident": ("XXXXXXXXXXX", "O XXSXOXX S", "X    P  1 X", "X2   P    X", "XXXXDXDXXXX"),
    "schelling_s": ("XSPDX", "X 1 X", "O   O", "X 2 X", "XDPSX"),
    "schelling": ("XXSPDXX", "X  1  X", "X  X  X", "O     O", "X  X  X", "X  2  X", "XXDPSXX"),
    "multiplayer_schelling": ("XXSPDXX", "X  1  X", "X  X  X", "O3   4O", "X  X  X", "X  2  X", "XXDPSXX"),
    "scenario1_s": ("XXOXDXX", "X 1X2 X", "X  X  X", "X     X", "XSXXPPX"),
    "scenario2_s": ("XXOXXXX", "S     O", "D 1 2 X", "XXXPXXX"),
    "scenario2": ("XXXXXOXXXX", "S        O", "D    1 2 X", "XXXXXXPXXX"),
    "scenario3": ("XXXXXOXXXX", "S     XXPX", "X    1   X", "D XXXXXX X", "X     2  O", "XXXXXXXXXX"),
    "scenario4": ("XXXXXOXXXX", "S      XPX", "D    1   X", "XXXXXXXX X", "XXXXXX2  O", "XXXXXXXXXX"),
    "small_corridor": ("XXXXXOXDXXXXX", "X  1  X  2  X", "X  XXXXXXX  X", "X           X", "XSXXXXXXXXPPX"),
    "corridor": ("XXXXXOXXDXXXXX", "X  1  XX  2  X", "X     XX     X", "X  XXXXXXXX  X", "X            X",
                 "X  XXXXXXXX  X", "X     XX     X", "X     XX     X", "XXXXXSXXPPXXXX"),
}


def builtin_layout_names() -> List[str]:
    return sorted(_GRIDS)


def builtin_layout_dict(name: str) -> dict:
    name = CLASSIC_LAYOUTS.get(name, name)
    if name not in _GRIDS:
        raise FileNotFoundError("unknown layout '%s'" % name)
    return {"grid": "\n".join(_GRIDS[name]), "start_order_list": None, "cook_time": 20,
            "num_items_for_soup": 3, "delivery_reward": 20, "rew_shaping_params": None}


class ocb_config(ctypes.Structure):
    """ctypes twin of ``struct ocb_config`` (include/ocb.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("num_players", ctypes.c_int32),
        ("horizon", ctypes.c_int32),
        ("placement_in_pot_rew", ctypes.c_int32),
        ("dish_pickup_rew", ctypes.c_int32),
        ("soup_pickup_rew", ctypes.c_int32),
        ("recipe_values", ctypes.c_int32 * NUM_RECIPES),
        ("recipe_times", ctypes.c_int32 * NUM_RECIPES),
        ("start_player_x", ctypes.c_int32 * MAX_PLAYERS),
        ("start_player_y", ctypes.c_int32 * MAX_PLAYERS),
        ("terrain", ctypes.c_uint8 * MAX_CELLS),
    ]


@dataclass
class LayoutParams:
    """Same fields (and names) as the dict returned by the reference parser."""
    height: int
    width: int
    terrain: List[int]
    num_players: int
    start_player_x: List[int]
    start_player_y: List[int]
    placement_in_pot_rew: int
    dish_pickup_rew: int
    soup_pickup_rew: int
    recipe_times: List[int]
    recipe_values: List[int]
    horizon: int
    extra: dict = field(default_factory=dict)

    @property
    def size(self) -> int:
        return self.width * self.height

    @property
    def channels(self) -> int:
        return 5 * self.num_players + 10

    def as_dict(self) -> dict:
        d = dict(self.extra)
        d.update(height=self.height, width=self.width, terrain=list(self.terrain),
                 num_players=self.num_players, start_player_x=list(self.start_player_x),
                 start_player_y=list(self.start_player_y),
                 placement_in_pot_rew=self.placement_in_pot_rew, dish_pickup_rew=self.dish_pickup_rew,
                 soup_pickup_rew=self.soup_pickup_rew, recipe_times=list(self.recipe_times),
                 recipe_values=list(self.recipe_values), horizon=self.horizon)
        return d

    def validate(self) -> None:
        """Constraints of the CUDA path (the Python reference is unbounded)."""
        if not (1 <= self.num_players <= MAX_PLAYERS):
            raise ValueError("num_players must be in 1..%d" % MAX_PLAYERS)
        if self.size > MAX_CELLS or self.width < 1 or self.height < 1:
            raise ValueError("grid must have 1..%d cells" % MAX_CELLS)
        if len(self.terrain) != self.size:
            raise ValueError("ragged grid: terrain has %d cells, expected %d" % (len(self.terrain), self.size))
        for t in self.recipe_times:
            if not (0 <= t <= MAX_COOK_TIME):
                raise ValueError("recipe_times must be in 0..%d" % MAX_COOK_TIME)
        W, H = self.width, self.height
        for y in range(H):
            for x in range(W):
                if (x in (0, W - 1) or y in (0, H - 1)) and self.terrain[y * W + x] == 0:
                    raise ValueError("border cell (%d,%d) is walkable; agents could leave the grid" % (x, y))
        for x, y in zip(self.start_player_x, self.start_player_y):
            if self.terrain[y * W + x] != 0:
                raise ValueError("player start cell is not AIR")
        starts = list(zip(self.start_player_x, self.start_player_y))[:self.num_players]
        if len(set(starts)) != len(starts):
            raise ValueError("two players start on the same cell")

    def to_config(self) -> ocb_config:
        self.validate()
        c = ocb_config()
        c.struct_size = ctypes.sizeof(ocb_config)
        c.width, c.height, c.num_players, c.horizon = self.width, self.height, self.num_players, self.horizon
        c.placement_in_pot_rew = self.placement_in_pot_rew
        c.dish_pickup_rew = self.dish_pickup_rew
        c.soup_pickup_rew = self.soup_pickup_rew
        for i in range(NUM_RECIPES):
            c.recipe_values[i] = int(self.recipe_values[i])
            c.recipe_times[i] = int(self.recipe_times[i])
        for i in range(self.num_players):
            c.start_player_x[i] = self.start_player_x[i]
            c.start_player_y[i] = self.start_player_y[i]
        for i, t in enumerate(self.terrain):
            c.terrain[i] = t
        return c


def _count(order: dict, what: str) -> int:
    return sum(1 for ing in order["ingredients"] if ing == what)


def parse_layout_dict(d: dict, horizon: int, max_num_players: Optional[int] = None) -> LayoutParams:
    d = dict(d)
    rows = [r.strip() for r in d.pop("grid").split("\n")]
    d.pop("start_order_list", None)
    d.pop("num_items_for_soup", None)

    cells = [list(r) for r in rows]
    starts: List[Optional[tuple]] = [None] * 64
    for y, row in enumerate(cells):
        for x, ch in enumerate(row):
            if ch in PLAYER_CHARS:
                row[x] = " "
                k = PLAYER_CHARS.index(ch)
                if max_num_players is None or k < max_num_players:
                    starts[k] = (x, y)
    n_players = sum(1 for s in starts if s is not None)
    starts = starts[:n_players]
    if any(s is None for s in starts):
        raise ValueError("player start markers must be contiguous from '1'")

    height, width = len(cells), len(cells[0])
    terrain = [TERRAIN_CHARS.index(ch) for row in cells for ch in row]

    shaping = d.pop("rew_shaping_params", None) or DEFAULT_SHAPING
    all_orders = d.pop("start_all_orders", None) or []
    d.pop("start_bonus_orders", None)
    d.pop("order_bonus", None)

    n = MAX_INGREDIENTS + 1
    times = [20] * (n * n)
    if "onion_time" in d and "tomato_time" in d:
        ot, tt = d.pop("onion_time"), d.pop("tomato_time")
        times = [o * ot + t * tt for o in range(n) for t in range(n)]
    if "recipe_times" in d:
        for order, tm in zip(all_orders, d["recipe_times"]):
            times[n * _count(order, "onion") + _count(order, "tomato")] = tm
    if "cook_time" in d:
        times = [d.pop("cook_time")] * (n * n)
    d.pop("recipe_times", None)

    values = [20] * (n * n)
    if "onion_value" in d and "tomato_value" in d:
        ov, tv = d.pop("onion_value"), d.pop("tomato_value")
        values = [o * ov + t * tv for o in range(n) for t in range(n)]
    if "recipe_values" in d:
        for order, val in zip(all_orders, d["recipe_values"]):
            values[n * _count(order, "onion") + _count(order, "tomato")] = val
    if "delivery_reward" in d:
        values = [d.pop("delivery_reward")] * (n * n)
    d.pop("recipe_values", None)

    return LayoutParams(
        height=height, width=width, terrain=terrain, num_players=n_players,
        start_player_x=[s[0] for s in starts], start_player_y=[s[1] for s in starts],
        placement_in_pot_rew=shaping["PLACEMENT_IN_POT_REW"], dish_pickup_rew=shaping["DISH_PICKUP_REWARD"],
        soup_pickup_rew=shaping["SOUP_PICKUP_REWARD"], recipe_times=times, recipe_values=values,
        horizon=horizon, extra=d)


def load_layout(layout_name: str, horizon: int, max_num_players: Optional[int] = None) -> LayoutParams:
    """Drop-in for ``get_base_layout_params(layout_name, horizon, max_num_players)``."""
    if layout_name.endswith(".layout"):
        with open(layout_name, "r") as f:
            d = ast.literal_eval(f.read())
    else:
        d = builtin_layout_dict(layout_name)
    return parse_layout_dict(d, horizon, max_num_players)


def get_base_layout_params(layout_name: str, horizon, max_num_players=None) -> dict:
    """Same name / return type as the reference function (envs/overcooked2_env.py:171)."""
    return load_layout(layout_name, horizon, max_num_players).as_dict()


def io_bytes_per_world_step(p: LayoutParams) -> int:
    """Algorithmic I/O bytes per world-step:
    obs int8 write + action int32 read + reward int32 write + done int32 write."""
    P = p.num_players
    return P * p.size * p.channels + 4 * P + 4 * P + 4
