"""Import shim for the UNMODIFIED Python reference (test infrastructure only).

The reference is looked up at ``/root/reference`` (the build container) and, failing
that, at ``baseline/_ref`` — the verbatim, git-ignored install made by
``tools/install_reference.py`` that travels to the GPU box with the tree.  Nothing reads
``/root/reference`` at run time on the GPU box.  Users: ``tests/golden/make_*.py``
(golden vectors), the ``reference``-marked tests (oracle restatement and the CUDA path
against the live reference classes, trainers included) and the reference arm of
``bench.py``.  The product package never imports this module.

The reference imports ``gym`` (not installed) and the compiled Madrona modules
(``build.madrona_*``, not buildable offline); neither is touched by the pure-Python
env (envs/overcooked2_env.py:294-343, envs/overcooked2_reimplement.py), so tiny stub
modules are injected before the import.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.abspath(os.path.join(_HERE, "..", "baseline", "_ref"))


def _resolve_root() -> str:
    env = os.environ.get("OCB_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", INSTALLED_ROOT):
        if os.path.isfile(os.path.join(cand, "envs", "overcooked2_reimplement.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "overcooked2_reimplement.py"))


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Space:
            pass

        class Discrete(Space):
            def __init__(self, n):
                self.n, self.shape = n, ()

        class MultiBinary(Space):
            def __init__(self, n):
                self.n, self.shape = n, tuple(int(x) for x in n)

        class MultiDiscrete(Space):
            def __init__(self, nvec):
                self.nvec, self.shape = nvec, (len(nvec),)

        gym.Env = type("Env", (), {})
        gym.spaces = spaces
        spaces.Space, spaces.Discrete = Space, Discrete
        spaces.MultiBinary, spaces.MultiDiscrete = MultiBinary, MultiDiscrete
        vector = types.ModuleType("gym.vector")
        vector_env = types.ModuleType("gym.vector.vector_env")
        vector_env.VectorEnv = type("VectorEnv", (), {"__init__": lambda self, *a, **k: None})
        vector.vector_env = vector_env
        gym.vector = vector
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.vector": vector,
                            "gym.vector.vector_env": vector_env})
    if "shutup" not in sys.modules:  # train/testing.py:1 silences warnings with it; not installed here
        sh = types.ModuleType("shutup")
        sh.please = lambda: None
        sys.modules["shutup"] = sh
    if "build" not in sys.modules:
        b = types.ModuleType("build")
        b.__path__ = []
        sys.modules["build"] = b
        for n in ("madrona_python", "madrona_simplecooked_example_python", "madrona_balance_example_python"):
            m = types.ModuleType("build." + n)
            sys.modules["build." + n] = m
            setattr(b, n, m)


def load():
    """Return a namespace with the reference classes used as the parity target."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    for p in (os.path.join(REFERENCE_ROOT, "train"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    ns = types.SimpleNamespace()
    from envs.overcooked2_env import SimplifiedOvercooked, get_base_layout_params  # noqa
    from envs import overcooked2_reimplement as reimpl  # noqa
    from pantheonrl_extension.vectorenv import SyncVectorEnv  # noqa
    ns.SimplifiedOvercooked = SimplifiedOvercooked
    ns.get_base_layout_params = get_base_layout_params
    ns.reimpl = reimpl
    ns.SyncVectorEnv = SyncVectorEnv
    ns.layouts_dir = os.path.join(REFERENCE_ROOT, "envs", "layouts")
    return ns


def load_balance():
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    for p in (os.path.join(REFERENCE_ROOT, "train"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from envs.balance_beam_env import PantheonLine  # noqa
    return PantheonLine


def load_policy():
    """R_Actor / R_Critic from the reference (train/MAPPO/r_actor_critic.py)."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    for p in (os.path.join(REFERENCE_ROOT, "train"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from MAPPO.r_actor_critic import R_Actor, R_Critic  # noqa
    return R_Actor, R_Critic


def load_trainers():
    """The reference's rollout-side trainer classes, imported unmodified (train/ uses script-style absolute imports):
    MainPlayer (train/MAPPO/main_player.py), CentralizedAgent / CentralizedMultiAgent / DecentralizedAgent
    (train/partner_agents.py), run_sim (train/testing.py:39-59), XDPlayer (train/XD/xd_player.py),
    SharedReplayBuffer (train/MAPPO/utils/shared_buffer.py), get_config (train/config.py)."""
    ns = load()
    from MAPPO.main_player import MainPlayer  # noqa
    from MAPPO.utils.shared_buffer import SharedReplayBuffer  # noqa
    from MAPPO.rMAPPOPolicy import R_MAPPOPolicy, rMAPPOWrapper  # noqa
    from partner_agents import CentralizedAgent, CentralizedMultiAgent, DecentralizedAgent  # noqa
    from XD.xd_player import XDPlayer  # noqa
    from XD.MCPolicy import MCPolicy  # noqa
    from config import get_config  # noqa
    import testing  # noqa
    ns.MainPlayer, ns.SharedReplayBuffer = MainPlayer, SharedReplayBuffer
    ns.R_MAPPOPolicy, ns.rMAPPOWrapper = R_MAPPOPolicy, rMAPPOWrapper
    ns.CentralizedAgent, ns.CentralizedMultiAgent, ns.DecentralizedAgent = CentralizedAgent, CentralizedMultiAgent, DecentralizedAgent
    ns.XDPlayer, ns.MCPolicy, ns.get_config, ns.run_sim = XDPlayer, MCPolicy, get_config, testing.run_sim
    return ns
