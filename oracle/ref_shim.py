"""Import shim for the UNMODIFIED Python reference (test infrastructure only).

Only usable where ``/root/reference`` exists (the build container); the GPU box does
not have it, so nothing under ``tests -m gpu``, ``smoke()`` or ``bench.py`` may call
this.  It is used by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and by the optional ``-m "not gpu"`` tests that pin the oracle restatement
against the live reference.

The reference imports ``gym`` (not installed) and the compiled Madrona modules
(``build.madrona_*``, not buildable offline); neither is touched by the pure-Python
env (envs/overcooked2_env.py:294-343, envs/overcooked2_reimplement.py), so tiny stub
modules are injected before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OCB_REFERENCE_ROOT", "/root/reference")


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
