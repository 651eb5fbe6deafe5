#!/usr/bin/env python
"""Install the UNMODIFIED Python reference under ``baseline/_ref`` (git-ignored, NOT gpurun-ignored, so it
travels to the GPU box with the tree; the GPU box has no /root/reference).

Why not pip: the reference's setup.py uses ``find_packages()`` but neither ``envs/``, ``pantheonrl_extension/``
nor ``train/`` carries an ``__init__.py`` (they are namespace packages / script directories with absolute
imports such as ``from MAPPO.main_player import MainPlayer``), so
``pip install --no-index --no-deps --target baseline/_ref /root/reference`` builds a wheel that contains only
the dist-info (tried, recorded in DESIGN.md).  The install is therefore a verbatim copy of the three pure-Python
trees the hot path and its callers live in; nothing is patched, nothing of it is committed:

    envs/                  SimplifiedOvercooked, PantheonLine, layouts     (the CPU parity target)
    pantheonrl_extension/  VectorMultiAgentEnv, SyncVectorEnv, VectorAgent (the drop-in boundary)
    train/                 MainPlayer, CentralizedAgent, XDPlayer, R_Actor/R_Critic, SharedReplayBuffer ...

Consumers (test infrastructure and the reference arm of bench.py only): ``oracle/ref_shim.py`` resolves
/root/reference first and baseline/_ref second.  The product package never imports either.
"""
import os
import shutil
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.environ.get("OCB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("envs", "pantheonrl_extension", "train")
KEEP_EXT = (".py", ".layout", ".json")


def _ignore(directory, names):
    drop = []
    for n in names:
        full = os.path.join(directory, n)
        if os.path.isdir(full):
            if n in ("results", "__pycache__", "wandb"):
                drop.append(n)
        elif not n.endswith(KEEP_EXT):
            drop.append(n)
    return drop


def install(force: bool = False) -> str:
    if not os.path.isfile(os.path.join(SRC, "envs", "overcooked2_reimplement.py")):
        if os.path.isfile(os.path.join(DST, "envs", "overcooked2_reimplement.py")):
            return DST  # GPU box: the travelled copy is all there is
        raise RuntimeError("reference checkout not found at %s" % SRC)
    stamp = os.path.join(DST, ".installed_from")
    if not force and os.path.isfile(stamp):
        return DST
    for t in TREES:
        dst = os.path.join(DST, t)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, t), dst, ignore=_ignore)
    with open(stamp, "w") as f:
        f.write(SRC + "\n")
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
