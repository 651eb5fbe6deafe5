"""``generate_env`` — same factory signature as the reference's train/env_utils.py:10-28.

``use_baseline=True`` selects the reference's pure-Python envs in the reference; this
package ships no CPU simulator (by design), so that flag raises here.
"""
from __future__ import annotations


def generate_env(name, num_envs, layout="simple", use_env_cpu=False, use_baseline=False, gpu_id=0, horizon=200, seed=0):
    if use_baseline:
        raise RuntimeError("use_baseline=True selects the reference's Python envs; this package is GPU only")
    if name == "balance":
        from .balance_env import B200BalanceBeam
        return B200BalanceBeam(num_envs, gpu_id, debug_compile=False, use_env_cpu=use_env_cpu, seed=seed)
    if name == "overcooked":
        from .overcooked_env import B200Overcooked
        return B200Overcooked(layout, num_envs, gpu_id, debug_compile=False, use_env_cpu=use_env_cpu, horizon=horizon,
                              seed=seed)
    raise Exception("Invalid environment name")
