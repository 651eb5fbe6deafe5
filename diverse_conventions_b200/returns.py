"""Returns / GAE / advantage normalisation on the device, over the seat-major rollout buffer
(``ocb_compute_returns`` / ``ocb_normalize_advantages``, csrc/ppo_kernels.cu).

Replaces ``SharedReplayBuffer.compute_returns`` (train/MAPPO/utils/shared_buffer.py:248-304: a
Python loop over the T steps), the ``ValueNorm.denormalize`` calls inside it
(train/MAPPO/utils/valuenorm.py:76-87) and the advantage normalisation at the top of
``R_MAPPO.train`` (train/MAPPO/r_mappo.py:174-182)."""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _native


class ocb_returns_cfg(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("use_gae", ctypes.c_int32), ("gamma", ctypes.c_double),
                ("gae_lambda", ctypes.c_double), ("vn_mean", ctypes.c_float), ("vn_std", ctypes.c_float)]


def valuenorm_mean_std(value_normalizer) -> Tuple[float, float]:
    """(debiased mean, sqrt of the clamped debiased variance) of a reference ``ValueNorm`` (or any
    object with ``running_mean_var()``, valuenorm.py:34-41) as Python floats; ``None`` -> (0, 1).
    Synchronises when the statistics live on a CUDA device — ``compute_returns`` uses
    ``valuenorm_mean_std_device`` for those instead."""
    if value_normalizer is None:
        return 0.0, 1.0
    mean, var = value_normalizer.running_mean_var()
    return float(mean.reshape(-1)[0]), float(torch.sqrt(var).reshape(-1)[0])


def valuenorm_mean_std_device(value_normalizer, device) -> Optional[torch.Tensor]:
    """float32 ``[mean, std]`` on ``device`` when the normaliser's statistics are CUDA tensors (the device-resident
    ``ValueNormState`` or a reference ``ValueNorm`` moved to the GPU): two tiny torch launches, no host
    synchronisation, so ``compute_returns`` stays asynchronous and graph-capturable.  ``None`` otherwise."""
    if value_normalizer is None:
        return None
    mean, var = value_normalizer.running_mean_var()
    if not (torch.is_tensor(mean) and mean.is_cuda):
        return None
    return torch.stack([mean.reshape(-1)[0], torch.sqrt(var).reshape(-1)[0]]).to(device=device, dtype=torch.float32)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def compute_returns(value_preds: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor, gamma: float = 0.99,
                    gae_lambda: float = 0.95, use_gae: bool = True, value_normalizer=None, normalize: bool = True,
                    out_returns: Optional[torch.Tensor] = None, out_advantages: Optional[torch.Tensor] = None,
                    group=None):
    """value_preds f32 ``[T+1,P,N]`` (slot T = bootstrap value), rewards int32 ``[T,P,N]``, dones int32
    ``[T,N]`` on one CUDA device -> ``(returns [T+1,P,N], advantages [T,P,N])``; the advantages are
    normalised in place ((a - mean) / (std + 1e-5)) when ``normalize``.  Asynchronous on torch's
    current stream; no host synchronisation (ValueNorm statistics that live on the GPU are read there).
    ``group``: a ``torch.distributed`` process group (or ``True`` for the default group) over which the
    advantage statistics (sum, sum of squares, count — three doubles) are all-reduced before the
    normalisation, so that world-sharded ranks normalise exactly like the unsharded run (r_mappo.py:174-182
    takes mean / std over the whole [T, N, 2] batch)."""
    if not value_preds.is_cuda:
        raise RuntimeError("compute_returns needs CUDA tensors; there is no CPU fallback")
    T, P, N = rewards.shape
    if tuple(value_preds.shape) != (T + 1, P, N) or tuple(dones.shape) != (T, N):
        raise ValueError("expected value_preds [T+1,P,N], rewards [T,P,N], dones [T,N]")
    if value_preds.dtype != torch.float32 or rewards.dtype != torch.int32 or dones.dtype != torch.int32:
        raise ValueError("expected float32 value_preds and int32 rewards / dones")
    for t in (value_preds, rewards, dones):
        if not t.is_contiguous():
            raise ValueError("buffers must be contiguous")
    dev = value_preds.device
    vn_dev = valuenorm_mean_std_device(value_normalizer, dev)
    mean, std = (0.0, 1.0) if vn_dev is not None else valuenorm_mean_std(value_normalizer)
    cfg = ocb_returns_cfg(ctypes.sizeof(ocb_returns_cfg), int(use_gae), gamma, gae_lambda, mean, std)
    returns = out_returns if out_returns is not None else torch.zeros((T + 1, P, N), dtype=torch.float32, device=dev)
    adv = out_advantages if out_advantages is not None else torch.empty((T, P, N), dtype=torch.float32, device=dev)
    stats = torch.empty((3,), dtype=torch.float64, device=dev)
    lib = _native.lib()
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _native.check(lib.ocb_compute_returns_dev(dev.index, ctypes.byref(cfg), T, P, N, _p(value_preds), _p(rewards),
                                                  _p(dones), _p(returns), _p(adv), _p(stats), _p(vn_dev), stream))
        if normalize and group is not None:
            import torch.distributed as dist
            dist.all_reduce(stats, group=None if group is True else group)
        if normalize:
            _native.check(lib.ocb_normalize_advantages(dev.index, _p(adv), adv.numel(), _p(stats), stream))
    return returns, adv
