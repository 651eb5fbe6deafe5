"""PPO minibatch path on the device: sampler, gather, evaluate_actions and loss over the seat-major
rollout buffer (``ocb_minibatch_gather`` / ``ocb_policy_evaluate`` / ``ocb_ppo_loss``).

Replaces the data side of the reference's update step:

* ``SharedReplayBuffer.feed_forward_generator`` (train/MAPPO/utils/shared_buffer.py:306-366):
  ``torch.randperm`` over the ``T*N*P`` samples, then fancy-indexing of twelve flattened tensors
  (two fp32 copies of the observations).  Here the permutation is converted once to agent-row indices
  of the seat-major buffer and either consumed in place by the tensor-core forward or gathered by one
  launch into dense minibatch tensors (int8 or the reference's fp32).
* ``R_MAPPO_Policy.evaluate_actions`` -> ``R_Actor.evaluate_actions`` / ``R_Critic.forward``
  (train/MAPPO/r_actor_critic.py:73-109,178-197) and the loss half of ``R_MAPPO.ppo_update`` /
  ``cal_value_loss`` (train/MAPPO/r_mappo.py:52-127), forward plus the analytic gradients with
  respect to the new log-probs and values.  Parameter gradients / the optimiser step stay with the
  trainer (out of scope).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterator, Optional, Sequence, Tuple

import torch

from . import _native
from .policy import FusedPolicy
from .rollout import RolloutBuffer

PPO_STATS = 16


class ocb_ppo_cfg(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("use_clipped_value_loss", ctypes.c_int32),
                ("use_huber_loss", ctypes.c_int32), ("use_valuenorm", ctypes.c_int32),
                ("use_value_active_masks", ctypes.c_int32), ("use_policy_active_masks", ctypes.c_int32),
                ("clip_param", ctypes.c_float), ("huber_delta", ctypes.c_float), ("vn_beta", ctypes.c_double),
                ("vn_epsilon", ctypes.c_double)]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def reference_flat_to_rows(idx: torch.Tensor, N: int, P: int) -> torch.Tensor:
    """The reference flattens ``[T, N, P]`` (shared_buffer.py:330-342: flat = (t*N + n)*P + p); the
    seat-major buffer's agent row is ``(t*P + p)*N + n``.  int64/int32 in, int32 out."""
    idx = idx.to(torch.int64)
    p = idx % P
    tn = idx // P
    n = tn % N
    t = tn // N
    return ((t * P + p) * N + n).to(torch.int32)


def minibatch_rows(T: int, N: int, P: int, num_mini_batch: int, generator: Optional[torch.Generator] = None,
                   device=None) -> Sequence[torch.Tensor]:
    """feed_forward_generator's sampler (shared_buffer.py:313-326): one CPU ``torch.randperm`` over
    ``T*N*P`` — the same permutation the reference draws under the same torch seed — split into
    ``num_mini_batch`` index lists, returned as seat-major agent rows (int32, on ``device``)."""
    batch_size = T * N * P
    if batch_size < num_mini_batch:
        raise ValueError("PPO requires T*N*P = %d >= the number of mini batches %d" % (batch_size, num_mini_batch))
    mbs = batch_size // num_mini_batch
    rand = torch.randperm(batch_size, generator=generator)
    out = []
    for i in range(num_mini_batch):
        rows = reference_flat_to_rows(rand[i * mbs:(i + 1) * mbs], N, P)
        out.append(rows.to(device, non_blocking=True) if device is not None else rows)
    return out


def gather_minibatch(buf: RolloutBuffer, rows: torch.Tensor, advantages: torch.Tensor, returns: torch.Tensor,
                     obs_dtype: torch.dtype = torch.int8) -> Dict[str, torch.Tensor]:
    """One launch: the dense minibatch the reference's generator yields (shared_buffer.py:339-366), with its names.
    ``obs_batch`` is int8 by default (the reference's consumers cast with ``.to(float32)``) or fp32;
    ``share_obs_batch`` is the same tensor (state == obs for Overcooked)."""
    dev = buf.obs.device
    if not buf.obs.is_cuda:
        raise RuntimeError("gather_minibatch needs CUDA tensors; there is no CPU fallback")
    if obs_dtype not in (torch.int8, torch.float32):
        raise ValueError("obs_dtype must be int8 or float32")
    rows = rows.to(dev).contiguous()
    assert rows.dtype == torch.int32
    B = rows.numel()
    W, H, C = buf.obs.shape[-3:]
    obs_out = torch.empty((B, W, H, C), dtype=obs_dtype, device=dev)
    f32_src = [buf.value_preds, returns, buf.action_log_probs, advantages]
    names = ["value_preds_batch", "return_batch", "old_action_log_probs_batch", "adv_targ"]
    f32_out = [torch.empty((B, 1), dtype=torch.float32, device=dev) for _ in f32_src]
    actions_out = torch.empty((B, 1), dtype=torch.int32, device=dev)
    FA, IA = ctypes.c_void_p * len(f32_src), ctypes.c_void_p * 1
    lib = _native.lib()
    with torch.cuda.device(dev):
        _native.check(lib.ocb_minibatch_gather(
            dev.index, _p(rows), B, W * H * C, _p(buf.obs), _p(obs_out), int(obs_dtype == torch.float32), len(f32_src),
            FA(*[t.data_ptr() for t in f32_src]), FA(*[t.data_ptr() for t in f32_out]), 1, IA(buf.actions.data_ptr()),
            IA(actions_out.data_ptr()), _stream(dev)))
    batch = dict(zip(names, f32_out))
    one = torch.ones((1, 1), dtype=torch.float32, device=dev)  # all agents active, all actions available: broadcast views
    batch.update(obs_batch=obs_out, share_obs_batch=obs_out, actions_batch=actions_out,
                 masks_batch=None, active_masks_batch=one.expand(B, 1), available_actions_batch=one.expand(B, 6))
    return batch


class ValueNormState:
    """Device-resident state of the reference's ``ValueNorm(1)`` (train/MAPPO/utils/valuenorm.py:9-41):
    float32 ``[running_mean, running_mean_sq, debiasing_term]``, updated by ``ppo_loss`` on the device."""

    def __init__(self, device, beta: float = 0.99999, epsilon: float = 1e-5):
        self.state = torch.zeros(3, dtype=torch.float32, device=device)
        self.beta, self.epsilon = beta, epsilon

    @classmethod
    def from_reference(cls, vn, device) -> "ValueNormState":
        s = cls(device, float(vn.beta), float(vn.epsilon))
        s.state.copy_(torch.stack([vn.running_mean.reshape(()), vn.running_mean_sq.reshape(()),
                                   vn.debiasing_term.reshape(())]).to(torch.float32))
        return s

    def running_mean_var(self) -> Tuple[torch.Tensor, torch.Tensor]:
        deb = self.state[2].clamp(min=self.epsilon)
        mean, mean_sq = self.state[0] / deb, self.state[1] / deb
        return mean.reshape(1), (mean_sq - mean ** 2).clamp(min=1e-2).reshape(1)


def ppo_loss(rows: Optional[torch.Tensor], logp_new: torch.Tensor, entropy: Optional[torch.Tensor], values_new: torch.Tensor,
             old_logp_src: torch.Tensor, adv_src: torch.Tensor, value_preds_src: torch.Tensor, returns_src: torch.Tensor,
             active_src: Optional[torch.Tensor] = None, value_norm: Optional[ValueNormState] = None,
             clip_param: float = 0.2, huber_delta: float = 10.0, use_clipped_value_loss: bool = True,
             use_huber_loss: bool = True, use_value_active_masks: bool = True, use_policy_active_masks: bool = True,
             want_grads: bool = True) -> Dict[str, torch.Tensor]:
    """The loss half of ``R_MAPPO.ppo_update`` (r_mappo.py:110-127) + ``cal_value_loss`` (52-89) on the device.
    ``*_src`` tensors are whole buffers addressed through ``rows`` (None = already dense).  Returns device
    tensors: ``policy_loss, value_loss, dist_entropy, ratio_mean`` (0-d, float64), ``imp_weights [B]`` and, with
    ``want_grads``, ``dlogp = d policy_loss / d logp_new`` and ``dvalues = d value_loss / d values_new``.
    ``value_norm`` is updated in place with the batch returns first, like the reference."""
    dev = logp_new.device
    if not logp_new.is_cuda:
        raise RuntimeError("ppo_loss needs CUDA tensors; there is no CPU fallback")
    B = logp_new.numel()
    for t in (logp_new, entropy, values_new, old_logp_src, adv_src, value_preds_src, returns_src, active_src):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("expected contiguous float32 tensors")
    if values_new.numel() != B or (entropy is not None and entropy.numel() != B):
        raise ValueError("logp_new, entropy and values_new must have one entry per minibatch row")
    if rows is not None:
        if rows.dtype != torch.int32 or rows.numel() != B or not rows.is_contiguous():
            raise ValueError("rows must be contiguous int32 [B]")
    cfg = ocb_ppo_cfg(ctypes.sizeof(ocb_ppo_cfg), int(use_clipped_value_loss), int(use_huber_loss),
                      int(value_norm is not None), int(use_value_active_masks), int(use_policy_active_masks),
                      clip_param, huber_delta, value_norm.beta if value_norm is not None else 0.99999,
                      value_norm.epsilon if value_norm is not None else 1e-5)
    f = lambda: torch.empty((B,), dtype=torch.float32, device=dev)
    imp, dlogp, dvalues = f(), (f() if want_grads else None), (f() if want_grads else None)
    stats = torch.empty((PPO_STATS,), dtype=torch.float64, device=dev)
    lib = _native.lib()
    with torch.cuda.device(dev):
        _native.check(lib.ocb_ppo_loss(
            dev.index, ctypes.byref(cfg), B, _p(rows), _p(logp_new), _p(entropy), _p(values_new), _p(old_logp_src),
            _p(adv_src), _p(value_preds_src), _p(returns_src), _p(active_src),
            _p(value_norm.state) if value_norm is not None else None, _p(imp), _p(dlogp), _p(dvalues), _p(stats),
            _stream(dev)))
    return {"policy_loss": stats[0], "value_loss": stats[1], "dist_entropy": stats[2], "ratio_mean": stats[3],
            "imp_weights": imp, "dlogp": dlogp, "dvalues": dvalues, "stats": stats}


class PPOMinibatches:
    """``feed_forward_generator`` + ``evaluate_actions`` + loss over one collected rollout, in place.

    ``for mb in PPOMinibatches(buf, policy, ...)`` yields, per minibatch, the dict of ``ppo_loss`` plus
    ``rows, logp, entropy, values`` — three launches per minibatch (forward, stats, loss), no observation
    copy and no host synchronisation."""

    def __init__(self, buf: RolloutBuffer, policy: FusedPolicy, num_mini_batch: int = 1,
                 value_norm: Optional[ValueNormState] = None, generator: Optional[torch.Generator] = None,
                 policy_index: int = 0, **loss_kwargs):
        if getattr(buf, "returns", None) is None or getattr(buf, "advantages", None) is None:
            raise ValueError("call buf.compute_returns() first")
        self.buf, self.policy, self.value_norm, self.loss_kwargs = buf, policy, value_norm, loss_kwargs
        self.rows = minibatch_rows(buf.T, buf.N, buf.P, num_mini_batch, generator, buf.obs.device)
        B = self.rows[0].numel()
        self.tile_policy = None if policy_index == 0 else torch.full(
            ((B + FusedPolicy.TILE - 1) // FusedPolicy.TILE,), policy_index, dtype=torch.int32, device=buf.obs.device)

    def __len__(self):
        return len(self.rows)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        b = self.buf
        for rows in self.rows:
            ev = self.policy.evaluate(b.obs, b.actions, rows, self.tile_policy)
            out = ppo_loss(rows, ev["logp"], ev["entropy"], ev["values"], b.action_log_probs, b.advantages,
                           b.value_preds, b.returns, None, self.value_norm, **self.loss_kwargs)
            out.update(rows=rows, logp=ev["logp"], entropy=ev["entropy"], values=ev["values"])
            yield out
