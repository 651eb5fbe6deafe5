"""CPU restatement of the reference's returns / GAE / advantage normalisation — TEST INFRASTRUCTURE
ONLY (imported by tests/ and the golden generator, never by the product path).

Follows, operation by operation in numpy float32:
  * SharedReplayBuffer.compute_returns, train/MAPPO/utils/shared_buffer.py:248-304 (bad_masks are all
    ones for these envs, train/MAPPO/main_player.py:274, so the proper-time-limits branches coincide
    with the plain ones);
  * ValueNorm.running_mean_var / denormalize, train/MAPPO/utils/valuenorm.py:34-41,76-87;
  * the advantage normalisation at the top of R_MAPPO.train, train/MAPPO/r_mappo.py:174-182
    (all entries active).
Pinned against the reference itself by tests/golden/returns.npz (tests/golden/make_returns_golden.py).
Layouts are the seat-major ones of the rollout buffer: value_preds [T+1,P,N], rewards [T,P,N],
done [T,N]."""
import numpy as np

f32 = np.float32


def valuenorm_mean_std(running_mean, running_mean_sq, debiasing_term, epsilon=1e-5):
    """valuenorm.py:34-41 -> (debiased mean, sqrt(clamped debiased var)) as float32 scalars"""
    d = max(f32(debiasing_term), f32(epsilon))
    mean = f32(running_mean) / d
    mean_sq = f32(running_mean_sq) / d
    var = max(f32(mean_sq - f32(mean * mean)), f32(1e-2))
    return f32(mean), f32(np.sqrt(var))


def compute_returns(value_preds, rewards, done, gamma=0.99, gae_lambda=0.95, use_gae=True, vn_mean=0.0, vn_std=1.0):
    """-> (returns [T+1,P,N] f32, advantages [T,P,N] f32 un-normalised)"""
    v = np.asarray(value_preds, dtype=f32)
    r = np.asarray(rewards).astype(f32)
    T = r.shape[0]
    masks = (f32(1.0) - np.asarray(done).astype(f32))[:, None, :]  # masks[t+1] = 1 - done[t], [T,1,N]
    g, gl = f32(gamma), f32(gamma * gae_lambda)  # python double product, then cast (tensor * python float)
    mean, std = f32(vn_mean), f32(vn_std)
    dn = v * std + mean  # valuenorm.py:83-85
    ret = np.zeros_like(v)
    ret[T] = v[T]
    if use_gae:
        gae = np.zeros_like(v[0])
        for t in reversed(range(T)):
            delta = r[t] + g * dn[t + 1] * masks[t] - dn[t]  # shared_buffer.py:283-285
            gae = delta + gl * masks[t] * gae                # :286
            ret[t] = gae + dn[t]                             # :287
    else:
        for t in reversed(range(T)):
            ret[t] = ret[t + 1] * g * masks[t] + r[t]        # :299
    return ret, ret[:-1] - dn[:-1]                           # r_mappo.py:175


def normalize_advantages(adv):
    """r_mappo.py:177-182 with every entry active: (adv - mean) / (unbiased std + 1e-5)"""
    a = np.asarray(adv, dtype=f32)
    mean = f32(a.astype(np.float64).mean())
    std = f32(a.astype(np.float64).std(ddof=1))
    return (a - mean) / (std + f32(1e-5))
