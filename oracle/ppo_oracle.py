"""CPU restatement of the reference's PPO minibatch path — TEST INFRASTRUCTURE ONLY (imported by tests/
and the golden generator, never by the product path).

Follows, in numpy float32:
  * SharedReplayBuffer.feed_forward_generator's flattening, train/MAPPO/utils/shared_buffer.py:328-342
    ([T,N,P] row-major) against the seat-major [T,P,N] rollout buffer;
  * FixedCategorical.log_probs / entropy of the Discrete head, train/MAPPO/utils/distributions.py:14-28 and
    ACTLayer.evaluate_actions, train/MAPPO/utils/act.py:164-175;
  * R_MAPPO.cal_value_loss and the surrogate of ppo_update, train/MAPPO/r_mappo.py:52-127, with
    huber_loss / mse_loss of train/MAPPO/utils/util.py:46-53 and ValueNorm.update / normalize of
    train/MAPPO/utils/valuenorm.py:34-74;
  * the gradients torch.autograd produces for those expressions (ties of torch.min / torch.max split evenly,
    clamp passes the gradient on its closed interval).
Pinned against the reference itself by tests/golden/ppo.npz (tests/golden/make_ppo_golden.py)."""
import numpy as np

f32 = np.float32


def flat_to_rows(idx, N, P):
    """reference flat sample index (t*N + n)*P + p -> seat-major agent row (t*P + p)*N + n"""
    idx = np.asarray(idx, dtype=np.int64)
    p, tn = idx % P, idx // P
    n, t = tn % N, tn // N
    return ((t * P + p) * N + n).astype(np.int32)


def evaluate_head(logits, actions):
    """-> (log pi(a) [B], entropy [B]) of Categorical(logits=logits)"""
    x = np.asarray(logits, dtype=f32)
    mx = x.max(-1, keepdims=True)
    e = np.exp(x - mx)
    s = e.sum(-1, keepdims=True)
    logp_all = (x - mx) - np.log(s)
    p = e / s
    a = np.asarray(actions).astype(np.int64).reshape(-1)
    return logp_all[np.arange(len(a)), a].astype(f32), (-(p * logp_all).sum(-1)).astype(f32)


def valuenorm_update(state, batch, beta=0.99999):
    """ValueNorm.update (valuenorm.py:43-60) on state = [running_mean, running_mean_sq, debiasing_term]"""
    b = np.asarray(batch, dtype=f32)
    w, omw = f32(beta), f32(1.0 - beta)
    m1, m2 = f32(b.mean(dtype=np.float64)), f32((b * b).mean(dtype=np.float64))
    s = np.asarray(state, dtype=f32).copy()
    s[0] = f32(s[0] * w) + f32(m1 * omw)
    s[1] = f32(s[1] * w) + f32(m2 * omw)
    s[2] = f32(s[2] * w) + f32(f32(1.0) * omw)
    return s


def valuenorm_mean_std(state, epsilon=1e-5):
    d = max(f32(state[2]), f32(epsilon))
    mean, mean_sq = f32(state[0]) / d, f32(state[1]) / d
    return f32(mean), f32(np.sqrt(max(f32(mean_sq - f32(mean * mean)), f32(1e-2))))


def _err_loss(e, d, huber):
    if not huber:
        return (e * e / f32(2)).astype(f32), e
    a = (np.abs(e) <= d).astype(f32)
    b = (e > d).astype(f32)  # util.py:46-49: errors below -d contribute nothing
    return (a * e * e / f32(2) + b * d * (np.abs(e) - d / f32(2))).astype(f32), (a * e + b * d).astype(f32)


def ppo_loss(logp_new, entropy, values_new, old_logp, adv, value_preds, returns, active=None, vn_state=None, clip=0.2,
             delta=10.0, use_clipped_value_loss=True, use_huber_loss=True, use_value_active_masks=True,
             use_policy_active_masks=True, beta=0.99999, epsilon=1e-5):
    """Dense [B] inputs -> dict(policy_loss, value_loss, dist_entropy, ratio_mean, imp_weights, dlogp, dvalues, vn_state)."""
    lp, olp, adv = (np.asarray(t, dtype=f32).reshape(-1) for t in (logp_new, old_logp, adv))
    v, vp, ret = (np.asarray(t, dtype=f32).reshape(-1) for t in (values_new, value_preds, returns))
    B = lp.size
    act = np.ones(B, dtype=f32) if active is None else np.asarray(active, dtype=f32).reshape(-1)
    clip, delta = f32(clip), f32(delta)
    pmask, vmask = use_policy_active_masks and active is not None, use_value_active_masks and active is not None
    pw, vw = (act if pmask else np.ones(B, f32)), (act if vmask else np.ones(B, f32))
    pden, vden = np.float64(pw.sum(dtype=np.float64)), np.float64(vw.sum(dtype=np.float64))
    # ---- actor
    ratio = np.exp(lp - olp).astype(f32)
    surr1 = ratio * adv
    lo, hi = f32(1.0) - clip, f32(1.0) + clip
    inside = (ratio >= lo) & (ratio <= hi)
    surr2 = np.clip(ratio, lo, hi) * adv
    m = np.minimum(surr1, surr2)
    policy_loss = -(m * pw).sum(dtype=np.float64) / pden
    g = np.where(surr1 < surr2, surr1, np.where(surr1 == surr2, f32(0.5) * surr1 + np.where(inside, f32(0.5) * surr1, f32(0)), f32(0)))
    dlogp = (-g * pw / f32(pden)).astype(f32)
    ent = None if entropy is None else (np.asarray(entropy, dtype=f32).reshape(-1) * pw).sum(dtype=np.float64) / pden
    # ---- critic
    out_state = None
    if vn_state is not None:
        out_state = valuenorm_update(vn_state, ret, beta)
        mean, std = valuenorm_mean_std(out_state, epsilon)
        ret = ((ret - mean) / std).astype(f32)
    dv = v - vp
    vin = (dv >= -clip) & (dv <= clip)
    vpc = vp + np.clip(dv, -clip, clip)
    l_c, de_c = _err_loss(ret - vpc, delta, use_huber_loss)
    l_o, de_o = _err_loss(ret - v, delta, use_huber_loss)
    l, gv = l_o, -de_o
    if use_clipped_value_loss:
        g_c = np.where(vin, -de_c, f32(0))
        gv = np.where(l_c > l_o, g_c, np.where(l_c == l_o, f32(0.5) * gv + f32(0.5) * g_c, gv))
        l = np.maximum(l_o, l_c)
    value_loss = (l * vw).sum(dtype=np.float64) / vden
    dvalues = (gv * vw / f32(vden)).astype(f32)
    return {"policy_loss": policy_loss, "value_loss": value_loss, "dist_entropy": ent, "ratio_mean": ratio.mean(dtype=np.float64),
            "imp_weights": ratio, "dlogp": dlogp, "dvalues": dvalues, "vn_state": out_state}
