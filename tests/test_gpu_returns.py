"""GPU parity of ocb_compute_returns / ocb_normalize_advantages (csrc/ppo_kernels.cu) against the
reference's own results (tests/golden/returns.npz) and the numpy oracle on larger seeded inputs.
Bar: returns and un-normalised advantages bit-identical (fp32, same operation order); normalised
advantages within 2e-6 (the mean / std reduction order differs)."""
import os

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import returns as R
from oracle import returns_oracle as ro

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "returns.npz"))
CASES = {"gae_vn": (True, True), "gae_plain": (True, False), "disc_vn": (False, True), "disc_plain": (False, False),
         "gae_vn_ptl": (True, True)}


class _VN:
    def __init__(self, mean, std):
        self.m, self.s = mean, std

    def running_mean_var(self):
        return torch.tensor([self.m], dtype=torch.float32), torch.tensor([self.s], dtype=torch.float32) ** 2


def _dev(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)


@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_reference_golden(name):
    use_gae, vn = CASES[name]
    mean, std = ro.valuenorm_mean_std(*G[name + "_vn_state"]) if vn else (0.0, 1.0)
    T = G["rewards"].shape[0]
    cfg_vn = _VN(float(mean), float(std)) if vn else None
    # sqrt(std^2) must round-trip: pass the std through the same float32 path the wrapper uses
    ret, adv = R.compute_returns(_dev(G["value_preds"], torch.float32), _dev(G["rewards"], torch.int32),
                                 _dev(G["done"], torch.int32), 0.99, 0.95, use_gae, cfg_vn, normalize=False)
    oret, oadv = ro.compute_returns(G["value_preds"], G["rewards"], G["done"], 0.99, 0.95, use_gae,
                                    *R.valuenorm_mean_std(cfg_vn))
    assert np.array_equal(ret.cpu().numpy()[:T], oret[:T]) and np.array_equal(adv.cpu().numpy(), oadv)
    np.testing.assert_allclose(ret.cpu().numpy()[:T], G[name + "_returns"][:T], rtol=1e-6, atol=1e-6)
    ret, advn = R.compute_returns(_dev(G["value_preds"], torch.float32), _dev(G["rewards"], torch.int32),
                                  _dev(G["done"], torch.int32), 0.99, 0.95, use_gae, cfg_vn, normalize=True)
    np.testing.assert_allclose(advn.cpu().numpy(), G[name + "_adv_norm"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("T,P,N,use_gae", [(400, 2, 1000, True), (13, 2, 77, False), (1, 1, 1, True), (9, 4, 130, True)])
def test_bit_identical_to_oracle_on_seeded_inputs(T, P, N, use_gae):
    rng = np.random.default_rng(T * 1000 + N)
    v = rng.normal(0, 2.0, size=(T + 1, P, N)).astype(np.float32)
    r = rng.choice([0, 0, 0, 3, 5, 20], size=(T, 1, N)).repeat(P, axis=1).astype(np.int32)
    d = (rng.random((T, N)) < 0.05).astype(np.int32)
    mean, std = np.float32(1.7), np.float32(3.1)
    ret, adv = R.compute_returns(_dev(v, torch.float32), _dev(r, torch.int32), _dev(d, torch.int32), 0.99, 0.95, use_gae,
                                 _VN(float(mean), float(std)), normalize=False)
    oret, oadv = ro.compute_returns(v, r, d, 0.99, 0.95, use_gae, *R.valuenorm_mean_std(_VN(float(mean), float(std))))
    assert np.array_equal(ret.cpu().numpy()[:T], oret[:T])
    assert np.array_equal(adv.cpu().numpy(), oadv)
    if T * P * N > 1:
        _, advn = R.compute_returns(_dev(v, torch.float32), _dev(r, torch.int32), _dev(d, torch.int32), 0.99, 0.95,
                                    use_gae, _VN(float(mean), float(std)), normalize=True)
        np.testing.assert_allclose(advn.cpu().numpy(), ro.normalize_advantages(oadv), rtol=2e-5, atol=2e-5)


def test_full_size_buffer_properties():
    """config-4 size (T=400, 8,192 worlds, 2 seats): zero rewards and zero values give zero returns;
    rewards only -> returns are the discounted reward-to-go, checked on a strided subset of columns."""
    T, P, N = 400, 2, 8192
    z = torch.zeros((T + 1, P, N), dtype=torch.float32, device="cuda")
    rew = torch.zeros((T, P, N), dtype=torch.int32, device="cuda")
    done = torch.zeros((T, N), dtype=torch.int32, device="cuda")
    ret, adv = R.compute_returns(z, rew, done, normalize=False)
    assert not ret.any() and not adv.any()
    g = torch.Generator(device="cuda").manual_seed(3)
    rew = (torch.rand((T, 1, N), device="cuda", generator=g) < 0.01).to(torch.int32).expand(T, P, N).contiguous() * 20
    done[199] = 1
    done[399] = 1
    ret, adv = R.compute_returns(z, rew, done, 0.99, 1.0, True, None, normalize=False)
    cols = slice(0, N, 257)
    oret, _ = ro.compute_returns(np.zeros((T + 1, P, len(range(N)[cols])), np.float32), rew[:, :, cols].cpu().numpy(),
                                 done[:, cols].cpu().numpy(), 0.99, 1.0, True)
    assert np.array_equal(ret[:T, :, cols].cpu().numpy(), oret[:T])
    assert torch.equal(ret[:T, 0], ret[:T, 1])


def test_device_resident_valuenorm_statistics_give_the_same_bits_without_a_host_sync():
    """a ValueNorm whose statistics are CUDA tensors is read on the device (ocb_compute_returns_dev): identical
    returns / advantages to the host-float path, and capturable into a CUDA graph"""
    T, P, N = 50, 2, 300
    rng = np.random.default_rng(5)
    v = _dev(rng.normal(0, 2.0, size=(T + 1, P, N)).astype(np.float32), torch.float32)
    r = _dev(rng.choice([0, 0, 3, 20], size=(T, 1, N)).repeat(P, axis=1).astype(np.int32), torch.int32)
    d = _dev((rng.random((T, N)) < 0.05).astype(np.int32), torch.int32)

    class DevVN(_VN):  # statistics live on the device, like ppo.ValueNormState or a reference ValueNorm moved to the GPU
        def __init__(self, mean, std):
            super().__init__(mean, std)
            m, var = _VN.running_mean_var(self)
            self.dm, self.dvar = m.cuda(), var.cuda()

        def running_mean_var(self):
            return self.dm, self.dvar

    host = R.compute_returns(v, r, d, 0.99, 0.95, True, _VN(1.25, 2.5), normalize=False)
    dev = R.compute_returns(v, r, d, 0.99, 0.95, True, DevVN(1.25, 2.5), normalize=False)
    assert torch.equal(host[0][:T], dev[0][:T]) and torch.equal(host[1], dev[1])
    vn = DevVN(1.25, 2.5)
    out_r, out_a = torch.zeros_like(host[0]), torch.zeros_like(host[1])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        R.compute_returns(v, r, d, 0.99, 0.95, True, vn, False, out_r, out_a)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_r[:T], host[0][:T]) and torch.equal(out_a, host[1])
