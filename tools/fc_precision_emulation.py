#!/usr/bin/env python
"""Would single-product tf32 (tcgen05.mma.kind::tf32) do for the FC layers of the policy networks?  CPU emulation of the
operand roundings (no GPU needed): the conv stays exact (int8 observations, fp32 weights), FC1 / FC2 run with operands
rounded to tf32 (cvt.rna), to bf16, or split into bf16 hi + lo with the three products the kernels issue; the head is
fp32 in every variant, as in the kernels.  Printed per layout / hidden size for three networks (reference-init actor with
gain 0.01, an actor with gain 2, a critic): max |err| / max |ref|  and  the worst element-wise relative error.

    python tools/fc_precision_emulation.py > profiles/r2_fc_precision_emulation.txt
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import layouts  # noqa: E402
from diverse_conventions_b200.policy import PolicyNet  # noqa: E402
from oracle.c_oracle import COracle  # noqa: E402


def tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def forward(net, obs, mode):
    x = F.relu(F.conv2d(obs.float().movedim(-1, -3), net.conv_w, net.conv_b)).flatten(1)

    def lin(x, w, b):
        if mode == "fp32":
            return F.linear(x, w, b)
        if mode == "tf32 x1":
            return F.linear(tf32(x).double(), tf32(w).double()).float() + b
        if mode == "bf16 x1":
            return F.linear(bf16(x).double(), bf16(w).double()).float() + b
        xh, wh = bf16(x), bf16(w)
        xl, wl = bf16(x - xh), bf16(w - wh)
        return (F.linear(xh.double(), wh.double()) + F.linear(xh.double(), wl.double()) + F.linear(xl.double(), wh.double())).float() + b

    x = F.relu(lin(x, net.fc1_w, net.fc1_b))
    x = F.relu(lin(x, net.fc2_w, net.fc2_b))
    return F.linear(x, net.head_w, net.head_b)


def main():
    print("layout hidden | network | mode: max-norm relative error / worst element-wise relative error (bar of the north star: 1e-3)")
    for layout in ("simple", "random1", "unident_s"):
        lp = layouts.load_layout(layout, 400)
        orc, rng = COracle(lp, 192), np.random.default_rng(0)
        for _ in range(50):
            o, _, _ = orc.step(rng.choice(6, size=(2, 192), p=[.15, .15, .15, .15, .05, .35]))
        obs = torch.from_numpy(o.reshape(-1, lp.width, lp.height, lp.channels).copy())
        for hidden in (64, 512):
            for name, seed, gain, kind in (("actor gain 0.01", 1, 0.01, "actor"), ("actor gain 2", 3, 2.0, "actor"), ("critic", 2, 1.0, "critic")):
                net = PolicyNet(kind, lp.width, lp.height, lp.channels, hidden).init_like_reference(seed, gain=gain)
                for b in (net.conv_b, net.fc1_b, net.fc2_b, net.head_b):
                    b.uniform_(-0.1, 0.1)
                ref = forward(net, obs, "fp32").double()
                cells = []
                for mode in ("tf32 x1", "bf16 x1", "bf16 hi/lo x3"):
                    out = forward(net, obs, mode).double()
                    cells.append("%s %.1e / %.1e" % (mode, float((out - ref).abs().max() / ref.abs().max()),
                                                     float(((out - ref).abs() / (ref.abs() + 1e-12)).max())))
                print("%-9s %3d | %-15s | %s" % (layout, hidden, name, " | ".join(cells)))


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
