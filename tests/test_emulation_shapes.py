"""
Architectures and libraries beyond the golden fixtures, checked by running the exact kernel source in emulation
(tests/emu.py) against the float64 oracle: one hidden layer, narrow/wide/ragged widths (padding to a multiple of 8, up to
the 64 limit), J = 1 (value only), K = 0 (no RHS term), an all-masked library, high powers, d = 3 with mixed derivatives.
"""
import zlib

import numpy as np
import pytest
import torch

from oracle import pde_oracle as O
from pde_learn_b200 import _lib as L
from tests import emu
from tests.helpers import rel_max

TOL = 1e-5
from tests.fixtures_meta import CASES            # plain data shared with __graft_entry__.build() (plan pre-compilation)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_kernel_vs_oracle(case):
    name, d, hidden, act, lhs, rhs, n, masked = case
    widths = [d] + hidden + [1]
    net32 = O.init_network(widths, act, seed=zlib.crc32(name.encode()) % 1000)
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for b in net32.biases:
            b.copy_(0.1 * torch.randn(b.shape, generator=gen))
        for a in net32.rat_a:
            a.add_(0.03 * torch.randn(4, generator=gen))
    K = len(rhs)
    xi = (0.3 * torch.randn(max(K, 1), generator=gen))[:K].to(torch.float32)
    mask = [k in masked for k in range(K)]
    pts = (torch.rand(n, d, generator=gen) * 2 - 1).to(torch.float32)
    encs = O.distinct_encodings(lhs, rhs)
    idx = {e: i for i, e in enumerate(encs)}
    terms = [[(idx[tuple(e)], p) for e, p in t] for t in [lhs] + rhs]
    desc, keep = L.make_desc(d, widths, L.PDL_ACT_RATIONAL if act == "rational" else L.PDL_ACT_TANH, 0, encs, terms)
    out = emu.run_emu(desc, pts.numpy(), [p.numpy() for p in net32.tensors()], xi.numpy(), np.array(mask, dtype=np.uint8), n_ctas=2)
    net = net32.to(torch.float64, True)
    xi64 = xi.double().requires_grad_(True)
    loss, resid, _ = O.coll_loss(lambda X: O.mlp_forward(net, X), xi64, mask, pts.double(), lhs, rhs)
    grads = torch.autograd.grad(loss, net.tensors() + [xi64], allow_unused=True)
    assert abs(out["loss"] - float(loss.detach())) <= TOL * abs(float(loss.detach()))
    assert rel_max(out["residual"], resid.detach().numpy()) <= TOL
    ref = np.concatenate([g.numpy().ravel() if g is not None else np.zeros(t.numel()) for g, t in zip(grads[:-1], net.tensors())])
    assert not np.isnan(out["grad_params"]).any()
    assert rel_max(out["grad_params"], ref) <= TOL
    if K:
        gxi = grads[-1].numpy() if grads[-1] is not None else np.zeros(K)
        assert np.max(np.abs(out["grad_xi"] - gxi)) <= TOL * max(1e-30, np.max(np.abs(gxi))) or np.max(np.abs(gxi)) == 0
        for k in masked:
            assert out["grad_xi"][k] == 0.0
