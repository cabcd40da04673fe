"""Known-answer tests of the MCMC-strategy oracle (oracle/gs_oracle.py mcmc_*; gsplat.MCMCStrategy restated,
reference call sites starster/gs.py:43-45,146-147,163-164).  Parity unpinned (gsplat is not installable here): these
closed forms are what pins the restatement; the CUDA kernels are compared with it in tests/test_mcmc_gpu.py."""
import math

import torch

from oracle import gs_oracle as go


def test_binoms_table():
    b = go.mcmc_binoms()
    assert b.shape == (51, 51)
    assert b[0, 0] == 1 and b[5, 2] == 10 and b[10, 5] == 252 and b[3, 4] == 0
    assert b[50, 25] == float(torch.tensor(float(math.comb(50, 25)), dtype=torch.float32))


def test_relocation_ratio_one_is_identity():
    o = torch.tensor([0.05, 0.3, 0.9])
    s = torch.rand(3, 3) + 0.1
    no, ns = go.mcmc_compute_relocation(o, s, torch.tensor([1, 1, 1]), go.mcmc_binoms())
    assert torch.allclose(no, o, atol=1e-7)
    assert torch.allclose(ns, s, rtol=1e-6)


def test_relocation_ratio_two_closed_form():
    # n = 2: o' = 1 - sqrt(1 - o); denominator = C(0,0) o' + [C(1,0) o' - C(1,1) o'^2 / sqrt 2]
    o = torch.tensor([0.2, 0.5, 0.8], dtype=torch.float64)
    s = torch.ones(3, 3)
    no, ns = go.mcmc_compute_relocation(o.float(), s, torch.tensor([2, 2, 2]), go.mcmc_binoms())
    op = 1 - torch.sqrt(1 - o)
    den = 2 * op - op ** 2 / math.sqrt(2)
    assert torch.allclose(no.double(), op, atol=1e-6)
    assert torch.allclose(ns[:, 0].double(), o / den, rtol=1e-5)
    # splitting conserves total opacity: 1 - (1 - o')^2 = o
    assert torch.allclose(1 - (1 - no.double()) ** 2, o, atol=1e-6)


def test_relocation_ratio_clamped_to_table():
    o = torch.tensor([0.5])
    s = torch.ones(1, 3)
    a = go.mcmc_compute_relocation(o, s, torch.tensor([51]), go.mcmc_binoms())
    b = go.mcmc_compute_relocation(o, s, torch.tensor([500]), go.mcmc_binoms())
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert 0 < float(a[1][0, 0]) < 1      # many-way split shrinks the scale


def _params(n, seed=0, n_dead=3):
    g = torch.Generator().manual_seed(seed)
    p = {"means": torch.randn(n, 3, generator=g), "scales": torch.randn(n, 3, generator=g) * 0.3 - 3.0,
         "quats": torch.randn(n, 4, generator=g), "opacities": torch.randn(n, generator=g),
         "sh0": torch.randn(n, 1, 3, generator=g), "shN": torch.randn(n, 24, 3, generator=g)}
    dead = torch.randperm(n, generator=g)[:n_dead]
    p["opacities"][dead] = -8.0          # sigmoid(-8) = 3.4e-4 <= 0.005
    m = {k: (torch.rand(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in p.items()}
    return p, m, dead.sort().values


def test_relocate_semantics():
    p, m, dead = _params(40, n_dead=4)
    before = {k: v.clone() for k, v in p.items()}
    sampled = torch.tensor([0, 5, 5, 17])           # indices into the alive list, one duplicate
    dead_out, src = go.mcmc_relocate(p, m, sampled, go.mcmc_binoms())
    assert torch.equal(dead_out, dead)
    alive = torch.tensor([i for i in range(40) if i not in dead.tolist()])
    assert torch.equal(src, alive[sampled])
    # dead rows are copies of their (updated) sources; untouched rows are unchanged
    for k in p:
        assert torch.equal(p[k][dead], p[k][src])
    untouched = [i for i in range(40) if i not in dead.tolist() and i not in src.tolist()]
    for k in p:
        assert torch.equal(p[k][untouched], before[k][untouched])
    # the duplicated source was split three ways (2 draws + itself), the others two ways
    o_old = torch.sigmoid(before["opacities"][src])
    o_new = torch.sigmoid(p["opacities"][src])
    ways = torch.tensor([2.0, 3.0, 3.0, 2.0])
    expect = (1 - (1 - o_old) ** (1 / ways)).clamp(min=0.005)
    assert torch.allclose(o_new, expect, atol=1e-6)
    # only the sources' Adam moments are reset
    for k, (ea, es) in m.items():
        assert ea[src].abs().max() == 0 and es[src].abs().max() == 0
        assert (ea[dead] != 0).any() and (ea[untouched] != 0).any()


def test_sample_add_semantics():
    p, m, _ = _params(20, n_dead=0)
    sampled = torch.tensor([3, 3, 7])
    out_p, out_m = go.mcmc_sample_add(p, m, sampled, go.mcmc_binoms())
    for k in p:
        assert out_p[k].shape[0] == 23
        assert torch.equal(out_p[k][20:], out_p[k][sampled])
    assert torch.equal(out_p["means"][:20], p["means"])
    for k, (ea, es) in out_m.items():
        assert torch.equal(ea[:20], m[k][0]) and ea[20:].abs().max() == 0 and es[20:].abs().max() == 0


def test_inject_noise_gate_and_covariance():
    p, _, _ = _params(16, n_dead=0)
    p["quats"][:] = torch.tensor([1.0, 0, 0, 0])
    p["scales"][:] = torch.log(torch.tensor([1.0, 2.0, 3.0]))
    p["opacities"][:8] = 8.0            # opaque: gate = op_sigmoid(1 - 0.9997) ~ e^-99.5 ~ 0
    p["opacities"][8:] = -8.0           # transparent: gate = op_sigmoid(0.99966) ~ 0.61
    noise = torch.ones(16, 3)
    out = go.mcmc_inject_noise(p, noise, scaler=0.5)
    d = out - p["means"]
    assert d[:8].abs().max() < 1e-30
    o = torch.sigmoid(torch.tensor(-8.0))
    gate = 1 / (1 + torch.exp(-100 * ((1 - o) - 0.995)))
    assert torch.allclose(d[8:], (gate * 0.5) * torch.tensor([1.0, 4.0, 9.0]).expand(8, 3), rtol=1e-5)
