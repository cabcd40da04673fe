"""CPU: the MCMC strategy kernels of the C ABI (st3r_mcmc_partition / _relocate / _compute_relocation / _inject_noise)
on the emulated library (tests/host/build_emu_lib.py), driven through starst3r_b200.gs.MCMCStrategy on CPU tensors:
the checks tests/test_mcmc_gpu.py makes on the B200 against the oracle restatement of gsplat's strategy ops."""
import pytest
import torch

CPU = torch.device("cpu")


@pytest.mark.parametrize("n", [1, 7, 1000])
def test_compute_relocation(emu_backend, n):
    import test_mcmc_gpu as t
    t.test_compute_relocation_vs_oracle(CPU, n)


@pytest.mark.parametrize("n,n_dead", [(64, 5), (3000, 150), (3000, 0)])
def test_relocate(emu_backend, n, n_dead):
    import test_mcmc_gpu as t
    t.test_relocate_vs_oracle(CPU, n, n_dead)


def test_partition_sample_add_noise(emu_backend):
    import test_mcmc_gpu as t
    t.test_partition_matches_nonzero(CPU)
    t.test_sample_add_vs_oracle(CPU, 10)
    t.test_sample_add_vs_oracle(CPU, 2000)
    t.test_cap_max_limits_growth(CPU)
    t.test_inject_noise_vs_oracle(CPU, 1)
    t.test_inject_noise_vs_oracle(CPU, 1000)
