"""CPU: the RASTER + ADAM path of the C ABI (projection / SH, scan, tile binning - fused and the generic radix chain -,
blend forward / backward, projection backward, fused SSIM + L1 loss, fused Adam) compiled for the host by
tests/host/build_emu_lib.py - kernel launches rewritten onto the SIMT emulator - and driven through the PRODUCT's own
Python glue (starst3r_b200.gs) on CPU tensors.  The glue refuses anything but CUDA by design, so this test (and only the
test) swaps the library handle, the stream getter and torch.cuda.device for host stand-ins.  The checks are the ones
tests/test_gs_gpu.py makes on the B200 against the oracle: bin indices bit-exact, RGB / alpha, gradients vs autograd,
loss, Adam, whole training steps - for both blend kernel pairs (fragment-pool, visit-list)."""
import pytest
import torch

CPU = torch.device("cpu")


@pytest.fixture
def backend(emu_backend):
    return emu_backend


@pytest.fixture(params=[0, 1], ids=["pool-kernels", "visit-list-kernels"])
def backend_bwd(request, emu_backend, monkeypatch):
    from starst3r_b200 import gs
    monkeypatch.setattr(gs, "RASTER_VARIANT", request.param)
    yield emu_backend
    emu_backend.st3r_gs_set_raster_variant(0)


def test_forward_indices_bit_exact_and_rgb(backend):
    import test_gs_gpu as t
    t.test_rasterization_indices_bit_exact_and_rgb(CPU)
    t.test_rasterization_ragged_image_and_empty(CPU)


def test_backward_vs_autograd(backend_bwd):
    import test_gs_gpu as t
    t.test_rasterization_backward_vs_autograd(CPU)


def test_loss_and_adam(backend):
    import test_gs_gpu as t
    t.test_loss_forward_backward_vs_oracle(CPU)
    t.test_fused_adam_vs_torch(CPU)
    t.test_adam_step_dev_equals_host_form(CPU)
    t.test_adam_layouts_bit_identical(CPU)
    for world in (2, 3, 8):
        t.test_peer_gradient_kernels_on_one_device(CPU, world)


def test_train_steps_vs_oracle(backend_bwd):
    import test_gs_gpu as t
    t.test_train_steps_vs_oracle(CPU)


def test_fused_binning_equals_radix_chain(backend):
    import test_gs_gpu as t
    t.test_fused_binning_equals_radix_chain(CPU, 500, 8.0, 3, 80, 48)
