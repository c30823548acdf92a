"""GPU parity of fgc_opt_step (--optimizer RMSprop | AdaDelta | AdaGrad) against the plain-torch operator.

The update rules themselves are checked against the oracle's restatement of TF-1's optimisers on the CPU
(tests/test_host_cpu.py)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("kind", ["rmsprop", "adadelta", "adagrad"])
def test_optimizer_step_matches_torch(kind):
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.params import ParamStore, discriminator_vars
    from torch_ops import TorchOps
    cu, ref = CudaOps("cuda:0", torch.float32), TorchOps(torch.float64, torch.device("cuda:0"))
    a, b = ParamStore(discriminator_vars(16), "cuda:0"), ParamStore(discriminator_vars(16), "cuda:0", torch.float64)
    a.initialize(5)
    b.load_state_dict(a.state_dict())
    a.set_optimizer(kind)
    b.set_optimizer(kind)
    g = torch.Generator(device="cuda").manual_seed(1)
    lr_dev = torch.tensor(3e-4, device="cuda")
    # the kernel walks the chunk table (the variables); the alignment padding between them carries no gradient
    live = torch.zeros(a.n_flat, device="cuda")
    for s0, n in zip(a.chunk_start.tolist(), a.chunk_len.tolist()):
        live[s0:s0 + n] = 1.0
    for step in range(3):
        grad = torch.randn(a.n_flat, device="cuda", generator=g) * 0.01 * live
        a.grad.copy_(grad)
        b.grad.copy_(grad.double())
        cu.optimizer_step(a, kind, 3e-4, lr_dev=lr_dev if step == 2 else None)
        ref.optimizer_step(b, kind, 3e-4)
        torch.cuda.synchronize()
        assert ((a.flat.double() - b.flat) * live).abs().max().item() <= 2e-6 * max(1.0, b.flat.abs().max().item()), (kind, step)
        assert ((a.adam_v.double() - b.adam_v) * live).abs().max().item() <= 1e-5 * max(1.0, b.adam_v.abs().max().item())
        if kind == "adadelta":
            assert ((a.opt_s2.double() - b.opt_s2) * live).abs().max().item() <= 1e-5 * max(1e-6, b.opt_s2.abs().max().item())
