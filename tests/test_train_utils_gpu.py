"""GPU: the fused clip + Adamax step (ruart_grad_sqnorm / ruart_adamax_step, ruart_b200/train_utils.py)
against torch.nn.utils.clip_grad_norm_ + torch.optim.Adamax, the pair SDNetTrainer.update uses
(Models/SDNetTrainer.py:313,363-365), including the TUNE_PARTIAL row reset (:367-371)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("max_norm", [10.0, 0.05, None])
def test_flat_adamax_matches_torch(max_norm):
    from ruart_b200.train_utils import FlatAdamax
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(1000, 300), (1200, 1388), (12,), (1, 1), (250, 801), (7, 3, 5)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    fixed = ours[0].detach()[600:].clone()
    opt_ref = torch.optim.Adamax(ref, lr=1e-3)
    opt = FlatAdamax(ours, lr=1e-3, max_norm=max_norm)
    for p, q in zip(ours, ref):
        assert torch.equal(p.detach(), q.detach())               # flattening preserved the values
    for step in range(4):
        grads = [torch.randn(s, device="cuda", generator=g) * (0.1 if step % 2 else 3.0) for s in shapes]
        for q, gr in zip(ref, grads):
            q.grad = gr.clone()
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt_ref.step()
        with torch.no_grad():
            ref[0].data[600:] = fixed
        opt.step(grads, reset=[(ours[0], 600, fixed)])
        if max_norm is not None:
            assert abs(opt.grad_norm() - float(total)) < 1e-4 * float(total)
        for p, q in zip(ours, ref):
            assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-7), (step, tuple(p.shape))
    assert torch.equal(ours[0].detach()[600:], fixed)


def test_grad_sqnorm_is_deterministic_and_handles_odd_sizes():
    from ruart_b200._lib import current_stream, ptr
    from ruart_b200.ops import call
    for n in (1, 3, 1027, 12246007):
        g = torch.randn(n + 1, device="cuda")[1:]                   # misaligned start
        ws = torch.empty(1025, dtype=torch.float64, device="cuda")
        outs = []
        for _ in range(2):
            call("ruart_grad_sqnorm", ptr(g), n, ptr(ws), ws.data_ptr() + 8 * 1024, current_stream())
            outs.append(float(ws[1024]))
        assert outs[0] == outs[1]
        want = float((g.double() ** 2).sum())
        assert abs(outs[0] - want) < 1e-10 * max(1.0, want)
