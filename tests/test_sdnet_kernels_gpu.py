"""GPU: SDNet-stack kernels against the CPU oracle's building blocks on odd shapes."""
import pytest
import torch

from oracle import sdnet_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,I,H,bidir", [(5, 7, 40, 125, True), (1, 1, 300, 125, True), (33, 20, 64, 128, True),
                                           (9, 13, 50, 64, False), (130, 3, 16, 17, True), (4, 6, 30, 300, False),
                                           (256, 9, 32, 125, True), (200, 5, 24, 89, False), (600, 3, 16, 125, True)])
def test_stacked_brnn_layer_matches_oracle(B, L, I, H, bidir):
    from ruart_b200.Models import Layers
    Layers.set_dropout_prob(0.0)
    Layers.set_sdnet_precision(3)
    torch.manual_seed(B * 100 + L)
    rnn = Layers.StackedBRNN(I, H, 1, bidirectional=bidir).cuda().eval()
    x = torch.randn(B, L, I, device="cuda")
    with torch.no_grad():
        got = rnn(x, None, LN=True)
    sd = {"r." + k: v.detach().cpu() for k, v in rnn.state_dict().items()}
    want = sdnet_oracle.stacked_brnn(sd, "r", x.cpu(), 1, bidirectional=bidir, whole_ln=True)[-1]
    assert got.shape == want.shape
    assert (got.cpu() - want).abs().max().item() < 2e-4


def test_attention_module_matches_oracle():
    from ruart_b200.Models import Layers
    Layers.set_dropout_prob(0.0)
    Layers.set_sdnet_precision(3)
    torch.manual_seed(3)
    att = Layers.Attention(70, 33, correlation_func=3).cuda().eval()
    x1 = torch.randn(6, 19, 70, device="cuda")
    x2 = torch.randn(6, 11, 70, device="cuda")
    x3 = torch.randn(6, 11, 45, device="cuda")
    mask = torch.ones(6, 11, dtype=torch.uint8, device="cuda")
    mask[0, 5:] = 0
    mask[3, 1:] = 0
    with torch.no_grad():
        got = att(x1, x2, mask, x3=x3)
    sd = {"a." + k: v.detach().cpu() for k, v in att.state_dict().items()}
    want = sdnet_oracle.attention(sd, "a", x1.cpu(), x2.cpu(), mask.cpu(), x3.cpu())
    assert (got.cpu() - want).abs().max().item() < 1e-4


def test_whole_layernorm_on_strided_rows():
    from ruart_b200 import sdnet_ops as K
    buf = torch.randn(7, 11, 40, device="cuda") * 3 + 1.5
    view = buf[:, :, 5:30]
    want = view.clone()
    m = want.mean()
    v = (want - m).pow(2).mean()
    want = (want - m) / torch.sqrt(v + 1e-5)
    keep = buf.clone()
    K.whole_layernorm_(view)
    assert (view - want).abs().max().item() < 1e-5
    assert torch.equal(buf[:, :, :5], keep[:, :, :5]) and torch.equal(buf[:, :, 30:], keep[:, :, 30:])
