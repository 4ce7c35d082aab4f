"""`parallel.DataParallel` — the single-process scatter / replicate / gather strategy the reference's README describes
next to DDP (ref: README.md:10-17).  CPU "devices" stand in for GPUs: the mechanics (chunking the *global* batch,
one replica per chunk on its own thread, gather, gradients flowing back to the one set of parameters) are the same."""
import pytest
import torch
import torch.nn as nn

import pytorch_distributed_train_b200 as pdt


def _mlp():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(12, 16), nn.ReLU(), nn.Linear(16, 5))


def test_forward_and_gradients_match_a_single_replica():
    ref, net = _mlp(), _mlp()
    dp = pdt.DataParallel(net, device_ids=["cpu", "cpu", "cpu"])
    x = torch.randn(10, 12)          # 10 = 4 + 4 + 2: uneven last chunk, like torch's scatter
    y = torch.randint(0, 5, (10,))
    out = dp(x)
    assert out.shape == (10, 5) and torch.allclose(out, ref(x), atol=1e-6)
    nn.functional.cross_entropy(out, y).backward()
    nn.functional.cross_entropy(ref(x), y).backward()
    for a, b in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(a.grad, b.grad, atol=1e-6)      # the global batch is split, the gradient is not


def test_structured_inputs_kwargs_and_module_attribute():
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(4, 3)

        def forward(self, pair, scale=None):
            a, b = pair
            out = self.fc(a + b)
            return {"logits": out * scale, "norm": out.norm(dim=1)}

    net = Net()
    dp = pdt.DataParallel(net, device_ids=["cpu", "cpu"])
    assert dp.module is net
    a, b, s = torch.randn(6, 4), torch.randn(6, 4), torch.rand(6, 1)
    got, want = dp((a, b), scale=s), net((a, b), scale=s)
    assert torch.allclose(got["logits"], want["logits"], atol=1e-6) and torch.allclose(got["norm"], want["norm"], atol=1e-6)


def test_single_device_is_a_passthrough_and_errors_propagate():
    net = _mlp()
    assert torch.equal(pdt.DataParallel(net, device_ids=["cpu"])(torch.ones(2, 12)), net(torch.ones(2, 12)))
    with pytest.raises(RuntimeError):
        pdt.DataParallel(net, device_ids=["cpu", "cpu"])(torch.ones(4, 7))   # shape error raised inside a replica thread
