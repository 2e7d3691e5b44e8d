"""CPU tests of the device-policy export (air_hockey_training_environment_b200/policy.py): BatchNorm folding, weight
layout, and the reference-shaped networks (lib/agents/ppo/model.py:56-110, a2c/model.py:50-60, dueling/model.py:50-66)."""
import numpy as np
import pytest
import torch
from torch import nn

from air_hockey_training_environment_b200 import _capi, policy


def chain_forward(chain, x):
    h = x.astype(np.float64)
    for W, b, act in chain:
        h = h @ W.detach().double().numpy().T + b.detach().double().numpy()
        if act == _capi.AH_ACTV_RELU:
            h = np.maximum(h, 0)
        elif act == _capi.AH_ACTV_TANH:
            h = np.tanh(h)
    return h


def randomize_bn(module):
    for m in module.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.uniform_(-0.5, 0.5)
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.uniform_(-0.3, 0.3)


@pytest.mark.parametrize("make,head", [(policy.ppo_actor_discrete, "softmax"), (policy.a2c_actor, "softmax"),
                                       (policy.ppo_actor_continuous, None)])
def test_folded_chain_equals_the_module(make, head):
    torch.manual_seed(0)
    net = make()
    for m in net:
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, std=0.5)
            nn.init.normal_(m.bias, std=0.2)
    randomize_bn(net)
    chain, implied = policy.dense_chain_of(net)
    assert implied == head
    x = torch.randn(64, 8) * 100
    want = net(x).detach().double().numpy()
    got = chain_forward(chain, x.numpy())
    if head == "softmax":
        e = np.exp(got - got.max(1, keepdims=True))
        got = e / e.sum(1, keepdims=True)
    assert np.abs(got - want).max() < 1e-5


def test_batchnorm_directly_after_linear_is_folded_into_it():
    torch.manual_seed(1)
    net = nn.Sequential(nn.Linear(8, 5), nn.BatchNorm1d(5), nn.Tanh(), nn.Linear(5, 3)).eval()
    randomize_bn(net)
    chain, head = policy.dense_chain_of(net)
    assert head is None and len(chain) == 2
    x = torch.randn(32, 8)
    assert np.abs(chain_forward(chain, x.numpy()) - net(x).detach().double().numpy()).max() < 1e-5


def test_dueling_towers_as_one_dense_chain():
    torch.manual_seed(2)
    net = policy.DuelingQNet(4)
    for p in net.parameters():
        nn.init.normal_(p, std=0.4)
    chain, head = policy.dense_chain_of(net)
    assert head == "dueling" and [tuple(W.shape) for W, _, _ in chain] == [(32, 8), (16, 32), (5, 16)]
    x = torch.randn(40, 8) * 50
    y = chain_forward(chain, x.numpy())
    q = y[:, :1] + y[:, 1:] - y[:, 1:].mean(1, keepdims=True)
    assert np.abs(q - net(x).detach().double().numpy()).max() < 1e-4


def test_export_rejects_what_the_kernel_cannot_run():
    with pytest.raises(ValueError):
        policy.dense_chain_of(nn.Sequential(nn.Linear(8, 4), nn.BatchNorm1d(4)))           # training-mode BatchNorm
    with pytest.raises(TypeError):
        policy.dense_chain_of(nn.Sequential(nn.Linear(8, 4), nn.GELU()))
    with pytest.raises(ValueError):
        policy.DevicePolicy(nn.Sequential(nn.Linear(8, 64), nn.ReLU(), nn.Linear(64, 4)), "cpu")      # wider than 32
    with pytest.raises(ValueError):
        policy.DevicePolicy(nn.Sequential(nn.Linear(7, 4)), "cpu")                                    # not the observation


def test_packed_weight_layout_is_transposed_and_padded():
    torch.manual_seed(3)
    net = nn.Sequential(nn.Linear(8, 6), nn.ReLU(), nn.Linear(6, 3)).eval()
    dp = policy.DevicePolicy(net, "cpu", head="values", select="greedy")
    assert dp.n_actions == 3 and list(dp.c.width)[:3] == [8, 6, 3] and dp.c.n_layers == 2
    w = dp.weights.numpy()
    assert w.size == 8 * 8 + 8 + 6 * 4 + 4
    wt1 = w[:64].reshape(8, 8)
    assert np.allclose(wt1[:, :6], net[0].weight.detach().numpy().T) and np.all(wt1[:, 6:] == 0)
    assert np.allclose(w[64:70], net[0].bias.detach().numpy()) and np.all(w[70:72] == 0)
    wt2 = w[72:96].reshape(6, 4)
    assert np.allclose(wt2[:, :3], net[2].weight.detach().numpy().T) and np.all(wt2[:, 3] == 0)
    ptr = dp.c.weights
    with torch.no_grad():
        net[2].bias.add_(1.0)
    dp.refresh()
    assert dp.c.weights == ptr and np.allclose(dp.weights.numpy()[96:99], net[2].bias.detach().numpy())   # same storage
