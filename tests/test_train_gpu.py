"""NNet::train (src/nnet.rs:38) on the device against torch autograd (float64, same parameters): loss, every gradient
tensor, the Adam step, and that training lowers the loss.  The device runs the tower in bf16 (fp32 accumulation), so the
gradients are compared by direction and relative norm, not bit for bit; the tolerances are written below."""
import numpy as np
import pytest

from test_nnet_gpu import random_features

pytestmark = pytest.mark.gpu


def torch_loss_and_grads(azb, params, blocks, boards, pis, vs):
    import torch
    import torch.nn.functional as F
    L = {k: v for k, v in azb.param_layout(blocks).items() if k != "total"}
    total = azb.param_layout(blocks)["total"]
    leaf = {}
    for name, (o, shape) in L.items():
        leaf[name] = torch.from_numpy(params[o:o + int(np.prod(shape))].reshape(shape).copy()).double().requires_grad_(True)
    x = torch.from_numpy(boards).double()
    x = F.relu(F.conv2d(x, leaf["stem_w"].reshape(3, 3, 2, 128).permute(3, 2, 0, 1), leaf["stem_b"], padding=1))
    for b in range(blocks):
        w1 = leaf["tower_w"][2 * b].reshape(3, 3, 128, 128).permute(3, 2, 0, 1)
        w2 = leaf["tower_w"][2 * b + 1].reshape(3, 3, 128, 128).permute(3, 2, 0, 1)
        y = F.relu(F.conv2d(x, w1, leaf["tower_b"][2 * b], padding=1))
        x = F.relu(F.conv2d(y, w2, leaf["tower_b"][2 * b + 1], padding=1) + x)
    n = len(boards)
    pol = F.relu(torch.einsum("bchw,cp->bphw", x, leaf["pol_w"]) + leaf["pol_b"].view(1, 2, 1, 1)).reshape(n, 84)
    logp = torch.log_softmax(pol @ leaf["pol_fc_w"] + leaf["pol_fc_b"], dim=1)
    val = F.relu(torch.einsum("bchw,c->bhw", x, leaf["val_w"]) + leaf["val_b"]).reshape(n, 42)
    h = F.relu(val @ leaf["val_fc1_w"] + leaf["val_fc1_b"])
    v = torch.tanh(h @ leaf["val_fc2_w"] + leaf["val_fc2_b"])
    loss_pi = -(torch.from_numpy(pis).double() * logp).sum(1).mean()
    loss_v = ((v - torch.from_numpy(vs).double()) ** 2).mean()
    (loss_pi + loss_v).backward()
    g = np.zeros(total, np.float64)
    for name, (o, shape) in L.items():
        g[o:o + int(np.prod(shape))] = leaf[name].grad.numpy().reshape(-1)
    return float(loss_pi.detach()), float(loss_v.detach()), g


def make_batch(oracle, n_games, seed):
    rng = np.random.default_rng(seed)
    boards = random_features(oracle, n_games, seed=seed)
    pis = rng.dirichlet(np.ones(7), len(boards)).astype(np.float32)
    vs = rng.choice(np.array([-1.0, 1.0, 1e-4], np.float32), len(boards))
    return boards, pis, vs


@pytest.mark.parametrize("blocks,n_games", [(1, 6), (2, 30)])
def test_loss_and_gradients_match_torch(azb, oracle, blocks, n_games):
    net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC)
    boards, pis, vs = make_batch(oracle, n_games, 11)
    lp, lv = net.train_begin(boards, pis, vs)
    g = net.grads().astype(np.float64)
    rp, rv, rg = torch_loss_and_grads(azb, net.get_params(), blocks, boards, pis, vs)
    assert abs(lp - rp) < 2e-2 * abs(rp) and abs(lv - rv) < 2e-2 * max(abs(rv), 0.1), (lp, rp, lv, rv)
    L = {k: v for k, v in azb.param_layout(blocks).items() if k != "total"}
    for name, (o, shape) in L.items():
        a, b = g[o:o + int(np.prod(shape))], rg[o:o + int(np.prod(shape))]
        na, nb = np.linalg.norm(a), np.linalg.norm(b)
        assert nb > 0, name
        cos = float(a @ b / (na * nb))
        # bf16 tower (8-bit mantissa) through up to 2*blocks layers, forward and backward: direction within 0.5 %
        # of a radian-ish, length within 5 %
        assert cos > 0.995, (name, cos)
        assert abs(na - nb) < 0.05 * nb, (name, na, nb)


def test_adam_step_and_training_lowers_the_loss(azb, oracle):
    net = azb.NNet(seed=3, blocks=1, precision=azb.NNET_BF16_TC)
    boards, pis, vs = make_batch(oracle, 12, 5)
    p0 = net.get_params().astype(np.float64)
    l0 = net.train_begin(boards, pis, vs)
    g = net.grads().astype(np.float64)
    net.train_apply(lr=1e-3)
    p1 = net.get_params().astype(np.float64)
    # first Adam step: m_hat = g, v_hat = g^2  ->  p -= lr * g / (|g| + eps)
    expect = p0 - 1e-3 * g / (np.abs(g) + 1e-8)
    assert np.abs(p1 - expect).max() < 2e-6
    # learnable labels: pi = one-hot of the fullest column (first maximum), v = +1 if the side to move owns at least as
    # many stones in the three centre columns as the opponent, else -1
    occ = boards[:, 0] + boards[:, 1]
    pis2 = np.eye(7, dtype=np.float32)[occ.sum(1).argmax(1)]
    vs2 = np.where(boards[:, 0, :, 2:5].sum((1, 2)) >= boards[:, 1, :, 2:5].sum((1, 2)), 1.0, -1.0).astype(np.float32)
    losses = [sum(net.train_begin(boards, pis2, vs2))]
    for _ in range(150):
        losses.append(sum(net.train((boards, pis2, vs2), lr=2e-3)))
    assert losses[-1] < 0.5 * losses[0] and losses[75] < losses[0], losses[::25]
    vs = vs2
    # the inference path sees the updated weights (tiles, stem table and head constants were rebuilt)
    pi, v = net.predict(boards)
    assert np.allclose(pi.sum(1), 1.0, atol=1e-5) and np.isfinite(v).all()
    assert np.mean((v - vs) ** 2) < 1.2 * losses[-1]


def test_default_step_size_is_justified(azb, oracle):
    """The reference's Adam runs at 1e-3 (connect_four_net.py:21) on a network with BatchNorm in front of every ReLU.  This
    network has none (folded away for inference): at 1e-3 the two-channel policy head dies within ~100 steps and the policy
    loss parks at ln 7 = 1.946 (uniform), at 1e-4 (azb_learn_config_default) the same batches train.  This is the evidence
    for the stated deviation; it fails if a future normalisation makes 1e-3 trainable (then the default should move back)."""
    feats = random_features(oracle, 60, seed=11)[:512]
    rng = np.random.default_rng(4)
    # a learnable target: the policy prefers the centre-most legal column of the position, the value its stone balance
    pis = np.zeros((len(feats), 7), np.float32)
    for i, f in enumerate(feats):
        free = [c for c in (3, 2, 4, 1, 5, 0, 6) if f[:, 0, c].sum() == 0]
        pis[i, free[0] if free else 3] = 1.0
    vs = np.tanh(feats[:, 0].sum((1, 2)) - feats[:, 1].sum((1, 2))).astype(np.float32)
    final = {}
    for lr in (1e-3, 1e-4):
        net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
        losses = []
        for step in range(160):
            idx = rng.permutation(len(feats))[:256]
            losses.append(net.train((feats[idx], pis[idx], vs[idx]), lr=lr))
        final[lr] = float(np.mean([l[0] for l in losses[-10:]]))
    # measured on B200: 1e-4 -> 0.0014 (the 512 positions are learnt), 1e-3 -> 1.71 (the policy head is dead: the loss sits at the
    # entropy of the targets' marginal, a little under ln 7 = 1.946)
    assert final[1e-4] < 0.2, final
    assert final[1e-3] > 1.2, final
