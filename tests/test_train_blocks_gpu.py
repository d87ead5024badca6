"""Building blocks of NNet::train (src/nnet.rs:38) on the tensor cores against torch (float64 with the
same bf16 quantisation points): the 3x3 convolution forward, backward data (the forward kernel with tap-mirrored,
transposed weight tiles and a d-ReLU epilogue) and backward weights (tcgen05 with MN-major operands)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def q(t):
    import torch
    return t.to(torch.bfloat16).to(torch.float64)


def setup(azb, n, seed, layer):
    import torch
    net = azb.NNet(seed=5, blocks=2, precision=azb.NNET_BF16_TC)
    L = azb.param_layout(2)
    prm = net.get_params()
    o, shape = L["tower_w"]
    tw = torch.from_numpy(prm[o:o + int(np.prod(shape))].reshape(shape).copy())
    o, shape = L["tower_b"]
    tb = torch.from_numpy(prm[o:o + int(np.prod(shape))].reshape(shape).copy()).double()
    w = q(tw[layer]).reshape(3, 3, 128, 128).permute(3, 2, 0, 1).contiguous()  # [co][ci][ky][kx]
    rng = np.random.default_rng(seed)
    mk = lambda: rng.standard_normal((n, 42, 128)).astype(np.float32)
    return net, w, tb[layer], mk


def to_nchw(a):
    import torch
    return q(torch.from_numpy(a)).reshape(len(a), 6, 7, 128).permute(0, 3, 1, 2)


def from_nchw(t):
    return t.permute(0, 2, 3, 1).reshape(len(t), 42, 128).numpy()


@pytest.mark.parametrize("n", [1, 5, 37, 300])
def test_conv_forward_hook(azb, n):
    import torch.nn.functional as F
    net, w, b, mk = setup(azb, n, 1, layer=1)
    x, res = mk(), mk()
    y = net.conv_hook(1, 0, x, residual=res)
    ref = F.relu(F.conv2d(to_nchw(x), w, b, padding=1) + to_nchw(res))
    ref = from_nchw(q(ref.float()))
    assert np.abs(y - ref).max() <= 2.0 ** -7 * max(1.0, np.abs(ref).max())  # one bf16 ulp of the largest value
    assert np.mean(y != ref) < 0.02  # (fp32 accumulation order: a few borderline roundings)


@pytest.mark.parametrize("n", [1, 5, 37, 300])
def test_conv_backward_data_hook(azb, n):
    import torch
    import torch.nn.functional as F
    net, w, b, mk = setup(azb, n, 2, layer=2)
    dz, add, act = mk(), mk(), mk()
    dx = net.conv_hook(2, 1, dz, residual=add, mask=act)
    xin = torch.zeros(n, 128, 6, 7, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, w, None, padding=1).backward(to_nchw(dz))
    ref = (xin.grad + to_nchw(add)) * (to_nchw(act) > 0)
    ref = from_nchw(q(ref.float()))
    assert np.abs(dx - ref).max() <= 2.0 ** -7 * max(1.0, np.abs(ref).max())
    assert np.mean(dx != ref) < 0.02
    # without the optional operands
    dx0 = net.conv_hook(2, 1, dz)
    ref0 = from_nchw(q(xin.grad.float()))
    assert np.abs(dx0 - ref0).max() <= 2.0 ** -7 * max(1.0, np.abs(ref0).max())


@pytest.mark.parametrize("n", [1, 5, 37, 300, 2000])
def test_conv_backward_weights_hook(azb, n):
    import torch
    import torch.nn.functional as F
    net, w, b, mk = setup(azb, n, 3, layer=0)
    x, dz = mk(), mk()
    dw = net.wgrad_hook(x, dz)
    wv = torch.zeros(128, 128, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(to_nchw(x), wv, None, padding=1).backward(to_nchw(dz))
    ref = wv.grad.permute(2, 3, 1, 0).reshape(9, 128, 128).numpy()  # [tap][ci][co]
    scale = np.abs(ref).max()
    assert np.abs(dw - ref).max() <= 1e-4 * scale + 1e-5 * np.sqrt(n * 42), (np.abs(dw - ref).max(), scale)
