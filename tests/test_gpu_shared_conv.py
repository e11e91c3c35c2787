"""shared_conv producer on tcgen05 vs the oracle's conv2d + BatchNorm(eval) + ReLU + NHWC permute."""
import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from shasta_b200 import synthetic
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _conv_weights(seed):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / np.sqrt(512 * 9)
    return {
        "shared_conv.0.weight": (torch.rand((64, 512, 3, 3), generator=g) * 2 - 1) * bound,
        "shared_conv.0.bias": (torch.rand(64, generator=g) * 2 - 1) * bound,
        "shared_conv.1.weight": torch.rand(64, generator=g) + 0.5,
        "shared_conv.1.bias": torch.randn(64, generator=g) * 0.1,
        "shared_conv.1.running_mean": torch.randn(64, generator=g) * 0.05,
        "shared_conv.1.running_var": torch.rand(64, generator=g) + 0.5,
    }


def _model_with(weights, M=6):
    model = G.make_model(M, (-4.8, -4.8), synthetic.make_weights(M, seed=1))
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(v)
    return model


@pytest.mark.parametrize("N,H,W", [(1, 8, 16), (2, 20, 37), (3, 9, 5), (1, 180, 180)])
def test_shared_conv_matches_oracle(N, H, W):
    """Patch tiles 16 x 8: exact fit, ragged right/bottom edges, maps smaller than a patch, the real 180 x 180 map.
    fp32-equivalent arithmetic (3xTF32, chains flushed every tap): 1e-5 of the output scale."""
    w = _conv_weights(3)
    model = _model_with(w)
    g = torch.Generator().manual_seed(10 + H)
    x = torch.relu(torch.randn((N, 512, H, W), generator=g))
    want = O.shared_conv_nhwc(w, x).numpy()
    with torch.no_grad():
        got = model.shared_conv_nhwc(x.to(G.DEV))
    assert tuple(got.shape) == (N, H, W, 64) and got.is_contiguous()
    got = got.cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    print("shared_conv %dx%dx%d: max err / scale = %.3g" % (N, H, W, err))
    assert err < 1e-5


def test_shared_conv_chunking_and_repack():
    """More maps than one launch takes; weights changed in place trigger a re-pack."""
    w = _conv_weights(4)
    model = _model_with(w)
    x = torch.relu(torch.randn((5, 512, 12, 20), generator=torch.Generator().manual_seed(2)))
    with torch.no_grad():
        a = model.shared_conv_nhwc(x.to(G.DEV), maps_per_launch=2).cpu().numpy()
        b = model.shared_conv_nhwc(x.to(G.DEV), maps_per_launch=8).cpu().numpy()
        assert np.array_equal(a, b)
        assert np.abs(a - O.shared_conv_nhwc(w, x).numpy()).max() / np.abs(a).max() < 1e-5
        model.shared_conv[0].weight.mul_(0.5)
        w2 = dict(w)
        w2["shared_conv.0.weight"] = w["shared_conv.0.weight"] * 0.5
        c = model.shared_conv_nhwc(x.to(G.DEV)).cpu().numpy()
        assert np.abs(c - O.shared_conv_nhwc(w2, x).numpy()).max() / np.abs(c).max() < 1e-5


def test_forward_from_512_channel_maps():
    """Shasta.forward with an attached trunk stub: shared_conv (CUDA) -> gather -> ... equals the oracle chain."""
    M, H, W, B = 6, 16, 16, 2
    pc_start = (-W * 0.3, -H * 0.3)
    hw = synthetic.make_weights(M, seed=2)
    cw = _conv_weights(5)
    model = G.make_model(M, pc_start, hw)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in cw.items():
            sd[k].copy_(v)
    d = synthetic.make_frame_pairs(B, M, H, W, 9, pc_start=pc_start, with_maps=False)
    g = torch.Generator().manual_seed(3)
    raw = torch.relu(torch.randn((B, 512, H, W), generator=g))
    raw_prev = torch.relu(torch.randn((B, 512, H, W), generator=g))
    model.extract_feat = lambda ex: (raw.to(G.DEV), None, raw_prev.to(G.DEV), None)
    ex = {"det_boxes": G.t(d["det_boxes"]), "prev_det_boxes": G.t(d["prev_det_boxes"])}
    with torch.no_grad():
        m1, m2, ex = model(ex, train_mode=False)
    wt = O.weights_to_torch(hw)
    o1, o2 = O.forward(wt, O.shared_conv_nhwc(cw, raw), O.shared_conv_nhwc(cw, raw_prev),
                       torch.from_numpy(d["det_boxes"].copy()), torch.from_numpy(d["prev_det_boxes"]),
                       pc_start=pc_start)
    assert G.rel_err(m1.cpu().numpy(), o1.numpy()) < 2e-4
    assert G.rel_err(m2.cpu().numpy(), o2.numpy()) < 2e-4
    assert "bev_feature" in ex
