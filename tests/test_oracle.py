"""CPU: the oracle restatement (oracle/shasta_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (tests/golden, oracle/make_golden.py), plus properties of the path."""
import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from tests.golden_util import golden_names, headline_names, load_golden


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # fixtures were generated single-threaded: same summation order
    yield
    torch.set_num_threads(n)


def _run(name):
    c, pc_start, data, weights, g = load_golden(name)
    w = O.weights_to_torch(weights)
    det = torch.from_numpy(data["det_boxes"].copy())
    m1, m2, inter = O.forward(w, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]), det,
                              torch.from_numpy(data["prev_det_boxes"]), pc_start=pc_start, return_intermediates=True)
    return c, data, g, det, m1, m2, inter


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden_bit_exact(name):
    c, data, g, det, m1, m2, inter = _run(name)
    assert np.array_equal(m1.numpy(), g["matched1"])
    assert np.array_equal(m2.numpy(), g["matched2"])
    assert np.array_equal(inter["feature"].numpy(), g["feature"])
    assert np.array_equal(inter["prev_feature"].numpy(), g["prev_feature"])
    assert np.array_equal(inter["residual"].numpy(), g["residual"])
    assert np.array_equal(inter["logits"].numpy(), g["logits"])
    for k in ("newborn", "fp", "dead_trk", "fn"):
        assert np.array_equal(inter[k].numpy(), g[k])
    assert np.array_equal(torch.stack(inter["aug_shape"]).numpy(), g["aug_shape"])
    # in-place back-projection of the caller's boxes (shasta.py:270)
    assert np.array_equal(det.numpy(), g["det_boxes_after"])
    assert not np.array_equal(det.numpy(), data["det_boxes"])


@pytest.mark.parametrize("name", headline_names())
def test_oracle_matches_reference_golden_at_headline_size(name):
    """M = 200 (BASELINE.json configs[0]): outputs of the unmodified reference, bit for bit."""
    c, pc_start, data, weights, g = load_golden(name)
    det = torch.from_numpy(data["det_boxes"].copy())
    m1, m2 = O.forward(O.weights_to_torch(weights), torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]),
                       det, torch.from_numpy(data["prev_det_boxes"]), pc_start=pc_start)
    assert np.array_equal(m1.numpy(), g["matched1"])
    assert np.array_equal(m2.numpy(), g["matched2"])
    assert np.array_equal(det.numpy(), g["det_boxes_after"])


@pytest.mark.parametrize("name", golden_names())
def test_softmax_normalisation_and_shapes(name):
    c, data, g, det, m1, m2, inter = _run(name)
    B, M = c["B"], c["M"]
    assert tuple(m1.shape) == (B, M, M + 2) and tuple(m2.shape) == (B, M + 2, M)
    assert torch.allclose(m1.sum(dim=2), torch.ones(B, M), atol=1e-5)
    assert torch.allclose(m2.sum(dim=1), torch.ones(B, M), atol=1e-5)


def test_first_layer_decomposition_is_the_same_function():
    """The identity the CUDA path relies on: Linear([a;b]) = W[:, :A] a + W[:, A:] b + bias."""
    c, pc_start, data, weights, g = load_golden("m6_16px_b2")
    w = O.weights_to_torch(weights, torch.float64)
    f_prev = torch.randn(8, 320, dtype=torch.float64)
    f_cur = torch.randn(8, 320, dtype=torch.float64)
    W, b = w["fuse_shape.0.weight"], w["fuse_shape.0.bias"]
    full = torch.nn.functional.linear(torch.cat([f_prev[:, None].expand(8, 8, 320), f_cur[None].expand(8, 8, 320)], -1), W, b)
    dec = (f_prev @ W[:, :320].T)[:, None] + (f_cur @ W[:, 320:].T + b)[None]
    assert torch.allclose(full, dec, atol=1e-12)


def test_bilinear_border_semantics():
    """center_utils.py:103-119: weights come from the CLAMPED integers, so points on/outside the last row or
    column give ~0, not the border pixel (SURVEY.md §7 hard parts)."""
    im = torch.rand(5, 7, 4) + 1.0
    x = torch.tensor([6.0, -1.5, 8.2, 2.25, 0.0])
    y = torch.tensor([1.0, 1.0, 2.0, 4.0, 0.0])
    out = O.bilinear_interpolate(im, x, y)
    assert torch.all(out[:4].abs() < 1e-5)
    assert torch.equal(out[4], im[0, 0])
    xi = torch.tensor([2.25])
    yi = torch.tensor([1.5])
    ref = (im[1, 2] * 0.75 * 0.5 + im[2, 2] * 0.75 * 0.5) + im[1, 3] * 0.25 * 0.5 + im[2, 3] * 0.25 * 0.5
    assert torch.allclose(O.bilinear_interpolate(im, xi, yi)[0], ref, atol=1e-6)


def test_padded_rows_are_processed_like_real_ones():
    """No padding mask anywhere: zero boxes sample the map at the ego origin and flow on."""
    c, pc_start, data, weights, g = load_golden("m20_32px_b2")
    n_det = int(data["n_det"][0])
    assert n_det < c["M"]
    pad = g["feature"][0, n_det:]
    assert np.all(pad == pad[0])          # all padded boxes are identical -> identical features
    assert np.isfinite(g["matched1"]).all() and np.isfinite(g["matched2"]).all()


def test_decode_thresholds():
    M = 4
    m1 = torch.full((M, M + 2), 0.01)
    m2 = torch.full((M + 2, M), 0.01)
    m1[0, 1] = 0.9       # prev 0 -> det 1 (kept)
    m1[1, M] = 0.8       # prev 1 dead
    m1[2, M + 1] = 0.6   # prev 2 false negative, score = 1 - P(dead)
    m1[2, M] = 0.1
    m2[0, 1] = 0.9
    m2[M, 0] = 0.55      # det 0 newborn
    m2[M + 1, 2] = 0.75  # det 2 false positive -> dropped
    m2[M + 1, 0] = 0.2
    d = O.decode(m1, m2, n_prev=3, n_det=3)
    assert d["dead"] == [1] and d["fn"] == [2] and d["keep_prev"] == [0]
    assert abs(d["fn_score"][0] - 0.9) < 1e-6
    assert d["keep_dets"] == [0, 1] and d["newborn"] == [True, False]
    assert abs(d["det_score"][0] - 0.8) < 1e-6
    # empty sides (eval.py:130,156)
    d0 = O.decode(m1, m2, n_prev=0, n_det=2)
    assert d0["keep_prev"] == [] and d0["keep_dets"] == [0, 1]
    d1 = O.decode(m1, m2, n_prev=2, n_det=0)
    assert d1["keep_dets"] == [] and d1["dead"] == [1]


def test_loss_matches_formula():
    torch.manual_seed(0)
    M = 5
    m1 = torch.softmax(torch.randn(2, M, M + 2), 2)
    m2 = torch.softmax(torch.randn(2, M + 2, M), 1)
    gt = (torch.rand(2, M + 2, M + 2) > 0.8).float()
    l = O.affinity_loss(m1, m2, gt)
    gt1, gt2 = gt[:, :-2, :], gt[:, :, :-2]
    ref = ((gt1 * -torch.log(m1 + 1e-10)).sum() / gt1.sum() + (gt2 * -torch.log(m2 + 1e-10)).sum() / gt2.sum()) / 2
    assert torch.allclose(l, ref)
    assert O.affinity_loss(m1, m2, torch.zeros_like(gt)) == 0
