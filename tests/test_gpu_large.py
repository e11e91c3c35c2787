"""-m gpu parity at the sizes the round-1 review found unpinned (VERDICT.md "What's missing" 1-2):

* BASELINE.json configs[2] / configs[3] shapes — M = 500 and M = 1000 — against the CPU oracle
  (``oracle/shasta_oracle.py``, restating /root/reference/det3d/models/tracker/shasta.py:213-327), default kernel
  flags, i.e. the streamed aff kernel, the split-K anchors GEMM at K = 160 000 / 320 000 and 252 004 / 1 004 004 pairs;
* the literal bench configuration — M = 200, 512 x 512 x 64 maps, B = 64 frame pairs, flags 0, automatic anchors
  path, CUDA-graph replay — in one piece.

Bars: affinities within 1e-3 relative (BASELINE.json north_star); association (the decode of eval.py:126-181):
every dead / FN / kept / newborn decision identical and every argmax identical, except argmaxes whose top-2 margin in
the ORACLE is below 2e-5 relative (fp32 rounding noise: the reference's own pick depends on its summation order there;
counted and printed). The oracle needs ~9 GB (M = 500) / ~40 GB (M = 1000) of host memory: the tests skip with a message when
the host has less.
"""
import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from shasta_b200 import synthetic
from tests import gpu_util as G
from tests.golden_util import headline_names, load_golden

pytestmark = pytest.mark.gpu



def _host_free_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:  # noqa: BLE001
        return 0.0


def _device_maps(B, H, W, seed):
    g = torch.Generator(device=G.DEV)
    g.manual_seed(seed)
    bev = torch.relu(torch.randn((B, H, W, 64), generator=g, device=G.DEV))
    prev = torch.relu(torch.randn((B, H, W, 64), generator=g, device=G.DEV))
    return bev, prev


def _check_against_oracle(M, w, bev, prev_bev, data, pc_start, m1, m2, det_after, chunk=1):
    B = bev.shape[0]
    m1c, m2c = m1.cpu(), m2.cpu()
    worst, ties, small = 0.0, 0, 0.0
    for b0 in range(0, B, chunk):
        sl = slice(b0, min(B, b0 + chunk))
        det_o = torch.from_numpy(data["det_boxes"][sl].copy())
        o1, o2 = O.forward(w, bev[sl].cpu(), prev_bev[sl].cpu(), det_o, torch.from_numpy(data["prev_det_boxes"][sl]),
                           pc_start=pc_start)
        e1, e2 = G.rel_err(m1c[sl].numpy(), o1.numpy()), G.rel_err(m2c[sl].numpy(), o2.numpy())
        worst = max(worst, e1, e2)
        assert e1 < 1e-3 and e2 < 1e-3, (b0, e1, e2)
        # the same bar with the denominator floored at 1e-6 (not 1e-3) of the scale: small probabilities relatively
        s1 = G.rel_err(m1c[sl].numpy(), o1.numpy(), 1e-6)
        s2 = G.rel_err(m2c[sl].numpy(), o2.numpy(), 1e-6)
        small = max(small, s1, s2)
        assert s1 < 1e-3 and s2 < 1e-3, (b0, s1, s2)
        assert np.array_equal(det_after[sl].cpu().numpy(), det_o.numpy()), "in-place back-projection (shasta.py:270)"
        for i, b in enumerate(range(sl.start, sl.stop)):
            ties += G.assert_same_association(O, o1[i], o2[i], m1c[b], m2c[b], int(data["n_prev"][b]),
                                              int(data["n_det"][b]), where="frame pair %d" % b)
    assert torch.allclose(m1c.sum(2), torch.ones(B, M), atol=1e-5)
    assert torch.allclose(m2c.sum(1), torch.ones(B, M), atol=1e-5)
    print("worst relative error with the 1e-6 floor: %.3g" % small)
    print("argmax decisions inside the oracle's fp32 noise (top-2 margin < 2e-5 relative): %d of %d"
          % (ties, 2 * B * M))
    return worst


@pytest.mark.parametrize("M,need_gb", [(500, 24), (1000, 90)])
def test_large_sizes_against_oracle(M, need_gb):
    """configs[2] (500 x 500) and configs[3] (1000 x 1000) against the oracle, B = 1, default flags."""
    if _host_free_gb() < need_gb:
        pytest.skip("host has %.0f GB free, the M = %d oracle needs ~%d GB" % (_host_free_gb(), M, need_gb))
    H = W = 96
    pc_start = (-W * 0.3, -H * 0.3)
    torch.manual_seed(1000 + M)
    model = G.make_model(M, pc_start)   # default nn.Linear init on the device (the weights are 6.4 / 25.6 GB)
    with torch.no_grad():
        model.aff[10].weight.mul_(300.0)   # "peaky": some affinities cross the 0.5 / 0.7 decode thresholds
    data = synthetic.make_frame_pairs(1, M, H, W, 77 + M, pc_start=pc_start, with_maps=False)
    bev, prev_bev = _device_maps(1, H, W, 5 + M)
    det = G.t(data["det_boxes"])
    with torch.no_grad():
        m1, m2 = model.affinity(bev, prev_bev, det, G.t(data["prev_det_boxes"]))
    torch.cuda.synchronize()
    head = set(synthetic.head_param_shapes(M))
    w = {k: v.detach().cpu() for k, v in model.state_dict().items() if k in head}
    del model
    torch.cuda.empty_cache()
    worst = _check_against_oracle(M, w, bev, prev_bev, data, pc_start, m1, m2, det)
    print("M = %d: worst relative affinity error vs the oracle %.3g" % (M, worst))


def test_bench_configuration_against_oracle():
    """The configuration bench.py times, in one piece: M = 200, 512 x 512 x 64 maps, B = 64, kernel flags 0, automatic
    anchors path, the forward captured into a CUDA graph and replayed."""
    M, H, W, B = 200, 512, 512, 64
    pc_start = (-W * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(B, M, H, W, 4242, pc_start=pc_start, with_maps=False)
    weights = synthetic.make_weights(M, seed=23, peaky=300.0)
    model = G.make_model(M, pc_start, weights)
    model.kernel_flags = 0
    model.cuda_graphs = True
    bev, prev_bev = _device_maps(B, H, W, 99)
    det0 = G.t(data["det_boxes"])
    prev = G.t(data["prev_det_boxes"])
    det = det0.clone()
    with torch.no_grad():
        a1, a2 = model.affinity(bev, prev_bev, det, prev)      # capture + first replay
        first = (a1.clone(), a2.clone(), det.clone())
        det.copy_(det0)
        m1, m2 = model.affinity(bev, prev_bev, det, prev)      # replay of the same graph entry
    torch.cuda.synchronize()
    assert len(model._graphs) == 1
    assert torch.equal(first[0], m1) and torch.equal(first[1], m2) and torch.equal(first[2], det), \
        "graph replay must reproduce the captured run bit for bit"
    w = O.weights_to_torch(weights)
    worst = _check_against_oracle(M, w, bev, prev_bev, data, pc_start, m1, m2, det, chunk=4)
    print("bench configuration (M=200, 512^2, B=64, graph replay): worst relative affinity error %.3g" % worst)


@pytest.mark.parametrize("name", headline_names())
def test_forward_matches_reference_golden_at_headline_size(name):
    """M = 200 against outputs of the UNMODIFIED reference (tests/golden/h200_*.npz, oracle/make_golden.py --headline):
    no oracle in between."""
    c, pc_start, data, weights, g = load_golden(name)
    model = G.make_model(c["M"], pc_start, weights)
    det = G.t(data["det_boxes"])
    with torch.no_grad():
        m1, m2 = model.affinity(G.t(data["bev"]), G.t(data["prev_bev"]), det, G.t(data["prev_det_boxes"]))
    m1, m2 = m1.cpu(), m2.cpu()
    assert G.rel_err(m1.numpy(), g["matched1"]) < 1e-3 and G.rel_err(m2.numpy(), g["matched2"]) < 1e-3
    assert np.array_equal(det.cpu().numpy(), g["det_boxes_after"])
    ties = 0
    for b in range(c["B"]):
        ties += G.assert_same_association(O, torch.from_numpy(g["matched1"][b]), torch.from_numpy(g["matched2"][b]),
                                          m1[b], m2[b], int(data["n_prev"][b]), int(data["n_det"][b]))
    print("%s: rel err %.3g / %.3g, fp32-level argmax ties %d" % (name, G.rel_err(m1.numpy(), g["matched1"]),
                                                                   G.rel_err(m2.numpy(), g["matched2"]), ties))
