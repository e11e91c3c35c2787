"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the reference goldens and the
CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): affinities within 1e-3 relative in fp32 (asserted at 2e-4 here), association
(argmax / decode) identical, index / copy work bit-exact (bilinear on given coordinates, in-place back-projection).
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from shasta_b200 import _cabi, synthetic
from tests import gpu_util as G
from tests.golden_util import golden_names, load_golden

pytestmark = pytest.mark.gpu

FP32_REL_TOL = 2e-4   # north_star allows 1e-3


@pytest.fixture(autouse=True)
def _lib(shasta_lib):
    return shasta_lib


# ------------------------------------------------------------------------------------------------
# a2: bilinear on explicit coordinates — bit exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,C,n", [(5, 7, 4, 64), (32, 48, 64, 1000), (180, 180, 64, 4096), (1, 1, 8, 10)])
def test_bilinear_bit_exact(H, W, C, n):
    rng = np.random.default_rng(H * 1000 + W)
    im = rng.standard_normal((H, W, C)).astype(np.float32)
    xs = rng.uniform(-3, W + 2, n).astype(np.float32)
    ys = rng.uniform(-3, H + 2, n).astype(np.float32)
    # exact integers, the last row/column, far outside
    xs[:6] = [0.0, W - 1.0, W - 1.0, -1.0, W + 0.5, 1e6]
    ys[:6] = [0.0, H - 1.0, 0.0, 2.0, -0.5, -1e6]
    want = O.bilinear_interpolate(torch.from_numpy(im), torch.from_numpy(xs), torch.from_numpy(ys)).numpy()
    lib = _cabi.lib()
    dim, dx, dy = G.t(im), G.t(xs), G.t(ys)
    out = torch.empty((n, C), device=G.DEV)
    rc = lib.shasta_bilinear_f32(dim.data_ptr(), H, W, C, dx.data_ptr(), dy.data_ptr(), n, out.data_ptr(), G.stream())
    _cabi.check(rc, "bilinear")
    got = out.cpu().numpy()
    assert np.array_equal(got, want), "max diff %g" % np.abs(got - want).max()


def test_bilinear_empty_and_bad_args():
    lib = _cabi.lib()
    im = torch.zeros((4, 4, 8), device=G.DEV)
    out = torch.zeros((1, 8), device=G.DEV)
    assert lib.shasta_bilinear_f32(im.data_ptr(), 4, 4, 8, None, None, 0, out.data_ptr(), G.stream()) == 0
    assert lib.shasta_bilinear_f32(im.data_ptr(), 4, 4, 6, None, None, 0, out.data_ptr(), G.stream()) < 0
    assert b"multiple of 4" in lib.shasta_last_error_string()


def test_bev_extractor_module_api():
    """BEVFeatureExtractor.forward(example, batch_centers, num_point) like bird_eye_view.py:24-41."""
    c, pc_start, data, weights, g = load_golden("m20_32px_b2")
    model = G.make_model(c["M"], pc_start, weights)
    bev = G.t(data["bev"])
    boxes = torch.from_numpy(data["det_boxes"][:, :, :7])
    centers = [O.box_points(boxes[b]).to(G.DEV) for b in range(c["B"])]
    out = model.bev_extractor({"bev_feature": bev}, centers, 5)
    assert isinstance(out, list) and len(out) == c["B"] and tuple(out[0].shape) == (c["M"], 320)
    got = torch.stack(out).cpu().numpy()
    # coordinates computed by torch on the GPU may differ by an ulp from the CPU's: tolerance, not bits
    assert np.allclose(got, g["feature"], rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# a1+a2: fused box gather, both samplers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("name", golden_names())
def test_gather_from_boxes(name, variant):
    c, pc_start, data, weights, g = load_golden(name)
    model = G.make_model(c["M"], pc_start, weights)
    st = G.Stages(model, c["B"])
    for bev_k, box_k, gold_k, region in (("bev", "det_boxes", "feature", _cabi.WS_FEAT_CUR),
                                         ("prev_bev", "prev_det_boxes", "prev_feature", _cabi.WS_FEAT_PREV)):
        feat = st.gather(G.t(data[bev_k]), G.t(data[box_k]), region, variant)
        got = feat[:, :c["M"], :].cpu().numpy()
        want = g[gold_k]
        exact = float(np.mean(got == want))
        assert np.allclose(got, want, rtol=1e-4, atol=1e-4), (gold_k, np.abs(got - want).max())
        assert exact > 0.5, "only %.3f of the gathered values are bit-identical" % exact


@pytest.mark.parametrize("M,H,W,B", [(7, 9, 11, 3), (200, 180, 180, 2), (33, 2, 2, 1), (5, 1, 6, 2)])
def test_gather_samplers_bit_identical(M, H, W, B):
    """The four samplers (direct loads, bulk-copy staged, narrow grid, one TMA box {64 ch, 2 px, 2 px} per point) run
    the same tap / blend arithmetic: identical bits, also for boxes on the border and far outside the map (clamped
    taps: center_utils.py:104-107; the TMA variant must not zero-fill them), ragged last tiles and maps too small
    for a 2 x 2 box."""
    rng = np.random.default_rng(M * 7 + H)
    pc_start = (-W * 0.6 / 2.0, -H * 0.6 / 2.0)
    model = G.make_model(M, pc_start)
    st = G.Stages(model, B)
    bev = rng.standard_normal((B, H, W, 64)).astype(np.float32)
    boxes = np.zeros((B, M, 11), np.float32)
    boxes[..., 0] = rng.uniform(-W * 0.45, W * 0.45, (B, M))
    boxes[..., 1] = rng.uniform(-H * 0.45, H * 0.45, (B, M))
    boxes[..., 3:6] = rng.uniform(0.5, 6.0, (B, M, 3))
    boxes[..., 6] = rng.uniform(-3.2, 3.2, (B, M))
    boxes[:, 0, 0] = -W * 0.3 + 0.01          # left border: x0 clamps
    boxes[:, 1, 1] = H * 0.3 - 0.01           # bottom border
    boxes[:, 2, :2] = 1e4                     # far outside
    boxes[:, 3, :2] = -1e4
    dbev, dbox = G.t(bev), G.t(boxes)
    got = {}
    for variant in (0, 1, 2, 3):
        st.ws.buf.zero_()
        got[variant] = st.gather(dbev, dbox, _cabi.WS_FEAT_CUR, variant)[:, :M].cpu().numpy().copy()
    for variant in (1, 2, 3):
        assert np.array_equal(got[variant], got[0]), (variant, np.abs(got[variant] - got[0]).max())
    # and against the oracle's sampler on the same pixel coordinates (bit-exact blend, coordinates to 1e-4)
    want = O.gather_box_features(torch.from_numpy(bev), torch.from_numpy(boxes[..., :7].copy()), pc_start,
                                 (0.075, 0.075), 8).numpy()
    assert np.allclose(got[3], want, rtol=1e-4, atol=2e-4), np.abs(got[3] - want).max()


def test_forward_with_tma_gather_bit_identical():
    """Whole forward with the TMA-box sampler (flags & 3 = 3, writes the features AND their tf32 low parts for the
    anchors GEMM) against the default sampler: same bits out."""
    c, pc_start, data, weights, g = load_golden("m20_32px_b3_peaky")
    outs = []
    for flags in (0, 3):
        model = G.make_model(c["M"], pc_start, weights)
        model.kernel_flags = flags
        model.cuda_graphs = False
        lib = _cabi.lib()
        lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, _cabi.ANCHOR_TC)   # the GEMM that reads FEATLO
        try:
            m1, m2 = model.affinity(G.t(data["bev"]), G.t(data["prev_bev"]), G.t(data["det_boxes"]), G.t(data["prev_det_boxes"]))
            torch.cuda.synchronize()
        finally:
            lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, _cabi.ANCHOR_AUTO)
        outs.append((m1.cpu().numpy(), m2.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


# ------------------------------------------------------------------------------------------------
# a3+a4: anchors
# ------------------------------------------------------------------------------------------------
@pytest.fixture(params=[_cabi.ANCHOR_TC, _cabi.ANCHOR_TC_GEN1], ids=["tc2", "tc1"])
def force_tc_anchors(request):
    lib = _cabi.lib()
    lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, request.param)
    yield
    lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, _cabi.ANCHOR_AUTO)


@pytest.mark.parametrize("name", golden_names())
def test_anchors_stage(name):
    _anchors_stage(name)


@pytest.mark.parametrize("name", golden_names())
def test_anchors_stage_tcgen05(name, force_tc_anchors):
    """Same check with the tcgen05 3xTF32 GEMM forced (tiles far larger than these shapes: exercises the TMA
    out-of-bounds zero fill on rows, the split-K partials and the TMEM epilogue masks)."""
    _anchors_stage(name)


@pytest.mark.parametrize("raw_hi", [0, 1])
@pytest.mark.parametrize("B", [12, 130])
def test_anchor_paths_agree_at_m200(B, raw_hi):
    """Streaming CUDA-core kernel vs tcgen05 kernel on the headline shape (K = 64000, 4 x 1000 rows), one and two
    batch tiles."""
    M = 200
    lib = _cabi.lib()
    model = G.make_model(M, (-54.0, -54.0))
    st = G.Stages(model, B)
    gen = torch.Generator(device=G.DEV).manual_seed(B)
    fc = st.region(_cabi.WS_FEAT_CUR, (B, M + 2, 320))
    fp = st.region(_cabi.WS_FEAT_PREV, (B, M + 2, 320))
    fc[:, :M] = torch.randn((B, M, 320), device=G.DEV, generator=gen).relu_()
    fp[:, :M] = torch.randn((B, M, 320), device=G.DEV, generator=gen).relu_()
    boxes = torch.randn((B, M, 11), device=G.DEV, generator=gen)
    out = {}
    for mode in (_cabi.ANCHOR_STREAM, _cabi.ANCHOR_TC, _cabi.ANCHOR_TC_GEN1):
        lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, mode)
        lib.shasta_set_option(_cabi.OPT_TC_RAW_HI, raw_hi)
        try:
            st.anchors(boxes, boxes)
        finally:
            lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, _cabi.ANCHOR_AUTO)
            lib.shasta_set_option(_cabi.OPT_TC_RAW_HI, 1)
        torch.cuda.synchronize()
        out[mode] = torch.cat([fc[:, M:], fp[:, M:]], dim=1).clone()
    a, b = out[_cabi.ANCHOR_STREAM].cpu().numpy(), out[_cabi.ANCHOR_TC].cpu().numpy()
    b1 = out[_cabi.ANCHOR_TC_GEN1].cpu().numpy()
    assert np.isfinite(b).all() and np.isfinite(b1).all()
    # float64 reference of all four anchors: rows [dead, fn] of FEAT_CUR then [newborn, fp] of FEAT_PREV
    xc = fc[:, :M].reshape(B, -1).double()
    xp = fp[:, :M].reshape(B, -1).double()
    ref = []
    with torch.no_grad():
        for i, x in ((2, xp), (3, xp), (0, xc), (1, xc)):
            l0, l2 = model.aug_shape[i][0], model.aug_shape[i][2]
            h = torch.relu(x @ l0.weight.double().T + l0.bias.double())
            ref.append(torch.abs(h @ l2.weight.double().T + l2.bias.double()))
    ref = torch.stack(ref, dim=1).cpu().numpy()
    scale = np.abs(ref).max()
    err_stream = np.abs(a - ref).max() / scale
    err_tc = np.abs(b - ref).max() / scale
    err_tc1 = np.abs(b1 - ref).max() / scale
    print("anchor L1+L2 max err / scale vs float64 (B=%d raw_hi=%d): streaming fp32 %.3g, tcgen05 (bounded chains) "
          "%.3g, tcgen05 gen1 %.3g" % (B, raw_hi, err_stream, err_tc, err_tc1))
    assert err_stream < 1e-5 and err_tc < 1e-5 and err_tc1 < 1e-4, (err_stream, err_tc, err_tc1)


def _anchors_stage(name):
    c, pc_start, data, weights, g = load_golden(name)
    B, M = c["B"], c["M"]
    model = G.make_model(M, pc_start, weights)
    st = G.Stages(model, B)
    st.region(_cabi.WS_FEAT_CUR, (B, M + 2, 320))[:, :M] = G.t(g["feature"])
    st.region(_cabi.WS_FEAT_PREV, (B, M + 2, 320))[:, :M] = G.t(g["prev_feature"])
    det, prev = G.t(data["det_boxes"]), G.t(data["prev_det_boxes"])
    st.anchors(det, prev)
    assert np.array_equal(det.cpu().numpy(), data["det_boxes"]), "anchors stage must not modify its inputs"
    fc = st.region(_cabi.WS_FEAT_CUR, (B, M + 2, 320)).cpu().numpy()
    fp = st.region(_cabi.WS_FEAT_PREV, (B, M + 2, 320)).cpu().numpy()
    aug = g["aug_shape"]  # (4,B,1,320): newborn, fp, dead, fn
    for got, want in ((fp[:, M], aug[0][:, 0]), (fp[:, M + 1], aug[1][:, 0]), (fc[:, M], aug[2][:, 0]),
                      (fc[:, M + 1], aug[3][:, 0])):
        assert np.allclose(got, want, rtol=2e-4, atol=2e-5), np.abs(got - want).max()
    ab = st.region(_cabi.WS_ANCHOR_BOX, (B, 4, 7)).cpu().numpy()
    for i, k in enumerate(("newborn", "fp", "dead_trk", "fn")):
        assert np.allclose(ab[:, i], g[k][:, 0], rtol=2e-4, atol=2e-5), k
    bc = st.region(_cabi.WS_BOX_CUR, (B, M + 2, 8)).cpu().numpy()
    bp = st.region(_cabi.WS_BOX_PREV, (B, M + 2, 8)).cpu().numpy()
    assert np.array_equal(bc[:, :M, :7], g["det_boxes_after"][:, :, :7])   # back-projection: bit exact
    assert np.array_equal(bp[:, :M, :7], data["prev_det_boxes"][:, :, :7])
    assert np.allclose(bc[:, M, :7], g["dead_trk"][:, 0], rtol=2e-4, atol=2e-5)
    assert np.allclose(bp[:, M + 1, :7], g["fp"][:, 0], rtol=2e-4, atol=2e-5)


# ------------------------------------------------------------------------------------------------
# per-object projections / aux / colnorm, then pairwise
# ------------------------------------------------------------------------------------------------
def _fill_aug(st, c, data, g):
    """Puts the reference's augmented features and boxes into the workspace."""
    B, M = c["B"], c["M"]
    fc = st.region(_cabi.WS_FEAT_CUR, (B, M + 2, 320))
    fp = st.region(_cabi.WS_FEAT_PREV, (B, M + 2, 320))
    fc[:, :M], fp[:, :M] = G.t(g["feature"]), G.t(g["prev_feature"])
    aug = g["aug_shape"]
    fp[:, M], fp[:, M + 1] = G.t(aug[0][:, 0]), G.t(aug[1][:, 0])
    fc[:, M], fc[:, M + 1] = G.t(aug[2][:, 0]), G.t(aug[3][:, 0])
    bc = st.region(_cabi.WS_BOX_CUR, (B, M + 2, 8))
    bp = st.region(_cabi.WS_BOX_PREV, (B, M + 2, 8))
    bc.zero_(), bp.zero_()
    bc[:, :M, :7] = G.t(g["det_boxes_after"][:, :, :7])
    bp[:, :M, :7] = G.t(data["prev_det_boxes"][:, :, :7])
    bc[:, M, :7], bc[:, M + 1, :7] = G.t(g["dead_trk"][:, 0]), G.t(g["fn"][:, 0])
    bp[:, M, :7], bp[:, M + 1, :7] = G.t(g["newborn"][:, 0]), G.t(g["fp"][:, 0])
    prev_aug = np.concatenate([data["prev_det_boxes"][:, :, :7], g["newborn"], g["fp"]], axis=1)
    det_aug = np.concatenate([g["det_boxes_after"][:, :, :7], g["dead_trk"], g["fn"]], axis=1)
    f_prev = np.concatenate([g["prev_feature"], aug[0], aug[1]], axis=1)
    f_cur = np.concatenate([g["feature"], aug[2], aug[3]], axis=1)
    return prev_aug, det_aug, f_prev, f_cur


@pytest.fixture(params=[1, 2], ids=["ffma", "tcgen05"])
def project_path(request):
    lib = _cabi.lib()
    lib.shasta_set_option(_cabi.OPT_PROJECT_PATH, request.param)
    yield request.param
    lib.shasta_set_option(_cabi.OPT_PROJECT_PATH, 0)


@pytest.mark.parametrize("name", golden_names())
def test_project_and_pairwise_stage(name, project_path):
    c, pc_start, data, weights, g = load_golden(name)
    B, M = c["B"], c["M"]
    T = M + 2
    model = G.make_model(M, pc_start, weights)
    st = G.Stages(model, B)
    prev_aug, det_aug, f_prev, f_cur = _fill_aug(st, c, data, g)
    det = G.t(data["det_boxes"])
    st.project(det)
    torch.cuda.synchronize()
    # in-place back-projection of the caller's boxes
    assert np.array_equal(det.cpu().numpy(), g["det_boxes_after"])
    # projections vs float64 numpy
    W = {k: v.astype(np.float64) for k, v in weights.items()}
    fs, rc, fd = W["fuse_shape.0.weight"], W["res_coeff.0.weight"], W["fuse_det.0.weight"]
    pp = np.concatenate([f_prev @ fs[:, :320].T,
                         f_prev @ rc[:, :320].T + prev_aug[:, :, :3] @ rc[:, 320:323].T,
                         prev_aug[:, :, :3] @ fd[:, :3].T], axis=2)
    pcur = np.concatenate([f_cur @ fs[:, 320:].T + W["fuse_shape.0.bias"],
                           f_cur @ rc[:, 323:643].T + det_aug[:, :, :3] @ rc[:, 643:].T + W["res_coeff.0.bias"],
                           det_aug[:, :, :3] @ fd[:, 3:].T + W["fuse_det.0.bias"]], axis=2)
    got_pp = st.region(_cabi.WS_PROJ_PREV, (B, T, 144)).cpu().numpy()
    got_pc = st.region(_cabi.WS_PROJ_CUR, (B, 144, st.DP)).cpu().numpy()[:, :, :T].transpose(0, 2, 1)
    assert np.allclose(got_pp, pp, rtol=1e-4, atol=1e-5), np.abs(got_pp - pp).max()
    assert np.allclose(got_pc, pcur, rtol=1e-4, atol=1e-5), np.abs(got_pc - pcur).max()
    # column norm of the squared-distance matrix
    dist = ((prev_aug[:, :, None, :3].astype(np.float64) - det_aug[:, None, :, :3]) ** 2).sum(-1)
    cn = np.sqrt((dist ** 2).sum(axis=1))
    got_cn = st.region(_cabi.WS_COLNORM, (B, T)).cpu().numpy()
    assert np.allclose(got_cn, cn, rtol=1e-5), np.abs(got_cn - cn).max()
    # object-major copy of the current-frame projections (operand of the tcgen05 tiles)
    got_pct = st.region(_cabi.WS_PROJ_CUR_T, (B, T, 144)).cpu().numpy()
    assert np.array_equal(got_pct, got_pc)
    # pairwise -> residual: CUDA-core fp32, tcgen05 3xTF32 (fp32-equivalent), tcgen05 bf16 (separate tolerance)
    want = g["residual"]
    # (variant 4 = the warp-specialised pipelined tcgen05 kernel of pairwise_tc3.cu - max-form outer sum, seeded
    # accumulators - with the FFMA2 and the scalar-FFMA epilogue)
    variants = [(3, 1e-5, 0), (1, 2e-5, 0), (0, 2e-5, 0), (2, 2e-2, 0)]
    if M >= 15:
        variants += [(4, 2e-5, 0), (4, 2e-5, 2)]
    for variant, tol, ffma2 in variants:
        st.region(_cabi.WS_RESIDUAL, (B, T, st.RS)).zero_()
        _cabi.lib().shasta_set_option(_cabi.OPT_PAIR_FFMA2, ffma2)
        try:
            st.pairwise(variant)
        finally:
            _cabi.lib().shasta_set_option(_cabi.OPT_PAIR_FFMA2, 0)
        res = st.region(_cabi.WS_RESIDUAL, (B, T, st.RS)).cpu().numpy()[:, :, :T]
        err = np.abs(res - want).max() / np.abs(want).max()
        print("pairwise variant %d (0 default, 1 tf32x3 v2, 2 bf16, 3 ffma, 4 tf32x3 v3; ffma2 opt %d): residual max err / "
              "scale = %.3g" % (variant, ffma2, err))
        assert err < tol, "variant %d residual max err / scale = %g" % (variant, err)


@pytest.fixture(params=[1, 2], ids=["ffma", "tcgen05"])
def aff_path(request):
    lib = _cabi.lib()
    lib.shasta_set_option(_cabi.OPT_AFF_PATH, request.param)
    yield request.param
    lib.shasta_set_option(_cabi.OPT_AFF_PATH, 0)


@pytest.mark.parametrize("M,B", [(200, 7), (222, 2), (90, 3), (20, 9), (6, 1), (223, 2), (300, 3), (500, 1), (1000, 1)])
def test_aff_paths_agree(M, B):
    """tcgen05 3xTF32 row tiles vs the CUDA-core kernel on random residuals: partial last tile, K and N padding at
    D = 202 / 224 / 92 / 22 / 8 (staged variant) and D = 225 (one last-layer half of 240), 302 (halves 256 + 48),
    502 (256 + 256), 1002 (256 x 3 + 240; 16 K chunks) for the streamed variant."""
    lib = _cabi.lib()
    T = M + 2
    if M <= 222:
        model = G.make_model(M, (-4.8, -4.8), synthetic.make_weights(M, seed=3, peaky=50.0))
    else:   # the big aug_shape matrices make numpy-generated weights impractical: default init on the device
        torch.manual_seed(3)
        model = G.make_model(M, (-4.8, -4.8))
        with torch.no_grad():
            model.aff[10].weight.mul_(50.0)
    st = G.Stages(model, B)
    gen = torch.Generator(device="cpu").manual_seed(5)
    res = (torch.randn((B, T, T), generator=gen) * 4.0).to(G.DEV)
    out = {}
    for mode in (1, 2):
        lib.shasta_set_option(_cabi.OPT_AFF_PATH, mode)
        try:
            st.region(_cabi.WS_RESIDUAL, (B, T, st.RS)).zero_()
            st.region(_cabi.WS_RESIDUAL, (B, T, st.RS))[:, :, :T] = res
            m1, m2 = st.aff_softmax()
            out[mode] = (st.region(_cabi.WS_LOGITS, (B, T, st.RS))[:, :, :T].cpu().numpy().copy(), m1.cpu().numpy(),
                         m2.cpu().numpy())
        finally:
            lib.shasta_set_option(_cabi.OPT_AFF_PATH, 0)
    scale = np.abs(out[1][0]).max()
    err = np.abs(out[1][0] - out[2][0]).max() / scale
    print("aff tcgen05 vs CUDA cores: logits max err / scale = %.3g" % err)
    # two fp32 evaluations with different summation orders: the gap grows with the K = D of the first layer
    # (1.2e-5 measured at D = 1002)
    assert err < (1e-5 if M <= 222 else 3e-5)
    assert G.rel_err(out[2][1], out[1][1]) < FP32_REL_TOL
    assert G.rel_err(out[2][2], out[1][2]) < FP32_REL_TOL
    if M > 222:
        del st, model
        torch.cuda.empty_cache()


@pytest.mark.parametrize("name", golden_names())
def test_aff_softmax_stage(name, aff_path):
    c, pc_start, data, weights, g = load_golden(name)
    B, M = c["B"], c["M"]
    T = M + 2
    model = G.make_model(M, pc_start, weights)
    st = G.Stages(model, B)
    st.region(_cabi.WS_RESIDUAL, (B, T, st.RS))[:, :, :T] = G.t(g["residual"])
    m1, m2 = st.aff_softmax()
    logits = st.region(_cabi.WS_LOGITS, (B, T, st.RS)).cpu().numpy()[:, :, :T]
    scale = np.abs(g["logits"]).max()
    assert np.abs(logits - g["logits"]).max() / scale < 1e-5
    assert G.rel_err(m1.cpu().numpy(), g["matched1"]) < FP32_REL_TOL
    assert G.rel_err(m2.cpu().numpy(), g["matched2"]) < FP32_REL_TOL


# ------------------------------------------------------------------------------------------------
# whole path through the module interface
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_reference_golden_tcgen05_anchors(name, force_tc_anchors):
    test_forward_matches_reference_golden(name, 0)


@pytest.mark.parametrize("name", golden_names())
def test_forward_bf16_pairwise_tolerance(name):
    """bf16 mode of the pairwise tiles (flags 0x20): tolerance stated separately from fp32 (north_star) —
    affinities within 2e-2 relative; association agreement is reported, not required to be identical."""
    c, pc_start, data, weights, g = load_golden(name)
    model = G.make_model(c["M"], pc_start, weights)
    model.kernel_flags = 0x20
    with torch.no_grad():
        m1, m2 = model.affinity(G.t(data["bev"]), G.t(data["prev_bev"]), G.t(data["det_boxes"]),
                                G.t(data["prev_det_boxes"]))
    m1, m2 = m1.cpu().numpy(), m2.cpu().numpy()
    e1, e2 = G.rel_err(m1, g["matched1"]), G.rel_err(m2, g["matched2"])
    agree = float(np.mean(m1.argmax(2) == g["matched1"].argmax(2)))
    print("bf16 pairwise: rel err m1 %.3g m2 %.3g, row-argmax agreement %.4f" % (e1, e2, agree))
    if c["peaky"] == 0:
        assert e1 < 2e-2 and e2 < 2e-2, (e1, e2)
    assert np.isfinite(m1).all() and np.isfinite(m2).all()


@pytest.mark.parametrize("M,B", [(200, 12), (20, 70), (50, 130)])
def test_forward_bf16_mode_tolerance(M, B):
    """shasta_forward_bf16 (bf16 aug_shape.i.0 weights + features, bf16 pairwise tiles) against the fp32 path on the
    same inputs: bf16 tolerance 2e-2 relative on the affinities (stated separately from the fp32 bar, north_star);
    the agreement of the row-wise argmax is reported. B > 4 so that the tcgen05 bf16 GEMM is the one that runs;
    B = 70 / 130 exercise two batch tiles of 64 / 128."""
    H = W = 32
    pc_start = (-W * 0.3, -H * 0.3)
    model = G.make_model(M, pc_start, synthetic.make_weights(M, seed=21))
    d = synthetic.make_frame_pairs(B, M, H, W, 71, pc_start=pc_start)
    args = [G.t(d[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
    with torch.no_grad():
        r1, r2 = model.affinity(args[0], args[1], args[2].clone(), args[3])
        model.bf16 = True
        det = args[2].clone()
        b1, b2 = model.affinity(args[0], args[1], det, args[3])
        assert model._w16 is not None
        # the anchor rows / columns are what the bf16 GEMM feeds
        e1, e2 = G.rel_err(b1.cpu().numpy(), r1.cpu().numpy()), G.rel_err(b2.cpu().numpy(), r2.cpu().numpy())
        agree = float((b1.argmax(2) == r1.argmax(2)).float().mean())
    print("bf16 mode M=%d B=%d: rel err m1 %.3g m2 %.3g, row-argmax agreement %.4f" % (M, B, e1, e2, agree))
    assert e1 < 2e-2 and e2 < 2e-2
    assert torch.isfinite(b1).all() and torch.isfinite(b2).all()


@pytest.mark.parametrize("flags", [0, 1, 0x30])
@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_reference_golden(name, flags):
    c, pc_start, data, weights, g = load_golden(name)
    model = G.make_model(c["M"], pc_start, weights)
    model.kernel_flags = flags
    example = {"det_boxes": G.t(data["det_boxes"]), "prev_det_boxes": G.t(data["prev_det_boxes"]),
               "bev_feature": G.t(data["bev"]), "prev_bev_feature": G.t(data["prev_bev"])}
    with torch.no_grad():
        m1, m2, ex = model(example, train_mode=False)
    assert ex is example and tuple(m1.shape) == g["matched1"].shape and tuple(m2.shape) == g["matched2"].shape
    m1, m2 = m1.cpu().numpy(), m2.cpu().numpy()
    e1, e2 = G.rel_err(m1, g["matched1"]), G.rel_err(m2, g["matched2"])
    assert e1 < FP32_REL_TOL and e2 < FP32_REL_TOL, (e1, e2)
    # side effects of the reference forward
    assert np.array_equal(example["det_boxes"].cpu().numpy(), g["det_boxes_after"])
    for k in ("newborn", "fp", "dead_trk", "fn"):
        assert np.allclose(getattr(model, k).cpu().numpy(), g[k], rtol=2e-4, atol=2e-5)
    # association: decode identical to the decode of the reference's own outputs
    for b in range(c["B"]):
        n_prev, n_det = int(data["n_prev"][b]), int(data["n_det"][b])
        want = O.decode(torch.from_numpy(g["matched1"][b]), torch.from_numpy(g["matched2"][b]), n_prev, n_det)
        got = O.decode(torch.from_numpy(m1[b]), torch.from_numpy(m2[b]), n_prev, n_det)
        for key in ("dead", "fn", "keep_prev", "keep_dets", "newborn", "row_argmax", "col_argmax"):
            assert got[key] == want[key], (name, b, key)


def test_forward_batch_rows_independent():
    """Frame pairs are independent: a batch of 5 equals five batches of 1."""
    M, H, W = 20, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(5, M, H, W, 77, pc_start=pc_start)
    model = G.make_model(M, pc_start, synthetic.make_weights(M, seed=9))
    args = [G.t(data[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
    with torch.no_grad():
        m1, m2 = model.affinity(args[0], args[1], args[2].clone(), args[3])
        for b in range(5):
            s1, s2 = model.affinity(args[0][b:b + 1], args[1][b:b + 1], args[2][b:b + 1].clone(), args[3][b:b + 1])
            assert torch.allclose(s1[0], m1[b], rtol=1e-6, atol=1e-9)
            assert torch.allclose(s2[0], m2[b], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("M,B", [(20, 3), (200, 12)])
def test_cuda_graph_replay_matches_eager(M, B):
    """cuda_graphs=True captures the launch sequence once and replays it: results (and the in-place back-projection)
    are bit-identical to the eager launches, also after the input buffers were refilled in place."""
    H = W = 32
    pc_start = (-W * 0.3, -H * 0.3)
    model = G.make_model(M, pc_start, synthetic.make_weights(M, seed=5))
    d0 = synthetic.make_frame_pairs(B, M, H, W, 31, pc_start=pc_start)
    d1 = synthetic.make_frame_pairs(B, M, H, W, 32, pc_start=pc_start)
    bufs = [G.t(d0[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
    with torch.no_grad():
        for rep, d in enumerate((d0, d1, d0)):
            src = [G.t(d[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
            model.cuda_graphs = False
            det_e = src[2].clone()
            e1, e2 = model.affinity(src[0], src[1], det_e, src[3])
            model.cuda_graphs = True
            for dst, s_ in zip(bufs, src):
                dst.copy_(s_)
            g1, g2 = model.affinity(*bufs)
            assert torch.equal(g1, e1) and torch.equal(g2, e2), rep
            assert torch.equal(bufs[2], det_e), rep
    assert len(model._graphs) == 1


def test_internal_side_stream_changes_nothing():
    """The box-only kernels run on a library-owned stream next to the anchors GEMM (fork / join inside
    shasta_forward_f32); flag 0x400 keeps everything on the caller's stream. Same kernels, same results."""
    M, B, H, W = 200, 12, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    model = G.make_model(M, pc_start, synthetic.make_weights(M, seed=8))
    d = synthetic.make_frame_pairs(B, M, H, W, 51, pc_start=pc_start)
    args = [G.t(d[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
    with torch.no_grad():
        out = {}
        for flags in (0, 0x400):
            model.kernel_flags = flags
            det = args[2].clone()
            for _ in range(3):   # repeated calls reuse the side stream and its events
                det.copy_(args[2])
                m1, m2 = model.affinity(args[0], args[1], det, args[3])
            out[flags] = (m1.clone(), m2.clone(), det.clone())
    for a, b in zip(out[0], out[0x400]):
        assert torch.equal(a, b)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_second_device_in_same_process():
    """Kernel attributes (dynamic shared memory, carve-out) are per device: a model on cuda:1 after one on cuda:0."""
    M, B, H, W = 20, 9, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    w = synthetic.make_weights(M, seed=12)
    d = synthetic.make_frame_pairs(B, M, H, W, 61, pc_start=pc_start)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        model = G.make_model(M, pc_start, w, device=dev)
        args = [G.t(d[k], device=dev) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
        with torch.no_grad(), torch.cuda.device(dev):
            m1, m2 = model.affinity(*args)
            outs.append((m1.cpu(), m2.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_pipelined_host_inputs_match_device_path():
    """Pinned host inputs run the gather on a side stream into alternating workspaces (two-stage pipeline over calls):
    every call must return what the plain device path returns, including the in-place back-projection."""
    M, B, H, W = 20, 3, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    model = G.make_model(M, pc_start, synthetic.make_weights(M, seed=6))
    outs = []
    with torch.no_grad():
        for seed in (41, 42, 43, 44, 45):
            d = synthetic.make_frame_pairs(B, M, H, W, seed, pc_start=pc_start)
            host = [torch.from_numpy(d[k]).pin_memory() for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
            m1, m2 = model.affinity(*host)          # asynchronous: do not synchronise between the calls
            outs.append((d, host, m1, m2))
        torch.cuda.synchronize()
        assert model._pipe is not None
        model.pipeline_host_inputs = False
        for d, host, m1, m2 in outs:
            dev = [G.t(d[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
            e1, e2 = model.affinity(*dev)
            assert torch.equal(m1, e1) and torch.equal(m2, e2)
            assert torch.equal(host[2], dev[2].cpu())


def test_forward_empty_batch_and_bad_inputs():
    M = 6
    model = G.make_model(M, (-4.8, -4.8), synthetic.make_weights(M, seed=1))
    z = lambda *s: torch.zeros(*s, device=G.DEV)  # noqa: E731
    m1, m2 = model.affinity(z(0, 16, 16, 64), z(0, 16, 16, 64), z(0, M, 11), z(0, M, 11))
    assert tuple(m1.shape) == (0, M, M + 2) and tuple(m2.shape) == (0, M + 2, M)
    with pytest.raises(_cabi.ShastaLibraryError):
        model.affinity(torch.zeros(1, 16, 16, 64), torch.zeros(1, 16, 16, 64), torch.zeros(1, M, 11),
                       torch.zeros(1, M, 11))
    with pytest.raises(ValueError):
        model.affinity(z(1, 16, 16, 32), z(1, 16, 16, 32), z(1, M, 11), z(1, M, 11))
    with pytest.raises(ValueError):
        model.affinity(z(1, 16, 16, 64), z(1, 16, 16, 64), z(1, M + 1, 11), z(1, M + 1, 11))


def test_all_zero_boxes_and_out_of_range_boxes():
    """Empty frames (all padding) and boxes far outside the map stay finite and normalised."""
    M, H, W = 20, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(2, M, H, W, 5, pc_start=pc_start)
    data["det_boxes"][0] = 0.0
    data["prev_det_boxes"][1] = 0.0
    data["det_boxes"][1, :4, :2] = [[1e4, 1e4], [-1e4, 3.0], [W * 0.3, H * 0.3], [-W * 0.3, -H * 0.3]]
    weights = synthetic.make_weights(M, seed=3)
    model = G.make_model(M, pc_start, weights)
    det = G.t(data["det_boxes"])
    with torch.no_grad():
        m1, m2 = model.affinity(G.t(data["bev"]), G.t(data["prev_bev"]), det, G.t(data["prev_det_boxes"]))
    w = O.weights_to_torch(weights)
    o1, o2 = O.forward(w, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]),
                       torch.from_numpy(data["det_boxes"].copy()), torch.from_numpy(data["prev_det_boxes"]),
                       pc_start=pc_start)
    assert torch.isfinite(m1).all() and torch.isfinite(m2).all()
    assert G.rel_err(m1.cpu().numpy(), o1.numpy()) < FP32_REL_TOL
    assert G.rel_err(m2.cpu().numpy(), o2.numpy()) < FP32_REL_TOL


# ------------------------------------------------------------------------------------------------
# headline size: M = 200 (BASELINE.json configs[0]/[1]) against the oracle, plus size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,B,flags,anchor", [(180, 180, 3, 0, 0), (512, 512, 1, 0x30, 0), (180, 180, 2, 0x10, 2),
                                                (512, 512, 5, 0x40, 2)])
def test_headline_size_against_oracle(H, W, B, flags, anchor):
    """anchor = 2 forces the tcgen05 anchors GEMM; flags 0 / 0x10 = persistent tcgen05 3xTF32 pairwise tiles (the
    default), 0x40 = the warp-specialised pipelined tcgen05 kernel, 0x30 = CUDA-core pairwise tiles: every fp32-mode combination must reproduce the oracle's association exactly."""
    M = 200
    _cabi.lib().shasta_set_option(_cabi.OPT_ANCHOR_PATH, anchor)
    try:
        _headline(M, H, W, B, flags)
    finally:
        _cabi.lib().shasta_set_option(_cabi.OPT_ANCHOR_PATH, 0)


def _headline(M, H, W, B, flags):
    pc_start = (-W * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(B, M, H, W, 2024 + H, pc_start=pc_start)
    weights = synthetic.make_weights(M, seed=21, peaky=300.0)
    model = G.make_model(M, pc_start, weights)
    model.kernel_flags = flags
    det = G.t(data["det_boxes"])
    with torch.no_grad():
        m1, m2 = model.affinity(G.t(data["bev"]), G.t(data["prev_bev"]), det, G.t(data["prev_det_boxes"]))
    w = O.weights_to_torch(weights)
    det_o = torch.from_numpy(data["det_boxes"].copy())
    o1, o2 = O.forward(w, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]), det_o,
                       torch.from_numpy(data["prev_det_boxes"]), pc_start=pc_start)
    m1, m2 = m1.cpu(), m2.cpu()
    assert G.rel_err(m1.numpy(), o1.numpy()) < 1e-3
    assert G.rel_err(m2.numpy(), o2.numpy()) < 1e-3
    assert np.array_equal(det.cpu().numpy(), det_o.numpy())
    assert torch.allclose(m1.sum(2), torch.ones(B, M), atol=1e-5)
    assert torch.allclose(m2.sum(1), torch.ones(B, M), atol=1e-5)
    for b in range(B):
        n_prev, n_det = int(data["n_prev"][b]), int(data["n_det"][b])
        want = O.decode(o1[b], o2[b], n_prev, n_det)
        got = O.decode(m1[b], m2[b], n_prev, n_det)
        for key in ("dead", "fn", "keep_prev", "keep_dets", "newborn", "row_argmax", "col_argmax"):
            assert got[key] == want[key], (b, key)


@pytest.mark.parametrize("M,B", [(500, 2), (1000, 1)])
def test_large_sizes_properties(M, B):
    """BASELINE.json configs[2]/[3] sizes (up to 500x500, 1000x1000 stress): size-independent properties —
    normalisation, finiteness, independence of the batch composition, determinism."""
    H = W = 64
    pc_start = (-W * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(B, M, H, W, 31 + M, pc_start=pc_start)
    model = G.make_model(M, pc_start)  # default nn.Linear init on the device (weights are 6.4 / 25.6 GB)
    args = [G.t(data[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
    with torch.no_grad():
        m1, m2 = model.affinity(args[0], args[1], args[2].clone(), args[3])
        r1, r2 = model.affinity(args[0], args[1], args[2].clone(), args[3])
    assert torch.isfinite(m1).all() and torch.isfinite(m2).all()
    assert torch.allclose(m1.sum(2), torch.ones(B, M, device=G.DEV), atol=1e-5)
    assert torch.allclose(m2.sum(1), torch.ones(B, M, device=G.DEV), atol=1e-5)
    assert torch.equal(m1, r1) and torch.equal(m2, r2), "the path must be bit-reproducible run to run"
    del model
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# a13: device decode
# ------------------------------------------------------------------------------------------------
def _device_decode(m1, m2, n_prev, n_det):
    lib = _cabi.lib()
    B, M = m1.shape[0], m1.shape[1]
    i32 = lambda: torch.empty((B, M), dtype=torch.int32, device=G.DEV)  # noqa: E731
    f32 = lambda: torch.empty((B, M), dtype=torch.float32, device=G.DEV)  # noqa: E731
    ps, pa, fs, ds, da, dsc = i32(), i32(), f32(), i32(), i32(), f32()
    npv = torch.tensor(n_prev, dtype=torch.int32, device=G.DEV)
    ndv = torch.tensor(n_det, dtype=torch.int32, device=G.DEV)
    rc = lib.shasta_decode_f32(m1.data_ptr(), m2.data_ptr(), npv.data_ptr(), ndv.data_ptr(), B, M, ps.data_ptr(),
                               pa.data_ptr(), fs.data_ptr(), ds.data_ptr(), da.data_ptr(), dsc.data_ptr(), G.stream())
    _cabi.check(rc, "decode")
    return [x.cpu().numpy() for x in (ps, pa, fs, ds, da, dsc)]


def _check_decode(m1, m2, n_prev, n_det):
    ps, pa, fs, ds, da, dsc = _device_decode(G.t(m1), G.t(m2), n_prev, n_det)
    for b in range(m1.shape[0]):
        want = O.decode(torch.from_numpy(m1[b]), torch.from_numpy(m2[b]), n_prev[b], n_det[b])
        assert [n for n in range(n_prev[b]) if ps[b, n] == 1] == want["dead"]
        assert [n for n in range(n_prev[b]) if ps[b, n] == 2] == want["fn"]
        assert [n for n in range(n_prev[b]) if ps[b, n] == 0] == want["keep_prev"]
        assert np.all(ps[b, n_prev[b]:] == -1) and np.all(ds[b, n_det[b]:] == -1)
        assert pa[b, :n_prev[b]].tolist() == want["row_argmax"]
        assert da[b, :n_det[b]].tolist() == want["col_argmax"]
        # the kernel stores the raw matched values; ref_detection_score = 1 - value in double must equal the oracle's
        assert (1.0 - fs[b, want["fn"]].astype(np.float64)).tolist() == want["fn_score"]
        keep = [k for k in range(n_det[b]) if ds[b, k] in (0, 1)]
        assert keep == want["keep_dets"]
        assert [bool(ds[b, k] == 1) for k in keep] == want["newborn"]
        assert (1.0 - dsc[b, keep].astype(np.float64)).tolist() == want["det_score"]


def test_decode_on_planted_matrices():
    rng = np.random.default_rng(5)
    B, M = 6, 40
    m1 = rng.uniform(0, 0.45, (B, M, M + 2)).astype(np.float32)
    m2 = rng.uniform(0, 0.45, (B, M + 2, M)).astype(np.float32)
    n_prev = [40, 17, 0, 25, 1, 33]
    n_det = [40, 22, 9, 0, 1, 30]
    for b in range(B):
        for n in range(n_prev[b]):
            r = rng.integers(0, 5)
            if r == 0:
                m1[b, n, M] = 0.9                      # dead
            elif r == 1:
                m1[b, n, M + 1] = 0.8                  # false negative
            elif r == 2 and n_det[b] > 0:
                m1[b, n, rng.integers(0, n_det[b])] = 0.95
            elif r == 3:
                m1[b, n, M + 1] = 0.5                  # exactly at the threshold: not > 0.5
        for k in range(n_det[b]):
            r = rng.integers(0, 5)
            if r == 0:
                m2[b, M + 1, k] = 0.75                 # false positive, dropped
            elif r == 1:
                m2[b, M + 1, k] = 0.7                  # float32(0.7) < 0.7: kept
            elif r == 2:
                m2[b, M, k] = 0.6                      # newborn
            elif r == 3 and n_prev[b] > 0:
                m2[b, rng.integers(0, n_prev[b]), k] = 0.99
    # ties: first maximum wins
    m1[0, 0, :] = 0.3
    m2[0, :, 0] = 0.3
    _check_decode(m1, m2, n_prev, n_det)


def test_decode_large_rows_compaction():
    """More previous objects than one 256-row compaction pass; kept rows interleaved with dead / FN ones."""
    rng = np.random.default_rng(9)
    B, M = 3, 600
    m1 = rng.uniform(0, 0.45, (B, M, M + 2)).astype(np.float32)
    m2 = rng.uniform(0, 0.45, (B, M + 2, M)).astype(np.float32)
    n_prev, n_det = [600, 257, 511], [600, 300, 64]
    for b in range(B):
        for n in range(n_prev[b]):
            r = rng.integers(0, 4)
            if r == 0:
                m1[b, n, M] = 0.9
            elif r == 1:
                m1[b, n, M + 1] = 0.8
        for k in range(n_det[b]):
            if rng.integers(0, 3) == 0:
                m2[b, rng.integers(0, n_prev[b]), k] = 0.99
    _check_decode(m1, m2, n_prev, n_det)


def test_decode_on_peaky_golden():
    c, pc_start, data, weights, g = load_golden("m20_32px_b3_peaky")
    _check_decode(g["matched1"], g["matched2"], [int(x) for x in data["n_prev"]], [int(x) for x in data["n_det"]])


# ------------------------------------------------------------------------------------------------
# decode fused into the softmax epilogues (shasta_forward_decode_f32) == decode kernel on the same outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,B,aff_path", [(20, 3, 0), (50, 4, 0), (200, 5, 0), (200, 2, 1), (300, 2, 0), (500, 1, 0)])
def test_fused_decode_equals_decode_kernel(M, B, aff_path):
    """aff_path 0 = tcgen05 aff kernels (staged up to D = 224, streamed + row_softmax_kernel beyond), 1 = CUDA-core
    aff_row_kernel. Every plane of the fused block must equal shasta_decode_f32 on the returned matched1 / matched2
    bit for bit; the ring slot follows the device counter, also under CUDA-graph replay."""
    H = W = 48
    pc_start = (-W * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(B, M, H, W, 900 + M, pc_start=pc_start)
    torch.manual_seed(M)
    model = G.make_model(M, pc_start) if M > 200 else G.make_model(M, pc_start, synthetic.make_weights(M, seed=M, peaky=300.0))
    if M > 200:
        with torch.no_grad():
            model.aff[10].weight.mul_(300.0)
    lib = _cabi.lib()
    lib.shasta_set_option(_cabi.OPT_AFF_PATH, aff_path)
    try:
        n_prev = torch.from_numpy(data["n_prev"].astype(np.int32)).to(G.DEV)
        n_det = torch.from_numpy(data["n_det"].astype(np.int32)).to(G.DEV)
        ring = torch.full((3, 6, B, M), -77, dtype=torch.int32, device=G.DEV)
        counter = torch.zeros(1, dtype=torch.int32, device=G.DEV)
        args = [G.t(data[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
        model.cuda_graphs = True
        decisions = 0
        with torch.no_grad():
            for call in range(4):      # call 0 captures (eager run + capture run + replay: three forwards), 1-3 replay
                m1, m2 = model.affinity(args[0], args[1], args[2].clone() if call else args[2], args[3],
                                        decode={"n_prev": n_prev, "n_det": n_det, "out": ring, "counter": counter})
                torch.cuda.synchronize()
                slot = (int(counter.item()) - 1) % 3
                want = model.decode(m1, m2, n_prev, n_det)
                got = ring[slot]
                for i, k in enumerate(("prev_state", "prev_argmax", "fn_dead_prob", "det_state", "det_argmax", "det_fp_prob")):
                    w = want[k].view(torch.int32) if k.endswith("prob") else want[k]
                    assert torch.equal(got[i], w), (call, k)
                decisions += int((want["prev_state"] > 0).sum() + (want["det_state"] > 0).sum())
        assert int(counter.item()) >= 4
        print("fused decode M=%d B=%d aff_path=%d: %d dead / FN / newborn / FP decisions compared" % (M, B, aff_path, decisions))
    finally:
        lib.shasta_set_option(_cabi.OPT_AFF_PATH, 0)
