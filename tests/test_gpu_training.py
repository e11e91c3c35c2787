"""GPU: training configuration (BASELINE.json config 5, tools/nusc_shasta/train.py:195-215) — the CUDA backward against
torch.autograd through the CPU oracle (float64) on the same seeded inputs, weights and ground-truth matrix."""
import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from shasta_b200 import loss as L
from shasta_b200 import training
from tests import gpu_util as G
from tests.golden_util import load_golden

pytestmark = pytest.mark.gpu


def _gt(B, M, n_prev, n_det, seed):
    """Augmented ground-truth affinity matrix like det3d/datasets/nuscenes/nuscenes.py:297-349: one 1 per real
    previous row (a detection, 'dead' or 'FN' column) and per real detection column (a row, 'newborn' or 'FP')."""
    rng = np.random.default_rng(seed)
    gt = np.zeros((B, M + 2, M + 2), np.float32)
    for b in range(B):
        for t in range(int(n_prev[b])):
            gt[b, t, rng.choice([rng.integers(0, max(1, int(n_det[b]))), M, M + 1])] = 1.0
        for d in range(int(n_det[b])):
            if gt[b, :, d].sum() == 0:
                gt[b, rng.choice([M, M + 1]), d] = 1.0
    return gt


def _gt_from_builder(B, M, n_prev, n_det, seed):
    """Targets built the way the reference's dataset builds them: a per-frame label (matched (N, K+2), newborn (K),
    preprocessing/make_gt_shasta.py via formats.label_affinity) pushed through formats.build_gt_affinity
    (det3d/datasets/nuscenes/nuscenes.py:297-349: dead-track / FP sub-sampling, compaction to the padded layout)."""
    import random

    from shasta_b200 import formats as F
    rng = np.random.default_rng(seed)
    out = np.zeros((B, M + 2, M + 2), np.float32)
    for b in range(B):
        N, K = int(n_prev[b]), int(n_det[b])
        # a synthetic association: instance ids for the GT boxes of both frames, detections as TPs / FPs
        prev_gt_ids = list(range(N + 2))
        prev_tp = {d: d for d in range(N) if rng.random() < 0.8}
        cur_gt_ids = [i for i in prev_gt_ids if rng.random() < 0.85] + [1000 + i for i in range(3)]
        cur_tp = {}
        free = list(range(len(cur_gt_ids)))
        rng.shuffle(free)
        for d in range(K):
            if free and rng.random() < 0.8:
                cur_tp[d] = free.pop()
        fn = [g for g in range(len(cur_gt_ids)) if g not in cur_tp.values()]
        matched, newborn = F.label_affinity(cur_tp, cur_gt_ids, fn, K, prev=(prev_tp, prev_gt_ids, N))
        gt, _, _ = F.build_gt_affinity(matched, newborn, list(range(N)), list(range(K)), M, True, fp_ratio=0.5,
                                       dead_trk_ratio=0.5, rng=random.Random(seed + b))
        out[b] = gt
    return out


def _oracle_grads(weights, data, pc_start, gt, names):
    w = {k: torch.from_numpy(v).double() for k, v in weights.items()}
    for n in names:
        w[n].requires_grad_(True)
    m1, m2 = O.forward(w, torch.from_numpy(data["bev"]).double(), torch.from_numpy(data["prev_bev"]).double(),
                       torch.from_numpy(data["det_boxes"].copy()).double(),
                       torch.from_numpy(data["prev_det_boxes"]).double(), pc_start=pc_start)
    loss = O.affinity_loss(m1, m2, torch.from_numpy(gt).double())
    loss.backward()
    return float(loss), {n: w[n].grad.numpy() for n in names}


@pytest.mark.parametrize("name", ["m6_16px_b2", "m20_32px_b2", "m50_48x40_b2_peaky"])
def test_gradients_match_autograd_through_oracle(name):
    c, pc_start, data, weights, g = load_golden(name)
    B, M = c["B"], c["M"]
    gt = _gt(B, M, data["n_prev"], data["n_det"], seed=c["seed"])
    names = training.differentiable_parameter_names()
    want_loss, want = _oracle_grads(weights, data, pc_start, gt, names)

    model = G.make_model(M, pc_start, weights)
    model.train()
    example = {"det_boxes": G.t(data["det_boxes"]), "prev_det_boxes": G.t(data["prev_det_boxes"]),
               "bev_feature": G.t(data["bev"]), "prev_bev_feature": G.t(data["prev_bev"])}
    m1, m2, _ = model(example, train_mode=True)
    assert m1.requires_grad and m2.requires_grad
    # the forward of the training path is the same kernels: same numbers as inference
    assert G.rel_err(m1.detach().cpu().numpy(), g["matched1"]) < 2e-4
    loss = L.affinity_loss(m1, m2, G.t(gt))
    loss.backward()
    assert abs(float(loss) - want_loss) < 1e-4 * max(1.0, abs(want_loss))
    sd = dict(model.named_parameters())
    for n in names:
        got = sd[n].grad.detach().cpu().numpy().astype(np.float64)
        scale = np.abs(want[n]).max() + 1e-12
        err = np.abs(got - want[n]).max() / scale
        print("%-22s grad max err / scale = %.3g (scale %.3g)" % (n, err, scale))
        assert err < 2e-3, "%s: grad max err / scale = %g" % (n, err)
    # the maps of this test are leaves without requires_grad: nothing flows to the producer
    assert sd["shared_conv.0.weight"].grad is None


@pytest.mark.parametrize("name", ["m6_16px_b2", "m20_32px_b2", "m50_48x40_b2_peaky"])
def test_map_gradients_match_autograd_through_oracle(name):
    """d loss / d bev_feature and d prev_bev_feature (what trains shared_conv, train.py:184-191) against float64
    autograd through the oracle: the gather's bilinear scatter, the first-layer path and the aug_shape.i.0 path."""
    c, pc_start, data, weights, g = load_golden(name)
    B, M = c["B"], c["M"]
    gt = _gt(B, M, data["n_prev"], data["n_det"], seed=c["seed"])
    w = {k: torch.from_numpy(v).double() for k, v in weights.items()}
    bev64 = torch.from_numpy(data["bev"]).double().requires_grad_(True)
    prev64 = torch.from_numpy(data["prev_bev"]).double().requires_grad_(True)
    m1, m2 = O.forward(w, bev64, prev64, torch.from_numpy(data["det_boxes"].copy()).double(),
                       torch.from_numpy(data["prev_det_boxes"]).double(), pc_start=pc_start)
    O.affinity_loss(m1, m2, torch.from_numpy(gt).double()).backward()

    model = G.make_model(M, pc_start, weights)
    model.train()
    bev = G.t(data["bev"]).requires_grad_(True)
    prev = G.t(data["prev_bev"]).requires_grad_(True)
    example = {"det_boxes": G.t(data["det_boxes"]), "prev_det_boxes": G.t(data["prev_det_boxes"]),
               "bev_feature": bev, "prev_bev_feature": prev}
    a1, a2, _ = model(example, train_mode=True)
    L.affinity_loss(a1, a2, G.t(gt)).backward()
    for got, want, nm in ((bev.grad, bev64.grad, "bev_feature"), (prev.grad, prev64.grad, "prev_bev_feature")):
        assert got is not None, nm
        got = got.cpu().numpy().astype(np.float64)
        want = want.numpy()
        scale = np.abs(want).max() + 1e-30
        err = np.abs(got - want).max() / scale
        nz = float((want != 0).mean())
        print("%-18s grad max err / scale = %.3g (scale %.3g, %.1f%% of the map touched)" % (nm, err, scale, 100 * nz))
        assert err < 2e-3, "%s: grad max err / scale = %g" % (nm, err)


def test_shared_conv_receives_gradients_in_train_mode():
    """End to end like train.py:195-215 with a trunk stub: (B,512,H,W) maps -> shared_conv (autograd, train-mode
    BatchNorm) -> CUDA head -> loss.backward(): shared_conv.0.weight / shared_conv.1.weight get gradients that match
    the same graph with the head replaced by the float64 oracle."""
    c, pc_start, data, weights, g = load_golden("m6_16px_b2")
    B, M, H, W = c["B"], c["M"], c["H"], c["W"]
    gt = _gt(B, M, data["n_prev"], data["n_det"], seed=5)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn((B, 512, H, W), generator=gen) * 0.1
    xp = torch.randn((B, 512, H, W), generator=gen) * 0.1
    model = G.make_model(M, pc_start, weights)
    model.train()
    torch.backends.cudnn.allow_tf32 = False      # fp32 convolution: the comparison below is against float64
    model.extract_feat = lambda ex: (G.t(x.numpy()), None, G.t(xp.numpy()), None)
    example = {"det_boxes": G.t(data["det_boxes"]), "prev_det_boxes": G.t(data["prev_det_boxes"])}
    m1, m2, _ = model(example, train_mode=True)
    L.affinity_loss(m1, m2, G.t(gt)).backward()
    got = {k: v.grad.detach().cpu().double() for k, v in model.shared_conv.named_parameters()}
    assert all(v is not None and torch.isfinite(v).all() for v in got.values())

    # reference graph on the CPU in float64: same shared_conv weights, train-mode BN, oracle head
    import copy
    conv = copy.deepcopy(model.shared_conv).cpu().double()
    conv.train()
    for p_ in conv.parameters():
        p_.grad = None
    w = {k: torch.from_numpy(v).double() for k, v in weights.items()}
    bev = conv(x.double()).permute(0, 2, 3, 1).contiguous()
    prev_bev = conv(xp.double()).permute(0, 2, 3, 1).contiguous()
    o1, o2 = O.forward(w, bev, prev_bev, torch.from_numpy(data["det_boxes"].copy()).double(),
                       torch.from_numpy(data["prev_det_boxes"]).double(), pc_start=pc_start)
    O.affinity_loss(o1, o2, torch.from_numpy(gt).double()).backward()
    top = max(float(p_.grad.abs().max()) for p_ in conv.parameters())
    for k, p_ in conv.named_parameters():
        # (the convolution bias has an exactly-zero gradient under train-mode BatchNorm - fp32 leaves rounding noise
        # there: the scale is floored at 1e-3 of the largest gradient)
        scale = max(float(p_.grad.abs().max()), 1e-3 * top)
        err = float((got[k] - p_.grad).abs().max()) / scale
        print("shared_conv.%-10s grad max err / scale = %.3g (scale %.3g)" % (k, err, scale))
        assert err < 5e-3, (k, err)


def test_training_steps_lower_the_loss():
    c, pc_start, data, weights, g = load_golden("m20_32px_b2")
    B, M = c["B"], c["M"]
    gt_np = _gt_from_builder(B, M, data["n_prev"], data["n_det"], seed=3)   # targets from the dataset-side builder
    assert gt_np[:, :-2, :].sum() > 0 and gt_np[:, :, :-2].sum() > 0
    gt = G.t(gt_np)
    model = G.make_model(M, pc_start, weights)
    model.train()
    opt = torch.optim.Adam(training.differentiable_parameters(model), lr=1e-3)
    args = [G.t(data[k]) for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes")]
    losses = []
    for _ in range(5):
        opt.zero_grad()
        m1, m2 = model.affinity(args[0], args[1], args[2].clone(), args[3])
        loss = L.affinity_loss(m1, m2, gt)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    # the kernel-side weight cache must follow the optimizer's in-place updates
    model.eval()
    with torch.no_grad():
        e1, _ = model.affinity(args[0], args[1], args[2].clone(), args[3])
    w = {k: v.detach().cpu() for k, v in model.state_dict().items() if not k.startswith("shared_conv")}
    o1, _ = O.forward(w, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]),
                      torch.from_numpy(data["det_boxes"].copy()), torch.from_numpy(data["prev_det_boxes"]),
                      pc_start=pc_start)
    assert G.rel_err(e1.cpu().numpy(), o1.numpy()) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 4096 + 3, 1 << 20])
def test_stream_adam_matches_torch_adam(n):
    """shasta_adam_step_f32 (training.StreamAdam) against torch.optim.Adam on the same parameters and gradients
    (train.py:146: lr, weight_decay): three steps, parameters and both moments to 1e-6 relative."""
    from shasta_b200 import training
    gen = torch.Generator(device="cpu").manual_seed(n)
    p0 = torch.randn(n, generator=gen)
    grads = [torch.randn(n, generator=gen) * (10.0 ** (k - 1)) for k in range(3)]
    kw = dict(lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8)
    pa = torch.nn.Parameter(p0.clone().to(G.DEV))
    pb = torch.nn.Parameter(p0.clone().to(G.DEV))
    oa = torch.optim.Adam([pa], **kw)
    ob = training.StreamAdam([pb], **kw)
    for g in grads:
        pa.grad = g.to(G.DEV)
        pb.grad = g.to(G.DEV)
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    sa, sb = oa.state[pa], ob.state[pb]
    for name, a, b in (("param", pa.data, pb.data), ("exp_avg", sa["exp_avg"], sb["exp_avg"]),
                       ("exp_avg_sq", sa["exp_avg_sq"], sb["exp_avg_sq"])):
        a, b = a.cpu().double(), b.cpu().double()
        # relative, floored at 1e-4 of the tensor's scale and of the quantity's natural scale (gradients of ~0.1 / 1 / 10:
        # moments of order 1 and 0.1; a parameter can be left arbitrarily close to zero by an update of ~lr, where one ulp
        # of the update is a large relative error - and with n = 1 the tensor's own maximum is that element)
        floor = max(1e-4 * float(a.abs().max()), {"param": 1e-3, "exp_avg": 1e-4, "exp_avg_sq": 1e-5}[name])
        err = float(((a - b).abs() / a.abs().clamp_min(floor)).max())
        assert err < 5e-6, (name, err)
    assert int(sb["step"]) == 3
