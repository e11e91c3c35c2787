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
    # the producer left of the path is not differentiated in this revision
    assert sd["shared_conv.0.weight"].grad is None


def test_training_steps_lower_the_loss():
    c, pc_start, data, weights, g = load_golden("m20_32px_b2")
    B, M = c["B"], c["M"]
    gt = G.t(_gt(B, M, data["n_prev"], data["n_det"], seed=3))
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
