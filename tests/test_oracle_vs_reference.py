"""CPU, authoring container only: the oracle restatement against the LIVE unmodified reference on fresh random
cases (skipped where /root/reference does not exist, e.g. the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import shasta_oracle as O
from shasta_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("M,H,W,B,seed", [(7, 12, 20, 2, 101), (33, 40, 40, 1, 102), (90, 64, 64, 1, 103)])
def test_oracle_equals_live_reference(M, H, W, B, seed):
    torch.set_num_threads(1)
    pc_start = (-W * 0.6 / 2.0, -H * 0.6 / 2.0)
    data = synthetic.make_frame_pairs(B, M, H, W, seed, pc_start=pc_start)
    weights = synthetic.make_weights(M, seed=seed)
    model = ref_loader.build_reference_head(M, 3, pc_start=pc_start)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
    det_ref = torch.from_numpy(data["det_boxes"].copy())
    m1r, m2r, ex = ref_loader.run_reference(model, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]),
                                           det_ref, torch.from_numpy(data["prev_det_boxes"]))
    det = torch.from_numpy(data["det_boxes"].copy())
    m1, m2 = O.forward(O.weights_to_torch(weights), torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]),
                       det, torch.from_numpy(data["prev_det_boxes"]), pc_start=pc_start)
    assert np.array_equal(m1.numpy(), m1r.numpy())
    assert np.array_equal(m2.numpy(), m2r.numpy())
    assert np.array_equal(det.numpy(), det_ref.numpy())


def test_state_dict_names_match_reference():
    model = ref_loader.build_reference_head(20, 3)
    ref_keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    ours = {k: tuple(v) for k, v in synthetic.head_param_shapes(20).items()}
    assert ours == ref_keys  # shared_conv was replaced by the parameter-free stand-in in the loader
