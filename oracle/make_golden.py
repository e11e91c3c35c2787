"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/det3d/models/tracker/shasta.py) on seeded synthetic frame pairs.

Run in the authoring container only (the reference tree does not travel):
    python -m oracle.make_golden
Inputs and weights are regenerated from (seed, shape) by shasta_b200.synthetic, so a fixture stores the
case description, an input checksum and the reference's outputs/intermediates.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader  # noqa: E402
from shasta_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> case description. pixel = voxel*stride = 0.6 m everywhere; pc_start = -H*0.6/2.
CASES = {
    "m6_16px_b2": dict(M=6, H=16, W=16, B=2, seed=11, wseed=1, peaky=0.0),
    "m20_32px_b2": dict(M=20, H=32, W=32, B=2, seed=12, wseed=2, peaky=0.0),
    "m20_32px_b3_peaky": dict(M=20, H=32, W=32, B=3, seed=13, wseed=3, peaky=400.0),
    "m20_180px_b1": dict(M=20, H=180, W=180, B=1, seed=14, wseed=4, peaky=0.0),
    "m50_48x40_b2_peaky": dict(M=50, H=48, W=40, B=2, seed=15, wseed=5, peaky=400.0),
}


# headline-size fixtures (BASELINE.json configs[0]/[1] shape, M = 200): outputs only - the intermediates of a 200 x 200
# case would be tens of MB. File names start with "h" so that the per-stage tests (which need intermediates) skip them.
HEADLINE_CASES = {
    "h200_180px_b1_peaky": dict(M=200, H=180, W=180, B=1, seed=16, wseed=6, peaky=400.0),
}


def case_inputs(c):
    pc_start = (-c["W"] * 0.6 / 2.0, -c["H"] * 0.6 / 2.0)
    data = synthetic.make_frame_pairs(c["B"], c["M"], c["H"], c["W"], c["seed"], pc_start=pc_start)
    weights = synthetic.make_weights(c["M"], seed=c["wseed"], peaky=c["peaky"])
    return pc_start, data, weights


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)  # single-thread: summation order independent of the host's core count
    for name, c in CASES.items():
        pc_start, data, weights = case_inputs(c)
        model = ref_loader.build_reference_head(c["M"], 3, pc_start=pc_start)
        sd = model.state_dict()
        for k, v in weights.items():
            assert tuple(sd[k].shape) == v.shape, (k, sd[k].shape, v.shape)
        missing = [k for k in sd if k not in weights and not k.startswith(("shared_conv",))]
        assert not missing, missing
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
        det = torch.from_numpy(data["det_boxes"].copy())
        prev = torch.from_numpy(data["prev_det_boxes"].copy())
        m1, m2, ex, inter = ref_loader.run_reference(
            model, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]), det, prev, capture=True)
        out = dict(
            matched1=m1.numpy(), matched2=m2.numpy(),
            det_boxes_after=ex["det_boxes"].numpy(),
            feature=inter["feature"].numpy(), prev_feature=inter["prev_feature"].numpy(),
            residual=inter["residual"].numpy(), logits=inter["logits"].numpy(),
            fuse_shape=inter["fuse_shape"].numpy(), fuse_det=inter["fuse_det"].numpy(),
            res_coeff=inter["res_coeff"].numpy(),
            newborn=inter["newborn"].numpy(), fp=inter["fp"].numpy(),
            dead_trk=inter["dead_trk"].numpy(), fn=inter["fn"].numpy(),
            aug_shape=np.stack([inter["aug_shape%d" % k].numpy() for k in range(4)]),
            n_det=data["n_det"], n_prev=data["n_prev"],
            input_checksum=np.array([synthetic.checksum(data["det_boxes"], data["prev_det_boxes"],
                                                        data["bev"], data["prev_bev"])], dtype=np.uint64),
            weight_checksum=np.array([synthetic.checksum(*[weights[k] for k in sorted(weights)])],
                                     dtype=np.uint64),
            case=np.array(json.dumps(c)),
        )
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "m1", m1.shape, "max", float(m1.max()), "m2 max", float(m2.max()),
              "bytes", os.path.getsize(path))


def main_headline():
    torch.set_num_threads(1)
    for name, c in HEADLINE_CASES.items():
        pc_start, data, weights = case_inputs(c)
        model = ref_loader.build_reference_head(c["M"], 3, pc_start=pc_start)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
        det = torch.from_numpy(data["det_boxes"].copy())
        prev = torch.from_numpy(data["prev_det_boxes"].copy())
        m1, m2, ex = ref_loader.run_reference(
            model, torch.from_numpy(data["bev"]), torch.from_numpy(data["prev_bev"]), det, prev, capture=False)
        out = dict(matched1=m1.numpy(), matched2=m2.numpy(), det_boxes_after=ex["det_boxes"].numpy(),
                   n_det=data["n_det"], n_prev=data["n_prev"],
                   input_checksum=np.array([synthetic.checksum(data["det_boxes"], data["prev_det_boxes"],
                                                               data["bev"], data["prev_bev"])], dtype=np.uint64),
                   weight_checksum=np.array([synthetic.checksum(*[weights[k] for k in sorted(weights)])],
                                            dtype=np.uint64),
                   case=np.array(json.dumps(c)))
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "m1", m1.shape, "max", float(m1.max()), "m2 max", float(m2.max()), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    if "--headline" in sys.argv:
        main_headline()
    else:
        main()
