"""Shared helpers: golden fixtures (tests/golden/*.npz, written by oracle/make_golden.py from the unmodified
reference) and their regenerated inputs."""
import glob
import json
import os

import numpy as np

from shasta_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "m*.npz")))  # head fixtures: m<M>_<map>_b<B>[_peaky].npz


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    c = json.loads(str(g["case"]))
    pc_start = (-c["W"] * 0.6 / 2.0, -c["H"] * 0.6 / 2.0)
    data = synthetic.make_frame_pairs(c["B"], c["M"], c["H"], c["W"], c["seed"], pc_start=pc_start)
    weights = synthetic.make_weights(c["M"], seed=c["wseed"], peaky=c["peaky"])
    assert synthetic.checksum(data["det_boxes"], data["prev_det_boxes"], data["bev"], data["prev_bev"]) == int(
        g["input_checksum"][0]), "synthetic generator drifted from the fixture"
    assert synthetic.checksum(*[weights[k] for k in sorted(weights)]) == int(g["weight_checksum"][0])
    return c, pc_start, data, weights, g


def headline_names():
    """Outputs-only fixtures of the headline size (M = 200), written by ``python -m oracle.make_golden --headline``."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "h*.npz")))
