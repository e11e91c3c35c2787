"""bench.py's byte accounting against SURVEY.md §8d (the figures `roofline.achieved` and `path_hbm_frac` rest on):
W(200) = 1 030 701 708 B of fp32 weights, IO(200) = 2 388 800 B per frame pair, and the parameter count of the head
derived independently from the state-dict shapes."""
import numpy as np

import bench
from shasta_b200 import synthetic


def _param_count(M):
    shapes = synthetic.head_param_shapes(M)
    return sum(int(np.prod(s)) for name, s in shapes.items() if not name.startswith("shared_conv"))


def test_path_bytes_match_survey():
    M = 200
    weights = bench.path_bytes(M, 0, 512)
    assert weights == 1_030_701_708
    io = bench.path_bytes(M, 1, 512) - weights
    assert io == 2_388_800
    assert bench.path_bytes(M, 64, 512) == weights + 64 * io


def test_weight_bytes_equal_four_times_the_parameter_count():
    for M in (20, 90, 200):
        assert bench.path_bytes(M, 0, 180) == 4 * _param_count(M)


def test_roofline_kernel_bytes():
    M, B = 200, 64
    w = 4 * (5 * M) * (320 * M) * 4
    assert w == 1_024_000_000
    assert bench.algorithmic_bytes_anchor_hidden(M, B) == w + 2 * B * 320 * M * 4 == 1_056_768_000
    assert bench.algorithmic_bytes_anchor_hidden(M, B, bf16=True) == 1_056_768_000 // 2
    # bf16 mode: only the aug_shape.i.0 weights shrink in the whole-path figure
    assert bench.path_bytes(M, B, 512) - bench.path_bytes(M, B, 512, bf16=True) == w // 2
