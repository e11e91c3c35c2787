"""Timing of the shared_conv producer (reported separately from the metric's timed region):
shasta_b200's tcgen05 implicit GEMM (3xTF32, fused BN/ReLU, channels-last output) next to the reference formulation on
the same GPU: cuDNN conv2d + BatchNorm2d + ReLU + permute(0,2,3,1).contiguous() in fp32 and with TF32 allowed."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shasta_b200 import build_track  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maps", type=int, default=8)
    ap.add_argument("--hw", type=int, default=180)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = dict(type="Shasta", reader=None, backbone=None, neck=None,
               bev_extractor=dict(type="BEVFeatureExtractor", pc_start=[-54, -54], voxel_size=[0.075, 0.075],
                                  out_stride=8), max_obj=20, num_feats=3)
    torch.manual_seed(0)
    with torch.device(dev):
        model = build_track(cfg)
    model.eval()
    from shasta_b200 import _cabi
    _cabi.lib().shasta_set_option(2, int(os.environ.get("SHASTA_DBG", "0"), 0))
    x = torch.relu(torch.randn((a.maps, 512, a.hw, a.hw), device=dev))
    flop = 2.0 * a.maps * a.hw * a.hw * 64 * 512 * 9
    out = {}
    with torch.no_grad():
        ours = model.shared_conv_nhwc(x, maps_per_launch=a.maps)
        out["tcgen05_3xtf32_ms_per_map"] = timed(lambda: model.shared_conv_nhwc(x, maps_per_launch=a.maps), a.iters) / a.maps
        ref = lambda: model.shared_conv(x).permute(0, 2, 3, 1).contiguous()  # noqa: E731
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        want = ref()
        out["cudnn_fp32_ms_per_map"] = timed(ref, a.iters) / a.maps
        torch.backends.cudnn.allow_tf32 = True
        out["cudnn_tf32_ms_per_map"] = timed(ref, a.iters) / a.maps
        got_tf32 = ref()
    scale = float(want.abs().max())
    out["max_err_vs_cudnn_fp32"] = float((ours - want).abs().max()) / scale
    out["cudnn_tf32_max_err_vs_fp32"] = float((got_tf32 - want).abs().max()) / scale
    out["tcgen05_tflops_fp32_equiv"] = flop / a.maps / (out["tcgen05_3xtf32_ms_per_map"] * 1e-3) / 1e12
    out["config"] = {"maps": a.maps, "hw": a.hw, "gflop_per_map": flop / a.maps / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
