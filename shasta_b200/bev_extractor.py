"""BEVFeatureExtractor behind the reference's SECOND_STAGE registry (det3d/models/second_stage/bird_eye_view.py:9-41).

Same constructor (``pc_start, voxel_size, out_stride``) and the same ``forward(example, batch_centers, num_point)``
contract: a python list (len B) of ``(M, num_point*C)`` tensors sampled from ``example['bev_feature']`` (B,H,W,C).
The sampling itself is the CUDA kernel behind ``shasta_bilinear_f32`` (bit-exact restatement of
center_utils.py:92-121); ``Shasta.forward`` does not go through this per-frame API but through the fused
box gather, which computes the sample points on the device as well.
"""
import ctypes

import torch
from torch import nn

from . import _cabi
from .registry import SECOND_STAGE


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@SECOND_STAGE.register_module
class BEVFeatureExtractor(nn.Module):
    def __init__(self, pc_start, voxel_size, out_stride):
        super().__init__()
        self.pc_start = pc_start
        self.voxel_size = voxel_size
        self.out_stride = out_stride

    def absl_to_relative(self, absolute):
        """bird_eye_view.py:18-22 (metric -> BEV pixel coordinates)."""
        a1 = (absolute[..., 0] - self.pc_start[0]) / self.voxel_size[0] / self.out_stride
        a2 = (absolute[..., 1] - self.pc_start[1]) / self.voxel_size[1] / self.out_stride
        return a1, a2

    def geom(self, height, width):
        return _cabi.ShastaGeom(float(self.pc_start[0]), float(self.pc_start[1]), float(self.voxel_size[0]),
                                float(self.voxel_size[1]), float(self.out_stride), int(height), int(width))

    def forward(self, example, batch_centers, num_point):
        bev = example["bev_feature"]
        if not bev.is_cuda:
            raise _cabi.ShastaLibraryError("BEVFeatureExtractor: bev_feature must be a CUDA tensor (no CPU path)")
        lib = _cabi.lib()
        ret_maps = []
        for batch_idx in range(len(bev)):
            im = bev[batch_idx]
            if im.dtype != torch.float32 or not im.is_contiguous():
                im = im.float().contiguous()
            H, W, C = im.shape
            xs, ys = self.absl_to_relative(batch_centers[batch_idx])
            xs = xs.float().contiguous()
            ys = ys.float().contiguous()
            n = xs.numel()
            out = torch.empty((n, C), dtype=torch.float32, device=im.device)
            with torch.cuda.device(im.device):
                rc = lib.shasta_bilinear_f32(im.data_ptr(), H, W, C, xs.data_ptr(), ys.data_ptr(), n,
                                             out.data_ptr(), _stream_ptr(im.device))
            _cabi.check(rc, "shasta_bilinear_f32")
            if num_point > 1:
                section = n // num_point
                out = torch.cat([out[i * section:(i + 1) * section] for i in range(num_point)], dim=1)
            ret_maps.append(out)
        return ret_maps
