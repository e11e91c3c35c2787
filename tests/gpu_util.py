"""Helpers for the -m gpu parity tests: build the CUDA head with given weights, call single stages of the C ABI on
torch tensors, read workspace regions."""
import ctypes

import numpy as np
import torch

from shasta_b200 import _cabi, build_track, load_matching_state_dict

DEV = "cuda:0"


def make_model(M, pc_start, weights_np=None, device=DEV):
    cfg = dict(type="Shasta", reader=None, backbone=None, neck=None,
               bev_extractor=dict(type="BEVFeatureExtractor", pc_start=list(pc_start), voxel_size=[0.075, 0.075],
                                  out_stride=8),
               max_obj=M, num_feats=3)
    with torch.device(device):
        model = build_track(cfg)
    model.eval()
    if weights_np is not None:
        skipped = load_matching_state_dict(model, {k: torch.from_numpy(v) for k, v in weights_np.items()})
        assert not skipped, skipped
    return model


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def t(a, device=DEV):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


class Stages:
    """Thin stage-by-stage driver over the C ABI for one (model, batch) pair."""

    def __init__(self, model, B):
        self.lib = _cabi.lib()
        self.model, self.B, self.M = model, B, model.max_obj
        model._ensure_packed(torch.device(DEV))
        self.ws = model._workspace(B, torch.device(DEV))
        self.ws.buf.zero_()
        self.T = self.M + 2
        self.DP = self.lib.shasta_proj_cur_stride(self.M)
        self.RS = self.lib.shasta_row_stride(self.M)

    def region(self, rid, shape):
        return self.ws.region(rid, int(np.prod(shape))).view(*shape)

    def geom(self, H, W):
        return self.model.bev_extractor.geom(H, W)

    def gather(self, bev, boxes, region, variant=0):
        B, H, W, _ = bev.shape
        g = self.geom(H, W)
        feat = self.region(region, (B, self.T, 320))
        rc = self.lib.shasta_gather_f32(bev.data_ptr(), boxes.data_ptr(), boxes.shape[-1], B, self.M, ctypes.byref(g),
                                        feat.data_ptr(), self.T * 320, variant, stream())
        _cabi.check(rc, "shasta_gather_f32")
        return feat

    def anchors(self, det, prev):
        rc = self.lib.shasta_anchors_f32(ctypes.byref(self.model._cparams), det.data_ptr(), prev.data_ptr(), self.B,
                                         self.ws.buf.data_ptr(), stream())
        _cabi.check(rc, "shasta_anchors_f32")

    def project(self, det_inout=None):
        rc = self.lib.shasta_project_f32(self.model._packed.data_ptr(), self.B, self.M, self.ws.buf.data_ptr(),
                                         det_inout.data_ptr() if det_inout is not None else None, stream())
        _cabi.check(rc, "shasta_project_f32")

    def pairwise(self, variant=0):
        rc = self.lib.shasta_pairwise_f32(self.model._packed.data_ptr(), self.B, self.M, self.ws.buf.data_ptr(),
                                          variant, stream())
        _cabi.check(rc, "shasta_pairwise_f32")

    def aff_softmax(self):
        m1 = torch.empty((self.B, self.M, self.M + 2), device=DEV)
        m2 = torch.empty((self.B, self.M + 2, self.M), device=DEV)
        rc = self.lib.shasta_aff_softmax_f32(self.model._packed.data_ptr(), self.B, self.M, self.ws.buf.data_ptr(),
                                             m1.data_ptr(), m2.data_ptr(), stream())
        _cabi.check(rc, "shasta_aff_softmax_f32")
        return m1, m2


def rel_err(a, b):
    """max |a-b| / max(|b|, floor) with a floor at 1e-3 of the tensor's scale (outputs are probabilities)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = max(1e-30, 1e-3 * float(np.abs(b).max()))
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())
