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


def rel_err(a, b, floor_frac=1e-3):
    """max |a-b| / max(|b|, floor) with a floor at ``floor_frac`` of the tensor's scale (outputs are probabilities;
    the large-size tests also check with floor_frac = 1e-6, i.e. the small probabilities of peaky cases relatively)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = max(1e-30, floor_frac * float(np.abs(b).max()))
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


DECODE_KEYS = ("dead", "fn", "keep_prev", "keep_dets", "newborn")


def assert_same_association(oracle_mod, o1, o2, m1, m2, n_prev, n_det, tie_rel=2e-5, where=""):
    """Decode (tools/nusc_shasta/eval.py:126-181) of the CUDA outputs against the decode of the oracle's outputs for ONE
    frame pair: every discrete decision (dead / FN / kept / newborn lists) must be identical, and so must every row /
    column argmax - except where the ORACLE's own top-2 margin is below ``tie_rel`` relative, i.e. inside fp32
    rounding noise (two candidates whose affinities agree to ~5 digits: the reference's pick then depends on its own
    summation order). Returns the number of such fp32-level ties (reported by the callers)."""
    want = oracle_mod.decode(o1, o2, n_prev, n_det)
    got = oracle_mod.decode(m1, m2, n_prev, n_det)
    for key in DECODE_KEYS:
        assert got[key] == want[key], (where, key)
    ties = 0
    a1 = o1.detach().cpu().numpy().astype(np.float64)
    a2 = o2.detach().cpu().numpy().astype(np.float64)
    if n_prev > 0:
        A = np.concatenate((a1[:n_prev, :n_det], a1[:n_prev, -2:]), axis=1)
        for n, (kg, kw) in enumerate(zip(got["row_argmax"], want["row_argmax"])):
            if kg != kw:
                assert abs(A[n, kg] - A[n, kw]) <= tie_rel * abs(A[n, kw]), (where, "row_argmax", n, kg, kw, A[n, kg], A[n, kw])
                ties += 1
    if n_det > 0:
        Bm = np.concatenate((a2[want["keep_prev"], :n_det], a2[-2:, :n_det]), axis=0)
        for k, (ng, nw) in enumerate(zip(got["col_argmax"], want["col_argmax"])):
            if ng != nw:
                assert abs(Bm[ng, k] - Bm[nw, k]) <= tie_rel * abs(Bm[nw, k]), (where, "col_argmax", k, ng, nw, Bm[ng, k], Bm[nw, k])
                ties += 1
    return ties
