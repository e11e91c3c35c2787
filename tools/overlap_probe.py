"""Does the aug_shape GEMM of one batch overlap the pairwise kernel of another? Times the two stage calls alone and
enqueued together on two streams (two model instances = two workspaces). Round-2 result (profiles/README.md): they do
not - 0.281 + 0.215 ms alone, 0.477 ms together; a variant of the GEMM cut down to share an SM with a pairwise CTA
(4-stage ring, 256 TMEM columns, 128 registers at launch + setmaxnreg) ran 0.533 ms together and 15 % slower alone.
Usage: python tools/overlap_probe.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shasta_b200 import _cabi, synthetic  # noqa: E402
from tests import gpu_util as G  # noqa: E402


def main():
    M, B, H = 200, 64, 128
    pc_start = (-H * 0.3, -H * 0.3)
    lib = _cabi.lib()
    data = synthetic.make_frame_pairs(B, M, H, H, 5, pc_start=pc_start)
    torch.manual_seed(0)
    st = []
    for k in range(2):
        model = G.make_model(M, pc_start)
        s = G.Stages(model, B)
        for key, reg in (("bev", _cabi.WS_FEAT_CUR), ("prev_bev", _cabi.WS_FEAT_PREV)):
            s.gather(G.t(data[key]), G.t(data["det_boxes" if key == "bev" else "prev_det_boxes"]), reg)
        det = G.t(data["det_boxes"])
        s.anchors(det, G.t(data["prev_det_boxes"]))
        s.project(det)
        s.pairwise(0)
        st.append((s, det, G.t(data["prev_det_boxes"])))
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def run_anchors(s, det, prev, stream):
        rc = lib.shasta_anchors_f32(ctypes.byref(s.model._cparams), det.data_ptr(), prev.data_ptr(), B,
                                    s.ws.buf.data_ptr(), ctypes.c_void_p(stream.cuda_stream))
        _cabi.check(rc, "anchors")

    def run_pairwise(s, stream, variant=0):
        rc = lib.shasta_pairwise_f32(s.model._packed.data_ptr(), B, M, s.ws.buf.data_ptr(), variant,
                                     ctypes.c_void_p(stream.cuda_stream))
        _cabi.check(rc, "pairwise")

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sa.wait_event(e0), sb.wait_event(e0)
        for _ in range(reps):
            fn()
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for stages in (0,):
        ta = timed(lambda: run_anchors(*st[0], sa))
        tp = timed(lambda: run_pairwise(st[1][0], sb))
        both = timed(lambda: (run_anchors(*st[0], sa), run_pairwise(st[1][0], sb)))
        both2 = timed(lambda: (run_pairwise(st[1][0], sb), run_anchors(*st[0], sa)))
        print("ring depth %d%s: anchors stage %.3f ms, pairwise %.3f ms, sum %.3f, together %.3f (pairwise first: %.3f)"
              % ((stages & 15) or 7, " (one GEMM CTA per SM)" if stages & 16 else "", ta, tp, ta + tp, both, both2))


if __name__ == "__main__":
    main()
