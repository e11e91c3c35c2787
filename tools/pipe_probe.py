"""Timeline of the pipelined host-input path: when do gather (side stream) and compute (main stream) of each call run?"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from shasta_b200 import _cabi
sys.argv = ["bench.py", "--no-cpu"]
a = bench.parse()
dev = torch.device("cuda:0")
pc_start, d, bev, prev_bev = bench.make_inputs(a, dev, seed=1000)
model = bench.build_model(a, pc_start, dev)
lib = _cabi.lib()
B, M = a.batch, a.max_obj
h_bev = torch.empty(bev.shape, dtype=torch.float32, pin_memory=True); h_bev.copy_(bev)
h_prev = torch.empty(bev.shape, dtype=torch.float32, pin_memory=True); h_prev.copy_(prev_bev)
det = torch.from_numpy(d["det_boxes"]).to(dev); prev = torch.from_numpy(d["prev_det_boxes"]).to(dev)
model._ensure_packed(dev)
from shasta_b200.shasta import _Workspace
ws = [_Workspace(B, M, dev) for _ in range(2)]
dets = [det.clone() for _ in range(2)]
geom = model.bev_extractor.geom(a.hw, a.hw)
side = torch.cuda.Stream(); cur = torch.cuda.current_stream()
m1 = torch.empty((B, M, M + 2), device=dev); m2 = torch.empty((B, M + 2, M), device=dev)
hm1 = torch.empty((B, M, M + 2), pin_memory=True); hm2 = torch.empty((B, M + 2, M), pin_memory=True)
def E():
    return torch.cuda.Event(enable_timing=True)
N = 8
ev = [[E() for _ in range(6)] for _ in range(N)]
gathered = [torch.cuda.Event() for _ in range(2)]; consumed = [None, None]
base = E(); torch.cuda.synchronize(); base.record()
narrow = int(os.environ.get("NARROW", "2"))
lib.shasta_set_option(6, int(os.environ.get("CTAS", "0")))
for i in range(N):
    k = i & 1
    if consumed[k] is not None: side.wait_event(consumed[k])
    with torch.cuda.stream(side):
        ev[i][0].record(side)
        dets[k].copy_(det)
        _cabi.check(lib.shasta_gather_pair_f32(h_bev.data_ptr(), h_prev.data_ptr(), dets[k].data_ptr(), prev.data_ptr(), B, M,
                    ctypes.byref(geom), ws[k].buf.data_ptr(), ws[k].nbytes, narrow, ctypes.c_void_p(side.cuda_stream)), "g")
        ev[i][1].record(side); gathered[k].record(side)
    cur.wait_event(gathered[k])
    ev[i][2].record(cur)
    _cabi.check(lib.shasta_forward_f32(ctypes.byref(model._cparams), model._packed.data_ptr(), h_bev.data_ptr(), h_prev.data_ptr(),
                dets[k].data_ptr(), prev.data_ptr(), B, ctypes.byref(geom), ws[k].buf.data_ptr(), ws[k].nbytes, m1.data_ptr(), m2.data_ptr(),
                0x200, ctypes.c_void_p(cur.cuda_stream)), "f")
    ev[i][3].record(cur)
    hm1.copy_(m1, non_blocking=True); hm2.copy_(m2, non_blocking=True)
    ev[i][4].record(cur)
    consumed[k] = torch.cuda.Event(); consumed[k].record(cur)
torch.cuda.synchronize()
for i in range(N):
    t = [base.elapsed_time(e) for e in ev[i][:5]]
    print("call %d: gather %.2f-%.2f | compute %.2f-%.2f | d2h -%.2f" % (i, t[0], t[1], t[2], t[3], t[4]))
