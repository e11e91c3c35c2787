"""PCIe probes for the host-input path: pinned H2D copy bandwidth and the zero-copy gather (LDG vs distinct-pixel)."""
import ctypes, os, sys, time
import numpy as np, torch
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
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
hb = torch.empty(64 * 1024 * 1024, dtype=torch.float32, pin_memory=True)
ms = timed(lambda: buf.copy_(hb, non_blocking=True)); print("H2D memcpy 256 MB: %.2f ms = %.1f GB/s" % (ms, 0.268435456 / ms * 1e3))
ms = timed(lambda: hb.copy_(buf, non_blocking=True)); print("D2H memcpy 256 MB: %.2f ms = %.1f GB/s" % (ms, 0.268435456 / ms * 1e3))
ws = model._workspace(B, dev); model._ensure_packed(dev)
geom = model.bev_extractor.geom(a.hw, a.hw)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
taps = 2 * B * 5 * M * 4 * 64 * 4
for name, src0, src1 in (("device maps", bev, prev_bev), ("host maps", h_bev, h_prev)):
    for flags in (0, 2):
        f = lambda: _cabi.check(lib.shasta_gather_pair_f32(src0.data_ptr(), src1.data_ptr(), det.data_ptr(), prev.data_ptr(), B, M,
                                                            ctypes.byref(geom), ws.buf.data_ptr(), ws.nbytes, flags, st), "g")
        ms = timed(f)
        print("%s gather flags %d: %.3f ms (nominal taps %.1f MB -> %.1f GB/s)" % (name, flags, ms, taps / 1e6, taps / ms / 1e6))
