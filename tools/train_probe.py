"""Where does a training step spend its time? (CPU enqueue time vs GPU time, per phase)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from shasta_b200 import loss as L, training
a = bench.parse.__wrapped__() if hasattr(bench.parse, "__wrapped__") else None
sys.argv = ["bench.py", "--train", "--no-cpu"]
a = bench.parse()
device = torch.device("cuda:0")
pc_start, d, bev, prev_bev = bench.make_inputs(a, device, seed=2000)
model = bench.build_model(a, pc_start, device); model.train()
params = training.differentiable_parameters(model)
for p_ in model.parameters(): p_.requires_grad_(False)
for p_ in params: p_.requires_grad_(True)
opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-2)
B, M = a.batch, a.max_obj
det0 = torch.from_numpy(d["det_boxes"]).to(device); prev = torch.from_numpy(d["prev_det_boxes"]).to(device); det = det0.clone()
gt = torch.zeros((B, M + 2, M + 2), device=device); gt[:, :100, 5] = 1.0
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(6):
    t0 = time.perf_counter(); e = [ev()]
    det.copy_(det0); opt.zero_grad(set_to_none=True)
    m1, m2 = model.affinity(bev, prev_bev, det, prev); e.append(ev()); t1 = time.perf_counter()
    loss = L.affinity_loss(m1, m2, gt); loss.backward(); e.append(ev()); t2 = time.perf_counter()
    opt.step(); e.append(ev()); t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("it %d  cpu: fwd %.1f bwd %.1f opt %.1f sync %.1f ms | gpu: fwd %.2f bwd %.2f opt %.2f ms" % (
        it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3,
        e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for it in range(3):
    det.copy_(det0); opt.zero_grad(set_to_none=True)
    m1, m2 = model.affinity(bev, prev_bev, det, prev)
    loss = L.affinity_loss(m1, m2, gt); loss.backward(); opt.step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
