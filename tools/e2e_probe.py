"""Which part of the e2e step costs what (variants of bench.run_e2e's step)?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
sys.argv = ["bench.py", "--no-cpu"]
a = bench.parse()
dev = torch.device("cuda:0")
pc_start, d, bev, prev_bev = bench.make_inputs(a, dev, seed=1000)
model = bench.build_model(a, pc_start, dev)
B, M = a.batch, a.max_obj
h_bev = torch.empty(bev.shape, dtype=torch.float32, pin_memory=True); h_bev.copy_(bev)
h_prev_bev = torch.empty(bev.shape, dtype=torch.float32, pin_memory=True); h_prev_bev.copy_(prev_bev)
h_det0 = torch.from_numpy(d["det_boxes"]).pin_memory(); h_prev = torch.from_numpy(d["prev_det_boxes"]).pin_memory()
NB = 3
sets = [{"det": h_det0.clone().pin_memory(), "m1": torch.empty((B, M, M + 2), pin_memory=True),
         "m2": torch.empty((B, M + 2, M), pin_memory=True), "ev": None} for _ in range(NB)]
def run(name, cpu_copy=True, d2h=True, ring=True, steps=40):
    c = [0]
    def step():
        st = sets[c[0] % NB]; c[0] += 1
        if ring and st["ev"] is not None: st["ev"].synchronize()
        if cpu_copy: st["det"].copy_(h_det0)
        ex = {"det_boxes": st["det"], "prev_det_boxes": h_prev, "bev_feature": h_bev, "prev_bev_feature": h_prev_bev}
        m1, m2, _ = model(ex, train_mode=False)
        if d2h:
            st["m1"].copy_(m1, non_blocking=True); st["m2"].copy_(m2, non_blocking=True)
        st["ev"] = torch.cuda.Event(); st["ev"].record()
    for _ in range(4): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): step()
    torch.cuda.synchronize()
    print("%-28s %.3f ms/step" % (name, (time.perf_counter() - t0) * 1e3 / steps))
with torch.no_grad():
    run("full")
    run("no cpu copy", cpu_copy=False)
    run("no d2h of m1/m2", d2h=False)
    run("no ring sync", ring=False)
    run("no copy, no d2h", cpu_copy=False, d2h=False)
