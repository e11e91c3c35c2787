#!/usr/bin/env python
"""Affinity frame-pairs/sec of the ShaSTA hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--max-obj M] [--hw H]

A *step* is one pass of the hot path (gather -> anchors -> pairwise MLPs -> aff + dual softmax) over one batch of B
synthetic frame pairs per GPU. Workload at the defaults = BASELINE.json configs[1]: M = 200 tracks x 200 detections,
512 x 512 x 64 channels-last BEV maps, fp32, random-init weights. Frame pairs are independent, so N GPUs run N
independent shards (weak scaling); NCCL only gathers the per-rank decode results.

`value`  : frame pairs / s with all inputs resident in HBM when the timed region starts (max over ranks, CUDA events).
`e2e`    : the same through Shasta.forward with HOST (pinned) buffers: boxes are copied H2D, BEV maps are sampled in
           place over PCIe by the gather kernel, results are copied D2H, all inside the timed region.
`roofline`: the dominant kernel (anchor_hidden_kernel, weight streaming of aug_shape.i.0) against measured HBM GB/s.
`cpu_baseline`: the CPU oracle port of the reference path on this box's host cores, bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "affinity frame-pairs/sec (200x200 pairs)"
UNIT = "frame-pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frame pairs per step per GPU")
    ap.add_argument("--max-obj", type=int, default=200)
    ap.add_argument("--hw", type=int, default=512, help="BEV map height = width")
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0, help="kernel variant flags (see shasta_b200.h)")
    ap.add_argument("--anchor-path", type=int, default=0, help="0 auto, 1 streaming CUDA-core, 2 tcgen05")
    ap.add_argument("--raw-hi", type=int, default=1)
    ap.add_argument("--dbg", type=lambda x: int(x, 0), default=0, help="kernel experiment bits (results invalid when non-zero)")
    ap.add_argument("--splits", type=int, default=0, help="force the split-K count of the anchors GEMM (experiment)")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE.json config 5: training step (forward + loss + CUDA backward + gradient all-reduce "
                         "+ Adam on the differentiated parameters) instead of inference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--streams", type=int, default=2,
                    help="independent batches in flight: N step graphs replayed round-robin on N streams, each with its "
                         "own model instance / workspace (default 2: the tail of one batch's kernels - 101-CTA aff "
                         "tiles, the second wave of the projection GEMM, small launches - overlaps the next batch)")
    ap.add_argument("--bf16", action="store_true", help="bf16 mode (shasta_forward_bf16): separate tolerance, dtype bf16")
    ap.add_argument("--opt", action="append", default=[], metavar="ID=VALUE",
                    help="shasta_set_option(ID, VALUE) before the run (experiment knob, repeatable)")
    ap.add_argument("--workload", default="head", choices=["head", "multiclass"],
                    help="head = the headline metric (BASELINE.json configs[1]); multiclass = configs[2]: the 10-class "
                         "sequence batch, scenes sharded over the ranks (tools/bench_multiclass.py, its own JSON line)")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_multiclass as _bm
    _bm.add_args(ap)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def config_dict(a, extra=None):
    c = {"workload": "ShaSTA affinity head, M=%d (T=D=%d), %dx%dx64 NHWC BEV maps, fp32, random-init weights "
                     "(BASELINE.json configs[1])" % (a.max_obj, a.max_obj + 2, a.hw, a.hw),
         "max_obj": a.max_obj, "frame_pairs_per_step_per_gpu": a.batch, "bev_hw": a.hw,
         "l2_policy": "inputs larger than L2: 1.03 GB of streamed weights + %.1f GB of BEV maps per step vs 126 MB L2"
                      % (2 * a.batch * a.hw * a.hw * 64 * 4 / 1e9),
         "parallelism": "independent frame-pair shards per GPU (no data-path collective)"}
    try:   # recorded B sweep of this code (profiles/r2_batch_sweep.json): which frames-per-step is the best throughput
        rows = json.load(open(os.path.join(ROOT, "profiles", "r2_batch_sweep.json")))["rows"]
        best = max(rows, key=lambda r: r["frame_pairs_per_s"])
        same = max((r for r in rows if r["batch"] == 64), key=lambda r: r["frame_pairs_per_s"])
        c["batch_sweep_recorded"] = {"file": "profiles/r2_batch_sweep.json", "best_batch": best["batch"],
                                     "best_frame_pairs_per_s": best["frame_pairs_per_s"],
                                     "batch_64_same_box": same["frame_pairs_per_s"],
                                     "note": "larger --batch amortises the 1.03 GB weight stream over more pairs "
                                             "(one BN = 128 GEMM tile up to 128; two passes above)"}
    except Exception:  # noqa: BLE001
        pass
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region (B200_PROFILING.md recipe), taken in-process through
    NVML every 20 ms. (An `nvidia-smi -lms` child was measurably perturbing short timed regions: its queries stall
    kernel launches for milliseconds.) Falls back to one nvidia-smi query if NVML is unavailable."""
    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.handle = None
        self.nv = None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            if os.environ.get("SHASTA_NO_SAMPLER"):
                raise RuntimeError("sampler disabled")
            import pynvml as nv
            nv.nvmlInit()
            try:
                uuid = torch.cuda.get_device_properties(self.gpu).uuid
                self.handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:  # noqa: BLE001
                self.handle = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nv = nv
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.time(), mhz, reasons))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(float(os.environ.get("SHASTA_SAMPLER_PERIOD", "0.02")))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken inside [t_begin, t_end]."""
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
                f = [float(x) for x in out.stdout.strip().split(",")]
                return {"sm_mhz": f[0], "sm_max_mhz": f[1], "reasons": ["nvml unavailable: one nvidia-smi sample after the run"],
                        "samples": 1}
            except Exception:  # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"], "samples": 0}
        self._stop.set()
        self.thread.join(1.0)
        rows = [r for r in self.rows if (t_begin is None or r[0] >= t_begin) and (t_end is None or r[0] <= t_end)]
        if not rows:
            rows = self.rows[-2:]
        reasons = set()
        for _, _, bits in rows:
            for name, bit in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                              ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP)):
                if bits & bit:
                    reasons.add(name)
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml"}


# --------------------------------------------------------------------------------------------------
def make_inputs(a, device, seed):
    """Boxes from the seeded nuScenes-shape generator; BEV maps relu(N(0,1)) drawn on the device."""
    from shasta_b200 import synthetic
    pc_start = (-a.hw * 0.3, -a.hw * 0.3)
    d = synthetic.make_frame_pairs(a.batch, a.max_obj, a.hw, a.hw, seed, pc_start=pc_start, with_maps=False)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    maps = []
    for _ in range(2):
        m = torch.empty((a.batch, a.hw, a.hw, 64), dtype=torch.float32, device=device)
        for b in range(a.batch):
            m[b].normal_(generator=g).relu_()
        maps.append(m)
    return pc_start, d, maps[0], maps[1]


def build_model(a, pc_start, device):
    from shasta_b200 import build_track
    cfg = dict(type="Shasta", reader=None, backbone=None, neck=None,
               bev_extractor=dict(type="BEVFeatureExtractor", pc_start=list(pc_start), voxel_size=[0.075, 0.075],
                                  out_stride=8),
               max_obj=a.max_obj, num_feats=3)
    torch.manual_seed(0)
    with torch.device(device):
        model = build_track(cfg)
    model.eval()
    model.kernel_flags = a.flags
    model.bf16 = bool(getattr(a, "bf16", False))
    model.cuda_graphs = not getattr(a, "no_graph", False)
    return model


def algorithmic_bytes_anchor_hidden(M, B, bf16=False):
    """Dominant kernel (anchor_hidden_kernel): the four aug_shape.i.0 matrices are read once (4 x 5M x 320M fp32),
    the gathered features of both frames once; split-K partials are an implementation artefact and
    are not counted."""
    w = 4 * (5 * M) * (320 * M) * 4
    x = 2 * B * (320 * M) * 4
    return (w + x) // 2 if bf16 else w + x   # bf16 mode streams bf16 copies of both operands


def path_bytes(M, B, hw, bf16=False):
    """bytes(M,B) per step = W(M) + B*IO(M); bf16 mode reads the aug_shape.i.0 weights as bf16."""
    params = 4 * ((5 * M) * (320 * M) + 5 * M + 320 * 5 * M + 320) + 4 * ((7 * M // 32) * 7 * M + 7 * M // 32 + 7 * (7 * M // 32) + 7)
    params += 640 * 40 + 40 + 800 + 20 + 200 + 10 + 10 + 1 + 6 * 32 + 32 + 256 + 8 + 8 + 1 + 646 * 72 + 72 + 72 * 18 + 18 + 54 + 3
    params += 2 * ((M + 2) * 128 + 128 * 64 + 64 * 32) + 128 + 64 + 32 + 64 + 128 + (M + 2)
    io = 2 * 5 * M * 4 * 64 * 4 + 2 * M * 11 * 4 + 2 * M * (M + 2) * 4
    return 4 * params + B * io - (2 * 4 * (5 * M) * (320 * M) if bf16 else 0)


# --------------------------------------------------------------------------------------------------
def cpu_reference_run(a, seconds, weights_state=None, sample_pairs=1):
    """Times the CPU oracle port of the reference path (oracle/shasta_oracle.py — the reference itself is Python and
    does not travel to the GPU box) with every host thread torch can use, on a bounded sample of the workload."""
    from oracle import shasta_oracle as O
    from shasta_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    M, hw = a.max_obj, a.hw
    pc_start = (-hw * 0.3, -hw * 0.3)
    d = synthetic.make_frame_pairs(sample_pairs, M, hw, hw, 1234, pc_start=pc_start, with_maps=False)
    g = torch.Generator().manual_seed(1)
    bev = torch.randn((sample_pairs, hw, hw, 64), generator=g).relu_()
    prev_bev = torch.randn((sample_pairs, hw, hw, 64), generator=g).relu_()
    if weights_state is None:
        torch.manual_seed(0)
        shapes = synthetic.head_param_shapes(M)
        weights_state = {}
        for k, shp in shapes.items():
            fan_in = shapes[k.rsplit(".", 1)[0] + ".weight"][1]
            weights_state[k] = (torch.rand(shp) * 2 - 1) / max(fan_in, 1) ** 0.5
    det = torch.from_numpy(d["det_boxes"])
    prev = torch.from_numpy(d["prev_det_boxes"])
    times = []
    fwd, kind = cpu_forward_fn(M, pc_start, weights_state)
    with torch.no_grad():
        fwd(bev, prev_bev, det.clone(), prev)  # warm-up
        t_end = time.perf_counter() + seconds
        while time.perf_counter() < t_end or len(times) < 3:
            t0 = time.perf_counter()
            fwd(bev, prev_bev, det.clone(), prev)
            times.append(time.perf_counter() - t0)
            if len(times) >= 200:
                break
    med = float(np.median(times))
    return {"value": sample_pairs / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%d iterations of %d frame pair(s), M=%d, %dx%d maps, median %.1f ms each; %s "
                      "(reference formulation, materialised pair tensors), torch %s CPU fp32"
                      % (len(times), sample_pairs, M, hw, hw, med * 1e3,
                         "the unmodified reference head (oracle/ref_loader.py)" if kind == "reference"
                         else "oracle/shasta_oracle.py", torch.__version__),
            "ms_per_frame_pair": med * 1e3 / sample_pairs}


def cpu_forward_fn(M, pc_start, weights_state):
    """The CPU implementation of the path that the reference arm / cpu_baseline time: the UNMODIFIED reference
    (oracle/ref_loader.py, kind "reference") where its tree is present (SHASTA_REF_ROOT, /root/reference or
    baseline/_ref), else the oracle port (kind "port" - bit-identical to it, tests/test_oracle*.py)."""
    from oracle import ref_loader
    if not ref_loader.available() and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "det3d")):
        ref_loader.REF_ROOT = os.path.join(ROOT, "baseline", "_ref")
    if ref_loader.available():
        try:
            model = ref_loader.build_reference_head(M, 3, pc_start=pc_start)
            model.load_state_dict(weights_state, strict=False)

            def fwd(bev, prev_bev, det, prev):
                return ref_loader.run_reference(model, bev, prev_bev, det, prev)[:2]
            return fwd, "reference"
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference tree present but not loadable (%s): timing the oracle port\n" % e)
    from oracle import shasta_oracle as O

    def fwd(bev, prev_bev, det, prev):
        return O.forward(weights_state, bev, prev_bev, det, prev, pc_start=pc_start)
    return fwd, "port"


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = a.steps, a.warmup
    # each step = a bounded sample (ONE frame pair) of the workload, so K steps finish within minutes on a CPU
    t = cpu_reference_timed(a, steps, warm)
    res = {"cores": torch.get_num_threads()}
    line = {"metric": METRIC, "value": t["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": config_dict(a, {"frame_pairs_per_step_per_gpu": 1,
                                      "note": "reference path on host cores (%s); one frame pair per step"
                                              % ("unmodified reference" if t["kind"] == "reference" else "CPU oracle port")}),
            "cpu_baseline": {"value": t["value"], "unit": UNIT, "cores": res["cores"], "kind": t["kind"],
                             "sample": "%d steps x 1 frame pair after %d warm-up" % (steps, warm)},
            "e2e": {"value": t["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_reference_timed(a, steps, warm):
    from oracle import shasta_oracle as O
    from shasta_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    M, hw = a.max_obj, a.hw
    pc_start = (-hw * 0.3, -hw * 0.3)
    d = synthetic.make_frame_pairs(1, M, hw, hw, 1234, pc_start=pc_start, with_maps=False)
    g = torch.Generator().manual_seed(1)
    bev = torch.randn((1, hw, hw, 64), generator=g).relu_()
    prev_bev = torch.randn((1, hw, hw, 64), generator=g).relu_()
    torch.manual_seed(0)
    shapes = synthetic.head_param_shapes(M)
    w = {}
    for k, shp in shapes.items():
        fan_in = shapes[k.rsplit(".", 1)[0] + ".weight"][1]
        w[k] = (torch.rand(shp) * 2 - 1) / max(fan_in, 1) ** 0.5
    det = torch.from_numpy(d["det_boxes"])
    prev = torch.from_numpy(d["prev_det_boxes"])
    fwd, kind = cpu_forward_fn(M, pc_start, w)
    with torch.no_grad():
        for _ in range(warm):
            fwd(bev, prev_bev, det.clone(), prev)
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd(bev, prev_bev, det.clone(), prev)
        dt = time.perf_counter() - t0
    return {"value": steps / dt, "ms_per_step": dt / steps * 1e3, "kind": kind}


# --------------------------------------------------------------------------------------------------
def run_train(a):
    """Config 5: data-parallel training step of the head. One step = forward (same kernels as inference) + the
    reference's loss (train.py:201-211) + CUDA backward + all-reduce of the gradients over the ranks + Adam."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # the gradient all-reduce runs next to the pairwise backward (whose CTAs hold whole SMs for ~2 ms each): NCCL's
        # stream gets priority so its few CTAs take the first slots that free up
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=device)
    from shasta_b200 import _cabi, loss as L, training
    lib = _cabi.lib()
    lib.shasta_set_option(2, a.dbg)
    pc_start, d, bev, prev_bev = make_inputs(a, device, seed=2000 + rank)
    model = build_model(a, pc_start, device)
    model.train()
    params = training.differentiable_parameters(model)
    for p_ in model.parameters():
        p_.requires_grad_(False)
    for p_ in params:
        p_.requires_grad_(True)
    # train.py:147 Adam. One fused multi-tensor instance per aug_shape.i generator (4 x 64.3 M parameters) + one for the
    # rest: with several ranks the update of generator i runs while the all-reduce of generator i+1 is still in flight
    adam = dict(lr=1e-4, weight_decay=1e-2, fused=True)
    big_params = [p_ for p_ in params if p_.numel() >= (1 << 22)]          # aug_shape.i.0.weight, i = 0..3
    # the big ones: the library's streaming Adam kernel (28 bytes per parameter at HBM speed); BENCH_TORCH_ADAM=1: torch's
    big_adam = dict(lr=adam["lr"], weight_decay=adam["weight_decay"])
    opts = [(torch.optim.Adam([p_], **adam) if os.environ.get("BENCH_TORCH_ADAM") else training.StreamAdam([p_], **big_adam))
            for p_ in big_params]
    opt_rest = torch.optim.Adam([p_ for p_ in params if p_.numel() < (1 << 22)], **adam)
    B, M = a.batch, a.max_obj
    det0 = torch.from_numpy(d["det_boxes"]).to(device)
    prev = torch.from_numpy(d["prev_det_boxes"]).to(device)
    det = det0.clone()
    rng = np.random.default_rng(7 + rank)
    gt = np.zeros((B, M + 2, M + 2), np.float32)
    for b_ in range(B):
        for t in range(int(d["n_prev"][b_])):
            gt[b_, t, rng.integers(0, M + 2)] = 1.0
    gt = torch.from_numpy(gt).to(device)
    flat = torch.zeros(sum(p_.numel() for p_ in params), device=device)
    # the aug_shape gradients (1.03 GB of the 1.03 GB + 1.6 MB) are final after ~15 % of the backward: their all-reduce
    # starts on a side stream behind the library's event and overlaps the rest of the backward
    pending = []
    if dist is not None and not os.environ.get("BENCH_NO_GRAD_OVERLAP"):
        side = torch.cuda.Stream()

        def grad_hook(aug_grads, ready):
            side.wait_event(ready)
            with torch.cuda.stream(side):
                for g_ in aug_grads:
                    if g_.numel() >= (1 << 22):
                        pending.append((g_, dist.all_reduce(g_, op=dist.ReduceOp.AVG, async_op=True)))
        model.grad_sync_hook = grad_hook

    marks = []

    def mark():
        if os.environ.get("BENCH_TRAIN_PHASES"):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)

    def step():
        mark()
        det.copy_(det0)
        for o_ in opts + [opt_rest]:
            o_.zero_grad(set_to_none=True)
        m1, m2 = model.affinity(bev, prev_bev, det, prev)
        loss = L.affinity_loss(m1, m2, gt)
        mark()
        loss.backward()
        mark()
        if dist is not None:   # DDP-style gradient averaging (train.py:154-156): the four 257 MB aug_shape.i.0 gradients
            # are averaged in place, everything else (1.6 MB in 60 tensors) through one flat bucket
            big = [p_ for p_ in params if p_.numel() >= (1 << 22)]
            small = [p_ for p_ in params if p_.numel() < (1 << 22)]
            if pending:   # started from the backward hook, in parameter order
                assert len(pending) == len(big)
                works = [w_ for _, w_ in pending]
            else:
                works = [dist.all_reduce(p_.grad, op=dist.ReduceOp.AVG, async_op=True) for p_ in big]
            fl = flat[:sum(p_.numel() for p_ in small)]
            torch.cat([p_.grad.reshape(-1) for p_ in small], out=fl)
            w_small = dist.all_reduce(fl, op=dist.ReduceOp.AVG, async_op=True)
            mark()
            for i_, (p_, w_) in enumerate(zip(big, works)):   # generator i: wait for its all-reduce, update it
                w_.wait()
                if pending and p_.grad.data_ptr() != pending[i_][0].data_ptr():
                    p_.grad.copy_(pending[i_][0])      # (autograd normally adopts the buffer the backward returned)
                opts[i_].step()
            pending.clear()
            w_small.wait()                             # queued behind the big ones
            o = 0
            for q_ in small:
                q_.grad.copy_(fl[o:o + q_.numel()].view_as(q_))
                o += q_.numel()
        else:
            mark()
            for o_ in opts:
                o_.step()
        opt_rest.step()
        mark()
        return loss

    sampler = ClockSampler(local_rank)
    sampler.start()
    # warm-up: >= W steps and >= 0.6 s. Every step contains collectives, so all ranks must run the SAME number of
    # steps: rank 0 decides after each step and broadcasts the decision (a per-rank clock would let one rank leave
    # the loop a step earlier than its peers - a deadlock, seen once in a while at N = 2)
    t_w, n_warm = time.time(), 0
    go = torch.ones(1, dtype=torch.int32, device=device)
    while True:
        step()
        n_warm += 1
        more = n_warm < max(a.warmup, 3) or time.time() - t_w < 0.6
        if dist is not None:
            go.fill_(1 if more else 0)
            dist.broadcast(go, 0)
            more = bool(go.item())
        if not more:
            break
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tms = torch.tensor([ms], device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    phases = None
    if marks:
        m5 = marks[-5 * a.steps:]
        names = ["forward+loss", "backward", "small_bucket_enqueue", "sync+adam"]
        phases = {n: round(sum(m5[5 * i + k].elapsed_time(m5[5 * i + k + 1]) for i in range(a.steps)) / a.steps, 3)
                  for k, n in enumerate(names)}
    if rank == 0:
        if phases:
            sys.stderr.write("train phases (ms, rank 0): %s\n" % json.dumps(phases))
        line = {"metric": "affinity-head training frame-pairs/sec (200x200 pairs)", "value": world * B * a.steps / (ms / 1e3),
                "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": n_warm, "ms_per_step": ms / a.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "impl": "b200", "mode": "train", "loss": float(loss.detach()),
                "config": config_dict(a, {"differentiated_parameters": training.differentiable_parameter_names(),
                                          "note": "BASELINE.json config 5: every head parameter trained (258 M), shared_conv / trunk frozen"}),
                "clocks": clocks, "gpu_launches": (lib.shasta_last_launch_count() + 6) * a.steps}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    if a.train:
        run_train(a)
        return
    if a.workload == "multiclass":
        import bench_multiclass
        bench_multiclass.run(a)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the shasta_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        pin_rank_to_cores(local_rank, world)

    from shasta_b200 import _cabi, sharding
    lib = _cabi.lib()
    lib.shasta_set_option(_cabi.OPT_ANCHOR_PATH, a.anchor_path)
    lib.shasta_set_option(_cabi.OPT_TC_RAW_HI, a.raw_hi)
    lib.shasta_set_option(2, a.dbg)
    lib.shasta_set_option(3, a.splits)
    for kv in a.opt:
        k, v = kv.split("=")
        _cabi.check(lib.shasta_set_option(int(k, 0), int(v, 0)), "shasta_set_option")
    pc_start, d, bev, prev_bev = make_inputs(a, device, seed=1000 + rank)
    model = build_model(a, pc_start, device)
    det0 = torch.from_numpy(d["det_boxes"]).to(device)
    prev = torch.from_numpy(d["prev_det_boxes"]).to(device)
    det = det0.clone()
    B, M = a.batch, a.max_obj
    n_prev = torch.from_numpy(d["n_prev"].astype(np.int32)).to(device)
    n_det = torch.from_numpy(d["n_det"].astype(np.int32)).to(device)
    # every step ends with the device decode (eval.py:126-181) of its affinities; the compact result (4 int32 + 2
    # float32 planes of (B,M)) lands in one block per rank. The work per step is the SAME at every N; with N > 1 the
    # blocks of all ranks are all-gathered ONCE at the end of the timed region - the only collective of the path
    # The decode is fused into the softmax kernels (shasta_forward_decode_f32) and writes straight into this step's
    # slot of the ring (device-side call counter): no decode kernel, no copy.
    nl_ = max(1, a.streams) if not (a.no_graph or os.environ.get("BENCH_INNER_GRAPH")) else 1
    slots = (a.steps + nl_ - 1) // nl_        # ring slots per lane (steps go to the lanes round-robin)
    dec_pack = torch.empty((nl_, slots, 6, B, M), dtype=torch.int32, device=device)
    # one CUDA graph per step: box refresh + forward (+ decode when results are gathered across ranks); the model's own
    # per-call graph cache is not needed on top of it
    inner = bool(os.environ.get("BENCH_INNER_GRAPH"))   # A/B knob: the model's per-call graph instead of the step graph
    model.cuda_graphs = inner and not a.no_graph

    # lanes: independent batches in flight (--streams N): each lane has its own model instance (workspace), box buffer,
    # step graph and stream; steps go to the lanes round-robin. The default is one lane on the current stream.
    nl = max(1, a.streams) if not (a.no_graph or inner) else 1
    lanes = []
    for k in range(nl):
        mk = model if k == 0 else build_model(a, pc_start, device)
        mk.cuda_graphs = model.cuda_graphs
        lanes.append({"model": mk, "det": det if k == 0 else det0.clone(),
                      "ring": dec_pack[k],
                      "counter": torch.zeros(1, dtype=torch.int32, device=device),
                      "stream": torch.cuda.current_stream() if k == 0 else torch.cuda.Stream(), "g": None, "out": None})

    def lane_body(ln):
        ln["det"].copy_(det0)  # fresh boxes every step (the forward back-projects det_boxes in place)
        if a.bf16:   # bf16 mode: separate decode kernel into the latest slot
            m1, m2 = ln["model"].affinity(bev, prev_bev, ln["det"], prev)
            ln["model"].decode(m1, m2, n_prev, n_det, out=ln["ring"][0])
            return m1, m2
        return ln["model"].affinity(bev, prev_bev, ln["det"], prev,
                                    decode={"n_prev": n_prev, "n_det": n_det, "out": ln["ring"], "counter": ln["counter"]})

    def step_body():
        return lane_body(lanes[0])

    counter = [0]

    def step():
        ln = lanes[counter[0] % nl]
        counter[0] += 1
        if a.no_graph or inner:
            out = lane_body(ln)
        else:
            if ln["g"] is None:
                with torch.cuda.stream(ln["stream"]):
                    for _ in range(2):   # eager runs first: one-time kernel attribute set-up must not be captured
                        lane_body(ln)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ln["out"] = lane_body(ln)
                ln["g"] = g
            with torch.cuda.stream(ln["stream"]):
                ln["g"].replay()
            return ln["out"]
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # the clock sampler starts BEFORE the warm-up: nvidia-smi start-up can stall the GPU for milliseconds
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.5)
        # warm-up: at least W (>= 3) steps AND at least 0.6 s of sustained work - a freshly started process finds the
        # GPU in an idle power state and the first ~100 ms of kernels run several times slower (measured: 3.3 ms vs
        # 0.93 ms per step right after start-up, SM clock already reading 1965 MHz)
        step()   # builds the step graph (eager runs + capture: host time with an idle GPU, kept out of the warm-up clock)
        torch.cuda.synchronize()
        t_w = time.time()
        n_warm = 0
        while n_warm < max(a.warmup, 3) or time.time() - t_w < 0.6:
            step()
            n_warm += 1
            if n_warm % 8 == 0:
                torch.cuda.synchronize()
        launches_per_step = lib.shasta_last_launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record()
        for ln in lanes[1:]:
            ln["stream"].wait_event(e0)
        for it in range(a.steps):
            m1, m2 = step()
        for ln in lanes[1:]:
            torch.cuda.current_stream().wait_stream(ln["stream"])
        if world > 1:  # NCCL only gathers the per-rank results
            gathered = sharding.gather_rank_blocks(dec_pack)  # (world, steps, 6, B, M) on every rank
            assert gathered.shape[0] == world
        e1.record()
        barrier()
        t_end = time.time()
        if t_end - t_begin < 0.35:   # keep the GPU under the same load until a few clock samples exist
            while time.time() - t_begin < 0.35:
                step()
            torch.cuda.synchronize()
            t_end = time.time()
        clocks = sampler.stop(t_begin, t_end)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tms = torch.tensor([ms], device=device)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        value = world * B * a.steps / (ms / 1e3)

        # ---- per-kernel durations, live, same loop with event records between the kernels -------------------
        model.kernel_flags = a.flags | 0x100
        _cabi.check(lib.shasta_profile_begin(a.steps), "profile_begin")
        for _ in range(a.steps):
            step_body()
        stage_ms = (ctypes.c_float * 7)()
        nsteps = ctypes.c_int(0)
        _cabi.check(lib.shasta_profile_end(stage_ms, ctypes.byref(nsteps)), "profile_end")
        model.kernel_flags = a.flags
        stage_names = ["gather", "anchor_hidden", "anchor_finish", "project", "pairwise", "aff_row", "col_softmax"]
        stages = {n: float(stage_ms[i]) for i, n in enumerate(stage_names)}

        # ---- end to end through the public API with host buffers -----------------------------------------------
        e2e = None
        if not a.no_e2e:
            e2e = run_e2e(a, model, bev, prev_bev, det0, prev, device, dist, world, n_prev, n_det)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    dom = max(stages, key=stages.get)
    tc_path = a.anchor_path == 2 or (a.anchor_path == 0 and B > 4)
    bf16_w = a.bf16 and tc_path
    step_s = ms / a.steps / 1e3
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1391.5))   # dense bf16, sustained (a kernel inside a long step)
    # ---- the HBM-bound kernel: aug_shape.i.0 weight stream ------------------------------------------------------
    ab = algorithmic_bytes_anchor_hidden(M, B, bf16_w)
    ah_ms = stages["anchor_hidden"]
    achieved = ab / (ah_ms / 1e3) / 1e9 if ah_ms > 0 else 0.0
    traffic = None
    if tc_path and M == 200 and B == 64 and not a.bf16:   # dram bytes of one launch from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[
                "anchor_hidden_tc2_kernel<64>"]["traffic_bytes_per_launch"]
        except Exception:  # noqa: BLE001
            traffic = None
    hbm_kernel = {"kernel": ("anchor_hidden_bf16_kernel (aug_shape.i.0, 0.51 GB of bf16 weights)" if a.bf16 else
                             "anchor_hidden_tc2_kernel<64> (aug_shape.i.0, the 1.03 GB weight stream)") if tc_path
                  else "anchor_hidden_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                  "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": ab,
                  "ms_per_launch": ah_ms, "step_share": ah_ms / max(sum(stages.values()), 1e-9)}
    # ---- the pairwise kernel (largest stage by time at the default configuration): tensor pipe -------------------
    # algorithmic flops per pair (decomposed formulation): outer sum + ReLU over the 144 first-layer
    # columns, the three second layers, the third / fourth layers, the hand-designed residuals
    T = M + 2
    flop_pair = 2 * (40 * 20 + 20 * 10 + 10) + 2 * (72 * 18 + 18 * 3) + 2 * (32 * 8 + 8) + 2 * 144 + 30
    pw_flop = float(B) * T * T * flop_pair
    pw_ms = stages["pairwise"]
    pw_tf = pw_flop / (pw_ms / 1e3) / 1e12 if pw_ms > 0 else 0.0
    # the bound this kernel actually runs against (tools/micro/mma_rate.cu, profiles/README.md): a tcgen05.mma with
    # M = 128 and a TMEM A operand issues every 32.6 clocks at N <= 32, 42 at N = 64; per 128-pair tile (3xTF32 with the
    # hi | lo weight images merged along N): fuse_det 4 x (32.6 + 32.6), fuse_shape 5 x (42 + 32.6), res_coeff 9 x (42 + 32.6)
    sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    tiles = B * ((T + 7) // 8) * ((T + 15) // 16)
    mma_floor_ms = tiles * (8 * 32.6 + 14 * (42.0 + 32.6)) / 148.0 / (sm_mhz * 1e3)
    pair_kernel = {"kernel": "pairwise_tc_kernel<false> (second pairwise layers on tcgen05, 3xTF32)", "bound": "tensor",
                   "achieved": pw_tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": pw_tf / tensor_peak,
                   "traffic": None, "algorithmic_flop_per_launch": pw_flop, "ms_per_launch": pw_ms,
                   "step_share": pw_ms / max(sum(stages.values()), 1e-9),
                   "mma_issue_floor_ms": mma_floor_ms, "frac_of_mma_issue_floor": mma_floor_ms / pw_ms if pw_ms > 0 else None,
                   "note": "small-N tf32 UMMAs are issue-bound, not FLOP-bound: 32.6 / 42 clk per instruction (measured), "
                           "so the dense bf16 peak is the contract's denominator, the issue floor the attainable one"}
    pb = path_bytes(M, B, a.hw, bf16_w)
    dominant = pair_kernel if dom == "pairwise" else hbm_kernel
    roofline = dict(dominant)
    roofline.update({"peak_source": peak_src, "dominant_stage_by_time": dom, "stage_ms": stages,
                     # whole path against the HBM roofline (bytes(M,B) = W(M) + B IO(M) over the step time)
                     "path_bytes_per_step": pb, "path_hbm_gbs": pb / step_s / 1e9, "path_hbm_frac": pb / step_s / 1e9 / hbm_peak,
                     "kernels": {"pairwise": pair_kernel, "anchor_hidden": hbm_kernel}})

    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu:
            state = {k: v.detach().cpu() for k, v in model.state_dict().items() if not k.startswith("shared_conv")}
            cpu = cpu_reference_run(a, a.cpu_seconds, weights_state=state)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": n_warm, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if a.bf16 else "f32", "data": "synthetic", "impl": "b200",
                "config": config_dict(a, {"flags": a.flags, "anchor_path": a.anchor_path, "cuda_graph": not a.no_graph, "batches_in_flight": nl}), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * a.steps}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def pin_rank_to_cores(local_rank, world):
    """One slice of the host cores per rank: the e2e path is host-driven (pinned buffers, per-step enqueue), and N
    ranks time-sharing all cores double the per-step enqueue time at N = 8 (round-1 SCALE record)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 1:
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
            torch.set_num_threads(max(1, per))
    except Exception:  # noqa: BLE001
        pass


def run_e2e(a, model, bev, prev_bev, det0, prev, device, dist, world, n_prev=None, n_det=None):
    """Shasta.forward(example) with pinned HOST tensors: per step the boxes go H2D, the gather kernel samples the
    host-resident BEV maps over PCIe (zero-copy, only the taps move), matched1/matched2 come back D2H."""
    B, M = a.batch, a.max_obj
    h_bev = torch.empty(bev.shape, dtype=torch.float32, pin_memory=True)
    h_prev_bev = torch.empty(bev.shape, dtype=torch.float32, pin_memory=True)
    h_bev.copy_(bev)
    h_prev_bev.copy_(prev_bev)
    h_det0 = det0.cpu().pin_memory()
    h_prev = prev.cpu().pin_memory()
    # a ring of pinned buffer sets (per-step inputs that the forward modifies in place + outputs): a set is reused only
    # after the step that last used it has delivered its results to the host - what a streaming consumer does. The
    # model overlaps the PCIe-bound gather of step i+1 with the remaining stages and the D2H copies of step i.
    NBUF = 3
    sets = [{"det": h_det0.clone().pin_memory(),
             "m1": torch.empty((B, M, M + 2), dtype=torch.float32, pin_memory=True),
             "m2": torch.empty((B, M + 2, M), dtype=torch.float32, pin_memory=True),
             "dec": torch.empty((6, B, M), dtype=torch.int32, pin_memory=True),
             "dec_dev": torch.empty((6, B, M), dtype=torch.int32, device=device), "ev": None} for _ in range(NBUF)]
    counter = [0]

    def step(decode):
        st = sets[counter[0] % NBUF]
        counter[0] += 1
        if st["ev"] is not None:
            st["ev"].synchronize()
        st["det"].copy_(h_det0)
        example = {"det_boxes": st["det"], "prev_det_boxes": h_prev, "bev_feature": h_bev,
                   "prev_bev_feature": h_prev_bev}
        m1, m2, _ = model(example, train_mode=False)
        if decode:   # what the tracker downstream consumes: the thresholded argmax of eval.py:126-181 (0.3 MB)
            model.decode(m1, m2, n_prev, n_det, out=st["dec_dev"])
            st["dec"].copy_(st["dec_dev"], non_blocking=True)
        else:
            st["m1"].copy_(m1, non_blocking=True)
            st["m2"].copy_(m2, non_blocking=True)
        st["ev"] = torch.cuda.Event()
        st["ev"].record()

    def timed(decode):
        for _ in range(3):
            step(decode)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            step(decode)
        enqueue_ms = (time.perf_counter() - t0) * 1e3
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), wall * 1e3)
        if dist is not None:
            tms = torch.tensor([ms], device=device)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, enqueue_ms

    ms, enqueue_ms = timed(False)
    taps = 2 * B * 5 * M * 4 * 64 * 4
    h2d = 2 * B * M * 11 * 4 + taps
    d2h = B * (M * (M + 2) + (M + 2) * M) * 4 + B * M * 11 * 4
    res = {"value": world * B * a.steps / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms / a.steps, "host_enqueue_ms_per_step": enqueue_ms / a.steps,
           "note": "pinned host inputs; BEV maps sampled in place over PCIe (tap bytes counted), boxes copied, "
                   "matched1/matched2 and the back-projected boxes copied back; box upload + gather of step i+1 "
                   "overlap the other stages and the D2H copies of step i (side stream, two workspaces); a ring of "
                   "%d pinned buffer sets, each reused only after its results reached the host" % NBUF}
    if n_prev is not None:
        ms_d, enq_d = timed(True)
        res["decode_terminated"] = {
            "value": world * B * a.steps / (ms_d / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 6 * B * M * 4 + B * M * 11 * 4, "ms_per_step": ms_d / a.steps,
            "host_enqueue_ms_per_step": enq_d / a.steps,
            "note": "same inputs; the step ends with the device decode (eval.py:126-181) and only its (6,B,M) block "
                    "and the back-projected boxes go back to the host - what the tracker consumes"}
    return res


if __name__ == "__main__":
    main()
