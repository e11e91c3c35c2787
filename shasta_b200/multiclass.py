"""The per-class sequence batch (BASELINE.json configs[2]).

The reference tracks every detection class with its own Shasta model (``configs/nusc/<class>.py:26-29`` fixes
``max_objects`` per class; ``tools/nusc_shasta/eval.py:82-112`` builds that model and walks all frames of all scenes
with batch size 1). Frame pairs are independent of each other (the "tracks" of a pair are the previous frame's
detections), so here every class model runs over BATCHES of the rank's frame pairs, scenes are dealt round-robin to
ranks (sharding.py) and the compact decode output (six (M,) vectors per frame pair: states, argmax indices, scores -
eval.py:126-181) is the only thing gathered, once per class, at the end.

Host logic only: the numbers come from ``Shasta.affinity`` / ``Shasta.decode`` (CUDA through the C ABI).
"""
import torch

from . import sharding

# configs/nusc/{car,pedestrian,bus,truck,trailer,bicycle,motorcycle}.py:26-29 (max_objects of the shipped class models)
NUSC_CLASS_MAX_OBJ = (("car", 90), ("pedestrian", 90), ("bus", 20), ("truck", 60), ("trailer", 60), ("bicycle", 50),
                      ("motorcycle", 50))
# the three nuScenes classes the reference ships no config for; sized up to the detector's max_per_img = 500
# (configs/nusc/car.py:86) for the synthetic 10-class workload
SYNTHETIC_CLASS_MAX_OBJ = (("construction_vehicle", 200), ("barrier", 500), ("traffic_cone", 500))

DECODE_FIELDS = ("prev_state", "prev_argmax", "fn_dead_prob", "det_state", "det_argmax", "det_fp_prob")


class ClassLane:
    """One class model plus the step graphs of its batches. ``step_graphs``: capture (box refresh -> forward -> decode)
    once per distinct set of input addresses and replay it; meant for providers that serve batches from fixed buffers
    (a loader ring), pointless when every batch arrives in fresh tensors."""

    def __init__(self, name, model, step_graphs=False):
        self.name, self.model, self.step_graphs = name, model, step_graphs
        self._graphs = {}
        self._bufs = {}

    def _buffers(self, B, device):
        b = self._bufs.get(B)
        if b is None:
            M = self.model.max_obj
            b = self._bufs[B] = (torch.empty((B, M, 11), dtype=torch.float32, device=device),
                                 torch.empty((len(DECODE_FIELDS), B, M), dtype=torch.int32, device=device))
        return b

    def _body(self, batch, det_work, dec_out):
        # the forward back-projects the detections in place (shasta.py:270): work on a copy so that a provider's
        # buffers can be served again
        det_work.copy_(batch["det_boxes"])
        if self.model.bf16:
            m1, m2 = self.model.affinity(batch["bev"], batch["prev_bev"], det_work, batch["prev_det_boxes"])
            self.model.decode(m1, m2, batch["n_prev"], batch["n_det"], out=dec_out)
        else:   # the decode runs inside the softmax kernels (shasta_forward_decode_f32): no separate pass
            self.model.affinity(batch["bev"], batch["prev_bev"], det_work, batch["prev_det_boxes"],
                                decode={"n_prev": batch["n_prev"], "n_det": batch["n_det"], "out": dec_out})

    def step(self, batch):
        """``batch``: dict of CUDA tensors bev, prev_bev (B,H,W,64), det_boxes, prev_det_boxes (B,M,11), n_prev,
        n_det (B,) int32. Returns the (6, B, M) int32 decode block (float fields bit-cast), valid until the next step
        of this lane with the same B."""
        B = batch["det_boxes"].shape[0]
        device = batch["det_boxes"].device
        det_work, dec_out = self._buffers(B, device)
        if not self.step_graphs:
            self._body(batch, det_work, dec_out)
            return dec_out
        # everything the captured launches bake in: input addresses, the packed-weight buffer (re-allocated whenever a
        # parameter changes), the workspace, kernel flags and precision mode. A stale entry would replay against freed
        # weights or a freed workspace.
        model = self.model
        model._ensure_packed(device)
        key = (tuple(batch[k].data_ptr() for k in ("bev", "prev_bev", "det_boxes", "prev_det_boxes", "n_prev", "n_det"))
               + (B, model._pack_key, model._packed.data_ptr(), model._workspace(B, device).buf.data_ptr(),
                  int(model.kernel_flags), bool(model.bf16)))
        graph = self._graphs.get(key)
        if graph is None:
            if len(self._graphs) >= 64:
                self._graphs.clear()
            keep = self.model.cuda_graphs
            self.model.cuda_graphs = False        # the step graph replaces the model's own forward graph
            try:
                self._body(batch, det_work, dec_out)   # eager once: one-time kernel set-up stays outside the capture
                torch.cuda.current_stream(device).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._body(batch, det_work, dec_out)
            finally:
                self.model.cuda_graphs = keep
            self._graphs[key] = graph
        graph.replay()
        return dec_out


def run_sequence_batch(lanes, scene_lengths, provider, batch_pairs, world_size=1, rank=0, group=None, gather=True,
                       on_class_done=None):
    """Runs every class lane over this rank's frame pairs and gathers the decode blocks.

    ``lanes``: list of ClassLane. ``scene_lengths[s]``: frame pairs of scene s. ``provider(name, items)`` returns the
    batch dict (see ClassLane.step) for ``items`` = [(scene, frame_pair_index)] of one class. ``batch_pairs``: int or
    ``{class name: int}``. ``on_class_done(name)``: called after a class's work has been enqueued (timing hooks).

    Returns ``{name: [per scene (len_s, 6, M) int32 tensor]}`` in scene order (every rank holds the full result, like
    the reference's single process does) or, with ``gather=False``, ``{name: (6, n_local, M)}`` of this rank only."""
    items = sharding.frame_pairs_for_rank(scene_lengths, world_size, rank)
    padded = sharding.padded_count(scene_lengths, world_size)
    out = {}
    for lane in lanes:
        M = lane.model.max_obj
        device = lane.model.aff[0].weight.device
        bp = batch_pairs[lane.name] if isinstance(batch_pairs, dict) else batch_pairs
        block = torch.zeros((len(DECODE_FIELDS), padded, M), dtype=torch.int32, device=device)
        row = 0
        with torch.no_grad():
            for chunk in sharding.batches(items, bp):
                dec = lane.step(provider(lane.name, chunk))
                block[:, row:row + len(chunk)].copy_(dec)
                row += len(chunk)
        if not gather:
            out[lane.name] = block[:, :len(items)]
            if on_class_done is not None:
                on_class_done(lane.name)
            continue
        gathered = sharding.gather_rank_blocks(block.transpose(0, 1).contiguous(), group=group)   # (world, padded, 6, M)
        out[lane.name] = sharding.scatter_to_scene_order(gathered, scene_lengths, world_size)
        if on_class_done is not None:
            on_class_done(lane.name)
    return out


def decode_fields(block):
    """(..., 6, M) int32 block -> dict of named views (scores bit-cast back to float32)."""
    res = {}
    for i, k in enumerate(DECODE_FIELDS):
        v = block[..., i, :]
        res[k] = v.view(torch.float32) if k.endswith("prob") else v
    return res


class SyntheticProvider:
    """Seeded nuScenes-shape inputs for ``run_sequence_batch`` (SURVEY.md §8d generator, synthetic.make_frame_pairs).

    ``ring=None``: every (class, scene, frame pair) has its own boxes and every frame its own map (pair f of a scene
    reads maps f and f+1 of that scene), batches are assembled by indexing - small parity cases.
    ``ring=K``: K resident batches of boxes per class and one resident run of ``max batch + 1`` maps shared by all
    classes (pair i of a batch reads maps i and i+1), served round-robin from fixed addresses - the throughput
    workload, where the inputs stand for the loader ring a deployment would fill."""

    def __init__(self, class_max_obj, scene_lengths, hw, device, seed=0, ring=None, batch_pairs=64, pc_start=None):
        import numpy as np
        from . import synthetic
        self.device, self.ring, self.hw = device, ring, hw
        self.pc_start = pc_start if pc_start is not None else (-hw * 0.3, -hw * 0.3)
        self.scene_lengths = list(scene_lengths)
        self.offset = np.concatenate([[0], np.cumsum(self.scene_lengths)]).astype(np.int64)
        total = int(self.offset[-1])
        self.boxes = {}
        g = torch.Generator(device=device)
        g.manual_seed(seed)

        def maps(n):
            m = torch.empty((n, hw, hw, 64), dtype=torch.float32, device=device)
            for i in range(n):
                m[i].normal_(generator=g).relu_()
            return m

        for ci, (name, M) in enumerate(class_max_obj):
            bp = batch_pairs[name] if isinstance(batch_pairs, dict) else batch_pairs
            n = total if ring is None else ring * bp
            d = synthetic.make_frame_pairs(n, M, hw, hw, seed * 1000 + ci, pc_start=self.pc_start, with_maps=False)
            self.boxes[name] = {
                "det_boxes": torch.from_numpy(d["det_boxes"]).to(device),
                "prev_det_boxes": torch.from_numpy(d["prev_det_boxes"]).to(device),
                "n_prev": torch.from_numpy(d["n_prev"].astype(np.int32)).to(device),
                "n_det": torch.from_numpy(d["n_det"].astype(np.int32)).to(device), "bp": bp}
        if ring is None:
            # one map per frame: scene s owns frames offset[s] + s ... (len_s + 1 of them)
            self.maps = {name: maps(total + len(self.scene_lengths)) for name, _ in class_max_obj}
        else:
            bmax = max(v["bp"] for v in self.boxes.values())
            self.maps = maps(bmax + 1)
        self._turn = {name: 0 for name, _ in class_max_obj}

    def pair_index(self, scene, frame):
        return int(self.offset[scene]) + frame

    def __call__(self, name, items):
        bx = self.boxes[name]
        B = len(items)
        if self.ring is None:
            idx = torch.tensor([self.pair_index(s, f) for s, f in items], device=self.device)
            fidx = torch.tensor([self.pair_index(s, f) + s for s, f in items], device=self.device)
            m = self.maps[name]
            return {"bev": m[fidx + 1], "prev_bev": m[fidx], "det_boxes": bx["det_boxes"][idx],
                    "prev_det_boxes": bx["prev_det_boxes"][idx], "n_prev": bx["n_prev"][idx], "n_det": bx["n_det"][idx]}
        k = self._turn[name]
        self._turn[name] = (k + 1) % self.ring
        lo = k * bx["bp"]
        return {"bev": self.maps[1:B + 1], "prev_bev": self.maps[0:B], "det_boxes": bx["det_boxes"][lo:lo + B],
                "prev_det_boxes": bx["prev_det_boxes"][lo:lo + B], "n_prev": bx["n_prev"][lo:lo + B],
                "n_det": bx["n_det"][lo:lo + B]}


class DetectionFileProvider:
    """Serves ``run_sequence_batch`` from a binary detection file (detfile.py): scene s / frame pair f = frame
    ``scenes[s].first + f`` of the file, packed per class with that class's ``det_type`` filter and ``max_obj``.
    ``maps_for(name, frame_indices)`` returns the two (B,H,W,64) CUDA maps (current, previous) of those frames - in the
    reference they come from the frozen trunk + that class model's shared_conv."""

    def __init__(self, det_file, class_det_type, class_max_obj, maps_for, device, seed=0):
        self.df, self.det_type, self.max_obj = det_file, dict(class_det_type), dict(class_max_obj)
        self.maps_for, self.device, self.seed = maps_for, device, seed
        self.scenes = det_file.scenes()
        self.scene_lengths = [n for _, n in self.scenes]

    def frame_of(self, scene, frame):
        return self.scenes[scene][0] + frame

    def rng_for(self, i):
        import random
        return random.Random(self.seed * 1000003 + i)

    def pack(self, name, frame_indices):
        return self.df.frame_pair_batch(frame_indices, self.max_obj[name], self.det_type[name], rng_for=self.rng_for)

    def __call__(self, name, items):
        idx = [self.frame_of(s, f) for s, f in items]
        b = self.pack(name, idx)
        bev, prev_bev = self.maps_for(name, idx)
        t = lambda a: torch.from_numpy(a).to(self.device, non_blocking=True)  # noqa: E731
        return {"bev": bev, "prev_bev": prev_bev, "det_boxes": t(b["det_boxes"]), "prev_det_boxes": t(b["prev_det_boxes"]),
                "n_prev": t(b["n_prev"]), "n_det": t(b["n_det"])}

    def annotations(self, name, per_scene_blocks):
        """Gathered decode blocks of one class (``run_sequence_batch(...)[name]``) -> ``{token: annos}`` like the
        reference's eval loop emits (tools/nusc_shasta/eval.py:126-181, incl. the ``dead`` post-pass)."""
        from . import formats
        df = self.df
        results, dead_tracker = {}, {}
        for s, (first, n) in enumerate(self.scenes):
            fields = decode_fields(per_scene_blocks[s].cpu())
            packed = self.pack(name, list(range(first, first + n)))
            for f in range(n):
                i = first + f
                token = df.tokens[i]
                p = int(df.prev_index[i])
                prev_token = "" if p < 0 else df.tokens[p]
                dead_tracker.setdefault(token, {"dead_idx": [], "keep_idx": []})
                prev_cls = df.cls_info(packed["prev_rows"][f], prev_token)
                cur_cls = df.cls_info(packed["rows"][f], token)
                time_lag = float(packed["prev_det_boxes"][f, 0, 9])
                annos, dead_idx, keep = formats.annos_from_decode(
                    prev_cls, cur_cls, fields["prev_state"][f].numpy(), fields["fn_dead_prob"][f].numpy(),
                    fields["det_state"][f].numpy(), fields["det_fp_prob"][f].numpy(), token, time_lag)
                if len(prev_cls) > 0:
                    dead_tracker.setdefault(prev_token, {"dead_idx": [], "keep_idx": []})["dead_idx"].extend(dead_idx)
                if len(cur_cls) > 0:
                    dead_tracker[token]["keep_idx"] = keep
                results[token] = annos
        return formats.mark_dead(results, dead_tracker)
