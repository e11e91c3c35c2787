"""``Shasta`` — the affinity head behind the reference's TRACK registry (det3d/models/tracker/shasta.py:9-327).

Drop-in contract kept: class name, constructor arguments, sub-module construction through the
builder, parameter / state_dict names in PyTorch ``(out, in)`` layout, ``forward(example, train_mode=True)`` returning
``(matched1 (B,M,M+2), matched2 (B,M+2,M), example)``, the in-place back-projection of ``example["det_boxes"]``,
``example['bev_feature']`` and the ``newborn / fp / dead_trk / fn`` attributes.

What changed is everything under ``forward``: from the 64-channel channels-last BEV maps onward the path is five
hand-written CUDA kernels (gather, anchors, per-object projections, pairwise tiles, aff + dual softmax) reached
through the C ABI in ``include/shasta_b200.h``. There is no PyTorch/CPU fallback for that part.
"""
import ctypes

import torch
from torch import nn

from . import _cabi
from . import registry as builder
from .registry import TRACK

FLAG_TMA_GATHER = 1
PAIRWISE_DEFAULT, PAIRWISE_TF32X3, PAIRWISE_BF16, PAIRWISE_FFMA = 0, 1, 2, 3   # kernel_flags bits 4-7


class _Workspace:
    """Per-(device, batch) scratch the kernels carve into regions (see enum shasta_region)."""

    def __init__(self, batch, max_obj, device):
        lib = _cabi.lib()
        self.batch, self.max_obj = batch, max_obj
        self.nbytes = lib.shasta_workspace_bytes(batch, max_obj)
        self.buf = torch.empty(max(self.nbytes // 4, 1), dtype=torch.float32, device=device)

    def region(self, rid, numel):
        off = _cabi.lib().shasta_workspace_offset(self.batch, self.max_obj, rid)
        return self.buf[off:off + numel]


@TRACK.register_module
class Shasta(nn.Module):
    def __init__(
        self,
        reader,
        backbone,
        neck,
        bev_extractor,
        train_cfg=None,
        test_cfg=None,
        pretrained=None,
        max_obj=100,
        num_feats=7,
        in_channels=512,
        share_conv_channel=64,
        num_point=5,
    ):
        super().__init__()
        if num_point != 5 or share_conv_channel != 64:
            raise NotImplementedError(
                "shasta_b200 implements num_point=5, share_conv_channel=64 (every shipped config); got "
                "num_point=%r share_conv_channel=%r" % (num_point, share_conv_channel))
        if num_feats != 3:
            raise NotImplementedError(
                "shasta_b200 implements num_feats=3 (configs/nusc/*.py); got %r" % (num_feats,))
        # frozen CenterPoint trunk: built through the registry exactly like the reference (shasta.py:28-31);
        # it is outside this package, a cfg of None leaves the slot empty
        self.reader = builder.build_reader(reader) if reader is not None else None
        self.backbone = builder.build_backbone(backbone) if backbone is not None else None
        self.neck = builder.build_neck(neck) if neck is not None else None
        self.bev_extractor = builder.build_second_stage_module(bev_extractor)

        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.num_feats = num_feats
        self.max_obj = max_obj
        self.num_point = num_point

        # ---- parameters: identical module tree => identical state_dict keys (shasta.py:42-106) ----
        self.shared_conv = nn.Sequential(
            nn.Conv2d(in_channels, share_conv_channel, kernel_size=3, padding=1, bias=True),
            nn.BatchNorm2d(share_conv_channel),
            nn.ReLU(inplace=True),
        )
        F = share_conv_channel * num_point
        self.aug_shape_input = max_obj * F
        self.aug_shape_output = F

        def mlp(widths):
            layers = []
            for i in range(len(widths) - 1):
                layers.append(nn.Linear(widths[i], widths[i + 1]))
                if i + 2 < len(widths):
                    layers.append(nn.ReLU(inplace=True))
            return nn.Sequential(*layers)

        self.aug_shape = nn.ModuleList([mlp([max_obj * F, (max_obj * F) // 64, F]) for _ in range(4)])
        self.fuse_shape = mlp([2 * F, F // 8, F // 16, F // 32, 1])
        self.aug_input = max_obj * 7
        self.aug_dets = nn.ModuleList([mlp([max_obj * 7, (max_obj * 7) // 32, 7]) for _ in range(4)])
        self.fuse_det = mlp([num_feats * 2, 32, 8, 1])
        self.res_coeff = mlp([num_feats * 2 + 2 * F, 32 + F // 8, 8 + F // 32, 3])
        self.aff = mlp([max_obj + 2, 128, 64, 32, 64, 128, max_obj + 2])
        self.softmax1 = nn.Softmax(dim=2)  # kept for module-tree parity; the kernels do both softmaxes
        self.softmax2 = nn.Softmax(dim=1)

        self.init_weights(pretrained=pretrained)

        # kernel-side state (never part of the state_dict)
        self.kernel_flags = 0
        # cuda_graphs = True: the forward launch sequence is captured once per (input addresses, shapes, flags, weight
        # version) into a CUDA graph and replayed afterwards; matched1/matched2 then live in buffers owned by the graph
        # entry and are overwritten by the next replay of the same entry. Off by default (fresh outputs per call).
        self.cuda_graphs = False
        self._graphs = {}
        # host-resident (pinned) BEV maps are sampled in place over PCIe; with pipeline_host_inputs the box upload and
        # the gather of a call run on a side stream into one of two workspaces, so they overlap the remaining stages
        # (and the caller's device-to-host copies) of the previous call
        self.pipeline_host_inputs = True
        self._pipe = None
        # bf16 = True: inference runs shasta_forward_bf16 (bf16 copy of the aug_shape.i.0 weights and of the gathered
        # features for that GEMM, bf16 pairwise tiles; everything else fp32). Tolerance stated separately (2e-2).
        self.bf16 = False
        self._w16 = None
        self._w16_key = None
        self._packed = None
        self._pack_key = None
        self._cparams = None
        self._ws = {}

    # ------------------------------------------------------------------------------------------
    def init_weights(self, pretrained=None):
        """shasta.py:111-119: best-effort load, failures are printed and ignored."""
        if pretrained is None:
            return
        try:
            checkpoint = torch.load(pretrained, map_location="cpu")
            load_matching_state_dict(self, checkpoint.get("state_dict", checkpoint))
            print("init weight from {}".format(pretrained))
        except Exception:  # noqa: BLE001 - reference behaviour
            print("no pretrained model at {}".format(pretrained))

    @property
    def with_neck(self):
        return getattr(self, "neck", None) is not None

    def extract_feat(self, data):
        """shasta.py:164-210 — the frozen spconv trunk; not part of this package. Callers either provide the
        64-channel maps in ``example`` or attach a trunk (reader/backbone/neck) that yields (B,512,H,W) maps."""
        if self.backbone is None:
            raise RuntimeError(
                "Shasta.extract_feat: no trunk attached; put 'bev_feature' and 'prev_bev_feature' (B,H,W,64) into "
                "the example or build the model with reader/backbone/neck configs")
        feats = self.reader(data["voxels"], data["num_points"])
        prev_feats = self.reader(data["prev_voxels"], data["prev_num_points"])
        x, vf = self.backbone(feats, data["coordinates"], len(data["points"]), data["shape"][0])
        px, pvf = self.backbone(prev_feats, data["prev_coordinates"], len(data["prev_points"]), data["prev_shape"][0])
        if self.with_neck:
            x, px = self.neck(x), self.neck(px)
        return x, vf, px, pvf

    # ------------------------------------------------------------------------------------------
    def _head_params(self):
        seqs = [("aug_shape", self.aug_shape, (0, 2)), ("aug_dets", self.aug_dets, (0, 2))]
        out = {}
        for name, mods, idxs in seqs:
            for i in range(4):
                for li in idxs:
                    out["%s.%d.%d" % (name, i, li)] = mods[i][li]
        for name, seq, idxs in (("fuse_shape", self.fuse_shape, (0, 2, 4, 6)), ("fuse_det", self.fuse_det, (0, 2, 4)),
                                ("res_coeff", self.res_coeff, (0, 2, 4)), ("aff", self.aff, (0, 2, 4, 6, 8, 10))):
            for li in idxs:
                out["%s.%d" % (name, li)] = seq[li]
        return out

    def _ensure_packed(self, device):
        layers = self._head_params()
        key = tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version)
                    for l in layers.values())
        if key == self._pack_key and self._packed is not None and self._packed.device == device:
            return
        lib = _cabi.lib()
        for name, l in layers.items():
            for t in (l.weight, l.bias):
                if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                    raise _cabi.ShastaLibraryError(
                        "parameter %s must be a contiguous float32 tensor on %s (got %s %s)" %
                        (name, device, t.dtype, t.device))
        p = _cabi.ShastaParams()
        p.max_obj, p.num_feats = self.max_obj, self.num_feats
        for i in range(4):
            p.aug_shape_w0[i] = layers["aug_shape.%d.0" % i].weight.data_ptr()
            p.aug_shape_b0[i] = layers["aug_shape.%d.0" % i].bias.data_ptr()
            p.aug_shape_w2[i] = layers["aug_shape.%d.2" % i].weight.data_ptr()
            p.aug_shape_b2[i] = layers["aug_shape.%d.2" % i].bias.data_ptr()
            p.aug_dets_w0[i] = layers["aug_dets.%d.0" % i].weight.data_ptr()
            p.aug_dets_b0[i] = layers["aug_dets.%d.0" % i].bias.data_ptr()
            p.aug_dets_w2[i] = layers["aug_dets.%d.2" % i].weight.data_ptr()
            p.aug_dets_b2[i] = layers["aug_dets.%d.2" % i].bias.data_ptr()
        for n, li in enumerate((0, 2, 4, 6)):
            p.fuse_shape_w[n] = layers["fuse_shape.%d" % li].weight.data_ptr()
            p.fuse_shape_b[n] = layers["fuse_shape.%d" % li].bias.data_ptr()
        for n, li in enumerate((0, 2, 4)):
            p.fuse_det_w[n] = layers["fuse_det.%d" % li].weight.data_ptr()
            p.fuse_det_b[n] = layers["fuse_det.%d" % li].bias.data_ptr()
            p.res_coeff_w[n] = layers["res_coeff.%d" % li].weight.data_ptr()
            p.res_coeff_b[n] = layers["res_coeff.%d" % li].bias.data_ptr()
        for n, li in enumerate((0, 2, 4, 6, 8, 10)):
            p.aff_w[n] = layers["aff.%d" % li].weight.data_ptr()
            p.aff_b[n] = layers["aff.%d" % li].bias.data_ptr()
        nbytes = lib.shasta_packed_weight_bytes(self.max_obj, self.num_feats)
        packed = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            rc = lib.shasta_pack_weights(ctypes.byref(p), packed.data_ptr(), nbytes,
                                         ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _cabi.check(rc, "shasta_pack_weights")
        self._packed, self._pack_key, self._cparams = packed, key, p

    def _ensure_bf16_anchor_weights(self, device):
        """bf16 copies of the four aug_shape.i.0 matrices (half of their 4*5M*320M*4 bytes), cached per weight version."""
        ws0 = [self.aug_shape[i][0].weight for i in range(4)]
        key = tuple((w.data_ptr(), w._version) for w in ws0)
        if key == self._w16_key and self._w16 is not None and self._w16.device == device:
            return
        lib = _cabi.lib()
        nbytes = lib.shasta_anchor_bf16_bytes(self.max_obj)
        if self._w16 is None or self._w16.numel() * 2 != nbytes or self._w16.device != device:
            self._w16 = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=device)
        with torch.cuda.device(device):
            rc = lib.shasta_pack_anchor_bf16(ctypes.byref(self._cparams), self._w16.data_ptr(), nbytes,
                                             ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _cabi.check(rc, "shasta_pack_anchor_bf16")
        self._w16_key = key

    def _workspace(self, batch, device):
        """One scratch buffer per (batch size, device). Workspaces are kept for the life of the model: captured CUDA
        graphs (the model's own and the class lanes') hold their addresses, so evicting one would leave a graph
        replaying into freed memory. ``release_workspaces()`` drops them together with the graphs."""
        k = (batch, device)
        ws = self._ws.get(k)
        if ws is None:
            ws = self._ws[k] = _Workspace(batch, self.max_obj, device)
        return ws

    def release_workspaces(self):
        """Frees every workspace and every captured graph that references one (call between workloads of very
        different batch sizes)."""
        self._graphs.clear()
        self._ws.clear()
        self._pipe = None

    def invalidate_packed(self):
        """Forces a re-pack of the kernel-side weight cache on the next call. The cache is keyed on each parameter's
        (data_ptr, _version); in-place updates through ``param.data`` (EMA, some AMP / clipping wrappers) do NOT bump
        ``_version`` - call this after such an update (``load_state_dict`` and ``train()`` do it automatically)."""
        self._pack_key = None
        self._w16_key = None
        self._graphs.clear()

    def load_state_dict(self, *args, **kwargs):
        res = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed()
        return res

    def train(self, mode=True):
        self.invalidate_packed()
        return super().train(mode)

    # ------------------------------------------------------------------------------------------
    def affinity(self, bev, prev_bev, det_boxes, prev_det_boxes, decode=None):
        """Hot path from the channels-last maps: (B,H,W,64) x2, (B,M,11) x2 -> matched1, matched2.
        ``det_boxes[:, :, :2]`` is back-projected in place (shasta.py:270).

        ``decode`` (inference with device-resident inputs): dict(n_prev=, n_det= int32 CUDA tensors (B,), out= int32
        CUDA tensor (nslots, 6, B, M) or (6, B, M), counter= optional int32 CUDA scalar). The consumer decode of
        tools/nusc_shasta/eval.py:126-181 then runs inside the softmax kernels (``shasta_forward_decode_f32``) and fills
        slot ``counter % nslots`` of ``out`` (fields as in ``Shasta.decode``); the counter is incremented on the device,
        so a captured graph of the call fills consecutive slots on consecutive replays."""
        device = self.aff[0].weight.device
        if device.type != "cuda":
            raise _cabi.ShastaLibraryError("Shasta parameters are on %s: shasta_b200 has no CPU path" % device)
        for name, t in (("bev_feature", bev), ("prev_bev_feature", prev_bev), ("det_boxes", det_boxes),
                        ("prev_det_boxes", prev_det_boxes)):
            if t.dtype != torch.float32:
                raise TypeError("%s must be float32, got %s" % (name, t.dtype))
            if not t.is_cuda and not t.is_pinned():
                # host inputs are accepted only as page-locked buffers: boxes are copied asynchronously, BEV maps are
                # sampled in place over PCIe (the gather touches ~2 MB of a 134 MB map pair)
                raise _cabi.ShastaLibraryError(
                    "%s must be a CUDA tensor or a pinned host tensor: shasta_b200 has no CPU path" % name)
        B, H, W, C = bev.shape
        M = self.max_obj
        if C != 64 or prev_bev.shape != bev.shape:
            raise ValueError("bev maps must both be (B,H,W,64); got %s / %s" % (tuple(bev.shape), tuple(prev_bev.shape)))
        if tuple(det_boxes.shape) != (B, M, 11) or tuple(prev_det_boxes.shape) != (B, M, 11):
            raise ValueError("boxes must be (B=%d, max_obj=%d, 11); got %s / %s" %
                             (B, M, tuple(det_boxes.shape), tuple(prev_det_boxes.shape)))
        if B == 0:
            return (torch.empty((0, M, M + 2), dtype=torch.float32, device=device),
                    torch.empty((0, M + 2, M), dtype=torch.float32, device=device))
        bev = bev if bev.is_contiguous() else bev.contiguous()
        prev_bev = prev_bev if prev_bev.is_contiguous() else prev_bev.contiguous()
        if (not bev.is_cuda or not prev_bev.is_cuda) and (self.kernel_flags & FLAG_TMA_GATHER):
            raise _cabi.ShastaLibraryError("host-resident BEV maps need the LDG sampler (kernel_flags bit 0 clear)")
        if decode is not None and (not bev.is_cuda or not prev_bev.is_cuda or (torch.is_grad_enabled() and self.training)):
            raise _cabi.ShastaLibraryError("the fused decode is an inference option for device-resident maps; use "
                                           "Shasta.decode on the outputs otherwise")
        if (self.pipeline_host_inputs and not self.bf16 and not bev.is_cuda and not prev_bev.is_cuda and not det_boxes.is_cuda
                and not prev_det_boxes.is_cuda and not (self.kernel_flags & 0x100)
                and not (torch.is_grad_enabled() and self.training)):
            return self._affinity_pipelined(bev, prev_bev, det_boxes, prev_det_boxes, device)
        # boxes: device copies of host inputs (async from pinned memory); the back-projection is written back below
        prev_c = prev_det_boxes.to(device, non_blocking=True).contiguous()
        det_c = det_boxes.to(device, non_blocking=True).contiguous()

        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.aff.parameters()):
            # training configuration: same forward kernels on a private workspace + CUDA backward (training.py)
            if (bev.requires_grad or prev_bev.requires_grad) and not _cabi.lib_has("shasta_backward_maps_f32"):
                # the reference trains shared_conv with the head (train.py:184-191 freezes backbone and neck only)
                raise _cabi.ShastaLibraryError(
                    "the BEV maps require grad (shared_conv is being trained) but this build of libshasta_b200 has no "
                    "map-gradient kernels: shared_conv would silently stay at its initialisation")
            from .training import affinity_with_grad
            m1, m2 = affinity_with_grad(self, bev, prev_bev, det_c, prev_c)
        else:
            m1, m2, _ = self._launch_forward(bev, prev_bev, det_c, prev_c, self._workspace(B, device), decode)
        if det_c is not det_boxes:
            if det_boxes.is_cuda:
                det_boxes[:, :, :2] = det_c[:, :, :2]
            else:  # pinned host input: asynchronous write-back of the back-projected boxes (shasta.py:270)
                det_boxes.copy_(det_c, non_blocking=True)
        return m1, m2

    def _affinity_pipelined(self, bev, prev_bev, det_boxes, prev_det_boxes, device):
        """Host inputs, two-stage software pipeline over calls: stage 1 (side stream) uploads the boxes and gathers the
        box features straight from the pinned maps; stage 2 (current stream) runs anchors ... softmax. The two stages
        of consecutive calls use alternating workspaces and overlap. Semantics are those of asynchronous CUDA work on
        pinned memory: inputs must stay unchanged until the gather ran, results (and the in-place back-projection of
        ``det_boxes``) are complete once the current stream reaches this point."""
        lib = _cabi.lib()
        B, H, W, _ = bev.shape
        M = self.max_obj
        self._ensure_packed(device)
        cur = torch.cuda.current_stream(device)
        key = (B, device)
        if self._pipe is None or self._pipe["key"] != key:
            self._pipe = {"key": key, "idx": 0, "stream": torch.cuda.Stream(device=device),
                          "ws": [_Workspace(B, M, device) for _ in range(2)],
                          "det": [torch.empty((B, M, 11), dtype=torch.float32, device=device) for _ in range(2)],
                          "prev": [torch.empty((B, M, 11), dtype=torch.float32, device=device) for _ in range(2)],
                          "gathered": [torch.cuda.Event() for _ in range(2)],
                          "consumed": [None, None]}
        P = self._pipe
        k = P["idx"]
        P["idx"] = 1 - k
        side, ws, det_c, prev_c = P["stream"], P["ws"][k], P["det"][k], P["prev"][k]
        geom = self.bev_extractor.geom(H, W)
        if P["consumed"][k] is not None:
            side.wait_event(P["consumed"][k])      # stage 2 of the call that used this buffer set has finished
        with torch.cuda.stream(side):
            det_c.copy_(det_boxes, non_blocking=True)
            prev_c.copy_(prev_det_boxes, non_blocking=True)
            rc = lib.shasta_gather_pair_f32(bev.data_ptr(), prev_bev.data_ptr(), det_c.data_ptr(), prev_c.data_ptr(), B, M,
                                            ctypes.byref(geom), ws.buf.data_ptr(), ws.nbytes,
                                            (int(self.kernel_flags) & ~3) | _cabi.FLAG_NARROW_GATHER,  # PCIe-bound: few CTAs suffice
                                            ctypes.c_void_p(side.cuda_stream))
            _cabi.check(rc, "shasta_gather_pair_f32")
            P["gathered"][k].record(side)
        cur.wait_event(P["gathered"][k])
        m1 = torch.empty((B, M, M + 2), dtype=torch.float32, device=device)
        m2 = torch.empty((B, M + 2, M), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            rc = lib.shasta_forward_f32(
                ctypes.byref(self._cparams), self._packed.data_ptr(), bev.data_ptr(), prev_bev.data_ptr(),
                det_c.data_ptr(), prev_c.data_ptr(), B, ctypes.byref(geom), ws.buf.data_ptr(), ws.nbytes,
                m1.data_ptr(), m2.data_ptr(), int(self.kernel_flags) | _cabi.FLAG_SKIP_GATHER,
                ctypes.c_void_p(cur.cuda_stream))
        _cabi.check(rc, "shasta_forward_f32")
        det_boxes.copy_(det_c, non_blocking=True)  # back-projected boxes to the caller's pinned tensor (shasta.py:270)
        ev = torch.cuda.Event()
        ev.record(cur)
        P["consumed"][k] = ev
        anchors = ws.region(_cabi.WS_ANCHOR_BOX, B * 4 * 7).view(B, 4, 7)
        self.newborn, self.fp = anchors[:, 0:1, :], anchors[:, 1:2, :]
        self.dead_trk, self.fn = anchors[:, 2:3, :], anchors[:, 3:4, :]
        return m1, m2

    # ------------------------------------------------------------------------------------------
    def shared_conv_nhwc(self, x, maps_per_launch=8):
        """``self.shared_conv(x).permute(0, 2, 3, 1).contiguous()`` (shasta.py:223-228) for inference: one tcgen05
        implicit-GEMM kernel (3xTF32, fp32-equivalent) with the bias / BatchNorm (running statistics) / ReLU folded
        into its epilogue, writing the channels-last map directly. x: (N,512,H,W) float32 CUDA tensor.
        In training mode BatchNorm needs batch statistics: that case stays on the nn.Sequential under autograd; the CUDA
        head hands the gradients of the two maps back (``shasta_backward_maps_f32``), so shared_conv is trained with
        the head like in the reference (train.py:184-191)."""
        conv, bn = self.shared_conv[0], self.shared_conv[1]
        if self.training or not x.is_cuda:
            if not x.is_cuda:
                raise _cabi.ShastaLibraryError("shared_conv_nhwc needs a CUDA tensor: shasta_b200 has no CPU path")
            return self.shared_conv(x).permute(0, 2, 3, 1).contiguous()
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 512 or conv.out_channels != 64:
            raise ValueError("shared_conv_nhwc expects a float32 (N,512,H,W) tensor and 64 output channels")
        lib = _cabi.lib()
        device = x.device
        x = x if x.is_contiguous() else x.contiguous()
        N, _, H, W = x.shape
        tensors = (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        if getattr(self, "_conv_key", None) != key:
            nbytes = lib.shasta_shared_conv_packed_bytes()
            packed = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                rc = lib.shasta_shared_conv_pack(*[t.detach().contiguous().data_ptr() for t in tensors],
                                                 ctypes.c_float(bn.eps), packed.data_ptr(), nbytes, stream)
            _cabi.check(rc, "shasta_shared_conv_pack")
            self._conv_packed, self._conv_key = packed, key
        out = torch.empty((N, H, W, 64), dtype=torch.float32, device=device)
        step = max(1, min(N, int(maps_per_launch)))
        sbytes = lib.shasta_shared_conv_scratch_bytes(step, H, W)
        scratch = getattr(self, "_conv_scratch", None)
        if scratch is None or scratch.numel() * 4 < sbytes or scratch.device != device:
            scratch = self._conv_scratch = torch.empty(sbytes // 4, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            for n0 in range(0, N, step):
                n = min(step, N - n0)
                rc = lib.shasta_shared_conv_f32(self._conv_packed.data_ptr(), x[n0:n0 + n].data_ptr(), n, H, W,
                                                scratch.data_ptr(), scratch.numel() * 4, out[n0:n0 + n].data_ptr(),
                                                stream)
                _cabi.check(rc, "shasta_shared_conv_f32")
        return out

    def decode(self, matched1, matched2, n_prev, n_det, out=None):
        """The thresholded argmax of the eval loop (tools/nusc_shasta/eval.py:126-181) for a whole batch on the device:
        returns int32/float32 CUDA tensors of shape (B, M): ``prev_state`` (0 keep, 1 dead, 2 FN, -1 padding),
        ``prev_argmax``, ``fn_dead_prob`` (matched1[n,-2] of FN rows), ``det_state`` (0 keep, 1 newborn, 2 dropped FP,
        -1 padding), ``det_argmax``, ``det_fp_prob`` (matched2[-1,k] of kept detections; the reference's
        ``ref_detection_score`` is ``1 - value``, formed on the host in double). ``n_prev`` / ``n_det``: real counts per
        frame pair (sequence or int32 tensor). ``out``: optional contiguous (6, B, M) int32 CUDA tensor the six fields
        are written into in the order above (float fields bit-cast); the returned dict then holds views of it."""
        if not matched1.is_cuda or not matched2.is_cuda:
            raise _cabi.ShastaLibraryError("decode needs CUDA tensors: shasta_b200 has no CPU path")
        dev = matched1.device
        B, M = matched1.shape[0], matched1.shape[1]
        m1, m2 = matched1.contiguous(), matched2.contiguous()
        npv = torch.as_tensor(n_prev, dtype=torch.int32).to(dev)
        ndv = torch.as_tensor(n_det, dtype=torch.int32).to(dev)
        fields = ("prev_state", "prev_argmax", "fn_dead_prob", "det_state", "det_argmax", "det_fp_prob")
        if out is None:
            out = {k: torch.empty((B, M), dtype=(torch.float32 if k.endswith("prob") else torch.int32), device=dev)
                   for k in fields}
        else:
            if (tuple(out.shape) != (6, B, M) or out.dtype != torch.int32 or out.device != dev
                    or not out.is_contiguous()):
                raise ValueError("decode: out must be a contiguous (6, %d, %d) int32 tensor on %s" % (B, M, dev))
            out = {k: (out[i].view(torch.float32) if k.endswith("prob") else out[i]) for i, k in enumerate(fields)}
        with torch.cuda.device(dev):
            rc = _cabi.lib().shasta_decode_f32(
                m1.data_ptr(), m2.data_ptr(), npv.data_ptr(), ndv.data_ptr(), B, M, out["prev_state"].data_ptr(),
                out["prev_argmax"].data_ptr(), out["fn_dead_prob"].data_ptr(), out["det_state"].data_ptr(),
                out["det_argmax"].data_ptr(), out["det_fp_prob"].data_ptr(),
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _cabi.check(rc, "shasta_decode_f32")
        return out

    def _decode_struct(self, decode, B, device):
        M = self.max_obj
        out = decode["out"]
        if out.dtype != torch.int32 or out.device != device or not out.is_contiguous() or \
                tuple(out.shape[-3:]) != (6, B, M) or out.dim() not in (3, 4):
            raise ValueError("decode['out'] must be a contiguous int32 CUDA tensor of shape ([nslots,] 6, %d, %d)" % (B, M))
        d = _cabi.ShastaDecodeOut()
        for k in ("n_prev", "n_det"):
            t = decode[k]
            if t.dtype != torch.int32 or t.device != device or t.numel() != B or not t.is_contiguous():
                raise ValueError("decode['%s'] must be a contiguous int32 CUDA tensor with %d entries" % (k, B))
            setattr(d, k, t.data_ptr())
        d.out = out.data_ptr()
        d.nslots = out.shape[0] if out.dim() == 4 else 1
        d.slot_stride = 6 * B * M
        counter = decode.get("counter")
        if counter is not None and (counter.dtype != torch.int32 or counter.device != device or counter.numel() != 1):
            raise ValueError("decode['counter'] must be an int32 CUDA scalar")
        d.counter = counter.data_ptr() if counter is not None else None
        return d

    def _launch_forward(self, bev, prev_bev, det_c, prev_c, ws, decode=None):
        """Enqueues the five forward kernels on the current stream. Inputs are validated, contiguous, boxes on the
        device; ``ws`` is the workspace the activations are left in (the backward pass reads them)."""
        device = det_c.device
        B, H, W, _ = bev.shape
        M = self.max_obj
        lib = _cabi.lib()
        self._ensure_packed(device)
        geom = self.bev_extractor.geom(H, W)

        use_bf16 = self.bf16 and not (torch.is_grad_enabled() and self.training)
        if use_bf16:
            self._ensure_bf16_anchor_weights(device)
        dstruct = None
        if decode is not None:
            if use_bf16 or not bev.is_cuda:
                raise _cabi.ShastaLibraryError("the fused decode needs fp32 mode and device-resident inputs")
            dstruct = self._decode_struct(decode, B, device)

        def enqueue(m1, m2):
            with torch.cuda.device(device):
                stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
                if use_bf16:
                    rc = lib.shasta_forward_bf16(
                        ctypes.byref(self._cparams), self._packed.data_ptr(), self._w16.data_ptr(), bev.data_ptr(),
                        prev_bev.data_ptr(), det_c.data_ptr(), prev_c.data_ptr(), B, ctypes.byref(geom),
                        ws.buf.data_ptr(), ws.nbytes, m1.data_ptr(), m2.data_ptr(), int(self.kernel_flags), stream)
                elif dstruct is not None:
                    rc = lib.shasta_forward_decode_f32(
                        ctypes.byref(self._cparams), self._packed.data_ptr(), bev.data_ptr(), prev_bev.data_ptr(),
                        det_c.data_ptr(), prev_c.data_ptr(), B, ctypes.byref(geom), ws.buf.data_ptr(), ws.nbytes,
                        m1.data_ptr(), m2.data_ptr(), int(self.kernel_flags), ctypes.byref(dstruct), stream)
                else:
                    rc = lib.shasta_forward_f32(
                        ctypes.byref(self._cparams), self._packed.data_ptr(), bev.data_ptr(), prev_bev.data_ptr(),
                        det_c.data_ptr(), prev_c.data_ptr(), B, ctypes.byref(geom), ws.buf.data_ptr(), ws.nbytes,
                        m1.data_ptr(), m2.data_ptr(), int(self.kernel_flags), stream)
            _cabi.check(rc, "shasta_forward_bf16" if use_bf16 else "shasta_forward_f32")

        # (not while training: the packed weights change every optimizer step, a capture would never be replayed)
        use_graph = (self.cuda_graphs and not (self.kernel_flags & 0x100) and bev.is_cuda and prev_bev.is_cuda
                     and not self.training and not torch.cuda.is_current_stream_capturing())
        if use_graph:
            key = (bev.data_ptr(), prev_bev.data_ptr(), det_c.data_ptr(), prev_c.data_ptr(), B, H, W,
                   int(self.kernel_flags), ws.buf.data_ptr(), self._pack_key, use_bf16,
                   None if decode is None else (decode["out"].data_ptr(), tuple(decode["out"].shape),
                                                decode["n_prev"].data_ptr(), decode["n_det"].data_ptr(),
                                                None if decode.get("counter") is None else decode["counter"].data_ptr()))
            entry = self._graphs.get(key)
            if entry is None:
                if len(self._graphs) >= 16:
                    self._graphs.clear()
                m1 = torch.empty((B, M, M + 2), dtype=torch.float32, device=device)
                m2 = torch.empty((B, M + 2, M), dtype=torch.float32, device=device)
                keep = det_c.clone()      # the forward back-projects det_c in place: capture must not change it twice
                enqueue(m1, m2)           # eager run: one-time kernel attribute set-up happens outside the capture
                det_c.copy_(keep)
                torch.cuda.current_stream(device).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    enqueue(m1, m2)
                entry = self._graphs[key] = (graph, m1, m2)
            graph, m1, m2 = entry
            graph.replay()
        else:
            m1 = torch.empty((B, M, M + 2), dtype=torch.float32, device=device)
            m2 = torch.empty((B, M + 2, M), dtype=torch.float32, device=device)
            enqueue(m1, m2)
        anchors = ws.region(_cabi.WS_ANCHOR_BOX, B * 4 * 7).view(B, 4, 7)
        self.newborn, self.fp = anchors[:, 0:1, :], anchors[:, 1:2, :]
        self.dead_trk, self.fn = anchors[:, 2:3, :], anchors[:, 3:4, :]
        return m1, m2, ws

    def forward(self, example, train_mode=True, **kwargs):
        """shasta.py:213-327. ``example`` needs ``det_boxes`` and ``prev_det_boxes`` (B,M,11) and either the
        precomputed channels-last maps (``bev_feature`` + ``prev_bev_feature``, the metric's timed-region entry)
        or inputs for an attached trunk."""
        if "bev_feature" in example and "prev_bev_feature" in example:
            bev, prev_bev = example["bev_feature"], example["prev_bev_feature"]
        else:
            bev_map, _, prev_bev_map, _ = self.extract_feat(example)
            bev = self.shared_conv_nhwc(bev_map)
            prev_bev = self.shared_conv_nhwc(prev_bev_map)
            example["bev_feature"] = bev
        matched1, matched2 = self.affinity(bev, prev_bev, example["det_boxes"], example["prev_det_boxes"])
        return matched1, matched2, example


def load_matching_state_dict(module, state_dict):
    """det3d/torchie/trainer/checkpoint.py:67-107 semantics: copy tensors whose name AND shape match, skip the
    rest silently; returns the list of skipped keys."""
    own = module.state_dict()
    skipped = []
    with torch.no_grad():
        for name, value in state_dict.items():
            if name.startswith("module."):
                name = name[len("module."):]
            if name in own and tuple(own[name].shape) == tuple(value.shape):
                own[name].copy_(value)
            else:
                skipped.append(name)
    return skipped
