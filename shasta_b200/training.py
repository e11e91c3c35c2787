"""Training configuration (BASELINE.json config 5): forward on a private workspace + hand-written CUDA backward.

The reference trains the head with plain autograd (tools/nusc_shasta/train.py:195-215: forward, the two masked
cross-entropy terms on matched1/matched2, ``loss.backward()``, Adam). Here the forward is the same set of kernels as
inference; ``torch.autograd`` only sees one node whose backward calls ``shasta_backward_f32``:
dual-softmax backward, the aff row-MLP backward, the pairwise-MLP backward and the first-layer weight gradients.

Gradient coverage of this revision: ``aff.*``, ``fuse_shape.*``, ``res_coeff.*``, ``fuse_det.*`` (the pairwise MLPs
incl. their decomposed first layers), ``aug_shape.*`` (the anchor shape generators, 99 % of the parameters) and
``aug_dets.*`` (anchor boxes, through the first-layer box columns and the hand-designed residual incl. the F.normalize
backward) - i.e. every parameter of the head - and the two channels-last BEV maps (``shasta_backward_maps_f32``:
d feature through the first layers and aug_shape.i.0, scattered back through the bilinear taps), so that autograd trains
``shared_conv`` with the head exactly as the reference does (train.py:184-191 freezes only backbone and neck).
"""
import ctypes

import torch

from . import _cabi

AFF_LAYERS = (0, 2, 4, 6, 8, 10)
GROUPS = (("aff", AFF_LAYERS), ("fuse_shape", (0, 2, 4, 6)), ("fuse_det", (0, 2, 4)), ("res_coeff", (0, 2, 4)))


def differentiable_parameters(model):
    """Parameters that receive gradients from the CUDA backward, in the order the autograd node expects them:
    aff.*, fuse_shape.*, fuse_det.*, res_coeff.* (weight, bias per layer), then aug_shape.{0..3}.{0,2} and
    aug_dets.{0..3}.{0,2} - every trainable tensor of the head except shared_conv."""
    out = []
    for name, layers in GROUPS:
        seq = getattr(model, name)
        for li in layers:
            out += [seq[li].weight, seq[li].bias]
    for mods in (model.aug_shape, model.aug_dets):
        for i in range(4):
            for li in (0, 2):
                out += [mods[i][li].weight, mods[i][li].bias]
    return out


def differentiable_parameter_names():
    names = ["%s.%d.%s" % (name, li, k) for name, layers in GROUPS for li in layers for k in ("weight", "bias")]
    for grp in ("aug_shape", "aug_dets"):
        names += ["%s.%d.%d.%s" % (grp, i, li, k) for i in range(4) for li in (0, 2) for k in ("weight", "bias")]
    return names


class _AffinityFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, bev, prev_bev, det_c, prev_c, *params):
        from .shasta import _Workspace
        B = bev.shape[0]
        ws = _Workspace(B, model.max_obj, det_c.device)   # private: the backward reads the saved activations
        m1, m2, _ = model._launch_forward(bev, prev_bev, det_c, prev_c, ws)
        ctx.model, ctx.ws, ctx.batch = model, ws, B
        ctx.boxes = (det_c, prev_c)          # det_c is back-projected by now; the raw x,y live in the workspace
        ctx.map_shape = tuple(bev.shape)
        ctx.save_for_backward(m1, m2)
        return m1, m2

    @staticmethod
    def backward(ctx, gm1, gm2):
        model, ws, B = ctx.model, ctx.ws, ctx.batch
        m1, m2 = ctx.saved_tensors
        gm1 = torch.zeros_like(m1) if gm1 is None else gm1.contiguous().float()
        gm2 = torch.zeros_like(m2) if gm2 is None else gm2.contiguous().float()
        lib = _cabi.lib()
        params = differentiable_parameters(model)
        n_small = 2 * sum(len(layers) for _, layers in GROUPS)
        # aff / pairwise gradients are accumulated by the kernels (zero-initialised here); the aug_shape gradients
        # (1 GB at M = 200) are assigned by the kernels, so they are allocated uninitialised
        grads = [torch.zeros_like(p) for p in params[:n_small]] + [torch.empty_like(p) for p in params[n_small:]]
        g = _cabi.ShastaGrads()
        it = iter(grads)
        for name, layers in GROUPS:
            gw, gb = getattr(g, name + "_w"), getattr(g, name + "_b")
            for n in range(len(layers)):
                gw[n] = next(it).data_ptr()
                gb[n] = next(it).data_ptr()
        for grp in ("aug_shape", "aug_dets"):
            for i in range(4):
                getattr(g, grp + "_w0")[i] = next(it).data_ptr()
                getattr(g, grp + "_b0")[i] = next(it).data_ptr()
                getattr(g, grp + "_w2")[i] = next(it).data_ptr()
                getattr(g, grp + "_b2")[i] = next(it).data_ptr()
        device = m1.device
        # data-parallel training: ``model.grad_sync_hook(aug_shape_grads, ready_event)`` (if set) is called as soon as
        # the backward is enqueued; the event fires when the four aug_shape.i gradient sets (99.8 % of the bytes) are
        # final - after ~15 % of the backward - so the hook can start their all-reduce on another stream while the
        # rest of the backward runs (apex DDP does the same for the reference, train.py:154-156)
        hook = getattr(model, "grad_sync_hook", None)
        ready = None
        with torch.cuda.device(device):
            cur = torch.cuda.current_stream(device)
            if hook is not None:
                ready = torch.cuda.Event()
                ready.record(cur)            # creates the underlying cudaEvent_t; the library records it again
            rc = lib.shasta_backward_overlap_f32(
                ctypes.byref(model._cparams), ctypes.byref(g), model._packed.data_ptr(), B, ws.buf.data_ptr(), ws.nbytes,
                m1.data_ptr(), m2.data_ptr(), gm1.data_ptr(), gm2.data_ptr(),
                ctypes.c_void_p(ready.cuda_event) if ready is not None else None, ctypes.c_void_p(cur.cuda_stream))
        _cabi.check(rc, "shasta_backward_overlap_f32")
        if hook is not None:
            hook(grads[n_small:n_small + 16], ready)     # aug_shape.{0..3}.{0,2}.{weight,bias}
        d_bev = d_prev = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            # the maps came out of shared_conv under autograd: hand their gradients back (the reference trains it)
            Bm, H, W, C = ctx.map_shape
            det_c, prev_c = ctx.boxes
            if ctx.needs_input_grad[1]:
                d_bev = torch.zeros((Bm, H, W, C), dtype=torch.float32, device=device)
            if ctx.needs_input_grad[2]:
                d_prev = torch.zeros((Bm, H, W, C), dtype=torch.float32, device=device)
            nbytes = lib.shasta_backward_maps_scratch_bytes(B, model.max_obj)
            scratch = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
            geom = model.bev_extractor.geom(H, W)
            with torch.cuda.device(device):
                rc = lib.shasta_backward_maps_f32(
                    ctypes.byref(model._cparams), B, ctypes.byref(geom), ws.buf.data_ptr(), ws.nbytes, det_c.data_ptr(),
                    prev_c.data_ptr(), scratch.data_ptr(), nbytes,
                    d_bev.data_ptr() if d_bev is not None else None, d_prev.data_ptr() if d_prev is not None else None,
                    ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
            _cabi.check(rc, "shasta_backward_maps_f32")
        return (None, d_bev, d_prev, None, None) + tuple(grads)


def affinity_with_grad(model, bev, prev_bev, det_c, prev_c):
    return _AffinityFunction.apply(model, bev, prev_bev, det_c, prev_c, *differentiable_parameters(model))


class StreamAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` semantics (train.py:146: lr, weight_decay as L2 on the gradient, betas (0.9, 0.999),
    eps 1e-8, no amsgrad) for LARGE fp32 CUDA tensors, one ``shasta_adam_step_f32`` launch per tensor: the four
    ``aug_shape.i.0.weight`` matrices are 99.8 % of the head's parameters and their update is a pure 28-bytes-per-parameter
    stream. Same constructor keywords and ``state_dict`` layout (``step``, ``exp_avg``, ``exp_avg_sq``) as torch's."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _cabi.lib()
        stream = None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise ValueError("StreamAdam: contiguous fp32 CUDA parameters only")
                if stream is None:
                    stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                _cabi.check(lib.shasta_adam_step_f32(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                                     st["exp_avg_sq"].data_ptr(), p.numel(), group["lr"], b1, b2,
                                                     group["eps"], group["weight_decay"], st["step"], stream),
                            "shasta_adam_step_f32")
        return loss
