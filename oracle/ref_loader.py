"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference head from /root/reference.

Only usable where the reference tree exists (the authoring container); it never travels to the
GPU box. It is used by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent) to pin the restatement in
``oracle/shasta_oracle.py`` against the real code.

The ``det3d`` package cannot be imported as a package here (``spconv``, ``terminaltables``,
``pycocotools``, ``addict`` are absent), so the nine files on the hot path are loaded one by one
with stub parent packages (recipe: SURVEY.md §8c).
"""
import importlib.util
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("SHASTA_REF_ROOT", "/root/reference")

_FILES = [
    # (module name, relative path, {attr exports onto other stub modules})
    ("det3d.utils.registry", "det3d/utils/registry.py"),
    ("det3d.models.registry", "det3d/models/registry.py"),
    ("det3d.models.builder", "det3d/models/builder.py"),
    ("det3d.core.utils.circle_nms_jit", "det3d/core/utils/circle_nms_jit.py"),
    ("det3d.core.utils.center_utils", "det3d/core/utils/center_utils.py"),
    ("det3d.core.bbox.box_torch_ops", "det3d/core/bbox/box_torch_ops.py"),
    ("det3d.models.second_stage.bird_eye_view", "det3d/models/second_stage/bird_eye_view.py"),
    ("det3d.models.tracker.base", "det3d/models/tracker/base.py"),
    ("det3d.models.tracker.shasta", "det3d/models/tracker/shasta.py"),
]

_STUB_PKGS = [
    "det3d", "det3d.utils", "det3d.core", "det3d.core.utils", "det3d.core.bbox", "det3d.models",
    "det3d.models.tracker", "det3d.models.second_stage", "det3d.torchie", "det3d.torchie.trainer",
    "det3d.ops", "det3d.ops.iou3d_nms", "pycocotools", "pycocotools.mask",
]

_loaded = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "det3d/models/tracker/shasta.py"))


def _load_file(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    parent, _, leaf = name.rpartition(".")
    spec.loader.exec_module(mod)
    setattr(sys.modules[parent], leaf, mod)
    return mod


def load_reference():
    """Returns a namespace with the reference's Shasta class, registries and builder."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REF_ROOT)
    for name in _STUB_PKGS:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, sys.modules[name])
    sys.modules["det3d.torchie"].is_str = lambda x: isinstance(x, str)
    sys.modules["det3d.torchie.trainer"].load_state_dict = lambda *a, **k: None

    mods = {}
    for name, rel in _FILES:
        mods[name] = _load_file(name, rel)
        if name == "det3d.utils.registry":
            sys.modules["det3d.utils"].Registry = mods[name].Registry
            sys.modules["det3d.utils"].build_from_cfg = mods[name].build_from_cfg
        if name == "det3d.core.bbox.box_torch_ops":
            sys.modules["det3d.core"].box_torch_ops = mods[name]

    reg = mods["det3d.models.registry"]

    class NullReader(nn.Module):
        def forward(self, *a, **k):
            return None

    class NullBackbone(NullReader):
        pass

    class NullNeck(NullReader):
        pass

    reg.READERS.register_module(NullReader)
    reg.BACKBONES.register_module(NullBackbone)
    reg.NECKS.register_module(NullNeck)

    ns = types.SimpleNamespace(
        Shasta=mods["det3d.models.tracker.shasta"].Shasta,
        BEVFeatureExtractor=mods["det3d.models.second_stage.bird_eye_view"].BEVFeatureExtractor,
        builder=mods["det3d.models.builder"],
        registry=reg,
        center_utils=mods["det3d.core.utils.center_utils"],
        box_torch_ops=mods["det3d.core.bbox.box_torch_ops"],
        Registry=mods["det3d.utils.registry"].Registry,
        build_from_cfg=mods["det3d.utils.registry"].build_from_cfg,
    )
    _loaded = ns
    return ns


def build_reference_head(max_obj, num_feats=3, pc_start=(-54, -54), voxel_size=(0.075, 0.075),
                         out_stride=8):
    """Builds the reference ``Shasta`` through its own registry, trunk stubbed, shared_conv = Identity
    (the timed region of the metric starts at the 64-channel NHWC map)."""
    ref = load_reference()
    cfg = dict(
        type="Shasta",
        reader=dict(type="NullReader"), backbone=dict(type="NullBackbone"), neck=dict(type="NullNeck"),
        bev_extractor=dict(type="BEVFeatureExtractor", pc_start=list(pc_start),
                           voxel_size=list(voxel_size), out_stride=out_stride),
        max_obj=max_obj, num_feats=num_feats,
    )
    model = ref.builder.build_track(cfg)
    model.shared_conv = _NCHWIdentity()
    model.eval()
    return model


class _NCHWIdentity(nn.Module):
    """Stands in for shared_conv: the stubbed trunk hands over the 64-channel map as (B,64,H,W) so the
    reference's own ``permute(0,2,3,1).contiguous()`` (shasta.py:224,228) runs unmodified."""

    def forward(self, x):
        return x


def run_reference(model, bev_nhwc, prev_bev_nhwc, det_boxes, prev_det_boxes, capture=False):
    """Runs the unmodified reference forward. ``bev_*`` are (B,H,W,64); boxes are (B,M,11) and
    ``det_boxes`` is mutated in place exactly as the reference does (shasta.py:270).
    With ``capture`` returns a dict of intermediates taken with forward hooks."""
    bev = bev_nhwc.permute(0, 3, 1, 2)
    prev_bev = prev_bev_nhwc.permute(0, 3, 1, 2)
    model.extract_feat = lambda ex: (bev, None, prev_bev, None)
    example = {"det_boxes": det_boxes, "prev_det_boxes": prev_det_boxes}
    inter = {}
    hooks = []
    if capture:
        feats = []

        def saver(key):
            def hook(mod, inp, out):
                inter[key] = out.detach().clone()  # returns None: the output is left untouched
            return hook

        def feat_hook(mod, inp, out):
            feats.append(torch.stack(out).detach().clone())

        def aff_hook(mod, inp, out):
            inter["residual"] = inp[0].detach().clone()
            inter["logits"] = out.detach().clone()

        hooks.append(model.bev_extractor.register_forward_hook(feat_hook))
        hooks.append(model.fuse_shape.register_forward_hook(saver("fuse_shape")))
        hooks.append(model.fuse_det.register_forward_hook(saver("fuse_det")))
        hooks.append(model.res_coeff.register_forward_hook(saver("res_coeff")))
        hooks.append(model.aff.register_forward_hook(aff_hook))
    with torch.no_grad():
        m1, m2, ex = model(example, train_mode=False)
    for h in hooks:
        h.remove()
    if capture:
        inter["feature"], inter["prev_feature"] = feats[0], feats[1]
        # shasta.py:241-244 calls aug_shape[i].forward() directly (hooks do not fire), so the four anchor
        # shape vectors are recomputed here with the reference's own modules on the captured features.
        with torch.no_grad():
            for k in range(4):
                src = feats[0] if k < 2 else feats[1]
                inter["aug_shape%d" % k] = torch.abs(
                    model.aug_shape[k](src.view(src.shape[0], -1))).reshape(src.shape[0], 1, -1)
        for k in ("newborn", "fp", "dead_trk", "fn"):
            inter[k] = getattr(model, k).detach().clone()
        return m1, m2, ex, inter
    return m1, m2, ex
