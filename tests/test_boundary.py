"""CPU: the drop-in boundary — registry/builder semantics (det3d/utils/registry.py, det3d/models/builder.py),
state_dict contract (shasta.py:42-106), and that the C-ABI library loads and exports every symbol the header
declares. No compute call is made (there is no GPU on the CPU test box)."""
import os
import re

import pytest
import torch
from torch import nn

import shasta_b200
from shasta_b200 import _cabi, registry, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_semantics():
    R = registry.Registry("thing")

    @R.register_module
    class A(nn.Module):
        def __init__(self, x=1, train_cfg=None):
            super().__init__()
            self.x, self.train_cfg = x, train_cfg

    assert R.get("A") is A and R.get("B") is None
    with pytest.raises(KeyError):
        R.register_module(A)                       # duplicate name
    with pytest.raises(TypeError):
        R.register_module(lambda: None)            # not a class
    obj = registry.build_from_cfg(dict(type="A", x=3), R, dict(train_cfg="t", x=9))
    assert obj.x == 3 and obj.train_cfg == "t"      # cfg wins over default_args
    assert isinstance(registry.build_from_cfg(dict(type=A), R), A)   # class instead of name
    with pytest.raises(KeyError, match="is not in the thing registry"):
        registry.build_from_cfg(dict(type="Nope"), R)
    with pytest.raises(TypeError):
        registry.build_from_cfg(dict(type=3), R)
    seq = registry.build([dict(type="A"), dict(type="A", x=2)], R)
    assert isinstance(seq, nn.Sequential) and len(seq) == 2


def _cfg(M=20):
    return dict(type="Shasta", reader=None, backbone=None, neck=None,
                bev_extractor=dict(type="BEVFeatureExtractor", pc_start=[-54, -54], voxel_size=[0.075, 0.075],
                                   out_stride=8),
                max_obj=M, num_feats=3)


def test_build_track_and_state_dict_contract():
    model = shasta_b200.build_simp_track(_cfg(20), train_cfg="tr", test_cfg="te")
    assert type(model).__name__ == "Shasta" and registry.TRACK.get("Shasta") is type(model)
    assert registry.SECOND_STAGE.get("BEVFeatureExtractor") is type(model.bev_extractor)
    assert model.train_cfg == "tr" and model.test_cfg == "te" and model.max_obj == 20
    sd = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    want = {k: tuple(v) for k, v in synthetic.head_param_shapes(20).items()}
    head = {k: v for k, v in sd.items() if not k.startswith("shared_conv")}
    assert head == want
    # shared_conv keeps the reference's names too (shasta.py:42-47)
    assert sd["shared_conv.0.weight"] == (64, 512, 3, 3) and "shared_conv.1.running_mean" in sd
    # load-by-name-and-shape semantics (checkpoint.py:67-107): mismatches are skipped, not fatal
    ck = {"aff.0.weight": torch.ones(128, 22), "aff.2.weight": torch.ones(3, 3), "module.aff.0.bias": torch.ones(128),
          "unknown": torch.ones(1)}
    skipped = shasta_b200.load_matching_state_dict(model, ck)
    assert sorted(skipped) == ["aff.2.weight", "unknown"]
    assert torch.all(model.aff[0].weight == 1) and torch.all(model.aff[0].bias == 1)


def test_unsupported_variants_raise_clearly():
    for kw in (dict(num_feats=7), dict(num_point=4), dict(share_conv_channel=32)):
        cfg = _cfg(20)
        cfg.update(kw)
        with pytest.raises(NotImplementedError):
            shasta_b200.build_track(cfg)


def test_no_cpu_fallback():
    model = shasta_b200.build_track(_cfg(6))
    z = torch.zeros
    with pytest.raises(_cabi.ShastaLibraryError, match="no CPU path"):
        model({"det_boxes": z(1, 6, 11), "prev_det_boxes": z(1, 6, 11), "bev_feature": z(1, 8, 8, 64),
               "prev_bev_feature": z(1, 8, 8, 64)}, train_mode=False)
    with pytest.raises(RuntimeError, match="no trunk attached"):
        model({"det_boxes": z(1, 6, 11), "prev_det_boxes": z(1, 6, 11)}, train_mode=False)


def test_cabi_exports_every_declared_symbol(shasta_lib):
    header = open(os.path.join(ROOT, "include", "shasta_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(shasta_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(shasta_lib, name), name


def test_cabi_host_side_queries(shasta_lib):
    lib = shasta_lib
    assert lib.shasta_abi_version() == 1
    assert lib.shasta_packed_weight_bytes(200, 3) > 0 and lib.shasta_packed_weight_bytes(200, 7) == 0
    assert lib.shasta_workspace_bytes(0, 200) >= 0
    n = lib.shasta_workspace_bytes(4, 200)
    offs = [lib.shasta_workspace_offset(4, 200, r) for r in range(21)]
    assert offs == sorted(offs) and offs[0] == 0 and offs[-1] * 4 < n
    assert all(o % 64 == 0 for o in offs)                      # 256-byte aligned regions
    assert lib.shasta_workspace_offset(4, 200, 99) == 2 ** 64 - 1
    assert lib.shasta_proj_cur_stride(200) == 256 and lib.shasta_row_stride(200) == 204
    assert lib.shasta_hidden_splits(200) == 32
    # argument errors are reported, not crashed on (no launch happens before validation)
    assert lib.shasta_bilinear_f32(None, 4, 4, 8, None, None, 0, None, None) < 0
    assert b"NULL" in lib.shasta_last_error_string()


def test_adam_entry_point_validates_on_the_host(shasta_lib):
    """shasta_adam_step_f32 (train.py:146 optim.Adam): argument errors come back as codes before any launch, an empty
    tensor is a no-op, and the torch-side wrapper refuses tensors the kernel cannot update."""
    lib = shasta_lib
    assert lib.shasta_adam_step_f32(None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None) == 0
    assert lib.shasta_adam_step_f32(None, None, None, None, 4, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None) < 0
    assert b"NULL" in lib.shasta_last_error_string()
    buf = torch.zeros(16)
    p = buf.data_ptr()
    assert lib.shasta_adam_step_f32(p, p, p, p, 4, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, None) < 0      # step is 1-based
    assert lib.shasta_adam_step_f32(p, p, p, p, 4, 1e-3, 1.0, 0.999, 1e-8, 0.0, 1, None) < 0      # beta1 < 1
    assert b"adam" in lib.shasta_last_error_string()
    from shasta_b200 import training
    with pytest.raises(ValueError):
        training.StreamAdam([torch.nn.Parameter(torch.zeros(4))], lr=-1.0)
    w = torch.nn.Parameter(torch.zeros(4))
    w.grad = torch.ones(4)
    with pytest.raises(ValueError, match="CUDA"):
        training.StreamAdam([w]).step()
