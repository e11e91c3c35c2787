"""Drop-in at the registry level against the LIVE reference: the CUDA head is
registered under the name ``Shasta`` in the REFERENCE's own ``TRACK`` registry and built through the reference's own
``det3d/models/builder.py:70-75`` (``build_simp_track``), exactly as the maintainer's one-file change would do it.
CPU only (construction, registry semantics, checkpoint compatibility); skipped where /root/reference is absent."""
import pytest

from oracle import ref_loader
import shasta_b200

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture()
def swapped_registry():
    """The reference registries with its own Shasta / BEVFeatureExtractor taken out and ours registered instead."""
    ref = ref_loader.load_reference()
    track, second = ref.registry.TRACK, ref.registry.SECOND_STAGE
    saved = (track._module_dict.pop("Shasta"), second._module_dict.pop("BEVFeatureExtractor"))

    class Shasta(shasta_b200.Shasta):                      # INTEGRATION.md §1, verbatim
        def __init__(self, reader, backbone, neck, bev_extractor, **kw):
            builder = ref.builder                          # det3d.models.builder of the reference tree
            super().__init__(None, None, None, bev_extractor, **kw)
            self.reader = builder.build_reader(reader)
            self.backbone = builder.build_backbone(backbone)
            self.neck = builder.build_neck(neck)

    track.register_module(Shasta)
    second.register_module(shasta_b200.BEVFeatureExtractor)
    try:
        yield ref, Shasta
    finally:
        track._module_dict.pop("Shasta", None)
        second._module_dict.pop("BEVFeatureExtractor", None)
        track._module_dict["Shasta"], second._module_dict["BEVFeatureExtractor"] = saved


def _cfg(M):
    return dict(type="Shasta", reader=dict(type="NullReader"), backbone=dict(type="NullBackbone"),
                neck=dict(type="NullNeck"),
                bev_extractor=dict(type="BEVFeatureExtractor", pc_start=[-54, -54], voxel_size=[0.075, 0.075],
                                   out_stride=8),
                max_obj=M, num_feats=3)


def test_build_through_the_reference_builder(swapped_registry):
    ref, Shasta = swapped_registry
    model = ref.builder.build_simp_track(_cfg(20), train_cfg=None, test_cfg=dict(score=0.1))
    assert isinstance(model, shasta_b200.Shasta) and type(model) is Shasta
    assert model.test_cfg == dict(score=0.1) and model.train_cfg is None     # builder.py:70-75 default_args
    # the trunk slots were built by the REFERENCE's registries, the sampler is ours
    assert type(model.reader).__name__ == "NullReader" and type(model.neck).__name__ == "NullNeck"
    assert isinstance(model.bev_extractor, shasta_b200.BEVFeatureExtractor)
    # registry semantics of det3d/utils/registry.py:28-46
    with pytest.raises(KeyError):
        ref.registry.TRACK.register_module(Shasta)
    with pytest.raises(KeyError):
        ref.builder.build_simp_track(dict(_cfg(20), type="NoSuchHead"))


def test_checkpoint_of_the_reference_loads_by_name_and_shape(swapped_registry):
    """det3d/torchie/trainer/checkpoint.py:67-107 copies tensors whose name and shape match: every parameter and buffer
    of the reference head (incl. shared_conv and its BatchNorm buffers) must find its twin."""
    ref, Shasta = swapped_registry
    M = 20
    ours = ref.builder.build_simp_track(_cfg(M))
    sd_ref = {k: tuple(v.shape) for k, v in ref_state_dict(ref, M).items()}
    sd_ours = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert sd_ours == sd_ref
    skipped = shasta_b200.load_matching_state_dict(ours, ref_state_dict(ref, M))
    assert skipped == []


def ref_state_dict(ref, M):
    """state_dict of the unmodified reference head with its real shared_conv (the class is instantiated directly: its
    registry entry is swapped out while the fixture is active, the sub-modules resolve through the registries)."""
    second = ref.registry.SECOND_STAGE
    ours = second._module_dict.pop("BEVFeatureExtractor")
    second._module_dict["BEVFeatureExtractor"] = ref.BEVFeatureExtractor
    try:
        model = ref.Shasta(reader=dict(type="NullReader"), backbone=dict(type="NullBackbone"),
                           neck=dict(type="NullNeck"),
                           bev_extractor=dict(type="BEVFeatureExtractor", pc_start=[-54, -54],
                                              voxel_size=[0.075, 0.075], out_stride=8),
                           max_obj=M, num_feats=3)
    finally:
        second._module_dict["BEVFeatureExtractor"] = ours
    return model.state_dict()
