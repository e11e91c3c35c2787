"""ctypes binding of libshasta_b200.so (the C ABI declared in include/shasta_b200.h).

There is no CPU or eager-PyTorch fallback: if the library is missing the import of the compute path fails
loudly (``ShastaLibraryError``), and every non-zero return code raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libshasta_b200.so")

F32P = ctypes.POINTER(ctypes.c_float)


class ShastaLibraryError(RuntimeError):
    pass


class ShastaParams(ctypes.Structure):
    """Mirror of shasta_params_t."""
    _fields_ = [
        ("max_obj", ctypes.c_int32),
        ("num_feats", ctypes.c_int32),
        ("aug_shape_w0", ctypes.c_void_p * 4),
        ("aug_shape_b0", ctypes.c_void_p * 4),
        ("aug_shape_w2", ctypes.c_void_p * 4),
        ("aug_shape_b2", ctypes.c_void_p * 4),
        ("aug_dets_w0", ctypes.c_void_p * 4),
        ("aug_dets_b0", ctypes.c_void_p * 4),
        ("aug_dets_w2", ctypes.c_void_p * 4),
        ("aug_dets_b2", ctypes.c_void_p * 4),
        ("fuse_shape_w", ctypes.c_void_p * 4),
        ("fuse_shape_b", ctypes.c_void_p * 4),
        ("fuse_det_w", ctypes.c_void_p * 3),
        ("fuse_det_b", ctypes.c_void_p * 3),
        ("res_coeff_w", ctypes.c_void_p * 3),
        ("res_coeff_b", ctypes.c_void_p * 3),
        ("aff_w", ctypes.c_void_p * 6),
        ("aff_b", ctypes.c_void_p * 6),
    ]


class ShastaGrads(ctypes.Structure):
    """Mirror of shasta_grads_t (same fields as ShastaParams without the two leading ints)."""
    _fields_ = [f for f in ShastaParams._fields_ if f[0] not in ("max_obj", "num_feats")]


class ShastaGeom(ctypes.Structure):
    """Mirror of shasta_geom_t."""
    _fields_ = [
        ("pc_start_x", ctypes.c_float), ("pc_start_y", ctypes.c_float),
        ("voxel_x", ctypes.c_float), ("voxel_y", ctypes.c_float),
        ("out_stride", ctypes.c_float),
        ("height", ctypes.c_int32), ("width", ctypes.c_int32),
    ]


class ShastaDecodeOut(ctypes.Structure):
    """Mirror of shasta_decode_out_t (fused decode ring of shasta_forward_decode_f32)."""
    _fields_ = [
        ("n_prev", ctypes.c_void_p), ("n_det", ctypes.c_void_p), ("out", ctypes.c_void_p),
        ("slot_stride", ctypes.c_size_t), ("nslots", ctypes.c_int32), ("counter", ctypes.c_void_p),
    ]


# region ids (enum shasta_region)
WS_FEAT_CUR, WS_FEAT_PREV, WS_BOX_CUR, WS_BOX_PREV, WS_HIDDEN_PART, WS_PROJ_PREV, WS_PROJ_CUR, WS_AUX_PREV, \
    WS_AUX_CUR, WS_COLNORM, WS_RESIDUAL, WS_LOGITS, WS_ANCHOR_BOX, WS_PROJ_CUR_T, WS_DPROJ_PREV, WS_DPROJ_CUR, WS_ANCH_H, \
    WS_ANCH_DY, WS_ANCH_DZ, WS_RAW_XY, WS_BOX_BWD, WS_FEATLO_CUR, WS_FEATLO_PREV, WS_COUNTERS, WS_HID, WS_HIDLO, WS_OUT_PART = range(27)
FLAG_TMA_GATHER, FLAG_NARROW_GATHER, FLAG_PROFILE, FLAG_SKIP_GATHER = 0x1, 0x2, 0x100, 0x200

OPT_ANCHOR_PATH = 0
OPT_TC_RAW_HI = 1
OPT_AFF_PATH = 4      # 0 auto, 1 CUDA cores, 2 tcgen05
OPT_PROJECT_PATH = 5  # same values
OPT_HOST_GATHER_CTAS = 6
OPT_PAIR_FFMA2 = 7    # 2 = scalar FFMA in the pairwise epilogue (comparison)
ANCHOR_AUTO, ANCHOR_STREAM, ANCHOR_TC, ANCHOR_TC_GEN1 = 0, 1, 2, 3

# every symbol include/shasta_b200.h declares: name -> (restype, argtypes)
_vp, _i, _sz, _u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint32
SYMBOLS = {
    "shasta_set_option": (_i, [_i, _i]),
    "shasta_get_option": (_i, [_i]),
    "shasta_abi_version": (_i, []),
    "shasta_last_error_string": (ctypes.c_char_p, []),
    "shasta_last_launch_count": (_i, []),
    "shasta_packed_weight_bytes": (_sz, [_i, _i]),
    "shasta_workspace_bytes": (_sz, [_i, _i]),
    "shasta_workspace_offset": (_sz, [_i, _i, _i]),
    "shasta_proj_cur_stride": (_i, [_i]),
    "shasta_row_stride": (_i, [_i]),
    "shasta_hidden_splits": (_i, [_i]),
    "shasta_pack_weights": (_i, [ctypes.POINTER(ShastaParams), _vp, _sz, _vp]),
    "shasta_bilinear_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "shasta_gather_f32": (_i, [_vp, _vp, _i, _i, _i, ctypes.POINTER(ShastaGeom), _vp, _sz, _i, _vp]),
    "shasta_anchors_f32": (_i, [ctypes.POINTER(ShastaParams), _vp, _vp, _i, _vp, _vp]),
    "shasta_project_f32": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "shasta_pairwise_f32": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "shasta_aff_softmax_f32": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "shasta_forward_f32": (_i, [ctypes.POINTER(ShastaParams), _vp, _vp, _vp, _vp, _vp, _i,
                                ctypes.POINTER(ShastaGeom), _vp, _sz, _vp, _vp, _u32, _vp]),
    "shasta_forward_decode_f32": (_i, [ctypes.POINTER(ShastaParams), _vp, _vp, _vp, _vp, _vp, _i,
                                       ctypes.POINTER(ShastaGeom), _vp, _sz, _vp, _vp, _u32,
                                       ctypes.POINTER(ShastaDecodeOut), _vp]),
    "shasta_anchor_bf16_bytes": (_sz, [_i]),
    "shasta_pack_anchor_bf16": (_i, [ctypes.POINTER(ShastaParams), _vp, _sz, _vp]),
    "shasta_forward_bf16": (_i, [ctypes.POINTER(ShastaParams), _vp, _vp, _vp, _vp, _vp, _vp, _i,
                                 ctypes.POINTER(ShastaGeom), _vp, _sz, _vp, _vp, _u32, _vp]),
    "shasta_gather_pair_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, ctypes.POINTER(ShastaGeom), _vp, _sz, _u32, _vp]),
    "shasta_greedy_assign_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "shasta_shared_conv_packed_bytes": (_sz, []),
    "shasta_shared_conv_scratch_bytes": (_sz, [_i, _i, _i]),
    "shasta_shared_conv_pack": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, _vp, _sz, _vp]),
    "shasta_shared_conv_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "shasta_backward_f32": (_i, [ctypes.POINTER(ShastaParams), ctypes.POINTER(ShastaGrads), _vp, _i, _vp, _sz, _vp, _vp,
                                 _vp, _vp, _vp]),
    "shasta_backward_overlap_f32": (_i, [ctypes.POINTER(ShastaParams), ctypes.POINTER(ShastaGrads), _vp, _i, _vp, _sz,
                                         _vp, _vp, _vp, _vp, _vp, _vp]),
    "shasta_backward_maps_scratch_bytes": (_sz, [_i, _i]),
    "shasta_backward_maps_f32": (_i, [ctypes.POINTER(ShastaParams), _i, ctypes.POINTER(ShastaGeom), _vp, _sz, _vp, _vp,
                                      _vp, _sz, _vp, _vp, _vp]),
    "shasta_adam_step_f32": (_i, [_vp, _vp, _vp, _vp, _sz, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_double, ctypes.c_double, _i, _vp]),
    "shasta_profile_begin": (_i, [_i]),
    "shasta_profile_end": (_i, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "shasta_decode_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Loads the shared library once; raises ShastaLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShastaLibraryError(
            "%s not found: build it with `python -m shasta_b200.build` (nvcc, sm_100a). "
            "shasta_b200 has no CPU or PyTorch fallback for the affinity path." % LIB_PATH)
    try:
        handle = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise ShastaLibraryError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            raise ShastaLibraryError("%s does not export %s (stale build?)" % (LIB_PATH, name))
        fn.restype = res
        fn.argtypes = args
    if handle.shasta_abi_version() != 1:
        raise ShastaLibraryError("ABI version mismatch: library %d, binding 1" % handle.shasta_abi_version())
    _lib = handle
    return _lib


def lib_has(symbol):
    """True when the loaded library exports ``symbol`` (entry points added after ABI version 1 are optional)."""
    try:
        getattr(lib(), symbol)
        return True
    except AttributeError:
        return False


def check(rc, what):
    if rc != 0:
        msg = lib().shasta_last_error_string().decode("utf-8", "replace")
        raise ShastaLibraryError("%s failed with code %d: %s" % (what, rc, msg))
