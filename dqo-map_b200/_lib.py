"""ctypes binding of libdqomap_b200.so (include/dqo_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  `lib()` raises if the shared object is
missing or its ABI version does not match this file.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DQO_B200_LIB: another build of the same library (kernel A/B experiments, tests/dev_*.py); there is still no fallback
LIB_PATH = os.environ.get("DQO_B200_LIB") or os.path.join(_HERE, "csrc", "libdqomap_b200.so")
ABI_VERSION = 14

ST_NUM_RENDERED, ST_TILE_NUM, ST_OVERFLOW, ST_NUM_VISIBLE, ST_R_FRONT, ST_R_BACK, ST_WALKED, ST_UNFINISHED = range(8)
ST_WORDS = 8
ADAM_MAX_TENSORS = 16
STEP_TERM_SSIM, STEP_TERM_SEMANTIC = 1, 2

c_p = C.c_void_p


class RastSettings(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("D", C.c_int32), ("M", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("scale_modifier", C.c_float), ("color_sigma", C.c_float), ("opaque_threshold", C.c_float),
        ("depth_threshold", C.c_float), ("normal_threshold", C.c_float), ("T_threshold", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32), ("need_n_touched", C.c_int32),
        ("front_instances", C.c_int32), ("back_instances", C.c_int32), ("geom_clean", C.c_int32),
    ]


class AdamTensor(C.Structure):
    _fields_ = [
        ("param", c_p), ("grad", c_p), ("exp_avg", c_p), ("exp_avg_sq", c_p),
        ("numel", C.c_int64), ("lr", C.c_double), ("row_width", C.c_int32), ("reserved", C.c_int32),
    ]


class MapParams(C.Structure):
    _fields_ = [("param", c_p * 6), ("exp_avg", c_p * 6), ("exp_avg_sq", c_p * 6), ("lr", C.c_double * 6),
                ("confidence", c_p), ("ever", c_p), ("ever_list", c_p), ("ever_count", c_p),
                ("init_xyz", c_p), ("init_scaling", c_p), ("init_rotation", c_p), ("init_opacity", c_p),
                ("attach_count", c_p), ("attach_weight", C.c_float), ("attach_opacity_thres", C.c_float),
                ("step_state", c_p),
                ("semantics", c_p), ("semantics_exp_avg", c_p), ("semantics_exp_avg_sq", c_p),
                ("lr_semantics", C.c_double), ("workspace_terms", C.c_int32), ("reserved", C.c_int32)]


class Keyframe(C.Structure):
    _fields_ = [("gt_color", c_p), ("gt_depth", c_p), ("render_mask", c_p), ("tile_mask", c_p), ("viewmatrix", c_p),
                ("projmatrix", c_p), ("campos", c_p), ("background", c_p), ("color_weight", C.c_float),
                ("depth_weight", C.c_float), ("depth_err_thres", C.c_float), ("ssim_weight", C.c_float),
                ("semantic_weight", C.c_float), ("gt_semantic", c_p)]


# name -> (restype, argtypes); mirrors include/dqo_b200.h declaration by declaration
PROTOTYPES = {
    "dqo_abi_version": (C.c_int, []),
    "dqo_last_error": (C.c_char_p, []),
    "dqo_launch_count": (C.c_longlong, []),
    "dqo_profile_enable": (None, [C.c_int]),
    "dqo_profile_read": (C.c_int, [C.POINTER(C.c_float), C.c_int]),
    "dqo_rast_geom_bytes": (C.c_size_t, [C.c_int32]),
    "dqo_rast_geom_init": (C.c_int, [C.c_int32, c_p, c_p]),
    "dqo_rast_binning_bytes": (C.c_size_t, [C.c_int64]),
    "dqo_rast_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dqo_rast_forward": (C.c_int, [C.POINTER(RastSettings)] + [c_p] * 12 + [c_p, c_p, C.c_int64, c_p] + [c_p] * 12),
    "dqo_rast_backward": (C.c_int, [C.POINTER(RastSettings)] + [c_p] * 11 + [c_p, c_p, C.c_int64, c_p, c_p]
                          + [c_p] * 3 + [c_p] * 9 + [c_p]),
    "dqo_mark_visible": (C.c_int, [C.c_int32, c_p, c_p, c_p, c_p, c_p]),
    "dqo_rast_export_state": (C.c_int, [C.POINTER(RastSettings), c_p, c_p, C.c_int64, c_p, c_p] + [c_p] * 10 + [c_p]),
    "dqo_rast_blend_extra": (C.c_int, [C.POINTER(RastSettings), c_p, c_p, c_p, c_p, C.c_int64, c_p, c_p, c_p, c_p]),
    "dqo_rast_blend_extra_backward": (C.c_int, [C.POINTER(RastSettings)] + [c_p] * 10 + [c_p, c_p, C.c_int64, c_p, c_p]
                                      + [c_p, c_p] + [c_p] * 8 + [c_p]),
    "dqo_sort_pairs_temp_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "dqo_sort_pairs_u32": (C.c_int, [c_p, c_p, c_p, c_p, C.c_int32, c_p, c_p, C.c_int64, C.c_int32, c_p, c_p]),
    "dqo_sort_pairs_u16": (C.c_int, [c_p, c_p, c_p, c_p, C.c_int32, c_p, c_p, C.c_int64, C.c_int32, c_p, c_p]),
    "dqo_knn_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "dqo_knn3": (C.c_int, [C.c_int32, c_p, c_p, c_p, c_p, C.c_size_t, c_p]),
    "dqo_bbox_mask": (C.c_int, [C.c_int32, c_p, C.c_int32, c_p, C.c_float, c_p, c_p, c_p, c_p]),
    "dqo_gaussian_radius": (C.c_int, [C.c_int32, c_p, c_p, c_p]),
    "dqo_scale_init": (C.c_int, [C.c_int32, C.c_int32, c_p, c_p, c_p] + [C.c_float] * 6 + [c_p, c_p, c_p, c_p]),
    "dqo_knn_cross3_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dqo_knn_cross3": (C.c_int, [C.c_int32, c_p, C.c_int32, c_p, c_p, c_p, c_p, C.c_size_t, c_p]),
    "dqo_inside_mask": (C.c_int, [C.c_int32, c_p, c_p, c_p, C.c_float, c_p, c_p]),
    "dqo_accumulate_error": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [c_p] * 5
                             + [C.c_float, C.c_float, C.c_float, C.c_int32] + [c_p] * 7 + [c_p]),
    "dqo_render_range": (C.c_int, [C.c_int32, C.c_int32, c_p, C.c_float, c_p, c_p, c_p, c_p]),
    "dqo_pixelmask_to_tilemask": (C.c_int, [C.c_int32, C.c_int32, c_p, C.c_float, c_p, c_p, c_p]),
    "dqo_color_error_map": (C.c_int, [C.c_int32, C.c_int32, c_p, c_p, c_p, c_p]),
    "dqo_topk_tilemask_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dqo_topk_tilemask": (C.c_int, [C.c_int32, C.c_int32, c_p, C.c_int32, C.c_int32, c_p, c_p, c_p, c_p]),
    "dqo_tilemask_to_pixelmask": (C.c_int, [C.c_int32, C.c_int32, c_p, c_p, c_p]),
    "dqo_render_error_maps": (C.c_int, [C.c_int32, C.c_int32] + [c_p] * 8 + [c_p]),
    "dqo_depth_maxpool": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_p, c_p, c_p]),
    "dqo_vertex_normal_workspace_bytes": (C.c_size_t, []),
    "dqo_vertex_normal_map": (C.c_int, [C.c_int32, C.c_int32, c_p] + [C.c_float] * 4 + [c_p, c_p, c_p, c_p]),
    "dqo_normal_map": (C.c_int, [C.c_int32, C.c_int32, c_p, c_p, c_p, c_p]),
    "dqo_icp_workspace_bytes": (C.c_size_t, []),
    "dqo_icp_level": (C.c_int, [C.c_int32, C.c_int32, c_p, c_p, c_p, c_p] + [C.c_float] * 7 + [C.c_int32, c_p, c_p, c_p, c_p]),
    "dqo_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dqo_masked_l1_loss": (C.c_int, [C.c_int32, C.c_int32] + [c_p] * 6 + [C.c_float, C.c_float, C.c_float]
                           + [c_p] * 5 + [c_p]),
    "dqo_ssim_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dqo_ssim_window": (None, [C.POINTER(C.c_float)]),
    "dqo_ssim_loss": (C.c_int, [C.c_int32, C.c_int32, c_p, c_p, C.c_float, c_p, C.c_int32, c_p, c_p, c_p]),
    "dqo_adam_step": (C.c_int, [C.POINTER(AdamTensor), C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                c_p, C.c_int32, c_p]),
    "dqo_mapping_step_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "dqo_mapping_step_workspace_init": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, c_p, c_p]),
    "dqo_mapping_step": (C.c_int, [C.POINTER(RastSettings), C.POINTER(MapParams), C.POINTER(Keyframe), C.c_int32,
                                   C.c_double, C.c_double, C.c_double, c_p, C.c_int64, c_p, c_p, c_p, c_p]),
    "dqo_attach_count": (C.c_int, [C.c_int32, c_p, C.c_float, c_p, c_p]),
    "dqo_mapping_step_outputs": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, c_p]
                                 + [C.POINTER(c_p)] * 4),
    "dqo_quadric_init": (C.c_int, [C.c_int32] + [c_p] * 7 + [c_p]),
    "dqo_quadric_project": (C.c_int, [C.c_int32] + [c_p] * 6 + [c_p]),
    "dqo_quadric_refine": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [c_p] * 4
                           + [C.c_float, C.c_float, C.c_float] + [c_p] * 4 + [c_p]),
}

_lib = None


class DqoError(RuntimeError):
    pass


def lib():
    """Load (once) and return the C-ABI library; fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DqoError(
            "libdqomap_b200.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if handle.dqo_abi_version() != ABI_VERSION:
        raise DqoError("libdqomap_b200.so ABI %d != binding ABI %d; rebuild" % (handle.dqo_abi_version(), ABI_VERSION))
    _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().dqo_last_error().decode("utf-8", "replace")
        raise DqoError("%s failed (code %d): %s" % (what, code, msg))


def ptr(t):
    """Device/host pointer of a tensor as an int (None -> NULL; empty tensors -> NULL like the reference's
    null data_ptr convention)."""
    if t is None:
        return None
    if t.numel() == 0:
        return None
    return t.data_ptr()
