"""ctypes binding of libhmsg_b200.so (include/hmsg_b200.h).  No CPU fallback: if the
library is missing or no B200 is present, every product call fails loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhmsg_b200.so")

_i32, _i64, _f32, _f64, _vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class VitDesc(C.Structure):
    _fields_ = [(n, _i32) for n in ("image", "patch", "width", "layers", "heads", "mlp", "out_dim", "quick_gelu")]


# name -> (restype, argtypes); mirrors include/hmsg_b200.h one to one
SIGNATURES = {
    "hmsg_version": (_i32, []),
    "hmsg_ctx_create": (_i32, [_i32, C.POINTER(_vp)]),
    "hmsg_ctx_destroy": (_i32, [_vp]),
    "hmsg_last_error": (C.c_char_p, [_vp]),
    "hmsg_sync": (_i32, [_vp]),
    "hmsg_stream": (_vp, [_vp]),
    "hmsg_launch_count": (_i64, [_vp]),
    "hmsg_set_option": (_i32, [_vp, C.c_char_p, _i32]),
    "hmsg_prof_enable": (_i32, [_vp, C.c_uint32]),
    "hmsg_prof_read": (_i32, [_vp, _i32, C.POINTER(_f64), C.POINTER(_i64), C.POINTER(_f64)]),
    "hmsg_scene_begin": (_i32, [_vp, _i32, _i32, _vp, _f32, _f64, _i64]),
    "hmsg_scene_add_frames": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32]),
    "hmsg_scene_put_frames": (_i32, [_vp, _i64, _vp, _vp, _vp, _i32, _i32]),
    "hmsg_scene_set_num_frames": (_i32, [_vp, _i64]),
    "hmsg_scene_set_intrinsics": (_i32, [_vp, _i64, _i32, _vp]),
    "hmsg_scene_put_rgb": (_i32, [_vp, _i64, _vp, _i32, _i32]),
    "hmsg_scene_num_frames": (_i64, [_vp]),
    "hmsg_unproject_frame": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "hmsg_voxel_build": (_i32, [_vp, C.POINTER(_i64), _vp]),
    "hmsg_voxel_bounds": (_i32, [_vp, _i64, _i64, _vp]),
    "hmsg_voxel_grid_set": (_i32, [_vp, _vp]),
    "hmsg_voxel_mark": (_i32, [_vp, _i64, _i64]),
    "hmsg_voxel_bitmap": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "hmsg_voxel_bitmap_or": (_i32, [_vp, _vp, _i32]),
    "hmsg_voxel_scan": (_i32, [_vp, C.POINTER(_i64)]),
    "hmsg_voxel_accumulate": (_i32, [_vp, _i64, _i64]),
    "hmsg_voxel_acc": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "hmsg_voxel_finalize": (_i32, [_vp]),
    "hmsg_voxels_read": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "hmsg_radius_filter": (_i32, [_vp, _i32, _f64, C.POINTER(_i64)]),
    "hmsg_radius_counts_read": (_i32, [_vp, _vp]),
    "hmsg_nodes_read": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "hmsg_num_nodes": (_i64, [_vp]),
    "hmsg_pixel_to_node": (_i32, [_vp, _i64, _vp, _vp]),
    "hmsg_points_to_node": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "hmsg_features_begin": (_i32, [_vp, _i32]),
    "hmsg_masks_dense": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32]),
    "hmsg_masks_boxes": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32]),
    "hmsg_fuse_scatter": (_i32, [_vp, _i64, _i32, _i32, _vp, _f32, _vp, _i32]),
    "hmsg_node_feats_finalize": (_i32, [_vp, _vp, _i32]),
    "hmsg_node_feats_raw": (_i32, [_vp, _vp, _vp]),
    "hmsg_mask_nodes": (_i32, [_vp, _i64, _f64, _vp, _vp, _vp, _vp]),
    "hmsg_encoder_load": (_i32, [_vp, C.POINTER(VitDesc), _vp, _i64]),
    "hmsg_encode_images": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32]),
    "hmsg_gemm_f16_debug": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32]),
    "hmsg_make_crops": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, C.POINTER(_vp)]),
    "hmsg_encode_crops": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp]),
    "hmsg_crops_read": (_i32, [_vp, _i64, _vp]),
    "hmsg_debug_pil_mma_table": (_i32, [_vp, _vp]),
    "hmsg_scene_reset_frames": (_i32, [_vp]),
    "hmsg_node_feats_pack": (_i32, [_vp, _vp, _vp, _i64]),
    "hmsg_node_feats_merge": (_i32, [_vp, _vp, _i32, _i64]),
    "hmsg_query_scores": (_i32, [_vp, _vp, _i32, _vp, _i32]),
    "hmsg_pixel_feature_map": (_i32, [_vp, _i64, _vp]),
    "hmsg_index_set": (_i32, [_vp, _vp, _i64, _i32, _i32]),
    "hmsg_query_topk": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32]),
    "hmsg_query_object": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32]),
    "hmsg_objects_begin": (_i32, [_vp, C.c_double, C.c_double, C.c_double]),
    "hmsg_objects_add_masks": (_i32, [_vp, _i32, _vp, _vp, _vp, _i32]),
    "hmsg_objects_add_frame": (_i32, [_vp, _i64, C.c_double, C.c_double]),
    "hmsg_objects_finish": (_i32, [_vp, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "hmsg_objects_read": (_i32, [_vp, _vp, _vp, _vp]),
    "hmsg_objects_count": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "hmsg_object_feats": (_i32, [_vp, _vp, _i32, C.c_double, C.c_double, C.c_float, _i32, _vp, _i32]),
    "hmsg_node_feats_device": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i32)]),
    "hmsg_masks_counts": (_i32, [_vp, _i64, _i32, _vp]),
    "hmsg_masks_labels": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32]),
    "hmsg_mask_nodes_batch": (_i32, [_vp, _i64, _i32, C.c_double, C.c_double, _i32]),
    "hmsg_mask_store_reset": (_i32, [_vp]),
    "hmsg_mask_store_count": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "hmsg_mask_store_read": (_i32, [_vp, _i64, C.POINTER(_i32), _vp, _vp, _vp, _vp]),
    "hmsg_objects_merge_stored": (_i32, [_vp, _i64, _i64]),
    "hmsg_comm_unique_id": (_i32, [_vp]),
    "hmsg_comm_init": (_i32, [_vp, _vp, _i32, _i32]),
    "hmsg_comm_attach": (_i32, [_vp, _vp]),
    "hmsg_comm_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_f64)]),
    "hmsg_voxel_build_sharded": (_i32, [_vp, _vp, _vp, _i32, C.POINTER(_i64), _vp]),
    "hmsg_radius_filter_sharded": (_i32, [_vp, _vp, _i32, _f64, C.POINTER(_i64)]),
    "hmsg_allgather_nodes": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64]),
}

_lib = None


def load(path: str | None = None):
    """dlopen the library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -m holoagent_b200.build` (nvcc, sm_100a). "
            "holoagent_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def ptr(a):
    """void* of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
