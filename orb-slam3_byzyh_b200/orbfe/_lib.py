"""ctypes binding of libORBfe_b200.so (the C ABI declared in include/orbfe.h).

The library is the product: if it is missing, or no CUDA device is usable, every entry point
fails loudly -- there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libORBfe_b200.so")

OK, EMPTY_IMAGE, ERR_INVALID, ERR_CUDA, ERR_CAPACITY = 0, -1, -2, -3, -4
NUM_STAGES = 10
STAGE_NAMES = ("h2d", "pyramid", "fast", "octree", "layout", "blur", "describe", "d2h")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])  # == cv::KeyPoint
assert KP_DTYPE.itemsize == 28


class OrbfeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"orbfe error {code}: {msg}")
        self.code = code


class FrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys", C.c_void_p), ("uright", C.c_void_p),
                ("desc", C.c_void_p), ("min_x", C.c_float), ("min_y", C.c_float),
                ("max_x", C.c_float), ("max_y", C.c_float), ("grid_w_inv", C.c_float),
                ("grid_h_inv", C.c_float)]


class ProjPoints(C.Structure):
    _fields_ = [("m", C.c_int32), ("u", C.c_void_p), ("v", C.c_void_p), ("ur", C.c_void_p),
                ("radius", C.c_void_p), ("min_level", C.c_void_p), ("max_level", C.c_void_p),
                ("angle", C.c_void_p), ("valid", C.c_void_p), ("blocks", C.c_void_p),
                ("desc", C.c_void_p)]


class SearchParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("th_accept", C.c_int32), ("nnratio", C.c_float),
                ("check_orientation", C.c_int32)]


class WindowParams(C.Structure):
    _fields_ = [("th_accept", C.c_int32), ("gate", C.c_int32), ("inv_level_sigma2", C.c_void_p),
                ("n_levels", C.c_int32)]


class FeatureVectorC(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_id", C.c_void_p), ("start", C.c_void_p), ("feat", C.c_void_p)]


class BowSide(C.Structure):
    _fields_ = [("n", C.c_int32), ("desc", C.c_void_p), ("angle", C.c_void_p), ("valid", C.c_void_p),
                ("fv", FeatureVectorC)]


class TriSide(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p),
                ("has_map_point", C.c_void_p), ("fv", FeatureVectorC)]


class TriCameraPair(C.Structure):
    _fields_ = [("params1", C.c_float * 8), ("params2", C.c_float * 8), ("precision1", C.c_float),
                ("precision2", C.c_float), ("R12", C.c_float * 9), ("t12", C.c_float * 3)]


class TriRig(C.Structure):
    _fields_ = [("n_left1", C.c_int32), ("n_left2", C.c_int32), ("level_sigma2_1", C.c_void_p), ("pair", TriCameraPair * 4)]


class TriParams(C.Structure):
    _fields_ = [("f12", C.c_float * 9), ("epipole", C.c_float * 2), ("scale_factors2", C.c_void_p),
                ("level_sigma2_2", C.c_void_p), ("n_levels", C.c_int32), ("only_stereo", C.c_int32),
                ("coarse", C.c_int32), ("check_orientation", C.c_int32), ("th_low", C.c_int32),
                ("rig", C.POINTER(TriRig))]


_lib = None

_vp, _i, _f, _sz, _ull = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_ulonglong
_SIGS = {
    "orbfe_extractor_create": (_i, [_i, _f, _i, _i, _i, _i, C.POINTER(_vp)]),
    "orbfe_extractor_destroy": (None, [_vp]),
    "orbfe_last_error": (C.c_char_p, []),
    "orbfe_version": (C.c_char_p, []),
    "orbfe_get_levels": (_i, [_vp]),
    "orbfe_get_scale_factor": (_f, [_vp]),
    "orbfe_scale_tables": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "orbfe_features_per_level": (_i, [_vp, _vp]),
    "orbfe_max_keypoints": (_i, [_vp]),
    "orbfe_max_keypoints_for": (_i, [_vp, _i, _i]),
    "orbfe_extract": (_i, [_vp, _vp, _i, _i, _sz, _i, _i, _vp, _vp, _i, C.POINTER(_i)]),
    "orbfe_extract_batch": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "orbfe_extract_batch_submit": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "orbfe_extract_batch_wait": (_i, [_vp]),
    "orbfe_extract_batch_device": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "orbfe_level_size": (_i, [_vp, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "orbfe_pyramid_level": (_i, [_vp, _i, _i, _i, _vp, _sz]),
    "orbfe_debug_candidates": (_i, [_vp, _i, _i, _vp, _i, C.POINTER(_i)]),
    "orbfe_debug_level_keypoints": (_i, [_vp, _i, _i, _vp, _i, C.POINTER(_i)]),
    "orbfe_debug_blurred": (_i, [_vp, _i, _i, _vp, _sz]),
    "orbfe_debug_octree": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, C.POINTER(_i)]),
    "orbfe_set_profiling": (_i, [_vp, _i]),
    "orbfe_stage_ms": (_i, [_vp, _vp]),
    "orbfe_launch_count": (C.c_longlong, [_vp]),
    "orbfe_extractor_set_rectification": (_i, [_vp, _vp, _vp, _i, _i]),
    "orbfe_set_max_bytes": (_i, [_vp, _ull]),
    "orbfe_frame_geometry": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_ull), C.POINTER(_ull)]),
    "orbfe_descriptor_distance": (_i, [_vp, _vp, _i, _vp, _i]),
    "orbfe_knn2": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i]),
    "orbfe_knn2_device": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "orbfe_knn2_merge": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i]),
    "orbfe_knn2_merge_device": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbfe_knn2_merge_peers_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbfe_knn2_merge_packed_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbfe_search_by_projection": (_i, [C.POINTER(FrameView), C.POINTER(ProjPoints), C.POINTER(SearchParams),
                                        _vp, _vp, _vp, _vp, _i]),
    "orbfe_map_shard_create": (_i, [C.POINTER(ProjPoints), _i, _i, C.POINTER(_vp)]),
    "orbfe_map_shard_destroy": (None, [_vp]),
    "orbfe_map_shard_set_frame": (_i, [_vp, C.POINTER(FrameView), _vp]),
    "orbfe_claims_init_device": (_i, [_vp, _i, _vp, _vp]),
    "orbfe_map_shard_pass": (_i, [_vp, C.POINTER(SearchParams), _vp, _vp, _vp]),
    "orbfe_claims_min_peers_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbfe_map_shard_finish": (_i, [_vp, _vp, _vp, _vp]),
    "orbfe_map_shard_results": (_i, [_vp, _vp, _vp, _vp]),
    "orbfe_search_by_projection_fisheye": (_i, [C.POINTER(FrameView), C.POINTER(FrameView), _vp, _vp,
                                                C.POINTER(ProjPoints), C.POINTER(ProjPoints), C.POINTER(SearchParams),
                                                _vp, _vp, _vp, _vp, _i]),
    "orbfe_search_for_initialization": (_i, [C.POINTER(FrameView), C.POINTER(FrameView), _vp, _i, _f, _i, _vp, _i]),
    "orbfe_search_window": (_i, [C.POINTER(FrameView), C.POINTER(ProjPoints), C.POINTER(WindowParams), _vp, _vp, _i]),
    "orbfe_search_by_sim3": (_i, [C.POINTER(FrameView), C.POINTER(FrameView), C.POINTER(ProjPoints),
                                  C.POINTER(ProjPoints), _i, _vp, _i]),
    "orbfe_vocabulary_create": (_i, [_i, _i, _i, _vp, _vp, _vp, _i, C.POINTER(_vp)]),
    "orbfe_vocabulary_destroy": (None, [_vp]),
    "orbfe_bow_transform": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "orbfe_bow_transform_device": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbfe_bow_fold_device": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "orbfe_search_by_bow": (_i, [C.POINTER(BowSide), C.POINTER(BowSide), _i, _i, _f, _i, _i, _vp, _vp, _i]),
    "orbfe_search_for_triangulation": (_i, [C.POINTER(TriSide), C.POINTER(TriSide), C.POINTER(TriParams), _vp, _i]),
    "orbfe_cvt_gray": (_i, [_vp, _i, _i, _sz, _i, _i, _vp, _sz, _i]),
    "orbfe_cvt_gray_device": (_i, [_vp, _i, _i, _sz, _i, _i, _vp, _sz, _vp]),
    "orbfe_remap_linear": (_i, [_vp, _i, _i, _sz, _vp, _vp, _i, _i, _vp, _sz, _i]),
    "orbfe_remap_linear_device": (_i, [_vp, _i, _i, _sz, _vp, _vp, _sz, _i, _i, _vp, _sz, _vp]),
    "orbfe_resize_linear": (_i, [_vp, _i, _i, _sz, _i, _i, _vp, _sz, _i]),
    "orbfe_undistort_keypoints": (_i, [_vp, _i, _f, _f, _f, _f, _vp, _i, _vp, _i]),
    "orbfe_distinctive_descriptors": (_i, [_vp, _vp, _i, _vp, _i]),
    "orbfe_kb8_project": (_i, [_vp, _vp, _i, _vp, _i]),
    "orbfe_kb8_unproject": (_i, [_vp, _f, _vp, _i, _vp, _i]),
    "orbfe_kb8_triangulate_matches": (_i, [_vp, _f, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i]),
    "orbfe_debug_kb8_null_vectors": (_i, [_vp, _i, _vp, _i]),
    "orbfe_stereo_match_batch_device": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "orbfe_knn2_batch_device": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbfe_stereo_match": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _f, _f, _vp, _vp]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load libORBfe_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # a fresh checkout (built artefacts are git-ignored): compile the CUDA sources once
            import shutil
            import subprocess
            if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
                subprocess.call(["make", "-s", "-j8", "-C", os.path.join(os.path.dirname(_HERE), "csrc")])
        if not os.path.exists(LIB_PATH):
            raise OrbfeError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                                       "g.build()'` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError here == the ABI in include/orbfe.h is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().orbfe_last_error().decode("utf-8", "replace")


def check(rc, allow=()):
    if rc < 0 and rc not in allow:
        raise OrbfeError(rc, last_error())
    return rc


def ptr(a):
    """void* of a numpy array (host) or a torch tensor (device or pinned host)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())
