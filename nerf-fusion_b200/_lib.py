"""ctypes binding of libdifusion_b200.so (the C-ABI in include/difusion_b200.h).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).  There is no CPU fallback: if the
shared object is missing or a call fails, this module raises.
"""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdifusion_b200.so"


class DfbError(RuntimeError):
    pass


class MapParams(C.Structure):
    """dfb_map_params (include/difusion_b200.h)."""
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("bound_min", C.c_float * 3), ("voxel_size", C.c_float), ("div_mode", C.c_int32),
                ("prune_min_vox_obs", C.c_int32), ("ignore_count_th", C.c_float), ("encoder_count_th", C.c_float)]


class GnConfig(C.Structure):
    """dfb_gn_config"""
    _fields_ = [("n_groups", C.c_int32), ("n_iter", C.c_int32 * 8), ("use_sdf", C.c_int32 * 8), ("rgb_level", C.c_int32 * 8),
                ("sdf_robust", C.c_int32), ("sdf_robust_k", C.c_float), ("rgb_robust", C.c_int32), ("rgb_robust_k", C.c_float),
                ("rgb_weight", C.c_float), ("rgb_min_grad_scale", C.c_float), ("rgb_max_depth_delta", C.c_float)]


class RgbLevel(C.Structure):
    """dfb_rgb_level"""
    _fields_ = [("prev_I", C.c_void_p), ("prev_D", C.c_void_p), ("cur_I", C.c_void_p), ("cur_D", C.c_void_p), ("cur_G", C.c_void_p),
                ("H", C.c_int32), ("W", C.c_int32)]


SHARD_MAX_WORLD = 8


class Shard(C.Structure):
    """dfb_shard (include/difusion_b200.h): one rank of the sharded map, all pointers are device pointers."""
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("bound_min", C.c_float * 3), ("voxel_size", C.c_float),
                ("div_mode", C.c_int32), ("prune_min_vox_obs", C.c_int32), ("encoder_count_th", C.c_float),
                ("world", C.c_int32), ("rank", C.c_int32),
                ("indexer_local", C.c_void_p), ("latent_vecs", C.c_void_p), ("latent_vecs_pos", C.c_void_p), ("voxel_obs_count", C.c_void_p),
                ("capacity", C.c_int32), ("cand_bits", C.c_void_p), ("grid_count", C.c_void_p), ("acc", C.c_void_p), ("acc_n", C.c_void_p),
                ("touched", C.c_void_p), ("counters", C.c_void_p), ("delta_list", C.c_void_p), ("next_delta", C.c_void_p),
                ("n_next_delta", C.c_void_p), ("delta_cap", C.c_int32),
                ("pts_inbox", C.c_void_p), ("pts_cap", C.c_int32), ("pts_count", C.c_void_p),
                ("ids_inbox", C.c_void_p), ("ids_cap", C.c_int32), ("ids_count", C.c_void_p),
                ("smp_inbox", C.c_void_p), ("smp_cap", C.c_int32), ("smp_count", C.c_void_p),
                ("dlt_inbox", C.c_void_p), ("dlt_cap", C.c_int32), ("dlt_count", C.c_void_p),
                ("peer_pts", C.c_void_p * SHARD_MAX_WORLD), ("peer_pts_count", C.c_void_p * SHARD_MAX_WORLD),
                ("peer_ids", C.c_void_p * SHARD_MAX_WORLD), ("peer_ids_count", C.c_void_p * SHARD_MAX_WORLD),
                ("peer_smp", C.c_void_p * SHARD_MAX_WORLD), ("peer_smp_count", C.c_void_p * SHARD_MAX_WORLD),
                ("peer_dlt", C.c_void_p * SHARD_MAX_WORLD), ("peer_dlt_count", C.c_void_p * SHARD_MAX_WORLD),
                ("flags", C.c_void_p), ("peer_flags", C.c_void_p * SHARD_MAX_WORLD), ("fuse_publish", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_SZ = C.c_size_t
_I64 = C.c_int64
_FP = C.POINTER(C.c_float)
_MP = C.POINTER(MapParams)

# name -> (restype, argtypes); every symbol include/difusion_b200.h declares
SIGNATURES = {
    "dfb_version": (_I, []),
    "dfb_last_error": (C.c_char_p, []),
    "dfb_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "dfb_decoder_blob_floats": (_SZ, []),
    "dfb_encoder_blob_floats": (_SZ, []),
    "dfb_set_decoder_engine": (_I, [_I]),
    "dfb_get_decoder_engine": (_I, []),
    "dfb_set_gn_reserved_sms": (_I, [_I]),
    "dfb_set_encoder_engine": (_I, [_I]),
    "dfb_get_encoder_engine": (_I, []),
    "dfb_ingest_frame": (_I, [_P, _P, _I, _I, _F, _I, _F, _F, _I, _P, _P, _P]),
    "dfb_transform_points": (_I, [_P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P]),
    "dfb_unproject_depth": (_I, [_P, _I, _I, _F, _F, _F, _F, _P, _P]),
    "dfb_pcproc_ws_bytes": (_SZ, [_I]),
    "dfb_remove_radius_outlier": (_I, [_P, _I, _I, _F, _P, _P, _SZ, _P]),
    "dfb_estimate_normals": (_I, [_P, _I, _I, _F, _FP, _P, _P, _SZ, _P]),
    "dfb_scatter_mean_ws_bytes": (_SZ, [_I, _I]),
    "dfb_scatter_mean": (_I, [_P, _P, _I, _I, _I, _P, _P, _SZ, _P]),
    "dfb_box_filter_ws_bytes": (_SZ, [_I]),
    "dfb_point_box_filter": (_I, [_P, _P, _I, _F, _I, _P, _P, _P, _P, _SZ, _P]),
    "dfb_preprocess_ws_bytes": (_SZ, [_I, _I]),
    "dfb_preprocess_frame": (_I, [_P, _I, _I, _F, _F, _F, _F, _I, _F, _I, _F, _FP, _F, _I, _P, _P, _P, _P, _SZ, _P]),
    "dfb_groupby_sum": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "dfb_gradient_xy": (_I, [_P, _I, _I, _P, _P]),
    "dfb_frame_images": (_I, [_P, _P, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dfb_rgb_odometry": (_I, [_P, _P, _P, _P, _P, _I, _I, _FP, _FP, _FP, _F, _F, _P, _P, _P]),
    "dfb_rgb_hg": (_I, [_P, _P, _P, _P, _P, _I, _I, _FP, _FP, _FP, _F, _F, _I, _F, _I, _P, _P]),
    "dfb_integrate_ws_bytes": (_SZ, [_I, _I64]),
    "dfb_integrate_plan": (_I, [_MP, _P, _P, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "dfb_integrate_commit": (_I, [_MP, _P, _P, _I, _P, _P, _P, _P, _P, _I64, _I64, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "dfb_encoder_forward": (_I, [_P, _I, _P, _P, _P]),
    "dfb_decoder_forward": (_I, [_P, _I, _P, _P, _P, _P]),
    "dfb_get_sdf": (_I, [_MP, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dfb_sdf_hg": (_I, [_MP, _P, _I, _FP, _P, _P, _P, _P, _I, _F, _I, _P, _P]),
    "dfb_gauss_newton": (_I, [_MP, C.POINTER(GnConfig), _P, _I, _P, _P, _P, _P, _P, C.POINTER(RgbLevel), C.POINTER(C.c_double),
                              C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P, C.POINTER(C.c_int32), _P]),
    "dfb_decode_cubes_ws_bytes": (_SZ, [_I, _I]),
    "dfb_decode_cubes": (_I, [_P, _P, _I, _I, _F, _P, _P, _P, _P, _SZ, _P]),
    "dfb_marching_cubes": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P]),
    "dfb_latent_adam_step": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _I, _F, _F, _P]),
    "dfb_shard_counter_ints": (_I, []),
    "dfb_shard_local_cells": (_I64, [_I, _I, _I, _I]),
    "dfb_shard_phase1": (_I, [C.POINTER(Shard), _P, _P, _I, _P]),
    "dfb_shard_phase2": (_I, [C.POINTER(Shard), _P]),
    "dfb_shard_phase3": (_I, [C.POINTER(Shard), _P]),
    "dfb_shard_phase4": (_I, [C.POINTER(Shard), _P]),
    "dfb_shard_phase5": (_I, [C.POINTER(Shard), _P, _P, _P]),
    "dfb_shard_barrier": (_I, [C.POINTER(Shard), _I, _I, _P]),
    "dfb_peer_alloc": (_I, [_SZ, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    "dfb_peer_open": (_I, [C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "dfb_peer_close": (_I, [_P]),
    "dfb_peer_free": (_I, [_P]),
}

# kernels (memset nodes excluded) each entry point launches; used for the bench's `gpu_launches` claim (dfb_preprocess_frame
# counted against the ncu launch list of the bench, profiles/r02_launches_final.txt, minus the second grid build)
KERNELS_PER_CALL = {
    "dfb_ingest_frame": 1, "dfb_transform_points": 1, "dfb_unproject_depth": 1, "dfb_remove_radius_outlier": 9, "dfb_estimate_normals": 9, "dfb_scatter_mean": 5,
    "dfb_point_box_filter": 13, "dfb_preprocess_frame": 25, "dfb_groupby_sum": 1, "dfb_gradient_xy": 1, "dfb_frame_images": 2, "dfb_rgb_odometry": 1, "dfb_rgb_hg": 2,
    "dfb_integrate_plan": 5, "dfb_integrate_commit": 4, "dfb_encoder_forward": 1, "dfb_decoder_forward": 1,
    "dfb_get_sdf": 1, "dfb_sdf_hg": 2, "dfb_gauss_newton": 0, "dfb_decode_cubes": 3, "dfb_marching_cubes": 1,
    "gn_kernel": 1,                    # launched inside dfb_gauss_newton (count reported through h_stats[7])
    "dfb_latent_adam_step": 2, "dfb_shard_phase1": 2, "dfb_shard_phase2": 3, "dfb_shard_phase3": 2, "dfb_shard_phase4": 3, "dfb_shard_phase5": 4, "dfb_shard_barrier": 1,
}
CALLS = {}


class _Api:
    """Typed entry points; every call is counted in CALLS (name -> number of calls)."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(cdll, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, self._counted(name, fn) if name in KERNELS_PER_CALL else fn)

    @staticmethod
    def _counted(name, fn):
        def call(*a):
            CALLS[name] = CALLS.get(name, 0) + 1
            return fn(*a)
        return call


def kernel_launches():
    return sum(KERNELS_PER_CALL[k] * v for k, v in CALLS.items())


_lib = None


def load():
    """dlopen the library and type every entry point.  Raises DfbError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("DFB_LIB", LIB_PATH))
    if not path.exists():
        raise DfbError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"(nvcc, sm_100a).  difusion_b200 has no CPU fallback.")
    _lib = _Api(C.CDLL(str(path)))
    return _lib


def check(rc):
    if rc != 0:
        raise DfbError(f"difusion_b200 call failed ({rc}): {load().dfb_last_error().decode()}")


def fptr(values):
    """host float array argument"""
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr
