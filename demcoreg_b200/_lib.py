"""ctypes binding of ``libdemcoreg_b200.so`` (the C-ABI declared in ``include/demcoreg_b200.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
visible when a compute entry point is called, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DCR_LIB: another build of the SAME library (A/B timing of kernel variants, profiles/experiments); never a fallback
LIB_PATH = os.environ.get("DCR_LIB") or os.path.join(_HERE, "csrc", "libdemcoreg_b200.so")

NBINS = 360
M_BASE, M_MAXDZ, M_MAD, M_SLOPE, M_ASPECT = 0x01, 0x02, 0x04, 0x08, 0x10
M_DIFF = M_BASE | M_MAXDZ | M_MAD | M_SLOPE
ITER_REMOVE_OUTLIERS, ITER_PROFILE, ITER_RADIX_ONLY, ITER_APPLY_Z, ITER_TAN_SLOPE = 1, 2, 4, 8, 16
ITER_RADIX_BINS = 64
ITER_NO_BINS = 128
STATUS_RERUN_RADIX = 100
SAD_PER_OFFSET = 1


class RasterDesc(C.Structure):
    _fields_ = [("gt", C.c_double * 6), ("width", C.c_int32), ("height", C.c_int32),
                ("ndv", C.c_double), ("has_ndv", C.c_int32), ("_pad", C.c_int32)]


class WarpGeom(C.Structure):
    _fields_ = [("ix0", C.c_int32), ("iy0", C.c_int32), ("nx0", C.c_int32), ("ny0", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("wx", C.c_double * 4), ("wy", C.c_double * 4),
                ("identity", C.c_int32), ("_pad", C.c_int32)]


class Grid(C.Structure):
    _fields_ = [("gt", C.c_double * 6), ("width", C.c_int32), ("height", C.c_int32)]


class FrontArgs(C.Structure):
    _fields_ = [
        ("ref", C.c_void_p), ("ref_stride", C.c_int64), ("ref_desc", RasterDesc), ("ref_geom", WarpGeom),
        ("src", C.c_void_p), ("src_stride", C.c_int64), ("src_desc", RasterDesc), ("src_geom", WarpGeom),
        ("smask", C.c_void_p), ("smask_stride", C.c_int64), ("smask_h", C.c_int32), ("smask_w", C.c_int32),
        ("smask_nx0", C.c_int32), ("smask_ny0", C.c_int32),
        ("h", C.c_int32), ("w", C.c_int32),
        ("ewres", C.c_double), ("nsres", C.c_double), ("dst_ndv", C.c_double),
        ("max_dz", C.c_float), ("use_max_dz", C.c_int32), ("slope_lo", C.c_float), ("slope_hi", C.c_float),
        ("dh", C.c_void_p), ("slope", C.c_void_p), ("aspect", C.c_void_p), ("maskbits", C.c_void_p),
        ("ref_w", C.c_void_p), ("src_w", C.c_void_p),
        ("row_begin", C.c_int32), ("row_end", C.c_int32),
        ("ref_row0", C.c_int32), ("ref_nrows", C.c_int32), ("src_row0", C.c_int32), ("src_nrows", C.c_int32),
        ("smask_row0", C.c_int32), ("smask_nrows", C.c_int32),
        ("src_ndz", C.c_int32), ("_pad2", C.c_int32), ("src_dz", C.c_float * 16),
    ]


class NuthStats(C.Structure):
    _fields_ = [("n_common", C.c_int64), ("n_final", C.c_int64), ("mean_y", C.c_float), ("std_y", C.c_float),
                ("rmin", C.c_float), ("rmax", C.c_float), ("bin_count", C.c_int64 * NBINS),
                ("bin_med", C.c_float * NBINS)]


class IterResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_valid_bins", C.c_int32), ("count_base", C.c_int64),
                ("count_maxdz", C.c_int64), ("count_mad", C.c_int64), ("count_final", C.c_int64),
                ("count_slope", C.c_int64), ("stats_stride", C.c_int64),
                ("med1", C.c_float), ("nmad1", C.c_float), ("mad_lo", C.c_float), ("mad_hi", C.c_float),
                ("dz_med", C.c_float), ("med_dh", C.c_float), ("med_slope", C.c_float), ("c_seed", C.c_float),
                ("nuth", NuthStats), ("fit", C.c_double * 3), ("dx", C.c_double), ("dy", C.c_double),
                ("dz", C.c_double), ("stage_ms", C.c_float * 8), ("select_engine", C.c_int32),
                ("z_applied", C.c_int32), ("bin_engine", C.c_int32), ("_pad", C.c_int32)]


ALIGN_MAX_PENDING = 4


class AlignIter(C.Structure):
    _fields_ = [("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("status", C.c_int32),
                ("select_engine", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("bin_engine", C.c_int32), ("_pad", C.c_int32), ("fit", C.c_double * 3), ("dz_med", C.c_float), ("med_slope", C.c_float)]


class AlignState(C.Structure):
    _fields_ = [
        ("ref", C.c_void_p), ("ref_stride", C.c_int64), ("ref_desc", RasterDesc),
        ("src", C.c_void_p), ("src_stride", C.c_int64), ("src_desc", RasterDesc),
        ("smask", C.c_void_p), ("smask_stride", C.c_int64), ("smask_desc", RasterDesc),
        ("dh", C.c_void_p), ("slope", C.c_void_p), ("aspect", C.c_void_p), ("maskbits", C.c_void_p),
        ("out_capacity", C.c_int64),
        ("tol", C.c_double), ("max_offset", C.c_double),
        ("max_dz", C.c_float), ("slope_lo", C.c_float), ("slope_hi", C.c_float),
        ("max_iter", C.c_int32), ("remove_outliers", C.c_int32),
        ("n_done", C.c_int32), ("finished", C.c_int32), ("exit_code", C.c_int32), ("src_ndz", C.c_int32),
        ("src_dz", C.c_float * 16),
        ("dx_total", C.c_double), ("dy_total", C.c_double), ("dz_total", C.c_float), ("_pad", C.c_float),
        ("last", IterResult),
    ]


class Poly2d(C.Structure):
    _fields_ = [("order", C.c_int32), ("_pad", C.c_int32), ("x0", C.c_double), ("y0", C.c_double), ("xs", C.c_double),
                ("ys", C.c_double), ("coeff_norm", C.c_double * 16), ("coeff", C.c_double * 16)]


# every symbol include/demcoreg_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "dcr_last_error": (C.c_char_p, []),
    "dcr_version": (C.c_int, []),
    "dcr_struct_size": (C.c_int, [C.c_int]),
    "dcr_plan_intersection": (C.c_int, [C.POINTER(RasterDesc), C.POINTER(RasterDesc), C.c_double, C.POINTER(Grid),
                                        C.POINTER(WarpGeom), C.POINTER(WarpGeom)]),
    "dcr_plan_onto_grid": (C.c_int, [C.POINTER(RasterDesc), C.POINTER(Grid), C.POINTER(WarpGeom)]),
    "dcr_warp_cubic": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_int, C.POINTER(WarpGeom), _P,
                                 C.c_int, C.c_int, C.c_int64, C.c_float, _P]),
    "dcr_horn_slope_aspect": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double,
                                        _P, _P, _P]),
    "dcr_warp_diff_horn": (C.c_int, [C.POINTER(FrontArgs), _P]),
    "dcr_select_workspace_bytes": (C.c_size_t, []),
    "dcr_select_quantiles": (C.c_int, [_P, _P, C.c_uint32, C.c_int64, C.c_int, C.c_float, C.POINTER(C.c_double),
                                       C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64), _P, C.c_size_t, _P]),
    "dcr_select_workspace_bytes_n": (C.c_size_t, [C.c_int64]),
    "dcr_select_quantiles_radix": (C.c_int, [_P, _P, C.c_uint32, C.c_int64, C.c_int, C.c_float, C.POINTER(C.c_double),
                                             C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64), _P, C.c_size_t, _P]),
    "dcr_moments": (C.c_int, [_P, _P, C.c_uint32, C.c_int64, C.POINTER(C.c_double), _P, C.c_size_t, _P]),
    "dcr_compress_strided": (C.c_int, [_P, _P, C.c_uint32, C.c_int64, C.c_int64, _P, C.c_int64,
                                       C.POINTER(C.c_int64), _P, C.c_size_t, _P]),
    "dcr_nuth_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "dcr_nuth_bin_stats": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64, C.c_int,
                                     C.POINTER(NuthStats), _P, C.c_size_t, _P]),
    "dcr_ncc_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dcr_ncc_corr": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P, _P, C.POINTER(C.c_double),
                               C.POINTER(C.c_double), _P, C.c_size_t, _P]),
    "dcr_sad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dcr_sad_search": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_double), _P,
                                 C.c_size_t, _P]),
    "dcr_sad_search_ex": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_double),
                                    C.POINTER(C.c_int), C.POINTER(C.c_float), _P, C.c_size_t, _P]),
    "dcr_resample_cubic": (C.c_int, [_P, C.POINTER(RasterDesc), C.c_int64, C.POINTER(Grid), _P, C.c_int64, C.c_double, _P]),
    "dcr_axis_median_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dcr_axis_median": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "dcr_axis_subtract": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _P, C.c_int, _P]),
    "dcr_binary_dilate": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dcr_mask_from_classes": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "dcr_mask_threshold": (C.c_int, [_P, C.c_int64, C.c_float, C.c_int, _P, _P]),
    "dcr_fit_nuth": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]),
    "dcr_apply_z_shift": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_float, _P, _P]),
    "dcr_apply_z_shift_seq": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_double, C.POINTER(C.c_float), C.c_int, _P]),
    "dcr_iteration_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "dcr_band_nphases": (C.c_int, []),
    "dcr_band_phase": (C.c_int, [C.POINTER(FrontArgs), C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t,
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "dcr_band_result": (C.c_int, [C.POINTER(FrontArgs), _P, C.c_size_t, C.POINTER(IterResult), _P]),
    "dcr_comm_unique_id": (C.c_int, [_P, C.c_size_t]),
    "dcr_comm_init": (C.c_int, [_P, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "dcr_comm_set": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "dcr_comm_destroy": (C.c_int, [_P]),
    "dcr_comm_rank": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dcr_band_exchange_halos": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, _P]),
    "dcr_band_iteration": (C.c_int, [C.POINTER(FrontArgs), C.c_int, _P, C.POINTER(IterResult), _P, C.c_size_t, _P]),
    "dcr_nuth_iteration": (C.c_int, [C.POINTER(FrontArgs), C.c_int, C.POINTER(IterResult), _P, C.c_size_t, _P]),
    "dcr_align_nuth": (C.c_int, [C.POINTER(AlignState), C.c_int, C.c_int, C.POINTER(AlignIter), C.c_int,
                                 C.POINTER(C.c_int), _P, C.c_size_t, _P]),
    "dcr_align_flush_z": (C.c_int, [C.POINTER(AlignState), _P]),
    "dcr_polyfit2d_workspace_bytes": (C.c_size_t, []),
    "dcr_polyfit2d": (C.c_int, [_P, _P, C.c_uint32, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_int,
                                C.POINTER(Poly2d), C.POINTER(C.c_int64), _P, C.c_size_t, _P]),
    "dcr_polyval2d": (C.c_int, [C.POINTER(Poly2d), C.POINTER(C.c_double), C.c_int, C.c_int, _P, C.c_int64, _P]),
    "dcr_polyval2d_apply": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_double, C.c_int,
                                      C.POINTER(Poly2d), C.c_double, _P]),
}

_lib = None


class DcrError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DcrError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(demcoreg_b200 has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().dcr_last_error()
        raise DcrError("demcoreg_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise DcrError("demcoreg_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch
