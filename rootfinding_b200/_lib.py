"""ctypes binding of include/yroots_b200.h.

The product library is rootfinding_b200/csrc/libyroots_b200.so (hand-written sm_100a
CUDA, built by __graft_entry__.build() / `make -C rootfinding_b200/csrc`).  There is no
CPU fallback: if the library or a CUDA device is missing, loading raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.environ.get("YR_LIB_OVERRIDE") or os.path.join(_HERE, "csrc", "libyroots_b200.so")   # override: kernel experiments only

u64, i64, i32 = C.c_uint64, C.c_int64, C.c_int32
vp, dp = C.c_void_p, C.c_void_p          # device pointers travel as integers


class YrParams(C.Structure):
    _fields_ = [("exact", i32), ("constant_check", i32), ("low_dim_quad_check", i32), ("all_dim_quad_check", i32),
                ("max_zoom_count", i32), ("use_fma", i32), ("trim", i32), ("trim_policy", i32),
                ("trim_rel_tol", C.c_double), ("trim_abs_tol", C.c_double)]


class YrLaunch(C.Structure):
    _fields_ = [("threads", i32), ("ctas", i32), ("ws_in_smem", i32), ("warp_per_box", i32),
                ("ws_doubles", u64), ("gws", dp), ("gws_stride", u64)]


class YrLevelDesc(C.Structure):
    _fields_ = [("dim", i32), ("reserved", i32),
                ("recs_in", vp), ("data_in", dp), ("n_in", u64),
                ("recs_out", vp), ("data_out", dp), ("cap_recs_out", u64), ("cap_data_out", u64),
                ("results", vp), ("cap_results", u64),
                ("nodes", vp), ("cap_nodes", u64),
                ("result_base", u64), ("node_base", u64),
                ("counters", vp), ("prm", YrParams), ("launch", YrLaunch),
                ("work", vp), ("work_bytes", u64), ("pipeline", i32), ("resume_subdivide", i32),
                ("max_box_doubles", u64), ("max_tensor_doubles", u64), ("max_extent", u64)]


class YrInitDesc(C.Structure):
    _fields_ = [("dim", i32), ("reserved", i32),
                ("coeffs", dp), ("coeff_off", vp), ("shapes", vp), ("errs", dp),
                ("job_system", vp), ("job_restrict", vp), ("job_alpha", dp), ("job_beta", dp), ("job_iv", dp),
                ("job_level", vp), ("job_next_mid", dp), ("job_parent", vp), ("njobs", u64),
                ("recs_out", vp), ("data_out", dp), ("cap_recs_out", u64), ("cap_data_out", u64),
                ("counters", vp), ("prm", YrParams), ("launch", YrLaunch)]


class YrCounters(C.Structure):
    _fields_ = [(n, u64) for n in ("next_box", "n_children", "data_used", "n_results", "n_nodes", "boxes", "zooms",
                                   "subdivisions", "discard_const", "discard_quad", "discard_zoom", "entry_coeffs",
                                   "restrict_macs", "max_child_doubles", "max_child_extent", "max_level", "overflow",
                                   "max_child_tensor", "no_axis", "reserved1")]


class YrFrontierDesc(C.Structure):
    _fields_ = [("dim", i32), ("time_kernels", i32), ("recs_in", vp), ("data_in", dp), ("n_in", u64),
                ("max_extent", u64), ("result_base", u64), ("node_base", u64), ("prm", YrParams)]


class YrFrontierOut(C.Structure):
    _fields_ = [("results", vp), ("n_results", u64), ("nodes", vp), ("n_nodes", u64), ("order", vp),
                ("counters", YrCounters),
                ("rounds", u64), ("launches", u64), ("slots", u64), ("tensor_ms", C.c_double), ("scalar_ms", C.c_double)]


class YrF2OpsDesc(C.Structure):
    _fields_ = [("op", i32), ("flat", i32), ("n", u64), ("coeffs", vp), ("off", vp), ("shape", vp), ("err", vp),
                ("alpha", vp), ("beta", vp), ("iv", vp), ("next_mid", vp), ("level", vp), ("prm", YrParams),
                ("cap_out", u64), ("cap_coeffs", u64), ("n_out", vp),
                ("out_coeffs", vp), ("out_off", vp), ("out_shape", vp), ("out_err", vp), ("out_sumabs", vp),
                ("out_trim_shape", vp), ("out_trim_err", vp), ("out_iv", vp), ("out_next_mid", vp), ("out_parent", vp)]


class YrFrontierProgress(C.Structure):
    _fields_ = [("pending", u64), ("pending_restrict", u64), ("pending_subdivide", u64), ("rounds", u64), ("boxes", u64),
                ("arena_doubles", u64)]


# yr_sorted_rec (include/yroots_b200.h): root[dim]
def sorted_dtype(dim):
    return np.dtype([("index", np.uint32), ("system", np.int32), ("flags", np.uint32), ("collector", np.int32),
                     ("root", np.float64, (dim,))])


SORTED_DTYPE = sorted_dtype(2)

COUNTER_FIELDS = ["next_box", "n_children", "data_used", "n_results", "n_nodes", "boxes", "zooms", "subdivisions",
                  "discard_const", "discard_quad", "discard_zoom", "entry_coeffs", "restrict_macs",
                  "max_child_doubles", "max_child_extent", "max_level", "overflow", "max_child_tensor",
                  "no_axis", "reserved1"]
COUNTER_BYTES = 8 * len(COUNTER_FIELDS)

# every symbol include/yroots_b200.h declares
EXPORTS = ["yr_abi_version", "yr_last_error", "yr_last_error_code", "yr_build_info", "yr_default_params", "yr_record_layout",
           "yr_workspace_doubles", "yr_level_work_bytes", "yr_static_smem_bytes", "yr_device_limits", "yr_init_boxes", "yr_process_level", "yr_bils_batched",
           "yr_quadcheck_batched", "yr_dct1_axis", "yr_poly_eval_axis", "yr_abs_axis_mean",
           "yr_frontier_solve", "yr_frontier_release", "yr_frontier_fits", "yr_frontier_trim", "yr_fetch",
           "yr_frontier_begin", "yr_frontier_advance", "yr_frontier_export_bound", "yr_frontier_export",
           "yr_frontier_import", "yr_frontier_finish", "yr_decay_probe", "yr_upper_axis_apply", "yr_f2_tensor_ops", "yr_frontier_fits_dim", "yr_gather_by_system", "yr_frontier_adopt", "yr_device_copy"]

_RESULT_FIELDS = ["root", "box", "comb_iv", "iv_lo", "cell_iv", "next_mid", "system", "parent_node", "collector",
                  "level", "flags", "path_hi", "path_lo"]
_NODE_FIELDS = ["iv", "next_mid", "system", "parent_node", "level", "nsplit"]
_BOX_FIELDS = ["data_off", "shape", "err", "iv", "A", "B", "next_mid", "pf_iv", "pfA", "pfB", "flags", "system",
               "parent_node", "collector", "level", "path_hi", "path_lo"]


def bind(lib):
    """Attach prototypes to a loaded library and return it."""
    lib.yr_abi_version.restype = C.c_int
    lib.yr_last_error.restype = C.c_char_p
    lib.yr_last_error_code.restype = C.c_int
    lib.yr_build_info.restype = C.c_char_p
    lib.yr_default_params.argtypes = [C.POINTER(YrParams)]
    lib.yr_default_params.restype = None
    lib.yr_record_layout.argtypes = [i32, i32, C.POINTER(i64), i32]
    lib.yr_record_layout.restype = i64
    lib.yr_workspace_doubles.argtypes = [i32, u64, u64, u64, i32]
    lib.yr_workspace_doubles.restype = u64
    lib.yr_level_work_bytes.argtypes = [i32, u64]
    lib.yr_level_work_bytes.restype = u64
    lib.yr_static_smem_bytes.argtypes = [i32]
    lib.yr_static_smem_bytes.restype = i64
    lib.yr_device_limits.argtypes = [C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    lib.yr_device_limits.restype = C.c_int
    lib.yr_init_boxes.argtypes = [C.POINTER(YrInitDesc), vp]
    lib.yr_init_boxes.restype = C.c_int
    lib.yr_process_level.argtypes = [C.POINTER(YrLevelDesc), vp]
    lib.yr_process_level.restype = C.c_int
    lib.yr_bils_batched.argtypes = [i32, i64, dp, dp, dp, dp, vp, dp, vp, vp]
    lib.yr_bils_batched.restype = C.c_int
    lib.yr_quadcheck_batched.argtypes = [i32, i64, dp, vp, vp, dp, i32, vp, vp]
    lib.yr_quadcheck_batched.restype = C.c_int
    lib.yr_dct1_axis.argtypes = [dp, dp, i64, i64, i64, vp]
    lib.yr_dct1_axis.restype = C.c_int
    lib.yr_poly_eval_axis.argtypes = [dp, dp, dp, i64, i64, i64, i64, i32, vp]
    lib.yr_poly_eval_axis.restype = C.c_int
    lib.yr_abs_axis_mean.argtypes = [dp, dp, dp, i64, i64, i64, vp]
    lib.yr_abs_axis_mean.restype = C.c_int
    lib.yr_frontier_solve.argtypes = [C.POINTER(YrFrontierDesc), C.POINTER(YrFrontierOut), vp]
    lib.yr_frontier_solve.restype = C.c_int
    lib.yr_frontier_release.argtypes = [C.POINTER(YrFrontierOut), vp]
    lib.yr_frontier_release.restype = C.c_int
    lib.yr_frontier_trim.argtypes = [vp]
    lib.yr_frontier_trim.restype = C.c_int
    lib.yr_frontier_fits.argtypes = [u64]
    lib.yr_frontier_fits.restype = C.c_int
    lib.yr_fetch.argtypes = [vp, vp, u64, vp]
    lib.yr_fetch.restype = C.c_int
    lib.yr_frontier_adopt.argtypes = [C.c_int32, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64), C.c_int32, C.c_int32,
                                      C.POINTER(YrFrontierOut), vp]
    lib.yr_frontier_adopt.restype = C.c_int
    lib.yr_device_copy.argtypes = [vp, vp, u64, vp]
    lib.yr_device_copy.restype = C.c_int
    lib.yr_gather_by_system.argtypes = [vp, u64, C.c_uint32, C.c_uint32, vp, u64, vp, vp, u64, vp, vp]
    lib.yr_gather_by_system.restype = C.c_int
    lib.yr_frontier_fits_dim.argtypes = [i32, u64, u64]
    lib.yr_frontier_fits_dim.restype = C.c_int
    lib.yr_f2_tensor_ops.argtypes = [C.POINTER(YrF2OpsDesc), vp]
    lib.yr_f2_tensor_ops.restype = C.c_int
    lib.yr_upper_axis_apply.argtypes = [dp, dp, dp, i64, i64, i64, vp]
    lib.yr_upper_axis_apply.restype = C.c_int
    lib.yr_decay_probe.argtypes = [i32, i32, i32, dp, vp, vp, dp, vp, vp, dp, i64, C.c_double, C.c_double, C.c_double, dp, vp, vp]
    lib.yr_decay_probe.restype = C.c_int
    lib.yr_frontier_begin.argtypes = [C.POINTER(YrFrontierDesc), vp]
    lib.yr_frontier_begin.restype = C.c_int
    lib.yr_frontier_advance.argtypes = [u64, C.POINTER(YrFrontierProgress), vp]
    lib.yr_frontier_advance.restype = C.c_int
    lib.yr_frontier_export_bound.argtypes = [u64]
    lib.yr_frontier_export_bound.restype = u64
    lib.yr_frontier_export.argtypes = [u64, vp, u64, C.POINTER(u64), C.POINTER(u64), vp]
    lib.yr_frontier_export.restype = C.c_int
    lib.yr_frontier_import.argtypes = [vp, u64, vp]
    lib.yr_frontier_import.restype = C.c_int
    lib.yr_frontier_finish.argtypes = [C.POINTER(YrFrontierOut), vp]
    lib.yr_frontier_finish.restype = C.c_int
    if lib.yr_abi_version() != 1:
        raise RuntimeError("yroots_b200: ABI version mismatch")
    return lib


NO_AXIS_TEXT = ("a box has no axis left to subdivide along (point-sized box at level <= 5); the reference raises here: "
                "IndexError in getInverseOrder, ChebyshevSubdivisionSolver.py:1010")


class NoAxisLeft(IndexError):
    """The input on which the reference's getInverseOrder raises IndexError (a zoom collapsed the box to a point before
    level 6, then the final step asks for a subdivision: getSubdivisionDims :1036-1049 keeps no axis)."""


class DeviceOutOfMemory(RuntimeError):
    """The device ran out of memory inside the library (yr_last_error_code() == 2): the batch can be retried in parts."""


def check(lib, rc, what):
    if rc != 0:
        msg = f"yroots_b200: {what} failed: {lib.yr_last_error().decode()}"
        code = lib.yr_last_error_code()
        raise (DeviceOutOfMemory if code == 2 else NoAxisLeft if code == 3 else RuntimeError)(msg)


def default_params(lib):
    p = YrParams()
    lib.yr_default_params(C.byref(p))
    return p


def _record_dtype(lib, dim, kind, names):
    buf = (i64 * (3 * 32))()
    size = lib.yr_record_layout(dim, kind, buf, 32)
    fields = {}
    for k, name in enumerate(names):
        fid, off, nbytes = buf[3 * k], buf[3 * k + 1], buf[3 * k + 2]
        assert fid == k
        if name in ("system", "parent_node", "collector", "level", "nsplit"):
            fmt = np.int32
        elif name == "flags":
            fmt = np.uint32
        elif name in ("path_hi", "path_lo", "data_off"):
            fmt = np.uint64
        elif name == "shape":
            fmt = (np.int32, (dim, dim))
        elif nbytes == 8 * dim:
            fmt = (np.float64, (dim,))
        elif nbytes == 16 * dim:
            fmt = (np.float64, (dim, 2))
        else:
            raise AssertionError((name, nbytes))
        fields[name] = (fmt, int(off))
    return np.dtype({"names": list(fields), "formats": [f[0] for f in fields.values()],
                     "offsets": [f[1] for f in fields.values()], "itemsize": int(size)})


def result_dtype(lib, dim):
    return _record_dtype(lib, dim, 1, _RESULT_FIELDS)


def node_dtype(lib, dim):
    return _record_dtype(lib, dim, 2, _NODE_FIELDS)


def box_dtype(lib, dim):
    return _record_dtype(lib, dim, 0, _BOX_FIELDS)


_cuda_lib = None


def load_cuda_library():
    """Load the CUDA library.  Raises when it has not been built -- there is no other path."""
    global _cuda_lib
    if _cuda_lib is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise RuntimeError(
                f"yroots_b200: CUDA library {CUDA_LIB_PATH} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        _cuda_lib = bind(C.CDLL(CUDA_LIB_PATH))
    return _cuda_lib
