"""ctypes binding of libbbm_b200.so -- the same C ABI (include/bbm_b200.h) a Julia package
extension would ``ccall``.  There is no CPU fallback: if the library is missing or no B200 is
present, calls raise."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BBM_B200_LIB") or os.path.join(_HERE, "lib", "libbbm_b200.so")  # override: A/B of builds
CSRC = os.path.join(_HERE, "csrc")

i32, i64, vp, f64, f32, sz = C.c_int32, C.c_int64, C.c_void_p, C.c_double, C.c_float, C.c_size_t
pp = C.POINTER(C.c_void_p)
pi64 = C.POINTER(C.c_int64)


class BBMError(RuntimeError):
    pass


class DimensionMismatch(ValueError):
    """Mirrors Julia's DimensionMismatch (BBM_ERR_DIM)."""


class BandError(IndexError):
    """Mirrors BandedMatrices.BandError (BBM_ERR_BAND)."""


# name -> (restype, argtypes); every symbol include/bbm_b200.h declares
SIGNATURES = {
    "bbm_version": (i32, []),
    "bbm_status_string": (C.c_char_p, [i32]),
    "bbm_create": (i32, [i32, pp]),
    "bbm_destroy": (i32, [vp]),
    "bbm_set_stream": (i32, [vp, vp]),
    "bbm_synchronize": (i32, [vp]),
    "bbm_last_error": (C.c_char_p, [vp]),
    "bbm_set_option": (i32, [vp, C.c_char_p, i64]),
    "bbm_get_option": (i32, [vp, C.c_char_p, pi64]),
    "bbm_launch_count": (i32, [vp, pi64]),
    "bbm_malloc": (i32, [vp, pp, sz]),
    "bbm_free": (i32, [vp, vp]),
    "bbm_host_alloc": (i32, [vp, pp, sz]),
    "bbm_host_free": (i32, [vp, vp]),
    "bbm_memcpy_h2d": (i32, [vp, vp, vp, sz]),
    "bbm_memcpy_d2h": (i32, [vp, vp, vp, sz]),
    "bbm_memcpy_d2d": (i32, [vp, vp, vp, sz]),
    "bbm_memset": (i32, [vp, vp, i32, sz]),
    "bbm_axpy_f64": (i32, [vp, i64, f64, vp, vp]),
    "bbm_axpy_f32": (i32, [vp, i64, f32, vp, vp]),
    "bbm_scal_f64": (i32, [vp, i64, f64, vp]),
    "bbm_scal_f32": (i32, [vp, i64, f32, vp]),
    "bbm_bbb_add_f64": (i32, [vp, vp, vp, vp, f64, vp, f64, vp, vp]),
    "bbm_bbb_add_f32": (i32, [vp, vp, vp, vp, f32, vp, f32, vp, vp]),
    "bbm_sky_add_f64": (i32, [vp, vp, vp, vp, vp, vp, f64, vp, f64, vp, vp]),
    "bbm_sky_add_f32": (i32, [vp, vp, vp, vp, vp, vp, f32, vp, f32, vp, vp]),
    "bbm_bbb_add_diags_f64": (i32, [vp, vp, vp, f64, f64, vp, vp, vp]),
    "bbm_bbb_add_diags_f32": (i32, [vp, vp, vp, f32, f32, vp, vp, vp]),
    "bbm_sky_add_diags_f64": (i32, [vp, vp, vp, f64, f64, vp, vp, vp]),
    "bbm_sky_add_diags_f32": (i32, [vp, vp, vp, f32, f32, vp, vp, vp]),
    "bbm_bbb_desc_create": (i32, [vp, i64, i64, vp, vp, i64, i64, i64, i64, pp]),
    "bbm_bbb_desc_destroy": (i32, [vp]),
    "bbm_bbb_desc_info": (i32, [vp, pi64, pi64, pi64, pi64]),
    "bbm_bbb_mul_f64": (i32, [vp, vp, f64, vp, vp, i64, i64, f64, vp, i64]),
    "bbm_bbb_mul_f32": (i32, [vp, vp, f32, vp, vp, i64, i64, f32, vp, i64]),
    "bbm_bbb_trimul_f64": (i32, [vp, vp, i32, i32, f64, vp, vp, i64, i64, f64, vp, i64]),
    "bbm_bbb_trimul_f32": (i32, [vp, vp, i32, i32, f32, vp, vp, i64, i64, f32, vp, i64]),
    "bbm_bbb_mul_t_f64": (i32, [vp, vp, f64, vp, vp, i64, i64, f64, vp, i64]),
    "bbm_bbb_mul_t_f32": (i32, [vp, vp, f32, vp, vp, i64, i64, f32, vp, i64]),
    "bbm_sky_mul_t_f64": (i32, [vp, vp, f64, vp, vp, i64, i64, f64, vp, i64]),
    "bbm_sky_mul_t_f32": (i32, [vp, vp, f32, vp, vp, i64, i64, f32, vp, i64]),
    "bbm_sky_ql_f64": (i32, [vp, vp, vp, vp]),
    "bbm_sky_ql_lmul_adj_f64": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_ql_ldiv_f64": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_qr_f64": (i32, [vp, vp, vp, vp]),
    "bbm_sky_qr_lmul_adj_f64": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_qr_ldiv_f64": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_ql_f32": (i32, [vp, vp, vp, vp]),
    "bbm_sky_ql_lmul_adj_f32": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_ql_ldiv_f32": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_qr_f32": (i32, [vp, vp, vp, vp]),
    "bbm_sky_qr_lmul_adj_f32": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_sky_qr_ldiv_f32": (i32, [vp, vp, vp, vp, vp, i64, i64]),
    "bbm_bbb_trsv_f64": (i32, [vp, vp, i32, i32, vp, vp, i64, i64]),
    "bbm_bbb_trsv_f32": (i32, [vp, vp, i32, i32, vp, vp, i64, i64]),
    "bbm_bbb_trsv_prepare_f64": (i32, [vp, vp, i32, i32, vp, pp]),
    "bbm_bbb_trsv_prepare_f32": (i32, [vp, vp, i32, i32, vp, pp]),
    "bbm_bbb_trsv_solve_f64": (i32, [vp, vp, i64, i64]),
    "bbm_bbb_trsv_solve_f32": (i32, [vp, vp, i64, i64]),
    "bbm_bbb_tri_plan_destroy": (i32, [vp]),
    "bbm_bbb_sparse_colptr": (i32, [vp, vp, i32, vp, C.POINTER(i64)]),
    "bbm_bbb_sparse_fill_f64": (i32, [vp, vp, vp, vp, i32, vp, vp]),
    "bbm_bbb_sparse_fill_f32": (i32, [vp, vp, vp, vp, i32, vp, vp]),
    "bbm_bbb_matmul_f64": (i32, [vp, vp, vp, vp, f64, vp, vp, f64, vp]),
    "bbm_bbb_matmul_f32": (i32, [vp, vp, vp, vp, f32, vp, vp, f32, vp]),
    "bbm_bbb_mul_host_f64": (i32, [vp, vp, f64, vp, vp, i64, i64, f64, vp, i64]),
    "bbm_bbb_mul_host_f32": (i32, [vp, vp, f32, vp, vp, i64, i64, f32, vp, i64]),
    "bbm_sky_desc_create": (i32, [vp, i64, i64, vp, vp, vp, vp, pp]),
    "bbm_sky_desc_destroy": (i32, [vp]),
    "bbm_sky_desc_info": (i32, [vp, pi64, pi64, pi64, pi64]),
    "bbm_sky_desc_layout": (i32, [vp, vp, vp, vp, vp]),
    "bbm_sky_mul_f64": (i32, [vp, vp, f64, vp, vp, i64, i64, f64, vp, i64]),
    "bbm_sky_mul_f32": (i32, [vp, vp, f32, vp, vp, i64, i64, f32, vp, i64]),
    "bbm_sky_matmul_f64": (i32, [vp, vp, vp, vp, f64, vp, vp, f64, vp]),
    "bbm_sky_matmul_f32": (i32, [vp, vp, vp, vp, f32, vp, vp, f32, vp]),
    "bbm_sky_rmul_f64": (i32, [vp, vp, f64, vp, i64, i64, vp, f64, vp, i64]),
    "bbm_sky_rmul_f32": (i32, [vp, vp, f32, vp, i64, i64, vp, f32, vp, i64]),
    "bbm_sky_trsm_f64": (i32, [vp, vp, i32, i32, vp, vp, i64, i64]),
    "bbm_sky_trsm_f32": (i32, [vp, vp, i32, i32, vp, vp, i64, i64]),
    "bbm_bbb_partition_rows": (i32, [i64, vp, i32, vp]),
    "bbm_bbb_dist_layout": (i32, [i64, i64, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp]),
    "bbm_comm_unique_id": (i32, [vp]),
    "bbm_comm_create": (i32, [vp, vp, i32, i32, pp]),
    "bbm_comm_destroy": (i32, [vp]),
    "bbm_dist_bbb_plan_create": (i32, [vp, i64, i64, vp, vp, i64, i64, i64, i64, vp, pp]),
    "bbm_dist_bbb_plan_destroy": (i32, [vp]),
    "bbm_dist_bbb_plan_layout": (i32, [vp, vp]),
    "bbm_dist_bbb_enable_peer": (i32, [vp, i32]),
    "bbm_dist_bbb_peer_status": (i32, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "bbm_dist_bbb_mul_f64": (i32, [vp, f64, vp, vp, i64, f64, vp]),
    "bbm_dist_bbb_mul_f32": (i32, [vp, f32, vp, vp, i64, f32, vp]),
    "bbm_dist_bbb_mul_host_f64": (i32, [vp, f64, vp, vp, f64, vp]),
    "bbm_dist_bbb_mul_host_f32": (i32, [vp, f32, vp, vp, f32, vp]),
    "bbm_sky_rowrange_structure": (i32, [i64, i64, i64, i64, i64, i64, pi64, pi64, vp, vp]),
    "bbm_dist_sky_mm_layout": (i32, [i64, i64, i64, vp, vp, i64, i64, i64, i64, i32, i32, vp, vp, vp, vp]),
    "bbm_dist_sky_plan_create": (i32, [vp, i64, i64, vp, vp, i64, i64, vp, pp]),
    "bbm_dist_sky_plan_destroy": (i32, [vp]),
    "bbm_dist_sky_plan_layout": (i32, [vp, vp, pi64, pi64, pi64, pi64]),
    "bbm_dist_sky_plan_bandwidths": (i32, [vp, vp, vp]),
    "bbm_dist_sky_mul_f64": (i32, [vp, f64, vp, vp, i64, f64, vp]),
    "bbm_dist_sky_mul_f32": (i32, [vp, f32, vp, vp, i64, f32, vp]),
    "bbm_dist_sky_mm_plan_create": (i32, [vp, i64, i64, i64, vp, vp, vp, i64, i64, i64, i64, vp, pp]),
    "bbm_dist_sky_mm_plan_destroy": (i32, [vp]),
    "bbm_dist_sky_mm_plan_info": (i32, [vp, vp, vp, vp, vp]),
    "bbm_dist_sky_mm_plan_bandwidths": (i32, [vp, i32, vp, vp]),
    "bbm_dist_sky_matmul_f64": (i32, [vp, f64, vp, vp, f64, vp]),
    "bbm_dist_sky_matmul_f32": (i32, [vp, f32, vp, vp, f32, vp]),
}


class DistLayout(C.Structure):
    """``bbm_dist_layout`` of include/bbm_b200.h."""
    _fields_ = [(k, C.c_int64) for k in ("K0", "K1", "J0", "J1", "Jlo", "Jhi", "Ka", "Kb", "m_own", "n_own", "n_local",
                                         "col_offset", "own_offset", "row_offset", "nsend", "nrecv")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}

_lib = None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lib/libbbm_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise BBMError("building libbbm_b200.so failed")
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BBMError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("BBM_B200_LIB") and not hasattr(lib, name):
                continue  # A/B runs against an older build: symbols added since are simply unavailable
            fn = getattr(lib, name)  # AttributeError if the header promises something the .so lacks
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, handle=None, what: str = ""):
    if rc == 0:
        return
    lib = load()
    msg = lib.bbm_status_string(rc).decode()
    if handle is not None:
        detail = lib.bbm_last_error(handle).decode()
        if detail:
            msg = f"{msg}: {detail}"
    if what:
        msg = f"{what}: {msg}"
    if rc == 2:
        raise DimensionMismatch(msg)
    if rc == 3:
        raise BandError(msg)
    raise BBMError(f"[status {rc}] {msg}")
