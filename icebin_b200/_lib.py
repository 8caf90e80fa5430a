"""ctypes loader for libicebin_b200.so (the C ABI in include/icebin_b200.h).

The library is built in-tree by icebin_b200/build.py.  There is no fallback: if the shared object is missing the
import of any compute entry point raises, and on a machine without a CUDA device every compute call returns
IB_ERR_CUDA (raised here as IceBinError).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ICEBIN_B200_LIB: an instrumented build of the same sources (tools/build_trace.py); the product always loads the in-tree library
SO_PATH = os.environ.get("ICEBIN_B200_LIB") or os.path.join(_HERE, "libicebin_b200.so")

i64, i32, f64, vp, cp = C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_char_p
PP = C.POINTER(vp)

# name -> (restype, argtypes): every symbol include/icebin_b200.h declares
SIGNATURES = {
    "ib_last_error": (cp, []),
    "ib_device_count": (C.c_int, []),
    "ib_set_device": (C.c_int, [C.c_int]),
    "ib_launch_count": (i64, []),
    "ib_tune": (C.c_int, [cp, C.c_int]),
    "ib_trim": (C.c_int, []),
    "ib_dev_alloc": (C.c_int, [PP, i64]),
    "ib_dev_free": (None, [vp]),
    "ib_dev_upload": (C.c_int, [vp, vp, i64]),
    "ib_dev_download": (C.c_int, [vp, vp, i64]),
    "ib_gcm_create": (C.c_int, [PP, i64, i32, vp, vp, i32, vp, C.c_int]),
    "ib_gcm_free": (None, [vp]),
    "ib_gcm_add_sheet": (C.c_int, [vp, cp, i64, i32, vp, vp, vp, i64, vp, vp, vp, C.c_int]),
    "ib_gcm_filter_cellsA": (C.c_int, [vp, vp]),
    "ib_gcm_nA_realized": (i64, [vp]),
    "ib_gcm_nA": (i64, [vp]),
    "ib_gcm_nE": (i64, [vp]),
    "ib_gcm_nhc": (i32, [vp]),
    "ib_gcm_nI": (i64, [vp, C.c_int]),
    "ib_gcm_nX": (i64, [vp, C.c_int]),
    "ib_gcm_nsheets": (C.c_int, [vp]),
    "ib_gcm_sheet_index": (C.c_int, [vp, cp]),
    "ib_gcm_correctA": (C.c_int, [vp]),
    "ib_gcm_wA": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "ib_regrid_matrices": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, vp, PP]),
    "ib_regrid_matrices_dev": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, PP]),
    "ib_rm_free": (None, [vp]),
    "ib_dim_create": (C.c_int, [PP]),
    "ib_dim_create_ids": (C.c_int, [PP, i64, i64, vp]),
    "ib_dim_create_identity": (C.c_int, [PP, i64]),
    "ib_dim_dense_extent": (i64, [vp]),
    "ib_dim_sparse_extent": (i64, [vp]),
    "ib_dim_get_ids": (C.c_int, [vp, vp]),
    "ib_dim_free": (None, [vp]),
    "ib_rm_matrix": (C.c_int, [vp, cp, PP]),
    "ib_rm_matrix_d": (C.c_int, [vp, cp, vp, vp, C.c_int, C.c_int, vp, PP]),
    "ib_w_free": (None, [vp]),
    "ib_w_create_csr": (C.c_int, [PP, i64, i64, vp, vp, vp, vp, vp, C.c_int]),
    "ib_w_aggregate": (C.c_int, [vp, C.c_int, vp, vp, i64, vp, PP]),
    "ib_w_set_weights": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "ib_w_set_dim": (C.c_int, [vp, C.c_int, i64, i64, vp]),
    "ib_w_info": (C.c_int, [vp, vp]),
    "ib_w_get_csr": (C.c_int, [vp, vp, vp, vp]),
    "ib_w_get_coo": (C.c_int, [vp, vp, vp, vp]),
    "ib_w_get_weights": (C.c_int, [vp, vp, vp]),
    "ib_w_get_dim": (C.c_int, [vp, C.c_int, vp]),
    "ib_w_dev_ptrs": (C.c_int, [vp, PP, PP, PP, PP, PP]),
    "ib_w_apply": (C.c_int, [vp, vp, i64, i32, vp, i64, C.c_int, C.c_int]),
    "ib_w_apply_dev": (C.c_int, [vp, vp, i64, i32, vp, i64, C.c_int, vp]),
    "ib_w_apply_many_dev": (C.c_int, [i32, vp, vp, i64, i32, vp, vp, C.c_int, vp]),
    "ib_w_apply_many": (C.c_int, [i32, vp, vp, i64, i32, vp, vp, C.c_int, C.c_int]),
    "ib_w_apply_rows_dev": (C.c_int, [vp, vp, i64, i32, vp, i64, i64, i64, vp]),
    "ib_w_apply_M": (C.c_int, [vp, vp, i32, vp, f64, C.c_int]),
    "ib_w_apply_weight": (C.c_int, [vp, C.c_int, vp, i32, vp]),
    "ib_coo_matvec": (C.c_int, [vp, vp, C.c_int, i64, i64, i64, vp, vp, vp]),
    "ib_multivec_to_dense": (C.c_int, [i64, vp, vp, vp, i32, i64, f64, vp, vp]),
    "ib_vartrans_dev": (C.c_int, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, i64, vp]),
    "ib_e1ve0c": (i64, [i32, vp, vp, i64, vp, vp, vp]),
    "ib_w_combine_chunks": (C.c_int, [i32, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "ib_exgrid_rectilinear": (i64, [i32, vp, vp, vp, i32, vp, i32, vp, C.c_int, vp, vp]),
    "ib_xchg_plan": (i64, [C.c_int, vp, vp, i64, C.c_int, vp, vp, vp]),
    "ib_xchg_create": (C.c_int, [PP, C.c_int, C.c_int, C.c_int, vp, i32]),
    "ib_xchg_free": (None, [vp]),
    "ib_xchg_handle_bytes": (C.c_int, []),
    "ib_xchg_get_handle": (C.c_int, [vp, vp]),
    "ib_xchg_open": (C.c_int, [vp, vp]),
    "ib_xchg_open_local": (C.c_int, [vp, vp]),
    "ib_xchg_set_part": (C.c_int, [vp, C.c_int, i64, vp, vp, vp, vp, C.c_int]),
    "ib_xchg_merge_dev": (C.c_int, [vp, vp, vp, i32, vp]),
    "ib_xchg_publish_dev": (C.c_int, [vp, vp, vp, i32, vp]),
    "ib_xchg_gather_dev": (C.c_int, [vp, vp, vp, i32, vp]),
    "ib_xchg_status": (C.c_int, [vp]),
    "ib_xchg_step": (i64, [vp]),
    "ib_merge_pack_dev": (C.c_int, [vp, i64, i32, vp, i64, vp, vp]),
    "ib_merge_unpack_dev": (C.c_int, [vp, i64, i32, i64, vp, vp, vp, vp, vp]),
    "ib_hntr_create": (C.c_int, [PP, C.c_int, C.c_int, f64, f64, C.c_int, C.c_int, f64, f64, f64]),
    "ib_hntr_free": (None, [vp]),
    "ib_hntr_get": (C.c_int, [vp, C.c_int, vp, vp]),
    "ib_hntr_regrid": (C.c_int, [vp, vp, f64, f64, vp, C.c_int, vp, C.c_int]),
    "ib_hntr_regrid_dev": (C.c_int, [vp, vp, f64, f64, vp, C.c_int, vp, C.c_int, vp]),
    "ib_hntr_matrix": (i64, [vp, f64, vp, vp, vp]),
}

_lib = None


class IceBinError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference throws through icebin_error)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "libicebin_b200.so is not built (run `python -m icebin_b200.build`); "
                "icebin_b200 has no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)           # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status < 0:
        raise IceBinError(status, lib().ib_last_error().decode(errors="replace"))
    return status
