"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product package icebin_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_gcm_new.restype = C.c_void_p
        L.orc_gcm_new.argtypes = [C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_gcm_free.argtypes = [C.c_void_p]
        L.orc_add_sheet.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_matrix.restype = C.c_void_p
        L.orc_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p,
                                 C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_w_free.argtypes = [C.c_void_p]
        L.orc_w_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_w_coo.argtypes = [C.c_void_p] * 4
        L.orc_w_weights.argtypes = [C.c_void_p] * 3
        L.orc_w_dim.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_w_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_multivec_to_dense.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_long,
                                            C.c_double, C.c_void_p, C.c_void_p]
        L.orc_vartrans.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.orc_e1ve0c.restype = C.c_long
        L.orc_e1ve0c.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_exgrid.restype = C.c_long
        L.orc_exgrid.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p]
        L.orc_filter_cellsA.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_nX.restype = C.c_long
        L.orc_nX.argtypes = [C.c_void_p, C.c_int]
        L.orc_combine_chunks.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.orc_linterp.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_nearest.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_hntr_new.restype = C.c_void_p
        L.orc_hntr_new.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                   C.c_double, C.c_double]
        L.orc_hntr_free.argtypes = [C.c_void_p]
        L.orc_hntr_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_hntr_regrid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_double, C.c_double]
        L.orc_hntr_matrix.restype = C.c_long
        L.orc_hntr_matrix.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    pass


def _err():
    return OracleError(lib().orc_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Weighted:
    """Result of one matrix(spec): dense-indexed COO (CSC order) + weights + dims (linear::Weighted_Eigen)."""

    def __init__(self, h):
        L = lib()
        self._h = h
        info = np.zeros(6, np.int64)
        L.orc_w_info(h, _p(info))
        self.shape = (int(info[0]), int(info[1]))
        self.nnz = int(info[2])
        self.conservative = bool(info[3])
        self.sparse_extent = (int(info[4]), int(info[5]))
        self.row = np.zeros(self.nnz, np.int32)
        self.col = np.zeros(self.nnz, np.int32)
        self.val = np.zeros(self.nnz, np.float64)
        L.orc_w_coo(h, _p(self.row), _p(self.col), _p(self.val))
        self.wM = np.zeros(self.shape[0])
        self.Mw = np.zeros(self.shape[1])
        L.orc_w_weights(h, _p(self.wM), _p(self.Mw))
        self.dims = [np.zeros(self.shape[0], np.int64), np.zeros(self.shape[1], np.int64)]
        L.orc_w_dim(h, 0, _p(self.dims[0]))
        L.orc_w_dim(h, 1, _p(self.dims[1]))

    def apply(self, X, force_conservation=True):
        """X: [nvar, ncol] dense-indexed, field-major.  Returns [nvar, nrow]."""
        X = _c(np.atleast_2d(X), np.float64)
        assert X.shape[1] == self.shape[1]
        Y = np.zeros((X.shape[0], self.shape[0]))
        if lib().orc_w_apply(self._h, _p(X), X.shape[0], _p(Y), int(force_conservation)) != 0:
            raise _err()
        return Y

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_w_free(self._h)
            self._h = None


class Gcm:
    """Oracle twin of GCMRegridder_Standard holding grids as flat arrays (see icebin_b200/synth.py)."""

    def __init__(self, nA, dimA, native_areaA, hcdefs):
        dimA = _c(dimA, np.int64); native_areaA = _c(native_areaA, np.float64); hcdefs = _c(hcdefs, np.float64)
        self._h = lib().orc_gcm_new(nA, len(dimA), _p(dimA), _p(native_areaA), len(hcdefs), _p(hcdefs))
        if not self._h:
            raise _err()
        self._keep = []

    def add_sheet(self, nI, dimI, native_areaI, centroid_xy, ex_ij, ex_area, gridA_proj_area, interp_style=0):
        dimI = _c(dimI, np.int64); native_areaI = _c(native_areaI, np.float64)
        centroid_xy = None if centroid_xy is None else _c(centroid_xy, np.float64)
        ex_ij = _c(ex_ij, np.int32); ex_area = _c(ex_area, np.float64); pa = _c(gridA_proj_area, np.float64)
        r = lib().orc_add_sheet(self._h, nI, len(dimI), _p(dimI), _p(native_areaI), _p(centroid_xy),
                                len(ex_area), _p(ex_ij), _p(ex_area), _p(pa), interp_style)
        if r < 0:
            raise _err()
        return r

    def filter_cellsA(self, keepA):
        """GCMRegridder_Standard::filter_cellsA (GCMRegridder.cpp:154-164).  keepA: bool [nA]."""
        k = _c(np.asarray(keepA) != 0, np.uint8)
        if lib().orc_filter_cellsA(self._h, _p(k)) != 0:
            raise _err()

    def nX(self, sheet):
        return int(lib().orc_nX(self._h, sheet))

    def matrix(self, sheet, elevmaskI, spec, scale=True, correctA=False, sigma=(0., 0., 0.), dims=(None, None)):
        em = _c(elevmaskI, np.float64)
        sg = _c(sigma, np.float64)
        d0 = None if dims[0] is None else _c(dims[0], np.int64)
        d1 = None if dims[1] is None else _c(dims[1], np.int64)
        h = lib().orc_matrix(self._h, sheet, _p(em), spec.encode(), int(scale), int(correctA), _p(sg),
                             -1 if d0 is None else len(d0), _p(d0), -1 if d1 is None else len(d1), _p(d1))
        if not h:
            msg = lib().orc_last_error().decode()
            if msg.startswith("map::at"):
                raise KeyError(spec)
            raise OracleError(msg)
        return Weighted(h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_gcm_free(self._h)
            self._h = None


def vartrans(O, M, b):
    """Z = O * M + b (IceCoupler.cpp:203-252, :404-446).  O: [ncell, nin]; M: scipy sparse [nin, nout]; b: [nout]."""
    import scipy.sparse
    O = _c(np.atleast_2d(O), np.float64)
    Mc = scipy.sparse.csc_matrix(M)
    Mc.sort_indices()
    assert Mc.shape[0] == O.shape[1]
    cp = _c(Mc.indptr, np.int32); row = _c(Mc.indices, np.int32); val = _c(Mc.data, np.float64); b = _c(b, np.float64)
    Z = np.zeros((O.shape[0], Mc.shape[1]))
    if lib().orc_vartrans(_p(O), O.shape[0], O.shape[1], Mc.shape[1], _p(cp), _p(row), _p(val), _p(b), _p(Z)) != 0:
        raise _err()
    return Z


def e1ve0c(XuE1s, XuE0s, nE):
    """e1ve0.cpp:55-106.  XuE*: lists of oracle Weighted (XvE, scale=False, correctA=True, rows = all exchange cells).
    Returns (iE1, iE0, val) sorted by (iE1, iE0), sparse E indices."""
    n = len(XuE1s)
    a = (C.c_void_p * n)(*[w._h for w in XuE1s]); b = (C.c_void_p * n)(*[w._h for w in XuE0s])
    cnt = lib().orc_e1ve0c(n, a, b, nE, None, None, None)
    if cnt < 0:
        raise _err()
    r = np.zeros(cnt, np.int64); c = np.zeros(cnt, np.int64); v = np.zeros(cnt)
    lib().orc_e1ve0c(n, a, b, nE, _p(r), _p(c), _p(v))
    return r, c, v


def exgrid_rectilinear(polys, iA, xe, ye, iI_major_x=True):
    """gridgen/GridGen_Exchange.cpp:175-316 for projected GCM polygons x a rectilinear ice grid.  Returns (ex_ij [n,2], ex_area [n])."""
    pptr = np.zeros(len(polys) + 1, np.int32)
    np.cumsum([len(p) for p in polys], out=pptr[1:])
    pxy = _c(np.concatenate([np.asarray(p, np.float64).reshape(-1, 2) for p in polys]), np.float64)
    iA = _c(iA, np.int64); xe = _c(xe, np.float64); ye = _c(ye, np.float64)
    args = (len(polys), _p(pptr), _p(pxy), _p(iA), len(xe) - 1, _p(xe), len(ye) - 1, _p(ye), int(iI_major_x))
    n = lib().orc_exgrid(*args, None, None)
    if n < 0:
        raise _err()
    ij = np.zeros((n, 2), np.int32); area = np.zeros(n)
    lib().orc_exgrid(*args, _p(ij), _p(area))
    return ij, area


def combine_chunks(chunks, scale=True):
    """modele/combine_global_ec.cpp:95-262.  chunks: oracle Weighted built with scale=False.  Returns (coo M, (wM idx, val), (Mw idx, val))."""
    import scipy.sparse
    n = len(chunks)
    a = (C.c_void_p * n)(*[w._h for w in chunks])
    nnz = sum(w.nnz for w in chunks); nr = sum(w.shape[0] for w in chunks); nc = sum(w.shape[1] for w in chunks)
    r = np.zeros(nnz, np.int64); c = np.zeros(nnz, np.int64); v = np.zeros(nnz)
    wi = np.zeros(nr, np.int64); wv = np.zeros(nr); mi = np.zeros(nc, np.int64); mv = np.zeros(nc)
    if lib().orc_combine_chunks(n, a, int(scale), _p(r), _p(c), _p(v), _p(wi), _p(wv), _p(mi), _p(mv)) != 0:
        raise _err()
    return scipy.sparse.coo_matrix((v, (r, c)), shape=chunks[0].sparse_extent), (wi, wv), (mi, mv)


def linterp(xp, xx):
    xp = _c(xp, np.float64)
    idx = np.zeros(2, np.int64); w = np.zeros(2)
    if lib().orc_linterp(_p(xp), len(xp), float(xx), _p(idx), _p(w)) != 0:
        raise _err()
    return idx, w


def nearest(xp, xx):
    xp = _c(xp, np.float64)
    return lib().orc_nearest(_p(xp), len(xp), float(xx))


def multivec_to_dense(index, weights, vals, nE, fill=np.nan):
    index = _c(index, np.int64); weights = _c(weights, np.float64); vals = _c(vals, np.float64)
    n, nvar = vals.shape
    scale = np.zeros(nE); dense = np.zeros((nvar, nE))
    if lib().orc_multivec_to_dense(n, _p(index), _p(weights), _p(vals), nvar, nE, fill, _p(scale), _p(dense)) != 0:
        raise _err()
    return scale, dense


class Hntr:
    """Oracle twin of icebin::modele::Hntr(yp17, Bspec, Aspec, DATMIS).  spec = (im, jm, offi, dlat_minutes)."""
    _NAMES = {"SINA": 0, "SINB": 1, "FMIN": 2, "FMAX": 3, "GMIN": 4, "GMAX": 5, "IMIN": 6, "IMAX": 7, "JMIN": 8,
              "JMAX": 9, "dxypA": 10, "dxypB": 11}

    def __init__(self, specB, specA, datmis=0.0):
        self.specB, self.specA = specB, specA
        self._h = lib().orc_hntr_new(specB[0], specB[1], specB[2], specB[3], specA[0], specA[1], specA[2], specA[3],
                                     datmis)
        if not self._h:
            raise _err()

    def get(self, name):
        w = self._NAMES[name]
        n = {0: self.specA[1] + 1, 1: self.specB[1] + 1, 2: self.specB[0], 3: self.specB[0], 4: self.specB[1],
             5: self.specB[1], 6: self.specB[0], 7: self.specB[0], 8: self.specB[1], 9: self.specB[1],
             10: self.specA[1], 11: self.specB[1]}[w]
        if 6 <= w < 10:
            out = np.zeros(n, np.int32)
            lib().orc_hntr_get(self._h, w, None, _p(out))
        else:
            out = np.zeros(n)
            lib().orc_hntr_get(self._h, w, _p(out), None)
        return out

    def regrid(self, WTA, A, mean_polar=False, wtm=1.0, wtb=0.0):
        """A: flat [jmA*imA] (i fastest), float64 or int16.  WTA: same shape float64 or None (constant 1)."""
        nA = self.specA[0] * self.specA[1]
        A = np.ascontiguousarray(A).reshape(-1)
        assert A.size == nA
        if A.dtype == np.int16:
            dt = 1
        else:
            A = _c(A, np.float64); dt = 0
        W = None if WTA is None else _c(np.asarray(WTA).reshape(-1), np.float64)
        B = np.zeros(self.specB[0] * self.specB[1])
        if lib().orc_hntr_regrid(self._h, _p(W), _p(A), dt, _p(B), int(mean_polar), wtm, wtb) != 0:
            raise _err()
        return B

    def matrix(self, R=None):
        """overlap(R) if R is given, else scaled_regrid_matrix.  Returns (rows, cols, vals) in emission order."""
        R2 = -1.0 if R is None else float(R) * float(R)
        n = lib().orc_hntr_matrix(self._h, R2, None, None, None)
        r = np.zeros(n, np.int64); c = np.zeros(n, np.int64); v = np.zeros(n)
        lib().orc_hntr_matrix(self._h, R2, _p(r), _p(c), _p(v))
        return r, c, v

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_hntr_free(self._h)
            self._h = None
