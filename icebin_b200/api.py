"""Python face of the backend, mirroring the reference's Cython module pylib/_icebin.pyx:50-228
(GCMRegridder, RegridMatrices, HntrSpec, Hntr) and ibmisc.linear_Weighted (sphinx/source/matrix_formats.rst:121-195).

Differences from the reference binding, all upstream of the hot path:
  * grids arrive as arrays (icebin_b200.synth.GridA / SheetGrid) instead of NetCDF file names -- netCDF, PROJ and
    CGAL are out of scope;
  * `matrix_d` and `SparseSet` expose the C++ RegridMatrices_Dynamic::matrix_d path that the coupler uses.
Everything computes on the GPU through the C ABI; numpy is only the host container.
"""
import ctypes as C

import numpy as np

from ._lib import IceBinError, check, lib, vp

__all__ = ["GCMRegridder", "RegridMatrices", "Weighted", "SparseSet", "HntrSpec", "Hntr", "IceBinError",
           "coo_matvec", "multivec_to_dense", "device_count", "launch_count", "apply_many", "apply_many_dev",
           "vartrans_dev", "vartrans", "e1ve0c", "exgrid_rectilinear", "combine_chunks"]


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def device_count():
    return lib().ib_device_count()


def tune(key, value):
    """Performance knobs of the library (ib_tune); results never depend on them."""
    check(lib().ib_tune(key.encode(), int(value)))


def trim():
    """Give cached device memory back to the driver (ib_trim)."""
    check(lib().ib_trim())


def launch_count():
    return int(lib().ib_launch_count())


class SparseSet:
    """SparseSetT (spsparse::SparseSet<long,int>): ordered sparse<->dense index bijection, resident on the GPU."""

    def __init__(self, ids=None, sparse_extent=-1, identity=None):
        h = vp()
        if identity is not None:
            check(lib().ib_dim_create_identity(C.byref(h), int(identity)))
        elif ids is None:
            check(lib().ib_dim_create(C.byref(h)))
        else:
            ids = _c(ids, np.int64)
            check(lib().ib_dim_create_ids(C.byref(h), int(sparse_extent), len(ids), _p(ids)))
        self._h = h

    @property
    def dense_extent(self):
        return int(lib().ib_dim_dense_extent(self._h))

    @property
    def sparse_extent(self):
        return int(lib().ib_dim_sparse_extent(self._h))

    def ids(self):
        out = np.zeros(self.dense_extent, np.int64)
        check(lib().ib_dim_get_ids(self._h, _p(out)))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ib_dim_free(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the loader module may already be gone
            pass


class Weighted:
    """ibmisc.linear_Weighted: regrid matrix M (dense-indexed) + weight vectors wM (rows) and Mw (columns)."""

    def __init__(self, h, keep=()):
        self._h = h
        self._keep = keep      # borrowed dims must outlive the matrix
        info = np.zeros(8, np.int64)
        check(lib().ib_w_info(h, _p(info)))
        self.shape = (int(info[0]), int(info[1]))
        self.nnz = int(info[2])
        self.conservative = bool(info[3])
        self.scaled = bool(info[4])
        self.sparse_extent = (int(info[5]), int(info[6]))

    @classmethod
    def from_csr(cls, nrow, ncol, rowptr, col, val, wM=None, Mw=None, conservative=True):
        """Wrap a dense-indexed CSR matrix (extension; e.g. matrices read back from a Weighted_Eigen NetCDF dump)."""
        rowptr = _c(rowptr, np.int32); col = _c(col, np.int32); val = _c(val, np.float64)
        wM = None if wM is None else _c(wM, np.float64)
        Mw = None if Mw is None else _c(Mw, np.float64)
        h = vp()
        check(lib().ib_w_create_csr(C.byref(h), nrow, ncol, _p(rowptr), _p(col), _p(val), _p(wM), _p(Mw), int(conservative)))
        return cls(h)

    # ---- extraction
    def to_csr(self):
        rp = np.zeros(self.shape[0] + 1, np.int32)
        col = np.zeros(self.nnz, np.int32)
        val = np.zeros(self.nnz)
        check(lib().ib_w_get_csr(self._h, _p(rp), _p(col), _p(val)))
        return rp, col, val

    def to_coo_dense(self):
        """(row, col, val) with DENSE indices, sorted by (row, col)."""
        row = np.zeros(self.nnz, np.int32)
        col = np.zeros(self.nnz, np.int32)
        val = np.zeros(self.nnz)
        check(lib().ib_w_get_coo(self._h, _p(row), _p(col), _p(val)))
        return row, col, val

    def to_coo(self):
        """scipy.sparse.coo_matrix in the ORIGINAL sparse index spaces (matrix_formats.rst:188-192)."""
        import scipy.sparse
        row, col, val = self.to_coo_dense()
        d0, d1 = self.dim(0), self.dim(1)
        return scipy.sparse.coo_matrix((val, (d0[row], d1[col])), shape=self.sparse_extent)

    def get_weights(self, dim):
        """Dense-indexed weight vector: 0 = wM (output space B), 1 = Mw (input space A)."""
        wM = np.zeros(self.shape[0])
        Mw = np.zeros(self.shape[1])
        check(lib().ib_w_get_weights(self._h, _p(wM), _p(Mw)))
        return wM if dim == 0 else Mw

    def dim(self, which):
        out = np.zeros(self.shape[which], np.int64)
        check(lib().ib_w_get_dim(self._h, which, _p(out)))
        return out

    # ---- application
    def apply(self, X, force_conservation=True, layout="field"):
        """Dense-indexed apply (Weighted_Eigen::apply).  layout='field': X is [nvar, ncol] as in the reference
        (IceCoupler.cpp:231-233) and the result is [nvar, nrow]; layout='cell': X is [ncol, nvar] -> [nrow, nvar]."""
        X = _c(np.atleast_2d(X), np.float64)
        if layout == "field":
            nvar = X.shape[0]
            assert X.shape[1] == self.shape[1]
            Y = np.zeros((nvar, self.shape[0]))
            check(lib().ib_w_apply(self._h, _p(X), X.shape[1], nvar, _p(Y), max(1, Y.shape[1]), 1, int(force_conservation)))
        else:
            nvar = X.shape[1]
            assert X.shape[0] == self.shape[1]
            Y = np.zeros((self.shape[0], nvar))
            check(lib().ib_w_apply(self._h, _p(X), nvar, nvar, _p(Y), nvar, 0, int(force_conservation)))
        return Y

    def apply_dev(self, dX_ptr, ldx, nvar, dY_ptr, ldy, force_conservation=True, stream=0):
        """Device-resident cell-major apply; pointers are raw CUDA addresses (e.g. torch.Tensor.data_ptr())."""
        check(lib().ib_w_apply_dev(self._h, vp(dX_ptr), ldx, nvar, vp(dY_ptr), ldy, int(force_conservation),
                                   vp(stream) if stream else None))

    def apply_rows_dev(self, dX_ptr, ldx, nvar, dY_ptr, ldy, row_begin, row_end, stream=0):
        """Row-block shard of apply_dev: only output rows [row_begin, row_end) (multiples of 32) are computed."""
        check(lib().ib_w_apply_rows_dev(self._h, vp(dX_ptr), ldx, nvar, vp(dY_ptr), ldy, row_begin, row_end,
                                        vp(stream) if stream else None))

    def apply_M(self, A_s, fill=np.nan, force_conservation=True):
        """Sparse-indexed apply (matrix_formats.rst:150-165).  A_s: [sparse_extent(A)] or [nvar, sparse_extent(A)]."""
        A = _c(A_s, np.float64)
        one = A.ndim == 1
        A = np.atleast_2d(A)
        if A.shape[1] != self.sparse_extent[1]:
            raise ValueError("apply_M: expected trailing dimension %d, got %r" % (self.sparse_extent[1], A.shape))
        B = np.zeros((A.shape[0], self.sparse_extent[0]))
        check(lib().ib_w_apply_M(self._h, _p(A), A.shape[0], _p(B), float(fill), int(force_conservation)))
        return B[0] if one else B

    def apply_weight(self, dim, A_s):
        A = _c(A_s, np.float64)
        one = A.ndim == 1
        A = np.atleast_2d(A)
        out = np.zeros(A.shape[0])
        check(lib().ib_w_apply_weight(self._h, dim, _p(A), A.shape[0], _p(out)))
        return out[0] if one else out

    def apply_wM(self, A_s):
        return self.apply_weight(0, A_s)

    def apply_Mw(self, A_s):
        return self.apply_weight(1, A_s)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ib_w_free(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the loader module may already be gone
            pass


def apply_many_dev(mats, dX_ptr, ldx, nvar, dY_ptrs, ldys, force_conservation=True, stream=0):
    """Apply several matrices to the SAME device-resident input fields concurrently (IceCoupler.cpp:445-446 applies
    EvI and AvI to one ice_ovalsI).  dY_ptrs[i] receives mats[i] @ X."""
    n = len(mats)
    Wp = (vp * n)(*[m._h for m in mats])
    Yp = (vp * n)(*[vp(p) for p in dY_ptrs])
    ld = (C.c_int64 * n)(*ldys)
    check(lib().ib_w_apply_many_dev(n, Wp, vp(dX_ptr), ldx, nvar, Yp, ld, int(force_conservation), vp(stream) if stream else None))


def apply_many(mats, X, force_conservation=True, layout="cell"):
    """Host-buffer form: X is uploaded once.  layout='cell': X [ncol, nvar] -> list of [nrow_i, nvar];
    layout='field': X [nvar, ncol] -> list of [nvar, nrow_i]."""
    X = _c(np.atleast_2d(X), np.float64)
    n = len(mats)
    if layout == "cell":
        nvar = X.shape[1]
        Ys = [np.zeros((m.shape[0], nvar)) for m in mats]
        ld = (C.c_int64 * n)(*[nvar] * n)
        ldx, lay = nvar, 0
    else:
        nvar = X.shape[0]
        Ys = [np.zeros((nvar, m.shape[0])) for m in mats]
        ld = (C.c_int64 * n)(*[max(1, m.shape[0]) for m in mats])
        ldx, lay = X.shape[1], 1
    Wp = (vp * n)(*[m._h for m in mats])
    Yp = (vp * n)(*[y.ctypes.data_as(vp) for y in Ys])
    check(lib().ib_w_apply_many(n, Wp, _p(X), ldx, nvar, Yp, ld, lay, int(force_conservation)))
    return Ys


class RegridMatrices:
    """RegridMatrices_Dynamic for one sheet + elevmaskI (pylib/_icebin.pyx:50-75)."""

    SPECS = ("AvI", "IvA", "AvX", "XvA", "EvI", "IvE", "EvX", "XvE", "EvA", "AvE")

    def __init__(self, h, gcm):
        self._h = h
        self._gcm = gcm        # closures hold a raw IceRegridder* upstream; keep the owner alive here

    def matrix(self, spec_name):
        h = vp()
        st = lib().ib_rm_matrix(self._h, spec_name.encode(), C.byref(h))
        if st < 0 and spec_name not in self.SPECS:
            raise KeyError(spec_name)
        check(st)
        return Weighted(h)

    def matrix_d(self, spec_name, dims, scale=True, correctA=False, sigma=(0., 0., 0.)):
        """RegridMatrices_Dynamic::matrix_d(spec, {dim0, dim1}, RegridParams): appends to the caller's SparseSets."""
        h = vp()
        sg = _c(sigma, np.float64)
        st = lib().ib_rm_matrix_d(self._h, spec_name.encode(), dims[0]._h, dims[1]._h, int(scale), int(correctA), _p(sg),
                                  C.byref(h))
        if st < 0 and spec_name not in self.SPECS:
            raise KeyError(spec_name)
        check(st)
        return Weighted(h, keep=tuple(dims))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ib_rm_free(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the loader module may already be gone
            pass


class GCMRegridder:
    """GCMRegridder_Standard (pylib/_icebin.pyx:77-175; GCMRegridder.hpp:312-400)."""

    INTERP = {"Z_INTERP": 0, "ELEV_CLASS_INTERP": 1, 0: 0, 1: 1}

    def __init__(self, gridA, hcdefs, correctA=True):
        dimA = _c(gridA.dim, np.int64)
        area = _c(gridA.native_area, np.float64)
        hc = _c(hcdefs, np.float64)
        h = vp()
        check(lib().ib_gcm_create(C.byref(h), int(gridA.nA), len(dimA), _p(dimA), _p(area), len(hc), _p(hc), int(correctA)))
        self._h = h
        self._n_real = len(dimA)
        self.sheets = {}

    @property
    def nA(self):
        return int(lib().ib_gcm_nA(self._h))

    @property
    def nE(self):
        return int(lib().ib_gcm_nE(self._h))

    @property
    def nhc(self):
        return int(lib().ib_gcm_nhc(self._h))

    def nI(self, sheet_name):
        return int(lib().ib_gcm_nI(self._h, self.sheets[sheet_name]))

    def nX(self, sheet_name):
        return int(lib().ib_gcm_nX(self._h, self.sheets[sheet_name]))

    def filter_cellsA(self, keepA):
        """GCMRegridder_Standard::filter_cellsA(keepA) (GCMRegridder.hpp:380-389): restrict every sheet's exchange grid and
        the GCM grid's realised set to the kept GCM cells (a rank's / a GPU's domain).  keepA: bool [nA] or a predicate on iA."""
        if callable(keepA):
            keepA = np.array([bool(keepA(i)) for i in range(self.nA)])
        k = _c(np.asarray(keepA) != 0, np.uint8)
        if k.shape != (self.nA,):
            raise ValueError("keepA must have nA=%d elements" % self.nA)
        check(lib().ib_gcm_filter_cellsA(self._h, _p(k)))
        self._n_real = int(lib().ib_gcm_nA_realized(self._h))

    def add_sheet(self, name, sheet, interp_style="Z_INTERP"):
        dimI = _c(sheet.dimI, np.int64)
        areaI = _c(sheet.native_areaI, np.float64)
        cen = None if sheet.centroid_xy is None else _c(sheet.centroid_xy, np.float64)
        ex_ij = _c(sheet.ex_ij, np.int32)
        ex_area = _c(sheet.ex_area, np.float64)
        proj = _c(sheet.gridA_proj_area, np.float64)
        if len(proj) != self._n_real:
            raise ValueError("gridA_proj_area must have one entry per realised A cell")
        ix = check(lib().ib_gcm_add_sheet(self._h, name.encode(), int(sheet.nI), len(dimI), _p(dimI), _p(areaI), _p(cen),
                                          len(ex_area), _p(ex_ij), _p(ex_area), _p(proj), self.INTERP[interp_style]))
        self.sheets[name] = ix
        return ix

    def wA(self, sheet_name, snative, fill=0.):
        if snative not in ("native", "proj"):
            raise ValueError("Invalid argument: snative must be 'native' or 'proj'")
        idx = np.zeros(self._n_real, np.int64)
        area = np.zeros(self._n_real)
        check(lib().ib_gcm_wA(self._h, self.sheets[sheet_name], int(snative == "native"), _p(idx), _p(area)))
        out = np.full(self.nA, float(fill))
        out[idx] = area           # accum::blitz_existing(..., DuplicatePolicy::REPLACE), icebin_cython.cpp:122-123
        return out

    def regrid_matrices(self, sheet_name, elevmaskI, scale=True, correctA=True, sigma=(0, 0, 0), conserve=True):
        """pylib/_icebin.pyx:164-175 (`conserve` is accepted and ignored, exactly as upstream: icebin_cython.cpp:215-236)."""
        if sheet_name not in self.sheets:
            raise KeyError(sheet_name)        # ice_regridders().index.at(name)
        em = _c(np.asarray(elevmaskI).reshape(-1), np.float64)
        ix = self.sheets[sheet_name]
        if len(em) != lib().ib_gcm_nI(self._h, ix):
            raise ValueError("elevmaskI must have nI=%d elements" % lib().ib_gcm_nI(self._h, ix))
        sg = _c(sigma, np.float64)
        h = vp()
        check(lib().ib_regrid_matrices(self._h, ix, _p(em), int(scale), int(correctA), _p(sg), C.byref(h)))
        return RegridMatrices(h, self)

    def regrid_matrices_dev(self, sheet_name, d_elevmask_ptr, scale=True, correctA=True, sigma=(0, 0, 0), stream=0):
        """The same with elevmaskI [nI] fp64 already in device memory (the ice model's state stays in HBM): no upload."""
        if sheet_name not in self.sheets:
            raise KeyError(sheet_name)
        sg = _c(sigma, np.float64)
        h = vp()
        check(lib().ib_regrid_matrices_dev(self._h, self.sheets[sheet_name], vp(d_elevmask_ptr), int(scale), int(correctA), _p(sg),
                                           vp(stream) if stream else None, C.byref(h)))
        return RegridMatrices(h, self)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ib_gcm_free(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the loader module may already be gone
            pass


class HntrSpec:
    """pylib/_icebin.pyx:183-211; dlat in MINUTES (GridSpec.hpp:143-170)."""

    def __init__(self, im, jm, offi, dlat):
        self.im, self.jm, self.offi, self.dlat = int(im), int(jm), float(offi), float(dlat)

    @property
    def size(self):
        return self.im * self.jm


# predefined ModelE grids (slib/icebin/modele/grids.cpp:9-18, grids.hpp:9-34)
_dLATM = 180. * 60. / 180
GRIDS = {
    "g1mx1m": HntrSpec(21600, 10800, 0., 1.), "g2mx2m": HntrSpec(10800, 5400, 0., 2.),
    "g10mx10m": HntrSpec(2160, 1080, 0., 10.), "gqxq": HntrSpec(1440, 720, 0., 15.),
    "ghxh": HntrSpec(720, 360, 0., 30.), "g1x1": HntrSpec(360, 180, 0., 60.),
    "g1qx1": HntrSpec(288, 180, 0., _dLATM), "g2hx2": HntrSpec(144, 90, 0., _dLATM * 2.),
    "g5x4": HntrSpec(72, 45, 0., _dLATM * 4.),
}


class Hntr:
    """icebin::modele::Hntr(yp17, Bgrid, Agrid, DATMIS) (pylib/_icebin.pyx:218-228); yp17 is ignored upstream too."""
    _SEL = {"SINA": 0, "SINB": 1, "FMIN": 2, "FMAX": 3, "GMIN": 4, "GMAX": 5, "IMIN": 6, "IMAX": 7, "JMIN": 8, "JMAX": 9,
            "dxypA": 10, "dxypB": 11}

    def __init__(self, yp17, Bgrid, Agrid, DATMIS=0.0):
        self.Bgrid, self.Agrid = Bgrid, Agrid
        h = vp()
        check(lib().ib_hntr_create(C.byref(h), Bgrid.im, Bgrid.jm, Bgrid.offi, Bgrid.dlat, Agrid.im, Agrid.jm, Agrid.offi,
                                   Agrid.dlat, float(DATMIS)))
        self._h = h

    def get(self, name):
        w = self._SEL[name]
        A, B = self.Agrid, self.Bgrid
        n = {0: A.jm + 1, 1: B.jm + 1, 2: B.im, 3: B.im, 4: B.jm, 5: B.jm, 6: B.im, 7: B.im, 8: B.jm, 9: B.jm, 10: A.jm, 11: B.jm}[w]
        if 6 <= w < 10:
            out = np.zeros(n, np.int32)
            check(lib().ib_hntr_get(self._h, w, None, _p(out)))
        else:
            out = np.zeros(n)
            check(lib().ib_hntr_get(self._h, w, _p(out), None))
        return out

    def regrid(self, WTA, A, mean_polar=False, wtm=1.0, wtb=0.0):
        """A: [jmA*imA] (i fastest; any shape with that many elements), float64 or int16.  WTA: float64 or None
        (None = constant weight 1).  Returns B[jmB*imB] float64.  (Hntr_regrid, icebin_cython.cpp:202-213)"""
        A = np.ascontiguousarray(A).reshape(-1)
        if A.size != self.Agrid.size:
            raise IceBinError(-5, "Error in dimensions: (%d) vs. (%d)" % (A.size, self.Agrid.size))
        if A.dtype == np.int16:
            dt = 1
        else:
            A = _c(A, np.float64)
            dt = 0
        W = None
        if WTA is not None:
            W = _c(np.asarray(WTA).reshape(-1), np.float64)
            if W.size != self.Agrid.size:
                raise IceBinError(-5, "Error in dimensions: WTA has %d elements vs. %d" % (W.size, self.Agrid.size))
        B = np.zeros(self.Bgrid.size)
        check(lib().ib_hntr_regrid(self._h, _p(W), float(wtm), float(wtb), _p(A), dt, _p(B), int(mean_polar)))
        return B

    def regrid_dev(self, d_wta, d_src, src_dtype, d_dst, mean_polar=False, wtm=1.0, wtb=0.0, stream=0):
        check(lib().ib_hntr_regrid_dev(self._h, vp(d_wta) if d_wta else None, float(wtm), float(wtb), vp(d_src), src_dtype,
                                       vp(d_dst), int(mean_polar), vp(stream) if stream else None))

    def overlap(self, R=1.0):
        return self._matrix(float(R))

    def scaled_regrid_matrix(self):
        return self._matrix(-1.0)

    def _matrix(self, R):
        n = lib().ib_hntr_matrix(self._h, R, None, None, None)
        check(n)
        r = np.zeros(n, np.int64); c = np.zeros(n, np.int64); v = np.zeros(n)
        check(lib().ib_hntr_matrix(self._h, R, _p(r), _p(c), _p(v)))
        return r, c, v

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ib_hntr_free(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the loader module may already be gone
            pass


def coo_matvec(yy, xx, ignore_nan, nrow, ncol, row, col, data):
    """pylib/icebin_cython.cpp:159-192: yy = M xx for a COO matrix; NaN inputs skipped when ignore_nan."""
    row = _c(row, np.int32); col = _c(col, np.int32); data = _c(data, np.float64); xx = _c(xx, np.float64)
    assert yy.dtype == np.float64 and yy.flags.c_contiguous
    check(lib().ib_coo_matvec(_p(yy), _p(xx), int(ignore_nan), nrow, ncol, len(data), _p(row), _p(col), _p(data)))
    return yy


def _csc_arrays(M):
    import scipy.sparse
    Mc = scipy.sparse.csc_matrix(M)
    Mc.sort_indices()
    return Mc.shape, _c(Mc.indptr, np.int32), _c(Mc.indices, np.int32), _c(Mc.data, np.float64)


def vartrans_dev(dO_ptr, ldo, ncell, M, b, dZ_ptr, ldz, stream=0):
    """Z = O * M + b on device-resident cell-major fields (IceCoupler.cpp:203-252, :404-446).  M: [nin, nout] sparse
    (the transposed VarTransformer matrix), b: [nout]."""
    (nin, nout), cp_, row, val = _csc_arrays(M)
    b = _c(b, np.float64)
    assert b.shape == (nout,)
    check(lib().ib_vartrans_dev(vp(dO_ptr), ldo, ncell, nin, nout, _p(cp_), _p(row), _p(val), _p(b), vp(dZ_ptr), ldz,
                                vp(stream) if stream else None))


def vartrans(O, M, b):
    """Host-array convenience around vartrans_dev (torch is only the device container).  O: [ncell, nin] -> [ncell, nout]."""
    import torch
    O = _c(np.atleast_2d(O), np.float64)
    nout = M.shape[1]
    dO = torch.from_numpy(O).cuda()
    dZ = torch.empty((O.shape[0], nout), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    vartrans_dev(dO.data_ptr(), O.shape[1], O.shape[0], M, b, dZ.data_ptr(), nout)
    return dZ.cpu().numpy()


def e1ve0c(XuE1s, XuE0s, nE):
    """e1ve0::compute_E1vE0c (e1ve0.cpp:55-106): correction matrix between the elevation-class spaces of two coupling
    steps.  XuE*: per-sheet "XvE" matrices (scale=False, correctA=True, identity X dim).  Returns (iE1, iE0, val) in sparse
    E indices, sorted by (iE1, iE0)."""
    n = len(XuE1s)
    assert n == len(XuE0s) and n > 0
    a = (vp * n)(*[w._h for w in XuE1s]); b = (vp * n)(*[w._h for w in XuE0s])
    cnt = lib().ib_e1ve0c(n, a, b, int(nE), None, None, None)
    check(min(cnt, 0))
    r = np.zeros(cnt, np.int64); c = np.zeros(cnt, np.int64); v = np.zeros(cnt)
    check(min(lib().ib_e1ve0c(n, a, b, int(nE), _p(r), _p(c), _p(v)), 0))
    return r, c, v


def combine_chunks(chunks, scale=True):
    """combine_chunks (modele/combine_global_ec.cpp:95-262): concatenate per-chunk UNSCALED matrices into one matrix in sparse
    indexing, scaled by 1 / (sum of wM over all chunks) per row.  Returns (M: scipy coo [sparse extents] with the entries in
    chunk order, duplicates NOT merged; (wM_index, wM_val); (Mw_index, Mw_val))."""
    import scipy.sparse
    n = len(chunks)
    Wp = (vp * n)(*[w._h for w in chunks])
    nnz = sum(w.nnz for w in chunks); nr = sum(w.shape[0] for w in chunks); nc = sum(w.shape[1] for w in chunks)
    r = np.zeros(nnz, np.int64); c = np.zeros(nnz, np.int64); v = np.zeros(nnz)
    wi = np.zeros(nr, np.int64); wv = np.zeros(nr); mi = np.zeros(nc, np.int64); mv = np.zeros(nc)
    check(lib().ib_w_combine_chunks(n, Wp, int(scale), _p(r), _p(c), _p(v), _p(wi), _p(wv), _p(mi), _p(mv)))
    return scipy.sparse.coo_matrix((v, (r, c)), shape=chunks[0].sparse_extent), (wi, wv), (mi, mv)


def exgrid_rectilinear(polys, iA, xe, ye, iI_major_x=True):
    """Exchange grid of projected GCM polygons with a rectilinear ice grid (gridgen/GridGen_Exchange.cpp:175-316).
    polys: list of [n_i, 2] vertex arrays in the ice grid's plane; iA: their GCM cell indices; xe, ye: ice cell edges.
    Returns (ex_ij int32 [n, 2] = (iA, iI), ex_area [n]) -- the contents of an ExchangeGrid (AbbrGrid.hpp:40-78)."""
    pptr = np.zeros(len(polys) + 1, np.int32)
    np.cumsum([len(p) for p in polys], out=pptr[1:])
    pxy = _c(np.concatenate([np.asarray(p, np.float64).reshape(-1, 2) for p in polys]) if len(polys) else np.zeros((0, 2)), np.float64)
    iA = _c(iA, np.int64); xe = _c(xe, np.float64); ye = _c(ye, np.float64)
    args = (len(polys), _p(pptr), _p(pxy), _p(iA), len(xe) - 1, _p(xe), len(ye) - 1, _p(ye), int(iI_major_x))
    n = lib().ib_exgrid_rectilinear(*args, None, None)
    check(min(n, 0))
    ij = np.zeros((n, 2), np.int32); area = np.zeros(n)
    check(min(lib().ib_exgrid_rectilinear(*args, _p(ij), _p(area)), 0))
    return ij, area


def multivec_to_dense(index, weights, vals, nE, fill=np.nan):
    """VectorMultivec::to_dense_scale/to_dense (multivec.cpp:35-81).  vals: [n, nvar].  Returns (scaleE, dense[nvar,nE])."""
    index = _c(index, np.int64); weights = _c(weights, np.float64); vals = _c(vals, np.float64)
    n, nvar = vals.shape
    scale = np.zeros(nE); dense = np.zeros((nvar, nE))
    check(lib().ib_multivec_to_dense(n, _p(index), _p(weights), _p(vals), nvar, nE, float(fill), _p(scale), _p(dense)))
    return scale, dense
