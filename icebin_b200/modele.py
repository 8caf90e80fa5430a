"""GCMRegridder_ModelE: "mismatched" regridding between ModelE's atmosphere grid and an ice sheet, through the ocean grid
the ice sheet was overlapped with (slib/icebin/modele/GCMRegridder_ModelE.cpp:92-121, 130-282, 285-307, 318-435, 490-575;
topo.cpp:46-110, 112-240; hntr.cpp:232-241).

Grids and their two masks (the reference's naming):
    O  = ocean grid (what `gcmO`, an ordinary GCMRegridder, calls its A grid); A = atmosphere grid = O coarsened 2:1;
    E  = elevation grid on either;  I = ice grid;  X = exchange grid;
    p  = "as the ice model (PISM) sees it": foceanAOp in [0, 1];   m = "as ModelE sees it": foceanAOm in {0, 1}.
The regridder offers, per sheet and elevation mask,
    EvI = EAmvIp, AvI = AAmvIp, EvX, AvX          XAmvGp = diag(sXAm) * XAmvXOm * crop(diag(1 / XOpvGp.wM) * XOpvGp)
    IvE = IpvEAm, IvA = IpvAAm, XvE, XvA          GpvXAm = crop(diag(1 / GpvXOp.wM) * GpvXOp) * diag(XAmvXOms) * XOmvXAm
    AOmvAAm, AAmvAOm                              Hntr overlap of the two grids, clipped to ModelE's continent
where XOpvGp / GpvXOp are the ocean-grid regridder's own unscaled matrices (built on the GPU as always) and XAmvXOm is a
SMALL matrix (A/E-sized) with one entry per ocean-grid cell.  The small factors are computed here on the host with numpy;
the product with the LARGE ice-side matrix is `ib_w_aggregate` on the device (csrc/modele.cu): re-index, scale, add up.

Not built: global (base-ice) elevation classes read from a global_ec file (`global_ecO`, :466-487) and `global_AvE`
(:577-646), which merge the TOPO generator's output; `GCMRegridder_WrapE` is a pair of stored focean arrays.
"""
import ctypes as C

import numpy as np

from . import api
from ._lib import check, lib, vp


def make_hntrA(hspecO):
    """hntr.cpp:232-241: the atmosphere grid is exactly twice the ocean grid."""
    if hspecO.im % 2 or hspecO.jm % 2:
        raise ValueError("Ocean grid must have even number of gridcells for im and jm (vs. %d %d)" % (hspecO.im, hspecO.jm))
    return api.HntrSpec(hspecO.im // 2, hspecO.jm // 2, hspecO.offi * 0.5, hspecO.dlat * 2.0)


class _Dim:
    """The part of SparseSetT the host factors need: ordered ids with ADD_DENSE / TO_DENSE_IGNORE_MISSING."""

    def __init__(self, ids=(), sparse_extent=-1):
        self.ids = [int(i) for i in ids]
        self.pos = {s: d for d, s in enumerate(self.ids)}
        self.sparse_extent = sparse_extent

    def add_dense(self, s):
        s = int(s)
        d = self.pos.get(s)
        if d is None:
            d = len(self.ids)
            self.pos[s] = d
            self.ids.append(s)
        return d

    def to_dense(self, s):
        return self.pos.get(int(s), -1)

    def __len__(self):
        return len(self.ids)


def _aggregate(W, side, map_, factor, n_new, other):
    h = vp()
    m = np.ascontiguousarray(map_, np.int32)
    f = None if factor is None else np.ascontiguousarray(factor, np.float64)
    o = None if other is None else np.ascontiguousarray(other, np.float64)
    check(lib().ib_w_aggregate(W._h, side, m.ctypes.data_as(vp), None if f is None else f.ctypes.data_as(vp), int(n_new),
                               None if o is None else o.ctypes.data_as(vp), C.byref(h)))
    return h


def _finish(h, dims, extents, wM, Mw, scaled):
    for k in range(2):
        ids = np.ascontiguousarray(dims[k], np.int64)
        check(lib().ib_w_set_dim(h, k, int(extents[k]), len(ids), ids.ctypes.data_as(vp)))
    wM = np.ascontiguousarray(wM, np.float64); Mw = np.ascontiguousarray(Mw, np.float64)
    check(lib().ib_w_set_weights(h, wM.ctypes.data_as(vp), Mw.ctypes.data_as(vp), 0, int(scaled)))      # not conservative (:331, :403)
    return api.Weighted(h)


class GCMRegridder_ModelE:
    """GCMRegridder_ModelE(global_ecO = "", gcmO) (GCMRegridder_ModelE.cpp:438-487)."""

    def __init__(self, gcmO, hspecO, eq_rad=6371000.0):
        self.gcmO, self.hspecO, self.eq_rad = gcmO, hspecO, float(eq_rad)
        self.hspecA = make_hntrA(hspecO)
        if gcmO.nA != hspecO.size:
            raise ValueError("gcmO's GCM grid has %d cells, the ocean HntrSpec %d" % (gcmO.nA, hspecO.size))
        self.nhc = gcmO.nhc
        self.nO, self.nA = hspecO.size, self.hspecA.size
        # AOvAA: Hntr(B = O, A = A).overlap(R = eq_rad): (iO, iA, area of the overlap); hntr.hpp:282-290
        r, c, v = api.Hntr(17.17, hspecO, self.hspecA, 0.0).overlap(self.eq_rad)
        self._ov = (r, c, v)

    def nAE(self, X):
        return self.nA if X == "A" else self.nA * self.nhc

    def regrid_matrices(self, sheet_name, foceanAOp, foceanAOm, elevmaskI, scale=True, sigma=(0, 0, 0)):
        return RegridMatrices_ModelE(self, sheet_name, np.asarray(foceanAOp, np.float64), np.asarray(foceanAOm, np.float64),
                                     elevmaskI, scale, sigma)


class RegridMatrices_ModelE:
    """The closures GCMRegridder_ModelE::regrid_matrices registers (:490-575)."""

    SPECS = {"EvI": ("XG", "E", "I"), "AvI": ("XG", "A", "I"), "EvX": ("XG", "E", "X"), "AvX": ("XG", "A", "X"),
             "EAmvIp": ("XG", "E", "I"), "AAmvIp": ("XG", "A", "I"),
             "IvE": ("GX", "I", "E"), "IvA": ("GX", "I", "A"), "XvE": ("GX", "X", "E"), "XvA": ("GX", "X", "A"),
             "IpvEAm": ("GX", "I", "E"), "IpvAAm": ("GX", "I", "A"),
             "AOmvAAm": ("OA", ".", ""), "AAmvAOm": ("OA", "T", "")}

    def __init__(self, gcmA, sheet_name, foceanAOp, foceanAOm, elevmaskI, scale, sigma):
        self.gcmA, self.sheet = gcmA, sheet_name
        self.foceanAOp, self.foceanAOm, self.scale = foceanAOp, foceanAOm, bool(scale)
        if foceanAOp.shape != (gcmA.nO,) or foceanAOm.shape != (gcmA.nO,):
            raise ValueError("focean arrays must have one entry per ocean-grid cell (%d)" % gcmA.nO)
        # rmO: the ocean-grid regridder's matrices for this sheet and mask; scale / correctA are set per matrix below
        self.rmO = gcmA.gcmO.regrid_matrices(sheet_name, elevmaskI, scale=False, correctA=False, sigma=sigma)

    def matrix(self, spec):
        if spec not in self.SPECS:
            raise KeyError(spec)
        kind, a, b = self.SPECS[spec]
        if kind == "XG":
            return self._XAmvGp(a, b)
        if kind == "GX":
            return self._GpvXAm(a, b)
        return self._AOmvAAm(a)

    # ---- ComputeXAmvGp_Helper (:130-282)
    def _helper(self, X, G):
        g, rmO = self.gcmA, self.rmO
        dimIp = api.SparseSet()
        dimAOp_s = api.SparseSet()
        # wAOp: weights of AOpvGp with correctA (:185-188)
        AOpvGp_c = rmO.matrix_d("Av" + G, (dimAOp_s, dimIp), scale=False, correctA=True)
        wAOp = AOpvGp_c.get_weights(0)
        dimAOp = _Dim(dimAOp_s.ids(), g.nO)
        # compute_wAOm (topo.cpp:84-110) with scaled_AOmvAOp (:46-80)
        dimAOm = _Dim(sparse_extent=g.nO)
        for s in dimAOp.ids:
            if self.foceanAOm[s] == 0:
                dimAOm.add_dense(s)
        wAOm = np.zeros(len(dimAOm))
        for d, s in enumerate(dimAOp.ids):
            fcont_p, fcont_m = 1.0 - self.foceanAOp[s], 1.0 - self.foceanAOm[s]
            if fcont_m == 0.0:
                continue
            if fcont_m != 1.0:
                raise ValueError("fcont_m[%d] = %g, must be 0 or 1" % (s, fcont_m))
            if fcont_p == 0.0:
                continue
            wAOm[dimAOm.to_dense(s)] = (1.0 / fcont_p) * wAOp[d]
        ovO, ovA, ovV = g._ov
        H = dict(dimIp=dimIp, dimAOp=dimAOp, dimAOm=dimAOm, wAOm=wAOm)
        if X == "A":
            # XOpvGp = AOpvGp without correctA, on the same dims (:252-259)
            H["XOpvGp"] = rmO.matrix_d("Av" + G, (dimAOp_s, dimIp), scale=False, correctA=False)
            H["dimXOp_s"] = dimAOp_s
            dimXOp, dimXOm, wXOm = dimAOp, dimAOm, wAOm
            # XAmvXOm = overlap(O <- A) clipped to dimAOm, transposed (:264-273): (iAA, iAOm, area)
            dimXAm = _Dim(sparse_extent=g.nA)
            T = []
            for io, ia, v in zip(ovO, ovA, ovV):
                dm = dimAOm.to_dense(io)
                if dm >= 0:
                    T.append((dimXAm.add_dense(ia), dm, v))
        else:
            dimEOp_s = api.SparseSet()
            H["XOpvGp"] = rmO.matrix_d("Ev" + G, (dimEOp_s, dimIp), scale=False, correctA=False)
            H["dimXOp_s"] = dimEOp_s
            dimXOp = _Dim(dimEOp_s.ids(), g.nO * g.nhc)
            # EOpvAOp on fresh dims (:219-221), cropped to ModelE's continent: EOmvAOm and dimEOm (topo.cpp:211-240)
            dE2, dA2 = api.SparseSet(), api.SparseSet()
            EOpvAOp = rmO.matrix_d("EvA", (dE2, dA2), scale=False, correctA=False)
            r, c, v = EOpvAOp.to_coo_dense()
            o = np.lexsort((r, c))                                   # Eigen's (column-major) iteration order
            idsE2, idsA2 = dE2.ids(), dA2.ids()
            dimXOm = _Dim(sparse_extent=g.nO * g.nhc)
            ent = []
            for rr, cc, vv in zip(r[o], c[o], v[o]):
                dm = dimAOm.to_dense(idsA2[cc])
                if dm >= 0:
                    ent.append((dimXOm.add_dense(idsE2[rr]), dm, vv))
            colsum = np.zeros(len(dimAOm))
            for e, am, vv in ent:
                colsum[am] += vv
            wXOm = np.zeros(len(dimXOm))
            for e, am, vv in ent:                                     # wEOm = EOmvAOm * diag(1 / colsum) * wAOm  (:226-229)
                wXOm[e] += vv * (1.0 / colsum[am]) * wAOm[am]
            # raw_EOvEA (topo.cpp:112-208), transposed: (iEAm, iEOm, wEOm)
            dimXAm = _Dim(sparse_extent=g.nA * g.nhc)
            T = []
            for io, ia, v in zip(ovO, ovA, ovV):
                if dimAOm.to_dense(io) < 0:
                    continue
                if abs(v) < 1e-8:
                    raise ValueError("Found a stray overlap; what should we do about it?")
                for ihc in range(g.nhc):
                    em = dimXOm.to_dense(ihc * g.nO + io)            # indexingHC: iE = ihc * nA + iA
                    if em < 0 or wXOm[em] == 0:
                        continue
                    T.append((dimXAm.add_dense(ihc * g.nA + ia), em, wXOm[em]))
        nXAm, nXOm = len(dimXAm), len(dimXOm)
        colsum = np.zeros(nXOm); rowsum = np.zeros(nXAm)
        for xa, xo, v in T:
            colsum[xo] += v
            rowsum[xa] += v
        with np.errstate(divide="ignore"):
            XAmvXOms = 1.0 / colsum                                  # sum(XAmvXOm, 1, '-')  (:277)
            sXAm = 1.0 / rowsum                                      # sum(XAmvXOm, 0, '-')  (:353)
        wXAm = np.zeros(nXAm)
        for xa, xo, v in T:                                          # wXAm = XAmvXOm * diag(XAmvXOms) * wXOm  (:281-282)
            wXAm[xa] += v * XAmvXOms[xo] * wXOm[xo]
        per_col = np.zeros(nXOm, np.int64)
        for xa, xo, v in T:
            per_col[xo] += 1
        if per_col.max(initial=0) > 1:
            raise NotImplementedError("an ocean-grid cell overlaps more than one atmosphere cell: the grids do not nest 2:1")
        H.update(dimXOp=dimXOp, dimXOm=dimXOm, dimXAm=dimXAm, T=T, XAmvXOms=XAmvXOms, sXAm=sXAm, wXAm=wXAm)
        return H

    def _maps(self, H):
        """ocean-grid dense index (p) -> atmosphere dense index and the entry of XAmvXOm, through dimXOm (crop_mvp, :285-307)"""
        dimXOp, dimXOm = H["dimXOp"], H["dimXOm"]
        xa_of_m = np.full(len(dimXOm), -1, np.int64); t_of_m = np.zeros(len(dimXOm))
        for xa, xo, v in H["T"]:
            xa_of_m[xo] = xa
            t_of_m[xo] = v
        m_of_p = np.array([dimXOm.to_dense(s) for s in dimXOp.ids], np.int64)
        ok = m_of_p >= 0
        amap = np.full(len(dimXOp), -1, np.int32)
        amap[ok] = xa_of_m[m_of_p[ok]]
        return amap, m_of_p, t_of_m

    def _XAmvGp(self, X, G):
        H = self._helper(X, G)
        g = self.gcmA
        W = H["XOpvGp"]
        amap, m_of_p, t_of_m = self._maps(H)
        with np.errstate(divide="ignore"):
            sXOpvGp = 1.0 / W.get_weights(0)                         # :349
        rowscale = H["sXAm"] * H["wXAm"] if not self.scale else H["sXAm"]            # :354-375
        factor = np.zeros(len(amap))
        ok = amap >= 0
        factor[ok] = rowscale[amap[ok]] * t_of_m[m_of_p[ok]]
        # M = diag(rowscale) * XAmvXOm * crop(diag(sXOpvGp) * XOpvGp): every entry of row p of the large matrix is scaled by
        # rowscale[xa(p)] * XAmvXOm[xa(p), p] * sXOpvGp[p] and moves to row xa(p)
        h = _aggregate(W, 0, amap, factor * np.where(ok, sXOpvGp, 0.0), len(H["dimXAm"]), None)
        nG = g.gcmO.nI(self.sheet) if G == "I" else g.gcmO.nX(self.sheet)
        return _finish(h, (H["dimXAm"].ids, H["dimIp"].ids()), (g.nAE(X), nG), H["wXAm"], W.get_weights(1), self.scale)

    def _GpvXAm(self, G, X):
        H = self._helper(X, G)
        g = self.gcmA
        # IpvXOp on (dimIp, dimXOp) (:407-413)
        W = self.rmO.matrix_d(G + "v" + X, (H["dimIp"], H["dimXOp_s"]), scale=False, correctA=False)
        dimXOp = _Dim(H["dimXOp_s"].ids(), H["dimXOp"].sparse_extent)
        H = dict(H, dimXOp=dimXOp)
        amap, m_of_p, t_of_m = self._maps(H)
        ok = amap >= 0
        factor = np.zeros(len(amap))
        factor[ok] = H["XAmvXOms"][m_of_p[ok]] * t_of_m[m_of_p[ok]]  # diag(sXOmvXAm) * XOmvXAm  (:420-431)
        other = None
        if self.scale:
            with np.errstate(divide="ignore"):
                other = 1.0 / W.get_weights(0)                       # sIpvXOp (:421)
        h = _aggregate(W, 1, amap, factor, len(H["dimXAm"]), other)
        nG = g.gcmO.nI(self.sheet) if G == "I" else g.gcmO.nX(self.sheet)
        return _finish(h, (H["dimIp"].ids(), H["dimXAm"].ids), (nG, g.nAE(X)), H["XOpvGp"].get_weights(1), H["wXAm"], self.scale)

    def _AOmvAAm(self, transpose):
        """compute_AOmvAAm (:92-121): the overlap matrix clipped to the ocean-grid cells in dimAOm -- here, as the reference's
        test entry point does with a caller-supplied dimAOm, the cells ModelE sees as continent (foceanAOm == 0)."""
        g = self.gcmA
        ovO, ovA, ovV = g._ov
        dimAOm = _Dim([s for s in range(g.nO) if self.foceanAOm[s] == 0], g.nO)
        dimAAm = _Dim(sparse_extent=g.nA)
        rows, cols, vals = [], [], []
        for io, ia, v in zip(ovO, ovA, ovV):
            dm = dimAOm.to_dense(io)
            if dm >= 0:
                rows.append(dm); cols.append(dimAAm.add_dense(ia)); vals.append(v)
        import scipy.sparse as sp
        M = sp.csr_matrix((vals, (rows, cols)), shape=(len(dimAOm), len(dimAAm)))
        dims, ext = (dimAOm.ids, dimAAm.ids), (g.nO, g.nA)
        if transpose == "T":
            M, dims, ext = M.T.tocsr(), dims[::-1], ext[::-1]
        M.sort_indices()
        wM, Mw = np.asarray(M.sum(1)).ravel(), np.asarray(M.sum(0)).ravel()
        W = api.Weighted.from_csr(M.shape[0], M.shape[1], M.indptr, M.indices, M.data, wM, Mw, True)
        h, W._h = W._h, None                      # re-wrap after the dims are in place (a Weighted caches its extents)
        for k in range(2):
            ids = np.ascontiguousarray(dims[k], np.int64)
            check(lib().ib_w_set_dim(h, k, ext[k], len(ids), ids.ctypes.data_as(vp)))
        return api.Weighted(h)
