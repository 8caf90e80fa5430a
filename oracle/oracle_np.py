"""Independent numpy/scipy twin of oracle/oracle.cpp.  TEST INFRASTRUCTURE ONLY.

Second restatement of the same reference algebra (RegridMatrices_Dynamic.cpp:50-332, IceRegridder_L0.cpp:100-214,
IceRegridder.cpp:41-65, modele/hntr.cpp:84-168, modele/hntr.hpp:204-435) written against scipy.sparse instead of
hand-rolled CSC, so that a misreading shared by both is less likely.  Pure-Python loops: small cases only.
"""
import math

import numpy as np
import scipy.sparse as sp


class SparseSet:
    def __init__(self, ids=None):
        self.to_d = {}
        self.to_s = []
        for s in ([] if ids is None else ids):
            self.add(int(s))

    def add(self, s):
        d = self.to_d.get(s)
        if d is None:
            d = len(self.to_s)
            self.to_d[s] = d
            self.to_s.append(s)
        return d

    def __len__(self):
        return len(self.to_s)


def _interp(hc, elev):
    n = len(hc)
    hi = int(np.searchsorted(hc, elev, side="left"))     # std::lower_bound
    hi = max(hi, 1)
    if hi >= n:
        raise ValueError("Elevation %g out of bounds" % elev)
    lo = hi - 1
    ratio = (elev - hc[lo]) / (hc[hi] - hc[lo])
    return (lo, hi), (1.0 - ratio, ratio)


def _gen(kind, gridA, sheet, hc, em):
    """Generator tuples (i_sparse, j_sparse, value) for gridG == 'X'."""
    out = []
    nA = gridA.nA
    for x in range(sheet.nX):
        a, i = int(sheet.ex_ij[x, 0]), int(sheet.ex_ij[x, 1])
        if math.isnan(em[i]):
            continue
        s = float(sheet.ex_area[x])
        if kind == "GvI":
            out.append((x, i, s))
        elif kind == "GvAp":
            if s > 0:
                out.append((x, a, s))
        else:
            (k0, k1), (w0, w1) = _interp(hc, max(float(em[i]), 0.0))
            if w0 != 0:
                out.append((x, k0 * nA + a, s * w0))
            if w1 != 0:
                out.append((x, k1 * nA + a, s * w1))
    return out


def _diag_gen(kind, gridA, sheet, nhc):
    out = []
    for d, a in enumerate(gridA.dim):
        r = gridA.native_area[d] / sheet.gridA_proj_area[d]
        if kind == "A":
            out.append((int(a), r))
        else:
            out.extend((k * gridA.nA + int(a), r) for k in range(nhc))
    return out


def _dense(gen, d0, d1):
    r = [d0.add(t[0]) for t in gen]
    c = [d1.add(t[1]) for t in gen]
    v = [t[2] for t in gen]
    return r, c, v


def _mat(rcv, n0, n1, transpose=False):
    r, c, v = rcv
    M = sp.coo_matrix((v, (r, c)), shape=(n0, n1)).tocsc()
    M.sum_duplicates()
    return M.T.tocsc() if transpose else M


def _diag(gen, dim):
    v = np.zeros(len(dim))
    for s, r in gen:
        d = dim.to_d.get(s)
        if d is not None:
            v[d] = r
    return v


def _rows(M):
    return np.asarray(M.sum(axis=1)).ravel()


def _cols(M):
    return np.asarray(M.sum(axis=0)).ravel()


def matrix(gridA, sheet, hc, em, spec, scale=True, correctA=False, dims=(None, None)):
    """Returns dict(M=csc, wM, Mw, dims=[ids0, ids1]) with the reference's dense numbering."""
    b, a = spec[0], spec[2]
    nhc = len(hc)
    dG = SparseSet()
    dB, dA = SparseSet(dims[0]), SparseSet(dims[1])
    gen = {k: None for k in ("GvI", "GvAp", "GvEp")}

    def g(kind):
        if gen[kind] is None:
            gen[kind] = _gen(kind, gridA, sheet, hc, em)
        return gen[kind]
    ur = {"A": ("GvAp", "A"), "E": ("GvEp", "E")}
    with np.errstate(divide="ignore", invalid="ignore"):
        if b in "AE" and a == "I":
            t1 = _dense(g(ur[b][0]), dG, dB)
            t2 = _dense(g("GvI"), dG, dA)
            ApvG = _mat(t1, len(dG), len(dB), True)
            GvI = _mat(t2, len(dG), len(dA))
            ApvI = (ApvG @ sp.diags(1.0 / _rows(GvI))) @ GvI
            Mw = _cols(ApvI)
            wApvI = _rows(ApvI)
            if correctA:
                w = _diag(_diag_gen(ur[b][1], gridA, sheet, nhc), dB)
                wM = w * wApvI
                M = sp.diags((1.0 / w) * (1.0 / wApvI)) @ ApvI if scale else ApvI
            else:
                wM = wApvI
                M = sp.diags(1.0 / wApvI) @ ApvI if scale else ApvI
        elif b == "I" and a in "AE":
            t1 = _dense(g(ur[a][0]), dG, dA)
            t2 = _dense(g("GvI"), dG, dB)
            GvAp = _mat(t1, len(dG), len(dA))
            IvG = _mat(t2, len(dG), len(dB), True)
            IvAp = (IvG @ sp.diags(1.0 / _rows(GvAp))) @ GvAp
            wM = _rows(IvAp)
            if correctA:
                s = _diag(_diag_gen(ur[a][1], gridA, sheet, nhc), dA)
                Mw = s * _cols(IvAp)
                M = (sp.diags(1.0 / wM) @ IvAp if scale else IvAp) @ sp.diags(s)
            else:
                Mw = _cols(IvAp)
                M = sp.diags(1.0 / wM) @ IvAp if scale else IvAp
        elif (b, a) in (("E", "A"), ("A", "E")):
            tA = _dense(g(ur[a][0]), dG, dA)
            tE = _dense(g(ur[b][0]), dG, dB)
            GvAp = _mat(tA, len(dG), len(dA))
            EpvG = _mat(tE, len(dG), len(dB), True)
            EpvAp = (EpvG @ sp.diags(1.0 / _rows(GvAp))) @ GvAp
            wE = _rows(EpvAp)
            if correctA:
                sA = _diag(_diag_gen(ur[a][1], gridA, sheet, nhc), dA)
                sE = _diag(_diag_gen(ur[b][1], gridA, sheet, nhc), dB)
                wM = sE * wE
                Mw = sA * _cols(EpvAp)
                M = (sp.diags(1.0 / wM) @ EpvAp if scale else EpvAp) @ sp.diags(sA)
            else:
                wM, Mw = wE, _cols(EpvAp)
                M = sp.diags(1.0 / wE) @ EpvAp if scale else EpvAp
        else:
            raise KeyError(spec)
    M = sp.csc_matrix(M)
    M.sort_indices()
    return dict(M=M, wM=wM, Mw=Mw, dims=[np.array(dB.to_s, np.int64), np.array(dA.to_s, np.int64)])


# ----------------------------------------------------------------------------------------------
class Hntr:
    """Loop-for-loop numpy twin of the Hntr partition + regrid; spec = (im, jm, offi, dlat_minutes)."""

    def __init__(self, specB, specA, datmis=0.0):
        (imB, jmB, offB, dlB), (imA, jmA, offA, dlA) = specB, specA
        self.specB, self.specA, self.datmis = specB, specA, datmis
        self.IMIN = np.zeros(imB, int); self.IMAX = np.zeros(imB, int)
        self.FMIN = np.zeros(imB); self.FMAX = np.zeros(imB)
        ia, ria, prev = 1, (1 + offA - imA) * imB, imB
        for ib in range(1, imB + 1):
            rib = (ib - 1 + offB) * imA
            while ria < rib:
                ia += 1; ria += imB
            if ria == rib:
                self.IMAX[prev - 1], self.FMAX[prev - 1] = ia, 0.0
                ia += 1; ria += imB
                self.IMIN[ib - 1], self.FMIN[ib - 1] = ia, 0.0
            else:
                self.IMAX[prev - 1], self.FMAX[prev - 1] = ia, (ria - rib) / imB
                self.IMIN[ib - 1], self.FMIN[ib - 1] = ia, 1 - self.FMAX[prev - 1]
            prev = ib
        self.IMAX[imB - 1] += imA
        k = (2 * math.pi) / (360 * 60)
        self.SINA = np.array([-1.0] + [math.sin((j + .5 - .5 * (1 + jmA)) * dlA * k) for j in range(1, jmA)] + [1.0])
        self.SINB = np.array([-1.0] + [math.sin((j + .5 - .5 * (1 + jmB)) * dlB * k) for j in range(1, jmB)] + [1.0])
        self.JMIN = np.zeros(jmB, int); self.JMAX = np.zeros(jmB, int)
        self.GMIN = np.zeros(jmB); self.GMAX = np.zeros(jmB)
        self.JMIN[0] = 1
        ja = 1
        for jb in range(1, jmB):
            while self.SINA[ja] < self.SINB[jb]:
                ja += 1
            if self.SINA[ja] == self.SINB[jb]:
                self.JMAX[jb - 1] = ja
                ja += 1
                self.JMIN[jb] = ja
            else:
                self.JMAX[jb - 1], self.GMAX[jb - 1] = ja, self.SINA[ja] - self.SINB[jb]
                self.JMIN[jb], self.GMIN[jb] = ja, self.SINB[jb] - self.SINA[ja - 1]
        self.JMAX[jmB - 1] = jmA

    def regrid(self, WTA, A, mean_polar=False):
        (imB, jmB, _, _), (imA, jmA, _, _) = self.specB, self.specA
        A = np.asarray(A, float).reshape(jmA, imA)
        W = np.ones_like(A) if WTA is None else np.asarray(WTA, float).reshape(jmA, imA)
        B = np.zeros((jmB, imB))
        for jb in range(jmB):
            for ib in range(imB):
                wsum = vsum = 0.0
                for ja in range(self.JMIN[jb], self.JMAX[jb] + 1):
                    G = self.SINA[ja] - self.SINA[ja - 1]
                    if ja == self.JMIN[jb]: G -= self.GMIN[jb]
                    if ja == self.JMAX[jb]: G -= self.GMAX[jb]
                    for ir in range(self.IMIN[ib], self.IMAX[ib] + 1):
                        F = 1.0
                        if ir == self.IMIN[ib]: F -= self.FMIN[ib]
                        if ir == self.IMAX[ib]: F -= self.FMAX[ib]
                        wt = (F * G) * W[ja - 1, (ir - 1) % imA]
                        wsum += wt
                        vsum += wt * A[ja - 1, (ir - 1) % imA]
                B[jb, ib] = self.datmis if wsum == 0 else vsum / wsum
        if mean_polar:
            for jb in {0, jmB - 1}:
                B[jb, :] = self.datmis if np.any(B[jb] == self.datmis) else B[jb].sum() / imB
        return B.ravel()


# ------------------------------------------------------------------------------------------------------------------
# GCMRegridder_ModelE (slib/icebin/modele/GCMRegridder_ModelE.cpp:130-435, topo.cpp:46-240): a literal restatement with
# scipy sparse products, on top of the C++ oracle's ocean-grid matrices.  TEST INFRASTRUCTURE.  Parity unpinned: the
# reference holds no test or golden vector for these matrices.
class ModelE:
    """go: oracle Gcm over the OCEAN grid; specO / specA: (im, jm, offi, dlat) of the ocean / atmosphere Hntr grids."""

    def __init__(self, go, sheet, elevmaskI, foceanAOp, foceanAOm, specO, eq_rad, nhc, nG):
        from oracle import orc
        self.go, self.sheet, self.em = go, sheet, elevmaskI
        self.fp, self.fm = np.asarray(foceanAOp, float), np.asarray(foceanAOm, float)
        self.nO = specO[0] * specO[1]
        self.specA = (specO[0] // 2, specO[1] // 2, specO[2] * 0.5, specO[3] * 2.0)       # make_hntrA, hntr.cpp:232-241
        self.nA = self.specA[0] * self.specA[1]
        self.nhc, self.nG = nhc, nG
        self.ov = orc.Hntr(specO, self.specA).matrix(eq_rad)          # Hntr(B = O, A = A).overlap: (iO, iA, area)

    @staticmethod
    def _sset(ids=()):
        d = {}
        for s in ids:
            d.setdefault(int(s), len(d))
        return d

    def helper(self, X, G, dimIp=None):
        import scipy.sparse as sp
        m = lambda spec, scale, correctA, dims: self.go.matrix(self.sheet, self.em, spec, scale=scale, correctA=correctA, dims=dims)
        AOpvGp_c = m("Av" + G, False, True, (None, dimIp))                                 # :185-188
        wAOp, dimAOp, dimIp = AOpvGp_c.wM, self._sset(AOpvGp_c.dims[0]), AOpvGp_c.dims[1]
        # compute_wAOm (topo.cpp:84-110)
        dimAOm = self._sset(s for s in dimAOp if self.fm[s] == 0)
        sc = sp.lil_matrix((len(dimAOm), len(dimAOp)))
        for s, dp in dimAOp.items():                                                       # scaled_AOmvAOp (topo.cpp:46-80)
            fcont_p, fcont_m = 1.0 - self.fp[s], 1.0 - self.fm[s]
            if fcont_m == 0.0:
                continue
            assert fcont_m == 1.0
            if fcont_p == 0.0:
                continue
            sc[dimAOm[s], dp] = 1.0 / fcont_p
        wAOm = sc.tocsr() @ wAOp
        if X == "A":
            XOpvGp = m("Av" + G, False, False, (AOpvGp_c.dims[0], dimIp))                  # :252-259
            dimXOp, dimXOm, wXOm = dimAOp, dimAOm, wAOm
            dimXAm = {}
            r, c, v = [], [], []
            for io, ia, a in zip(*self.ov):                                                # :264-273, transposed
                if int(io) in dimAOm:
                    r.append(dimXAm.setdefault(int(ia), len(dimXAm))); c.append(dimAOm[int(io)]); v.append(a)
        else:
            XOpvGp = m("Ev" + G, False, False, (None, dimIp))                              # :208-213
            dimXOp = self._sset(XOpvGp.dims[0])
            EOpvAOp = m("EvA", False, False, (None, None))                                 # :219-221
            dimXOm = {}
            er, ec, ev = [], [], []
            for rr, cc, vv in zip(EOpvAOp.row, EOpvAOp.col, EOpvAOp.val):                  # column-major order (topo.cpp:211-240)
                sA = int(EOpvAOp.dims[1][cc])
                if sA in dimAOm:
                    er.append(dimXOm.setdefault(int(EOpvAOp.dims[0][rr]), len(dimXOm))); ec.append(dimAOm[sA]); ev.append(vv)
            EOmvAOm = sp.csr_matrix((ev, (er, ec)), shape=(len(dimXOm), len(dimAOm)))
            with np.errstate(divide="ignore"):
                EOmvAOms = 1.0 / np.asarray(EOmvAOm.sum(0)).ravel()                        # sum(EOmvAOm, 1, '-')
            wXOm = EOmvAOm @ sp.diags(EOmvAOms) @ wAOm                                     # :226-229
            dimXAm = {}
            r, c, v = [], [], []
            for io, ia, a in zip(*self.ov):                                                # raw_EOvEA (topo.cpp:112-208), transposed
                if int(io) not in dimAOm:
                    continue
                for ihc in range(self.nhc):
                    sEO = ihc * self.nO + int(io)
                    if sEO not in dimXOm or wXOm[dimXOm[sEO]] == 0:
                        continue
                    r.append(dimXAm.setdefault(ihc * self.nA + int(ia), len(dimXAm))); c.append(dimXOm[sEO]); v.append(wXOm[dimXOm[sEO]])
        XAmvXOm = sp.csr_matrix((v, (r, c)), shape=(len(dimXAm), len(dimXOm)))
        with np.errstate(divide="ignore"):
            XAmvXOms = 1.0 / np.asarray(XAmvXOm.sum(0)).ravel()                            # :277
        wXAm = XAmvXOm @ sp.diags(XAmvXOms) @ wXOm                                         # :281-282
        return dict(XOpvGp=XOpvGp, dimXOp=dimXOp, dimXOm=dimXOm, dimXAm=dimXAm, dimIp=dimIp, XAmvXOm=XAmvXOm, XAmvXOms=XAmvXOms,
                    wXAm=wXAm)

    @staticmethod
    def crop_mvp(dimAm, dimAp_ids, index, M):
        """crop_mvp (:285-307): re-index dimension `index` of M from dense Ap to dense Am, dropping what is not in dimAm."""
        import scipy.sparse as sp
        M = M.tocoo()
        idx = [M.row, M.col]
        new = np.array([dimAm.get(int(dimAp_ids[i]), -1) for i in idx[index]], dtype=np.int64)
        keep = new >= 0
        idx[index] = new
        shape = list(M.shape)
        shape[index] = len(dimAm)
        return sp.csr_matrix((M.data[keep], (idx[0][keep], idx[1][keep])), shape=shape)

    def XAmvGp(self, X, G, scale=True):
        """compute_XAmvGp (:318-386) -> (dims[0] ids, dims[1] ids, M csr, wM, Mw)"""
        import scipy.sparse as sp
        H = self.helper(X, G)
        W = H["XOpvGp"]
        Mo = sp.csr_matrix((W.val, (W.row, W.col)), shape=W.shape)
        with np.errstate(divide="ignore"):
            sXOpvGp = 1.0 / W.wM
            sXAm = 1.0 / np.asarray(H["XAmvXOm"].sum(1)).ravel()
        rowscale = sXAm if scale else H["wXAm"] * sXAm
        M = (H["XAmvXOm"] @ self.crop_mvp(H["dimXOm"], W.dims[0], 0, sp.diags(sXOpvGp) @ Mo)).tocsr()
        # Eigen's diagonal * sparse keeps the pattern where the scale is 0 (unscaled rows with wXAm == 0: continent for ModelE,
        # ocean for the ice model); scipy's diags() product would drop those entries, so scale the stored values in place
        M.data *= np.repeat(rowscale, np.diff(M.indptr))
        return np.array(list(H["dimXAm"]), np.int64), np.asarray(H["dimIp"]), M, H["wXAm"], W.Mw

    def GpvXAm(self, G, X, scale=True):
        """compute_GpvXAm (:388-435)"""
        import scipy.sparse as sp
        H = self.helper(X, G)
        W = self.go.matrix(self.sheet, self.em, G + "v" + X, scale=False, correctA=False, dims=(H["dimIp"], H["XOpvGp"].dims[0]))
        Mo = sp.csr_matrix((W.val, (W.row, W.col)), shape=W.shape)
        if scale:
            with np.errstate(divide="ignore"):
                Mo = sp.diags(1.0 / W.wM) @ Mo
        M = self.crop_mvp(H["dimXOm"], W.dims[1], 1, Mo) @ sp.diags(H["XAmvXOms"]) @ H["XAmvXOm"].T
        return np.asarray(W.dims[0]), np.array(list(H["dimXAm"]), np.int64), M.tocsr(), H["XOpvGp"].Mw, H["wXAm"]
