"""combine_chunks (modele/combine_global_ec.cpp:95-262): per-chunk unscaled matrices -> one matrix in sparse indexing.
CPU: the oracle's combination of GCM-cell domains reproduces the undivided scaled matrix, and the combination of two sheets
applied to a field equals the VectorMultivec merge of the sheets' contributions.  GPU: parity with the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_gpu, make_oracle
from couple_common import two_sheet_config
from icebin_b200 import multi


def sparse_of(W):
    return sp.csr_matrix((W.val, (W.dims[0][W.row], W.dims[1][W.col])), shape=W.sparse_extent)


def test_oracle_domains_combine_to_the_scaled_matrix(cfg1):
    from oracle import orc
    em, sh, nA = cfg1["elevmask"][0], cfg1["sheets"][0], cfg1["gridA"].nA
    keep = multi.partition_cellsA(sh.ex_ij[:, 0], nA, 3)
    chunks = []
    for r in range(3):
        g = make_oracle(cfg1)
        g.filter_cellsA(keep[r])
        chunks.append(g.matrix(0, em, "EvI", scale=False, correctA=False))
    M, (wi, wv), (mi, mv) = orc.combine_chunks(chunks, scale=True)
    full = make_oracle(cfg1).matrix(0, em, "EvI", scale=True, correctA=False)
    F = sparse_of(full)
    assert abs(M.tocsr() - F).max() <= 1e-13 * abs(F).max()
    assert M.nnz == sum(c.nnz for c in chunks) and len(np.unique(wi)) == len(wi)          # domains own disjoint rows
    wfull = np.zeros(full.sparse_extent[0]); wfull[full.dims[0]] = full.wM
    wcomb = np.zeros_like(wfull); wcomb[wi] = wv
    np.testing.assert_allclose(wcomb, wfull, rtol=1e-13)
    # ice cells on a domain boundary appear in two chunks: their Mw entries add up to the undivided weight
    mfull = np.zeros(full.sparse_extent[1]); mfull[full.dims[1]] = full.Mw
    mcomb = np.zeros_like(mfull); np.add.at(mcomb, mi, mv)
    np.testing.assert_allclose(mcomb, mfull, rtol=1e-12)
    Mu, _, _ = orc.combine_chunks(chunks, scale=False)
    U = sparse_of(make_oracle(cfg1).matrix(0, em, "EvI", scale=False, correctA=False))
    assert abs(Mu.tocsr() - U).max() <= 1e-13 * abs(U).max()


def test_oracle_two_sheets_combine_like_the_multivec_merge():
    """Rows shared by two ice sheets: (combined, scaled M) x = sum_sheets (M_sheet x_sheet) / sum_sheets wM = to_dense()."""
    from oracle import orc
    cfg = two_sheet_config()
    go = make_oracle(cfg)
    nA = cfg["gridA"].nA
    Ws = [go.matrix(i, cfg["elevmask"][i], "AvI", scale=False, correctA=True) for i in range(2)]
    rng = np.random.default_rng(0)
    xs = [rng.uniform(0, 1, (1, W.shape[1])) for W in Ws]
    idx = np.concatenate([W.dims[0] for W in Ws]); w = np.concatenate([W.wM for W in Ws])
    vals = np.concatenate([W.apply(x).T for W, x in zip(Ws, xs)])
    with np.errstate(divide="ignore"):
        _, dense = orc.multivec_to_dense(idx, w, vals, nA, fill=0.0)
    assert len(np.intersect1d(Ws[0].dims[0], Ws[1].dims[0])) > 0
    # the two sheets' I spaces differ, so combine them on a common (concatenated) input space by hand
    out = np.zeros(nA)
    sM = np.zeros(nA); np.add.at(sM, idx, w)
    for W, x in zip(Ws, xs):
        M1, _, _ = orc.combine_chunks([W], scale=False)
        xi = np.zeros(W.sparse_extent[1]); xi[W.dims[1]] = x[0]
        out += (M1.tocsr() @ xi)
    touched = sM > 0
    np.testing.assert_allclose(out[touched] / sM[touched], dense[0][touched], rtol=1e-12)


@pytest.mark.gpu
def test_gpu_combine_chunks_parity(cfg2):
    from icebin_b200 import api
    from oracle import orc
    em, sh, nA = cfg2["elevmask"][0], cfg2["sheets"][0], cfg2["gridA"].nA
    keep = multi.partition_cellsA(sh.ex_ij[:, 0], nA, 4)
    for spec in ("EvI", "AvI", "IvE"):
        co, cg = [], []
        for r in range(4):
            go, gg = make_oracle(cfg2), make_gpu(cfg2)
            go.filter_cellsA(keep[r]); gg.filter_cellsA(keep[r])
            co.append(go.matrix(0, em, spec, scale=False, correctA=True))
            cg.append(gg.regrid_matrices(sh.name, em, scale=False, correctA=True).matrix(spec))
        for scale in (True, False):
            Mo, (wio, wvo), (mio, mvo) = orc.combine_chunks(co, scale)
            Mg, (wig, wvg), (mig, mvg) = api.combine_chunks(cg, scale)
            assert np.array_equal(wig, wio) and np.array_equal(mig, mio)
            np.testing.assert_allclose(wvg, wvo, rtol=1e-12); np.testing.assert_allclose(mvg, mvo, rtol=1e-12)
            assert Mg.nnz == Mo.nnz
            # entries are not merged: compare as sorted (row, col, value) lists
            ko = np.lexsort((Mo.data, Mo.col, Mo.row)); kg = np.lexsort((Mg.data, Mg.col, Mg.row))
            assert np.array_equal(Mg.row[kg], Mo.row[ko]) and np.array_equal(Mg.col[kg], Mo.col[ko])
            np.testing.assert_allclose(Mg.data[kg], Mo.data[ko], rtol=1e-12)
    with pytest.raises(api.IceBinError):
        api.combine_chunks([make_gpu(cfg2).regrid_matrices(sh.name, em, scale=True).matrix("AvI")])     # scaled chunk
