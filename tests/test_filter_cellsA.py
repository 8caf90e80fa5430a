"""filter_cellsA (GCMRegridder.cpp:154-164): domain restriction of the exchange grids to a set of GCM cells.
CPU: the oracle's filtered matrices are the corresponding blocks of the unfiltered ones.  GPU: parity with the oracle and
the two-domain decomposition adding up to the whole."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_gpu, make_oracle


def keep_mask(cfg, frac=0.5):
    """Keep the western part of the GCM cells the sheet touches (a rank's longitude band)."""
    gA, sh = cfg["gridA"], cfg["sheets"][0]
    used = np.unique(sh.ex_ij[:, 0])
    cols = used % gA.im
    cut = np.sort(np.unique(cols))[int(len(np.unique(cols)) * frac)]
    keep = np.zeros(gA.nA, bool)
    keep[used[cols < cut]] = True
    return keep


def sparse_of(W):
    return sp.csr_matrix((W.val, (W.dims[0][W.row], W.dims[1][W.col])), shape=W.sparse_extent)


def test_oracle_filtered_matrices_are_blocks_of_the_full_ones(cfg1):
    em = cfg1["elevmask"][0]
    full, part = make_oracle(cfg1), make_oracle(cfg1)
    keep = keep_mask(cfg1)
    nX0 = part.nX(0)
    part.filter_cellsA(keep)
    assert 0 < part.nX(0) < nX0 and part.nX(0) == int(keep[cfg1["sheets"][0].ex_ij[:, 0]].sum())
    nA = cfg1["gridA"].nA
    keepE = np.tile(keep, len(cfg1["hcdefs"]))
    for spec, rows in (("AvI", keep), ("EvI", keepE)):
        F, P = sparse_of(full.matrix(0, em, spec, True, True)), sparse_of(part.matrix(0, em, spec, True, True))
        # a GCM cell's row depends only on its own exchange cells: kept rows are unchanged, bit for bit
        assert (F[rows] != P[rows]).nnz == 0 and P[~rows].nnz == 0
    # unscaled I-side rows split between the domains: IvA(full) = IvA(kept) + IvA(dropped)
    other = make_oracle(cfg1)
    other.filter_cellsA(~keep)
    F = sparse_of(full.matrix(0, em, "IvA", False, False))
    S = sparse_of(part.matrix(0, em, "IvA", False, False)) + sparse_of(other.matrix(0, em, "IvA", False, False))
    assert abs(F - S).max() <= 1e-12 * abs(F).max()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["cfg1", "cfg2"])
def test_gpu_filter_cellsA_parity(which, request):
    cfg = request.getfixturevalue(which)
    em = cfg["elevmask"][0]
    keep = keep_mask(cfg, 0.4)
    go, gg = make_oracle(cfg), make_gpu(cfg)
    name = cfg["sheets"][0].name
    before = gg.regrid_matrices(name, em).matrix("AvI")          # built before the restriction: unaffected by it
    go.filter_cellsA(keep)
    gg.filter_cellsA(keep)
    assert gg.nX(name) == go.nX(0)
    assert np.array_equal(gg.wA(name, "native") != 0, keep & np.isin(np.arange(cfg["gridA"].nA), cfg["gridA"].dim))
    for spec in ("AvI", "EvI", "IvE", "IvA", "EvA", "AvE", "XvE"):
        Wo = go.matrix(0, em, spec, True, True)
        Wg = gg.regrid_matrices(name, em, scale=True, correctA=True).matrix(spec)
        assert Wg.shape == Wo.shape and Wg.nnz == Wo.nnz, spec
        assert np.array_equal(Wg.dim(0), Wo.dims[0]) and np.array_equal(Wg.dim(1), Wo.dims[1]), spec
        r, c, v = Wg.to_coo_dense()
        o = np.lexsort((Wo.col, Wo.row))
        assert np.array_equal(r, Wo.row[o]) and np.array_equal(c, Wo.col[o]), spec
        np.testing.assert_allclose(v, Wo.val[o], rtol=1e-12, err_msg=spec)
        np.testing.assert_allclose(Wg.get_weights(0), Wo.wM, rtol=1e-12)
        np.testing.assert_allclose(Wg.get_weights(1), Wo.Mw, rtol=1e-12)
    assert before.nnz > gg.regrid_matrices(name, em).matrix("AvI").nnz
    # a predicate works like the reference's std::function<bool(long)>
    g2 = make_gpu(cfg)
    g2.filter_cellsA(lambda iA: bool(keep[iA]))
    assert g2.nX(name) == go.nX(0)


def test_partition_cellsA_is_a_balanced_partition(cfg2):
    from icebin_b200 import multi
    sh, nA = cfg2["sheets"][0], cfg2["gridA"].nA
    used = np.unique(sh.ex_ij[:, 0])
    for world in (1, 2, 3, 8):
        keep = multi.partition_cellsA(sh.ex_ij[:, 0], nA, world)
        assert keep.shape == (world, nA) and np.array_equal(keep.sum(0) > 0, np.isin(np.arange(nA), used))
        assert keep.sum(0).max() == 1                                   # every used GCM cell on exactly one rank
        load = np.array([keep[r][sh.ex_ij[:, 0]].sum() for r in range(world)])
        assert load.sum() == sh.nX and load.max() <= sh.nX / world + np.bincount(sh.ex_ij[:, 0]).max()
        for r in range(world):                                          # contiguous runs of GCM cells
            ids = np.nonzero(keep[r])[0]
            if len(ids):
                assert np.array_equal(ids, used[(used >= ids[0]) & (used <= ids[-1])])


@pytest.mark.gpu
def test_gpu_sharded_build_concatenates_to_the_full_matrices(cfg2):
    """One sheet's build cut over 4 domains by GCM-cell ranges (each domain is what one GPU would hold): the AvI / EvI rows
    of the domains are disjoint and, put together, are the full matrices bit for bit; the unscaled IvA add up."""
    from icebin_b200 import multi
    sh, em, nA = cfg2["sheets"][0], cfg2["elevmask"][0], cfg2["gridA"].nA
    full = make_gpu(cfg2).regrid_matrices(sh.name, em, scale=True, correctA=True)
    keep = multi.partition_cellsA(sh.ex_ij[:, 0], nA, 4)

    def sparse(W):
        r, c, v = W.to_coo_dense()
        return sp.csr_matrix((v, (W.dim(0)[r], W.dim(1)[c])), shape=W.sparse_extent)
    parts = []
    for r in range(4):
        g = make_gpu(cfg2)
        g.filter_cellsA(keep[r])
        parts.append(g.regrid_matrices(sh.name, em, scale=True, correctA=True))
    for spec in ("AvI", "EvI"):
        F = sparse(full.matrix(spec))
        S = sum(sparse(p.matrix(spec)) for p in parts)
        assert (F != S).nnz == 0, spec
        rows = [np.unique(p.matrix(spec).dim(0)) for p in parts]
        assert sum(len(x) for x in rows) == len(np.unique(np.concatenate(rows)))        # disjoint row sets
    fu = make_gpu(cfg2).regrid_matrices(sh.name, em, scale=False, correctA=False)
    F = sparse(fu.matrix("IvA"))
    S = None
    for r in range(4):
        g = make_gpu(cfg2)
        g.filter_cellsA(keep[r])
        M = sparse(g.regrid_matrices(sh.name, em, scale=False, correctA=False).matrix("IvA"))
        S = M if S is None else S + M
    assert abs(F - S).max() <= 1e-12 * abs(F).max()
