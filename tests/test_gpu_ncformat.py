"""GPU: the global_ec output sequence (modele/global_ec.cpp:566-659) through the library -- matrix_d with shared, growing
dims, stored in Eigen format -- read back and applied on the GPU."""
import numpy as np
import pytest

from conftest import make_gpu, make_oracle
from icebin_b200 import api, ncformat as nf

pytestmark = pytest.mark.gpu


def test_global_ec_file_round_trip(cfg1, tmp_path):
    gg, go = make_gpu(cfg1), make_oracle(cfg1)
    sh, em, gA = cfg1["sheets"][0], cfg1["elevmask"][0], cfg1["gridA"]
    nhc = len(cfg1["hcdefs"])
    rm = gg.regrid_matrices(sh.name, em)
    dims = {"A": api.SparseSet(), "E": api.SparseSet(), "I": api.SparseSet()}
    ds = nf.write_global_ec(nf.Dataset("global_ec"), rm, ("AvI", "EvI", "IvE", "IvA", "AvE", "EvA"), dims,
                            dict(scale=True, correctA=True), shapes={"A": (gA.jm, gA.im), "E": (nhc, gA.jm, gA.im), "I": (sh.nx, sh.ny)})
    assert list(ds.vars["dimE"].attrs["shape"]) == [nhc, gA.jm, gA.im] and ds.vars["dimI"].attrs["description"] == "Fine-scale ('Ice') Grid"
    p = str(tmp_path / "global_ec.cdf5")
    ds.to_cdf5(p)
    back = nf.Dataset.from_cdf5(p)
    X = np.random.default_rng(0).uniform(0, 1, (3, sh.nI))
    for spec in ("AvI", "EvI"):
        H = nf.get_weighted(back, spec)
        Wo = go.matrix(0, em, spec, scale=True, correctA=True)
        assert np.array_equal(H.dims[0], Wo.dims[0]) and np.array_equal(H.dims[1], Wo.dims[1]) and H.sparse_extents[1] == sh.nI
        o = np.lexsort((H.row, H.col))
        assert np.array_equal(H.row[o], Wo.row) and np.array_equal(H.col[o], Wo.col) and np.allclose(H.val[o], Wo.val, rtol=1e-12, atol=0)
        Wd = H.to_device()
        Y = Wd.apply(X[:, H.dims[1]])
        assert np.allclose(Y, Wo.apply(X[:, Wo.dims[1]]), rtol=1e-12, atol=0)
    # IvE was written against the dims as they stood after AvI and EvI: same dimE / dimI ids as EvI's
    assert np.array_equal(nf.get_weighted(back, "IvE").dims[0], nf.get_weighted(back, "EvI").dims[1])
