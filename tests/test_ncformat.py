"""The on-disk twins of the reference's matrix files (icebin_b200/ncformat.py): Eigen format (matrix_formats.rst:9-63),
compressed format (:65-117), GCMRegridder::ncio (GCMRegridder.cpp:104-150), in CDL, CDF-5 and NPZ.  CPU: the matrices come
from the oracle; the GPU round trip is tests/test_gpu_ncformat.py."""
import os
import re

import numpy as np
import pytest

from conftest import make_oracle
from icebin_b200 import ncformat as nf

DOC_EXAMPLE = """netcdf __linear1 {
dimensions:
        dimB.dense_extent = 5 ;
        dimA.dense_extent = 6 ;
        BvA.M.nnz = 6 ;
        BvA.M.rank = 2 ;
variables:
        int64 dimB(dimB.dense_extent) ;
                dimB:sparse_extent = 40LL ;
        int64 dimA(dimA.dense_extent) ;
                dimA:sparse_extent = 50LL ;
        int BvA.info ;
                BvA.info:type = "EIGEN" ;
                BvA.info:conservative = 0 ;
                string BvA.info:dim_names = "BvA.dimB", "BvA.dimA" ;
        int64 BvA.M.info ;
                BvA.M.info:shape = 5LL, 6LL ;
                BvA.M.info:conservative = "f" ;
        int BvA.M.indices(BvA.M.nnz, BvA.M.rank) ;
        double BvA.M.values(BvA.M.nnz) ;
        double BvA.Mw(dimA.dense_extent) ;
        double BvA.wM(dimB.dense_extent) ;
data:
 dimB = 3, 7, 11, 20, 39 ;
 dimA = 0, 1, 17, 18, 30, 49 ;
 BvA.M.indices = 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5 ;
 BvA.M.values = 0.5, 0.5, 1, 1, 0.25, 0.75 ;
 BvA.Mw = 28.9, 1, 2, 3, 4, 5 ;
 BvA.wM = 1, 2, 3, 4, 5 ;
}
"""


def host_weighted(Wo, sparse_extents):
    return nf.HostWeighted(Wo.shape, Wo.row, Wo.col, Wo.val, Wo.wM, Wo.Mw, Wo.dims, sparse_extents, True)


def same(a, b):
    assert a.shape == b.shape and a.sparse_extents == b.sparse_extents and a.conservative == b.conservative
    for x, y in ((a.row, b.row), (a.col, b.col), (a.val, b.val), (a.wM, b.wM), (a.Mw, b.Mw), (a.dims[0], b.dims[0]), (a.dims[1], b.dims[1])):
        assert x.dtype == y.dtype and np.array_equal(x, y)


def test_reads_the_documented_example():
    """The CDL of matrix_formats.rst:17-42 (with data added) parses into the documented objects."""
    ds = nf.Dataset.from_cdl(DOC_EXAMPLE)
    assert ds.dims["BvA.M.nnz"] == 6 and ds.vars["dimA"].dtype == np.int64 and ds.vars["dimA"].attrs["sparse_extent"] == 50
    assert ds.vars["BvA.info"].attrs["dim_names"] == ["BvA.dimB", "BvA.dimA"] and ds.vars["BvA.M.info"].attrs["conservative"] == "f"
    W = nf.get_weighted(ds, "BvA")
    assert W.shape == (5, 6) and W.sparse_extents == [40, 50] and not W.conservative
    # "if BvA.Mw[0] == 28.9 and dimA[0] == 17 ..." (the doc's sentence, with this file's dimA): dense 2 <-> sparse 17
    assert W.dims[1][2] == 17 and W.Mw[0] == 28.9
    coo = W.to_coo()
    assert coo.shape == (40, 50) and coo.nnz == 6 and coo.tocsr()[39, 49] == 0.75


@pytest.mark.parametrize("fmt", ["cdl", "cdf5", "npz"])
def test_eigen_format_round_trip(cfg1, fmt, tmp_path):
    go = make_oracle(cfg1)
    em, sh, gA = cfg1["elevmask"][0], cfg1["sheets"][0], cfg1["gridA"]
    nhc = len(cfg1["hcdefs"])
    ext = {"A": gA.nA, "E": gA.nA * nhc, "I": sh.nI}
    ds = nf.Dataset("global_ec")
    Ws = {}
    for spec in ("AvI", "EvI", "IvE"):
        Wo = go.matrix(0, em, spec, scale=True, correctA=True)
        Ws[spec] = host_weighted(Wo, [ext[spec[0]], ext[spec[2]]])
        nf.put_weighted(ds, Ws[spec], spec, ["dim" + spec[0], "dim" + spec[2]])
    for k in "AEI":
        src = Ws["AvI"] if k in "AI" else Ws["EvI"]
        nf.put_dim(ds, "dim" + k, src.dims[0] if k in "AE" else src.dims[1], ext[k], description="grid " + k)
    # the reference's names (matrix_formats.rst:17-42)
    for name in ("EvI.info", "EvI.M.info", "EvI.M.indices", "EvI.M.values", "EvI.Mw", "EvI.wM", "dimE", "dimI"):
        assert name in ds.vars, name
    assert ds.vars["EvI.M.indices"].dims == ("EvI.M.nnz", "EvI.M.rank") and ds.vars["EvI.wM"].dims == ("dimE.dense_extent",)
    assert ds.vars["EvI.info"].attrs["type"] == "EIGEN" and ds.vars["dimE"].attrs["sparse_extent"] == ext["E"]
    path = str(tmp_path / ("m." + fmt))
    if fmt == "cdl":
        text = ds.to_cdl(path)
        assert re.search(r"int64 dimE\(dimE.dense_extent\) ;\s+dimE:sparse_extent = %dLL ;" % ext["E"], text)
        assert 'string EvI.info:dim_names = "dimE", "dimI" ;' in text
        back = nf.Dataset.from_cdl(path)
    elif fmt == "cdf5":
        ds.to_cdf5(path)
        assert open(path, "rb").read(4) == b"CDF\x05" and os.path.getsize(path) % 4 == 0
        back = nf.Dataset.from_cdf5(path)
    else:
        ds.to_npz(path)
        back = nf.Dataset.from_npz(path + ("" if path.endswith(".npz") else ".npz"))
    assert list(back.dims.items()) == list(ds.dims.items()) and list(back.vars) == list(ds.vars)
    for spec in Ws:
        same(nf.get_weighted(back, spec), Ws[spec])


def test_compressed_format_round_trip(cfg1, tmp_path):
    """combine_global_ec's output (combine_global_ec.cpp:256-261): sparse-indexed M, wM, Mw as zlib-compressed tensors."""
    go = make_oracle(cfg1)
    gA, sh = cfg1["gridA"], cfg1["sheets"][0]
    Wo = go.matrix(0, cfg1["elevmask"][0], "AvI", scale=True, correctA=True)
    r, c = Wo.dims[0][Wo.row], Wo.dims[1][Wo.col]
    ds = nf.Dataset("combined")
    nf.put_compressed(ds, "AvI", (gA.nA, sh.nI), r, c, Wo.val, Wo.dims[0], Wo.wM, Wo.dims[1], Wo.Mw)
    for name, dims in (("AvI.M.indices", ("AvI.M.indices.zsize",)), ("AvI.wM.values", ("AvI.wM.values.zsize",))):
        assert ds.vars[name].dtype == np.uint8 and ds.vars[name].dims == dims
    assert ds.vars["AvI.info"].attrs["type"] == "COMPRESSED" and ds.vars["AvI.M.info"].attrs["format"] == "ZARRAY"
    assert int(ds.vars["AvI.M.info"].attrs["nnz"]) == Wo.nnz and list(ds.vars["AvI.M.info"].attrs["shape"]) == [gA.nA, sh.nI]
    p = str(tmp_path / "c.cdf5")
    ds.to_cdf5(p)
    got = nf.get_compressed(nf.Dataset.from_cdf5(p), "AvI")
    assert got["shape"] == (gA.nA, sh.nI) and got["conservative"]
    assert np.array_equal(got["M"][0], r) and np.array_equal(got["M"][1], c) and np.array_equal(got["M"][2], Wo.val)
    assert np.array_equal(got["wM"][0], Wo.dims[0]) and np.array_equal(got["wM"][1], Wo.wM) and np.array_equal(got["Mw"][1], Wo.Mw)
    assert ds.vars["AvI.M.values"].data.size < 8 * Wo.nnz        # it does compress


def test_gcm_regridder_round_trip(cfg1, tmp_path):
    """GCMRegridder_Standard::ncio: what the hot path reads of the regridder survives the file, under the reference's names;
    the regridder rebuilt from the file gives the same matrices."""
    from conftest import make_oracle as mk
    gA, sh, hc = cfg1["gridA"], cfg1["sheets"][0], cfg1["hcdefs"]
    ds = nf.put_gcm(nf.Dataset("icebin_in"), gA, hc, [sh], correctA=True)
    for name in ("m.info", "m.hcdefs", "m.agridA.dim", "m.agridA.native_area", "m.%s.info" % sh.name, "m.%s.gridA_proj_area" % sh.name,
                 "m.%s.agridI.dim" % sh.name, "m.%s.agridI.centroid_xy" % sh.name, "m.%s.aexgrid.indices" % sh.name,
                 "m.%s.aexgrid.overlaps" % sh.name):
        assert name in ds.vars, name
    assert ds.vars["m.info"].attrs["sheets"] == [sh.name] and ds.dims["m.%s.aexgrid.nindices" % sh.name] == 2 * sh.nX
    p = str(tmp_path / "in.cdl")
    ds.to_cdl(p)
    gA2, hc2, sheets2, correctA = nf.get_gcm(nf.Dataset.from_cdl(p))
    s2 = sheets2[0]
    assert correctA and np.array_equal(hc2, hc) and np.array_equal(gA2.dim, gA.dim) and np.array_equal(gA2.native_area, gA.native_area)
    assert np.array_equal(s2.ex_ij, sh.ex_ij) and np.array_equal(s2.ex_area, sh.ex_area) and np.array_equal(s2.centroid_xy, sh.centroid_xy)
    a = mk(cfg1).matrix(0, cfg1["elevmask"][0], "EvI", scale=True, correctA=True)
    b = mk(dict(gridA=gA2, sheets=sheets2, hcdefs=hc2)).matrix(0, cfg1["elevmask"][0], "EvI", scale=True, correctA=True)
    assert np.array_equal(a.val, b.val) and np.array_equal(a.row, b.row) and np.array_equal(a.dims[0], b.dims[0])
