"""Generates tests/golden/*.npz.  Run in the BUILD container (needs /root/reference); the outputs are committed.

  1. elevmask_searise20km.npz -- the reference's own fixture tests/test_conserv/elev_mask.cdl (CDL text dump of a PISM
     state, dims time=1,x=76,y=141) turned into elevmaskI by the rule of slib/icebin/ElevMask.cpp:56-71
     (ice: topg+thk; bare land: topg for emI_land / NaN for emI_ice; ocean: NaN).
  2. oracle_golden_cfg1.npz -- regression vectors produced by oracle/oracle.cpp on config 1 (the reference holds no
     golden vectors for the six matrices; these pin the ORACLE against accidental change, not against upstream).
  3. hntr_kat.npz -- Hntr partition tables + regrid results of the oracle for the grid pairs of tests/test_hntr.cpp:174-184.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def parse_cdl(path):
    txt = open(path).read()
    data = txt[txt.index("data:"):]
    out = {}
    for name in ("mask", "thk", "topg"):
        m = re.search(r"\n\s*%s\s*=(.*?);" % name, data, re.S)
        vals = [v for v in re.split(r"[,\s]+", m.group(1).strip()) if v]
        out[name] = np.array([float(v) for v in vals])
    return out


def elevmask():
    d = parse_cdl("/root/reference/tests/test_conserv/elev_mask.cdl")
    nx, ny = 76, 141
    for k in d:
        assert d[k].size == nx * ny, (k, d[k].size)
    mask, thk, topg = d["mask"], d["thk"], d["topg"]
    # IceMask (ElevMask.hpp): 0 ice_free_bedrock, 2 grounded_ice, 3 floating_ice, 4 ice_free_ocean
    ice = (mask == 2) | (mask == 3)
    em_ice = np.where(ice, topg + thk, np.nan)
    em_land = np.where(ice, topg + thk, np.where(mask == 0, topg, np.nan))
    np.savez_compressed(os.path.join(OUT, "elevmask_searise20km.npz"), nx=nx, ny=ny, mask=mask.astype(np.int8),
                        emI_ice=em_ice, emI_land=em_land)
    print("elevmask: ice cells", int(ice.sum()), "land", int((mask == 0).sum()), "max elev", np.nanmax(em_ice))


def oracle_cfg1():
    from icebin_b200 import synth
    from oracle import orc
    cfg = synth.config(1)
    gA, sh, hc, em = cfg["gridA"], cfg["sheets"][0], cfg["hcdefs"], cfg["elevmask"][0]
    g = orc.Gcm(gA.nA, gA.dim, gA.native_area, hc)
    g.add_sheet(sh.nI, sh.dimI, sh.native_areaI, sh.centroid_xy, sh.ex_ij, sh.ex_area, sh.gridA_proj_area)
    out = {}
    for spec in ("AvI", "IvA", "EvI", "IvE", "EvA", "AvE"):
        W = g.matrix(0, em, spec, scale=True, correctA=True)
        out[spec + "_row"], out[spec + "_col"], out[spec + "_val"] = W.row, W.col, W.val
        out[spec + "_wM"], out[spec + "_Mw"], out[spec + "_d0"], out[spec + "_d1"] = W.wM, W.Mw, W.dims[0], W.dims[1]
    np.savez_compressed(os.path.join(OUT, "oracle_golden_cfg1.npz"), **out)


def hntr():
    from oracle import orc
    pairs = {"g8_to_g4": ((8, 4, 0., 2700.), (16, 8, 0., 1350.)), "g1qx1_to_g2hx2": ((144, 90, 0., 120.), (288, 180, 0., 60.)),
             "ghxh_to_g1qx1": ((288, 180, 0., 60.), (720, 360, 0., 30.)), "g1x1_to_g1qx1": ((288, 180, 0., 60.), (360, 180, 0., 60.))}
    out = {}
    for name, (B, A) in pairs.items():
        h = orc.Hntr(B, A, 0.0)
        for t in ("FMIN", "FMAX", "GMIN", "GMAX", "IMIN", "IMAX", "JMIN", "JMAX"):
            out["%s_%s" % (name, t)] = h.get(t)
        src = np.random.default_rng(5).uniform(0, 1, A[0] * A[1])
        out[name + "_src_seed5_B"] = h.regrid(None, src)
    np.savez_compressed(os.path.join(OUT, "hntr_kat.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    elevmask()
    oracle_cfg1()
    hntr()
    print("wrote", sorted(os.listdir(OUT)))
