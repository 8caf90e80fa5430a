import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cfg1():
    from icebin_b200 import synth
    return synth.config(1)


@pytest.fixture(scope="session")
def cfg2():
    from icebin_b200 import synth
    return synth.config(2)


def make_oracle(cfg, interp_style=0):
    from oracle import orc
    gA = cfg["gridA"]
    g = orc.Gcm(gA.nA, gA.dim, gA.native_area, cfg["hcdefs"])
    for sh in cfg["sheets"]:
        g.add_sheet(sh.nI, sh.dimI, sh.native_areaI, sh.centroid_xy, sh.ex_ij, sh.ex_area, sh.gridA_proj_area, interp_style)
    return g


def make_gpu(cfg, interp_style="Z_INTERP"):
    from icebin_b200 import api
    g = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
    for sh in cfg["sheets"]:
        g.add_sheet(sh.name, sh, interp_style)
    return g
