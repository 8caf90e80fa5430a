"""One process, two devices: kernel attributes (opt-in shared memory) are per device, and a library that sets them once
would fail -- or silently fall back -- on the second GPU.  Needs a box with two GPUs (`gpurun --gpus 2`); skipped otherwise."""
import numpy as np
import pytest

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def test_every_kernel_family_on_the_second_device(cfg2):
    import torch
    from icebin_b200 import api
    from icebin_b200._lib import check, lib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    go = make_oracle(cfg2)
    em = cfg2["elevmask"][0]
    sigma = (50e3, 50e3, 100.0)
    X = {}
    ref = {}
    for spec in ("EvI", "AvI", "IvE", "IvA"):
        Wo = go.matrix(0, em, spec, scale=True, correctA=True)
        X[spec] = np.random.default_rng(7).uniform(0, 1, (128, Wo.shape[1]))
        ref[spec] = Wo.apply(X[spec], force_conservation=False)
    Ws = go.matrix(0, em, "IvE", scale=True, correctA=True, sigma=sigma)
    ref["IvE_s"] = Ws.apply(X["IvE"], force_conservation=True)
    try:
        for dev in (0, 1, 0):
            check(lib().ib_set_device(dev))
            with torch.cuda.device(dev):
                gg = make_gpu(cfg2)
                rm = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True)
                for spec in ("EvI", "AvI", "IvE", "IvA"):      # one-pass stream kernel, distinct-column group kernel
                    np.testing.assert_allclose(rm.matrix(spec).apply(X[spec], force_conservation=False), ref[spec], rtol=1e-12, atol=0)
                IvEs = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True, sigma=sigma).matrix("IvE")
                np.testing.assert_allclose(IvEs.apply(X["IvE"], force_conservation=True), ref["IvE_s"], rtol=1e-12, atol=0)   # DMMA group kernel
                del gg, rm, IvEs
    finally:
        check(lib().ib_set_device(0))
