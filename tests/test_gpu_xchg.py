"""GPU: the peer-memory merge (csrc/xchg.cu, ib_xchg_*) with the ranks of several sheets living in ONE process on one GPU,
driven one after the other (all ranks publish, then all ranks gather) so that no kernel ever waits for a kernel that is
not running: flags, double buffering over many steps, owner masks, graph replay.  The multi-process IPC path over NVLink
is exercised by tools/xchg_ranks.py / bench.py --gpus N on a multi-GPU box."""
import numpy as np
import pytest
import torch

from icebin_b200 import api, multi, synth

pytestmark = pytest.mark.gpu


def three_sheets():
    hc = synth.hcdefs_conserv(40)
    im, jm = 144, 90
    ids = np.arange(im * jm, dtype=np.int64)
    gA = synth.GridA(im, jm, im * jm, ids, np.full(im * jm, synth.A_W * synth.A_H))
    # neighbouring sheets overlap in a column of A cells; sheet 2 also meets sheet 0's far side: cells with 2 and 3 owners
    anchors = [(10, 40), (20, 40), (15, 40)]
    sheets = [synth.make_gcm_and_sheet("greenland_20km", anchor=a, gridA=gA, name="s%d" % r)[1] for r, a in enumerate(anchors)]
    gcm = api.GCMRegridder(gA, hc, True)
    mats = []
    for r, sh in enumerate(sheets):
        gcm.add_sheet(sh.name, sh)
        em = synth.make_elevmask(sh, hc[-1], seed=r, frac=0.95)
        rm = gcm.regrid_matrices(sh.name, em, scale=False, correctA=True)
        mats.append((rm.matrix("EvI"), rm.matrix("AvI")))
    return gcm, mats


@pytest.mark.parametrize("nvar,scaled", [(6, False), (65, False), (64, True)])
def test_three_ranks_in_one_process(nvar, scaled):
    """scaled=False: the coupler's unscaled matrices, every row divided by its weight in the merge; scaled=True: scaled
    matrices, only the shared cells travel (as Y * wM) and are rewritten."""
    gcm, mats = three_sheets()
    dev = torch.device("cuda")
    ctx = multi.PeerMerge.local([[(gcm.nE, E), (gcm.nA, A)] for E, A in mats], nvar, scaled=scaled)
    assert ctx[0].nshared > 0 and any((p["owners"] == 7).any() for p in ctx[0].parts), "some cell should have three owners"
    rng = np.random.default_rng(0)
    for step in range(5):                       # > 2 steps: both buffers are reused
        Ys, num = [], []
        for r, (E, A) in enumerate(mats):
            X = torch.from_numpy(rng.uniform(0, 1, (E.shape[1], nvar))).to(dev)
            ye = torch.empty((E.shape[0], nvar), dtype=torch.float64, device=dev)
            ya = torch.empty((A.shape[0], nvar), dtype=torch.float64, device=dev)
            E.apply_dev(X.data_ptr(), nvar, nvar, ye.data_ptr(), nvar)
            A.apply_dev(X.data_ptr(), nvar, nvar, ya.data_ptr(), nvar)
            torch.cuda.synchronize()
            num.append([ye.cpu().numpy().copy(), ya.cpu().numpy().copy()])
            if scaled:          # what the scaled matrices would have produced
                ye /= torch.from_numpy(E.get_weights(0)).to(dev)[:, None]
                ya /= torch.from_numpy(A.get_weights(0)).to(dev)[:, None]
            Ys.append([ye, ya])
        for r in range(3):
            ctx[r].publish(Ys[r])
        torch.cuda.synchronize()
        for r in range(3):
            ctx[r].gather(Ys[r])
        torch.cuda.synchronize()
        for p, n_sparse in ((0, gcm.nE), (1, gcm.nA)):
            ref = multi.merge_reference([(mats[r][p].dim(0), mats[r][p].get_weights(0), num[r][p]) for r in range(3)], n_sparse)
            for r in range(3):
                np.testing.assert_allclose(Ys[r][p].cpu().numpy(), ref[mats[r][p].dim(0)], rtol=1e-13 if not scaled else 1e-12, atol=0)
        # every owner of a shared cell holds the same bits
        for p in range(2):
            a, b = mats[0][p].dim(0), mats[1][p].dim(0)
            common, ia, ib_ = np.intersect1d(a, b, return_indices=True)
            assert len(common) and np.array_equal(Ys[0][p].cpu().numpy()[ia], Ys[1][p].cpu().numpy()[ib_])
    for c in ctx:
        c.status()
        assert c.steps_done() == 5


def test_plan_rejects_bad_input():
    from icebin_b200._lib import lib
    import ctypes as C
    ids = np.array([1, 99], np.int64)
    ptr = (C.c_void_p * 1)(ids.ctypes.data)
    n = (C.c_int64 * 1)(2)
    assert lib().ib_xchg_plan(1, ptr, n, 10, 0, None, None, None) < 0
    assert "outside the grid" in lib().ib_last_error().decode()
