"""GPU: the device-resident multi-sheet merge kernels (csrc/merge.cu) on one GPU holding two sheets; the all-reduce
between the pack and the unpack is emulated by adding the two packed buffers (a multi-rank NCCL run is bench.py --gpus N)."""
import numpy as np
import pytest
import torch

from conftest import make_gpu
from icebin_b200 import api, multi, synth
from icebin_b200._lib import check, lib

pytestmark = pytest.mark.gpu


def test_two_sheets_share_boundary_cells():
    hc = synth.hcdefs_conserv(40)
    im, jm = 144, 90
    ids = np.arange(im * jm, dtype=np.int64)
    gA = synth.GridA(im, jm, im * jm, ids, np.full(im * jm, synth.A_W * synth.A_H))
    sheets = [synth.make_gcm_and_sheet("greenland_20km", anchor=(10 + 10 * r, 40), gridA=gA, name="s%d" % r)[1] for r in range(2)]
    gcm = api.GCMRegridder(gA, hc, True)
    dev = torch.device("cuda")
    nvar = 6
    W, Y, num = [], [], []
    for r, sh in enumerate(sheets):
        gcm.add_sheet(sh.name, sh)
        em = synth.make_elevmask(sh, hc[-1], seed=r, frac=0.95)
        M = gcm.regrid_matrices(sh.name, em, scale=False, correctA=True).matrix("AvI")
        X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device=dev)
        y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device=dev)
        M.apply_dev(X.data_ptr(), nvar, nvar, y.data_ptr(), nvar)
        torch.cuda.synchronize()
        W.append(M); Y.append(y); num.append(y.cpu().numpy().copy())
    idsets = [w.dim(0) for w in W]
    shared = np.intersect1d(idsets[0], idsets[1])
    assert len(shared) > 0, "the two sheets were meant to overlap in a column of A cells"
    # per-'rank' plans (computed as plan_shared would with an all-reduce of the membership counts)
    plans = []
    for r in range(2):
        pos = {int(s): k for k, s in enumerate(idsets[r])}
        loc = np.array([pos[int(s)] for s in shared], np.int32)
        slot = np.full(len(idsets[r]), -1, np.int32); slot[loc] = np.arange(len(shared))
        plans.append((torch.from_numpy(loc).to(dev), torch.from_numpy(slot).to(dev)))
    L = lib()
    bufs = []
    for r in range(2):
        b = torch.empty((len(shared), nvar), dtype=torch.float64, device=dev)
        check(L.ib_merge_pack_dev(Y[r].data_ptr(), nvar, nvar, plans[r][0].data_ptr(), len(shared), b.data_ptr(), None))
        bufs.append(b)
    torch.cuda.synchronize()
    tot = bufs[0] + bufs[1]                                          # == NCCL all-reduce(sum)
    wl = [torch.from_numpy(w.get_weights(0)).to(dev) for w in W]
    wtot = wl[0][plans[0][0].long()] + wl[1][plans[1][0].long()]
    for r in range(2):
        check(L.ib_merge_unpack_dev(Y[r].data_ptr(), nvar, nvar, W[r].shape[0], plans[r][1].data_ptr(), tot.data_ptr(),
                                    wtot.data_ptr(), wl[r].data_ptr(), None))
    torch.cuda.synchronize()
    ref = multi.merge_reference([(idsets[r], W[r].get_weights(0), num[r]) for r in range(2)], gA.nA)
    for r in range(2):
        np.testing.assert_allclose(Y[r].cpu().numpy(), ref[idsets[r]], rtol=1e-14)
