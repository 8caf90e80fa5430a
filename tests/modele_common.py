"""A small ModelE-shaped case: ocean grid 288x180 (1 x 1.25 degrees), atmosphere grid 144x90, a 50 km ice sheet overlapped
with the OCEAN grid, and the two ocean masks (ice model: fractions; ModelE: 0 / 1)."""
import numpy as np

from icebin_b200 import synth

SPEC_O = (288, 180, 0.0, 60.0)


def case(seed=0, nhc=10):
    hc = synth.hcdefs_example1(nhc)
    gO, sh = synth.make_gcm_and_sheet("greenland_50km", anchor=(80, 120), im=288, jm=180, cell=(synth.A_W / 2, synth.A_H / 2), margin=2)
    em = synth.make_elevmask(sh, hc[-1], seed)
    rng = np.random.default_rng(seed)
    nO = gO.nA
    # ice model's view: continent with fractional coast; ModelE's view: rounded, and a few cells it calls ocean outright
    fp = np.clip(rng.uniform(-0.6, 0.9, nO), 0.0, 1.0)
    fm = (fp > 0.6).astype(np.float64)
    touched = np.unique(sh.ex_ij[:, 0])
    fp[touched[::17]] = 1.0                     # all ocean for the ice model: nothing to extend there
    fm[touched[::11]] = 1.0                     # ModelE says ocean where the ice model has ice
    return dict(gridA=gO, sheets=[sh], hcdefs=hc, elevmask=[em]), fp, fm
