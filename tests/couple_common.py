"""Shared set-up of the coupling-step tests: two ice sheets on one GCM grid whose GCM cells overlap (so that they share
elevation-class cells), and an ice-model step that changes the elevation mask."""
import numpy as np

from icebin_b200 import synth


def two_sheet_config(seed=0):
    hc = synth.hcdefs_conserv(40)
    im, jm = 144, 90
    ids = np.arange(im * jm, dtype=np.int64)
    full = synth.GridA(im, jm, im * jm, ids, (synth.A_W * synth.A_H) * (1.0 + 0.08 * np.sin(0.37 * (ids % im) + 0.21 * (ids // im))))
    _, s0 = synth.make_gcm_and_sheet("greenland_20km", anchor=(40, 60), gridA=full, name="north")
    _, s1 = synth.make_gcm_and_sheet("greenland_50km", anchor=(45, 63), gridA=full, name="south")     # overlaps s0's GCM cells
    em = [synth.make_elevmask(s0, hc[-1], seed), synth.make_elevmask(s1, hc[-1], seed + 1)]
    return dict(gridA=full, sheets=[s0, s1], hcdefs=hc, elevmask=em)


def stepped_elevmask(em, seed):
    """What an ice-model step does to elevmaskI: thinning/thickening, margin retreat and advance."""
    rng = np.random.default_rng(seed)
    out = em.copy()
    ice = ~np.isnan(em)
    out[ice] = np.clip(em[ice] - 35.0 + rng.normal(0, 20.0, ice.sum()), 0.0, None)
    idx = np.nonzero(ice)[0]
    out[rng.choice(idx, max(1, len(idx) // 50), replace=False)] = np.nan           # retreat
    bare = np.nonzero(~ice)[0]
    if len(bare):
        out[rng.choice(bare, max(1, len(bare) // 80), replace=False)] = 12.5        # advance onto bare cells
    return out


def var_transform(nin, nout, seed=0):
    """A VarTransformer matrix (apply_scalars(..., 'T')): sparse [nin, nout] with a few entries per output + bias."""
    import scipy.sparse
    rng = np.random.default_rng(seed)
    M = np.zeros((nin, nout))
    for k in range(nout):
        rows = rng.choice(nin, size=min(nin, 1 + k % 3), replace=False)
        M[rows, k] = rng.uniform(-2, 2, len(rows))
    b = rng.uniform(-1, 1, nout)
    b[::3] = 0.0
    return scipy.sparse.csc_matrix(M), b
