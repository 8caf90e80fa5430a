"""Shared set-up of the exchange-grid generation tests: GCM cells as polygons in the ice grid's plane."""
import numpy as np

from icebin_b200 import synth


def rect_polys(shape="greenland_50km", anchor=(40, 60), im=144):
    """The rectangles synth.make_gcm_and_sheet models GCM cells with, as 4-vertex polygons + their iA, and the ice edges."""
    nx, ny, dxkm = synth.SHEET_SHAPES[shape]
    dx = dxkm * synth.KM
    xe = np.arange(nx + 1) * dx
    ye = np.arange(ny + 1) * dx
    npx = int(np.floor((xe[-1] - synth.A_OFF[0]) / synth.A_W)) + 1
    npy = int(np.floor((ye[-1] - synth.A_OFF[1]) / synth.A_H)) + 1
    polys, iA = [], []
    for q in range(npy):
        for p in range(npx):
            x0 = synth.A_OFF[0] + p * synth.A_W; x1 = synth.A_OFF[0] + (p + 1) * synth.A_W
            y0 = synth.A_OFF[1] + q * synth.A_H; y1 = synth.A_OFF[1] + (q + 1) * synth.A_H
            polys.append(np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]]))
            iA.append((anchor[1] + q) * im + anchor[0] + p)
    return polys, np.array(iA, np.int64), xe, ye


def warped_mesh(ncx=7, ncy=5, nsub=6, extent=(0.0, 900e3, 0.0, 700e3), seed=0):
    """A mesh of 'projected lon-lat cells': a regular mesh pushed through a smooth map, every side subdivided into nsub
    segments BEFORE the map (as the reference does, so sides are curved and cells slightly non-convex).  Neighbouring
    cells share their side vertices exactly, so the polygons tile the mapped region."""
    rng = np.random.default_rng(seed)
    a, b = rng.uniform(0.02, 0.05, 2)

    def fmap(u, v):       # u, v in [0, 1]
        x = extent[0] + (extent[1] - extent[0]) * (0.1 + 0.8 * (u + a * np.sin(2.1 * v + 0.3) * (1 - u) * u * 4))
        y = extent[2] + (extent[3] - extent[2]) * (0.1 + 0.8 * (v + b * np.sin(1.7 * u + 1.1) + 0.03 * u * u))
        return x, y
    us = np.linspace(0, 1, ncx * nsub + 1)
    vs = np.linspace(0, 1, ncy * nsub + 1)
    polys, iA = [], []
    for j in range(ncy):
        for i in range(ncx):
            iu = np.arange(i * nsub, (i + 1) * nsub + 1); iv = np.arange(j * nsub, (j + 1) * nsub + 1)
            side = [(us[k], vs[iv[0]]) for k in iu[:-1]] + [(us[iu[-1]], vs[k]) for k in iv[:-1]] + \
                   [(us[k], vs[iv[-1]]) for k in iu[:0:-1]] + [(us[iu[0]], vs[k]) for k in iv[:0:-1]]
            uv = np.array(side)
            x, y = fmap(uv[:, 0], uv[:, 1])
            polys.append(np.stack([x, y], axis=1))
            iA.append(j * 100 + i)
    return polys, np.array(iA, np.int64)


def shoelace(p):
    x, y = p[:, 0], p[:, 1]
    return 0.5 * abs(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))
