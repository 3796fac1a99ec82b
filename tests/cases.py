"""Seeded synthetic inputs shared by the parity tests and bench.py."""
import numpy as np

STEEL = dict(gg=(82000.0, 82000.0), poiss=(0.28, 0.28))
# grid of perfc_test/spence71_8281pt.inp:12 with its quadratic undeformed distance (:15)
HERTZ91 = dict(mx=91, my=91, xl=-4.55, yl=-4.55, dx=0.1, dy=0.1, ibase=1,
               prmudf=[0.004116, 0.0, 0.004116, 0.0, 0.0, 0.0])
# examples/cattaneo.inp:37-46 (second case)
CATTANEO2 = dict(mx=19, my=19, xl=-1.26667, yl=-1.26667, dx=.13333, dy=.13333, ibase=1,
                 prmudf=[0.0100, 0.0, 0.0100, 0.0, 0.0, 0.0], gg=(200.0, 200.0), poiss=(0.42, 0.42), fn=9.1954,
                 maxgs=100, maxin=100, eps=1e-4)


def hertz91_fn(ncase, fn0=2000.0, seed=20240229):
    """FN = FN0 (1 + 0.2 u), u ~ U(-1,1); case i uses draws [4i:4i+4] (first draw = load)."""
    u = np.random.default_rng(seed).uniform(-1.0, 1.0, size=4 * ncase).reshape(ncase, 4)
    return fn0 * (1.0 + 0.2 * u[:, 0]), u


def grid_xy(mx, my, xl, yl, dx, dy):
    x = xl + 0.5 * dx + dx * np.arange(mx)
    y = yl + 0.5 * dy + dy * np.arange(my)
    X, Y = np.meshgrid(x, y)           # (my, mx), x fastest
    return X.ravel(), Y.ravel()


def quadratic_h(g):
    X, Y = grid_xy(g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"])
    b = g["prmudf"]
    return b[0] * X * X + b[1] * X * Y + b[2] * Y * Y + b[3] * X + b[4] * Y + b[5]


def disk_mask(mx, my, frac=0.4):
    iy, ix = np.mgrid[0:my, 0:mx]
    r = frac * min(mx, my)
    return (((ix - (mx - 1) / 2.0) ** 2 + (iy - (my - 1) / 2.0) ** 2) <= r * r).astype(np.int32).ravel()


def microbench_p(mx, my, ncase, seed=7):
    """p ~ N(0,1) on a disk of radius 0.4 min(mx,my), zero elsewhere (kernel microbenchmark input)."""
    rng = np.random.default_rng(seed)
    el = disk_mask(mx, my)
    p = np.zeros((ncase, 3, mx * my))
    p[:] = rng.standard_normal((ncase, 3, mx * my)) * el
    return p, np.tile(el, (ncase, 1))
