"""CPU model of the cut of the split extrapolation (fluid-sim_b200/csrc/projection.cu nearLayersKernel /
particleCellDistKernel): the particle stages that run beside the far layer fill -- grid-to-particle transfer
(reference src/FluidSim2D.cpp:552-568, bilinear gather include/Array2D.h:402-420) and RK3 advection (:570-604,
Catmull-Rom include/Array2D.h:244-359 with the MAC mapping include/MACGrid2D.h:80-98) -- may only read faces whose
BFS layer (include/Array2D.h:552-591) is at most K = 2 * (ceil(c) + 3) + 2 + D.  The index arithmetic of the
samplers is restated here (same truncations, clamps and half-cell shifts as csrc/sampling.cuh) and run against
exact L1 distance transforms of random label fields."""
import numpy as np
import pytest
from scipy.ndimage import distance_transform_cdt

FLUID, EMPTY, SOLID = 0, 1, 2


def face_layers(cell):
    """BFS layers of the u and v faces: 0 = known (a FLUID cell on either side, or the last column / row, which the
    reference's loops never visit), else the L1 distance to the nearest known face of the same array."""
    ny, nx = cell.shape
    fl = cell == FLUID
    ku = np.zeros((ny, nx + 1), bool)
    ku[:, :nx] |= fl
    ku[:, 1:nx] |= fl[:, :-1]
    ku[:, nx] = True
    kv = np.zeros((ny + 1, nx), bool)
    kv[:ny, :] |= fl
    kv[1:ny, :] |= fl[:-1, :]
    kv[ny, :] = True
    du = distance_transform_cdt(~ku, metric="taxicab")
    dv = distance_transform_cdt(~kv, metric="taxicab")
    return du, dv


def clamp(v, lo, hi):
    return max(lo, min(v, hi))


def bicubic_faces(px, py, NX, NY):
    """indices the Catmull-Rom sampler reads (csrc/sampling.cuh bicubic)"""
    x, y = int(px), int(py)
    if x < 0 or x >= NX or y < 0 or y >= NY:
        return []
    xs = [clamp(x - 1 + k, 0, NX - 1) for k in range(4)]
    ys = [clamp(y - 1 + k, 0, NY - 1) for k in range(4)]
    return [(i, j) for j in ys for i in xs]


def bilinear_faces(px, py, NX, NY):
    ui, uj = int(px), int(py)
    xs = [clamp(ui, 0, NX - 1), clamp(ui + 1, 0, NX - 1)]
    ys = [clamp(uj, 0, NY - 1), clamp(uj + 1, 0, NY - 1)]
    return [(i, j) for j in ys for i in xs]


def sample_u(x, y, nx, ny):  # positions in cells (x / dx, y / dx)
    px = clamp(x, 1e-6, (nx - 1) - 1e-6)
    py = clamp(y - 0.5, 1e-6, (ny - 1) - 1e-6)
    return bicubic_faces(px, py, nx + 1, ny)


def sample_v(x, y, nx, ny):
    px = clamp(x - 0.5, 1e-6, (nx - 1) - 1e-6)
    py = clamp(y, 1e-6, (ny - 1) - 1e-6)
    return bicubic_faces(px, py, nx, ny + 1)


@pytest.mark.parametrize("seed", range(6))
def test_particle_stages_stay_inside_the_cut(seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(12, 48)), int(rng.integers(12, 48))
    cell = np.full((ny, nx), EMPTY, np.uint8)
    cell[[0, -1], :] = SOLID
    cell[:, [0, -1]] = SOLID
    # a few fluid blobs and an interior solid block
    for _ in range(int(rng.integers(1, 4))):
        ci, cj, r = rng.integers(2, nx - 2), rng.integers(2, ny - 2), rng.integers(1, 6)
        jj, ii = np.ogrid[:ny, :nx]
        blob = (abs(ii - ci) + abs(jj - cj) <= r) & (cell != SOLID)
        cell[blob] = FLUID
    bi, bj = rng.integers(2, nx - 4), rng.integers(2, ny - 4)
    cell[bj:bj + 3, bi:bi + 3] = SOLID
    du, dv = face_layers(cell)
    off = 1e-3
    worst = 0
    for _ in range(400):
        # a particle anywhere clampPos (src/FluidSim2D.cpp:645-651) allows, in a cell of any label
        x = rng.uniform(1 + off, nx - 1 - off)
        y = rng.uniform(1 + off, ny - 1 - off)
        ci, cj = int(x), int(y)
        D = 0
        if cell[cj, ci] != FLUID:
            D = max(du[cj, ci], du[cj, ci + 1], dv[cj, ci], dv[cj + 1, ci])
        c = rng.choice([0.0, 0.3, 1.0, 2.7, 6.2])       # bound of the displacement in cells (1.5625 vmax dt / dx)
        K = 2 * (int(np.ceil(c)) + 3) + 2 + D
        touched_u = bilinear_faces(x, y - 0.5, nx + 1, ny)  # G2P (csrc/particles.cu g2pKernel)
        touched_v = bilinear_faces(x - 0.5, y, nx, ny + 1)
        for _stage in range(3):                            # the three RK3 stage positions, anywhere within c cells
            sx, sy = x + rng.uniform(-c, c), y + rng.uniform(-c, c)
            touched_u += sample_u(sx, sy, nx, ny)
            touched_v += sample_v(sx, sy, nx, ny)
        lu = max(du[j, i] for i, j in touched_u)
        lv = max(dv[j, i] for i, j in touched_v)
        assert lu <= K and lv <= K, (seed, x, y, c, D, K, lu, lv)
        worst = max(worst, max(lu, lv) - (K - 2))
    assert worst <= 0  # the +2 of the formula is pure margin


def test_layers_are_l1_distances_like_the_bfs():
    """the BFS of Array2D::extrapolate on the full rectangle assigns layer = L1 distance to the nearest known face"""
    rng = np.random.default_rng(1)
    known = rng.random((17, 23)) < 0.05
    known[3, 4] = True
    d = distance_transform_cdt(~known, metric="taxicab")
    # plain breadth-first search
    layer = np.where(known, 0, -1)
    frontier = list(zip(*np.nonzero(known)))
    k = 0
    while frontier:
        k += 1
        nxt = []
        for j, i in frontier:
            for dj, di in ((0, -1), (0, 1), (-1, 0), (1, 0)):
                a, b = j + dj, i + di
                if 0 <= a < known.shape[0] and 0 <= b < known.shape[1] and layer[a, b] < 0:
                    layer[a, b] = k
                    nxt.append((a, b))
        frontier = nxt
    assert np.array_equal(layer, d)
