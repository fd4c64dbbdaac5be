"""Synthetic scenes of BASELINE.json's configurations (shared by bench.py, the CLI driver and the tests)."""
import numpy as np

FS_EMPTY, FS_FLUID, FS_SOLID = 0, 1, 2


def dam_break_cells(n, ny=None):
    """Scene of reference demo/App.cpp:147-160 in row-major cell[j, i]: solid border,
    FLUID where i + j < 3N/4, EMPTY elsewhere."""
    ny = n if ny is None else ny
    j, i = np.meshgrid(np.arange(ny), np.arange(n), indexing="ij")
    c = np.where(i + j < ny * 3 // 4, FS_FLUID, FS_EMPTY).astype(np.uint8)
    c[0, :] = FS_SOLID
    c[-1, :] = FS_SOLID
    c[:, 0] = FS_SOLID
    c[:, -1] = FS_SOLID
    return c


def dam_break_params(n):
    """dx = 1.28/N; dt scaled to keep the demo's CFL number above 128 cells"""
    return dict(dt=0.005 * 128.0 / n if n > 128 else 0.005, dx=1.28 / n)


def splitmix_uniform(count, seed):
    """U(-1, 1) from SplitMix64, vectorised"""
    idx = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def projection_stress(n, variant="3a"):
    """BASELINE config 3: closed tank (3a: border SOLID, interior FLUID) or free surface (3b: rows j >= 3N/4 EMPTY with
    phi = (j - 3N/4 + 1/2) dx, fluid phi = -dx), interior face velocities U(-1, 1), faces touching SOLID zero.
    Returns cells, phi, u, v, dx (dt = dx, rho = 997)."""
    dx = 1.0 / n
    cells = np.full((n, n), FS_FLUID, np.uint8)
    cells[0, :] = cells[-1, :] = FS_SOLID
    cells[:, 0] = cells[:, -1] = FS_SOLID
    phi = np.full((n, n), -dx)
    if variant.startswith("3b"):
        top = 3 * n // 4
        cells[top:-1, 1:-1] = FS_EMPTY
        phi[top:, :] = ((np.arange(top, n) - top + 0.5) * dx)[:, None]
    u = splitmix_uniform(n * (n + 1), 0x5EED).reshape(n, n + 1)
    v = splitmix_uniform((n + 1) * n, 0x5EED + 1).reshape(n + 1, n)
    u[:, :2] = 0; u[:, -2:] = 0; v[:2, :] = 0; v[-2:, :] = 0
    return cells, phi, u, v, dx
