"""Synthetic geometry for the benchmark configurations (BASELINE.json ``configs``).

Every generator returns the reference's mesh format: ``[(name, V float32[N,3], F int32[M,3]), ...]``
(README.md:55-64 of the reference).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

Mesh = Tuple[str, np.ndarray, np.ndarray]


def parallel_unit_squares() -> List[Mesh]:
    """C1: the README example, two unit squares at z=0 and z=1 (reference README.md:55-64)."""
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=float)
    F = np.array([[0, 1, 2], [0, 2, 3]])
    return [("A", V, F), ("B", V + [0, 0, 1], F)]


def _quad(name: str, corners, flip: bool) -> Mesh:
    V = np.asarray(corners, dtype=np.float32)
    F = np.asarray([[0, 2, 1], [0, 3, 2]] if flip else [[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    return name, V, F


def street_canyon() -> List[Mesh]:
    """C2/C3: the 11-mesh street canyon of the reference's examples (ex00_street_canyon_geometry.py:68-102):
    five 10 m x 4 m facade panels per side at x = -4 / +4 and an 8 m x 10 m road at z = 0."""
    meshes: List[Mesh] = []
    for i in range(5):
        z0, z1 = 4.0 * i, 4.0 * (i + 1)
        for name, x, flip in ((f"east_side_{i}", -4.0, False), (f"west_side_{i}", 4.0, True)):
            meshes.append(_quad(name, [(x, -5.0, z0), (x, 5.0, z0), (x, 5.0, z1), (x, -5.0, z1)], flip))
    meshes.append(_quad("road", [(-4.0, -5.0, 0.0), (4.0, -5.0, 0.0), (4.0, 5.0, 0.0), (-4.0, 5.0, 0.0)], False))
    return meshes


def unit_cube_enclosure() -> List[Mesh]:
    """C4: closed unit cube, six quads with outward normals (ex04_inside_enclosure.py:33-65)."""
    O = np.zeros(3, np.float32)
    X = np.array([1, 0, 0], np.float32)
    Y = np.array([0, 1, 0], np.float32)
    Z = np.array([0, 0, 1], np.float32)

    def oriented(name, p0, p1, p2, p3, want):
        V = np.array([p0, p1, p2, p3], dtype=np.float32)
        n = np.cross(V[1] - V[0], V[2] - V[0])
        return _quad(name, V, bool(np.dot(n, np.asarray(want, np.float32)) < 0.0))

    return [
        oriented("Bottom", O, X, X + Y, Y, (0, 0, -1)),
        oriented("Top", O + Z, Y + Z, X + Y + Z, X + Z, (0, 0, 1)),
        oriented("Front", O, O + Z, X + Z, X, (0, -1, 0)),
        oriented("Back", Y, X + Y, X + Y + Z, Y + Z, (0, 1, 0)),
        oriented("Left", O, Y, Y + Z, O + Z, (-1, 0, 0)),
        oriented("Right", X, X + Z, X + Y + Z, X + Y, (1, 0, 0)),
    ]


def quad_grid(name: str, p0, du, dv, n: int) -> Mesh:
    """(n x n) quad grid; vertex (i,j) = p0 + i*du/n + j*dv/n at index i*(n+1)+j; per cell the triangles
    [a,b,c],[a,c,d] with a=(i,j), b=(i+1,j), c=(i+1,j+1), d=(i,j+1), so the normal is du x dv."""
    p0 = np.asarray(p0, np.float64)
    du = np.asarray(du, np.float64)
    dv = np.asarray(dv, np.float64)
    i, j = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    V = (p0[None, None, :] + i[..., None] * du / n + j[..., None] * dv / n).reshape(-1, 3).astype(np.float32)
    ci, cj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (ci * (n + 1) + cj).reshape(-1)
    b = a + (n + 1)
    c = b + 1
    d = a + 1
    F = np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1).reshape(-1, 3).astype(np.int32)
    return name, V, F


def urban_block(n_side: int = 20, face_grid: int = 16, ground_grid: int = 32, seed: int = 0) -> List[Mesh]:
    """C5 family: ``n_side x n_side`` buildings on a 20 m pitch, 12 m x 12 m footprints, heights
    ``default_rng(seed).uniform(8, 60)`` drawn per building in ``for bx: for by:`` order; each building is the
    five meshes S, N, W, E, roof (each a ``face_grid``^2 quad grid, outward normals); the last mesh is the ground.
    Defaults give SURVEY.md's C5: 2001 meshes, 1 026 048 triangles."""
    rng = np.random.default_rng(seed)
    meshes: List[Mesh] = []
    X = np.array([12.0, 0.0, 0.0])
    Y = np.array([0.0, 12.0, 0.0])
    for bx in range(n_side):
        for by in range(n_side):
            h = float(rng.uniform(8.0, 60.0))
            o = np.array([20.0 * bx, 20.0 * by, 0.0])
            Zv = np.array([0.0, 0.0, h])
            tag = f"b{bx:02d}_{by:02d}"
            meshes.append(quad_grid(f"{tag}_S", o, X, Zv, face_grid))
            meshes.append(quad_grid(f"{tag}_N", o + Y, Zv, X, face_grid))
            meshes.append(quad_grid(f"{tag}_W", o, Zv, Y, face_grid))
            meshes.append(quad_grid(f"{tag}_E", o + X, Y, Zv, face_grid))
            meshes.append(quad_grid(f"{tag}_roof", o + Zv, X, Y, face_grid))
    ext = 20.0 * n_side + 20.0
    meshes.append(quad_grid("ground", (-20.0, -20.0, 0.0), (ext, 0.0, 0.0), (0.0, ext, 0.0), ground_grid))
    return meshes


def tilted_pair() -> List[Mesh]:
    """Two non-axis-aligned, non-planar-split meshes (exercises the general tangent-frame branch)."""
    V1 = np.array([[0, 0, 0], [2, 0.3, 0.1], [2.2, 1.7, 0.9], [0.1, 1.5, 0.7], [1.0, 0.8, 0.2]], np.float32)
    F1 = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.int32)
    V2 = np.array([[0.2, 0.1, 2.0], [1.9, 0.0, 2.4], [2.1, 1.8, 1.6], [0.0, 1.6, 2.2]], np.float32)
    F2 = np.array([[0, 2, 1], [0, 3, 2]], np.int32)
    return [("bowl", V1, F1), ("lid", V2, F2)]
