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


def _height_grid(name: str, x0: float, y0: float, size: float, n: int, height) -> Mesh:
    """(n x n) quad grid over [x0, x0+size] x [y0, y0+size] lifted to z = height(x, y); same indexing as quad_grid."""
    i, j = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    x = x0 + i * (size / n)
    y = y0 + j * (size / n)
    V = np.stack([x, y, height(x, y)], -1).reshape(-1, 3).astype(np.float32)
    ci, cj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (ci * (n + 1) + cj).reshape(-1)
    b, d = a + (n + 1), a + 1
    c = b + 1
    F = np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1).reshape(-1, 3).astype(np.int32)
    return name, V, F


def _box(name: str, centre, half) -> Mesh:
    cx, cy, cz = centre
    hx, hy, hz = half
    V = np.array([[cx + sx * hx, cy + sy * hy, cz + sz * hz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], np.float32)
    F = np.array([[0, 2, 1], [1, 2, 3], [4, 5, 6], [5, 7, 6], [0, 1, 4], [1, 5, 4], [2, 6, 3], [3, 6, 7],
                  [0, 4, 2], [2, 4, 6], [1, 3, 5], [3, 7, 5]], np.int32)           # outward normals
    return name, V, F


def _uv_sphere(name: str, centre, radius: float, n_lon: int = 16, n_lat: int = 8) -> Mesh:
    th = np.linspace(0.0, np.pi, n_lat + 1)[1:-1]
    ph = np.linspace(0.0, 2.0 * np.pi, n_lon, endpoint=False)
    ring = np.stack([np.outer(np.sin(th), np.cos(ph)), np.outer(np.sin(th), np.sin(ph)), np.outer(np.cos(th), np.ones(n_lon))], -1)
    V = np.concatenate([[[0.0, 0.0, 1.0]], ring.reshape(-1, 3), [[0.0, 0.0, -1.0]]]) * radius + np.asarray(centre)
    F = []
    for k in range(n_lon):
        F.append([0, 1 + k, 1 + (k + 1) % n_lon])
    for r in range(n_lat - 2):
        for k in range(n_lon):
            a, b = 1 + r * n_lon + k, 1 + r * n_lon + (k + 1) % n_lon
            F += [[a, a + n_lon, b], [b, a + n_lon, b + n_lon]]
    last = 1 + (n_lat - 2) * n_lon
    for k in range(n_lon):
        F.append([last + k, last + n_lon, last + (k + 1) % n_lon])
    return name, V.astype(np.float32), np.asarray(F, np.int32)


def terrain_with_objects(n_tiles: int = 6, tile_grid: int = 118, n_boxes: int = 500, n_spheres: int = 200, n_slabs: int = 12,
                         seed: int = 0) -> List[Mesh]:
    """A scene with strongly varying triangle sizes (the opposite of the regular urban grid): rolling terrain in
    ``n_tiles`` x ``n_tiles`` non-planar patch meshes of 100 m (``tile_grid``^2 quads each, ~0.85 m triangles at the
    default), boxes of 0.3-3 m (12 triangles each), finely tessellated spheres of 0.2-0.6 m radius (224 triangles of a
    few centimetres), and a few 40-60 m free-standing slabs of two triangles.  Defaults: 748 meshes, 1 053 352 triangles."""
    rng = np.random.default_rng(seed)
    size = 100.0
    ext = n_tiles * size

    def height(x, y):
        return 6.0 * np.sin(x / 37.0) * np.cos(y / 53.0) + 2.5 * np.sin((x + 2.0 * y) / 19.0) + 0.02 * x

    meshes: List[Mesh] = []
    for tx in range(n_tiles):
        for ty in range(n_tiles):
            meshes.append(_height_grid(f"terrain_{tx}_{ty}", tx * size, ty * size, size, tile_grid, height))
    for k in range(n_boxes):
        x, y = rng.uniform(5.0, ext - 5.0, 2)
        half = rng.uniform(0.15, 1.5, 3)
        meshes.append(_box(f"box_{k:03d}", (x, y, float(height(x, y)) + half[2] + 0.05), half))
    for k in range(n_spheres):
        x, y = rng.uniform(5.0, ext - 5.0, 2)
        r = float(rng.uniform(0.2, 0.6))
        meshes.append(_uv_sphere(f"sphere_{k:03d}", (x, y, float(height(x, y)) + r + rng.uniform(0.1, 3.0)), r))
    for k in range(n_slabs):
        x, y = rng.uniform(60.0, ext - 60.0, 2)
        L, H = rng.uniform(40.0, 60.0), rng.uniform(8.0, 20.0)
        ang = rng.uniform(0.0, np.pi)
        dx, dy = 0.5 * L * np.cos(ang), 0.5 * L * np.sin(ang)
        z0 = float(height(x, y)) - 8.0
        meshes.append(_quad(f"slab_{k:02d}", [(x - dx, y - dy, z0), (x + dx, y + dy, z0), (x + dx, y + dy, z0 + H + 16.0),
                                               (x - dx, y - dy, z0 + H + 16.0)], False))
    return meshes
