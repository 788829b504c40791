"""Scene / emitter preparation and the ``PreparedSolver`` cache.

Mirrors the public surface of the reference's ``raystrack.utils.prepared`` (utils/prepared.py:13-431):
``PreparedSolver(meshes)`` with ``get_scene / get_emitters / get_emitter / get_mesh_bounds /
clear_device_cache / get_device_scene / get_device_emitter`` and the attributes ``meshes`` / ``total_faces``.
The per-triangle Python loops of the reference (115 s for one million triangles) are replaced by vectorised
NumPy that yields the same float32 arrays; the BVH and the Halton tables are built on the GPU
(``csrc/rsk_bvh.cu``, ``csrc/rsk_qmc.cu``), so host copies of them exist only on request.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import _native

Mesh = Tuple[str, np.ndarray, np.ndarray]


def grid_from_density(area: float, density: float) -> int:
    """Halton grid side for a surface area and a sample density (reference utils/helpers.py:8-11)."""
    return max(int(np.ceil(np.sqrt(max(area, 0.0) * density))), 4)


@dataclass(frozen=True)
class PreparedScene:
    """Flattened scene in mesh order (reference utils/prepared.py:13-26).  The BVH arrays of the reference
    (``bb_min`` ... ``count``) stay ``None``: the tree is built and kept on the device."""
    v0: np.ndarray
    e1: np.ndarray
    e2: np.ndarray
    normals: np.ndarray
    sid: np.ndarray
    bb_min: Optional[np.ndarray] = None
    bb_max: Optional[np.ndarray] = None
    left: Optional[np.ndarray] = None
    right: Optional[np.ndarray] = None
    start: Optional[np.ndarray] = None
    count: Optional[np.ndarray] = None
    use_bvh: bool = False


@dataclass(frozen=True)
class PreparedEmitter:
    """Per-mesh emission data (reference utils/prepared.py:29-55)."""
    tri_a: np.ndarray
    tri_e1: np.ndarray
    tri_e2: np.ndarray
    tri_u: np.ndarray
    tri_v: np.ndarray
    tri_n: np.ndarray
    tri_origin_eps: np.ndarray
    plane_origin: np.ndarray
    plane_normal: np.ndarray
    plane_tol: float
    plane_is_planar: bool
    cdf: np.ndarray
    total_area: float
    g: int
    rays: int = 0

    @property
    def n_cells(self) -> int:
        return int(self.g) * int(self.g)

    # The reference stores the QMC tables on every emitter; here they live on the GPU.  Host copies are
    # produced on demand with the same float64 recurrence (utils/halton.py:9-39).
    # A zero-area emitter has all-zero tables in the reference (prepared.py:278-287).
    @property
    def u_grid(self) -> np.ndarray:
        return halton_grid_host(self.g)[0] if self.total_area > 0.0 else np.zeros(self.n_cells, np.float32)

    @property
    def v_grid(self) -> np.ndarray:
        return halton_grid_host(self.g)[1] if self.total_area > 0.0 else np.zeros(self.n_cells, np.float32)

    def _dim(self, base: int) -> np.ndarray:
        if not self.total_area > 0.0:
            return np.zeros(self.n_cells * self.rays, np.float32)
        return halton_dim_host(self.n_cells * self.rays, base)

    halton_tri = property(lambda self: self._dim(5))
    halton_u = property(lambda self: self._dim(2))
    halton_v = property(lambda self: self._dim(3))
    halton_r1 = property(lambda self: self._dim(7))
    halton_r2 = property(lambda self: self._dim(11))


def halton_dim_host(length: int, base: int) -> np.ndarray:
    """float32(H_base(i+1)), i < length, with the reference's recurrence f /= b; r += f*digit (halton.py:9-18)."""
    i = np.arange(1, int(length) + 1, dtype=np.int64)
    f = np.ones(i.shape, np.float64)
    r = np.zeros(i.shape, np.float64)
    while np.any(i > 0):
        live = i > 0
        f = np.where(live, f / base, f)
        r = np.where(live, r + f * (i % base), r)
        i = i // base
    return r.astype(np.float32)


def halton_grid_host(g: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-cell jitter grid (halton.py:21-31)."""
    g = int(g)
    c = np.arange(g * g, dtype=np.int64)
    h2 = halton_dim_host_f64(g * g, 2)
    h3 = halton_dim_host_f64(g * g, 3)
    return ((h2 + (c // g)) / g).astype(np.float32), ((h3 + (c % g)) / g).astype(np.float32)


def halton_dim_host_f64(length: int, base: int) -> np.ndarray:
    i = np.arange(1, int(length) + 1, dtype=np.int64)
    f = np.ones(i.shape, np.float64)
    r = np.zeros(i.shape, np.float64)
    while np.any(i > 0):
        live = i > 0
        f = np.where(live, f / base, f)
        r = np.where(live, r + f * (i % base), r)
        i = i // base
    return r


def _unit_rows(v: np.ndarray) -> np.ndarray:
    n = np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)      # prepared.py:93-96
    return v / n


def _corners(V: np.ndarray, F: np.ndarray):
    a = np.asarray(V[F[:, 0]], dtype=np.float32)
    e1 = (np.asarray(V[F[:, 1]], dtype=np.float32) - a).astype(np.float32, copy=False)
    e2 = (np.asarray(V[F[:, 2]], dtype=np.float32) - a).astype(np.float32, copy=False)
    return a, e1, e2


def _triangle_frames(tri_n: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Tangent frames (reference prepared.py:99-122), all triangles at once: u = normalise(ref x n) with
    ref = x unless |n.x| >= 0.9, v = n x u; degenerate normals fall back to (x, y)."""
    ax = np.asarray([1.0, 0.0, 0.0], np.float32)
    ay = np.asarray([0.0, 1.0, 0.0], np.float32)
    t = tri_n.shape[0]
    if t == 0:
        return np.empty((0, 3), np.float32), np.empty((0, 3), np.float32)
    use_x = np.abs(tri_n[:, 0].astype(np.float64)) < 0.9
    ref = np.where(use_x[:, None], ax[None, :], ay[None, :]).astype(np.float32)
    u = np.cross(ref, tri_n).astype(np.float32)
    length = np.linalg.norm(u, axis=1)
    retry = length.astype(np.float64) <= 1e-12
    if np.any(retry):
        ref2 = np.where(use_x[:, None], ay[None, :], ax[None, :]).astype(np.float32)
        u2 = np.cross(ref2, tri_n).astype(np.float32)
        u = np.where(retry[:, None], u2, u)
        length = np.where(retry, np.linalg.norm(u2, axis=1), length)
    dead = length.astype(np.float64) <= 1e-12
    safe = np.where(dead, np.float32(1.0), length).astype(np.float32)
    u = (u / safe[:, None]).astype(np.float32)
    v = np.cross(tri_n, u).astype(np.float32)
    if np.any(dead):
        u[dead] = ax
        v[dead] = ay
    return u, v


def _origin_eps(e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    """Ray-origin offset per triangle: 1e-6 of the longest edge, at least 1e-8 (prepared.py:125-130)."""
    scale = np.maximum(np.linalg.norm(e1, axis=1), np.maximum(np.linalg.norm(e2, axis=1), np.linalg.norm(e2 - e1, axis=1)))
    return np.maximum(scale * 1.0e-6, 1.0e-8).astype(np.float32, copy=False)


def _emitter_plane(a, e1, e2, n, eps):
    """Planarity record used for back-face culling of whole meshes (prepared.py:133-167)."""
    origin = np.zeros(3, np.float32)
    normal = np.zeros(3, np.float32)
    tol = float(max(1.0e-7, np.max(eps) if eps.size else 0.0))
    if a.shape[0] == 0:
        return origin, normal, tol, False
    origin = np.asarray(a[0], dtype=np.float32)
    normal = np.asarray(n[0], dtype=np.float32)
    nl = float(np.linalg.norm(normal))
    if nl <= 1.0e-12:
        return origin, normal, tol, False
    normal = (normal / nl).astype(np.float32, copy=False)
    if np.any(n @ normal < (1.0 - 1.0e-4)):
        return origin, normal, tol, False
    worst = 0.0
    for p in (a, a + e1, a + e2):
        d = np.abs((p - origin) @ normal)
        worst = max(worst, float(np.max(d)) if d.size else 0.0)
    return origin, normal, tol, worst <= tol


def prepare_scene(meshes: List[Mesh], *, use_bvh: bool) -> PreparedScene:
    """Flatten all meshes into triangle arrays in mesh order (prepared.py:170-243, without the host BVH)."""
    parts = []
    for sid, (_, V, F) in enumerate(meshes):
        a, e1, e2 = _corners(V, F)
        n = _unit_rows(np.cross(e1, e2).astype(np.float32, copy=False)).astype(np.float32, copy=False)
        parts.append((a, e1, e2, n, np.full(F.shape[0], sid, np.int32)))
    if not parts:
        z3 = np.empty((0, 3), np.float32)
        return PreparedScene(z3, z3, z3, z3, np.empty(0, np.int32), use_bvh=False)
    v0, e1, e2, nn, sid = (np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=0)) for k in range(5))
    return PreparedScene(v0, e1, e2, nn, sid, use_bvh=bool(use_bvh and v0.shape[0] > 0))


def prepare_emitters(meshes: List[Mesh], *, samples: int, rays: int, flip_faces: bool) -> List[PreparedEmitter]:
    """Per-mesh emission data (prepared.py:246-321): frames, origin offsets, area CDF, grid size."""
    out: List[PreparedEmitter] = []
    for _, V, F in meshes:
        Fe = F[:, [0, 2, 1]] if flip_faces else F
        a, e1, e2 = _corners(V, Fe)
        n_raw = np.cross(e1, e2).astype(np.float32, copy=False)
        twice = np.linalg.norm(n_raw, axis=1)
        n = _unit_rows(n_raw).astype(np.float32, copy=False)
        tu, tv = _triangle_frames(n)
        eps = _origin_eps(e1, e2)
        po, pn, ptol, planar = _emitter_plane(a, e1, e2, n, eps)
        areas = 0.5 * twice
        total = float(areas.sum())
        if total <= 0.0:
            cdf = np.ones(Fe.shape[0], np.float32)
            g = 4
        else:
            cdf = np.cumsum(areas, dtype=np.float64)
            cdf = (cdf / cdf[-1]).astype(np.float32)
            g = grid_from_density(total, samples)
        out.append(PreparedEmitter(a, e1, e2, tu, tv, n, eps, po, pn, ptol, planar, cdf, total, g, int(rays)))
    return out


@dataclass(frozen=True)
class EmitterSummary:
    """What the solve drivers need to know about an emitter besides its device records: the fields of
    ``PreparedEmitter`` that are per mesh, not per triangle."""
    plane_origin: np.ndarray
    plane_normal: np.ndarray
    plane_tol: float
    plane_is_planar: bool
    total_area: float
    g: int
    rays: int = 0

    @property
    def n_cells(self) -> int:
        return int(self.g) * int(self.g)


class SummaryList(list):
    """List of EmitterSummary that also carries the plane records as arrays: ``soa`` = (planar, origin, normal, tol)."""
    soa = None


def flatten_meshes(meshes: List[Mesh]):
    """All meshes back to back: (verts float32[nv,3], vert_offset int64[n+1], faces int32[nt,3], tri_offset int64[n+1])
    -- the input of the device-side preparation (csrc/rsk_prepare.cu)."""
    n = len(meshes)
    vo = np.zeros(n + 1, np.int64)
    to = np.zeros(n + 1, np.int64)
    if n == 0:
        return np.empty((0, 3), np.float32), vo, np.empty((0, 3), np.int32), to
    vs = [np.asarray(V, dtype=np.float32).reshape(-1, 3) for _, V, _ in meshes]
    fs = [np.asarray(F).reshape(-1, 3) for _, _, F in meshes]
    np.cumsum([v.shape[0] for v in vs], out=vo[1:])
    np.cumsum([f.shape[0] for f in fs], out=to[1:])
    verts = np.concatenate(vs, axis=0)
    faces64 = np.concatenate(fs, axis=0)
    narrow = faces64.dtype == np.int32 or (faces64.dtype.kind in "iu" and faces64.dtype.itemsize < 4)     # cannot overflow
    if not narrow and faces64.size and (faces64.max() > 0x7fffffff or faces64.min() < -0x80000000):
        raise IndexError("face index does not fit in int32")
    return np.ascontiguousarray(verts), vo, np.ascontiguousarray(faces64.astype(np.int32, copy=False)), to


def summaries_from_emitters(emitters: List[PreparedEmitter]) -> List[EmitterSummary]:
    return [EmitterSummary(e.plane_origin, e.plane_normal, e.plane_tol, e.plane_is_planar, e.total_area, e.g, e.rays) for e in emitters]


def device_plane_verdict(row, tol: float) -> Optional[bool]:
    """Planarity of a mesh from the device's float64 statistics, or None when they lie within the rounding distance
    of a threshold of ``_emitter_plane`` (float32 BLAS products there: ~4e-7 relative to the size of the terms)."""
    lo = 1.0 - 1.0e-4
    band = 1.0e-6 * float(row["worst_mag"])
    if row["min_dot"] < lo - 2.0e-6 or row["worst"] - band > tol:
        return False
    if row["min_dot"] >= lo + 2.0e-6 and row["worst"] + band <= tol:
        return True
    return None


def summaries_from_device(summary: np.ndarray, g: np.ndarray, meshes: List[Mesh], *, samples: int, rays: int,
                          flip_faces: bool) -> List[EmitterSummary]:
    """Finish the planarity record of ``_emitter_plane`` (reference prepared.py:133-167) from the float64 statistics
    the device returns.  The reference decides with float32 BLAS products; a mesh whose statistics lie within the
    rounding distance of a threshold is prepared once more on the host with the reference arithmetic, so the flag is
    the reference's in every case."""
    n = int(summary.shape[0])
    n0 = np.ascontiguousarray(summary["normal0"], np.float32).reshape(n, 3)
    origin = np.ascontiguousarray(summary["origin"], np.float32).reshape(n, 3).copy()
    tol = np.maximum(1.0e-7, summary["eps_max"].astype(np.float64))
    empty = np.asarray([np.shape(F)[0] == 0 for _, _, F in meshes], bool) if n else np.zeros(0, bool)
    # nl = float(np.linalg.norm(normal)): a BLAS dot in the reference.  With at most one non-zero component the result
    # is |component| in any summation order (sqrt(fl(a*a)) == |a|); other normals go through the same NumPy call.
    nl = np.abs(n0).max(axis=1).astype(np.float64) if n else np.zeros(0)
    for i in np.nonzero(((n0 != 0).sum(axis=1) > 1) & ~empty)[0]:
        nl[i] = float(np.linalg.norm(n0[i]))
    unit = (nl > 1.0e-12) & ~empty
    normal = n0.copy()
    normal[unit] = n0[unit] / nl[unit].astype(np.float32)[:, None]
    normal[empty] = 0.0
    origin[empty] = 0.0
    lo = 1.0 - 1.0e-4
    band = 1.0e-6 * summary["worst_mag"]
    no = (summary["min_dot"] < lo - 2.0e-6) | (summary["worst"] - band > tol)
    yes = ~no & (summary["min_dot"] >= lo + 2.0e-6) & (summary["worst"] + band <= tol)
    planar = (unit & yes).tolist()
    for i in np.nonzero(unit & ~yes & ~no)[0].tolist():      # within rounding distance of a threshold: the reference arithmetic decides
        planar[i] = bool(prepare_emitters([meshes[i]], samples=samples, rays=rays, flip_faces=flip_faces)[0].plane_is_planar)
    rays = int(rays)
    out = SummaryList(EmitterSummary(o, nv, t, p, a, gi, rays)
                      for o, nv, t, p, a, gi in zip(origin, normal, tol.tolist(), planar, summary["total_area"].tolist(), np.asarray(g).tolist()))
    out.soa = (np.asarray(planar, bool), origin, normal, tol.astype(np.float32))      # the same fields as arrays (surface masks)
    return out


@dataclass
class PreparedDeviceScene:
    """Device-resident scene: triangles + wide BVH owned by librsk_b200 (reference prepared.py:58-71)."""
    native: Any
    use_bvh: bool

    def info(self) -> dict:
        return self.native.info()


@dataclass
class PreparedDeviceEmitters:
    """Device-resident emitter set (all meshes of one samples/rays/flip_faces key)."""
    native: Any
    n_rays_once: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))


class PreparedSolver:
    """Cache prepared geometry and device uploads across solves (reference prepared.py:324-431).

    Reusing one instance avoids rebuilding triangle buffers, the GPU BVH, the QMC tables and the uploads;
    ``seed`` is not part of any key, so re-solves with a new seed reuse everything."""

    def __init__(self, meshes: List[Mesh]):
        self.meshes = list(meshes)
        self.total_faces = int(sum(F.shape[0] for _, _, F in self.meshes))
        self._scene_cache: Dict[bool, PreparedScene] = {}
        self._emitter_cache: Dict[Tuple[int, int, bool], List[PreparedEmitter]] = {}
        self._device_scene_cache: Dict[Tuple[int, bool], PreparedDeviceScene] = {}
        self._device_emitter_cache: Dict[Tuple[int, int, int, bool], PreparedDeviceEmitters] = {}
        self._mesh_bounds_cache: Optional[Tuple[np.ndarray, np.ndarray]] = None
        self._emitter_pack_cache: Dict[Tuple[int, int, bool], tuple] = {}
        self._summary_cache: Dict[Tuple[int, int, bool], List[EmitterSummary]] = {}
        self._flat_cache: Optional[tuple] = None                       # flatten_meshes(self.meshes)
        self._geometry_cache: Dict[int, Any] = {}                      # device -> _native.DeviceGeometry
        self._derived_cache: Dict[Any, Any] = {}      # host-side results derived from the prepared state (surface masks)

    def get_scene(self, *, use_bvh: bool) -> PreparedScene:
        key = bool(use_bvh)
        if key not in self._scene_cache:
            self._scene_cache[key] = prepare_scene(self.meshes, use_bvh=key)
        return self._scene_cache[key]

    def get_emitters(self, *, samples: int, rays: int, flip_faces: bool) -> List[PreparedEmitter]:
        key = (int(samples), int(rays), bool(flip_faces))
        if key not in self._emitter_cache:
            self._emitter_cache[key] = prepare_emitters(self.meshes, samples=samples, rays=rays, flip_faces=flip_faces)
        return self._emitter_cache[key]

    def get_emitter(self, index: int, *, samples: int, rays: int, flip_faces: bool) -> PreparedEmitter:
        return self.get_emitters(samples=samples, rays=rays, flip_faces=flip_faces)[int(index)]

    def get_mesh_bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        """AABB centre / half extent per mesh in float32 (prepared.py:359-375)."""
        if self._mesh_bounds_cache is None:
            n = len(self.meshes)
            centers = np.zeros((n, 3), np.float32)
            extents = np.zeros((n, 3), np.float32)
            verts, vo, _, _ = self._flat()
            if n and np.all(vo[1:] > vo[:-1]):              # every mesh has vertices: one segmented min/max
                cols = np.ascontiguousarray(verts.T)        # per coordinate: reduceat over a contiguous row is ~2x faster
                lo = np.stack([np.minimum.reduceat(cols[k], vo[:-1]) for k in range(3)], axis=1)
                hi = np.stack([np.maximum.reduceat(cols[k], vo[:-1]) for k in range(3)], axis=1)
                centers[:] = 0.5 * (lo + hi)
                extents[:] = 0.5 * (hi - lo)
            else:
                for i in range(n):
                    v = verts[vo[i]:vo[i + 1]]
                    if v.size == 0:
                        continue
                    lo, hi = np.min(v, axis=0), np.max(v, axis=0)
                    centers[i] = 0.5 * (lo + hi)
                    extents[i] = 0.5 * (hi - lo)
            self._mesh_bounds_cache = (centers, extents)
        return self._mesh_bounds_cache

    def _flat(self) -> tuple:
        if self._flat_cache is None:
            self._flat_cache = flatten_meshes(self.meshes)
        return self._flat_cache

    def _geometry(self, ctx: _native.Context):
        got = self._geometry_cache.get(ctx.device)
        if got is None:
            verts, vo, faces, to = self._flat()
            try:
                got = _native.DeviceGeometry(ctx, verts, vo, faces, to)
            except _native.NativeError as exc:
                raise ValueError(str(exc)) from exc
            self._geometry_cache[ctx.device] = got
        return got

    @staticmethod
    def _host_prepare_forced() -> bool:
        """RSK_HOST_PREPARE=1: prepare on the host with NumPy and upload the prepared arrays (A/B checks)."""
        return os.environ.get("RSK_HOST_PREPARE", "") not in ("", "0")

    def clear_device_cache(self) -> None:
        for s in self._device_scene_cache.values():
            s.native.close()
        for e in self._device_emitter_cache.values():
            e.native.close()
        for g in self._geometry_cache.values():
            g.close()
        self._device_scene_cache.clear()
        self._device_emitter_cache.clear()
        self._geometry_cache.clear()

    def get_device_scene(self, *, use_bvh: bool, ctx: Optional[_native.Context] = None) -> PreparedDeviceScene:
        """Device scene.  Built on the GPU from the raw meshes (csrc/rsk_prepare.cu) unless the host arrays of
        ``get_scene`` already exist, in which case those are uploaded; both give the same device records."""
        ctx = ctx or _native.Context.for_device()
        key = (ctx.device, bool(use_bvh))
        got = self._device_scene_cache.get(key)
        if got is None:
            hs = self._scene_cache.get(bool(use_bvh)) or self._scene_cache.get(not use_bvh)
            if hs is None and self._host_prepare_forced():
                hs = self.get_scene(use_bvh=use_bvh)
            if hs is not None:
                nat = _native.DeviceScene(ctx, hs.v0, hs.e1, hs.e2, hs.normals, hs.sid, len(self.meshes), bool(use_bvh and hs.v0.shape[0] > 0))
            else:
                nat = self._from_geometry(lambda: _native.DeviceScene.from_geometry(self._geometry(ctx), bool(use_bvh)))
            got = PreparedDeviceScene(nat, nat.use_bvh)
            self._device_scene_cache[key] = got
        return got

    @staticmethod
    def _from_geometry(make):
        try:
            return make()
        except _native.NativeError as exc:
            if "face indices" in str(exc):            # what NumPy's V[F] raises in the reference (prepared.py:182-188)
                raise IndexError(str(exc)) from exc
            raise

    def get_device_emitters(self, *, samples: int, rays: int, flip_faces: bool,
                            ctx: Optional[_native.Context] = None) -> PreparedDeviceEmitters:
        """Device emitter set of one (samples, rays, flip_faces) key; GPU-prepared unless ``get_emitters`` has
        already produced the host arrays for the key."""
        ctx = ctx or _native.Context.for_device()
        key = (ctx.device, int(samples), int(rays), bool(flip_faces))
        got = self._device_emitter_cache.get(key)
        if got is None:
            hkey = (int(samples), int(rays), bool(flip_faces))
            if hkey not in self._emitter_cache and not self._host_prepare_forced():
                nat, summary = self._from_geometry(
                    lambda: _native.DeviceEmitters.from_geometry(self._geometry(ctx), samples, int(rays), bool(flip_faces)))
                if hkey not in self._summary_cache:
                    self._summary_cache[hkey] = summaries_from_device(summary, nat.g, self.meshes, samples=samples,
                                                                       rays=int(rays), flip_faces=bool(flip_faces))
                got = PreparedDeviceEmitters(nat, nat.g.astype(np.int64) ** 2 * int(rays))
                self._device_emitter_cache[key] = got
                return got
            ems = self.get_emitters(samples=samples, rays=rays, flip_faces=flip_faces)
            pack = self._emitter_pack_cache.get(hkey)
            if pack is None:                      # concatenated host arrays: host preparation, kept across device resets
                counts = np.asarray([e.tri_a.shape[0] for e in ems], np.int64)
                off = np.zeros(len(ems) + 1, np.int64)
                np.cumsum(counts, out=off[1:])

                def cat(name, width):
                    if not ems:
                        return np.empty((0, width) if width else (0,), np.float32)
                    return np.ascontiguousarray(np.concatenate([getattr(e, name) for e in ems], axis=0), np.float32)

                pack = (off, cat("tri_a", 3), cat("tri_e1", 3), cat("tri_e2", 3), cat("tri_u", 3), cat("tri_v", 3),
                        cat("tri_n", 3), cat("tri_origin_eps", 0), cat("cdf", 0),
                        np.asarray([e.g if e.total_area > 0.0 else -e.g for e in ems], np.int32))      # -g: zero-area emitter
                self._emitter_pack_cache[hkey] = pack
            nat = _native.DeviceEmitters(ctx, *pack, int(rays))
            got = PreparedDeviceEmitters(nat, np.asarray([e.n_cells * int(rays) for e in ems], np.int64))
            self._device_emitter_cache[key] = got
        return got

    def get_emitter_summaries(self, *, samples: int, rays: int, flip_faces: bool,
                              ctx: Optional[_native.Context] = None) -> List[EmitterSummary]:
        """Per-mesh emitter facts (plane record, area, grid side) without the per-triangle host arrays: taken from
        ``get_emitters`` when that has run for the key, otherwise a by-product of the device-side preparation."""
        hkey = (int(samples), int(rays), bool(flip_faces))
        got = self._summary_cache.get(hkey)
        if got is None:
            if hkey in self._emitter_cache or self._host_prepare_forced():
                got = summaries_from_emitters(self.get_emitters(samples=samples, rays=rays, flip_faces=flip_faces))
                self._summary_cache[hkey] = got
            else:
                self.get_device_emitters(samples=samples, rays=rays, flip_faces=flip_faces, ctx=ctx)
                got = self._summary_cache[hkey]
        return got

    def get_device_emitter(self, index: int, *, samples: int, rays: int, flip_faces: bool) -> PreparedDeviceEmitters:
        """Reference signature (prepared.py:405-431); all emitters of a key share one device object."""
        return self.get_device_emitters(samples=samples, rays=rays, flip_faces=flip_faces)


__all__ = ["EmitterSummary", "PreparedScene", "PreparedEmitter", "PreparedDeviceScene", "PreparedDeviceEmitters", "PreparedSolver",
           "prepare_scene", "prepare_emitters", "grid_from_density"]
