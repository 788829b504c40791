"""ctypes binding of librsk_b200.so (include/raystrack_b200.h) -- the only bridge between the Python host
code and the CUDA implementation.  There is no CPU fallback: every entry point raises if the library or a
B200 is missing."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "librsk_b200.so"
_lib: Optional[C.CDLL] = None


class SolveParams(C.Structure):
    _fields_ = [("max_iters", C.c_int32), ("min_iters", C.c_int32), ("interval", C.c_int32),
                ("tol_mode", C.c_int32), ("tol", C.c_double)]


EXPORTS = (
    "rsk_last_error", "rsk_abi_version", "rsk_device_count",
    "rsk_ctx_create", "rsk_ctx_destroy", "rsk_ctx_synchronize", "rsk_ctx_timer_start", "rsk_ctx_timer_stop",
    "rsk_ctx_launch_count", "rsk_ctx_device_info",
    "rsk_scene_create", "rsk_scene_destroy", "rsk_scene_info", "rsk_scene_download_bvh",
    "rsk_emitters_create", "rsk_emitters_destroy", "rsk_emitters_download_tables",
    "rsk_geometry_create", "rsk_geometry_destroy", "rsk_scene_from_geometry", "rsk_emitters_from_geometry",
    "rsk_emitters_info", "rsk_emitters_download_records", "rsk_scene_download_triangles", "rsk_surface_masks",
    "rsk_trace_rays",
    "rsk_matrix_begin", "rsk_matrix_step", "rsk_matrix_read", "rsk_solve_read_block", "rsk_matrix_device_tallies",
    "rsk_sky_begin", "rsk_sky_step", "rsk_sky_read", "rsk_dual_begin", "rsk_dual_begin_sliced", "rsk_dual_step", "rsk_dual_sky_part",
    "rsk_solve_enqueue_trace", "rsk_solve_enqueue_fold", "rsk_solve_poll", "rsk_solve_device_iter_tallies", "rsk_solve_set_iter_tally_buffer",
    "rsk_solve_destroy", "rsk_solve_rays_traced", "rsk_trace_counters", "rsk_reciprocity_rowsum",
    "rsk_solve_read_block_view", "rsk_source_hash",
    "rsk_comm_unique_id", "rsk_comm_init", "rsk_comm_destroy", "rsk_comm_info", "rsk_allreduce_i64", "rsk_allreduce_host_i64",
    "rsk_solve_allreduce_iter_tallies",
    "rsk_tally_block_create", "rsk_tally_block_add_solve", "rsk_tally_block_allreduce", "rsk_tally_block_device",
    "rsk_tally_block_download", "rsk_tally_block_destroy",
    "rsk_solve_csr", "rsk_tally_block_csr", "rsk_csr_fetch", "rsk_ctx_set_l2_flush", "rsk_emitter_costs",
)
COMM_ID_BYTES = 128


class MeshSummary(C.Structure):
    """``rsk_mesh_summary``: per-mesh by-products of the device-side emitter preparation."""
    _fields_ = [("total_area", C.c_double), ("origin", C.c_float * 3), ("normal0", C.c_float * 3), ("eps_max", C.c_float),
                ("reserved", C.c_float), ("min_dot", C.c_double), ("worst", C.c_double), ("worst_mag", C.c_double)]


MESH_SUMMARY_DTYPE = np.dtype([("total_area", "<f8"), ("origin", "<f4", 3), ("normal0", "<f4", 3), ("eps_max", "<f4"),
                               ("reserved", "<f4"), ("min_dot", "<f8"), ("worst", "<f8"), ("worst_mag", "<f8")])
assert MESH_SUMMARY_DTYPE.itemsize == C.sizeof(MeshSummary) == 64


class NativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the in-tree library, building it with nvcc when it is missing (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import _build
    # The library is rebuilt whenever it is missing or was compiled from other sources than the ones in the tree (a
    # hash of csrc/ + include/ is embedded at build time and kept next to the .so): a stale binary never runs
    # against newer Python.  Without nvcc a stale library is an error, not a silent mismatch.
    if os.environ.get("RSK_REBUILD") or _build.stale():
        _build.build(force=True)
    # RSK_LIB: load an experimental build of the same library (kernel tuning, scripts/kernel_variants.py)
    variant = os.environ.get("RSK_LIB")
    lib = C.CDLL(variant or str(LIB_PATH))
    lib.rsk_last_error.restype = C.c_char_p
    lib.rsk_source_hash.restype = C.c_char_p
    for name in EXPORTS:
        if name not in ("rsk_last_error", "rsk_source_hash"):
            getattr(lib, name).restype = C.c_int
    if not variant:
        built_from = lib.rsk_source_hash().decode()
        if built_from != _build.source_hash():
            raise NativeError(f"{LIB_PATH} was built from other sources ({built_from}) than this tree ({_build.source_hash()})")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().rsk_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'librsk_b200'} failed (status {rc}): {msg}")


def ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    check(load().rsk_device_count(C.byref(n)), "rsk_device_count")
    return int(n.value)


CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the handle that names the legacy default stream explicitly


def torch_stream_handle(device: int) -> int:
    """cudaStream_t of torch's current stream on ``device`` in a form rsk_ctx_create accepts: torch reports the
    default stream as 0, which the C ABI reads as "create a private stream", so it is passed as cudaStreamLegacy."""
    import torch
    handle = int(torch.cuda.current_stream(device).cuda_stream)
    return handle if handle != 0 else CUDA_STREAM_LEGACY


class Context:
    """One CUDA context/stream per (process, GPU) -- wraps ``rsk_ctx``."""

    _by_device: dict = {}

    def __init__(self, device: int = 0, stream: int = 0):
        self.lib = load()
        self.handle = C.c_void_p()
        check(self.lib.rsk_ctx_create(C.c_int(device), C.c_void_p(stream or None), C.byref(self.handle)), "rsk_ctx_create")
        self.device = device

    @classmethod
    def for_device(cls, device: Optional[int] = None, stream: int = 0) -> "Context":
        """Cached context per (device, stream).  Default device: RSK_DEVICE or LOCAL_RANK (torchrun), else 0."""
        if device is None:
            device = int(os.environ.get("RSK_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            n = device_count()
            if n > 0:
                device %= n
        key = (device, int(stream or 0))
        ctx = cls._by_device.get(key)
        if ctx is None:
            ctx = cls(device, stream)
            cls._by_device[key] = ctx
        return ctx

    def synchronize(self) -> None:
        check(self.lib.rsk_ctx_synchronize(self.handle))

    def set_l2_flush(self, n_bytes: int) -> None:
        """Benchmark aid: write ``n_bytes`` of scratch (> L2) before every trace launch, on its stream; 0 = off."""
        check(self.lib.rsk_ctx_set_l2_flush(self.handle, C.c_int64(n_bytes)), "rsk_ctx_set_l2_flush")

    def timer_start(self) -> None:
        check(self.lib.rsk_ctx_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self.lib.rsk_ctx_timer_stop(self.handle, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_int64(0)
        check(self.lib.rsk_ctx_launch_count(self.handle, C.byref(n)))
        return int(n.value)

    def trace_counters(self, reset: bool = True) -> dict:
        """Work counters of the trace kernels (all zero unless the library is an RSK_COUNTERS=1 build)."""
        out = (C.c_int64 * 8)()
        check(self.lib.rsk_trace_counters(self.handle, out, C.c_int32(1 if reset else 0)))
        names = ("node_visits", "tri_tests", "tri_masked", "rays", "flush_trips", "stack_pushes")
        return {k: int(out[i]) for i, k in enumerate(names)}

    def device_info(self) -> dict:
        name = C.create_string_buffer(256)
        info = (C.c_int64 * 4)()
        check(self.lib.rsk_ctx_device_info(self.handle, name, info))
        return {"name": name.value.decode(), "sm_count": int(info[0]), "cc": (int(info[1]), int(info[2])), "mem": int(info[3])}

    def surface_masks(self, planar, plane_origin, plane_normal, plane_tol, centers, extents) -> np.ndarray:
        """``rsk_surface_masks``: uint8 [n_emit, n_surf] activity masks of all emitters."""
        planar = np.ascontiguousarray(planar, np.uint8)
        po, pn = np.ascontiguousarray(plane_origin, np.float32), np.ascontiguousarray(plane_normal, np.float32)
        tol = np.ascontiguousarray(plane_tol, np.float32)
        c, x = np.ascontiguousarray(centers, np.float32), np.ascontiguousarray(extents, np.float32)
        ne, ns = int(planar.shape[0]), int(c.shape[0])
        out = np.empty((ne, ns), np.uint8)
        check(self.lib.rsk_surface_masks(self.handle, C.c_int32(ne), C.c_int32(ns), ptr(planar), ptr(po), ptr(pn), ptr(tol),
                                         ptr(c), ptr(x), ptr(out)), "rsk_surface_masks")
        return out

    # ---- multi-GPU (rsk_comm.cu)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        check(load().rsk_comm_unique_id(buf), "rsk_comm_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError(f"NCCL unique id must be {COMM_ID_BYTES} bytes")
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        check(self.lib.rsk_comm_init(self.handle, buf, C.c_int32(rank), C.c_int32(nranks)), "rsk_comm_init")

    def comm_destroy(self) -> None:
        check(self.lib.rsk_comm_destroy(self.handle), "rsk_comm_destroy")

    def comm_info(self) -> dict:
        r, n, v = C.c_int32(0), C.c_int32(1), C.c_int32(0)
        check(self.lib.rsk_comm_info(self.handle, C.byref(r), C.byref(n), C.byref(v)), "rsk_comm_info")
        return {"rank": int(r.value), "nranks": int(n.value), "nccl_version": int(v.value)}

    def allreduce_device(self, device_ptr: int, n: int, op: str = "sum") -> None:
        check(self.lib.rsk_allreduce_i64(self.handle, C.c_void_p(device_ptr), C.c_int64(n), C.c_int32(0 if op == "sum" else 1)),
              "rsk_allreduce_i64")

    def allreduce_host(self, values: np.ndarray, op: str = "sum") -> np.ndarray:
        """In-place all-reduce of a small int64 host array over the context's communicator (synchronises)."""
        assert values.dtype == np.int64 and values.flags.c_contiguous
        check(self.lib.rsk_allreduce_host_i64(self.handle, ptr(values), C.c_int64(values.size), C.c_int32(0 if op == "sum" else 1)),
              "rsk_allreduce_host_i64")
        return values

    def reciprocity_rowsum(self, area: np.ndarray, F: np.ndarray, target: Optional[np.ndarray] = None,
                           tol: float = 1e-10, max_iter: int = 500) -> int:
        area = np.ascontiguousarray(area, np.float64)
        assert F.dtype == np.float64 and F.flags.c_contiguous
        tgt = None if target is None else np.ascontiguousarray(target, np.float64)
        sweeps = C.c_int32(0)
        check(self.lib.rsk_reciprocity_rowsum(self.handle, C.c_int32(area.shape[0]), ptr(area), ptr(tgt), ptr(F),
                                              C.c_double(tol), C.c_int32(max_iter), C.byref(sweeps)), "rsk_reciprocity_rowsum")
        return int(sweeps.value)


class DeviceGeometry:
    """Wraps ``rsk_geometry``: the raw meshes (float32 vertices, int32 faces) of a scene on the device."""

    def __init__(self, ctx: Context, verts: np.ndarray, vert_offset: np.ndarray, faces: np.ndarray, tri_offset: np.ndarray):
        self.ctx = ctx
        self.handle = C.c_void_p()
        vo = np.ascontiguousarray(vert_offset, np.int64)
        to = np.ascontiguousarray(tri_offset, np.int64)
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.int32)
        self.n_mesh = int(vo.shape[0]) - 1
        self.n_tri = int(to[-1])
        self.h2d_bytes = int(verts.nbytes + faces.nbytes + vo.nbytes + to.nbytes)
        check(ctx.lib.rsk_geometry_create(ctx.handle, C.c_int32(self.n_mesh), ptr(verts), ptr(vo), ptr(faces), ptr(to),
                                          C.byref(self.handle)), "rsk_geometry_create")

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.rsk_geometry_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceScene:
    """Wraps ``rsk_scene`` (triangles + GPU-built wide BVH)."""

    def __init__(self, ctx: Context, v0, e1, e2, normals, sid, n_surf: int, use_bvh: bool):
        self.ctx = ctx
        self.handle = C.c_void_p()
        self.n_tri = int(v0.shape[0])
        self.n_surf = int(n_surf)
        self.use_bvh = bool(use_bvh and self.n_tri > 0)
        arrs = [np.ascontiguousarray(a, np.float32) for a in (v0, e1, e2, normals)]
        sid = np.ascontiguousarray(sid, np.int32)
        check(ctx.lib.rsk_scene_create(ctx.handle, *(ptr(a) for a in arrs), ptr(sid), C.c_int64(self.n_tri),
                                       C.c_int32(n_surf), C.c_int32(1 if use_bvh else 0), C.byref(self.handle)), "rsk_scene_create")

    @classmethod
    def from_geometry(cls, geometry: "DeviceGeometry", use_bvh: bool) -> "DeviceScene":
        """``rsk_scene_from_geometry``: the triangle records are computed on the GPU from the raw meshes."""
        self = cls.__new__(cls)
        self.ctx = geometry.ctx
        self.handle = C.c_void_p()
        self.n_tri = geometry.n_tri
        self.n_surf = geometry.n_mesh
        self.use_bvh = bool(use_bvh and self.n_tri > 0)
        check(self.ctx.lib.rsk_scene_from_geometry(geometry.handle, C.c_int32(1 if use_bvh else 0), C.byref(self.handle)),
              "rsk_scene_from_geometry")
        return self

    def download_triangles(self):
        """(tri float32[n,12], normals float32[n,4]) in traversal order (test hook)."""
        tri = np.empty((self.n_tri, 12), np.float32)
        nrm = np.empty((self.n_tri, 4), np.float32)
        check(self.ctx.lib.rsk_scene_download_triangles(self.handle, ptr(tri), ptr(nrm)))
        return tri, nrm

    def info(self) -> dict:
        info = (C.c_int64 * 8)()
        check(self.ctx.lib.rsk_scene_info(self.handle, info))
        keys = ("n_tri", "n_surf", "use_bvh", "n_nodes", "node_bytes", "tri_bytes", "depth", "build_us")
        return dict(zip(keys, (int(x) for x in info)))

    def download_bvh(self):
        inf = self.info()
        nodes = np.empty((inf["n_nodes"], 96), np.uint8)
        order = np.empty(inf["n_tri"], np.int32)
        check(self.ctx.lib.rsk_scene_download_bvh(self.handle, ptr(nodes), ptr(order)))
        return nodes, order

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.rsk_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceEmitters:
    """Wraps ``rsk_emitters``: all emitter meshes of one (samples, rays, flip_faces) configuration."""

    def __init__(self, ctx: Context, tri_offset, tri_a, tri_e1, tri_e2, tri_u, tri_v, tri_n, tri_eps, cdf, g, rays: int):
        self.ctx = ctx
        self.handle = C.c_void_p()
        off = np.ascontiguousarray(tri_offset, np.int64)
        gs = np.ascontiguousarray(g, np.int32)
        f = [np.ascontiguousarray(a, np.float32) for a in (tri_a, tri_e1, tri_e2, tri_u, tri_v, tri_n, tri_eps, cdf)]
        self.n_emit = int(gs.shape[0])
        self.g = gs
        self.rays = int(rays)
        check(ctx.lib.rsk_emitters_create(ctx.handle, C.c_int32(self.n_emit), ptr(off), *(ptr(a) for a in f), ptr(gs),
                                          C.c_int32(rays), C.byref(self.handle)), "rsk_emitters_create")

    @classmethod
    def from_geometry(cls, geometry: "DeviceGeometry", samples, rays: int, flip_faces: bool):
        """``rsk_emitters_from_geometry``; returns (emitters, summary) with summary a MESH_SUMMARY_DTYPE array."""
        self = cls.__new__(cls)
        self.ctx = geometry.ctx
        self.handle = C.c_void_p()
        self.n_emit = geometry.n_mesh
        self.rays = int(rays)
        self.n_tri_total = geometry.n_tri
        summary = np.zeros(self.n_emit, MESH_SUMMARY_DTYPE)
        check(self.ctx.lib.rsk_emitters_from_geometry(geometry.handle, C.c_double(float(samples)), C.c_int32(int(rays)),
                                                      C.c_int32(1 if flip_faces else 0), C.byref(self.handle), ptr(summary)),
              "rsk_emitters_from_geometry")
        self.g = np.empty(self.n_emit, np.int32)
        check(self.ctx.lib.rsk_emitters_info(self.handle, ptr(self.g), None), "rsk_emitters_info")
        return self, summary

    def download_records(self, n_tri: int):
        """(records float32[n_tri,20], cdf float32[n_tri]) as stored on the device (test hook)."""
        rec = np.empty((n_tri, 20), np.float32)
        cdf = np.empty(n_tri, np.float32)
        check(self.ctx.lib.rsk_emitters_download_records(self.handle, ptr(rec), ptr(cdf)))
        return rec, cdf

    def download_tables(self, n: int, g: int):
        dims = np.empty((5, n), np.float32)
        gu = np.empty(g * g, np.float32)
        gv = np.empty(g * g, np.float32)
        check(self.ctx.lib.rsk_emitters_download_tables(self.handle, C.c_int64(n), ptr(dims), C.c_int32(g), ptr(gu), ptr(gv)))
        return dims, gu, gv

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.rsk_emitters_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def trace_rays(ctx: Context, scene: DeviceScene, em: DeviceEmitters, emitter: int, surf_active: np.ndarray,
               emit_sid: int, min_sid: int, cp: np.ndarray, mode: int = 0, first_ray: int = 0,
               n_rays: Optional[int] = None, want_rays: bool = True):
    """Per-ray hook (``rsk_trace_rays``): returns (orig, dirs, hit_sid, hit_front) for one emitter iteration."""
    n_once = int(em.g[emitter]) ** 2 * em.rays
    n = n_once - first_ray if n_rays is None else int(n_rays)
    orig = np.empty((n, 3), np.float32) if want_rays else None
    dirs = np.empty((n, 3), np.float32) if want_rays else None
    hit = np.empty(n, np.int32)
    front = np.empty(n, np.uint8)
    act = np.ascontiguousarray(surf_active, np.uint8)
    cpv = np.ascontiguousarray(cp, np.float32)
    check(ctx.lib.rsk_trace_rays(ctx.handle, scene.handle, em.handle, C.c_int32(emitter), ptr(act), C.c_int32(emit_sid),
                                 C.c_int32(min_sid), ptr(cpv), C.c_int32(mode), C.c_int64(first_ray), C.c_int64(n),
                                 ptr(orig), ptr(dirs), ptr(hit), ptr(front)), "rsk_trace_rays")
    return orig, dirs, hit, front


def emitter_costs(ctx: Context, scene: "DeviceScene", em: "DeviceEmitters", emit_ids, surf_active, emit_sid, min_sid, cp,
                  sample_rays: int = 2048):
    """``rsk_emitter_costs``: (ticks int64 [n], rays int64 [n]) -- SM clock ticks spent on the first ``sample_rays`` rays of
    every listed emitter (closest hit with the masks of a matrix solve)."""
    ids = np.ascontiguousarray(emit_ids, np.int32)
    n = int(ids.shape[0])
    act = np.ascontiguousarray(surf_active, np.uint8).reshape(n, scene.n_surf)
    es, ms = np.ascontiguousarray(emit_sid, np.int32), np.ascontiguousarray(min_sid, np.int32)
    cpr = np.ascontiguousarray(cp, np.float32).reshape(7)
    ticks, rays = np.zeros(n, np.int64), np.zeros(n, np.int64)
    check(ctx.lib.rsk_emitter_costs(ctx.handle, scene.handle, em.handle, ptr(ids), C.c_int32(n), ptr(act), ptr(es), ptr(ms), ptr(cpr),
                                    C.c_int64(sample_rays), ptr(ticks), ptr(rays)), "rsk_emitter_costs")
    return ticks, rays


class Solve:
    """Wraps ``rsk_solve`` for the matrix (``sky=False``) or sky variant."""

    def __init__(self, ctx: Context, scene: DeviceScene, em: DeviceEmitters, emit_ids, surf_active, cp_table, rot_base,
                 *, max_iters: int, min_iters: int, interval: int, tol_mode: str, tol: float,
                 emit_sid=None, min_sid=None, sky: bool = False, discrete: bool = False, ray_range=None):
        if tol_mode not in ("stderr", "delta"):
            raise ValueError(f"Unknown tol_mode: {tol_mode}")
        self.ctx, self.scene, self.em = ctx, scene, em
        self.sky, self.discrete = bool(sky), bool(discrete)
        self.handle = C.c_void_p()
        self.emit_ids = np.ascontiguousarray(emit_ids, np.int32)
        self.n_local = int(self.emit_ids.shape[0])
        act = np.ascontiguousarray(surf_active, np.uint8).reshape(self.n_local, scene.n_surf)
        cpt = np.ascontiguousarray(cp_table, np.float32).reshape(-1, 7)
        rb = np.ascontiguousarray(rot_base, np.int32)
        rr = None if ray_range is None else np.ascontiguousarray(ray_range, np.int64).reshape(self.n_local, 2)
        p = SolveParams(int(max_iters), int(min_iters), int(interval), 0 if tol_mode == "stderr" else 1, float(tol))
        if sky:
            check(ctx.lib.rsk_sky_begin(ctx.handle, scene.handle, em.handle, ptr(self.emit_ids), C.c_int32(self.n_local),
                                        ptr(act), ptr(cpt), C.c_int32(cpt.shape[0]), ptr(rb), ptr(rr), C.byref(p),
                                        C.c_int32(1 if discrete else 0), C.byref(self.handle)), "rsk_sky_begin")
        else:
            es = np.ascontiguousarray(emit_sid, np.int32)
            ms = np.ascontiguousarray(min_sid, np.int32)
            check(ctx.lib.rsk_matrix_begin(ctx.handle, scene.handle, em.handle, ptr(self.emit_ids), C.c_int32(self.n_local),
                                           ptr(act), ptr(es), ptr(ms), ptr(cpt), C.c_int32(cpt.shape[0]), ptr(rb), ptr(rr),
                                           C.byref(p), C.byref(self.handle)), "rsk_matrix_begin")

    def step(self, n_iters: int) -> int:
        n_active = C.c_int32(0)
        fn = self.ctx.lib.rsk_sky_step if self.sky else self.ctx.lib.rsk_matrix_step
        check(fn(self.handle, C.c_int32(n_iters), C.byref(n_active)), "solve step")
        return int(n_active.value)

    def enqueue_trace(self) -> None:
        check(self.ctx.lib.rsk_solve_enqueue_trace(self.handle), "rsk_solve_enqueue_trace")

    def enqueue_fold(self) -> None:
        check(self.ctx.lib.rsk_solve_enqueue_fold(self.handle), "rsk_solve_enqueue_fold")

    def poll(self) -> int:
        n_active = C.c_int32(0)
        check(self.ctx.lib.rsk_solve_poll(self.handle, C.byref(n_active)), "rsk_solve_poll")
        return int(n_active.value)

    def device_iter_tallies(self):
        """(device pointer, elements per job) of the uint64 per-iteration tally block."""
        p = C.c_void_p()
        n = C.c_int64(0)
        check(self.ctx.lib.rsk_solve_device_iter_tallies(self.handle, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def set_iter_tally_buffer(self, device_ptr: int, n_elements: int) -> None:
        check(self.ctx.lib.rsk_solve_set_iter_tally_buffer(self.handle, C.c_void_p(device_ptr), C.c_int64(n_elements)),
              "rsk_solve_set_iter_tally_buffer")

    def read_matrix(self, want_stderr: bool = False):
        ns = self.scene.n_surf
        hf = np.zeros((self.n_local, ns), np.int64)
        hb = np.zeros((self.n_local, ns), np.int64)
        iters = np.zeros(self.n_local, np.int32)
        total = np.zeros(self.n_local, np.int64)
        sf = np.zeros((self.n_local, ns), np.float64) if want_stderr else None
        sb = np.zeros((self.n_local, ns), np.float64) if want_stderr else None
        check(self.ctx.lib.rsk_matrix_read(self.handle, ptr(hf), ptr(hb), ptr(iters), ptr(total), ptr(sf), ptr(sb)), "rsk_matrix_read")
        return hf, hb, iters, total, sf, sb

    def read_block(self):
        """(tallies int64 [n_local, n_hist], iterations, total rays) in one contiguous device-to-host copy."""
        nh = (145 if self.discrete else 1) if self.sky else 2 * self.scene.n_surf
        tallies = np.empty((self.n_local, nh), np.int64)
        iters = np.zeros(self.n_local, np.int32)
        total = np.zeros(self.n_local, np.int64)
        check(self.ctx.lib.rsk_solve_read_block(self.handle, ptr(tallies), ptr(iters), ptr(total)), "rsk_solve_read_block")
        return tallies, iters, total

    def read_block_view(self):
        """``read_block`` without the second host copy: the tallies are a NumPy view of the context's pinned staging
        area, valid until the next staged download on this context."""
        nh = (145 if self.discrete else 1) if self.sky else 2 * self.scene.n_surf
        iters = np.zeros(self.n_local, np.int32)
        total = np.zeros(self.n_local, np.int64)
        view = C.POINTER(C.c_int64)()
        check(self.ctx.lib.rsk_solve_read_block_view(self.handle, C.byref(view), ptr(iters), ptr(total)), "rsk_solve_read_block_view")
        if self.n_local * nh == 0:
            return np.zeros((self.n_local, nh), np.int64), iters, total
        return np.ctypeslib.as_array(view, shape=(self.n_local, nh)), iters, total

    def read_csr(self):
        """The result rows in compressed form, built on the device: (row_ptr int64 [n_local + 1], cols int32 [nnz],
        vals float64 [nnz]) with vals = tally / total rays of the job, non-zero bins only, columns ascending."""
        row_ptr = np.zeros(self.n_local + 1, np.int64)
        check(self.ctx.lib.rsk_solve_csr(self.handle, ptr(row_ptr)), "rsk_solve_csr")
        return _fetch_csr(self.ctx, row_ptr)

    def allreduce_iter_tallies(self, n_jobs: int) -> None:
        """Sum the iteration tallies of the first ``n_jobs`` (ray-split) jobs over the context's communicator."""
        check(self.ctx.lib.rsk_solve_allreduce_iter_tallies(self.handle, C.c_int32(n_jobs)), "rsk_solve_allreduce_iter_tallies")

    def read_counters(self):
        """(iterations int32 [n_local], total rays int64 [n_local]) without the tally block."""
        iters = np.zeros(self.n_local, np.int32)
        total = np.zeros(self.n_local, np.int64)
        check(self.ctx.lib.rsk_solve_read_block(self.handle, None, ptr(iters), ptr(total)), "rsk_solve_read_block")
        return iters, total

    def read_sky(self):
        nb = 145 if self.discrete else 1
        counts = np.zeros((self.n_local, nb), np.int64)
        iters = np.zeros(self.n_local, np.int32)
        total = np.zeros(self.n_local, np.int64)
        check(self.ctx.lib.rsk_sky_read(self.handle, ptr(counts), ptr(iters), ptr(total)), "rsk_sky_read")
        return counts, iters, total

    def device_tallies(self):
        p = C.c_void_p()
        n = C.c_int64(0)
        check(self.ctx.lib.rsk_matrix_device_tallies(self.handle, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def rays_traced(self) -> int:
        n = C.c_int64(0)
        check(self.ctx.lib.rsk_solve_rays_traced(self.handle, C.byref(n)))
        return int(n.value)

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.rsk_solve_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _fetch_csr(ctx: "Context", row_ptr: np.ndarray):
    nnz = int(row_ptr[-1])
    cols = np.empty(nnz, np.int32)
    vals = np.empty(nnz, np.float64)
    check(ctx.lib.rsk_csr_fetch(ctx.handle, ptr(cols), ptr(vals)), "rsk_csr_fetch")
    return row_ptr, cols, vals


class TallyBlock:
    """Wraps ``rsk_tally_block``: the device-resident [n_rows, n_cols] int64 block in which the ranks of a sharded solve
    assemble and sum their results (scatter kernel + NCCL all-reduce + pinned download, all inside the library)."""

    def __init__(self, ctx: Context, n_rows: int, n_cols: int):
        self.ctx, self.shape = ctx, (int(n_rows), int(n_cols))
        self.handle = C.c_void_p()
        check(ctx.lib.rsk_tally_block_create(ctx.handle, C.c_int64(n_rows), C.c_int64(n_cols), C.byref(self.handle)), "rsk_tally_block_create")

    def add_solve(self, solve, keep: Optional[np.ndarray] = None) -> None:
        k = None if keep is None else np.ascontiguousarray(keep, np.uint8)
        check(self.ctx.lib.rsk_tally_block_add_solve(self.handle, solve.handle, ptr(k)), "rsk_tally_block_add_solve")

    def allreduce(self) -> None:
        check(self.ctx.lib.rsk_tally_block_allreduce(self.handle), "rsk_tally_block_allreduce")

    def device_pointer(self):
        p, n = C.c_void_p(), C.c_int64(0)
        check(self.ctx.lib.rsk_tally_block_device(self.handle, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def read_csr(self, total_rays: np.ndarray):
        """Rows of the (rank-summed) block in compressed form: vals = tally / total_rays[row]."""
        tot = np.ascontiguousarray(total_rays, np.int64)
        assert tot.shape[0] == self.shape[0]
        row_ptr = np.zeros(self.shape[0] + 1, np.int64)
        check(self.ctx.lib.rsk_tally_block_csr(self.handle, ptr(tot), ptr(row_ptr)), "rsk_tally_block_csr")
        return _fetch_csr(self.ctx, row_ptr)

    def download(self, copy: bool = False) -> np.ndarray:
        """The block on the host: a view of the context's pinned staging area (valid until the next staged download on
        this context) or, with ``copy``, an array of its own."""
        if self.shape[0] * self.shape[1] == 0:
            return np.zeros(self.shape, np.int64)
        if copy:
            out = np.empty(self.shape, np.int64)
            check(self.ctx.lib.rsk_tally_block_download(self.handle, ptr(out), None), "rsk_tally_block_download")
            return out
        view = C.POINTER(C.c_int64)()
        check(self.ctx.lib.rsk_tally_block_download(self.handle, None, C.byref(view)), "rsk_tally_block_download")
        return np.ctypeslib.as_array(view, shape=self.shape)

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.rsk_tally_block_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _SolvePart:
    """Read-only view of one side of a dual solve (shares Solve.read_block)."""

    def __init__(self, ctx, scene, handle, n_local, sky, discrete):
        self.ctx, self.scene, self.handle, self.n_local, self.sky, self.discrete = ctx, scene, handle, n_local, sky, discrete

    read_block = Solve.read_block
    read_block_view = Solve.read_block_view
    read_counters = Solve.read_counters
    read_csr = Solve.read_csr
    allreduce_iter_tallies = Solve.allreduce_iter_tallies
    enqueue_fold = Solve.enqueue_fold
    poll = Solve.poll
    device_iter_tallies = Solve.device_iter_tallies
    set_iter_tally_buffer = Solve.set_iter_tally_buffer


class DualSolve:
    """Wraps a shared-ray solve (``rsk_dual_*``): one traversal per ray feeds the matrix and the sky tallies."""

    def __init__(self, ctx: Context, scene: DeviceScene, em: DeviceEmitters, emit_ids, surf_active, cp_table, rot_base,
                 emit_sid, min_sid, matrix: dict, sky: dict, discrete: bool, ray_range=None):
        for side in (matrix, sky):
            if side["tol_mode"] not in ("stderr", "delta"):
                raise ValueError(f"Unknown tol_mode: {side['tol_mode']}")
        self.ctx, self.scene = ctx, scene
        self.handle = C.c_void_p()
        ids = np.ascontiguousarray(emit_ids, np.int32)
        self.n_local = int(ids.shape[0])
        act = np.ascontiguousarray(surf_active, np.uint8).reshape(self.n_local, scene.n_surf)
        cpt = np.ascontiguousarray(cp_table, np.float32).reshape(-1, 7)
        rb = np.ascontiguousarray(rot_base, np.int32)
        es = np.ascontiguousarray(emit_sid, np.int32)
        ms = np.ascontiguousarray(min_sid, np.int32)

        def pack(d):
            return SolveParams(int(d["max_iters"]), int(d["min_iters"]), int(d["interval"]), 0 if d["tol_mode"] == "stderr" else 1, float(d["tol"]))

        pm, pk = pack(matrix), pack(sky)
        rr = None if ray_range is None else np.ascontiguousarray(ray_range, np.int64).reshape(self.n_local, 2)
        check(ctx.lib.rsk_dual_begin_sliced(ctx.handle, scene.handle, em.handle, ptr(ids), C.c_int32(self.n_local), ptr(act), ptr(es),
                                            ptr(ms), ptr(cpt), C.c_int32(cpt.shape[0]), ptr(rb), ptr(rr), C.byref(pm), C.byref(pk),
                                            C.c_int32(1 if discrete else 0), C.byref(self.handle)), "rsk_dual_begin_sliced")
        sky_handle = C.c_void_p()
        check(ctx.lib.rsk_dual_sky_part(self.handle, C.byref(sky_handle)), "rsk_dual_sky_part")
        self.matrix_part = _SolvePart(ctx, scene, self.handle, self.n_local, False, False)
        self.sky_part = _SolvePart(ctx, scene, sky_handle, self.n_local, True, bool(discrete))

    def enqueue_trace(self) -> None:
        """One traversal per ray for both sides (split-phase stepping; fold each part afterwards)."""
        check(self.ctx.lib.rsk_solve_enqueue_trace(self.handle), "rsk_solve_enqueue_trace")

    def step(self, n_iters: int) -> int:
        n_active = C.c_int32(0)
        check(self.ctx.lib.rsk_dual_step(self.handle, C.c_int32(n_iters), C.byref(n_active)), "rsk_dual_step")
        return int(n_active.value)

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.rsk_solve_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
