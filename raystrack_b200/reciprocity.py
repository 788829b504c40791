"""Reciprocity / row-sum enforcement (reference utils/helpers.py:14-257).

``enforce_reciprocity_and_rowsum`` keeps the reference's dict-in / dict-out contract; its dense core
(symmetrise G = 1/2 (A F + (A F)^T), up to 500 sweeps of symmetric diagonal scaling, F' = D G D / A) runs on
the GPU through ``rsk_reciprocity_rowsum``."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from . import _native

Mesh = Tuple[str, np.ndarray, np.ndarray]


def _mesh_areas(meshes: List[Mesh]) -> np.ndarray:
    out = []
    for _, V, F in meshes:
        a = 0.5 * np.linalg.norm(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), axis=1)
        out.append(float(a.sum()))
    return np.asarray(out, np.float64)


def _base(key: str) -> str:
    if key.endswith("_front"):
        return key[:-6]
    if key.endswith("_back"):
        return key[:-5]
    return key


def _totals_matrix(result, names) -> np.ndarray:
    index = {n: i for i, n in enumerate(names)}
    F = np.zeros((len(names), len(names)), np.float64)
    for i, sname in enumerate(names):
        row = result.get(sname, {})
        if not isinstance(row, dict):
            continue
        acc: Dict[str, float] = {}
        for k, v in row.items():
            b = _base(k)
            acc[b] = acc.get(b, 0.0) + float(v)
        for b, v in acc.items():
            j = index.get(b)
            if j is not None:
                F[i, j] = v
    return F


def _write_back(result, names, Fp: np.ndarray) -> None:
    """Split the adjusted totals back into front/back proportionally to the old split (helpers.py:98-140)."""
    for i, sname in enumerate(names):
        row = result.get(sname, {})
        fb: Dict[str, Tuple[float, float]] = {}
        for k, v in row.items():
            f, b = fb.get(_base(k), (0.0, 0.0))
            if k.endswith("_front"):
                fb[_base(k)] = (f + float(v), b)
            else:
                fb[_base(k)] = (f, b + float(v))
        for j, rname in enumerate(names):
            t_new = float(max(Fp[i, j], 0.0))
            f, b = fb.get(rname, (0.0, 0.0))
            t_old = f + b
            if t_old > 0.0:
                s = t_new / t_old
                nf, nb = f * s, b * s
            else:
                nf, nb = 0.0, t_new
            for key, val in ((f"{rname}_front", nf), (f"{rname}_back", nb)):
                if val > 0.0:
                    row[key] = val
                elif key in row:
                    del row[key]
        result[sname] = row


def enforce_reciprocity_and_rowsum(result: Dict[str, Dict[str, float]], meshes: List[Mesh], areas: Optional[List[float]],
                                   row_targets: Optional[Iterable[float]] = None, tol: float = 1e-10, max_iter: int = 500,
                                   *, ctx: Optional[_native.Context] = None) -> None:
    """In place: rows sum to their targets (default 1) and A_i F_ij = A_j F_ji (helpers.py:14-140)."""
    names = [m[0] for m in meshes]
    A = _mesh_areas(meshes) if areas is None else np.asarray(areas, np.float64)
    target = None
    if row_targets is not None:
        target = np.asarray(list(row_targets), np.float64)
        if target.shape != A.shape:
            raise ValueError("row_targets must match number of meshes")
    F = np.ascontiguousarray(_totals_matrix(result, names))
    ctx = ctx or _native.Context.for_device()
    ctx.reciprocity_rowsum(A, F, target, tol=tol, max_iter=max_iter)
    _write_back(result, names, F)


def enforce_reciprocity_only(result: Dict[str, Dict[str, float]], meshes: List[Mesh], tol: float = 1e-12) -> None:
    """In place: A_i F_ij = A_j F_ji without touching the row sums (reference helpers.py:143-257):
    ``g = (A_i F_ij + A_j F_ji) / 2``, ``F'_ij = g / A_i``, ``F'_ji = g / A_j``; pairs with both entries <= tol are
    cleared; front/back keep their old proportions; entries <= tol are removed."""
    if tol <= 0.0:
        tol = 1e-12
    names = [m[0] for m in meshes]
    n = len(names)
    A = _mesh_areas(meshes)
    F = _totals_matrix(result, names)
    G = 0.5 * (A[:, None] * F + (A[:, None] * F).T)               # symmetric by construction (same summation order)
    with np.errstate(divide="ignore", invalid="ignore"):
        Fn = np.where(A[:, None] > 0.0, np.maximum(G / A[:, None], 0.0), 0.0)
    dead = (F <= tol) & (F.T <= tol)
    Fn[dead] = 0.0
    for i, sname in enumerate(names):
        row = result.get(sname, {})
        if not isinstance(row, dict):
            row = {}
        fb: Dict[str, Tuple[float, float]] = {}
        for k, v in row.items():
            f, b = fb.get(_base(k), (0.0, 0.0))
            if k.endswith("_front"):
                fb[_base(k)] = (f + float(v), b)
            else:
                fb[_base(k)] = (f, b + float(v))
        touched = set(np.nonzero(Fn[i])[0].tolist())
        index = {nm: j for j, nm in enumerate(names)}
        touched.update(index[b] for b in fb if b in index)
        touched.discard(i)
        for j in sorted(touched):
            rname = names[j]
            t_new = float(max(Fn[i, j], 0.0))
            f, b = fb.get(rname, (0.0, 0.0))
            t_old = f + b
            if t_old > 0.0:
                s_ = t_new / t_old
                nf, nb = f * s_, b * s_
            else:
                nf, nb = 0.0, t_new
            for key, val in ((f"{rname}_front", nf), (f"{rname}_back", nb)):
                if val > tol:
                    row[key] = val
                elif key in row:
                    del row[key]
        result[sname] = row


__all__ = ["enforce_reciprocity_and_rowsum", "enforce_reciprocity_only"]
