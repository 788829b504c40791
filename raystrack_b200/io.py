"""JSON persistence of view-factor matrices and meshes, behaviour-compatible with the reference's
``raystrack.io`` (src/raystrack/io.py:23-238): same file layout, same key handling, same errors."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Tuple, Union

import numpy as np

VFDict = Dict[str, Dict[str, float]]
VFInput = Union[VFDict, List[VFDict]]
Mesh = Tuple[str, np.ndarray, np.ndarray]


def _base_name(key: str) -> str:
    for suffix in ("_front", "_back"):
        if key.endswith(suffix):
            return key[: -len(suffix)]
    return key


def merge_vf_matrix(vf_matrix: VFInput) -> VFDict:
    """One ``{sender: {receiver: value}}`` dict from a dict or a list of dicts; rows of the same sender are
    merged, later entries win per receiver (io.py:23-66)."""
    if isinstance(vf_matrix, dict):
        return vf_matrix
    if isinstance(vf_matrix, list):
        merged: VFDict = {}
        for part in vf_matrix:
            if not isinstance(part, dict):
                raise TypeError("All elements of vf_matrix list must be dicts")
            for sender, row in part.items():
                merged.setdefault(sender, {}).update(row)
        return merged
    raise TypeError("vf_matrix must be a dict or list of dicts")


def _target_path(save_path: str) -> Path:
    path = Path(save_path)
    if path.suffix.lower() == "":
        path = path.with_suffix(".json")
    if path.parent and not path.parent.exists():
        path.parent.mkdir(parents=True, exist_ok=True)
    return path


def save_vf_matrix_json(vf_matrix: VFInput, save_path: str, *, strip_dir: bool = False) -> str:
    """Write a view-factor matrix as JSON (io.py:69-116): exact zeros are dropped, ``strip_dir`` sums
    ``*_front``/``*_back`` per base receiver, keys sorted, indent 2; returns the resolved path."""
    flat = merge_vf_matrix(vf_matrix)
    for sender, row in flat.items():
        if not isinstance(sender, str):
            raise TypeError("Sender keys must be strings")
        if not isinstance(row, dict):
            raise TypeError(f"Row for '{sender}' must be a dict mapping receiver->value")
        for recv, val in row.items():
            if not isinstance(recv, str):
                raise TypeError("Receiver keys must be strings")
            try:
                float(val)
            except Exception:
                raise TypeError(f"Value for '{sender}'->'{recv}' must be numeric")
    path = _target_path(save_path)
    cleaned: VFDict = {}
    for sender, row in flat.items():
        kept: Dict[str, float] = {}
        for key, value in row.items():
            v = float(value)
            if v == 0.0:
                continue
            name = _base_name(key) if strip_dir else key
            kept[name] = kept.get(name, 0.0) + v
        cleaned[sender] = kept
    with path.open("w", encoding="utf-8") as fh:
        json.dump(cleaned, fh, ensure_ascii=False, indent=2, sort_keys=True)
    return str(path.resolve())


def load_vf_matrix_json(load_path: str) -> VFDict:
    """Read a matrix written by :func:`save_vf_matrix_json` (io.py:119-146)."""
    path = Path(load_path)
    if not path.exists():
        raise FileNotFoundError(f"File not found: {load_path}")
    with path.open("r", encoding="utf-8") as fh:
        data = json.load(fh)
    if not isinstance(data, dict):
        raise TypeError("Invalid view-factor JSON: expected an object")
    out: VFDict = {}
    for sender, row in data.items():
        if not isinstance(row, dict):
            raise TypeError(f"Row for '{sender}' must be an object")
        out[str(sender)] = {str(k): float(v) for k, v in row.items()}
    return out


def save_meshes_json(meshes: List[Mesh], save_path: str) -> str:
    """``{"meshes": [{"name", "vertices", "faces"}, ...]}`` (io.py:153-199)."""
    if not isinstance(meshes, list):
        raise TypeError("meshes must be a list of (name, V, F) tuples")
    payload = {"meshes": []}
    for item in meshes:
        if not (isinstance(item, tuple) and len(item) == 3):
            raise TypeError("Each mesh must be a (name, V, F) tuple")
        name, V, F = item
        if not isinstance(name, str) or name.strip() == "":
            raise TypeError("Mesh name must be a non-empty string")
        V = np.asarray(V, dtype=np.float32)
        F = np.asarray(F, dtype=np.int32)
        if V.ndim != 2 or V.shape[1] != 3:
            raise ValueError(f"Vertices for '{name}' must have shape (N,3)")
        if F.ndim != 2 or F.shape[1] != 3:
            raise ValueError(f"Faces for '{name}' must have shape (M,3) of triangles")
        payload["meshes"].append({"name": name, "vertices": V.tolist(), "faces": F.tolist()})
    path = _target_path(save_path)
    with path.open("w", encoding="utf-8") as fh:
        json.dump(payload, fh, ensure_ascii=False, indent=2)
    return str(path.resolve())


def load_meshes_json(load_path: str) -> List[Mesh]:
    """Inverse of :func:`save_meshes_json` (io.py:202-238): float32 vertices, int32 faces."""
    path = Path(load_path)
    if not path.exists():
        raise FileNotFoundError(f"File not found: {load_path}")
    with path.open("r", encoding="utf-8") as fh:
        data = json.load(fh)
    if not isinstance(data, dict) or "meshes" not in data:
        raise TypeError("Invalid mesh JSON: expected an object with 'meshes' list")
    if not isinstance(data["meshes"], list):
        raise TypeError("'meshes' must be a list")
    out: List[Mesh] = []
    for i, entry in enumerate(data["meshes"]):
        if not isinstance(entry, dict):
            raise TypeError("Each entry in 'meshes' must be an object")
        name = entry.get("name")
        if not isinstance(name, str) or name.strip() == "":
            raise TypeError(f"Entry {i}: 'name' must be a non-empty string")
        V = np.asarray(entry.get("vertices"), dtype=np.float32)
        F = np.asarray(entry.get("faces"), dtype=np.int32)
        if V.ndim != 2 or V.shape[1] != 3:
            raise ValueError(f"Entry {i} ('{name}'): vertices must have shape (N,3)")
        if F.ndim != 2 or F.shape[1] != 3:
            raise ValueError(f"Entry {i} ('{name}'): faces must have shape (M,3)")
        out.append((name, V, F))
    return out


__all__ = ["save_vf_matrix_json", "load_vf_matrix_json", "save_meshes_json", "load_meshes_json", "merge_vf_matrix"]
