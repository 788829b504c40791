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


def _write_json(payload, save_path: str, **dump_options) -> str:
    path = _target_path(save_path)
    with path.open("w", encoding="utf-8") as fh:
        json.dump(payload, fh, ensure_ascii=False, indent=2, **dump_options)
    return str(path.resolve())


def _read_json(load_path: str):
    path = Path(load_path)
    if not path.exists():
        raise FileNotFoundError(f"File not found: {load_path}")
    with path.open("r", encoding="utf-8") as fh:
        return json.load(fh)


def _require(condition: bool, error, message: str) -> None:
    if not condition:
        raise error(message)


def _mesh_arrays(vertices, faces, vertex_message: str, face_message: str):
    """float32 (N,3) vertices and int32 (M,3) faces, or ValueError with the caller's wording."""
    V, F = np.asarray(vertices, dtype=np.float32), np.asarray(faces, dtype=np.int32)
    _require(V.ndim == 2 and V.shape[1] == 3, ValueError, vertex_message)
    _require(F.ndim == 2 and F.shape[1] == 3, ValueError, face_message)
    return V, F


def save_vf_matrix_json(vf_matrix: VFInput, save_path: str, *, strip_dir: bool = False) -> str:
    """Write a view-factor matrix as JSON (io.py:69-116): exact zeros are dropped, ``strip_dir`` sums
    ``*_front``/``*_back`` per base receiver, keys sorted, indent 2; returns the resolved path."""
    flat = merge_vf_matrix(vf_matrix)
    for sender, row in flat.items():                     # validate everything before the file is touched
        _require(isinstance(sender, str), TypeError, "Sender keys must be strings")
        _require(isinstance(row, dict), TypeError, f"Row for '{sender}' must be a dict mapping receiver->value")
        for recv, val in row.items():
            _require(isinstance(recv, str), TypeError, "Receiver keys must be strings")
            try:
                float(val)
            except Exception:
                raise TypeError(f"Value for '{sender}'->'{recv}' must be numeric")
    cleaned: VFDict = {}
    for sender, row in flat.items():
        kept = cleaned.setdefault(sender, {})
        for key, value in row.items():
            if float(value) != 0.0:
                name = _base_name(key) if strip_dir else key
                kept[name] = kept.get(name, 0.0) + float(value)
    return _write_json(cleaned, save_path, sort_keys=True)


def load_vf_matrix_json(load_path: str) -> VFDict:
    """Read a matrix written by :func:`save_vf_matrix_json` (io.py:119-146)."""
    data = _read_json(load_path)
    _require(isinstance(data, dict), TypeError, "Invalid view-factor JSON: expected an object")
    for sender, row in data.items():
        _require(isinstance(row, dict), TypeError, f"Row for '{sender}' must be an object")
    return {str(sender): {str(k): float(v) for k, v in row.items()} for sender, row in data.items()}


def save_meshes_json(meshes: List[Mesh], save_path: str) -> str:
    """``{"meshes": [{"name", "vertices", "faces"}, ...]}`` (io.py:153-199)."""
    _require(isinstance(meshes, list), TypeError, "meshes must be a list of (name, V, F) tuples")
    entries = []
    for item in meshes:
        _require(isinstance(item, tuple) and len(item) == 3, TypeError, "Each mesh must be a (name, V, F) tuple")
        name = item[0]
        _require(isinstance(name, str) and name.strip() != "", TypeError, "Mesh name must be a non-empty string")
        V, F = _mesh_arrays(item[1], item[2], f"Vertices for '{name}' must have shape (N,3)",
                            f"Faces for '{name}' must have shape (M,3) of triangles")
        entries.append({"name": name, "vertices": V.tolist(), "faces": F.tolist()})
    return _write_json({"meshes": entries}, save_path)


def load_meshes_json(load_path: str) -> List[Mesh]:
    """Inverse of :func:`save_meshes_json` (io.py:202-238): float32 vertices, int32 faces."""
    data = _read_json(load_path)
    _require(isinstance(data, dict) and "meshes" in data, TypeError, "Invalid mesh JSON: expected an object with 'meshes' list")
    _require(isinstance(data["meshes"], list), TypeError, "'meshes' must be a list")
    out: List[Mesh] = []
    for i, entry in enumerate(data["meshes"]):
        _require(isinstance(entry, dict), TypeError, "Each entry in 'meshes' must be an object")
        name = entry.get("name")
        _require(isinstance(name, str) and name.strip() != "", TypeError, f"Entry {i}: 'name' must be a non-empty string")
        out.append((name, *_mesh_arrays(entry.get("vertices"), entry.get("faces"), f"Entry {i} ('{name}'): vertices must have shape (N,3)",
                                        f"Entry {i} ('{name}'): faces must have shape (M,3)")))
    return out


__all__ = ["save_vf_matrix_json", "load_vf_matrix_json", "save_meshes_json", "load_meshes_json", "merge_vf_matrix"]
