"""Parameter containers of the public API -- field for field the reference's ``MatrixParams`` /
``SkyParams`` (reference src/raystrack/params.py:7-129), so existing call sites keep working."""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Any, Dict


@dataclass
class MatrixParams:
    """Configuration of a scene-to-scene view-factor solve (reference params.py:7-68).

    samples: QMC grid density (cells per unit area^0.5); rays: rays per cell; seed: base seed of the
    per-(emitter, iteration) Cranley-Patterson rotations; bvh: "auto" | "off" | "builtin";
    device: "auto" | "gpu" | "cpu" (accepted for compatibility -- this package always runs on the B200;
    "cpu" selects the reference's CPU convergence schedule, i.e. a check after every iteration);
    cuda_async / gpu_raygen: accepted, ignored (rays never leave the GPU); max_iters / min_iters / tol /
    tol_mode ("stderr" | "delta") / convergence_interval: stopping rule; reciprocity: fill F_ji from F_ij;
    enforce_reciprocity_rowsum: symmetric diagonal scaling to unit row sums; flip_faces: flip emitter winding.
    """
    samples: int = 16
    rays: int = 128
    seed: int = 1
    bvh: str = "auto"
    device: str = "auto"
    cuda_async: bool = True
    gpu_raygen: bool = True
    max_iters: int = 100
    tol: float = 1e-4
    tol_mode: str = "stderr"
    min_iters: int = 5
    convergence_interval: int = 1
    reciprocity: bool = True
    enforce_reciprocity_rowsum: bool = False
    flip_faces: bool = False

    def as_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "MatrixParams":
        return cls(**data)


@dataclass
class SkyParams:
    """Configuration of a sky view-factor solve (reference params.py:71-126); ``discrete=True`` returns the
    145 Tregenza patches, otherwise one merged "Sky" entry."""
    samples: int = 16
    rays: int = 128
    seed: int = 1
    bvh: str = "auto"
    device: str = "auto"
    cuda_async: bool = True
    gpu_raygen: bool = True
    max_iters: int = 100
    tol: float = 1e-4
    tol_mode: str = "stderr"
    min_iters: int = 5
    convergence_interval: int = 1
    discrete: bool = False

    def as_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "SkyParams":
        return cls(**data)


__all__ = ["MatrixParams", "SkyParams"]
