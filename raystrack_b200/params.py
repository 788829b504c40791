"""Parameter containers of the public API.

``MatrixParams`` and ``SkyParams`` accept exactly the fields, order and defaults of the reference's classes
(reference src/raystrack/params.py:47-61 and :107-119), so existing call sites -- keyword or positional -- keep
working.  The twelve fields the two solves have in common live in one base class here.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Any, Dict


@dataclass
class _SolveParams:
    """Sampling, execution and stopping-rule settings shared by both solves.

    samples              QMC cell density: an emitter of area A is sampled on a g x g grid, g = max(ceil(sqrt(A * samples)), 4)
    rays                 rays per cell and iteration
    seed                 base seed of the Cranley-Patterson rotations (emitter i, iteration k uses seed + i + k)
    bvh                  "auto" (BVH from 512 faces on) | "off" | "builtin"
    device               "auto" | "gpu" | "cpu": accepted for compatibility; every value runs on the B200.  "cpu" selects
                         the reference's CPU convergence schedule (a check after every iteration)
    cuda_async           accepted, ignored (everything is stream-ordered)
    gpu_raygen           accepted, ignored (rays are generated in registers and never leave the GPU)
    max_iters, min_iters bounds on the number of Monte-Carlo iterations (statistical replicates) per emitter
    tol, tol_mode        "stderr": stop when the replicate standard error of every tracked bin is <= tol;
                         "delta": stop when the cumulative estimate moved by < tol since the previous checkpoint
    convergence_interval check the stopping rule every N iterations after ``min_iters``
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

    def as_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @classmethod
    def from_dict(cls, data: Dict[str, Any]):
        return cls(**data)


@dataclass
class MatrixParams(_SolveParams):
    """Scene-to-scene view-factor solve (reference params.py:7-68).

    reciprocity                 trace only receivers j > i and fill F_ji = F_ij * A_i / A_j (front hits)
    enforce_reciprocity_rowsum  afterwards scale symmetrically so that rows sum to 1 and A_i F_ij = A_j F_ji
    flip_faces                  emit from the reversed winding of every emitter (inside-enclosure solves)
    """
    reciprocity: bool = True
    enforce_reciprocity_rowsum: bool = False
    flip_faces: bool = False


@dataclass
class SkyParams(_SolveParams):
    """Sky view-factor solve (reference params.py:71-126).

    discrete   True: the 145 Tregenza patches ("Sky_Patch_1" .. "Sky_Patch_145"); False: one merged "Sky" entry
    """
    discrete: bool = False


__all__ = ["MatrixParams", "SkyParams"]
