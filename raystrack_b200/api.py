"""Outside workflow: scene matrix + sky + residual so that scene + sky + rest = 1 per emitter (reference
src/raystrack/api.py:24-194)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .main import outside_workflow_shareable, view_factor_matrix, view_factor_matrix_and_sky, view_factor_to_tregenza_sky
from .params import MatrixParams, SkyParams
from .prepared import PreparedSolver
from .reciprocity import enforce_reciprocity_and_rowsum, enforce_reciprocity_only

Mesh = Tuple[str, np.ndarray, np.ndarray]
VF = Dict[str, Dict[str, float]]


def _row_sum(row: Dict[str, float]) -> float:
    return float(sum(float(v) for v in row.values()))


def _sky_total(row: Dict[str, float], discrete: bool) -> float:
    return float(sum(float(v) for v in row.values())) if discrete else float(row.get("Sky", 0.0))


def _clip_sky(sky_row: Dict[str, float], scene_sum: float, sky_total: float, discrete: bool, zero_when_full: bool):
    """Scale the sky entries of one emitter so that scene + sky <= 1 (api.py:138-153 and 171-186)."""
    allowed = max(0.0, 1.0 - scene_sum)
    if zero_when_full and allowed <= 0.0:
        return {k: 0.0 for k in sky_row}, 0.0
    scale = min(1.0, allowed / sky_total) if sky_total else 0.0
    if discrete:
        sky_row = {k: float(v) * scale for k, v in sky_row.items()}
        return sky_row, float(sum(float(v) for v in sky_row.values()))
    sky_row["Sky"] = float(sky_row.get("Sky", 0.0)) * scale
    return sky_row, float(sky_row.get("Sky", 0.0))


def view_factor_outside_workflow(meshes: List[Mesh], *, matrix_params: MatrixParams, sky_params: SkyParams,
                                 prepared: Optional[PreparedSolver] = None) -> Tuple[VF, VF, VF]:
    """Returns ``(vf_scene, sky_vf, rest_vf)``; compatible parameter sets share one ray set
    (:func:`view_factor_matrix_and_sky`), otherwise the two solves run separately (api.py:101-110)."""
    if not isinstance(matrix_params, MatrixParams):
        raise TypeError("matrix_params must be a MatrixParams instance")
    if not isinstance(sky_params, SkyParams):
        raise TypeError("sky_params must be a SkyParams instance")
    threshold = 1e-6
    enforce_scene = bool(matrix_params.enforce_reciprocity_rowsum)
    reciprocity_flag = bool(matrix_params.reciprocity)
    discrete = bool(sky_params.discrete)
    mp = MatrixParams(**matrix_params.as_dict())
    mp.enforce_reciprocity_rowsum = False                         # rows are enforced below, after the sky is known

    if outside_workflow_shareable(mp, sky_params):
        vf_scene, sky_vf = view_factor_matrix_and_sky(meshes, matrix_params=mp, sky_params=sky_params, prepared=prepared)
    else:
        vf_scene = view_factor_matrix(meshes, params=mp, prepared=prepared)
        sky_vf = view_factor_to_tregenza_sky(meshes, params=sky_params, prepared=prepared)

    names = [name for name, _, _ in meshes]
    if enforce_scene:                                             # api.py:120-122
        targets = [max(0.0, _row_sum(vf_scene.get(n, {}))) for n in names]
        enforce_reciprocity_and_rowsum(vf_scene, meshes, None, row_targets=targets)

    sky_totals = {n: 0.0 for n in names}
    for name in names:                                            # api.py:129-155
        scene_sum = _row_sum(vf_scene.get(name, {}))
        sky_row = dict(sky_vf.get(name, {}))
        total = _sky_total(sky_row, discrete)
        if scene_sum + total > 1.0 + threshold:
            if total > 0.0:
                sky_row, total = _clip_sky(sky_row, scene_sum, total, discrete, zero_when_full=False)
                sky_vf[name] = sky_row
            else:
                total = 0.0
        sky_totals[name] = max(0.0, total)

    if enforce_scene:                                             # api.py:157-161
        targets = [max(0.0, 1.0 - sky_totals.get(n, 0.0)) for n in names]
        enforce_reciprocity_and_rowsum(vf_scene, meshes, None, row_targets=targets)
    elif reciprocity_flag:
        enforce_reciprocity_only(vf_scene, meshes)

    rest_vf: VF = {}
    for name in names:                                            # api.py:163-192
        scene_sum = _row_sum(vf_scene.get(name, {}))
        sky_row = dict(sky_vf.get(name, {}))
        total = _sky_total(sky_row, discrete)
        combined = scene_sum + total
        if combined > 1.0 + threshold and total > 0.0:
            sky_row, total = _clip_sky(sky_row, scene_sum, total, discrete, zero_when_full=True)
            sky_vf[name] = sky_row
            combined = scene_sum + total
        residual = 1.0 - combined
        if abs(residual) <= threshold:
            residual = 0.0
        rest_vf[name] = {"Rest": residual}
    return vf_scene, sky_vf, rest_vf


__all__ = ["view_factor_outside_workflow"]
