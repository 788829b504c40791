"""raystrack_b200 -- B200-native (sm_100a CUDA) implementation of Raystrack's Monte-Carlo view-factor path.

Drop-in for the reference's public API (reference src/raystrack/__init__.py:1-30)::

    from raystrack_b200 import view_factor_matrix, view_factor_to_tregenza_sky, MatrixParams, SkyParams
"""
from .api import view_factor_outside_workflow
from .io import load_meshes_json, load_vf_matrix_json, merge_vf_matrix, save_meshes_json, save_vf_matrix_json
from .main import view_factor, view_factor_matrix, view_factor_to_tregenza_sky
from .params import MatrixParams, SkyParams
from .prepared import PreparedSolver

__all__ = [
    "view_factor_matrix", "view_factor", "view_factor_to_tregenza_sky", "view_factor_outside_workflow", "MatrixParams", "SkyParams",
    "PreparedSolver", "save_vf_matrix_json", "load_vf_matrix_json", "save_meshes_json", "load_meshes_json",
    "merge_vf_matrix",
]
__version__ = "0.1.0"
