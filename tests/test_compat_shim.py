"""``PYTHONPATH=compat`` turns ``import raystrack`` into the B200 implementation with the reference's module layout."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_import_raystrack_resolves_to_b200_package():
    code = (
        "import raystrack, raystrack.main, raystrack.params, raystrack.io, raystrack.api\n"
        "from raystrack import view_factor_matrix, view_factor, view_factor_to_tregenza_sky, view_factor_outside_workflow\n"
        "from raystrack import MatrixParams, SkyParams, PreparedSolver, save_vf_matrix_json, load_meshes_json\n"
        "from raystrack.params import MatrixParams as MP\n"
        "from raystrack.io import load_meshes_json, save_vf_matrix_json\n"
        "from raystrack.utils.prepared import PreparedSolver as PS\n"
        "from raystrack.utils.helpers import hold_console_open, grid_from_density, enforce_reciprocity_only\n"
        "import raystrack_b200, raystrack_b200.main\n"
        "assert raystrack.main is raystrack_b200.main and MP is raystrack_b200.MatrixParams and PS is raystrack_b200.PreparedSolver\n"
        "logs = []\n"
        "raystrack.main._log = logs.append\n"
        "assert raystrack_b200.main._log is logs.append or raystrack_b200.main._log == logs.append\n"
        "assert grid_from_density(40.0, 16) == 26 and hold_console_open() is None\n"
        "print('shim ok')\n"
    )
    env = {"PYTHONPATH": str(ROOT / "compat"), "PATH": "/usr/bin:/bin"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "shim ok" in r.stdout, r.stdout + r.stderr
