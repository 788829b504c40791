"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "raystrack_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("rsk_ctx_create", "rsk_scene_create", "rsk_emitters_create", "rsk_trace_rays", "rsk_matrix_begin",
                 "rsk_matrix_step", "rsk_matrix_read", "rsk_sky_begin", "rsk_sky_step", "rsk_sky_read",
                 "rsk_reciprocity_rowsum", "rsk_solve_enqueue_trace", "rsk_solve_enqueue_fold"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from raystrack_b200 import _native
    lib = _native.load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_native.EXPORTS) == set(declared_symbols())
    assert lib.rsk_abi_version() == 2


def test_sm100a_code_is_embedded():
    """The shared object carries sm_100a SASS (cuobjdump lists the ELF images)."""
    import shutil
    import subprocess
    from raystrack_b200 import _native
    _native.load()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-lelf", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_device_fails_loudly():
    """Without a GPU the product refuses to run (no CPU fallback, no oracle on the product path)."""
    import numpy as np
    import raystrack_b200 as rb
    from raystrack_b200 import _native, synthetic
    if _native.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError):
        rb.view_factor_matrix(synthetic.parallel_unit_squares(), rb.MatrixParams())
    with pytest.raises(_native.NativeError):
        _native.Context(0)
    # argument errors still surface before any device work
    with pytest.raises(TypeError):
        rb.view_factor_matrix(synthetic.parallel_unit_squares(), rb.SkyParams())
    with pytest.raises(ValueError):
        rb.view_factor_matrix(synthetic.parallel_unit_squares(), rb.MatrixParams(device="tpu"))


def test_product_never_imports_the_oracle():
    for path in (ROOT / "raystrack_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in text.replace("# oracle", ""), path
    for path in (ROOT / "raystrack_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(), path
