"""Import shim: ``PYTHONPATH=<repo>/compat`` makes ``import raystrack`` resolve to the B200 implementation, so scripts
written against philip-ba/raystrack (same module layout: raystrack, raystrack.main, .params, .io, .api,
.utils.prepared, .utils.helpers) run unchanged.  Everything is re-exported from ``raystrack_b200``."""
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[2]
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

from raystrack_b200 import *  # noqa: E402,F401,F403
from raystrack_b200 import __all__, __version__  # noqa: E402,F401
