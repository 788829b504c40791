"""Alias of raystrack_b200.prepared."""
import sys

import raystrack  # noqa: F401
import raystrack_b200.prepared as _impl

sys.modules[__name__] = _impl
