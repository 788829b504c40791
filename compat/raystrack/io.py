"""Alias of raystrack_b200.io (see compat/raystrack/__init__.py)."""
import sys

import raystrack  # noqa: F401  (puts the repository on sys.path)
import raystrack_b200.io as _impl

sys.modules[__name__] = _impl
