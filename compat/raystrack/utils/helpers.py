"""raystrack.utils.helpers of the reference (utils/helpers.py): grid sizing and the reciprocity enforcers."""
import raystrack  # noqa: F401
from raystrack_b200.prepared import grid_from_density  # noqa: F401
from raystrack_b200.reciprocity import enforce_reciprocity_and_rowsum, enforce_reciprocity_only  # noqa: F401


def hold_console_open(prompt: str = "Press Enter to close...") -> None:
    """The reference keeps a spawned log console open (utils/helpers.py:260-275); there is no console here."""
    return None
