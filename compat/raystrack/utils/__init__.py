"""raystrack.utils of the reference: only the names that belong to the hot path are provided."""
