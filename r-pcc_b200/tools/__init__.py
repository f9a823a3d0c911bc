"""Command-line drivers mirroring the reference's tools/*.py (same flags, same printed metrics)."""
