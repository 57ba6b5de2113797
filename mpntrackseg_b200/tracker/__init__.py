from .mpn_tracker import MPNTracker  # noqa: F401
