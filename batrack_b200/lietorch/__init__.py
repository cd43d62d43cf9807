from .groups import SE3  # noqa: F401

__all__ = ["SE3"]
