from .mlp import MLPHead

__all__ = ["MLPHead"]
