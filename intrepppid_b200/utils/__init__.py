from .weightdrop import WeightDrop
from .embedding_do import embedding_dropout

__all__ = ["WeightDrop", "embedding_dropout"]
