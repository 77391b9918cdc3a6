from .registry import DATA_GENERATOR, DATASET, HOOKS, LAYER, Registry

__all__ = ["DATA_GENERATOR", "DATASET", "HOOKS", "LAYER", "Registry"]
