from .sgd import SGD

__all__ = ["SGD"]
