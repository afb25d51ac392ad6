"""PS-side optimizers (parity: ``/root/reference/src/optim/__init__.py``)."""
from .sgd import SGD
from .adam import Adam

__all__ = ["SGD", "Adam"]
