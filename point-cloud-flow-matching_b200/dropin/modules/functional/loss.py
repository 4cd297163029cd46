import _bootstrap  # noqa: F401
from pcfm.functional import huber_loss, kl_loss

__all__ = ["kl_loss", "huber_loss"]
