from .pfgst_loss import PFGSTLoss, LOSS_KEYS

__all__ = ["PFGSTLoss", "LOSS_KEYS"]
