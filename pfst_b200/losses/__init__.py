from .cross_entropy import decode_head_losses, upsample_cross_entropy
from .pfgst_loss import PFGSTLoss, LOSS_KEYS

__all__ = ["PFGSTLoss", "LOSS_KEYS", "decode_head_losses", "upsample_cross_entropy"]
