"""A tiny segmentor with the EncoderDecoder API the UDA trainer calls
(rsiseg/models/segmentors/encoder_decoder.py:72-84, 166-217): stride-8 'decoded
features', stride-4 logits, pixel-weighted CE (decode_head.py:253-279,
cross_entropy_loss.py:45-63). Used by both the reference PFGST (loaded by path) and
the B200 drop-in so that the whole step can be compared."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class TinySegmentor(nn.Module):
    def __init__(self, num_classes=6, dim=16, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.num_classes = num_classes
        self.train_cfg, self.test_cfg = {}, {}
        self.backbone = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.ReLU(), nn.Conv2d(8, dim, 3, 2, 1), nn.ReLU(),
                                      nn.Conv2d(dim, dim, 3, 2, 1), nn.ReLU())
        self.drop = nn.Dropout2d(0.0)   # deterministic: CPU and CUDA RNG streams differ
        self.head = nn.Conv2d(dim, num_classes, 1)
        self.scalar = nn.Parameter(torch.tensor(1.0))      # 0-dim parameter (pfgst.py:111-112)
        with torch.no_grad():
            for p in self.parameters():
                if p.dim() > 0:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.3)

    def extract_feat(self, img):
        return self.backbone(img)

    def _decode(self, x):
        logits = F.interpolate(self.head(self.drop(x)) * self.scalar, scale_factor=2, mode='bilinear',
                               align_corners=False)
        return logits

    def encode_decode(self, img, img_metas):
        x = self.extract_feat(img)
        logits = self._decode(x)
        out = F.interpolate(logits, size=img.shape[2:], mode='bilinear', align_corners=False)
        return out, {'feats': x, 'decoded_features': x, 'seg_logits': out}

    def forward_train(self, img, img_metas, gt_semantic_seg, seg_weight=None, return_feats=False,
                      return_decoded_feats=False, return_logits=False, return_states=False):
        x = self.extract_feat(img)
        logits = self._decode(x)
        up = F.interpolate(logits, size=gt_semantic_seg.shape[2:], mode='bilinear', align_corners=False)
        loss = F.cross_entropy(up, gt_semantic_seg.squeeze(1), reduction='none', ignore_index=255)
        if seg_weight is not None:
            loss = loss * seg_weight
        losses = {'decode.loss_ce': loss.mean(),
                  'decode.acc_seg': (up.argmax(1) == gt_semantic_seg.squeeze(1)).float().mean() * 100}
        if return_feats:
            losses['features'] = x
        if return_logits:
            losses['logits'] = logits
        if return_decoded_feats:
            losses['decoded_features'] = x
        return losses
