from .metrics import (ConfusionMeter, confusion_matrix, eval_metrics, intersect_and_union,
                      intersect_and_union_batch, mean_dice, mean_fscore, mean_iou, pre_eval_logits,
                      pre_eval_to_metrics, seg_argmax, shard_range, total_area_to_metrics,
                      total_intersect_and_union)

__all__ = ["ConfusionMeter", "confusion_matrix", "eval_metrics", "intersect_and_union",
           "intersect_and_union_batch", "mean_dice", "mean_fscore", "mean_iou", "pre_eval_logits",
           "pre_eval_to_metrics", "seg_argmax", "shard_range",
           "total_area_to_metrics", "total_intersect_and_union"]
