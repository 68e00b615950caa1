from .metrics import (ConfusionMeter, confusion_matrix, eval_metrics, intersect_and_union,
                      intersect_and_union_batch, mean_dice, mean_fscore, mean_iou, pre_eval_logits,
                      pre_eval_to_metrics, seg_argmax, shard_range, total_area_to_metrics,
                      total_intersect_and_union)

from .inference import (aug_test, inference, inference_logits, make_test_cfg, simple_test, slide_inference, slide_logits,
                        window_grid)

__all__ = ["aug_test", "inference", "inference_logits", "make_test_cfg", "simple_test", "slide_inference", "slide_logits", "window_grid",
           "ConfusionMeter", "confusion_matrix", "eval_metrics", "intersect_and_union",
           "intersect_and_union_batch", "mean_dice", "mean_fscore", "mean_iou", "pre_eval_logits",
           "pre_eval_to_metrics", "seg_argmax", "shard_range",
           "total_area_to_metrics", "total_intersect_and_union"]
