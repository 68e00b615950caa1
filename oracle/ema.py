"""EMA mean-teacher update — restates rsiseg/models/uda/pfgst.py:105-127."""
from __future__ import annotations

import torch


def alpha_teacher(it: int, alpha: float) -> float:
    # pfgst.py:117 — python doubles
    return min(1 - 1 / (it + 1), alpha)


def ema_init(teacher: list[torch.Tensor], student: list[torch.Tensor]) -> None:
    # pfgst.py:105-114 — teacher <- clone(student), tensor by tensor
    for t, s in zip(teacher, student):
        if not s.shape:
            t.copy_(s.clone())
        else:
            t[:] = s[:].clone()


def ema_update(teacher: list[torch.Tensor], student: list[torch.Tensor], it: int, alpha: float) -> None:
    # pfgst.py:116-127 — three separately rounded fp32 ATen ops per tensor
    a = alpha_teacher(it, alpha)
    for t, s in zip(teacher, student):
        if not s.shape:
            t.copy_(a * t + (1 - a) * s)
        else:
            t[:] = a * t[:] + (1 - a) * s[:]
