"""Log variables without per-variable host syncs.

`BaseSegmentor._parse_losses` (rsiseg/models/segmentors/base.py:177-222) calls `.item()` on
every log variable — ~17 device->host syncs and, when distributed, ~17 scalar all-reduces per
iteration. Here every `_parse_losses` call is ONE single-thread kernel (`pfst_gather_scalars`)
that copies the 0-dim device scalars into a row of a persistent device ledger and forms the
left-to-right sum of the loss entries; the row is all-reduced ONCE per iteration
(`LogLedger.end`) and the ledger is copied to the host only when somebody reads a value
(`LazyScalar.__float__`, i.e. once per log interval with mmcv's logger hooks). The values,
keys and ordering are the reference's.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib

MAX_SCALARS = 32


class LazyScalar:
    """A log value that still lives on the device. Behaves like the Python float the reference
    returns (`float()`, arithmetic, comparisons, numpy conversion, formatting, pickling);
    the first use copies the outstanding ledger rows to the host in one transfer."""
    __slots__ = ("_ledger", "_seq", "_idx", "_val")

    def __init__(self, ledger, seq: int, idx: int):
        self._ledger, self._seq, self._idx, self._val = ledger, seq, idx, None

    def _get(self) -> float:
        if self._val is None:
            self._val = self._ledger.value(self._seq, self._idx)
            self._ledger = None
        return self._val

    def item(self) -> float:
        return self._get()

    __float__ = _get

    def __int__(self):
        return int(self._get())

    def __bool__(self):
        return bool(self._get())

    def __repr__(self):
        return repr(self._get())

    def __format__(self, spec):
        return format(self._get(), spec)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._get(), dtype=dtype or np.float64)

    def __reduce__(self):
        return (float, (self._get(),))

    def __hash__(self):
        return hash(self._get())

    def __eq__(self, o):
        return self._get() == o

    def __ne__(self, o):
        return self._get() != o

    def __lt__(self, o):
        return self._get() < o

    def __le__(self, o):
        return self._get() <= o

    def __gt__(self, o):
        return self._get() > o

    def __ge__(self, o):
        return self._get() >= o

    def __neg__(self):
        return -self._get()

    def __abs__(self):
        return abs(self._get())

    def __add__(self, o):
        return self._get() + o

    __radd__ = __add__

    def __sub__(self, o):
        return self._get() - o

    def __rsub__(self, o):
        return o - self._get()

    def __mul__(self, o):
        return self._get() * o

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._get() / o

    def __rtruediv__(self, o):
        return o / self._get()

    def __pow__(self, o):
        return self._get() ** o


class _SumFn(torch.autograd.Function):
    """total = sum_i w_i * v_i over the masked 0-dim inputs, formed left to right in fp32 by
    the gather kernel (the same chain of adds as Python's sum()); d total / d v_i = w_i."""

    @staticmethod
    def forward(ctx, ledger, mask, weights, record, *vals):
        total = torch.empty((), dtype=torch.float32, device=vals[0].device)
        ledger._gather(vals, mask, weights, total, record)
        ctx.mask, ctx.weights = mask, weights
        return total

    @staticmethod
    def backward(ctx, g):
        out = []
        for i in range(len(ctx.needs_input_grad) - 4):
            if not (ctx.mask >> i) & 1 or not ctx.needs_input_grad[4 + i]:
                out.append(None)
            else:
                w = 1.0 if ctx.weights is None else ctx.weights[i]
                out.append(g if w == 1.0 else g * w)
        return (None, None, None, None, *out)


class LogLedger:
    ROWS = 128            # iterations kept on the device between host reads (log interval: 50)
    KEEP = 4096           # materialised host rows kept for late readers

    def __init__(self, device):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.rows = torch.zeros((self.ROWS, MAX_SCALARS * 2), dtype=torch.float32, device=self.device)
        self.host = torch.zeros_like(self.rows, device="cpu")
        if self.cuda:
            self.host = self.host.pin_memory()
        self.seq, self.fill, self.open = 0, 0, False     # current row (iteration), entries used, begin() active
        self.first_unread = 0                            # rows [first_unread, seq] are not on the host yet
        self._vals: "OrderedDict[int, np.ndarray]" = OrderedDict()
        self._checked = set()                            # key tuples whose cross-rank length check has run
        self.d2h_bytes = 0
        self.world = 1

    # ------------------------------------------------------------------ iteration protocol
    def begin(self) -> None:
        """Opens the row of a new iteration (forward_train calls it once)."""
        if self.seq + 1 - self.first_unread >= self.ROWS:
            self.flush()
        self.seq += 1
        self.fill, self.open = 0, True
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def end(self) -> None:
        """Closes the row: ONE all-reduce of the iteration's log variables (each was divided by the
        world size when it was written, like base.py:217-218)."""
        if self.open and self.world > 1 and self.fill:
            dist.all_reduce(self.rows[self.seq % self.ROWS, :self.fill])
        self.open = False

    # ------------------------------------------------------------------------- parse_losses
    def parse(self, losses):
        """BaseSegmentor._parse_losses: -> (loss tensor for backward, OrderedDict of log values)."""
        names, vals = [], []
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                v = value if value.dim() == 0 and value.dtype == torch.float32 else value.mean()
            elif isinstance(value, list):
                v = sum(_l.mean() for _l in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
            if v.dtype != torch.float32:
                v = v.float()
            names.append(name)
            vals.append(v)
        if len(vals) + 1 > MAX_SCALARS:
            raise ValueError(f"_parse_losses: more than {MAX_SCALARS - 1} log variables in one call")
        standalone = not self.open
        if standalone:
            self.begin()
        if self.fill + len(vals) + 1 > self.rows.shape[1]:
            raise ValueError("_parse_losses: too many log variables in one iteration")
        self._length_check(names, vals[0].device if vals else self.device)
        mask = 0
        for i, n in enumerate(names):
            if 'loss' in n:
                mask |= 1 << i
        base = self.fill
        if vals:
            loss = _SumFn.apply(self, mask, None, True, *vals)
        else:
            loss = torch.zeros((), dtype=torch.float32, device=self.device)
            self._gather((), 0, None, loss, True)
        log_vars = OrderedDict((n, LazyScalar(self, self.seq, base + i)) for i, n in enumerate(names))
        log_vars['loss'] = LazyScalar(self, self.seq, base + len(names))
        if standalone:
            self.end()
        return loss, log_vars

    def weighted_total(self, parts, weights):
        """total_loss = 0 + p0*w0 + p1*w1 + ... (pfgst.py:237,310,342) in one launch; not logged."""
        return _SumFn.apply(self, (1 << len(parts)) - 1, tuple(float(w) for w in weights), False, *parts)

    def _length_check(self, names, device) -> None:
        """base.py:204-212 — 'to prevent GPUs from infinite waiting': every rank must log the same
        number of variables. Checked once per distinct key set (it costs a host sync)."""
        if self.world <= 1:
            return
        key = tuple(names)
        if key in self._checked:
            return
        n = torch.tensor(len(names), device=device)
        dist.all_reduce(n)
        assert int(n) == len(names) * self.world, \
            'loss log variables are different across GPUs!\n' + \
            f'rank {dist.get_rank()} len(log_vars): {len(names)} keys: ' + ','.join(names)
        self._checked.add(key)

    def _gather(self, vals, mask, weights, total, record) -> None:
        n = len(vals)
        row = self.rows[self.seq % self.ROWS] if record else None
        if not self.cuda:
            # host tensors (the gloo tests of the multi-rank logic): the same left-to-right chain
            t = torch.zeros((), dtype=torch.float32)
            for i, v in enumerate(vals):
                if record:
                    row[self.fill + i] = v.detach() / self.world
                if (mask >> i) & 1:
                    t = t + (v.detach() if weights is None else v.detach() * weights[i])
            total.copy_(t)
            if record:
                row[self.fill + n] = t / self.world
                self.fill += n + 1
            return
        ptrs = (C.c_void_p * max(n, 1))(*[v.data_ptr() for v in vals])
        for v in vals:
            if not v.is_cuda or v.dtype != torch.float32 or v.numel() != 1:
                raise TypeError("_parse_losses: log variables must be fp32 CUDA scalars")
        w = None if weights is None else (C.c_float * n)(*weights)
        _lib.call("pfst_gather_scalars", ptrs, w, n, mask, float(self.world),
                  None if row is None else row.data_ptr() + 4 * self.fill, total.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        if record:
            self.fill += n + 1

    # ----------------------------------------------------------------------------- host reads
    def flush(self) -> None:
        """One device->host copy of the ledger rows that have not been read yet."""
        lo, hi = self.first_unread, self.seq
        if hi < lo:
            return
        if self.cuda:
            self.host.copy_(self.rows, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.d2h_bytes += self.rows.numel() * 4
        else:
            self.host.copy_(self.rows)
        h = self.host.numpy()
        for s in range(max(lo, hi - self.ROWS + 1), hi + 1):
            self._vals[s] = h[s % self.ROWS].copy()
        while len(self._vals) > self.KEEP:
            self._vals.popitem(last=False)
        self.first_unread = hi + 1

    def value(self, seq: int, idx: int) -> float:
        row = self._vals.get(seq)
        if row is None or (seq == self.seq and self.open):
            self.flush()
            if seq == self.seq and self.open:
                self.first_unread = seq                 # the open row may still grow: read it again later
            row = self._vals.get(seq)
            if row is None:
                raise RuntimeError("log value is older than the ledger keeps (read it within "
                                   f"{self.ROWS} iterations)")
        return float(row[idx])


class StepRecord:
    """The _parse_losses calls of ONE forward_train iteration, gathered by a single launch
    (`pfst_gather_segments`): `add()` only registers pointers and returns the lazy log values;
    `launch()` — issued by engine.StepTotalFn once every scalar has been produced — writes the
    ledger row and the iteration's total loss."""

    def __init__(self, ledger: "LogLedger"):
        self.ledger = ledger
        self.tensors, self.flags, self.weights, self.seg_of = [], [], [], []
        self.n_seg = 0
        self.seg_weights = []
        self.aux_weights = (1.0, 1.0)      # segment weights of the PFGSTLoss terms / the prototype distance
        self.base = ledger.fill

    def add(self, losses, weight: float = 1.0):
        led = self.ledger
        names = []
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                v = value if value.dim() == 0 and value.dtype == torch.float32 else value.mean()
            elif isinstance(value, list):
                v = sum(_l.mean() for _l in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
            if v.dtype != torch.float32:
                v = v.float()
            if not v.is_cuda or v.numel() != 1:
                raise TypeError("_parse_losses: log variables must be CUDA scalars on the fused path")
            names.append(name)
            self.tensors.append(v)
            self.flags.append(1 if 'loss' in name else 0)
            self.weights.append(0.0)
            self.seg_of.append(self.n_seg)
        if not names:
            raise ValueError("_parse_losses: empty loss dict")
        if len(self.tensors) > MAX_SCALARS or led.fill + len(self.tensors) + self.n_seg + 1 > led.rows.shape[1]:
            raise ValueError("_parse_losses: too many log variables in one iteration")
        led._length_check(names, self.tensors[-1].device)
        self.flags[-1] |= 2
        self.weights[-1] = float(weight)
        self.seg_weights.append(float(weight))
        first = self.base + len(self.tensors) - len(names) + self.n_seg
        self.n_seg += 1
        log_vars = OrderedDict((n, LazyScalar(led, led.seq, first + i)) for i, n in enumerate(names))
        log_vars['loss'] = LazyScalar(led, led.seq, first + len(names))
        return log_vars

    def segment_weight(self, i: int) -> float:
        return self.seg_weights[self.seg_of[i]] if self.flags[i] & 1 else 0.0

    def launch(self, total: torch.Tensor) -> None:
        led, n = self.ledger, len(self.tensors)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in self.tensors])
        w = (C.c_float * n)(*self.weights)
        fl = (C.c_uint8 * n)(*self.flags)
        row = led.rows[led.seq % led.ROWS]
        _lib.call("pfst_gather_segments", ptrs, w, fl, n, float(led.world), row.data_ptr() + 4 * self.base,
                  total.data_ptr(), torch.cuda.current_stream().cuda_stream)
        led.fill = self.base + n + self.n_seg


_ledgers: dict = {}


def ledger_for(device) -> LogLedger:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    led = _ledgers.get(device)
    if led is None:
        led = _ledgers[device] = LogLedger(device)
    return led
