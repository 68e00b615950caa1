"""World-size-2 `gloo` tests (CPU) of the host-side multi-rank logic: the packed prototype
all-reduce buffer, the confusion-matrix all-reduce, the batched `_parse_losses` reduction and
the contiguous sharding of an evaluation sweep. The arithmetic each rank feeds in comes from
the CPU oracle (no kernel runs here); what is tested is that the reduced quantities equal
the single-process result over the union of the shards."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORLD = 2
C, D = 6, 16


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_inputs(rank: int):
    g = torch.Generator().manual_seed(1234 + rank)          # per-rank seeds, as bench.py
    feats = torch.relu(torch.randn((2, D, 8, 8), generator=g))
    labels = torch.randint(0, C + 1, (2, 8, 8), generator=g)
    labels[labels == C] = 255
    pred = torch.randint(0, C, (3, 32, 32), generator=g)
    gt = torch.randint(0, C, (3, 32, 32), generator=g)
    gt[:, :2] = 255
    return feats, labels, pred, gt


def _worker(rank: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from oracle import metrics as OM, prototypes as OP
        from pfst_b200.evaluation.metrics import ConfusionMeter, shard_range
        from pfst_b200.prototypes import PrototypeBank
        from pfst_b200.uda.uda_decorator import UDADecorator

        feats, labels, pred, gt = _rank_inputs(rank)
        # --- prototypes: each rank's kernel would write sums|counts straight into `packed`
        bank = PrototypeBank(C, D, torch.device("cpu"))
        sums, counts = OP.proto_accumulate(feats, labels, C)
        bank.packed[:C * D] = sums.float().reshape(-1)
        bank.packed[C * D:] = counts.float()
        work = bank.all_reduce()
        assert work is not None
        work.wait()
        # --- confusion matrix
        meter = ConfusionMeter(C, device=torch.device("cpu"))
        meter.conf[0, :C, :C] = torch.as_tensor(OM.confusion(pred.numpy(), gt.numpy(), C, 255))
        meter.all_reduce()
        # --- log-var reduction (one vector, one all-reduce)
        losses = {"decode.loss_ce": torch.tensor(1.0 + rank), "decode.acc_seg": torch.tensor(10.0 * (rank + 1)),
                  "loss_aux": [torch.tensor(0.5), torch.tensor(0.25 * rank)]}
        loss, log_vars = UDADecorator._parse_losses(losses)
        # --- sharding of an evaluation sweep
        lo, hi = shard_range(10, rank, WORLD)
        torch.save(dict(packed=bank.packed.clone(), conf=meter.matrix().clone(), log_vars=dict(log_vars),
                        loss=float(loss), shard=(lo, hi)), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_reductions_match_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    from oracle import metrics as OM, prototypes as OP
    outs = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(WORLD)]
    ins = [_rank_inputs(r) for r in range(WORLD)]
    # prototypes: all-reduced packed buffer == accumulate over the concatenated batch
    sums, counts = OP.proto_accumulate(torch.cat([i[0] for i in ins]), torch.cat([i[1] for i in ins]), C)
    for o in outs:
        assert torch.equal(o["packed"][C * D:].long(), counts)
        assert (o["packed"][:C * D].double().view(C, D) - sums).abs().max() <= 1e-5 * sums.abs().max()
        assert torch.equal(o["packed"], outs[0]["packed"])                 # identical on every rank
    # confusion matrix: integer, exact
    want = OM.confusion(np.concatenate([i[2].numpy() for i in ins]),
                               np.concatenate([i[3].numpy() for i in ins]), C, 255)
    for o in outs:
        assert torch.equal(o["conf"], torch.as_tensor(want))
    # log vars: mean over ranks, 'loss' = sum of the keys containing 'loss' (base.py:200-220)
    lv = outs[0]["log_vars"]
    assert list(lv) == ["decode.loss_ce", "decode.acc_seg", "loss_aux", "loss"]
    assert abs(lv["decode.loss_ce"] - 1.5) < 1e-6 and abs(lv["decode.acc_seg"] - 15.0) < 1e-6
    assert abs(lv["loss_aux"] - (0.5 + 0.125)) < 1e-6 and abs(lv["loss"] - (1.5 + 0.625)) < 1e-6
    assert outs[1]["log_vars"] == lv
    assert abs(outs[0]["loss"] - 1.5) < 1e-6 and abs(outs[1]["loss"] - 2.75) < 1e-6   # local loss drives backward
    # contiguous shards cover the sweep exactly once
    assert outs[0]["shard"] == (0, 5) and outs[1]["shard"] == (5, 10)


def test_shard_range_is_a_partition():
    from pfst_b200.evaluation.metrics import shard_range
    for n in (0, 1, 7, 10, 10000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
