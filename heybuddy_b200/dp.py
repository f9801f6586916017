"""
Data-parallel classifier training -- the one place on this path with a real exchange step.

The reference trains on one device (``trainer.py:405-462``): forward, high-loss selection, weighted BCE **averaged over the
selected rows**, backward, Adam.  Sharding the batch over ranks keeps that arithmetic if the average is taken over the
rows selected on ALL ranks, so a step is

    select (local)  ->  all-reduce SUM [n_selected]            4 bytes
    backward (local, loss and gradients divided by the global n_selected)
    all-reduce SUM [gradients | loss part]                       ~1 MB (256,417 parameters + 1)
    Adam (identical on every rank: replicas never diverge)

Two NCCL collectives per step, both latency-bound at this size (NVLink is otherwise idle on this path).  The engine is any
object with ``dp_select / dp_backward / dp_grads / dp_adam`` (``WakeWordMLPModel`` on the GPU; the CPU test drives the same
function with the oracle classifier over gloo).

Loss and gradients are linear in 1 / n_selected, so an engine that also has ``dp_local_step / dp_apply`` (the GPU model outside
its parity modes) does the step with ONE collective:

    local step: forward, select, UNNORMALISED loss sum and gradients  ->  [gradients | loss sum | n_selected]   one C call
    all-reduce SUM of that buffer                                                                                   ~1 MB
    apply: divide by the global n_selected, Adam                                                                  one C call
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

__all__ = ["distributed_train_step", "shard_batch"]


def shard_batch(x, y, rank: int, world: int):
    """Rank r's rows of a global batch: a contiguous block (the last ranks get one row less when it does not divide)."""
    n = x.shape[0]
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return x[lo:hi], y[lo:hi]


def _one_collective(engine: Any) -> bool:
    import os

    staged = os.environ.get("HB_MLP_STAGED", "")[:1] == "1" or os.environ.get("HB_MLP_FMA", "")[:1] == "1"
    return hasattr(engine, "dp_local_step") and hasattr(engine, "dp_apply") and not staged


def distributed_train_step(engine: Any, x, y, lr: float, negative_weight: float = 1.0, high_loss_threshold: float = 1e-4,
                           min_selected: int = 128, group: Optional[Any] = None, grad_buffer: Optional[Any] = None,
                           one_collective: Optional[bool] = None) -> Tuple[Any, Any]:
    """
    One training step on this rank's shard ``(x, y)`` of the global batch.  Returns ``(prob, stats)`` like
    ``WakeWordMLPModel.train_step`` with GLOBAL statistics: stats = [mean loss over all selected rows, rows selected on all
    ranks, stepped (0/1), this rank's high-loss rate].  With an uninitialised / single-rank process group it reduces to the
    fused single-device step's arithmetic.
    """
    import torch
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if one_collective is None:
        one_collective = multi and _one_collective(engine)
    if one_collective:
        prob, stats, exchange = engine.dp_local_step(x, y, negative_weight, high_loss_threshold)
        if multi:
            dist.all_reduce(exchange, op=dist.ReduceOp.SUM, group=group)
        engine.dp_apply(exchange, lr, min_selected, stats)
        return prob, stats
    prob, stats = engine.dp_select(x, y, high_loss_threshold)
    n_total = stats[1:2].clone()
    if multi:
        dist.all_reduce(n_total, op=dist.ReduceOp.SUM, group=group)
    stats = engine.dp_backward(n_total, negative_weight, high_loss_threshold, min_selected)
    if multi:
        buf = engine.dp_grads(grad_buffer)                    # [n_params] (+ 1 slot for the loss when the caller provides it)
        packed = torch.cat([buf, stats[0:1]])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        engine.dp_grads(packed[:-1].contiguous(), to_model=True)
        stats[0] = packed[-1]
    stats[1] = n_total[0]
    engine.dp_adam(lr, stats)
    return prob, stats
