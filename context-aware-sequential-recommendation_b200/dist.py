"""Data-parallel training over the GPUs of one box: one process per GPU, `torch.distributed` (NCCL over NVLink 5 /
NVSwitch) for the single exchange the path has — the gradient all-reduce.

The reference is single-device; the semantics kept here are those of its loss: the gradient is the
derivative of  sum(loss terms) / sum(istarget)  over the WHOLE batch (models/sasrec.py:105-108).  Each rank therefore
contributes un-normalised gradient numerators plus its local sum(istarget); one flat `all_reduce(SUM)` over
[gradients | loss_sum, auc_sum, count] makes every rank hold the global numerators and the global denominator, and
the Adam kernel divides.  N ranks with B/N sequences each reproduce the single-GPU step on B sequences up to fp32
summation order.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun / torch.distributed.run environment -> process group.  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def attach(engine, group=None):
    """Make `engine.launch_train_step` data parallel over `group` (default: the world group)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return engine
    engine.world_size = dist.get_world_size(group)
    engine.rank = dist.get_rank(group)
    # identical replicas: rank 0's parameters and optimizer state win
    for t in (engine.w, engine.m, engine.v, engine.adam_state):
        dist.broadcast(t, src=0, group=group)
    # independent dropout streams per rank (one global batch, different positions)
    engine.seed = (engine.seed + 0x9E3779B1 * engine.rank) & 0xFFFFFFFFFFFF

    def allreduce(c):
        dist.all_reduce(engine.gbuf, op=dist.ReduceOp.SUM, group=group)

    engine.grad_allreduce = allreduce
    return engine


def shard_users(n_users: int, rank: int, world: int):
    """Contiguous user shard for evaluation (users are independent)."""
    per = (n_users + world - 1) // world
    lo = min(n_users, rank * per)
    return lo, min(n_users, lo + per)


def reduce_rank_histogram(hist: torch.Tensor, group=None):
    """Evaluation's only exchange: int64 histogram of ranks 0..9 + valid-user count (integers => exact HR@10)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
