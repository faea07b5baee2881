"""Multi-GPU plumbing: one process per GPU, clips sharded by index, no collective in the forward path.

The path shards naturally by clip: every stage is per-clip, weights are replicated
(ViT 582 MB bf16 + projector 8-10 MB).  The only exchange step is in the training-shaped variant: one all-reduce
(mean) of ``mm_projector.{weight,bias}.grad`` -- 4.2 M (7B) / 5.2 M (13B) elements -- over NCCL / NVLink.
``embed_tokens`` gradients belong to the LLM and are not reduced here (ZeRO-3 handles them in the reference,
scripts/zero3.json).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: Optional[str] = None) -> None:
    """Reads RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment (torchrun contract)."""
    if dist.is_initialized():
        return
    rank, world, local = env_rank_world()
    if world == 1:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend, rank=rank, world_size=world)


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Round-robin clip assignment: clip i -> rank i % world (EPIC-eval sweep)."""
    return list(range(rank, n_clips, world))


def allreduce_projector_grads(projector: torch.nn.Module, group=None, async_op: bool = False):
    """Mean all-reduce of the projector gradients as ONE flat fp32 bucket (a single latency-bound collective
    instead of two).  Returns the work handle when async_op=True (call ``finish`` on it)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    grads = [p.grad for p in (projector.weight, projector.bias) if p.grad is not None]
    if not grads:
        return None
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if async_op:
            work.wait()
        flat.div_(dist.get_world_size(group))
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off: off + n].reshape(g.shape).to(g.dtype))
            off += n

    if async_op:
        return finish
    finish()
    return None


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: every multi-GPU number is the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
