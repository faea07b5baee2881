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


def _flat_sizes(projector):
    nW = projector.weight.numel()
    nb = projector.bias.numel() if getattr(projector, "bias", None) is not None else 0
    return nW, nb


def allreduce_projector_grads(projector: torch.nn.Module, group=None, async_op: bool = False):
    """Mean all-reduce of the projector gradients as ONE flat fp32 bucket of FIXED size ``weight.numel() + bias.numel()``
    (a single latency-bound collective instead of two).  Every rank of the group issues the same collective every time:
    a missing gradient (text-only micro-batch) contributes zeros and is created by the reduction, so ranks can never
    disagree on the size of, or skip, the call.  Returns a ``finish`` callable when async_op=True.
    Simple synchronous entry; the training loop uses :class:`ProjectorGradReducer` (preallocated bucket, overlap)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    params = [projector.weight] + ([projector.bias] if getattr(projector, "bias", None) is not None else [])
    nW, nb = _flat_sizes(projector)
    dev = projector.weight.device
    flat = torch.zeros(nW + nb, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off: off + p.numel()].copy_(p.grad.reshape(-1))
        off += p.numel()
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if async_op:
            work.wait()
        flat.div_(dist.get_world_size(group))
        off = 0
        for p in params:
            n = p.numel()
            g = flat[off: off + n].reshape(p.shape)
            if p.grad is None:
                p.grad = g.to(p.dtype)
            else:
                p.grad.copy_(g)
            off += n

    if async_op:
        return finish
    finish()
    return None


_own_group = None


def _dedicated_nccl_group():
    """The collective runs on a side stream WHILE the next step's ViT forward runs on the main stream, and that forward is
    a chain of persistent kernels that own every SM (one 225 KB CTA per SM): an NCCL kernel that is resident and spinning
    on slower ranks takes SMs away from them.  The reducer therefore uses its own communicator, created with
    ncclConfig ctaPolicy = EFFICIENCY on a high-priority stream.  Measured on one 8-GPU B200 box, training-shaped variant,
    same box back to back (profiles/r2_train_nccl_sweep.txt; 8-GPU step / collective alone / collective inside the step):
        ctaPolicy EFFICIENCY      60.08 ms  0.086 ms  3.7 ms   (0.993 of 8x the 1-GPU rate)   <- default
        default config            60.13 ms  0.085 ms  5.5 ms   (0.992)
        max 8 / 2 / 1 CTAs        60.26 / 60.56 / 60.88 ms; 0.22 / 0.75 / 1.47 ms alone       (0.990 / 0.985 / 0.980)
    i.e. capping the CTAs only makes the collective slower (16.8 MB is latency-bound at any width) and holds its SMs longer.
    HVLM_NCCL_CTA_POLICY (default 1), HVLM_NCCL_MAX_CTAS (default 0 = unset), HVLM_NCCL_HIGH_PRIO (default 1) for A/B runs.
    Collective call: every rank constructs its reducer at the same point.  Returns None (default group) off NCCL."""
    global _own_group
    if not dist.is_initialized() or dist.get_world_size() == 1 or dist.get_backend() != "nccl":
        return None
    if _own_group is None:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = os.environ.get("HVLM_NCCL_HIGH_PRIO", "1") != "0"
        max_ctas = int(os.environ.get("HVLM_NCCL_MAX_CTAS", "0"))
        if max_ctas > 0:
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = 1
        pol = int(os.environ.get("HVLM_NCCL_CTA_POLICY", "1"))      # NCCL_CTA_POLICY_EFFICIENCY
        if pol >= 0:
            opts.config.cta_policy = pol
        _own_group = dist.new_group(backend="nccl", pg_options=opts)
    return _own_group


class ProjectorGradReducer:
    """The training-shaped variant's only exchange step (the reference leaves it to ZeRO's reduce,
    scripts/zero3.json:16-27): mean all-reduce of ``mm_projector.{weight,bias}.grad`` over NCCL / NVLink.

    * ONE flat fp32 bucket ``[D*1024 + D]`` allocated once.  With ``use_sink=True`` the wgrad GEMM and the bias column
      sum of ``hvlm::linear``'s backward write their fp32 results STRAIGHT into it (``ops.register_wgrad_sink``): no
      ``cat``, no packing pass.  (Gradient accumulation over several backwards per reduction needs ``use_sink=False``:
      the bucket is then packed from ``.grad``.)
    * ``reduce_async()`` right after ``backward()``: the collective is enqueued on a side stream behind the backward
      kernels (``ReduceOp.AVG`` on NCCL: no separate division pass) and the call returns immediately.
    * ``wait()`` before the gradients are consumed (optimizer step): the current stream waits for the collective and the
      reduced values are cast into ``.grad``.  The vision tower is frozen, so the NEXT step's ViT forward does not depend
      on the projector update and may be enqueued before ``wait()`` -- ``arch`` calls ``projector._hvlm_pre_forward()``
      right before the projector GEMM, which is where :meth:`attach` hooks the wait (+ an optional update callback).
    Every rank issues exactly one fixed-size collective per ``reduce_async`` call, whatever gradients it has."""

    def __init__(self, projector: torch.nn.Module, group=None, use_sink: bool = True):
        self.projector = projector
        if group is None:
            group = _dedicated_nccl_group()
        self.group = group
        self.params = [projector.weight] + ([projector.bias] if getattr(projector, "bias", None) is not None else [])
        self.nW, self.nb = _flat_sizes(projector)
        dev = projector.weight.device
        self.bucket = torch.zeros(self.nW + self.nb, dtype=torch.float32, device=dev)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.pending = None       # (work handle | None, done event | None)
        self.use_sink = use_sink and dev.type == "cuda"
        self.on_reduced = None    # optional callback run by wait() after the gradients are in place (optimizer step)
        self.last_ms = None
        self._ev = None
        if self.use_sink:
            from . import ops
            ops.register_wgrad_sink(tuple(projector.weight.shape), dev, self.bucket[: self.nW].view(projector.weight.shape),
                                    self.bucket[self.nW:] if self.nb else None)

    def close(self):
        if self.use_sink:
            from . import ops
            ops.unregister_wgrad_sink(tuple(self.projector.weight.shape), self.projector.weight.device)
        if getattr(self.projector, "_hvlm_pre_forward", None) == self.wait:
            self.projector._hvlm_pre_forward = None

    def attach(self, on_reduced=None):
        """Defer ``wait()`` to the moment the projector is next used (see class docstring)."""
        self.on_reduced = on_reduced
        self.projector._hvlm_pre_forward = self.wait
        return self

    def _pack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.bucket[off: off + n].zero_()
            else:
                self.bucket[off: off + n].copy_(p.grad.reshape(-1))
            off += n

    def reduce_async(self, timed: bool = False):
        if self.pending is not None:
            self.wait()
        if not self.use_sink:
            self._pack()
        elif any(p.grad is None for p in self.params):
            self._pack()                      # no backward reached the projector on this rank: contribute zeros
        if self.world == 1:
            self.pending = (None, None, None)
            return
        if self.stream is None:               # gloo / CPU
            work = dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending = (work, None, None)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            e0 = e1 = None
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            op = dist.ReduceOp.AVG if self.backend == "nccl" else dist.ReduceOp.SUM
            dist.all_reduce(self.bucket, op=op, group=self.group)
            if op != dist.ReduceOp.AVG:
                self.bucket.div_(self.world)
            if timed:
                e1.record()
            done = torch.cuda.Event()
            done.record()
        self.pending = (None, done, (e0, e1) if timed else None)

    def wait(self):
        if self.pending is None:
            return
        work, done, tim = self.pending
        self.pending = None
        if work is not None:
            work.wait()
            self.bucket.div_(self.world)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        self._ev = tim
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.bucket[off: off + n].view(p.shape)
            if p.grad is None:
                p.grad = g.to(p.dtype)
            else:
                p.grad.copy_(g)
            off += n
        if self.on_reduced is not None:
            self.on_reduced()

    def last_allreduce_ms(self):
        """Device time of the last timed collective (CUDA events on the side stream); synchronises."""
        if self._ev is None:
            return None
        e0, e1 = self._ev
        e1.synchronize()
        return e0.elapsed_time(e1)


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
