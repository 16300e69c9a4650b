"""Batch sharding across GPUs: one process per GPU, every op on the hot path is per-sample, so
inference needs no collective; training adds ONE gradient all-reduce (sum, then / world) per optimizer step.

The reference's train scripts call `backward()` and `optimizer.step()` inside `GANOptimizer.__call__`
(modules/loss.py:126-132), so the all-reduce cannot be inserted by the caller. `GradientAllReducer` therefore hangs on
 * `Tensor.register_post_accumulate_grad_hook` of every trainable parameter — when all parameters of a bucket have their
   gradient, the bucket is flattened and `all_reduce` is launched asynchronously (NCCL over NVLink/NVSwitch on GPUs, gloo
   in the CPU tests), overlapping the rest of backward;
 * `Optimizer.register_step_pre_hook` — flushes buckets whose members never produced a gradient (14 tensors in PICNet:
   `Auto_Attn.alpha`, `Auto_Attn.model.*`), waits, divides by the world size and writes the averaged
   gradients back before the optimizer reads them.
torch.distributed is plumbing here; no arithmetic of the hot path lives in this file.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the default group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_batch(n_items: int, rank: int, world: int) -> range:
    """Contiguous, near-even split of `n_items` samples; the first (n_items % world) ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def all_ranks_finite(value: torch.Tensor, group=None) -> bool:
    """Every rank must take the same skip/step decision for a non-finite loss (train_psp.py:328-335)."""
    ok = torch.isfinite(value.detach()).all().to(torch.float32).reshape(1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return bool(ok.item() > 0)


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Parameters AND buffers (incl. SpectralNorm u/v, which mutate every forward) from rank `src`."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class _Bucket:
    __slots__ = ("params", "pending", "flat", "work", "launched", "averaged")

    def __init__(self, params):
        self.params: List[torch.nn.Parameter] = params
        self.pending = len(params)
        self.flat = None
        self.work = None
        self.launched = False
        self.averaged = False


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        ps = [p for p in params if p.requires_grad]
        ps.reverse()  # gradients arrive roughly in reverse registration order
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in ps:
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(_Bucket(cur))
        self._owner = {id(p): b for b in self.buckets for p in b.params}
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in ps]
        self._comm_stream = torch.cuda.Stream() if ps and ps[0].is_cuda else None
        self.enabled = True

    # ------------------------------------------------------------------ hooks
    def _on_grad(self, p):
        if not self.enabled or self.world == 1:
            return
        b = self._owner[id(p)]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        live = [p for p in b.params if p.grad is not None]
        b.launched = True
        if not live:
            return
        b.flat = torch.cat([p.grad.reshape(-1) for p in live])
        # NCCL averages inside the collective; gloo has no AVG: sum here, one division per bucket in finish()
        avg = b.flat.is_cuda and dist.get_backend(self.group) == "nccl"
        b.averaged = avg
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def attach(self, optimizer: torch.optim.Optimizer):
        optimizer.register_step_pre_hook(lambda opt, args, kwargs: self.finish())
        return self

    @torch.no_grad()
    def finish(self):
        """Flush, wait, average, write back, re-arm. Called right before optimizer.step()."""
        if self.world == 1:
            return
        for b in self.buckets:
            if not b.launched:
                self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                if self._comm_stream is not None:
                    torch.cuda.current_stream().wait_stream(self._comm_stream)
                live = [p for p in b.params if p.grad is not None]
                if not b.averaged:
                    b.flat.div_(self.world)
                # one multi-tensor copy per bucket (a division + a copy per parameter was ~400 tiny launches per step: 3 ms
                # of a 10 ms StyleGAN2 decoder step)
                views = [v.view_as(p.grad) for v, p in zip(b.flat.split([p.numel() for p in live]), live)]
                torch._foreach_copy_([p.grad for p in live], views)
            b.pending, b.flat, b.work, b.launched = len(b.params), None, None, False

    def remove(self):
        for h in self._handles:
            h.remove()
