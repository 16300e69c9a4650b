"""Batch sharding across GPUs: one process per GPU, every op on the hot path is per-sample, so
inference needs no collective; training adds ONE gradient all-reduce (sum, then / world) per optimizer step.

The reference's train scripts call `backward()` and `optimizer.step()` inside `GANOptimizer.__call__`
(modules/loss.py:126-132), so the all-reduce cannot be inserted by the caller. `GradientAllReducer` therefore hangs on
 * `Tensor.register_post_accumulate_grad_hook` of every trainable parameter — when every parameter a bucket expects has its
   gradient, the bucket's PERSISTENT flat buffer is all-reduced in place, asynchronously (NCCL over NVLink/NVSwitch on GPUs,
   gloo in the CPU tests), overlapping the rest of backward. `.grad` of every parameter is a VIEW into that buffer, so there is
   no flatten and no copy back; a gradient that autograd allocated afresh (after `zero_grad(set_to_none=True)`) is moved into
   its view with one multi-tensor copy per bucket;
 * `Tensor.register_hook` (fires BEFORE accumulation) — a second backward between two optimizer steps (the GANOptimizer order:
   G_loss.backward() runs through the unfrozen discriminator, then zero_grad, then D_loss.backward(); or plain gradient
   accumulation) makes the compute stream wait for the bucket's in-flight collective and marks the bucket dirty;
 * `Optimizer.register_step_pre_hook` — launches buckets whose members never produced a gradient (14 tensors in PICNet:
   `Auto_Attn.alpha`, `Auto_Attn.model.*`), RE-reduces dirty buckets from the current gradients (averaging is
   linear and idempotent on values already equal on all ranks, so avg(avg(g1) + g2) = avg(g1) + avg(g2)), and makes the
   optimizer's stream wait for the collectives.
Collectives are issued in bucket order on every rank. torch.distributed is plumbing here; no arithmetic of the hot path lives
in this file.
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


def replicas_in_sync(params, group=None, rtol: float = 0.0) -> bool:
    """True when every rank holds the same values in `params` (data-parallel replicas after N optimizer steps on averaged
    gradients): max over ranks == min over ranks of a per-tensor fingerprint (sum and sum of squares in double). Used by the
    bench after the timed region — a gradient all-reduce that silently dropped out of a captured CUDA graph would show here."""
    ps = [p.detach() for p in params]
    if not ps or not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return True
    fp = torch.stack([torch.stack((p.double().sum(), (p.double() ** 2).sum())) for p in ps]).reshape(-1)
    hi, lo = fp.clone(), fp.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    return bool(((hi - lo).abs() <= rtol * hi.abs().clamp_min(1e-30)).all().item())


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Parameters AND buffers (incl. SpectralNorm u/v, which mutate every forward) from rank `src`."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def enable_for_scripts() -> tuple[int, int, int]:
    """Data-parallel training of an UNCHANGED reference script under torchrun (run.py): the scripts build their optimizers
    themselves and call backward() / step() inside callees (modules/loss.py:126-132, train_psp.py:333-335), so the hooks go
    on `torch.optim.Optimizer.__init__`: every optimizer gets its parameters broadcast from rank 0 (SpectralNorm's u / v are
    parameters there, external_function.py:36-41) and a GradientAllReducer. Model construction runs under one common seed;
    after the first optimizer exists every rank is re-seeded with its own seed, so DataLoader shuffling, rsample() and the
    StyleGAN2 noise differ per rank (each rank trains on its own random batches). Only rank 0 writes checkpoints."""
    import random
    rank, local_rank, world = init_from_env()
    if world == 1:
        return rank, local_rank, world
    torch.manual_seed(0)
    random.seed(0)
    plain_init = torch.optim.Optimizer.__init__

    def init(self, params, defaults):
        plain_init(self, params, defaults)
        ps = [p for g in self.param_groups for p in g["params"]]
        for p in ps:
            dist.broadcast(p.data, src=0)
        self._fmi_reducer = GradientAllReducer([p for p in ps if p.requires_grad]).attach(self)
        torch.manual_seed(1000 + rank)
        random.seed(1000 + rank)

    torch.optim.Optimizer.__init__ = init
    if rank != 0:
        torch.save = lambda *a, **k: None
    return rank, local_rank, world


class _Bucket:
    __slots__ = ("params", "flat", "views", "arrived", "n_arrived", "expect", "work", "launched", "dirty", "index")

    def __init__(self, params, index):
        self.params: List[torch.nn.Parameter] = params
        self.index = index
        p0 = params[0]
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=p0.dtype, device=p0.device)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split([p.numel() for p in params]), params)]
        self.arrived = [False] * len(params)
        self.n_arrived = 0
        self.expect = len(params)      # how many members must report before the bucket is launched from the hooks
        self.work = None
        self.launched = False
        self.dirty = False


class GradientAllReducer:
    """See the module docstring. `bucket_bytes` caps a bucket; the first bucket (the LAST layers, whose gradients arrive
    first) and the last one (the first layers, whose collective cannot overlap anything) are capped at `edge_bytes`, so the
    first collective starts early and the exposed tail after backward is short."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None,
                 edge_bytes: Optional[int] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        ps = [p for p in params if p.requires_grad]
        ps.reverse()  # gradients arrive roughly in reverse registration order
        self.buckets: List[_Bucket] = []
        self._handles = []
        self.enabled = True
        self._comm_stream = None
        self._next = 0
        if self.world == 1 or not ps:
            return
        edge = min(bucket_bytes, edge_bytes if edge_bytes is not None else max(bucket_bytes // 16, 1 << 20))
        sizes = [p.numel() * p.element_size() for p in ps]
        # members of the tail bucket: the last parameters (in arrival order) that fit into `edge`
        tail_from, acc = len(ps), 0
        while tail_from > 1 and acc + sizes[tail_from - 1] <= edge:
            tail_from -= 1
            acc += sizes[tail_from]
        groups, cur, cur_bytes = [], [], 0
        for i, p in enumerate(ps):
            cap = edge if not groups else bucket_bytes
            if cur and (cur_bytes + sizes[i] > cap or i == tail_from or p.dtype != cur[0].dtype
                        or p.device != cur[0].device):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += sizes[i]
        if cur:
            groups.append(cur)
        self.buckets = [_Bucket(g, i) for i, g in enumerate(groups)]
        self._owner = {id(p): (b, i) for b in self.buckets for i, p in enumerate(b.params)}
        for p in ps:
            self._handles.append(p.register_hook(self._make_pre(p)))
            self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._comm_stream = torch.cuda.Stream(ps[0].device) if ps[0].is_cuda else None
        self._nccl_avg = ps[0].is_cuda and dist.get_backend(self.group) == "nccl"

    # ------------------------------------------------------------------ hooks
    def _make_pre(self, p):
        b, _ = self._owner[id(p)]

        def pre(grad):                      # before autograd accumulates `grad` into p.grad
            if self.enabled and b.launched:
                self._wait(b)               # p.grad may be the buffer an in-flight collective is working on
                b.dirty = True
            return None
        return pre

    def _on_grad(self, p):
        if not self.enabled:
            return
        b, i = self._owner[id(p)]
        if b.launched:
            b.dirty = True
            return
        if not b.arrived[i]:
            b.arrived[i] = True
            b.n_arrived += 1
        # in bucket order only, so that every rank issues the same sequence of collectives
        while self._next < len(self.buckets):
            nb = self.buckets[self._next]
            if nb.launched:
                self._next += 1
            elif nb.n_arrived >= nb.expect and nb.n_arrived > 0:
                self._launch(nb)
                self._next += 1
            else:
                break

    def _adopt(self, b: _Bucket):
        """Make every existing gradient of the bucket the view into its flat buffer (one multi-tensor copy for those that
        are not). Members without a gradient keep `.grad is None` (the optimizer skips them); their slice holds zeros."""
        src, dst = [], []
        for p, v in zip(b.params, b.views):
            g = p.grad
            if g is None:
                continue
            if g.data_ptr() != v.data_ptr() or g.shape != v.shape or not g.is_contiguous():
                src.append(g.detach())
                dst.append(v)
                p.grad = v
        if src:
            torch._foreach_copy_(dst, src)

    def _wait(self, b: _Bucket):
        if b.work is not None:
            b.work.wait()                   # NCCL: the current stream waits for the collective; no host block
            b.work = None

    def _launch(self, b: _Bucket):
        self._wait(b)
        self._adopt(b)
        b.launched, b.dirty = True, False
        if self._nccl_avg:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            # gloo (CPU tests) has no AVG: sum, then divide at once so that the buffer always holds averaged values
            if self._comm_stream is not None:
                torch.cuda.current_stream().synchronize()
            dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group)
            b.flat.div_(self.world)

    def attach(self, optimizer: torch.optim.Optimizer, keep_views: bool = True):
        """finish() before every optimizer.step(); and (keep_views) `optimizer.zero_grad()` keeps the gradient views alive (it
        zeroes the flat buffers — one memset per bucket — instead of dropping the gradients) when every parameter of the
        optimizer is managed here. Parameters that never had a gradient keep `.grad is None` either way."""
        optimizer.register_step_pre_hook(lambda opt, args, kwargs: self.finish())
        if self.world > 1 and keep_views:
            mine = {id(p) for b in self.buckets for p in b.params}
            theirs = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
            if all(id(p) in mine for p in theirs) and len(theirs) == len(mine):
                plain = optimizer.zero_grad

                def zero_grad(set_to_none: bool = True):
                    if not self.enabled:
                        return plain(set_to_none=set_to_none)
                    self.zero_grad()
                optimizer.zero_grad = zero_grad
        return self

    @torch.no_grad()
    def zero_grad(self):
        for b in self.buckets:
            self._wait(b)
            stray = [p for p, v in zip(b.params, b.views) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
            for p in stray:                 # gradients that are not views (assigned by the caller): drop them
                p.grad = None
            b.flat.zero_()
            # whatever was reduced so far is gone: re-arm, so that the next backward launches from its hooks again
            b.arrived = [False] * len(b.params)
            b.n_arrived, b.launched, b.dirty = 0, False, False
        self._next = 0

    @torch.no_grad()
    def finish(self):
        """Launch what the hooks could not, re-reduce dirty buckets, make the current stream wait for every collective, and
        re-arm. Called right before optimizer.step() (and callable by hand: gradients are final when it returns)."""
        if self.world == 1 or not self.enabled:
            return
        for b in self.buckets:
            if not b.launched or b.dirty or any(
                    p.grad is not None and p.grad.data_ptr() != v.data_ptr() for p, v in zip(b.params, b.views)):
                self._launch(b)
        for b in self.buckets:
            self._wait(b)
            got = b.n_arrived
            b.expect = got if got > 0 else len(b.params)
            b.arrived = [False] * len(b.params)
            b.n_arrived, b.launched, b.dirty = 0, False, False
        self._next = 0

    def remove(self):
        for h in self._handles:
            h.remove()
