"""N > 1 host logic on CPU with the gloo backend, world_size 2: batch sharding, gradient all-reduce through the
post-accumulate-grad / optimizer-pre-step hooks (incl. a parameter that never gets a gradient and an optimizer step
hidden inside a callee, as in modules/loss.py:126-132), state broadcast, and the agreed non-finite decision."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 8)
        self.b = torch.nn.Linear(8, 3)
        self.unused = torch.nn.Parameter(torch.ones(4))  # like Auto_Attn.alpha / model.* (never receives a gradient)
        self.register_buffer("u", torch.randn(5))

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from face_mask_inpaint_b200 import dist as fd
    r, lr, w = fd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)          # different init per rank ...
    net = Net()
    fd.broadcast_module_state(net)          # ... made identical, buffers included
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    idx = list(fd.shard_batch(8, rank, world))
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    red = fd.GradientAllReducer(net.parameters(), bucket_bytes=256).attach(opt)   # tiny buckets: several of them

    def hidden_step():                      # backward + step inside a callee, like GANOptimizer.__call__
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(net(x[idx]), y[idx])
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        loss = hidden_step()
    finite = fd.all_ranks_finite(loss if rank == 0 else torch.tensor(float("nan")))
    red.remove()
    q.put((rank, {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}, len(red.buckets), finite, idx))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process_step():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the concatenated batch, starting from rank 0's initial state
    torch.manual_seed(100)
    net = Net()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(x), y).backward()
        opt.step()
    want = net.state_dict()
    assert res[0][4] == [0, 1, 2, 3] and res[1][4] == [4, 5, 6, 7]
    assert res[0][2] >= 2                      # more than one bucket was exercised
    for rank, sd, _, finite, _ in res:
        assert finite is False                 # one rank saw a NaN -> every rank agrees to skip
        for k in want:
            assert torch.allclose(torch.from_numpy(sd[k]), want[k], atol=1e-6), (rank, k)


class _G(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 8)
        self.b = torch.nn.Linear(8, 5)
        self.alpha = torch.nn.Parameter(torch.zeros(1))     # never used, like Auto_Attn.alpha

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


class _D(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(5, 7)
        self.b = torch.nn.Linear(7, 1)

    def forward(self, x):
        return self.b(torch.nn.functional.leaky_relu(self.a(x), 0.1))


def _gan_steps(G, D, optG, optD, x, real, steps, set_to_none=True):
    """The GANOptimizer.__call__ order (modules/loss.py:120-134): G_loss.backward() runs through the UNFROZEN discriminator
    (freeze=False), then optimizer_D.zero_grad(), D_loss.backward(), optimizer_D.step()."""
    for _ in range(steps):
        fake = G(x)
        g_loss = (D(fake) - 1).square().mean() + (fake - real).abs().mean()
        optG.zero_grad(set_to_none=set_to_none)
        g_loss.backward()
        optG.step()
        d_loss = 0.5 * ((D(real) - 1).square().mean() + D(fake.detach()).square().mean())
        optD.zero_grad(set_to_none=set_to_none)
        d_loss.backward()
        optD.step()


def _accum_steps(net, opt, x, y, steps, micro):
    """Gradient accumulation: `micro` backwards, one step."""
    for _ in range(steps):
        opt.zero_grad()
        for xs, ys in zip(x.chunk(micro), y.chunk(micro)):
            (torch.nn.functional.mse_loss(net(xs), ys) / micro).backward()
        opt.step()


def _gan_data():
    g = torch.Generator().manual_seed(3)
    return torch.randn(8, 6, generator=g), torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)


def _worker_gan(rank, world, port, q, keep_views, set_to_none):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from face_mask_inpaint_b200 import dist as fd
    fd.init_from_env("gloo")
    torch.manual_seed(7)
    G, D, net = _G(), _D(), Net()
    x, real, y = _gan_data()
    idx = list(fd.shard_batch(8, rank, world))
    optG, optD = torch.optim.SGD(G.parameters(), lr=0.1), torch.optim.SGD(D.parameters(), lr=0.1)
    fd.GradientAllReducer(G.parameters(), bucket_bytes=128).attach(optG, keep_views=keep_views)
    rd = fd.GradientAllReducer(D.parameters(), bucket_bytes=128).attach(optD, keep_views=keep_views)
    _gan_steps(G, D, optG, optD, x[idx], real[idx], 3, set_to_none)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    fd.GradientAllReducer(net.parameters(), bucket_bytes=128).attach(opt, keep_views=keep_views)
    # every rank accumulates over 2 micro-batches of its shard; shard mean = mean of the two micro means
    _accum_steps(net, opt, x[idx], y[idx], 2, 2)
    q.put((rank, {f"{n}.{k}": v.detach().numpy().copy() for n, m in (("G", G), ("D", D), ("net", net))
                  for k, v in m.state_dict().items()}, len(rd.buckets)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("keep_views,set_to_none", [(True, True), (False, True), (False, False)])
def test_gan_order_and_gradient_accumulation(keep_views, set_to_none):
    """a parameter that gets more than one backward between optimizer steps. The discriminator's buckets are
    reduced during G_loss.backward() (stale), then zero_grad + D_loss.backward(): the step must use the averaged D-loss
    gradient. Same for gradient accumulation. Equality with the single-process step on the whole batch."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_gan, args=(r, world, port, q, keep_views, set_to_none)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(7)
    G, D, net = _G(), _D(), Net()
    x, real, y = _gan_data()
    optG, optD = torch.optim.SGD(G.parameters(), lr=0.1), torch.optim.SGD(D.parameters(), lr=0.1)
    _gan_steps(G, D, optG, optD, x, real, 3)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    _accum_steps(net, opt, x, y, 2, 1)
    want = {f"{n}.{k}": v for n, m in (("G", G), ("D", D), ("net", net)) for k, v in m.state_dict().items()}
    assert res[0][2] >= 2
    for rank, sd, _ in res:
        for k in want:
            assert torch.allclose(torch.from_numpy(sd[k]), want[k], atol=2e-6), (rank, k)


def test_shard_batch_covers_everything():
    from face_mask_inpaint_b200.dist import shard_batch
    for n in (0, 1, 7, 8, 13):
        for w in (1, 2, 4, 8):
            got = [i for r in range(w) for i in shard_batch(n, r, w)]
            assert got == list(range(n))


def _sync_worker(rank, world, port, q):
    import torch.distributed as dist
    from face_mask_inpaint_b200 import dist as fd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    same = [torch.nn.Parameter(torch.arange(6.0)), torch.nn.Parameter(torch.ones(3, 2))]
    diff = [torch.nn.Parameter(torch.arange(6.0) + (1e-3 if rank else 0.0))]
    q.put((rank, fd.replicas_in_sync(same), fd.replicas_in_sync(same + diff)))
    dist.destroy_process_group()


def test_replicas_in_sync_detects_divergence():
    """dist.replicas_in_sync (the bench's post-run check that the captured gradient all-reduce really ran): equal replicas pass,
    a rank whose parameters drifted is detected — on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 37
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert got == [(0, True, False), (1, True, False)]
