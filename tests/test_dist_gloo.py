"""Host-side multi-rank logic on CPU (gloo, world_size 2): batch sharding, packed-logits all-gather order,
prompt-gradient averaging with DDP semantics (T:786-787).  The GPU path uses the same functions over NCCL."""
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


def _worker(rank, world, port, fn_name):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _load_dist_module():
    # loaded by path: the package __init__ needs liblecb.so, this module is pure torch.distributed plumbing
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "language-enhanced-clip-for-multi-label-image-recognition_b200", "dist.py")
    spec = importlib.util.spec_from_file_location("_lecb200_dist", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _check_gather(rank, world):
    D = _load_dist_module()
    n_total, k = 10, 6
    full = torch.arange(n_total * k, dtype=torch.float32).view(n_total, k)
    lo, hi = D.shard_range(n_total, rank, world)
    assert hi - lo == n_total // world
    lg, ll = D.all_gather_logits(full[lo:hi].clone(), -full[lo:hi].clone())
    assert torch.equal(lg, full) and torch.equal(ll, -full)


def _check_broadcast(rank, world):
    D = _load_dist_module()
    p = [torch.nn.Parameter(torch.full((2, 3), float(rank + 1))), torch.nn.Parameter(torch.tensor(float(10 * rank)))]
    D.broadcast_params(p)
    assert torch.equal(p[0].data, torch.ones(2, 3)) and p[1].item() == 0.0 and p[0].requires_grad


def _check_grad_mean(rank, world):
    D = _load_dist_module()
    torch.manual_seed(0)
    p1 = torch.nn.Parameter(torch.zeros(3, 4))
    p2 = torch.nn.Parameter(torch.zeros(5))         # "unused" parameter: grad None on every rank
    p3 = torch.nn.Parameter(torch.zeros(()))
    p1.grad = torch.full((3, 4), float(rank + 1))
    p3.grad = torch.tensor(float(10 * (rank + 1))) if rank == 0 else None
    D.allreduce_mean_grads([p1, p2, p3])
    assert torch.allclose(p1.grad, torch.full((3, 4), sum(range(1, world + 1)) / world))
    assert torch.equal(p2.grad, torch.zeros(5))
    assert torch.allclose(p3.grad, torch.tensor(10.0 / world))


class _ShardedRows(torch.autograd.Function):
    """The protocol of train_path._DualPromptHead with a differentiable stand-in for the text tower: forward runs the
    per-row function on this rank's chunk and gathers the rows; backward sums the row gradients over ranks, then
    back-propagates this rank's chunk only (zeros elsewhere)."""

    @staticmethod
    def forward(ctx, D, x, w):
        lo, hi, _ = D.my_chunk(x.shape[0])
        ctx.D, ctx.rows = D, (lo, hi, x.shape)
        ctx.save_for_backward(x[lo:hi], w)
        return D.gather_rows(torch.tanh(x[lo:hi] @ w), x.shape[0])

    @staticmethod
    def backward(ctx, d_t):
        own, w = ctx.saved_tensors
        lo, hi, shape = ctx.rows
        d_t = ctx.D.sum_over_ranks(d_t.contiguous())
        d_pre = d_t[lo:hi] * (1 - torch.tanh(own @ w) ** 2)
        dx = torch.zeros(shape)
        dx[lo:hi] = d_pre @ w.t()
        return None, dx, None


def _check_sharded_prompt_branch(rank, world):
    """Class-sharded prompt branch: gradients of the shared context after the flat average equal the
    replicated branch's, for a row count the ranks share unevenly (7 rows, chunks of 4 + 3)."""
    D = _load_dist_module()
    g = torch.Generator().manual_seed(5)
    n_rows, n_ctx, width, dim = 7, 3, 6, 5
    ctx0 = torch.randn((n_ctx, width), generator=g)
    cls = torch.randn((n_rows, width), generator=g)             # per-class token embeddings (frozen)
    w = torch.randn((width, dim), generator=g)                  # the frozen "tower"
    caps = torch.randn((4, dim), generator=torch.Generator().manual_seed(100 + rank))      # this rank's captions
    y = (torch.rand((4, n_rows), generator=torch.Generator().manual_seed(200 + rank)) < 0.3).float()

    def run(sharded):
        ctx = torch.nn.Parameter(ctx0.clone())
        x = ctx.sum(0, keepdim=True) + cls                      # the shared context enters every row (CSC = False)
        t = _ShardedRows.apply(D, x, w) if sharded else torch.tanh(x @ w)
        logits = caps @ torch.nn.functional.normalize(t, dim=-1).t()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)      # mean over the LOCAL batch
        loss.backward()
        D.allreduce_mean_grads([ctx])
        return logits.detach(), ctx.grad.clone()

    lg_rep, g_rep = run(False)
    lg_sh, g_sh = run(True)
    assert torch.allclose(lg_sh, lg_rep, atol=1e-6)
    assert torch.allclose(g_sh, g_rep, atol=1e-6), (g_sh - g_rep).abs().max()
    assert D.chunk_range(7, 0, 2) == (0, 4, 4) and D.chunk_range(7, 1, 2) == (4, 7, 4) and D.chunk_range(5, 3, 4) == (5, 5, 2)


@pytest.mark.parametrize("fn", ["_check_gather", "_check_broadcast", "_check_grad_mean", "_check_sharded_prompt_branch"])
def test_two_rank_gloo(fn):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), fn), nprocs=world, join=True)


def test_shard_range_covers_everything():
    D = _load_dist_module()
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
