"""Multi-GPU plumbing for the scoring path: one process per GPU, torch.distributed (NCCL over NVLink).

Inference shards by image batch: every rank scores its own contiguous slice with replicated frozen
weights, then ONE all-gather of a packed [B_local, 2K] fp32 tile (global logits | local logits)
assembles the global result — the only exchange on the path (the reference itself never
shards inference, §2a).  Prompt tuning shards by caption batch and all-reduces (averages) one flat
prompt-gradient buffer, the semantics of the reference's DDP wrapper (T:786-787).

Class-sharded prompt branch (SURVEY §8e "optional later", §8f-2; opt-in, `DenseCLIPB200.shard_prompt_branch`): the
2-3 x K prompt sequences are split over the ranks instead of replicated — `gather_rows` assembles the [n*K, D] text
features after each rank has run the text tower on its own rows, `sum_over_ranks` adds up the per-rank feature gradients
before each rank back-propagates its rows only.  With the flat gradient average above the result is the replicated one:
mean_r( J_r^T (sum_r' dT_r')[rows_r] ) = (1/G) J^T sum_r' dT_r'."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_logits(logits: torch.Tensor, logits_local: torch.Tensor) -> torch.Tensor:
    return torch.cat([logits, logits_local], dim=1).contiguous()


def unpack_logits(packed: torch.Tensor):
    k = packed.shape[1] // 2
    return packed[:, :k], packed[:, k:]


def all_gather_logits(logits, logits_local, group=None):
    """-> (logits [G*B,K], logits_local [G*B,K]) in rank order; equal B on every rank."""
    packed = pack_logits(logits, logits_local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_logits(packed)
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.shape[0], packed.shape[1]), device=packed.device, dtype=packed.dtype)
    dist.all_gather_into_tensor(out, packed, group=group)
    return unpack_logits(out)


def flat_grads(params):
    """One flat fp32 buffer holding every prompt-parameter gradient (zeros where a grad is None: the
    reference needs find_unused_parameters=True for ctx_evidence / the scalars)."""
    chunks = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params]
    return torch.cat(chunks)


def allreduce_mean_grads(params, group=None):
    """DDP semantics (T:787): average the prompt gradients over ranks with a single all-reduce."""
    params = [p for p in params if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    if all(p.is_cuda and p.dtype == torch.float32 for p in params):
        GradBucket.for_params(params).allreduce_mean(group)
        return
    # CPU tensors: only the gloo protocol tests get here (the product runs on CUDA tensors, branch above)
    flat = flat_grads(params)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


class GradBucket:
    """One persistent flat fp32 buffer for the prompt-parameter gradients (the DDP bucket of T:786-787): a multi-tensor
    pack kernel fills it (a parameter without gradient contributes zeros, like find_unused_parameters=True), ONE all-reduce
    sums it over the ranks, and either `unpack` writes the averaged slices back into `.grad` or `PromptSGD.step` consumes
    the flat buffer directly.  Three launches + one collective per step instead of cat / divide / one copy per parameter."""

    _cache = {}

    def __init__(self, params):
        self.params = list(params)
        self.flat = torch.zeros((sum(p.numel() for p in self.params),), device=self.params[0].device, dtype=torch.float32)

    @classmethod
    def for_params(cls, params):
        key = tuple(id(p) for p in params)
        b = cls._cache.get(key)
        if b is None or b.flat.device != params[0].device:
            b = cls._cache[key] = cls(params)
        return b

    def pack(self):
        from . import ops
        grads = [None if p.grad is None else p.grad.contiguous() for p in self.params]
        ops.pack_f32(grads, [p.data for p in self.params], flat=self.flat)
        return self.flat

    def unpack(self, scale=1.0):
        from . import ops
        for p in self.params:
            if p.grad is None:
                p.grad = torch.empty_like(p.data)
        ops.unpack_scale_f32(self.flat, [p.grad for p in self.params], scale)

    def allreduce_sum(self, group=None):
        """pack + all-reduce; returns the 1 / world factor that turns the flat sum into DDP's average."""
        self.pack()
        if multi_rank(group):
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0

    def allreduce_mean(self, group=None):
        self.unpack(self.allreduce_sum(group))


class PromptSGD:
    """torch.optim.SGD (momentum, weight decay, dampening 0, no nesterov — the optimiser the reference builds for the
    prompt learner, T:773 / dassl/optim/optimizer.py) as ONE multi-tensor launch reading the flat gradient bucket."""

    def __init__(self, params, lr, momentum=0.9, weight_decay=0.0):
        self.params = [p for p in params]
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.bufs = [torch.zeros_like(p.data) for p in self.params] if momentum else None
        self.bucket = GradBucket.for_params(self.params)

    @torch.no_grad()
    def step(self, group=None):
        self.pack()
        self.reduce_and_update(group)

    @torch.no_grad()
    def pack(self):
        """First half of step(): gradients -> flat bucket (one launch; safe inside a CUDA-graph capture of the step)."""
        self.bucket.pack()

    @torch.no_grad()
    def reduce_and_update(self, group=None):
        """Second half: all-reduce the bucket over the ranks (if any) and apply the fused SGD update from it."""
        from . import ops
        scale = 1.0
        if multi_rank(group):
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=group)
            scale = 1.0 / dist.get_world_size(group)
        ops.sgd_step(self.bucket.flat, [p.data for p in self.params], self.bufs, self.lr, self.momentum, self.weight_decay,
                     grad_scale=scale)

    def zero_grad(self):
        for p in self.params:
            p.grad = None


def broadcast_params(params, src: int = 0, group=None):
    """DDP's construction-time broadcast (T:786-787): every rank starts from rank `src`'s prompt
    parameters — `ctx*` are drawn from each process's own RNG (T:134-151).  Required by the class-sharded prompt branch,
    where a rank consumes text features computed from another rank's copy of the parameters."""
    if not multi_rank(group):
        return
    with torch.no_grad():
        for p in params:
            dist.broadcast(p.data, src=src, group=group)


def multi_rank(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def my_chunk(n_rows: int, group=None):
    """chunk_range of the calling rank."""
    return chunk_range(n_rows, dist.get_rank(group), dist.get_world_size(group))


def chunk_range(n_rows: int, rank: int, world: int):
    """Equal chunks of ceil(n_rows / world) rows (the last ranks may get fewer, or none): the layout
    all_gather_into_tensor needs.  -> (lo, hi, rows per chunk)."""
    per = -(-n_rows // world)
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows), per


def gather_rows(own: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """own = this rank's chunk (chunk_range) of an [n_rows, D] matrix -> the whole matrix on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi, per = chunk_range(n_rows, rank, world)
    assert own.shape[0] == hi - lo, (own.shape, lo, hi)
    buf = own.new_zeros((per,) + tuple(own.shape[1:]))
    buf[:hi - lo] = own
    out = own.new_empty((world * per,) + tuple(own.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    return out[:n_rows].contiguous()


def sum_over_ranks(t: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
