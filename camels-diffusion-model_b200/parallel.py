"""One process per GPU (torch.distributed over NCCL/NVLink for the plumbing).

Sampling, likelihood and ELBO evaluation shard the batch contiguously over the ranks and need NO
data-path collective (independent samples / maps); only the per-forward shortcut table has to
be identical on every rank, and results are gathered once at the end.  Training is data-parallel: gradient
all-reduce + cross-rank BatchNorm statistics (see train.py).
"""
import inspect

import torch
import torch.distributed as dist


def _accepts(fn, name):
    try:
        ps = inspect.signature(fn).parameters
    except (TypeError, ValueError):
        return False
    return name in ps or any(p.kind == inspect.Parameter.VAR_KEYWORD for p in ps.values())


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total, rank, world_size):
    """Contiguous shard [start, end) of n_total items; the first n_total % world ranks get one extra item."""
    base, rem = divmod(n_total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_from_rank0(t, device=None):
    """Make `t` (e.g. the shortcut table drawn from rank 0's CPU generator) identical on every rank."""
    rank, ws = world()
    if ws == 1:
        return t
    backend = dist.get_backend()
    buf = t.to(device) if (backend == "nccl" and device is not None) else t.clone()
    dist.broadcast(buf, src=0)
    return buf.to(t.device)


def gather_shards(local, n_total, dim=0):
    """All-gather ragged contiguous shards back into the full batch (every rank gets the result)."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    max_n = max(e - s for s, e in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = max_n
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(outs, padded)
    return torch.cat([o.narrow(dim, 0, e - s) for o, (s, e) in zip(outs, sizes)], dim)


def sample_sharded(sample_fn, x_T_all, params_all, shortcut_tab, gather=True, device=None):
    """Batch-sharded sampling: rank r runs `sample_fn(x_T_shard, params_shard, shortcut_tab)` on its contiguous
    shard.  `shortcut_tab` is taken from rank 0.  Returns the gathered [n_total, ...] samples (or the local shard).
    A `sample_fn` that takes a `sample_offset` keyword receives the shard's first global sample index: forwarded to
    the sampler (`_SamplerRun(..., sample_offset=)`), it keys the in-kernel noise by the GLOBAL sample, so ranks never
    share noise and the sharded run draws what the single-process run draws."""
    rank, ws = world()
    n = x_T_all.shape[0]
    s, e = shard_range(n, rank, ws)
    tab = broadcast_from_rank0(shortcut_tab, device)
    prm = None if params_all is None else params_all[s:e]
    kw = {"sample_offset": s} if _accepts(sample_fn, "sample_offset") else {}
    local = sample_fn(x_T_all[s:e], prm, tab, **kw)
    return gather_shards(local, n) if gather else local


def reduce_mean_scalar(total, count, device=None):
    """Dataset-level mean of per-rank (sum, count) pairs — the single scalar exchange of the sharded NLL / ELBO."""
    rank, ws = world()
    if ws == 1:
        return total / max(count, 1)
    t = torch.tensor([float(total), float(count)], dtype=torch.float64,
                     device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t)
    return float(t[0] / t[1])


def evaluate_sharded(eval_fn, maps, params, batch_size=32, device=None):
    """Dataset-level NLL / ELBO over all ranks (BASELINE config 5: 4096 maps on 8 GPUs): rank r evaluates its
    contiguous shard of `maps` / `params` in batches of `batch_size` — `eval_fn(loader)` is e.g.
    `lambda dl: calculate_likelihood(model, dl, T, dev, ab_t, b_t, a_t)` or the ELBO/BPD variant and returns the
    shard MEAN (a float, or a tuple of floats) — and the per-rank (sum, count) pairs meet in one scalar all-reduce.
    No other communication: the maps are independent.  An `eval_fn` that takes a `sample_offset` keyword receives the
    shard's first global map index (pass it on: `calculate_likelihood(..., sample_offset=sample_offset)`), which keys
    the in-kernel noise by the global map so that no two ranks evaluate with the same noise."""
    rank, ws = world()
    s, e = shard_range(maps.shape[0], rank, ws)
    loader = [(maps[i:min(i + batch_size, e)], None if params is None else params[i:min(i + batch_size, e)])
              for i in range(s, e, batch_size)]
    n_local = e - s
    kw = {"sample_offset": s} if _accepts(eval_fn, "sample_offset") else {}
    res = eval_fn(loader, **kw) if n_local > 0 else ()  # a rank without maps contributes (0, 0) to every output
    vals = [float(v) for v in res] if isinstance(res, tuple) else [float(res)]
    arity = len(vals)
    if ws > 1:  # ranks with an empty shard do not know how many outputs eval_fn has
        a = torch.tensor([arity], device=device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        arity = int(a.item())
    vals += [0.0] * (arity - len(vals))
    out = tuple(reduce_mean_scalar(v * n_local, n_local, device) for v in vals)
    return out if (isinstance(res, tuple) and arity > 1) or arity > 1 else out[0]


class PeerExchange:
    """Peer group for the fused reduce + cross-rank exchange kernels (cdm_xrank in include/cdm_b200.h).

    torch's symmetric memory is the plumbing (one buffer per rank, mapped into every peer over NVLink); the
    exchange itself is our kernel: P2P stores into every rank's slot, a sequence-numbered flag per peer, a
    rank-ordered sum — so the 36 latency-bound [2C] all-reduces of a data-parallel training step (cross-rank
    BatchNorm statistics, forward and backward) cost no NCCL launch and give bit-identical sums on every rank."""
    MAX_N = 512

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        n_slot = 2 * self.world * self.MAX_N
        err = None
        try:
            self.buf = symm.empty(n_slot + 64, dtype=torch.float32, device=device)
            self.buf.zero_()
        except Exception as ex:  # noqa: BLE001
            err = ex
        # agree before the (collective) rendezvous: a rank that could not allocate must not leave the others waiting
        ok = torch.tensor([0 if err is not None else 1], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError(f"symmetric memory unavailable on at least one rank ({err})")
        handle = symm.rendezvous(self.buf, group.group_name)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        self.slot_ptrs = torch.tensor(ptrs, dtype=torch.int64).to(device)
        self.flag_ptrs = torch.tensor([p + 4 * n_slot for p in ptrs], dtype=torch.int64).to(device)
        self.seq = torch.zeros(1, dtype=torch.int32, device=device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        self._handle = handle
        self.args = L.XrankArgs(self.rank, self.world, self.slot_ptrs.data_ptr(), self.flag_ptrs.data_ptr(),
                                self.seq.data_ptr(), self.ticket.data_ptr())
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's buffer is zeroed and mapped before the first exchange
