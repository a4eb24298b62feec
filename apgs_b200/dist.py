"""Multi-GPU plumbing for the APG sweep: one process per GPU, the instance batch B is sharded
across ranks with NO data-path collective (every op normalises / scans / gathers over S within one
instance); the only exchange is one all-reduce of the stacked per-sweep metric sums.
The reference has no distributed code at all -- this is new functionality.
"""
import torch
import torch.distributed as dist


def shard_bounds(B, rank, world):
    """contiguous instance range [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_instances(x, rank, world, dim=0):
    lo, hi = shard_bounds(x.shape[dim], rank, world)
    return x.narrow(dim, lo, hi - lo)


def shard_uniforms(u_global, rank, world):
    """slice of ONE global (B,...) uniform vector: a B-sharded run then draws exactly the ancestors of
    the single-GPU run."""
    lo, hi = shard_bounds(u_global.shape[0], rank, world)
    return u_global[lo:hi]


def metric_sums(trace):
    """(num_sweeps, 4) tensor of [sum_b ess, sum_b mean_s density, sum_b log Z-hat, B_local] from a sweep trace
    (log Z-hat_b = logsumexp_s log_w - log S, trace['log_z']; zeros when the objective did not record it)."""
    def _get(k):
        return trace[k] if k in trace and torch.is_tensor(trace[k]) else None
    ess, dens, logz = _get('ess'), _get('density'), _get('log_z')
    ref = ess if ess is not None else dens
    n = ref.shape[0]
    out = torch.zeros(n, 4, dtype=torch.float64, device=ref.device)
    if ess is not None:
        out[:, 0] = ess.double().sum(-1)
    if dens is not None:
        out[:, 1] = dens.double().mean(1).sum(-1)
    if logz is not None and logz.shape[0] == n:
        out[:, 2] = logz.double().sum(-1)
    out[:, 3] = ref.shape[-1]
    return out


def allreduce_metrics(trace, group=None):
    """global per-sweep means {'ess', 'density', 'log_z': (num_sweeps,)} over ALL instances of all ranks: one
    all-reduce(SUM) of a (num_sweeps, 4) fp64 tensor -- latency-bound, issued once per objective (north_star: "NCCL
    over NVLink is used only for the log-marginal/ESS metric allreduce")."""
    sums = metric_sums(trace)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    cnt = sums[:, 3].clamp_min(1)
    return {'ess': sums[:, 0] / cnt, 'density': sums[:, 1] / cnt, 'log_z': sums[:, 2] / cnt, 'count': sums[:, 3]}
