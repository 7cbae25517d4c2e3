"""Read sharding across ranks (one process per GPU).  Reads are independent units: rank r of
`world` takes a contiguous block of the batch; there is no data-path collective.  The only
cross-rank traffic is the barrier and the max-over-ranks of the step time (bench) and, when a
caller wants one PAF stream, a gather of the small result records."""
import numpy as np


def shard_bounds(n_reads, world, rank):
    """Contiguous, balanced [lo, hi) block of `n_reads` for `rank` (first n_reads % world ranks get one more)."""
    base, rem = divmod(int(n_reads), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_files(files, world, rank):
    """fast5 files of a rank when `uncalled map` runs as one process per GPU: file i goes to rank i mod world
    (files are the natural unit: each is opened and decoded by exactly one process)."""
    return [f for i, f in enumerate(files) if i % int(world) == int(rank)]


def channel_owner(channel, world):
    """Streaming path: channel `c` (0-based) lives on rank c mod world -- its persistent device state (detector,
    normaliser, path buffers, seed clusters) never moves, so chunks need no cross-rank exchange either."""
    return int(channel) % int(world)


def local_channels(n_channels, world, rank):
    """The channels a rank owns, and their local (dense) indices on that rank's StreamMapper."""
    chans = list(range(int(rank), int(n_channels), int(world)))
    return chans, {c: i for i, c in enumerate(chans)}


def max_over_ranks(value, dist=None, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_records(local_recs, dist=None):
    """Concatenate per-rank numpy record arrays in rank order on every rank."""
    if dist is None or not dist.is_initialized():
        return local_recs
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local_recs)
    return np.concatenate(out)
