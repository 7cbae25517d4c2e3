"""CPU tier, world_size 2 over gloo: the N>1 host logic of bench.py / MapPool sharding --
contiguous read blocks per rank, max-over-ranks timing, rank-ordered gather of PAF records.
The mapping itself is replaced by the oracle here (this test is about the sharding plumbing;
GPU parity is covered by tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import orclib, synth, synthdata
    from uncalled_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 11, 2000, seed=4)
    lo, hi = shard.shard_bounds(len(sig), world, rank)
    O = orclib.Oracle(prefix)
    recs = np.array([orclib.paf_tuple(O.map_read(sig[i]))[:4] + (i,) for i in range(lo, hi)], dtype=np.int64)
    allrecs = shard.gather_records(recs, dist)
    t = shard.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        q.put((allrecs.tolist(), t, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_sharding():
    from uncalled_b200 import shard
    assert [shard.shard_bounds(11, 2, r) for r in range(2)] == [(0, 6), (6, 11)]
    assert [shard.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard.shard_bounds(0, 3, 1) == (0, 0)
    # streaming: channel c lives on rank c mod world; every channel has exactly one owner and a dense local index
    owners = [shard.channel_owner(c, 8) for c in range(512)]
    assert owners[:10] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1] and all(owners.count(r) == 64 for r in range(8))
    chans, local = shard.local_channels(512, 8, 3)
    assert chans[:3] == [3, 11, 19] and local[19] == 2 and len(chans) == 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    allrecs, t, (lo, hi) = q.get(timeout=300)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0                                   # max over ranks
    assert [r[-1] for r in allrecs] == list(range(11))  # rank-ordered gather keeps read order


def test_fast5_files_are_dealt_round_robin():
    from uncalled_b200.shard import local_files
    files = ["f%d.fast5" % i for i in range(11)]
    parts = [local_files(files, 4, r) for r in range(4)]
    assert sorted(sum(parts, [])) == sorted(files) and [len(p) for p in parts] == [3, 3, 3, 2]
    assert parts[1] == ["f1.fast5", "f5.fast5", "f9.fast5"] and local_files(files, 1, 0) == files
