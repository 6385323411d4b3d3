"""Multi-rank host logic on CPU (gloo, world_size 2): streams are sharded with no data-path
collective; the only collective is the end-of-run gather of per-rank statistics, and the
per-stream results must not depend on how streams were sharded."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total_streams, frames, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import posebyte_b200 as pb
    import oracle_py as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = pb.Shard(rank, world, total_streams)
    cfg = pb.synth_config(canvas=640, persons=8, period=16)
    heads = pb.synth_heads(cfg, sh.start, sh.count, 0, frames, frame_major=True, threads=1)
    r = orc.run_streams(heads, True)                 # stands in for the GPU path on this CPU-only box
    stats = torch.zeros(total_streams + 2, dtype=torch.int64)
    stats[sh.start: sh.start + sh.count] = torch.from_numpy(r["hashes"].view(np.int64))
    stats[-2] = r["tracks_total"]; stats[-1] = sh.count * frames
    gathered = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)                 # the ONLY collective: final statistics
    if rank == 0:
        tot = torch.stack(gathered).sum(0)
        q.put(tot.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_is_invariant():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import posebyte_b200 as pb
    import oracle_py as orc
    total, frames = 6, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = pb.synth_config(canvas=640, persons=8, period=16)
    heads = pb.synth_heads(cfg, 0, total, 0, frames, frame_major=True, threads=1)
    ref = orc.run_streams(heads, True)
    assert np.array_equal(got[:total].astype(np.int64), ref["hashes"].view(np.int64))
    assert got[-2] == ref["tracks_total"] and got[-1] == total * frames
