"""N > 1 host logic on CPU: world_size-2 gloo run of the clip sharding + timing reduction that
bench.py and the multi-GPU configs use (no data-path collective exists to test)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clip_shard_partitions():
    from whisper_context_biasing_b200.sharding import clip_shard

    for n in (0, 1, 7, 128, 256, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [clip_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert clip_shard(1024, 3, 8) == (384, 512)        # BASELINE configs[2]: 1024 / G contiguous clips
    with pytest.raises(ValueError):
        clip_shard(8, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from whisper_context_biasing_b200.sharding import clip_shard, gather_timings

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    start, stop = clip_shard(1024, rank, world)
    # every rank "processes" its shard: checksum of clip indices stands in for per-clip results
    local = torch.arange(start, stop, dtype=torch.int64)
    gathered = [torch.zeros(stop - start, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, local)                       # report-only gather, not a data-path step
    slowest = gather_timings(10.0 + rank, dist)            # max over ranks
    dist.barrier()
    q.put((rank, start, stop, torch.cat(gathered).tolist() == list(range(1024)), slowest))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1:3] == (0, 512) and out[1][1:3] == (512, 1024)
    assert all(o[3] for o in out)
    assert all(o[4] == 11.0 for o in out)
