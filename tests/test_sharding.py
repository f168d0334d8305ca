"""Stream sharding across ranks (replicas only) -- world_size 2 over gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cuda_vp9_b200 as vb
from cuda_vp9_b200 import sharding


def test_round_robin_assignment_covers_every_stream_once():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 64):
            owned = [sharding.streams_of_rank(n, r, world) for r in range(world)]
            flat = sorted(s for o in owned for s in o)
            assert flat == list(range(n))
            for r, o in enumerate(owned):
                assert all(sharding.device_for_stream(s, world) == r for s in o)
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    with pytest.raises(ValueError):
        sharding.streams_of_rank(4, 2, 2)


def test_host_threads_split():
    assert sharding.host_threads_for_rank(1, cores=16) == 16
    assert sharding.host_threads_for_rank(8, cores=16) == 2
    assert sharding.host_threads_for_rank(8, cores=4) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, clip, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.dist_env() == (rank, world, rank)
        mine = sharding.streams_of_rank(5, rank, world)
        # every owned stream is parsed independently by this rank (host half of the decoder)
        w, h, chunks = vb.ivf_read(clip)
        frames = 0
        for _ in mine:
            p = vb.HostParser()
            for ch in chunks:
                for fr in vb.parse_superframe(ch):
                    _, show = p.parse(fr)
                    frames += show >= 0
            p.close()
        total, tmax = sharding.combine(frames, 1.0 + rank)
        dist.barrier()
        q.put((rank, mine, frames, total, tmax))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    from conftest import clip_path
    world = 2
    clip = clip_path("odd_rt")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, clip, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    per_stream = res[1][2] // 2
    assert per_stream == 8 and res[0][2] == 3 * per_stream
    for r in res:
        assert r[3] == 5 * per_stream      # frames summed over ranks
        assert r[4] == 2.0                 # time = max over ranks
