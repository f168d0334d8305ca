"""Multi-GPU layout: replicas only.

Frames of one VP9 stream depend serially on their reference frames
(swap_frame_buffers, vp9/decoder/vp9_decoder.c:334-367), streams are independent, so work is
sharded by stream: stream i runs on GPU (i mod G) with its own decoder context.  Nothing is
exchanged on the data path; torch.distributed is used only for the barrier around a timed region
and to combine per-rank results (frames: sum, time: max over ranks).
"""
import os


def dist_env():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def device_for_stream(stream_id, n_gpus):
    return stream_id % n_gpus


def streams_of_rank(n_streams_total, rank, world):
    """Global stream ids owned by `rank` (round-robin, stream_i -> GPU i mod G)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return list(range(rank, n_streams_total, world))


def host_threads_for_rank(world, cores=None):
    """Host parse threads one rank may use when all ranks share one box."""
    cores = cores or os.cpu_count() or 1
    return max(1, cores // max(world, 1))


def combine(frames_local, seconds_local, device=None):
    """Whole-job (frames, seconds): frames summed over ranks, time = max over ranks.
    Works with any initialised backend (nccl needs `device`, gloo uses CPU tensors)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(frames_local), float(seconds_local)
    f = torch.tensor([int(frames_local)], dtype=torch.int64, device=device)
    t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(f.item()), float(t.item())
