#!/usr/bin/env python3
"""bench.py -- decoded frames/sec of the VP9 reconstruction hot path on B200.

A "step" is one pass over the bench clip for every stream of this rank (S independent decoder
contexts per GPU, each on its own CUDA stream -- multi-GPU work shards independent streams,
there is no collective on the data path).

  value     recon-only replay: the per-frame packed buffers already live in HBM; every step
            re-submits them through vp9gpu_frame_submit_device (C ABI) -- device-event timed.
  e2e       whole decoder through the C ABI with HOST buffers: compressed chunk -> host parse ->
            packed buffer H2D -> kernels -> D2H of every shown frame (vp9b_decoder_decode /
            vp9b_decoder_read_frame), one host thread per stream.
  roofline  dominant kernel of the step: algorithmic bytes / device time vs measured HBM copy peak.
  cpu_baseline / --impl reference   libvpx v1.15.2 (x86 SIMD; upstream of the reference fork)
            decoding the same clip on the host cores, one single-thread decoder per core.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# One hardware work queue per decoder context: with the default of 8, streams beyond the 8th alias
# onto the same queue and serialise (profiles/r01_summary.md).  Must be set before CUDA starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS = {"1440p": "bench_1440p", "4k": "bench_4k", "hbd1080": "bench_hbd1080"}


def clip_file(name):
    return os.path.join(ROOT, "tests", "golden", "clips", CLIPS[name] + ".ivf")


def load_clip(name):
    import cuda_vp9_b200 as vb
    path = clip_file(name)
    if not os.path.exists(path):
        raise SystemExit("bench clip %s missing: run `python tools/make_clips.py bench`" % path)
    w, h, chunks = vb.ivf_read(path)
    return w, h, chunks


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md), sampled through NVML
    every few milliseconds (nvidia-smi's own loop is too slow to start for sub-second regions)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def sample(self):
        nv = self.nv
        self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        except Exception:
            mask = 0
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:
                break
            time.sleep(0.004)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        if self.nv is None:   # last resort: one nvidia-smi query
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [int(x) for x in out.strip().split(",")]
                self.sm, self.max_mhz = [a], b
            except Exception:
                pass
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def dist_setup(args):
    from cuda_vp9_b200 import sharding
    return sharding.dist_env()


# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: libvpx v1.15.2 SIMD decoders, one single-thread instance per host core."""
    if rank != 0:
        return
    from oracle import vpx_oracle as vo
    w, h, chunks = load_clip(args.clip)
    cores = os.cpu_count() or 1
    par = cores

    def one_pass(counter, idx):
        d = vo.Decoder(threads=1)
        n = 0
        for ch in chunks:
            n += len(d.decode(ch, want="none"))
        d.close()
        counter[idx] = n

    def step():
        counter = [0] * par
        th = [threading.Thread(target=one_pass, args=(counter, i)) for i in range(par)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return sum(counter)

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        frames += step()
    dt = time.perf_counter() - t0
    fps = frames / dt
    out = {"impl": "reference", "metric": "decoded frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": workload_name(args, w, h), "streams": par, "frames_per_stream": len(chunks)},
           "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                            "sample": "libvpx v1.15.2 (upstream of the reference fork, x86 SIMD, prebuilt in this image): "
                                      "%d parallel single-thread decoders x %d steps over the clip, decode + get_frame, "
                                      "no MD5" % (par, args.steps)},
           "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def workload_name(args, w, h):
    bits = {"hbd1080": "10-bit profile 2"}.get(args.clip, "8-bit")
    return ("synthetic %dx%d %s 4:2:0 VP9 clip (4 tile columns, alt-ref + compound, loop filter on; stand-in for "
            "Bravia.1440.ivf / FoodMarket2.ivf which are absent from the reference checkout)" % (w, h, bits))


# --------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local):
    import numpy as np
    import torch
    import cuda_vp9_b200 as vb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the reconstruction path)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, chunks = load_clip(args.clip)
    S = args.streams
    from cuda_vp9_b200 import sharding
    # host parse threads: all ranks share this box's cores
    E = args.e2e_streams or sharding.host_threads_for_rank(world)
    hdr_size = C.sizeof(vb.FrameHdr)

    # ---- host parse once: packed buffer per coded frame
    parser = vb.HostParser()
    packed, shown = [], 0
    t_parse0 = time.perf_counter()
    for ch in chunks:
        for fr in vb.parse_superframe(ch):
            p, show = parser.parse(fr)
            if p is not None:
                packed.append(p)
            if show >= 0:
                shown += 1
    parse_s = time.perf_counter() - t_parse0
    hdrs = [vb.FrameHdr.from_buffer_copy(p[:hdr_size]) for p in packed]
    n_coded = len(packed)

    # ---- S independent reconstruction contexts, packed buffers resident in HBM
    ctxs = [vb.GpuRecon(local) for _ in range(S)]
    if args.stop_stage:     # debugging aid: marginal cost of each stage under stream concurrency
        for c in ctxs:
            c.set_stop_stage(args.stop_stage)
    dptr = [[c.upload(p)[0] for p in packed] for c in ctxs]
    streams = [torch.cuda.ExternalStream(c.stream()) for c in ctxs]
    master = torch.cuda.Stream()

    K = max(1, min(args.batch, 4))
    if K == 1:
        batches = [ctxs[ci].make_batch([(dptr[ci][f], hdrs[f]) for f in range(n_coded)]) for ci in range(S)]
        n_workers = S

        def submit_all(ci):
            ctxs[ci].submit_batch(batches[ci])      # one C-ABI call queues the whole clip pass
    else:
        # K streams share each launch (the hardware has 32 work queues): groups of K contexts
        groups = [list(range(g, min(g + K, S))) for g in range(0, S, K)]
        packs = [[vb.GpuRecon.make_multi([ctxs[ci] for ci in grp], [(dptr[ci][f], hdrs[f]) for ci in grp])
                  for f in range(n_coded)] for grp in groups]
        n_workers = len(groups)

        def submit_all(gi):
            for pk in packs[gi]:
                vb.GpuRecon.submit_multi(pk)

    def replay_step(threads=True):
        if threads and S > 1:
            th = [threading.Thread(target=submit_all, args=(i,)) for i in range(n_workers)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        else:
            for f in range(n_coded):
                for ci in range(S):
                    ctxs[ci].submit_device(dptr[ci][f], hdrs[f])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        replay_step()
    barrier()
    launches0 = sum(c.launch_count() for c in ctxs)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(master):
        e0.record()
    for s in streams:
        s.wait_event(e0)
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        replay_step()
    host_submit_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        master.wait_event(ev)
    with torch.cuda.stream(master):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.finish()
    launches = sum(c.launch_count() for c in ctxs) - launches0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    frames_total = shown * S * args.steps * world
    value = frames_total / (ms * 1e-3)

    # ---- per-stage device time on one stream (profiling events inside the C ABI) -> roofline
    prof = ctxs[0]
    prof.sync()
    prof.set_profiling(True)
    prof.stage_ms(reset=True)
    reps = 2
    for _ in range(reps):
        for f in range(n_coded):
            prof.submit_device(dptr[0][f], hdrs[f])
    prof.sync()
    stage = prof.stage_ms(reset=True)
    prof.set_profiling(False)
    bps = 2 if hdrs[0].bit_depth > 8 else 1
    mc_bytes = sum((sum(hh.n_inter_px) + sum(hh.n_compound_px)) * bps + sum(hh.n_inter_px) * bps for hh in hdrs)
    lf_frames = [hh for hh in hdrs if hh.lf_level]
    frame_bytes = lambda hh: int(hh.mi_cols * 8 * hh.mi_rows * 8 * 1.5 * bps)  # noqa: E731
    lf_bytes = sum(2 * frame_bytes(hh) for hh in lf_frames)
    stage_names = ["expand", "inter_pred", "inter_residual", "intra", "loop_filter"]
    per_stage = {n: stage[i] / reps for i, n in enumerate(stage_names)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    cand = {"k_inter_pred": (per_stage["inter_pred"], mc_bytes, sum(1 for hh in hdrs if not hh.intra_only) * 3),
            "k_loopfilter": (per_stage["loop_filter"], lf_bytes, len(lf_frames))}
    dom = max(cand, key=lambda k: cand[k][0])
    dms, dbytes, dlaunch = cand[dom]
    achieved = dbytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
    # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get(args.clip, {}).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dbytes / max(dlaunch, 1),
                "avg_launch_ms": dms / max(dlaunch, 1),
                "per_stage_ms_per_clip_pass": per_stage,
                "all": {k: {"ms": v[0], "bytes": v[1], "GBps": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0)}
                        for k, v in cand.items()}}

    # ---- e2e: compressed chunks in host memory -> frames in host memory, one thread per stream
    RING = 6
    res = [0] * E
    t_marks = {}
    ready_bar = threading.Barrier(E + 1)
    start_bar = threading.Barrier(E + 1)
    end_bar = threading.Barrier(E + 1)

    def e2e_worker(i):
        # decoder + pinned output ring are created once (allocation is not part of the steady state)
        d = vb.Decoder(local)
        ring = [vb.PinnedFrame(w, h, hdrs[0].bit_depth, hdrs[0].ss_x, hdrs[0].ss_y) for _ in range(RING)]
        tickets = [0] * RING

        def run(passes):
            n = 0
            for _ in range(passes):
                # a new pass restarts at the key frame, so the same decoder instance can be reused
                for ch in chunks:
                    d.decode(ch)                            # host parse + H2D + kernels (async)
                    k = n % RING
                    if tickets[k]:
                        d.wait(tickets[k])                  # the frame that used this pinned buffer RING frames ago
                        tickets[k] = 0                      # is complete in host memory: hand it to the consumer
                    if d.get_frame_into(ring[k]):           # D2H of the shown frame into pinned memory (async)
                        tickets[k] = d.mark()
                        n += 1
            d.sync()                                        # every frame is in host memory
            for k in range(RING):
                tickets[k] = 0
            return n

        run(1)                  # warm-up: sizes every pool
        ready_bar.wait()
        start_bar.wait()
        res[i] = run(args.steps)
        end_bar.wait()
        d.close()
        for r in ring:
            r.close()

    th = [threading.Thread(target=e2e_worker, args=(i,)) for i in range(E)]
    for t in th:
        t.start()
    ready_bar.wait()
    barrier()                   # every rank has warmed up: start the timed region together
    start_bar.wait()
    t0 = time.perf_counter()
    end_bar.wait()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for t in th:
        t.join()
    n_e2e = sum(res)
    n_e2e, e2e_s = sharding.combine(n_e2e, e2e_s, device="cuda")
    e2e_fps = n_e2e / e2e_s
    h2d = sum(len(p) for p in packed) * E * world
    d2h = shown * int(w * h * 1.5 * bps) * E * world

    out = {"metric": "decoded frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8" if bps == 1 else "u16", "data": "synthetic",
           "config": {"workload": workload_name(args, w, h), "streams_per_gpu": S, "coded_frames": n_coded,
                      "shown_frames": shown, "parallelism": "independent streams, %d per GPU, %d per kernel launch, no collective" % (S, K),
                      "l2": "working set %d MB per GPU (> 126 MB L2): %d streams x (frame pool + packed buffers)" %
                            (S * (12 * frame_bytes(hdrs[0]) + sum(len(p) for p in packed)) >> 20, S),
                      "host_parse_fps_1thread": shown / parse_s,
                      "host_submit_ms_per_step": host_submit_ms},
           "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
           "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "what": "vp9b_decoder_decode (host parse + H2D + kernels) + vp9b_decoder_read_frame (D2H) per shown "
                           "frame into pinned host memory, %d host threads / streams" % E}}

    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(chunks)
    for c in ctxs:
        c.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(chunks, budget_s=12.0):
    """libvpx v1.15.2 SIMD on the host cores: one single-thread decoder per core, bounded sample."""
    from oracle import vpx_oracle as vo
    cores = os.cpu_count() or 1
    stop_at = time.perf_counter() + budget_s
    counts = [0] * cores

    def work(i):
        while time.perf_counter() < stop_at:
            d = vo.Decoder(threads=1)
            for ch in chunks:
                counts[i] += len(d.decode(ch, want="none"))
            d.close()

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return {"value": sum(counts) / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "libvpx v1.15.2 (upstream of the reference fork, x86 SIMD build shipped in this image), %d parallel "
                      "single-thread decoders looping over the bench clip for %.0f s; decode + get_frame, no MD5"
                      % (cores, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="independent decoder contexts per GPU")
    ap.add_argument("--e2e-streams", type=int, default=0, help="host threads for the e2e leg (default: host cores)")
    ap.add_argument("--clip", default="1440p", choices=sorted(CLIPS))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--stop-stage", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--batch", type=int, default=2, help="streams per kernel launch (1..4)")
    args = ap.parse_args()
    rank, world, local = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
