#!/usr/bin/env python3
"""bench.py -- decoded frames/sec of the VP9 reconstruction hot path on B200.

Workload: the 4K bench clip (BASELINE.json configs[3]: 3840x2160 8-bit, 4 tile columns, hidden
alt-ref + compound prediction, scaled references, loop filter on), S independent streams per GPU.
A "step" is `passes_per_step` passes over the clip for every stream of this rank; multi-GPU work
shards independent streams, there is no collective on the data path.

  value     recon-only replay: the per-frame packed buffers already live in HBM; every step
            re-submits them through vp9gpu_frame_submit_device_multi (C ABI), K streams per
            launch -- device-event timed.
  e2e       whole decoder through the C ABI with HOST buffers: compressed chunk -> host parse ->
            packed buffer H2D -> kernels -> D2H of every shown frame (vp9b_decoder_decode /
            vp9b_decoder_read_frame), one host thread per stream.
  roofline  dominant kernel of the step (chosen among ALL stages): algorithmic bytes of one launch
            (a batch of K frames) / its device time, measured with CUDA events inside the C ABI on
            the launching stream, vs the measured HBM copy peak.
  other_workloads   the same replay measurement for the 1440p, 1080p 10-bit and lossless clips.
  cpu_baseline / --impl reference   the reference tree's own C decoder (oracle/_ref/vpxdec_ref,
            built from /root/reference/libvpx by tools/build_ref_vpxdec.py, VP9B_SHIM=cpu), one
            process per host core; libvpx v1.15.2 with x86 SIMD is timed next to it
            (cpu_baseline_simd) as the harder yardstick.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# One hardware work queue per launch group.  Must be set before CUDA starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS = {"4k": "bench_4k", "1440p": "bench_1440p", "hbd1080": "bench_hbd1080", "lossless": "bench_lossless"}
WORKLOADS = {
    "4k": "synthetic 3840x2160 8-bit 4:2:0 VP9 clip: two-pass alt-ref segment (hidden ARF + compound prediction) spliced with a "
          "one-pass segment whose coded size changes 4K -> 4/5 -> 3/5 -> 4K (scaled references), 4 tile columns, loop filter on "
          "(BASELINE.json configs[3]; stand-in for FoodMarket2.ivf, absent from the reference checkout)",
    "1440p": "synthetic 2560x1440 8-bit 4:2:0 VP9 clip, 4 tile columns, alt-ref + compound, loop filter on (stand-in for "
             "Bravia.1440.ivf, absent from the reference checkout)",
    "hbd1080": "synthetic 1920x1080 10-bit profile 2 VP9 clip, 4 tile columns, alt-ref + compound, loop filter on",
    "lossless": "synthetic 1920x1080 8-bit lossless VP9 clip (4x4 WHT for every block), 4 tile columns",
}
VPXDEC_REF = os.path.join(ROOT, "oracle", "_ref", "vpxdec_ref")


def clip_file(name):
    return os.path.join(ROOT, "tests", "golden", "clips", CLIPS[name] + ".ivf")


def load_clip(name):
    import cuda_vp9_b200 as vb
    path = clip_file(name)
    if not os.path.exists(path):
        raise SystemExit("bench clip %s missing: run `python tools/make_clips.py bench`" % path)
    return vb.ivf_read(path)


def make_config(args):
    """The same dict in both arms (the driver compares them key by key)."""
    md5 = os.path.join(ROOT, "tests", "golden", "md5", CLIPS[args.clip] + ".md5")
    return {"workload": WORKLOADS[args.clip], "clip": CLIPS[args.clip], "shown_frames": sum(1 for _ in open(md5)),
            "streams_per_gpu": args.streams, "streams_per_launch": max(1, min(args.batch, args.streams, 32)),
            "passes_per_step": args.passes}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md), sampled through NVML
    every few milliseconds."""

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


# --------------------------------------------------------------------------------------------
# CPU arms
def time_vpxdec_ref(clip, procs, budget_s=None, passes=1):
    """The reference tree's C decoder (VP9B_SHIM=cpu): `procs` parallel processes, each decoding the
    clip once per pass without writing frames.  Returns (frames, seconds)."""
    shown = sum(1 for _ in open(os.path.join(ROOT, "tests", "golden", "md5", CLIPS[clip] + ".md5")))
    env = dict(os.environ, VP9B_SHIM="cpu")
    cmd = [VPXDEC_REF, "--noblit", clip_file(clip)]
    t0 = time.perf_counter()
    frames = 0
    n = 0
    while True:
        ps = [subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(procs)]
        bad = [p.wait() for p in ps]
        if any(bad):
            raise RuntimeError("vpxdec_ref failed (rc %r)" % bad)
        frames += shown * procs
        n += 1
        if budget_s is None:
            if n >= passes:
                break
        elif time.perf_counter() - t0 >= budget_s:
            break
    return frames, time.perf_counter() - t0


def time_libvpx_simd(chunks, threads, budget_s):
    """libvpx v1.15.2 (x86 SIMD build inside this image's OpenCV wheel): one single-thread decoder
    per host core looping over the clip."""
    from oracle import vpx_oracle as vo
    stop_at = time.perf_counter() + budget_s
    counts = [0] * threads

    def work(i):
        while time.perf_counter() < stop_at:
            d = vo.Decoder(threads=1)
            for ch in chunks:
                counts[i] += len(d.decode(ch, want="none"))
            d.close()

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return sum(counts), time.perf_counter() - t0


def ref_usable(clip):
    return os.path.exists(VPXDEC_REF) and clip != "hbd1080"    # the reference tree builds without high bit depth


def cpu_baselines(clip, chunks, budget_s):
    cores = os.cpu_count() or 1
    out = {}
    if ref_usable(clip):
        f, s = time_vpxdec_ref(clip, cores, budget_s=budget_s)
        out["cpu_baseline"] = {"value": f / s, "unit": "frames/s", "cores": cores, "kind": "reference",
                               "sample": "oracle/_ref/vpxdec_ref (the reference tree's libvpx, portable C, VP9B_SHIM=cpu) --noblit: "
                                         "%d parallel processes x %d passes over the clip (%.0f s)" % (cores, f // (cores * len(chunks)), s)}
    f, s = time_libvpx_simd(chunks, cores, min(budget_s, 8.0))
    simd = {"value": f / s, "unit": "frames/s", "cores": cores, "kind": "libvpx v1.15.2 x86 SIMD (prebuilt in this image; not the reference tree)",
            "sample": "%d parallel single-thread decoders looping over the clip for %.0f s; decode + get_frame, no MD5" % (cores, s)}
    if "cpu_baseline" in out:
        out["cpu_baseline_simd"] = simd
    else:
        simd["kind"] = "port"
        out["cpu_baseline"] = simd
    return out


def run_reference(args, rank, world):
    """CPU arm: the reference tree's own decoder on all host cores, on the same clip."""
    if rank != 0:
        return
    _, _, chunks = load_clip(args.clip)
    cores = os.cpu_count() or 1
    shown = len(chunks)
    if ref_usable(args.clip):
        for _ in range(min(args.warmup, 1)):
            time_vpxdec_ref(args.clip, cores)
        frames, dt = 0, 0.0
        for _ in range(args.steps):
            f, s = time_vpxdec_ref(args.clip, cores)
            frames += f
            dt += s
        kind = "reference"
        sample = ("oracle/_ref/vpxdec_ref (reference tree's libvpx, portable C, VP9B_SHIM=cpu) --noblit: each step = %d parallel "
                  "processes x 1 pass over the clip" % cores)
    else:
        frames, dt = time_libvpx_simd(chunks, cores, max(3.0, 0.5 * args.steps))
        kind = "port"
        sample = "libvpx v1.15.2 x86 SIMD, %d parallel single-thread decoders (reference tree not buildable for this clip)" % cores
    fps = frames / dt
    out = {"impl": "reference", "metric": "decoded frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8" if args.clip != "hbd1080" else "u16", "data": "synthetic",
           "config": make_config(args),
           "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference" and not args.no_cpu:
        f, s = time_libvpx_simd(chunks, cores, 6.0)
        out["cpu_baseline_simd"] = {"value": f / s, "unit": "frames/s", "cores": cores,
                                    "kind": "libvpx v1.15.2 x86 SIMD (prebuilt in this image; not the reference tree)"}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------
def parse_clip(vb, chunks):
    parser = vb.HostParser()
    packed, shown = [], 0
    t0 = time.perf_counter()
    for ch in chunks:
        for fr in vb.parse_superframe(ch):
            p, show = parser.parse(fr)
            if p is not None:
                packed.append(p)
            if show >= 0:
                shown += 1
    return packed, shown, time.perf_counter() - t0


def stage_bytes(np, packed, hdrs):
    """Algorithmic bytes per stage, summed over the clip: what each kernel has
    to read and write at least once."""
    tot = {"expand": 0, "inter_pred": 0, "inter_residual": 0, "intra_residual": 0, "intra": 0, "loop_filter": 0}
    per_frame = []
    for p, h in zip(packed, hdrs):
        bps = 2 if h.bit_depth > 8 else 1
        cb, rb = (4, 4) if bps == 2 else (2, 2)
        raw = np.frombuffer(p, dtype=np.uint8)
        f = dict.fromkeys(tot, 0)

        def txb(sec):
            a = raw[h.sec[sec].offset:h.sec[sec].offset + h.sec[sec].count * 12].reshape(-1, 12)
            eob = a[:, 8].astype(np.int64) | (a[:, 9].astype(np.int64) << 8)
            bs = 4 << (a[:, 10] & 3).astype(np.int64)
            return eob, bs
        if not h.intra_only:
            f["expand"] = 16 * h.sec[0].count + 16 * h.mi_cols * h.mi_rows
            f["inter_pred"] = (sum(h.n_inter_px) + sum(h.n_compound_px)) * bps + sum(h.n_inter_px) * bps
            eob, bs = txb(2)
            f["inter_residual"] = int(cb * eob.sum() + 12 * len(eob) + 2 * bps * (bs * bs).sum())
        for sec in (3, 4, 5):
            eob, bs = txb(sec)
            coded = eob > 0
            f["intra_residual"] += int(cb * eob.sum() + 12 * len(eob) + rb * (bs * bs)[coded].sum())
            f["intra"] += int(12 * len(eob) + bps * (bs * bs).sum() + rb * (bs * bs)[coded].sum() + bps * (2 * bs + 1).sum())
        if h.lf_level:
            f["loop_filter"] = int(2 * h.mi_cols * 8 * h.mi_rows * 8 * 1.5 * bps) + 160 * h.sb_cols * h.sb_rows
        per_frame.append(f)
        for k in tot:
            tot[k] += f[k]
    return tot, per_frame


class Replay:
    """S reconstruction contexts replaying one clip's packed buffers from HBM, K streams per launch."""

    def __init__(self, vb, torch, local, clip, S, K, stop_stage=0, parsed=None):
        import numpy as np
        self.vb, self.torch, self.S = vb, torch, S
        if parsed is None:
            w, h, chunks = load_clip(clip)
            parsed = (w, h, chunks) + parse_clip(vb, chunks)
        self.w, self.h, self.chunks, self.packed, self.shown, self.parse_s = parsed
        hs = C.sizeof(vb.FrameHdr)
        self.hdrs = [vb.FrameHdr.from_buffer_copy(p[:hs]) for p in self.packed]
        self.n_coded = len(self.packed)
        self.ctxs = [vb.GpuRecon(local) for _ in range(S)]
        if stop_stage:
            for c in self.ctxs:
                c.set_stop_stage(stop_stage)
        self.dptr = [[c.upload(p)[0] for p in self.packed] for c in self.ctxs]
        self.streams = [torch.cuda.ExternalStream(c.stream()) for c in self.ctxs]
        K = max(1, min(K, S, vb.MAX_BATCH))
        self.K = K
        self.groups = [list(range(g, min(g + K, S))) for g in range(0, S, K)]
        self.packs = [[vb.GpuRecon.make_multi([self.ctxs[ci] for ci in grp], [(self.dptr[ci][f], self.hdrs[f]) for ci in grp])
                       for f in range(self.n_coded)] for grp in self.groups]
        self.bytes_tot, self.bytes_frame = stage_bytes(np, self.packed, self.hdrs)

    def submit_pass(self, gi):
        for pk in self.packs[gi]:
            self.vb.GpuRecon.submit_multi(pk)

    def step(self, passes):
        def work(gi):
            for _ in range(passes):
                self.submit_pass(gi)
        if len(self.groups) == 1:
            work(0)
            return
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(self.groups))]
        for t in th:
            t.start()
        for t in th:
            t.join()

    def launches(self):
        return sum(c.launch_count() for c in self.ctxs)

    def timed(self, steps, warmup, passes, barrier, sampler=None):
        torch = self.torch
        master = torch.cuda.Stream()
        for _ in range(warmup):
            self.step(passes)
        barrier()
        l0 = self.launches()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(master):
            e0.record()
        for s in self.streams:
            s.wait_event(e0)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step(passes)
        host_ms = (time.perf_counter() - t0) * 1e3 / steps
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            master.wait_event(ev)
        with torch.cuda.stream(master):
            e1.record()
        barrier()
        return e0.elapsed_time(e1), host_ms, self.launches() - l0

    def profile_batch(self, reps=2):
        """Per-stage device time of one launch group running alone (events inside the C ABI)."""
        lead = self.ctxs[self.groups[0][0]]
        self.torch.cuda.synchronize()
        lead.set_profiling(True)
        lead.stage_ms(reset=True)
        for _ in range(reps):
            self.submit_pass(0)
        self.torch.cuda.synchronize()
        ms = lead.stage_ms(reset=True)
        lead.set_profiling(False)
        names = ["expand", "inter_pred", "inter_residual", "intra_residual", "intra", "loop_filter"]
        return {n: ms[i] / reps for i, n in enumerate(names)}

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []


def run_ours(args, rank, world, local):
    import torch
    import cuda_vp9_b200 as vb
    from cuda_vp9_b200 import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the reconstruction path)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    S, P = args.streams, args.passes
    # ---- the clip, parsed once on the host
    w, h, chunks = load_clip(args.clip)
    parsed = (w, h, chunks) + parse_clip(vb, chunks)
    packed_all, shown = parsed[3], parsed[4]
    hs = C.sizeof(vb.FrameHdr)
    hdrs_all = [vb.FrameHdr.from_buffer_copy(p[:hs]) for p in packed_all]
    hdr0 = hdrs_all[0]
    bps = 2 if hdr0.bit_depth > 8 else 1
    packed_bytes = sum(len(p) for p in packed_all)
    # bytes of the frames a pass hands out (frames of the scaled segment are smaller than the stream size)
    shown_bytes = sum(int(hh.width * hh.height + 2 * ((hh.width + hh.ss_x) >> hh.ss_x) * ((hh.height + hh.ss_y) >> hh.ss_y)) * bps
                      for hh in hdrs_all if hh.show_frame)

    # ---- e2e first (the host is the bottleneck of this leg: measure it before the long replay phases):
    # compressed chunks in host memory -> frames in host memory, one thread per stream
    e2e = None
    if not args.no_e2e:
        E = args.e2e_streams or sharding.host_threads_for_rank(world)
        RING = 6
        res = [0] * E
        ready_bar = threading.Barrier(E + 1)
        start_bar = threading.Barrier(E + 1)
        end_bar = threading.Barrier(E + 1)
        e2e_passes = max(1, args.steps * P // 2)

        def e2e_worker(i):
            # decoder + pinned output ring are created once (allocation is not part of the steady state)
            d = vb.Decoder(local)
            ring = [vb.PinnedFrame(w, h, hdr0.bit_depth, hdr0.ss_x, hdr0.ss_y) for _ in range(RING)]
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
            res[i] = run(e2e_passes)
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
        n_e2e, e2e_s = sharding.combine(sum(res), e2e_s, device="cuda")

        # one stream alone (latency view): the whole decoder with 1 and with 4 tile-column parser threads
        def single_stream(threads, passes=2):
            d = vb.Decoder(local)
            d.set_parse_threads(threads)
            ring = [vb.PinnedFrame(w, h, hdr0.bit_depth, hdr0.ss_x, hdr0.ss_y) for _ in range(2)]
            n, t0 = 0, 0.0
            for ps in range(passes + 1):
                if ps == 1:
                    d.sync()
                    t0, n = time.perf_counter(), 0
                for ch in chunks:
                    d.decode(ch)
                    if d.get_frame_into(ring[n & 1], sync=True):
                        n += 1
            dt = time.perf_counter() - t0
            d.close()
            for r in ring:
                r.close()
            return n / dt
        single = {"parse_threads_1": single_stream(1), "parse_threads_4": single_stream(4)} if rank == 0 else None
        e2e = {"value": n_e2e / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": int(packed_bytes * E * world * e2e_passes / args.steps),
               "d2h_bytes_per_step": int(shown_bytes * E * world * e2e_passes / args.steps),
               "timed_region_s": e2e_s, "host_threads": E * world, "single_stream_fps": single,
               "what": "vp9b_decoder_decode (host parse + H2D + kernels) + vp9b_decoder_read_frame (D2H) per shown "
                       "frame into pinned host memory, %d host threads / streams per rank, %d clip passes each" % (E, e2e_passes)}

    # ---- recon-only replay
    rp = Replay(vb, torch, local, args.clip, S, args.batch, args.stop_stage, parsed=parsed)
    args.batch = rp.K
    sampler = ClockSampler(local)
    ms, host_submit_ms, launches = rp.timed(args.steps, args.warmup, P, barrier, sampler)
    clocks = sampler.finish()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    frames_total = rp.shown * S * P * args.steps * world
    value = frames_total / (ms * 1e-3)

    # ---- roofline: one launch group (K frames per launch) alone on the GPU, per-stage events
    stage = rp.profile_batch()
    K = rp.K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernel_of = {"expand": "k_expand_blocks", "inter_pred": "k_inter_pred_sb", "inter_residual": "k_residual",
                 "intra_residual": "k_intra_residual", "intra": "k_intra", "loop_filter": "k_lf_sb"}
    n_launch = {k: max(1, sum(1 for f in rp.bytes_frame if f[k] > 0)) for k in kernel_of}
    allk = {}
    for k, name in kernel_of.items():
        b = rp.bytes_tot[k] * K            # every launch covers K frames
        allk[name] = {"ms_per_clip_pass": stage[k], "bytes_per_clip_pass": b, "launches_per_clip_pass": n_launch[k],
                      "GBps": (b / (stage[k] * 1e-3) / 1e9 if stage[k] > 0 else 0.0)}
    dom_stage = max(kernel_of, key=lambda k: stage[k])
    dom = kernel_of[dom_stage]
    d = allk[dom]
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tr.get(CLIPS[args.clip], {}).get(dom)
        if ent and ent.get("streams_per_launch") == K:
            traffic, traffic_src = ent["dram_bytes_per_launch"], ent.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": d["GBps"], "peak": peak, "unit": "GB/s",
                "frac": d["GBps"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "frames_per_launch": K,
                "algorithmic_bytes_per_launch": d["bytes_per_clip_pass"] / d["launches_per_clip_pass"],
                "avg_launch_ms": d["ms_per_clip_pass"] / d["launches_per_clip_pass"],
                "how": "one launch group of %d streams alone on the GPU, CUDA events between the stages on the launching stream" % K,
                "step_share": {kernel_of[k]: stage[k] / max(sum(stage.values()), 1e-9) for k in kernel_of},
                "all": allk,
                "aggregate_mc_lf_GBps_in_timed_region": (rp.bytes_tot["inter_pred"] + rp.bytes_tot["loop_filter"]) * S * P * args.steps
                                                        / (ms * 1e-3) / 1e9}
    config = make_config(args)
    frame_bytes = int(rp.hdrs[0].mi_cols * 8 * rp.hdrs[0].mi_rows * 8 * 1.5 * bps)
    run_info = {"coded_frames": rp.n_coded,
                "parallelism": "independent streams, %d per GPU, %d per kernel launch, no collective" % (S, K),
                "l2": ("working set %d MB per GPU (> 126 MB L2): %d streams x (frame pool + packed buffers)" %
                       (S * (12 * frame_bytes + sum(len(p) for p in rp.packed)) >> 20, S)),
                "host_parse_fps_1thread": rp.shown / rp.parse_s, "host_submit_ms_per_step": host_submit_ms,
                "timed_region_s": ms * 1e-3}
    rp.close()
    del rp

    # ---- the other bench clips: same replay measurement, short
    others = {}
    if not args.no_others and world == 1:
        for clip in CLIPS:
            if clip == args.clip or not os.path.exists(clip_file(clip)):
                continue
            So = min(S, 8) if clip == "lossless" else min(S, 32)
            r2 = Replay(vb, torch, local, clip, So, args.batch)
            ms2, _, _ = r2.timed(3, 2, 1, barrier)
            st2 = r2.profile_batch(1)
            others[clip] = {"workload": WORKLOADS[clip], "value": r2.shown * So * 3 / (ms2 * 1e-3), "unit": "frames/s",
                            "streams_per_gpu": So, "streams_per_launch": r2.K, "timed_region_s": ms2 * 1e-3,
                            "stage_ms_per_clip_pass": st2}
            r2.close()
            del r2

    out = {"metric": "decoded frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8" if bps == 1 else "u16", "data": "synthetic", "config": config, "run": run_info,
           "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "other_workloads": others}
    if rank == 0 and world == 1 and not args.no_cpu:
        out.update(cpu_baselines(args.clip, chunks, 12.0))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="independent decoder contexts per GPU")
    ap.add_argument("--e2e-streams", type=int, default=0, help="host threads for the e2e leg (default: host cores)")
    ap.add_argument("--clip", default="4k", choices=sorted(CLIPS))
    ap.add_argument("--passes", type=int, default=2, help="clip passes per step (sizes the timed region: >= 3 s at the defaults)")
    ap.add_argument("--batch", type=int, default=32, help="streams per kernel launch")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--stop-stage", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    from cuda_vp9_b200 import sharding
    rank, world, local = sharding.dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
