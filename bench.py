#!/usr/bin/env python
"""bench.py -- DVB-T2 modulator hot path on B200: T2 baseband MS/s (x real-time) and FECFRAMEs/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the fused chain (TS bytes -> complex baseband) over one batch of synthetic
transport streams: BASELINE.json config 5 = 64 independent 32K / 256QAM-rotated / CR 2/3 channels
(config 3), one T2 frame per channel per step, per GPU (weak scaling: every rank processes its own 64
channels, seeds 0x12345678 + global channel index; no data-path collective).

value      = whole-job output Msamples/s with the TS already resident in HBM (CUDA events, max over ranks)
e2e        = same metric through dvbt2ll_chain_run_host(): pinned HOST TS in, HOST samples out, copies timed
roofline   = the dominant kernel (k_ofdm: carrier fill + IFFT + GI) against the measured HBM copy peak
cpu_baseline / --impl reference = the UNMODIFIED reference flowgraph (oracle/_ref, GNU Radio shim) on host cores

PyTorch is only plumbing here (device buffers, streams/events, torch.distributed); the hot path is
libdvbt2ll_cuda.so called through its C ABI.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-dvbt2ll_b200", "python"))

import numpy as np  # noqa: E402

METRIC = "t2_baseband_msps"
UNIT = "Msamples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", help="per-channel configuration (c1..c4)")
    ap.add_argument("--channels", type=int, default=64, help="independent channels per GPU per step")
    ap.add_argument("--frames", type=int, default=1, help="T2 frames per channel per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=24, help="T2 frames timed for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(args):
    return "c5: %d x %s channels (32K ext, 256QAM rot, CR 2/3, GI 1/128, PP7, 202 FECFRAMEs/T2 frame), %d T2 frame/channel/step per GPU" % (
        args.channels, args.config, args.frames) if args.config == "c3" else "%d x %s channels, %d T2 frame/channel/step per GPU" % (
        args.channels, args.config, args.frames)


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm: unmodified reference flowgraph (oracle/_ref) on host cores
# --------------------------------------------------------------------------------------------------
def _ref_worker(job):
    cfg, seed, nframes = job
    from oracle import ref
    from dvbt2ll_b200 import configs as K
    ref.lib().ref_set_quiet(1)
    ref.lib().ref_set_fft_fast(1)      # float Stockham FFT stand-in for FFTW (oracle/shim/gnuradio/fft/fft.h)
    ch = ref.Chain(cfg)
    ts = K.make_ts((nframes + 1) * ch.ts_bytes_per_t2_frame() + 1000, seed=seed)
    timers = {}
    t0 = time.perf_counter()
    n = 0
    for _ in range(nframes):
        n += ch.run_frame(ts, timers=timers)["samples"].size
    return n, time.perf_counter() - t0, timers


def cpu_reference_run(cfg, nframes_total, procs):
    """Times the reference chain on `procs` host processes (independent channels). Returns dict."""
    from oracle import ref
    if not ref.available():
        return None
    from dvbt2ll_b200 import configs as K
    per = max(1, nframes_total // procs)
    jobs = [(cfg, K.TS_SEED + i, per) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_ref_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    samples = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    stage = {}
    for r in res:
        for k, v in r[2].items():
            stage[k] = stage.get(k, 0.0) + v
    return {"samples": samples, "seconds": busy, "wall": wall, "frames": per * procs,
            "stage_seconds": {k: round(v, 4) for k, v in stage.items()}}


def run_reference_arm(args):
    from dvbt2ll_b200 import configs as K
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = K.resolve(args.config)
    procs = max(1, min(os.cpu_count() or 1, 64))
    frames_per_step = procs          # one T2 frame per process per step (bounded sample)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_run(cfg, procs, procs)
    tot_s, tot_t = 0, 0.0
    stage = {}
    for _ in range(args.steps):
        r = cpu_reference_run(cfg, frames_per_step, procs)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdvbt2ll_ref.so not built"}))
            return
        tot_s += r["samples"]; tot_t += r["seconds"]
        for k, v in r["stage_seconds"].items():
            stage[k] = stage.get(k, 0.0) + v
    value = tot_s / tot_t / 1e6
    F = cfg["fecblocks"]
    ch_samples = r["samples"] / r["frames"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": "%d T2 frames of %s per step on %d processes" % (frames_per_step, args.config, procs)},
        "x_realtime": value / K.REALTIME_MSPS, "fecframes_per_s": value * 1e6 / ch_samples * F,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference",
                         "sample": "%d T2 frames (%s) per step x %d steps, one chain per process, unmodified reference sources + GNU Radio shim, "
                                   "single-precision Stockham FFT stand-in for FFTW; stage CPU-seconds %s" % (
                                       frames_per_step, args.config, args.steps, json.dumps({k: round(v, 2) for k, v in stage.items()}))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(torch, local):
    """Run this rank on the CPU cores NVML reports as local to its GPU, so the pinned host buffers of the e2e path
    are first-touched on the GPU's NUMA node.  Returns (cores bound, original affinity) -- (0, None) if unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(local).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        orig = os.sched_getaffinity(0)
        cpus = [i for i in range(words * 64) if (mask[i // 64] >> (i % 64)) & 1 and i in orig]
        if cpus and len(cpus) < len(orig):
            os.sched_setaffinity(0, cpus)
            return len(cpus), orig
    except Exception:
        pass
    return 0, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import dvbt2ll_b200 as T
    from dvbt2ll_b200 import configs as K

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores, orig_affinity = bind_to_gpu_cpus(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = K.resolve(args.config)
    cell_size = (64800 if cfg["framesize"] else 16200) // (2 * (cfg["constellation"] + 1))
    nch, nfr = args.channels, args.frames
    frames = nch * nfr
    chain = T.Chain(cfg, max_frames=frames, device=local)
    n_ts, S, F = chain.ts_bytes_per_frame, chain.samples_per_frame, chain.fecframes_per_frame
    pitch = (nfr * n_ts + 255) // 256 * 256

    # synthetic TS: one independent stream per channel (seed + global channel index), pinned on the host
    ts_host = torch.empty((nch, pitch), dtype=torch.uint8).pin_memory()
    ts_np = ts_host.numpy()
    for c in range(nch):
        ts_np[c, :nfr * n_ts] = K.make_ts(nfr * n_ts, seed=K.TS_SEED + rank * nch + c)
    d_ts = ts_host.to(dev)
    d_out = torch.empty((nch, nfr * S), dtype=torch.complex64, device=dev)
    # a dedicated (non-default) stream: its handle is what the C ABI launches on and what the CUDA events time
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def step():
        chain.run_device(d_ts.data_ptr(), pitch, nch, nfr, 0, d_out.data_ptr(), sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from before the warm-up until after the timed region (nvidia-smi needs ~0.1 s to
    # start; the timed region itself can be shorter than one sampling period)
    clocks = ClockSampler(local)
    clocks.start()
    chain.enable_timing(False)
    t_w = time.perf_counter()
    n_w = 0
    while n_w < max(3, args.warmup) or time.perf_counter() - t_w < 0.4:
        step()
        n_w += 1
        if n_w % 8 == 0:
            torch.cuda.synchronize()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks
    chain.enable_timing(True)
    launches0 = T.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    stage_n = min(args.steps, 64)
    stage_acc = {k: v * stage_n for k, v in chain.stage_ms().items()}    # per-run CUDA events, read after the timed region
    launches = T.kernel_launches() - launches0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    chain.enable_timing(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    total_samples = frames * S * world
    value = total_samples / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: HOST TS in, HOST samples out through the C ABI (pinned buffers), copies inside the timed region
    out_host = torch.empty((nch, nfr * S), dtype=torch.complex64).pin_memory()
    out_np = out_host.numpy()
    chain.run_host(ts_np, nch, nfr, 0, out=out_np)      # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        chain.run_host(ts_np, nch, nfr, 0, out=out_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_samples / e2e_s / 1e6
    checksum = float(np.abs(out_np[0, :4096]).sum())

    # ---- extra: the flowgraph's sink side folded into the last kernel -- x0.2 gain and
    # 16-bit I/Q output, which halves the device-to-host bytes.  Reported beside e2e, not instead of it.
    out16_host = torch.empty((nch, nfr * S, 2), dtype=torch.int16).pin_memory()
    out16_np = out16_host.numpy()
    chain.set_sink(1, 0.2)
    chain.run_host(ts_np, nch, nfr, 0, out=out16_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        chain.run_host(ts_np, nch, nfr, 0, out=out16_np)
    torch.cuda.synchronize()
    e2e16_s = (time.perf_counter() - t0) / args.e2e_steps
    chain.set_sink(0, 1.0)
    if world > 1:
        t = torch.tensor([e2e16_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e16_s = float(t.item())
    e2e16_value = total_samples / e2e16_s / 1e6

    # ---- optional ordered gather of the finished frames to rank 0 (north_star: NCCL only for that)
    gather = None
    if world > 1:
        d_real = torch.view_as_real(d_out)
        bufs = [torch.empty_like(d_real) for _ in range(world)] if rank == 0 else None
        torch.cuda.synchronize(); dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.gather(d_real, bufs, dst=0)
        torch.cuda.synchronize(); dist.barrier()
        g0.record()
        for _ in range(3):
            dist.gather(d_real, bufs, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather = {"ms_per_step": g0.elapsed_time(g1) / 3.0, "bytes_into_root": (world - 1) * d_out.numel() * 8,
                  "note": "ordered NCCL gather of all ranks' frames to rank 0, timed separately (not in value)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    dims = chain.plan("ofdm.dims", np.int32)
    active_items = int(dims[15])
    stage_ms = {k: v / stage_n for k, v in stage_acc.items()}
    ofdm_bytes = frames * 8 * (active_items + S)              # 8*mapped_items + 8*samples per T2 frame
    # mapper kernel in chain mode: packed codewords in, 16-bit cell codes out
    map_bytes = frames * F * ((64800 if cfg["framesize"] else 16200) // 8 + 2 * cell_size)
    ach = ofdm_bytes / (stage_ms["ofdm"] * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ofdm_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_ofdm (cell staging + carrier fill + IFFT + scale + guard interval + P1)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": ofdm_bytes,
                "kernel_ms": stage_ms["ofdm"],
                "stage_ms": stage_ms,
                "map_kernel_gbs": map_bytes / (stage_ms["map"] * 1e-3) / 1e9}

    cpu = None
    if orig_affinity is not None:
        os.sched_setaffinity(0, orig_affinity)
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(cfg, args.cpu_frames, 1)
        if r is not None:
            v = r["samples"] / r["seconds"] / 1e6
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "%d consecutive T2 frames of one %s channel, unmodified reference sources via oracle/_ref "
                             "(GNU Radio shim, single-precision Stockham FFT stand-in for FFTW), one frame per general_work call; "
                             "stage CPU-seconds %s" % (r["frames"], args.config, json.dumps(r["stage_seconds"]))}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "channels_per_gpu": nch, "t2_frames_per_channel_per_step": nfr,
                   "fecframes_per_step": frames * F * world, "samples_per_step": total_samples,
                   "l2": "working set per step (%.0f MB 16-bit cells + %.0f MB samples per GPU) exceeds the 126 MB L2; no explicit flush" % (
                       frames * F * cell_size * 2 / 1e6, frames * S * 8 / 1e6)},
        "x_realtime": value / K.REALTIME_MSPS, "x_realtime_per_gpu": value / K.REALTIME_MSPS / world,
        "fecframes_per_s": frames * F * world / (ms_per_step * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nch * nfr * n_ts), "d2h_bytes_per_step": int(frames * S * 8),
                "api": "dvbt2ll_chain_run_host (pinned host buffers)", "checksum": checksum,
                "host_cores_local_to_gpu": numa_cores},
        "e2e_int16_sink": {"value": e2e16_value, "unit": UNIT, "d2h_bytes_per_step": int(frames * S * 4),
                           "note": "same call with dvbt2ll_chain_set_sink(format=int16 I/Q, gain=0.2): the flowgraph's multiply_const + sc16 conversion fused into the last kernel"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clk,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if gather is not None:
        line["gather"] = gather
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
