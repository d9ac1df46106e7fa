#!/usr/bin/env python
"""bench.py -- 48 kHz stereo PEAQ frames/s on synthetic (ref,test) batches.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (our arm)
  python bench.py --impl reference [...]                        (reference CPU arm)
  torchrun ... bench.py --gpus N ...                            (N > 1, one rank per GPU)

Workload (BASELINE.json configs[1]): batch = 4096 pairs per GPU of 10 s, 48 kHz,
stereo, basic mode; pairs are independent, so N GPUs process N*4096 pairs (weak
scaling) and the only collective is the gather of the per-pair results.
A step = one pass of the hot path over the whole batch.  One PEAQ frame = one
1024-sample step of one pair (468 per 10 s pair incl. the padded last frame).

The JSON line carries `value` (device-resident inputs), `e2e` (host buffers
through the C-ABI batch call, H2D/D2H inside the timed region), `roofline`
(frame kernel vs the measured HBM copy bandwidth; algorithmic bytes = 16384 B
read per frame) and `cpu_baseline` (the reference's own C code on
the host cores, bounded sample).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIR_SECONDS = 10
N_SAMPLES = 48000 * PAIR_SECONDS
CHANNELS = 2
PAIRS_PER_GPU = 4096
BYTES_PER_FRAME = 1024 * CHANNELS * 4 * 2        # algorithmic HBM read per PEAQ frame
# CPU arm: pairs per host core and step (basic: ~3 s of work per core, advanced ~20 s)
CPU_PAIRS_PER_CORE = 12
METRIC = "48 kHz stereo PEAQ frames/sec"
UNIT = "frames/s"


def frames_per_pair():
    import gstpeaq_b200 as G
    return G.frames_for_samples(N_SAMPLES)


# --------------------------------------------------------------------------
# CPU arm: the reference's own C path (oracle/_ref) or, if it could not be
# compiled, our C restatement (oracle/).  The ONLY place outside tests/ and
# smoke() that executes anything under oracle/.

def _cpu_worker(args):
    first, count, kind, mode = args
    global MODE
    MODE = mode
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import gstpeaq_b200 as G
    import refharness as H
    ref, test = G.synth_pairs_host(first, count, N_SAMPLES, CHANNELS)   # untimed input generation
    t0 = time.perf_counter()
    frames = 0
    odgs = []
    for p in range(count):
        if kind == "reference":
            r = H.RefPeaq(MODE == "advanced", 92.0, CHANNELS).run(ref[p], test[p])
        else:
            r = H.oracle_run_pair(ref[p], test[p], CHANNELS, advanced=(MODE == "advanced"))
        frames += r["frames_fft"]
        odgs.append(r["odg"])
    return frames, time.perf_counter() - t0, odgs


def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpeaq_ref.so")) else "port"


def run_cpu_sample(n_pairs, cores):
    """frames/s of the CPU path over `n_pairs` pairs spread over `cores` processes
    (processes, not threads: the reference lazily initialises static FFT plans
    and GTypes without locking)."""
    kind = cpu_kind()
    per = max(1, n_pairs // cores)
    jobs = []
    first = 0
    while first < n_pairs:
        c = min(per, n_pairs - first)
        jobs.append((first, c, kind, MODE))
        first += c
    ctx = mp.get_context("spawn")
    with ctx.Pool(min(cores, len(jobs))) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    # throughput over the slowest worker's compute time (generation excluded)
    return frames / busy, frames, busy, wall, kind


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_pairs = max(cores, min(PAIRS_PER_GPU, cores * CPU_PAIRS_PER_CORE))
    times = []
    frames = 0
    kind = cpu_kind()
    for i in range(args.warmup + args.steps):
        fps, frames, busy, wall, kind = run_cpu_sample(n_pairs, cores)
        if i >= args.warmup:
            times.append(busy)
    ms = 1e3 * sum(times) / max(len(times), 1)
    value = frames / (ms / 1e3)
    sample = "%d of %d pairs x %d s per step on %d processes" % (n_pairs, PAIRS_PER_GPU, PAIR_SECONDS, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


MODE = "basic"


def workload_config(n_gpus):
    return {"workload": "batch=%d synthetic 48 kHz stereo %d s pairs per GPU, %s mode (BASELINE configs[%d])"
                        % (PAIRS_PER_GPU, PAIR_SECONDS, MODE, 1 if MODE == "basic" else 2),
            "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": PAIRS_PER_GPU * n_gpus,
            "frames_per_pair": 468, "channels": CHANNELS, "mode": MODE,
            "parallelism": "pairs sharded over %d GPU(s), result gather only" % n_gpus,
            "l2": "inputs (31.5 GB per GPU) far exceed the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_frame():
    """dram bytes per PEAQ frame of the mode's dominant kernel from the committed ncu capture, or None"""
    p = os.path.join(ROOT, "profiles", "ncu_frame_kernel.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_frame" if MODE == "basic" else "bank_dram_bytes_per_frame"])
        except Exception:
            return None
    return None


def our_arm(args):
    import numpy as np
    import gstpeaq_b200 as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if G.device_count() <= local_rank:
        raise G.PeaqError("no CUDA device for rank %d: the engine has no CPU fallback" % rank)

    from gstpeaq_b200 import parallel
    n_global = PAIRS_PER_GPU * world
    first, count = parallel.shard_range(n_global, rank, world)
    fpp = frames_per_pair()
    L = G.load_library()
    eng = G.Engine(local_rank, advanced=(MODE == "advanced"))
    stride = N_SAMPLES * CHANNELS
    nbytes = count * stride * 4
    dref = G.DeviceBuffer(local_rank, nbytes)
    dtest = G.DeviceBuffer(local_rank, nbytes)
    G._check(L.peaq_b200_synth_pairs(local_rank, dref.ptr, dtest.ptr, stride, count, first, N_SAMPLES, CHANNELS))

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        return eng.run_device(dref.ptr, dtest.ptr, count, stride, CHANNELS, N_SAMPLES)

    for _ in range(args.warmup):
        out = step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    barrier()
    t0 = time.perf_counter()
    dev_ms = k1_ms = k2_ms = bank_ms = 0.0
    for _ in range(args.steps):
        out = step()
        dev_ms += eng.last_ms(0)       # CUDA events on the engine's stream
        k1_ms += eng.last_ms(1)
        k2_ms += eng.last_ms(2)
        bank_ms += eng.last_ms(5)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # gather of the per-pair results (the path's only collective), outside the
    # timed steps: it moves 128 B per pair once per job
    if dist is not None:
        full = parallel.gather_results(out, n_global, torch.device("cuda", local_rank))
        t = torch.tensor([dev_ms, wall_ms, k1_ms, k2_ms, bank_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, k1_ms, k2_ms, bank_ms = [float(x) for x in t.tolist()]
    else:
        full = out
    frames_step = int(full["frames_fft"].sum())
    assert frames_step == n_global * fpp, (frames_step, n_global * fpp)
    nan_odg = int(np.isnan(full["odg"]).sum())
    ms_per_step = dev_ms / args.steps
    value = frames_step / (ms_per_step / 1e3)

    # ---- end to end: host (pinned) buffers through the batch call ------------
    e2e = None
    # pin this rank to the CPUs next to its GPU while the host buffers are allocated and used:
    # with the default first-touch policy they then live on the GPU's NUMA node, and 8 ranks do
    # not pull 250 GB per step across the socket interconnect
    old_affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        if cpus:
            old_affinity = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus & old_affinity or old_affinity)
    except Exception:
        old_affinity = None
    try:
        e2e_pairs = count
        try:
            import psutil
            # every rank of the node pins its own buffers at the same time
            avail = psutil.virtual_memory().available / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
            while e2e_pairs > 64 and 2 * e2e_pairs * stride * 4 * 1.3 > avail:
                e2e_pairs //= 2
        except Exception:
            pass
        hb = e2e_pairs * stride * 4
        pr = G.C.c_void_p()
        pt = G.C.c_void_p()
        G._check(L.peaq_b200_host_alloc_pinned(hb, G.C.byref(pr)))
        G._check(L.peaq_b200_host_alloc_pinned(hb, G.C.byref(pt)))
        G._check(L.peaq_b200_memcpy_d2h(local_rank, pr.value, dref.ptr, hb))
        G._check(L.peaq_b200_memcpy_d2h(local_rank, pt.value, dtest.ptr, hb))
        dref.free()
        dtest.free()
        e2e_steps = max(1, min(args.steps, 3))
        eng._run(pr.value, pt.value, e2e_pairs, stride, CHANNELS, None, N_SAMPLES, on_device=False)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            o2 = eng._run(pr.value, pt.value, e2e_pairs, stride, CHANNELS, None, N_SAMPLES, on_device=False)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        if dist is not None:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": e2e_pairs * world * fpp / (e2e_ms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": 2 * hb * world, "d2h_bytes_per_step": e2e_pairs * world * 128,
               "ms_per_step": e2e_ms, "pairs_per_gpu": e2e_pairs, "steps": e2e_steps,
               "host_memory": "pinned", "odg_equal_to_resident_run": bool(np.array_equal(o2["odg"], out["odg"][:e2e_pairs], equal_nan=True))}
        L.peaq_b200_host_free_pinned(pr.value)
        L.peaq_b200_host_free_pinned(pt.value)
    except Exception as exc:   # report, never hide
        e2e = {"value": None, "unit": UNIT, "error": str(exc)}
    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)   # the CPU baseline below uses every core

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    # dominant kernel: fft_frames_kernel (basic) / fb_bank_rec_kernel (advanced).  Algorithmic
    # bytes per launch = 16384 B x the frames the launch covers (both modes read the
    # PCM once per ear model); duration = CUDA-event time of those launches inside the timed
    # steps (per GPU: the frames of one rank)
    frames_rank = count * fpp
    dom_ms = (k1_ms if MODE == "basic" else bank_ms) / args.steps
    achieved = frames_rank * BYTES_PER_FRAME / (dom_ms / 1e3) / 1e9
    traffic = ncu_traffic_per_frame()
    roofline = {"bound": "hbm", "kernel": "fft_frames_kernel" if MODE == "basic" else "fb_bank_rec_kernel",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic * frames_rank if traffic else None,
                "algorithmic_bytes_per_launch_set": frames_rank * BYTES_PER_FRAME,
                "kernel_ms_per_step": dom_ms, "frame_kernel_ms_per_step": k1_ms / args.steps,
                "scan_kernel_ms_per_step": k2_ms / args.steps,
                "kernel_share_of_step": dom_ms * args.steps / dev_ms,
                "note": "FP64 latency/issue bound in practice (DESIGN.md 3); HBM fraction reported as the north star asks"}
    # the binding resource next to it: the FP64 pipe, against a measured DFMA loop
    try:
        fp64_peak = float(L.peaq_b200_fp64_peak_tflops(local_rank))
    except Exception:
        fp64_peak = -1.0
    if fp64_peak > 0:
        roofline["fp64_peak_tflops_measured"] = fp64_peak
        if MODE == "advanced":
            # fb_bank_rec_kernel: 384 FMAs per band and 32-sample sub-step, 40 bands, 32 sub-steps
            # and 4 streams per PEAQ frame
            flop = 2.0 * 384 * 40 * 32 * 2 * CHANNELS * frames_rank
            roofline["fp64_algorithmic_flop_per_launch_set"] = flop
            roofline["fp64_achieved_tflops"] = flop / (dom_ms / 1e3) / 1e12
            roofline["fp64_frac"] = roofline["fp64_achieved_tflops"] / fp64_peak

    # ---- CPU baseline (N = 1 only): bounded sample on the host cores ------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = max(cores, min(PAIRS_PER_GPU, cores * CPU_PAIRS_PER_CORE))
        fps, fr, busy, wall, kind = run_cpu_sample(n_cpu, cores)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "first %d of %d pairs x %d s, one process per core, %.1f s wall" % (n_cpu, PAIRS_PER_GPU, PAIR_SECONDS, wall)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms / args.steps, "nan_odg_pairs": nan_odg,
        "odg_min": float(np.nanmin(full["odg"])), "odg_max": float(np.nanmax(full["odg"])),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="basic", choices=["basic", "advanced"],
                    help="basic = BASELINE configs[1] (default, the headline); advanced = configs[2]")
    args = ap.parse_args()
    global MODE
    MODE = args.mode
    if args.impl == "reference":
        return reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--mode", args.mode]
        return subprocess.call(cmd)
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
