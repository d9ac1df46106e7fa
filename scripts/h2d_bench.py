"""Pure pinned host->device copy bandwidth with 1..N GPUs copying at the same time: the ceiling of
bench.py's end-to-end number (16 KB of F32 PCM per PEAQ frame have to cross PCIe).

  python scripts/h2d_bench.py --gpus 1,2,4,8 [--gb 4] [--seconds 3]

One process, one host thread per GPU (like peaq_b200_multi); every thread pins its own buffer
(cudaHostAlloc through the C ABI, first touched by the copying thread itself, pinned to the CPUs
next to its GPU when NVML reports them) and copies it to its device in a loop of 256 MiB
cudaMemcpy calls.  Prints one JSON line per GPU count: aggregate and per-GPU GB/s."""
import argparse, ctypes as C, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gstpeaq_b200 as G


def cpus_near(dev):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        return {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
    except Exception:
        return set()


def worker(dev, nbytes, seconds, start, out, pin_cpu):
    L = G.load_library()
    if pin_cpu:
        near = cpus_near(dev)
        if near:
            try:
                os.sched_setaffinity(threading.get_native_id(), near)
            except Exception:
                pass
    hp = C.c_void_p()
    dp = C.c_void_p()
    G._check(L.peaq_b200_host_alloc_pinned(nbytes, C.byref(hp)))
    C.memset(hp.value, 1, nbytes)
    G._check(L.peaq_b200_device_alloc(dev, nbytes, C.byref(dp)))
    chunk = 256 << 20
    G._check(L.peaq_b200_memcpy_h2d(dev, dp.value, hp.value, min(chunk, nbytes)))   # warm-up
    start.wait()
    t0 = time.perf_counter()
    moved = 0
    while time.perf_counter() - t0 < seconds:
        off = 0
        while off < nbytes:
            n = min(chunk, nbytes - off)
            G._check(L.peaq_b200_memcpy_h2d(dev, dp.value + off, hp.value + off, n))
            off += n
            moved += n
    out[dev] = (moved, time.perf_counter() - t0)
    L.peaq_b200_device_free(dev, dp.value)
    L.peaq_b200_host_free_pinned(hp.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--no-affinity", action="store_true")
    a = ap.parse_args()
    have = G.device_count()
    for n in [int(x) for x in a.gpus.split(",")]:
        if n > have:
            print(json.dumps({"gpus": n, "skipped": "only %d devices visible" % have}))
            continue
        out = {}
        start = threading.Barrier(n)
        ts = [threading.Thread(target=worker, args=(d, int(a.gb * (1 << 30)), a.seconds, start, out, not a.no_affinity))
              for d in range(n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        per = [out[d][0] / out[d][1] / 1e9 for d in range(n)]
        print(json.dumps({"gpus": n, "aggregate_GBps": sum(per), "per_gpu_GBps_min": min(per), "per_gpu_GBps_max": max(per),
                          "buffer_GiB_per_gpu": a.gb, "seconds": a.seconds, "host_memory": "cudaHostAlloc (pinned)",
                          "affinity": "GPU-local CPUs" if not a.no_affinity else "none", "host_cpus": os.cpu_count()}))


if __name__ == "__main__":
    main()
