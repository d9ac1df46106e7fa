"""GPU diagnostic (development aid): compares the engine's per-frame records and
final results with the CPU oracle and prints per-field maximum errors, then
times a medium batch.  Run under gpurun."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gstpeaq_b200 as G
from refharness import OraclePeaq, audiotestsrc, as_interleaved

def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d[both_nan] = 0
    d[(a == b)] = 0
    return float(np.nanmax(d)) if d.size else 0.0, int(np.sum(np.isnan(a) != np.isnan(b)))

def compare(name, ref, test, channels, eng):
    ref = np.ascontiguousarray(ref, dtype=np.float32); test = np.ascontiguousarray(test, dtype=np.float32)
    n = ref.size // channels
    nf = G.frames_for_samples(n)
    o = OraclePeaq(False, 92.0, channels, fft_trace=nf)
    ores = o.run(ref, test)
    tr = o.fft_trace
    eng.keep_records(True)
    res = eng.run_host(ref[None, :], test[None, :], channels)
    rec = eng.records(1, nf)
    B = 109
    print("==", name, "frames", nf, res["frames_fft"][0], ores["frames_fft"])
    print(" unsmeared", relerr(rec["unsmeared"][0], tr["unsmeared"][:, :, :channels, :B]))
    print(" noise", relerr(rec["noise_in_bands"][0], tr["noise_in_bands"][:, :channels, :B]))
    fl = rec["flags"][0]
    print(" above mismatches", int(np.sum((fl & 1) != tr["above_threshold"])), " ehs_valid mismatches", int(np.sum(((fl >> 1) & 1) != tr["ehs_valid"])))
    print(" bw_ref mism", int(np.sum(rec["bw_ref"][0] != tr["bw_ref"][:, :channels])), " bw_test mism", int(np.sum(rec["bw_test"][0] != tr["bw_test"][:, :channels])))
    valid = tr["ehs_valid"].astype(bool)
    print(" ehs", relerr(rec["ehs"][0][valid], tr["ehs"][valid][:, :channels]), "max abs", float(np.max(np.abs(rec["ehs"][0][valid] - tr["ehs"][valid][:, :channels]))) if valid.any() else 0)
    print(" snr cum", relerr(np.cumsum(rec["snr"][0][:, 0]), tr["signal_energy"]), relerr(np.cumsum(rec["snr"][0][:, 1]), tr["noise_energy"]))
    print(" movs gpu   ", np.array2string(res["movs"][0], precision=6))
    print(" movs oracle", np.array2string(ores["movs"], precision=6))
    print(" movs rel", relerr(res["movs"][0], ores["movs"]), "di", res["di"][0], ores["di"], "odg", res["odg"][0], ores["odg"],
          "lrf", res["loudness_reached_frame"][0], ores["loudness_reached_frame"], "snr", res["totalsnr"][0], ores["totalsnr"])
    eng.keep_records(False)

def main():
    print("devices", G.device_count())
    eng = G.Engine(0, False, 92.0)
    n = 128 * 1024
    s = audiotestsrc("sine", n); saw = audiotestsrc("saw", n); tri = audiotestsrc("triangle", n)
    compare("sine/sine mono", s, s, 1, eng)
    compare("saw/tri mono", saw, tri, 1, eng)
    compare("saw/tri stereo", as_interleaved(saw, 2), as_interleaved(tri, 2), 2, eng)
    ref, test = G.synth_pairs_host(0, 4, 120000, 2)
    for p in range(2):
        compare("synth %d" % p, ref[p], test[p], 2, eng)
    # silence lead-in / tail, ragged length
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(50000 * 2) * 0.05).astype(np.float32)
    y = x + (rng.standard_normal(x.size) * 0.005).astype(np.float32)
    x[:20000] = 0; y[:20000] = 0; x[-16000:] = 0; y[-16000:] = 0
    compare("noise with silence", x[:2 * 48765], y[:2 * 48765], 2, eng)
    # device-side generator == host generator
    L = G.load_library()
    npairs, ns, ch = 3, 50000, 2
    dref = G.DeviceBuffer(0, npairs * ns * ch * 4); dtest = G.DeviceBuffer(0, npairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, npairs, 5, ns, ch))
    hr = np.zeros((npairs, ns * ch), np.float32); ht = np.zeros_like(hr)
    G._check(L.peaq_b200_memcpy_d2h(0, hr.ctypes.data, dref.ptr, hr.nbytes)); G._check(L.peaq_b200_memcpy_d2h(0, ht.ctypes.data, dtest.ptr, ht.nbytes))
    r2, t2 = G.synth_pairs_host(5, npairs, ns, ch)
    print("synth dev==host", bool(np.array_equal(hr, r2)), bool(np.array_equal(ht, t2)))
    # timing: 512 pairs x 10 s resident
    npairs, ns = 512, 480000
    dref = G.DeviceBuffer(0, npairs * ns * ch * 4); dtest = G.DeviceBuffer(0, npairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, npairs, 0, ns, ch))
    for it in range(3):
        t0 = time.time()
        out = eng.run_device(dref.ptr, dtest.ptr, npairs, ns * ch, ch, ns)
        dt = time.time() - t0
        fr = int(out["frames_fft"].sum())
        print("batch %d pairs: wall %.3fs total %.1f ms frames_k %.1f ms scan %.1f ms -> %.3f Mframes/s ; odg range %.3f..%.3f nan=%d"
              % (npairs, dt, eng.last_ms(0), eng.last_ms(1), eng.last_ms(2), fr / eng.last_ms(0) / 1e3,
                 np.nanmin(out["odg"]), np.nanmax(out["odg"]), int(np.isnan(out["odg"]).sum())))
    # oracle on a few of them
    r2, t2 = G.synth_pairs_host(0, 2, ns, ch)
    from refharness import oracle_run_pair
    for p in range(2):
        t0 = time.time(); o = oracle_run_pair(r2[p], t2[p], ch); dt = time.time() - t0
        print("pair", p, "gpu odg", out["odg"][p], "oracle", o["odg"], "movs rel", relerr(out["movs"][p], o["movs"]), "cpu s", dt)

if __name__ == "__main__":
    main()
