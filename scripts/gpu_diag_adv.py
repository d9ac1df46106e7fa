"""GPU diagnostic for advanced mode (development aid)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gstpeaq_b200 as G
from refharness import OraclePeaq, audiotestsrc, as_interleaved, oracle_run_pair
from gpu_diag import relerr

def compare(name, ref, test, ch, eng):
    ref = np.ascontiguousarray(ref, np.float32); test = np.ascontiguousarray(test, np.float32)
    n = ref.size // ch
    nf = G.frames_for_samples(n); nfb = (n + 191) // 192
    o = OraclePeaq(True, 92.0, ch, fft_trace=nf, fb_trace=nfb)
    ores = o.run(ref, test)
    eng.keep_records(True)
    res = eng.run_host(ref[None, :], test[None, :], ch)
    exc, movs = eng.fb_debug(1, ch)
    rec = eng.records(1, nf)
    eng.keep_records(False)
    tr = o.fb_trace; ft = o.fft_trace
    print("==", name, "fft frames", nf, res["frames_fft"][0], ores["frames_fft"], "fb frames", nfb, res["frames_fb"][0], ores["frames_fb"])
    # exc[0]: [frame, stream(2c+side), U|E, 40]; trace: [frame, side, ch, 40]
    for c in range(ch):
        for side in range(2):
            u = exc[0][:nfb, 2 * c + side, 0]; e = exc[0][:nfb, 2 * c + side, 1]
            print("  ch%d side%d U %s  E %s" % (c, side, relerr(u, tr["unsmeared"][:, side, c]), relerr(e, tr["excitation"][:, side, c])))
    print("  above mism", int(np.sum(movs[0][:nfb, 0, 5] != tr["above_threshold"])))
    for k, nm in enumerate(["mod_diff", "temp_wt", "noise_loud", "missing_comp", "lin_dist"]):
        print("  %-12s %s" % (nm, relerr(movs[0][:nfb, :, k], tr[nm][:, :ch])))
    print("  fft: unsmeared(ref)", relerr(rec["unsmeared"][0][:, 0], ft["unsmeared"][:, 0, :ch, :55]), "noise", relerr(rec["noise_in_bands"][0], ft["noise_in_bands"][:, :ch, :55]))
    print("  movs gpu   ", np.array2string(res["movs"][0][:5], precision=8))
    print("  movs oracle", np.array2string(ores["movs"], precision=8))
    print("  movs rel", relerr(res["movs"][0][:5], ores["movs"]), "di", res["di"][0], ores["di"], "odg", res["odg"][0], ores["odg"], "lrf", res["loudness_reached_frame"][0], ores["loudness_reached_frame"])

def main():
    eng = G.Engine(0, True, 92.0)
    n = 128 * 1024
    s = audiotestsrc("sine", n); saw = audiotestsrc("saw", n); tri = audiotestsrc("triangle", n)
    compare("sine/sine mono", s, s, 1, eng)
    compare("saw/tri stereo", as_interleaved(saw, 2), as_interleaved(tri, 2), 2, eng)
    ref, test = G.synth_pairs_host(0, 2, 60000, 2)
    compare("synth0", ref[0], test[0], 2, eng)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(48765 * 2) * 0.05).astype(np.float32)
    y = x + (rng.standard_normal(x.size) * 0.005).astype(np.float32)
    x[:40000] = 0; y[:40000] = 0; x[-32000:] = 0; y[-32000:] = 0
    compare("noise with silence", x, y, 2, eng)
    # chunked fb clock must equal single chunk
    os.environ["PEAQ_B200_FB_BUDGET_MB"] = "1"; os.environ["PEAQ_B200_RECORD_BUDGET_MB"] = "1"
    e2 = G.Engine(0, True, 92.0)
    r, t = G.synth_pairs_host(10, 3, 70000, 2)
    a = eng.run_host(r, t, 2); b = e2.run_host(r, t, 2)
    print("chunked == single:", bool(np.array_equal(a["movs"], b["movs"])), a["odg"], b["odg"])
    del os.environ["PEAQ_B200_FB_BUDGET_MB"]; del os.environ["PEAQ_B200_RECORD_BUDGET_MB"]
    # timing
    L = G.load_library(); npairs, ns, ch = 256, 480000, 2
    dref = G.DeviceBuffer(0, npairs * ns * ch * 4); dtest = G.DeviceBuffer(0, npairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, npairs, 0, ns, ch))
    for it in range(2):
        out = eng.run_device(dref.ptr, dtest.ptr, npairs, ns * ch, ch, ns)
        fr = int(out["frames_fft"].sum())
        print("adv batch %d pairs: total %.1f ms frames_k %.1f scan %.1f fb %.1f -> %.3f Mframes/s odg %.3f..%.3f nan=%d"
              % (npairs, eng.last_ms(0), eng.last_ms(1), eng.last_ms(2), eng.last_ms(4), fr / eng.last_ms(0) / 1e3,
                 np.nanmin(out["odg"]), np.nanmax(out["odg"]), int(np.isnan(out["odg"]).sum())))
    r2, t2 = G.synth_pairs_host(0, 1, ns, ch)
    t0 = time.time(); o = oracle_run_pair(r2[0], t2[0], ch, advanced=True); dt = time.time() - t0
    print("pair0 gpu odg", out["odg"][0], "oracle", o["odg"], "movs rel", relerr(out["movs"][0][:5], o["movs"]), "cpu s", dt)

if __name__ == "__main__":
    main()
