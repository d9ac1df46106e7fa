"""Development aid: resident-step timings of the headline batches for the library named by
PEAQ_B200_LIBRARY (default: the in-tree build).  usage: time_modes.py [basic|advanced|both] [pairs] [seconds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gstpeaq_b200 as G

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "basic"
    n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    seconds = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    ns, ch = 48000 * seconds, 2
    L = G.load_library()
    dref = G.DeviceBuffer(0, n_pairs * ns * ch * 4); dtest = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 0, ns, ch))
    for adv in ([0, 1] if which == "both" else [1 if which == "advanced" else 0]):
        eng = G.Engine(0, advanced=bool(adv))
        best = None
        every = []
        for i in range(5):
            out = eng.run_device(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, ns)
            t = [eng.last_ms(k) for k in range(7)]
            every.append(round(t[0], 1))
            if i and (best is None or t[0] < best[0]):
                best = t
        print("  totals of the passes (ms):", every)
        fr = int(out["frames_fft"].sum())
        print("%s adv %d pairs %d x %d s: total %.1f ms (frames %.1f scan %.1f fb-all %.1f bank %.1f spread+scan %.1f) -> %.3f M frames/s  odg[0..2] %s" % (
            os.path.basename(G.library_path()), adv, n_pairs, seconds, best[0], best[1], best[2], best[4], best[5], best[6],
            fr / best[0] / 1e3, out["odg"][:3]))
        eng.close()

if __name__ == "__main__":
    main()
