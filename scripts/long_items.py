"""BASELINE configs[4]-like workload on one GPU: few, long pairs (default 32 pairs x 10 min)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gstpeaq_b200 as G

def main():
    n_pairs = int(os.environ.get("PEAQ_LONG_PAIRS", "32"))
    seconds = int(os.environ.get("PEAQ_LONG_SECONDS", "600"))
    advanced = int(os.environ.get("PEAQ_PROFILE_ADVANCED", "0"))
    ns, ch = 48000 * seconds, 2
    L = G.load_library()
    eng = G.Engine(0, advanced=bool(advanced))
    dref = G.DeviceBuffer(0, n_pairs * ns * ch * 4); dtest = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 0, ns, ch))
    for _ in range(2):
        out = eng.run_device(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, ns)
        fr = int(out["frames_fft"].sum())
        print("pairs %d x %d s adv %d: ms total %.1f frames-kernel %.1f scan %.1f fb %.1f -> %.2f M frames/s, odg %.3f..%.3f" % (
            n_pairs, seconds, advanced, eng.last_ms(0), eng.last_ms(1), eng.last_ms(2), eng.last_ms(4),
            fr / eng.last_ms(0) / 1e3, out["odg"].min(), out["odg"].max()))

if __name__ == "__main__":
    main()
