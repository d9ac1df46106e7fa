"""Small resident workload for ncu captures (592 pairs x 10 s stereo, basic):
one warm-up pass and two measured passes of the batch entry."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gstpeaq_b200 as G

def main():
    n_pairs = int(os.environ.get("PEAQ_PROFILE_PAIRS", "592"))
    advanced = int(os.environ.get("PEAQ_PROFILE_ADVANCED", "0"))
    ns, ch = 48000 * int(os.environ.get("PEAQ_PROFILE_SECONDS", "10")), 2
    L = G.load_library()
    eng = G.Engine(0, advanced=bool(advanced))
    dref = G.DeviceBuffer(0, n_pairs * ns * ch * 4); dtest = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 0, ns, ch))
    for _ in range(3):
        out = eng.run_device(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, ns)
        print("ms total %.2f frames %.2f scan %.2f fb-all %.2f bank %.2f spread+scan %.2f" % tuple(eng.last_ms(k) for k in (0, 1, 2, 4, 5, 6)))

if __name__ == "__main__":
    main()
