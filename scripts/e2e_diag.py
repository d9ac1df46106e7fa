"""Where does the end-to-end (host buffers) step spend its time?  Engine timers of a host-batch run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gstpeaq_b200 as G

def main():
    n_pairs = int(os.environ.get("PEAQ_PROFILE_PAIRS", "4096"))
    advanced = int(os.environ.get("PEAQ_PROFILE_ADVANCED", "1"))
    ns, ch = 48000 * int(os.environ.get("PEAQ_PROFILE_SECONDS", "10")), 2
    L = G.load_library()
    eng = G.Engine(0, advanced=bool(advanced))
    nbytes = n_pairs * ns * ch * 4
    dref = G.DeviceBuffer(0, nbytes); dtest = G.DeviceBuffer(0, nbytes)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 0, ns, ch))
    pr = G.C.c_void_p(); pt = G.C.c_void_p()
    G._check(L.peaq_b200_host_alloc_pinned(nbytes, G.C.byref(pr)))
    G._check(L.peaq_b200_host_alloc_pinned(nbytes, G.C.byref(pt)))
    G._check(L.peaq_b200_memcpy_d2h(0, pr.value, dref.ptr, nbytes))
    G._check(L.peaq_b200_memcpy_d2h(0, pt.value, dtest.ptr, nbytes))
    for k in range(3):
        eng.run_device(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, ns)
        print("resident: total %.1f K1 %.1f scan %.1f fb-all %.1f bank %.1f spread+scan %.1f" % tuple(eng.last_ms(i) for i in (0, 1, 2, 4, 5, 6)))
    dref.free(); dtest.free()
    for k in range(3):
        t0 = time.perf_counter()
        eng._run(pr.value, pt.value, n_pairs, ns * ch, ch, None, ns, on_device=False)
        w = (time.perf_counter() - t0) * 1e3
        print("host    : wall %.1f total %.1f K1 %.1f scan %.1f fb-all %.1f bank %.1f spread+scan %.1f" % ((w,) + tuple(eng.last_ms(i) for i in (0, 1, 2, 4, 5, 6))))

if __name__ == "__main__":
    main()
