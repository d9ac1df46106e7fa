"""Development aid: runs the same seeded workloads through two builds of libpeaq_b200.so (each in
its own process) and compares the result rows byte for byte and numerically.

  python scripts/compare_builds.py gstpeaq_b200/libpeaq_b200_r1.so [gstpeaq_b200/libpeaq_b200.so]
"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(out_path):
    sys.path.insert(0, ROOT)
    import gstpeaq_b200 as G
    res = {}
    L = G.load_library()
    for adv in (0, 1):
        eng = G.Engine(0, advanced=bool(adv))
        for name, n_pairs, ns in (("batch", 96, 480000), ("short", 37, 30000)):
            ch = 2
            dref = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
            dtest = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
            G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 1000 * adv, ns, ch))
            out = eng.run_device(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, ns)
            res["%s_adv%d" % (name, adv)] = out
            if name == "short":   # ragged lengths + mono view of the same memory
                lens = np.array([(7919 * (p + 1)) % ns for p in range(n_pairs)], dtype=np.uint64)
                res["ragged_adv%d" % adv] = eng._run(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, lens, ns, True)
                res["mono_adv%d" % adv] = eng._run(dref.ptr, dtest.ptr, n_pairs, ns * ch, 1, None, ns, True)
            dref.free(); dtest.free()
        eng.close()
    np.savez(out_path, **res)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        return worker(sys.argv[2])
    libs = [os.path.abspath(p) for p in (sys.argv[1:] + [os.path.join(ROOT, "gstpeaq_b200", "libpeaq_b200.so")])[:2]]
    outs = []
    for lib in libs:
        f = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, PEAQ_B200_LIBRARY=lib)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--worker", f], env=env)
        outs.append(np.load(f))
    ok = True
    for k in outs[0].files:
        a, b = outs[0][k], outs[1][k]
        same = a.tobytes() == b.tobytes()
        dodg = np.nanmax(np.abs(a["odg"] - b["odg"])) if len(a) else 0.
        n = int(a["n_movs"][0])
        rel = np.nanmax(np.abs(a["movs"][:, :n] - b["movs"][:, :n]) / np.maximum(np.abs(a["movs"][:, :n]), 1e-9))
        print("%-14s bytes equal: %-5s max|dODG| %.3e max rel dMOV %.3e  nan odg %d/%d" %
              (k, same, dodg, rel, int(np.isnan(a["odg"]).sum()), int(np.isnan(b["odg"]).sum())))
        ok &= same
    print("ALL BYTES EQUAL" if ok else "DIFFERENT")


if __name__ == "__main__":
    main()
