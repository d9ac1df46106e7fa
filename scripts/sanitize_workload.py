"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck):
  python scripts/sanitize_workload.py [short] [segments] [session] [fused]
short: 2 ragged pairs per mode and channel count through the batch entry (all batch kernels, the
block-parallel DC-reject passes included); segments: one 50 s mono item per mode (two segments:
seg_init / seg_combine / gather, base offsets into the PCM); session: streaming pushes of odd
sizes in both modes (sample-by-sample DC-reject kernel, pinned staging); fused: the fused
persistent kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gstpeaq_b200 as G

what = sys.argv[1:] or ["short"]
if "short" in what:
    for adv in (False, True):
        for ch in (2, 1):
            n = 12000 + 777
            r, t = G.synth_pairs_host(0, 2, n, ch)
            e = G.Engine(0, advanced=adv)
            out = e.run_host(r, t, ch, n_samples=np.array([n, n - 4321], np.uint64))
            print("adv", adv, "ch", ch, out["odg"])
            e.close()
if "segments" in what:
    for adv in (False, True):
        n = 48000 * 50
        assert G.segment_plan(n)[0] == 2
        r, t = G.synth_pairs_host(3, 1, n, 1)
        e = G.Engine(0, advanced=adv)
        out = e.run_host(r, t, 1)
        print("segments adv", adv, out["odg"], out["frames_fft"], out["frames_fb"])
        e.close()
if "session" in what:
    rng = np.random.default_rng(1)
    for adv in (False, True):
        r, t = G.synth_pairs_host(5, 1, 30000, 2)
        p = G.Peaq(0, advanced=adv, console_output=False)
        p.set_caps(2)
        a = b = 0
        while a < r.size or b < t.size:
            na, nb = 2 * int(rng.integers(1, 5000)), 2 * int(rng.integers(1, 5000))
            p.chain_ref(r[0][a:a + na]); p.chain_test(t[0][b:b + nb])
            a += na; b += nb
        print("session adv", adv, p.stop()["odg"])
        p.close()
if "fused" in what:
    os.environ["PEAQ_B200_FUSED"] = "1"
    r, t = G.synth_pairs_host(0, 3, 20000, 2)
    e = G.Engine(0, advanced=False)
    print("fused", e.run_host(r, t, 2)["odg"])
    e.close()
