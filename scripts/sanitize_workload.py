"""Tiny workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gstpeaq_b200 as G
for adv in (False, True):
    for ch in (2, 1):
        r, t = G.synth_pairs_host(0, 2, 12000 + 777, ch)
        e = G.Engine(0, advanced=adv)
        out = e.run_host(r, t, ch)
        print("adv", adv, "ch", ch, out["odg"])
        e.close()
