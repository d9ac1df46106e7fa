"""Default (exact sequential) DC-reject scan against the opt-in time-parallel one
(PEAQ_B200_HP_PARALLEL=1): deviation of per-frame excitations / MOVs / ODG and timing."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import json, sys, time, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gstpeaq_b200 as G
from signals import synth_pair
ch = 2
ref, test = synth_pair(77, 96000, ch)
e = G.Engine(0, advanced=True)
e.keep_records(True)
out = e.run_host(ref, test, ch)
exc, movs = e.fb_debug(1, ch)
e.keep_records(False)
# timing: one 5-minute pair
n = 48000 * 300
L = G.load_library()
d1 = G.DeviceBuffer(0, n * ch * 4); d2 = G.DeviceBuffer(0, n * ch * 4)
G._check(L.peaq_b200_synth_pairs(0, d1.ptr, d2.ptr, n * ch, 1, 0, n, ch))
for _ in range(2):
    o2 = e.run_device(d1.ptr, d2.ptr, 1, n * ch, ch, n)
ms = e.last_ms(0)
print(json.dumps({"movs": out["movs"][0][:5].tolist(), "odg": float(out["odg"][0]), "exc": np.asarray(exc[0]).ravel().tolist(),
                  "ms_5min": ms, "odg_5min": float(o2["odg"][0]), "movs_5min": o2["movs"][0][:5].tolist()}))
''' % (ROOT, os.path.join(ROOT, "tests"))
res = {}
for par in ("0", "1"):
    env = dict(os.environ, PEAQ_B200_HP_PARALLEL=par)
    p = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=900)
    if p.returncode != 0:
        print(p.stderr[-3000:]); sys.exit(1)
    res[par] = json.loads(p.stdout.strip().splitlines()[-1])
import numpy as np
a, b = res["0"], res["1"]
ea, eb = np.array(a["exc"]), np.array(b["exc"])
rel = np.abs(ea - eb) / np.maximum(np.abs(ea), 1e-300)
print("excitation max rel dev %.3e  (99.9 pct %.3e)" % (rel.max(), np.quantile(rel, 0.999)))
print("movs rel dev", np.abs(np.array(a["movs"]) - np.array(b["movs"])) / np.abs(np.array(a["movs"])))
print("odg dev %.3e" % abs(a["odg"] - b["odg"]))
print("5 min pair: sequential %.1f ms, parallel %.1f ms; odg dev %.3e; movs rel dev %s" % (
    a["ms_5min"], b["ms_5min"], abs(a["odg_5min"] - b["odg_5min"]),
    np.abs(np.array(a["movs_5min"]) - np.array(b["movs_5min"])) / np.abs(np.array(a["movs_5min"]))))
