"""The GStreamer element shell (gstpeaq_b200/gst/gstpeaqb200.c) against a stand-in for the
GStreamer / GObject API (tests/gst_stub/: the real signatures, a minimal implementation).

GStreamer is not installed in the image, so this is what can be verified about SURVEY 8(f) rank 1:
 - CPU: the element compiles against the API with -Wall -Wextra -Werror (type / signature rot),
 - GPU: a harness plays pipeline -- plugin_init + factory, construct-time property defaults, caps
   queries (a pad offers what the peer of the OTHER pad can do, gstpeaq.c:215-244), CAPS events,
   buffers of unequal sizes on the two pads, EOS aggregation (gstpeaq.c:668-688), PAUSED->READY
   with the console output, property reads -- and the known answers of runtest-1.0.sh come out."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "gst_stub")
CFLAGS = ["-std=gnu11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", STUB, "-I", os.path.join(ROOT, "include")]


def build_harness(tmp_path):
    objs = []
    for src in (os.path.join(ROOT, "gstpeaq_b200", "gst", "gstpeaqb200.c"), os.path.join(STUB, "gst_stub.c"),
                os.path.join(STUB, "harness.c")):
        obj = str(tmp_path / (os.path.basename(src) + ".o"))
        subprocess.check_call(["gcc"] + CFLAGS + ["-c", src, "-o", obj])
        objs.append(obj)
    exe = str(tmp_path / "gst_harness")
    libdir = os.path.join(ROOT, "gstpeaq_b200")
    subprocess.check_call(["gcc", "-o", exe] + objs + ["-L", libdir, "-lpeaq_b200", "-Wl,-rpath," + libdir, "-lm"])
    return exe


def test_element_compiles_against_the_gstreamer_api(tmp_path):
    build_harness(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("channels,advanced,want_odg", [(1, 0, "-2.007"), (2, 0, "-2.007"), (2, 1, "-3.612")])
def test_element_runs_like_in_a_pipeline(tmp_path, channels, advanced, want_odg):
    """runtest-1.0.sh:21-48 (saw against triangle) through the element"""
    import refharness as H
    exe = build_harness(tmp_path)
    n = 128 * 1024
    ref = H.as_interleaved(H.audiotestsrc("saw", n), channels)
    test = H.as_interleaved(H.audiotestsrc("triangle", n), channels)
    fr, ft = str(tmp_path / "ref.f32"), str(tmp_path / "test.f32")
    ref.astype(np.float32).tofile(fr)
    test.astype(np.float32).tofile(ft)
    p = subprocess.run([exe, fr, ft, str(channels), str(advanced)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "HARNESS OK" in p.stdout and "FAIL" not in p.stdout, p.stdout
    console = p.stdout.split("--- console output ---")[1].split("--- end ---")[0]
    assert "Objective Difference Grade: %s" % want_odg in console, console
    assert ("RmsModDiffA = " in console) if advanced else ("   BandwidthRefB: " in console)
    m = re.search(r"RESULT odg (\S+) di (\S+) totalsnr (\S+) odg_mid (\S+)", p.stdout)
    odg, di = float(m.group(1)), float(m.group(2))
    want = H.oracle_run_pair(ref, test, channels, advanced=bool(advanced))
    assert abs(odg - want["odg"]) < 1e-4 and abs(di - want["di"]) < 1e-4
