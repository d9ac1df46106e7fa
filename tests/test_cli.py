"""The `peaq` command line (gstpeaq_b200/cli/peaq.c) against the reference's
CLI contract (/root/reference/src/peaq.c:35-45, :110-133, :146-150, :217-220;
doc/man/peaq.xml): options, the two output lines, exit codes."""
import os
import struct
import subprocess

import numpy as np
import pytest

import gstpeaq_b200 as G
from signals import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAQ = os.path.join(ROOT, "gstpeaq_b200", "peaq")


def write_wav(path, data, channels, rate=48000, kind="float32"):
    data = np.asarray(data, dtype=np.float32).reshape(-1)
    if kind == "float32":
        tag, bits, raw = 3, 32, data.astype("<f4").tobytes()
    elif kind == "int16":
        tag, bits, raw = 1, 16, np.round(data * 32768).clip(-32768, 32767).astype("<i2").tobytes()
    elif kind == "int24":
        v = np.round(data.astype(np.float64) * 8388608).clip(-8388608, 8388607).astype("<i4")
        raw = b"".join(struct.pack("<i", int(x))[:3] for x in v)
        tag, bits = 1, 24
    else:
        raise ValueError(kind)
    block = channels * bits // 8
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, tag, channels, rate, rate * block, block, bits) + b"data" + struct.pack("<I", len(raw))
    with open(path, "wb") as f:
        f.write(hdr + raw)


def run(*args):
    p = subprocess.run([PEAQ] + list(args), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return p.returncode, p.stdout


def test_cli_version_and_usage():
    rc, out = run("--version")
    assert rc == 0 and out.startswith("peaq-b200")
    rc, out = run()                       # no files: help text, exit 1 (peaq.c:123-133)
    assert rc == 1 and "REFFILE TESTFILE" in out
    rc, out = run("a.wav", "b.wav", "c.wav")
    assert rc == 1
    rc, out = run("--bogus", "a.wav", "b.wav")
    assert rc == 1 and "Failed to initialize" in out      # peaq.c:110-115


def test_cli_rejects_bad_input(tmp_path):
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wav file at all")
    rc, out = run(str(bad), str(bad))
    assert rc == 1 and out.startswith("Error:")
    w = tmp_path / "r44.wav"
    write_wav(str(w), np.zeros(1000, np.float32), 1, rate=44100)
    rc, out = run(str(w), str(w))
    assert rc == 1 and "48000" in out


@pytest.mark.skipif(G.device_count() > 0, reason="a GPU is present")
def test_cli_without_gpu_reports_missing_engine(tmp_path):
    w = tmp_path / "s.wav"
    write_wav(str(w), np.zeros(4800, np.float32), 1)
    rc, out = run(str(w), str(w))
    assert rc == 2 and "no CPU fallback" in out          # exit 2 = element missing (peaq.c:146-150)


@pytest.mark.gpu
def test_cli_known_answers(tmp_path):
    """runtest-1.0.sh through the CLI: 0.171, -2.007 (mono, and mono vs stereo)"""
    cases = golden_cases()
    sine, _, _ = cases["kat_sine_sine_mono"]
    saw, tri, _ = cases["kat_saw_tri_mono"]
    saw2, tri2, _ = cases["kat_saw_tri_stereo"]
    p = lambda n: str(tmp_path / n)
    write_wav(p("sine.wav"), sine, 1)
    write_wav(p("saw.wav"), saw, 1)
    write_wav(p("tri.wav"), tri, 1)
    write_wav(p("saw2.wav"), saw2, 2)
    write_wav(p("tri2.wav"), tri2, 2)
    rc, out = run(p("sine.wav"), p("sine.wav"))
    assert rc == 0
    lines = out.strip().splitlines()
    assert lines[0] == "Objective Difference Grade: 0.171" and lines[1].startswith("Distortion Index: 4.437")
    for a, b in (("saw.wav", "tri.wav"), ("saw2.wav", "tri.wav"), ("saw.wav", "tri2.wav"), ("saw2.wav", "tri2.wav")):
        rc, out = run("--basic", p(a), p(b))
        assert rc == 0 and out.splitlines()[0] == "Objective Difference Grade: -2.007", (a, b, out)
    rc, out = run("--advanced", p("saw.wav"), p("tri.wav"))
    assert rc == 0 and out.splitlines()[0] == "Objective Difference Grade: -3.612"   # SURVEY 8c anchor


@pytest.mark.gpu
def test_cli_integer_pcm_matches_float(tmp_path):
    """16-bit full scale must map to 1.0 (audioconvert scaling); the synthetic
    signals are exactly representable in 16 bit, so both files give the same ODG"""
    r, t = G.synth_pairs_host(3, 1, 48000, 2)
    p = lambda n: str(tmp_path / n)
    write_wav(p("rf.wav"), r[0], 2, kind="float32")
    write_wav(p("tf.wav"), t[0], 2, kind="float32")
    write_wav(p("ri.wav"), r[0], 2, kind="int16")
    write_wav(p("ti.wav"), t[0], 2, kind="int16")
    write_wav(p("r24.wav"), r[0], 2, kind="int24")
    write_wav(p("t24.wav"), t[0], 2, kind="int24")
    outs = [run(p(a), p(b)) for a, b in (("rf.wav", "tf.wav"), ("ri.wav", "ti.wav"), ("r24.wav", "t24.wav"))]
    assert all(rc == 0 for rc, _ in outs)
    assert outs[0][1] == outs[1][1] == outs[2][1]
