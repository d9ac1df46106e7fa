"""CPU: pins the oracle (oracle/peaq_oracle.c) to the reference.

  - the reference's own golden vectors (testpeaq.c:37-599) with the reference's
    own tolerance (abs 5e-6 or rel 5e-5, testpeaq.c:32-35,606-621);
  - its known-answer ODGs 0.171 / -2.007 (runtest-1.0.sh:18,28,38,48);
  - outputs of the reference itself (oracle/_ref) committed in
    tests/golden/ref_outputs.npz, and -- when oracle/_ref is present -- live,
    frame by frame, including chunked pushes and unequal lengths.
"""
import math
import os

import numpy as np
import pytest

import refharness as H
from signals import golden_cases, noise_pair, synth_pair

DELTA, RELDELTA = 0.000005, 0.00005


def assert_testpeaq_close(dut, ref, name):
    """assertArrayEquals of testpeaq.c:606-621"""
    dut = np.asarray(dut)
    ref = np.asarray(ref)
    diff = dut - ref
    rel = 2 * (dut - ref) / (dut + ref)
    bad = (np.abs(diff) > DELTA) & (np.abs(rel) > RELDELTA)
    assert not bad.any(), "%s: %d values off, first at %d" % (name, bad.sum(), np.argmax(bad))


def test_struct_layouts_match_numpy_mirrors():
    L = H.oracle_lib()
    L.peaq_oracle_sizeof.restype = H.C.c_size_t
    assert L.peaq_oracle_sizeof(0) == H.FFT_TRACE_DTYPE.itemsize
    assert L.peaq_oracle_sizeof(1) == H.FB_TRACE_DTYPE.itemsize
    assert L.peaq_oracle_sizeof(2) == H.C.sizeof(H.OracleResult)


def test_golden_fft_ear_model(golden_vectors):
    """test_ear of testpeaq.c:655-705: step frame, ramp frame, SPL of a sine"""
    L = H.oracle_lib()
    s = L.peaq_oracle_stage_new(109)
    x = np.empty(2048, np.float32)
    x[:1024] = -1
    x[1024] = 0
    x[1025:] = 1
    ps = np.zeros(1025); pw = np.zeros(1025); un = np.zeros(109); ex = np.zeros(109)
    L.peaq_oracle_stage_fft_ear(s, x.ctypes.data, None, None, None, None)
    x = ((np.arange(2048) - 1024) / np.float32(1024)).astype(np.float32)
    x = (np.arange(2048, dtype=np.float32) - 1024) / np.float32(1024)
    L.peaq_oracle_stage_fft_ear(s, x.ctypes.data, ps.ctypes.data, pw.ctypes.data, un.ctypes.data, ex.ctypes.data)
    assert_testpeaq_close(ps, golden_vectors["fft_ref_data"] ** 2, "absolute_spectrum")
    assert_testpeaq_close(pw, golden_vectors["weighted_fft_ref_data"] ** 2, "weighted_fft")
    assert_testpeaq_close(un, golden_vectors["unsmeared_excitation_ref"], "unsmeared_excitation")
    assert_testpeaq_close(ex, golden_vectors["excitation_ref"], "excitation")
    for frame in range(10):
        i = np.arange(2048)
        x = np.sin(2 * math.pi * 1019.5 / 48000. * (i + frame * 1024)).astype(np.float32)
        L.peaq_oracle_stage_fft_ear(s, x.ctypes.data, ps.ctypes.data, None, None, None)
        spl = 10 * math.log10(ps[43])
        assert 91.9999 < spl < 92.0001
    L.peaq_oracle_stage_free(s)


def test_golden_loudness_scalars():
    """testpeaq.c:707-744: 1 kHz at 40 dB SPL -> 0.58..0.59 (FFT), 1.03..1.04 (filter bank)"""
    L = H.oracle_lib()
    s = L.peaq_oracle_stage_new(109)
    scale = 10. ** ((40. - 92.) / 20)
    for frame in range(50):
        i = np.arange(2048)
        x = (scale * np.sin(2 * math.pi * 1000. / 48000. * (i + frame * 1024))).astype(np.float32)
        L.peaq_oracle_stage_fft_ear(s, x.ctypes.data, None, None, None, None)
    assert 0.58 < L.peaq_oracle_stage_loudness(s, 0) < 0.59
    for frame in range(250):
        i = np.arange(192)
        x = (scale * np.sin(2 * math.pi * 1000. / 48000. * (i + frame * 192))).astype(np.float32)
        L.peaq_oracle_stage_fb_ear(s, x.ctypes.data, None, None)
    assert 1.03 < L.peaq_oracle_stage_loudness(s, 1) < 1.04
    L.peaq_oracle_stage_free(s)


def test_golden_level_adapter(golden_vectors):
    """test_leveladapt of testpeaq.c:747-784"""
    L = H.oracle_lib()
    s = L.peaq_oracle_stage_new(109)
    ref = np.arange(1, 110, dtype=np.float64)
    test = np.arange(109, 0, -1, dtype=np.float64)
    o1 = np.zeros(109); o2 = np.zeros(109)
    for k in (1, 2):
        L.peaq_oracle_stage_level_adapt(s, ref.ctypes.data, test.ctypes.data, o1.ctypes.data, o2.ctypes.data)
        assert_testpeaq_close(o1, golden_vectors["spectrally_adapted_ref_patterns%d_ref" % k], "ref%d" % k)
        assert_testpeaq_close(o2, golden_vectors["spectrally_adapted_test_patterns%d_ref" % k], "test%d" % k)
    L.peaq_oracle_stage_free(s)


def test_golden_modulation_processor(golden_vectors):
    """test_modulationproc of testpeaq.c:786-810"""
    L = H.oracle_lib()
    s = L.peaq_oracle_stage_new(109)
    x = np.arange(1, 110, dtype=np.float64)
    m = np.zeros(109); l = np.zeros(109)
    for k in (1, 2):
        L.peaq_oracle_stage_modulation(s, x.ctypes.data, m.ctypes.data, l.ctypes.data)
        assert_testpeaq_close(m, golden_vectors["modulation%d_ref" % k], "modulation%d" % k)
        assert_testpeaq_close(l, golden_vectors["loudness%d_ref" % k], "loudness%d" % k)
    L.peaq_oracle_stage_free(s)


def test_known_answer_odgs():
    """runtest-1.0.sh: 0.171 for sine vs itself, -2.007 for saw vs triangle,
    mono and stereo (printed with %.3f)"""
    cases = golden_cases()
    r = H.oracle_run_pair(*cases["kat_sine_sine_mono"][:2], channels=1)
    assert "%.3f" % r["odg"] == "0.171"
    r = H.oracle_run_pair(*cases["kat_saw_tri_mono"][:2], channels=1)
    assert "%.3f" % r["odg"] == "-2.007"
    r = H.oracle_run_pair(*cases["kat_saw_tri_stereo"][:2], channels=2)
    assert "%.3f" % r["odg"] == "-2.007"


def _unpack(v):
    n = int(v[6])
    return {"odg": v[0], "di": v[1], "totalsnr": v[2], "frames_fft": int(v[3]), "frames_fb": int(v[4]),
            "loudness_reached_frame": int(v[5]), "movs": v[7:7 + n]}


def assert_results_close(got, want, rtol, what):
    assert got["frames_fft"] == want["frames_fft"], what
    assert got["frames_fb"] == want["frames_fb"], what
    assert got["loudness_reached_frame"] == want["loudness_reached_frame"], what
    np.testing.assert_allclose(got["movs"], want["movs"], rtol=rtol, atol=1e-12, equal_nan=True, err_msg=what)
    np.testing.assert_allclose([got["di"], got["odg"]], [want["di"], want["odg"]], rtol=0, atol=max(rtol, 1e-12) * 10,
                               equal_nan=True, err_msg=what)
    if math.isfinite(want["totalsnr"]):
        assert abs(got["totalsnr"] - want["totalsnr"]) < 1e-9, what
    else:
        assert got["totalsnr"] == want["totalsnr"] or (math.isnan(got["totalsnr"]) and math.isnan(want["totalsnr"]))


@pytest.mark.parametrize("advanced", [False, True])
def test_oracle_matches_committed_reference_outputs(golden_ref_outputs, advanced):
    for name, (ref, test, ch) in golden_cases().items():
        want = _unpack(golden_ref_outputs["%s|%s" % (name, "advanced" if advanced else "basic")])
        got = H.oracle_run_pair(ref, test, ch, advanced=advanced)
        # same compiler, same FFT: the restatement reproduces the reference to rounding
        assert_results_close(got, want, 1e-12, "%s advanced=%s" % (name, advanced))


needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("advanced", [False, True])
def test_oracle_matches_live_reference_chunked_and_ragged(advanced):
    """arbitrary buffer sizes and unequal stream lengths (pad_chain / do_flush)"""
    rng = np.random.default_rng(7)
    ref, test = synth_pair(21, 30000, 2)
    test = test[:2 * 29000]          # test stream ends early
    r = H.RefPeaq(advanced, 92.0, 2)
    o = H.OraclePeaq(advanced, 92.0, 2)
    pr = pt = 0
    while pr < ref.size or pt < test.size:
        a = 2 * int(rng.integers(1, 3000))
        b = 2 * int(rng.integers(1, 3000))
        cr, ct = ref[pr:pr + a], test[pt:pt + b]
        r.push(cr, ct)
        o.push(cr, ct)
        pr += a
        pt += b
    r.finish()
    o.finish()
    assert_results_close(o.result(), r.result(), 1e-12, "chunked advanced=%s" % advanced)


@needs_ref
def test_oracle_state_matches_live_reference_per_frame():
    """taps through the reference's own accessors after every frame"""
    ref, test = synth_pair(2, 2048 + 1024 * 30, 2)
    r = H.RefPeaq(False, 92.0, 2)
    nf = 31
    o = H.OraclePeaq(False, 92.0, 2, fft_trace=nf)
    frame = 2048 * 2
    step = 1024 * 2
    r.push(ref[:frame], test[:frame])
    o.push(ref[:frame], test[:frame])
    for f in range(nf):
        for side in (0, 1):
            for c in (0, 1):
                np.testing.assert_allclose(o.fft_trace["unsmeared"][f, side, c], r.tap(2, side, c), rtol=1e-13)
                np.testing.assert_allclose(o.fft_trace["excitation"][f, side, c], r.tap(3, side, c), rtol=1e-13)
        lo = frame + f * step
        r.push(ref[lo:lo + step], test[lo:lo + step])
        o.push(ref[lo:lo + step], test[lo:lo + step])


@needs_ref
@pytest.mark.parametrize("advanced", [False, True])
def test_oracle_tables_match_reference(advanced):
    r = H.RefPeaq(advanced, 92.0, 1)
    o = H.OraclePeaq(advanced, 92.0, 1)
    for model in (0, 1):
        for which in range(7):
            a, b = o.table(model, which), r.table(model, which)
            assert a.shape == b.shape
            if a.size:
                np.testing.assert_array_equal(a, b)
    # anchors recorded in SURVEY.md 8c
    fc = o.table(0, 0)
    if advanced:
        assert abs(fc[0] - 103.4454474) < 1e-6
    else:
        assert abs(fc[0] - 91.7081015) < 1e-6 and abs(fc[108] - 17690.04359) < 1e-4
    fb = o.table(1, 0)
    assert abs(fb[0] - 50) < 1e-9 and abs(fb[39] - 18000) < 1e-6
