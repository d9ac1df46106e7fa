"""GPU: the CUDA engine against the CPU oracle, through the C ABI.

Bar (SURVEY 8d / BASELINE north star): integer-valued quantities exact (frame
counts, per-frame bandwidth bins, above-threshold / EHS-valid flags,
loudness-reached frame); floating MOVs rel <= 1e-6; |dODG|, |dDI| <= 1e-4
(expected ~1e-12: the engine is FP64 end to end); known-answer ODGs print
identically at %.3f."""
import math

import numpy as np
import pytest

import gstpeaq_b200 as G
import refharness as H
from signals import golden_cases, noise_pair, synth_pair

pytestmark = pytest.mark.gpu

MOV_RTOL = 1e-6
ODG_ATOL = 1e-4


@pytest.fixture(scope="module")
def engine():
    e = G.Engine(0, advanced=False)
    yield e
    e.close()


def check_result(got, want, what):
    """got: one row of the engine's result array; want: oracle/reference dict"""
    assert int(got["frames_fft"]) == want["frames_fft"], what
    assert int(got["loudness_reached_frame"]) == want["loudness_reached_frame"], what
    n = len(want["movs"])
    assert int(got["n_movs"]) == n
    np.testing.assert_allclose(got["movs"][:n], want["movs"], rtol=MOV_RTOL, atol=1e-9, equal_nan=True,
                               err_msg=what)
    for k in ("di", "odg"):
        if math.isnan(want[k]):
            assert math.isnan(got[k]), what
        else:
            assert abs(got[k] - want[k]) <= ODG_ATOL, (what, k, got[k], want[k])
    if math.isfinite(want["totalsnr"]):
        assert abs(got["totalsnr"] - want["totalsnr"]) < 1e-6, what


def test_device_generator_is_bit_identical_to_host_generator():
    L = G.load_library()
    n_pairs, ns, ch = 3, 30000, 2
    dref = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    dtest = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 5, ns, ch))
    hr = np.zeros((n_pairs, ns * ch), np.float32)
    ht = np.zeros_like(hr)
    G._check(L.peaq_b200_memcpy_d2h(0, hr.ctypes.data, dref.ptr, hr.nbytes))
    G._check(L.peaq_b200_memcpy_d2h(0, ht.ctypes.data, dtest.ptr, ht.nbytes))
    r2, t2 = G.synth_pairs_host(5, n_pairs, ns, ch)
    np.testing.assert_array_equal(hr, r2)
    np.testing.assert_array_equal(ht, t2)


def test_known_answer_odgs_print_like_the_reference(engine):
    """runtest-1.0.sh:18,28,38,48"""
    cases = golden_cases()
    for name, want in (("kat_sine_sine_mono", "0.171"), ("kat_saw_tri_mono", "-2.007"),
                       ("kat_saw_tri_stereo", "-2.007")):
        ref, test, ch = cases[name]
        out = engine.run_host(ref, test, ch)
        assert "%.3f" % out["odg"][0] == want, name


@pytest.mark.parametrize("name", sorted(golden_cases().keys()))
def test_batch_matches_committed_reference_outputs(engine, golden_ref_outputs, name):
    """fixtures = outputs of the reference's own C code (tests/golden/make_golden.py)"""
    ref, test, ch = golden_cases()[name]
    v = golden_ref_outputs[name + "|basic"]
    n = int(v[6])
    want = {"odg": v[0], "di": v[1], "totalsnr": v[2], "frames_fft": int(v[3]),
            "loudness_reached_frame": int(v[5]), "movs": v[7:7 + n]}
    out = engine.run_host(ref, test, ch)
    check_result(out[0], want, name)


@pytest.mark.parametrize("name", ["synth0_stereo", "noise_silence_stereo", "kat_saw_tri_mono", "synth5_mono"])
def test_per_frame_records_match_oracle(engine, name):
    ref, test, ch = golden_cases()[name]
    nf = G.frames_for_samples(ref.size // ch)
    o = H.OraclePeaq(False, 92.0, ch, fft_trace=nf)
    o.run(ref, test)
    tr = o.fft_trace
    engine.keep_records(True)
    try:
        engine.run_host(ref, test, ch)
        rec = engine.records(1, nf)
    finally:
        engine.keep_records(False)
    B = 109
    # integer-valued: exact
    np.testing.assert_array_equal(rec["flags"][0] & 1, tr["above_threshold"])
    np.testing.assert_array_equal((rec["flags"][0] >> 1) & 1, tr["ehs_valid"])
    np.testing.assert_array_equal(rec["bw_ref"][0], tr["bw_ref"][:, :ch])
    np.testing.assert_array_equal(rec["bw_test"][0], tr["bw_test"][:, :ch])
    # floating point
    np.testing.assert_allclose(rec["unsmeared"][0], tr["unsmeared"][:, :, :ch, :B], rtol=1e-10)
    # noise in bands = P_ref - 2 sqrt(P_ref P_test) + P_test cancels heavily: compare
    # against the band's signal power scale as well
    np.testing.assert_allclose(rec["noise_in_bands"][0], tr["noise_in_bands"][:, :ch, :B], rtol=1e-4, atol=1e-9)
    valid = tr["ehs_valid"].astype(bool)
    np.testing.assert_allclose(rec["ehs"][0][valid], tr["ehs"][valid][:, :ch], rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(np.cumsum(rec["snr"][0][:, 0]), tr["signal_energy"], rtol=1e-12)
    np.testing.assert_allclose(np.cumsum(rec["snr"][0][:, 1]), tr["noise_energy"], rtol=1e-12)


@pytest.mark.parametrize("name", ["synth0_stereo", "synth5_mono", "noise_silence_stereo", "noise_loud_mono"])
def test_per_frame_scan_terms_match_oracle(engine, name):
    """K2's per-frame taps against the oracle's trace: time-smeared excitation patterns, and the
    band-summed terms the MOV functions hand to the accumulators (modulation differences and
    temporal weight from frame 24 on, noise loudness once the loudness latch + 3 frames allows,
    N/M mean and maximum, binaural detection probability and steps)."""
    ref, test, ch = golden_cases()[name]
    nf = G.frames_for_samples(ref.size // ch)
    o = H.OraclePeaq(False, 92.0, ch, fft_trace=nf)
    o.run(ref, test)
    tr = o.fft_trace
    engine.keep_records(True)
    try:
        out = engine.run_host(ref, test, ch)
        exc, terms = engine.scan_debug(1, ch)
    finally:
        engine.keep_records(False)
    B = 109
    exc, terms = exc[0, :nf], terms[0, :nf]
    np.testing.assert_allclose(exc, tr["excitation"][:, :, :ch, :B], rtol=1e-10)
    f = np.arange(nf)
    md = f >= 24
    np.testing.assert_allclose(terms[md, :, 0], tr["mod_diff1"][md][:, :ch], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(terms[md, :, 1], tr["mod_diff2"][md][:, :ch], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(terms[md, :, 2], tr["temp_wt"][md][:, :ch], rtol=1e-9)
    loud = int(out["loudness_reached_frame"][0])
    nl = md & (f.astype(np.int64) - 3 >= loud)
    if nl.any():
        np.testing.assert_allclose(terms[nl, :, 3], tr["noise_loud"][nl][:, :ch], rtol=1e-8, atol=1e-11)
    # the noise-to-mask ratios inherit the cancellation of noise_in_bands (see the record test)
    np.testing.assert_allclose(terms[:, :, 4], tr["nmr"][:, :ch], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(terms[:, :, 5], tr["nmr_max"][:, :ch], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(terms[:, 0, 6], tr["det_prob"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(terms[:, 0, 7], tr["adb_steps"], rtol=1e-8, atol=1e-9)


def test_batch_of_ragged_synthetic_pairs_matches_oracle(engine):
    """one call, pairs of different lengths (incl. empty and sub-frame items)"""
    ch = 2
    lengths = [48000, 30001, 2048, 1, 0, 70000, 4096 + 512, 1023]
    stride = max(lengths) * ch
    ref = np.zeros((len(lengths), stride), np.float32)
    test = np.zeros_like(ref)
    for p, n in enumerate(lengths):
        if n:
            r, t = synth_pair(100 + p, n, ch)
            ref[p, :n * ch] = r
            test[p, :n * ch] = t
    out = engine.run_host(ref, test, ch, n_samples=np.array(lengths, np.uint64))
    for p, n in enumerate(lengths):
        want = H.oracle_run_pair(ref[p, :n * ch], test[p, :n * ch], ch)
        check_result(out[p], want, "pair %d len %d" % (p, n))


def test_chunked_frame_loop_is_bit_identical(engine, monkeypatch):
    """records budget forces several K1/K2 chunks: the recurrent state handed
    over in memory must give bit-identical results to the single-chunk run"""
    ch = 2
    r, t = G.synth_pairs_host(40, 6, 60000, ch)
    a = engine.run_host(r, t, ch)
    monkeypatch.setenv("PEAQ_B200_RECORD_BUDGET_MB", "1")
    e2 = G.Engine(0, advanced=False)
    try:
        b = e2.run_host(r, t, ch)
    finally:
        e2.close()
    for k in ("odg", "di", "totalsnr", "movs", "frames_fft", "loudness_reached_frame"):
        np.testing.assert_array_equal(a[k], b[k])


def test_device_resident_batch_equals_host_batch(engine):
    L = G.load_library()
    n_pairs, ns, ch = 5, 40000, 2
    dref = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    dtest = G.DeviceBuffer(0, n_pairs * ns * ch * 4)
    G._check(L.peaq_b200_synth_pairs(0, dref.ptr, dtest.ptr, ns * ch, n_pairs, 9, ns, ch))
    a = engine.run_device(dref.ptr, dtest.ptr, n_pairs, ns * ch, ch, ns)
    r, t = G.synth_pairs_host(9, n_pairs, ns, ch)
    b = engine.run_host(r, t, ch)
    np.testing.assert_array_equal(a["odg"], b["odg"])
    np.testing.assert_array_equal(a["movs"], b["movs"])
    assert engine.launch_count() > 0


def test_session_streaming_matches_oracle_with_arbitrary_buffers():
    """element surface: caps, chain on both pads with arbitrary buffer sizes,
    unequal stream lengths, stop -> odg/di (gstpeaq.c:614-661, :716-745, :764-778)"""
    rng = np.random.default_rng(3)
    ch = 2
    ref, test = synth_pair(21, 30000, ch)
    test = test[:ch * 29000]
    o = H.OraclePeaq(False, 92.0, ch)
    p = G.Peaq(0, console_output=False)
    p.set_caps(ch)
    assert math.isnan(p.odg)          # nothing processed yet: 0/0 like the reference
    pr = pt = 0
    while pr < ref.size or pt < test.size:
        a = ch * int(rng.integers(1, 4000))
        b = ch * int(rng.integers(1, 4000))
        cr, ct = ref[pr:pr + a], test[pt:pt + b]
        p.chain_ref(cr)
        p.chain_test(ct)
        o.push(cr, ct)
        pr += a
        pt += b
        mid = p.result()
        want_mid = o.result()
        assert mid["frames_fft"] == want_mid["frames_fft"]
    res = p.stop()
    o.finish()
    want = o.result()
    got = {"frames_fft": res["frames_fft"], "loudness_reached_frame": res["loudness_reached_frame"],
           "n_movs": 11, "movs": np.concatenate([res["movs"], np.zeros(0)]), "di": res["di"], "odg": res["odg"],
           "totalsnr": res["totalsnr"]}
    check_result(got, want, "session")
    assert abs(p.odg - want["odg"]) < ODG_ATOL and abs(p.di - want["di"]) < ODG_ATOL
    p.close()


def test_session_console_output_format(capsys):
    ref, test, ch = golden_cases()["kat_saw_tri_mono"]
    p = G.Peaq(0, console_output=True)
    p.set_caps(ch)
    p.chain_ref(ref)
    p.chain_test(test)
    p.stop()
    out = capsys.readouterr().out
    assert out.splitlines()[0].startswith("   BandwidthRefB: 921.000000")
    assert out.splitlines()[-1] == "Objective Difference Grade: -2.007"
    p.close()


def test_linearity_property_playback_level():
    """size-independent property: +6.0206 dB playback level == input scaled by 2
    (level enters only through the FFT level factor, fftearmodel.c:304-314)"""
    ch = 2
    r, t = G.synth_pairs_host(77, 2, 40000, ch)
    e1 = G.Engine(0, False, 92.0 + 20 * math.log10(2.0))
    e2 = G.Engine(0, False, 92.0)
    try:
        a = e1.run_host(r * np.float32(0.5), t * np.float32(0.5), ch)
        b = e2.run_host(r, t, ch)
    finally:
        e1.close()
        e2.close()
    # thresholds on raw samples (above-threshold, energy flag) are level independent
    # here because the signals are loud; MOVs agree to rounding
    np.testing.assert_allclose(a["movs"], b["movs"], rtol=1e-9)


# ---------------------------------------------------------------------------
# advanced mode (filter-bank ear model + 55-band FFT model, 5 MOVs)

@pytest.fixture(scope="module")
def engine_adv():
    e = G.Engine(0, advanced=True)
    yield e
    e.close()


@pytest.mark.parametrize("name", sorted(golden_cases().keys()))
def test_advanced_batch_matches_committed_reference_outputs(engine_adv, golden_ref_outputs, name):
    ref, test, ch = golden_cases()[name]
    v = golden_ref_outputs[name + "|advanced"]
    n = int(v[6])
    want = {"odg": v[0], "di": v[1], "totalsnr": v[2], "frames_fft": int(v[3]), "frames_fb": int(v[4]),
            "loudness_reached_frame": int(v[5]), "movs": v[7:7 + n]}
    out = engine_adv.run_host(ref, test, ch)
    assert int(out["frames_fb"][0]) == want["frames_fb"]
    check_result(out[0], want, name)


@pytest.mark.parametrize("name", ["synth0_stereo", "noise_silence_stereo", "synth5_mono"])
def test_advanced_per_frame_filter_bank_matches_oracle(engine_adv, name):
    ref, test, ch = golden_cases()[name]
    n = ref.size // ch
    nfb = (n + 191) // 192
    o = H.OraclePeaq(True, 92.0, ch, fb_trace=nfb)
    o.run(ref, test)
    tr = o.fb_trace
    engine_adv.keep_records(True)
    try:
        engine_adv.run_host(ref, test, ch)
        exc, movs = engine_adv.fb_debug(1, ch)
    finally:
        engine_adv.keep_records(False)
    np.testing.assert_array_equal(movs[0][:nfb, 0, 5], tr["above_threshold"])     # integer: exact
    for c in range(ch):
        for side in range(2):
            np.testing.assert_allclose(exc[0][:nfb, 2 * c + side, 0], tr["unsmeared"][:, side, c], rtol=1e-9)
            np.testing.assert_allclose(exc[0][:nfb, 2 * c + side, 1], tr["excitation"][:, side, c], rtol=1e-9)
    for k, nm in enumerate(["mod_diff", "temp_wt", "noise_loud", "missing_comp", "lin_dist"]):
        np.testing.assert_allclose(movs[0][:nfb, :, k], tr[nm][:, :ch], rtol=1e-7, atol=1e-9, err_msg=nm)


def test_advanced_ragged_batch_and_chunking(engine_adv, monkeypatch):
    ch = 2
    lengths = [48000, 30001, 192, 1, 0, 40000, 191, 2048]
    stride = max(lengths) * ch
    ref = np.zeros((len(lengths), stride), np.float32)
    test = np.zeros_like(ref)
    for p, n in enumerate(lengths):
        if n:
            r, t = synth_pair(200 + p, n, ch)
            ref[p, :n * ch] = r
            test[p, :n * ch] = t
    ns = np.array(lengths, np.uint64)
    out = engine_adv.run_host(ref, test, ch, n_samples=ns)
    for p, n in enumerate(lengths):
        want = H.oracle_run_pair(ref[p, :n * ch], test[p, :n * ch], ch, advanced=True)
        assert int(out["frames_fb"][p]) == want["frames_fb"]
        check_result(out[p], want, "adv pair %d len %d" % (p, n))
    monkeypatch.setenv("PEAQ_B200_FB_BUDGET_MB", "1")
    monkeypatch.setenv("PEAQ_B200_RECORD_BUDGET_MB", "1")
    e2 = G.Engine(0, advanced=True)
    try:
        b = e2.run_host(ref, test, ch, n_samples=ns)
    finally:
        e2.close()
    for k in ("odg", "di", "movs", "frames_fft", "frames_fb", "loudness_reached_frame"):
        np.testing.assert_array_equal(out[k], b[k])


def test_advanced_session_matches_oracle_with_unequal_streams():
    """element with advanced=TRUE: arbitrary buffers, test stream ends early"""
    rng = np.random.default_rng(5)
    ch = 2
    ref, test = synth_pair(31, 26000, ch)
    test = test[:ch * 24500]
    o = H.OraclePeaq(True, 92.0, ch)
    p = G.Peaq(0, advanced=True, console_output=False)
    p.set_caps(ch)
    pr = pt = 0
    k = 0
    while pr < ref.size or pt < test.size:
        a = ch * int(rng.integers(1, 6000))
        b = ch * int(rng.integers(1, 6000))
        cr, ct = ref[pr:pr + a], test[pt:pt + b]
        p.chain_ref(cr)
        p.chain_test(ct)
        o.push(cr, ct)
        pr += a
        pt += b
        k += 1
        if k == 3:      # mid-stream property read
            mid, want_mid = p.result(), o.result()
            assert mid["frames_fft"] == want_mid["frames_fft"] and mid["frames_fb"] == want_mid["frames_fb"]
    res = p.stop()
    o.finish()
    want = o.result()
    assert res["frames_fb"] == want["frames_fb"]
    got = {"frames_fft": res["frames_fft"], "loudness_reached_frame": res["loudness_reached_frame"],
           "n_movs": 5, "movs": res["movs"], "di": res["di"], "odg": res["odg"], "totalsnr": res["totalsnr"]}
    check_result(got, want, "advanced session")
    p.close()


# ---------------------------------------------------------------------------
# long items / batch properties (BASELINE configs[4]: hour-long pairs run as a
# sequence of chunks with the recurrent state handed over in HBM)

def test_long_item_chunked_matches_oracle(monkeypatch):
    """2 minutes of audio, records budget of 8 MiB => dozens of K1/K2 chunks"""
    ch = 2
    n = 48000 * 120
    r, t = G.synth_pairs_host(500, 1, n, ch)
    monkeypatch.setenv("PEAQ_B200_RECORD_BUDGET_MB", "8")
    e = G.Engine(0, advanced=False)
    try:
        out = e.run_host(r, t, ch)
    finally:
        e.close()
    want = H.oracle_run_pair(r[0], t[0], ch)
    assert int(out["frames_fft"][0]) == G.frames_for_samples(n) == want["frames_fft"]
    check_result(out[0], want, "2 min item")


def test_long_item_advanced_chunked_matches_oracle(monkeypatch):
    ch = 2
    n = 48000 * 30
    r, t = G.synth_pairs_host(501, 1, n, ch)
    monkeypatch.setenv("PEAQ_B200_FB_BUDGET_MB", "4")
    monkeypatch.setenv("PEAQ_B200_RECORD_BUDGET_MB", "4")
    e = G.Engine(0, advanced=True)
    try:
        out = e.run_host(r, t, ch)
    finally:
        e.close()
    want = H.oracle_run_pair(r[0], t[0], ch, advanced=True)
    assert int(out["frames_fb"][0]) == want["frames_fb"] == 7500
    check_result(out[0], want, "30 s advanced item")


def test_batch_properties_order_and_replication(engine):
    """size-independent properties: results do not depend on the position of a
    pair in the batch, nor on its neighbours (pairs are closed computations)"""
    ch = 2
    r, t = G.synth_pairs_host(600, 7, 30000, ch)
    a = engine.run_host(r, t, ch)
    perm = np.array([3, 0, 6, 1, 5, 2, 4])
    b = engine.run_host(r[perm], t[perm], ch)
    np.testing.assert_array_equal(a["movs"][perm], b["movs"])
    np.testing.assert_array_equal(a["odg"][perm], b["odg"])
    rep = engine.run_host(np.repeat(r[:1], 64, axis=0), np.repeat(t[:1], 64, axis=0), ch)
    assert np.all(rep["odg"] == a["odg"][0]) and np.all(rep["movs"] == a["movs"][0])
    # identical signals: noise floor only (NMR at its -118.7 dB class floor, detection probability 0)
    same = engine.run_host(r[:2], r[:2], ch)
    assert np.all(same["movs"][:, 9] == 0) and np.all(same["movs"][:, 2] < -100)


@pytest.mark.parametrize("advanced", [False, True])
def test_session_tiny_buffers_and_midstream_reads(advanced):
    """buffers much smaller than a frame (and than the FIR history of the
    filter bank); every few pushes the running result must equal the
    reference's running value (properties are readable at any time)"""
    rng = np.random.default_rng(11)
    ch = 2
    ref, test = synth_pair(41, 14000, ch)
    o = H.OraclePeaq(advanced, 92.0, ch)
    p = G.Peaq(0, advanced=advanced, console_output=False)
    p.set_caps(ch)
    pos = 0
    k = 0
    while pos < ref.size:
        a = ch * int(rng.integers(50, 900))
        cr, ct = ref[pos:pos + a], test[pos:pos + a]
        p.chain_ref(cr)
        p.chain_test(ct)
        o.push(cr, ct)
        pos += a
        k += 1
        if k % 7 == 0:
            mid, want = p.result(), o.result()
            if want["frames_fft"] or want["frames_fb"]:
                assert mid["frames_fft"] == want["frames_fft"] and mid["frames_fb"] == want["frames_fb"]
                np.testing.assert_allclose(mid["movs"], want["movs"], rtol=MOV_RTOL, atol=1e-9, equal_nan=True)
    res = p.stop()
    o.finish()
    want = o.result()
    np.testing.assert_allclose(res["movs"], want["movs"], rtol=MOV_RTOL, atol=1e-9, equal_nan=True)
    assert res["frames_fft"] == want["frames_fft"] and res["frames_fb"] == want["frames_fb"]
    p.close()


def test_filter_bank_recursion_against_direct_fir():
    """the sliding-window-DFT form of the 40 filters (fb_bank_rec_kernel) against
    the polyphase direct-FIR kernel (PEAQ_B200_FB_DIRECT=1, read once per
    process, hence the subprocess): per-frame excitations and MOVs"""
    import json
    import os
    import subprocess
    import sys
    code = r'''
import json, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gstpeaq_b200 as G
from signals import synth_pair
ch = 2
ref, test = synth_pair(77, 30000, ch)
e = G.Engine(0, advanced=True)
e.keep_records(True)
out = e.run_host(ref, test, ch)
exc, movs = e.fb_debug(1, ch)
print(json.dumps({"movs": out["movs"][0][:5].tolist(), "odg": float(out["odg"][0]),
                  "exc": np.asarray(exc[0]).ravel().tolist()}))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for direct in ("0", "1"):
        env = dict(os.environ, PEAQ_B200_FB_DIRECT=direct)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res[direct] = json.loads(p.stdout.strip().splitlines()[-1])
    a, b = res["0"], res["1"]
    np.testing.assert_allclose(a["exc"], b["exc"], rtol=4e-9)
    np.testing.assert_allclose(a["movs"], b["movs"], rtol=1e-9, atol=1e-12)
    assert abs(a["odg"] - b["odg"]) < 1e-9


def test_filter_bank_coefficient_sources_same_bits():
    """fb_bank_rec_kernel reads the entering side's coefficients from constant memory (uniform
    datapath) by default and from shared memory with PEAQ_B200_FB_SMEM_COEF=1 (read once per
    process, hence the subprocess): same products in the same order, so the per-frame excitations
    and the results must be the same bits -- also for a second engine of the process (the constant
    table is uploaded once per device)"""
    import json
    import os
    import subprocess
    import sys
    code = r'''
import json, sys, hashlib, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gstpeaq_b200 as G
from signals import synth_pair
ch = 2
ref, test = synth_pair(91, 40000, ch)
digest = []
for level in (92.0, 80.0):
    e = G.Engine(0, advanced=True, playback_level=level)
    e.keep_records(True)
    out = e.run_host(ref, test, ch)
    exc, movs = e.fb_debug(1, ch)
    digest.append(hashlib.sha256(np.asarray(exc[0]).tobytes() + out.tobytes()).hexdigest())
    e.close()
print(json.dumps(digest))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for smem in ("0", "1"):
        env = dict(os.environ, PEAQ_B200_FB_SMEM_COEF=smem)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res[smem] = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["0"] == res["1"]
    assert res["0"][0] != res["0"][1]   # the two playback levels do differ


@pytest.mark.parametrize("advanced", [False, True])
def test_unaligned_mono_batch_matches_oracle(advanced):
    """mono pairs at an odd stride: every pair but the first starts off the 16-byte grid, so the
    frame kernel's guarded staging path (no TMA) and the scalar loads of the filter-bank input
    are the ones that run"""
    ch = 1
    lengths = [30001, 29999, 30001]
    stride = 30001
    ref = np.zeros((len(lengths), stride), np.float32)
    test = np.zeros_like(ref)
    for p, n in enumerate(lengths):
        r, t = synth_pair(300 + p, n, ch)
        ref[p, :n] = r
        test[p, :n] = t
    e = G.Engine(0, advanced=advanced)
    try:
        out = e.run_host(ref, test, ch, n_samples=np.array(lengths, np.uint64))
    finally:
        e.close()
    for p, n in enumerate(lengths):
        want = H.oracle_run_pair(ref[p, :n], test[p, :n], ch, advanced=advanced)
        check_result(out[p], want, "mono pair %d len %d adv %d" % (p, n, advanced))


def test_dc_reject_block_scan_same_bits_sequential_and_parallel():
    """FB1 is a block scan over absolute 512-sample blocks with two implementations -- one thread
    per stream walking the samples (sessions, big batches) and blocks in parallel (few long
    items).  Both must produce the SAME bits, whatever the chunking: per-frame excitations of a
    ragged batch, whole-item and chunked, forced through either kernel; and the oracle's (exact
    sequential) excitations are matched to 5e-9 (the block scan's rounding sequence differs)."""
    import json
    import os
    import subprocess
    import sys
    code = r'''
import json, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gstpeaq_b200 as G
from signals import synth_pair
ch = 2
lengths = [96000, 70001, 50000]
stride = max(lengths) * ch
ref = np.zeros((3, stride), np.float32); test = np.zeros_like(ref)
for p, n in enumerate(lengths):
    r, t = synth_pair(77 + p, n, ch)
    ref[p, :n * ch] = r; test[p, :n * ch] = t
e = G.Engine(0, advanced=True)
e.keep_records(True)
out = e.run_host(ref, test, ch, n_samples=np.array(lengths, np.uint64))
exc, movs = e.fb_debug(3, ch)
e2 = G.Engine(0, advanced=True)      # chunked (PEAQ_B200_FB_BUDGET_MB from the environment applies to both)
out2 = e2.run_host(ref, test, ch, n_samples=np.array(lengths, np.uint64))
print(json.dumps({"movs": out["movs"][:, :5].tolist(), "odg": out["odg"].tolist(), "odg2": out2["odg"].tolist(),
                  "exc": np.asarray(exc).ravel().tolist()}))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for par, budget in (("0", ""), ("1", ""), ("1", "1"), ("0", "1")):
        env = dict(os.environ, PEAQ_B200_HP_PARALLEL=par)
        if budget:
            env["PEAQ_B200_FB_BUDGET_MB"] = budget      # a few frames per chunk
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res[par + budget] = json.loads(p.stdout.strip().splitlines()[-1])
    a = res["0"]
    for k in ("1", "11", "01"):
        assert res[k]["odg"] == a["odg"] and res[k]["movs"] == a["movs"], k
        assert res[k]["odg2"] == a["odg"], k
    assert res["1"]["exc"] == a["exc"]
    # against the oracle's sequential recurrence
    ch = 2
    r, t = synth_pair(77, 96000, ch)
    o = H.OraclePeaq(True, 92.0, ch, fb_trace=600)
    o.push(r, t)
    o.finish()
    nfb = o.result()["frames_fb"]
    exc = np.asarray(a["exc"]).reshape(3, -1, 2 * ch, 2, 40)[0, :nfb]
    want = o.fb_trace["excitation"][:nfb]           # [frame][side][channel][band]
    got = exc[:, :, 1, :].reshape(nfb, ch, 2, 40).transpose(0, 2, 1, 3)    # stream = 2 * channel + side
    np.testing.assert_allclose(got, want, rtol=5e-9)


def _engine_with_env(monkeypatch, advanced=False, **env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    e = G.Engine(0, advanced=advanced)
    for k in env:
        monkeypatch.delenv(k)
    return e


@pytest.mark.parametrize("ch", [2, 1])
def test_fused_persistent_kernel_is_bit_identical_to_two_kernel_path(monkeypatch, ch):
    """the fused kernel (one persistent CTA per pair, nothing per-frame through HBM) must give
    the same bits as K1 + records + K2: ragged lengths (incl. empty, sub-frame, exact multiples),
    leading / trailing silence (INIT / TENTATIVE accumulators), stereo and mono (TMA prefetch on / off)"""
    lengths = [48000, 30001, 2048, 1, 0, 70000, 4096 + 512, 1023, 2048 + 1024 * 7, 65536]
    stride = max(lengths) * ch
    ref = np.zeros((len(lengths) + 1, stride), np.float32)
    test = np.zeros_like(ref)
    for p, n in enumerate(lengths):
        if n:
            r, t = synth_pair(300 + p, n, ch)
            ref[p, :n * ch] = r
            test[p, :n * ch] = t
    x, y = noise_pair(7, 48765, ch, lead=20000, tail=16000)
    ref[-1, :x.size] = x
    test[-1, :y.size] = y
    ns = np.array(lengths + [48765], np.uint64)
    fused = _engine_with_env(monkeypatch, PEAQ_B200_FUSED="1")
    split = _engine_with_env(monkeypatch, PEAQ_B200_FUSED="0")
    try:
        a = fused.run_host(ref, test, ch, n_samples=ns)
        b = split.run_host(ref, test, ch, n_samples=ns)
        assert fused.last_ms(2) == 0.0 and split.last_ms(2) > 0.0   # really two different paths
    finally:
        fused.close()
        split.close()
    assert a.tobytes() == b.tobytes()
    for p, n in enumerate(ns):
        want = H.oracle_run_pair(ref[p, :int(n) * ch], test[p, :int(n) * ch], ch)
        check_result(a[p], want, "fused pair %d len %d" % (p, n))


def test_threshold_replay_amplitude_sweep(engine):
    """is_frame_above_threshold (gstpeaq.c:1081-1099): amplitudes around 200/32768 / {5, 1} put
    frames into the window where K1 cannot decide from the largest |x| and replays the reference's
    float recurrence literally; flags must equal the oracle's frame by frame"""
    ch = 1
    n = 2048 + 1024 * 6
    thr = 200.0 / 32768.0
    rng = np.random.default_rng(11)
    rows_r, rows_t, amps = [], [], []
    for base in (thr / 5, thr):
        for f in (0.90, 0.985, 0.995, 0.99999, 0.999999, 1.0, 1.000001, 1.00001, 1.004, 1.009, 1.02, 1.2):
            amp = np.float32(base * f)
            sign = np.where(rng.random(n) < 0.5, -1.0, 1.0).astype(np.float32)
            r = (sign * amp).astype(np.float32)                  # constant magnitude: sliding sums sit at 5 * amp
            r[::7] *= np.float32(0.5)
            t = (r * np.float32(0.9)).astype(np.float32)
            rows_r.append(r)
            rows_t.append(t)
            amps.append(float(amp))
    ref = np.stack(rows_r)
    test = np.stack(rows_t)
    engine.keep_records(True)
    try:
        out = engine.run_host(ref, test, ch)
        rec = engine.records(len(amps), G.frames_for_samples(n))
    finally:
        engine.keep_records(False)
    seen = set()
    for p in range(len(amps)):
        o = H.OraclePeaq(False, 92.0, ch, fft_trace=16)
        o.push(ref[p], test[p])
        o.finish()
        nf = G.frames_for_samples(n)
        want = o.fft_trace["above_threshold"][:nf]
        got = rec["flags"][p, :nf] & 1
        np.testing.assert_array_equal(got, want, err_msg="amp %g" % amps[p])
        seen.update(int(v) for v in want)
        check_result(out[p], o.result(), "amp %g" % amps[p])
    assert seen == {0, 1}   # the sweep really crosses the threshold


@pytest.mark.parametrize("advanced", [False, True])
def test_bench_shaped_batch_matches_oracle(advanced):
    """the bench workload's shape -- 10 s synthetic stereo pairs, one batch call -- against the oracle
    (24 pairs basic, 8 advanced: the oracle's filter bank runs ~0.7 s per 10 s pair)"""
    ch, ns = 2, 480000
    n_pairs = 8 if advanced else 24
    ref, test = G.synth_pairs_host(2000, n_pairs, ns, ch)
    e = G.Engine(0, advanced=advanced)
    try:
        out = e.run_host(ref, test, ch)
    finally:
        e.close()
    for p in range(n_pairs):
        want = H.oracle_run_pair(ref[p], test[p], ch, advanced=advanced)
        check_result(out[p], want, "pair %d" % p)


def _stream(p, ref, test, ch, lo, hi, step):
    for a in range(lo, hi, step):
        b = min(a + step, hi)
        p.chain_ref(ref[a * ch:b * ch])
        p.chain_test(test[a * ch:b * ch])


@pytest.mark.parametrize("advanced", [False, True])
def test_session_snapshot_restore_continues_bit_identically(advanced):
    """a snapshot taken mid-stream (samples waiting in the adapters + recurrent
    state on the device) restored into ANOTHER session continues to the same bits as the
    uninterrupted session; the snapshotted session itself is not disturbed either"""
    ch, n = 2, 70000
    ref, test = synth_pair(77, n, ch)
    cut = 33333                       # not a multiple of any frame size: both adapters hold samples
    whole = G.Peaq(0, advanced=advanced, console_output=False)
    whole.set_caps(ch)
    _stream(whole, ref, test, ch, 0, n, 5000)
    want = whole.stop()
    whole.close()

    a = G.Peaq(0, advanced=advanced, console_output=False)
    a.set_caps(ch)
    _stream(a, ref, test, ch, 0, cut, 5000)
    blob = a.snapshot()
    assert len(blob) > 1000
    b = G.Peaq(0, console_output=False)          # mode, level and channels come from the snapshot
    b.restore(blob)
    assert b.advanced == advanced
    _stream(b, ref, test, ch, cut, n, 7000)
    got_b = b.stop()
    _stream(a, ref, test, ch, cut, n, 3000)      # the original continues as well
    got_a = a.stop()
    a.close()
    b.close()
    for got in (got_a, got_b):
        for k in ("odg", "di", "totalsnr", "frames_fft", "frames_fb", "loudness_reached_frame"):
            assert got[k] == want[k] or (math.isnan(got[k]) and math.isnan(want[k])), k
        np.testing.assert_array_equal(got["movs"], want["movs"])
    with pytest.raises(G.PeaqError):
        b2 = G.Peaq(0, console_output=False)
        try:
            b2.restore(blob[:100])
        finally:
            b2.close()


def test_session_pushes_do_not_wait_for_the_gpu():
    """pad_chain returns at once (gstpeaq.c:660); here a push only queues copies and kernels --
    results are bit-identical to a session that reads the result after every push (= synchronises)"""
    ch, n = 2, 120000
    ref, test = synth_pair(5, n, ch)
    res = []
    for read_every_push in (False, True):
        p = G.Peaq(0, console_output=False)
        p.set_caps(ch)
        for a in range(0, n, 1500):              # 80 pushes per pad: more than the 64 the session lets queue up
            p.chain_ref(ref[a * ch:(a + 1500) * ch])
            p.chain_test(test[a * ch:(a + 1500) * ch])
            if read_every_push:
                p.odg
        res.append(p.stop())
        p.close()
    assert res[0]["odg"] == res[1]["odg"] and res[0]["di"] == res[1]["di"]
    np.testing.assert_array_equal(res[0]["movs"], res[1]["movs"])
    want = H.oracle_run_pair(ref, test, ch)
    assert abs(res[0]["odg"] - want["odg"]) < ODG_ATOL


def test_session_playback_level_can_change_mid_stream():
    """property `playback_level` is writable at any time and keeps the model state (gstpeaq.c:509-514)"""
    ch, n = 1, 60000
    ref, test = synth_pair(8, n, ch)

    def run(levels):
        p = G.Peaq(0, console_output=False, playback_level=levels[0])
        p.set_caps(ch)
        _stream(p, ref, test, ch, 0, n // 2, 4000)
        p.playback_level = levels[1]
        assert p.playback_level == levels[1]
        _stream(p, ref, test, ch, n // 2, n, 4000)
        r = p.stop()
        p.close()
        return r

    same = run((92.0, 92.0))
    want = H.oracle_run_pair(ref, test, ch)
    assert abs(same["odg"] - want["odg"]) < ODG_ATOL and same["frames_fft"] == want["frames_fft"]
    mixed, low = run((92.0, 70.0)), run((70.0, 70.0))
    assert same["frames_fft"] == mixed["frames_fft"] == low["frames_fft"]     # no restart of the stream
    assert math.isfinite(mixed["odg"]) and mixed["odg"] != same["odg"] and mixed["odg"] != low["odg"]


def test_multi_engine_shards_a_batch_over_devices():
    """peaq_b200_multi: contiguous blocks of pairs per device, rows gathered into one host array;
    with one GPU in the box both engines sit on device 0, which exercises the same host logic"""
    ch = 2
    lengths = [48000, 30001, 2048, 70000, 1023, 0, 5000]
    stride = max(lengths) * ch
    ref = np.zeros((len(lengths), stride), np.float32)
    test = np.zeros_like(ref)
    for p, n in enumerate(lengths):
        if n:
            r, t = synth_pair(500 + p, n, ch)
            ref[p, :n * ch] = r
            test[p, :n * ch] = t
    ns = np.array(lengths, np.uint64)
    n_dev = G.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    m = G.MultiEngine(devices)
    e = G.Engine(0)
    try:
        assert m.device_count() == len(devices)
        a = m.run_host(ref, test, ch, n_samples=ns)
        b = e.run_host(ref, test, ch, n_samples=ns)
        one = m.run_host(ref[:1], test[:1], ch, n_samples=ns[:1])     # fewer pairs than devices
    finally:
        m.close()
        e.close()
    assert a.tobytes() == b.tobytes()
    assert one.tobytes() == b[:1].tobytes()


def test_pipelined_chunks_are_bit_identical_to_the_serial_loop(monkeypatch):
    """few-pair batches run K2 of chunk i underneath K1 of chunk i + 1 on a second stream
    (double-buffered records): same bits as the serial K1 -> K2 loop, ragged lengths included"""
    ch = 2
    lengths = [400000, 262144 + 2048, 300001, 1000, 0, 350000]
    stride = max(lengths) * ch
    ref = np.zeros((len(lengths), stride), np.float32)
    test = np.zeros_like(ref)
    for p, n in enumerate(lengths):
        if n:
            r, t = synth_pair(700 + p, n, ch)
            ref[p, :n * ch] = r
            test[p, :n * ch] = t
    ns = np.array(lengths, np.uint64)
    piped = _engine_with_env(monkeypatch, PEAQ_B200_PIPELINE="1", PEAQ_B200_RECORD_BUDGET_MB="8")
    serial = _engine_with_env(monkeypatch, PEAQ_B200_PIPELINE="0")
    try:
        a = piped.run_host(ref, test, ch, n_samples=ns)
        a2 = piped.run_host(ref, test, ch, n_samples=ns)     # and again: events / buffers are reusable
        b = serial.run_host(ref, test, ch, n_samples=ns)
    finally:
        piped.close()
        serial.close()
    assert a.tobytes() == b.tobytes() and a2.tobytes() == b.tobytes()
    want = H.oracle_run_pair(ref[2, :lengths[2] * ch], test[2, :lengths[2] * ch], ch)
    check_result(a[2], want, "pipelined pair 2")


# ---------------------------------------------------------------------------
# long items as segments (peaq_segments.cu; BASELINE configs[4])

def _run_one(monkeypatch, advanced, ref, test, ch, ns=None, **env):
    e = _engine_with_env(monkeypatch, advanced=advanced, **env)
    try:
        return e.run_host(ref, test, ch, n_samples=ns)
    finally:
        e.close()


@pytest.mark.parametrize("advanced", [False, True])
def test_segmented_long_items_match_oracle_and_the_whole_run(monkeypatch, advanced):
    """Items longer than 49 s run as ~33 s segments with a 4.1 s warm-up each.  A ragged batch --
    100 s (3 segments), 75 s with a silent stretch across the first segment boundary and a silent
    tail (TENTATIVE accumulators across segments), 20 s (not cut) -- against the oracle, and
    against the same engine with segments switched off: MOVs to 1e-9 (observed <= 1e-13; the
    warm-up leaves the recurrences within an ulp or two of the sequential run)."""
    ch = 2
    lengths = [48000 * 100, 48000 * 75, 48000 * 20]
    stride = max(lengths) * ch
    ref = np.zeros((3, stride), np.float32)
    test = np.zeros_like(ref)
    r, t = synth_pair(900, lengths[0], ch)
    ref[0, :r.size] = r
    test[0, :t.size] = t
    x, y = noise_pair(11, lengths[1], ch, tail=48000 * 9)
    x[ch * 48000 * 30:ch * 48000 * 36] = 0      # silence over the boundary at 32.8 s
    y[ch * 48000 * 30:ch * 48000 * 36] = 0
    ref[1, :x.size] = x
    test[1, :y.size] = y
    r, t = synth_pair(901, lengths[2], ch)
    ref[2, :r.size] = r
    test[2, :t.size] = t
    ns = np.array(lengths, np.uint64)
    seg = _run_one(monkeypatch, advanced, ref, test, ch, ns, PEAQ_B200_SEGMENTS="1")
    whole = _run_one(monkeypatch, advanced, ref, test, ch, ns, PEAQ_B200_SEGMENTS="0")
    assert seg[2].tobytes() == whole[2].tobytes()          # the short item is not touched
    # where an item is cut depends on its length alone: same bits in any batch, at any position
    rev = _run_one(monkeypatch, advanced, ref[::-1].copy(), test[::-1].copy(), ch, ns[::-1].copy(),
                   PEAQ_B200_SEGMENTS="1")
    assert rev[::-1].tobytes() == seg.tobytes()
    alone = _run_one(monkeypatch, advanced, ref[:1, :lengths[0] * ch].copy(), test[:1, :lengths[0] * ch].copy(), ch,
                     None, PEAQ_B200_SEGMENTS="1")
    assert alone[0].tobytes() == seg[0].tobytes()
    n = 5 if advanced else 11
    for p in range(3):
        for k in ("frames_fft", "frames_fb", "loudness_reached_frame", "n_movs"):
            assert seg[k][p] == whole[k][p], (p, k)
        np.testing.assert_allclose(seg["movs"][p][:n], whole["movs"][p][:n], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose([seg["odg"][p], seg["di"][p], seg["totalsnr"][p]],
                                   [whole["odg"][p], whole["di"][p], whole["totalsnr"][p]], rtol=0, atol=1e-9,
                                   equal_nan=True)
        want = H.oracle_run_pair(ref[p, :lengths[p] * ch], test[p, :lengths[p] * ch], ch, advanced=advanced)
        check_result(seg[p], want, "segmented item %d adv %d" % (p, advanced))


@pytest.mark.parametrize("advanced", [False, True])
def test_segments_fall_back_to_the_whole_item_when_their_assumptions_fail(monkeypatch, advanced):
    """An item whose first 40 s are digital silence: no frame above the threshold and no loudness
    latch before the second segment's warm-up, so the segments' start assumptions are wrong; the
    engine notices (redo flag of seg_combine_*) and runs the item as a whole -- same bits as
    with segments switched off -- also as part of a batch, from device-resident PCM too."""
    ch = 2
    n = 48000 * 70
    x, y = noise_pair(12, n, ch, lead=48000 * 40)
    r, t = synth_pair(902, n, ch)
    ref = np.stack([r, x])
    test = np.stack([t, y])
    seg = _run_one(monkeypatch, advanced, ref, test, ch, None, PEAQ_B200_SEGMENTS="1")
    whole = _run_one(monkeypatch, advanced, ref, test, ch, None, PEAQ_B200_SEGMENTS="0")
    assert seg[1].tobytes() == whole[1].tobytes()
    want = H.oracle_run_pair(x, y, ch, advanced=advanced)
    check_result(seg[1], want, "silent start adv %d" % advanced)
    e = _engine_with_env(monkeypatch, advanced=advanced, PEAQ_B200_SEGMENTS="1")
    try:
        dr = G.DeviceBuffer(0, ref.nbytes)
        dt = G.DeviceBuffer(0, test.nbytes)
        L = G.load_library()
        G._check(L.peaq_b200_memcpy_h2d(0, dr.ptr, ref.ctypes.data, ref.nbytes))
        G._check(L.peaq_b200_memcpy_h2d(0, dt.ptr, test.ctypes.data, test.nbytes))
        dev = e.run_device(dr.ptr, dt.ptr, 2, n * ch, ch, n)
    finally:
        e.close()
    assert dev.tobytes() == seg.tobytes()


def test_ten_minute_advanced_item_bounds_the_filter_bank_drift(monkeypatch):
    """The filter bank runs as marginally stable recursions (sliding windowed DFTs, |r| = 1) whose
    rounding error random-walks with the item length.  A 10-minute mono item (150 000 filter-bank
    frames, 900 000 sub-steps) against the oracle's direct FIR filters, once as ONE sequential
    chain (segments off: the recursion never restarts) and once as 18 segments: MOVs to 1e-9
    (observed ~1e-12; the walk grows with the square root of the length, so an hour-long
    sequential session stays below 3e-9), ODG to 1e-9."""
    ch = 1
    n = 48000 * 600
    r, t = G.synth_pairs_host(950, 1, n, ch)
    want = H.oracle_run_pair(r[0], t[0], ch, advanced=True)
    for seg in ("0", "1"):
        out = _run_one(monkeypatch, True, r, t, ch, None, PEAQ_B200_SEGMENTS=seg)
        assert int(out["frames_fb"][0]) == want["frames_fb"] == 150000
        assert int(out["frames_fft"][0]) == want["frames_fft"]
        np.testing.assert_allclose(out["movs"][0][:5], want["movs"], rtol=1e-9, atol=1e-12, err_msg="segments " + seg)
        assert abs(out["odg"][0] - want["odg"]) < 1e-9 and abs(out["di"][0] - want["di"]) < 1e-9
