"""CPU: host-side logic of the product -- the C-ABI library loads and exports
every symbol of include/peaq_b200.h, its constant tables equal the oracle's,
the integer generator is deterministic, and compute calls fail loudly without
a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import gstpeaq_b200 as G
import refharness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "peaq_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(peaq_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(G.ABI_SYMBOLS)
    L = G.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert L.peaq_b200_version().startswith(b"peaq-b200")


def test_result_struct_layout():
    assert G.C.sizeof(G.Result) == 128 == G.RESULT_DTYPE.itemsize


def test_frames_for_samples_matches_reference_framing():
    # gstpeaq.c:596-611 + :716-745 (10 s -> 468, 1 h -> 168750)
    assert G.frames_for_samples(480000) == 468
    assert G.frames_for_samples(172800000) == 168750
    assert G.frames_for_samples(131072) == 128
    assert G.frames_for_samples(0) == 0
    assert G.frames_for_samples(1) == 1
    assert G.frames_for_samples(2048) == 2
    assert G.frames_for_samples(2047) == 1
    for n in (1500, 2049, 3072, 3073, 48765):
        o = H.oracle_run_pair(np.zeros(n, np.float32), np.zeros(n, np.float32), 1)
        assert o["frames_fft"] == G.frames_for_samples(n)


@pytest.mark.skipif(G.device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(G.PeaqError, match="no CPU fallback"):
        G.Engine(0)
    with pytest.raises(G.PeaqError, match="no CPU fallback"):
        G.Peaq(0)


def test_synth_generator_is_deterministic_and_random_access():
    a_r, a_t = G.synth_pairs_host(3, 2, 5000, 2)
    b_r, b_t = G.synth_pairs_host(4, 1, 5000, 2)
    np.testing.assert_array_equal(a_r[1], b_r[0])
    np.testing.assert_array_equal(a_t[1], b_t[0])
    # 16-bit grid, exact /32768
    assert np.all(a_r * 32768 == np.round(a_r * 32768))
    assert np.abs(a_r).max() < 1.0 and np.abs(a_r).max() > 0.05
    assert not np.array_equal(a_r[0], a_t[0])
    m_r, _ = G.synth_pairs_host(3, 1, 5000, 1)
    np.testing.assert_array_equal(m_r[0], a_r[0][0::2])   # channel 0 is the mono signal


def test_synth_pairs_have_finite_spread_odg():
    """SURVEY 8d acceptance: every synthetic pair gives a finite ODG"""
    odgs = []
    for p in range(14):
        r, t = G.synth_pairs_host(p, 1, 48000, 2)
        o = H.oracle_run_pair(r[0], t[0], 2)
        assert np.isfinite(o["odg"]), p
        odgs.append(o["odg"])
    assert min(odgs) < -1.5 and max(odgs) > -0.6


def test_engine_tables_match_oracle_tables():
    """every table the kernels read equals the oracle's (built by the same
    formulas as the reference's constructors)"""
    for advanced in (0, 1):
        o = H.OraclePeaq(bool(advanced), 92.0, 1)
        for model in (0, 1):
            for which in range(15):
                want = o.table(model, which)
                got = G.table(advanced, model, which)
                assert got.size == want.size, (advanced, model, which)
                if want.size:
                    np.testing.assert_allclose(got, want, rtol=2e-15, atol=0,
                                               err_msg=str((advanced, model, which)))


def test_filter_bank_recursion_tables_reproduce_the_taps():
    """host-only check of the sliding-DFT form of the filter bank (fb_bank_rec_kernel): for every
    band the three coefficient sets sum to the reference's taps, the removal coefficients are the
    entry coefficients times -e^{jwN}, and the rotations are unit-modulus powers"""
    import gstpeaq_b200 as G
    for band in range(40):
        t = G.fb_filter_tables(band)
        N, ph, rp, h = t["N"], t["ph"], t["rpow"], t["h"]
        scale = np.abs(ph).max()   # the three terms cancel (2 - 1 - 1) near the window's edge
        k = np.arange(32)
        # full-length taps: even / odd symmetry about N/2 (fbearmodel.c:418-425)
        h = np.concatenate([h, np.conj(h[-2::-1])])
        # sum_f P_f[k] = (Wt/N)(2 - e^{jdk} - e^{-jdk}) e^{jw(k - N/2)} = h[k]   (N >= 52 > 2 * 31)
        # (the reference's taps carry the rounding of their cosine arguments, ~1700 rad: 2e-13 relative)
        np.testing.assert_allclose(ph[:, :3].sum(axis=1), h[k], rtol=1e-12, atol=2e-15 * scale)
        # Q_f[k] = -c P_f[k] with one unit-modulus c per band
        c = -ph[:, 3:] / ph[:, :3]
        np.testing.assert_allclose(np.abs(c), 1.0, rtol=0, atol=1e-13)
        np.testing.assert_allclose(c, c[1, 0], rtol=0, atol=1e-12)
        # rotations: |r| = 1, r_f^(i+1); and P_f[k+1] / P_f[k] = e^{j w_f} = r_f^(1/32)
        np.testing.assert_allclose(np.abs(rp), 1.0, rtol=0, atol=1e-15)
        for f in range(3):
            np.testing.assert_allclose(rp[f], rp[f, 0] ** np.arange(1, 7), rtol=0, atol=1e-14)
            step = ph[1:, f] / ph[:-1, f]
            np.testing.assert_allclose(step ** 32, rp[f, 0], rtol=0, atol=1e-12)
        assert t["D"] == 1 + (1456 - N) // 2


def test_filter_bank_recursion_equals_direct_fir_in_numpy():
    """the algorithm of fb_bank_rec_kernel replayed in numpy from the library's own tables: the
    three one-step recursions reproduce the direct FIR (taps of fbearmodel.c:213-220 at delays
    D + n, fbearmodel.c:408) on a random signal"""
    import gstpeaq_b200 as G
    rng = np.random.default_rng(5)
    n_sub = 120
    x = rng.standard_normal(32 * n_sub + 1) * 1e3

    def at(i):   # x[i], zero before the start of the signal
        i = np.asarray(i)
        return np.where(i >= 0, x[np.clip(i, 0, None)], 0.0)

    for band in (1, 7, 19, 30, 39):   # (band 0 additionally carries the reference's ring-buffer alias)
        t = G.fb_filter_tables(band)
        N, D, ph, rp = t["N"], t["D"], t["ph"], t["rpow"]
        h = np.concatenate([t["h"], np.conj(t["h"][-2::-1])])   # taps 0..N, h[0] = h[N] = 0
        n = np.arange(1, N)
        k = np.arange(32)
        S = np.zeros(3, complex)
        worst, scale = 0.0, 0.0
        for s in range(n_sub):
            direct = np.sum(h[n] * at(32 * s - D - n))
            S = rp[:, 0] * S + ph[:, :3].T @ at(32 * s - D - k) + ph[:, 3:].T @ at(32 * s - D - k - N)
            worst = max(worst, abs(S.sum() - direct))
            scale = max(scale, abs(direct))
        assert worst < 1e-12 * scale, (band, worst, scale)


def test_time_parallel_dc_reject_scan_algorithm_in_numpy():
    """the algorithm of the opt-in fb_hp_par_* kernels (block scan + refinement) replayed in
    plain Python on the reference's recurrence (fbearmodel.c:292-303): without the refinement
    pass the start states are ~1e-10 off (the homogeneous transition cancels ~30:1), with it the
    output agrees with the sequential run to a few 1e-12 of the signal level"""
    L, N = 512, 512 * 24
    rng = np.random.default_rng(3)
    t = np.arange(N)
    x = (0.3 * np.sin(2 * np.pi * 40 / 48000 * t) + 0.2 * np.sin(2 * np.pi * 1000 / 48000 * t)
         + 0.01 * rng.standard_normal(N) + 0.05).astype(np.float32).astype(np.float64) * 39810.7

    def run(seg, s, x1, x2):
        y1a, y2a, y1b, y2b = s
        out = np.empty(len(seg))
        for i, xi in enumerate(seg):
            h1 = xi - 2. * x1 + x2 + 1.99517 * y1a - 0.995174 * y2a
            h2 = h1 - 2. * y1a + y2a + 1.99799 * y1b - 0.997998 * y2b
            x2, x1, y2a, y1a, y2b, y1b = x1, xi, y1a, h1, y1b, h2
            out[i] = h2
        return out, np.array([y1a, y2a, y1b, y2b])

    ref, _ = run(x, (0., 0., 0., 0.), 0., 0.)
    M = np.stack([run(np.zeros(L), tuple(np.eye(4)[k]), 0., 0.)[1] for k in range(4)], axis=1)
    nb = N // L
    hist = lambda j: (x[j * L - 1], x[j * L - 2]) if j else (0., 0.)
    blocks = [x[j * L:(j + 1) * L] for j in range(nb)]
    e0 = [run(blocks[j], (0., 0., 0., 0.), *hist(j))[1] for j in range(nb)]
    s0 = [np.zeros(4)]
    for j in range(nb - 1):
        s0.append(e0[j] + M @ s0[j])
    e1 = [run(blocks[j], tuple(s0[j]), *hist(j))[1] for j in range(nb)]
    d = np.zeros(4)
    s = [s0[0]]
    for j in range(nb - 1):
        d = (e1[j] - s0[j + 1]) + M @ d
        s.append(s0[j + 1] + d)
    rms = np.sqrt(np.mean(ref ** 2))
    coarse = np.concatenate([run(blocks[j], tuple(s0[j]), *hist(j))[0] for j in range(nb)])
    fine = np.concatenate([run(blocks[j], tuple(s[j]), *hist(j))[0] for j in range(nb)])
    assert np.abs(fine - ref).max() < 5e-11 * rms
    assert np.abs(fine - ref).max() < 0.2 * np.abs(coarse - ref).max()   # the refinement is what gets it there
    assert np.abs(M).max() > 50      # the ill-conditioning that makes it necessary


def test_segment_plan_of_long_items():
    """peaq_segments.cu: items are cut by their length alone, into equally long segments whose
    boundaries lie on whole frames of both clocks (1024 / 192 samples) and on the DC-reject scan's
    512-sample blocks; the warm-up is 4.096 s; at most 512 segments per item."""
    for n in (0, 1, 480000, 48000 * 49, 48000 * 50, 48000 * 100, 48000 * 600, 48000 * 3600, 48000 * 3600 * 8 + 12345):
        k, seg, warm = G.segment_plan(n)
        assert warm == 196608 and warm % 1024 == 0 and warm % 192 == 0 and warm % 512 == 0
        assert 1 <= k <= 512
        if n <= 48000 * 49:
            assert k == 1                      # BASELINE configs[1..3] items are never cut
            continue
        assert seg % 3072 == 0 and seg >= 4 * warm
        assert (k - 1) * seg < n <= k * seg    # equal lengths, a non-empty last segment
        if k < 512:
            assert abs(seg / 48000. - 32.768) < 11.    # about 33 s each
    assert G.segment_plan(48000 * 600) == (18, 1600512, 196608)
    assert G.segment_plan(48000 * 3600)[0] == 110


def test_segment_combination_rule_replayed_in_python():
    """The rule seg_combine_* applies (peaq_segments.cu), replayed with integers: an accumulator
    in the reference commits everything from the first frame above the threshold up to the LAST
    one (movaccum.c:317-354: frames after it stay tentative).  Segments whose sums restart at
    their first frame report `cur` (all their frames), `saved` (up to their own last frame above
    the threshold, when they end tentative) and whether they own such a frame; with T = sum of
    `cur` of the segments before s, the item's value is T + M(s) at the last owning segment."""
    rng = np.random.default_rng(3)
    for trial in range(200):
        n, seg = int(rng.integers(20, 200)), int(rng.integers(5, 40))
        above = rng.random(n) < rng.choice([0.02, 0.3, 0.9])
        above[0] = True                       # assumption A1: the item has started in segment 0
        val = rng.integers(1, 1000, n)
        # the sequential state machine
        num = saved = 0
        status = 0                            # 0 INIT, 1 NORMAL, 2 TENTATIVE
        for f in range(n):
            if not above[f]:
                if status == 1:
                    saved, status = num, 2
            else:
                status = 1
            if status != 0:
                num += int(val[f])
        want = saved if status == 2 else num
        # segments: status NORMAL at their start (A1), sums restart at their first frame
        T = committed = 0
        for s0 in range(0, n, seg):
            cur = sv = 0
            st, owned = 1, False
            for f in range(s0, min(s0 + seg, n)):
                if not above[f]:
                    if st == 1:
                        sv, st = cur, 2
                else:
                    st, owned = 1, True
                cur += int(val[f])
            m = sv if st == 2 else cur
            if owned:
                committed = T + m
            T += cur
        assert committed == want, (trial, n, seg)


def test_frame_kernel_fft_two_trip_form_replayed_on_the_host(tmp_path):
    """The 1024-point transform of the frame kernel (peaq_fft.cuh: two radix-4 levels per trip through
    shared memory, mirror exchange, swizzled twiddle table) and the register-resident first levels of
    the EHS transforms, run thread by thread on the host: bit-equal to the single-level passes, equal
    to a long-double DFT, every 128-bit shared-memory access at its minimum wavefront count."""
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft_check")
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-fmad=false", "-w", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "fft_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


def test_sass_of_the_built_kernels_has_what_design_md_claims():
    """static SASS facts of the in-tree build (scripts/sass_summary.py over the kernel objects): the
    frame kernel stages PCM with TMA bulk copies behind an mbarrier, the filter bank reads half of
    its coefficients from constant memory on the uniform datapath (LDCU from bank 3, DFMAs with a
    uniform-register operand), its shared-memory cross-check form does not, and nothing on the path
    uses tensor cores"""
    import glob
    import shutil
    import subprocess
    import sys
    if not shutil.which("cuobjdump") or not glob.glob(os.path.join(ROOT, "gstpeaq_b200", "csrc", "build", "*.o")):
        pytest.skip("no cuobjdump or no kernel objects (library not built in this tree)")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sass_summary
    facts = sass_summary.collect()
    k1 = facts["fft_frames_kernel"]
    assert k1.get("UBLKCP (TMA bulk)", 0) >= 2 and k1.get("SYNCS (mbarrier)", 0) >= 1, k1
    assert k1.get("DFMA", 0) > 300 and k1.get("LDS", 0) > 100
    fb_const, fb_smem = facts["fb_bank_rec_kernel<1>"], facts["fb_bank_rec_kernel<0>"]
    assert fb_const.get("LDCU c[0x3] (uniform datapath)", 0) >= 48 and fb_const.get("DFMA with UR operand", 0) >= 400, fb_const
    assert fb_smem.get("LDCU c[0x3] (uniform datapath)", 0) == 0
    assert fb_const.get("LDGSTS (cp.async)", 0) >= 1
    for name, f in facts.items():
        assert not any(k.startswith("tensor") and v for k, v in f.items()), (name, f)
