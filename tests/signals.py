"""Seeded test signals shared by the golden-fixture generator and the tests
(TEST INFRASTRUCTURE ONLY)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refharness import audiotestsrc, as_interleaved  # noqa: E402


def noise_pair(seed, n, channels, level=0.05, noise=0.005, lead=0, tail=0):
    """band-unlimited gaussian reference + additive noise; optional digital
    silence at both ends (exercises the INIT/TENTATIVE accumulator paths)"""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n * channels) * level).astype(np.float32)
    y = (x + (rng.standard_normal(n * channels) * noise).astype(np.float32)).astype(np.float32)
    if lead:
        x[:lead * channels] = 0
        y[:lead * channels] = 0
    if tail:
        x[-tail * channels:] = 0
        y[-tail * channels:] = 0
    return x, y


def synth_pair(index, n, channels):
    import gstpeaq_b200 as G
    r, t = G.synth_pairs_host(index, 1, n, channels)
    return r[0], t[0]


def golden_cases():
    """name -> (ref, test, channels); all deterministic"""
    n = 128 * 1024
    sine = audiotestsrc("sine", n)
    saw = audiotestsrc("saw", n)
    tri = audiotestsrc("triangle", n)
    cases = {
        "kat_sine_sine_mono": (sine, sine, 1),                      # runtest-1.0.sh:7-18  -> 0.171
        "kat_saw_tri_mono": (saw, tri, 1),                          # runtest-1.0.sh:21-28 -> -2.007
        "kat_saw_tri_stereo": (as_interleaved(saw, 2), as_interleaved(tri, 2), 2),  # :31-48
    }
    for i in (0, 3, 13):
        r, t = synth_pair(i, 60000, 2)
        cases["synth%d_stereo" % i] = (r, t, 2)
    r, t = synth_pair(5, 40000, 1)
    cases["synth5_mono"] = (r, t, 1)
    x, y = noise_pair(1, 48765, 2, lead=20000, tail=16000)
    cases["noise_silence_stereo"] = (x, y, 2)
    x, y = noise_pair(2, 30000, 1, level=0.2, noise=0.05)
    cases["noise_loud_mono"] = (x, y, 1)
    x, y = noise_pair(3, 2048 + 1024 * 5, 2)       # exact number of frames, no padded frame... plus one
    cases["noise_exact_stereo"] = (x, y, 2)
    x, y = noise_pair(4, 1500, 1)                  # shorter than one frame: a single padded frame
    cases["noise_short_mono"] = (x, y, 1)
    return cases
