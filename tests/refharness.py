"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE ONLY).

  RefPeaq    -> oracle/_ref/libpeaq_ref.so : the reference's own C code
                (compiled from /root/reference/src by oracle/Makefile).
  audiotestsrc -> restatement of gst-plugins-base `audiotestsrc` (the signal
                source of /root/reference/src/runtest-1.0.sh), un-vendored
                third-party code, version unpinned: phase accumulator advanced
                by 2*pi*f/fs BEFORE use, wrapped at 2*pi, volume 0.8, result
                cast to float32.

Nothing in the product package imports this file.
"""
import ctypes as C
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpeaq_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libpeaq_oracle.so")


def have_ref():
    return os.path.exists(REF_SO)


def audiotestsrc(wave, n_samples, freq=440.0, rate=48000, volume=0.8):
    two_pi = 2.0 * math.pi
    step = two_pi * freq / rate
    acc = 0.0
    out = np.empty(n_samples, dtype=np.float32)
    for i in range(n_samples):
        acc += step
        if acc >= two_pi:
            acc -= two_pi
        if wave == "sine":
            v = math.sin(acc) * volume
        elif wave == "saw":
            amp = volume / math.pi
            v = acc * amp if acc < math.pi else (two_pi - acc) * -amp
        elif wave == "triangle":
            amp = volume / (math.pi / 2)
            if acc < math.pi / 2:
                v = acc * amp
            elif acc < math.pi * 1.5:
                v = (acc - math.pi) * -amp
            else:
                v = (two_pi - acc) * -amp
        else:
            raise ValueError(wave)
        out[i] = v
    return out


def as_interleaved(x, channels):
    """mono float32 [n] -> interleaved [n*channels] by channel duplication
    (what audioconvert does when caps force channels=2 on a mono source)."""
    x = np.asarray(x, dtype=np.float32)
    if channels == 1:
        return np.ascontiguousarray(x)
    return np.ascontiguousarray(np.repeat(x[:, None], channels, axis=1).reshape(-1))


class RefPeaq:
    """The reference's per-frame path (gstpeaq.c) behind a tiny C driver."""

    def __init__(self, advanced=False, playback_level=92.0, channels=1):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.peaq_ref_new.restype = C.c_void_p
        L.peaq_ref_new.argtypes = [C.c_int, C.c_double, C.c_int]
        L.peaq_ref_free.argtypes = [C.c_void_p]
        L.peaq_ref_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.peaq_ref_finish.argtypes = [C.c_void_p]
        L.peaq_ref_result.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.peaq_ref_tap.restype = C.c_int
        L.peaq_ref_tap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.peaq_ref_table.restype = C.c_int
        L.peaq_ref_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.advanced = bool(advanced)
        self.channels = channels
        self.h = L.peaq_ref_new(int(advanced), float(playback_level), int(channels))

    def close(self):
        if self.h:
            self.lib.peaq_ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, ref, test):
        ref = np.ascontiguousarray(ref, dtype=np.float32).reshape(-1)
        test = np.ascontiguousarray(test, dtype=np.float32).reshape(-1)
        self.lib.peaq_ref_push(self.h, ref.ctypes.data, ref.size // self.channels,
                               test.ctypes.data, test.size // self.channels)

    def finish(self):
        self.lib.peaq_ref_finish(self.h)

    def result(self):
        odg = C.c_double()
        di = C.c_double()
        snr = C.c_double()
        movs = (C.c_double * 11)()
        f1 = C.c_uint()
        f2 = C.c_uint()
        lr = C.c_uint()
        self.lib.peaq_ref_result(self.h, C.byref(odg), C.byref(di), movs, C.byref(snr),
                                 C.byref(f1), C.byref(f2), C.byref(lr))
        n = 5 if self.advanced else 11
        return {"odg": odg.value, "di": di.value, "totalsnr": snr.value,
                "movs": np.array(movs[:n]), "frames_fft": f1.value,
                "frames_fb": f2.value, "loudness_reached_frame": lr.value}

    def tap(self, which, side_test, channel):
        buf = np.zeros(1025, dtype=np.float64)
        n = self.lib.peaq_ref_tap(self.h, which, int(side_test), channel, buf.ctypes.data)
        return buf[:n].copy()

    def table(self, model, which):
        buf = np.zeros(128, dtype=np.float64)
        n = self.lib.peaq_ref_table(self.h, model, which, buf.ctypes.data)
        return buf[:n].copy()

    def run(self, ref, test):
        self.push(ref, test)
        self.finish()
        return self.result()
