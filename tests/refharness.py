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


# --------------------------------------------------------------------------
# oracle/libpeaq_oracle.so : our plain-C restatement (oracle/peaq_oracle.c)

MAXB = 109
TRC = 2

FFT_TRACE_DTYPE = np.dtype([
    ("frame", np.int32), ("above_threshold", np.int32),
    ("energy_flag", np.int32, (2, TRC)),
    ("bw_ref", np.int32, (TRC,)), ("bw_test", np.int32, (TRC,)),
    ("ehs_valid", np.int32), ("pad_", np.int32),
    ("unsmeared", np.float64, (2, TRC, MAXB)),
    ("excitation", np.float64, (2, TRC, MAXB)),
    ("noise_in_bands", np.float64, (TRC, MAXB)),
    ("nmr", np.float64, (TRC,)), ("nmr_max", np.float64, (TRC,)),
    ("ehs", np.float64, (TRC,)),
    ("mod_diff1", np.float64, (TRC,)), ("mod_diff2", np.float64, (TRC,)),
    ("temp_wt", np.float64, (TRC,)), ("noise_loud", np.float64, (TRC,)),
    ("adb_steps", np.float64), ("det_prob", np.float64),
    ("signal_energy", np.float64), ("noise_energy", np.float64),
], align=True)

FB_TRACE_DTYPE = np.dtype([
    ("frame", np.int32), ("above_threshold", np.int32),
    ("unsmeared", np.float64, (2, TRC, 40)),
    ("excitation", np.float64, (2, TRC, 40)),
    ("mod_diff", np.float64, (TRC,)), ("temp_wt", np.float64, (TRC,)),
    ("noise_loud", np.float64, (TRC,)), ("missing_comp", np.float64, (TRC,)),
    ("lin_dist", np.float64, (TRC,)),
], align=True)


class OracleResult(C.Structure):
    _fields_ = [("odg", C.c_double), ("di", C.c_double), ("totalsnr", C.c_double),
                ("movs", C.c_double * 11), ("n_movs", C.c_int),
                ("frames_fft", C.c_uint), ("frames_fb", C.c_uint),
                ("loudness_reached_frame", C.c_uint)]

    def as_dict(self):
        return {"odg": self.odg, "di": self.di, "totalsnr": self.totalsnr,
                "movs": np.array(self.movs[:self.n_movs]),
                "frames_fft": self.frames_fft, "frames_fb": self.frames_fb,
                "loudness_reached_frame": self.loudness_reached_frame}


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        L = C.CDLL(ORACLE_SO)
        L.peaq_oracle_new.restype = C.c_void_p
        L.peaq_oracle_new.argtypes = [C.c_int, C.c_double, C.c_int]
        L.peaq_oracle_free.argtypes = [C.c_void_p]
        L.peaq_oracle_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.peaq_oracle_finish.argtypes = [C.c_void_p]
        L.peaq_oracle_result.argtypes = [C.c_void_p, C.POINTER(OracleResult)]
        L.peaq_oracle_set_fft_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.peaq_oracle_set_fb_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.peaq_oracle_run_pair.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_size_t, C.POINTER(OracleResult)]
        L.peaq_oracle_table.restype = C.c_int
        L.peaq_oracle_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.peaq_oracle_stage_new.restype = C.c_void_p
        L.peaq_oracle_stage_new.argtypes = [C.c_int]
        L.peaq_oracle_stage_free.argtypes = [C.c_void_p]
        L.peaq_oracle_stage_fft_ear.argtypes = [C.c_void_p] * 6
        L.peaq_oracle_stage_fb_ear.argtypes = [C.c_void_p] * 4
        L.peaq_oracle_stage_loudness.restype = C.c_double
        L.peaq_oracle_stage_loudness.argtypes = [C.c_void_p, C.c_int]
        L.peaq_oracle_stage_level_adapt.argtypes = [C.c_void_p] * 5
        L.peaq_oracle_stage_modulation.argtypes = [C.c_void_p] * 4
        _oracle_lib = L
    return _oracle_lib


class OraclePeaq:
    def __init__(self, advanced=False, playback_level=92.0, channels=1,
                 fft_trace=0, fb_trace=0):
        self.lib = oracle_lib()
        self.advanced = bool(advanced)
        self.channels = channels
        self.h = self.lib.peaq_oracle_new(int(advanced), float(playback_level), int(channels))
        self.fft_trace = None
        self.fb_trace = None
        if fft_trace:
            self.fft_trace = np.zeros(fft_trace, dtype=FFT_TRACE_DTYPE)
            self.lib.peaq_oracle_set_fft_trace(self.h, self.fft_trace.ctypes.data, fft_trace)
        if fb_trace:
            self.fb_trace = np.zeros(fb_trace, dtype=FB_TRACE_DTYPE)
            self.lib.peaq_oracle_set_fb_trace(self.h, self.fb_trace.ctypes.data, fb_trace)

    def close(self):
        if self.h:
            self.lib.peaq_oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, ref, test):
        ref = np.ascontiguousarray(ref, dtype=np.float32).reshape(-1)
        test = np.ascontiguousarray(test, dtype=np.float32).reshape(-1)
        self.lib.peaq_oracle_push(self.h, ref.ctypes.data, ref.size // self.channels,
                                  test.ctypes.data, test.size // self.channels)

    def finish(self):
        self.lib.peaq_oracle_finish(self.h)

    def result(self):
        r = OracleResult()
        self.lib.peaq_oracle_result(self.h, C.byref(r))
        return r.as_dict()

    def run(self, ref, test):
        self.push(ref, test)
        self.finish()
        return self.result()

    def table(self, model, which):
        buf = np.zeros(128, dtype=np.float64)
        n = self.lib.peaq_oracle_table(self.h, model, which, buf.ctypes.data)
        return buf[:n].copy()


def oracle_run_pair(ref, test, channels, advanced=False, playback_level=92.0):
    ref = np.ascontiguousarray(ref, dtype=np.float32).reshape(-1)
    test = np.ascontiguousarray(test, dtype=np.float32).reshape(-1)
    r = OracleResult()
    oracle_lib().peaq_oracle_run_pair(int(advanced), float(playback_level), int(channels),
                                      ref.ctypes.data, ref.size // channels,
                                      test.ctypes.data, test.size // channels, C.byref(r))
    return r.as_dict()
