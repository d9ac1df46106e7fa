"""gstpeaq_b200 -- host-side mirror of the GStreamer `peaq` element on top of
the B200-native CUDA engine (libpeaq_b200.so, C ABI in include/peaq_b200.h).

The reference's host side is compiled C bound to GStreamer (absent in this
image); this module is the thin binding used by the tests, the bench and the
CLI wrapper.  It mirrors the element's surface (/root/reference/src/gstpeaq.c):

  Peaq              one element instance: properties `advanced`,
                    `playback_level`, `console_output` (gstpeaq.c:273-317),
                    sink pads `ref` / `test` fed through chain_ref/chain_test
                    (pad_chain, :614-661), caps via set_caps (:569-593),
                    PAUSED->READY via stop() (:764-778), result properties
                    `odg`, `di`, `totalsnr` (:484-497).
  Engine            the batch entry the reference lacks: thousands of
                    independent (ref,test) pairs per call.

There is NO CPU fallback: every compute call goes through the CUDA library and
raises PeaqError if the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

__all__ = ["Peaq", "Engine", "MultiEngine", "PeaqError", "Result", "library_path", "load_library",
           "device_count", "synth_pairs_host", "frames_for_samples", "segment_plan", "fb_filter_tables", "ABI_SYMBOLS"]

_HERE = os.path.dirname(os.path.abspath(__file__))

FFT_FRAME = 2048
FFT_STEP = 1024

# every symbol include/peaq_b200.h declares
ABI_SYMBOLS = [
    "peaq_b200_last_error", "peaq_b200_version", "peaq_b200_device_count",
    "peaq_b200_engine_create", "peaq_b200_engine_destroy", "peaq_b200_engine_run_batch",
    "peaq_b200_synth_pairs", "peaq_b200_device_alloc", "peaq_b200_device_free",
    "peaq_b200_memcpy_h2d", "peaq_b200_memcpy_d2h", "peaq_b200_host_alloc_pinned",
    "peaq_b200_host_free_pinned", "peaq_b200_engine_last_ms", "peaq_b200_engine_launch_count",
    "peaq_b200_engine_keep_records", "peaq_b200_engine_record_layout",
    "peaq_b200_engine_copy_records", "peaq_b200_engine_copy_fb_debug", "peaq_b200_table",
    "peaq_b200_session_create", "peaq_b200_session_destroy", "peaq_b200_session_set_advanced",
    "peaq_b200_session_set_playback_level", "peaq_b200_session_get_playback_level",
    "peaq_b200_session_set_channels", "peaq_b200_session_push", "peaq_b200_session_finish",
    "peaq_b200_session_get_result", "peaq_b200_fp64_peak_tflops",
    "peaq_b200_session_snapshot", "peaq_b200_session_restore",
    "peaq_b200_multi_create", "peaq_b200_multi_destroy", "peaq_b200_multi_device_count",
    "peaq_b200_multi_run_batch", "peaq_b200_segment_plan",
]

MOV_NAMES_BASIC = ["BandwidthRefB", "BandwidthTestB", "Total NMRB", "WinModDiff1B", "ADBB", "EHSB",
                   "AvgModDiff1B", "AvgModDiff2B", "RmsNoiseLoudB", "MFPDB", "RelDistFramesB"]
MOV_NAMES_ADVANCED = ["RmsModDiffA", "RmsNoiseLoudAsymA", "SegmentalNMRB", "EHSB", "AvgLinDistA"]


class PeaqError(RuntimeError):
    pass


class Result(C.Structure):
    """peaq_b200_result"""
    _fields_ = [("odg", C.c_double), ("di", C.c_double), ("totalsnr", C.c_double),
                ("movs", C.c_double * 11), ("n_movs", C.c_int32),
                ("frames_fft", C.c_uint32), ("frames_fb", C.c_uint32),
                ("loudness_reached_frame", C.c_uint32)]

    def as_dict(self):
        return {"odg": self.odg, "di": self.di, "totalsnr": self.totalsnr,
                "movs": np.array(self.movs[:self.n_movs]),
                "frames_fft": self.frames_fft, "frames_fb": self.frames_fb,
                "loudness_reached_frame": self.loudness_reached_frame}


RESULT_DTYPE = np.dtype([("odg", "f8"), ("di", "f8"), ("totalsnr", "f8"), ("movs", "f8", (11,)),
                         ("n_movs", "i4"), ("frames_fft", "u4"), ("frames_fb", "u4"),
                         ("loudness_reached_frame", "u4")], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class _Batch(C.Structure):
    """peaq_b200_batch"""
    _fields_ = [("n_pairs", C.c_int32), ("channels", C.c_int32), ("ref", C.c_void_p),
                ("test", C.c_void_p), ("pair_stride", C.c_size_t), ("n_samples", C.c_void_p),
                ("n_samples_all", C.c_uint64), ("on_device", C.c_int32)]


def library_path():
    """libpeaq_b200.so next to this file (PEAQ_B200_LIBRARY overrides it: development aid for
    comparing two builds of the engine)"""
    return os.environ.get("PEAQ_B200_LIBRARY") or os.path.join(_HERE, "libpeaq_b200.so")


_lib = None


def load_library():
    """Loads libpeaq_b200.so; raises PeaqError if it has not been built
    (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise PeaqError("%s is missing: build it with `make -C gstpeaq_b200/csrc` "
                        "(no CPU fallback exists)" % path)
    L = C.CDLL(path)
    L.peaq_b200_last_error.restype = C.c_char_p
    L.peaq_b200_version.restype = C.c_char_p
    L.peaq_b200_engine_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double]
    L.peaq_b200_engine_destroy.argtypes = [C.c_void_p]
    L.peaq_b200_engine_run_batch.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p]
    L.peaq_b200_synth_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32,
                                        C.c_uint64, C.c_uint64, C.c_int32]
    L.peaq_b200_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    L.peaq_b200_device_free.argtypes = [C.c_int, C.c_void_p]
    L.peaq_b200_memcpy_h2d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.peaq_b200_memcpy_d2h.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.peaq_b200_host_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.peaq_b200_host_free_pinned.argtypes = [C.c_void_p]
    L.peaq_b200_engine_last_ms.restype = C.c_double
    L.peaq_b200_engine_last_ms.argtypes = [C.c_void_p, C.c_int]
    L.peaq_b200_engine_launch_count.restype = C.c_uint64
    L.peaq_b200_engine_launch_count.argtypes = [C.c_void_p]
    L.peaq_b200_fp64_peak_tflops.restype = C.c_double
    L.peaq_b200_fp64_peak_tflops.argtypes = [C.c_int]
    L.peaq_b200_engine_keep_records.argtypes = [C.c_void_p, C.c_int]
    L.peaq_b200_engine_record_layout.argtypes = [C.c_void_p, C.c_void_p]
    L.peaq_b200_engine_copy_records.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t,
                                                C.POINTER(C.c_size_t)]
    L.peaq_b200_engine_copy_fb_debug.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t,
                                                 C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
    L.peaq_b200_table.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.peaq_b200_session_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.peaq_b200_session_destroy.argtypes = [C.c_void_p]
    L.peaq_b200_session_set_advanced.argtypes = [C.c_void_p, C.c_int]
    L.peaq_b200_session_set_playback_level.argtypes = [C.c_void_p, C.c_double]
    L.peaq_b200_session_get_playback_level.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.peaq_b200_session_set_channels.argtypes = [C.c_void_p, C.c_int]
    L.peaq_b200_session_push.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.peaq_b200_session_finish.argtypes = [C.c_void_p]
    L.peaq_b200_session_get_result.argtypes = [C.c_void_p, C.POINTER(Result)]
    if hasattr(L, "peaq_b200_session_snapshot"):   # (an older build loaded through PEAQ_B200_LIBRARY lacks these)
        L.peaq_b200_session_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.peaq_b200_session_restore.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.peaq_b200_multi_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.peaq_b200_multi_destroy.argtypes = [C.c_void_p]
        L.peaq_b200_multi_device_count.argtypes = [C.c_void_p]
        L.peaq_b200_multi_run_batch.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise PeaqError("libpeaq_b200: %s (status %d)" %
                        (load_library().peaq_b200_last_error().decode(), rc))


def device_count():
    return int(load_library().peaq_b200_device_count())


def frames_for_samples(n):
    """FFT-clock frames the element processes for n samples per channel
    (do_processing + do_flush, gstpeaq.c:596-611, :716-745)."""
    full = (n - FFT_FRAME) // FFT_STEP + 1 if n >= FFT_FRAME else 0
    left = n - full * FFT_STEP
    return full + (1 if left > 0 else 0)


def segment_plan(n_samples):
    """(segments, segment length, warm-up) in samples for a batch item of n_samples per channel
    (host code; see peaq_b200_segment_plan)."""
    L = load_library()
    L.peaq_b200_segment_plan.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.peaq_b200_segment_plan.restype = C.c_int
    seg, warm = C.c_uint64(), C.c_uint64()
    k = L.peaq_b200_segment_plan(int(n_samples), C.byref(seg), C.byref(warm))
    return int(k), int(seg.value), int(warm.value)


def synth_pairs_host(first_pair, n_pairs, n_samples, channels=2):
    """Synthetic pairs generated on the HOST (integer generator, identical to
    the device one).  Returns (ref, test) float32 arrays [n_pairs, n_samples*channels]."""
    L = load_library()
    ref = np.empty((n_pairs, n_samples * channels), dtype=np.float32)
    test = np.empty_like(ref)
    _check(L.peaq_b200_synth_pairs(-1, ref.ctypes.data, test.ctypes.data, n_samples * channels,
                                   n_pairs, first_pair, n_samples, channels))
    return ref, test


def table(advanced, model, which, playback_level=92.0):
    """Constant table of the engine (host code; see peaq_b200_table)."""
    if model == 2:
        raise PeaqError("model 2 (filter-bank taps) has its own accessor: fb_filter_tables()")
    buf = np.zeros(128, dtype=np.float64)
    n = load_library().peaq_b200_table(int(advanced), float(playback_level), model, which,
                                       buf.ctypes.data)
    return buf[:max(n, 0)].copy()   # n < 0: this model has no such table


def fb_filter_tables(band, playback_level=92.0):
    """Filter-bank tables of one band (host code): dict with N, D, the recursion coefficients
    ph[k, 6] = {P_0, P_+, P_-, Q_0, Q_+, Q_-}[k] and rotations rpow[f, i] = e^{j 32 w_f (i+1)} of
    fb_bank_rec_kernel, and the reference-form taps h[0..N/2] (fbearmodel.c:213-220)."""
    buf = np.zeros(1880, dtype=np.float64)
    n = load_library().peaq_b200_table(1, float(playback_level), 2, int(band), buf.ctypes.data)
    if n < 0:
        _check(n)
    N, D = int(buf[0]), int(buf[1])
    o = 2
    ph = buf[o:o + 384].reshape(32, 6, 2)
    o += 384
    rp = buf[o:o + 36].reshape(3, 6, 2)
    o += 36
    hre = buf[o:o + N // 2 + 1]
    him = buf[o + N // 2 + 1:o + 2 * (N // 2 + 1)]
    return {"N": N, "D": D, "ph": ph[..., 0] + 1j * ph[..., 1], "rpow": rp[..., 0] + 1j * rp[..., 1],
            "h": hre + 1j * him}


class DeviceBuffer:
    """Device memory owned through the C ABI (no torch involved)."""

    def __init__(self, device, nbytes):
        self.device = device
        self.nbytes = nbytes
        p = C.c_void_p()
        _check(load_library().peaq_b200_device_alloc(device, nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            load_library().peaq_b200_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """Batch engine: one per GPU and mode (peaq_b200_engine)."""

    def __init__(self, device=0, advanced=False, playback_level=92.0):
        self.lib = load_library()
        self.device = device
        self.advanced = bool(advanced)
        h = C.c_void_p()
        _check(self.lib.peaq_b200_engine_create(C.byref(h), device, int(advanced),
                                                float(playback_level)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.peaq_b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_host(self, ref, test, channels, n_samples=None):
        """ref/test: float32 arrays [n_pairs, stride] (host), interleaved.
        n_samples: per-pair samples per channel (default: stride // channels)."""
        ref = np.ascontiguousarray(ref, dtype=np.float32)
        test = np.ascontiguousarray(test, dtype=np.float32)
        if ref.ndim == 1:
            ref = ref[None, :]
            test = test[None, :]
        n_pairs, stride = ref.shape
        return self._run(ref.ctypes.data, test.ctypes.data, n_pairs, stride, channels, n_samples,
                         stride // channels, on_device=False, keep=(ref, test))

    def run_device(self, ref_ptr, test_ptr, n_pairs, pair_stride, channels, n_samples):
        """ref_ptr/test_ptr: device pointers (int) on this engine's GPU."""
        return self._run(ref_ptr, test_ptr, n_pairs, pair_stride, channels, None, n_samples,
                         on_device=True)

    def _run(self, ref_ptr, test_ptr, n_pairs, stride, channels, n_samples, n_all, on_device,
             keep=None):
        b = _Batch()
        b.n_pairs = n_pairs
        b.channels = channels
        b.ref = ref_ptr
        b.test = test_ptr
        b.pair_stride = stride
        ns = None
        if n_samples is not None and not np.isscalar(n_samples):
            ns = np.ascontiguousarray(n_samples, dtype=np.uint64)
            assert ns.shape == (n_pairs,)
            b.n_samples = ns.ctypes.data
        else:
            b.n_samples = None
            if n_samples is not None:
                n_all = int(n_samples)
        b.n_samples_all = n_all
        b.on_device = 1 if on_device else 0
        out = np.zeros(n_pairs, dtype=RESULT_DTYPE)
        _check(self.lib.peaq_b200_engine_run_batch(self.h, C.byref(b), out.ctypes.data))
        return out

    def last_ms(self, which=0):
        return float(self.lib.peaq_b200_engine_last_ms(self.h, which))

    def launch_count(self):
        return int(self.lib.peaq_b200_engine_launch_count(self.h))

    def keep_records(self, enable=True):
        _check(self.lib.peaq_b200_engine_keep_records(self.h, int(enable)))

    def fb_debug(self, n_pairs, channels):
        """Advanced mode taps of the last run (needs keep_records(True)):
        (exc [pair, frame, stream, U|E, 40], movs [pair, frame, channel, 8])"""
        cap = 1 << 26
        buf = np.zeros(cap, dtype=np.float64)
        n = C.c_size_t()
        fr = C.c_uint32()
        _check(self.lib.peaq_b200_engine_copy_fb_debug(self.h, buf.ctypes.data, cap, C.byref(n),
                                                       C.byref(fr)))
        F = fr.value
        n_exc = n_pairs * F * 2 * channels * 2 * 40
        exc = buf[:n_exc].reshape(n_pairs, F, 2 * channels, 2, 40)
        movs = buf[n_exc:n_exc + n_pairs * F * channels * 8].reshape(n_pairs, F, channels, 8)
        return exc, movs

    def scan_debug(self, n_pairs, channels, bands=109):
        """Basic mode taps of the scan kernel for the last run (needs keep_records(True)):
        (excitation [pair, frame, ref|test, channel, band], terms [pair, frame, channel, 8])"""
        cap = 1 << 26
        buf = np.zeros(cap, dtype=np.float64)
        n = C.c_size_t()
        fr = C.c_uint32()
        _check(self.lib.peaq_b200_engine_copy_fb_debug(self.h, buf.ctypes.data, cap, C.byref(n),
                                                       C.byref(fr)))
        F = fr.value
        per = 2 * channels * bands + channels * 8
        rows = buf[:n_pairs * F * per].reshape(n_pairs, F, per)
        exc = rows[:, :, :2 * channels * bands].reshape(n_pairs, F, 2, channels, bands)
        terms = rows[:, :, 2 * channels * bands:].reshape(n_pairs, F, channels, 8)
        return exc, terms

    def records(self, n_pairs, n_frames):
        """Per-frame records of the last run (needs keep_records(True))."""
        lay = np.zeros(9, dtype=np.int32)
        _check(self.lib.peaq_b200_engine_record_layout(self.h, lay.ctypes.data))
        Cn, B, off_noise, off_ehs, off_snr, off_ints, stride = [int(x) for x in lay[:7]]
        buf = np.zeros(n_pairs * n_frames * stride, dtype=np.float64)
        n = C.c_size_t()
        _check(self.lib.peaq_b200_engine_copy_records(self.h, buf.ctypes.data, buf.size, C.byref(n)))
        rec = buf[:n.value].reshape(n_pairs, -1, stride)
        ints = np.ascontiguousarray(rec[:, :, off_ints:]).view(np.int32)
        return {
            "unsmeared": rec[:, :, :2 * Cn * B].reshape(n_pairs, -1, 2, Cn, B),
            "noise_in_bands": rec[:, :, off_noise:off_noise + Cn * B].reshape(n_pairs, -1, Cn, B),
            "ehs": rec[:, :, off_ehs:off_ehs + Cn],
            "snr": rec[:, :, off_snr:off_snr + 2],
            "flags": ints[:, :, 0],
            "bw_ref": ints[:, :, 1:1 + 2 * Cn:2],
            "bw_test": ints[:, :, 2:2 + 2 * Cn:2],
        }



class MultiEngine:
    """Batch engine over several GPUs of one process (peaq_b200_multi): contiguous blocks of
    pairs per device, results gathered into one host array."""

    def __init__(self, devices=None, advanced=False, playback_level=92.0):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            _check(self.lib.peaq_b200_multi_create(C.byref(h), None, 0, int(advanced), float(playback_level)))
        else:
            arr = (C.c_int * len(devices))(*devices)
            _check(self.lib.peaq_b200_multi_create(C.byref(h), arr, len(devices), int(advanced),
                                                   float(playback_level)))
        self.h = h

    def device_count(self):
        return int(self.lib.peaq_b200_multi_device_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.peaq_b200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_host(self, ref, test, channels, n_samples=None):
        ref = np.ascontiguousarray(ref, dtype=np.float32)
        test = np.ascontiguousarray(test, dtype=np.float32)
        n_pairs, stride = ref.shape
        b = _Batch()
        b.n_pairs, b.channels, b.ref, b.test, b.pair_stride = n_pairs, channels, ref.ctypes.data, test.ctypes.data, stride
        ns = None
        if n_samples is not None:
            ns = np.ascontiguousarray(n_samples, dtype=np.uint64)
            b.n_samples = ns.ctypes.data
        b.n_samples_all = stride // channels
        b.on_device = 0
        out = np.zeros(n_pairs, dtype=RESULT_DTYPE)
        _check(self.lib.peaq_b200_multi_run_batch(self.h, C.byref(b), out.ctypes.data))
        return out


class Peaq:
    """One `peaq` element instance (struct _GstPeaq, gstpeaq.c:110-139)."""

    def __init__(self, device=0, advanced=False, playback_level=92.0, console_output=True):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.peaq_b200_session_create(C.byref(h), device))
        self.h = h
        self._advanced = False
        self.console_output = console_output
        self._channels = 0
        if playback_level != 92.0:
            self.playback_level = playback_level
        if advanced:
            self.advanced = True

    def close(self):
        if getattr(self, "h", None):
            self.lib.peaq_b200_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- properties (gstpeaq.c:273-317) ------------------------------------
    @property
    def advanced(self):
        return self._advanced

    @advanced.setter
    def advanced(self, value):
        _check(self.lib.peaq_b200_session_set_advanced(self.h, int(bool(value))))
        self._advanced = bool(value)

    @property
    def playback_level(self):
        v = C.c_double()
        _check(self.lib.peaq_b200_session_get_playback_level(self.h, C.byref(v)))
        return v.value

    @playback_level.setter
    def playback_level(self, level):
        _check(self.lib.peaq_b200_session_set_playback_level(self.h, float(level)))

    def _result(self):
        r = Result()
        _check(self.lib.peaq_b200_session_get_result(self.h, C.byref(r)))
        return r

    @property
    def odg(self):
        return self._result().odg

    @property
    def di(self):
        return self._result().di

    @property
    def totalsnr(self):
        return self._result().totalsnr

    def result(self):
        return self._result().as_dict()

    # -- pads ------------------------------------------------------------------
    def set_caps(self, channels):
        """CAPS event on either pad (set_caps, gstpeaq.c:569-593)."""
        _check(self.lib.peaq_b200_session_set_channels(self.h, int(channels)))
        self._channels = int(channels)

    def _chain(self, pad, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float32).reshape(-1)
        if self._channels == 0:
            raise PeaqError("caps not negotiated: call set_caps(channels) first")
        _check(self.lib.peaq_b200_session_push(self.h, pad, buf.ctypes.data,
                                               buf.size // self._channels))

    def chain_ref(self, buf):
        """buffer on the `ref` sink pad (interleaved F32)"""
        self._chain(0, buf)

    def chain_test(self, buf):
        """buffer on the `test` sink pad"""
        self._chain(1, buf)

    def snapshot(self):
        """bytes holding the whole state of the running session (peaq_b200_session_snapshot)"""
        n = C.c_size_t()
        _check(self.lib.peaq_b200_session_snapshot(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(self.lib.peaq_b200_session_snapshot(self.h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def restore(self, blob):
        """continue from a snapshot (possibly taken by another session / process)"""
        _check(self.lib.peaq_b200_session_restore(self.h, blob, len(blob)))
        r = Result()
        self._channels = 0
        # mode and channels come from the snapshot
        import struct
        _, _, adv, ch = struct.unpack_from("<IIii", blob, 0)
        self._advanced = bool(adv)
        self._channels = ch

    def stop(self):
        """PAUSED->READY: flush the last partial frame and evaluate
        (change_state, gstpeaq.c:764-778); prints like calculate_odg when
        console_output is set (:1022-1036, :1050-1061, :1074-1076)."""
        _check(self.lib.peaq_b200_session_finish(self.h))
        r = self._result()
        if self.console_output:
            print(format_console_output(r, self._advanced), end="")
        return r.as_dict()


def format_console_output(r, advanced):
    """The element's console output, byte for byte (gstpeaq.c:1023-1035,
    :1051-1060, :1075)."""
    m = r.movs
    if advanced:
        s = ("RmsModDiffA = %f\nRmsNoiseLoudAsymA = %f\nSegmentalNMRB = %f\nEHSB = %f\n"
             "AvgLinDistA = %f\n" % (m[0], m[1], m[2], m[3], m[4]))
    else:
        s = ("   BandwidthRefB: %f\n  BandwidthTestB: %f\n      Total NMRB: %f\n"
             "    WinModDiff1B: %f\n            ADBB: %f\n            EHSB: %f\n"
             "    AvgModDiff1B: %f\n    AvgModDiff2B: %f\n   RmsNoiseLoudB: %f\n"
             "           MFPDB: %f\n  RelDistFramesB: %f\n" % tuple(m[i] for i in range(11)))
    return s + "Objective Difference Grade: %.3f\n" % r.odg
