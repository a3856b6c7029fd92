"""speedy_b200 — B200-native nonlinear speech speed-up (google/speedy hot path).

The product is ``libspeedy_b200.so`` (CUDA kernels for sm_100a behind the C ABI of
``include/speedy_b200.h``).  This package is the thin ctypes mirror of that ABI
used by the tests and by bench.py; it adds nothing of its own.  There is no CPU
fallback: loading fails loudly when the library has not been built, and every
compute entry point fails without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPEEDY_B200_LIB") or os.path.join(HERE, "libspeedy_b200.so")

TAP_TENSION, TAP_SPEED, TAP_FEATURES, TAP_SPECTROGRAM, TAP_ENERGY = 1, 2, 4, 8, 16
STATUS_OUTPUT_OVERFLOW, STATUS_FLUSHED, STATUS_INPUT_OVERFLOW, STATUS_READ_TRUNCATED = 1, 2, 4, 8
FEATURE_COUNT = 15


class BatchConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("num_channels", C.c_int32),
                ("num_streams", C.c_int32), ("match_matlab", C.c_int32),
                ("speed", C.c_float), ("nonlinear_factor", C.c_float),
                ("feedback_strength", C.c_float), ("device", C.c_int32),
                ("max_write_frames", C.c_int64), ("out_capacity", C.c_int64),
                ("taps", C.c_int32), ("threads_per_stream", C.c_int32),
                ("analysis_frame_step", C.c_int32)]


tensionFunction = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_float)
speedFunction = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_float)
featuresFunction = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_float))
spectrogramFunction = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_float))

_lib = None


class SessionPoolConfig(C.Structure):  # speedySessionPoolConfig
    _fields_ = [("sample_rate", C.c_int32), ("num_channels", C.c_int32), ("max_sessions", C.c_int32),
                ("device", C.c_int32), ("max_pending_frames", C.c_int32), ("min_speed", C.c_float),
                ("auto_step_sessions", C.c_int32)]


class SessionPoolStats(C.Structure):  # speedySessionPoolStats
    _fields_ = [("steps", C.c_int64), ("session_writes", C.c_int64), ("sessions_served", C.c_int64),
                ("open_sessions", C.c_int32), ("pending_sessions", C.c_int32), ("last_step_ms", C.c_double)]


def lib():
    """The loaded C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "speedy_b200: %s is missing - build it with `python -m speedy_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32p, i16p, fp = C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p
    sig = {
        "speedyBatchDefaultConfig": (None, [C.POINTER(BatchConfig)]),
        "speedyBatchCreate": (vp, [C.POINTER(BatchConfig)]),
        "speedyBatchDestroy": (None, [vp]),
        "speedyBatchLastError": (C.c_char_p, []),
        "speedyBatchReset": (C.c_int, [vp, vp]),
        "speedyBatchSetSpeed": (C.c_int, [vp, fp, C.c_float]),
        "speedyBatchSetNonlinear": (C.c_int, [vp, fp, C.c_float]),
        "speedyBatchSetFeedback": (C.c_int, [vp, fp, C.c_float]),
        "speedyBatchOverrideSpeeds": (C.c_int, [vp, fp, C.c_int64]),
        "speedyBatchWriteDevice": (C.c_int, [vp, i16p, C.c_int64, C.c_int64, vp, vp]),
        "speedyBatchWrite": (C.c_int, [vp, i16p, C.c_int64, C.c_int64, vp]),
        "speedyBatchFlushDevice": (C.c_int, [vp, vp]),
        "speedyBatchFlush": (C.c_int, [vp]),
        "speedyBatchReadDevice": (C.c_int, [vp, i16p, C.c_int64, vp, vp]),
        "speedyBatchRead": (C.c_int, [vp, i16p, C.c_int64, vp]),
        "speedyBatchPeekOutputDevice": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64)]),
        "speedyBatchDiscardOutput": (C.c_int, [vp, vp]),
        "speedyBatchProcess": (C.c_int, [vp, i16p, C.c_int64, i16p, C.c_int64, vp]),
        "speedyBatchSetProfiling": (C.c_int, [vp, C.c_int]),
        "speedyBatchGetKernelTimes": (C.c_int, [vp, fp]),
        "speedyBatchGetTaps": (C.c_int, [vp, C.c_int64, vp, vp, fp, fp, fp, fp, fp]),
        "speedyBatchGetStatus": (C.c_int, [vp, vp]),
        "speedyBatchFrameGeometry": (C.c_int, [C.c_int, i32p, i32p, i32p]),
        "speedyBatchNumStreams": (C.c_int, [vp]),
        "speedyBatchKernelLaunches": (C.c_int64, []),
        "speedyBatchBuildInfo": (C.c_char_p, []),
        "speedyBatchSynthDevice": (C.c_int, [i16p, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp]),
        "speedyBatchHostAlloc": (C.c_void_p, [C.c_size_t, C.c_int]),
        "speedyBatchHostFree": (None, [C.c_void_p]),
        "speedyBatchFlushStreams": (C.c_int, [vp, vp]),
        "speedyBatchResetStreams": (C.c_int, [vp, vp]),
        # session pool (many drop-in handles on one batch)
        "speedySessionPoolDefaultConfig": (None, [C.POINTER(SessionPoolConfig)]),
        "speedySessionPoolCreate": (vp, [C.POINTER(SessionPoolConfig)]),
        "speedySessionPoolDestroy": (None, [vp]),
        "speedySessionPoolOpen": (vp, [vp]),
        "speedySessionPoolStep": (C.c_int, [vp]),
        "speedySessionPoolGetStats": (C.c_int, [vp, C.POINTER(SessionPoolStats)]),
        # Sonic / Speedy drop-in
        "sonicCreateStream": (vp, [C.c_int, C.c_int]),
        "sonicDestroyStream": (None, [vp]),
        "sonicWriteShortToStream": (C.c_int, [vp, i16p, C.c_int]),
        "sonicReadShortFromStream": (C.c_int, [vp, i16p, C.c_int]),
        "sonicWriteFloatToStream": (C.c_int, [vp, fp, C.c_int]),
        "sonicReadFloatFromStream": (C.c_int, [vp, fp, C.c_int]),
        "sonicSetRate": (None, [vp, C.c_float]),
        "sonicSetSpeed": (None, [vp, C.c_float]),
        "sonicFlushStream": (C.c_int, [vp]),
        "sonicEnableNonlinearSpeedup": (None, [vp, C.c_float]),
        "sonicSetDurationFeedbackStrength": (None, [vp, C.c_float]),
        "getSonicBufferSize": (C.c_int, [vp]),
        "sonicSpectrogramSize": (C.c_int, [vp]),
        "sonicTensionCallback": (None, [vp, tensionFunction]),
        "getSonicTensionCallback": (tensionFunction, [vp]),
        "sonicSpeedCallback": (None, [vp, speedFunction]),
        "getSonicSpeedCallback": (speedFunction, [vp]),
        "sonicFeaturesCallback": (None, [vp, featuresFunction]),
        "getSonicFeaturesCallback": (featuresFunction, [vp]),
        "sonicSpectrogramCallback": (None, [vp, spectrogramFunction]),
        "getSonicSpectrogramCallback": (spectrogramFunction, [vp]),
        "sonicNormalizedSpectrogramCallback": (None, [vp, spectrogramFunction]),
        "getSonicNormalizedSpectrogramCallback": (spectrogramFunction, [vp]),
        "sonicIntGetNumChannels": (C.c_int, [vp]),
        "sonicIntGetSampleRate": (C.c_int, [vp]),
        "sonicIntGetSpeed": (C.c_float, [vp]),
        "sonicIntSamplesAvailable": (C.c_int, [vp]),
        "sonicIntSetSpeed": (None, [vp, C.c_float]),
        "sonicIntWriteShortToStream": (C.c_int, [vp, i16p, C.c_int]),
        "sonicIntReadShortFromStream": (C.c_int, [vp, i16p, C.c_int]),
        "sonicIntFlushStream": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L


def last_error():
    return lib().speedyBatchLastError().decode()


def frame_geometry(rate):
    w, n, s = C.c_int32(), C.c_int32(), C.c_int32()
    if not lib().speedyBatchFrameGeometry(rate, C.byref(w), C.byref(n), C.byref(s)):
        raise ValueError("unsupported sample rate %d" % rate)
    return w.value, n.value, s.value


def _ptr(x):
    """Raw address of a torch tensor / numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError(type(x))


class Batch:
    """speedyBatch* handle (include/speedy_b200.h section 2)."""

    def __init__(self, num_streams, sample_rate=16000, num_channels=1, speed=1.0,
                 nonlinear=0.0, feedback=0.1, match_matlab=False, device=0,
                 max_write_frames=16000, out_capacity=0, taps=0, threads_per_stream=0,
                 analysis_frame_step=0):
        L = lib()
        cfg = BatchConfig()
        L.speedyBatchDefaultConfig(C.byref(cfg))
        cfg.sample_rate, cfg.num_channels, cfg.num_streams = sample_rate, num_channels, num_streams
        cfg.match_matlab = int(match_matlab)
        cfg.speed, cfg.nonlinear_factor, cfg.feedback_strength = speed, nonlinear, feedback
        cfg.device, cfg.max_write_frames, cfg.out_capacity = device, max_write_frames, out_capacity
        cfg.taps, cfg.threads_per_stream = taps, threads_per_stream
        self.cfg = cfg
        self.n = num_streams
        self.channels = num_channels
        self.window, self.fft, self.step = frame_geometry(sample_rate)
        if analysis_frame_step > 0:  # white-box hook (speedy_b200.h)
            cfg.analysis_frame_step, self.step = analysis_frame_step, analysis_frame_step
        self.max_rows = max_write_frames // self.step + 2
        self.h = L.speedyBatchCreate(C.byref(cfg))
        if not self.h:
            raise RuntimeError("speedyBatchCreate failed: " + last_error())

    def _ok(self, rc, what):
        if not rc:
            raise RuntimeError("%s failed: %s" % (what, last_error()))

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.speedyBatchDestroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, stream=None):
        self._ok(lib().speedyBatchReset(self.h, stream), "speedyBatchReset")

    def set_speed(self, value):
        self._set(lib().speedyBatchSetSpeed, value)

    def set_nonlinear(self, value):
        self._set(lib().speedyBatchSetNonlinear, value)

    def set_feedback(self, value):
        self._set(lib().speedyBatchSetFeedback, value)

    def _set(self, fn, value):
        if np.isscalar(value):
            self._ok(fn(self.h, None, float(value)), fn.__name__)
        else:
            v = np.ascontiguousarray(value, np.float32)
            assert v.shape == (self.n,)
            self._ok(fn(self.h, v.ctypes.data, 0.0), fn.__name__)

    def override_speeds(self, speeds):
        if speeds is None:
            self._ok(lib().speedyBatchOverrideSpeeds(self.h, None, 0), "override")
            return
        v = np.ascontiguousarray(speeds, np.float32)
        assert v.ndim == 2 and v.shape[0] == self.n
        self._ok(lib().speedyBatchOverrideSpeeds(self.h, v.ctypes.data, v.shape[1]), "override")

    # -- host buffers (numpy int16 [n, frames, channels]) -------------------
    def write(self, pcm, counts=None):
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(self.n, -1, self.channels)
        frames = pcm.shape[1]
        c = None if counts is None else np.ascontiguousarray(counts, np.int32)
        self._ok(lib().speedyBatchWrite(self.h, pcm.ctypes.data, frames, frames, _ptr(c)), "speedyBatchWrite")

    def flush(self):
        self._ok(lib().speedyBatchFlush(self.h), "speedyBatchFlush")

    def read(self, max_frames):
        out = np.zeros((self.n, max_frames, self.channels), np.int16)
        counts = np.zeros(self.n, np.int32)
        self._ok(lib().speedyBatchRead(self.h, out.ctypes.data, max_frames, counts.ctypes.data), "speedyBatchRead")
        return out, counts

    def process(self, pcm, out_frames):
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(self.n, -1, self.channels)
        out = np.zeros((self.n, out_frames, self.channels), np.int16)
        counts = np.zeros(self.n, np.int32)
        self._ok(lib().speedyBatchProcess(self.h, pcm.ctypes.data, pcm.shape[1], out.ctypes.data, out_frames,
                                          counts.ctypes.data), "speedyBatchProcess")
        return out, counts

    # -- device buffers (torch tensors or raw addresses) --------------------
    def write_device(self, d_in, stride_frames, frames, d_counts=None, stream=None):
        self._ok(lib().speedyBatchWriteDevice(self.h, _ptr(d_in), stride_frames, frames, _ptr(d_counts), stream),
                 "speedyBatchWriteDevice")

    def flush_device(self, stream=None):
        self._ok(lib().speedyBatchFlushDevice(self.h, stream), "speedyBatchFlushDevice")

    def read_device(self, d_out, stride_frames, d_counts, stream=None):
        self._ok(lib().speedyBatchReadDevice(self.h, _ptr(d_out), stride_frames, _ptr(d_counts), stream),
                 "speedyBatchReadDevice")

    def discard_output(self, stream=None):
        self._ok(lib().speedyBatchDiscardOutput(self.h, stream), "speedyBatchDiscardOutput")

    def set_profiling(self, on):
        self._ok(lib().speedyBatchSetProfiling(self.h, int(on)), "speedyBatchSetProfiling")

    def kernel_times(self):
        ms = np.zeros(5, np.float32)
        self._ok(lib().speedyBatchGetKernelTimes(self.h, ms.ctypes.data), "speedyBatchGetKernelTimes")
        return dict(zip(("spectral", "tension", "sonic", "tail", "flush_sonic"), ms.tolist()))

    def write_ptr(self, h_in, stride_frames, frames, offset_frames=0):
        """speedyBatchWrite on a raw host address (e.g. a pinned torch tensor): stream s starts
        at h_in + (s * stride_frames + offset_frames) frames."""
        addr = _ptr(h_in) + offset_frames * self.channels * 2
        self._ok(lib().speedyBatchWrite(self.h, addr, stride_frames, frames, None), "speedyBatchWrite")

    def read_ptr(self, h_out, stride_frames, h_counts):
        """speedyBatchRead into raw host addresses (int16 [n, stride_frames, channels], int32 [n])."""
        self._ok(lib().speedyBatchRead(self.h, _ptr(h_out), stride_frames, _ptr(h_counts)), "speedyBatchRead")

    def process_ptr(self, h_in, frames, h_out, out_stride, h_counts):
        """speedyBatchProcess on raw host addresses (e.g. pinned torch tensors)."""
        self._ok(lib().speedyBatchProcess(self.h, _ptr(h_in), frames, _ptr(h_out), out_stride, _ptr(h_counts)),
                 "speedyBatchProcess")

    def status(self):
        st = np.zeros(self.n, np.int32)
        self._ok(lib().speedyBatchGetStatus(self.h, st.ctypes.data), "speedyBatchGetStatus")
        return st

    def taps(self):
        """Taps of the last write: dict of per-stream lists trimmed to the frames
        that write produced."""
        rows = self.max_rows
        t = self.cfg.taps
        na = np.zeros(self.n, np.int32)
        nt = np.zeros(self.n, np.int32)
        spec = np.zeros((self.n, rows, self.fft), np.float32) if t & TAP_SPECTROGRAM else None
        energy = np.zeros((self.n, rows), np.float32) if t & TAP_ENERGY else None
        feat = np.zeros((self.n, rows, FEATURE_COUNT), np.float32) if t & TAP_FEATURES else None
        tens = np.zeros((self.n, rows), np.float32) if t & TAP_TENSION else None
        spd = np.zeros((self.n, rows), np.float32) if t & TAP_SPEED else None
        self._ok(lib().speedyBatchGetTaps(self.h, rows, na.ctypes.data, nt.ctypes.data, _ptr(spec), _ptr(energy),
                                          _ptr(feat), _ptr(tens), _ptr(spd)), "speedyBatchGetTaps")
        res = {"n_analysis": na, "n_tension": nt}
        if spec is not None:
            res["spectrogram"] = [spec[s, :na[s]] for s in range(self.n)]
        if energy is not None:
            res["energy"] = [energy[s, :na[s]] for s in range(self.n)]
        if feat is not None:
            res["features"] = [feat[s, :nt[s]] for s in range(self.n)]
        if tens is not None:
            res["tension"] = [tens[s, :nt[s]] for s in range(self.n)]
        if spd is not None:
            res["speed"] = [spd[s, :nt[s]] for s in range(self.n)]
        return res


def synth_device(d_out, first_id, num_streams, rate, channels, frames, stream=None):
    if not lib().speedyBatchSynthDevice(_ptr(d_out), first_id, num_streams, rate, channels, frames, stream):
        raise RuntimeError("speedyBatchSynthDevice failed: " + last_error())


def kernel_launches():
    return lib().speedyBatchKernelLaunches()


def host_alloc(shape, dtype=np.int16, write_combined=False):
    """Page-locked numpy array from speedyBatchHostAlloc (free with host_free(arr))."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().speedyBatchHostAlloc(nbytes, int(write_combined))
    if not p:
        raise RuntimeError("speedyBatchHostAlloc failed: " + last_error())
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


def host_free(arr):
    lib().speedyBatchHostFree(arr.ctypes.data)



class SessionPool:
    """Mirror of the session pool (speedy_b200.h section 1b): handles from open() are plain
    sonicStream handles for the drop-in calls (lib().sonicWriteShortToStream ...)."""

    def __init__(self, rate, channels=1, max_sessions=1024, device=0, max_pending_frames=0, min_speed=0.25,
                 auto_step_sessions=0):
        cfg = SessionPoolConfig()
        lib().speedySessionPoolDefaultConfig(C.byref(cfg))
        cfg.sample_rate, cfg.num_channels, cfg.max_sessions, cfg.device = rate, channels, max_sessions, device
        cfg.max_pending_frames, cfg.min_speed, cfg.auto_step_sessions = max_pending_frames, min_speed, auto_step_sessions
        self.h = lib().speedySessionPoolCreate(C.byref(cfg))
        if not self.h:
            raise RuntimeError("speedySessionPoolCreate failed: " + last_error())

    def open(self):
        h = lib().speedySessionPoolOpen(self.h)
        if not h:
            raise RuntimeError("speedySessionPoolOpen: pool is full")
        return h

    def step(self):
        n = lib().speedySessionPoolStep(self.h)
        if n < 0:
            raise RuntimeError("speedySessionPoolStep failed: " + last_error())
        return n

    def stats(self):
        st = SessionPoolStats()
        lib().speedySessionPoolGetStats(self.h, C.byref(st))
        return {k: getattr(st, k) for k, _ in SessionPoolStats._fields_}

    def close(self):
        if self.h:
            lib().speedySessionPoolDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
