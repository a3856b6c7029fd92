"""ctypes bindings for the CPU checkers under oracle/ (test infrastructure).

  * ``port()``  -> oracle/liboracle.so: our restatement (speedy_oracle.c).
  * ``ref(kind)`` -> oracle/_ref/libspeedy_ref_{kiss,fftw}.so: the reference's own
    speedy.c + soniclib.c compiled unmodified (see oracle/Makefile), driven
    through oracle/ref_driver.c.

Nothing outside tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke()
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
N_FEATURES = 15

c_float_p = C.POINTER(C.c_float)
c_short_p = C.POINTER(C.c_short)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)


def build():
    """Compile liboracle.so and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


class OracleCfg(C.Structure):
    _fields_ = [("rate", C.c_int), ("channels", C.c_int),
                ("match_matlab", C.c_int), ("fft_double", C.c_int),
                ("speed", C.c_float), ("nonlinear", C.c_float),
                ("feedback", C.c_float)]


class OracleGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("window", "fft", "step", "partial", "future", "past",
                 "min_period", "max_period", "max_required", "skip")]


class OracleTaps(C.Structure):
    _fields_ = [("max_frames", C.c_int), ("n_analysis", C.c_int),
                ("n_tension", C.c_int), ("spectrogram", c_float_p),
                ("energy", c_float_p), ("normalized", c_float_p),
                ("features", c_float_p), ("tension", c_float_p),
                ("speed", c_float_p)]


class RefTaps(C.Structure):
    _fields_ = [("max_frames", C.c_int), ("fft", C.c_int),
                ("n_tension", C.c_int), ("n_speed", C.c_int),
                ("n_features", C.c_int), ("n_spec", C.c_int),
                ("n_norm", C.c_int), ("tension", c_float_p),
                ("speed", c_float_p), ("features", c_float_p),
                ("spectrogram", c_float_p), ("normalized", c_float_p),
                ("tension_time", c_int_p), ("spec_time", c_int_p)]


_port = None
_refs = {}


def port():
    global _port
    if _port is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.oracle_geometry.argtypes = [C.c_int, C.c_int, C.POINTER(OracleGeom)]
        lib.oracle_frames_analyzed.argtypes = [C.POINTER(OracleGeom), C.c_long]
        lib.oracle_tensions_ready.argtypes = [C.POINTER(OracleGeom), C.c_int]
        lib.oracle_hamming.argtypes = [C.c_int, c_float_p]
        lib.oracle_analyze.argtypes = [C.POINTER(OracleCfg), c_short_p, C.c_long,
                                       C.POINTER(OracleTaps)]
        lib.oracle_resynthesize.restype = C.c_long
        lib.oracle_resynthesize.argtypes = [C.POINTER(OracleCfg), c_short_p,
                                            C.c_long, c_float_p, C.c_int, C.c_int,
                                            c_short_p, C.c_long]
        lib.oracle_process.restype = C.c_long
        lib.oracle_process.argtypes = [C.POINTER(OracleCfg), c_short_p, C.c_long,
                                       c_float_p, c_short_p, C.c_long,
                                       C.POINTER(OracleTaps)]
        lib.oracle_process_batch.argtypes = [C.POINTER(OracleCfg), c_short_p,
                                             C.c_long, C.c_int, c_short_p,
                                             C.c_long, c_long_p, C.c_int]
        lib.oracle_synth_fill.argtypes = [c_short_p, C.c_ulonglong, C.c_int,
                                          C.c_int, C.c_int, C.c_long]
        lib.oracle_synth_fill.restype = None
        for name in ("sonicIntCreateStream",):
            getattr(lib, name).restype = C.c_void_p
        _port = lib
    return _port


def ref_available(kind="kiss"):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref",
                                       "libspeedy_ref_%s.so" % kind))


def ref(kind="kiss"):
    """kind: 'kiss' (-DKISS_FFT -DMATCH_MATLAB, how the reference tests build)
    or 'fftw' (the shipped-library configuration)."""
    if kind not in _refs:
        path = os.path.join(ORACLE_DIR, "_ref", "libspeedy_ref_%s.so" % kind)
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.ref_run_stream.restype = C.c_long
        lib.ref_run_stream.argtypes = [c_short_p, C.c_long, C.c_int, C.c_int,
                                       C.c_float, C.c_float, C.c_float, C.c_int,
                                       c_short_p, C.c_long, C.POINTER(RefTaps)]
        lib.ref_run_batch.argtypes = [c_short_p, C.c_long, C.c_int, C.c_int,
                                      C.c_int, C.c_float, C.c_float, C.c_float,
                                      C.c_int, c_short_p, C.c_long, c_long_p,
                                      C.c_int]
        lib.ref_quiet.argtypes = [C.c_int]
        # speedy.h white-box surface (speedy.h:61-133)
        lib.speedyCreateStream.restype = C.c_void_p
        lib.speedyCreateStream.argtypes = [C.c_int]
        lib.speedyDestroyStream.argtypes = [C.c_void_p]
        for n in ("speedyInputFrameSize", "speedyInputFrameStep", "speedyFFTSize"):
            getattr(lib, n).argtypes = [C.c_void_p]
        lib.speedyFreqToBin.argtypes = [C.c_void_p, C.c_float]
        lib.speedyAddData.argtypes = [C.c_void_p, c_float_p, C.c_int64]
        lib.speedyAddDataShort.argtypes = [C.c_void_p, c_short_p, C.c_int64]
        lib.speedyComputeTension.argtypes = [C.c_void_p, C.c_int64, c_float_p]
        lib.speedyComputeSpeedFromTension.restype = C.c_float
        lib.speedyComputeSpeedFromTension.argtypes = [C.c_float, C.c_float,
                                                      C.c_float, C.c_void_p]
        lib.speedyGetCurrentTime.restype = C.c_int64
        lib.speedyGetCurrentTime.argtypes = [C.c_void_p]
        lib.speedySpectrogram.restype = c_float_p
        lib.speedySpectrogram.argtypes = [C.c_void_p, c_float_p]
        for n in ("speedyGetSpectrogram", "speedyGetNormalizedSpectrogram",
                  "speedyGetInternalState", "speedyGetInternalSpectrogram"):
            getattr(lib, n).restype = c_float_p
            getattr(lib, n).argtypes = [C.c_void_p]
        lib.speedyGetSpectrogramAtTime.restype = c_float_p
        lib.speedyGetSpectrogramAtTime.argtypes = [C.c_void_p, C.c_int64]
        lib.speedyEvaluateHysteresis.restype = C.c_float
        lib.speedyEvaluateHysteresis.argtypes = [C.c_void_p, C.c_int64]
        lib.speedyAddToHysteresisBuffer.argtypes = [C.c_void_p, C.c_float, C.c_int64]
        lib.speedyComputeSpectralDifference.argtypes = [C.c_void_p, c_float_p,
                                                        c_float_p, C.c_int64]
        lib.speedyComputeLocalEnergy.argtypes = [C.c_void_p, c_float_p, C.c_int64]
        lib.speedyPreemphasisFilter.argtypes = [C.c_void_p, c_float_p, C.c_int]
        lib.speedyGetEnergyCompressed.restype = C.c_float
        lib.speedyGetEnergyCompressed.argtypes = [C.c_void_p]
        lib.speedyGetSpeechChanges.restype = C.c_float
        lib.speedyGetSpeechChanges.argtypes = [C.c_void_p]
        lib.speedyNormalizeByEnergy.restype = C.c_float
        lib.speedyNormalizeByEnergy.argtypes = [c_float_p, c_float_p, C.c_int]
        lib.CreateFirstOrderFilter.restype = C.c_void_p
        lib.CreateFirstOrderFilter.argtypes = [C.c_float]
        lib.IterateFirstOrderFilter.restype = C.c_float
        lib.IterateFirstOrderFilter.argtypes = [C.c_void_p, C.c_float]
        lib.ResetFirstOrderFilter.argtypes = [C.c_void_p]
        lib.DeleteFirstOrderFilter.argtypes = [C.c_void_p]
        # Sonic API as the reference exports it (sonic2.h:54-125)
        lib.sonicCreateStream.restype = C.c_void_p
        lib.sonicCreateStream.argtypes = [C.c_int, C.c_int]
        lib.sonicIntCreateStream.restype = C.c_void_p
        lib.sonicIntCreateStream.argtypes = [C.c_int, C.c_int]
        for pre in ("sonic", "sonicInt"):
            getattr(lib, pre + "DestroyStream").argtypes = [C.c_void_p]
            getattr(lib, pre + "SetSpeed").argtypes = [C.c_void_p, C.c_float]
            getattr(lib, pre + "WriteShortToStream").argtypes = [C.c_void_p, c_short_p, C.c_int]
            getattr(lib, pre + "ReadShortFromStream").argtypes = [C.c_void_p, c_short_p, C.c_int]
            getattr(lib, pre + "WriteFloatToStream").argtypes = [C.c_void_p, c_float_p, C.c_int]
            getattr(lib, pre + "ReadFloatFromStream").argtypes = [C.c_void_p, c_float_p, C.c_int]
            getattr(lib, pre + "FlushStream").argtypes = [C.c_void_p]
        lib.sonicEnableNonlinearSpeedup.argtypes = [C.c_void_p, C.c_float]
        lib.sonicSetDurationFeedbackStrength.argtypes = [C.c_void_p, C.c_float]
        lib.getSonicBufferSize.argtypes = [C.c_void_p]
        lib.sonicSpectrogramSize.argtypes = [C.c_void_p]
        _refs[kind] = lib
    return _refs[kind]


def fptr(a):
    return a.ctypes.data_as(c_float_p) if a is not None else None


def sptr(a):
    return a.ctypes.data_as(c_short_p) if a is not None else None


def geometry(rate, match_matlab=False):
    g = OracleGeom()
    port().oracle_geometry(rate, int(match_matlab), C.byref(g))
    return g


def cfg(rate, channels=1, speed=2.0, nonlinear=1.0, feedback=0.1,
        match_matlab=False, fft_double=False):
    return OracleCfg(rate, channels, int(match_matlab), int(fft_double),
                     speed, nonlinear, feedback)


def synth(first_id, n_streams, rate, channels, n_frames):
    out = np.empty((n_streams, n_frames, channels), dtype=np.int16)
    port().oracle_synth_fill(sptr(out), first_id, n_streams, rate, channels,
                             n_frames)
    return out


def default_out_cap(n, speed, nonlinear, g):
    """Output frames that always suffice: speeds never drop below 1 when R_g > 1
    (speedy.c:774) and never below kMinimumSpeed = 0.01 otherwise (speedy.c:776)."""
    if speed > 1.0:
        worst = 1.0
    elif nonlinear != 0:
        worst = 0.01
    else:
        worst = speed
    return int(n / worst * 1.02) + 4 * g.max_required


def port_process(c, pcm, speed_override=None, taps=True, out_cap=None):
    """Run one stream through the restatement.  pcm: int16 [frames, channels]
    (or [frames] for mono).  Returns dict(out, spectrogram, energy, normalized,
    features, tension, speed)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, c.channels)
    n = pcm.shape[0]
    g = geometry(c.rate, c.match_matlab)
    nA = port().oracle_frames_analyzed(C.byref(g), n)
    nT = port().oracle_tensions_ready(C.byref(g), nA)
    if out_cap is None:
        out_cap = default_out_cap(n, c.speed, c.nonlinear, g)
    out = np.zeros((out_cap, c.channels), dtype=np.int16)
    res = {}
    t = None
    if taps and c.nonlinear != 0:
        res["spectrogram"] = np.zeros((nA, g.fft), np.float32)
        res["energy"] = np.zeros(nA, np.float32)
        res["normalized"] = np.zeros((nT, g.fft // 2), np.float32)
        res["features"] = np.zeros((nT, N_FEATURES), np.float32)
        res["tension"] = np.zeros(nT, np.float32)
        res["speed"] = np.zeros(nT, np.float32)
        t = OracleTaps(max(nA, 1), 0, 0, fptr(res["spectrogram"]),
                       fptr(res["energy"]), fptr(res["normalized"]),
                       fptr(res["features"]), fptr(res["tension"]),
                       fptr(res["speed"]))
    ov = None
    if speed_override is not None:
        ov = np.ascontiguousarray(speed_override, dtype=np.float32)
        assert ov.shape[0] >= nT
    produced = port().oracle_process(C.byref(c), sptr(pcm), n, fptr(ov), sptr(out),
                                     out_cap, C.byref(t) if t is not None else None)
    assert 0 <= produced <= out_cap, produced
    res["out"] = out[:produced].copy()
    return res


def ref_process(kind, pcm, rate, channels=1, speed=2.0, nonlinear=1.0,
                feedback=0.1, chunk=0, taps=True, out_cap=None):
    """Run one stream through the reference's own API (oracle/_ref)."""
    lib = ref(kind)
    pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, channels)
    n = pcm.shape[0]
    g = geometry(rate, kind == "kiss")
    maxf = n // g.step + 2
    if out_cap is None:
        out_cap = default_out_cap(n, speed, nonlinear, g)
    out = np.zeros((out_cap, channels), dtype=np.int16)
    res = {}
    t = None
    if taps:
        res["spectrogram"] = np.zeros((maxf, g.fft), np.float32)
        res["normalized"] = np.zeros((maxf, g.fft // 2), np.float32)
        res["features"] = np.zeros((maxf, N_FEATURES), np.float32)
        res["tension"] = np.zeros(maxf, np.float32)
        res["speed"] = np.zeros(maxf, np.float32)
        res["tension_time"] = np.zeros(maxf, np.int32)
        res["spec_time"] = np.zeros(maxf, np.int32)
        t = RefTaps(maxf, 0, 0, 0, 0, 0, 0, fptr(res["tension"]),
                    fptr(res["speed"]), fptr(res["features"]),
                    fptr(res["spectrogram"]), fptr(res["normalized"]),
                    res["tension_time"].ctypes.data_as(c_int_p),
                    res["spec_time"].ctypes.data_as(c_int_p))
    lib.ref_quiet(1)
    try:
        produced = lib.ref_run_stream(sptr(pcm), n, rate, channels, speed,
                                      nonlinear, feedback, chunk, sptr(out),
                                      out_cap, C.byref(t) if t is not None else None)
    finally:
        lib.ref_quiet(0)
    assert 0 <= produced <= out_cap, produced
    res["out"] = out[:produced].copy()
    if taps:
        nT, nS = t.n_tension, t.n_spec
        assert nT <= maxf and nS <= maxf
        res["n_norm"] = t.n_norm
        for k in ("tension", "speed", "features", "tension_time"):
            res[k] = res[k][:nT]
        for k in ("spectrogram", "spec_time", "normalized"):
            res[k] = res[k][:nS]
    return res
