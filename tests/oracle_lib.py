"""ctypes bindings to the CPU oracle (oracle/liboracle.so) and to the compiled-reference harness
(oracle/_ref/tube_ref).  Test infrastructure only -- never imported by the product package."""
import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "tube_ref")

WAVETABLE_ANALYTIC = 1
SRC_STATELESS = 2


class OracleInputParameters(C.Structure):
    _fields_ = [
        ("outputFileFormat", C.c_int32), ("outputRate", C.c_float), ("controlRate", C.c_float),
        ("volume", C.c_double), ("channels", C.c_int32), ("balance", C.c_double), ("waveform", C.c_int32),
        ("tp", C.c_double), ("tnMin", C.c_double), ("tnMax", C.c_double), ("breathiness", C.c_double),
        ("length", C.c_double), ("temperature", C.c_double), ("lossFactor", C.c_double), ("apScale", C.c_double),
        ("mouthCoef", C.c_double), ("noseCoef", C.c_double), ("noseRadius", C.c_double * 6),
        ("throatCutoff", C.c_double), ("throatVol", C.c_double), ("usesModulation", C.c_int32),
        ("mixOffset", C.c_double),
    ]


class OracleInfo(C.Structure):
    _fields_ = [("controlPeriod", C.c_int32), ("sampleRate", C.c_int32), ("actualTubeLength", C.c_double),
                ("numberTaps", C.c_int32), ("padSize", C.c_int32), ("timeRegisterIncrement", C.c_uint32),
                ("numberSamples", C.c_int32), ("maximumSampleValue", C.c_double), ("finalNoiseSeed", C.c_double),
                ("tubeSamples", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(ORACLE_SO)
        L.oracle_synthesize.restype = C.c_int
        L.oracle_synthesize.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                        C.POINTER(C.POINTER(C.c_double)), C.POINTER(OracleInfo)]
        L.oracle_derive.restype = C.c_int
        L.oracle_derive.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(OracleInfo)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_pcm16.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_int, C.c_void_p]
        L.oracle_wav_bytes.restype = C.c_void_p
        L.oracle_wav_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.POINTER(C.c_size_t)]
        L.oracle_parse_input_file.restype = C.c_int
        L.oracle_parse_input_file.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.oracle_fir_design.restype = C.c_int
        L.oracle_fir_design.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int32)]
        L.oracle_src_filter.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_noise_draws.restype = C.c_double
        L.oracle_noise_draws.argtypes = [C.c_double, C.c_size_t, C.c_void_p]
        L.oracle_amplitude.restype = C.c_double
        L.oracle_amplitude.argtypes = [C.c_double]
        L.oracle_frequency.restype = C.c_double
        L.oracle_frequency.argtypes = [C.c_double]
        L.oracle_synthesize_batch.restype = C.c_int
        L.oracle_synthesize_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                              C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_div_known_mismatches.restype = C.c_int64
        L.oracle_div_known_mismatches.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int64, C.c_uint64]
        _lib = L
    return _lib


EVENT_DTYPE = np.dtype([("time", np.int64), ("value", np.float64, (36,))])   # oracle_event


class OracleFrameGen(C.Structure):
    _fields_ = [("useMacroIntonation", C.c_int32), ("useMicroIntonation", C.c_int32), ("useSmoothIntonation", C.c_int32),
                ("useDrift", C.c_int32), ("driftDeviation", C.c_double), ("driftCutoff", C.c_double), ("pitch", C.c_double),
                ("driftSeed", C.c_float)]


def frame_count(events):
    events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
    L = lib()
    L.oracle_frame_count.restype = C.c_int64
    L.oracle_frame_count.argtypes = [C.c_void_p, C.c_int64]
    return int(L.oracle_frame_count(events.ctypes.data_as(C.c_void_p), len(events)))


def generate_frames(events, fg):
    """Control frames of one event list by the oracle's restatement of EventList.m:883-1061.  fg: any object with the
    TRMFrameGeneration fields.  Returns (frames (n, 16), drift seed at exit)."""
    events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
    L = lib()
    L.oracle_generate_frames.restype = C.c_int64
    L.oracle_generate_frames.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    o = OracleFrameGen(*[getattr(fg, n) for n, _ in OracleFrameGen._fields_])
    n = frame_count(events)
    out = np.zeros((max(n, 1), 16), np.float64)
    seed = C.c_float(0)
    got = L.oracle_generate_frames(events.ctypes.data_as(C.c_void_p), len(events), C.byref(o), out.ctypes.data_as(C.c_void_p), n,
                                   C.byref(seed))
    assert got == n
    return out[:n], seed.value


def synthetic_event_list(seed, seconds, smooth=True):
    """An event list shaped like Monet's (EventList.m applyRule / applyIntonation): posture events every 40-160 ms
    carry all 16 parameters, transition events in between carry a few, 'special' offsets (tracks 16..31) and macro
    intonation (32; with slopes 33..35 when smooth intonation is used) appear at some events; values are floats."""
    rng = np.random.default_rng(seed)
    lo = np.array([-22, 0, 0, 0, 0, 864, 500, 0.8, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.1])
    hi = np.array([-2, 60, 10, 24, 7, 5500, 4500, 0.8, 2.61, 2.61, 2.61, 2.61, 2.61, 2.61, 2.61, 1.5])
    times, vals = [0], []
    t = 0
    while t < seconds * 1000:
        t += int(rng.integers(3, 40)) * 4 + int(rng.integers(0, 4))      # not always on the 4 ms grid
        times.append(t)
    n = len(times)
    v = np.full((n, 36), np.nan)
    posture = np.zeros(n, bool)
    posture[0] = posture[-1] = True
    posture[rng.random(n) < 0.45] = True
    for i in range(n):
        if posture[i]:
            v[i, :16] = lo + (hi - lo) * rng.random(16)
        else:
            pick = rng.random(16) < 0.3
            v[i, :16][pick] = (lo + (hi - lo) * rng.random(16))[pick]
        if rng.random() < 0.2:
            k = 16 + int(rng.integers(0, 16))
            v[i, k] = (hi - lo)[k - 16] * 0.05 * (rng.random() - 0.5)
        if rng.random() < 0.3 or i == 0:
            v[i, 32] = 6.0 * (rng.random() - 0.5)
            if smooth:
                v[i, 33:36] = (rng.random(3) - 0.5) * [0.2, 0.01, 0.0005]
    v = v.astype(np.float32).astype(np.float64)                          # Monet's values are floats
    ev = np.zeros(n, EVENT_DTYPE)
    ev["time"] = times
    ev["value"] = v
    return ev


def as_oracle_ip(ip):
    """Reinterprets any struct with the TRMInputParameters layout (e.g. gnuspeech_b200.TRMInputParameters)."""
    o = OracleInputParameters()
    assert C.sizeof(o) == C.sizeof(ip)
    C.memmove(C.byref(o), C.byref(ip), C.sizeof(o))
    return o


def male_voice(outputRate=44100.0, **kw):
    """MMSynthesisParameters.m:163-187 defaults, mono, 250 Hz control rate."""
    ip = OracleInputParameters(0, outputRate, 250.0, 60.0, 1, 0.0, 0, 40.0, 16.0, 32.0, 1.0, 17.5, 25.0, 0.5, 3.05,
                               5000.0, 5000.0, (C.c_double * 6)(0, 1.35, 1.96, 1.91, 1.3, 0.73), 1500.0, 6.0, 1, 54.0)
    for k, v in kw.items():
        setattr(ip, k, v)
    return ip


class OracleResult(object):
    pass


def synthesize(ip, frames, flags=0, want_tube=True):
    frames = np.ascontiguousarray(frames, dtype=np.float64).reshape(-1, 16)
    n = frames.shape[0]
    ip = as_oracle_ip(ip)
    info = OracleInfo()
    rc = lib().oracle_derive(C.byref(ip), n, C.byref(info))
    if rc != 0:
        raise RuntimeError("oracle_derive rc=%d" % rc)
    tube = np.zeros(max(1, info.tubeSamples)) if want_tube else None
    out = C.POINTER(C.c_double)()
    rc = lib().oracle_synthesize(C.byref(ip), frames.ctypes.data_as(C.c_void_p), n, flags,
                                 tube.ctypes.data_as(C.c_void_p) if want_tube else None, C.byref(out), C.byref(info))
    if rc != 0:
        raise RuntimeError("oracle_synthesize rc=%d" % rc)
    r = OracleResult()
    r.info = info
    r.numberSamples = info.numberSamples
    r.maximumSampleValue = info.maximumSampleValue
    r.samples = np.ctypeslib.as_array(out, (max(info.numberSamples, 1),))[: info.numberSamples].copy()
    lib().oracle_free(out)
    r.tube = tube[: info.tubeSamples] if want_tube else None
    r.ip = ip
    return r


def pcm16(ip, samples, maximum, file_variant=False):
    ip = as_oracle_ip(ip)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    ch = 2 if ip.channels == 2 else 1
    out = np.zeros(max(1, samples.shape[0] * ch), dtype=np.int16)
    lib().oracle_pcm16(C.byref(ip), samples.ctypes.data_as(C.c_void_p), samples.shape[0], maximum, int(file_variant),
                       out.ctypes.data_as(C.c_void_p))
    return out[: samples.shape[0] * ch]


def wav_bytes(ip, samples, maximum):
    ip = as_oracle_ip(ip)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    ln = C.c_size_t(0)
    p = lib().oracle_wav_bytes(C.byref(ip), samples.ctypes.data_as(C.c_void_p), samples.shape[0], maximum, C.byref(ln))
    data = C.string_at(p, ln.value)
    lib().oracle_free(p)
    return data


def parse_input_file(path):
    ip = OracleInputParameters()
    fr = C.c_void_p()
    n = C.c_size_t(0)
    rc = lib().oracle_parse_input_file(path.encode(), C.byref(ip), C.byref(fr), C.byref(n))
    if rc != 0:
        raise RuntimeError("oracle_parse_input_file rc=%d" % rc)
    frames = np.ctypeslib.as_array(C.cast(fr, C.POINTER(C.c_double)), (n.value, 16)).copy()
    lib().oracle_free(fr)
    return ip, frames


def synthesize_batch(ip, frames, n_frames, flags=0, threads=1):
    """One utterance per thread; returns (numberSamples, max, checksum) arrays."""
    frames = np.ascontiguousarray(frames, dtype=np.float64).reshape(-1, 16)
    n_frames = np.ascontiguousarray(n_frames, dtype=np.int32)
    n = n_frames.shape[0]
    off = np.concatenate(([0], np.cumsum(n_frames[:-1], dtype=np.int64))).astype(np.int64)
    ns = np.zeros(n, np.int32)
    mx = np.zeros(n, np.float64)
    cs = np.zeros(n, np.float64)
    ip = as_oracle_ip(ip)
    rc = lib().oracle_synthesize_batch(C.byref(ip), 1, frames.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                       n_frames.ctypes.data_as(C.c_void_p), n, flags, threads,
                                       ns.ctypes.data_as(C.c_void_p), mx.ctypes.data_as(C.c_void_p),
                                       cs.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle_synthesize_batch rc=%d" % rc)
    return ns, mx, cs


def have_reference_binary():
    return os.path.exists(REF_BIN)


def run_reference(ip, frames):
    """Runs the compiled reference C (TRAcT/tube.c) harness in a fresh process; see oracle/ref_harness.c."""
    frames = np.ascontiguousarray(frames, dtype=np.float64).reshape(-1, 16)
    ip = as_oracle_ip(ip)
    with tempfile.TemporaryDirectory() as d:
        req, resp = os.path.join(d, "req.bin"), os.path.join(d, "resp.bin")
        with open(req, "wb") as f:
            f.write(struct.pack("<ii", 0x514D5254, frames.shape[0]))
            f.write(bytes(ip))
            f.write(frames.tobytes())
        subprocess.run([REF_BIN, req, resp], check=True, timeout=600)
        b = open(resp, "rb").read()
    cp, sr, taps, pad = struct.unpack_from("<4i", b, 0)
    nt, no = struct.unpack_from("<2q", b, 16)
    (mx,) = struct.unpack_from("<d", b, 32)
    off = 40
    tube = np.frombuffer(b, np.float64, nt, off).copy(); off += 8 * nt
    out = np.frombuffer(b, np.float32, no, off).copy(); off += 4 * no
    fir = np.frombuffer(b, np.float64, taps, off).copy()
    return dict(controlPeriod=cp, sampleRate=sr, numberTaps=taps, padSize=pad, tube=tube, out=out, fir=fir, max=mx)


def snr_db(ref, test):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    den = float(np.sum(err * err))
    num = float(np.sum(ref * ref))
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(num / den)
