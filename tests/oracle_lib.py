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
        _lib = L
    return _lib


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
