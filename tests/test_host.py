"""CPU tests of the host side: the C-ABI libraries load and export every symbol the headers declare, the host
library's derived values / parser / planning agree with the oracle, error behaviour mirrors the reference, and
without a CUDA device the synthesis path fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _g():
    import gnuspeech_b200 as g
    return g


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:TRM|trm_cuda_)\w+)\s*\(", txt)))


def _exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "gnuspeech_b200", "lib", lib)],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    return set(ln.split()[-1] for ln in out.splitlines() if " T " in ln)


def test_c_abi_exports_every_declared_symbol():
    _g()
    exp_host, exp_cuda = _exported("libtrm.so"), _exported("libtrm_cuda.so")
    decl = _declared("trm.h") + _declared("trm_workload.h")
    assert len(decl) > 45
    missing = [s for s in decl if s not in exp_host]
    assert not missing, missing
    decl_cuda = _declared("trm_cuda.h")
    assert len(decl_cuda) >= 14
    missing = [s for s in decl_cuda if s not in exp_cuda]
    assert not missing, missing
    # the product never links the oracle
    ldd = subprocess.run(["ldd", os.path.join(ROOT, "gnuspeech_b200", "lib", "libtrm.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in ldd


def test_struct_layouts_match_headers():
    g = _g()
    from gnuspeech_b200 import _native as N
    assert C.sizeof(N.TRMInputParametersStruct) == 208 == C.sizeof(O.OracleInputParameters)
    assert g.TRMParameters().values.nbytes == 128


@pytest.mark.parametrize("kw", [dict(outputRate=44100.0), dict(outputRate=22050.0), dict(outputRate=22050.0, length=15.0),
                                dict(outputRate=44100.0, length=7.5), dict(outputRate=22050.0, length=12.5, temperature=32.0)])
@pytest.mark.parametrize("nframes", [0, 1, 2, 251, 15001])
def test_derived_values_match_oracle(kw, nframes):
    g = _g()
    ip = g.TRMInputParameters(**kw)
    dv = g.derive(ip, nframes)
    info = O.OracleInfo()
    oip = O.as_oracle_ip(ip)
    assert O.lib().oracle_derive(C.byref(oip), nframes, C.byref(info)) == 0
    assert (dv.controlPeriod, dv.sampleRate, dv.padSize, dv.timeRegisterIncrement, dv.tubeSamples, dv.numberSamples) == (
        info.controlPeriod, info.sampleRate, info.padSize, info.timeRegisterIncrement, info.tubeSamples, info.numberSamples)
    assert dv.actualTubeLength == info.actualTubeLength


def test_defaults_are_monets_male_voice():
    g = _g()
    ip = g.TRMInputParameters(44100.0)
    ref = O.male_voice(44100.0)
    assert bytes(O.as_oracle_ip(ip)) == bytes(ref)


def test_data_list_parser_and_writer(tmp_path):
    g = _g()
    path = os.path.join(GOLDEN, "gnuspeech.input")
    dl = g.TRMDataList(path)
    oip, oframes = O.parse_input_file(path)
    assert dl.count == 344 and np.array_equal(dl.values, oframes)
    assert bytes(O.as_oracle_ip(dl.inputParameters)) == bytes(oip)
    out = str(tmp_path / "roundtrip.input")
    dl.writeToFile(out)
    dl2 = g.TRMDataList(out)
    assert dl2.count == 345                                   # the parser doubles the last line again
    assert np.array_equal(dl2.values[:344], oframes)           # %.3f is what Monet writes (TRMParameters.m:26-45)
    with pytest.raises(g.TRMError) as e:
        g.TRMDataList(str(tmp_path / "missing.input"))
    assert e.value.code == -6
    dl.removeAllParameters()
    assert dl.count == 0
    dl.addParameters(g.TRMParameters(glottalPitch=-12, radius=[0.8] * 8, velum=0.1))
    assert dl.count == 1 and dl.values[0, 0] == -12 and dl.values[0, 15] == 0.1


def test_error_behaviour_mirrors_reference():
    g = _g()
    dl = g.TRMDataList()
    dl.inputParameters.length = 0.0                           # -initWithInputData: returns nil (TRMTubeModel.m:204-207)
    assert g.TRMTubeModel.initWithInputData(dl) is None
    with pytest.raises(g.TRMError) as e:
        g.TRMTubeModel(dl)
    assert e.value.code == -1
    dl.inputParameters.length = 17.5
    dl.inputParameters.tnMin = 0.0                            # closure point would leave the table: documented deviation
    with pytest.raises(g.TRMError) as e:
        g.TRMTubeModel(dl)
    assert e.value.code == -4
    dl.inputParameters.tnMin = 16.0
    dl.inputParameters.channels = 3
    with pytest.raises(g.TRMError) as e:
        g.TRMTubeModel(dl)
    assert e.value.code == -4
    dl.inputParameters.channels = 1
    # rate pairs whose converter window (2*(padSize+1)+3 halo rows + the outputs' span) cannot fit the TRM_SRC_ROWS = 128
    # rows the kernel stages are refused when the descriptor is derived; a 5 cm tube at 22.05 kHz still fits
    dl.inputParameters.outputRate = 22050.0
    dl.inputParameters.length = 4.0                           # 87.5 kHz tube rate: ratio 0.25, 52 pad samples per wing
    with pytest.raises(g.TRMError) as e:
        g.TRMTubeModel(dl)
    assert e.value.code == -4
    dl.inputParameters.length = 5.0
    assert g.TRMTubeModel(dl) is not None
    dl.inputParameters.length = 17.5
    dl.inputParameters.outputRate = 44100.0
    m = g.TRMTubeModel(dl)                                    # zero frames: synthesize returns without output
    m.synthesize()
    assert m.numberSamples == 0
    with pytest.raises(g.TRMError) as e:
        m.synthesize()                                        # single-use, like the reference's model
    assert e.value.code == -7
    with pytest.raises(g.TRMError) as e:
        m.generateWAVData()                                   # NSParameterAssert(maximumSampleValue != 0)
    assert e.value.code == -8


def test_batch_layout_and_offsets():
    g = _g()
    ip = g.TRMInputParameters(44100.0)
    nf = [251, 2, 1, 501, 126]
    b = g.TRMBatch(ip, nf)
    ns = b.numberSamples
    for u, n in enumerate(nf):
        assert ns[u] == g.derive(ip, n).numberSamples
    po, oo, to = b.pcmOffsets, b.outOffsets, b.tubeOffsets
    assert all(x % 32 == 0 for x in list(po) + list(oo) + list(to))
    for u in range(len(nf) - 1):
        assert oo[u + 1] >= oo[u] + ns[u] and po[u + 1] >= po[u] + ns[u]
    lay = b.layout
    assert lay.total_frames == sum(nf) and lay.out_samples == int(ns.sum())
    assert abs(lay.audio_seconds - sum(max(n - 1, 0) for n in nf) / 250.0) < 1e-12
    # per-utterance parameters, stereo doubles the PCM footprint
    ips = [g.TRMInputParameters(44100.0), g.TRMInputParameters(22050.0, channels=2), g.TRMInputParameters(22050.0, length=10.0)]
    b2 = g.TRMBatch(ips, [51, 51, 51])
    assert b2.pcmOffsets[2] - b2.pcmOffsets[1] >= 2 * b2.numberSamples[1]
    with pytest.raises(g.TRMError):
        g.TRMBatch([g.TRMInputParameters(44100.0, length=-1.0)], [10])


def test_workload_generators_are_deterministic_and_in_range():
    from gnuspeech_b200 import workloads as W
    a = W.random_walk(6, 40, seed=3, threads=1)
    b = W.random_walk(6, 40, seed=3, threads=4)
    assert np.array_equal(a, b)
    c = W.random_walk(3, 40, seed=3, first_index=3)
    assert np.array_equal(a[3 * 40:], c)                      # utterance streams are keyed by absolute index
    assert np.array_equal(a, a.astype(np.float32).astype(np.float64))   # float-rounded, as Monet produces them
    lo = np.array([-22, 0, 0, 0, 0, 864, 500, 0.8, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.1]) - 1e-5
    hi = np.array([-2, 60, 10, 24, 7, 5500, 4500, 0.8, 2.61, 2.61, 2.61, 2.61, 2.61, 2.61, 2.61, 1.5]) + 1e-3
    assert (a >= lo).all() and (a <= hi).all()
    gr = W.grid([0, 1, 4 ** 7 - 1, 2 ** 14, 2 ** 15], 3)
    assert gr.shape == (15, 16) and gr[0, 8] == np.float32(0.4) and gr[3, 8] == np.float32(0.9)
    assert gr[9, 15] == np.float32(0.8) and gr[12, 0] == -5.0
    sv = W.static_vowel(4, 1)
    assert sv[0, 14] == 2.61 and (sv == sv[0]).all()


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly -- never synthesize on the CPU."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    dl = g.TRMDataList()
    dl.addParameters(W.static_vowel(3, 0))
    m = g.TRMTubeModel(dl)
    with pytest.raises(g.TRMError) as e:
        m.synthesize()
    assert e.value.code == -5
    b = g.TRMBatch(g.TRMInputParameters(), [3])
    with pytest.raises(g.TRMError) as e:
        b.synthesize(W.static_vowel(3, 0), pcm_out=np.zeros(b.layout.total_pcm_samples, np.int16))
    assert e.value.code == -5


def test_voice_parameters_defaults_header_and_named_voices(tmp_path):
    """MMSynthesisParameters layer: the registered defaults are the male voice; -parameterString
    reproduces the 26 header lines of the reference's own sample input (Applications/Monet/samples/gnuspeech.input,
    committed as tests/golden/gnuspeech.input) byte for byte; the header round-trips through the TRM file parser into
    the TRMInputParameters that -setupSynthesisParameters: builds; the named voices of Other/voices.config give the
    tube rates of SURVEY 8 (control period / sample rate follow from the tract length)."""
    g = _g()
    sp = g.MMSynthesisParameters()
    assert (sp.masterVolume, sp.vocalTractLength, sp.temperature, sp.pitch) == (60.0, 17.5, 25.0, -12.0)
    assert (sp.tp, sp.tnMin, sp.tnMax, sp.glottalPulseShape, sp.shouldUseNoiseModulation) == (40.0, 16.0, 32.0, 0, 1)
    assert (sp.samplingRate, sp.outputChannels) == (1, 1)               # 44.1 kHz stereo are Monet's defaults
    # the reference's sample file was written with 22.05 kHz mono
    sp.samplingRate, sp.outputChannels = 0, 0
    want = open(os.path.join(GOLDEN, "gnuspeech.input")).read().split("\n")[:26]
    got = sp.parameterString.split("\n")
    # line 2: the sample file predates the "%g" the current -parameterString uses for the rate (MMSynthesisParameters.m:283)
    assert got[1] == "22050\t\t; output sample rate (22050.0, 44100.0)" and want[1].startswith("22050.000000")
    assert got[:1] + got[2:] == want[:1] + want[2:]
    # round trip: header text -> TRM file parser == -setupSynthesisParameters:
    path = tmp_path / "voice.input"
    path.write_text(sp.parameterString + "\n" + " ".join(["0.0"] * 16) + "\n")
    dl = g.TRMDataList.initWithContentsOfFile(str(path))
    a, b = dl.inputParameters, sp.inputParameters()
    for name, _ in a._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        if name == "noseRadius":
            assert list(va)[1:] == list(vb)[1:]
        else:
            assert va == vb, name
    assert b.channels == 1 and b.controlRate == 250.0 and b.noseRadius[0] == 0.0
    # named voices
    want_rates = {"Male": (79, 19750), "Female": (92, 23000), "LgChild": (111, 27750), "SmChild": (139, 34750), "Baby": (185, 46250)}
    for name, (cp, sr) in want_rates.items():
        v = g.MMSynthesisParameters(name.lower())
        d = g.derive(v.inputParameters(), 251)
        assert (d.controlPeriod, d.sampleRate) == (cp, sr), name
    assert g.MMSynthesisParameters("Female").tnMin == 32.0 and g.MMSynthesisParameters("Baby").pitch == 7.5
    with pytest.raises(g.TRMError):
        g.MMSynthesisParameters("tenor")


def test_reference_flush_bug_is_flagged():
    """TRMSampleRateConverter.m:160-168: when DOWN-sampling, a final drain that finds fewer new
    inputs than the previous pass overshot by makes the reference run over a whole ring of stale data and append spurious
    samples.  The oracle restates the streaming converter and reproduces that; this implementation computes the defined
    output and FLAGS the condition: flag set <=> the reference's sample count differs from the closed form."""
    g = _g()
    from gnuspeech_b200 import _native as N
    from gnuspeech_b200 import workloads as W
    seen = 0
    for kw in (dict(length=7.5805, temperature=32.0), dict(length=10.0, temperature=32.0), dict(length=6.5)):
        ip = g.TRMInputParameters(22050.0, **kw)
        lengths = list(range(2, 260, 7)) + [209, 210, 211]
        b = g.TRMBatch(ip, lengths)
        flags = b.referenceFlushBugFlags
        for k, nf in enumerate(lengths):
            assert bool(N.lib().TRMReferenceFlushBug(C.byref(ip), nf)) == bool(flags[k])
            ref = O.synthesize(ip, W.static_vowel(nf, 0), want_tube=False)
            assert (ref.numberSamples != b.numberSamples[k]) == bool(flags[k]), (kw, nf, ref.numberSamples, b.numberSamples[k])
            seen += int(flags[k])
    assert seen >= 1                                                         # (7.58 cm tract, 210 frames: +488 samples)
    up = g.TRMInputParameters(44100.0)                                       # up-sampling is immune
    assert not any(N.lib().TRMReferenceFlushBug(C.byref(up), nf) for nf in range(2, 400))
