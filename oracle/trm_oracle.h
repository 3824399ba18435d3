/*
 * trm_oracle.h -- CPU ORACLE for the Tube Resonance Model (TRM) synthesis loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the algorithm in
 * the reference's Frameworks/Tube sources (.m, Objective-C, not compilable here).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (gnuspeech_b200/, libtrm.so, libtrm_cuda.so)
 * never links, imports or calls anything in oracle/.
 *
 * Parity pinning: the reference ships NO tests and NO golden vectors (SURVEY.md
 * section 4).  The oracle is therefore pinned against outputs of the reference's own
 * plain-C copy of the same DSP, Applications/TRAcT/tube.c, compiled unmodified
 * from /root/reference into oracle/_ref/ (see oracle/Makefile, oracle/ref_harness.c)
 * and against known-answer values measured from that compiled reference
 * (tests/golden/).
 *
 * Struct layouts here intentionally match include/trm.h field for field so the
 * same byte buffers can be handed to both sides by the tests, but the two
 * headers are independent.
 */
#ifndef TRM_ORACLE_H
#define TRM_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Utterance-rate parameters: mirrors TRMInputParameters.h:26-54. */
typedef struct {
    int32_t outputFileFormat;   /* 0 = AU, 1 = AIFF, 2 = WAVE */
    float   outputRate;         /* 22050 / 44100 */
    float   controlRate;        /* control frames per second */
    double  volume;             /* master volume 0-60 dB */
    int32_t channels;           /* 1 or 2 */
    double  balance;            /* -1..+1 */
    int32_t waveform;           /* 0 = pulse, 1 = sine */
    double  tp, tnMin, tnMax;   /* glottal pulse shape, % of period */
    double  breathiness;        /* % */
    double  length;             /* nominal tube length, cm */
    double  temperature;        /* deg C */
    double  lossFactor;         /* % */
    double  apScale;            /* aperture scaling radius, cm */
    double  mouthCoef, noseCoef;/* aperture coefficients, Hz */
    double  noseRadius[6];      /* [0] unused, TRMTubeModel.m:695-706 */
    double  throatCutoff, throatVol;
    int32_t usesModulation;
    double  mixOffset;
} oracle_input_parameters;

/* One control frame: mirrors TRMParameters.h:9-17 (16 doubles = 128 bytes).
 * v[0]=glottalPitch v[1]=glottalVolume v[2]=aspirationVolume v[3]=fricationVolume
 * v[4]=fricationPosition v[5]=fricationCenterFrequency v[6]=fricationBandwidth
 * v[7..14]=radius[0..7]  v[15]=velum */
typedef struct { double v[16]; } oracle_frame;

typedef struct {
    int32_t controlPeriod;
    int32_t sampleRate;
    double  actualTubeLength;
    int32_t numberTaps;          /* FIR taps */
    int32_t padSize;             /* SRC ring pad */
    uint32_t timeRegisterIncrement;
    int32_t numberSamples;       /* output-rate sample frames */
    double  maximumSampleValue;
    double  finalNoiseSeed;      /* noise generator seed after the last draw */
    int64_t tubeSamples;         /* tube-rate samples synthesized */
} oracle_result_info;

/* flags for oracle_synthesize */
#define ORACLE_WAVETABLE_ANALYTIC 1   /* evaluate glottal table on lookup instead of rewriting it each sample
                                         (identical values; the reference rewrites, TRMWavetable.m:117-162) */
#define ORACLE_SRC_STATELESS      2   /* use the closed-form gather SRC instead of the streaming ring buffer */

/*
 * Run the whole path for one utterance: TRMTubeModel -initWithInputData: + -synthesize
 * (TRMTubeModel.m:186-260, 272-361).
 *   tube_out  : optional, receives (n_frames-1)*controlPeriod tube-rate samples (caller sized), may be NULL
 *   out       : *out is malloc'ed and receives numberSamples doubles at the output rate (caller frees with oracle_free)
 * Returns 0 on success, <0 on error (-1 illegal tube length, -2 FIR design failure, -3 allocation failure).
 */
int oracle_synthesize(const oracle_input_parameters *ip, const oracle_frame *frames, size_t n_frames,
                      int flags, double *tube_out, double **out, oracle_result_info *info);

void oracle_free(void *p);

/* Derived values only (no synthesis): controlPeriod/sampleRate/... and the predicted SRC output count
 * for n_frames frames (stateless formula). */
int oracle_derive(const oracle_input_parameters *ip, size_t n_frames, oracle_result_info *info);

/* Output stage: mono or stereo int16 PCM (TRMTubeModel.m:370-383 file variant when file_variant != 0,
 * 515-557 WAV variant otherwise).  dst holds numberSamples*channels int16. */
void oracle_pcm16(const oracle_input_parameters *ip, const double *samples, int32_t numberSamples,
                  double maximumSampleValue, int file_variant, int16_t *dst);

/* -generateWAVData (TRMTubeModel.m:509-593): returns malloc'ed RIFF bytes, *len set. */
uint8_t *oracle_wav_bytes(const oracle_input_parameters *ip, const double *samples, int32_t numberSamples,
                          double maximumSampleValue, size_t *len);

/* TRM input file parser (TRMDataList.m:43-247); frames malloc'ed (last frame duplicated as the reference does). */
int oracle_parse_input_file(const char *path, oracle_input_parameters *ip, oracle_frame **frames, size_t *n_frames);

/* Primitives exposed for primitive-level cross-checks against oracle/_ref. */
int    oracle_fir_design(double beta, double gamma, double cutoff, double *coef /*>=401*/, int32_t *numberTaps);
void   oracle_src_filter(double *h /*3328*/, double *deltaH /*3328*/);
double oracle_noise_draws(double seed, size_t n, double *out /* may be NULL */);   /* returns final seed */
double oracle_amplitude(double dB);
double oracle_frequency(double pitch);
double oracle_izero(double x);

/* One utterance per thread on n_threads host threads (CPU baseline).
 * Utterance u uses ip[u] (or ip[0] if shared_ip != 0) and frames[frame_offset[u] .. +n_frames[u]).
 * Outputs only numberSamples / maximumSampleValue / a checksum per utterance. */
int oracle_synthesize_batch(const oracle_input_parameters *ip, int shared_ip,
                            const oracle_frame *frames, const int64_t *frame_offset, const int32_t *n_frames,
                            int n_utterances, int flags, int n_threads,
                            int32_t *numberSamples, double *maximumSampleValue, double *checksum);

/* ------------------------------------------------------------------------------------------------
 * Control-frame generator: -[EventList generateOutputInTimeRange:forSynthesizer:parameterLogger:]
 * (Frameworks/GnuSpeech/MonetModel/EventList.m:883-1061, full time range) + MMDriftGenerator
 * (MMDriftGenerator.m:41-78).  SURVEY.md 8(f) rank 1.  Reference tests / golden vectors for it: none -> parity unpinned.
 * ---------------------------------------------------------------------------------------------- */
#define ORACLE_EVENT_VALUES 36
typedef struct { int64_t time; double value[ORACLE_EVENT_VALUES]; } oracle_event;   /* Event.m: time in ms, NaN = unset */
typedef struct {
    int32_t useMacroIntonation, useMicroIntonation, useSmoothIntonation, useDrift;   /* MMIntonation.m:74-80 */
    double  driftDeviation, driftCutoff;
    double  pitch;                 /* model.synthesisParameters.pitch, added to every frame's glottal pitch (m:983) */
    float   driftSeed;             /* MMDriftGenerator seed at entry (0.7892347 for a fresh generator) */
} oracle_framegen;
/* number of frames the loop emits */
int64_t oracle_frame_count(const oracle_event *events, int64_t n_events);
/* writes at most max_frames frames, returns the number the loop emits; *seed_out = drift seed at exit (may be NULL) */
int64_t oracle_generate_frames(const oracle_event *events, int64_t n_events, const oracle_framegen *fg,
                               oracle_frame *out, int64_t max_frames, float *seed_out);

/* ------------------------------------------------------------------------------------------------
 * Checker for a kernel shortcut, not reference behaviour: the CUDA feed-forward code divides by divisors known in
 * advance (20, 12, the tube sample rate) as q = a*rc, r = fma(-c, q, a), q' = fma(r, rc, q) with rc = RN(1/c)
 * (div_known(), gnuspeech_b200/csrc/tube_kernel.cuh).  Returns how many of n pseudo-random numerators in [lo, hi)
 * give a result different from the IEEE division a / c (expected: 0).
 * ---------------------------------------------------------------------------------------------- */
int64_t oracle_div_known_mismatches(double c, double lo, double hi, int64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
