/*
 * trm_cuda.h -- thin C-ABI shim between the C host library (libtrm) and the sm_100a kernels
 * (libtrm_cuda).  Plain pointers and sizes only.
 *
 * What each entry point replaces in the reference (all under /root/reference/Frameworks/Tube/):
 *   trm_cuda_stage TRM_STAGE_TUBE  -> the sample-rate loop of -[TRMTubeModel synthesize]
 *                                     TRMTubeModel.m:292-354 (+611-688, 712-853), TRMWavetable.m:117-195,
 *                                     TRMFIRFilter.m:116-146, TRMFilters.m:9-86, TRMUtility.m:26-47,71-85
 *   trm_cuda_stage TRM_STAGE_SRC   -> -[TRMSampleRateConverter processDataFromRingBuffer:]
 *                                     TRMSampleRateConverter.m:155-298 + TRMRingBuffer.m:47-93 (stateless form)
 *   trm_cuda_stage TRM_STAGE_PCM   -> the scaling loops of -saveOutputToFile: / -generateWAVData
 *                                     TRMTubeModel.m:370-383,420-484 / 515-559
 * The derived constants in trm_cuda_utterance are computed by the host library exactly as
 * -initWithInputData: does (TRMTubeModel.m:196-241).
 */
#ifndef TRM_CUDA_H
#define TRM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRM_FIR_MAX_TAPS     64
#define TRM_SRC_FILTER_LEN   3328      /* 13 zero crossings x 256 phases (TRMSampleRateConverter.m:11-21) */
#define TRM_TABLE_LENGTH     512       /* TRMWavetable.m:22 */
#define TRM_NOISE_JUMP       16        /* noise powers 377^0..377^16 mod 2^44 */
#define TRM_ALIGN_ELEMS      32        /* every per-utterance buffer offset is a multiple of this */
#define TRM_SRC_ROWS         128       /* tube-rate samples per utterance the converter kernel stages for one work item;
                                          bounds the rate ratios libtrm accepts (an item needs 2*(padSize+1)+3 rows of
                                          halo plus its outputs' span) */

enum { TRM_STAGE_TUBE = 0, TRM_STAGE_SRC = 1, TRM_STAGE_PCM = 2, TRM_STAGE_COUNT = 3 };

/* One utterance as the kernels see it.  Offsets are element offsets into the job's buffers. */
typedef struct trm_cuda_utterance {
    int64_t  frame_offset;      /* first control frame (unit: frames of 16 doubles)                     */
    int64_t  tube_offset;       /* first tube-rate sample                                              */
    int64_t  out_offset;        /* first output-rate sample (also PCM frame offset)                     */
    int64_t  pcm_offset;        /* first int16 element of this utterance in the PCM buffer             */
    int64_t  n_tube;            /* (n_frames-1)*controlPeriod                                          */
    int64_t  n_out;             /* numberSamples                                                       */
    int32_t  n_frames;
    int32_t  controlPeriod;
    int32_t  waveform;          /* 0 pulse, 1 sine                                                     */
    int32_t  usesModulation;
    int32_t  voice;             /* row of the wavetable array                                          */
    int32_t  padSize;
    int32_t  upsample;          /* sampleRateRatio >= 1                                                */
    int32_t  channels;
    int32_t  div1, div2;        /* tableDiv1 / tableDiv2 (TRMWavetable.m:71-72)                        */
    uint32_t tri;               /* timeRegisterIncrement                                               */
    uint32_t phaseIncrement;
    double   sampleRate;        /* (double)_sampleRate                                                 */
    double   sampleRateRatio;
    double   dampingFactor;
    double   breathinessFactor;
    double   crossmixFactor;
    double   basicIncrement;
    double   tnDelta;
    double   mouth[5];          /* a10 b11 a20 a21 b21 (TRMFilters.m:34-45)                            */
    double   nose[5];
    double   nasal_coeff[5];    /* NC2..NC6 (TRMTubeModel.m:692-707)                                   */
    double   nr1sq;             /* noseRadius[1]^2 (TRMTubeModel.m:741)                                */
    double   apScale2;          /* apScale^2 (TRMTubeModel.m:724)                                      */
    double   ta0, tb1;          /* throat low-pass (TRMFilters.m:64-68)                                */
    double   throatGain;
    double   volumeAmp;         /* amplitude(volume)                                                   */
    double   leftGain, rightGain; /* stereo factors applied on top of scale (TRMTubeModel.m:532-533)   */
    /* streaming (trm_cuda_stream_*): a call continues an utterance.  All zero for whole utterances.              */
    int64_t  out_start;         /* first output-rate sample this call produces (earlier ones exist already)       */
    int64_t  in_start;          /* first tube-rate sample still in memory (multiple of 4; earlier ones are gone)  */
    int32_t  jc0;               /* tube samples already produced inside the first control interval of `frames`    */
    int32_t  reserved;
} trm_cuda_utterance;

/* Constant tables shared by all utterances; computed on the host by libtrm. */
typedef struct trm_cuda_tables {
    int32_t  fir_taps;
    double   fir_coef[TRM_FIR_MAX_TAPS];
    double   src_h[TRM_SRC_FILTER_LEN];
    double   src_dh[TRM_SRC_FILTER_LEN];
    uint64_t noise_pow[TRM_NOISE_JUMP + 1];   /* 377^i mod 2^44                                        */
    uint64_t noise_k0;                         /* state before the first draw: k1 * 377^-1 mod 2^44     */
} trm_cuda_tables;

/* Control-frame generator inputs (EventList.m:883-1061; include/trm.h TRMEvent / TRMFrameGeneration have the same layout) */
#define TRM_EVENT_VALUES 36
typedef struct trm_cuda_event { int64_t time; double value[TRM_EVENT_VALUES]; } trm_cuda_event;
typedef struct trm_cuda_framegen {
    int32_t useMacroIntonation, useMicroIntonation, useSmoothIntonation, useDrift;
    double  driftDeviation, driftCutoff, pitch;
    float   driftSeed;
    int32_t reserved;
} trm_cuda_framegen;

typedef struct trm_cuda_ctx trm_cuda_ctx;
typedef struct trm_cuda_resident trm_cuda_resident;

int         trm_cuda_device_count(void);
const char *trm_cuda_last_error(void);
void       *trm_cuda_host_alloc(size_t bytes);      /* pinned */
void        trm_cuda_host_free(void *p);

/* One context per (process, device): uploads the tables, owns streams and scratch. */
int  trm_cuda_ctx_create(int device, const trm_cuda_tables *tables, trm_cuda_ctx **ctx);
void trm_cuda_ctx_destroy(trm_cuda_ctx *ctx);
/* Glottal wavetables as built at init time (TRMWavetable.m:78-102), n_voices x 512 doubles; utterances pick a
 * row with trm_cuda_utterance.voice.  Replaces the previous set (blocking). */
int  trm_cuda_set_wavetables(trm_cuda_ctx *ctx, const double *tables, int n_voices);

/*
 * Host-buffer path (what TRMTubeModelSynthesize / TRMBatchSynthesize call): copies the frames of the
 * n utterances described by desc[] (offsets relative to the host arrays) to the device in chunks,
 * runs the three stages, copies PCM / samples / max back.  Chunks are double-buffered over CUDA
 * streams so copies overlap kernels.  Blocking.
 *   pcm_host     int16 [..]  or NULL      samples_host  double/float [..] or NULL
 *   max_host     double[n]   or NULL      tube_host     double/float tube-rate signal or NULL
 *   launches     receives the number of kernel launches issued (may be NULL)
 */
int trm_cuda_synthesize_host(trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc,
                             const double *frames_host, int16_t *pcm_host, void *samples_host,
                             double *max_host, void *tube_host, int64_t *launches);

/* Same call; `enqueued(arg)` (may be NULL) is invoked once, from the calling thread, as soon as the first chunk's copies
 * and kernels are in the device queues -- libtrm's asynchronous tickets use it to start concurrent calls in submission
 * order (the queues are shared per device, so the order of arrival there is the order of execution). */
int trm_cuda_synthesize_host_ex(trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc,
                                const double *frames_host, int16_t *pcm_host, void *samples_host,
                                double *max_host, void *tube_host, int64_t *launches,
                                void (*enqueued)(void *), void *arg);

/* The same call with the frame format given: 0 = TRMParameters rows of 16 doubles (128 bytes), 1 = rows of 16 floats
 * (64 bytes; what Monet's generator holds, EventList.m:968-1002), staged as they are and widened by the waveguide kernel
 * when a parameter lane reads its value -- half the upload, identical results. */
int trm_cuda_synthesize_host_fmt(trm_cuda_ctx *ctx, int precision, int frame_format, int n, const trm_cuda_utterance *desc,
                                 const void *frames_host, int16_t *pcm_host, void *samples_host,
                                 double *max_host, void *tube_host, int64_t *launches,
                                 void (*enqueued)(void *), void *arg);

/*
 * Device-resident path (bench `value`, roofline timing): inputs uploaded once, stages launched on the
 * caller's stream (a cudaStream_t passed as void*, NULL = the legacy default stream) without any
 * host<->device copy, so they can be bracketed by CUDA events on that stream.
 */
int  trm_cuda_resident_create(trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc,
                              const double *frames_host, trm_cuda_resident **res);
void trm_cuda_resident_destroy(trm_cuda_resident *res);
int  trm_cuda_resident_stage(trm_cuda_resident *res, int stage, void *stream);   /* one kernel stage      */
int  trm_cuda_resident_run(trm_cuda_resident *res, void *stream);                /* all stages, in order  */
int  trm_cuda_resident_fetch(trm_cuda_resident *res, int16_t *pcm_host, void *samples_host,
                             double *max_host, void *tube_host);                 /* blocking D2H          */
int  trm_cuda_resident_fetch_utterance(trm_cuda_resident *res, int u, void *samples_host, int16_t *pcm_host,
                                       double *max_host);                        /* one utterance, blocking */
int  trm_cuda_stage_launches(int stage);     /* kernel launches one stage issues */
/* Sweep (BASELINE configs[4]): n utterances of one voice x n_frames frames, control tracks = the walk2 workload
 * (include/trm_workload.h) generated on the device (utterance k: index first_index + k of stream seed), synthesized in
 * chunks of one full wave; per utterance an 8-byte PCM checksum sum(pcm[i] * (2 i + 1)) mod 2^64 and the maximum come back,
 * plus the PCM of the utterances listed in probe_utt (sorted; rows of probe_stride int16).  kernel_ms: device time of
 * all chunks (CUDA events around generator + 3 stages + checksum). */
int  trm_cuda_sweep(trm_cuda_ctx *ctx, int precision, const trm_cuda_utterance *voice, int32_t n_frames, uint64_t seed,
                    uint64_t first_index, int64_t n, uint64_t *checksums_host, double *max_host, int64_t n_probe,
                    const int64_t *probe_utt, int16_t *probe_pcm, int64_t probe_stride, int64_t *launches, double *kernel_ms);
/* Copy-only probe: h2d_bytes up and d2h_bytes down concurrently, `reps` times; *ms = average per repetition. */
int  trm_cuda_copy_probe(int device, const void *host_in, size_t h2d_bytes, void *host_out, size_t d2h_bytes, int reps,
                         double *ms);

/*
 * Streaming synthesis (TRAcT-style, Applications/TRAcT/tube.c:1096-1191: output as it is produced, no normalisation):
 * n_streams voices with the SAME rates (one trm_cuda_utterance template: controlPeriod, converter signature ...) are
 * advanced together.  Every push appends `m` control frames per stream (frames_host: [stream][m][16] doubles) and
 * returns the un-normalised output-rate samples that became computable (samples_host: [stream][*n_samples], row
 * stride = capacity given at creation; double or float by precision).  Calls cover whole 16-sample blocks of the
 * waveguide, the state of every recurrence is carried on the device: the concatenation of all pushes plus the flush
 * is bit-identical to synthesizing the whole utterance at once.  flush != 0 ends the streams (converter tail, as
 * -synthesize does at the end of an utterance); the object can then be destroyed.
 */
typedef struct trm_cuda_stream trm_cuda_stream;
int  trm_cuda_stream_create(trm_cuda_ctx *ctx, int precision, int n_streams, const trm_cuda_utterance *voice,
                            int max_frames_per_push, trm_cuda_stream **out);
int64_t trm_cuda_stream_capacity(const trm_cuda_stream *s);          /* samples per stream a push can return */
int  trm_cuda_stream_push(trm_cuda_stream *s, const double *frames_host, int m, int flush, void *samples_host,
                          int64_t *n_samples);
void trm_cuda_stream_destroy(trm_cuda_stream *s);
/*
 * Control frames on the device (replaces -[EventList generateOutputInTimeRange:...], EventList.m:883-1061): uploads
 * the event lists of the n utterances (utterance u: events[ev_offset[u] .. +ev_count[u]); fg: one entry per
 * utterance, or one for all if shared_fg), runs the generator kernel and leaves the frames -- desc[u].n_frames of
 * them at frame desc[u].frame_offset -- in a device buffer owned by the context.  *frames_dev receives that buffer
 * (valid until the next call on this context; usable as the `frames_host` argument of trm_cuda_synthesize_host,
 * which copies with cudaMemcpyDefault).  frames_host / seed_host, if not NULL, receive copies.  Blocking.
 */
int  trm_cuda_generate_frames(trm_cuda_ctx *ctx, int n, const trm_cuda_utterance *desc, const trm_cuda_event *events,
                              const int64_t *ev_offset, const int32_t *ev_count, const trm_cuda_framegen *fg, int shared_fg,
                              const double **frames_dev, double *frames_host, float *seed_host);
/* Measures the device's FMA peak (TFLOP/s, 2 flops per FMA) with a register-resident FMA chain:
 * precision 0 = FP64, 1 = FP32.  The waveguide kernel's roofline denominator. */
int  trm_cuda_fp_peak(int device, int precision, int reps, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* TRM_CUDA_H */
