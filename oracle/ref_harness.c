/*
 * ref_harness.c -- drives the REFERENCE'S OWN C copy of the TRM DSP
 * (/root/reference/Applications/TRAcT/tube.c, compiled unmodified from where it lies into
 * oracle/_ref/tube_ref.o by oracle/Makefile) in the operation order of the authoritative
 * Objective-C loop (Frameworks/Tube/TRMTubeModel.m:272-361).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied here: this file only CALLS the
 * reference's primitives (noise, noiseFilter, updateWavetable, oscillator, vocalTract, throat,
 * bandpassFilter, calculateTubeCoefficients, setFricationTaps, calculateBandpassCoefficients,
 * dataFill/flushBuffer ...) and sets its global parameters.  tube.c keeps all state in globals and
 * function-statics, so this is a one-utterance-per-process executable.
 *
 * Known differences of tube.c from Frameworks/Tube that the harness works around (Appendix D):
 *   D.1 setFricationTaps() uses 10*amplitude(fricVol): taps are divided by 10 after the call.
 *   D.2 synthesize() multiplies the signal by 100 before the SRC: the harness does not call it.
 *   D.3 updateWavetable() uses 1-(j/L)^2 instead of the vDSP order 1-(j*j)*(1/L^2): <= ~1 ulp/entry.
 *   D.5 the SRC emits (float) samples into a ring buffer: SRC cross-check is float precision.
 *
 * usage: tube_ref <request.bin> <response.bin>
 *   request : int32 magic 'TRMQ', int32 n_frames, then the 200-byte parameter block
 *             (same layout as oracle_input_parameters), then n_frames*16 doubles.
 *   response: int32 controlPeriod, int32 sampleRate, int32 numberTaps, int32 padSize,
 *             int64 n_tube, int64 n_out, double max, then n_tube doubles (tube-rate signal, before
 *             the SRC), then n_out floats (SRC output as the reference emits it), then numberTaps doubles
 *             (FIR coefficients), then 5 doubles (first noise draws of a *separate* generator are not
 *             available -- tube.c's noise() is a single static stream; the first 5 lp-noise-free draws
 *             are captured from the synthesis itself when fricVol=aspVol=0 is not required).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trm_oracle.h" /* only for the parameter struct layout */

/* ---- the reference's globals and functions (declared here, defined in tube.c / its headers) ---- */
extern int controlPeriod, sampleRate;
extern double actualTubeLength, dampingFactor, crossmixFactor, breathinessFactor;
extern float controlRate, outputRate;
extern double volume, balance, tp, tnMin, tnMax, breathiness, length, lossFactor, apScale, mouthCoef, noseCoef;
extern double noseRadius[6], throatCutoff, throatVol, mixOffset;
extern int channels, waveform, modulation, outputFileFormat;
extern double fricationTap[8];
extern double *FIRCoef;
extern int numberTaps, padSize;
extern double maximumSampleValue;
extern long int numberSamples;
extern int circBuff2Count;
extern unsigned int timeRegisterIncrement;

double amplitude(double);
double frequency(double);
void initializeWavetable(void);
void updateWavetable(double);
void initializeFIR(double, double, double);
double noise(void);
double noiseFilter(double);
void initializeMouthCoefficients(double);
void initializeNasalFilterCoefficients(double);
void initializeNasalCavity(void);
void initializeThroat(void);
void calculateTubeCoefficients(void);
void setFricationTaps(void);
void calculateBandpassCoefficients(void);
double oscillator(double);
double vocalTract(double, double);
double throat(double);
double bandpassFilter(double);
void initializeConversion(void);
void dataFill(double);
void flushBuffer(void);
void initCircBuff(void);
void initCircBuff2(void);
float getCircBuff2(void);
double *getGlotPitch(void);

/* layout of tube.c's static `current` (tube.c:292-312), reached through getGlotPitch() */
typedef struct {
    double pair[7][2];       /* value, delta for pitch, glotVol, aspVol, fricVol, fricPos, fricCF, fricBW */
    double radius[8], radiusDelta[8];
    double velum, velumDelta;
} ref_current;

static float *g_out;
static size_t g_out_n, g_out_cap;

static void drain(void)
{
    while (circBuff2Count > 0) {
        float v = getCircBuff2();
        if (g_out_n >= g_out_cap) {
            g_out_cap = g_out_cap ? g_out_cap * 2 : (1 << 16);
            g_out = (float *)realloc(g_out, g_out_cap * sizeof(float));
        }
        g_out[g_out_n++] = v;
    }
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s request.bin response.bin\n", argv[0]); return 2; }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) { perror("request"); return 2; }
    int32_t magic, n_frames;
    oracle_input_parameters ip;
    if (fread(&magic, 4, 1, fp) != 1 || fread(&n_frames, 4, 1, fp) != 1 || fread(&ip, sizeof(ip), 1, fp) != 1 ||
        magic != 0x514d5254 || n_frames < 1) { fprintf(stderr, "bad request\n"); return 2; }
    oracle_frame *frames = (oracle_frame *)malloc((size_t)n_frames * sizeof(oracle_frame));
    if (fread(frames, sizeof(oracle_frame), (size_t)n_frames, fp) != (size_t)n_frames) { fprintf(stderr, "short request\n"); return 2; }
    fclose(fp);

    /* tube.c printf()s from several primitives: silence stdout */
    if (!freopen("/dev/null", "w", stdout)) return 2;

    /* utterance-rate parameters -> the reference's globals */
    outputFileFormat = ip.outputFileFormat; outputRate = ip.outputRate; controlRate = ip.controlRate;
    volume = ip.volume; channels = ip.channels; balance = ip.balance; waveform = ip.waveform;
    tp = ip.tp; tnMin = ip.tnMin; tnMax = ip.tnMax; breathiness = ip.breathiness; length = ip.length;
    lossFactor = ip.lossFactor; apScale = ip.apScale; mouthCoef = ip.mouthCoef; noseCoef = ip.noseCoef;
    for (int i = 0; i < 6; i++) noseRadius[i] = ip.noseRadius[i];
    throatCutoff = ip.throatCutoff; throatVol = ip.throatVol; modulation = ip.usesModulation; mixOffset = ip.mixOffset;

    /* derived values: the arithmetic of initializeSynthesizer() (tube.c:595-650), which itself cannot be
       called (it spawns the real-time thread and uses a file-static temperature) */
    double c = 331.4 + (0.6 * ip.temperature);
    controlPeriod = rint((c * 10 * 100.0) / (length * controlRate));
    sampleRate = controlRate * controlPeriod;
    actualTubeLength = (c * 10 * 100.0) / sampleRate;
    double nyquist = (double)sampleRate / 2.0;
    breathinessFactor = breathiness / 100.0;
    crossmixFactor = 1.0 / amplitude(mixOffset);
    dampingFactor = (1.0 - (lossFactor / 100.0));
    initializeWavetable();
    initializeFIR(.2, .1, .00000001);
    initializeMouthCoefficients((nyquist - mouthCoef) / nyquist);
    initializeNasalFilterCoefficients((nyquist - noseCoef) / nyquist);
    initializeNasalCavity();
    initializeThroat();
    initializeConversion();
    initCircBuff();
    initCircBuff2();
    circBuff2Count = 0;

    ref_current *cur = (ref_current *)getGlotPitch();
    int64_t n_tube = (int64_t)(n_frames - 1) * controlPeriod, k = 0;
    double *tube = (double *)malloc((size_t)(n_tube > 0 ? n_tube : 1) * sizeof(double));

    for (int32_t f = 1; f < n_frames; f++) {
        const double *prev = frames[f - 1].v, *next = frames[f].v;
        for (int q = 0; q < 7; q++) {
            cur->pair[q][0] = prev[q];
            cur->pair[q][1] = (next[q] - cur->pair[q][0]) / (double)controlPeriod;
        }
        for (int q = 0; q < 8; q++) {
            cur->radius[q] = prev[7 + q];
            cur->radiusDelta[q] = (next[7 + q] - cur->radius[q]) / (double)controlPeriod;
        }
        cur->velum = prev[15];
        cur->velumDelta = (next[15] - cur->velum) / (double)controlPeriod;

        for (int j = 0; j < controlPeriod; j++) {
            double f0 = frequency(cur->pair[0][0]);
            double ax = amplitude(cur->pair[1][0]);
            double ah1 = amplitude(cur->pair[2][0]);
            calculateTubeCoefficients();
            setFricationTaps();
            for (int q = 0; q < 8; q++) fricationTap[q] /= 10.0;   /* Appendix D.1 */
            calculateBandpassCoefficients();
            double lp_noise = noiseFilter(noise());
            if (waveform == 0) updateWavetable(ax);
            double pulse = oscillator(f0);
            double pulsed_noise = lp_noise * pulse;
            pulse = ax * ((pulse * (1.0 - breathinessFactor)) + (pulsed_noise * breathinessFactor));
            double sig;
            if (modulation) {
                double crossmix = ax * crossmixFactor;
                crossmix = (crossmix < 1.0) ? crossmix : 1.0;
                sig = (pulsed_noise * crossmix) + (lp_noise * (1.0 - crossmix));
            } else
                sig = lp_noise;
            sig = vocalTract(((pulse + (ah1 * sig)) * 0.125), bandpassFilter(sig));
            sig += throat(pulse * 0.125);
            tube[k++] = sig;
            dataFill(sig);
            drain();
            /* sampleRateInterpolation() */
            for (int q = 0; q < 7; q++) cur->pair[q][0] += cur->pair[q][1];
            for (int q = 0; q < 8; q++) cur->radius[q] += cur->radiusDelta[q];
            cur->velum += cur->velumDelta;
        }
    }
    flushBuffer();
    drain();

    fp = fopen(argv[2], "wb");
    if (!fp) { perror("response"); return 2; }
    int32_t hdr[4] = {controlPeriod, sampleRate, numberTaps, padSize};
    int64_t cnt[2] = {n_tube, (int64_t)g_out_n};
    fwrite(hdr, 4, 4, fp);
    fwrite(cnt, 8, 2, fp);
    fwrite(&maximumSampleValue, 8, 1, fp);
    fwrite(tube, 8, (size_t)n_tube, fp);
    fwrite(g_out, 4, g_out_n, fp);
    fwrite(FIRCoef, 8, (size_t)numberTaps, fp);
    fclose(fp);
    return 0;
}
