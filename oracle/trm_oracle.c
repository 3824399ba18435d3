/*
 * trm_oracle.c -- CPU ORACLE (test infrastructure, see trm_oracle.h).
 *
 * Restates, operation by operation and in the same order, the arithmetic of the
 * reference's Frameworks/Tube (Objective-C).  Every function cites the reference
 * file:line it follows.  Build with -O2 -ffp-contract=off (the reference build
 * has neither FMA contraction nor fast-math).
 *
 * Parity pinning: cross-checked against the compiled reference C copy
 * Applications/TRAcT/tube.c (oracle/_ref, oracle/ref_harness.c) and the
 * known-answer values in tests/golden/ that were produced by that harness.
 */
#include "trm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * TRMUtility.m
 * ---------------------------------------------------------------------------------------------- */

/* TRMUtility.m:20-23 */
static double speed_of_sound(double t_celsius) { return 331.4 + (0.6 * t_celsius); }

/* TRMUtility.m:26-41 */
double oracle_amplitude(double dB)
{
    dB -= 60.0;
    if (dB <= -60.0) return 0.0;
    if (dB >= 0.0) return 1.0;
    return pow(10.0, dB / 20.0);
}

/* TRMUtility.m:44-47 */
double oracle_frequency(double pitch) { return 220.0 * pow(2.0, (pitch + 3.0) / 12.0); }

/* TRMUtility.m:50-66 */
double oracle_izero(double x)
{
    double sum = 1, u = 1, n = 1, halfx = x / 2.0;
    do {
        double temp = halfx / n;
        n += 1;
        temp *= temp;
        u *= temp;
        sum += u;
    } while (u >= (1E-21 * sum));
    return sum;
}

/* TRMUtility.m:71-85 : seed = frac(seed*377), sample = seed-0.5, seed0 = 0.7892347 */
static inline double noise_next(double *seed)
{
    double product = *seed * 377.0;
    *seed = product - (int)product;
    return *seed - 0.5;
}

double oracle_noise_draws(double seed, size_t n, double *out)
{
    for (size_t i = 0; i < n; i++) {
        double s = noise_next(&seed);
        if (out) out[i] = s;
    }
    return seed;
}

/* ------------------------------------------------------------------------------------------------
 * TRMFIRFilter.m
 * ---------------------------------------------------------------------------------------------- */
#define COEF_LIMIT 200

/* TRMFIRFilter.m:265-310 */
static void rational_approximation(double number, int32_t *order, int32_t *numerator, int32_t *denominator)
{
    if (*order <= 0) { *numerator = 0; *denominator = 0; *order = -1; return; }
    double frac = fabs(number - (int)number);
    int32_t order_max = 2 * (*order);
    if (order_max > COEF_LIMIT) order_max = COEF_LIMIT;
    int32_t modulus = 0;
    double min_err = 1.0;
    for (int32_t i = *order; i <= order_max; i++) {
        double ps = i * frac;
        int ip = (int)(ps + 0.5);
        double err = fabs((ps - (double)ip) / (double)i);
        if (err < min_err) { min_err = err; modulus = ip; *denominator = i; }
    }
    *numerator = (int)fabs(number) * (*denominator) + modulus;
    if (number < 0) *numerator *= -1;
    *order = *denominator - 1;
    if (*numerator == *denominator) {
        *denominator = order_max;
        *order = *numerator = *denominator - 1;
    }
}

/* TRMFIRFilter.m:161-233 ; coefficient[] is 1-based as in the reference */
static int maximally_flat(double beta, double gamma, int32_t *np, double *coefficient)
{
    double a[COEF_LIMIT + 1], c[COEF_LIMIT + 1];
    const double two_pi = 2.0 * M_PI;
    *np = 0;
    if (beta <= 0.0 || beta >= 0.5) return 1;
    double beta_min = ((2.0 * beta) < (1.0 - 2.0 * beta)) ? (2.0 * beta) : (1.0 - 2.0 * beta);
    if (gamma <= 0.0 || gamma >= beta_min) return 2;
    int32_t nt = (int32_t)(1.0 / (4.0 * gamma * gamma));
    if (nt > 160) return 3;
    double ac = (1.0 + cos(two_pi * beta)) / 2.0;
    int32_t numerator;
    rational_approximation(ac, &nt, &numerator, np);
    int32_t n = (2 * (*np)) - 1;
    if (numerator == 0) numerator = 1;
    c[1] = a[1] = 1.0;
    int32_t ll = nt - numerator;
    for (int32_t i = 2; i <= *np; i++) {
        double sum = 1.0;
        c[i] = cos(two_pi * ((double)(i - 1) / (double)n));
        double x = (1.0 - c[i]) / 2.0;
        double y = x;
        if (numerator == nt) continue;
        for (int32_t j = 1; j <= ll; j++) {
            double z = y;
            if (numerator != 1)
                for (int32_t jj = 1; jj <= (numerator - 1); jj++) z *= 1.0 + ((double)j / (double)jj);
            y *= x;
            sum += z;
        }
        a[i] = sum * pow(1.0 - x, numerator);
    }
    for (int32_t i = 1; i <= *np; i++) {
        coefficient[i] = a[1] / 2.0;
        for (int32_t j = 2; j <= *np; j++) {
            int m = ((i - 1) * (j - 1)) % n;
            if (m > nt) m = n - m;
            coefficient[i] += c[m + 1] * a[j];
        }
        coefficient[i] *= 2.0 / (double)n;
    }
    return 0;
}

/* TRMFIRFilter.m:37-98 (design + tap layout) and :236-244 (trim) */
int oracle_fir_design(double beta, double gamma, double cutoff, double *coef, int32_t *numberTaps)
{
    int32_t nc;
    double coefficient[COEF_LIMIT + 1];
    for (int i = 0; i <= COEF_LIMIT; i++) coefficient[i] = 0;
    if (maximally_flat(beta, gamma, &nc, coefficient) != 0) return -2;
    for (int32_t i = nc; i > 0; i--)
        if (fabs(coefficient[i]) >= fabs(cutoff)) { nc = i; break; }
    *numberTaps = (nc * 2) - 1;
    int32_t inc = -1, ptr = nc;
    for (int32_t i = 0; i < *numberTaps; i++) {
        coef[i] = coefficient[ptr];
        ptr += inc;
        if (ptr <= 0) { ptr = 2; inc = 1; }
    }
    return 0;
}

typedef struct {
    double data[2 * COEF_LIMIT + 1];
    double coef[2 * COEF_LIMIT + 1];
    int32_t ptr, taps;
} fir_t;

/* TRMFIRFilter.m:116-146 */
static inline double fir_filter(fir_t *f, double input, int need_output)
{
    if (need_output) {
        double output = 0.0;
        f->data[f->ptr] = input;
        for (int32_t i = 0; i < f->taps; i++) {
            output += f->data[f->ptr] * f->coef[i];
            if (++f->ptr >= f->taps) f->ptr = 0;
        }
        if (--f->ptr < 0) f->ptr = f->taps - 1;
        return output;
    }
    f->data[f->ptr] = input;
    if (--f->ptr < 0) f->ptr = f->taps - 1;
    return 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * TRMWavetable.m
 * ---------------------------------------------------------------------------------------------- */
#define TABLE_LENGTH 512
#define TABLE_MODULUS (TABLE_LENGTH - 1)

typedef struct {
    fir_t fir;
    double table[TABLE_LENGTH];
    int32_t div1, div2;
    double tnLength, tnDelta, basicIncrement, position;
    int waveform;
    /* analytic mode: current closure point / scale (pure function of the last update amplitude) */
    int analytic;
    double a_newDiv2, a_scale;
    double rise[TABLE_LENGTH];     /* init-time table, used by the analytic lookup outside [div1,div2) */
} wavetable_t;

/* TRMWavetable.m:28-34 : wraps only when value > 511 */
static inline double mod0(double value)
{
    if (value > TABLE_MODULUS) value -= TABLE_LENGTH;
    return value;
}

/* TRMWavetable.m:56-106 */
static int wavetable_init(wavetable_t *w, int waveform, double tp, double tnMin, double tnMax, double sampleRate,
                          int analytic)
{
    memset(w, 0, sizeof(*w));
    if (oracle_fir_design(.2, .1, .00000001, w->fir.coef, &w->fir.taps) != 0) return -2;   /* TRMFIRFilter.h:7-9 */
    w->fir.ptr = 0;
    w->div1 = rint(TABLE_LENGTH * (tp / 100.0));
    w->div2 = rint(TABLE_LENGTH * ((tp + tnMax) / 100.0));
    w->tnLength = w->div2 - w->div1;
    w->tnDelta = rint(TABLE_LENGTH * ((tnMax - tnMin) / 100.0));
    w->basicIncrement = (double)TABLE_LENGTH / sampleRate;
    w->position = 0;
    w->waveform = waveform;
    w->analytic = analytic;
    if (waveform == 0) {
        int32_t i, j;
        for (i = 0; i < w->div1; i++) {
            double x = (double)i / (double)w->div1;
            double x2 = x * x;
            double x3 = x2 * x;
            w->table[i] = (3.0 * x2) - (2.0 * x3);
        }
        for (i = w->div1, j = 0; i < w->div2; i++, j++) {
            double x = (double)j / w->tnLength;
            w->table[i] = 1.0 - (x * x);
        }
        for (i = w->div2; i < TABLE_LENGTH; i++) w->table[i] = 0.0;
    } else {
        for (int32_t i = 0; i < TABLE_LENGTH; i++)
            w->table[i] = sin(((double)i / (double)TABLE_LENGTH) * 2.0 * M_PI);
    }
    memcpy(w->rise, w->table, sizeof(w->table));
    w->a_newDiv2 = w->div2;
    w->a_scale = 1.0 / (w->tnLength * w->tnLength);
    return 0;
}

/* TRMWavetable.m:117-162, vDSP operation order: wt[div1+i] = 1 - (i*i)*(1/(L*L)) */
static inline void wavetable_update(wavetable_t *w, double amplitude)
{
    double newDiv2 = w->div2 - rint(amplitude * w->tnDelta);
    double newTnLength = newDiv2 - w->div1;
    double scale = 1.0 / (newTnLength * newTnLength);
    if (w->analytic) { w->a_newDiv2 = newDiv2; w->a_scale = scale; return; }
    int32_t len = newTnLength;
    double j = 0.0;
    for (int32_t i = 0; i < len; i++, j += 1.0) {
        double ajj = j * j;               /* vDSP_vsqD   */
        double aj = ajj * scale;          /* vDSP_vsmulD */
        w->table[w->div1 + i] = 1.0 - aj; /* vDSP_vsubD  */
    }
    for (int32_t i = newDiv2; i < w->div2; i++) w->table[i] = 0.0;
}

static inline double wavetable_at(const wavetable_t *w, int32_t i)
{
    if (!w->analytic || w->waveform != 0) return w->table[i];
    if (i < w->div1 || i >= w->div2) return w->rise[i];
    if ((double)i >= w->a_newDiv2) return 0.0;
    double j = (double)(i - w->div1);
    return 1.0 - ((j * j) * w->a_scale);
}

/* TRMWavetable.m:165-195 (2x oversampling oscillator) */
static inline double wavetable_oscillator(wavetable_t *w, double frequency)
{
    double output = 0.0;
    for (int index = 0; index < 2; index++) {
        w->position = mod0(w->position + ((frequency / 2.0) * w->basicIncrement));
        int32_t lower = w->position;
        int32_t upper = mod0(lower + 1);
        double lo = wavetable_at(w, lower);
        double interpolated = lo + ((w->position - lower) * (wavetable_at(w, upper) - lo));
        output = fir_filter(&w->fir, interpolated, index == 1);
    }
    return output;
}

/* ------------------------------------------------------------------------------------------------
 * TRMFilters.m
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double alpha, beta, gamma, xn1, xn2, yn1, yn2; } bandpass_t;
typedef struct { double a10, b11, a20, a21, b21, reflY, radX, radY; } radrefl_t;

/* TRMFilters.m:9-17 */
static inline void bandpass_coefficients(bandpass_t *f, int32_t sampleRate, double cf, double bw)
{
    double tanValue = tan((M_PI * bw) / sampleRate);
    double cosValue = cos((2.0 * M_PI * cf) / sampleRate);
    f->beta = (1.0 - tanValue) / (2.0 * (1.0 + tanValue));
    f->gamma = (0.5 + f->beta) * cosValue;
    f->alpha = (0.5 - f->beta) / 2.0;
}

/* TRMFilters.m:19-29 */
static inline double bandpass_filter(bandpass_t *f, double input)
{
    double output = 2.0 * ((f->alpha * (input - f->xn2)) + (f->gamma * f->yn1) - (f->beta * f->yn2));
    f->xn2 = f->xn1; f->xn1 = input; f->yn2 = f->yn1; f->yn1 = output;
    return output;
}

/* TRMFilters.m:34-45 */
static void radrefl_init(radrefl_t *f, double coeff)
{
    f->b11 = -coeff;
    f->a10 = 1.0 - fabs(f->b11);
    f->a20 = coeff;
    f->a21 = f->b21 = -(f->a20);
    f->reflY = f->radX = f->radY = 0;
}

/* TRMFilters.m:47-52 */
static inline double reflection_filter(radrefl_t *f, double input)
{
    double output = (f->a10 * input) - (f->b11 * f->reflY);
    f->reflY = output;
    return output;
}

/* TRMFilters.m:54-60 */
static inline double radiation_filter(radrefl_t *f, double input)
{
    double output = (f->a20 * input) + (f->a21 * f->radX) - (f->b21 * f->radY);
    f->radX = input;
    f->radY = output;
    return output;
}

/* ------------------------------------------------------------------------------------------------
 * TRMRingBuffer.m + TRMSampleRateConverter.m
 * ---------------------------------------------------------------------------------------------- */
#define ZERO_CROSSINGS 13
#define LP_CUTOFF (11.0 / 13.0)
#define L_RANGE 256
#define M_RANGE 256
#define M_BITS 8
#define FRACTION_BITS 16
#define FRACTION_RANGE 65536
#define FILTER_LENGTH (ZERO_CROSSINGS * L_RANGE)
#define FILTER_LIMIT (FILTER_LENGTH - 1)
#define N_MASK 0xFFFF0000u
#define L_MASK 0x0000FF00u
#define M_MASK 0x000000FFu
#define FRACTION_MASK 0x0000FFFFu
#define nValue(x) (((x) & N_MASK) >> FRACTION_BITS)
#define lValue(x) (((x) & L_MASK) >> M_BITS)
#define mValue(x) ((x) & M_MASK)
#define fractionValue(x) ((x) & FRACTION_MASK)
#define KAISER_BETA 5.658
#define RING_SIZE 1024

typedef struct {
    double sampleRateRatio;
    double h[FILTER_LENGTH], deltaH[FILTER_LENGTH];
    uint32_t timeRegisterIncrement, filterIncrement, phaseIncrement, timeRegister;
    double maximumSampleValue;
    int32_t numberSamples;
    /* output "memory stream" */
    double *out;
    size_t out_cap;
    int oom;
    /* ring buffer (TRMRingBuffer.m:27-44) */
    double buffer[RING_SIZE];
    int32_t padSize, fillSize, fillPtr, emptyPtr, fillCounter;
} src_t;

/* TRMSampleRateConverter.m:110-131 */
void oracle_src_filter(double *h, double *deltaH)
{
    h[0] = LP_CUTOFF;
    double x = M_PI / (double)L_RANGE;
    for (int index = 1; index < FILTER_LENGTH; index++) {
        double y = (double)index * x;
        h[index] = sin(y * LP_CUTOFF) / y;
    }
    double IBeta = 1.0 / oracle_izero(KAISER_BETA);
    for (int index = 0; index < FILTER_LENGTH; index++) {
        double temp = (double)index / FILTER_LENGTH;
        h[index] *= oracle_izero(KAISER_BETA * sqrt(1.0 - (temp * temp))) * IBeta;
    }
    for (int index = 0; index < FILTER_LIMIT; index++) deltaH[index] = h[index + 1] - h[index];
    deltaH[FILTER_LIMIT] = 0.0 - h[FILTER_LIMIT];
}

/* TRMSampleRateConverter.m:69-106 (without the ring/stream allocation) */
static void src_rates(src_t *s, double inputRate, double outputRate)
{
    s->sampleRateRatio = outputRate / inputRate;
    s->timeRegisterIncrement = (int)rint(pow(2.0, FRACTION_BITS) / s->sampleRateRatio);
    double roundedSampleRateRatio = pow(2.0, FRACTION_BITS) / (double)s->timeRegisterIncrement;
    if (s->sampleRateRatio >= 1.0) s->filterIncrement = L_RANGE;
    else s->phaseIncrement = (uint32_t)rint(s->sampleRateRatio * (double)FRACTION_RANGE);
    s->padSize = (s->sampleRateRatio >= 1.0) ? ZERO_CROSSINGS
                                             : (int32_t)((float)ZERO_CROSSINGS / roundedSampleRateRatio) + 1;
}

static int src_init(src_t *s, double inputRate, double outputRate)
{
    memset(s, 0, sizeof(*s));
    oracle_src_filter(s->h, s->deltaH);
    src_rates(s, inputRate, outputRate);
    /* TRMRingBuffer.m:27-44 */
    s->fillSize = RING_SIZE - (2 * s->padSize);
    s->fillPtr = s->padSize;
    s->emptyPtr = 0;
    s->fillCounter = 0;
    s->out_cap = 1 << 16;
    s->out = (double *)malloc(s->out_cap * sizeof(double));
    return s->out ? 0 : -3;
}

static inline void src_emit(src_t *s, double output)
{
    double a = fabs(output);
    if (a > s->maximumSampleValue) s->maximumSampleValue = a;
    if ((size_t)s->numberSamples >= s->out_cap) {
        size_t cap = s->out_cap * 2;
        double *p = (double *)realloc(s->out, cap * sizeof(double));
        if (!p) { s->oom = 1; return; }
        s->out = p;
        s->out_cap = cap;
    }
    s->out[s->numberSamples++] = output;
}

static inline void ring_inc(int32_t *i) { if (++(*i) >= RING_SIZE) (*i) -= RING_SIZE; }
static inline void ring_dec(int32_t *i) { if (--(*i) < 0) (*i) += RING_SIZE; }

/* TRMSampleRateConverter.m:155-298 */
static void src_process(src_t *s)
{
    int32_t endPtr = s->fillPtr - s->padSize;
    if (endPtr < 0) endPtr += RING_SIZE;
    if (endPtr < s->emptyPtr) endPtr += RING_SIZE;

    if (s->sampleRateRatio >= 1.0) {
        while (s->emptyPtr < endPtr) {
            double output = 0.0;
            double interpolation = (double)mValue(s->timeRegister) / (double)M_RANGE;
            int32_t index = s->emptyPtr;
            for (uint32_t fi = lValue(s->timeRegister); fi < FILTER_LENGTH; ring_dec(&index), fi += s->filterIncrement)
                output += s->buffer[index] * (s->h[fi] + s->deltaH[fi] * interpolation);
            s->timeRegister = ~s->timeRegister;
            interpolation = (double)mValue(s->timeRegister) / (double)M_RANGE;
            index = s->emptyPtr;
            ring_inc(&index);
            for (uint32_t fi = lValue(s->timeRegister); fi < FILTER_LENGTH; ring_inc(&index), fi += s->filterIncrement)
                output += s->buffer[index] * (s->h[fi] + s->deltaH[fi] * interpolation);
            src_emit(s, output);
            s->timeRegister = ~s->timeRegister;
            s->timeRegister += s->timeRegisterIncrement;
            s->emptyPtr += nValue(s->timeRegister);
            if (s->emptyPtr >= RING_SIZE) { s->emptyPtr -= RING_SIZE; endPtr -= RING_SIZE; }
            s->timeRegister &= (~N_MASK);
        }
    } else {
        while (s->emptyPtr < endPtr) {
            double output = 0.0;
            uint32_t phaseIndex = (uint32_t)rint(((double)fractionValue(s->timeRegister)) * s->sampleRateRatio);
            uint32_t impulseIndex;
            int32_t index = s->emptyPtr;
            while ((impulseIndex = (phaseIndex >> M_BITS)) < FILTER_LENGTH) {
                double impulse = s->h[impulseIndex] + (s->deltaH[impulseIndex] * (((double)mValue(phaseIndex)) / (double)M_RANGE));
                output += (s->buffer[index] * impulse);
                ring_dec(&index);
                phaseIndex += s->phaseIncrement;
            }
            phaseIndex = (unsigned int)rint(((double)fractionValue(~s->timeRegister)) * s->sampleRateRatio);
            index = s->emptyPtr;
            ring_inc(&index);
            while ((impulseIndex = (phaseIndex >> M_BITS)) < FILTER_LENGTH) {
                double impulse = s->h[impulseIndex] + (s->deltaH[impulseIndex] * (((double)mValue(phaseIndex)) / (double)M_RANGE));
                output += (s->buffer[index] * impulse);
                ring_inc(&index);
                phaseIndex += s->phaseIncrement;
            }
            src_emit(s, output);
            s->timeRegister += s->timeRegisterIncrement;
            s->emptyPtr += nValue(s->timeRegister);
            if (s->emptyPtr >= RING_SIZE) { s->emptyPtr -= RING_SIZE; endPtr -= RING_SIZE; }
            s->timeRegister &= (~N_MASK);
        }
    }
}

/* TRMRingBuffer.m:47-60 */
static inline void src_data_fill(src_t *s, double data)
{
    s->buffer[s->fillPtr] = data;
    if (++s->fillPtr >= RING_SIZE) s->fillPtr -= RING_SIZE;
    if (++s->fillCounter >= s->fillSize) {
        src_process(s);
        s->fillCounter = 0;
    }
}

/* TRMRingBuffer.m:85-93 */
static void src_flush(src_t *s)
{
    for (int32_t i = 0; i < (s->padSize * 2); i++) src_data_fill(s, 0.0);
    src_process(s);
}

/* Stateless closed form of the same converter (verified equal to the streaming form by
 * tests/test_oracle.py).  xb[p] = x[p-pad], zero outside [0,n_in). */
static int64_t src_stateless_count(const src_t *s, int64_t n_in)
{
    int64_t total = n_in + 2 * (int64_t)s->padSize;
    int64_t tri = s->timeRegisterIncrement;
    return (total * 65536 + tri - 1) / tri;
}

static void src_stateless(src_t *s, const double *x, int64_t n_in)
{
    int64_t n_out = src_stateless_count(s, n_in);
    int64_t pad = s->padSize;
    for (int64_t n = 0; n < n_out && !s->oom; n++) {
        uint64_t T = (uint64_t)n * s->timeRegisterIncrement;
        int64_t P = (int64_t)(T >> 16);
        uint32_t F = (uint32_t)(T & 0xFFFF);
        double output = 0.0;
#define XB(p) ((((p) - pad) >= 0 && ((p) - pad) < n_in) ? x[(p) - pad] : 0.0)
        if (s->sampleRateRatio >= 1.0) {
            double interpolation = (double)mValue(F) / (double)M_RANGE;
            int64_t idx = P;
            for (uint32_t fi = lValue(F); fi < FILTER_LENGTH; idx--, fi += L_RANGE)
                output += XB(idx) * (s->h[fi] + s->deltaH[fi] * interpolation);
            uint32_t G = ~F;
            interpolation = (double)mValue(G) / (double)M_RANGE;
            idx = P + 1;
            for (uint32_t fi = lValue(G); fi < FILTER_LENGTH; idx++, fi += L_RANGE)
                output += XB(idx) * (s->h[fi] + s->deltaH[fi] * interpolation);
        } else {
            uint32_t phaseIndex = (uint32_t)rint(((double)fractionValue(F)) * s->sampleRateRatio), impulseIndex;
            int64_t idx = P;
            while ((impulseIndex = (phaseIndex >> M_BITS)) < FILTER_LENGTH) {
                double impulse = s->h[impulseIndex] + (s->deltaH[impulseIndex] * (((double)mValue(phaseIndex)) / (double)M_RANGE));
                output += (XB(idx) * impulse);
                idx--;
                phaseIndex += s->phaseIncrement;
            }
            phaseIndex = (unsigned int)rint(((double)fractionValue(~F)) * s->sampleRateRatio);
            idx = P + 1;
            while ((impulseIndex = (phaseIndex >> M_BITS)) < FILTER_LENGTH) {
                double impulse = s->h[impulseIndex] + (s->deltaH[impulseIndex] * (((double)mValue(phaseIndex)) / (double)M_RANGE));
                output += (XB(idx) * impulse);
                idx++;
                phaseIndex += s->phaseIncrement;
            }
        }
#undef XB
        src_emit(s, output);
    }
}

/* ------------------------------------------------------------------------------------------------
 * TRMTubeModel.m
 * ---------------------------------------------------------------------------------------------- */
enum { R1, R2, R3, R4, R5, R6, R7, R8, TOTAL_REGIONS };
enum { S1, S2, S3, S4, S5, S6, S7, S8, S9, S10, TOTAL_SECTIONS };
enum { N1, N2, N3, N4, N5, N6, TOTAL_NASAL };
enum { C1, C2, C3, C4, C5, C6, C7, C8 };
enum { NC1, NC2, NC3, NC4, NC5, NC6 };
enum { FC1, FC2, FC3, FC4, FC5, FC6, FC7, FC8, TOTAL_FRIC };
enum { LEFT, RIGHT, UPPER };
enum { TOP, BOTTOM };
#define VT_SCALE 0.125

enum { P_PITCH, P_GLOTVOL, P_ASPVOL, P_FRICVOL, P_FRICPOS, P_FRICCF, P_FRICBW, P_RADIUS, P_VELUM = 15 };

typedef struct {
    const oracle_input_parameters *ip;
    int32_t controlPeriod, sampleRate;
    double actualTubeLength, dampingFactor, crossmixFactor, breathinessFactor;
    double noiseSeed, noiseFilterX;
    radrefl_t mouth, nasalPair;
    double ta0, tb1, throatY, throatGain;
    bandpass_t bp;
    double oropharynx[TOTAL_SECTIONS][2][2];
    double oropharynx_coeff[8];
    double nasal[TOTAL_NASAL][2][2];
    double nasal_coeff[TOTAL_NASAL];
    double alpha[3];
    unsigned cur, prev;
    double fricationTap[TOTAL_FRIC];
    double current[16], delta[16];
    wavetable_t wt;
    src_t src;
} tube_t;

/* TRMTubeModel.m:692-707 */
static void initialize_nasal_cavity(tube_t *t)
{
    const double *nr = t->ip->noseRadius;
    for (int index = N2, j = NC2; index < N6; index++, j++) {
        double radA2 = nr[index] * nr[index];
        double radB2 = nr[index + 1] * nr[index + 1];
        t->nasal_coeff[j] = (radA2 - radB2) / (radA2 + radB2);
    }
    double radA2 = nr[N6] * nr[N6];
    double radB2 = t->ip->apScale * t->ip->apScale;
    t->nasal_coeff[NC6] = (radA2 - radB2) / (radA2 + radB2);
}

/* TRMTubeModel.m:186-260 ; derive-only part at :196-203 */
static int tube_derive(const oracle_input_parameters *ip, int32_t *controlPeriod, int32_t *sampleRate, double *actualLen)
{
    if (!(ip->length > 0.0)) return -1;
    double c = speed_of_sound(ip->temperature);
    *controlPeriod = rint((c * TOTAL_SECTIONS * 100.0) / (ip->length * ip->controlRate));
    *sampleRate = ip->controlRate * *controlPeriod;
    *actualLen = (c * TOTAL_SECTIONS * 100.0) / *sampleRate;
    return 0;
}

static int tube_init(tube_t *t, const oracle_input_parameters *ip, int flags)
{
    memset(t, 0, sizeof(*t));
    t->ip = ip;
    int rc = tube_derive(ip, &t->controlPeriod, &t->sampleRate, &t->actualTubeLength);
    if (rc) return rc;
    double nyquist = (double)t->sampleRate / 2.0;
    t->breathinessFactor = ip->breathiness / 100.0;
    t->crossmixFactor = 1.0 / oracle_amplitude(ip->mixOffset);
    t->dampingFactor = (1.0 - (ip->lossFactor / 100.0));
    rc = wavetable_init(&t->wt, ip->waveform, ip->tp, ip->tnMin, ip->tnMax, t->sampleRate,
                        (flags & ORACLE_WAVETABLE_ANALYTIC) != 0);
    if (rc) return rc;
    radrefl_init(&t->mouth, (nyquist - ip->mouthCoef) / nyquist);
    radrefl_init(&t->nasalPair, (nyquist - ip->noseCoef) / nyquist);
    initialize_nasal_cavity(t);
    t->noiseSeed = 0.7892347;
    t->noiseFilterX = 0;
    /* TRMFilters.m:64-68 */
    t->ta0 = (ip->throatCutoff * 2.0) / t->sampleRate;
    t->tb1 = 1.0 - t->ta0;
    t->throatGain = oracle_amplitude(ip->throatVol);
    rc = src_init(&t->src, t->sampleRate, ip->outputRate);
    if (rc) return rc;
    t->cur = 1;
    t->prev = 0;
    return 0;
}

/* TRMTubeModel.m:712-744 */
static inline void calculate_tube_coefficients(tube_t *t)
{
    const double *radius = &t->current[P_RADIUS];
    double velum = t->current[P_VELUM];
    for (int index = 0; index < (TOTAL_REGIONS - 1); index++) {
        double radA2 = radius[index] * radius[index];
        double radB2 = radius[index + 1] * radius[index + 1];
        t->oropharynx_coeff[index] = (radA2 - radB2) / (radA2 + radB2);
    }
    {
        double radA2 = radius[R8] * radius[R8];
        double radB2 = t->ip->apScale * t->ip->apScale;
        t->oropharynx_coeff[C8] = (radA2 - radB2) / (radA2 + radB2);
    }
    double r0_2 = radius[R4] * radius[R4];
    double r1_2 = r0_2;
    double r2_2 = velum * velum;
    double sum = 2.0 / (r0_2 + r1_2 + r2_2);
    t->alpha[LEFT] = sum * r0_2;
    t->alpha[RIGHT] = sum * r1_2;
    t->alpha[UPPER] = sum * r2_2;
    {
        double radA2 = velum * velum;
        double radB2 = t->ip->noseRadius[N2] * t->ip->noseRadius[N2];
        t->nasal_coeff[NC1] = (radA2 - radB2) / (radA2 + radB2);
    }
}

/* TRMTubeModel.m:748-765 */
static inline void set_frication_taps(tube_t *t)
{
    double fricationAmplitude = oracle_amplitude(t->current[P_FRICVOL]);
    int32_t integerPart = (int32_t)t->current[P_FRICPOS];
    double complement = t->current[P_FRICPOS] - (double)integerPart;
    double remainder = 1.0 - complement;
    for (unsigned long index = FC1; index < TOTAL_FRIC; index++) {
        if (index == (unsigned long)(long)integerPart) {   /* NSUInteger vs int32_t compare: sign-extended then unsigned */
            t->fricationTap[index] = remainder * fricationAmplitude;
            if ((index + 1) < TOTAL_FRIC) t->fricationTap[++index] = complement * fricationAmplitude;
        } else
            t->fricationTap[index] = 0.0;
    }
}

/* TRMTubeModel.m:778-853 */
static inline double update_vocal_tract(tube_t *t, double input, double frication)
{
    t->cur = (t->cur + 1) % 2;
    t->prev = (t->prev + 1) % 2;
    const unsigned c = t->cur, p = t->prev;
    const double d = t->dampingFactor;
    double (*o)[2][2] = t->oropharynx;
    double (*n)[2][2] = t->nasal;
    const double *k = t->oropharynx_coeff, *nk = t->nasal_coeff, *tap = t->fricationTap;
    double delta;

    o[S1][TOP][c] = (o[S1][BOTTOM][p] * d) + input;

    delta = k[C1] * (o[S1][TOP][p] - o[S2][BOTTOM][p]);
    o[S2][TOP][c] = (o[S1][TOP][p] + delta) * d;
    o[S1][BOTTOM][c] = (o[S2][BOTTOM][p] + delta) * d;

    for (int i = S2, j = C2, f = FC1; i < S4; i++, j++, f++) {
        delta = k[j] * (o[i][TOP][p] - o[i + 1][BOTTOM][p]);
        o[i + 1][TOP][c] = ((o[i][TOP][p] + delta) * d) + (tap[f] * frication);
        o[i][BOTTOM][c] = ((o[i + 1][BOTTOM][p] + delta) * d);
    }

    double junctionPressure = (t->alpha[LEFT] * o[S4][TOP][p]) + (t->alpha[RIGHT] * o[S5][BOTTOM][p]) + (t->alpha[UPPER] * n[N1][BOTTOM][p]);
    o[S4][BOTTOM][c] = ((junctionPressure - o[S4][TOP][p]) * d);
    o[S5][TOP][c] = ((junctionPressure - o[S5][BOTTOM][p]) * d) + (tap[FC3] * frication);
    n[N1][TOP][c] = ((junctionPressure - n[N1][BOTTOM][p]) * d);

    delta = k[C4] * (o[S5][TOP][p] - o[S6][BOTTOM][p]);
    o[S6][TOP][c] = ((o[S5][TOP][p] + delta) * d) + (tap[FC4] * frication);
    o[S5][BOTTOM][c] = ((o[S6][BOTTOM][p] + delta) * d);

    o[S7][TOP][c] = (o[S6][TOP][p] * d) + (tap[FC5] * frication);
    o[S6][BOTTOM][c] = (o[S7][BOTTOM][p] * d);

    for (int i = S7, j = C5, f = FC6; i < S10; i++, j++, f++) {
        delta = k[j] * (o[i][TOP][p] - o[i + 1][BOTTOM][p]);
        o[i + 1][TOP][c] = ((o[i][TOP][p] + delta) * d) + (tap[f] * frication);
        o[i][BOTTOM][c] = ((o[i + 1][BOTTOM][p] + delta) * d);
    }

    o[S10][BOTTOM][c] = d * reflection_filter(&t->mouth, k[C8] * o[S10][TOP][p]);
    double output = radiation_filter(&t->mouth, (1.0 + k[C8]) * o[S10][TOP][p]);

    for (int i = N1, j = NC1; i < N6; i++, j++) {
        delta = nk[j] * (n[i][TOP][p] - n[i + 1][BOTTOM][p]);
        n[i + 1][TOP][c] = (n[i][TOP][p] + delta) * d;
        n[i][BOTTOM][c] = (n[i + 1][BOTTOM][p] + delta) * d;
    }

    n[N6][BOTTOM][c] = d * reflection_filter(&t->nasalPair, nk[NC6] * n[N6][TOP][p]);
    output += radiation_filter(&t->nasalPair, (1.0 + nk[NC6]) * n[N6][TOP][p]);
    return output;
}

/* TRMTubeModel.m:272-361 */
static void tube_synthesize(tube_t *t, const oracle_frame *frames, size_t n_frames, int flags, double *tube_out,
                            int64_t *n_tube)
{
    *n_tube = 0;
    if (n_frames == 0) return;
    const int stateless = (flags & ORACLE_SRC_STATELESS) != 0;
    double *xs = NULL;
    int64_t total = (int64_t)(n_frames - 1) * t->controlPeriod;
    if (stateless) xs = (double *)malloc((size_t)(total > 0 ? total : 1) * sizeof(double));

    for (size_t fi = 1; fi < n_frames; fi++) {
        const double *prev = frames[fi - 1].v, *next = frames[fi].v;
        /* TRMTubeModel.m:611-672 */
        for (int q = 0; q < 16; q++) {
            t->current[q] = prev[q];
            t->delta[q] = (next[q] - t->current[q]) / (double)t->controlPeriod;
        }
        for (int32_t j = 0; j < t->controlPeriod; j++) {
            double f0 = oracle_frequency(t->current[P_PITCH]);
            double ax = oracle_amplitude(t->current[P_GLOTVOL]);
            double ah1 = oracle_amplitude(t->current[P_ASPVOL]);
            calculate_tube_coefficients(t);
            set_frication_taps(t);
            bandpass_coefficients(&t->bp, t->sampleRate, t->current[P_FRICCF], t->current[P_FRICBW]);

            /* TRMFilters.m:81-86 on the noise draw */
            double nz = noise_next(&t->noiseSeed);
            double lp_noise = nz + t->noiseFilterX;
            t->noiseFilterX = nz;

            if (t->ip->waveform == 0) wavetable_update(&t->wt, ax);
            double pulse = wavetable_oscillator(&t->wt, f0);
            double pulsed_noise = lp_noise * pulse;
            pulse = ax * ((pulse * (1.0 - t->breathinessFactor)) + (pulsed_noise * t->breathinessFactor));
            double signal;
            if (t->ip->usesModulation) {
                double crossmix = ax * t->crossmixFactor;
                crossmix = (crossmix < 1.0) ? crossmix : 1.0;
                signal = (pulsed_noise * crossmix) + (lp_noise * (1.0 - crossmix));
            } else
                signal = lp_noise;

            signal = update_vocal_tract(t, ((pulse + (ah1 * signal)) * VT_SCALE), bandpass_filter(&t->bp, signal));

            /* TRMFilters.m:72-77 + TRMTubeModel.m:341 */
            t->throatY = (t->ta0 * (pulse * VT_SCALE)) + (t->tb1 * t->throatY);
            signal += t->throatY * t->throatGain;

            if (tube_out) tube_out[*n_tube] = signal;
            if (stateless) xs[*n_tube] = signal;
            else src_data_fill(&t->src, signal);
            (*n_tube)++;

            /* TRMTubeModel.m:676-688 */
            for (int q = 0; q < 16; q++) t->current[q] += t->delta[q];
        }
    }
    if (stateless) {
        src_stateless(&t->src, xs, *n_tube);
        free(xs);
    } else
        src_flush(&t->src);
}

int oracle_derive(const oracle_input_parameters *ip, size_t n_frames, oracle_result_info *info)
{
    memset(info, 0, sizeof(*info));
    int rc = tube_derive(ip, &info->controlPeriod, &info->sampleRate, &info->actualTubeLength);
    if (rc) return rc;
    src_t *s = (src_t *)calloc(1, sizeof(src_t));
    if (!s) return -3;
    src_rates(s, info->sampleRate, ip->outputRate);
    info->padSize = s->padSize;
    info->timeRegisterIncrement = s->timeRegisterIncrement;
    info->tubeSamples = n_frames ? (int64_t)(n_frames - 1) * info->controlPeriod : 0;
    info->numberSamples = n_frames ? (int32_t)src_stateless_count(s, info->tubeSamples) : 0;
    free(s);
    return 0;
}

int oracle_synthesize(const oracle_input_parameters *ip, const oracle_frame *frames, size_t n_frames, int flags,
                      double *tube_out, double **out, oracle_result_info *info)
{
    tube_t *t = (tube_t *)malloc(sizeof(tube_t));
    if (!t) return -3;
    int rc = tube_init(t, ip, flags);
    if (rc) { free(t->src.out); free(t); return rc; }
    int64_t n_tube = 0;
    tube_synthesize(t, frames, n_frames, flags, tube_out, &n_tube);
    if (t->src.oom) { free(t->src.out); free(t); return -3; }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->controlPeriod = t->controlPeriod;
        info->sampleRate = t->sampleRate;
        info->actualTubeLength = t->actualTubeLength;
        info->numberTaps = t->wt.fir.taps;
        info->padSize = t->src.padSize;
        info->timeRegisterIncrement = t->src.timeRegisterIncrement;
        info->numberSamples = t->src.numberSamples;
        info->maximumSampleValue = t->src.maximumSampleValue;
        info->finalNoiseSeed = t->noiseSeed;
        info->tubeSamples = n_tube;
    }
    if (out) *out = t->src.out;
    else free(t->src.out);
    free(t);
    return 0;
}

void oracle_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * Output stage: TRMTubeModel.m:365-490 (file) and 509-593 (WAV bytes)
 * ---------------------------------------------------------------------------------------------- */
void oracle_pcm16(const oracle_input_parameters *ip, const double *samples, int32_t numberSamples,
                  double maximumSampleValue, int file_variant, int16_t *dst)
{
    double scale = (32767.0 / maximumSampleValue) * oracle_amplitude(ip->volume);
    if (ip->channels == 2) {
        double leftScale = -((ip->balance / 2.0) - 0.5) * scale;
        double rightScale = ((ip->balance / 2.0) + 0.5) * scale;
        if (file_variant) { leftScale *= 2.0; rightScale *= 2.0; }   /* TRMTubeModel.m:382-383 vs 532-533 */
        for (int32_t i = 0; i < numberSamples; i++) {
            dst[2 * i] = (int16_t)rint(samples[i] * leftScale);
            dst[2 * i + 1] = (int16_t)rint(samples[i] * rightScale);
        }
    } else {
        for (int32_t i = 0; i < numberSamples; i++) dst[i] = (int16_t)rint(samples[i] * scale);
    }
}

static void put_le16(uint8_t **p, uint16_t v) { (*p)[0] = v & 0xFF; (*p)[1] = v >> 8; *p += 2; }
static void put_le32(uint8_t **p, uint32_t v) { for (int i = 0; i < 4; i++) (*p)[i] = (v >> (8 * i)) & 0xFF; *p += 4; }
static void put_be32(uint8_t **p, uint32_t v) { for (int i = 0; i < 4; i++) (*p)[i] = (v >> (8 * (3 - i))) & 0xFF; *p += 4; }

/* TRMTubeModel.m:562-590 : 18-byte fmt chunk, header 46 bytes */
uint8_t *oracle_wav_bytes(const oracle_input_parameters *ip, const double *samples, int32_t numberSamples,
                          double maximumSampleValue, size_t *len)
{
    int channels = ip->channels == 2 ? 2 : 1;
    size_t data_bytes = (size_t)numberSamples * channels * 2;
    uint8_t *buf = (uint8_t *)malloc(46 + data_bytes);
    if (!buf) return NULL;
    int frameSize = (int)ceil(ip->channels * ((double)16 / 8));
    int bytesPerSecond = (int)ceil(ip->outputRate * frameSize);
    uint32_t sub1 = 18, sub2 = (uint32_t)data_bytes;
    uint8_t *p = buf;
    put_be32(&p, 0x52494646);
    put_le32(&p, 4 + (8 + sub1) + (8 + sub2));
    put_be32(&p, 0x57415645);
    put_be32(&p, 0x666d7420);
    put_le32(&p, sub1);
    put_le16(&p, 1);
    put_le16(&p, (uint16_t)ip->channels);
    put_le32(&p, (uint32_t)ip->outputRate);
    put_le32(&p, (uint32_t)bytesPerSecond);
    put_le16(&p, (uint16_t)frameSize);
    put_le16(&p, 16);
    put_le16(&p, 0);
    put_be32(&p, 0x64617461);
    put_le32(&p, sub2);
    int16_t *pcm = (int16_t *)malloc(data_bytes ? data_bytes : 2);
    if (!pcm) { free(buf); return NULL; }
    oracle_pcm16(ip, samples, numberSamples, maximumSampleValue, 0, pcm);
    for (size_t i = 0; i < data_bytes / 2; i++) put_le16(&p, (uint16_t)pcm[i]);   /* host-endian append on LE hosts */
    free(pcm);
    *len = 46 + data_bytes;
    return buf;
}

/* ------------------------------------------------------------------------------------------------
 * TRMDataList.m:43-247
 * ---------------------------------------------------------------------------------------------- */
int oracle_parse_input_file(const char *path, oracle_input_parameters *ip, oracle_frame **frames, size_t *n_frames)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    char line[128];
    memset(ip, 0, sizeof(*ip));
#define NEXT() do { if (fgets(line, 128, fp) == NULL) { fclose(fp); return -2; } } while (0)
    NEXT(); ip->outputFileFormat = (int32_t)strtol(line, NULL, 10);
    NEXT(); ip->outputRate = strtod(line, NULL);
    NEXT(); ip->controlRate = strtod(line, NULL);
    NEXT(); ip->volume = strtod(line, NULL);
    NEXT(); ip->channels = (int32_t)strtol(line, NULL, 10);
    NEXT(); ip->balance = strtod(line, NULL);
    NEXT(); ip->waveform = (int32_t)strtol(line, NULL, 10);
    NEXT(); ip->tp = strtod(line, NULL);
    NEXT(); ip->tnMin = strtod(line, NULL);
    NEXT(); ip->tnMax = strtod(line, NULL);
    NEXT(); ip->breathiness = strtod(line, NULL);
    NEXT(); ip->length = strtod(line, NULL);
    NEXT(); ip->temperature = strtod(line, NULL);
    NEXT(); ip->lossFactor = strtod(line, NULL);
    NEXT(); ip->apScale = strtod(line, NULL);
    NEXT(); ip->mouthCoef = strtod(line, NULL);
    NEXT(); ip->noseCoef = strtod(line, NULL);
    for (int i = 1; i < TOTAL_NASAL; i++) { NEXT(); ip->noseRadius[i] = strtod(line, NULL); }
    NEXT(); ip->throatCutoff = strtod(line, NULL);
    NEXT(); ip->throatVol = strtod(line, NULL);
    NEXT(); ip->usesModulation = (strtol(line, NULL, 10) != 0);
    NEXT(); ip->mixOffset = strtod(line, NULL);
#undef NEXT
    size_t cap = 256, n = 0;
    oracle_frame *fr = (oracle_frame *)malloc(cap * sizeof(oracle_frame));
    if (!fr) { fclose(fp); return -3; }
    while (fgets(line, 128, fp)) {
        if (n + 2 > cap) {
            cap *= 2;
            oracle_frame *q = (oracle_frame *)realloc(fr, cap * sizeof(oracle_frame));
            if (!q) { free(fr); fclose(fp); return -3; }
            fr = q;
        }
        char *ptr = line;
        for (int q = 0; q < 16; q++) fr[n].v[q] = strtod(ptr, &ptr);
        n++;
    }
    if (n > 0) { fr[n] = fr[n - 1]; n++; }   /* TRMDataList.m:239-241 doubles the last frame */
    fclose(fp);
    *frames = fr;
    *n_frames = n;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline: one utterance per thread (dynamic schedule over an atomic counter).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const oracle_input_parameters *ip;
    int shared_ip;
    const oracle_frame *frames;
    const int64_t *frame_offset;
    const int32_t *n_frames;
    int n, flags;
    int *next;
    pthread_mutex_t *mu;
    int32_t *numberSamples;
    double *maximumSampleValue, *checksum;
    int rc;
} batch_ctx;

static void *batch_worker(void *arg)
{
    batch_ctx *c = (batch_ctx *)arg;
    for (;;) {
        pthread_mutex_lock(c->mu);
        int u = (*c->next)++;
        pthread_mutex_unlock(c->mu);
        if (u >= c->n) break;
        double *out = NULL;
        oracle_result_info info;
        int rc = oracle_synthesize(c->shared_ip ? c->ip : c->ip + u, c->frames + c->frame_offset[u],
                                   (size_t)c->n_frames[u], c->flags, NULL, &out, &info);
        if (rc) { c->rc = rc; continue; }
        double sum = 0.0;
        for (int32_t i = 0; i < info.numberSamples; i++) sum += out[i];
        if (c->numberSamples) c->numberSamples[u] = info.numberSamples;
        if (c->maximumSampleValue) c->maximumSampleValue[u] = info.maximumSampleValue;
        if (c->checksum) c->checksum[u] = sum;
        free(out);
    }
    return NULL;
}

int oracle_synthesize_batch(const oracle_input_parameters *ip, int shared_ip, const oracle_frame *frames,
                            const int64_t *frame_offset, const int32_t *n_frames, int n_utterances, int flags,
                            int n_threads, int32_t *numberSamples, double *maximumSampleValue, double *checksum)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t th[1024];
    batch_ctx ctx[1024];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = 0;
    for (int i = 0; i < n_threads; i++) {
        ctx[i] = (batch_ctx){ip, shared_ip, frames, frame_offset, n_frames, n_utterances, flags, &next, &mu,
                             numberSamples, maximumSampleValue, checksum, 0};
        if (pthread_create(&th[i], NULL, batch_worker, &ctx[i]) != 0) {
            n_threads = i;
            break;
        }
    }
    int rc = 0;
    for (int i = 0; i < n_threads; i++) {
        pthread_join(th[i], NULL);
        if (ctx[i].rc) rc = ctx[i].rc;
    }
    return n_threads > 0 ? rc : -3;
}

/* ------------------------------------------------------------------------------------------------
 * Control-frame generator (EventList.m:883-1061, MMDriftGenerator.m:41-78).  Operation by operation in the
 * reference's order and types: currentValues / currentDeltas are double, the output table and the drift generator
 * are float.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float pitchDeviation, pitchOffset, a0, b1, seed, previousSample; } drift_t;

/* MMDriftGenerator.m:41-58 */
static void drift_configure(drift_t *g, float deviation, float sampleRate, float lowpassCutoff)
{
    g->pitchDeviation = deviation * 2.0;
    g->pitchOffset = deviation;
    if (lowpassCutoff < 0.0) lowpassCutoff = 0.0;
    else if (lowpassCutoff > (sampleRate / 2.0)) lowpassCutoff = sampleRate / 2.0;
    g->a0 = (lowpassCutoff * 2.0) / sampleRate;
    g->b1 = 1.0 - g->a0;
    g->previousSample = 0.0;
}

/* MMDriftGenerator.m:65-78 */
static float drift_generate(drift_t *g)
{
    float temp = g->seed * 377.0f;
    g->seed = temp - (int32_t)temp;
    temp = (g->seed * g->pitchDeviation) - g->pitchOffset;
    g->previousSample = (g->a0 * temp) + (g->b1 * g->previousSample);
    return g->previousSample;
}

static int64_t generate_frames(const oracle_event *ev, int64_t count, const oracle_framegen *fg, oracle_frame *out,
                               int64_t max_frames, float *seed_out)
{
    if (seed_out && fg) *seed_out = fg->driftSeed;
    if (count == 0) return 0;                                       /* m:890-891 */
    drift_t drift = {0, 0, 0, 0, fg ? fg->driftSeed : 0.7892347f, 0};
    const double millisecondsPerInterval = 1000.0 / 250.0;
    if (fg && fg->useDrift) drift_configure(&drift, (float)fg->driftDeviation, (float)(1000 / 4), (float)fg->driftCutoff);   /* m:903-907 */

    double currentValues[36], currentDeltas[36], temp;
    int64_t emitted = 0;
    if (count < 2) return 0;     /* the reference indexes _mutableEvents[1] unconditionally: a one-event list is undefined */
    for (int i = 0; i < 16; i++) {                                  /* m:919-926 */
        int64_t j = 1;
        while (isnan(temp = ev[j].value[i])) j++;
        currentValues[i] = ev[0].value[i];
        currentDeltas[i] = ((temp - currentValues[i]) / (double)(ev[j].time)) * millisecondsPerInterval;
    }
    for (int i = 16; i < 36; i++) currentValues[i] = currentDeltas[i] = 0.0;   /* m:929-930 */
    if (fg && fg->useSmoothIntonation) {                            /* m:932-943 */
        int64_t j = 0;
        while (isnan(temp = ev[j].value[32])) {
            j++;
            if (j >= count) break;
        }
        currentValues[32] = j < count ? ev[j].value[32] : NAN;      /* the reference reads out of bounds here */
        currentDeltas[32] = 0.0;
    } else if (fg) {                                                /* m:944-961 */
        int64_t j = 1;
        while (isnan(temp = ev[j].value[32])) {
            j++;
            if (j >= count) break;
        }
        currentValues[32] = ev[0].value[32];
        if (j < count) currentDeltas[32] = ((temp - currentValues[32]) / (double)(ev[j].time)) * millisecondsPerInterval;
        else currentDeltas[32] = 0;
        currentValues[32] = -20.0;
    }

    int64_t i = 1;
    uint64_t currentTime_ms = 0;
    uint64_t nextTime = (uint64_t)ev[1].time;
    float table[16];
    while (i < count) {                                             /* m:973 */
        if (fg) {
            for (int j = 0; j < 16; j++) table[j] = (float)currentValues[j] + (float)currentValues[j + 16];
            if (!fg->useMicroIntonation) table[0] = 0.0;
            if (fg->useDrift) table[0] += drift_generate(&drift);
            if (fg->useMacroIntonation) table[0] += currentValues[32];
            table[0] += fg->pitch;
            if (out && emitted < max_frames)
                for (int j = 0; j < 16; j++) out[emitted].v[j] = table[j];
        }
        emitted++;
        if (fg) {
            for (int j = 0; j < 32; j++)
                if (currentDeltas[j]) currentValues[j] += currentDeltas[j];
            if (fg->useSmoothIntonation) {
                currentDeltas[34] += currentDeltas[35];
                currentDeltas[33] += currentDeltas[34];
                currentValues[32] += currentDeltas[33];
            } else {
                if (currentDeltas[32]) currentValues[32] += currentDeltas[32];
            }
        }
        currentTime_ms += millisecondsPerInterval;

        if (currentTime_ms >= nextTime) {                           /* m:1025 */
            i++;
            if (i == count) break;
            nextTime = (uint64_t)ev[i].time;
            if (!fg) continue;
            for (int j = 0; j < 33; j++) {
                if (!isnan(ev[i - 1].value[j])) {
                    int64_t k = i;
                    while (isnan(temp = ev[k].value[j])) {
                        if (k >= count - 1) {
                            currentDeltas[j] = 0.0;
                            break;
                        }
                        k++;
                    }
                    if (!isnan(temp))
                        currentDeltas[j] = (temp - currentValues[j]) / (double)((uint64_t)ev[k].time - currentTime_ms) * millisecondsPerInterval;
                }
            }
            if (fg->useSmoothIntonation) {
                if (!isnan(ev[i - 1].value[33])) {
                    currentValues[32] = ev[i - 1].value[32];
                    currentDeltas[32] = 0.0;
                    currentDeltas[33] = ev[i - 1].value[33];
                    currentDeltas[34] = ev[i - 1].value[34];
                    currentDeltas[35] = ev[i - 1].value[35];
                }
            }
        }
    }
    if (seed_out) *seed_out = drift.seed;
    return emitted;
}

int64_t oracle_frame_count(const oracle_event *events, int64_t n_events) { return generate_frames(events, n_events, NULL, NULL, 0, NULL); }

int64_t oracle_generate_frames(const oracle_event *events, int64_t n_events, const oracle_framegen *fg, oracle_frame *out,
                               int64_t max_frames, float *seed_out)
{
    return generate_frames(events, n_events, fg, out, max_frames, seed_out);
}

/* see trm_oracle.h: exactness check of the kernels' known-divisor division */
int64_t oracle_div_known_mismatches(double c, double lo, double hi, int64_t n, uint64_t seed)
{
    const volatile double rc = 1.0 / c;
    int64_t bad = 0;
    uint64_t s = seed * 2862933555777941757ULL + 3037000493ULL;
    for (int64_t i = 0; i < n; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
        const volatile double a = lo + (hi - lo) * u;
        const volatile double q = a * rc;
        const double r = fma(-c, q, a);
        const double q2 = fma(r, rc, q);
        const volatile double ref = a / c;
        if (q2 != ref) ++bad;
    }
    return bad;
}
