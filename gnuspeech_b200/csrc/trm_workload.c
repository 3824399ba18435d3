/* trm_workload.c -- synthetic control-frame generators (include/trm_workload.h). */
#include "trm_workload.h"

#include <math.h>
#include <pthread.h>
#include <string.h>

static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uniform01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }
/* standard normal, Box-Muller */
static double normal(uint64_t *s)
{
    double u1 = uniform01(s), u2 = uniform01(s);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static double as_float(double v) { return (double)(float)v; }

static const double k_posture[2][9] = {
    {0.8, 0.65, 0.65, 0.65, 1.31, 1.23, 1.31, 1.67, 0.1},    /* "a"  */
    {0.8, 0.65, 0.84, 1.15, 1.31, 1.59, 1.59, 2.61, 0.1},    /* "aa" */
};

static void base_frame(TRMParameters *f, int posture, double pitch)
{
    f->glottalPitch = pitch; f->glottalVolume = 60; f->aspirationVolume = 0; f->fricationVolume = 0;
    f->fricationPosition = 5.5; f->fricationCenterFrequency = 2500; f->fricationBandwidth = 500;
    for (int i = 0; i < 8; i++) f->radius[i] = k_posture[posture][i];
    f->velum = k_posture[posture][8];
}

void TRMWorkloadStaticVowel(int posture, double pitch, size_t n_frames, TRMParameters *out)
{
    TRMParameters f;
    base_frame(&f, posture ? 1 : 0, pitch);
    for (size_t i = 0; i < n_frames; i++) out[i] = f;
}

static const double k_lo[16] = {-22, 0, 0, 0, 0, 864, 500, 0.8, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.1};
static const double k_hi[16] = {-2, 60, 10, 24, 7, 5500, 4500, 0.8, 2.61, 2.61, 2.61, 2.61, 2.61, 2.61, 2.61, 1.5};

void TRMWorkloadRandomWalk(uint64_t seed, uint64_t index, size_t n_frames, TRMParameters *out)
{
    uint64_t s = seed * 0xD1342543DE82EF95ull + index * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    splitmix64(&s);
    double x[16];
    for (int q = 0; q < 16; q++) x[q] = k_lo[q] + (k_hi[q] - k_lo[q]) * uniform01(&s);
    for (size_t i = 0; i < n_frames; i++) {
        double *v = (double *)&out[i];
        for (int q = 0; q < 16; q++) {
            const double range = k_hi[q] - k_lo[q];
            if (i > 0 && range > 0) {
                x[q] += normal(&s) * (range / 50.0);
                /* reflect into [lo, hi] */
                for (int it = 0; it < 4 && (x[q] < k_lo[q] || x[q] > k_hi[q]); it++) {
                    if (x[q] < k_lo[q]) x[q] = 2 * k_lo[q] - x[q];
                    if (x[q] > k_hi[q]) x[q] = 2 * k_hi[q] - x[q];
                }
            }
            v[q] = as_float(x[q]);
        }
    }
}

/* walk2: see include/trm_workload.h.  Every operation below is one IEEE double operation in a fixed order (this file is
 * compiled with -ffp-contract=off); workload_walk2_kernel performs the same sequence with __dadd_rn / __dmul_rn. */
void TRMWorkloadWalk2(uint64_t seed, uint64_t index, size_t n_frames, TRMParameters *out)
{
    for (int q = 0; q < 16; q++) {
        uint64_t s = seed * 0xD1342543DE82EF95ull + index * 0x9E3779B97F4A7C15ull + (uint64_t)q * 0xC2B2AE3D27D4EB4Full + 0x632BE59BD9B4E019ull;
        splitmix64(&s);
        const double lo = k_lo[q], hi = k_hi[q], range = hi - lo;
        const double step = (range * 1.7320508075688772) * 0.02;
        double x = lo + range * uniform01(&s);
        for (size_t i = 0; i < n_frames; i++) {
            if (i > 0 && range > 0) {
                double g = uniform01(&s);
                g = g + uniform01(&s);
                g = g + uniform01(&s);
                g = g + uniform01(&s);
                g = g - 2.0;
                x = x + g * step;
                for (int it = 0; it < 4 && (x < lo || x > hi); it++) {
                    if (x < lo) x = 2 * lo - x;
                    if (x > hi) x = 2 * hi - x;
                }
            }
            ((double *)&out[i])[q] = as_float(x);
        }
    }
}

typedef struct { uint64_t seed, first; size_t n, n_frames, t, nt; TRMParameters *out; } walk_job;
static void *walk_main(void *arg)
{
    walk_job *j = arg;
    for (size_t u = j->t; u < j->n; u += j->nt)
        TRMWorkloadRandomWalk(j->seed, j->first + u, j->n_frames, j->out + u * j->n_frames);
    return NULL;
}

void TRMWorkloadRandomWalkBatch(uint64_t seed, uint64_t first_index, size_t n, size_t n_frames, TRMParameters *out,
                                int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    walk_job jobs[256];
    int started = 0;
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (walk_job){seed, first_index, n, n_frames, (size_t)t, (size_t)n_threads, out};
        if (t == n_threads - 1 || pthread_create(&th[t], NULL, walk_main, &jobs[t]) != 0) {
            /* run the remaining stripes inline */
            for (int r = t; r < n_threads; r++) {
                jobs[r] = (walk_job){seed, first_index, n, n_frames, (size_t)r, (size_t)n_threads, out};
                walk_main(&jobs[r]);
            }
            break;
        }
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

void TRMWorkloadGridPoint(uint64_t index, size_t n_frames, TRMParameters *out)
{
    static const double level[4] = {0.4, 0.9, 1.4, 1.9};
    TRMParameters f;
    base_frame(&f, 0, (index >> 15) & 1 ? -5.0 : -12.0);
    f.velum = (index >> 14) & 1 ? 0.8 : 0.1;
    for (int i = 0; i < 7; i++) f.radius[1 + i] = level[(index >> (2 * i)) & 3];
    for (int q = 0; q < 16; q++) ((double *)&f)[q] = as_float(((double *)&f)[q]);
    for (size_t i = 0; i < n_frames; i++) out[i] = f;
}
