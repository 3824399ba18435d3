/*
 * trm_workload.h -- synthetic control-frame generators for the benchmark configurations of
 * BASELINE.json.  Deterministic, counter-based (SplitMix64 keyed by the
 * utterance index), so CPU oracle, GPU path and every rank see identical tracks without exchanging data.
 * Values are rounded to float and widened, as Monet produces them (EventList.m:968-1002).
 */
#ifndef TRM_WORKLOAD_H
#define TRM_WORKLOAD_H

#include <stddef.h>
#include <stdint.h>

#include "trm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Postures of diphones.mxml used by config 1: 0 = "a" (:196-211), 1 = "aa" (:260-275). */
void TRMWorkloadStaticVowel(int posture, double pitch, size_t n_frames, TRMParameters *out);

/* Config 2/4/5: bounded reflecting random walk, one step per frame, sigma = range/50, utterance `index`
 * of stream `seed`.  Ranges: pitch -12+-10, glotVol 0..60, aspVol 0..10, fricVol 0..24, fricPos 0..7,
 * fricCF 864..5500, fricBW 500..4500, r1 = 0.8, r2..r8 0.05..2.61, velum 0.1..1.5. */
void TRMWorkloadRandomWalk(uint64_t seed, uint64_t index, size_t n_frames, TRMParameters *out);

/* Config 5 (10^6 utterances: the tracks cannot cross PCIe, they are generated where they are consumed): the same bounded
 * reflecting walk with ranges as above, defined so that the host and a GPU kernel produce IDENTICAL bits -- one SplitMix64
 * stream per (seed, utterance index, parameter), steps = (u1 + u2 + u3 + u4 - 2) * sqrt(3) * range / 50 (Irwin-Hall, unit
 * variance) instead of Box-Muller: only IEEE additions and multiplications, no libm.  The device twin is
 * workload_walk2_kernel (kernels_aux.cu, no FMA contraction); tests/test_gpu_sweep.py compares them bit for bit. */
void TRMWorkloadWalk2(uint64_t seed, uint64_t index, size_t n_frames, TRMParameters *out);

/* n utterances of n_frames frames each, written back to back; uses n_threads host threads. */
void TRMWorkloadRandomWalkBatch(uint64_t seed, uint64_t first_index, size_t n, size_t n_frames, TRMParameters *out,
                                int n_threads);

/* Config 3: static grid point `index` in [0, 65536): r2..r8 in {0.4,0.9,1.4,1.9}^7 x velum {0.1,0.8} x
 * pitch {-12,-5}; other parameters as config 1. */
void TRMWorkloadGridPoint(uint64_t index, size_t n_frames, TRMParameters *out);

#ifdef __cplusplus
}
#endif
#endif
