// src_kernel.cuh -- batch-wide sample-rate converter and PCM scaling kernels for sm_100a.
//
// src_kernel replaces -[TRMSampleRateConverter processDataFromRingBuffer:]
// (/root/reference/Frameworks/Tube/TRMSampleRateConverter.m:155-298) together with the ring buffer that
// feeds it (TRMRingBuffer.m:27-105).  The streaming converter is equivalent to a stateless gather
// (checked on the CPU by tests/test_oracle.py): with xb[p] = x[p - pad]
// (zero outside [0, n_in)), output n has time register T = n*TRI, P = T>>16, F = T&0xFFFF and
//   up-sampling  : y = sum_{k=0..12} xb[P-k]  *(h[l +256k] + dH[l +256k]*m /256)        (l ,m ) = (F>>8, F&255)
//                    + sum_{k=0..12} xb[P+1+k]*(h[l'+256k] + dH[l'+256k]*m'/256)        (l',m') from (~F)&0xFFFF
//   down-sampling: phase walks of TRMSampleRateConverter.m:246-270.
// Accumulation order (left wing first, from 0.0) is the reference's.  Also produces the per-utterance
// maximumSampleValue (m:206-208) with an order-independent integer atomicMax on the bit pattern.
//
// pcm_kernel replaces the scaling loops of -generateWAVData (TRMTubeModel.m:515-559):
//   scale = (32767 / max) * amplitude(volume);  int16 = rint(sample * scale)  (mono)
//   stereo: left = rint(sample * leftGain*scale), right = rint(sample * rightGain*scale), interleaved.
//
// Persistent CTAs: the 3328-entry (h, deltaH) table is staged once per CTA in shared memory; input
// windows are staged per tile with coalesced loads; outputs are written coalesced.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kernel_args.h"
#include "trm_cuda.h"

namespace trm {

template <typename R> __device__ __forceinline__ R r_abs(R x);
template <> __device__ __forceinline__ double r_abs<double>(double x) { return fabs(x); }
template <> __device__ __forceinline__ float r_abs<float>(float x) { return fabsf(x); }

// One work item = one tile of 32 utterances with the same converter signature (time-register increment, pad,
// direction, phase increment) x one run of `nt` consecutive output samples.  Lane = utterance: every lane of a
// warp computes the SAME output index n, so the time register, the filter phase and all 26 interpolated
// coefficients are warp-uniform (one broadcast table read per tap), and the input window is staged transposed
// ([input row][utterance], padded) so each tap is one conflict-free shared-memory read.  Outputs are transposed
// back through a small per-warp tile so global stores are coalesced 128-byte rows.
template <typename R>
__global__ void __launch_bounds__(SRC_THREADS) src_kernel(SrcArgs args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HD<R> *tab = reinterpret_cast<HD<R> *>(smem_raw);
    R *xT = reinterpret_cast<R *>(tab + TRM_SRC_FILTER_LEN);             // [SRC_ROWS][SRC_LD]
    R *yT = xT + SRC_ROWS * SRC_LD;                                      // [warps][SRC_CHUNK][SRC_LD]
    __shared__ long long s_tube_off[32], s_out_off[32], s_n_in[32], s_n_out[32];
    __shared__ int s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const HD<R> *g = reinterpret_cast<const HD<R> *>(args.table);
        for (int i = threadIdx.x; i < TRM_SRC_FILTER_LEN; i += SRC_THREADS) tab[i] = g[i];
    }

    for (long long item = blockIdx.x; item < args.total_items; item += gridDim.x) {
        __syncthreads();                       // previous item's window / descriptors no longer read
        if (threadIdx.x == 0) {
            int lo = 0, hi = args.n_tiles;     // largest tile with item_base[tile] <= item
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (args.item_base[mid] <= item) lo = mid; else hi = mid;
            }
            s_tile = lo;
        }
        __syncthreads();
        const int tile = s_tile;
        if (threadIdx.x < 32) {
            const int u = args.tile_utt[tile * 32 + threadIdx.x];
            const trm_cuda_utterance *D = args.desc + (u >= 0 ? u : 0);
            s_tube_off[threadIdx.x] = D->tube_offset;
            s_out_off[threadIdx.x] = D->out_offset;
            s_n_in[threadIdx.x] = u >= 0 ? D->n_tube : -1;
            s_n_out[threadIdx.x] = u >= 0 ? D->n_out : 0;
        }
        // signature of the tile (row 0 is always a real utterance)
        const trm_cuda_utterance *__restrict__ D0 = args.desc + args.tile_utt[tile * 32];
        const unsigned long long tri = D0->tri;
        const int pad = D0->padSize, reach = pad + 1;
        const bool up = D0->upsample != 0;
        const double ratio = D0->sampleRateRatio;
        const unsigned phaseIncrement = D0->phaseIncrement;
        const int nt = args.tile_nt[tile];
        const long long n_s = (item - args.item_base[tile]) * nt;
        const long long tile_max = args.tile_max_out[tile];
        const long long n_e = (n_s + nt < tile_max) ? n_s + nt : tile_max;
        const long long win_lo = (long long)(((unsigned long long)n_s * tri) >> 16) - reach;
        const int rows = (int)((long long)(((unsigned long long)(n_e - 1) * tri) >> 16) + reach + 2 - win_lo);
        __syncthreads();

        // stage the input window transposed: xT[i][r] = xb_r[win_lo + i], xb[p] = x[p - pad] (0 outside)
        for (int r = warp; r < 32; r += SRC_THREADS / 32) {
            const long long n_in = s_n_in[r];
            const R *__restrict__ x = reinterpret_cast<const R *>(args.tube) + s_tube_off[r];
            for (int i = lane; i < rows; i += 32) {
                const long long q = win_lo + i - pad;
                xT[i * SRC_LD + r] = (q >= 0 && q < n_in) ? x[q] : (R)0;
            }
        }
        __syncthreads();

        const long long my_n_out = s_n_out[lane];
        R local_max = (R)0;
        R *yw = yT + warp * (SRC_CHUNK * SRC_LD);
        for (long long n0 = n_s + (long long)warp * SRC_CHUNK; n0 < n_e; n0 += (SRC_THREADS / 32) * SRC_CHUNK) {
            if (up) {
                // SRC_ILP outputs at a time: their 26-tap accumulation chains are independent, which is what hides
                // the FP latency with only 8-16 warps per SM (the taps themselves stay in the reference's order)
#pragma unroll 1
                for (int j = 0; j < SRC_CHUNK; j += SRC_ILP) {
                    const R *xp[SRC_ILP];
                    unsigned F[SRC_ILP];
                    R acc[SRC_ILP];
#pragma unroll
                    for (int o = 0; o < SRC_ILP; ++o) {
                        const unsigned long long T = (unsigned long long)(n0 + j + o) * tri;
                        xp[o] = xT + (int)((long long)(T >> 16) - win_lo) * SRC_LD + lane;
                        F[o] = (unsigned)(T & 0xFFFFull);
                        acc[o] = (R)0;
                    }
                    {
                        R interp[SRC_ILP];
                        unsigned fi[SRC_ILP];
#pragma unroll
                        for (int o = 0; o < SRC_ILP; ++o) { interp[o] = (R)(F[o] & 255u) / (R)256; fi[o] = F[o] >> 8; }
#pragma unroll
                        for (int k = 0; k < SRC_ZC; ++k) {
#pragma unroll
                            for (int o = 0; o < SRC_ILP; ++o) {
                                const HD<R> c = tab[fi[o] + 256u * k];
                                acc[o] += xp[o][-k * SRC_LD] * (c.h + c.dh * interp[o]);
                            }
                        }
                    }
                    {
                        R interp[SRC_ILP];
                        unsigned fi[SRC_ILP];
#pragma unroll
                        for (int o = 0; o < SRC_ILP; ++o) {
                            const unsigned G = (~F[o]) & 0xFFFFu;
                            interp[o] = (R)(G & 255u) / (R)256;
                            fi[o] = G >> 8;
                        }
#pragma unroll
                        for (int k = 0; k < SRC_ZC; ++k) {
#pragma unroll
                            for (int o = 0; o < SRC_ILP; ++o) {
                                const HD<R> c = tab[fi[o] + 256u * k];
                                acc[o] += xp[o][(1 + k) * SRC_LD] * (c.h + c.dh * interp[o]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < SRC_ILP; ++o) {
                        yw[(j + o) * SRC_LD + lane] = acc[o];
                        const R av = r_abs<R>(acc[o]);
                        if (n0 + j + o < my_n_out && av > local_max) local_max = av;   // NaN never wins, like the reference
                    }
                }
            } else {
                for (int j = 0; j < SRC_CHUNK; ++j) {
                    const unsigned long long T = (unsigned long long)(n0 + j) * tri;
                    const int base = (int)((long long)(T >> 16) - win_lo);
                    const unsigned F = (unsigned)(T & 0xFFFFull);
                    const R *xp = xT + base * SRC_LD + lane;
                    R acc = (R)0;
                    unsigned ph = (unsigned)rint((double)F * ratio), ii;
                    const R *xq = xp;
                    while ((ii = (ph >> 8)) < (unsigned)TRM_SRC_FILTER_LEN) {
                        const HD<R> c = tab[ii];
                        const R impulse = c.h + (c.dh * ((R)(ph & 255u) / (R)256));
                        acc += (*xq * impulse);
                        xq -= SRC_LD;
                        ph += phaseIncrement;
                    }
                    ph = (unsigned)rint((double)((~F) & 0xFFFFu) * ratio);
                    xq = xp + SRC_LD;
                    while ((ii = (ph >> 8)) < (unsigned)TRM_SRC_FILTER_LEN) {
                        const HD<R> c = tab[ii];
                        const R impulse = c.h + (c.dh * ((R)(ph & 255u) / (R)256));
                        acc += (*xq * impulse);
                        xq += SRC_LD;
                        ph += phaseIncrement;
                    }
                    yw[j * SRC_LD + lane] = acc;
                    const R av = r_abs<R>(acc);
                    if (n0 + j < my_n_out && av > local_max) local_max = av;
                }
            }
            __syncwarp();
            // transposed write-back: each half-warp stores SRC_CHUNK consecutive samples of one utterance
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = 2 * i + (lane >> 4), cc = lane & 15;
                if (n0 + cc < s_n_out[r])
                    (reinterpret_cast<R *>(args.out) + s_out_off[r])[n0 + cc] = yw[cc * SRC_LD + r];
            }
            __syncwarp();
        }
        // per-utterance maximum: integer atomicMax on the bit pattern (order independent for non-negative doubles)
        if (local_max > (R)0) {
            const int u = args.tile_utt[tile * 32 + lane];
            if (u >= 0) atomicMax(args.maxbits + u, (unsigned long long)__double_as_longlong((double)local_max));
        }
    }
}

template <typename R>
__global__ void __launch_bounds__(PCM_THREADS) pcm_kernel(PcmArgs args)
{
    const int u = blockIdx.y;
    const trm_cuda_utterance *__restrict__ D = args.desc + u;
    const long long n_out = D->n_out;
    const long long n0 = ((long long)blockIdx.x * PCM_THREADS + threadIdx.x) * PCM_PER_THREAD;
    if (n0 >= n_out) return;
    const double mx = __longlong_as_double((long long)args.maxbits[u]);
    const double scale = (32767.0 / mx) * D->volumeAmp;                  // TRMTubeModel.m:515
    const R *__restrict__ z = reinterpret_cast<const R *>(args.out) + D->out_offset + n0;
    alignas(16) R v[PCM_PER_THREAD];
    const bool full = n0 + PCM_PER_THREAD <= n_out;
    if (full) {
        // out_offset and n0 are multiples of 8 elements: 128-bit loads
        constexpr int NV = (int)(PCM_PER_THREAD * sizeof(R) / 16);
        const float4 *zv = reinterpret_cast<const float4 *>(z);
        float4 *vv = reinterpret_cast<float4 *>(v);
#pragma unroll
        for (int i = 0; i < NV; ++i) vv[i] = zv[i];
    } else {
        for (int i = 0; i < PCM_PER_THREAD; ++i) v[i] = (n0 + i < n_out) ? z[i] : (R)0;
    }
    if (D->channels == 2) {
        const double ls = D->leftGain * scale, rs = D->rightGain * scale;  // TRMTubeModel.m:532-533
        int16_t *p = args.pcm + D->pcm_offset + 2 * n0;
        alignas(16) short2 q[PCM_PER_THREAD];
#pragma unroll
        for (int i = 0; i < PCM_PER_THREAD; ++i) {
            q[i].x = (short)__double2int_rn((double)v[i] * ls);
            q[i].y = (short)__double2int_rn((double)v[i] * rs);
        }
        if (full) {
            int4 *pv = reinterpret_cast<int4 *>(p);
            const int4 *qv = reinterpret_cast<const int4 *>(q);
            pv[0] = qv[0];
            pv[1] = qv[1];
        } else {
            for (int i = 0; i < PCM_PER_THREAD && n0 + i < n_out; ++i) reinterpret_cast<short2 *>(p)[i] = q[i];
        }
    } else {
        int16_t *p = args.pcm + D->pcm_offset + n0;
        alignas(16) short q[PCM_PER_THREAD];
#pragma unroll
        for (int i = 0; i < PCM_PER_THREAD; ++i) q[i] = (short)__double2int_rn((double)v[i] * scale);
        if (full) {
            *reinterpret_cast<int4 *>(p) = *reinterpret_cast<const int4 *>(q);
        } else {
            for (int i = 0; i < PCM_PER_THREAD && n0 + i < n_out; ++i) p[i] = q[i];
        }
    }
}

}  // namespace trm
