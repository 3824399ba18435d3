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

template <typename R>
__global__ void __launch_bounds__(SRC_THREADS) src_kernel(SrcArgs args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HD<R> *tab = reinterpret_cast<HD<R> *>(smem_raw);
    R *xw = reinterpret_cast<R *>(tab + TRM_SRC_FILTER_LEN);
    __shared__ int s_u;

    {
        const HD<R> *g = reinterpret_cast<const HD<R> *>(args.table);
        for (int i = threadIdx.x; i < TRM_SRC_FILTER_LEN; i += SRC_THREADS) tab[i] = g[i];
    }

    for (long long tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
        __syncthreads();                       // previous tile's window no longer read; table staged
        if (threadIdx.x == 0) {
            int lo = 0, hi = args.n_utt;       // largest u with tile_base[u] <= tile
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (args.tile_base[mid] <= tile) lo = mid; else hi = mid;
            }
            s_u = lo;
        }
        __syncthreads();
        const int u = s_u;
        const trm_cuda_utterance *__restrict__ D = args.desc + u;
        const long long n_out = D->n_out, n_in = D->n_tube;
        const long long n_s = (tile - args.tile_base[u]) * SRC_TILE;
        const long long n_e = (n_s + SRC_TILE < n_out) ? n_s + SRC_TILE : n_out;
        const unsigned long long tri = D->tri;
        const int pad = D->padSize;
        const int reach = pad + 1;
        const long long P_lo = (long long)(((unsigned long long)n_s * tri) >> 16);
        const long long P_hi = (long long)(((unsigned long long)(n_e - 1) * tri) >> 16);
        const long long win_lo = P_lo - reach;
        const int win_len = (int)(P_hi + reach + 1 - win_lo + 1);
        const R *__restrict__ x = reinterpret_cast<const R *>(args.tube) + D->tube_offset;
        R *__restrict__ y = reinterpret_cast<R *>(args.out) + D->out_offset;

        for (int i = threadIdx.x; i < win_len && i < SRC_XW; i += SRC_THREADS) {
            const long long q = win_lo + i - pad;
            xw[i] = (q >= 0 && q < n_in) ? x[q] : (R)0;
        }
        __syncthreads();

        R local_max = (R)0;
        const bool up = D->upsample != 0;
        const double ratio = D->sampleRateRatio;
        const unsigned phaseIncrement = D->phaseIncrement;
        for (long long n = n_s + threadIdx.x; n < n_e; n += SRC_THREADS) {
            const unsigned long long T = (unsigned long long)n * tri;
            const int base = (int)((long long)(T >> 16) - win_lo);
            const unsigned F = (unsigned)(T & 0xFFFFull);
            R acc = (R)0;
            if (up) {
                R interp = (R)(F & 255u) / (R)256;
                unsigned fi = F >> 8;
#pragma unroll
                for (int k = 0; k < SRC_ZC; ++k) {
                    const HD<R> c = tab[fi + 256u * k];
                    acc += xw[base - k] * (c.h + c.dh * interp);
                }
                const unsigned G = (~F) & 0xFFFFu;
                interp = (R)(G & 255u) / (R)256;
                fi = G >> 8;
#pragma unroll
                for (int k = 0; k < SRC_ZC; ++k) {
                    const HD<R> c = tab[fi + 256u * k];
                    acc += xw[base + 1 + k] * (c.h + c.dh * interp);
                }
            } else {
                unsigned ph = (unsigned)rint((double)F * ratio), ii;
                int idx = base;
                while ((ii = (ph >> 8)) < (unsigned)TRM_SRC_FILTER_LEN) {
                    const HD<R> c = tab[ii];
                    const R impulse = c.h + (c.dh * ((R)(ph & 255u) / (R)256));
                    acc += (xw[idx] * impulse);
                    --idx;
                    ph += phaseIncrement;
                }
                ph = (unsigned)rint((double)((~F) & 0xFFFFu) * ratio);
                idx = base + 1;
                while ((ii = (ph >> 8)) < (unsigned)TRM_SRC_FILTER_LEN) {
                    const HD<R> c = tab[ii];
                    const R impulse = c.h + (c.dh * ((R)(ph & 255u) / (R)256));
                    acc += (xw[idx] * impulse);
                    ++idx;
                    ph += phaseIncrement;
                }
            }
            y[n] = acc;
            const R av = r_abs<R>(acc);
            if (av > local_max) local_max = av;          // NaN never wins, like the reference's compare
        }
        // per-utterance maximum: warp max, then one integer atomic per warp (bit order == value order for
        // non-negative doubles, so the result does not depend on the order of the atomics)
        double m = (double)local_max;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xFFFFFFFFu, m, o);
            m = (other > m) ? other : m;
        }
        if ((threadIdx.x & 31) == 0 && m > 0.0)
            atomicMax(args.maxbits + u, (unsigned long long)__double_as_longlong(m));
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
