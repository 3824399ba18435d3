// framegen_kernel.cuh -- control-frame generator on the device.
//
// Replaces -[EventList generateOutputInTimeRange:forSynthesizer:parameterLogger:]
// (/root/reference/Frameworks/GnuSpeech/MonetModel/EventList.m:883-1061, full time range) and MMDriftGenerator
// (MMDriftGenerator.m:41-78): piece-wise linear interpolation of the 33 event tracks to 4 ms frames by repeated
// adds, micro / macro / smooth intonation sums, float rounding of the output table, drift (float MCG x 377 + one-pole
// low-pass).  The frames are written where the waveguide kernel reads them: a batch then needs only its sparse event
// lists uploaded (a few hundred bytes per event) instead of 128 B per 4 ms frame.
//
// One warp per utterance, lane j = event-track index j (0..31); tracks 32..35 (macro / smooth intonation) are
// carried redundantly by every lane.  The loop over frames is sequential, as in the reference; all control flow is
// warp-uniform except the per-track searches for the next non-NaN event value.  Arithmetic in the reference's
// order and types (double tracks, float table and drift; this TU is compiled with -fmad=false).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernel_args.h"
#include "trm_cuda.h"

namespace trm {

__global__ void __launch_bounds__(128) framegen_kernel(FrameGenArgs a)
{
    const int lane = threadIdx.x & 31;
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= a.n_utt) return;
    const trm_cuda_event *__restrict__ ev = a.events + a.ev_offset[u];
    const long long count = a.ev_count[u];
    const trm_cuda_framegen fg = a.fg[a.shared_fg ? 0 : u];
    const long long n_frames = a.desc[u].n_frames;
    double *__restrict__ out = a.frames + a.desc[u].frame_offset * 16;
    float seed = fg.driftSeed;
    if (count < 2) {
        if (a.seed_out && lane == 0) a.seed_out[u] = seed;
        return;
    }
    const double ms = 1000.0 / 250.0;
    // MMDriftGenerator -configureWithDeviation:sampleRate:lowpassCutoff: (m:41-58)
    float pitchDeviation = 0.0f, pitchOffset = 0.0f, a0 = 0.0f, b1 = 0.0f, previousSample = 0.0f;
    if (fg.useDrift) {
        const float deviation = (float)fg.driftDeviation, sampleRate = (float)(1000 / 4);
        float cutoff = (float)fg.driftCutoff;
        pitchDeviation = (float)((double)deviation * 2.0);
        pitchOffset = deviation;
        if ((double)cutoff < 0.0) cutoff = 0.0f;
        else if ((double)cutoff > ((double)sampleRate / 2.0)) cutoff = (float)((double)sampleRate / 2.0);
        a0 = (float)(((double)cutoff * 2.0) / (double)sampleRate);
        b1 = (float)(1.0 - (double)a0);
    }
    // tracks 0..31: this lane's current value and delta (m:919-930)
    double cv = 0.0, cd = 0.0;
    if (lane < 16) {
        long long j = 1;
        double temp;
        // (bounded: the reference walks past the end of the array when a track has no later value; libtrm refuses such
        //  lists -- check_event_lists -- and the kernel never reads outside this utterance's events)
        while (isnan(temp = ev[j].value[lane]) && j < count - 1) j++;
        cv = ev[0].value[lane];
        cd = isnan(temp) ? 0.0 : ((temp - cv) / (double)ev[j].time) * ms;
    }
    // tracks 32..35 (m:932-961)
    double cv32 = 0.0, cd32 = 0.0, cd33 = 0.0, cd34 = 0.0, cd35 = 0.0;
    {
        double temp;
        if (fg.useSmoothIntonation) {
            long long j = 0;
            while (isnan(temp = ev[j].value[32])) {
                j++;
                if (j >= count) break;
            }
            cv32 = j < count ? ev[j].value[32] : __longlong_as_double(0x7ff8000000000000ll);
            cd32 = 0.0;
        } else {
            long long j = 1;
            while (isnan(temp = ev[j].value[32])) {
                j++;
                if (j >= count) break;
            }
            cv32 = ev[0].value[32];
            if (j < count) cd32 = ((temp - cv32) / (double)ev[j].time) * ms;
            else cd32 = 0.0;
            cv32 = -20.0;
        }
    }

    long long i = 1, f = 0;
    unsigned long long currentTime_ms = 0ull, nextTime = (unsigned long long)ev[1].time;
    while (i < count) {                                              // m:973
        // table[j] = (float)currentValues[j] + (float)currentValues[j+16]  (m:974-976)
        const double hi = __shfl_down_sync(0xFFFFFFFFu, cv, 16);
        float t = __fadd_rn((float)cv, (float)hi);
        if (lane == 0) {
            if (!fg.useMicroIntonation) t = 0.0f;
        }
        {
            // drift (MMDriftGenerator.m:65-78): every lane keeps the generator, lane 0 uses it
            float d = 0.0f;
            if (fg.useDrift) {
                float temp = __fmul_rn(seed, 377.0f);
                seed = __fsub_rn(temp, (float)(int)temp);
                temp = __fsub_rn(__fmul_rn(seed, pitchDeviation), pitchOffset);
                previousSample = __fadd_rn(__fmul_rn(a0, temp), __fmul_rn(b1, previousSample));
                d = previousSample;
            }
            if (lane == 0) {
                if (fg.useDrift) t = __fadd_rn(t, d);
                if (fg.useMacroIntonation) t = (float)((double)t + cv32);
                t = (float)((double)t + fg.pitch);
            }
        }
        if (lane < 16 && f < n_frames) out[f * 16 + lane] = (double)t;      // 128 B per frame, coalesced
        ++f;
        if (cd != 0.0) cv += cd;                                         // m:1007-1010 (tracks 0..31)
        if (fg.useSmoothIntonation) {                                    // m:1011-1018
            cd34 += cd35;
            cd33 += cd34;
            cv32 += cd33;
        } else {
            if (cd32 != 0.0) cv32 += cd32;
        }
        currentTime_ms = (unsigned long long)((double)currentTime_ms + ms);

        if (currentTime_ms >= nextTime) {                                // m:1025
            i++;
            if (i == count) break;
            nextTime = (unsigned long long)ev[i].time;
            // new deltas for every track whose previous event carried a value (m:1031-1047): lane j < 32 its own
            // track, then track 32 by everybody
            for (int pass = 0; pass < 2; ++pass) {
                const int j = pass == 0 ? lane : 32;
                if (!isnan(ev[i - 1].value[j])) {
                    long long k = i;
                    double temp;
                    bool zeroed = false;
                    while (isnan(temp = ev[k].value[j])) {
                        if (k >= count - 1) { zeroed = true; break; }
                        k++;
                    }
                    const double cur = pass == 0 ? cv : cv32;
                    double nd = 0.0;
                    bool set = false;
                    if (zeroed) { nd = 0.0; set = true; }
                    if (!isnan(temp)) {
                        nd = (temp - cur) / (double)((unsigned long long)ev[k].time - currentTime_ms) * ms;
                        set = true;
                    }
                    if (set) { if (pass == 0) cd = nd; else cd32 = nd; }
                }
            }
            if (fg.useSmoothIntonation) {                                // m:1049-1057
                if (!isnan(ev[i - 1].value[33])) {
                    cv32 = ev[i - 1].value[32];
                    cd32 = 0.0;
                    cd33 = ev[i - 1].value[33];
                    cd34 = ev[i - 1].value[34];
                    cd35 = ev[i - 1].value[35];
                }
            }
        }
    }
    if (a.seed_out && lane == 0) a.seed_out[u] = seed;
}

}  // namespace trm
