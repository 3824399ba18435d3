// kernel_args.h -- plain structs and tile constants shared by the kernels (device) and the shim (host).
#pragma once

#include <stdint.h>

#include "trm_cuda.h"

// Every kernels_*.cu translation unit instantiates the kernel templates under its own compiler flags (FMA contraction
// on / off), so each one puts them into its own namespace: identical template instantiations from two TUs would otherwise
// be merged by the linker.  TRM_STRICT = 1 selects the reference-order FP64 arithmetic (tube_wide.cuh).
#ifndef TRM_KERNEL_NS
#define TRM_KERNEL_NS trm_kernels
#endif
#ifndef TRM_STRICT
#define TRM_STRICT 0
#endif

namespace trm {

constexpr int TB = 16;           // samples per block == lanes per utterance
constexpr int FIR_TAPS = 49;     // TRMFIRFilter.h:7-9 fixed design -> 49 taps (checked on the host)
constexpr int FIR_HIST = 24;     // (FIR_TAPS-1)/2 previous even / odd oscillator values
constexpr int FRAME_CHUNK = 2;   // control frames per bulk copy (256 B)


constexpr int SRC_ROWS = TRM_SRC_ROWS; // staged input rows per work item (include/trm_cuda.h: the host checks rate ratios against it)
constexpr int SRC_CHUNK = 8;         // consecutive outputs a warp finishes before the transposed write-back
constexpr int SRC_ZC = 13;           // zero crossings -> 13 taps per wing when up-sampling
constexpr int SRC_TAPS = 2 * SRC_ZC; // coefficients per output
constexpr int SRC_CLD = 28;          // row stride of the coefficient table (16-byte aligned rows in both precisions)
constexpr int PCM_THREADS = 256;
constexpr int PCM_PER_THREAD = 8;

// Carried state of one utterance for streaming synthesis (tube_wide.cuh): 4 eight-byte header words
// {oscillator position (double), the same in 2^-55 table entries (fast mode), noise generator state, "no sample
// produced yet" flag} followed by STATE_R values of the arithmetic type: oscillator history even / odd (24 + 24),
// band-pass input x[n-1], x[n-2], then the 41 recurrence values (waves, reflection / radiation memories, band-pass
// and throat outputs).
constexpr int STATE_HDR = 4;
constexpr int STATE_R = 96;
constexpr int STATE_HE = 0, STATE_HO = 24, STATE_XM = 48, STATE_SER = 50;
inline size_t tube_state_bytes(size_t esz) { return STATE_HDR * 8 + STATE_R * esz; }

struct TubeArgs {
    void *state;                 // streaming: [n_utt] carried states (loaded at start, stored at the end), else null
    const trm_cuda_utterance *desc;
    const int *order;            // slot -> utterance (longest first), may be null
    int n_utt;
    const void *frames;          // [frame][16] doubles, or floats when frames_f32
    int frames_f32;
    void *tube;                  // Real[...]
    const double *wavetables;    // [voice][512]
    uint64_t noise_k0;
};

template <typename R> struct alignas(2 * sizeof(R)) HD { R h, dh; };   // one vector load per entry

struct SrcArgs {
    const trm_cuda_utterance *desc;
    int n_utt;
    const void *tube;                // Real[]
    void *out;                       // Real[]
    unsigned long long *maxbits;     // [n_utt] bit pattern of the running max |y| as double
    const void *table;               // HD<Real>[256][13]: filter index l + 256 k at [l][k]
    const void *ctab;                // Real[65536][SRC_CLD]: interpolated coefficients per time-register fraction
    // work decomposition: tiles of <= 32 utterances that share the converter signature
    const int *tile_utt;             // [n_tiles][tile width] utterance index or -1 (width: KernelInfo.src[shape].tile)
    const int *tile_nt;              // [n_tiles] outputs per work item of that tile (window fits SRC_ROWS)
    const long long *tile_max_out;   // [n_tiles] longest utterance of the tile
    const long long *tile_first_out; // [n_tiles] first output of the tile's first work item (0 unless streaming)
    const long long *item_base;      // [n_tiles+1] prefix sum of work items
    int n_tiles;
    long long total_items;
    long long item_begin, item_end;  // work items this launch processes ([0, total_items) when item_end == 0)
};

struct PcmArgs {
    const trm_cuda_utterance *desc;
    int n_utt;
    const void *out;                 // Real[]
    const unsigned long long *maxbits;
    int16_t *pcm;
    int u_begin;                     // utterances [u_begin, n_utt) are scaled by this launch
};

// control-frame generator (framegen_kernel.cuh)
struct FrameGenArgs {
    const trm_cuda_utterance *desc;       // frame_offset / n_frames of every utterance
    int n_utt;
    const trm_cuda_event *events;         // all utterances back to back
    const long long *ev_offset;           // [n_utt] first event of utterance u
    const int *ev_count;                  // [n_utt]
    const trm_cuda_framegen *fg;          // [n_utt] or [1] if shared
    int shared_fg;
    double *frames;                       // [frame][16]
    float *seed_out;                      // [n_utt] drift seed at exit, may be null
};

struct KernelInfo {
    // resampler shapes (src_kernel.cuh SrcCfg): [0] handles every converter signature, [1] (if n_src_shapes == 2) is
    // the faster shape for up-sampling signatures whose work-item window fits its smaller staging buffer
    struct SrcShape {
        int smem_bytes, threads, ctas_per_sm, regs;
        int tile;                  // utterances per tile (32 x utterances per lane)
        int rows;                  // tube-rate samples per utterance staged for one work item
        int nt_max;                // outputs per work item, at most
    } src[2];
    int n_src_shapes;
    int pcm_regs;
    int wide_smem_bytes, wide_threads, wide_max_utt, wide_regs;   // batch-throughput waveguide mapping (tube_wide.cuh)
};

}  // namespace trm
