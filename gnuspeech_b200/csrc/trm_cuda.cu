// trm_cuda.cu -- C-ABI shim (include/trm_cuda.h): device memory, streams, chunked copy/compute
// pipeline and kernel launches.  No torch types, no C++ in the signatures.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>

#include "kernel_args.h"
#include "trm_cuda.h"

// launchers of the three kernel translation units (launch.cuh): FP64 conformance, FP32 fast, FP64 strict
#define TRM_DECLARE_LAUNCHERS(SUF)                                              \
    int trm_k_configure_##SUF(trm::KernelInfo *);                               \
    int trm_k_upload_##SUF(const double *, int, const unsigned long long *);    \
    int trm_k_tube_wide_##SUF(const trm::TubeArgs *, int, cudaStream_t);        \
    int trm_k_src_##SUF(const trm::SrcArgs *, int, int, cudaStream_t);          \
    int trm_k_src_ctab_##SUF(const void *, void *, cudaStream_t);               \
    int trm_k_pcm_##SUF(const trm::PcmArgs *, long long, cudaStream_t);
extern "C" {
TRM_DECLARE_LAUNCHERS(f64)
TRM_DECLARE_LAUNCHERS(f32)
TRM_DECLARE_LAUNCHERS(f64s)
int trm_k_framegen(const trm::FrameGenArgs *, cudaStream_t);
int trm_k_workload_walk2(unsigned long long, unsigned long long, long long, int, double *, cudaStream_t);
int trm_k_pcm_checksum(const trm_cuda_utterance *, int, const int16_t *, unsigned long long *, cudaStream_t);
}

// precision codes of include/trm.h: 0 = FP64 conformance, 1 = FP32 fast, 2 = FP64 strict
static inline bool prec_is_f64(int precision) { return precision != 1; }
static inline size_t prec_esz(int precision) { return precision != 1 ? sizeof(double) : sizeof(float); }
struct KernelSet {
    int (*tube_wide)(const trm::TubeArgs *, int, cudaStream_t);
    int (*src)(const trm::SrcArgs *, int, int, cudaStream_t);
    int (*pcm)(const trm::PcmArgs *, long long, cudaStream_t);
};
static const KernelSet g_kernels[3] = {
    {trm_k_tube_wide_f64, trm_k_src_f64, trm_k_pcm_f64},
    {trm_k_tube_wide_f32, trm_k_src_f32, trm_k_pcm_f32},
    {trm_k_tube_wide_f64s, trm_k_src_f64s, trm_k_pcm_f64s},
};

// FMA-chain kernels used to MEASURE the FP32 / FP64 CUDA-core peak of the device the bench runs on
// (MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers; the waveguide kernel is bound by neither).
template <typename T> __global__ void fp_peak_kernel(T *out, int iters)
{
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T m = (T)0.999999, c = (T)1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

namespace {

thread_local std::string g_err;

int fail(const char *what, cudaError_t e)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%d)", what, cudaGetErrorString(e), (int)e);
    g_err = buf;
    return -5;   // TRM_ERR_CUDA
}
int fail_msg(const char *what)
{
    g_err = what;
    return -5;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) return fail(#call, _e);                                             \
    } while (0)

constexpr int MAX_SLOTS = 16;  // upper bound of chunks in flight (each on its own stream and arena)
constexpr int MAX_OUT_GROUPS = 4;   // output groups of a chunk (ChunkPlan::groups)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Arena {
    unsigned char *base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&base, bytes);
        if (e != cudaSuccess) return fail("cudaMalloc(arena)", e);
        cap = bytes;
        return 0;
    }
    void reset() { used = 0; }
    void *take(size_t bytes)
    {
        size_t off = align_up(used, 256);
        used = off + bytes;
        return base + off;
    }
    void release()
    {
        if (base) cudaFree(base);
        base = nullptr;
        cap = used = 0;
    }
};

struct HostStage {
    unsigned char *base = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (base) cudaFreeHost(base);
        base = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost((void **)&base, bytes);
        if (e != cudaSuccess) return fail("cudaMallocHost(stage)", e);
        cap = bytes;
        return 0;
    }
    void release()
    {
        if (base) cudaFreeHost(base);
        base = nullptr;
        cap = 0;
    }
};

// Everything one chunk needs on the device, carved from an arena.
struct DeviceChunk {
    int n = 0;
    trm_cuda_utterance *desc = nullptr;
    int *order = nullptr;
    int *tile_utt = nullptr, *tile_nt = nullptr;
    long long *tile_max_out = nullptr, *tile_first_out = nullptr, *item_base = nullptr;
    unsigned long long *maxbits = nullptr;
    double *frames = nullptr;               // (holds float32 rows when the plan says f32_frames: the kernels widen on read)
    bool f32_frames = false;
    void *tube = nullptr, *out = nullptr;
    int16_t *pcm = nullptr;
    long long total_items = 0, max_n_out = 0;
    int n_tiles = 0;
    int src_shape = 0;                      // resampler shape the tiles were built for (KernelInfo::src[])
    // time split (ChunkPlan::split_frame > 0): descriptors of the two waveguide launches and the state carried between them
    trm_cuda_utterance *desc_t[2] = {nullptr, nullptr};
    void *state = nullptr;
    size_t tube_elems = 0, out_elems = 0, pcm_elems = 0, frame_rows = 0;
    const double *wavetables = nullptr;     // own copy of the glottal tables (device-resident batches); null = the context's
};

// Host-side plan of one chunk: rebased descriptors + where its spans live in the caller's arrays.
struct ChunkPlan {
    int u0 = 0, u1 = 0;
    std::vector<trm_cuda_utterance> desc;   // rebased to chunk-local offsets
    std::vector<int> order;
    // resampler work decomposition (src_kernel.cuh): tiles of <= 32 utterances with one converter signature
    std::vector<int> tile_utt, tile_nt;
    std::vector<long long> tile_max_out, tile_first_out, item_base;
    bool frames_dense = true;
    long long frames_lo = 0;                // host frame index of the span start (dense case)
    size_t frame_rows = 0;
    long long tube_lo = 0, out_lo = 0, pcm_lo = 0;   // host element offsets of the spans
    size_t tube_elems = 0, out_elems = 0, pcm_elems = 0;
    int src_shape = 0;
    // Output groups: contiguous utterance ranges whose resampling + scaling is launched separately, so that the PCM of
    // one group can leave for the host while the next group is still being resampled (tiles never span groups).
    struct Group { int u_begin, u_end; long long item_begin, item_end, max_n_out, pcm_lo, pcm_hi; };
    // Time split: when every utterance of the chunk has the same number of frames, the waveguide runs as two launches
    // (tube samples before / from control frame `split_frame`, recurrence state carried in `state` exactly as
    // trm_cuda_stream_push does), so that the upload of the later frames overlaps the first launch.  0 = one launch.
    int split_frame = 0, uniform_frames = 0;
    std::vector<trm_cuda_utterance> desc_t[2];
    std::vector<Group> groups;
    long long total_items = 0, max_n_out = 0;
    size_t n_tiles() const { return tile_nt.size(); }
    bool f32_frames = false;                // the caller's frames are float32 rows (64 bytes)
    size_t arena_bytes(size_t esz, bool want_pcm) const
    {
        size_t n = desc.size(), b = 0;
        auto add = [&](size_t x) { b = align_up(b, 256) + x; };
        add(n * sizeof(trm_cuda_utterance));
        add(n * sizeof(int));
        add(tile_utt.size() * sizeof(int));
        add(tile_nt.size() * sizeof(int));
        add(tile_max_out.size() * sizeof(long long));
        add(tile_first_out.size() * sizeof(long long));
        add(item_base.size() * sizeof(long long));
        add(n * sizeof(unsigned long long));
        add(frame_rows * 128);
        add(tube_elems * esz);
        add(out_elems * esz);
        if (want_pcm) add(pcm_elems * sizeof(int16_t));
        if (split_frame > 0) { add(n * sizeof(trm_cuda_utterance)); add(n * sizeof(trm_cuda_utterance)); add(n * trm::tube_state_bytes(esz)); }
        return b + 256;
    }
    size_t stage_bytes() const
    {
        size_t n = desc.size();
        return align_up(n * sizeof(trm_cuda_utterance), 256) + align_up(n * sizeof(int), 256) +
               align_up(tile_utt.size() * sizeof(int), 256) + align_up(tile_nt.size() * sizeof(int), 256) +
               align_up(tile_max_out.size() * sizeof(long long), 256) + align_up(tile_first_out.size() * sizeof(long long), 256) +
               align_up(item_base.size() * sizeof(long long), 256) +
               (split_frame > 0 ? 2 * align_up(n * sizeof(trm_cuda_utterance), 256) + align_up(n * trm::tube_state_bytes(8), 256) : 0) +
               align_up(n * sizeof(unsigned long long), 256);     // (last: the maxima come back here)
    }
};

}  // namespace

// One compute queue per device, shared by every context (lane) on it: the kernels of concurrent calls run in
// submission order -- waveguide, resampler, PCM of call k, then those of call k+1 -- instead of fighting for the SMs
// (a waveguide launch of call k+1 that slips in front of call k's resampler would delay k's copy-out by a whole
// waveguide kernel).  Copies stay on per-context streams.  The mutex covers the enqueue of one chunk's three kernels.
static cudaStream_t g_run_stream[64];
static std::mutex g_run_mu[64];
// Likewise one copy-in queue per device: the frames of concurrent calls go up one call after the other at full PCIe
// rate (the first call's kernels start after ITS upload, not after everybody's).
static cudaStream_t g_in_stream[64];
static std::mutex g_in_mu[64];

struct trm_cuda_ctx {
    int device = 0;
    int sm_count = 0;
    double *d_wavetables = nullptr;
    int wt_capacity = 0;
    std::vector<double> wt_host;      // what d_wavetables holds
    Arena gen;                        // frame generator: events, descriptors, generated frames
    void *d_tab_f64 = nullptr, *d_tab_f32 = nullptr;
    void *d_ctab_f64 = nullptr, *d_ctab_f32 = nullptr;   // interpolated converter coefficients per time-register fraction
    uint64_t noise_k0 = 0;
    trm::KernelInfo info[3]{};        // by precision code
    const trm::KernelInfo &ki(int precision) const { return info[precision]; }
    cudaStream_t streams[MAX_SLOTS]{};
    Arena arenas[MAX_SLOTS];
    HostStage stages[MAX_SLOTS];
    int n_slots = 3;              // chunks in flight: one uploading, one computing, one downloading
    cudaEvent_t ev_in[MAX_SLOTS]{}, ev_run[MAX_SLOTS]{}, ev_out[MAX_SLOTS]{};
    cudaEvent_t ev_in2[MAX_SLOTS]{};                    // time split: the later frames are on the device
    cudaEvent_t ev_grp[MAX_SLOTS][MAX_OUT_GROUPS]{};   // a group's PCM is complete (copy-out of the group may start)
};

struct trm_cuda_resident {
    trm_cuda_ctx *ctx = nullptr;
    int precision = 0;
    ChunkPlan plan;
    Arena arena;
    DeviceChunk dc;
    double *d_wavetables = nullptr;     // the batch's glottal tables: the context's set may be replaced by later calls
};

namespace {

int plan_chunk(const trm_cuda_utterance *desc, int u0, int u1, const trm::KernelInfo &ki, ChunkPlan &p, int want_groups = 1)
{
    const int n = u1 - u0;
    // The faster up-sampling shape needs every utterance of the chunk to up-sample with a work-item window that fits it.
    int shape = ki.n_src_shapes > 1 ? 1 : 0;
    for (int u = u0; u < u1 && shape == 1; ++u) {
        const auto &d = desc[u];
        const trm::KernelInfo::SrcShape &s1 = ki.src[1];
        const long long unit = (long long)trm::SRC_CHUNK * (s1.threads / 32);
        if (!d.upsample || d.tri == 0 || (long long)((double)(s1.rows - 3 - 2 * (d.padSize + 1)) * 65536.0 / (double)d.tri) < unit) shape = 0;
    }
    const trm::KernelInfo::SrcShape &sh = ki.src[shape];
    const int tile_width = sh.tile;
    p.src_shape = shape;
    p.u0 = u0;
    p.u1 = u1;
    p.desc.assign(desc + u0, desc + u1);
    // spans in the caller's arrays
    long long f_lo = INT64_MAX, f_hi = 0, f_sum = 0;
    long long t_lo = INT64_MAX, t_hi = 0, o_lo = INT64_MAX, o_hi = 0, c_lo = INT64_MAX, c_hi = 0;
    for (const auto &d : p.desc) {
        f_lo = std::min<long long>(f_lo, d.frame_offset);
        f_hi = std::max<long long>(f_hi, d.frame_offset + d.n_frames);
        f_sum += d.n_frames;
        t_lo = std::min<long long>(t_lo, d.tube_offset);
        t_hi = std::max<long long>(t_hi, d.tube_offset + d.n_tube);
        o_lo = std::min<long long>(o_lo, d.out_offset);
        o_hi = std::max<long long>(o_hi, d.out_offset + d.n_out);
        c_lo = std::min<long long>(c_lo, d.pcm_offset);
        c_hi = std::max<long long>(c_hi, d.pcm_offset + d.n_out * d.channels);
    }
    if (n == 0) { f_lo = t_lo = o_lo = c_lo = 0; }
    p.frames_dense = (f_hi - f_lo) <= f_sum + f_sum / 2 + 64;
    p.frames_lo = f_lo;
    p.tube_lo = t_lo; p.out_lo = o_lo; p.pcm_lo = c_lo;
    p.tube_elems = (size_t)align_up((size_t)(t_hi - t_lo), TRM_ALIGN_ELEMS);
    p.out_elems = (size_t)align_up((size_t)(o_hi - o_lo), TRM_ALIGN_ELEMS);
    p.pcm_elems = (size_t)align_up((size_t)(c_hi - c_lo), TRM_ALIGN_ELEMS);
    long long compact = 0;
    p.max_n_out = 0;
    for (int i = 0; i < n; ++i) {
        auto &d = p.desc[i];
        if (p.frames_dense) d.frame_offset -= f_lo;
        else { d.frame_offset = compact; compact += d.n_frames; }
        d.tube_offset -= t_lo;
        d.out_offset -= o_lo;
        d.pcm_offset -= c_lo;
        p.max_n_out = std::max<long long>(p.max_n_out, d.n_out);
    }
    {
        // resampler tiles: utterances that share (time-register increment, pad, direction, phase increment,
        // ratio) share every filter coefficient; longest first so the rows of a tile finish together
        auto key = [&](int a) {
            const auto &d = p.desc[a];
            unsigned long long rb;
            memcpy(&rb, &d.sampleRateRatio, sizeof rb);
            return std::make_tuple(d.tri, d.padSize, d.upsample, d.phaseIncrement, rb);
        };
        // group boundaries: contiguous index ranges with about equal shares of the output samples
        long long out_sum = 0;
        for (const auto &d : p.desc) out_sum += d.n_out;
        const int n_groups = std::max(1, std::min(want_groups, n / std::max(1, tile_width)));
        p.groups.clear();
        p.tile_utt.clear(); p.tile_nt.clear(); p.tile_max_out.clear(); p.tile_first_out.clear(); p.item_base.assign(1, 0);
        int g_at = 0;
        long long acc = 0;
        for (int g = 0; g < n_groups; ++g) {
            int g_end = g_at;
            if (g == n_groups - 1) g_end = n;
            else
                while (g_end < n - (n_groups - 1 - g) && (g_end == g_at || acc + p.desc[g_end].n_out <= out_sum * (g + 1) / n_groups)) acc += p.desc[g_end++].n_out;
            ChunkPlan::Group grp{g_at, g_end, p.item_base.back(), 0, 0, INT64_MAX, 0};
            std::vector<int> idx(g_end - g_at);
            std::iota(idx.begin(), idx.end(), g_at);
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
                const auto ka = key(a), kb = key(b);
                if (ka != kb) return ka < kb;
                return p.desc[a].n_out > p.desc[b].n_out;
            });
            const int m = (int)idx.size();
            for (int at = 0; at < m;) {
                int end = at;
                while (end < m && end - at < tile_width && key(idx[end]) == key(idx[at])) ++end;
                const auto &d0 = p.desc[idx[at]];
                if (d0.n_out > 0) {
                    const int reach = d0.padSize + 1;
                    // outputs per work item: the input window must fit the staged rows, every warp of the CTA gets the
                    // same whole number of SRC_CHUNK-sized runs
                    const long long unit = d0.upsample ? (long long)trm::SRC_CHUNK * (sh.threads / 32) : (long long)trm::SRC_CHUNK;
                    long long nt = (long long)((double)(sh.rows - 3 - 2 * reach) * 65536.0 / (double)d0.tri);
                    // (libtrm refuses such rate pairs when it derives the descriptor; this guards the C-ABI itself)
                    if (nt < unit) return fail_msg("resampler: the input window of one work item does not fit the staged rows (rate ratio too small)");
                    nt = std::min<long long>(nt, sh.nt_max);
                    nt = nt / unit * unit;
                    long long first = d0.n_out;              // streaming: the tile starts at its earliest missing output
                    for (int r = at; r < end; ++r) first = std::min<long long>(first, p.desc[idx[r]].out_start);
                    first = first / nt * nt;
                    for (int r = 0; r < tile_width; ++r) p.tile_utt.push_back(at + r < end ? idx[at + r] : -1);
                    p.tile_nt.push_back((int)nt);
                    p.tile_max_out.push_back(d0.n_out);
                    p.tile_first_out.push_back(first);
                    p.item_base.push_back(p.item_base.back() + (d0.n_out - first + nt - 1) / nt);
                }
                at = end;
            }
            grp.item_end = p.item_base.back();
            for (int u = g_at; u < g_end; ++u) {
                const auto &d = p.desc[u];
                grp.max_n_out = std::max<long long>(grp.max_n_out, d.n_out);
                grp.pcm_lo = std::min<long long>(grp.pcm_lo, d.pcm_offset);
                grp.pcm_hi = std::max<long long>(grp.pcm_hi, d.pcm_offset + d.n_out * d.channels);
            }
            if (grp.pcm_lo > grp.pcm_hi) grp.pcm_lo = grp.pcm_hi = 0;
            p.groups.push_back(grp);
            g_at = g_end;
        }
        p.total_items = p.item_base.back();
    }
    p.frame_rows = (size_t)(p.frames_dense ? (f_hi - f_lo) : compact);
    // longest utterances first, so the two utterances of a warp (and the warps of a CTA) finish together
    p.order.resize(n);
    std::iota(p.order.begin(), p.order.end(), 0);
    std::stable_sort(p.order.begin(), p.order.end(),
                     [&](int a, int b) { return p.desc[a].n_tube > p.desc[b].n_tube; });
    // time split candidates: dense frames, the same frame count everywhere (so that a frame range is one strided copy) and
    // one control period (so that every utterance of a CTA reaches the split after the same number of 16-sample blocks:
    // the kernels save an utterance's state when its CTA's loop ends)
    p.split_frame = 0;
    p.uniform_frames = 0;
    if (n > 0 && p.frames_dense) {
        const int nf = p.desc[0].n_frames;
        bool uniform = nf > 0;
        for (int i = 0; i < n && uniform; ++i)
            uniform = p.desc[i].n_frames == nf && p.desc[i].frame_offset == (long long)i * nf && p.desc[i].jc0 == 0 && p.desc[i].n_tube > 0 &&
                      p.desc[i].controlPeriod == p.desc[0].controlPeriod;
        if (uniform) p.uniform_frames = nf;
    }
    return 0;
}

// Arms the time split of a planned chunk: launch 0 covers the tube samples before control frame f1 (rounded down to a
// whole 16-sample block per utterance), launch 1 the rest, continuing from the saved state in the middle of a control
// interval (jc0) exactly like a streaming push.
void arm_time_split(ChunkPlan &p, int f1)
{
    const size_t n = p.desc.size();
    p.split_frame = f1;
    p.desc_t[0].assign(p.desc.begin(), p.desc.end());
    p.desc_t[1].assign(p.desc.begin(), p.desc.end());
    for (size_t u = 0; u < n; ++u) {
        const trm_cuda_utterance &d = p.desc[u];
        const long long cp = d.controlPeriod;
        const long long t1 = ((long long)f1 * cp) / 16 * 16;          // first tube sample of launch 1
        const long long f0 = t1 / cp;                                 // control interval it lies in
        trm_cuda_utterance &a = p.desc_t[0][u], &b = p.desc_t[1][u];
        a.n_tube = t1;
        a.n_frames = f1 + 1;
        b.frame_offset = d.frame_offset + f0;
        b.n_frames = d.n_frames - (int)f0;
        b.jc0 = (int)(t1 - f0 * cp);
        b.tube_offset = d.tube_offset + t1;
        b.n_tube = d.n_tube - t1;
    }
}

void carve(Arena &a, const ChunkPlan &p, size_t esz, bool want_pcm, DeviceChunk &dc)
{
    const size_t n = p.desc.size();
    a.reset();
    dc.n = (int)n;
    dc.desc = (trm_cuda_utterance *)a.take(n * sizeof(trm_cuda_utterance));
    dc.order = (int *)a.take(n * sizeof(int));
    dc.tile_utt = (int *)a.take(p.tile_utt.size() * sizeof(int));
    dc.tile_nt = (int *)a.take(p.tile_nt.size() * sizeof(int));
    dc.tile_max_out = (long long *)a.take(p.tile_max_out.size() * sizeof(long long));
    dc.tile_first_out = (long long *)a.take(p.tile_first_out.size() * sizeof(long long));
    dc.item_base = (long long *)a.take(p.item_base.size() * sizeof(long long));
    dc.maxbits = (unsigned long long *)a.take(n * sizeof(unsigned long long));
    dc.frames = (double *)a.take(p.frame_rows * 128);
    dc.f32_frames = p.f32_frames;
    dc.tube = a.take(p.tube_elems * esz);
    dc.out = a.take(p.out_elems * esz);
    dc.pcm = want_pcm ? (int16_t *)a.take(p.pcm_elems * sizeof(int16_t)) : nullptr;
    if (p.split_frame > 0) {
        dc.desc_t[0] = (trm_cuda_utterance *)a.take(n * sizeof(trm_cuda_utterance));
        dc.desc_t[1] = (trm_cuda_utterance *)a.take(n * sizeof(trm_cuda_utterance));
        dc.state = a.take(n * trm::tube_state_bytes(esz));
    } else {
        dc.desc_t[0] = dc.desc_t[1] = nullptr;
        dc.state = nullptr;
    }
    dc.total_items = p.total_items;
    dc.n_tiles = (int)p.n_tiles();
    dc.src_shape = p.src_shape;
    dc.max_n_out = p.max_n_out;
    dc.tube_elems = p.tube_elems; dc.out_elems = p.out_elems; dc.pcm_elems = p.pcm_elems; dc.frame_rows = p.frame_rows;
}

// small tables (descriptors, order, tile prefix) -> device, through pinned staging when given
int upload_plan(const ChunkPlan &p, const DeviceChunk &dc, unsigned char *stage, cudaStream_t s, size_t esz = 8, unsigned long long noise_k0 = 0)
{
    const size_t n = p.desc.size();
    if (n == 0) return 0;
    unsigned char *q = stage;
    auto put = [&](void *dst, const void *src, size_t bytes) -> int {
        if (bytes == 0) return 0;
        const void *from = src;
        if (stage) {                       // pinned staging keeps the copy asynchronous
            memcpy(q, src, bytes);
            from = q;
            q += align_up(bytes, 256);
        }
        CK(cudaMemcpyAsync(dst, from, bytes, cudaMemcpyHostToDevice, s));
        return 0;
    };
    int rc;
    if ((rc = put(dc.desc, p.desc.data(), n * sizeof(trm_cuda_utterance))) != 0) return rc;
    if ((rc = put(dc.order, p.order.data(), n * sizeof(int))) != 0) return rc;
    if ((rc = put(dc.tile_utt, p.tile_utt.data(), p.tile_utt.size() * sizeof(int))) != 0) return rc;
    if ((rc = put(dc.tile_nt, p.tile_nt.data(), p.tile_nt.size() * sizeof(int))) != 0) return rc;
    if ((rc = put(dc.tile_max_out, p.tile_max_out.data(), p.tile_max_out.size() * sizeof(long long))) != 0) return rc;
    if ((rc = put(dc.tile_first_out, p.tile_first_out.data(), p.tile_first_out.size() * sizeof(long long))) != 0) return rc;
    if ((rc = put(dc.item_base, p.item_base.data(), p.item_base.size() * sizeof(long long))) != 0) return rc;
    if (p.split_frame > 0 && dc.state) {
        if ((rc = put(dc.desc_t[0], p.desc_t[0].data(), n * sizeof(trm_cuda_utterance))) != 0) return rc;
        if ((rc = put(dc.desc_t[1], p.desc_t[1].data(), n * sizeof(trm_cuda_utterance))) != 0) return rc;
        // fresh recurrence state: everything zero, noise generator at its start, "no sample yet" flag set
        const size_t sb = trm::tube_state_bytes(esz);
        std::vector<unsigned char> init(n * sb, 0);
        for (size_t u = 0; u < n; ++u) {
            unsigned long long *h = (unsigned long long *)(init.data() + u * sb);
            h[2] = noise_k0;
            h[3] = 1ull;
        }
        if ((rc = put(dc.state, init.data(), init.size())) != 0) return rc;
    }
    return 0;
}

int upload_frames(const ChunkPlan &p, const DeviceChunk &dc, const trm_cuda_utterance *desc_global,
                  const void *frames_host, cudaStream_t s)
{
    if (p.frame_rows == 0) return 0;
    const size_t row = p.f32_frames ? 64 : 128;                     // bytes per frame in the caller's array and on the device
    unsigned char *dst = (unsigned char *)dc.frames;
    const unsigned char *src = (const unsigned char *)frames_host;
    if (p.frames_dense) {
        CK(cudaMemcpyAsync(dst, src + (size_t)p.frames_lo * row, p.frame_rows * row, cudaMemcpyDefault, s));
    } else {
        for (size_t i = 0; i < p.desc.size(); ++i) {
            const auto &g = desc_global[p.u0 + i];
            CK(cudaMemcpyAsync(dst + (size_t)p.desc[i].frame_offset * row, src + (size_t)g.frame_offset * row,
                               (size_t)g.n_frames * row, cudaMemcpyDefault, s));
        }
    }
    return 0;
}

// CTAs of the waveguide kernel for a chunk of n utterances: one per SM and wave, at most wide_max_utt utterances each; small
// chunks are spread two utterances per CTA (one feed-forward warp each) over as many SMs as there are pairs.
int wide_groups(const trm_cuda_ctx *ctx, const trm::KernelInfo &ki, int n)
{
    if (n <= 0) return 0;
    const int sm = ctx->sm_count, gmax = ki.wide_max_utt;
    if ((long long)n <= (long long)sm * gmax) return std::max(1, std::min(sm, (n + 1) / 2));
    const int waves = (int)(((long long)n + (long long)sm * gmax - 1) / ((long long)sm * gmax));
    return waves * sm;
}

int launch_stage(trm_cuda_ctx *ctx, int precision, int stage, const DeviceChunk &dc, cudaStream_t s, const ChunkPlan::Group *grp = nullptr,
                 int time_part = -1)
{
    const bool f64 = prec_is_f64(precision);
    const KernelSet &K = g_kernels[precision];
    int rc = 0;
    if (stage == TRM_STAGE_TUBE) {
        trm::TubeArgs a{};
        a.desc = dc.desc; a.order = dc.order; a.n_utt = dc.n; a.frames = dc.frames; a.frames_f32 = dc.f32_frames ? 1 : 0; a.tube = dc.tube;
        a.wavetables = dc.wavetables ? dc.wavetables : ctx->d_wavetables; a.noise_k0 = ctx->noise_k0;
        if (time_part >= 0) { a.desc = dc.desc_t[time_part]; a.state = dc.state; }     // (lane-per-utterance mapping only)
        const trm::KernelInfo &ki = ctx->ki(precision);
        const int groups = wide_groups(ctx, ki, dc.n);
        if (groups > 0) rc = K.tube_wide(&a, groups, s);
    } else if (stage == TRM_STAGE_SRC) {
        // (the running maxima are cleared once per chunk: by the ungrouped launch, or by the first group's)
        if (!grp || grp->u_begin == 0) CK(cudaMemsetAsync(dc.maxbits, 0, (size_t)dc.n * sizeof(unsigned long long), s));
        trm::SrcArgs a{};
        a.desc = dc.desc; a.n_utt = dc.n; a.tube = dc.tube; a.out = dc.out; a.maxbits = dc.maxbits;
        a.table = f64 ? ctx->d_tab_f64 : ctx->d_tab_f32;
        a.ctab = f64 ? ctx->d_ctab_f64 : ctx->d_ctab_f32;
        a.tile_utt = dc.tile_utt; a.tile_nt = dc.tile_nt; a.tile_max_out = dc.tile_max_out; a.tile_first_out = dc.tile_first_out;
        a.item_base = dc.item_base;
        a.n_tiles = dc.n_tiles; a.total_items = dc.total_items;
        if (grp) { a.item_begin = grp->item_begin; a.item_end = grp->item_end; if (a.item_end <= a.item_begin) return 0; }
        const trm::KernelInfo &ki = ctx->ki(precision);
        const int grid = ctx->sm_count * std::max(1, ki.src[dc.src_shape].ctas_per_sm);
        rc = K.src(&a, grid, dc.src_shape, s);
    } else if (stage == TRM_STAGE_PCM) {
        if (!dc.pcm) return 0;
        trm::PcmArgs a{};
        a.desc = dc.desc; a.n_utt = grp ? grp->u_end : dc.n; a.u_begin = grp ? grp->u_begin : 0;
        a.out = dc.out; a.maxbits = dc.maxbits; a.pcm = dc.pcm;
        const long long longest = grp ? grp->max_n_out : dc.max_n_out;
        rc = K.pcm(&a, longest, s);
    }
    if (rc != 0) return fail("kernel launch", (cudaError_t)rc);
    return 0;
}

int chunk_utterances(const trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc, size_t esz)
{
    // Chunk size: the kernels of one chunk run alone on the device, so a chunk must be large enough to fill it --
    // for the batch-throughput waveguide mapping that is >= ~14 utterances per SM (below that its recurrence warp,
    // not the feed-forward work, sets the time) -- and small enough that there are at least two chunks, so that the
    // PCM of one chunk leaves for the host while the next one is computed.  Bounded by scratch memory.
    const char *env = getenv("TRM_CHUNK_UTTERANCES");
    if (env && atoi(env) > 0) return atoi(env);
    // scratch of a chunk = frames + tube-rate + output-rate samples + PCM of its utterances; the bound uses the LARGEST
    // utterance of the batch (ragged batches: an average over a prefix would let one chunk of long utterances overshoot)
    // and what the device can actually give three in-flight chunks per context lane
    size_t per_utt = 1;
    for (int i = 0; i < n; ++i)
        per_utt = std::max(per_utt, (size_t)desc[i].n_frames * 128 + ((size_t)desc[i].n_tube + (size_t)desc[i].n_out) * esz +
                                        (size_t)desc[i].n_out * 2 * desc[i].channels + 1024);
    size_t free_b = 0, total_b = 0, reserved = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { free_b = (size_t)96 << 30; cudaGetLastError(); }
    for (int i = 0; i < MAX_SLOTS; ++i) reserved += ctx->arenas[i].cap;          // already ours: reused, not needed again
    const size_t budget = std::min<size_t>((size_t)32 << 30, (size_t)(0.85 * (double)(free_b + reserved)) / 3);
    const long long by_mem = std::max<long long>(1, (long long)(budget / per_utt));
    const trm::KernelInfo &ki = ctx->ki(precision);
    const long long lo = (long long)ctx->sm_count * (ki.wide_max_utt / 2), hi = (long long)ctx->sm_count * ki.wide_max_utt;
    // FP64: a launch below ~28 utterances per SM is bound by the latency of its feed-forward warps, so halving a batch
    // costs more kernel time than the overlapped copy-out saves -> one chunk up to the device's capacity.  FP32: the
    // kernels are short next to the PCIe time of their PCM -> two chunks, the first one's PCM hides behind the second.
    long long want = prec_is_f64(precision) ? hi : std::min<long long>(hi, std::max<long long>(lo, (n + 1) / 2));
    long long c = std::max<long long>(1, std::min<long long>(std::min<long long>(want, by_mem), n));
    const long long n_chunks = (n + c - 1) / c;                 // balance the chunks
    return (int)((n + n_chunks - 1) / n_chunks);
}

}  // namespace

extern "C" {

const char *trm_cuda_last_error(void) { return g_err.c_str(); }

// The pipeline keeps up to MAX_SLOTS streams busy; with the default of 8 hardware work queues, streams alias
// onto the same queue and one chunk's kernels wait behind another chunk's copies.  Must be set before the
// CUDA context exists, so it is done on the first entry into the library (no effect if the host application
// already initialised CUDA: such callers set CUDA_DEVICE_MAX_CONNECTIONS=32 themselves).
static void want_many_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

int trm_cuda_device_count(void)
{
    want_many_connections();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { fail("cudaGetDeviceCount", e); return 0; }
    return n;
}

void *trm_cuda_host_alloc(size_t bytes)
{
    want_many_connections();
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail("cudaMallocHost", e); return nullptr; }
    return p;
}

void trm_cuda_host_free(void *p) { if (p) cudaFreeHost(p); }

/* Measured FMA peak in TFLOP/s (2 flops per FMA) for precision 0 = FP64, 1 = FP32; best of `reps` runs. */
int trm_cuda_fp_peak(int device, int precision, int reps, double *tflops)
{
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = precision != 1 ? 1 << 14 : 1 << 16;
    void *buf = nullptr;
    CK(cudaMalloc(&buf, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0;
    for (int r = 0; r < reps + 1; ++r) {
        CK(cudaEventRecord(e0, 0));
        if (precision != 1) fp_peak_kernel<double><<<blocks, threads>>>((double *)buf, iters);
        else fp_peak_kernel<float><<<blocks, threads>>>((float *)buf, iters);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops = best;
    return 0;
}

int trm_cuda_stage_launches(int stage) { return (stage >= 0 && stage < TRM_STAGE_COUNT) ? 1 : 0; }

int trm_cuda_ctx_create(int device, const trm_cuda_tables *t, trm_cuda_ctx **out)
{
    *out = nullptr;
    want_many_connections();
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail_msg("trm_cuda_ctx_create: no such CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail_msg("trm_cuda_ctx_create: kernels are built for sm_100a only");
    trm_cuda_ctx *c = new trm_cuda_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->noise_k0 = t->noise_k0;
    int rc;
    if ((rc = trm_k_upload_f64(t->fir_coef, t->fir_taps, (const unsigned long long *)t->noise_pow)) != 0 ||
        (rc = trm_k_upload_f32(t->fir_coef, t->fir_taps, (const unsigned long long *)t->noise_pow)) != 0 ||
        (rc = trm_k_upload_f64s(t->fir_coef, t->fir_taps, (const unsigned long long *)t->noise_pow)) != 0) {
        delete c;
        return rc < 0 ? fail_msg("FIR design is not the 49-tap filter the kernels are built for") : fail("constant upload", (cudaError_t)rc);
    }
    if ((rc = trm_k_configure_f64(&c->info[0])) != 0 || (rc = trm_k_configure_f32(&c->info[1])) != 0 ||
        (rc = trm_k_configure_f64s(&c->info[2])) != 0) {
        delete c;
        return fail("kernel configuration", (cudaError_t)rc);
    }
    // interleaved (h, deltaH) tables in both precisions
    {
        std::vector<trm::HD<double>> td(TRM_SRC_FILTER_LEN);
        std::vector<trm::HD<float>> tf(TRM_SRC_FILTER_LEN);
        // phase-major: entry (l, k) = filter index l + 256 k sits at l * 13 + k, so the 13 taps of one wing are contiguous
        for (int i = 0; i < TRM_SRC_FILTER_LEN; ++i) {
            const int at = (i & 255) * trm::SRC_ZC + (i >> 8);
            td[at].h = t->src_h[i]; td[at].dh = t->src_dh[i];
            tf[at].h = (float)t->src_h[i]; tf[at].dh = (float)t->src_dh[i];
        }
        CK(cudaMalloc(&c->d_tab_f64, td.size() * sizeof(td[0])));
        CK(cudaMalloc(&c->d_tab_f32, tf.size() * sizeof(tf[0])));
        CK(cudaMemcpy(c->d_tab_f64, td.data(), td.size() * sizeof(td[0]), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_tab_f32, tf.data(), tf.size() * sizeof(tf[0]), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&c->d_ctab_f64, (size_t)65536 * trm::SRC_CLD * sizeof(double)));
        CK(cudaMalloc(&c->d_ctab_f32, (size_t)65536 * trm::SRC_CLD * sizeof(float)));
        if ((rc = trm_k_src_ctab_f64s(c->d_tab_f64, c->d_ctab_f64, 0)) != 0 || (rc = trm_k_src_ctab_f32(c->d_tab_f32, c->d_ctab_f32, 0)) != 0) {
            delete c;
            return fail("converter coefficient table", (cudaError_t)rc);
        }
        CK(cudaDeviceSynchronize());
    }
    {
        // streams[2] = copy-out of the chunk pipeline (trm_cuda_synthesize_host); frames and kernels go to the device's
        // shared copy-in and compute queues
        {
            std::lock_guard<std::mutex> lk(g_run_mu[device]);
            if (!g_run_stream[device]) CK(cudaStreamCreateWithFlags(&g_run_stream[device], cudaStreamNonBlocking));
            if (!g_in_stream[device]) CK(cudaStreamCreateWithFlags(&g_in_stream[device], cudaStreamNonBlocking));
        }
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));          // lo = least (numerically largest)
        const char *env = getenv("TRM_SLOTS");
        if (env && atoi(env) > 0) c->n_slots = std::min(MAX_SLOTS, atoi(env));
        for (int i = 0; i < MAX_SLOTS; ++i) {
            const int prio = std::min(lo, hi + i);
            CK(cudaStreamCreateWithPriority(&c->streams[i], cudaStreamNonBlocking, prio));
            CK(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_run[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_in2[i], cudaEventDisableTiming));
            for (int g = 0; g < MAX_OUT_GROUPS; ++g) CK(cudaEventCreateWithFlags(&c->ev_grp[i][g], cudaEventDisableTiming));
        }
    }
    *out = c;
    return 0;
}

int trm_cuda_set_wavetables(trm_cuda_ctx *c, const double *tables, int n_voices)
{
    if (n_voices <= 0) return fail_msg("trm_cuda_set_wavetables: no voices");
    const size_t n = (size_t)n_voices * TRM_TABLE_LENGTH;
    // unchanged since the last call (the usual case: one voice set per application): nothing to do, and in particular
    // no synchronisation that would serialise this call behind another lane's work on the same device
    if (c->wt_host.size() == n && memcmp(c->wt_host.data(), tables, n * sizeof(double)) == 0) return 0;
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < 3; ++i) CK(cudaStreamSynchronize(c->streams[i]));   // this context's own work only
    if (n_voices > c->wt_capacity) {
        if (c->d_wavetables) cudaFree(c->d_wavetables);
        c->d_wavetables = nullptr;
        c->wt_capacity = 0;
        CK(cudaMalloc((void **)&c->d_wavetables, n * sizeof(double)));
        c->wt_capacity = n_voices;
    }
    CK(cudaMemcpy(c->d_wavetables, tables, n * sizeof(double), cudaMemcpyHostToDevice));
    c->wt_host.assign(tables, tables + n);
    return 0;
}

void trm_cuda_ctx_destroy(trm_cuda_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &s : c->streams) if (s) cudaStreamDestroy(s);
    for (int i = 0; i < MAX_SLOTS; ++i) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_run[i]) cudaEventDestroy(c->ev_run[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
        if (c->ev_in2[i]) cudaEventDestroy(c->ev_in2[i]);
        for (int g = 0; g < MAX_OUT_GROUPS; ++g) if (c->ev_grp[i][g]) cudaEventDestroy(c->ev_grp[i][g]);
    }
    for (auto &a : c->arenas) a.release();
    c->gen.release();
    for (auto &h : c->stages) h.release();
    if (c->d_wavetables) cudaFree(c->d_wavetables);
    if (c->d_tab_f64) cudaFree(c->d_tab_f64);
    if (c->d_tab_f32) cudaFree(c->d_tab_f32);
    if (c->d_ctab_f64) cudaFree(c->d_ctab_f64);
    if (c->d_ctab_f32) cudaFree(c->d_ctab_f32);
    delete c;
}

int trm_cuda_synthesize_host(trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc,
                             const double *frames_host, int16_t *pcm_host, void *samples_host, double *max_host,
                             void *tube_host, int64_t *launches)
{
    return trm_cuda_synthesize_host_ex(ctx, precision, n, desc, frames_host, pcm_host, samples_host, max_host, tube_host, launches,
                                       nullptr, nullptr);
}

int trm_cuda_synthesize_host_ex(trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc,
                                const double *frames_host, int16_t *pcm_host, void *samples_host, double *max_host,
                                void *tube_host, int64_t *launches, void (*enqueued)(void *), void *enqueued_arg)
{
    return trm_cuda_synthesize_host_fmt(ctx, precision, 0, n, desc, frames_host, pcm_host, samples_host, max_host, tube_host, launches,
                                        enqueued, enqueued_arg);
}

int trm_cuda_synthesize_host_fmt(trm_cuda_ctx *ctx, int precision, int frame_format, int n, const trm_cuda_utterance *desc,
                                 const void *frames_host, int16_t *pcm_host, void *samples_host, double *max_host,
                                 void *tube_host, int64_t *launches, void (*enqueued)(void *), void *enqueued_arg)
{
    if (frame_format != 0 && frame_format != 1) return fail_msg("unknown frame format");
    const bool f32_frames = frame_format == 1;
    if (launches) *launches = 0;
    if (n <= 0) return 0;
    CK(cudaSetDevice(ctx->device));
    if (precision < 0 || precision > 2) return fail_msg("unknown precision code");
    const size_t esz = prec_esz(precision);
    const bool want_pcm = pcm_host != nullptr;
    const int per_chunk = chunk_utterances(ctx, precision, n, desc, esz);
    const int n_chunks = (n + per_chunk - 1) / per_chunk;
    const int N_SLOTS = std::min(ctx->n_slots, std::max(n_chunks, 1));
    // Three-stage pipeline over chunks: frames go up on the copy-in stream, the three kernels of a chunk run on the
    // compute stream (one chunk at a time: a waveguide launch fills every SM), PCM / samples come back on the copy-out
    // stream.  H2D of chunk k+1 and D2H of chunk k-1 overlap the kernels of chunk k (full-duplex PCIe).
    cudaStream_t s_in = g_in_stream[ctx->device], s_run = g_run_stream[ctx->device], s_out = ctx->streams[2];
    // TRM_TRACE=1: per-chunk timeline of the pipeline stages (CUDA events), printed to stderr
    const bool trace = getenv("TRM_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    static cudaEvent_t t_origin = nullptr;          // one time axis for every traced call of the process
    static std::mutex t_mutex;
    static double t_host0 = 0.0;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec; };
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    if (trace) {
        std::lock_guard<std::mutex> lk(t_mutex);
        if (!t_origin) { cudaEventCreate(&t_origin); cudaEventRecord(t_origin, s_out); cudaEventSynchronize(t_origin); t_host0 = now_ms(); }
    }
    const double t_enter = trace ? now_ms() - t_host0 : 0.0;
    std::vector<ChunkPlan> plans(N_SLOTS);
    std::vector<int> slot_chunk(N_SLOTS, -1);
    int64_t n_launch = 0;

    auto finish_slot = [&](int slot) -> int {
        // wait for the slot's copy-out, then hand the per-utterance maxima to the caller
        CK(cudaEventSynchronize(ctx->ev_out[slot]));
        const int ci = slot_chunk[slot];
        if (ci >= 0 && max_host) {
            const ChunkPlan &p = plans[slot];
            const unsigned long long *mb =
                (const unsigned long long *)(ctx->stages[slot].base + p.stage_bytes() - align_up(p.desc.size() * sizeof(unsigned long long), 256));
            for (size_t i = 0; i < p.desc.size(); ++i) {
                double v;
                memcpy(&v, &mb[i], sizeof v);
                max_host[p.u0 + i] = v;
            }
        }
        slot_chunk[slot] = -1;
        return 0;
    };

    // An error inside the chunk loop must not return while earlier chunks' copies are still landing in the caller's buffers
    // (the caller may free them as soon as the call fails): drain this call's queues first.
    auto drain = [&]() {
        cudaStreamSynchronize(s_in);
        cudaStreamSynchronize(s_run);
        cudaStreamSynchronize(s_out);
    };
#define CKD(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) { drain(); return fail(#call, _e); }                                \
    } while (0)
#define RCD(expr)                                                                                  \
    do {                                                                                           \
        if ((rc = (expr)) != 0) { drain(); return rc; }                                            \
    } while (0)
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int slot = ci % N_SLOTS;
        int rc;
        if (slot_chunk[slot] >= 0) RCD(finish_slot(slot));
        ChunkPlan &p = plans[slot];
        const int u0 = ci * per_chunk, u1 = std::min(n, u0 + per_chunk);
        // large chunks with PCM output are resampled and scaled in up to MAX_OUT_GROUPS output groups
        long long chunk_out = 0;
        for (int u = u0; u < u1; ++u) chunk_out += desc[u].n_out;
        const int want_groups = (want_pcm && !getenv("TRM_NO_OUT_GROUPS")) ? (int)std::min<long long>(MAX_OUT_GROUPS, chunk_out / (64ll << 20)) : 1;
        RCD(plan_chunk(desc, u0, u1, ctx->ki(precision), p, want_groups));
        p.f32_frames = f32_frames;
        const bool grouped = p.groups.size() > 1;
        // long uniform chunks: the waveguide as two launches in time, the later frames uploaded behind the first
        if (p.uniform_frames >= 512 && !getenv("TRM_NO_TIME_SPLIT") &&
            wide_groups(ctx, ctx->ki(precision), u1 - u0) > 0)
            arm_time_split(p, std::max(64, (int)(0.28 * p.uniform_frames)));
        const bool split = p.split_frame > 0;
        RCD(ctx->arenas[slot].reserve(p.arena_bytes(esz, want_pcm)));
        RCD(ctx->stages[slot].reserve(p.stage_bytes()));
        DeviceChunk dc;
        carve(ctx->arenas[slot], p, esz, want_pcm, dc);
        // ---- copy-in -----------------------------------------------------------------------------------------
        {
            std::lock_guard<std::mutex> lk(g_in_mu[ctx->device]);
            mark(s_in);
            RCD(upload_plan(p, dc, ctx->stages[slot].base, s_in, esz, ctx->noise_k0));
            if (!split) {
                RCD(upload_frames(p, dc, desc, frames_host, s_in));
                CKD(cudaEventRecord(ctx->ev_in[slot], s_in));
            } else {
                // frames [0, f1] of every utterance, then the rest: two strided copies
                const size_t row = f32_frames ? 64 : 128;
                const size_t pitch = (size_t)p.uniform_frames * row, first = (size_t)(p.split_frame + 1) * row;
                const unsigned char *src = (const unsigned char *)frames_host + (size_t)p.frames_lo * row;
                unsigned char *dst = (unsigned char *)dc.frames;
                CKD(cudaMemcpy2DAsync(dst, pitch, src, pitch, first, (size_t)(u1 - u0), cudaMemcpyDefault, s_in));
                CKD(cudaEventRecord(ctx->ev_in[slot], s_in));
                CKD(cudaMemcpy2DAsync(dst + first, pitch, src + first, pitch, pitch - first, (size_t)(u1 - u0), cudaMemcpyDefault, s_in));
                CKD(cudaEventRecord(ctx->ev_in2[slot], s_in));
            }
            mark(s_in);
        }
        // ---- kernels -----------------------------------------------------------------------------------------
        {
            std::lock_guard<std::mutex> lk(g_run_mu[ctx->device]);
            CKD(cudaStreamWaitEvent(s_run, ctx->ev_in[slot], 0));
            auto waveguide = [&]() -> int {
                if (!split) { ++n_launch; return launch_stage(ctx, precision, TRM_STAGE_TUBE, dc, s_run); }
                int r = launch_stage(ctx, precision, TRM_STAGE_TUBE, dc, s_run, nullptr, 0);
                if (r != 0) return r;
                CKD(cudaStreamWaitEvent(s_run, ctx->ev_in2[slot], 0));
                n_launch += 2;
                return launch_stage(ctx, precision, TRM_STAGE_TUBE, dc, s_run, nullptr, 1);
            };
            if (!grouped) {
                for (int st = 0; st < TRM_STAGE_COUNT; ++st) {
                    if (st == TRM_STAGE_PCM && !want_pcm) { mark(s_run); continue; }
                    if (st == TRM_STAGE_TUBE) rc = waveguide();
                    else { rc = launch_stage(ctx, precision, st, dc, s_run); ++n_launch; }
                    if (rc != 0) { drain(); return rc; }
                    mark(s_run);
                }
            } else {
                // the waveguide over the whole chunk (it needs all of it to fill the device), then resampling + scaling
                // group by group: group g's PCM crosses PCIe while group g+1 is resampled
                RCD(waveguide());
                mark(s_run);
                for (size_t g = 0; g < p.groups.size(); ++g) {
                    RCD(launch_stage(ctx, precision, TRM_STAGE_SRC, dc, s_run, &p.groups[g]));
                    if (g + 1 == p.groups.size()) mark(s_run);
                    RCD(launch_stage(ctx, precision, TRM_STAGE_PCM, dc, s_run, &p.groups[g]));
                    CKD(cudaEventRecord(ctx->ev_grp[slot][g], s_run));
                    n_launch += 2;
                }
                mark(s_run);
            }
            CKD(cudaEventRecord(ctx->ev_run[slot], s_run));
        }
        // ---- copy-out ----------------------------------------------------------------------------------------
        if (grouped) {
            for (size_t g = 0; g < p.groups.size(); ++g) {
                const ChunkPlan::Group &grp = p.groups[g];
                CKD(cudaStreamWaitEvent(s_out, ctx->ev_grp[slot][g], 0));
                if (grp.pcm_hi > grp.pcm_lo)
                    CKD(cudaMemcpyAsync(pcm_host + p.pcm_lo + grp.pcm_lo, dc.pcm + grp.pcm_lo, (size_t)(grp.pcm_hi - grp.pcm_lo) * sizeof(int16_t),
                                       cudaMemcpyDeviceToHost, s_out));
            }
        }
        CKD(cudaStreamWaitEvent(s_out, ctx->ev_run[slot], 0));
        if (!grouped && want_pcm && p.pcm_elems) {
            long long c_hi = 0;
            for (const auto &d : p.desc) c_hi = std::max<long long>(c_hi, d.pcm_offset + d.n_out * d.channels);
            CKD(cudaMemcpyAsync(pcm_host + p.pcm_lo, dc.pcm, (size_t)c_hi * sizeof(int16_t), cudaMemcpyDeviceToHost, s_out));
        }
        if (samples_host && p.out_elems) {
            long long o_hi = 0;
            for (const auto &d : p.desc) o_hi = std::max<long long>(o_hi, d.out_offset + d.n_out);
            CKD(cudaMemcpyAsync((unsigned char *)samples_host + (size_t)p.out_lo * esz, dc.out, (size_t)o_hi * esz, cudaMemcpyDeviceToHost, s_out));
        }
        if (tube_host && p.tube_elems) {
            long long t_hi = 0;
            for (const auto &d : p.desc) t_hi = std::max<long long>(t_hi, d.tube_offset + d.n_tube);
            CKD(cudaMemcpyAsync((unsigned char *)tube_host + (size_t)p.tube_lo * esz, dc.tube, (size_t)t_hi * esz, cudaMemcpyDeviceToHost, s_out));
        }
        if (max_host) {
            unsigned char *mb = ctx->stages[slot].base + p.stage_bytes() - align_up(p.desc.size() * sizeof(unsigned long long), 256);
            CKD(cudaMemcpyAsync(mb, dc.maxbits, p.desc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_out));
        }
        CKD(cudaEventRecord(ctx->ev_out[slot], s_out));
        mark(s_out);
        slot_chunk[slot] = ci;
        if (ci == 0 && enqueued) enqueued(enqueued_arg);
    }
    const double t_enqueued = trace ? now_ms() - t_host0 : 0.0;
    for (int slot = 0; slot < N_SLOTS; ++slot) {
        int rc;
        if (slot_chunk[slot] >= 0 && (rc = finish_slot(slot)) != 0) return rc;
    }
    if (trace) {
        std::lock_guard<std::mutex> lk(t_mutex);
        fprintf(stderr, "[trm trace] ctx %p chunk: start h2d_done tube_done src_done pcm_done d2h_done (ms since the first traced call)\n", (void *)ctx);
        for (size_t i = 0; i + 6 <= tev.size(); i += 6) {
            float t[6];
            for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], t_origin, tev[i + k]);
            fprintf(stderr, "[trm trace] %2zu: %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f\n", i / 6, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        fprintf(stderr, "[trm trace] host: entered %.2f, everything enqueued %.2f, outputs complete %.2f\n", t_enter, t_enqueued, now_ms() - t_host0);
        for (auto e : tev) cudaEventDestroy(e);
    }
    if (launches) *launches = n_launch;
    return 0;
#undef CKD
#undef RCD
}

int trm_cuda_generate_frames(trm_cuda_ctx *ctx, int n, const trm_cuda_utterance *desc, const trm_cuda_event *events,
                             const int64_t *ev_offset, const int32_t *ev_count, const trm_cuda_framegen *fg, int shared_fg,
                             const double **frames_dev, double *frames_host, float *seed_host)
{
    if (frames_dev) *frames_dev = nullptr;
    if (n <= 0) return 0;
    CK(cudaSetDevice(ctx->device));
    long long n_events = 0, frames_hi = 0;
    for (int u = 0; u < n; ++u) {
        n_events = std::max<long long>(n_events, ev_offset[u] + ev_count[u]);
        frames_hi = std::max<long long>(frames_hi, desc[u].frame_offset + desc[u].n_frames);
    }
    const size_t n_fg = shared_fg ? 1 : (size_t)n;
    std::vector<long long> off(ev_offset, ev_offset + n);
    size_t bytes = 0;
    auto add = [&](size_t x) { bytes = align_up(bytes, 256) + x; };
    add((size_t)n * sizeof(trm_cuda_utterance)); add((size_t)n_events * sizeof(trm_cuda_event)); add((size_t)n * sizeof(long long));
    add((size_t)n * sizeof(int)); add(n_fg * sizeof(trm_cuda_framegen)); add((size_t)n * sizeof(float)); add((size_t)frames_hi * 128);
    int rc;
    if ((rc = ctx->gen.reserve(bytes + 256)) != 0) return rc;
    ctx->gen.reset();
    trm::FrameGenArgs a{};
    auto *d_desc = (trm_cuda_utterance *)ctx->gen.take((size_t)n * sizeof(trm_cuda_utterance));
    auto *d_ev = (trm_cuda_event *)ctx->gen.take((size_t)n_events * sizeof(trm_cuda_event));
    auto *d_off = (long long *)ctx->gen.take((size_t)n * sizeof(long long));
    auto *d_cnt = (int *)ctx->gen.take((size_t)n * sizeof(int));
    auto *d_fg = (trm_cuda_framegen *)ctx->gen.take(n_fg * sizeof(trm_cuda_framegen));
    auto *d_seed = (float *)ctx->gen.take((size_t)n * sizeof(float));
    auto *d_frames = (double *)ctx->gen.take((size_t)frames_hi * 128);
    cudaStream_t s = ctx->streams[2];
    CK(cudaMemcpyAsync(d_desc, desc, (size_t)n * sizeof(trm_cuda_utterance), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ev, events, (size_t)n_events * sizeof(trm_cuda_event), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_off, off.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_cnt, ev_count, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_fg, fg, n_fg * sizeof(trm_cuda_framegen), cudaMemcpyHostToDevice, s));
    a.desc = d_desc; a.n_utt = n; a.events = d_ev; a.ev_offset = d_off; a.ev_count = d_cnt; a.fg = d_fg; a.shared_fg = shared_fg;
    a.frames = d_frames; a.seed_out = d_seed;
    if ((rc = trm_k_framegen(&a, s)) != 0) return fail("frame generator launch", (cudaError_t)rc);
    if (frames_host) CK(cudaMemcpyAsync(frames_host, d_frames, (size_t)frames_hi * 128, cudaMemcpyDeviceToHost, s));
    if (seed_host) CK(cudaMemcpyAsync(seed_host, d_seed, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (frames_dev) *frames_dev = d_frames;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * streaming
 * ---------------------------------------------------------------------------------------------------------- */
struct trm_cuda_stream {
    trm_cuda_ctx *ctx = nullptr;
    int precision = 0, n = 0, max_m = 0;
    size_t esz = 8;
    trm_cuda_utterance voice{};
    int cp = 0;
    long long frames_seen = 0;     // control frames pushed so far (per stream)
    long long s_done = 0;          // tube-rate samples produced so far (multiple of 16 until the flush)
    long long out_done = 0;        // output-rate samples returned so far
    long long in_start = 0;        // first tube-rate sample still in the device buffer (multiple of 4)
    bool flushed = false;
    std::vector<double> pending;   // [stream][pend_frames][16]: frames from the control interval in progress on
    long long pend_first = 0;      // index of the first pending frame
    int pend_frames = 0;
    long long cap_tube = 0, cap_out = 0;     // elements per stream
    unsigned char *d_tube[2] = {nullptr, nullptr};
    int cur = 0;
    unsigned char *d_out = nullptr, *d_state = nullptr;
    double *d_frames = nullptr;
    double *d_wavetables = nullptr;    // own copy: the stream does not hold a context lane between pushes
    Arena scratch;                 // descriptors + resampler plan of a push
    HostStage stage;
    cudaStream_t st = nullptr;
};

static long long src_safe_outputs(long long n_in, unsigned tri)
{
    // outputs whose right filter wing lies inside the first n_in input samples: P(n) = (n*tri)>>16 <= n_in-1
    if (n_in <= 0) return 0;
    return (long long)(((unsigned long long)n_in * 65536ull + tri - 1) / tri);
}

int trm_cuda_stream_create(trm_cuda_ctx *ctx, int precision, int n_streams, const trm_cuda_utterance *voice,
                           int max_frames_per_push, trm_cuda_stream **out)
{
    *out = nullptr;
    if (n_streams <= 0 || max_frames_per_push <= 0) return fail_msg("trm_cuda_stream_create: bad sizes");
    CK(cudaSetDevice(ctx->device));
    trm_cuda_stream *s = new trm_cuda_stream();
    s->ctx = ctx; s->precision = precision; s->n = n_streams; s->max_m = max_frames_per_push;
    if (precision < 0 || precision > 2) return fail_msg("unknown precision code");
    s->esz = prec_esz(precision);
    s->voice = *voice;
    s->cp = voice->controlPeriod;
    const long long max_new = (long long)(max_frames_per_push + 1) * s->cp + 32;
    s->cap_tube = (long long)align_up((size_t)(max_new + 128), 32);
    s->cap_out = (long long)align_up((size_t)(src_safe_outputs(max_new + 64, voice->tri) + 64), 32);
    const size_t st_bytes = trm::tube_state_bytes(s->esz);
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaMalloc((void **)&s->d_tube[k], (size_t)n_streams * s->cap_tube * s->esz);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_out, (size_t)n_streams * s->cap_out * s->esz);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_state, (size_t)n_streams * st_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_frames, (size_t)n_streams * (max_frames_per_push + 3) * 128);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    if (e == cudaSuccess && !ctx->wt_host.empty()) {
        e = cudaMalloc((void **)&s->d_wavetables, ctx->wt_host.size() * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(s->d_wavetables, ctx->wt_host.data(), ctx->wt_host.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        // fresh state: everything zero, "no sample yet" flag set, noise generator at its start
        std::vector<unsigned char> init((size_t)n_streams * st_bytes, 0);
        for (int u = 0; u < n_streams; ++u) {
            unsigned long long *h = (unsigned long long *)(init.data() + (size_t)u * st_bytes);
            h[2] = ctx->noise_k0;
            h[3] = 1ull;
        }
        e = cudaMemcpy(s->d_state, init.data(), init.size(), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { trm_cuda_stream_destroy(s); return fail("trm_cuda_stream_create", e); }
    *out = s;
    return 0;
}

int64_t trm_cuda_stream_capacity(const trm_cuda_stream *s) { return s->cap_out; }

void trm_cuda_stream_destroy(trm_cuda_stream *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->st) { cudaStreamSynchronize(s->st); cudaStreamDestroy(s->st); }
    for (auto &p : s->d_tube) if (p) cudaFree(p);
    if (s->d_out) cudaFree(s->d_out);
    if (s->d_state) cudaFree(s->d_state);
    if (s->d_frames) cudaFree(s->d_frames);
    if (s->d_wavetables) cudaFree(s->d_wavetables);
    s->scratch.release();
    s->stage.release();
    delete s;
}

int trm_cuda_stream_push(trm_cuda_stream *s, const double *frames_host, int m, int flush, void *samples_host, int64_t *n_samples)
{
    if (n_samples) *n_samples = 0;
    if (s->flushed) return fail_msg("trm_cuda_stream_push: the streams were flushed");
    if (m < 0 || m > s->max_m) return fail_msg("trm_cuda_stream_push: more frames than the stream was created for");
    trm_cuda_ctx *ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    const int n = s->n, cp = s->cp;
    // ---- frames: keep everything from the control interval in progress on ------------------------------------
    {
        const int keep = s->pend_frames;
        std::vector<double> next((size_t)n * (keep + m) * 16);
        for (int u = 0; u < n; ++u) {
            if (keep) memcpy(&next[(size_t)u * (keep + m) * 16], &s->pending[(size_t)u * keep * 16], (size_t)keep * 128);
            if (m) memcpy(&next[((size_t)u * (keep + m) + keep) * 16], frames_host + (size_t)u * m * 16, (size_t)m * 128);
        }
        s->pending.swap(next);
        s->pend_frames = keep + m;
        s->frames_seen += m;
    }
    const long long avail = s->frames_seen >= 2 ? (s->frames_seen - 1) * cp : 0;      // tube samples the frames define
    const long long target = flush ? avail : avail / trm::TB * trm::TB;               // whole blocks until the end
    const long long n_new = target - s->s_done;
    const unsigned tri = s->voice.tri;
    const int pad = s->voice.padSize;
    const long long out_total = flush ? (long long)(((unsigned long long)(target + 2 * pad) * 65536ull + tri - 1) / tri)
                                      : src_safe_outputs(target, tri);
    if (flush && s->frames_seen < 1) { s->flushed = true; return 0; }
    if (n_new <= 0 && !(flush && out_total > s->out_done)) { if (flush) s->flushed = true; return 0; }
    cudaStream_t st = s->st;
    const bool f64 = prec_is_f64(s->precision);
    const KernelSet &K = g_kernels[s->precision];
    const trm::KernelInfo &ki = ctx->ki(s->precision);
    // ---- waveguide: samples [s_done, target) -----------------------------------------------------------------
    const long long f0 = s->s_done / cp;                          // control interval the call starts in
    const int jc0 = (int)(s->s_done % cp);
    const int skip = (int)(f0 - s->pend_first);                   // pending frames before it are no longer needed
    const int nf_call = s->pend_frames - skip;
    std::vector<trm_cuda_utterance> dt(n), ds(n);
    for (int u = 0; u < n; ++u) {
        trm_cuda_utterance d = s->voice;
        d.frame_offset = (long long)u * nf_call;
        d.n_frames = nf_call;
        d.n_tube = n_new > 0 ? n_new : 0;
        d.tube_offset = (long long)u * s->cap_tube + (s->s_done - s->in_start);
        d.jc0 = jc0;
        d.out_start = 0; d.in_start = 0;
        dt[u] = d;
        trm_cuda_utterance r = s->voice;                          // resampler view: global sample indices
        r.n_tube = target;
        r.tube_offset = (long long)u * s->cap_tube - s->in_start;
        r.n_out = out_total;
        r.out_start = s->out_done;
        r.in_start = s->in_start;
        r.out_offset = (long long)u * s->cap_out - (s->out_done / 4 * 4);
        r.pcm_offset = 0;
        ds[u] = r;
    }
    if (target - s->in_start > s->cap_tube || out_total - s->out_done / 4 * 4 > s->cap_out) return fail_msg("trm_cuda_stream_push: internal capacity");
    ChunkPlan plan;
    if (plan_chunk(ds.data(), 0, n, ki, plan) != 0) return -1;
    // plan_chunk rebases offsets to the chunk's span: streaming keeps its own (virtual) offsets
    for (int u = 0; u < n; ++u) { plan.desc[u].tube_offset = ds[u].tube_offset; plan.desc[u].out_offset = ds[u].out_offset; plan.desc[u].frame_offset = 0; }
    size_t bytes = plan.stage_bytes() + align_up((size_t)n * sizeof(trm_cuda_utterance), 256) + 1024;
    int rc;
    if ((rc = s->scratch.reserve(bytes + 4096)) != 0) return rc;
    if ((rc = s->stage.reserve(bytes + (size_t)n * nf_call * 128 + 4096)) != 0) return rc;
    s->scratch.reset();
    DeviceChunk dc;
    dc.n = n;
    dc.desc = (trm_cuda_utterance *)s->scratch.take((size_t)n * sizeof(trm_cuda_utterance));
    dc.order = nullptr;
    dc.tile_utt = (int *)s->scratch.take(plan.tile_utt.size() * sizeof(int));
    dc.tile_nt = (int *)s->scratch.take(plan.tile_nt.size() * sizeof(int));
    dc.tile_max_out = (long long *)s->scratch.take(plan.tile_max_out.size() * sizeof(long long));
    dc.tile_first_out = (long long *)s->scratch.take(plan.tile_first_out.size() * sizeof(long long));
    dc.item_base = (long long *)s->scratch.take(plan.item_base.size() * sizeof(long long));
    dc.maxbits = (unsigned long long *)s->scratch.take((size_t)n * sizeof(unsigned long long));
    auto *d_dt = (trm_cuda_utterance *)s->scratch.take((size_t)n * sizeof(trm_cuda_utterance));
    dc.total_items = plan.total_items; dc.n_tiles = (int)plan.n_tiles(); dc.max_n_out = plan.max_n_out; dc.src_shape = plan.src_shape;
    {
        unsigned char *q = s->stage.base;
        auto put = [&](void *dst, const void *src, size_t b) -> int {
            if (!b) return 0;
            memcpy(q, src, b);
            CK(cudaMemcpyAsync(dst, q, b, cudaMemcpyHostToDevice, st));
            q += align_up(b, 256);
            return 0;
        };
        if ((rc = put(dc.desc, plan.desc.data(), (size_t)n * sizeof(trm_cuda_utterance))) != 0) return rc;
        if ((rc = put(dc.tile_utt, plan.tile_utt.data(), plan.tile_utt.size() * sizeof(int))) != 0) return rc;
        if ((rc = put(dc.tile_nt, plan.tile_nt.data(), plan.tile_nt.size() * sizeof(int))) != 0) return rc;
        if ((rc = put(dc.tile_max_out, plan.tile_max_out.data(), plan.tile_max_out.size() * sizeof(long long))) != 0) return rc;
        if ((rc = put(dc.tile_first_out, plan.tile_first_out.data(), plan.tile_first_out.size() * sizeof(long long))) != 0) return rc;
        if ((rc = put(dc.item_base, plan.item_base.data(), plan.item_base.size() * sizeof(long long))) != 0) return rc;
        if ((rc = put(d_dt, dt.data(), (size_t)n * sizeof(trm_cuda_utterance))) != 0) return rc;
        if (n_new > 0) {
            // frames of this call: pending[skip ..) of every stream, back to back
            double *fq = (double *)q;
            for (int u = 0; u < n; ++u)
                memcpy(fq + (size_t)u * nf_call * 16, &s->pending[((size_t)u * s->pend_frames + skip) * 16], (size_t)nf_call * 128);
            CK(cudaMemcpyAsync(s->d_frames, fq, (size_t)n * nf_call * 128, cudaMemcpyHostToDevice, st));
        }
    }
    if (n_new > 0) {
        trm::TubeArgs a{};
        a.state = s->d_state; a.desc = d_dt; a.order = nullptr; a.n_utt = n; a.frames = s->d_frames; a.tube = s->d_tube[s->cur];
        a.wavetables = s->d_wavetables ? s->d_wavetables : ctx->d_wavetables; a.noise_k0 = ctx->noise_k0;
        const int gmax = ki.wide_max_utt;
        const int groups = std::max(1, std::min(ctx->sm_count, (n + 1) / 2));
        const int g2 = (n + groups - 1) / groups > gmax ? (n + gmax - 1) / gmax : groups;
        rc = K.tube_wide(&a, g2, st);
        if (rc != 0) return fail("stream waveguide launch", (cudaError_t)rc);
    }
    // ---- resampler: outputs [out_done, out_total) ------------------------------------------------------------
    const long long n_ret = out_total - s->out_done;
    if (n_ret > 0) {
        trm::SrcArgs a{};
        a.desc = dc.desc; a.n_utt = n; a.tube = s->d_tube[s->cur]; a.out = s->d_out; a.maxbits = dc.maxbits;
        a.table = f64 ? ctx->d_tab_f64 : ctx->d_tab_f32;
        a.ctab = f64 ? ctx->d_ctab_f64 : ctx->d_ctab_f32;
        a.tile_utt = dc.tile_utt; a.tile_nt = dc.tile_nt; a.tile_max_out = dc.tile_max_out; a.tile_first_out = dc.tile_first_out;
        a.item_base = dc.item_base; a.n_tiles = dc.n_tiles; a.total_items = dc.total_items;
        const int grid = ctx->sm_count * std::max(1, ki.src[dc.src_shape].ctas_per_sm);
        rc = K.src(&a, grid, dc.src_shape, st);
        if (rc != 0) return fail("stream resampler launch", (cudaError_t)rc);
        if (samples_host)
            CK(cudaMemcpy2DAsync(samples_host, (size_t)s->cap_out * s->esz, s->d_out + (size_t)(s->out_done % 4) * s->esz,
                                 (size_t)s->cap_out * s->esz, (size_t)n_ret * s->esz, (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    // ---- keep what the next call's filter wings reach back to: x[P(next) - 2*pad - 1 ..] -----------------------
    long long next_in = flush ? target : (long long)(((unsigned long long)out_total * tri) >> 16) - 2 * pad - 4;
    if (next_in < 0) next_in = 0;
    next_in = next_in / 4 * 4;
    if (next_in < s->in_start) next_in = s->in_start;
    if (!flush) {
        const long long keep = target - next_in;
        CK(cudaMemcpy2DAsync(s->d_tube[s->cur ^ 1], (size_t)s->cap_tube * s->esz,
                             s->d_tube[s->cur] + (size_t)(next_in - s->in_start) * s->esz, (size_t)s->cap_tube * s->esz,
                             (size_t)keep * s->esz, (size_t)n, cudaMemcpyDeviceToDevice, st));
        s->cur ^= 1;
    }
    CK(cudaStreamSynchronize(st));
    s->in_start = next_in;
    s->s_done = target;
    s->out_done = out_total;
    // frames before the interval the next call starts in are done with
    {
        const long long nf0 = s->s_done / cp;
        const int drop = (int)(nf0 - s->pend_first);
        if (drop > 0) {
            const int keep = s->pend_frames - drop;
            std::vector<double> next((size_t)n * keep * 16);
            for (int u = 0; u < n; ++u)
                memcpy(&next[(size_t)u * keep * 16], &s->pending[((size_t)u * s->pend_frames + drop) * 16], (size_t)keep * 128);
            s->pending.swap(next);
            s->pend_frames = keep;
            s->pend_first = nf0;
        }
    }
    if (flush) s->flushed = true;
    if (n_samples) *n_samples = n_ret > 0 ? n_ret : 0;
    return 0;
}

int trm_cuda_resident_create(trm_cuda_ctx *ctx, int precision, int n, const trm_cuda_utterance *desc,
                             const double *frames_host, trm_cuda_resident **out)
{
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    if (precision < 0 || precision > 2) return fail_msg("unknown precision code");
    const size_t esz = prec_esz(precision);
    trm_cuda_resident *r = new trm_cuda_resident();
    r->ctx = ctx;
    r->precision = precision;
    int rc;
    if ((rc = plan_chunk(desc, 0, n, ctx->ki(precision), r->plan)) != 0) { delete r; return rc; }
    if ((rc = r->arena.reserve(r->plan.arena_bytes(esz, true))) != 0) { delete r; return rc; }
    carve(r->arena, r->plan, esz, true, r->dc);
    if ((rc = upload_plan(r->plan, r->dc, nullptr, 0)) != 0 || (rc = upload_frames(r->plan, r->dc, desc, frames_host, 0)) != 0) {
        r->arena.release();
        delete r;
        return rc;
    }
    CK(cudaMemset(r->dc.maxbits, 0, (size_t)n * sizeof(unsigned long long)));
    if (!ctx->wt_host.empty()) {
        if (cudaMalloc((void **)&r->d_wavetables, ctx->wt_host.size() * sizeof(double)) != cudaSuccess ||
            cudaMemcpy(r->d_wavetables, ctx->wt_host.data(), ctx->wt_host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            if (r->d_wavetables) cudaFree(r->d_wavetables);
            r->arena.release();
            delete r;
            return fail("resident wavetables", cudaGetLastError());
        }
        r->dc.wavetables = r->d_wavetables;
    }
    CK(cudaDeviceSynchronize());
    *out = r;
    return 0;
}

void trm_cuda_resident_destroy(trm_cuda_resident *r)
{
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaDeviceSynchronize();
    r->arena.release();
    if (r->d_wavetables) cudaFree(r->d_wavetables);
    delete r;
}

int trm_cuda_resident_stage(trm_cuda_resident *r, int stage, void *stream)
{
    CK(cudaSetDevice(r->ctx->device));
    return launch_stage(r->ctx, r->precision, stage, r->dc, (cudaStream_t)stream);
}

int trm_cuda_resident_run(trm_cuda_resident *r, void *stream)
{
    for (int st = 0; st < TRM_STAGE_COUNT; ++st) {
        int rc = trm_cuda_resident_stage(r, st, stream);
        if (rc) return rc;
    }
    return 0;
}

int trm_cuda_resident_fetch(trm_cuda_resident *r, int16_t *pcm_host, void *samples_host, double *max_host, void *tube_host)
{
    CK(cudaSetDevice(r->ctx->device));
    CK(cudaDeviceSynchronize());
    const size_t esz = prec_esz(r->precision);
    const ChunkPlan &p = r->plan;
    long long c_hi = 0, o_hi = 0, t_hi = 0;
    for (const auto &d : p.desc) {
        c_hi = std::max<long long>(c_hi, d.pcm_offset + d.n_out * d.channels);
        o_hi = std::max<long long>(o_hi, d.out_offset + d.n_out);
        t_hi = std::max<long long>(t_hi, d.tube_offset + d.n_tube);
    }
    if (pcm_host && c_hi) CK(cudaMemcpy(pcm_host + p.pcm_lo, r->dc.pcm, (size_t)c_hi * sizeof(int16_t), cudaMemcpyDeviceToHost));
    if (samples_host && o_hi) CK(cudaMemcpy((unsigned char *)samples_host + (size_t)p.out_lo * esz, r->dc.out, (size_t)o_hi * esz, cudaMemcpyDeviceToHost));
    if (tube_host && t_hi) CK(cudaMemcpy((unsigned char *)tube_host + (size_t)p.tube_lo * esz, r->dc.tube, (size_t)t_hi * esz, cudaMemcpyDeviceToHost));
    if (max_host && !p.desc.empty()) {
        std::vector<unsigned long long> mb(p.desc.size());
        CK(cudaMemcpy(mb.data(), r->dc.maxbits, mb.size() * sizeof(mb[0]), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < mb.size(); ++i) memcpy(&max_host[i], &mb[i], sizeof(double));
    }
    return 0;
}

/*
 * Sweep (BASELINE configs[4]): n utterances of one voice and n_frames frames each, whose control
 * tracks are the walk2 workload (include/trm_workload.h) generated ON THE DEVICE -- utterance k uses index first_index + k
 * of stream `seed` -- synthesized chunk by chunk; the audio never leaves the device: 8 bytes of PCM checksum (and the
 * maximum) per utterance come back.  probe_utt[] (sorted, relative to this call) names utterances whose PCM is copied to
 * probe_pcm (n_probe rows of probe_stride int16) for oracle spot checks.  One stream, no copies inside the loop.
 */
int trm_cuda_sweep(trm_cuda_ctx *ctx, int precision, const trm_cuda_utterance *voice, int32_t n_frames, uint64_t seed,
                   uint64_t first_index, int64_t n, uint64_t *checksums_host, double *max_host, int64_t n_probe,
                   const int64_t *probe_utt, int16_t *probe_pcm, int64_t probe_stride, int64_t *launches, double *kernel_ms)
{
    if (launches) *launches = 0;
    if (kernel_ms) *kernel_ms = 0.0;
    if (n <= 0) return 0;
    if (precision < 0 || precision > 2) return fail_msg("unknown precision code");
    CK(cudaSetDevice(ctx->device));
    const size_t esz = prec_esz(precision);
    const trm::KernelInfo &ki = ctx->ki(precision);
    const int64_t cap = (int64_t)ctx->sm_count * ki.wide_max_utt;          // one full wave of the waveguide kernel
    const int C = (int)std::min<int64_t>(n, cap);
    // one chunk plan, reused: C utterances of the template voice, frames back to back
    std::vector<trm_cuda_utterance> desc((size_t)C, *voice);
    long long tube_at = 0, out_at = 0, pcm_at = 0;
    auto up = [](long long v) { return (v + TRM_ALIGN_ELEMS - 1) / TRM_ALIGN_ELEMS * TRM_ALIGN_ELEMS; };
    for (int u = 0; u < C; ++u) {
        trm_cuda_utterance &d = desc[u];
        d.n_frames = n_frames;
        d.frame_offset = (long long)u * n_frames;
        d.tube_offset = tube_at; d.out_offset = out_at; d.pcm_offset = pcm_at;
        tube_at += up(d.n_tube); out_at += up(d.n_out); pcm_at += up(d.n_out * d.channels);
    }
    ChunkPlan plan;
    int rc;
    if ((rc = plan_chunk(desc.data(), 0, C, ki, plan)) != 0) return rc;
    Arena &arena = ctx->arenas[0];
    if ((rc = arena.reserve(plan.arena_bytes(esz, true) + (size_t)C * sizeof(unsigned long long) + 4096)) != 0) return rc;
    DeviceChunk dc;
    carve(arena, plan, esz, true, dc);
    unsigned long long *d_sums = (unsigned long long *)arena.take((size_t)C * sizeof(unsigned long long));
    cudaStream_t st = ctx->streams[2];
    if ((rc = upload_plan(plan, dc, nullptr, st)) != 0) return rc;
    std::vector<unsigned long long> mb((size_t)C);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int64_t n_launch = 0, probe_at = 0;
    double ms_total = 0.0;
    for (int64_t at = 0; at < n; at += C) {
        const int m = (int)std::min<int64_t>(C, n - at);
        // (a short last chunk runs the full plan: the utterances beyond m keep the previous chunk's frames and are ignored)
        CK(cudaEventRecord(e0, st));
        if (trm_k_workload_walk2(seed, first_index + (uint64_t)at, m, n_frames, dc.frames, st) != 0) return fail_msg("walk2 launch");
        for (int stage = 0; stage < TRM_STAGE_COUNT; ++stage)
            if ((rc = launch_stage(ctx, precision, stage, dc, st)) != 0) return rc;
        if (trm_k_pcm_checksum(dc.desc, m, dc.pcm, d_sums, st) != 0) return fail_msg("checksum launch");
        CK(cudaEventRecord(e1, st));
        n_launch += 5;
        CK(cudaMemcpyAsync(checksums_host + at, d_sums, (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(mb.data(), dc.maxbits, (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        while (probe_at < n_probe && probe_utt[probe_at] < at + m) {
            const int64_t k = probe_utt[probe_at] - at;
            if (k >= 0) {
                const trm_cuda_utterance &d = plan.desc[(size_t)k];
                const int64_t cnt = std::min<int64_t>(d.n_out * d.channels, probe_stride);
                CK(cudaMemcpyAsync(probe_pcm + probe_at * probe_stride, dc.pcm + d.pcm_offset, (size_t)cnt * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
            }
            ++probe_at;
        }
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms_total += ms;
        if (max_host)
            for (int i = 0; i < m; ++i) memcpy(&max_host[at + i], &mb[(size_t)i], sizeof(double));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (launches) *launches = n_launch;
    if (kernel_ms) *kernel_ms = ms_total;
    return 0;
}

/* Output-rate samples (and PCM) of ONE utterance of a device-resident batch: what bench.py's in-run oracle check reads. */
int trm_cuda_resident_fetch_utterance(trm_cuda_resident *r, int u, void *samples_host, int16_t *pcm_host, double *max_host)
{
    CK(cudaSetDevice(r->ctx->device));
    CK(cudaDeviceSynchronize());
    if (u < 0 || u >= (int)r->plan.desc.size()) return fail_msg("trm_cuda_resident_fetch_utterance: no such utterance");
    const size_t esz = prec_esz(r->precision);
    const trm_cuda_utterance &d = r->plan.desc[u];
    if (samples_host && d.n_out) CK(cudaMemcpy(samples_host, (unsigned char *)r->dc.out + (size_t)d.out_offset * esz, (size_t)d.n_out * esz, cudaMemcpyDeviceToHost));
    if (pcm_host && d.n_out) CK(cudaMemcpy(pcm_host, r->dc.pcm + d.pcm_offset, (size_t)d.n_out * d.channels * sizeof(int16_t), cudaMemcpyDeviceToHost));
    if (max_host) {
        unsigned long long mb = 0;
        CK(cudaMemcpy(&mb, r->dc.maxbits + u, sizeof mb, cudaMemcpyDeviceToHost));
        memcpy(max_host, &mb, sizeof(double));
    }
    return 0;
}

/* Pure copy traffic of one synthesis step, no kernels: `h2d_bytes` from pinned host memory up and `d2h_bytes` down, at the
 * same time on two streams, `reps` times; *ms receives the average time of one repetition.  The ceiling the end-to-end
 * number of a step with these byte counts can reach on this host / PCIe path (tools/pcie_ceiling.py, bench.py e2e.ceiling). */
int trm_cuda_copy_probe(int device, const void *host_in, size_t h2d_bytes, void *host_out, size_t d2h_bytes, int reps, double *ms)
{
    CK(cudaSetDevice(device));
    void *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_up, s_down;
    cudaEvent_t e0, e1, e2;
    CK(cudaMalloc(&d_in, h2d_bytes ? h2d_bytes : 1));
    CK(cudaMalloc(&d_out, d2h_bytes ? d2h_bytes : 1));
    CK(cudaMemset(d_out, 0, d2h_bytes ? d2h_bytes : 1));
    CK(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, s_up));
    CK(cudaStreamWaitEvent(s_down, e0, 0));
    for (int r = 0; r < reps; ++r) {
        if (h2d_bytes) CK(cudaMemcpyAsync(d_in, host_in, h2d_bytes, cudaMemcpyHostToDevice, s_up));
        if (d2h_bytes) CK(cudaMemcpyAsync(host_out, d_out, d2h_bytes, cudaMemcpyDeviceToHost, s_down));
    }
    CK(cudaEventRecord(e1, s_up));
    CK(cudaEventRecord(e2, s_down));
    CK(cudaStreamSynchronize(s_up));
    CK(cudaStreamSynchronize(s_down));
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, e0, e1));
    CK(cudaEventElapsedTime(&b, e0, e2));
    *ms = (double)(a > b ? a : b) / (reps > 0 ? reps : 1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    cudaStreamDestroy(s_up); cudaStreamDestroy(s_down);
    cudaFree(d_in); cudaFree(d_out);
    return 0;
}

}  // extern "C"
