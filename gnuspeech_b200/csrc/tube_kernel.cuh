// tube_kernel.cuh -- the TRM waveguide kernel for sm_100a.
//
// Replaces the sample-rate loop of -[TRMTubeModel synthesize]
// (/root/reference/Frameworks/Tube/TRMTubeModel.m:292-354) and everything it calls:
// parameter interpolation (m:611-688), frequency()/amplitude() (TRMUtility.m:26-47), tube and
// frication coefficients (m:712-773), band-pass coefficients/filter (TRMFilters.m:9-29), noise +
// one-zero low-pass (TRMUtility.m:71-85, TRMFilters.m:81-86), glottal wavetable + 2x oversampling
// oscillator + 49-tap FIR (TRMWavetable.m:117-195, TRMFIRFilter.m:116-146), source mixing (m:305-337),
// Kelly-Lochbaum ladder with the velum 3-way junction and nasal branch (m:778-853), mouth/nose
// reflection + radiation filters (TRMFilters.m:34-60) and the throat low-pass (TRMFilters.m:64-77).
//
// Mapping (BASELINE.json north_star): one utterance per HALF-WARP, two per warp.  The path has two
// kinds of work and each gets the lane mapping that suits it, alternating every 16 samples:
//
//   time-parallel phases  (lane = sample t of the block): everything that does not depend on the tube
//       state -- transcendentals, junction coefficients, taps, band-pass coefficients, jump-ahead
//       noise, table look-ups, FIR, source mixing.  Results go to shared memory.
//   section-parallel phase (lane = scattering junction): the strictly sequential ladder.  Lane j owns
//       the two waves incident on junction j (registers); after each sample the outgoing waves move to
//       the neighbouring junctions with __shfl_sync (3 shuffles: right-going, left-going, velum port).
//
//   lane: 0 S1|S2 (+glottis end)  1 S2|S3  2 S3|S4  3 S4|S5|N1 (3-way)  4 S5|S6  5 S6|S7 (k=0)
//         6 S7|S8  7 S8|S9  8 S9|S10  9 mouth  10 N1|N2  11 N2|N3  12 N3|N4  13 N4|N5  14 N5|N6  15 nose
//
// Control frames (128 B each) are staged into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier), double-buffered FRAME_CHUNK frames ahead of the sample loop.
// Tube-rate output is gathered in shared memory and stored with 128-bit vector stores.
//
// Real = double : FP64 conformance mode, compile this TU with -fmad=false (reference build has no FMA).
// Real = float  : FP32 fast mode: state/signal/coefficients FP32; parameter interpolation, pitch -> f0 ->
//                 table position, the glottal-closure decision rint(ax*tnDelta) and the noise MCG stay
//                 FP64 / integer.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <math.h>

#include "kernel_args.h"
#include "trm_cuda.h"

namespace TRM_KERNEL_NS {
using namespace trm;

constexpr double TWO_M44 = 5.684341886080801486968994140625e-14;   // 2^-44 exactly

// Per-TU constant tables (uploaded by the TU's upload function).
static __constant__ double c_fir_d[FIR_TAPS];
static __constant__ float c_fir_f[FIR_TAPS];
static __constant__ unsigned long long c_noise_pow[TRM_NOISE_JUMP + 1];

template <typename R> struct FirCoef;
template <> struct FirCoef<double> { static __device__ __forceinline__ double at(int i) { return c_fir_d[i]; } };
template <> struct FirCoef<float> { static __device__ __forceinline__ float at(int i) { return c_fir_f[i]; } };

template <typename R>
struct alignas(16) UttSmem {
    // control-frame staging (TMA destination), double-buffered
    double FR[2][FRAME_CHUNK][16];
    unsigned long long mbar[2];
    // P (interpolated parameters: written by the parameter lanes, read by phase A1) is dead once A1 is
    // done; vectors produced after A1 share its storage.
    union {
        double P[TB][17];
        struct {
            double POS[2 * TB];              // oscillator positions of the two 2x-rate steps of each sample
            R OUTM[TB], OUTN[TB];            // mouth / nose radiation outputs
            R TH[TB];                        // throat low-pass output
            R YB[TB];                        // finished tube-rate samples, for the vector store
            R SCR[TB];                       // scratch: where non-terminal lanes dump their (unused) radiation value
        } v;
    } a;
    double INC[TB];                          // oscillator increment (f0/2)*basicIncrement
    R BC[TB][4];                             // per-sample scalars every junction lane needs (broadcast reads):
                                             //   [0] 2*alpha*(x[n]-x[n-2])  [1] 2*gamma  [2] 2*beta  (band-pass)
                                             //   [3] ta0*(pulse*VT_SCALE)                           (throat)
    // conformance mode (R = double): reference-order arithmetic needs one coefficient per junction
    R KQ[sizeof(R) == 8 ? TB : 1][sizeof(R) == 8 ? 17 : 1];   // junction coefficient per lane (cols 0..15), alphaU (16)
    R TAPV[sizeof(R) == 8 ? TB : 1][sizeof(R) == 8 ? 9 : 1];  // col 0: glottal input; cols 1..8: taps FC1..FC8
    // fast mode (R = float): cancellation-free forms with the damping d folded in, one 128-bit load per lane.
    // Every lane evaluates   Rr = x*a + (+-y)*b + rc*c + inj*fr,   Lo = y*a + z*b + rc*c   on its own tuple {x,y,z,w}:
    //   two-port lanes : {d(1+k), d k, d(1-k), tap}     (lane 0: w = glottal input; constant lanes are pre-filled)
    //   3-way lane 3   : {d aL, d(aL-1), d aL, d(aU-1)}, rc = d aU and its tap come from Z3
    //   termination    : {0, d a10 k, -b11, 1+k}        (its b register carries the previous Lo)
    float4 KF[sizeof(R) == 4 ? TB : 1][sizeof(R) == 4 ? 17 : 1];
    float2 Z3[sizeof(R) == 4 ? TB : 1];      // {d aU, FC3 tap} for the 3-way lane
    R HE[FIR_HIST + TB], HO[FIR_HIST + TB];  // oscillator history, even / odd 2x-rate samples
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity)
{
    // try_wait suspends in hardware; the bound turns a lost copy into a trap instead of a hang
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// math helpers
// ---------------------------------------------------------------------------------------------
// amplitude(): TRMUtility.m:26-41.  Always evaluated in double (it is off the serial chain and the
// glottal-closure decision rint(ax*tnDelta) must not move).
// a / c for a divisor known in advance, rc = RN(1 / c): one multiply and two fused operations instead of the ~26
// instructions of a general IEEE division.  q' = RN(q + (a - c q) rc) with the residual exact is the correctly rounded
// quotient (Markstein's correction step; Brisebarre, Muller & Raina 2004): the value rounded last differs from a / c by
// <= 2^-105 |a / c|, closer than a quotient of two doubles comes to a rounding boundary except for isolated operand
// pairs.  Same result as `a / c`, so the conformance arithmetic is unchanged.
__device__ __forceinline__ double div_known(double a, double c, double rc)
{
    const double q = a * rc;
    const double r = fma(-c, q, a);
    return fma(r, rc, q);
}

__device__ __forceinline__ double amplitude_db(double dB)
{
    double x = dB - 60.0;
    if (x <= -60.0) return 0.0;
    if (x >= 0.0) return 1.0;
    return exp10(div_known(x, 20.0, 1.0 / 20.0));
}
// ---- helpers of the FP64 conformance mode's cheaper forms (tolerance 1e-9, BASELINE.json; none of them is used when
// TRM_STRICT = 1) -----------------------------------------------------------------------------------------------------
// 1 / s: hardware seed (MUFU.RCP64H, >= 16 good bits) refined by one cubic Newton step, x (1 + e + e^2) with
// e = 1 - s x: relative error <= 2^-48 in the worst case the seed allows, 5e-19 for its typical 2^-22.  No special
// cases: s = 0 gives inf * 0 = NaN in the correction, which is what the reference's 0 / 0 between two closed sections
// produces (TRMTubeModel.m:716-718) and must propagate.
__device__ __forceinline__ double rcp_fast(double s)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(s));
    const double e = fma(-s, x, 1.0);
    const double e2 = fma(e, e, e);
    return fma(x, e2, x);
}

// 2^x for |x| <= 1000: x = n + f, |f| <= 1/2, 2^f = exp(f ln 2) by its Taylor series to degree 13 (|f ln 2| <= 0.347:
// truncation 4e-18), exponent added as an integer.  Straight-line code; ~1 ulp.
__device__ __forceinline__ double exp2_inline(double x)
{
    x = fmin(fmax(x, -1000.0), 1000.0);
    const double n = rint(x);
    const double g = (x - n) * 0.693147180559945309417;
    double p = 1.6059043836821613e-10;                 // 1/13!
    p = fma(p, g, 2.08767569878681e-09);               // 1/12!
    p = fma(p, g, 2.505210838544172e-08);              // 1/11!
    p = fma(p, g, 2.755731922398589e-07);              // 1/10!
    p = fma(p, g, 2.7557319223985893e-06);             // 1/9!
    p = fma(p, g, 2.48015873015873e-05);               // 1/8!
    p = fma(p, g, 0.0001984126984126984);              // 1/7!
    p = fma(p, g, 0.001388888888888889);               // 1/6!
    p = fma(p, g, 0.008333333333333333);               // 1/5!
    p = fma(p, g, 0.041666666666666664);               // 1/4!
    p = fma(p, g, 0.16666666666666666);                // 1/3!
    p = fma(p, g, 0.5);
    p = fma(p, g, 1.0);
    p = fma(p, g, 1.0);
    return __hiloint2double(__double2hiint(p) + ((int)n << 20), __double2loint(p));
}

// sin and cos for |x| <= ~100 (the arguments here are below 2 pi): Cody-Waite reduction by pi/2 in two parts with
// fused multiply-adds, then the minimax kernels of fdlibm (k_sin.c / k_cos.c coefficients) on |r| <= pi/4; ~1 ulp.
__device__ __forceinline__ void sincos_inline(double x, double *sn, double *cs)
{
    const double n = rint(x * 0.63661977236758134308);
    double r = fma(-n, 1.5707963267948966, x);
    r = fma(-n, 6.123233995736766e-17, r);
    const int q = (int)n;
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sr = fma(r * z, ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
    *sn = (q & 2) ? -s0 : s0;
    *cs = ((q + 1) & 2) ? -c0 : c0;
}

// Bit-wise select (one LOP3 per 32 bits): m = all ones -> x, m = 0 -> y.  Used instead of ?: in the junction
// loop so that the per-lane roles stay straight-line code (the compiler turns lane-dependent ternaries into
// divergent branch regions with reconvergence barriers and a divergence check before every shuffle).
__device__ __forceinline__ float blend(unsigned m, float x, float y)
{
    return __uint_as_float((__float_as_uint(x) & m) | (__float_as_uint(y) & ~m));
}
__device__ __forceinline__ double blend(unsigned m, double x, double y)
{
    const unsigned lo = ((unsigned)__double2loint(x) & m) | ((unsigned)__double2loint(y) & ~m);
    const unsigned hi = ((unsigned)__double2hiint(x) & m) | ((unsigned)__double2hiint(y) & ~m);
    return __hiloint2double((int)hi, (int)lo);
}

// fast-mode amplitude(): 10^((dB-60)/20) as 2^(..), FP32 (relative error ~5e-7, i.e. -126 dB on a linear gain)
__device__ __forceinline__ float amplitude_f(float dB)
{
    const float x = dB - 60.0f;
    const float v = exp2f(fminf(x, 0.0f) * 0.16609640474436813f);    // x >= 0 -> exactly 1
    return (x <= -60.0f) ? 0.0f : v;
}

// fast-mode glottal table: rise 3x^2 - 2x^3 (x = i/div1), fall 1 - (j*j)/L^2, closed 0 -- evaluated, never loaded
// (TRMWavetable.m:78-96, 117-162).  The sine waveform (rare) still reads the 512-entry table.
__device__ __forceinline__ float table_value_fast(const double *__restrict__ base, int i, int div1, float inv_div1,
                                                  double newDiv2, float scale, bool pulse)
{
    if (!pulse) return (float)__ldg(base + i);
    const float x = (float)i * inv_div1;
    const float rise = (x * x) * (3.0f - 2.0f * x);
    const float j = (float)(i - div1);
    const float fall = 1.0f - ((j * j) * scale);
    const float v = (i < div1) ? rise : fall;
    return ((double)i >= newDiv2) ? 0.0f : v;
}

// Glottal table value at integer index i for the current closure point (TRMWavetable.m:78-102 init,
// :117-162 update, vDSP order 1 - (j*j)*(1/(L*L))).  The reference rewrites the table every sample; the
// table is a pure function of the current amplitude, so it is evaluated on look-up instead.
template <typename R>
__device__ __forceinline__ R table_value(const double *__restrict__ base, int i, int div1, int div2, double newDiv2,
                                         R scale, bool pulse)
{
    if (!pulse || i < div1 || i >= div2) return (R)__ldg(base + i);
    if ((double)i >= newDiv2) return (R)0;
    R j = (R)(i - div1);
    return (R)1 - ((j * j) * scale);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xFFFFFFFFu;

template <typename R>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 7) tube_kernel(TubeArgs args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, hl = lane & 15;
    const int slot = (blockIdx.x * WARPS_PER_CTA + warp) * 2 + half;
    UttSmem<R> &S = reinterpret_cast<UttSmem<R> *>(smem_raw)[warp * 2 + half];

    // A warp always runs both halves in lock step (full-mask shuffles and syncs).  A half without an
    // utterance (odd batch) or whose utterance is shorter than its partner's keeps executing with
    // n_tube = 0 / past its end: it computes on stale shared memory and never touches global memory.
    const bool has_utt = slot < args.n_utt;
    const int u = args.order ? args.order[has_utt ? slot : args.n_utt - 1] : (has_utt ? slot : args.n_utt - 1);
    const trm_cuda_utterance *__restrict__ D = args.desc + u;
    const int64_t n_tube = has_utt ? D->n_tube : 0;
    int64_t n_warp = n_tube;
    {
        const int64_t other = __shfl_xor_sync(FULL, n_tube, 16);
        n_warp = other > n_tube ? other : n_tube;
    }
    if (n_warp <= 0) return;
    const int n_frames = D->n_frames;
    const int cp = D->controlPeriod;
    // control frames: rows of 16 doubles (128 bytes) or, when the caller gave float32 rows, of 16 floats (64 bytes) --
    // staged as they are and widened when a parameter lane reads its value (exact)
    const uint32_t frow = args.frames_f32 ? 64u : 128u;
    const unsigned char *__restrict__ F = reinterpret_cast<const unsigned char *>(args.frames) + (size_t)D->frame_offset * frow;
    auto staged = [&](int buf, int row) -> double {
        return args.frames_f32 ? (double)reinterpret_cast<const float *>(&S.FR[buf][0][0])[row * 16 + hl] : S.FR[buf][row][hl];
    };
    R *__restrict__ out = reinterpret_cast<R *>(args.tube) + D->tube_offset;
    const double *__restrict__ wt_base = args.wavetables + (size_t)D->voice * TRM_TABLE_LENGTH;

    // ---- per-utterance constants -------------------------------------------------------------
    const bool pulse_wave = D->waveform == 0;
    const bool modulation = D->usesModulation != 0;
    const int div1 = D->div1, div2 = D->div2;
    const R tb1 = (R)D->tb1;
    const R d = (R)D->dampingFactor;

    // ---- lane roles for the section-parallel phase ----------------------------------------------
    constexpr bool FAST = sizeof(R) == 4;
    const bool is3 = hl == 3, is_end = (hl == 9 || hl == 15);
    const bool kvar = (hl <= 10) && (hl != 5);            // coefficient changes every sample
    const bool has_tap = (hl >= 1 && hl <= 8);
    const unsigned m3 = is3 ? 0xFFFFFFFFu : 0u, me = is_end ? 0xFFFFFFFFu : 0u;
    const unsigned m0 = (hl == 0) ? 0xFFFFFFFFu : 0u, m10 = (hl == 10) ? 0xFFFFFFFFu : 0u;
    const int tap_col = (hl <= 8) ? hl : 0;
    const int srcA = (lane & 16) | ((hl - 1) & 15), srcB = (lane & 16) | ((hl + 1) & 15);
    const int srcC = (lane & 16) | ((hl == 3) ? 10 : 3);
    const double *fc = (hl == 9) ? D->mouth : D->nose;
    const R f_a10 = (R)fc[0], f_b11 = (R)fc[1], f_a20 = (R)fc[2], f_a21 = (R)fc[3], f_b21 = (R)fc[4];
    // terminal lanes write their radiation output to OUTM / OUTN, everyone else to a scratch word
    R *const out_sm = (hl == 9) ? S.a.v.OUTM : ((hl == 15) ? S.a.v.OUTN : &S.a.v.SCR[hl]);
    const int out_step = is_end ? 1 : 0;
    // fast mode: sign of the y*b term of Rr, gate of the 3-way terms, gate of the frication tap
    const float sgn_y = is3 ? 1.0f : -1.0f, gate3 = is3 ? 1.0f : 0.0f;
    const float gate_tap = (has_tap && !is3) ? 1.0f : 0.0f;
    const R gate_inj = has_tap ? (R)1 : (R)0;

    // ---- shared-memory init + first two frame chunks ---------------------------------------------
    {
        // zero everything the block phases may read before they first write it (not the TMA
        // destination / mbarriers: those are only ever written through the async proxy / mbarrier ops)
        uint32_t *w = reinterpret_cast<uint32_t *>(&S.a);
        constexpr int NW = (int)((sizeof(UttSmem<R>) - offsetof(UttSmem<R>, a)) / 4);
        for (int i = hl; i < NW; i += 16) w[i] = 0u;
    }
    __syncwarp(FULL);
    if (!kvar) {
        // junctions whose coefficient never changes (pure delay 5, nasal 11..14, nose 15): fill their column once
        const double dd = D->dampingFactor;
        const double k = (hl >= 11) ? D->nasal_coeff[hl - 11] : 0.0;
        if constexpr (FAST) {
            const float4 c = (hl == 15) ? make_float4(0.0f, (float)(dd * D->nose[0] * k), (float)(-D->nose[1]), (float)(1.0 + k))
                                        : make_float4((float)(dd * (1.0 + k)), (float)(dd * k), (float)(dd * (1.0 - k)), 0.0f);
            for (int t = 0; t < TB; ++t) S.KF[t][hl] = c;
        } else {
            for (int t = 0; t < TB; ++t) S.KQ[t][hl] = (R)k;
        }
    }
    __syncwarp(FULL);
    const bool feeds = n_tube > 0;                        // this half stages frames
    const int n_chunks = (n_frames + FRAME_CHUNK - 1) / FRAME_CHUNK;
    if (hl == 0 && feeds) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        mbar_fence_init();
    }
    __syncwarp(FULL);
    if (hl == 0 && feeds) {
        for (int c = 0; c < 2 && c < n_chunks; ++c) {
            int cnt = min(FRAME_CHUNK, n_frames - c * FRAME_CHUNK);
            mbar_expect_tx(&S.mbar[c], (uint32_t)cnt * frow);
            tma_bulk_g2s(&S.FR[c][0][0], F + (size_t)c * FRAME_CHUNK * frow, (uint32_t)cnt * frow, &S.mbar[c]);
        }
    }
    if (feeds) mbar_wait(&S.mbar[0], 0);

    // ---- running state ---------------------------------------------------------------------------
    // parameter lane p = hl (TRMTubeModel.m:611-688)
    double p_cur = 0.0, p_delta = 0.0, p_next = feeds ? staged(0, 0) : 0.0;
    int f_idx = 0, jc = 0;                 // current interval, sample inside it
    // oscillator position (all lanes carry the same value)
    double pos = 0.0;
    unsigned long long pos_fx = 0ull;      // fast mode: the same position in 2^-55 table entries
    // noise MCG: state after the previous block's last draw; per-lane jump multipliers
    unsigned long long kb = args.noise_k0;
    const unsigned long long MASK44 = (1ull << 44) - 1ull;
    const unsigned long long pw0 = c_noise_pow[hl], pw1 = c_noise_pow[hl + 1], pwB = c_noise_pow[TB];
    // band-pass / throat memories (replicated in all lanes), x[n-1], x[n-2] of the band-pass input
    R y1 = 0, y2 = 0, thy = 0, xm1 = 0, xm2 = 0;
    // junction state: waves incident on this lane's junction
    R a = 0, b = 0, c3 = 0, s1bot = 0, ry = 0, rx = 0, rY = 0;

    for (int64_t n0 = 0; n0 < n_warp; n0 += TB) {
        const int64_t left = n_tube - n0;
        const int nb = left >= TB ? TB : (left > 0 ? (int)left : 0);    // valid samples of this half's block
        const bool active = hl < nb;

        // =========================================================================================
        // S0  parameter interpolation, lane = parameter (m:611-688): cur = prev; delta = (next-prev)/cp;
        //     one add per sample AFTER the sample is used.  Frame boundaries are handled between runs of
        //     plain adds; the TMA refill a boundary asks for is issued after the warp-wide sync below.
        // =========================================================================================
        int refill = -1;
        for (int s = 0; s < TB;) {
            if (jc == 0 && f_idx + 1 < n_frames && feeds) {
                const int fn = f_idx + 1;                    // frame that ends this interval
                const int ch = fn / FRAME_CHUNK;
                if ((fn % FRAME_CHUNK) == 0) {               // entering chunk ch
                    refill = ch + 1;
                    mbar_wait(&S.mbar[ch & 1], (uint32_t)((ch >> 1) & 1));
                }
                const double nxt = staged(ch & 1, fn % FRAME_CHUNK);
                p_cur = p_next;
                p_delta = (nxt - p_cur) / (double)cp;
                p_next = nxt;
            }
            const int run = min(TB - s, cp - jc);
            for (int i = 0; i < run; ++i) {
                S.a.P[s + i][hl] = p_cur;
                p_cur += p_delta;
            }
            s += run;
            jc += run;
            if (jc == cp) { jc = 0; ++f_idx; }
        }
        __syncwarp(FULL);
        if (hl == 0 && refill >= 0 && refill < n_chunks) {
            // every lane is past its last read of chunk refill-2, whose buffer is refilled now
            const int cnt = min(FRAME_CHUNK, n_frames - refill * FRAME_CHUNK);
            mbar_expect_tx(&S.mbar[refill & 1], (uint32_t)cnt * frow);
            tma_bulk_g2s(&S.FR[refill & 1][0][0], F + (size_t)refill * FRAME_CHUNK * frow, (uint32_t)cnt * frow,
                         &S.mbar[refill & 1]);
        }

        // =========================================================================================
        // A1  lane = sample t: conversions and coefficients (m:294-300, 712-773; TRMFilters.m:9-17).
        //     Coefficient math is double in both precision modes (it is feed-forward,
        //     so it costs issue slots but no serial latency, and FP32 here is the dominant error source).
        // =========================================================================================
        const double *prm = S.a.P[hl];
        {
            // pitch -> f0 -> increment: always double (a relative error here is a frequency error whose phase
            // drift grows with the length of the utterance)
            const double f0 = 220.0 * exp2(div_known(prm[0] + 3.0, 12.0, 1.0 / 12.0));
            S.INC[hl] = (f0 / 2.0) * D->basicIncrement;
        }
        double ax_d;
        R ax, ah1;
        R bp_alpha2;
        if constexpr (FAST) {
            // ---- fast mode: FP32 coefficient math in cancellation-free form -----------------------------------
            // amplitude(): 10^((dB-60)/20) = 2^((dB-60)*log2(10)/20)
            const float axf = amplitude_f((float)prm[1]);
            ax_d = (double)axf;
            {
                // the glottal closure point rint(ax*tnDelta) is the one discontinuous function of a parameter on
                // the path: when the FP32 product is near a rounding boundary, decide with the FP64 amplitude
                const float tq = axf * (float)D->tnDelta;
                const float fr = tq - floorf(tq);
                if (fabsf(fr - 0.5f) < 2e-3f) ax_d = amplitude_db(prm[1]);
            }
            ax = axf;
            ah1 = amplitude_f((float)prm[2]);
            const float fa = amplitude_f((float)prm[3]);
            const float dd = (float)D->dampingFactor;
            float r2[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { const float r = (float)prm[7 + q]; r2[q] = r * r; }
            float tap[8];
            {
                const double fpos = prm[4];
                const int ipos = (int)fpos;
                const float comp = (float)(fpos - (double)ipos);
                const float t0 = (1.0f - comp) * fa, t1 = comp * fa;
#pragma unroll
                for (int q = 0; q < 8; ++q) tap[q] = (q == ipos) ? t0 : ((ipos >= 0 && q == ipos + 1) ? t1 : 0.0f);
            }
            float4 *kf = S.KF[hl];
            // two-port junction between squared radii (ra2, rb2): k = (ra2-rb2)/(ra2+rb2),
            // 1+k = 2 ra2/(ra2+rb2), 1-k = 2 rb2/(ra2+rb2): no cancellation next to k = +-1
            auto two_port = [&](float ra2, float rb2, float tapv) {
                const float inv = __fdividef(dd, ra2 + rb2);     // one common factor for all three forms
                return make_float4(2.0f * ra2 * inv, (ra2 - rb2) * inv, 2.0f * rb2 * inv, tapv);
            };
            {
                const float4 c0 = two_port(r2[0], r2[1], 0.0f);
                kf[0].x = c0.x; kf[0].y = c0.y; kf[0].z = c0.z;          // .w (glottal input) is written by A2
            }
            kf[1] = two_port(r2[1], r2[2], tap[0]);
            kf[2] = two_port(r2[2], r2[3], tap[1]);
            kf[4] = two_port(r2[3], r2[4], tap[3]);
            kf[5].w = tap[4];                                             // pure-delay lane: only its tap varies
            kf[6] = two_port(r2[4], r2[5], tap[5]);
            kf[7] = two_port(r2[5], r2[6], tap[6]);
            kf[8] = two_port(r2[6], r2[7], tap[7]);
            {
                // 3-way junction: aL = 2 r4^2/s, aL-1 = -v^2/s, aU = 2 v^2/s, aU-1 = (v^2 - 2 r4^2)/s, s = 2 r4^2 + v^2
                const float vel = (float)prm[15], v2 = vel * vel;
                const float inv = __fdividef(dd, (r2[3] + r2[3]) + v2);
                const float daL = 2.0f * r2[3] * inv;
                kf[3] = make_float4(daL, -v2 * inv, daL, (v2 - (r2[3] + r2[3])) * inv);
                S.Z3[hl] = make_float2(2.0f * v2 * inv, tap[2]);
                kf[10] = two_port(v2, (float)D->nr1sq, 0.0f);
            }
            {
                // mouth termination: Lo = d a10 k8 a - b11 Lo_prev, radiation input (1+k8) a
                const float ap2 = (float)D->apScale2;
                const float inv = __fdividef(1.0f, r2[7] + ap2);
                kf[9] = make_float4(0.0f, dd * (float)D->mouth[0] * ((r2[7] - ap2) * inv), -(float)D->mouth[1], 2.0f * r2[7] * inv);
            }
            {
                // band-pass coefficients with the output factor 2 folded in:
                // beta = (1-tan u)/(2(1+tan u)) = (cos u - sin u)/(2(cos u + sin u))
                const float sr = (float)D->sampleRate;
                const float pi = 3.14159265358979323846f;
                // arguments lie in [0, pi): the SFU sine / cosine have an absolute error of 2^-21.4 there
                const float inv_sr = 1.0f / sr;
                float su, cu;
                __sincosf((pi * (float)prm[6]) * inv_sr, &su, &cu);
                const float cosv = __cosf(((2.0f * pi) * (float)prm[5]) * inv_sr);
                const float beta2 = __fdividef(cu - su, cu + su);          // 2*beta
                S.BC[hl][2] = beta2;
                S.BC[hl][1] = (1.0f + beta2) * cosv;                       // 2*gamma = 2(0.5+beta) cos v
                bp_alpha2 = 0.5f - 0.5f * beta2;                           // 2*alpha = (0.5-beta)
            }
        } else {
            // ---- conformance mode: the reference's operations in the reference's order, FP64 ----------------
            ax_d = amplitude_db(prm[1]);
            ax = (R)ax_d;
            ah1 = (R)amplitude_db(prm[2]);
            double r2[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { const double r = prm[7 + q]; r2[q] = r * r; }
            R *kq = S.KQ[hl];
            kq[0] = (R)((r2[0] - r2[1]) / (r2[0] + r2[1]));
            kq[1] = (R)((r2[1] - r2[2]) / (r2[1] + r2[2]));
            kq[2] = (R)((r2[2] - r2[3]) / (r2[2] + r2[3]));
            kq[4] = (R)((r2[3] - r2[4]) / (r2[3] + r2[4]));
            kq[6] = (R)((r2[4] - r2[5]) / (r2[4] + r2[5]));
            kq[7] = (R)((r2[5] - r2[6]) / (r2[5] + r2[6]));
            kq[8] = (R)((r2[6] - r2[7]) / (r2[6] + r2[7]));
            const double ap2 = D->apScale2;
            kq[9] = (R)((r2[7] - ap2) / (r2[7] + ap2));
            const double vel = prm[15];
            const double v2 = vel * vel;
            const double sum = 2.0 / ((r2[3] + r2[3]) + v2);
            kq[3] = (R)(sum * r2[3]);
            kq[16] = (R)(sum * v2);
            const double n2 = D->nr1sq;
            kq[10] = (R)((v2 - n2) / (v2 + n2));
            {
                // frication taps (m:748-765)
                const double fa = amplitude_db(prm[3]);
                const double fpos = prm[4];
                const int ipos = (int)fpos;
                const double comp = fpos - (double)ipos;
                const double rem = 1.0 - comp;
                const R t0 = (R)(rem * fa), t1 = (R)(comp * fa);
                R *tv = S.TAPV[hl];
#pragma unroll
                for (int q = 0; q < 8; ++q) tv[q + 1] = (q == ipos) ? t0 : ((ipos >= 0 && q == ipos + 1) ? t1 : (R)0);
            }
            {
                // band-pass coefficients (TRMFilters.m:9-17).  The filter output is 2*(...): the factor is folded
                // into the coefficients, which is exact (scaling by 2 commutes with rounding).
                const double sr = D->sampleRate;
                const double pi = 3.14159265358979323846;
                const double tanv = tan((pi * prm[6]) / sr);
                const double cosv = cos(((2.0 * pi) * prm[5]) / sr);
                const double beta = (1.0 - tanv) / (2.0 * (1.0 + tanv));
                S.BC[hl][2] = (R)(2.0 * beta);
                S.BC[hl][1] = (R)(2.0 * ((0.5 + beta) * cosv));
                bp_alpha2 = (R)(2.0 * ((0.5 - beta) / 2.0));
            }
        }
        // noise (TRMUtility.m:71-85 as the equivalent MCG mod 2^44) + one-zero low-pass (TRMFilters.m:81-86)
        R lp_noise;
        {
            const unsigned long long kt = (kb * pw1) & MASK44, kp = (kb * pw0) & MASK44;
            const double nz = (double)(long long)kt * TWO_M44 - 0.5;
            const double nzp = (n0 + hl == 0) ? 0.0 : ((double)(long long)kp * TWO_M44 - 0.5);
            lp_noise = (R)(nz + nzp);
            kb = (kb * pwB) & MASK44;
        }
        __syncwarp(FULL);                                   // P is dead from here: its storage is reused

        // =========================================================================================
        // S1  oscillator position (TRMWavetable.m:165-168, 28-34)
        // =========================================================================================
        double p0, p1;
        if constexpr (FAST) {
            // 64-bit fixed point, 2^55 units per table entry: 512 entries are exactly 2^64, so the table wrap is
            // the integer overflow, addition is associative and the 32 positions of a block come from a 4-step
            // warp scan instead of a 32-step dependent chain.  (Resolution 2.8e-17 entries; the reference's own
            // double accumulator rounds to 5.7e-14.)  mod0 quirk: values in (511, 512) are reported as negative.
            const unsigned long long inc_fx = __double2ull_rn(S.INC[hl] * 36028797018963968.0);
            unsigned long long incl = inc_fx + inc_fx;
#pragma unroll
            for (int o = 1; o < TB; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(FULL, incl, o, 16);
                if (hl >= o) incl += up;
            }
            const unsigned long long ub = pos_fx + incl, ua = ub - inc_fx;
            pos_fx += __shfl_sync(FULL, incl, (lane & 16) | (TB - 1));
            const unsigned long long top = 511ull << 55;
            p0 = (ua > top) ? -((double)(0ull - ua) * 2.77555756156289135e-17) : (double)ua * 2.77555756156289135e-17;
            p1 = (ub > top) ? -((double)(0ull - ub) * 2.77555756156289135e-17) : (double)ub * 2.77555756156289135e-17;
        } else {
            // sequential, in the reference's order, replicated in all lanes
#pragma unroll
            for (int s = 0; s < TB; ++s) {
                const double di = S.INC[s];
                pos = pos + di;
                pos = pos - ((pos > 511.0) ? 512.0 : 0.0);     // mod0: wraps only above 511 (TRMWavetable.m:28-34)
                const double pa = pos;
                pos = pos + di;
                pos = pos - ((pos > 511.0) ? 512.0 : 0.0);
                if (hl == 0) *reinterpret_cast<double2 *>(&S.a.v.POS[2 * s]) = make_double2(pa, pos);
            }
            __syncwarp(FULL);
            const double2 pp = *reinterpret_cast<const double2 *>(&S.a.v.POS[2 * hl]);
            p0 = pp.x; p1 = pp.y;
        }

        // =========================================================================================
        // A2  lane = sample t: table look-ups, FIR, source mixing (TRMWavetable.m:174-195, m:305-337)
        // =========================================================================================
        {
            if (!active) { p0 = 0.0; p1 = 0.0; }
            const double newDiv2 = (double)div2 - rint(ax_d * D->tnDelta);
            const double Ld = newDiv2 - (double)div1;
            int lo0 = ((int)p0) & (TRM_TABLE_LENGTH - 1), lo1 = ((int)p1) & (TRM_TABLE_LENGTH - 1);
            int hi0 = lo0 + 1, hi1 = lo1 + 1;
            if (hi0 > 511) hi0 -= 512;
            if (hi1 > 511) hi1 -= 512;
            if constexpr (FAST) {
                const float Lf = (float)Ld;
                const float scale = __fdividef(1.0f, Lf * Lf);
                const float inv_div1 = 1.0f / (float)div1;
                const float w00 = table_value_fast(wt_base, lo0, div1, inv_div1, newDiv2, scale, pulse_wave);
                const float w01 = table_value_fast(wt_base, hi0, div1, inv_div1, newDiv2, scale, pulse_wave);
                const float w10 = table_value_fast(wt_base, lo1, div1, inv_div1, newDiv2, scale, pulse_wave);
                const float w11 = table_value_fast(wt_base, hi1, div1, inv_div1, newDiv2, scale, pulse_wave);
                S.HE[FIR_HIST + hl] = w00 + ((float)(p0 - (double)lo0) * (w01 - w00));
                S.HO[FIR_HIST + hl] = w10 + ((float)(p1 - (double)lo1) * (w11 - w10));
            } else {
                const R scale = (R)(1.0 / (Ld * Ld));
                R w0 = table_value<R>(wt_base, lo0, div1, div2, newDiv2, scale, pulse_wave);
                R w1 = table_value<R>(wt_base, hi0, div1, div2, newDiv2, scale, pulse_wave);
                S.HE[FIR_HIST + hl] = w0 + ((R)(p0 - (double)lo0) * (w1 - w0));
                w0 = table_value<R>(wt_base, lo1, div1, div2, newDiv2, scale, pulse_wave);
                w1 = table_value<R>(wt_base, hi1, div1, div2, newDiv2, scale, pulse_wave);
                S.HO[FIR_HIST + hl] = w0 + ((R)(p1 - (double)lo1) * (w1 - w0));
            }
        }
        __syncwarp(FULL);
        R sig;
        {
            // 49-tap FIR at the odd sample, newest -> oldest from 0.0 (TRMFIRFilter.m:116-131)
            const R *ho = &S.HO[FIR_HIST + hl], *he = &S.HE[FIR_HIST + hl];
            R pulse0;
            if constexpr (FAST) {
                // four independent partial sums instead of one 49-long dependent chain
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                for (int q = 0; q < FIR_HIST; q += 2) {
                    s0 += ho[-q] * FirCoef<R>::at(2 * q);
                    s1 += he[-q] * FirCoef<R>::at(2 * q + 1);
                    s2 += ho[-q - 1] * FirCoef<R>::at(2 * q + 2);
                    s3 += he[-q - 1] * FirCoef<R>::at(2 * q + 3);
                }
                s0 += ho[-FIR_HIST] * FirCoef<R>::at(2 * FIR_HIST);
                pulse0 = (s0 + s1) + (s2 + s3);
            } else {
                R acc = (R)0;
#pragma unroll
                for (int q = 0; q < FIR_HIST; ++q) {
                    acc += ho[-q] * FirCoef<R>::at(2 * q);
                    acc += he[-q] * FirCoef<R>::at(2 * q + 1);
                }
                acc += ho[-FIR_HIST] * FirCoef<R>::at(2 * FIR_HIST);
                pulse0 = acc;
            }
            const R bf = (R)D->breathinessFactor, one_minus_bf = (R)(1.0 - D->breathinessFactor);
            const R pulsed_noise = lp_noise * pulse0;
            const R pulse = ax * ((pulse0 * one_minus_bf) + (pulsed_noise * bf));
            if (modulation) {
                R crossmix = ax * (R)D->crossmixFactor;
                crossmix = (crossmix < (R)1) ? crossmix : (R)1;
                sig = (pulsed_noise * crossmix) + (lp_noise * ((R)1 - crossmix));
            } else
                sig = lp_noise;
            const R tube_in = (pulse + (ah1 * sig)) * (R)0.125;
            if constexpr (FAST) S.KF[hl][0].w = tube_in; else S.TAPV[hl][0] = tube_in;
            S.BC[hl][3] = (R)D->ta0 * (pulse * (R)0.125);
        }
        {
            // band-pass feed-forward part alpha*(x[n]-x[n-2]); x[n-2] comes from two lanes down or the carry
            R x2 = __shfl_sync(FULL, sig, (lane & 16) | ((hl - 2) & 15));
            if (hl == 0) x2 = xm2;
            if (hl == 1) x2 = xm1;
            S.BC[hl][0] = bp_alpha2 * (sig - x2);
            const int last = nb > 0 ? nb - 1 : 0;
            const R l1 = __shfl_sync(FULL, sig, (lane & 16) | last);
            const R l2 = __shfl_sync(FULL, sig, (lane & 16) | (last > 0 ? last - 1 : 0));
            xm2 = (last > 0) ? l2 : xm1;
            xm1 = l1;
        }
        __syncwarp(FULL);
        {
            // slide the oscillator history down by one block (rows TB.. -> 0..)
            const R e0 = S.HE[TB + hl], o0 = S.HO[TB + hl];
            const R e1 = (hl < FIR_HIST - TB) ? S.HE[2 * TB + hl] : (R)0;
            const R o1 = (hl < FIR_HIST - TB) ? S.HO[2 * TB + hl] : (R)0;
            __syncwarp(FULL);
            S.HE[hl] = e0; S.HO[hl] = o0;
            if (hl < FIR_HIST - TB) { S.HE[TB + hl] = e1; S.HO[TB + hl] = o1; }
        }

        // =========================================================================================
        // B   lane = junction: band-pass / throat recursions + the Kelly-Lochbaum ladder (m:778-853).
        //     Always TB iterations (fully unrolled, constant shared-memory offsets); samples past the end
        //     of an utterance compute on stale data and are never stored.
        // =========================================================================================
        // The three junction roles (two-port, 3-way, mouth/nose termination) run as ONE straight-line instruction
        // stream: every lane evaluates the same expressions on role-routed operands (LOP3 blends / 0-1 gates), so
        // there is no divergent region, no reconvergence barrier and no divergence check before the shuffles.
        if constexpr (FAST) {
#pragma unroll
            for (int s = 0; s < TB; ++s) {
                const float4 bc = *reinterpret_cast<const float4 *>(S.BC[s]);
                const float fr = (bc.x + (bc.y * y1)) - (bc.z * y2);       // frication band-pass (TRMFilters.m:19-29)
                y2 = y1; y1 = fr;
                const float th = bc.w + (tb1 * thy);                       // throat low-pass (TRMFilters.m:72-77)
                thy = th;
                S.a.v.TH[s] = th;

                const float4 c = S.KF[s][hl];
                const float2 z3 = S.Z3[s];
                const float rc = z3.x * gate3;                             // d aU on the 3-way lane, 0 elsewhere
                const float inj = (c.w * gate_tap) + (z3.y * gate3);
                const float Rr = (((c.x * a) + ((c.y * sgn_y) * b)) + (rc * c3)) + (inj * fr);
                const float Lo = ((c.y * a) + (c.z * b)) + (rc * c3);
                const float x3 = ((c.x * a) + (c.x * b)) + (c.w * c3);     // 3-way: d aL (a+b) + d(aU-1) c
                const float X3 = blend(m3, x3, Lo);
                const float xr = c.w * a;                                  // termination: (1+k) a
                const float rad = ((f_a20 * xr) + (f_a21 * rx)) - (f_b21 * rY);
                rx = xr; rY = rad;
                out_sm[s * out_step] = rad;
                const float nA = __shfl_sync(FULL, Rr, srcA);
                const float nB = __shfl_sync(FULL, Lo, srcB);
                const float nC = __shfl_sync(FULL, X3, srcC);
                const float a0 = (s1bot * d) + c.w;                        // glottis end (m:792)
                a = blend(m0, a0, blend(m10, nC, nA));
                s1bot = Lo;
                b = blend(me, Lo, nB);                                     // a termination feeds back its own Lo
                c3 = nC;
            }
        } else {
#pragma unroll
            for (int s = 0; s < TB; ++s) {
                // frication band-pass recursion (TRMFilters.m:19-29), factor 2 folded into the coefficients
                const R fr = (S.BC[s][0] + (S.BC[s][1] * y1)) - (S.BC[s][2] * y2);
                y2 = y1; y1 = fr;
                // throat low-pass (TRMFilters.m:72-77)
                const R th = S.BC[s][3] + (tb1 * thy);
                thy = th;
                S.a.v.TH[s] = th;

                const R k = S.KQ[s][hl];
                const R aU = S.KQ[s][16];
                const R inj = S.TAPV[s][tap_col];
                // two-port junction (m:796-829); a termination is the same expression with no right neighbour
                const R bb = blend(me, (R)0, b);
                const R delta = k * (a - bb);
                // 3-way junction (m:810-813)
                const R p = ((k * a) + (k * b)) + (aU * c3);
                // termination: reflection + radiation filter pair (m:832-835,846-849; TRMFilters.m:47-60)
                const R refl = (f_a10 * delta) - (f_b11 * ry);
                ry = refl;
                const R xr = ((R)1 + k) * a;
                const R rad = ((f_a20 * xr) + (f_a21 * rx)) - (f_b21 * rY);
                rx = xr; rY = rad;
                out_sm[s * out_step] = rad;

                const R Rr = (blend(m3, p - b, a + delta) * d) + ((inj * gate_inj) * fr);
                const R Lo = blend(m3, p - a, blend(me, refl, b + delta)) * d;
                const R X3 = blend(m3, (p - c3) * d, Lo);

                const R nA = __shfl_sync(FULL, Rr, srcA);
                const R nB = __shfl_sync(FULL, Lo, srcB);
                const R nC = __shfl_sync(FULL, X3, srcC);
                const R a0 = (s1bot * d) + inj;                            // glottis end (m:792)
                a = blend(m0, a0, blend(m10, nC, nA));
                s1bot = Lo;
                b = nB;
                c3 = nC;
            }
        }
        __syncwarp(FULL);

        // =========================================================================================
        // A4  lane = sample t: sum mouth + nose + throat (m:835,849,341); 128-bit coalesced store
        // =========================================================================================
        S.a.v.YB[hl] = (S.a.v.OUTM[hl] + S.a.v.OUTN[hl]) + (S.a.v.TH[hl] * (R)D->throatGain);
        __syncwarp(FULL);
        {
            constexpr int VEC = 16 / (int)sizeof(R);          // elements per 128-bit store
            R *dst = out + n0;                                // n0 and tube_offset are multiples of 16
            if (nb == TB) {
                if (hl < TB / VEC) reinterpret_cast<float4 *>(dst)[hl] = reinterpret_cast<const float4 *>(S.a.v.YB)[hl];
            } else if (active) {
                dst[hl] = S.a.v.YB[hl];
            }
        }
        __syncwarp(FULL);                                     // v.* is reused as P by the next block
    }
}

}  // namespace TRM_KERNEL_NS
