// jet_common.cuh -- what the two fused-kernel engines (jet_mlp_kernel.cuh: FP32 pipe, jet_mlp_kernel_mma.cuh: tensor pipe) share:
// the CUDA / CPU-emulation switch, the plan constants that do not depend on the engine, memory helpers, the per-warp jet-buffer and
// L2-scratch accessors, the activation derivative recurrences (tanh / Rational, Code/Classes/Network.py:32-59,161-162) with the
// Rational coefficient adjoint, the layer-0 pre-activation jets, and the CTA-record reduction arguments.
// Included by the engine headers at the END of a generated plan translation unit (plan_codegen.cpp), which defines
//   PDL_D, PDL_J, PDL_WP, PDL_H, PDL_K, PDL_ACT, PDL_NS, PDL_NW, PDL_MODE, ...
//   pdl_width(l), pdl_comp_axis(c), pdl_tanh_sigmas(), pdl_act_fwd(), pdl_act_adj(), pdl_epilogue()
// The same source compiles as plain C++ with -DPDL_EMULATE (32 lanes in a loop): TEST INFRASTRUCTURE only.
#pragma once
#include <stdint.h>
#include <math.h>
#include "kernel_params.h"

#ifdef PDL_EMULATE
#include <string.h>
#include <vector>
#define PDL_DEV static inline
#define PDL_HD static inline
#define PDL_NL 32
#define PDL_FOR_LANES for (int lane = 0; lane < 32; ++lane)
#define PDL_L lane
#define PDL_SYNCWARP() do {} while (0)
struct pdl_f2 { float x, y; };
struct pdl_f4 { float x, y, z, w; };
#else
#define PDL_DEV __device__ __forceinline__
#define PDL_HD __host__ __device__ __forceinline__
#define PDL_NL 1
#define PDL_FOR_LANES
#define PDL_L 0
#define PDL_SYNCWARP() __syncwarp()
typedef float2 pdl_f2;
typedef float4 pdl_f4;
#endif

namespace pdl {

constexpr int D = PDL_D, J = PDL_J, WP = PDL_WP, H = PDL_H, K = PDL_K, NS = PDL_NS, NW = PDL_NW;
constexpr int KX = (K > 0) ? K : 1;
constexpr int NPL = WP / 4;                 // neurons per lane group (fwd / dgrad micro-tile rows)
constexpr int NPN = WP / 8;                 // neurons per lane group along n in wgrad
constexpr int JA = (J >= 3) ? 4 : J;        // plane A width (comps 0..3)
constexpr int JB = (J <= 4) ? 0 : ((J == 5) ? 1 : ((J == 6) ? 2 : ((J <= 8) ? 4 : 8)));   // plane B width (comps 4..)
constexpr int pad_stride(int n) { int m = (n + 3) / 4 * 4; while (((m / 4) & 1) == 0) m += 4; return m; }
constexpr int SA = pad_stride(WP * JA);     // floats per point in plane A (odd number of 16B chunks)
constexpr int SB = (JB > 0) ? pad_stride(WP * JB) : 0;
constexpr int NHH = (H > 1) ? (H - 1) : 0;  // hidden->hidden linear layers
constexpr int ZS_LAYER = NPL * 32 * (JA + JB);
constexpr int ZSCR = NHH * ZS_LAYER;        // floats of pre-activation scratch per warp

// per-warp small accumulators (shared memory, exported per CTA)
constexpr int SO_W0 = 0;                    // [WP][D]
constexpr int SO_B0 = SO_W0 + WP * D;       // [WP]
constexpr int SO_WH = SO_B0 + WP;           // [WP]
constexpr int SO_BH = SO_WH + WP;           // [1]
constexpr int SO_RAT = SO_BH + 1;           // [H][7]
constexpr int SO_XI = SO_RAT + 7 * H;       // [KX][8] one slot per point-lane
constexpr int SACC = (SO_XI + KX * 8 + 3) / 4 * 4;
constexpr int SACC_OUT = SO_XI + KX;        // what a CTA exports (xi slots summed)
constexpr int SM_W0 = 0;                    // shared-memory offsets every engine agrees on: layer-0 weights and bias come first
constexpr int SM_B0 = SM_W0 + WP * D;

// flat parameter layout (reference order: all (W,b) pairs, then (a,b) per hidden layer)
constexpr int NPRM = 2 * (H + 1) + ((PDL_ACT == 1) ? 2 * H : 0);

typedef PdlKernelParams Params;

// ---------------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------------
PDL_DEV pdl_f4 ld4(const float* p) { return *reinterpret_cast<const pdl_f4*>(p); }
PDL_DEV pdl_f2 ld2(const float* p) { return *reinterpret_cast<const pdl_f2*>(p); }
PDL_DEV void st4(float* p, float a, float b, float c, float d) { pdl_f4 v; v.x = a; v.y = b; v.z = c; v.w = d; *reinterpret_cast<pdl_f4*>(p) = v; }
PDL_DEV void st2(float* p, float a, float b) { pdl_f2 v; v.x = a; v.y = b; *reinterpret_cast<pdl_f2*>(p) = v; }

#ifdef PDL_EMULATE
PDL_DEV pdl_f4 ldg4(const float* p) { return ld4(p); }
PDL_DEV pdl_f2 ldg2(const float* p) { return ld2(p); }
PDL_DEV float ldg1(const float* p) { return *p; }
PDL_DEV void stg4(float* p, pdl_f4 v) { *reinterpret_cast<pdl_f4*>(p) = v; }
PDL_DEV void stg2(float* p, pdl_f2 v) { *reinterpret_cast<pdl_f2*>(p) = v; }
PDL_DEV void stg1(float* p, float v) { *p = v; }
#else
// scratch traffic is cached at L2 only: it is private to one warp and must not evict L1/shared carve-out
PDL_DEV pdl_f4 ldg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
PDL_DEV pdl_f2 ldg2(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }
PDL_DEV float ldg1(const float* p) { return __ldcg(p); }
PDL_DEV void stg4(float* p, pdl_f4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
PDL_DEV void stg2(float* p, pdl_f2 v) { __stcg(reinterpret_cast<float2*>(p), v); }
PDL_DEV void stg1(float* p, float v) { __stcg(p, v); }
#endif

// vector access of W floats (W in {1,2,4,8}); shared/generic address space
template <int W> PDL_DEV void ldw(const float* p, float* v) {
    if (W == 8) { pdl_f4 a = ld4(p), b = ld4(p + 4); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
    else if (W == 4) { pdl_f4 a = ld4(p); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
    else if (W == 2) { pdl_f2 a = ld2(p); v[0] = a.x; v[1] = a.y; }
    else if (W == 1) { v[0] = p[0]; }
}
template <int W> PDL_DEV void stw(float* p, const float* v) {
    if (W == 8) { st4(p, v[0], v[1], v[2], v[3]); st4(p + 4, v[4], v[5], v[6], v[7]); }
    else if (W == 4) st4(p, v[0], v[1], v[2], v[3]);
    else if (W == 2) st2(p, v[0], v[1]);
    else if (W == 1) p[0] = v[0];
}
// same, global memory cached at L2 only (scratch)
template <int W> PDL_DEV void ldgw(const float* p, float* v) {
    if (W == 8) { pdl_f4 a = ldg4(p), b = ldg4(p + 4); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
    else if (W == 4) { pdl_f4 a = ldg4(p); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
    else if (W == 2) { pdl_f2 a = ldg2(p); v[0] = a.x; v[1] = a.y; }
    else if (W == 1) { v[0] = ldg1(p); }
}
template <int W> PDL_DEV void stgw(float* p, const float* v) {
    if (W == 8) { pdl_f4 a, b; a.x = v[0]; a.y = v[1]; a.z = v[2]; a.w = v[3]; b.x = v[4]; b.y = v[5]; b.z = v[6]; b.w = v[7]; stg4(p, a); stg4(p + 4, b); }
    else if (W == 4) { pdl_f4 a; a.x = v[0]; a.y = v[1]; a.z = v[2]; a.w = v[3]; stg4(p, a); }
    else if (W == 2) { pdl_f2 a; a.x = v[0]; a.y = v[1]; stg2(p, a); }
    else if (W == 1) stg1(p, v[0]);
}
constexpr int JT = JA + JB;                 // padded component count (>= J)
// one (point, neuron) entry of a jet buffer: comps 0..JA-1 in plane A, the rest in plane B
PDL_DEV void load_entry(const float* A, const float* B, float (&v)[J]) {
    float t[JT];
    ldw<JA>(A, t);
    if (JB > 0) ldw<JB>(B, t + JA);
#pragma unroll
    for (int c = 0; c < J; ++c) v[c] = t[c];
}
PDL_DEV void store_entry(float* A, float* B, const float (&v)[J]) {
    float t[JT];
#pragma unroll
    for (int c = 0; c < JT; ++c) t[c] = (c < J) ? v[c < J ? c : 0] : 0.f;
    stw<JA>(A, t);
    if (JB > 0) stw<JB>(B, t + JA);
}
// scratch (global, L2) entry for (layer slot, neuron i, lane): coalesced across lanes
PDL_DEV void scratch_store(float* zs, int slot, int i, int lane, const float (&v)[J]) {
    float t[JT];
#pragma unroll
    for (int c = 0; c < JT; ++c) t[c] = (c < J) ? v[c < J ? c : 0] : 0.f;
    stgw<JA>(zs + slot * ZS_LAYER + (i * 32 + lane) * JA, t);
    if (JB > 0) stgw<JB>(zs + slot * ZS_LAYER + NPL * 32 * JA + (i * 32 + lane) * JB, t + JA);
}
PDL_DEV void scratch_load(const float* zs, int slot, int i, int lane, float (&v)[J]) {
    float t[JT];
    ldgw<JA>(zs + slot * ZS_LAYER + (i * 32 + lane) * JA, t);
    if (JB > 0) ldgw<JB>(zs + slot * ZS_LAYER + NPL * 32 * JA + (i * 32 + lane) * JB, t + JA);
#pragma unroll
    for (int c = 0; c < J; ++c) v[c] = t[c];
}

// cross-lane sum over lanes whose index differs in bit m (butterfly step)
template <int N>
PDL_DEV void xor_add(float (&v)[PDL_NL][N], int m) {
#ifdef PDL_EMULATE
    float t[32][N];
    for (int l = 0; l < 32; ++l) for (int n = 0; n < N; ++n) t[l][n] = v[l][n];
    for (int l = 0; l < 32; ++l) for (int n = 0; n < N; ++n) v[l][n] = t[l][n] + t[l ^ m][n];
#else
#pragma unroll
    for (int n = 0; n < N; ++n) v[0][n] += __shfl_xor_sync(0xffffffffu, v[0][n], m);
#endif
}

// ---------------------------------------------------------------------------------------------
// activation sigmas: s[m] = sigma^(m)(z0), m = 0..NS-1
// ---------------------------------------------------------------------------------------------
// reciprocal of the Rational denominator: MUFU.RCP (1 ulp for normal inputs) instead of the ~9-instruction IEEE division with its
// slow-path branch; see jet_mlp_kernel_mma.cuh
PDL_DEV float rcp_fast(float d) {
#if defined(PDL_EMULATE) || defined(PDL_RCP_IEEE)
    return 1.0f / d;
#else
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
#endif
}
PDL_DEV float rat_fact(int k) { float f = 1.f; for (int i = 2; i <= k; ++i) f *= (float)i; return f; }

PDL_DEV void sigmas(float z0, const float* rp, float (&s)[NS], float (&r)[NS], float& inv_d0) {
#if PDL_ACT == 0
    (void)rp; (void)r; inv_d0 = 0.f;
    pdl_tanh_sigmas(tanhf(z0), s);
#else
    // Rational R = N/D, N cubic, D quadratic (Network.py:55-59).  r[k] = Taylor coefficient of R at z0.
    const float a0 = rp[0], a1 = rp[1], a2 = rp[2], a3 = rp[3], b0 = rp[4], b1 = rp[5], b2 = rp[6];
    float nk[4];
    nk[0] = a0 + z0 * (a1 + z0 * (a2 + a3 * z0));
    nk[1] = a1 + z0 * (2.f * a2 + 3.f * a3 * z0);
    nk[2] = a2 + 3.f * a3 * z0;
    nk[3] = a3;
    const float d0 = b0 + z0 * (b1 + b2 * z0), d1 = b1 + 2.f * b2 * z0, d2 = b2;
    inv_d0 = rcp_fast(d0);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        float t = (k < 4) ? nk[k < 4 ? k : 0] : 0.f;
        if (k >= 1) t -= d1 * r[k >= 1 ? k - 1 : 0];
        if (k >= 2) t -= d2 * r[k >= 2 ? k - 2 : 0];
        r[k] = t * inv_d0;
        s[k] = rat_fact(k) * r[k];
    }
#endif
}

#if PDL_ACT == 1
// adjoint of the Taylor recurrence w.r.t. the 7 coefficients; sb[m] = dL/d sigma_m for m = 0..NS-2
PDL_DEV void rational_param_adjoint(float z0, const float* rp, const float (&r)[NS], float inv_d0,
                                    const float (&sb)[NS - 1], float (&ga)[7]) {
    const float b1 = rp[5], b2 = rp[6];
    const float d1 = b1 + 2.f * b2 * z0, d2 = b2;
    float rb[NS - 1];
#pragma unroll
    for (int k = 0; k < NS - 1; ++k) rb[k] = rat_fact(k) * sb[k];
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    float d0b = 0.f, d1b = 0.f, d2b = 0.f;
#pragma unroll
    for (int k = NS - 2; k >= 0; --k) {
        const float t = rb[k] * inv_d0;
        if (k < 4) nb[k < 4 ? k : 0] += t;
        d0b -= t * r[k];
        if (k >= 1) { d1b -= t * r[k >= 1 ? k - 1 : 0]; rb[k >= 1 ? k - 1 : 0] -= t * d1; }
        if (k >= 2) { d2b -= t * r[k >= 2 ? k - 2 : 0]; rb[k >= 2 ? k - 2 : 0] -= t * d2; }
    }
    const float z2 = z0 * z0, z3 = z2 * z0;
    ga[0] += nb[0];
    ga[1] += nb[0] * z0 + nb[1];
    ga[2] += nb[0] * z2 + 2.f * nb[1] * z0 + nb[2];
    ga[3] += nb[0] * z3 + 3.f * nb[1] * z2 + 3.f * nb[2] * z0 + nb[3];
    ga[4] += d0b;
    ga[5] += d0b * z0 + d1b;
    ga[6] += d0b * z2 + 2.f * d1b * z0 + d2b;
}
#endif

// layer-0 pre-activation jets of neuron n: value = b + W x, first-order comps = weight columns, rest 0
PDL_DEV void layer0_preact(const float* sm, int n, const float (&x)[D], float (&z)[J]) {
    float v = sm[SM_B0 + n];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) v += sm[SM_W0 + n * D + dd] * x[dd];
    z[0] = v;
#pragma unroll
    for (int c = 1; c < J; ++c) z[c] = (pdl_comp_axis(c) >= 0) ? sm[SM_W0 + n * D + (pdl_comp_axis(c) >= 0 ? pdl_comp_axis(c) : 0)] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// final reduction over CTAs: arguments, flat parameter offsets, loss statistics (the per-engine record layout lives in reduce_store)
// ---------------------------------------------------------------------------------------------
struct ReduceArgs {
    const float* cta_out; const double* cta_stats; int n_ctas;
    float* grad_params;      // flat, reference order (W0,b0,...,WH,bH,a0,b0,...)
    float* grad_xi;          // [K]
    double* stats;           // [0] = mean loss (sum r^2 * inv_n), [1] = sum r^2, [2] = sum |r|
    double inv_n;
    const long long* n_points_dev; long long capacity;
};

PDL_HD int prm_offset(int idx) {          // flat offset of parameter tensor idx
    int off = 0;
    for (int t = 0; t < idx; ++t) {
        int sz;
        if (t < 2 * (H + 1)) { const int l = t / 2; sz = (t & 1) ? pdl_width(l + 1) : pdl_width(l + 1) * pdl_width(l); }
        else sz = ((t - 2 * (H + 1)) & 1) ? 3 : 4;
        off += sz;
    }
    return off;
}
PDL_DEV void reduce_stats(const ReduceArgs& R) {
    double s2 = 0.0, s1 = 0.0;
    for (int c = 0; c < R.n_ctas; ++c) { s2 += R.cta_stats[(size_t)c * 4]; s1 += R.cta_stats[(size_t)c * 4 + 1]; }
    double inv_n = R.inv_n;
    if (R.n_points_dev) { long long n = *R.n_points_dev; if (n > R.capacity) n = R.capacity; inv_n = n > 0 ? 1.0 / (double)n : 0.0; }
    R.stats[0] = s2 * inv_n; R.stats[1] = s2; R.stats[2] = s1;
}

}  // namespace pdl
