// jet_mlp_kernel_mma.cuh -- fused jet-MLP + library-residual forward/adjoint kernel for sm_100a, tensor-pipe variant.
//
// Same contract, tiling (8 points per warp tile), scratch, reductions and helpers as jet_mlp_kernel.cuh (see there for the
// reference map).  The contractions whose A operand is produced by the lane itself -- forward  z = W h  and, optionally, the
// data gradient  hb = zb W  -- run as split-precision mma.sync on the tensor pipe, concurrently with the FP32 pipe that keeps
// the jet activations and the epilogue (and the weight gradient unless PDL_WGRAD_MMA moves it to the tensor pipe as well).
// FP32-class accuracy comes from the split x = hi + lo (hi = TF32(x), lo = x - hi):  D += hi*Bhi (TF32 m16n8k8) + the two cross
// terms lo*B + hi*Blo either as two more TF32 MMAs (3xTF32) or as ONE 16-bit m16n8k16 MMA (PDL_CROSS_BF16 bit mask: scaled fp16 in
// the forward pass, bf16 in the adjoint: the defaults).
//
// Included at the END of a generated plan translation unit when the plan compiler picks this engine (PDL_ENGINE=mma).
// The same source compiles as plain C++ with -DPDL_EMULATE (mma.sync is modelled lane-exactly): TEST INFRASTRUCTURE only.
#pragma once
#include "jet_common.cuh"

namespace pdl {

constexpr int SF = WP;                      // row stride of the fragment-ordered weight copies (rows 8 banks apart for WP = 40)
constexpr int NT8 = WP / 8;                 // 8-wide neuron tiles (MMA n-tiles and k-steps)
constexpr int NM = (J + 1) / 2;             // component pairs = m16 row blocks per neuron tile
constexpr bool DGRAD_MMA = (PDL_DGRAD_MMA != 0);   // data gradient on the tensor pipe too
constexpr bool DGRAD_T = (PDL_DGRAD_MMA == 1);     // ... with its B fragments from W^T hi/lo copies (conflict-free LDS.64); mode 2 reads
                                                   // them from the forward copies instead (two LDS.32, 2-way conflicts, no extra memory)
#ifndef PDL_CROSS_BF16
#define PDL_CROSS_BF16 0
#endif
// XBF: the two cross terms lo*Bhi + hi*Blo of a split product run as ONE bf16 m16n8k16 MMA (k slots 0-7 carry lo(x) against
// bf16(W), slots 8-15 carry bf16(x) against lo(W)); with the TF32 hi*hi MMA that is 2 instead of 3 tensor-pipe instructions per
// product.  The cross terms are 2^-12 of the product, so bf16's 2^-9 rounding leaves an error of ~2^-21 per product.
// PDL_CROSS_BF16 is a bit mask: 1 = forward, 2 = data gradient, 4 = weight gradient.
// bit 8: the forward cross terms as ONE scaled fp16 m16n8k16 MMA (11-bit significands: error at the 3xTF32 level, see split_cross_f16);
// it uses the same packed cross plane of W as bit 1, so XBF_F ("forward cross terms in one 16-bit MMA") covers both
constexpr bool XF16_F = (PDL_CROSS_BF16 & 8) != 0;
constexpr bool XBF_F = (PDL_CROSS_BF16 & 1) != 0 || XF16_F, XBF_D = (PDL_CROSS_BF16 & 2) != 0, XBF_W = (PDL_CROSS_BF16 & 4) != 0;
constexpr bool WGRAD_MMA = (PDL_WGRAD_MMA != 0);   // weight gradient on the tensor pipe (k = the tile's 8 points, one k-step per component)
constexpr int NMT = (WP + 15) / 16;         // m16 tiles over the output neurons of a weight-gradient block
constexpr int NFRAG = NMT * NT8;            // C fragments of one layer's weight gradient
#ifndef PDL_WGRAD_NTB
#define PDL_WGRAD_NTB 3
#endif
#ifndef PDL_WGRAD_ROLL_MT
#define PDL_WGRAD_ROLL_MT (PDL_J >= 5)      // measured: from J = 5 on the unrolled m-tile loop costs more in instruction fetch than it gains
#endif
constexpr int NTB = PDL_WGRAD_NTB;          // n8 tiles per weight-gradient step (independent MMA chains vs registers)
#ifndef PDL_ADJ_FUSED
#define PDL_ADJ_FUSED 1
#endif
// order of the hidden-layer adjoint: 1 = one activation pass per layer shared by the recomputed forward and the adjoint (one tanh /
// Rational recurrence less per neuron and layer), 0 = the classic order (adjoint of layer l and recompute of layer l-1 in one loop)
constexpr bool ADJ_FUSED = (PDL_ADJ_FUSED != 0);
constexpr int GS = PDL_GS;                  // group stride of the dgrad weight copy  Wd[n][g4][i]
constexpr int SD = 4 * GS;
constexpr int NACC = NPN * NPL + NPN;       // wgrad accumulators per lane (weights + bias)
constexpr int NACC4 = (NACC + 3) / 4 * 4;
constexpr int WPRIV_LAYER = WGRAD_MMA ? (NFRAG + 1) * 128 : NACC4 * 32;   // floats per warp per hidden->hidden layer (MMA: C fragments + bias slot)
constexpr int WPRIV = NHH * WPRIV_LAYER;

// small per-warp accumulators (shared memory), in floats

// shared-memory map (floats).  Shared part first, then NW private parts.
constexpr int SM_HH = (SM_B0 + WP + 3) / 4 * 4;              // per hidden->hidden layer: Wf, Wd, bias
// per hidden->hidden layer: W hi/lo [n][k] (forward B fragments), then either W^T hi/lo [k][n] (dgrad B fragments) or the
// FP32-pipe dgrad copy Wd[n][t][i]; bias
// With XBF the "lo" planes hold the packed bf16 cross operands instead: word (n, 2j) = {bf16 W[n][2j], bf16 W[n][2j+1]}, word
// (n, 2j+1) = {bf16 lo(W[n][2j]), bf16 lo(W[n][2j+1])} -- one LDS.64 per B fragment, as before.  Mode 2 then needs the transposed
// cross plane as its third plane (HH_THI is unused there).
constexpr int HH_WHI = 0, HH_WLO = WP * SF, HH_THI = 2 * WP * SF, HH_TLO = DGRAD_T ? 3 * WP * SF : 2 * WP * SF, HH_WD = 2 * WP * SF;
constexpr int HH_B = DGRAD_T ? 4 * WP * SF : (DGRAD_MMA ? ((XBF_D ? 3 : 2) * WP * SF) : (2 * WP * SF + WP * SD));
static_assert(!(XBF_F && DGRAD_MMA && !DGRAD_T && !XBF_D), "mode-2 data gradient reads the TF32 lo plane the bf16 forward replaces");
constexpr int HH_SIZE = (HH_B + WP + 3) / 4 * 4;
constexpr int SM_WH = SM_HH + NHH * HH_SIZE;
constexpr int SM_BH = SM_WH + WP;
constexpr int SM_RAT = (SM_BH + 1 + 3) / 4 * 4;              // [H][8]: a0..a3, b0..b2, pad
constexpr int SM_XI = SM_RAT + 8 * H;                        // xi with masked entries zeroed
constexpr int SM_GM = SM_XI + KX;                            // 1 - mask
constexpr int SM_SHARED = (SM_GM + KX + 3) / 4 * 4;
constexpr int BUF = (8 * (SA + SB) > 128) ? 8 * (SA + SB) : 128;   // also hosts 64 doubles at warp exit
constexpr int SM_WARP = 2 * BUF + SACC;                      // bufH, bufZ, small accumulators
constexpr int SM_TOTAL = SM_SHARED + NW * SM_WARP;
constexpr int NT = NW * 32;

constexpr int CTA_OUT = WPRIV + SACC_OUT;                    // floats exported per CTA

// ---------------------------------------------------------------------------------------------
// warp-level tensor-core step: D(16x8) += A(16x8) * B(8x8), TF32 inputs, FP32 accumulate (mma.sync.m16n8k8)
//   lane = 4*g + t:  A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B: b0 (k=t, n=g) b1 (k=t+4, n=g);
//                    C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// Rows 0-7 are jet component 2m of the tile's 8 points, rows 8-15 component 2m+1 of the SAME points (m = component pair), so a
// lane (g, t) holds every component of point g for its neurons and the jet activation stays lane-local.  The k index is a
// summation index, so it is relabelled: slot t <-> neuron 8s+2t and slot t+4 <-> neuron 8s+2t+1, which makes a lane's C
// fragment of neuron tile s exactly its A fragment of k-step s: activations never pass through shared memory between layers.
// ---------------------------------------------------------------------------------------------
PDL_DEV uint32_t f2u(float x) {
#ifdef PDL_EMULATE
    uint32_t u; memcpy(&u, &x, 4); return u;
#else
    return __float_as_uint(x);
#endif
}
PDL_DEV float u2f(uint32_t u) {
#ifdef PDL_EMULATE
    float x; memcpy(&x, &u, 4); return x;
#else
    return __uint_as_float(u);
#endif
}
// round-to-nearest (ties away) FP32 -> TF32 bit pattern (low 13 mantissa bits zero).  Written as integer arithmetic on purpose:
// cvt.rna.tf32.f32 expands to four SASS instructions (FSETP/IADD3/LOP3/SEL, the NaN path), this is two and identical for every
// finite input (inf stays inf; a NaN still ends up as a NaN in the accumulator through the lo part).
PDL_DEV uint32_t tf32_rna(float x) { return (f2u(x) + 0x1000u) & 0xFFFFE000u; }
// FP32 value -> (hi, lo): hi = TF32(x), lo = TF32(x - hi); x - hi is exact in FP32, so hi + lo carries ~22 mantissa bits
PDL_DEV void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
#ifdef PDL_SPLIT_RAW_ALL
    lo = f2u(x - u2f(hi));
#else
    lo = tf32_rna(x - u2f(hi));
#endif
}
// same without rounding the low part: the tensor pipe reads only the upper 19 bits of an operand register, i.e. it truncates lo
// itself (error < 2^-21 |x|, the size of the dropped lo*lo term).  Used where splits dominate the instruction count and the
// result is a long sum in which the truncation errors average out (weight gradient).
PDL_DEV void split_tf32_raw(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = f2u(x - u2f(hi));
}
PDL_DEV void split_tf32_dg(float x, uint32_t& hi, uint32_t& lo) {
#ifdef PDL_SPLIT_ROUND_DG
    split_tf32(x, hi, lo);
#else
    split_tf32_raw(x, hi, lo);
#endif
}
// two FP32 values -> packed bf16x2, round to nearest even; `k0` lands in the low half (the lower k index of an MMA fragment)
PDL_DEV uint32_t pack_bf16(float k0, float k1) {
#ifdef PDL_EMULATE
    const uint32_t a = f2u(k0), b = f2u(k1);
    const uint32_t ra = (a + 0x7FFFu + ((a >> 16) & 1u)) >> 16, rb = (b + 0x7FFFu + ((b >> 16) & 1u)) >> 16;
    return (rb << 16) | (ra & 0xFFFFu);
#else
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(k1), "f"(k0));
    return r;
#endif
}
// A-operand registers of one component pair for the 2-MMA scheme: hi[] = TF32 fragment, x[] = bf16 cross fragment.
//   v0/v1 = the lane's two rows (g, g+8) at its first k element, v2/v3 = the same rows at its second k element
PDL_DEV void split_cross(float v0, float v1, float v2, float v3, uint32_t (&hi)[4], uint32_t (&x)[4]) {
    hi[0] = tf32_rna(v0); hi[1] = tf32_rna(v1); hi[2] = tf32_rna(v2); hi[3] = tf32_rna(v3);
    x[0] = pack_bf16(v0 - u2f(hi[0]), v2 - u2f(hi[2]));
    x[1] = pack_bf16(v1 - u2f(hi[1]), v3 - u2f(hi[3]));
    x[2] = pack_bf16(v0, v2);
    x[3] = pack_bf16(v1, v3);
}
// ---- scaled fp16 cross terms (forward pass) ----
// fp16 has the 11-bit significand the cross terms need (bf16's 8 bits leave 1.4e-5 on the derivative features) but only 5 exponent
// bits.  The two products of the cross MMA are scaled apart by exact powers of two:  slots 0-7  (lo(x) * 2^11) * (W * 2^-11),
// slots 8-15  x * lo(W).  lo(x) * 2^11 ~ x/2 is a normal fp16 for |x| > 2^-13; W * 2^-11 is normal for |W| >= 1/8 and subnormal below
// (absolute error 2^-25, i.e. 2^-14 in W): the term lo(x) * W then carries an error of |x| * 2^-26 -- below FP32's own rounding of the
// sum.  x and lo(W) in the other product are rounded to 11 bits (or, tiny, to fp16 subnormals: absolute 2^-25), which changes a
// term of size 2^-12 |x W| by 2^-23 |x W| at most.  Conversions saturate (.satfinite): |x| > 65504 degrades to TF32 accuracy
// instead of producing inf.
constexpr float F16_UP = 2048.f, F16_DOWN = 1.f / 2048.f;
#ifdef PDL_EMULATE
PDL_DEV uint32_t f32_to_f16_bits(float f) {            // round to nearest even, subnormals, saturating
    uint32_t x = f2u(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x > 0x7F800000u) return sign | 0x7E00u;                          // NaN
    if (x >= 0x477FF000u) return sign | 0x7BFFu;                         // >= 65520 (rounds past the largest finite): saturate
    if (x < 0x33000001u) return sign;                                    // < 2^-25 (or exactly 2^-25: ties to even = 0)
    const int e = (int)(x >> 23) - 127;
    uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    int shift = (e >= -14) ? 13 : (13 + (-14 - e));                      // bits dropped from the 24-bit significand
    const uint32_t half = 1u << (shift - 1), rem = m & ((1u << shift) - 1u);
    uint32_t q = m >> shift;
    if (rem > half || (rem == half && (q & 1u))) ++q;
    uint32_t out = (e >= -14) ? ((uint32_t)(e + 15) << 10) + (q - 0x400u) : q;   // a carry out of the significand bumps the exponent
    return sign | out;
}
PDL_DEV float f16_bits_to_f32(uint32_t h) {
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    if (e == 0) { const float v = (float)m * 5.9604644775390625e-08f; return u2f(f2u(v) | sign); }
    if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}
#endif
// two FP32 values -> packed f16x2, round to nearest even, saturating; `k0` lands in the low half
PDL_DEV uint32_t pack_f16(float k0, float k1) {
#ifdef PDL_EMULATE
    return (f32_to_f16_bits(k1) << 16) | (f32_to_f16_bits(k0) & 0xFFFFu);
#else
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(k1), "f"(k0));
    return r;
#endif
}
// A-operand registers of one component pair: hi[] = TF32 fragment, x[] = scaled fp16 cross fragment (layout as split_cross)
PDL_DEV void split_cross_f16(float v0, float v1, float v2, float v3, uint32_t (&hi)[4], uint32_t (&x)[4]) {
    hi[0] = tf32_rna(v0); hi[1] = tf32_rna(v1); hi[2] = tf32_rna(v2); hi[3] = tf32_rna(v3);
    x[0] = pack_f16((v0 - u2f(hi[0])) * F16_UP, (v2 - u2f(hi[2])) * F16_UP);
    x[1] = pack_f16((v1 - u2f(hi[1])) * F16_UP, (v3 - u2f(hi[3])) * F16_UP);
    x[2] = pack_f16(v0, v2);
    x[3] = pack_f16(v1, v3);
}
// the same MMA shape with fp16 inputs
PDL_DEV void warp_mma_f16(float (&d)[PDL_NL][4], const uint32_t (&a)[PDL_NL][4], const uint32_t (&b)[PDL_NL][2]) {
#ifdef PDL_EMULATE
    float A[16][16], B[16][8];
    for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        for (int e = 0; e < 2; ++e) {
            const int sh = 16 * e;
            A[g][2 * t + e] = f16_bits_to_f32((a[l][0] >> sh) & 0xFFFFu); A[g + 8][2 * t + e] = f16_bits_to_f32((a[l][1] >> sh) & 0xFFFFu);
            A[g][2 * t + 8 + e] = f16_bits_to_f32((a[l][2] >> sh) & 0xFFFFu); A[g + 8][2 * t + 8 + e] = f16_bits_to_f32((a[l][3] >> sh) & 0xFFFFu);
            B[2 * t + e][g] = f16_bits_to_f32((b[l][0] >> sh) & 0xFFFFu); B[2 * t + 8 + e][g] = f16_bits_to_f32((b[l][1] >> sh) & 0xFFFFu);
        }
    }
    for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        for (int q = 0; q < 4; ++q) {
            const int r = g + ((q & 2) ? 8 : 0), c = 2 * t + (q & 1);
            float s = d[l][q];
            for (int k = 0; k < 16; ++k) s += A[r][k] * B[k][c];
            d[l][q] = s;
        }
    }
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0][0]), "+f"(d[0][1]), "+f"(d[0][2]), "+f"(d[0][3])
                 : "r"(a[0][0]), "r"(a[0][1]), "r"(a[0][2]), "r"(a[0][3]), "r"(b[0][0]), "r"(b[0][1]));
#endif
}
// D(16x8) += A(16x16) * B(16x8), bf16 inputs, FP32 accumulate (mma.sync.m16n8k16).  lane = 4g + t:
//   A: a0 (g; k = 2t, 2t+1)  a1 (g+8; 2t, 2t+1)  a2 (g; 2t+8, 2t+9)  a3 (g+8; 2t+8, 2t+9);  B: b0 (k = 2t, 2t+1; n = g)  b1 (k = 2t+8, 2t+9; n = g)
PDL_DEV void warp_mma_bf16(float (&d)[PDL_NL][4], const uint32_t (&a)[PDL_NL][4], const uint32_t (&b)[PDL_NL][2]) {
#ifdef PDL_EMULATE
    float A[16][16], B[16][8];
    for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        for (int e = 0; e < 2; ++e) {
            const int sh = 16 * e;
            A[g][2 * t + e] = u2f(((a[l][0] >> sh) & 0xFFFFu) << 16); A[g + 8][2 * t + e] = u2f(((a[l][1] >> sh) & 0xFFFFu) << 16);
            A[g][2 * t + 8 + e] = u2f(((a[l][2] >> sh) & 0xFFFFu) << 16); A[g + 8][2 * t + 8 + e] = u2f(((a[l][3] >> sh) & 0xFFFFu) << 16);
            B[2 * t + e][g] = u2f(((b[l][0] >> sh) & 0xFFFFu) << 16); B[2 * t + 8 + e][g] = u2f(((b[l][1] >> sh) & 0xFFFFu) << 16);
        }
    }
    for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        for (int q = 0; q < 4; ++q) {
            const int r = g + ((q & 2) ? 8 : 0), c = 2 * t + (q & 1);
            float s = d[l][q];
            for (int k = 0; k < 16; ++k) s += A[r][k] * B[k][c];
            d[l][q] = s;
        }
    }
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0][0]), "+f"(d[0][1]), "+f"(d[0][2]), "+f"(d[0][3])
                 : "r"(a[0][0]), "r"(a[0][1]), "r"(a[0][2]), "r"(a[0][3]), "r"(b[0][0]), "r"(b[0][1]));
#endif
}
PDL_DEV void warp_mma(float (&d)[PDL_NL][4], const uint32_t (&a)[PDL_NL][4], const uint32_t (&b)[PDL_NL][2]) {
#ifdef PDL_EMULATE
    float A[16][8], B[8][8];
    for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        A[g][t] = u2f(a[l][0] & 0xFFFFE000u); A[g + 8][t] = u2f(a[l][1] & 0xFFFFE000u);
        A[g][t + 4] = u2f(a[l][2] & 0xFFFFE000u); A[g + 8][t + 4] = u2f(a[l][3] & 0xFFFFE000u);
        B[t][g] = u2f(b[l][0] & 0xFFFFE000u); B[t + 4][g] = u2f(b[l][1] & 0xFFFFE000u);
    }
    for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        for (int q = 0; q < 4; ++q) {
            const int r = g + ((q & 2) ? 8 : 0), c = 2 * t + (q & 1);
            float s = d[l][q];
            for (int k = 0; k < 8; ++k) s += A[r][k] * B[k][c];
            d[l][q] = s;
        }
    }
#else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0][0]), "+f"(d[0][1]), "+f"(d[0][2]), "+f"(d[0][3])
                 : "r"(a[0][0]), "r"(a[0][1]), "r"(a[0][2]), "r"(a[0][3]), "r"(b[0][0]), "r"(b[0][1]));
#endif
}
// neuron owned by lane-column t as element e (0/1) of neuron tile j, and its local index i = 2j + e
PDL_DEV int own_neuron(int t, int i) { return 8 * (i >> 1) + 2 * t + (i & 1); }

#ifndef PDL_TANH_CACHE
#define PDL_TANH_CACHE 0
#endif
// tanh nets: keep tanh(z0) of the layer whose adjoint comes next in a register per neuron (the forward pass / the recompute of the
// layer below evaluates it anyway), so the adjoint's activation derivatives cost the polynomials only
constexpr bool TANH_CACHE = (PDL_ACT == 0) && (PDL_TANH_CACHE != 0);

// forward activation of one neuron: jets z -> jets h; tc receives tanh(z0) (tanh nets)
PDL_DEV void act_forward(const float (&z)[J], const float* rp, float (&h)[J], float& tc) {
    float s[NS], r[NS], inv;
    sigmas(z[0], rp, s, r, inv);
    tc = s[0];
    pdl_act_fwd(z, s, h);
}
// adjoint of one neuron's activation: (z, hb) -> zb ; accumulates rational coefficient gradients; tc = cached tanh(z0) if TANH_CACHE
PDL_DEV void act_adjoint(const float (&z)[J], const float (&hb)[J], const float* rp, float (&zb)[J], float (&ga)[7], const float tc) {
    float s[NS], r[NS], inv, sb[NS - 1];
#if PDL_ACT == 0
    if (TANH_CACHE) { (void)r; inv = 0.f; pdl_tanh_sigmas(tc, s); }
    else sigmas(z[0], rp, s, r, inv);
#else
    (void)tc;
    sigmas(z[0], rp, s, r, inv);
#endif
    pdl_act_adj(z, hb, s, zb, sb);
    float z0b = 0.f;
#pragma unroll
    for (int m = 0; m < NS - 1; ++m) z0b += sb[m] * s[m + 1];
    zb[0] = z0b;
#if PDL_ACT == 1
    rational_param_adjoint(z[0], rp, r, inv, sb, ga);
#else
    (void)ga;
#endif
}

// both at once for one neuron of the layer below: h = act(z) (forward, recomputed) and zb = act'(z, hb) (adjoint), sharing the
// activation derivatives s[] -- the tanh / the Rational recurrence is evaluated once
PDL_DEV void act_forward_adjoint(const float (&z)[J], const float (&hb)[J], const float* rp, float (&h)[J], float (&zb)[J], float (&ga)[7]) {
    float s[NS], r[NS], inv, sb[NS - 1];
    sigmas(z[0], rp, s, r, inv);
    pdl_act_fwd(z, s, h);
    pdl_act_adj(z, hb, s, zb, sb);
    float z0b = 0.f;
#pragma unroll
    for (int m = 0; m < NS - 1; ++m) z0b += sb[m] * s[m + 1];
    zb[0] = z0b;
#if PDL_ACT == 1
    rational_param_adjoint(z[0], rp, r, inv, sb, ga);
#else
    (void)ga;
#endif
}

// ---------------------------------------------------------------------------------------------
// CTA prologue: stage parameters into shared memory (both GEMM layouts), zero the accumulators
// ---------------------------------------------------------------------------------------------
PDL_DEV void stage_thread(const Params& P, float* sm, int tid) {
    for (int e = tid; e < SM_TOTAL; e += NT) sm[e] = 0.f;
}
PDL_DEV void stage_thread2(const Params& P, float* sm, int tid) {
    // layer 0: W [w1][D], b [w1]
    const int w1 = pdl_width(1);
    for (int e = tid; e < w1 * D; e += NT) sm[SM_W0 + e] = P.prm[0][e];
    for (int e = tid; e < w1; e += NT) sm[SM_B0 + e] = P.prm[1][e];
    // hidden -> hidden layers l = 1..H-1 : W [w_{l+1}][w_l]
    for (int l = 1; l < H; ++l) {
        const int win = pdl_width(l), wout = pdl_width(l + 1);
        float* base = sm + SM_HH + (l - 1) * HH_SIZE;
        const float* W = P.prm[2 * l];
        for (int e = tid; e < wout * win; e += NT) {
            const int n = e / win, k = e - n * win;
            const float w = W[e];
            const uint32_t uh = tf32_rna(w), ul = tf32_rna(w - u2f(uh));   // staged once per CTA: round the low part too
            base[HH_WHI + n * SF + k] = u2f(uh);
            if (!XBF_F) base[HH_WLO + n * SF + k] = u2f(ul);
            else if ((k & 1) == 0) {                                       // packed bf16 cross operands of the pair (k, k+1)
                const float w1 = (k + 1 < win) ? W[e + 1] : 0.f;
                if (XF16_F) {
                    base[HH_WLO + n * SF + k] = u2f(pack_f16(w * F16_DOWN, w1 * F16_DOWN));
                    base[HH_WLO + n * SF + k + 1] = u2f(pack_f16(w - u2f(uh), w1 - u2f(tf32_rna(w1))));
                } else {
                    base[HH_WLO + n * SF + k] = u2f(pack_bf16(w, w1));
                    base[HH_WLO + n * SF + k + 1] = u2f(pack_bf16(w - u2f(uh), w1 - u2f(tf32_rna(w1))));
                }
            }
            if (DGRAD_T) base[HH_THI + k * SF + n] = u2f(uh);
            if (DGRAD_T && !XBF_D) base[HH_TLO + k * SF + n] = u2f(ul);
            if (DGRAD_MMA && XBF_D && (n & 1) == 0) {                        // transposed cross plane: pairs along n
                const float w1 = (n + 1 < wout) ? W[e + win] : 0.f;
                base[HH_TLO + k * SF + n] = u2f(pack_bf16(w, w1));
                base[HH_TLO + k * SF + n + 1] = u2f(pack_bf16(w - u2f(uh), w1 - u2f(tf32_rna(w1))));
            }
            if (!DGRAD_MMA) {
                // FP32-pipe dgrad copy grouped by the lane column that owns input neuron k: t = (k%8)/2, i = 2*(k/8) + k%2
                base[HH_WD + n * SD + ((k & 7) >> 1) * GS + 2 * (k >> 3) + (k & 1)] = w;
            }
        }
        for (int e = tid; e < wout; e += NT) base[HH_B + e] = P.prm[2 * l + 1][e];
    }
    // output layer: W [1][w_H], b [1]
    const int wH = pdl_width(H);
    for (int e = tid; e < wH; e += NT) sm[SM_WH + e] = P.prm[2 * H][e];
    if (tid == 0) sm[SM_BH] = P.prm[2 * H + 1][0];
#if PDL_ACT == 1
    for (int e = tid; e < H * 7; e += NT) {
        const int l = e / 7, q = e - l * 7;
        sm[SM_RAT + l * 8 + q] = (q < 4) ? P.prm[2 * (H + 1) + 2 * l][q] : P.prm[2 * (H + 1) + 2 * l + 1][q - 4];
    }
#endif
    for (int e = tid; e < K; e += NT) {
        const bool off = P.mask[e] != 0;
        sm[SM_XI + e] = off ? 0.f : P.xi[e];
        sm[SM_GM + e] = off ? 0.f : 1.f;
    }
}

// running per-warp partial sum of a weight-gradient fragment in the warp's private global buffer: ONE fire-and-forget vector reduction
// at the L2 (red.global.v4.f32.add) per fragment and tile instead of load + rounded add + store.  The buffer is private to the warp and
// a thread's reductions to one address are performed in program order, so the sum is the same sequence of rounded additions (launches
// stay bit-identical: tested) -- but the kernel no longer waits for 32 KB of partial sums per tile to come back from the L2 and holds
// 19 registers less (heat: 255 with 44 B of spills -> 236, none).  B200, 2^20 rows (profiles/r02_ab_wgrad_red.txt): heat 2.95 -> 2.80 ms,
// KdV 3.82 -> 3.71, KS 4.68 -> 4.58, rd2d 4.11 -> 4.00, heat Rational 3.37 -> 3.29, KS Rational unchanged.  PDL_WGRAD_RED=0: the old form.
#ifndef PDL_WGRAD_RED
#define PDL_WGRAD_RED 1
#endif
#if !defined(PDL_EMULATE) && PDL_WGRAD_RED
PDL_DEV void red4(float* q, pdl_f4 r, bool first) {
    if (first) { stg4(q, r); return; }
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------
// the per-warp tile loop
// ---------------------------------------------------------------------------------------------
template <bool BWD, bool DEVN>
PDL_DEV void warp_body(const Params& P, float* sm, int warp_in_cta, int gwarp, int n_warps
#ifndef PDL_EMULATE
                       , const int lane
#endif
) {
    float* bufH = sm + SM_SHARED + warp_in_cta * SM_WARP;
    float* bufZ = bufH + BUF;
    float* sacc = bufZ + BUF;
    float* zs = BWD ? (P.zscratch + (size_t)gwarp * ZSCR) : nullptr;
    float* wp = BWD ? (P.wpriv + (size_t)gwarp * WPRIV) : nullptr;
    // device-side row count (captured epochs: the number of targeted points changes from replay to replay)
    // (a separate instantiation: with the count in a kernel parameter the compiler keeps it in the constant bank, with the count
    // loaded from memory it costs four live registers in a kernel that has none to spare -- measured 8 % on heat)
    long long NPTS = P.n_points;
    double INVN = P.inv_n;
    if (DEVN) {
        NPTS = *P.n_points_dev;
        if (NPTS < 0) NPTS = 0;
        if (NPTS > P.n_points) NPTS = P.n_points;
        INVN = NPTS > 0 ? 1.0 / (double)NPTS : 0.0;
    }
    const long long n_tiles = (NPTS + 7) / 8;
    const float two_inv_n = (float)(2.0 * INVN);

    double lsum[PDL_NL], asum[PDL_NL];
    PDL_FOR_LANES { lsum[PDL_L] = 0.0; asum[PDL_L] = 0.0; }

    float x[PDL_NL][D];
    float h[PDL_NL][NPL][J];          // activation jets of the lane's neurons (layer being produced)
    float hb[PDL_NL][NPL][J];         // adjoint jets flowing backwards
    float uu[PDL_NL][J];
    float ub[PDL_NL][J];
    float valid[PDL_NL];
    float tcache[PDL_NL][NPL];        // tanh(z0) of the layer whose adjoint comes next (TANH_CACHE)
    bool first = true;

    for (long long tile = gwarp; tile < n_tiles; tile += n_warps, first = false) {
        // ---------------- layer 0 forward ----------------
        PDL_FOR_LANES {
            const int pg = lane >> 2, g4 = lane & 3;
            const long long p = tile * 8 + pg;
            valid[PDL_L] = (p < NPTS) ? 1.f : 0.f;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) x[PDL_L][dd] = (p < NPTS) ? P.points[p * D + dd] : 0.f;
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                float z[J];
                layer0_preact(sm, own_neuron(g4, i), x[PDL_L], z);
                act_forward(z, sm + SM_RAT, h[PDL_L][i], tcache[PDL_L][i]);
            }
        }
        // ---------------- hidden -> hidden layers, forward ----------------
        for (int l = 1; l < H; ++l) {
            // 3xTF32 on the tensor pipe: accumulator fragments [neuron tile][component pair][lane][2*(c&1) + (i&1)]
            const float* Whi = sm + SM_HH + (l - 1) * HH_SIZE + HH_WHI;
            const float* Wlo = sm + SM_HH + (l - 1) * HH_SIZE + HH_WLO;
            const float* bl = sm + SM_HH + (l - 1) * HH_SIZE + HH_B;
            float accf[NT8][NM][PDL_NL][4];
            PDL_FOR_LANES {
                const int t = lane & 3;
#pragma unroll
                for (int q = 0; q < NT8; ++q) {
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        accf[q][m][PDL_L][0] = 0.f; accf[q][m][PDL_L][1] = 0.f; accf[q][m][PDL_L][2] = 0.f; accf[q][m][PDL_L][3] = 0.f;
                    }
                    accf[q][0][PDL_L][0] = bl[8 * q + 2 * t];            // bias enters component 0 only
                    accf[q][0][PDL_L][1] = bl[8 * q + 2 * t + 1];
                }
            }
#pragma unroll
            for (int s = 0; s < NT8; ++s) {
                uint32_t ahi[NM][PDL_NL][4], alo[NM][PDL_NL][4];        // alo = the bf16 cross fragment when XBF_F
                PDL_FOR_LANES {
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        // a0 (comp 2m, neuron 8s+2t)  a1 (comp 2m+1, same neuron)  a2 (comp 2m, neuron 8s+2t+1)  a3 (comp 2m+1, ...)
                        const float v0 = h[PDL_L][2 * s][2 * m], v2 = h[PDL_L][2 * s + 1][2 * m];
                        const float v1 = (2 * m + 1 < J) ? h[PDL_L][2 * s][(2 * m + 1 < J) ? 2 * m + 1 : 0] : 0.f;
                        const float v3 = (2 * m + 1 < J) ? h[PDL_L][2 * s + 1][(2 * m + 1 < J) ? 2 * m + 1 : 0] : 0.f;
                        if (XF16_F) split_cross_f16(v0, v1, v2, v3, ahi[m][PDL_L], alo[m][PDL_L]);
                        else if (XBF_F) split_cross(v0, v1, v2, v3, ahi[m][PDL_L], alo[m][PDL_L]);
                        else {
                            split_tf32(v0, ahi[m][PDL_L][0], alo[m][PDL_L][0]);
                            split_tf32(v1, ahi[m][PDL_L][1], alo[m][PDL_L][1]);
                            split_tf32(v2, ahi[m][PDL_L][2], alo[m][PDL_L][2]);
                            split_tf32(v3, ahi[m][PDL_L][3], alo[m][PDL_L][3]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < NT8; ++q) {
                    uint32_t bhi[PDL_NL][2], blo[PDL_NL][2];
                    PDL_FOR_LANES {
                        const int g = lane >> 2, t = lane & 3;
                        const pdl_f2 vh = ld2(Whi + (8 * q + g) * SF + 8 * s + 2 * t);
                        const pdl_f2 vl = ld2(Wlo + (8 * q + g) * SF + 8 * s + 2 * t);
                        bhi[PDL_L][0] = f2u(vh.x); bhi[PDL_L][1] = f2u(vh.y); blo[PDL_L][0] = f2u(vl.x); blo[PDL_L][1] = f2u(vl.y);
                    }
                    if (XBF_F) {
#pragma unroll
                        for (int m = 0; m < NM; ++m) warp_mma(accf[q][m], ahi[m], bhi);
#pragma unroll
                        for (int m = 0; m < NM; ++m) { if (XF16_F) warp_mma_f16(accf[q][m], alo[m], blo); else warp_mma_bf16(accf[q][m], alo[m], blo); }
                    } else {
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            warp_mma(accf[q][m], alo[m], bhi);
                            warp_mma(accf[q][m], ahi[m], blo);
                            warp_mma(accf[q][m], ahi[m], bhi);
                        }
                    }
                }
            }
            PDL_FOR_LANES {
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    float z[J];
#pragma unroll
                    for (int c = 0; c < J; ++c) z[c] = accf[i >> 1][c >> 1][PDL_L][2 * (c & 1) + (i & 1)];
                    if (BWD) scratch_store(zs, l - 1, i, lane, z);
                    act_forward(z, sm + SM_RAT + l * 8, h[PDL_L][i], tcache[PDL_L][i]);
                }
            }
        }
        // ---------------- output layer (identity activation) ----------------
        PDL_FOR_LANES {
            const int g4 = lane & 3;
#pragma unroll
            for (int c = 0; c < J; ++c) uu[PDL_L][c] = 0.f;
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                const float w = sm[SM_WH + own_neuron(g4, i)];
#pragma unroll
                for (int c = 0; c < J; ++c) uu[PDL_L][c] += w * h[PDL_L][i][c];
            }
        }
        xor_add<J>(uu, 1);
        xor_add<J>(uu, 2);
        // ---------------- library terms, residual, loss, seeds ----------------
        PDL_FOR_LANES {
            const int pg = lane >> 2, g4 = lane & 3;
            const long long p = tile * 8 + pg;
            float u[J];
#pragma unroll
            for (int c = 0; c < J; ++c) u[c] = uu[PDL_L][c];
            u[0] += sm[SM_BH];
            float r, du[J], T[KX];
            pdl_epilogue(u, sm + SM_XI, r, du, T);
            double rd = (double)r;
#if PDL_MODE == 1
            rd = (p < NPTS) ? ((double)r - P.targets[p]) : 0.0;      // f32 prediction - f64 target (Loss.py:56-59)
            r = (float)rd;
#endif
            const float v = valid[PDL_L];
            if (g4 == 0 && v != 0.f) {
                if (P.residual) P.residual[p] = r;
                if (P.features) {
#pragma unroll
                    for (int c = 0; c < J; ++c) P.features[(size_t)c * P.n_points + p] = u[c];
                }
                lsum[PDL_L] += rd * rd;
                asum[PDL_L] += fabs(rd);
            }
            if (BWD) {
#if PDL_MODE == 1
                const float rs = (float)(2.0 * INVN * rd) * v;
#else
                const float rs = two_inv_n * r * v;
#endif
#pragma unroll
                for (int c = 0; c < J; ++c) ub[PDL_L][c] = rs * du[c];
                if (g4 == 0) {
#pragma unroll
                    for (int k = 0; k < K; ++k) sacc[SO_XI + k * 8 + pg] -= rs * T[k] * sm[SM_GM + k];
                }
            }
        }
        if (!BWD) continue;
        // ---------------- output layer adjoint ----------------
        {
            float g[PDL_NL][NPL + 1];
            PDL_FOR_LANES {
                const int g4 = lane & 3;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    float t = 0.f;
#pragma unroll
                    for (int c = 0; c < J; ++c) t += ub[PDL_L][c] * h[PDL_L][i][c];
                    g[PDL_L][i] = t;
                    const float w = sm[SM_WH + own_neuron(g4, i)];
#pragma unroll
                    for (int c = 0; c < J; ++c) hb[PDL_L][i][c] = w * ub[PDL_L][c];
                }
                g[PDL_L][NPL] = ub[PDL_L][0];
            }
            xor_add<NPL + 1>(g, 4); xor_add<NPL + 1>(g, 8); xor_add<NPL + 1>(g, 16);
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
                if (pg == 0) {
#pragma unroll
                    for (int i = 0; i < NPL; ++i) sacc[SO_WH + own_neuron(g4, i)] += g[PDL_L][i];
                    if (g4 == 0) sacc[SO_BH] += g[PDL_L][NPL];
                }
            }
        }
        // ---------------- hidden -> hidden layers, adjoint ----------------
        // Order per layer l (top down): data gradient of layer l (zb_l -> hb_{l-1}), then ONE pass over the lane's neurons of layer
        // l-1 that evaluates the activation derivatives once and uses them twice -- h_{l-1} = act(z_{l-1}) (the weight gradient's
        // other operand, recomputed instead of stored) and zb_{l-1} = act'(z_{l-1}, hb_{l-1}) in place -- then the weight gradient
        // of layer l.  hb[] therefore holds hb_l or zb_l alternately; bufZ holds zb_l for the shared-memory consumers.
        auto wgrad_layer = [&](const int l) {
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3, g8 = lane >> 2;
                // ---- weight gradient: Wbar[n][k] += sum_p sum_c zb[p][n][c] * h[p][k][c]   (n in g8 group, k in g4 group)
                if (!WGRAD_MMA) {
                    float acc[NACC4];
                    float* wpl = wp + (l - 1) * WPRIV_LAYER;
                    if (first) {
#pragma unroll
                        for (int a = 0; a < NACC4; ++a) acc[a] = 0.f;
                    } else {
#pragma unroll
                        for (int a = 0; a < NACC4; a += 4) {
                            const pdl_f4 t = ldg4(wpl + (a / 4 * 32 + lane) * 4);
                            acc[a] = t.x; acc[a + 1] = t.y; acc[a + 2] = t.z; acc[a + 3] = t.w;
                        }
                    }
#pragma unroll 2
                    for (int q = 0; q < 8; ++q) {
                        float zq[NPN][J], hq[NPL][J];
#pragma unroll
                        for (int a = 0; a < NPN; ++a) { const int n = g8 * NPN + a; load_entry(bufZ + q * SA + n * JA, bufZ + 8 * SA + q * SB + n * JB, zq[a]); }
#pragma unroll
                        for (int b = 0; b < NPL; ++b) { const int k = g4 * NPL + b; load_entry(bufH + q * SA + k * JA, bufH + 8 * SA + q * SB + k * JB, hq[b]); }
#pragma unroll
                        for (int a = 0; a < NPN; ++a) {
#pragma unroll
                            for (int b = 0; b < NPL; ++b) {
#pragma unroll
                                for (int c = 0; c < J; ++c) acc[a * NPL + b] += zq[a][c] * hq[b][c];
                            }
                            acc[NPN * NPL + a] += zq[a][0];
                        }
                    }
#pragma unroll
                    for (int a = 0; a < NACC4; a += 4) {
                        pdl_f4 t; t.x = acc[a]; t.y = acc[a + 1]; t.z = acc[a + 2]; t.w = acc[a + 3];
                        stg4(wpl + (a / 4 * 32 + lane) * 4, t);
                    }
                }
            }
            if (WGRAD_MMA) {
                // ---- weight gradient on the tensor pipe.  Wbar[n][k] = sum_c sum_p zb[p][n][c] h[p][k][c]: M = n (m16 tiles), N = k
                // (n8 tiles), one k8-step per jet component with the tile's 8 points as the summation index (slot t <-> point 2t,
                // slot t+4 <-> point 2t+1: conflict-free 16-byte reads of the [point][neuron][comp] buffers).  Each C fragment is
                // accumulated from zero on the tensor pipe (which truncates) and added to the running FP32 partial sum with a
                // rounded FADD, so the long sum over tiles keeps round-to-nearest behaviour.
                float* wpl = wp + (l - 1) * WPRIV_LAYER;
                uint32_t bhi[NTB][J][PDL_NL][2], blo[NTB][J][PDL_NL][2];
                auto make_b = [&](const int slot, const int nt) {           // B fragments of n8 tile nt -> slot (blo = bf16 cross fragment when XBF_W)
                    PDL_FOR_LANES {
                        const int g = lane >> 2, t = lane & 3;
                        const int k = 8 * nt + g;
                        float w0[J], w1[J];
                        load_entry(bufH + (2 * t) * SA + k * JA, bufH + 8 * SA + (2 * t) * SB + k * JB, w0);
                        load_entry(bufH + (2 * t + 1) * SA + k * JA, bufH + 8 * SA + (2 * t + 1) * SB + k * JB, w1);
#pragma unroll
                        for (int c = 0; c < J; ++c) {
                            if (XBF_W) {                          // k slots 2t, 2t+1 = points 2t, 2t+1
                                const uint32_t h0 = tf32_rna(w0[c]), h1 = tf32_rna(w1[c]);
                                bhi[slot][c][PDL_L][0] = h0; bhi[slot][c][PDL_L][1] = h1;
                                blo[slot][c][PDL_L][0] = pack_bf16(w0[c], w1[c]);
                                blo[slot][c][PDL_L][1] = pack_bf16(w0[c] - u2f(h0), w1[c] - u2f(h1));
                            } else {
                                split_tf32_raw(w0[c], bhi[slot][c][PDL_L][0], blo[slot][c][PDL_L][0]);
                                split_tf32_raw(w1[c], bhi[slot][c][PDL_L][1], blo[slot][c][PDL_L][1]);
                            }
                        }
                    }
                };
                float bsel[PDL_NL][2];
                PDL_FOR_LANES { bsel[PDL_L][0] = 0.f; bsel[PDL_L][1] = 0.f; }
#if PDL_WGRAD_ROLL_MT
#pragma unroll 1
#else
#pragma unroll
#endif
                for (int mt = 0; mt < NMT; ++mt) {
                    uint32_t ahi[J][PDL_NL][4], alo[J][PDL_NL][4];
                    float bs[PDL_NL][2];
                    PDL_FOR_LANES {
                        const int g = lane >> 2, t = lane & 3;
                        const int n0 = 16 * mt + g, n1 = n0 + 8;
                        float v[4][J];
                        load_entry(bufZ + (2 * t) * SA + n0 * JA, bufZ + 8 * SA + (2 * t) * SB + n0 * JB, v[0]);
                        load_entry(bufZ + (2 * t + 1) * SA + n0 * JA, bufZ + 8 * SA + (2 * t + 1) * SB + n0 * JB, v[2]);
                        if (16 * mt + 8 < WP) {
                            load_entry(bufZ + (2 * t) * SA + n1 * JA, bufZ + 8 * SA + (2 * t) * SB + n1 * JB, v[1]);
                            load_entry(bufZ + (2 * t + 1) * SA + n1 * JA, bufZ + 8 * SA + (2 * t + 1) * SB + n1 * JB, v[3]);
                        } else {
#pragma unroll
                            for (int c = 0; c < J; ++c) { v[1][c] = 0.f; v[3][c] = 0.f; }
                        }
#pragma unroll
                        for (int c = 0; c < J; ++c) {
                            if (XBF_W) split_cross(v[0][c], v[1][c], v[2][c], v[3][c], ahi[c][PDL_L], alo[c][PDL_L]);
                            else {
#pragma unroll
                                for (int q = 0; q < 4; ++q) split_tf32_raw(v[q][c], ahi[c][PDL_L][q], alo[c][PDL_L][q]);
                            }
                        }
                        bs[PDL_L][0] = v[0][0] + v[2][0];                 // bias gradient: sum over points of component 0
                        bs[PDL_L][1] = v[1][0] + v[3][0];
                    }
                    xor_add<2>(bs, 1); xor_add<2>(bs, 2);
                    PDL_FOR_LANES { if ((lane & 3) == mt) { bsel[PDL_L][0] = bs[PDL_L][0]; bsel[PDL_L][1] = bs[PDL_L][1]; } }
#pragma unroll
                    for (int nt0 = 0; nt0 < NT8; nt0 += NTB) {
                        // NTB n8 tiles at a time, one accumulator per MMA kind (lo*hi, hi*lo, hi*hi -- or hi*hi and the bf16 cross MMA):
                        // NACW*NTB independent MMA chains of length J
                        constexpr int NACW = XBF_W ? 2 : 3;
                        float c4[NTB][NACW][PDL_NL][4];
#pragma unroll
                        for (int b = 0; b < NTB; ++b) {
                            if (nt0 + b < NT8) make_b(b, nt0 + b);
                            PDL_FOR_LANES {
#pragma unroll
                                for (int q = 0; q < NACW; ++q) { c4[b][q][PDL_L][0] = 0.f; c4[b][q][PDL_L][1] = 0.f; c4[b][q][PDL_L][2] = 0.f; c4[b][q][PDL_L][3] = 0.f; }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < J; ++c) {
#pragma unroll
                            for (int b = 0; b < NTB; ++b) {
                                if (nt0 + b < NT8) {
                                    if (XBF_W) {
                                        warp_mma(c4[b][0], ahi[c], bhi[b][c]);
                                        warp_mma_bf16(c4[b][1], alo[c], blo[b][c]);
                                    } else {
                                        warp_mma(c4[b][0], alo[c], bhi[b][c]);
                                        warp_mma(c4[b][1], ahi[c], blo[b][c]);
                                        warp_mma(c4[b][NACW - 1], ahi[c], bhi[b][c]);
                                    }
                                }
                            }
                        }
                        PDL_FOR_LANES {
#pragma unroll
                            for (int b = 0; b < NTB; ++b) {
                                if (nt0 + b < NT8) {
                                    float* q = wpl + ((mt * NT8 + nt0 + b) * 32 + lane) * 4;
                                    pdl_f4 r;
                                    if (XBF_W) {
                                        r.x = c4[b][0][PDL_L][0] + c4[b][1][PDL_L][0];
                                        r.y = c4[b][0][PDL_L][1] + c4[b][1][PDL_L][1];
                                        r.z = c4[b][0][PDL_L][2] + c4[b][1][PDL_L][2];
                                        r.w = c4[b][0][PDL_L][3] + c4[b][1][PDL_L][3];
                                    } else {
                                        r.x = (c4[b][0][PDL_L][0] + c4[b][1][PDL_L][0]) + c4[b][NACW - 1][PDL_L][0];
                                        r.y = (c4[b][0][PDL_L][1] + c4[b][1][PDL_L][1]) + c4[b][NACW - 1][PDL_L][1];
                                        r.z = (c4[b][0][PDL_L][2] + c4[b][1][PDL_L][2]) + c4[b][NACW - 1][PDL_L][2];
                                        r.w = (c4[b][0][PDL_L][3] + c4[b][1][PDL_L][3]) + c4[b][NACW - 1][PDL_L][3];
                                    }
#if !defined(PDL_EMULATE) && PDL_WGRAD_RED
                                    red4(q, r, first);
#else
                                    if (!first) { const pdl_f4 o = ldg4(q); r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
                                    stg4(q, r);
#endif
                                }
                            }
                        }
                    }
                }
                PDL_FOR_LANES {                                            // bias slot: lane (g, t) owns neurons 16t+g and 16t+g+8
                    float* q = wpl + (NFRAG * 32 + lane) * 4;
                    pdl_f4 r; r.x = bsel[PDL_L][0]; r.y = bsel[PDL_L][1]; r.z = 0.f; r.w = 0.f;
#if !defined(PDL_EMULATE) && PDL_WGRAD_RED
                    red4(q, r, first);
#else
                    if (!first) { const pdl_f4 o = ldg4(q); r.x += o.x; r.y += o.y; }
                    stg4(q, r);
#endif
                }
            }
        };
        auto dgrad_layer = [&](const int l) {
            const float* Wd = sm + SM_HH + (l - 1) * HH_SIZE + HH_WD;
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
                // ---- data gradient on the FP32 pipe: hb[p][k] = sum_n zb[p][n] * W[n][k]   (k = the lane's own neurons)
                if (!DGRAD_MMA) {
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
#pragma unroll
                    for (int c = 0; c < J; ++c) hb[PDL_L][i][c] = 0.f;
                }
                const float* zA = bufZ + pg * SA;
                const float* zB = bufZ + 8 * SA + pg * SB;
#pragma unroll 2
                for (int nc = 0; nc < WP; nc += 4) {
                    float a[4][J];
#pragma unroll
                    for (int nn = 0; nn < 4; ++nn) load_entry(zA + (nc + nn) * JA, zB + (nc + nn) * JB, a[nn]);
#pragma unroll
                    for (int nn = 0; nn < 4; ++nn) {
                        const float* wr = Wd + (nc + nn) * SD + g4 * GS;
                        float w[NPL];
#pragma unroll
                        for (int i = 0; i + 3 < NPL; i += 4) { const pdl_f4 t = ld4(wr + i); w[i] = t.x; w[i + 1] = t.y; w[i + 2] = t.z; w[i + 3] = t.w; }
                        if ((NPL & 3) >= 2) { const pdl_f2 t = ld2(wr + (NPL & ~3)); w[(NPL & ~3)] = t.x; w[(NPL & ~3) + 1 < NPL ? (NPL & ~3) + 1 : 0] = t.y; }
                        if (NPL & 1) w[NPL - 1] = wr[NPL - 1];
#pragma unroll
                        for (int i = 0; i < NPL; ++i) {
#pragma unroll
                            for (int c = 0; c < J; ++c) hb[PDL_L][i][c] += a[nn][c] * w[i];
                        }
                    }
                }
                }
            }
            if (DGRAD_MMA) {
                // ---- data gradient on the tensor pipe: hb = zb W, A fragments straight from the registers, B = W^T hi/lo
                const float* Thi = sm + SM_HH + (l - 1) * HH_SIZE + HH_THI;
                const float* Tlo = sm + SM_HH + (l - 1) * HH_SIZE + HH_TLO;
                const float* Whi_ = sm + SM_HH + (l - 1) * HH_SIZE + HH_WHI;
                const float* Wlo_ = sm + SM_HH + (l - 1) * HH_SIZE + HH_WLO;
                float hbfr[NT8][NM][PDL_NL][4];
                PDL_FOR_LANES {
#pragma unroll
                    for (int q = 0; q < NT8; ++q) {
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            hbfr[q][m][PDL_L][0] = 0.f; hbfr[q][m][PDL_L][1] = 0.f; hbfr[q][m][PDL_L][2] = 0.f; hbfr[q][m][PDL_L][3] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < NT8; ++s) {
                    uint32_t ahi[NM][PDL_NL][4], alo[NM][PDL_NL][4];
                    PDL_FOR_LANES {
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            const float v0 = hb[PDL_L][2 * s][2 * m], v2 = hb[PDL_L][2 * s + 1][2 * m];
                            const float v1 = (2 * m + 1 < J) ? hb[PDL_L][2 * s][(2 * m + 1 < J) ? 2 * m + 1 : 0] : 0.f;
                            const float v3 = (2 * m + 1 < J) ? hb[PDL_L][2 * s + 1][(2 * m + 1 < J) ? 2 * m + 1 : 0] : 0.f;
                            if (XBF_D) split_cross(v0, v1, v2, v3, ahi[m][PDL_L], alo[m][PDL_L]);
                            else {
                                split_tf32_dg(v0, ahi[m][PDL_L][0], alo[m][PDL_L][0]);
                                split_tf32_dg(v1, ahi[m][PDL_L][1], alo[m][PDL_L][1]);
                                split_tf32_dg(v2, ahi[m][PDL_L][2], alo[m][PDL_L][2]);
                                split_tf32_dg(v3, ahi[m][PDL_L][3], alo[m][PDL_L][3]);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NT8; ++q) {
                        uint32_t bhi[PDL_NL][2], blo[PDL_NL][2];
                        PDL_FOR_LANES {
                            const int g = lane >> 2, t = lane & 3;
                            if (DGRAD_T) {
                                const pdl_f2 vh = ld2(Thi + (8 * q + g) * SF + 8 * s + 2 * t);
                                bhi[PDL_L][0] = f2u(vh.x); bhi[PDL_L][1] = f2u(vh.y);
                            } else {                            // B[n][k] = W[n][k] straight from the forward copies
                                const int o = (8 * s + 2 * t) * SF + 8 * q + g;
                                bhi[PDL_L][0] = f2u(Whi_[o]); bhi[PDL_L][1] = f2u(Whi_[o + SF]);
                                if (!XBF_D) { blo[PDL_L][0] = f2u(Wlo_[o]); blo[PDL_L][1] = f2u(Wlo_[o + SF]); }
                            }
                            if (DGRAD_T || XBF_D) {               // lo plane of W^T, or its packed bf16 cross plane (the only transposed plane in mode 2)
                                const pdl_f2 vl = ld2(Tlo + (8 * q + g) * SF + 8 * s + 2 * t);
                                blo[PDL_L][0] = f2u(vl.x); blo[PDL_L][1] = f2u(vl.y);
                            }
                        }
                        if (XBF_D) {
#pragma unroll
                            for (int m = 0; m < NM; ++m) warp_mma(hbfr[q][m], ahi[m], bhi);
#pragma unroll
                            for (int m = 0; m < NM; ++m) warp_mma_bf16(hbfr[q][m], alo[m], blo);
                        } else {
#pragma unroll
                            for (int m = 0; m < NM; ++m) {
                                warp_mma(hbfr[q][m], alo[m], bhi);
                                warp_mma(hbfr[q][m], ahi[m], blo);
                                warp_mma(hbfr[q][m], ahi[m], bhi);
                            }
                        }
                    }
                }
                PDL_FOR_LANES {
#pragma unroll
                    for (int i = 0; i < NPL; ++i) {
#pragma unroll
                        for (int c = 0; c < J; ++c) hb[PDL_L][i][c] = hbfr[i >> 1][c >> 1][PDL_L][2 * (c & 1) + (i & 1)];
                    }
                }
            }
        };
        auto store_zb = [&]() {                      // hb[] (= zb of the current layer) -> bufZ
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int n = own_neuron(g4, i);
                    store_entry(bufZ + pg * SA + n * JA, bufZ + 8 * SA + pg * SB + n * JB, hb[PDL_L][i]);
                }
            }
        };
        auto reduce_rational = [&](float (&ga)[PDL_NL][7], const int layer) {
#if PDL_ACT == 1
            xor_add<7>(ga, 1); xor_add<7>(ga, 2); xor_add<7>(ga, 4); xor_add<7>(ga, 8); xor_add<7>(ga, 16);
            PDL_FOR_LANES { if (lane == 0) { for (int q = 0; q < 7; ++q) sacc[SO_RAT + layer * 7 + q] += ga[PDL_L][q]; } }
#else
            (void)ga; (void)layer;
#endif
        };
        // ---------------- hidden layers adjoint: activation passes (top layer, classic and fused loops) ----------------
        if (ADJ_FUSED && H >= 2) {                   // top hidden layer: zb_{H-1} = act'(z_{H-1}, hb_{H-1}) in place
            float ga[PDL_NL][7];
            PDL_FOR_LANES {
#pragma unroll
                for (int q = 0; q < 7; ++q) ga[PDL_L][q] = 0.f;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    float z[J], zb[J];
                    scratch_load(zs, H - 2, i, lane, z);
                    act_adjoint(z, hb[PDL_L][i], sm + SM_RAT + (H - 1) * 8, zb, ga[PDL_L], tcache[PDL_L][i]);
#pragma unroll
                    for (int c = 0; c < J; ++c) hb[PDL_L][i][c] = zb[c];
                }
            }
            store_zb();                              // read by the weight gradient (after the next __syncwarp) and by the FP32-pipe data gradient
            reduce_rational(ga, H - 1);
            if (!DGRAD_MMA) PDL_SYNCWARP();
        }
        if (!ADJ_FUSED) {
            for (int l = H - 1; l >= 1; --l) {
                float ga[PDL_NL][7];
                PDL_FOR_LANES {
                    const int pg = lane >> 2, g4 = lane & 3;
#pragma unroll
                    for (int q = 0; q < 7; ++q) ga[PDL_L][q] = 0.f;
#pragma unroll
                    for (int i = 0; i < NPL; ++i) {
                        const int n = own_neuron(g4, i);
                        float z[J], zb[J];
                        scratch_load(zs, l - 1, i, lane, z);
                        act_adjoint(z, hb[PDL_L][i], sm + SM_RAT + l * 8, zb, ga[PDL_L], tcache[PDL_L][i]);
                        store_entry(bufZ + pg * SA + n * JA, bufZ + 8 * SA + pg * SB + n * JB, zb);
#pragma unroll
                        for (int c = 0; c < J; ++c) hb[PDL_L][i][c] = zb[c];
                        // activations of the layer below (input of this linear layer), recomputed
                        float zi[J], hi[J];
                        if (l >= 2) scratch_load(zs, l - 2, i, lane, zi);
                        else layer0_preact(sm, n, x[PDL_L], zi);
                        act_forward(zi, sm + SM_RAT + (l - 1) * 8, hi, tcache[PDL_L][i]);
                        store_entry(bufH + pg * SA + n * JA, bufH + 8 * SA + pg * SB + n * JB, hi);
                    }
                }
                PDL_SYNCWARP();
                reduce_rational(ga, l);
                wgrad_layer(l);
                dgrad_layer(l);
                PDL_SYNCWARP();
            }
        }
        for (int l = H - 1; ADJ_FUSED && l >= 1; --l) {
            dgrad_layer(l);                          // hb: zb_l -> hb_{l-1}
            float ga[PDL_NL][7];
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
#pragma unroll
                for (int q = 0; q < 7; ++q) ga[PDL_L][q] = 0.f;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int n = own_neuron(g4, i);
                    float zi[J], hi[J], zb[J];
                    if (l >= 2) scratch_load(zs, l - 2, i, lane, zi);
                    else layer0_preact(sm, n, x[PDL_L], zi);
                    act_forward_adjoint(zi, hb[PDL_L][i], sm + SM_RAT + (l - 1) * 8, hi, zb, ga[PDL_L]);
                    store_entry(bufH + pg * SA + n * JA, bufH + 8 * SA + pg * SB + n * JB, hi);
#pragma unroll
                    for (int c = 0; c < J; ++c) hb[PDL_L][i][c] = zb[c];
                }
            }
            reduce_rational(ga, l - 1);
            PDL_SYNCWARP();
            wgrad_layer(l);                          // bufZ = zb_l, bufH = h_{l-1}
            PDL_SYNCWARP();
            if (l >= 2) { store_zb(); if (!DGRAD_MMA) PDL_SYNCWARP(); }   // the tensor-pipe data gradient reads hb[], not bufZ
        }
        // ---------------- layer 0 adjoint ----------------
        {
            float g0[PDL_NL][NPL * (D + 1)];
            float ga[PDL_NL][7];
            PDL_FOR_LANES {
                const int g4 = lane & 3;
#pragma unroll
                for (int q = 0; q < 7; ++q) ga[PDL_L][q] = 0.f;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int n = own_neuron(g4, i);
                    float zb[J];
                    if (ADJ_FUSED && H >= 2) {         // zb_0 was produced in place by the last pass of the loop above
#pragma unroll
                        for (int c = 0; c < J; ++c) zb[c] = hb[PDL_L][i][c];
                    } else {
                        float z[J];
                        layer0_preact(sm, n, x[PDL_L], z);
                        act_adjoint(z, hb[PDL_L][i], sm + SM_RAT, zb, ga[PDL_L], tcache[PDL_L][i]);
                    }
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        float t = zb[0] * x[PDL_L][dd];
#pragma unroll
                        for (int c = 1; c < J; ++c) if (pdl_comp_axis(c) == dd) t += zb[c];
                        g0[PDL_L][i * (D + 1) + dd] = t;
                    }
                    g0[PDL_L][i * (D + 1) + D] = zb[0];
                }
            }
            xor_add<NPL*(D + 1)>(g0, 4); xor_add<NPL*(D + 1)>(g0, 8); xor_add<NPL*(D + 1)>(g0, 16);
#if PDL_ACT == 1
            if (!(ADJ_FUSED && H >= 2)) { xor_add<7>(ga, 1); xor_add<7>(ga, 2); xor_add<7>(ga, 4); xor_add<7>(ga, 8); xor_add<7>(ga, 16); }
#endif
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
                if (pg == 0) {
#pragma unroll
                    for (int i = 0; i < NPL; ++i) {
                        const int n = own_neuron(g4, i);
#pragma unroll
                        for (int dd = 0; dd < D; ++dd) sacc[SO_W0 + n * D + dd] += g0[PDL_L][i * (D + 1) + dd];
                        sacc[SO_B0 + n] += g0[PDL_L][i * (D + 1) + D];
                    }
                }
#if PDL_ACT == 1
                if (!(ADJ_FUSED && H >= 2) && lane == 0) { for (int q = 0; q < 7; ++q) sacc[SO_RAT + q] += ga[PDL_L][q]; }
#endif
            }
        }
        PDL_SYNCWARP();
    }

    // ---- warp epilogue: a warp that never ran a tile must still publish zeros; loss partials via shuffles ----
    if (BWD && first) {
        PDL_FOR_LANES {
            pdl_f4 zero; zero.x = zero.y = zero.z = zero.w = 0.f;
            for (int e = lane * 4; e < WPRIV; e += 128) stg4(wp + e, zero);
        }
    }
    {
        // loss partials stay in double: park them in the (now free) jet buffer for the CTA tail
        PDL_FOR_LANES {
            double* slot = reinterpret_cast<double*>(bufH);       // bufH is free now: 64 doubles fit easily
            slot[lane] = lsum[PDL_L];
            slot[32 + lane] = asum[PDL_L];
        }
        PDL_SYNCWARP();
    }
}

// CTA tail: sum the warps' partial buffers and export one record per CTA
PDL_DEV void cta_tail_thread(const Params& P, const float* sm, int cta, int tid, bool bwd) {
    if (bwd) {
        float* out = P.cta_out + (size_t)cta * CTA_OUT;
        for (int e = tid; e < WPRIV; e += NT) {
            float s = 0.f;
            for (int w = 0; w < NW; ++w) s += ldg1(P.wpriv + (size_t)(cta * NW + w) * WPRIV + e);
            out[e] = s;
        }
        for (int e = tid; e < SACC_OUT; e += NT) {
            float s = 0.f;
            for (int w = 0; w < NW; ++w) {
                const float* sacc = sm + SM_SHARED + w * SM_WARP + 2 * BUF;
                if (e < SO_XI) s += sacc[e];
                else { const int k = e - SO_XI; for (int q = 0; q < 8; ++q) s += sacc[SO_XI + k * 8 + q]; }
            }
            out[WPRIV + e] = s;
        }
    }
    if (tid < 2) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) {
            const double* slot = reinterpret_cast<const double*>(sm + SM_SHARED + w * SM_WARP);
            for (int q = 0; q < 32; ++q) s += slot[tid * 32 + q];
        }
        P.cta_stats[(size_t)cta * 4 + tid] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// final reduction over CTAs: gradients in reference parameter order, xi gradient, loss
// ---------------------------------------------------------------------------------------------
constexpr int REDUCE_ITEMS = CTA_OUT;

// The sum over CTAs is split into RED_Y interleaved partial sums (one per warp of a reduce block on the device) that are
// combined in a fixed order: deterministic, and 8x shorter dependent chains than one serial loop over 148 records.
constexpr int RED_Y = 8;
PDL_DEV double reduce_partial(const ReduceArgs& R, int e, int y) {
    double s = 0.0;
    for (int c = y; c < R.n_ctas; c += RED_Y) s += (double)R.cta_out[(size_t)c * CTA_OUT + e];
    return s;
}
PDL_DEV void reduce_store(const ReduceArgs& R, int e, double s);
PDL_DEV void reduce_item(const ReduceArgs& R, int e) {
    double s = 0.0;
    for (int y = 0; y < RED_Y; ++y) s += reduce_partial(R, e, y);
    reduce_store(R, e, s);
}
PDL_DEV void reduce_store(const ReduceArgs& R, int e, double s) {
    const float v = (float)s;
    if (e < WPRIV) {
        const int l = e / WPRIV_LAYER + 1;                   // hidden->hidden layer index 1..H-1
        const int q = e % WPRIV_LAYER;
        const int a = (q / 128) * 4 + (q & 3), lane = (q >> 2) & 31;
        const int g8 = lane >> 2, g4 = lane & 3;
        const int win = pdl_width(l), wout = pdl_width(l + 1);
        if (WGRAD_MMA) {                                     // C-fragment order: [m16 tile][n8 tile][lane][4], then the bias slot
            const int frag = q / 128, i = q & 3;
            if (frag < NFRAG) {
                const int mt = frag / NT8, nt = frag % NT8;
                const int n = 16 * mt + g8 + ((i & 2) ? 8 : 0), k = 8 * nt + 2 * g4 + (i & 1);
                if (n < wout && k < win) R.grad_params[prm_offset(2 * l) + n * win + k] = v;
            } else if (i < 2) {
                const int n = 16 * g4 + g8 + 8 * i;
                if (g4 < NMT && n < wout) R.grad_params[prm_offset(2 * l + 1) + n] = v;
            }
            return;
        }
        if (a < NPN * NPL) {
            const int n = g8 * NPN + a / NPL, k = g4 * NPL + a % NPL;
            if (n < wout && k < win) R.grad_params[prm_offset(2 * l) + n * win + k] = v;
        } else if (a < NACC && g4 == 0) {
            const int n = g8 * NPN + (a - NPN * NPL);
            if (n < wout) R.grad_params[prm_offset(2 * l + 1) + n] = v;
        }
        return;
    }
    const int t = e - WPRIV;
    if (t < SO_B0) { const int n = t / D, dd = t % D; if (n < pdl_width(1)) R.grad_params[prm_offset(0) + n * D + dd] = v; }
    else if (t < SO_WH) { const int n = t - SO_B0; if (n < pdl_width(1)) R.grad_params[prm_offset(1) + n] = v; }
    else if (t < SO_BH) { const int n = t - SO_WH; if (n < pdl_width(H)) R.grad_params[prm_offset(2 * H) + n] = v; }
    else if (t < SO_RAT) { R.grad_params[prm_offset(2 * H + 1)] = v; }
    else if (t < SO_XI) {
#if PDL_ACT == 1
        const int l = (t - SO_RAT) / 7, q = (t - SO_RAT) % 7;
        if (q < 4) R.grad_params[prm_offset(2 * (H + 1) + 2 * l) + q] = v;
        else R.grad_params[prm_offset(2 * (H + 1) + 2 * l + 1) + (q - 4)] = v;
#endif
    } else { const int k = t - SO_XI; if (k < K) R.grad_xi[k] = v; }
}

}  // namespace pdl

// =============================================================================================
// entry points
// =============================================================================================
#ifndef PDL_EMULATE

template <bool BWD, bool DEVN>
__global__ void __launch_bounds__(pdl::NT, 1) pdl_main_kernel(const pdl::Params P) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    pdl::stage_thread(P, sm, tid);
    __syncthreads();
    pdl::stage_thread2(P, sm, tid);
    __syncthreads();
    const int warp = tid >> 5;
    pdl::warp_body<BWD, DEVN>(P, sm, warp, blockIdx.x * pdl::NW + warp, gridDim.x * pdl::NW, tid & 31);
    __syncthreads();
    pdl::cta_tail_thread(P, sm, blockIdx.x, tid, BWD);
}

// block = 32 items x RED_Y partial sums: warp y adds the records c = y, y + RED_Y, ... (coalesced 128-byte rows)
__global__ void __launch_bounds__(32 * pdl::RED_Y) pdl_reduce_kernel(const pdl::ReduceArgs R, int with_grads) {
    __shared__ double part[pdl::RED_Y][32];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + x;
    if (with_grads) {
        part[y][x] = (e < pdl::REDUCE_ITEMS) ? pdl::reduce_partial(R, e, y) : 0.0;
        __syncthreads();
        if (y == 0 && e < pdl::REDUCE_ITEMS) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < pdl::RED_Y; ++q) s += part[q][x];
            pdl::reduce_store(R, e, s);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) pdl::reduce_stats(R);
}

extern "C" {
// static description of the compiled plan (read by the runtime library)
void pdl_k_info(int* out) {
    out[0] = pdl::SM_TOTAL * 4; out[1] = pdl::NT; out[2] = pdl::ZSCR; out[3] = pdl::WPRIV; out[4] = pdl::CTA_OUT;
    out[5] = pdl::NPRM; out[6] = pdl::J; out[7] = pdl::K; out[8] = pdl::D; out[9] = pdl::NW;
    int np = 0; for (int t = 0; t < pdl::NPRM; ++t) np = pdl::prm_offset(t + 1); out[10] = np;
    out[11] = PDL_MODE; out[12] = 8;
}
// returns a cudaError_t
int pdl_k_launch(const pdl::Params* P, int n_ctas, int bwd, float* grad_params, float* grad_xi, double* stats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_done[64] = {false};          // the opt-in shared-memory size is a per-device function attribute
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return (int)e0;
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pdl_main_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pdl::SM_TOTAL * 4);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(pdl_main_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pdl::SM_TOTAL * 4);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(pdl_main_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pdl::SM_TOTAL * 4);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(pdl_main_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pdl::SM_TOTAL * 4);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const bool devn = P->n_points_dev != nullptr;
    // the collocation adjoint's workspace in a persisting L2 window for the duration of this launch, see kernel_params.h
    const PdlL2Window win = (bwd && PDL_MODE == 0) ? pdl_l2_window(P, n_ctas, dev) : PdlL2Window{nullptr, 0, 0.f};
    pdl_l2_window_apply(st, win, true);
    if (bwd && !devn) pdl_main_kernel<true, false><<<n_ctas, pdl::NT, pdl::SM_TOTAL * 4, st>>>(*P);
    else if (bwd) pdl_main_kernel<true, true><<<n_ctas, pdl::NT, pdl::SM_TOTAL * 4, st>>>(*P);
    else if (!devn) pdl_main_kernel<false, false><<<n_ctas, pdl::NT, pdl::SM_TOTAL * 4, st>>>(*P);
    else pdl_main_kernel<false, true><<<n_ctas, pdl::NT, pdl::SM_TOTAL * 4, st>>>(*P);
    cudaError_t e = cudaGetLastError();
    pdl_l2_window_apply(st, win, false);
    if (e != cudaSuccess) return (int)e;
    pdl::ReduceArgs R;
    R.cta_out = P->cta_out; R.cta_stats = P->cta_stats; R.n_ctas = n_ctas;
    R.grad_params = grad_params; R.grad_xi = grad_xi; R.stats = stats; R.inv_n = P->inv_n;
    R.n_points_dev = P->n_points_dev; R.capacity = P->n_points;
    const int items = bwd ? pdl::REDUCE_ITEMS : 1;
    pdl_reduce_kernel<<<(items + 31) / 32, 32 * pdl::RED_Y, 0, st>>>(R, bwd);
    return (int)cudaGetLastError();
}
}

#else  // ---------------------------- CPU emulation (tests only) ----------------------------

extern "C" {
void pdl_k_info(int* out) {
    out[0] = pdl::SM_TOTAL * 4; out[1] = pdl::NT; out[2] = pdl::ZSCR; out[3] = pdl::WPRIV; out[4] = pdl::CTA_OUT;
    out[5] = pdl::NPRM; out[6] = pdl::J; out[7] = pdl::K; out[8] = pdl::D; out[9] = pdl::NW;
    int np = 0; for (int t = 0; t < pdl::NPRM; ++t) np = pdl::prm_offset(t + 1); out[10] = np;
    out[11] = PDL_MODE; out[12] = 8;
}
int pdl_k_launch(const pdl::Params* P, int n_ctas, int bwd, float* grad_params, float* grad_xi, double* stats, void* stream) {
    (void)stream;
    std::vector<float> smv(pdl::SM_TOTAL + 4);
    float* sm = smv.data();
    while (((uintptr_t)sm) & 15) ++sm;
    for (int cta = 0; cta < n_ctas; ++cta) {
        for (int tid = 0; tid < pdl::NT; ++tid) pdl::stage_thread(*P, sm, tid);
        for (int tid = 0; tid < pdl::NT; ++tid) pdl::stage_thread2(*P, sm, tid);
        for (int w = 0; w < pdl::NW; ++w) {
            const bool devn = P->n_points_dev != nullptr;
            if (bwd && devn) pdl::warp_body<true, true>(*P, sm, w, cta * pdl::NW + w, n_ctas * pdl::NW);
            else if (bwd) pdl::warp_body<true, false>(*P, sm, w, cta * pdl::NW + w, n_ctas * pdl::NW);
            else if (devn) pdl::warp_body<false, true>(*P, sm, w, cta * pdl::NW + w, n_ctas * pdl::NW);
            else pdl::warp_body<false, false>(*P, sm, w, cta * pdl::NW + w, n_ctas * pdl::NW);
        }
        for (int tid = 0; tid < pdl::NT; ++tid) pdl::cta_tail_thread(*P, sm, cta, tid, bwd != 0);
    }
    pdl::ReduceArgs R;
    R.cta_out = P->cta_out; R.cta_stats = P->cta_stats; R.n_ctas = n_ctas;
    R.grad_params = grad_params; R.grad_xi = grad_xi; R.stats = stats; R.inv_n = P->inv_n;
    R.n_points_dev = P->n_points_dev; R.capacity = P->n_points;
    if (bwd) for (int e = 0; e < pdl::REDUCE_ITEMS; ++e) pdl::reduce_item(R, e);
    pdl::reduce_stats(R);
    return 0;
}
}
#endif
