// jet_mlp_kernel.cuh -- fused jet-MLP + library-residual forward/adjoint kernel for sm_100a.
//
// Included at the END of a generated plan translation unit (see plan_codegen.cpp), which defines
//   PDL_D, PDL_J, PDL_WP, PDL_H, PDL_K, PDL_ACT, PDL_NS, PDL_NW, PDL_MODE, PDL_SF, PDL_GS
//   pdl_width(l), pdl_comp_axis(c), pdl_tanh_sigmas(), pdl_act_fwd(), pdl_act_adj(), pdl_epilogue()
//
// What it replaces in the reference (file:line under /root/reference):
//   Network.forward                Code/Classes/Network.py:287-312     -> layer loop below (jets, not values)
//   Derivative_From_Derivative     Code/Evaluate_Derivatives.py:20-207 -> forward jet propagation (no autograd)
//   Coll_Loss                      Code/Loss.py:66-248                 -> pdl_epilogue + reductions
//   Total.backward()               Code/Test_Train.py:187-188          -> hand-written adjoint (act_adj/dgrad/wgrad)
//
// Execution model: persistent grid, PDL_NW warps per CTA, every warp owns tiles of 8 collocation points.
// lane = pg*4 + g4:  pg = point within the tile, g4 = group of NPL = WP/4 output neurons.
// Weights live in shared memory (two layouts, one per GEMM direction); jets are exchanged between lanes
// through per-warp shared buffers; pre-activation jets needed by the adjoint are parked in an L2-resident
// per-warp scratch; weight-gradient partial sums live in a per-warp private buffer (no atomics).
//
// The same source compiles as plain C++ with -DPDL_EMULATE: every "per-lane" block becomes a loop over
// 32 lanes and shuffles become array permutations.  That build is TEST INFRASTRUCTURE (tests/ only).
#pragma once
#include "jet_common.cuh"

namespace pdl {

constexpr int SF = PDL_SF;                  // row stride of the forward weight copy  Wf[n][k]
constexpr int GS = PDL_GS;                  // group stride of the dgrad weight copy  Wd[n][g4][i]
constexpr int SD = 4 * GS;
constexpr int NACC = NPN * NPL + NPN;       // wgrad accumulators per lane (weights + bias)
constexpr int NACC4 = (NACC + 3) / 4 * 4;
constexpr int WPRIV_LAYER = NACC4 * 32;     // floats per warp per hidden->hidden layer
constexpr int WPRIV = NHH * WPRIV_LAYER;

// small per-warp accumulators (shared memory), in floats

// shared-memory map (floats).  Shared part first, then NW private parts.
constexpr int SM_HH = (SM_B0 + WP + 3) / 4 * 4;              // per hidden->hidden layer: Wf, Wd, bias
constexpr int HH_WF = 0, HH_WD = WP * SF, HH_B = HH_WD + WP * SD, HH_SIZE = (HH_B + WP + 3) / 4 * 4;
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

// forward activation of one neuron: jets z -> jets h
PDL_DEV void act_forward(const float (&z)[J], const float* rp, float (&h)[J]) {
    float s[NS], r[NS], inv;
    sigmas(z[0], rp, s, r, inv);
    pdl_act_fwd(z, s, h);
}
// adjoint of one neuron's activation: (z, hb) -> zb ; accumulates rational coefficient gradients
PDL_DEV void act_adjoint(const float (&z)[J], const float (&hb)[J], const float* rp, float (&zb)[J], float (&ga)[7]) {
    float s[NS], r[NS], inv, sb[NS - 1];
    sigmas(z[0], rp, s, r, inv);
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
            base[HH_WF + n * SF + k] = w;
            base[HH_WD + n * SD + (k / NPL) * GS + (k % NPL)] = w;
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
                const int n = g4 * NPL + i;
                float z[J];
                layer0_preact(sm, n, x[PDL_L], z);
                act_forward(z, sm + SM_RAT, h[PDL_L][i]);
                store_entry(bufH + pg * SA + n * JA, bufH + 8 * SA + pg * SB + n * JB, h[PDL_L][i]);
            }
        }
        PDL_SYNCWARP();
        // ---------------- hidden -> hidden layers, forward ----------------
        for (int l = 1; l < H; ++l) {
            const float* Wf = sm + SM_HH + (l - 1) * HH_SIZE + HH_WF;
            const float* bl = sm + SM_HH + (l - 1) * HH_SIZE + HH_B;
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
                float acc[NPL][J];
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    acc[i][0] = bl[g4 * NPL + i];
#pragma unroll
                    for (int c = 1; c < J; ++c) acc[i][c] = 0.f;
                }
                const float* hA = bufH + pg * SA;
                const float* hB = bufH + 8 * SA + pg * SB;
                const float* wrow = Wf + (g4 * NPL) * SF;
#pragma unroll 2
                for (int kc = 0; kc < WP; kc += 4) {
                    float a[4][J];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) load_entry(hA + (kc + kk) * JA, hB + (kc + kk) * JB, a[kk]);
#pragma unroll
                    for (int i = 0; i < NPL; ++i) {
                        const pdl_f4 w = ld4(wrow + i * SF + kc);
#pragma unroll
                        for (int c = 0; c < J; ++c) {
                            acc[i][c] += a[0][c] * w.x;
                            acc[i][c] += a[1][c] * w.y;
                            acc[i][c] += a[2][c] * w.z;
                            acc[i][c] += a[3][c] * w.w;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    if (BWD) scratch_store(zs, l - 1, i, lane, acc[i]);
                    act_forward(acc[i], sm + SM_RAT + l * 8, h[PDL_L][i]);
                }
            }
            PDL_SYNCWARP();
            if (l < H - 1) {                 // the last hidden layer feeds the output layer from registers
                PDL_FOR_LANES {
                    const int pg = lane >> 2, g4 = lane & 3;
#pragma unroll
                    for (int i = 0; i < NPL; ++i) {
                        const int n = g4 * NPL + i;
                        store_entry(bufH + pg * SA + n * JA, bufH + 8 * SA + pg * SB + n * JB, h[PDL_L][i]);
                    }
                }
                PDL_SYNCWARP();
            }
        }
        // ---------------- output layer (identity activation) ----------------
        PDL_FOR_LANES {
            const int g4 = lane & 3;
#pragma unroll
            for (int c = 0; c < J; ++c) uu[PDL_L][c] = 0.f;
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                const float w = sm[SM_WH + g4 * NPL + i];
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
                    const float w = sm[SM_WH + g4 * NPL + i];
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
                    for (int i = 0; i < NPL; ++i) sacc[SO_WH + g4 * NPL + i] += g[PDL_L][i];
                    if (g4 == 0) sacc[SO_BH] += g[PDL_L][NPL];
                }
            }
        }
        // ---------------- hidden -> hidden layers, adjoint ----------------
        for (int l = H - 1; l >= 1; --l) {
            float ga[PDL_NL][7];
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
#pragma unroll
                for (int q = 0; q < 7; ++q) ga[PDL_L][q] = 0.f;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int n = g4 * NPL + i;
                    float z[J], zb[J];
                    scratch_load(zs, l - 1, i, lane, z);
                    act_adjoint(z, hb[PDL_L][i], sm + SM_RAT + l * 8, zb, ga[PDL_L]);
                    store_entry(bufZ + pg * SA + n * JA, bufZ + 8 * SA + pg * SB + n * JB, zb);
                    // activations of the layer below (input of this linear layer), recomputed
                    float zi[J], hi[J];
                    if (l >= 2) scratch_load(zs, l - 2, i, lane, zi);
                    else layer0_preact(sm, n, x[PDL_L], zi);
                    act_forward(zi, sm + SM_RAT + (l - 1) * 8, hi);
                    store_entry(bufH + pg * SA + n * JA, bufH + 8 * SA + pg * SB + n * JB, hi);
                }
            }
            PDL_SYNCWARP();
#if PDL_ACT == 1
            xor_add<7>(ga, 1); xor_add<7>(ga, 2); xor_add<7>(ga, 4); xor_add<7>(ga, 8); xor_add<7>(ga, 16);
            PDL_FOR_LANES { if (lane == 0) { for (int q = 0; q < 7; ++q) sacc[SO_RAT + l * 7 + q] += ga[PDL_L][q]; } }
#endif
            const float* Wd = sm + SM_HH + (l - 1) * HH_SIZE + HH_WD;
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3, g8 = lane >> 2;
                // ---- weight gradient: Wbar[n][k] += sum_p sum_c zb[p][n][c] * h[p][k][c]   (n in g8 group, k in g4 group)
                {
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
                // ---- data gradient: hb[p][k] = sum_n zb[p][n] * W[n][k]   (k in own group)
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
            PDL_SYNCWARP();
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
                    const int n = g4 * NPL + i;
                    float z[J], zb[J];
                    layer0_preact(sm, n, x[PDL_L], z);
                    act_adjoint(z, hb[PDL_L][i], sm + SM_RAT, zb, ga[PDL_L]);
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
            xor_add<7>(ga, 1); xor_add<7>(ga, 2); xor_add<7>(ga, 4); xor_add<7>(ga, 8); xor_add<7>(ga, 16);
#endif
            PDL_FOR_LANES {
                const int pg = lane >> 2, g4 = lane & 3;
                if (pg == 0) {
#pragma unroll
                    for (int i = 0; i < NPL; ++i) {
                        const int n = g4 * NPL + i;
#pragma unroll
                        for (int dd = 0; dd < D; ++dd) sacc[SO_W0 + n * D + dd] += g0[PDL_L][i * (D + 1) + dd];
                        sacc[SO_B0 + n] += g0[PDL_L][i * (D + 1) + D];
                    }
                }
#if PDL_ACT == 1
                if (lane == 0) { for (int q = 0; q < 7; ++q) sacc[SO_RAT + q] += ga[PDL_L][q]; }
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
