// jet_mlp_fwd_tc.cuh -- forward-only jet-MLP + library residual on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Serves every launch without gradients of a collocation plan (Coll_Loss under no_grad, Testing, Evaluate_Residual,
// Evaluate_Derivatives -- Code/Loss.py:66-248, Code/Test_Train.py:205-330, Plot/Plot_One_Spatial_Dimension.py:149-178) whose jet set
// fits the tensor memory (see the budget below); everything else, and every launch with gradients, stays on jet_mlp_kernel_mma.cuh.
// Included at the end of a generated plan translation unit when plan_codegen.cpp defines PDL_FWD_TC; replaces Network.forward
// (Code/Classes/Network.py:287-312) + Derivative_From_Derivative (Code/Evaluate_Derivatives.py:20-207) for that case.
//
// Shape of the computation (measured building block: scripts/ubench/tc_jet.cu, profiles/r02_tc_jet*.txt):
//   * one persistent CTA per SM; a tile = 128 collocation points; THREAD = POINT: TMEM lane p of an accumulator is point p, so after
//     tcgen05.ld (32x32b) a thread holds all jet components of its point's neurons and the Faa-di-Bruno activation is thread-local;
//   * hidden->hidden layer  z_a = W h_a  for the J jet components a:  D_a[128 x NP] (TMEM, FP32) += A_a[128 x 8] (TMEM, TF32) *
//     B[8 x NP] (shared memory, W as hi/lo TF32 planes, K-major 8x16-byte core matrices), one k-step = 8 input neurons;
//     FP32 accuracy from the split  D += hi*Bhi + hi*Blo + lo*Bhi  (3 tcgen05.mma.kind::tf32 of 128 x NP x 8 per component and k-step);
//   * the activation warps stream: tcgen05.ld 8 neurons x J components of z -> jets h = act(z) (generated code) -> split hi/lo ->
//     tcgen05.st into a ring slot of the A operand -> mbarrier arrive; ONE issuer warp (elect.sync, warp-uniform operands so that
//     ptxas keeps every descriptor in uniform registers: 44 instead of 127-149 cycles per MMA, profiles/r02_tc_jet_t5/t6.txt) waits
//     for the slot, issues the 3J MMAs of that k-step and tcgen05.commit's the slot back; the MMAs of k-step g run under the
//     activation work of k-step g+1;
//   * accumulators are double buffered (layer l+1 accumulates while layer l is still being read);
//   * 2 warpgroups of activation warps, each taking 4 of the 8 neurons of every k-step of the same 128 points = 2 warps per scheduler.
// TMEM budget (512 columns): 2 x [J accumulators at a pitch of NP (or WP: the NP - WP padding columns of component a then overlap the
// first columns of component a+1, which only ever adds the zero rows of W)] + NSLOT x 16J ring columns.  heat (J = 4, w = 40):
// 2 x 192 + 2 x 64 = 512; J = 5: 2 x 208 + 80 = 496; J = 6 at w = 40 (KS, rd2d): ONE accumulator set, 248 + 2 x 96 = 440 (the
// activation warps drain a whole layer into registers at the layer boundary, see DOUBLE_D below).
#pragma once
#include <cstddef>
#if !defined(PDL_EMULATE) && defined(PDL_FWD_TC) && PDL_FWD_TC

namespace pdl_tc {
using namespace pdl;

#ifndef PDL_TC_NWG
#define PDL_TC_NWG 4      // measured (profiles/r02_tc_fwd_j6.txt, r02_tc_fwd_ring.txt): 4 warpgroups gain 2-6 % over 2 with the deep ring
#endif
#ifndef PDL_TC_SPLIT
#define PDL_TC_SPLIT 2
#endif
// how an activation x becomes the two TF32 operands (the tensor core reads the upper 19 bits of a 32-bit operand, i.e. it truncates):
//   0: hi = rna(x), lo = rna(x - hi)      5 instructions per value (what the mma.sync forward does)
//   1: hi = rna(x), lo = x - hi (raw)     3
//   2: hi = x (raw), lo = x - trunc(x)    2   -- all three are at FP32-matmul level in a 4-layer tanh chain (8e-7 .. 9.7e-7 relative,
//                                               plain FP32 matmul 7.4e-7 .. 8.1e-7: numpy emulation) and in the parity tests
constexpr int NWG = PDL_TC_NWG;                    // activation warpgroups
constexpr int NACT = 4 * NWG;                      // activation warps; warp NACT is the issuer
constexpr int THREADS = 32 * (NACT + 1);
constexpr int NP = (WP + 15) / 16 * 16;            // MMA N (M = 128 needs a multiple of 16)
constexpr int NKS = WP / 8;                        // k-steps per layer
constexpr int SLOT = 16 * J;                       // ring slot: per component 8 hi + 8 lo columns
constexpr int NMM = H - 1;                         // tensor-core layers per tile
constexpr int NPW = 8 / NWG;                       // neurons of a k-step per warpgroup
// Two ways to spend the 512 columns.  DOUBLE: two accumulator sets (layer l+1 accumulates while layer l is still being read) + a ring
// of 1-2 slots.  SINGLE: one accumulator set + a ring of up to 6 slots; the activation warps then read their whole share of a layer
// (NKS k-steps x J components x NPW neurons) into registers before they produce the first k-step of the next layer.  No extra barrier
// is needed for that: every activation warp takes part in every k-step, so the a_full completion of the next layer's first k-step
// already implies that all of them have drained the accumulators the first MMA overwrites.  A layer cannot start before the previous
// one has completed anyway (its first k-step needs the finished pre-activations), so the second accumulator set only overlaps the
// remaining tcgen05.ld's with MMAs.  Measured (profiles/r02_tc_fwd_ring.txt): J = 5, where DOUBLE leaves ONE ring slot, gains 13 % from
// SINGLE with three; J = 4 is neutral (5 slots vs 2: the chain activation -> MMA tail -> next layer is serial either way) and gains 2-4 %
// together with 4 warpgroups.  PDL_TC_SINGLE_D = 0 / 1 forces one of them (A/B runs), default: SINGLE when it gives >= 3 slots.
#ifndef PDL_TC_SINGLE_D
#define PDL_TC_SINGLE_D -1
#endif
constexpr int imin(int a, int b) { return a < b ? a : b; }
constexpr int PITCH2 = (2 * NP * J + 2 * SLOT <= 512) ? NP : WP;
constexpr int DC2 = PITCH2 * (J - 1) + NP;
constexpr int NS2 = (2 * DC2 + SLOT <= 512) ? imin(2, (512 - 2 * DC2) / SLOT) : 0;
constexpr int PITCH1 = (NP * J + 2 * SLOT <= 512) ? NP : WP;
constexpr int DC1 = PITCH1 * (J - 1) + NP;
constexpr int NS1 = (NKS * J * NPW <= 128 && DC1 + SLOT <= 512) ? imin(6, (512 - DC1) / SLOT) : 0;
constexpr bool DOUBLE_D = (PDL_TC_SINGLE_D == 0 && NS2 > 0) || (PDL_TC_SINGLE_D != 1 && !(NS1 >= 3 || NS2 == 0)) || NS1 == 0;
constexpr int NDSET = DOUBLE_D ? 2 : 1;
constexpr int PITCH = DOUBLE_D ? PITCH2 : PITCH1;  // column pitch of the per-component accumulators (WP: the NP - WP padding columns of
constexpr int DCOLS = DOUBLE_D ? DC2 : DC1;        // component a overlap the first columns of component a+1: only zero rows of W land there)
constexpr int NSLOT = DOUBLE_D ? NS2 : NS1;
static_assert(NSLOT >= 1 && NDSET * DCOLS + NSLOT * SLOT <= 512, "jet set does not fit the tensor memory (plan_codegen.cpp must not select this engine)");
static_assert(H >= 2, "needs a hidden->hidden layer");
constexpr uint32_t TM_D0 = 0, TM_D1 = DOUBLE_D ? DCOLS : 0, TM_A = NDSET * DCOLS;
static_assert(NWG == 1 || NWG == 2 || NWG == 4, "1, 2 or 4 activation warpgroups");

struct Smem {
    float B[NMM][2][WP / 4][NP][4];                // [layer][hi/lo][k chunk][n][4]: K-major, 8 x 16-byte core matrices
    float W0[WP][D];
    float b0[WP];
    float bias[NMM][WP];
    float wH[WP];
    float rat[H][8];
    float xi[KX];
    float bH;
    float upart[2][NWG > 1 ? NWG - 1 : 1][128][J]; // partial output sums of warpgroups 1.., double buffered by tile parity
    double red[2][4];
    unsigned long long a_full[8], a_empty[8], d_full[2];
    uint32_t tmem_base;
    int failed;
};
constexpr int SMEM_BYTES = (int)sizeof(Smem) + 128;

// ---- PTX wrappers ----
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute/arch/mma_sm100_desc.hpp: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout (0 = no swizzle) << 61
    return (unsigned long long)((saddr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor, kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// D[tmem] (+)= A[tmem] * B[smem]; every lane executes the statement, elect.sync picks the issuing lane (warp-uniform code)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, unsigned long long bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(IDESC),
                 "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(unsigned long long* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(s32(bar)) : "memory");
}
#ifdef PDL_TC_PROFILE
#define TC_T0() const long long t0_ = clock64()
#define TC_ACC(x) (x) += clock64() - t0_
#else
#define TC_T0() do {} while (0)
#define TC_ACC(x) do {} while (0)
#endif
// bounded wait: a wrong barrier protocol must end in an error, not in a hung GPU
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, uint32_t parity, volatile int* failed) {
    const uint32_t a = s32(bar);
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
        if ((it & 255) == 255 && *failed) return false;
    }
    *failed = 1;
    __trap();              // loud: the launch ends with an error (a thread parked in bar.sync could not see the flag)
    return false;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tm_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t taddr, float (&v)[2]) {
    uint32_t r[2];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
    v[0] = __uint_as_float(r[0]); v[1] = __uint_as_float(r[1]);
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, const uint32_t (&r)[2]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
template <int N> __device__ __forceinline__ void tm_ldn(uint32_t taddr, float (&v)[N]);
template <> __device__ __forceinline__ void tm_ldn<2>(uint32_t taddr, float (&v)[2]) { tm_ld2(taddr, v); }

template <> __device__ __forceinline__ void tm_ldn<8>(uint32_t taddr, float (&v)[8]) { tm_ld8(taddr, v); }
template <> __device__ __forceinline__ void tm_ldn<4>(uint32_t taddr, float (&v)[4]) { tm_ld4(taddr, v); }
template <int N> __device__ __forceinline__ void tm_stn(uint32_t taddr, const uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tm_stn<8>(uint32_t taddr, const uint32_t (&r)[8]) { tm_st8(taddr, r); }
template <> __device__ __forceinline__ void tm_stn<4>(uint32_t taddr, const uint32_t (&r)[4]) { tm_st4(taddr, r); }
template <> __device__ __forceinline__ void tm_stn<2>(uint32_t taddr, const uint32_t (&r)[2]) { tm_st2(taddr, r); }
__device__ __forceinline__ void tm_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// ---- CTA prologue: parameters -> shared memory (W as hi/lo TF32 planes in the tensor core's K-major layout) ----
__device__ void stage(const Params& P, Smem& S, int tid) {
    float* raw = reinterpret_cast<float*>(&S);
    for (int e = tid; e < (int)(offsetof(Smem, a_full) / 4); e += THREADS) raw[e] = 0.f;
    __syncthreads();
    const int w1 = pdl_width(1);
    for (int e = tid; e < w1 * D; e += THREADS) S.W0[e / D][e % D] = P.prm[0][e];
    for (int e = tid; e < w1; e += THREADS) S.b0[e] = P.prm[1][e];
    for (int l = 1; l < H; ++l) {
        const int win = pdl_width(l), wout = pdl_width(l + 1);
        const float* W = P.prm[2 * l];
        for (int e = tid; e < wout * win; e += THREADS) {
            const int n = e / win, k = e - n * win;
            const float w = W[e];
            const uint32_t uh = tf32_rna(w), ul = tf32_rna(w - u2f(uh));
            S.B[l - 1][0][k >> 2][n][k & 3] = u2f(uh);
            S.B[l - 1][1][k >> 2][n][k & 3] = u2f(ul);
        }
        for (int e = tid; e < wout; e += THREADS) S.bias[l - 1][e] = P.prm[2 * l + 1][e];
    }
    const int wH = pdl_width(H);
    for (int e = tid; e < wH; e += THREADS) S.wH[e] = P.prm[2 * H][e];
    if (tid == 0) S.bH = P.prm[2 * H + 1][0];
#if PDL_ACT == 1
    for (int e = tid; e < H * 7; e += THREADS) {
        const int l = e / 7, q = e - l * 7;
        S.rat[l][q] = (q < 4) ? P.prm[2 * (H + 1) + 2 * l][q] : P.prm[2 * (H + 1) + 2 * l + 1][q - 4];
    }
#endif
    for (int e = tid; e < K; e += THREADS) S.xi[e] = (P.mask[e] != 0) ? 0.f : P.xi[e];
}

// ring bookkeeping: the k-steps of a CTA use the slots round-robin; every role keeps its own (slot, wraps) pair and advances it once
// per k-step, `wraps` = how often the ring has been filled completely = the use count of the current slot.  Every activation warp
// takes part in every k-step (its warpgroup's NPW neurons), so a waiter is never more than one phase behind the barrier it waits on.
struct Ring {
    int slot; uint32_t wraps;
    int pending;                                   // slots written but not yet handed to the issuer (activation warps only)
    __device__ __forceinline__ void next() { if (++slot == NSLOT) { slot = 0; ++wraps; } }
};
// k-steps handed to the issuer per tcgen05.wait::st + fence + arrive sequence.  Sharing one hand-off between two k-steps (needs a ring
// of at least twice the batch) was measured and is OFF: heat 0.553 -> 0.605 ms (profiles/r02_tc_fwd_batch.txt) -- the MMAs of the
// first k-step of a pair start later and the layer's MMA tail, which the next layer waits for, moves out by the same amount.
#ifndef PDL_TC_BATCH
#define PDL_TC_BATCH 1
#endif
constexpr int BATCH = (PDL_TC_BATCH >= 2 && NSLOT >= 2 * PDL_TC_BATCH) ? PDL_TC_BATCH : 1;

// activation warps: jets h of this warpgroup's NPW neurons -> hi/lo TF32 -> the next ring slot
__device__ __forceinline__ void produce(Smem& S, uint32_t lane_addr, Ring& ring, int wg, const float (&h)[NPW][J], int lane,
                                        long long& t_wait, bool hand_over) {
    const int slot = ring.slot;
    { TC_T0(); mbar_wait(&S.a_empty[slot], (ring.wraps & 1u) ^ 1u, &S.failed); TC_ACC(t_wait); }   // consumed by the MMAs of its previous use
    fence_after();
    const uint32_t a_out = lane_addr + TM_A + slot * SLOT + wg * NPW;
#pragma unroll
    for (int c = 0; c < J; ++c) {
        uint32_t whi[NPW], wlo[NPW];
#pragma unroll
        for (int i = 0; i < NPW; ++i) {
            const float x = h[i][c];
            if (PDL_TC_SPLIT == 2) {
                whi[i] = f2u(x);
                wlo[i] = f2u(x - u2f(f2u(x) & 0xFFFFE000u));
            } else {
                whi[i] = tf32_rna(x);
                const float lo = x - u2f(whi[i]);
                wlo[i] = (PDL_TC_SPLIT == 0) ? tf32_rna(lo) : f2u(lo);
            }
        }
        tm_stn<NPW>(a_out + 16 * c, whi);
        tm_stn<NPW>(a_out + 16 * c + 8, wlo);
    }
    ring.next();
    ++ring.pending;
    if (!hand_over) return;
    tm_st_wait();
    fence_before();
    __syncwarp();
    if (lane == 0) {
        int sl = ring.slot;
        for (int q = 0; q < ring.pending; ++q) { sl = (sl == 0) ? NSLOT - 1 : sl - 1; mbar_arrive(&S.a_full[sl]); }
    }
    ring.pending = 0;
}

// issuer warp: the NMM tensor-core layers of one tile.  MPAR = parity of the tile's first layer index, so that every TMEM address
// and shared-memory descriptor offset is a compile-time constant (uniform registers, no per-MMA register shuffling).
template <int MPAR>
__device__ __forceinline__ bool issue_tile(Smem& S, long long m0, uint32_t b_base, Ring& ring, long long& t_wait) {
#pragma unroll
    for (int l = 0; l < NMM; ++l) {
        const uint32_t d_out = (((MPAR + l) & 1) ? TM_D1 : TM_D0);
#pragma unroll
        for (int g = 0; g < NKS; ++g) {
            const int slot = ring.slot;            // warp-uniform run-time value: the addresses below stay in uniform registers
            { TC_T0(); if (!mbar_wait(&S.a_full[slot], ring.wraps & 1u, &S.failed)) return false; TC_ACC(t_wait); }
            fence_after();
            const uint32_t a_in = TM_A + slot * SLOT;
            const uint32_t off = (uint32_t)((((l * 2 + 0) * (WP / 4) + 2 * g) * NP) * 16);
            const uint32_t off_lo = (uint32_t)((((l * 2 + 1) * (WP / 4) + 2 * g) * NP) * 16);
            const unsigned long long bhi = smem_desc(b_base + off, NP * 16, 128), blo = smem_desc(b_base + off_lo, NP * 16, 128);
#pragma unroll
            for (int c = 0; c < J; ++c) {
                mma_ts(d_out + PITCH * c, a_in + 16 * c, bhi, g > 0 ? 1u : 0u);
                mma_ts(d_out + PITCH * c, a_in + 16 * c, blo, 1u);
                mma_ts(d_out + PITCH * c, a_in + 16 * c + 8, bhi, 1u);
            }
            commit(&S.a_empty[slot]);
            ring.next();
        }
        commit(&S.d_full[(MPAR + l) & 1]);
    }
    return true;
}

// activation warps, hidden layers 2..H: one k-step's share of pre-activations (this warpgroup's NPW neurons, all components) ->
// jets h -> ring slot of the next tensor-core layer, or -- after the last hidden layer -- the output layer's partial sums
__device__ __forceinline__ void layer_slice(Smem& S, const float (&zc)[J][NPW], int l, int g, int wg, Ring& ring, uint32_t lane_addr, int lane,
                                            float (&u)[J], long long& t_empty) {
    float h[NPW][J];
#pragma unroll
    for (int i = 0; i < NPW; ++i) {
        float z[J], tc;
#pragma unroll
        for (int c = 0; c < J; ++c) z[c] = zc[c][i];
        z[0] += S.bias[l - 1][8 * g + wg * NPW + i];
        act_forward(z, S.rat[l], h[i], tc);
    }
    if (l < H - 1) {
        produce(S, lane_addr, ring, wg, h, lane, t_empty, (g % BATCH) == BATCH - 1 || g == NKS - 1);
    } else {
#pragma unroll
        for (int i = 0; i < NPW; ++i) {
            const float w = S.wH[8 * g + wg * NPW + i];
#pragma unroll
            for (int c = 0; c < J; ++c) u[c] += w * h[i][c];
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) pdl_fwd_tc_kernel(const Params P) {
    extern __shared__ __align__(128) unsigned char tc_raw[];
    Smem& S = *reinterpret_cast<Smem*>(tc_raw + ((128u - (s32(tc_raw) & 127u)) & 127u));
    // warp index through a shuffle: provably warp-uniform for ptxas, so the role branches below are uniform control flow and the issuer's
    // addresses / descriptors can live in uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    stage(P, S, tid);
    if (tid == 0) {
        S.failed = 0;
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&S.a_full[s], NACT); mbar_init(&S.a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) mbar_init(&S.d_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == NACT) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s32(&S.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // the staged W planes are read by the tensor core (async proxy)
    fence_before();
    __syncthreads();
    fence_after();
    // one CTA per SM owns the whole tensor memory: the allocation starts at column 0, and the issuer relies on it (constant addresses)
    if (S.tmem_base != 0) __trap();

    long long NPTS = P.n_points;
    if (P.n_points_dev) { NPTS = *P.n_points_dev; if (NPTS < 0) NPTS = 0; if (NPTS > P.n_points) NPTS = P.n_points; }
    const long long n_tiles = (NPTS + 127) / 128;
    double lsum = 0.0, asum = 0.0;

    if (warp < NACT) {
        // ================= activation warps: thread = point =================
        const int wg = warp >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane;                       // TMEM lane = point within the tile
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        long long t_empty = 0, t_dfull = 0, t_ld = 0, t_bar = 0;
        const long long t_begin = clock64();
        long long it = 0;
        Ring ring = {0, 0u, 0};
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long m0 = it * NMM;
            const long long p = tile * 128 + row;
            const bool valid = p < NPTS;
            float x[D];
#pragma unroll
            for (int dd = 0; dd < D; ++dd) x[dd] = valid ? P.points[p * D + dd] : 0.f;
            // ---- hidden layer 1 from the input layer (no tensor-core work: the first-order jets are weight columns) ----
#pragma unroll 1
            for (int g = 0; g < NKS; ++g) {
                float h[NPW][J];
#pragma unroll
                for (int i = 0; i < NPW; ++i) {
                    const int n = 8 * g + wg * NPW + i;
                    float z[J], tc;
                    float v = S.b0[n];
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) v += S.W0[n][dd] * x[dd];
                    z[0] = v;
#pragma unroll
                    for (int c = 1; c < J; ++c) z[c] = (pdl_comp_axis(c) >= 0) ? S.W0[n][(pdl_comp_axis(c) >= 0) ? pdl_comp_axis(c) : 0] : 0.f;
                    act_forward(z, S.rat[0], h[i], tc);
                }
                produce(S, lane_addr, ring, wg, h, lane, t_empty, (g % BATCH) == BATCH - 1 || g == NKS - 1);
            }
            // ---- hidden layers 2..H: pre-activations come out of the tensor core ----
            float u[J];
#pragma unroll
            for (int c = 0; c < J; ++c) u[c] = 0.f;
#pragma unroll 1
            for (int l = 1; l < H; ++l) {
                const long long m = m0 + l - 1;                    // the tensor-core layer that produced z_{l+1}
                { TC_T0(); mbar_wait(&S.d_full[m & 1], (uint32_t)((m >> 1) & 1), &S.failed); TC_ACC(t_dfull); }
                fence_after();
                const uint32_t d_in = lane_addr + ((m & 1) ? TM_D1 : TM_D0) + wg * NPW;
                if (DOUBLE_D) {
#pragma unroll 1
                    for (int g = 0; g < NKS; ++g) {
                        float zc[J][NPW];
#pragma unroll
                        for (int c = 0; c < J; ++c) tm_ldn<NPW>(d_in + PITCH * c + 8 * g, zc[c]);
                        { TC_T0(); tm_ld_wait(); TC_ACC(t_ld); }
                        layer_slice(S, zc, l, g, wg, ring, lane_addr, lane, u, t_empty);
                    }
                } else {
                    // single accumulator set: drain this thread's whole share of the layer first (fully unrolled: register arrays)
                    float zall[NKS][J][NPW];
#pragma unroll
                    for (int g = 0; g < NKS; ++g)
#pragma unroll
                        for (int c = 0; c < J; ++c) tm_ldn<NPW>(d_in + PITCH * c + 8 * g, zall[g][c]);
                    { TC_T0(); tm_ld_wait(); TC_ACC(t_ld); }
#pragma unroll
                    for (int g = 0; g < NKS; ++g) layer_slice(S, zall[g], l, g, wg, ring, lane_addr, lane, u, t_empty);
                }
            }
            // ---- output layer + library terms + residual (thread-local); warpgroup 1 hands its partial sums to warpgroup 0 ----
            if (NWG > 1) {
                if (wg > 0) {
#pragma unroll
                    for (int c = 0; c < J; ++c) S.upart[it & 1][wg - 1][row][c] = u[c];
                }
                { TC_T0(); asm volatile("bar.sync 1, %0;\n" ::"n"(NACT * 32) : "memory"); TC_ACC(t_bar); }
                if (wg == 0) {
#pragma unroll
                    for (int q = 0; q < NWG - 1; ++q)
#pragma unroll
                        for (int c = 0; c < J; ++c) u[c] += S.upart[it & 1][q][row][c];
                }
            }
            if (wg == 0) {
                u[0] += S.bH;
                float r, du[J], T[KX];
                pdl_epilogue(u, S.xi, r, du, T);
                if (valid) {
                    if (P.residual) P.residual[p] = r;
#ifndef PDL_TC_PROFILE
                    if (P.features) {
#pragma unroll
                        for (int c = 0; c < J; ++c) P.features[(size_t)c * P.n_points + p] = u[c];
                    }
#endif
                    lsum += (double)r * (double)r;
                    asum += fabs((double)r);
                }
            }
        }
#ifdef PDL_TC_PROFILE
        if (P.features && lane == 0 && (warp == 0 || warp == 4)) {        // cycles: total, waiting for a free ring slot, for D, for tcgen05.ld, at the tile barrier
            float* o = P.features + blockIdx.x * 16 + (warp ? 8 : 0);
            o[0] = (float)(clock64() - t_begin); o[1] = (float)t_empty; o[2] = (float)t_dfull; o[3] = (float)t_ld; o[4] = (float)t_bar; o[5] = (float)it;
        }
#endif
    } else {
        // ================= issuer warp (all 32 lanes run the loop; elect.sync inside the wrappers) =================
        const uint32_t b_base = s32(&S.B[0][0][0][0][0]);
        long long it = 0, t_full = 0;
        Ring ring = {0, 0u, 0};
        const long long t_begin = clock64();
        bool ok = true;
        for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
            const long long m0 = it * NMM;
            ok = (m0 & 1) ? issue_tile<1>(S, m0, b_base, ring, t_full) : issue_tile<0>(S, m0, b_base, ring, t_full);
        }
#ifdef PDL_TC_PROFILE
        if (P.features && lane == 0) { float* o = P.features + blockIdx.x * 16; o[6] = (float)(clock64() - t_begin); o[7] = (float)t_full; }
#endif
    }
    // ---- CTA tail: loss partials (double), TMEM release ----
    for (int o = 16; o > 0; o >>= 1) { lsum += __shfl_xor_sync(0xffffffffu, lsum, o); asum += __shfl_xor_sync(0xffffffffu, asum, o); }
    if (warp < 4 && lane == 0) { S.red[0][warp] = lsum; S.red[1][warp] = asum; }
    fence_before();
    __syncthreads();
    fence_after();
    if (S.failed) __trap();                                       // loud: the launch ends with an error instead of wrong numbers
    if (tid < 2) P.cta_stats[(size_t)blockIdx.x * 4 + tid] = S.red[tid][0] + S.red[tid][1] + S.red[tid][2] + S.red[tid][3];
    if (warp == NACT) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(0u));
}

}  // namespace pdl_tc

// ---- module entry points: launches without gradients run here, everything else is forwarded to the mma.sync engine ----
extern "C" {
void pdl_k_info(int* out) {
    pdl_k_info_mma(out);
    out[13] = 1;                                  // the tcgen05 forward engine is present
}
int pdl_k_launch(const pdl::Params* P, int n_ctas, int bwd, float* grad_params, float* grad_xi, double* stats, void* stream) {
    if (bwd) return pdl_k_launch_mma(P, n_ctas, bwd, grad_params, grad_xi, stats, stream);
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return (int)e0;
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pdl_tc::pdl_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pdl_tc::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const long long tiles = (P->n_points + 127) / 128;
    int grid = n_ctas;
    if (tiles < grid) grid = (int)(tiles > 0 ? tiles : 1);
    pdl_tc::pdl_fwd_tc_kernel<<<grid, pdl_tc::THREADS, pdl_tc::SMEM_BYTES, st>>>(*P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    pdl::ReduceArgs R;                            // sums the CTAs' loss partials (double) into stats, as for the mma.sync forward
    R.cta_out = P->cta_out; R.cta_stats = P->cta_stats; R.n_ctas = grid;
    R.grad_params = grad_params; R.grad_xi = grad_xi; R.stats = stats; R.inv_n = P->inv_n;
    R.n_points_dev = P->n_points_dev; R.capacity = P->n_points;
    pdl_reduce_kernel<<<1, 32 * pdl::RED_Y, 0, st>>>(R, 0);
    return (int)cudaGetLastError();
}
}

#endif  // PDL_FWD_TC
