// Ceiling of register-tiled FFMA code on B200: an 8x8 outer-product tile per thread, operands reloaded from shared memory
// (LDS.128, broadcast-style addresses) every step, at 8/12/16 warps per SM.  Also a pure-register variant.
#include <cuda_runtime.h>
#include <stdio.h>

template <int WITH_LDS, int TM, int TN>
__global__ void tile_kernel(float* out, int iters) {
    __shared__ float4 sa[512], sb[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { sa[i] = make_float4(i, 1, 2, 3); sb[i] = make_float4(1, i, 2, 3); }
    __syncthreads();
    float acc[TM][TN];
    for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float a[TM], b[TN];
    for (int i = 0; i < TM; ++i) a[i] = threadIdx.x + i;
    for (int j = 0; j < TN; ++j) b[j] = threadIdx.x * 0.5f + j;
    const int lane = threadIdx.x & 31;
    int ia = (lane >> 2), ib = (lane & 3);
    for (int it = 0; it < iters; ++it) {
        if (WITH_LDS) {
#pragma unroll
            for (int i = 0; i < TM; i += 4) { float4 v = sa[(ia + i * 8 + it) & 511]; a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w; }
#pragma unroll
            for (int j = 0; j < TN; j += 4) { float4 v = sb[(ib + j * 8 + it) & 511]; b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w; }
        } else {
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] += 1.0f;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    float s = 0.f;
    for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) s += acc[i][j];
    if (s == 1234.5f) out[0] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    float* out; cudaMalloc(&out, 4);
    for (int warps = 4; warps <= 16; warps += 4) {
        const int thr = warps * 32;
        float m0 = time_ms([&] { tile_kernel<0, 8, 8><<<sms, thr>>>(out, iters); });
        float m1 = time_ms([&] { tile_kernel<1, 8, 8><<<sms, thr>>>(out, iters); });
        float m2 = time_ms([&] { tile_kernel<1, 4, 12><<<sms, thr>>>(out, iters); });
        double f0 = 2.0 * 64 * iters * (double)thr * sms, f2 = 2.0 * 48 * iters * (double)thr * sms;
        printf("warps/SM=%2d  8x8 regs only: %6.1f TF | 8x8 + 4 LDS.128/step: %6.1f TF | 4x12 + 4 LDS.128/step (16:1... 12 FFMA per LDS): %6.1f TF\n",
               warps, f0 / m0 / 1e9, f0 / m1 / 1e9, f2 / m2 / 1e9);
    }
    return 0;
}
