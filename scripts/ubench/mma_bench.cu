// Microbenchmarks that decide the contraction path (run on a B200 via gpurun):
//   1. mma.sync.m16n8k8 tf32 (legacy warp-level tensor path) dense throughput
//   2. mma.sync.m16n8k16 bf16
//   3. FFMA with register operands (reference)
//   4. LDS.128 broadcast cost: wavefront behaviour seen through time
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int CHAINS>
__global__ void mma_kernel(float* out, int iters) {
    float d[CHAINS][4];
    uint32_t a[4], b[2];
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    for (int i = 0; i < 4; ++i) a[i] = 0x3f800000u + threadIdx.x + i;
    for (int i = 0; i < 2; ++i) b[i] = 0x3f000000u + threadIdx.x * 3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (KIND == 0) mma_tf32(d[c], a, b); else mma_bf16(d[c], a, b);
        }
    }
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
    if (s == 1234.5f) out[0] = s;
}

// mixed: per iteration CHAINS tf32 MMAs and NF FFMAs per thread (do the pipes overlap?)
template <int CHAINS, int NF>
__global__ void mixed_kernel(float* out, int iters) {
    float d[CHAINS][4];
    uint32_t a[4], b[2];
    float x[NF > 0 ? NF : 1];
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    for (int i = 0; i < 4; ++i) a[i] = 0x3f800000u + threadIdx.x + i;
    for (int i = 0; i < 2; ++i) b[i] = 0x3f000000u + threadIdx.x * 3 + i;
    x[0] = 0.f; for (int i = 0; i < NF; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma_tf32(d[c], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) x[i] = fmaf(x[i], 0.999f, 0.001f);
    }
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
    for (int i = 0; i < NF; ++i) s += x[i];
    if (s == 1234.5f) out[0] = s;
}

// LDS.128 where all lanes of a quarter-warp share the address (broadcast) vs distinct addresses
template <int MODE>
__global__ void lds_kernel(float* out, int iters) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, i, i, i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = (MODE == 0) ? 0 : (MODE == 1) ? (lane >> 2) : lane;      // uniform / 8 distinct / 32 distinct
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v = buf[(idx + u * 32) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 1) & 1023;
    }
    if (acc.x + acc.y + acc.z + acc.w == 1234.5f) out[0] = acc.x;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, 4);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int thr = warps * 32;
        float ms = time_ms([&] { mma_kernel<0, 8><<<sms, thr>>>(out, iters); });
        double fl = 2.0 * 16 * 8 * 8 * 8 * (double)iters * warps * sms;
        printf("tf32 m16n8k8  warps/SM=%2d chains=8 : %8.1f TFLOP/s (%.1f FMA/clk/SM @1.965GHz)\n", warps, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / sms / 1.965e9);
        ms = time_ms([&] { mma_kernel<0, 4><<<sms, thr>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 8 * 4 * (double)iters * warps * sms;
        printf("tf32 m16n8k8  warps/SM=%2d chains=4 : %8.1f TFLOP/s\n", warps, fl / ms / 1e9);
        ms = time_ms([&] { mma_kernel<1, 8><<<sms, thr>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 16 * 8 * (double)iters * warps * sms;
        printf("bf16 m16n8k16 warps/SM=%2d chains=8 : %8.1f TFLOP/s\n", warps, fl / ms / 1e9);
    }
    {
        const int warps = 8, thr = 256;
        float ms0 = time_ms([&] { mixed_kernel<8, 0><<<sms, thr>>>(out, iters); });
        float ms1 = time_ms([&] { mixed_kernel<8, 16><<<sms, thr>>>(out, iters); });
        float ms2 = time_ms([&] { mixed_kernel<8, 32><<<sms, thr>>>(out, iters); });
        float ms3 = time_ms([&] { mixed_kernel<8, 64><<<sms, thr>>>(out, iters); });
        printf("mixed 8 MMA + {0,16,32,64} FFMA per iter (8 warps/SM): %.3f %.3f %.3f %.3f ms\n", ms0, ms1, ms2, ms3);
        (void)warps;
    }
    for (int mode = 0; mode < 3; ++mode) {
        float ms;
        if (mode == 0) ms = time_ms([&] { lds_kernel<0><<<sms, 256>>>(out, 4000); });
        else if (mode == 1) ms = time_ms([&] { lds_kernel<1><<<sms, 256>>>(out, 4000); });
        else ms = time_ms([&] { lds_kernel<2><<<sms, 256>>>(out, 4000); });
        double n = 16.0 * 4000 * 8;     // LDS.128 warp-instructions per SM
        printf("LDS.128 mode %d (0 uniform, 1: 8 addrs/warp, 2: 32 addrs): %.2f clk per warp-instruction per SM\n", mode, ms * 1e-3 * 1.965e9 / n);
    }
    return 0;
}
