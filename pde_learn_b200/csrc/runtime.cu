// runtime.cu -- libpdelearn_b200.so: the C ABI declared in include/pdelearn_b200.h.
//
//   * plan cache + JIT: generated plan source -> nvcc (sm_100a) -> plan_<hash>.so -> dlopen
//   * launch of the fused loss kernels with caller-owned workspace
//   * small fixed kernels: L^p and L2 penalties, Philox point sampler, targeted-point compaction,
//     fused Adam, FP32 peak microbenchmark
//
// No CPU fallback anywhere: without nvcc / a CUDA device the calls fail with an error message.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/pdelearn_b200.h"
#include "kernel_params.h"
#include "plan_codegen.hpp"

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define PDL_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

typedef void (*k_info_fn)(int*);
typedef int (*k_launch_fn)(const PdlKernelParams*, int, int, float*, float*, double*, void*);

struct pdl_plan {
    pdl_host::PlanMeta meta;
    void* handle = nullptr;
    k_info_fn info = nullptr;
    k_launch_fn launch = nullptr;
    int kinfo[16] = {0};
    std::string so_path;
};

static std::string lib_dir() {
    Dl_info di;
    if (dladdr((void*)&lib_dir, &di) && di.dli_fname) {
        std::string p(di.dli_fname);
        size_t s = p.rfind('/');
        return s == std::string::npos ? "." : p.substr(0, s);
    }
    return ".";
}
static std::string csrc_dir() {
    const char* e = getenv("PDL_CSRC_DIR");
    if (e && *e) return e;
    return lib_dir() + "/csrc";
}
static std::string plans_dir() {
    const char* e = getenv("PDL_PLAN_DIR");
    if (e && *e) return e;
    return csrc_dir() + "/_plans";
}
static bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && st.st_size > 0; }
static std::string read_file(const std::string& p) {
    std::ifstream f(p);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

static int desc_to_spec(const pdl_plan_desc* d, pdl_host::PlanSpec* s) {
    if (!d) return fail("null plan description");
    if (d->n_layers < 2 || !d->widths) return fail("plan needs >= 2 linear layers and a widths array");
    s->d = d->d;
    s->widths.assign(d->widths, d->widths + d->n_layers + 1);
    s->activation = d->activation;
    s->mode = d->mode;
    for (int i = 0; i < d->n_derivs; ++i) {
        std::array<int, 4> e{d->encodings[4 * i], d->encodings[4 * i + 1], d->encodings[4 * i + 2], d->encodings[4 * i + 3]};
        s->encodings.push_back(e);
    }
    if (d->n_terms < 1 || !d->term_ptr) return fail("plan needs at least the LHS term");
    for (int t = 0; t < d->n_terms; ++t) {
        std::vector<std::pair<int, int>> term;
        for (int q = d->term_ptr[t]; q < d->term_ptr[t + 1]; ++q) term.push_back({d->term_deriv[q], d->term_pow[q]});
        s->terms.push_back(term);
    }
    const char* nw = getenv("PDL_NW");
    if (nw && *nw) s->warps_per_cta = atoi(nw);
    const char* dg = getenv("PDL_DGRAD_MMA");
    if (dg && *dg) s->dgrad_mma = atoi(dg);
    const char* wg = getenv("PDL_WGRAD_MMA");
    if (wg && *wg) s->wgrad_mma = atoi(wg);
    const char* af = getenv("PDL_ADJ_FUSED");
    if (af && *af) s->adj_fused = atoi(af);
    const char* xb = getenv("PDL_CROSS_BF16");
    if (xb && *xb) s->cross_bf16 = atoi(xb);
    const char* tc = getenv("PDL_FWD_TC");
    if (tc && *tc) s->fwd_tc = atoi(tc);
    const char* eng = getenv("PDL_ENGINE");
    if (eng && *eng) s->engine = (eng[0] == 'm' || eng[0] == 'M' || eng[0] == '1') ? 1 : 0;
    return 0;
}

extern "C" {

const char* pdl_last_error(void) { return g_err.c_str(); }
int pdl_abi_version(void) { return PDL_ABI_VERSION; }

int pdl_plan_source(const pdl_plan_desc* desc, char* buf, size_t cap, size_t* needed) {
    try {
        pdl_host::PlanSpec spec;
        if (desc_to_spec(desc, &spec)) return 1;
        std::string src = pdl_host::generate_plan_source(spec, nullptr);
        if (needed) *needed = src.size() + 1;
        if (buf && cap) {
            size_t n = src.size() < cap - 1 ? src.size() : cap - 1;
            memcpy(buf, src.data(), n);
            buf[n] = 0;
        }
        return 0;
    } catch (const std::exception& e) { return fail(e.what()); }
}

int pdl_plan_create(const pdl_plan_desc* desc, pdl_plan** out) {
    if (!out) return fail("null output");
    *out = nullptr;
    try {
        pdl_host::PlanSpec spec;
        if (desc_to_spec(desc, &spec)) return 1;
        pdl_host::PlanMeta meta;
        const std::string src = pdl_host::generate_plan_source(spec, &meta);
        const std::string cdir = csrc_dir(), pdir = plans_dir();
        const std::string hdr = read_file(cdir + (meta.engine == 1 ? "/jet_mlp_kernel_mma.cuh" : "/jet_mlp_kernel.cuh")) +
                                read_file(cdir + "/jet_common.cuh") + read_file(cdir + "/kernel_params.h") + (meta.fwd_tc ? read_file(cdir + "/jet_mlp_fwd_tc.cuh") : std::string());
        if (hdr.size() < 100) return fail("kernel template not found in " + cdir + " (set PDL_CSRC_DIR)");
        const char* xf = getenv("PDL_NVCC_FLAGS");
        const std::string flags = std::string("-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo ") + (xf ? xf : "");
        const char* nv = getenv("PDL_NVCC");
        const std::string nvcc = (nv && *nv) ? nv : "/usr/local/cuda/bin/nvcc";
        // the cache key covers the generated source, the flags, the kernel templates and the toolkit this library was built with
        // (a stale module must not survive a toolkit upgrade); the compiler path is part of it too
        char hx[32];
        snprintf(hx, sizeof hx, "%016llx", (unsigned long long)pdl_host::fnv1a(src + "\n//" + flags + "\n//cuda " + std::to_string(CUDART_VERSION) +
                                                                                " " + nvcc + "\n" + hdr));
        const std::string base = pdir + "/plan_" + hx;
        const std::string so = base + ".so";
        static std::mutex mu;                              // threads of this process
        std::lock_guard<std::mutex> lock(mu);
        if (!file_exists(so)) {
            mkdir(pdir.c_str(), 0755);
            // ranks that start together all miss on the same plan: one of them compiles, the others wait on the plan's lock file
            // and find the installed module (different plans compile side by side: build() uses one process per core).  Source, log
            // and module are written under per-process names and renamed into place.
            const std::string lockp = base + ".lock";
            const int lfd = open(lockp.c_str(), O_CREAT | O_RDWR, 0644);
            if (lfd >= 0) flock(lfd, LOCK_EX);
            struct Unlock { int fd; ~Unlock() { if (fd >= 0) { flock(fd, LOCK_UN); close(fd); } } } unlock{lfd};
            if (!file_exists(so)) {
                const std::string pid = std::to_string((long long)getpid());
                const std::string cu_tmp = base + "." + pid + ".tmp.cu", log_tmp = base + "." + pid + ".tmp.log", tmp = base + "." + pid + ".tmp.so";
                { std::ofstream f(cu_tmp); f << src; }
                auto q = [](const std::string& x) { std::string r = "'"; for (char c : x) { if (c == '\'') r += "'\\''"; else r += c; } return r + "'"; };
                const std::string cmd = q(nvcc) + " " + flags + " -Xptxas -v --shared -Xcompiler -fPIC -I" + q(cdir) + " -o " + q(tmp) + " " +
                                        q(cu_tmp) + " > " + q(log_tmp) + " 2>&1";
                const int rc = system(cmd.c_str());
                if (rc != 0 || !file_exists(tmp)) {
                    std::string log = read_file(log_tmp);
                    if (log.size() > 2000) log = log.substr(log.size() - 2000);
                    return fail("nvcc failed for plan " + std::string(hx) + " (no fallback): " + log);
                }
                rename(cu_tmp.c_str(), (base + ".cu").c_str());
                rename(log_tmp.c_str(), (base + ".log").c_str());
                if (rename(tmp.c_str(), so.c_str()) != 0 && !file_exists(so)) return fail("could not install " + so);
            }
        }
        void* h = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) return fail(std::string("dlopen failed: ") + dlerror());
        pdl_plan* p = new pdl_plan();
        p->meta = meta;
        p->handle = h;
        p->so_path = so;
        p->info = (k_info_fn)dlsym(h, "pdl_k_info");
        p->launch = (k_launch_fn)dlsym(h, "pdl_k_launch");
        if (!p->info || !p->launch) { delete p; return fail("plan module lacks its entry points"); }
        p->info(p->kinfo);
        if (p->kinfo[5] != meta.n_param_tensors || p->kinfo[10] != meta.n_param_scalars || p->kinfo[0] != meta.smem_bytes) {
            std::string m = "plan module disagrees with the generator (params " + std::to_string(p->kinfo[5]) + "/" +
                            std::to_string(meta.n_param_tensors) + ", scalars " + std::to_string(p->kinfo[10]) + "/" +
                            std::to_string(meta.n_param_scalars) + ", smem " + std::to_string(p->kinfo[0]) + "/" +
                            std::to_string(meta.smem_bytes) + ")";
            delete p;
            return fail(m);
        }
        *out = p;
        return 0;
    } catch (const std::exception& e) { return fail(e.what()); }
}

void pdl_plan_destroy(pdl_plan* plan) {
    if (!plan) return;
    // the module stays mapped: kernels may still be in flight on some stream
    delete plan;
}

int pdl_plan_info_get(const pdl_plan* p, pdl_plan_info* o) {
    if (!p || !o) return fail("null argument");
    memset(o, 0, sizeof *o);
    o->d = p->meta.d; o->n_jets = p->meta.n_jets; o->n_rhs_terms = p->meta.n_rhs;
    o->n_param_tensors = p->meta.n_param_tensors; o->n_param_scalars = p->meta.n_param_scalars;
    o->threads_per_cta = p->kinfo[1]; o->smem_bytes = p->kinfo[0]; o->warps_per_cta = p->kinfo[9];
    o->padded_width = p->meta.wp; o->hidden_layers = p->meta.hidden_layers; o->max_order = p->meta.max_order;
    o->mode = p->meta.mode; o->flops_per_point = p->meta.flops_per_point;
    for (int c = 0; c < p->meta.n_jets && c < 16; ++c)
        for (int a = 0; a < 4; ++a) o->jet_encodings[c][a] = p->meta.jets[c][a];
    return 0;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t pdl_plan_workspace_bytes(const pdl_plan* p, int32_t n_ctas) {
    if (!p || n_ctas <= 0) return 0;
    const size_t nw = (size_t)n_ctas * p->kinfo[9];
    return align256(nw * p->kinfo[2] * 4) + align256(nw * p->kinfo[3] * 4) + align256((size_t)n_ctas * p->kinfo[4] * 4) +
           align256((size_t)n_ctas * 4 * 8) + 256;
}

int pdl_device_sm_count(int32_t* out) {
    int dev = 0, n = 0;
    PDL_CUDA(cudaGetDevice(&dev));
    PDL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    *out = n;
    return 0;
}

int pdl_loss_launch(pdl_plan* p, const pdl_launch_args* a) {
    if (!p || !a) return fail("null argument");
    if (a->n_params != p->meta.n_param_tensors)
        return fail("expected " + std::to_string(p->meta.n_param_tensors) + " parameter tensors, got " + std::to_string(a->n_params));
    if (a->n_points < 0) return fail("negative point count");
    if (!a->stats) return fail("stats output is required");
    if (a->backward && (!a->grad_params || (p->meta.n_rhs > 0 && !a->grad_xi))) return fail("backward needs gradient outputs");
    if (p->meta.n_rhs > 0 && (!a->xi || !a->mask)) return fail("xi and mask are required");
    if (p->meta.mode == PDL_MODE_DATA && !a->targets) return fail("data mode needs targets");
    if (a->n_points > 0 && !a->points) return fail("points is null");
    int n_ctas = a->n_ctas;
    if (n_ctas <= 0) { int32_t sm = 0; if (pdl_device_sm_count(&sm)) return 1; n_ctas = sm; }
    const int tp = p->kinfo[12] > 0 ? p->kinfo[12] : 8;
    const long long tiles = (a->n_points + tp - 1) / tp;
    const long long need_ctas = (tiles + p->kinfo[9] - 1) / p->kinfo[9];
    if (need_ctas < n_ctas) n_ctas = (int)(need_ctas > 0 ? need_ctas : 1);
    const size_t need = pdl_plan_workspace_bytes(p, n_ctas);
    if (!a->workspace || a->workspace_bytes < need)
        return fail("workspace too small: need " + std::to_string(need) + " bytes, got " + std::to_string(a->workspace_bytes));
    PdlKernelParams P;
    memset(&P, 0, sizeof P);
    P.points = a->points; P.n_points = a->n_points;
    for (int i = 0; i < a->n_params; ++i) {
        if (!a->params[i]) return fail("null parameter pointer");
        P.prm[i] = a->params[i];
    }
    P.xi = a->xi; P.mask = a->mask; P.targets = a->targets; P.inv_n = a->inv_n;
    P.residual = a->residual; P.features = a->features;
    P.n_points_dev = (const long long*)a->n_points_dev;
    char* w = (char*)(((uintptr_t)a->workspace + 255) & ~(uintptr_t)255);
    const size_t nw = (size_t)n_ctas * p->kinfo[9];
    P.zscratch = (float*)w; w += align256(nw * p->kinfo[2] * 4);
    P.wpriv = (float*)w; w += align256(nw * p->kinfo[3] * 4);
    P.cta_out = (float*)w; w += align256((size_t)n_ctas * p->kinfo[4] * 4);
    P.cta_stats = (double*)w;
    const int rc = p->launch(&P, n_ctas, a->backward ? 1 : 0, a->grad_params, a->grad_xi, a->stats, a->stream);
    if (rc != 0) return fail(std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    return 0;
}

}  // extern "C"

// =============================================================================================
// small fixed kernels
// =============================================================================================

// ---- L^p surrogate (Code/Loss.py:252-312) ----
__global__ void lp_kernel(const float* __restrict__ xi, const unsigned char* __restrict__ mask, int k, double p,
                          float* __restrict__ out_loss, float* __restrict__ out_grad) {
    __shared__ double part[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        float w = 0.f;
        if (!mask[i]) {
            const double a = fabs((double)xi[i]);
            const double den = fmax(1e-7, pow(a, 2.0 - p));
            w = (float)(1.0 / den);                          // stored as float32 in the reference (W tensor)
        }
        const float x2 = xi[i] * xi[i];
        acc += (double)(w * x2);
        if (out_grad) out_grad[i] = 2.f * w * xi[i];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += part[w];
        out_loss[0] = (float)s;
    }
}

// ---- L2 penalty (Code/Loss.py:316-361) ----
struct L2Args { const float* ptr[PDL_MAX_PARAM_TENSORS]; int size[PDL_MAX_PARAM_TENSORS]; int off[PDL_MAX_PARAM_TENSORS]; int n; int total; };
__global__ void l2_kernel(const L2Args A, float* __restrict__ out_loss, float* __restrict__ out_grad) {
    __shared__ double part[32];
    double acc = 0.0;
    for (int t = 0; t < A.n; ++t) {
        for (int i = threadIdx.x; i < A.size[t]; i += blockDim.x) {
            const float v = A.ptr[t][i];
            acc += (double)(v * v);
            if (out_grad) out_grad[A.off[t] + i] = 2.f * v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += part[w];
        out_loss[0] = (float)s;
    }
}

// ---- Philox4x32-10 sampler (replaces Code/Points.py:45-57) ----
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
// raw Philox4x32-10 blocks (known-answer tests, Random123 kat_vectors): in [n][6] = counter[4], key[2]; out [n][4]
__global__ void philox_raw_kernel(const uint32_t* __restrict__ in, long long n, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c[4] = {in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3]};
    philox4x32_10(c, in[6 * i + 4], in[6 * i + 5]);
    for (int j = 0; j < 4; ++j) out[4 * i + j] = c[j];
}
struct SampleArgs { float lo[4]; float span[4]; float hi[4]; };
__global__ void sample_kernel(const SampleArgs A, int d, long long n, unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                              unsigned long long first, float* __restrict__ out) {
    if (seed_dev) seed = seed_dev[0];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long idx = first + (unsigned long long)i;
        uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int j = 0; j < d; ++j) {
            const float u = (float)(c[j] >> 8) * (1.0f / 16777216.0f);          // [0, 1)
            out[i * d + j] = fminf(fmaf(u, A.span[j], A.lo[j]), A.hi[j]);
        }
    }
}

// ---- targeted collocation points (Code/main.py:365-384) ----
// Two stages, no floating-point atomics: block b sums its grid-strided rows in a fixed order and writes (sum |r|, sum r^2) to
// partial[b]; one warp then adds the partials in index order.  Two launches over the same rows give the same bits, so the targeted
// cutoff is deterministic like everything else on the path.
__global__ void moments_partial_kernel(const float* __restrict__ r, long long n, double* __restrict__ partial) {
    double s1 = 0.0, s2 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = fabs((double)r[i]);
        s1 += a; s2 += a * a;
    }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    __shared__ double p1[32], p2[32];
    if ((threadIdx.x & 31) == 0) { p1[threadIdx.x >> 5] = s1; p2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) { a += p1[w]; b += p2[w]; }
        partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
    }
}
__global__ void moments_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ stats) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    // fixed butterfly order
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (threadIdx.x == 0) { stats[0] = a; stats[1] = b; }
}
__global__ void cutoff_kernel(const double* __restrict__ stats, long long n_random, float* __restrict__ out) {
    // mean + 3 * unbiased std of |r| over the random rows (Code/main.py:369-380)
    const double n = (double)n_random;
    const double mean = n > 0 ? stats[0] / n : 0.0;
    double var = n > 1 ? (stats[1] - n * mean * mean) / (n - 1.0) : 0.0;
    if (var < 0.0) var = 0.0;
    out[0] = (float)(mean + 3.0 * sqrt(var));
}
// the same cutoff from moments that live in device memory as three doubles (sum |r|, sum r^2, row count): the all-reduced moments
// of a batch whose random rows are sharded over ranks (captured multi-process epochs)
__global__ void cutoff_from_moments_kernel(const double* __restrict__ moments, float* __restrict__ out) {
    const double n = moments[2];
    const double mean = n > 0 ? moments[0] / n : 0.0;
    double var = n > 1 ? (moments[1] - n * mean * mean) / (n - 1.0) : 0.0;
    if (var < 0.0) var = 0.0;
    out[0] = (float)(mean + 3.0 * sqrt(var));
}
__global__ void select_count_kernel(const float* __restrict__ r, long long n, const long long* __restrict__ n_dev, float cutoff,
                                    const float* __restrict__ cutoff_dev, int* __restrict__ counts) {
    if (cutoff_dev) cutoff = cutoff_dev[0];
    if (n_dev) { const long long m = n_dev[0]; n = m < n ? (m < 0 ? 0 : m) : n; }
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int flag = (i < n && fabsf(r[i]) >= cutoff) ? 1 : 0;
    const int c = __syncthreads_count(flag);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}
__global__ void select_scan_kernel(int* __restrict__ counts, int n_blocks, long long* __restrict__ total, long long cap,
                                   long long* __restrict__ unclamped) {
    // single block: exclusive scan of counts[0..n_blocks) in place, total to *total
    __shared__ long long carry;
    __shared__ int wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < n_blocks) ? counts[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
        const long long c0 = carry;
        if (i < n_blocks) counts[i] = (int)(c0 + before + x - v);
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { *total = carry < cap ? carry : cap; if (unclamped) *unclamped = carry; }
}
__global__ void select_scatter_kernel(const float* __restrict__ pts, const float* __restrict__ r, long long n,
                                      const long long* __restrict__ n_dev, int d, float cutoff, const float* __restrict__ cutoff_dev,
                                      const int* __restrict__ offsets, float* __restrict__ out, long long cap) {
    __shared__ int wsum[32];
    if (cutoff_dev) cutoff = cutoff_dev[0];
    if (n_dev) { const long long m = n_dev[0]; n = m < n ? (m < 0 ? 0 : m) : n; }
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int flag = (i < n && fabsf(r[i]) >= cutoff) ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int within = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = wsum[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
        wsum[threadIdx.x] = w;
    }
    __syncthreads();
    if (flag) {
        const long long dst = (long long)offsets[blockIdx.x] + (warp ? wsum[warp - 1] : 0) + within;
        if (dst < cap) { for (int j = 0; j < d; ++j) out[dst * d + j] = pts[i * d + j]; }
    }
}

// ---- fused Adam (torch.optim.Adam, default flags) ----
__global__ void adam_kernel(float* __restrict__ th, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_c, float b1, float b2, float inv_sqrt_c2, float eps) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        th[i] -= lr_c * mi / (sqrtf(vi) * inv_sqrt_c2 + eps);
    }
}

// same update over a table of tensors; the step count lives on the device (captured epochs: the bias corrections change per replay)
#define PDL_ADAM_MAX_TENSORS 48
struct AdamMultiArgs {
    float* th[PDL_ADAM_MAX_TENSORS]; const float* g[PDL_ADAM_MAX_TENSORS]; float* m[PDL_ADAM_MAX_TENSORS]; float* v[PDL_ADAM_MAX_TENSORS];
    long long n[PDL_ADAM_MAX_TENSORS];
};
__global__ void adam_multi_kernel(const AdamMultiArgs A, const float* __restrict__ step_dev, double lr, double b1d, double b2d, float eps) {
    const int t = blockIdx.y;
    const double step = (double)step_dev[0];
    const double c1 = 1.0 - pow(b1d, step), c2 = 1.0 - pow(b2d, step);
    const float lr_c = (float)(lr / c1), inv_sqrt_c2 = (float)(1.0 / sqrt(c2)), b1 = (float)b1d, b2 = (float)b2d;
    float* th = A.th[t]; const float* g = A.g[t]; float* m = A.m[t]; float* v = A.v[t];
    const long long n = A.n[t];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        th[i] -= lr_c * mi / (sqrtf(vi) * inv_sqrt_c2 + eps);
    }
}

// ---- out = wa*a + wb*b + wc*c (b, c optional): combines the per-term gradients of one closure ----
__global__ void axpby3_kernel(float* __restrict__ out, long long n, const float* __restrict__ a, float wa,
                              const float* __restrict__ b, float wb, const float* __restrict__ c, float wc) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = wa * a[i];
        if (b) v += wb * b[i];
        if (c) v += wc * c[i];
        out[i] = v;
    }
}

// ---- the scalar tail of one closure (Code/Test_Train.py:167-183): per-network totals, the reported loss vector, the summed total and
// the weighted xi gradient  w_coll * sum_i dColl_i/dxi + lp_scale * dLp/dxi  -- one launch instead of a dozen tensor-library kernels ----
__global__ void closure_tail_kernel(const double* __restrict__ stats_c, const double* __restrict__ stats_d, const float* __restrict__ l2,
                                    const float* __restrict__ lp, const float* __restrict__ gx, const float* __restrict__ g_lp, int S, int K,
                                    int ldk, double w_data, double w_coll, double w_lp, double w_l2, float lp_scale,
                                    double* __restrict__ scalars, float* __restrict__ total, float* __restrict__ grad_xi) {
    if (grad_xi) {
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float s = 0.f;
            for (int i = 0; i < S; ++i) s += gx[(size_t)i * ldk + k];
            grad_xi[k] = s * (float)w_coll + lp_scale * g_lp[k];
        }
    }
    if (threadIdx.x == 0) {
        const double lpd = (double)lp[0];
        double sum = 0.0;
        for (int i = 0; i < S; ++i) {
            const double coll = stats_c[4 * i], data = stats_d[4 * i], l2d = (double)l2[i];
            const double tot = w_data * data + w_coll * coll + w_lp * lpd + w_l2 * l2d;
            scalars[i] = coll; scalars[S + i] = data; scalars[2 * S + i] = l2d; scalars[3 * S + i] = tot;
            sum += tot;
        }
        scalars[4 * S] = lpd;
        total[0] = (float)sum;
    }
}

// ---- FP32 peak: 8 independent FFMA chains per thread ----
__global__ void fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) out[0] = s;
}

extern "C" {

int pdl_lp_loss(const float* xi, const uint8_t* mask, int32_t k, double p, float* out_loss, float* out_grad, void* stream) {
    if (!(p > 0.0 && p < 2.0)) return fail("Lp_Loss requires 0 < p < 2 (Loss.py:279)");
    if (k <= 0 || !xi || !mask || !out_loss) return fail("bad Lp arguments");
    lp_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(xi, mask, k, p, out_loss, out_grad);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

int pdl_l2_squared_loss(const float* const* params, const int32_t* sizes, int32_t n, float* out_loss, float* out_grad, void* stream) {
    if (n <= 0 || n > PDL_MAX_PARAM_TENSORS || !params || !sizes || !out_loss) return fail("bad L2 arguments");
    L2Args A;
    memset(&A, 0, sizeof A);
    int off = 0;
    for (int t = 0; t < n; ++t) { A.ptr[t] = params[t]; A.size[t] = sizes[t]; A.off[t] = off; off += sizes[t]; }
    A.n = n; A.total = off;
    l2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A, out_loss, out_grad);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

static int sample_points_impl(const double* bounds, int32_t d, int64_t n, uint64_t seed, const uint64_t* seed_dev, uint64_t first,
                              float* out, void* stream) {
    if (d < 1 || d > 4 || !bounds || (n > 0 && !out)) return fail("bad sampler arguments");
    SampleArgs A;
    for (int j = 0; j < 4; ++j) { A.lo[j] = 0.f; A.span[j] = 0.f; A.hi[j] = 0.f; }
    for (int j = 0; j < d; ++j) {
        if (!(bounds[2 * j] <= bounds[2 * j + 1])) return fail("Generate_Points requires Bounds[j,0] <= Bounds[j,1] (Points.py:41-42)");
        A.lo[j] = (float)bounds[2 * j]; A.hi[j] = (float)bounds[2 * j + 1]; A.span[j] = A.hi[j] - A.lo[j];
    }
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sample_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(A, d, n, seed, (const unsigned long long*)seed_dev, first, out);
    PDL_CUDA(cudaGetLastError());
    return 0;
}
int pdl_sample_points(const double* bounds, int32_t d, int64_t n, uint64_t seed, uint64_t first, float* out, void* stream) {
    return sample_points_impl(bounds, d, n, seed, nullptr, first, out, stream);
}
int pdl_sample_points_dev(const double* bounds, int32_t d, int64_t n, const uint64_t* seed_dev, uint64_t first, float* out, void* stream) {
    if (!seed_dev) return fail("seed_dev is null");
    return sample_points_impl(bounds, d, n, 0, seed_dev, first, out, stream);
}

// partial: device scratch of 2 * max_blocks doubles (8-byte aligned)
static int moments_impl(const float* residual, int64_t n, double* stats, double* partial, long long max_blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { PDL_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st)); return 0; }
    long long blocks = (n + 4095) / 4096;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    moments_partial_kernel<<<(int)blocks, 1024, 0, st>>>(residual, n, partial);
    moments_final_kernel<<<1, 32, 0, st>>>(partial, (int)blocks, stats);
    PDL_CUDA(cudaGetLastError());
    return 0;
}
int pdl_abs_residual_moments(const float* residual, int64_t n, double* stats, void* stream) {
    if (!residual || !stats || n < 0) return fail("bad arguments");
    return moments_impl(residual, n, stats, stats + 2, 1, stream);      // stats[2..3] is the one-block partial record
}

int pdl_philox4x32_10_raw(const uint32_t* counters_keys, int64_t n, uint32_t* out, void* stream) {
    if (!counters_keys || !out || n < 0) return fail("bad arguments");
    if (n == 0) return 0;
    philox_raw_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(counters_keys, n, out);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

int pdl_select_targeted(const float* points, const float* residual, int64_t n, int32_t d, float cutoff, float* out_points,
                        int64_t* out_count, int32_t* block_scratch, void* stream) {
    if (!out_count || !block_scratch || n < 0) return fail("bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { PDL_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st)); return 0; }
    const int nb = (int)((n + 1023) / 1024);
    select_count_kernel<<<nb, 1024, 0, st>>>(residual, n, nullptr, cutoff, nullptr, block_scratch);
    select_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nb, (long long*)out_count, n, nullptr);
    select_scatter_kernel<<<nb, 1024, 0, st>>>(points, residual, n, nullptr, d, cutoff, nullptr, block_scratch, out_points, n);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

static int targeted_update_impl(const float* points, const float* residual, int64_t n, const int64_t* n_dev, int64_t n_random, int32_t d,
                                float* out_points, int64_t out_cap, int64_t* out_count, double* stats, int32_t* block_scratch, void* stream) {
    if (!out_count || !block_scratch || !stats || n < 0 || n_random < 0 || n_random > n || out_cap < 0) return fail("bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { PDL_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st)); return 0; }
    {   // the block scratch (int32[(n+1023)/1024 + 1], free until the selection below) hosts the per-block partial sums
        const long long words = (n + 1023) / 1024 + 1;
        double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(block_scratch) + 7) & ~(uintptr_t)7);
        const long long fit = (words * 4 - 8) / 16;
        if (fit >= 1) { if (moments_impl(residual, n_random, stats, partial, fit, stream)) return 1; }
        else if (moments_impl(residual, n_random, stats, stats + 2, 1, stream)) return 1;
    }
    float* cutoff_dev = reinterpret_cast<float*>(stats + 2);
    cutoff_kernel<<<1, 1, 0, st>>>(stats, n_random, cutoff_dev);
    const int nb = (int)((n + 1023) / 1024);
    select_count_kernel<<<nb, 1024, 0, st>>>(residual, n, (const long long*)n_dev, 0.f, cutoff_dev, block_scratch);
    select_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nb, (long long*)out_count, out_cap, reinterpret_cast<long long*>(stats + 3));
    select_scatter_kernel<<<nb, 1024, 0, st>>>(points, residual, n, (const long long*)n_dev, d, 0.f, cutoff_dev, block_scratch, out_points, out_cap);
    PDL_CUDA(cudaGetLastError());
    return 0;
}
int pdl_select_targeted_dev(const float* points, const float* residual, int64_t capacity, const int64_t* n_dev, int32_t d,
                            const double* moments_dev, float* out_points, int64_t out_capacity, int64_t* out_count, double* stats,
                            int32_t* block_scratch, void* stream) {
    if (!n_dev || !moments_dev || !out_count || !block_scratch || !stats || capacity < 0 || out_capacity < 0) return fail("bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (capacity == 0) { PDL_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st)); return 0; }
    float* cutoff_dev = reinterpret_cast<float*>(stats + 2);
    cutoff_from_moments_kernel<<<1, 1, 0, st>>>(moments_dev, cutoff_dev);
    const int nb = (int)((capacity + 1023) / 1024);
    select_count_kernel<<<nb, 1024, 0, st>>>(residual, capacity, (const long long*)n_dev, 0.f, cutoff_dev, block_scratch);
    select_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nb, (long long*)out_count, out_capacity, reinterpret_cast<long long*>(stats + 3));
    select_scatter_kernel<<<nb, 1024, 0, st>>>(points, residual, capacity, (const long long*)n_dev, d, 0.f, cutoff_dev, block_scratch, out_points,
                                               out_capacity);
    PDL_CUDA(cudaGetLastError());
    return 0;
}
int pdl_targeted_update(const float* points, const float* residual, int64_t n, int64_t n_random, int32_t d, float* out_points,
                        int64_t* out_count, double* stats, int32_t* block_scratch, void* stream) {
    return targeted_update_impl(points, residual, n, nullptr, n_random, d, out_points, n, out_count, stats, block_scratch, stream);
}
int pdl_targeted_update_dev(const float* points, const float* residual, int64_t capacity, const int64_t* n_dev, int64_t n_random, int32_t d,
                            float* out_points, int64_t out_capacity, int64_t* out_count, double* stats, int32_t* block_scratch,
                            void* stream) {
    if (!n_dev) return fail("n_dev is null");
    return targeted_update_impl(points, residual, capacity, n_dev, n_random, d, out_points, out_capacity, out_count, stats, block_scratch, stream);
}

int pdl_release_l2_setaside(void) {
    // undo the process-wide side effect of the adjoint launches (kernel_params.h: pdl_l2_window reserves the persisting-L2 set-aside once
    // per device): hand the whole L2 back to normal caching and drop lines that are still marked persisting
    PDL_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
    PDL_CUDA(cudaCtxResetPersistingL2Cache());
    return 0;
}

int pdl_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n, double lr, double b1, double b2, double eps,
                  int64_t step, void* stream) {
    if (!theta || !grad || !m || !v || n <= 0 || step < 1) return fail("bad Adam arguments");
    const double c1 = 1.0 - pow(b1, (double)step), c2 = 1.0 - pow(b2, (double)step);
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, n, (float)(lr / c1), (float)b1, (float)b2,
                                                               (float)(1.0 / sqrt(c2)), (float)eps);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

int pdl_adam_step_multi(float* const* theta, const float* const* grad, float* const* m, float* const* v, const int64_t* sizes,
                        int32_t n_tensors, double lr, double b1, double b2, double eps, const float* step_dev, void* stream) {
    if (!theta || !grad || !m || !v || !sizes || !step_dev || n_tensors <= 0 || n_tensors > PDL_ADAM_MAX_TENSORS)
        return fail("bad Adam arguments (at most 48 tensors per call)");
    AdamMultiArgs A;
    memset(&A, 0, sizeof A);
    long long nmax = 0;
    for (int t = 0; t < n_tensors; ++t) {
        if (!theta[t] || !grad[t] || !m[t] || !v[t] || sizes[t] <= 0) return fail("bad Adam tensor");
        A.th[t] = theta[t]; A.g[t] = grad[t]; A.m[t] = m[t]; A.v[t] = v[t]; A.n[t] = sizes[t];
        if (sizes[t] > nmax) nmax = sizes[t];
    }
    long long bx = (nmax + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)n_tensors), 256, 0, (cudaStream_t)stream>>>(A, step_dev, lr, b1, b2, (float)eps);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

int pdl_axpby3(float* out, int64_t n, const float* a, double wa, const float* b, double wb, const float* c, double wc, void* stream) {
    if (!out || !a || n <= 0) return fail("bad axpby3 arguments");
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    axpby3_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, n, a, (float)wa, b, (float)wb, c, (float)wc);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

int pdl_closure_tail(const double* stats_coll, const double* stats_data, const float* l2, const float* lp, const float* grad_xi_coll,
                     const float* grad_xi_lp, int32_t n_networks, int32_t k, int32_t ld_grad_xi, double w_data, double w_coll, double w_lp,
                     double w_l2, double lp_grad_scale, double* scalars, float* total, float* grad_xi, void* stream) {
    if (!stats_coll || !stats_data || !l2 || !lp || !scalars || !total || n_networks < 1 || k < 0) return fail("bad arguments");
    if (grad_xi && (!grad_xi_coll || !grad_xi_lp || ld_grad_xi < k)) return fail("bad gradient arguments");
    closure_tail_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(stats_coll, stats_data, l2, lp, grad_xi_coll, grad_xi_lp, n_networks, k, ld_grad_xi,
                                                             w_data, w_coll, w_lp, w_l2, (float)lp_grad_scale, scalars, total, grad_xi);
    PDL_CUDA(cudaGetLastError());
    return 0;
}

int pdl_measure_fp32_peak(double* tflops, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int32_t sms = 0;
    if (pdl_device_sm_count(&sms)) return 1;
    float* dummy = nullptr;
    PDL_CUDA(cudaMalloc(&dummy, 4));
    cudaEvent_t e0, e1;
    PDL_CUDA(cudaEventCreate(&e0)); PDL_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        PDL_CUDA(cudaEventRecord(e0, st));
        fp32_peak_kernel<<<blocks, threads, 0, st>>>(dummy, iters, 0.999f, 0.001f);
        PDL_CUDA(cudaEventRecord(e1, st));
        PDL_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PDL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dummy);
    *tflops = best;
    return 0;
}

}  // extern "C"
