/*
 * pdelearn_b200.h -- C ABI of the B200-native collocation-loss engine (libpdelearn_b200.so).
 *
 * The reference (punkduckable/PDE-LEARN) has no FFI: its boundary for this path is a set of Python
 * functions.  Each entry point below names the reference interface it replaces (file:line under the
 * reference tree).  All pointers documented "device" are CUDA device pointers owned by the caller; the
 * library never allocates device memory, never synchronises the host and enqueues everything on the
 * stream the caller passes.  Every function returns 0 on success and a non-zero code on failure;
 * pdl_last_error() then returns a human-readable message (thread-local).
 *
 * There is no CPU implementation behind this ABI.  If no CUDA device / compiler is available the calls
 * fail; nothing falls back.
 */
#ifndef PDELEARN_B200_H
#define PDELEARN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDL_ABI_VERSION 3

#define PDL_ACT_TANH 0      /* torch.nn.Tanh            (Code/Classes/Network.py:161-162) */
#define PDL_ACT_RATIONAL 1  /* Rational (3,2), trainable (Code/Classes/Network.py:7-59)   */

#define PDL_MODE_COLLOCATION 0  /* r = T_0(U) - sum_k xi_k T_k(U)   (Code/Loss.py:185-245) */
#define PDL_MODE_DATA 1         /* r = U(X) - target (float64)       (Code/Loss.py:56-62)   */

#define PDL_MAX_PARAM_TENSORS 40

typedef struct pdl_plan pdl_plan;

/*
 * Everything that is fixed for a run: the network architecture and the library of terms.
 * Replaces the (Derivatives, LHS_Term, RHS_Terms) objects built by Read_Library
 * (Code/Readers/Library_Reader.py:201-277) and the Widths / activation of Network.__init__
 * (Code/Classes/Network.py:63-129).
 */
typedef struct {
    int32_t d;                  /* input dimension: 1 time + (d-1) space, 2..4 (Code/Classes/Derivative.py:46) */
    int32_t n_layers;           /* linear layers = hidden layers + 1                                    */
    const int32_t* widths;      /* [n_layers+1]: d, hidden widths..., 1                                  */
    int32_t activation;         /* PDL_ACT_*                                                             */
    int32_t mode;               /* PDL_MODE_*                                                            */
    int32_t n_derivs;           /* distinct derivative operators used by the terms                       */
    const int32_t* encodings;   /* [n_derivs][4]  orders in (t,x,y,z), unused slots 0 (Derivative.Encoding) */
    int32_t n_terms;            /* 1 + number of RHS terms; term 0 is the LHS term                        */
    const int32_t* term_ptr;    /* [n_terms+1] CSR offsets into term_deriv / term_pow                    */
    const int32_t* term_deriv;  /* index into encodings for every sub-term (Term.Derivatives)            */
    const int32_t* term_pow;    /* power >= 1 of every sub-term (Term.Powers)                            */
} pdl_plan_desc;

typedef struct {
    int32_t d, n_jets, n_rhs_terms, n_param_tensors;
    int32_t n_param_scalars;    /* length of the flat gradient vector (reference parameter order)        */
    int32_t threads_per_cta, smem_bytes, warps_per_cta;
    int32_t padded_width, hidden_layers, max_order, mode;
    int64_t flops_per_point;    /* algorithmic flops of one fwd+bwd point: 6[(H-1)w^2 J + wJ] + 4dw */
    int32_t jet_encodings[16][4];  /* the downward-closed jet set, component order used by `features`   */
} pdl_plan_info;

/* One evaluation of the path.  Replaces one Coll_Loss(...)+backward (Code/Loss.py:66-248,
 * Code/Test_Train.py:151-188) or one Data_Loss(...)+backward (Code/Loss.py:23-62). */
typedef struct {
    const float* points;        /* device [n_points][d] row-major float32 (Coll_Points / Inputs)          */
    int64_t n_points;
    const float* const* params; /* HOST array of n_params DEVICE pointers, order of list(U.parameters()):
                                   W_0,b_0,...,W_H,b_H, then (a_i,b_i) per hidden layer when rational     */
    int32_t n_params;
    const float* xi;            /* device [K] float32 (ignored in data mode)                              */
    const uint8_t* mask;        /* device [K]; non-zero = term excluded, its xi gradient is exactly 0     */
    const double* targets;      /* device [n_points] float64, data mode only                              */
    double inv_n;               /* 1/N of the mean; pass 1/N_global when the rows are sharded over ranks  */
    int32_t backward;           /* 1: loss + gradients, 0: loss (+ residual/features) only                */
    float* residual;            /* device [n_points] or NULL                                              */
    float* features;            /* device [n_jets][n_points] or NULL: D_alpha U for every jet component   */
    float* grad_params;         /* device [n_param_scalars] flat, reference order (backward only)         */
    float* grad_xi;             /* device [K] (backward only)                                             */
    double* stats;              /* device [4]: mean loss (sum r^2 * inv_n), sum r^2, sum |r|, reserved    */
    void* workspace;            /* device scratch, >= pdl_plan_workspace_bytes(plan, n_ctas)              */
    size_t workspace_bytes;
    int32_t n_ctas;             /* 0 = one persistent CTA per SM of the current device                    */
    void* stream;               /* cudaStream_t                                                           */
    const int64_t* n_points_dev;/* optional, device int64[1] (ABI 2): the row count is read on the device at kernel start
                                   (clamped to n_points, which becomes the capacity used for grid and workspace sizing) and
                                   the mean uses 1/rows instead of inv_n.  Lets one captured CUDA graph serve epochs whose
                                   number of targeted collocation points differs (Code/main.py:290-297,365-384).        */
} pdl_launch_args;

const char* pdl_last_error(void);
int pdl_abi_version(void);

/* Generated CUDA source of the plan (what gets compiled).  Writes at most cap bytes, always reports the
 * full length in *needed (including the terminating NUL). */
int pdl_plan_source(const pdl_plan_desc* desc, char* buf, size_t cap, size_t* needed);

/* Builds (nvcc, sm_100a) or loads from the on-disk cache the kernels specialised for `desc`.
 * Needs no GPU: compilation is ahead of time, the module is loaded lazily at first launch. */
int pdl_plan_create(const pdl_plan_desc* desc, pdl_plan** out);
void pdl_plan_destroy(pdl_plan* plan);
int pdl_plan_info_get(const pdl_plan* plan, pdl_plan_info* out);
size_t pdl_plan_workspace_bytes(const pdl_plan* plan, int32_t n_ctas);
int pdl_device_sm_count(int32_t* out);

/* Fused forward (+ adjoint) of the collocation or data loss.  See pdl_launch_args. */
int pdl_loss_launch(pdl_plan* plan, const pdl_launch_args* args);

/* L^p surrogate sum_k w_k xi_k^2, w_k = 1/max(1e-7,|xi_k|^(2-p)) detached, 0 where masked, and its
 * gradient 2 w_k xi_k.  Replaces Lp_Loss (Code/Loss.py:252-312).  out_loss: device float[1]. */
int pdl_lp_loss(const float* xi, const uint8_t* mask, int32_t k, double p, float* out_loss, float* out_grad,
                void* stream);

/* Sum of squares of every parameter scalar and gradient 2*theta (flat, reference order).
 * Replaces L2_Squared_Loss (Code/Loss.py:316-361). */
int pdl_l2_squared_loss(const float* const* params, const int32_t* sizes, int32_t n_params, float* out_loss,
                        float* out_grad_flat, void* stream);

/* Uniform points in prod_j [lo_j, hi_j]: Philox4x32-10, counter = first_index + row, so a shard of a
 * batch is the same set of rows for any number of ranks.  Replaces Generate_Points (Code/Points.py:7-59).
 * bounds: HOST [d][2] float64.  out: device [n][d] float32. */
int pdl_sample_points(const double* bounds, int32_t d, int64_t n, uint64_t seed, uint64_t first_index, float* out,
                      void* stream);

/* Same sampler with the Philox key read from device memory at kernel start, so that a captured CUDA graph draws
 * fresh points on every replay (the caller advances *seed_dev between replays).  seed_dev: device uint64[1]. */
int pdl_sample_points_dev(const double* bounds, int32_t d, int64_t n, const uint64_t* seed_dev, uint64_t first_index,
                          float* out, void* stream);

/* The Philox4x32-10 block function the sampler is built on, exposed for known-answer tests (Random123 kat_vectors):
 * counters_keys: device uint32 [n][6] = counter[4], key[2]; out: device uint32 [n][4].  pdl_sample_points uses
 * counter = {row index low, high, 0, 0}, key = {seed low, high} and turns word j into coordinate j as
 * min(fma((word >> 8) * 2^-24, hi_j - lo_j, lo_j), hi_j). */
int pdl_philox4x32_10_raw(const uint32_t* counters_keys, int64_t n, uint32_t* out, void* stream);

/* Targeted collocation points (Code/main.py:365-384): sum and sum of squares of |r| over the first
 * n_random rows (stats: device double[4]; [0], [1] = the sums, [2], [3] = scratch; two-stage reduction in a fixed
 * order, no floating-point atomics: deterministic) and compaction of rows with |r| >= cutoff (order preserved).
 * out_count: device int64[1]; out_points must hold n rows; block_scratch: device int32[(n+1023)/1024 + 1]. */
int pdl_abs_residual_moments(const float* residual, int64_t n_random, double* stats, void* stream);
int pdl_select_targeted(const float* points, const float* residual, int64_t n, int32_t d, float cutoff,
                        float* out_points, int64_t* out_count, int32_t* block_scratch, void* stream);
/* The whole update without a host round trip: moments over the first n_random rows, cutoff = mean + 3*std
 * (unbiased) computed on the device, then the compaction.  stats: device double[4] (sum|r|, sum r^2, cutoff as
 * float in stats[2], and in stats[3] the number of selected rows BEFORE clamping to the output capacity, as an int64 bit
 * pattern: selected > capacity means rows were dropped). */
int pdl_targeted_update(const float* points, const float* residual, int64_t n, int64_t n_random, int32_t d,
                        float* out_points, int64_t* out_count, double* stats, int32_t* block_scratch, void* stream);

/* Same with the row count on the device (n_dev: device int64[1], clamped to `capacity`) and a bounded output:
 * at most out_capacity rows are written and *out_count = min(selected, out_capacity).  For captured epochs. */
int pdl_targeted_update_dev(const float* points, const float* residual, int64_t capacity, const int64_t* n_dev,
                            int64_t n_random, int32_t d, float* out_points, int64_t out_capacity, int64_t* out_count,
                            double* stats, int32_t* block_scratch, void* stream);

/* The selection step alone with everything on the device, for captured epochs that run data-parallel: the moments of |r| over the
 * random rows of ALL ranks arrive as moments_dev = device double[3] {sum |r|, sum r^2, number of random rows} (the caller
 * all-reduces pdl_abs_residual_moments' output and the row counts), the cutoff mean + 3*std (Code/main.py:369-380) is formed on the
 * device, and this rank's rows (count in *n_dev) are compacted as in pdl_targeted_update_dev.  stats: device double[4], [2] receives
 * the cutoff (float), [3] the selected row count before clamping (int64 bit pattern). */
int pdl_select_targeted_dev(const float* points, const float* residual, int64_t capacity, const int64_t* n_dev, int32_t d,
                            const double* moments_dev, float* out_points, int64_t out_capacity, int64_t* out_count,
                            double* stats, int32_t* block_scratch, void* stream);

/* Fused Adam update over a flat parameter vector (torch.optim.Adam semantics, no weight decay / amsgrad):
 * the optimiser step of Code/Test_Train.py:193 for the Adam branch of Code/main.py:232-246. */
int pdl_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, int64_t step, void* stream);

/* The same update for a table of up to 48 tensors in one launch, with the step count read from device memory
 * (step_dev: device float[1], the value torch.optim.Adam keeps in state["step"]; the caller increments it before the
 * call).  theta/grad/exp_avg/exp_avg_sq/sizes: HOST arrays of n_tensors entries (device pointers / element counts). */
int pdl_adam_step_multi(float* const* theta, const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq,
                        const int64_t* sizes, int32_t n_tensors, double lr, double beta1, double beta2, double eps,
                        const float* step_dev, void* stream);

/* out[i] = wa*a[i] + wb*b[i] + wc*c[i] (b, c may be NULL): the weighted sum wC*dColl + wD*dData + wL2*dL2 of one
 * closure (Code/Test_Train.py:167-170) in a single pass over the flat gradient vectors. */
int pdl_axpby3(float* out, int64_t n, const float* a, double wa, const float* b, double wb, const float* c, double wc,
               void* stream);

/* The scalar tail of one closure (Code/Test_Train.py:167-183) in one launch: per network i the total
 * w_data*Data_i + w_coll*Coll_i + w_lp*Lp + w_l2*L2_i (float64), the vector the training loop reports,
 * scalars[4S+1] = {Coll_0.., Data_0.., L2_0.., Total_0.., Lp}, *total = sum_i Total_i (float32, the value .backward() is called on) and
 * -- when grad_xi is not NULL -- the xi gradient  w_coll * sum_i grad_xi_coll[i][:] + lp_grad_scale * grad_xi_lp[:].
 * stats_coll / stats_data: device double [S][4] as written by pdl_loss_launch ([.][0] = the mean); l2: float[S]; lp: float[1];
 * grad_xi_coll: float [S][ld_grad_xi].  Pass the weights already divided by the world size where the term is replicated. */
int pdl_closure_tail(const double* stats_coll, const double* stats_data, const float* l2, const float* lp,
                     const float* grad_xi_coll, const float* grad_xi_lp, int32_t n_networks, int32_t k, int32_t ld_grad_xi,
                     double w_data, double w_coll, double w_lp, double w_l2, double lp_grad_scale, double* scalars, float* total,
                     float* grad_xi, void* stream);

/* Side effect control.  Adjoint launches of >= 2^15 rows keep their workspace in a persisting-L2 window; for that the first such
 * launch on a device reserves the device's persisting-L2 set-aside (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max)), which
 * every other kernel of the process then sees as a smaller L2.  This call returns the set-aside (limit 0) and resets lines still
 * marked persisting -- for a host application that is done training and goes on to other GPU work.  Later adjoint launches stay
 * correct; they run without the window (more DRAM write-back, same numbers) until the process restarts.  PDL_L2_PERSIST=0 in the
 * environment disables the reservation altogether.  No counterpart in the reference (it has no device-side state). */
int pdl_release_l2_setaside(void);

/* Register-resident FFMA-chain microbenchmark: measured FP32 peak of the current device in TFLOP/s
 * (the roofline denominator).  Synchronises the stream. */
int pdl_measure_fp32_peak(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDELEARN_B200_H */
