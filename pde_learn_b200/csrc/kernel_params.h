// kernel_params.h -- argument block shared by the runtime library and every generated plan module.
#pragma once
#define PDL_MAX_PRM 40
struct PdlKernelParams {
    const float* points;            // [n_points][D]
    long long n_points;
    const float* prm[PDL_MAX_PRM];  // device pointers in reference parameter order
    const float* xi;                // [K]
    const unsigned char* mask;      // [K]   nonzero = term excluded (Loss.py:216-218)
    const double* targets;          // [n_points] (data mode only)
    double inv_n;                   // 1 / (rows the mean runs over; the global count when sharded)
    float* residual;                // [n_points] or null
    float* features;                // [J][n_points] or null
    float* zscratch;                // [n_warps][ZSCR]   pre-activation jets parked for the adjoint (L2)
    float* wpriv;                   // [n_warps][WPRIV]  per-warp weight-gradient partial sums
    float* cta_out;                 // [n_ctas][CTA_OUT] per-CTA gradient record
    double* cta_stats;              // [n_ctas][4]       sum r^2, sum |r|
    const long long* n_points_dev;  // optional: row count read on the device (<= n_points = capacity); the mean then uses 1/rows
};
