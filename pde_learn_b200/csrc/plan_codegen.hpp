// plan_codegen.hpp -- host-side plan description and the code generator interface.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <utility>
#include <vector>

namespace pdl_host {

struct PlanSpec {
    int d = 2;
    std::vector<int> widths;                                 // d, hidden..., 1
    int activation = 0;                                      // 0 tanh, 1 rational
    int mode = 0;                                            // 0 collocation, 1 data
    std::vector<std::array<int, 4>> encodings;               // library derivatives, padded to 4 axes
    std::vector<std::vector<std::pair<int, int>>> terms;     // term 0 = LHS; (derivative index, power)
    int warps_per_cta = 0;                                   // 0 = default
    int dgrad_mma = -1;                                      // tensor-pipe engine: -1 = if it fits, 0 = FP32-pipe data gradient
    int wgrad_mma = -1;                                      // tensor-pipe engine: weight gradient as 3xTF32 MMAs too (-1 = choose)
    int engine = -1;                                         // -1 = choose, 0 = FP32-pipe kernel, 1 = tensor-pipe (mma.sync) kernel
    int adj_fused = -1;                                      // tensor-pipe engine: order of the hidden-layer adjoint (-1 = choose)
    int cross_bf16 = -1;                                     // tensor-pipe engine: hi*lo + lo*hi as ONE bf16 m16n8k16 MMA (-1 = choose)
    int fwd_tc = -1;                                         // launches without gradients on the tcgen05/TMEM forward engine (-1 = if it fits)
};

struct PlanMeta {
    int d = 0, n_jets = 0, n_rhs = 0, hidden_layers = 0, wp = 0, max_order = 0, mode = 0, activation = 0;
    int n_param_tensors = 0, n_param_scalars = 0, sf = 0, gs = 0, nw = 0, smem_bytes = 0;
    int engine = 0, dgrad_mma = 0, wgrad_mma = 0, cross_bf16 = 0, adj_fused = 0, tile_points = 8, fwd_tc = 0;
    long long flops_per_point = 0;
    std::vector<std::array<int, 4>> jets;
};

uint64_t fnv1a(const std::string& s);
int smem_floats(const PlanMeta& m, int nw);
std::string generate_plan_source(const PlanSpec& spec, PlanMeta* meta_out);

}  // namespace pdl_host
