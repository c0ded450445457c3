// plan_codegen.cpp -- plan compiler: (architecture, library of terms) -> CUDA source specialised for it.
//
// Input is what the reference builds in Read_Library (Code/Readers/Library_Reader.py:201-277) and
// Network.__init__ (Code/Classes/Network.py:63-129).  Output is a translation unit that defines
//   * the downward-closed jet set A (every D_beta U with beta <= some library derivative),
//   * straight-line multivariate Faa-di-Bruno code for the activation jets and their adjoint
//     (replaces the nested autograd.grad chain of Code/Evaluate_Derivatives.py:104-205 and its backward),
//   * straight-line code for the library terms / residual / seeds (Code/Loss.py:185-245),
// and then includes jet_mlp_kernel.cuh, the hand-written kernel template.
#include "plan_codegen.hpp"

#include <algorithm>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

namespace pdl_host {

namespace {

struct Partition {
    long long coef;
    std::vector<int> blocks;   // jet component index of every block, sorted
};

std::string prod_expr(const std::vector<int>& comps, const char* name) {
    if (comps.empty()) return "1.f";
    std::ostringstream o;
    for (size_t i = 0; i < comps.size(); ++i) {
        if (i) o << "*";
        o << name << "[" << comps[i] << "]";
    }
    return o.str();
}

std::string fnum(long long c) {
    std::ostringstream o;
    o << c << ".f";
    return o.str();
}

// set partitions of the multiset of differentiation variables of `enc`, grouped by block multiset
std::vector<Partition> partitions_of(const std::array<int, 4>& enc, const std::map<std::array<int, 4>, int>& index) {
    std::vector<int> vars;
    for (int a = 0; a < 4; ++a)
        for (int q = 0; q < enc[a]; ++q) vars.push_back(a);
    const int n = (int)vars.size();
    std::map<std::vector<int>, long long> count;
    std::vector<int> rgs(n, 0);
    // enumerate restricted growth strings
    std::function<void(int, int)> rec = [&](int pos, int nblocks) {
        if (pos == n) {
            std::vector<std::array<int, 4>> blk(nblocks, std::array<int, 4>{0, 0, 0, 0});
            for (int i = 0; i < n; ++i) blk[rgs[i]][vars[i]] += 1;
            std::vector<int> comps;
            for (auto& b : blk) {
                auto it = index.find(b);
                if (it == index.end()) throw std::runtime_error("jet set is not downward closed");
                comps.push_back(it->second);
            }
            std::sort(comps.begin(), comps.end());
            count[comps] += 1;
            return;
        }
        for (int b = 0; b <= nblocks; ++b) {
            rgs[pos] = b;
            rec(pos + 1, std::max(nblocks, b + 1));
        }
    };
    if (n > 0) rec(0, 0);
    std::vector<Partition> out;
    for (auto& kv : count) out.push_back(Partition{kv.second, kv.first});
    return out;
}

// Q_k with sigma_k = (1 - s^2) Q_k(s) for tanh:  Q_1 = 1,  Q_{k+1} = -2 s Q_k + (1 - s^2) Q_k'
std::vector<std::vector<long long>> tanh_Q(int kmax) {
    std::vector<std::vector<long long>> Q(kmax + 1);
    if (kmax >= 1) Q[1] = {1};
    for (int k = 1; k < kmax; ++k) {
        const auto& q = Q[k];
        std::vector<long long> nq(q.size() + 1, 0);
        for (size_t i = 0; i < q.size(); ++i) nq[i + 1] += -2 * q[i];          // -2 s Q
        for (size_t i = 1; i < q.size(); ++i) {                                // (1 - s^2) Q'
            const long long dq = (long long)i * q[i];
            nq[i - 1] += dq;
            nq[i + 1] -= dq;
        }
        Q[k + 1] = nq;
    }
    return Q;
}

int find_sf(int wp, int npl) {
    for (int m = wp; m < wp + 64; m += 4) {
        std::set<int> q;
        for (int g = 0; g < 4; ++g) q.insert((g * npl * (m / 4)) % 8);
        if (q.size() == 4) return m;
    }
    return wp + 4;
}
int find_gs(int npl) {
    const int npl4 = (npl + 3) / 4 * 4;
    for (int m = npl4; m < npl4 + 64; m += 4) {
        std::set<int> q;
        for (int g = 0; g < 4; ++g) q.insert((g * (m / 4)) % 8);
        if (q.size() == 4) return m;
    }
    return npl4;
}
int pad_stride(int n) {
    int m = (n + 3) / 4 * 4;
    while (((m / 4) & 1) == 0) m += 4;
    return m;
}

}  // namespace

uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

int smem_floats(const PlanMeta& m, int nw) {
    const int WP = m.wp, D = m.d, J = m.n_jets, H = m.hidden_layers, KX = std::max(m.n_rhs, 1);
    if (m.engine == 1) {
        const int JA = (J >= 3) ? 4 : J;
        const int JB = (J <= 4) ? 0 : ((J == 5) ? 1 : ((J == 6) ? 2 : ((J <= 8) ? 4 : 8)));
        const int SA = pad_stride(WP * JA), SB = JB ? pad_stride(WP * JB) : 0;
        const int SD = 4 * m.gs, SF = WP, TP = 8;
        const int NHH = H > 1 ? H - 1 : 0;
        const int SO_XI = WP * D + WP + WP + 1 + 7 * H;
        const int SACC = (SO_XI + KX * TP + 3) / 4 * 4;
        const int SM_HH = (WP * D + WP + 3) / 4 * 4;
        // weight planes per hidden->hidden layer: W hi + W lo (or the packed bf16 cross plane), then W^T hi + lo/cross (mode 1), or only
        // the W^T cross plane (mode 2 with bf16 cross terms), or the FP32-pipe dgrad copy
        const int HH_B = (m.dgrad_mma == 1) ? 4 * WP * SF
                       : (m.dgrad_mma ? (((m.cross_bf16 & 2) ? 3 : 2) * WP * SF) : (2 * WP * SF + WP * SD));
        const int HH_SIZE = (HH_B + WP + 3) / 4 * 4;
        const int SM_WH = SM_HH + NHH * HH_SIZE;
        const int SM_RAT = (SM_WH + WP + 1 + 3) / 4 * 4;
        const int SM_SHARED = (SM_RAT + 8 * H + 2 * KX + 3) / 4 * 4;
        const int BUF = std::max(8 * (SA + SB), 128);
        return SM_SHARED + nw * (2 * BUF + SACC);
    }
    const int NPL = WP / 4;
    const int JA = (J >= 3) ? 4 : J;
    const int JB = (J <= 4) ? 0 : ((J == 5) ? 1 : ((J == 6) ? 2 : ((J <= 8) ? 4 : 8)));
    const int SA = pad_stride(WP * JA), SB = JB ? pad_stride(WP * JB) : 0;
    const int SD = 4 * m.gs;
    const int NHH = H > 1 ? H - 1 : 0;
    const int SO_XI = WP * D + WP + WP + 1 + 7 * H;
    const int SACC = (SO_XI + KX * 8 + 3) / 4 * 4;
    const int SM_HH = (WP * D + WP + 3) / 4 * 4;
    const int HH_SIZE = (WP * m.sf + WP * SD + WP + 3) / 4 * 4;
    const int SM_WH = SM_HH + NHH * HH_SIZE;
    const int SM_RAT = (SM_WH + WP + 1 + 3) / 4 * 4;
    const int SM_SHARED = (SM_RAT + 8 * H + 2 * KX + 3) / 4 * 4;
    const int BUF = std::max(8 * (SA + SB), 128);
    (void)NPL;
    return SM_SHARED + nw * (2 * BUF + SACC);
}

std::string generate_plan_source(const PlanSpec& spec, PlanMeta* meta_out) {
    if (spec.d < 2 || spec.d > 4) throw std::runtime_error("input dimension must be 2, 3 or 4 (Derivative.py:46)");
    if (spec.widths.size() < 3) throw std::runtime_error("need at least one hidden layer");
    if (spec.widths.front() != spec.d) throw std::runtime_error("widths[0] must equal the input dimension");
    if (spec.widths.back() != 1) throw std::runtime_error("the system-response network has one output");
    if (spec.activation != 0 && spec.activation != 1)
        throw std::runtime_error("unsupported activation (tanh and rational are implemented; no fallback)");
    const int H = (int)spec.widths.size() - 2;
    int wmax = 0;
    for (int l = 1; l <= H; ++l) {
        if (spec.widths[l] < 1) throw std::runtime_error("layer widths must be positive");
        wmax = std::max(wmax, spec.widths[l]);
    }
    if (wmax > 64) throw std::runtime_error("hidden width above 64 is not supported by this kernel family");
    const int WP = (wmax + 7) / 8 * 8;
    if (2 * (H + 1) + (spec.activation == 1 ? 2 * H : 0) > 40) throw std::runtime_error("too many layers");

    // ---- jet set: downward closure of the library derivatives ----
    std::set<std::array<int, 4>> closure;
    closure.insert({0, 0, 0, 0});
    for (const auto& e : spec.encodings) {
        for (int a = spec.d; a < 4; ++a)
            if (e[a] != 0) throw std::runtime_error("derivative along an axis the network does not have");
        for (int a = 0; a < 4; ++a)
            if (e[a] < 0) throw std::runtime_error("negative derivative order");
        for (int t = 0; t <= e[0]; ++t)
            for (int x = 0; x <= e[1]; ++x)
                for (int y = 0; y <= e[2]; ++y)
                    for (int z = 0; z <= e[3]; ++z) closure.insert({t, x, y, z});
    }
    std::vector<std::array<int, 4>> jets(closure.begin(), closure.end());
    auto order = [](const std::array<int, 4>& e) { return e[0] + e[1] + e[2] + e[3]; };
    std::sort(jets.begin(), jets.end(), [&](const std::array<int, 4>& a, const std::array<int, 4>& b) {
        if (order(a) != order(b)) return order(a) < order(b);
        return a > b;   // t before x before y before z inside one order
    });
    const int J = (int)jets.size();
    if (J > 12) throw std::runtime_error("more than 12 jet components (derivative set too rich for this kernel family)");
    std::map<std::array<int, 4>, int> index;
    int kmax = 0;
    for (int c = 0; c < J; ++c) { index[jets[c]] = c; kmax = std::max(kmax, order(jets[c])); }
    if (kmax > 6) throw std::runtime_error("derivative order above 6 is not supported");
    const int NS = kmax + 2;

    const int K = (int)spec.terms.size() - 1;
    if (K < 0) throw std::runtime_error("need at least the LHS term");

    PlanMeta meta;
    meta.d = spec.d; meta.n_jets = J; meta.n_rhs = K; meta.hidden_layers = H; meta.wp = WP; meta.max_order = kmax;
    meta.mode = spec.mode; meta.activation = spec.activation;
    meta.jets = jets;
    meta.n_param_tensors = 2 * (H + 1) + (spec.activation == 1 ? 2 * H : 0);
    meta.n_param_scalars = 0;
    for (int l = 0; l <= H; ++l) meta.n_param_scalars += spec.widths[l + 1] * spec.widths[l] + spec.widths[l + 1];
    if (spec.activation == 1) meta.n_param_scalars += 7 * H;
    meta.sf = find_sf(WP, WP / 4);
    meta.gs = find_gs(WP / 4);
    {   // algorithmic flops per point, with the true widths
        long long f = 0;
        for (int l = 1; l < H; ++l) f += 6ll * spec.widths[l] * spec.widths[l + 1] * J;
        f += 6ll * spec.widths[H] * J + 4ll * spec.d * spec.widths[1];
        meta.flops_per_point = f;
    }
    int nw = spec.warps_per_cta > 0 ? spec.warps_per_cta : 8;
    // engine: the tensor-pipe kernel (forward and, when the extra weight copies fit, data gradient as 3xTF32 mma.sync) is
    // 1.07-1.18x faster than the FP32-pipe kernel for forward+adjoint and 1.5x forward-only (measured, experiments/README.md);
    // it needs a hidden->hidden layer to have anything to contract.  PDL_ENGINE=ffma|mma overrides.
    meta.engine = spec.engine >= 0 ? spec.engine : ((H >= 2) ? 1 : 0);
    if (meta.engine == 1) {
        meta.tile_points = 8;
        // data gradient on the tensor pipe: 1 = B fragments from W^T hi/lo copies (needs four weight copies in shared memory),
        // 2 = B fragments from the forward copies (no extra memory, 2-way bank conflicts); wide nets would spill either way
        if (spec.dgrad_mma >= 0) meta.dgrad_mma = spec.dgrad_mma;
        else {
            meta.dgrad_mma = (WP <= 48) ? 1 : 0;
            if (meta.dgrad_mma == 1 && smem_floats(meta, nw) * 4 > 227 * 1024) meta.dgrad_mma = 2;
        }
        if (meta.dgrad_mma == 1 && smem_floats(meta, nw) * 4 > 227 * 1024) meta.dgrad_mma = 2;
        // weight gradient on the tensor pipe: faster or equal everywhere (Rational nets with the W^T-copy data gradient used to lose
        // to register pressure; with the bf16 cross terms and MUFU.RCP it is a tie: heat 4.16 vs 4.20 ms, burgers 5.63 vs 5.62; Rational
        // J = 6: KS 7.87 -> 7.17 ms), see experiments/README.md
        meta.wgrad_mma = (spec.wgrad_mma >= 0) ? (spec.wgrad_mma ? 1 : 0) : 1;
        // order of the hidden-layer adjoint: the fused order (data gradient, then one activation pass shared by the recomputed
        // forward and the adjoint, then the weight gradient, with a rolled m-tile loop) saves one tanh / Rational recurrence per
        // neuron and layer and is 3-17 % faster everywhere except tanh nets with J <= 4, where the classic order with the unrolled
        // m-tile loop wins by 5 % (heat: 3.04 vs 3.16-3.21 ms; a register-allocation effect, measured in both directions)
        meta.adj_fused = (spec.adj_fused >= 0) ? (spec.adj_fused ? 1 : 0) : ((spec.activation == 0 && J <= 4) ? 0 : 1);
        // cross terms hi*lo + lo*hi of the split products as ONE bf16 m16n8k16 MMA (2 tensor-pipe instructions per product instead of
        // 3); bit mask: 1 = forward, 2 = data gradient, 4 = weight gradient.  Default 6: measured 8-11 % faster with unchanged parity
        // errors; the forward (bit 1) would gain another 3-7 % but its errors reach 1.4e-5 on the derivative features (gate: 1e-5).
        // default 14 = data + weight gradient cross terms as one bf16 MMA (bits 2, 4) and the forward cross terms as one SCALED fp16 MMA
        // (bit 8: 11-bit significands; measured on B200 -2.6 ... -6.8 % kernel time with parity errors equal or smaller than 3xTF32,
        // profiles/r02_ab_fp16_forward.txt).  Bit 1 (forward cross terms in bf16) stays an opt-in the parity gate rejects (1.4e-5).
        meta.cross_bf16 = (spec.cross_bf16 >= 0) ? (spec.cross_bf16 & 15) : 14;
        if (!meta.dgrad_mma) meta.cross_bf16 &= ~2;
        if (!meta.wgrad_mma) meta.cross_bf16 &= ~4;
        // mode 2 reads the TF32 lo plane of W that the bf16 forward replaces: forward bf16 then needs the data gradient bf16 as well
        if (meta.dgrad_mma == 2 && (meta.cross_bf16 & 9)) meta.cross_bf16 |= 2;
        if (meta.dgrad_mma == 2 && (meta.cross_bf16 & 2) && smem_floats(meta, nw) * 4 > 227 * 1024) meta.cross_bf16 &= ~11;
    }
    // launches without gradients on the tcgen05 / TMEM forward engine (jet_mlp_fwd_tc.cuh) when the jet set fits the 512 tensor-memory
    // columns: two accumulator sets (J components at a pitch of NP, or WP with overlapping padding columns) + one or two ring slots,
    // or ONE accumulator set + ring when a thread's share of a layer (WP/8 k-steps x J components x 4 neurons) fits 128 registers
    meta.fwd_tc = 0;
    if (meta.engine == 1 && spec.mode == 0 && H >= 2 && spec.fwd_tc != 0) {
        const int NP = (WP + 15) / 16 * 16, SLOT = 16 * J;
        const bool roomy = 2 * NP * J + 2 * SLOT <= 512;
        const int dcols = (roomy ? NP : WP) * (J - 1) + NP;
        if (2 * dcols + SLOT <= 512) meta.fwd_tc = 1;
        else if (dcols + SLOT <= 512 && (WP / 8) * J * 4 <= 128) meta.fwd_tc = 1;
    }
    while (nw > 1 && smem_floats(meta, nw) * 4 > 227 * 1024) --nw;
    if (smem_floats(meta, nw) * 4 > 227 * 1024) throw std::runtime_error("plan does not fit in shared memory");
    meta.nw = nw;
    meta.smem_bytes = smem_floats(meta, nw) * 4;

    std::ostringstream o;
    o << "// GENERATED by pde_learn_b200/csrc/plan_codegen.cpp -- do not edit.\n";
    o << "// jets:";
    for (int c = 0; c < J; ++c) o << " [" << c << "]=(" << jets[c][0] << "," << jets[c][1] << "," << jets[c][2] << "," << jets[c][3] << ")";
    o << "\n";
    o << "#define PDL_D " << spec.d << "\n#define PDL_J " << J << "\n#define PDL_WP " << WP << "\n#define PDL_H " << H
      << "\n#define PDL_K " << K << "\n#define PDL_ACT " << spec.activation << "\n#define PDL_NS " << NS
      << "\n#define PDL_NW " << nw << "\n#define PDL_MODE " << spec.mode << "\n#define PDL_SF " << meta.sf
      << "\n#define PDL_GS " << meta.gs << "\n#define PDL_DGRAD_MMA " << meta.dgrad_mma
      << "\n#define PDL_WGRAD_MMA " << meta.wgrad_mma << "\n#define PDL_CROSS_BF16 " << meta.cross_bf16
      << "\n#define PDL_ADJ_FUSED " << meta.adj_fused
      // (#ifndef: A/B runs override these two with PDL_NVCC_FLAGS=-D...)
      << "\n#ifndef PDL_WGRAD_ROLL_MT\n#define PDL_WGRAD_ROLL_MT " << ((J >= 5 || meta.adj_fused) ? 1 : 0) << "\n#endif"
      // tanh(z0) cached in a register per neuron for the next adjoint: +1-2 % with the fused order (kdv 4.19 -> 4.09 ms), -4 % for
      // the classic-order J <= 4 kernels (heat 3.03 -> 3.16 ms: spills)
      << "\n#ifndef PDL_TANH_CACHE\n#define PDL_TANH_CACHE " << ((spec.activation == 0 && meta.adj_fused) ? 1 : 0) << "\n#endif\n";
    // n8 tiles per weight-gradient step (independent MMA chains vs registers).  3 since round 1; re-swept in round 2 after the fp16 forward
    // and the L2-reduction flush changed the register picture (profiles/r02_knob_sweep{,2,3,4,5}.txt, A-B-A on one box): all 5 tiles of
    // a 40-wide layer in one step for tanh plans with J <= 5 (heat 2.80 -> 2.70 ms, Burgers / KdV 3.95 -> 3.83) and for Rational J = 5
    // (KdV 4.47 -> 4.42); J = 6 (KS, rd2d) loses 2.5-3.5 % with 5 and Rational J = 4 is indifferent: they keep 3.
    if (meta.engine == 1 && WP == 40 && ((spec.activation == 0 && J <= 5) || J == 5)) o << "#ifndef PDL_WGRAD_NTB\n#define PDL_WGRAD_NTB 5\n#endif\n";
    o << "#ifdef PDL_EMULATE\n#define PDL_GEN static inline\n#else\n#define PDL_GEN __host__ __device__ __forceinline__\n#endif\n";

    // widths / first-order component -> axis
    o << "PDL_GEN int pdl_width(int l) {\n  switch (l) {\n";
    for (size_t l = 0; l < spec.widths.size(); ++l) o << "    case " << l << ": return " << spec.widths[l] << ";\n";
    o << "    default: return 0;\n  }\n}\n";
    o << "PDL_GEN int pdl_comp_axis(int c) {\n  switch (c) {\n";
    for (int c = 1; c < J; ++c)
        if (order(jets[c]) == 1)
            for (int a = 0; a < 4; ++a)
                if (jets[c][a] == 1) o << "    case " << c << ": return " << a << ";\n";
    o << "    default: return -1;\n  }\n}\n";

    // tanh sigmas
    {
        auto Q = tanh_Q(NS - 1);
        o << "PDL_GEN void pdl_tanh_sigmas(const float t, float (&s)[" << NS << "]) {\n";
        o << "  const float t2 = t * t;\n  const float c1 = 1.f - t2;\n  s[0] = t;\n";
        for (int k = 1; k < NS; ++k) {
            const auto& q = Q[k];
            const int par = (k - 1) & 1;
            std::vector<long long> r;                          // coefficients in t2
            for (size_t i = par; i < q.size(); i += 2) r.push_back(q[i]);
            std::string e = fnum(r.back());
            for (int i = (int)r.size() - 2; i >= 0; --i) e = "(" + fnum(r[i]) + " + t2 * " + e + ")";
            o << "  s[" << k << "] = c1 * " << (par ? "t * " : "") << e << ";\n";
        }
        o << "}\n";
    }

    // Faa di Bruno forward and adjoint
    std::vector<std::vector<Partition>> parts(J);
    for (int c = 1; c < J; ++c) parts[c] = partitions_of(jets[c], index);
    o << "PDL_GEN void pdl_act_fwd(const float (&z)[" << J << "], const float (&s)[" << NS << "], float (&h)[" << J << "]) {\n";
    o << "  h[0] = s[0];\n";
    for (int c = 1; c < J; ++c) {
        o << "  h[" << c << "] = ";
        bool firstt = true;
        for (const auto& p : parts[c]) {
            if (!firstt) o << " + ";
            firstt = false;
            if (p.coef != 1) o << fnum(p.coef) << "*";
            o << "s[" << p.blocks.size() << "]*" << prod_expr(p.blocks, "z");
        }
        o << ";\n";
    }
    o << "}\n";
    o << "PDL_GEN void pdl_act_adj(const float (&z)[" << J << "], const float (&hb)[" << J << "], const float (&s)[" << NS
      << "], float (&zb)[" << J << "], float (&sb)[" << NS - 1 << "]) {\n";
    {
        std::vector<std::vector<std::string>> sb_terms(NS - 1), zb_terms(J);
        sb_terms[0].push_back("hb[0]");
        for (int c = 1; c < J; ++c) {
            for (const auto& p : parts[c]) {
                const int m = (int)p.blocks.size();
                std::ostringstream t;
                t << "hb[" << c << "]*";
                if (p.coef != 1) t << fnum(p.coef) << "*";
                t << prod_expr(p.blocks, "z");
                sb_terms[m].push_back(t.str());
                std::set<int> distinct(p.blocks.begin(), p.blocks.end());
                for (int b : distinct) {
                    const long long mu = std::count(p.blocks.begin(), p.blocks.end(), b);
                    std::vector<int> rest = p.blocks;
                    rest.erase(std::find(rest.begin(), rest.end(), b));
                    std::ostringstream u;
                    u << "hb[" << c << "]*";
                    if (p.coef * mu != 1) u << fnum(p.coef * mu) << "*";
                    u << "s[" << m << "]";
                    if (!rest.empty()) u << "*" << prod_expr(rest, "z");
                    zb_terms[b].push_back(u.str());
                }
            }
        }
        for (int m = 0; m < NS - 1; ++m) {
            o << "  sb[" << m << "] = ";
            if (sb_terms[m].empty()) o << "0.f";
            for (size_t i = 0; i < sb_terms[m].size(); ++i) o << (i ? " + " : "") << sb_terms[m][i];
            o << ";\n";
        }
        o << "  zb[0] = 0.f;\n";
        for (int c = 1; c < J; ++c) {
            o << "  zb[" << c << "] = ";
            if (zb_terms[c].empty()) o << "0.f";
            for (size_t i = 0; i < zb_terms[c].size(); ++i) o << (i ? " + " : "") << zb_terms[c][i];
            o << ";\n";
        }
    }
    o << "}\n";

    // library terms, residual, d residual / d feature
    {
        const int KX = std::max(K, 1);
        auto comp_of = [&](int deriv) {
            if (deriv < 0 || deriv >= (int)spec.encodings.size()) throw std::runtime_error("term refers to an unknown derivative");
            return index.at(spec.encodings[deriv]);
        };
        // per term: map comp -> power (merging repeated sub-terms)
        std::vector<std::map<int, int>> tp(spec.terms.size());
        for (size_t t = 0; t < spec.terms.size(); ++t)
            for (const auto& st : spec.terms[t]) {
                if (st.second < 1) throw std::runtime_error("sub-term powers must be >= 1 (Term.py:57-58)");
                tp[t][comp_of(st.first)] += st.second;
            }
        auto term_expr = [&](const std::map<int, int>& pw, int skip_comp) {
            // product of u[c]^p, with one factor of skip_comp removed and multiplied by its power (derivative)
            std::vector<std::string> f;
            for (auto& kv : pw) {
                int p = kv.second;
                if (kv.first == skip_comp) { f.push_back(fnum(p)); --p; }
                for (int q = 0; q < p; ++q) { std::ostringstream u; u << "u[" << kv.first << "]"; f.push_back(u.str()); }
            }
            if (f.empty()) return std::string("1.f");
            std::string e = f[0];
            for (size_t i = 1; i < f.size(); ++i) e += "*" + f[i];
            return e;
        };
        o << "PDL_GEN void pdl_epilogue(const float (&u)[" << J << "], const float* xi, float& r, float (&du)[" << J
          << "], float (&T)[" << KX << "]) {\n";
        o << "  const float T0 = " << term_expr(tp[0], -1) << ";\n";
        if (K == 0) o << "  T[0] = 0.f; (void)xi;\n";
        for (int k = 0; k < K; ++k) o << "  T[" << k << "] = " << term_expr(tp[k + 1], -1) << ";\n";
        o << "  float lxi = 0.f;\n";
        for (int k = 0; k < K; ++k) o << "  lxi += xi[" << k << "]*T[" << k << "];\n";
        o << "  r = T0 - lxi;\n";
        for (int c = 0; c < J; ++c) {
            o << "  du[" << c << "] = ";
            bool any = false;
            if (tp[0].count(c)) { o << term_expr(tp[0], c); any = true; }
            else o << "0.f";
            (void)any;
            for (int k = 0; k < K; ++k)
                if (tp[k + 1].count(c)) o << " - xi[" << k << "]*(" << term_expr(tp[k + 1], c) << ")";
            o << ";\n";
        }
        o << "}\n";
    }
    if (meta.fwd_tc) {
        // the tcgen05 forward engine wraps the module's entry points: it takes the launches without gradients and forwards the rest
        // (the mma.sync template itself stays untouched, so plans without this engine keep their cached modules)
        o << "#define PDL_FWD_TC 1\n#ifndef PDL_EMULATE\n#define pdl_k_launch pdl_k_launch_mma\n#define pdl_k_info pdl_k_info_mma\n#endif\n"
          << "#include \"jet_mlp_kernel_mma.cuh\"\n#ifndef PDL_EMULATE\n#undef pdl_k_launch\n#undef pdl_k_info\n#endif\n"
          << "#include \"jet_mlp_fwd_tc.cuh\"\n";
    } else {
        o << (meta.engine == 1 ? "#include \"jet_mlp_kernel_mma.cuh\"\n" : "#include \"jet_mlp_kernel.cuh\"\n");
    }
    if (meta_out) *meta_out = meta;
    return o.str();
}

}  // namespace pdl_host
