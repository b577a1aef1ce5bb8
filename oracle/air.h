// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED.
//
// The three AIRs on the hot path, restated from the reference:
//   TRAINING     src/training/air.rs:105-151 (context, assertions), :154-287 with src/helper.rs:141-146
//                (current_step() == 0  =>  every transition evaluation is zero)
//   AGGREGATION  src/aggregation/air.rs:93-99 (context), :101-119 (transition), :121-147 (assertions)
//   MIMC         defined by this build: W independent chains
//                next_j = (cur_j + rc[i mod 64])^7 with rc = src/helper.rs:404-406 as a periodic column,
//                round function from src/helper.rs:213-220; assertions col_j[0], col_j[n-1].
// plus the winter-air 0.12 AirContext arithmetic they imply.
#pragma once
#include "stark.h"

namespace orc {

enum AirId { AIR_TRAINING = 1, AIR_AGGREGATION = 2, AIR_MIMC = 3 };

struct Options {  // winter-air ProofOptions::new argument order, src/main.rs:98-107
    uint32_t num_queries = 40, blowup = 16, grinding = 21, field_extension = 1 /* None */,
             folding = 16, rem_max_degree = 7, batching_constraints = 1 /* Algebraic */, batching_deep = 1;
    size_t num_fri_layers(size_t domain) const {
        size_t r = 0, max_rem = (size_t)(rem_max_degree + 1) * blowup;
        while (domain > max_rem) { domain /= folding; r++; }
        return r;
    }
};

struct Assertion { uint32_t col; uint64_t step; Fe value; };

struct Air {
    uint32_t id = 0;
    size_t n = 0, w = 0;  // trace length / width
    Options opt;
    std::vector<Fe> pub_elems;          // PublicInputs::to_elements()
    std::vector<Assertion> assertions;  // Air::get_assertions(), any order
    std::vector<Fe> params;             // aggregation: [k]; mimc: round constants (periodic column)

    // --- AirContext ---
    size_t num_transition() const { return id == AIR_AGGREGATION ? w / 2 : w; }
    size_t degree_base() const { return id == AIR_MIMC ? 7 : 1; }
    // MiMC declares TransitionConstraintDegree::new(7): (cur + rc)^7 has degree exactly 7(n-1) because
    // the periodic polynomial's degree (n/L)(L-1) is below n-1, so no cycle term is declared.
    size_t num_cycles() const { return 0; }
    size_t cycle_len() const { return 0; }
    size_t ce_blowup() const {  // TransitionConstraintDegree::min_blowup_factor
        size_t d = degree_base() + num_cycles(), b = 1;
        while (b < d) b <<= 1;
        return b < 2 ? 2 : b;
    }
    size_t eval_degree() const {  // get_evaluation_degree(trace_len)
        size_t r = degree_base() * (n - 1);
        if (num_cycles()) r += (n / cycle_len()) * (cycle_len() - 1);
        return r;
    }
    size_t num_composition_columns() const {
        size_t div_degree = n - 1;  // trace_len - num_transition_exemptions(=1)
        size_t c = (eval_degree() - div_degree + n - 1) / n;
        return c < 1 ? 1 : c;
    }
    size_t num_constraints() const { return assertions.size() + num_transition(); }
    size_t lde_size() const { return n * opt.blowup; }
    size_t ce_size() const { return n * ce_blowup(); }

    // Air::evaluate_transition over the base field
    void evaluate_transition(const Fe* cur, const Fe* next, const Fe* periodic, Fe* result) const {
        if (id == AIR_TRAINING) {
            for (size_t i = 0; i < w; i++) result[i] = FE_ZERO;  // src/training/air.rs:274-278
        } else if (id == AIR_AGGREGATION) {
            size_t d = w / 2;
            Fe k = params[0];
            for (size_t i = 0; i < d; i++)  // src/aggregation/air.rs:110-115
                result[i] = sub(sub(mul(k, next[i]), mul(k, cur[i])), next[i + d]);
        } else {
            for (size_t i = 0; i < w; i++) {
                Fe a = add(cur[i], periodic[0]);
                Fe a2 = mul(a, a), a4 = mul(a2, a2), a6 = mul(a4, a2), a7 = mul(a6, a);
                result[i] = sub(next[i], a7);
            }
        }
    }

    // assertions sorted like winter-air prepare_assertions: (stride=0, first_step, column)
    std::vector<Assertion> sorted_assertions() const {
        std::vector<Assertion> a = assertions;
        std::stable_sort(a.begin(), a.end(), [](const Assertion& x, const Assertion& y) {
            if (x.step != y.step) return x.step < y.step;
            return x.col < y.col;
        });
        for (size_t i = 0; i + 1 < a.size(); i++)
            if (a[i].step == a[i + 1].step && a[i].col == a[i + 1].col) throw std::runtime_error("air: overlapping assertions");
        for (auto& x : a) if (x.col >= w || x.step >= n) throw std::runtime_error("air: assertion out of range");
        return a;
    }

    void validate() const {
        if (id < 1 || id > 3) throw std::runtime_error("air: unknown air id");
        if (n < 8 || (n & (n - 1))) throw std::runtime_error("air: trace length must be a power of two >= 8");
        if (w < 1 || w > 255) throw std::runtime_error("air: trace width must be in 1..=255");
        if (assertions.empty()) throw std::runtime_error("air: at least one assertion is required");
        if (opt.blowup < ce_blowup()) throw std::runtime_error("air: blowup factor too small for constraint degree");
        if (opt.blowup < 2 || opt.blowup > 128 || (opt.blowup & (opt.blowup - 1))) throw std::runtime_error("options: bad blowup");
        if (opt.folding != 2 && opt.folding != 4 && opt.folding != 8 && opt.folding != 16) throw std::runtime_error("options: bad folding factor");
        if (opt.num_queries < 1 || opt.num_queries > 255) throw std::runtime_error("options: bad query count");
        if (opt.grinding > 32) throw std::runtime_error("options: bad grinding factor");
        if (((opt.rem_max_degree + 1) & opt.rem_max_degree) != 0 || opt.rem_max_degree > 255) throw std::runtime_error("options: bad remainder degree");
        if (opt.field_extension != 1) throw std::runtime_error("options: only FieldExtension::None is supported");
        if (opt.batching_constraints != 1 || opt.batching_deep != 1) throw std::runtime_error("options: only algebraic batching is supported");
        if (id == AIR_AGGREGATION && (w % 2 || params.size() != 1)) throw std::runtime_error("air: aggregation needs even width and k");
        if (id == AIR_MIMC) {
            size_t L = params.size();
            if (L < 2 || (L & (L - 1)) || L > n) throw std::runtime_error("air: mimc needs a power-of-two round-constant cycle");
        }
        if (lde_size() > ((size_t)1 << 32)) throw std::runtime_error("air: lde domain too large");
    }

    // Context::to_elements() ++ pub_inputs.to_elements()  
    std::vector<Fe> coin_seed_elements() const {
        std::vector<Fe> e;
        e.push_back(fe_raw(((u128)w << 8) | 0));  // main width, 0 aux segments
        e.push_back(fe_raw((u128)n));             // trace length
        e.push_back(fe_raw((u128)(u64)P));        // modulus low 8 bytes
        e.push_back(fe_raw((u128)(u64)(P >> 64)));  // modulus high 8 bytes
        e.push_back(fe_raw((u128)num_constraints()));
        uint32_t buf = opt.field_extension;
        buf = (buf << 8) | opt.folding;
        buf = (buf << 8) | opt.rem_max_degree;
        buf = (buf << 8) | opt.blowup;
        e.push_back(fe_raw(buf));
        e.push_back(fe_raw(opt.grinding));
        e.push_back(fe_raw(opt.num_queries));
        e.insert(e.end(), pub_elems.begin(), pub_elems.end());
        return e;
    }
};

// (x^n - 1)/(x - g^(n-1)) and (x - g^step): ConstraintDivisor of winter-air
struct BoundaryGroup { uint64_t step; Fe g_step; std::vector<uint32_t> cols; std::vector<Fe> values, coeffs; };

static inline std::vector<BoundaryGroup> boundary_groups(const Air& air, const std::vector<Fe>& b_coeffs) {
    std::vector<Assertion> a = air.sorted_assertions();
    Fe g = get_root_of_unity(ilog2(air.n));
    std::vector<BoundaryGroup> groups;
    for (size_t i = 0; i < a.size(); i++) {
        if (groups.empty() || groups.back().step != a[i].step) {
            BoundaryGroup bg; bg.step = a[i].step; bg.g_step = pow(g, a[i].step);
            groups.push_back(bg);
        }
        groups.back().cols.push_back(a[i].col);
        groups.back().values.push_back(a[i].value);
        groups.back().coeffs.push_back(b_coeffs[i]);
    }
    return groups;
}

// periodic column polynomial: interpolate the cycle values over <w_L>
static inline std::vector<Fe> periodic_poly(const Air& air) {
    std::vector<Fe> p = air.params;
    interpolate_poly(p.data(), p.size());
    return p;
}

}  // namespace orc
