// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED.
//
// CPU restatement of winter-prover 0.12.0 `Prover::prove` / `generate_proof` with the Default*
// associated types the reference selects (src/training/prover.rs:225-233, :273-300;
// src/aggregation/prover.rs:198-206, :216-248), driven from src/main.rs:228,424,468.
// Stage numbering follows SURVEY §3.2; upstream module per stage is named inline.
#include "proof.h"

namespace orc {

int g_threads = 1;
// test hook: when set, prove() copies the constraint evaluations over the ce domain (CompositionPolyTrace, before interpolation) here
Fe* g_dump_comp_trace = nullptr;

// RowMatrix::evaluate_polys_over: LDE of every column polynomial over 3*<w_N>, row-major [N][w]
static std::vector<Fe> evaluate_polys_over(const std::vector<std::vector<Fe>>& polys, size_t n, size_t blowup) {
    size_t w = polys.size(), N = n * blowup;
    std::vector<Fe> lde(N * w);
    Fe* L = lde.data();
    parallel_for(w, [&](size_t a, size_t b) {
        for (size_t j = a; j < b; j++) {
            std::vector<Fe> col = evaluate_poly_with_offset(polys[j].data(), n, fe_raw(GENERATOR), blowup);
            for (size_t r = 0; r < N; r++) L[r * w + j] = col[r];
        }
    });
    return lde;
}
// RowMatrix::commit_to_rows (single partition): leaf = hash_elements(row); MerkleTree::new
static MerkleTree commit_to_rows(const std::vector<Fe>& m, size_t rows, size_t w) {
    std::vector<Digest> leaves(rows);
    const Fe* M = m.data();
    Digest* Lv = leaves.data();
    parallel_for(rows, [&](size_t a, size_t b) { for (size_t r = a; r < b; r++) Lv[r] = hash_elements(M + r * w, w); });
    return merkle_new(std::move(leaves));
}

static std::vector<uint8_t> rows_to_bytes(const std::vector<Fe>& m, size_t w, const std::vector<size_t>& pos) {
    std::vector<uint8_t> out(pos.size() * w * 16);
    for (size_t i = 0; i < pos.size(); i++) memcpy(&out[i * w * 16], &m[pos[i] * w], w * 16);
    return out;
}

// winter-fri folding::fold_positions
std::vector<size_t> fold_positions(const std::vector<size_t>& positions, size_t source_domain, size_t folding) {
    size_t target = source_domain / folding;
    std::vector<size_t> r;
    for (size_t p : positions) { size_t q = p % target; if (std::find(r.begin(), r.end(), q) == r.end()) r.push_back(q); }
    return r;
}

// winter-fri folding::apply_drp for one row: interpolate through (x*w_F^j, row[j]) and evaluate at alpha
Fe fold_row(const Fe* row, size_t F, Fe x, Fe alpha) {
    std::vector<Fe> c(row, row + F);
    interpolate_poly(c.data(), F);  // coefficients in y where point_j = w_F^j
    Fe t = mul(alpha, inv(x)), f = FE_ONE, r = FE_ZERO;  // p(alpha) = sum c_k (alpha/x)^k
    for (size_t k = 0; k < F; k++) { r = add(r, mul(c[k], f)); f = mul(f, t); }
    return r;
}

Proof prove(const Air& air, const Fe* trace_colmajor, Transcript* ts, uint64_t force_nonce) {
    air.validate();
    const size_t n = air.n, w = air.w, beta = air.opt.blowup, N = n * beta;
    const size_t ce = air.ce_blowup(), ce_n = n * ce, c = air.num_composition_columns();
    const Fe offset = fe_raw(GENERATOR);
    const Fe g = get_root_of_unity(ilog2(n));
    Transcript local; if (!ts) ts = &local;
    Proof proof;
    proof.trace_width = (uint32_t)w; proof.trace_len = (uint32_t)n; proof.opt = air.opt;

    // 0 ---- ProverChannel::new: coin seeded with Context ++ public inputs ----------------------------
    Coin coin = Coin::create(air.coin_seed_elements());

    // 1 ---- DefaultTraceLde::new: interpolate, extend, hash rows, Merkle --------------------------
    std::vector<std::vector<Fe>> polys(w);
    parallel_for(w, [&](size_t a, size_t b) {
        for (size_t j = a; j < b; j++) {
            polys[j].assign(trace_colmajor + j * n, trace_colmajor + (j + 1) * n);
            interpolate_poly(polys[j].data(), n);  // ColMatrix::interpolate_columns
        }
    });
    std::vector<Fe> lde = evaluate_polys_over(polys, n, beta);
    MerkleTree trace_tree = commit_to_rows(lde, N, w);
    ts->trace_root = trace_tree.root();
    proof.commitments.push_back(trace_tree.root());
    coin.reseed(trace_tree.root());  // channel.commit_trace

    // 2 ---- DefaultConstraintEvaluator::evaluate -----------------------------------------------------
    Fe alpha = coin.draw();  // ConstraintCompositionCoefficients::draw_algebraic
    ts->constraint_alpha = alpha;
    size_t nt = air.num_transition(), na = air.assertions.size();
    std::vector<Fe> tc(nt), bc(na);
    { Fe cur = FE_ONE; for (auto& x : tc) { x = cur; cur = mul(cur, alpha); } for (auto& x : bc) { x = cur; cur = mul(cur, alpha); } }
    std::vector<BoundaryGroup> groups = boundary_groups(air, bc);
    // periodic value table over the ce domain (PeriodicValueTable)
    std::vector<Fe> ptable; size_t pl = 0;
    if (air.id == AIR_MIMC) {
        std::vector<Fe> pp = periodic_poly(air);
        pl = pp.size();
        ptable = evaluate_poly_with_offset(pp.data(), pl, pow(offset, n / pl), ce);
    }
    std::vector<Fe> comp(ce_n);
    {
        const Fe gce = get_root_of_unity(ilog2(ce_n));
        const size_t lde_step = beta / ce;
        // 1/(x^n - 1) takes only `ce` distinct values over the ce domain
        std::vector<Fe> zinv(ce);
        { Fe on = pow(offset, n), wce = get_root_of_unity(ilog2(ce)), f = FE_ONE;
          for (size_t i = 0; i < ce; i++) { zinv[i] = sub(mul(on, f), FE_ONE); f = mul(f, wce); }
          batch_inv(zinv.data(), ce); }
        const Fe g_last = pow(g, n - 1);
        Fe* C = comp.data();
        parallel_for(ce_n, [&](size_t a, size_t b) {
            std::vector<Fe> tev(nt), den((b - a) * groups.size());
            std::vector<Fe> bnum((b - a) * groups.size());
            Fe x = mul(offset, pow(gce, a));
            for (size_t i = a; i < b; i++) {
                size_t r = i * lde_step;
                const Fe* cur = &lde[r * w];
                const Fe* nxt = &lde[((r + beta) % N) * w];
                Fe per = pl ? ptable[i % (pl * ce)] : FE_ZERO;
                air.evaluate_transition(cur, nxt, &per, tev.data());
                Fe t = FE_ZERO;
                for (size_t k = 0; k < nt; k++) t = add(t, mul(tc[k], tev[k]));
                // transition divisor (x^n - 1)/(x - g^(n-1))
                C[i] = mul(mul(t, sub(x, g_last)), zinv[i % ce]);
                for (size_t q = 0; q < groups.size(); q++) {
                    const BoundaryGroup& bg = groups[q];
                    Fe s = FE_ZERO;
                    for (size_t k = 0; k < bg.cols.size(); k++) s = add(s, mul(bg.coeffs[k], sub(cur[bg.cols[k]], bg.values[k])));
                    bnum[(i - a) * groups.size() + q] = s;
                    den[(i - a) * groups.size() + q] = sub(x, bg.g_step);
                }
                x = mul(x, gce);
            }
            batch_inv(den.data(), den.size());
            for (size_t i = a; i < b; i++)
                for (size_t q = 0; q < groups.size(); q++)
                    C[i] = add(C[i], mul(bnum[(i - a) * groups.size() + q], den[(i - a) * groups.size() + q]));
        });
    }

    if (g_dump_comp_trace) memcpy(g_dump_comp_trace, comp.data(), ce_n * sizeof(Fe));
    // 3 ---- DefaultConstraintCommitment::new / CompositionPoly::new --------------------------------------
    interpolate_poly_with_offset(comp.data(), ce_n, offset);
    for (size_t i = c * n; i < ce_n; i++) if (comp[i].v != 0) { ts->comp_degree_ok = 0; break; }
    std::vector<std::vector<Fe>> hcols(c);
    for (size_t i = 0; i < c; i++) hcols[i].assign(comp.begin() + i * n, comp.begin() + (i + 1) * n);
    std::vector<Fe> comp_lde = evaluate_polys_over(hcols, n, beta);
    MerkleTree comp_tree = commit_to_rows(comp_lde, N, c);
    ts->constraint_root = comp_tree.root();
    proof.commitments.push_back(comp_tree.root());
    coin.reseed(comp_tree.root());  // channel.commit_constraints

    // 4 ---- out-of-domain frame + DEEP composition polynomial (composer/mod.rs) ---------------------------
    Fe z = coin.draw();
    ts->z = z;
    Fe zg = mul(z, g);
    std::vector<Fe> ood_cur(w), ood_next(w), ood_h(c);
    parallel_for(w, [&](size_t a, size_t b) {
        for (size_t j = a; j < b; j++) { ood_cur[j] = poly_eval(polys[j].data(), n, z); ood_next[j] = poly_eval(polys[j].data(), n, zg); }
    });
    proof.ood_trace_states.resize(2 * w);
    for (size_t j = 0; j < w; j++) { proof.ood_trace_states[2 * j] = ood_cur[j]; proof.ood_trace_states[2 * j + 1] = ood_next[j]; }
    coin.reseed(hash_elements(proof.ood_trace_states.data(), 2 * w));  // send_ood_trace_states
    for (size_t i = 0; i < c; i++) ood_h[i] = poly_eval(hcols[i].data(), n, z);
    proof.ood_constraint_evals = ood_h;
    coin.reseed(hash_elements(ood_h.data(), c));  // send_ood_constraint_evaluations

    Fe dalpha = coin.draw();  // DeepCompositionCoefficients::draw_algebraic
    ts->deep_alpha = dalpha;
    std::vector<Fe> cct(w), ccc(c);
    { Fe cur = FE_ONE; for (auto& x : cct) { x = cur; cur = mul(cur, dalpha); } for (auto& x : ccc) { x = cur; cur = mul(cur, dalpha); } }
    std::vector<Fe> deep(n);
    {
        std::vector<Fe> t1(n, FE_ZERO), t2(n, FE_ZERO);
        for (size_t j = 0; j < w; j++) {  // add_trace_polys / acc_trace_poly
            for (size_t m = 0; m < n; m++) { Fe v = mul(polys[j][m], cct[j]); t1[m] = add(t1[m], v); t2[m] = add(t2[m], v); }
            t1[0] = sub(t1[0], mul(ood_cur[j], cct[j]));
            t2[0] = sub(t2[0], mul(ood_next[j], cct[j]));
        }
        syn_div_in_place(t1.data(), n, z);
        syn_div_in_place(t2.data(), n, zg);
        for (size_t m = 0; m < n; m++) deep[m] = add(t1[m], t2[m]);
        for (size_t i = 0; i < c; i++) {  // add_composition_poly
            std::vector<Fe> h = hcols[i];
            h[0] = sub(h[0], ood_h[i]);
            syn_div_in_place(h.data(), n, z);
            for (size_t m = 0; m < n; m++) deep[m] = add(deep[m], mul(h[m], ccc[i]));
        }
    }
    // 5 ---- DeepCompositionPoly::evaluate ------------------------------------------------------------
    std::vector<Fe> evals = evaluate_poly_with_offset(deep.data(), n, offset, beta);

    // 6 ---- FriProver::build_layers (winter-fri prover/mod.rs) ------------------------------------------
    const size_t F = air.opt.folding;
    size_t nlayers = air.opt.num_fri_layers(N);
    std::vector<std::vector<Fe>> layer_rows;  // transposed evaluations, row i = [e[i + j*M/F]]
    std::vector<MerkleTree> layer_trees;
    for (size_t l = 0; l < nlayers; l++) {
        size_t M = evals.size(), rows = M / F;
        std::vector<Fe> tr(M);
        for (size_t i = 0; i < rows; i++) for (size_t j = 0; j < F; j++) tr[i * F + j] = evals[i + j * rows];
        MerkleTree t = commit_to_rows(tr, rows, F);
        proof.commitments.push_back(t.root());
        ts->fri_roots.push_back(t.root());
        coin.reseed(t.root());  // commit_fri_layer
        Fe a = coin.draw();     // draw_fri_alpha
        ts->fri_alphas.push_back(a);
        std::vector<Fe> next(rows);
        Fe gm = get_root_of_unity(ilog2(M));
        parallel_for(rows, [&](size_t lo, size_t hi) {
            Fe x = mul(offset, pow(gm, lo));  // the offset is NOT raised to F between layers
            for (size_t i = lo; i < hi; i++) { next[i] = fold_row(&tr[i * F], F, x, a); x = mul(x, gm); }
        });
        layer_rows.push_back(std::move(tr));
        layer_trees.push_back(std::move(t));
        evals.swap(next);
    }
    {   // set_remainder: coefficients of the last layer, first M/blowup of them, reversed
        size_t M = evals.size();
        interpolate_poly_with_offset(evals.data(), M, offset);
        size_t rs = M / beta;
        proof.fri_remainder.assign(evals.begin(), evals.begin() + rs);
        std::reverse(proof.fri_remainder.begin(), proof.fri_remainder.end());
        Digest rc = hash_elements(proof.fri_remainder.data(), rs);
        ts->remainder_commitment = rc;
        proof.commitments.push_back(rc);
        coin.reseed(rc);
    }

    // 7 ---- grind_query_seed + get_query_positions (channel.rs) -------------------------------------------
    uint64_t nonce = force_nonce;
    if (!nonce) {  // non-`concurrent` semantics: the smallest valid nonce >= 1
        std::vector<uint64_t> found(g_threads > 0 ? g_threads : 1, 0);
        uint64_t base = 1;
        const uint64_t window = 1 << 16;
        while (!nonce) {
            int T = (int)found.size();
            parallel_for(T, [&](size_t a, size_t b) {
                for (size_t t = a; t < b; t++) {
                    found[t] = 0;
                    uint64_t lo = base + t * window;
                    for (uint64_t v = lo; v < lo + window; v++)
                        if (coin.check_leading_zeros(v) >= air.opt.grinding) { found[t] = v; break; }
                }
            });
            for (int t = 0; t < T && !nonce; t++) if (found[t]) nonce = found[t];
            base += (uint64_t)T * window;
        }
    }
    proof.pow_nonce = nonce; ts->pow_nonce = nonce;
    std::vector<size_t> positions = coin.draw_integers(air.opt.num_queries, N, nonce);
    std::sort(positions.begin(), positions.end());
    positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
    ts->positions = positions;
    proof.num_unique_queries = (uint8_t)positions.size();

    // 8 ---- FriProver::build_proof, TraceLde::query, ConstraintCommitment::query ------------------------
    {
        std::vector<size_t> pos = positions;
        size_t dom = N;
        for (size_t l = 0; l < nlayers; l++) {
            pos = fold_positions(pos, dom, F);
            proof.fri_layer_paths.push_back(batch_proof_to_bytes(merkle_prove_batch(layer_trees[l], pos)));
            proof.fri_layer_values.push_back(rows_to_bytes(layer_rows[l], F, pos));
            dom /= F;
        }
    }
    proof.trace_query_values = rows_to_bytes(lde, w, positions);
    proof.trace_query_proof = batch_proof_to_bytes(merkle_prove_batch(trace_tree, positions));
    proof.constraint_query_values = rows_to_bytes(comp_lde, c, positions);
    proof.constraint_query_proof = batch_proof_to_bytes(merkle_prove_batch(comp_tree, positions));
    return proof;
}

}  // namespace orc
