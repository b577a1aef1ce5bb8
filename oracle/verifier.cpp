// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED.
//
// CPU restatement of winter-verifier 0.12.0 `verify` / `perform_verification`, the only behavioural
// acceptance check the reference has (src/main.rs:251-257, 430-436, 478-484;
// tests/integration_tests.rs:102-111).  Closes prove -> verify inside this container: a proof from
// either the oracle prover or the CUDA prover must be accepted here.
#include "proof.h"

namespace orc {

std::vector<size_t> fold_positions(const std::vector<size_t>& positions, size_t source_domain, size_t folding);
Fe fold_row(const Fe* row, size_t F, Fe x, Fe alpha);

static void fail(const char* m) { throw std::runtime_error(std::string("verify: ") + m); }

void verify(const Air& air, const uint8_t* bytes, size_t len, Transcript* ts) {
    air.validate();
    Transcript local; if (!ts) ts = &local;
    Proof p = proof_from_bytes(bytes, len);
    const size_t n = air.n, w = air.w, beta = air.opt.blowup, N = n * beta, c = air.num_composition_columns();
    const size_t F = air.opt.folding;
    // proof context must match the AIR / AcceptableOptions::OptionSet
    if (p.trace_width != w || p.trace_len != n) fail("trace info mismatch");
    if (p.opt.num_queries != air.opt.num_queries || p.opt.blowup != air.opt.blowup || p.opt.grinding != air.opt.grinding ||
        p.opt.folding != air.opt.folding || p.opt.rem_max_degree != air.opt.rem_max_degree ||
        p.opt.field_extension != air.opt.field_extension || p.opt.batching_constraints != air.opt.batching_constraints ||
        p.opt.batching_deep != air.opt.batching_deep)
        fail("unacceptable proof options");
    const size_t nlayers = air.opt.num_fri_layers(N);
    if (p.commitments.size() != 2 + nlayers + 1) fail("wrong number of commitments");
    if (p.fri_layer_values.size() != nlayers) fail("wrong number of FRI layers");
    if (p.ood_trace_states.size() != 2 * w || p.ood_constraint_evals.size() != c) fail("bad OOD frame");

    const Fe offset = fe_raw(GENERATOR);
    const Fe g = get_root_of_unity(ilog2(n));
    Coin coin = Coin::create(air.coin_seed_elements());

    // 1 ---- trace commitment, constraint coefficients ------------------------------------------------
    coin.reseed(p.commitments[0]);
    ts->trace_root = p.commitments[0];
    Fe alpha = coin.draw();
    ts->constraint_alpha = alpha;
    size_t nt = air.num_transition(), na = air.assertions.size();
    std::vector<Fe> tc(nt), bc(na);
    { Fe cur = FE_ONE; for (auto& x : tc) { x = cur; cur = mul(cur, alpha); } for (auto& x : bc) { x = cur; cur = mul(cur, alpha); } }

    // 2 ---- constraint commitment, OOD point -----------------------------------------------------------
    coin.reseed(p.commitments[1]);
    ts->constraint_root = p.commitments[1];
    Fe z = coin.draw();
    ts->z = z;

    // 3 ---- OOD consistency check (verifier evaluator.rs::evaluate_constraints) --------------------------
    std::vector<Fe> ood_cur(w), ood_next(w);
    for (size_t j = 0; j < w; j++) { ood_cur[j] = p.ood_trace_states[2 * j]; ood_next[j] = p.ood_trace_states[2 * j + 1]; }
    Fe eval1;
    {
        Fe per = FE_ZERO;
        if (air.id == AIR_MIMC) { std::vector<Fe> pp = periodic_poly(air); per = poly_eval(pp.data(), pp.size(), pow(z, n / pp.size())); }
        std::vector<Fe> tev(nt);
        air.evaluate_transition(ood_cur.data(), ood_next.data(), &per, tev.data());
        Fe t = FE_ZERO;
        for (size_t k = 0; k < nt; k++) t = add(t, mul(tc[k], tev[k]));
        Fe zn = pow(z, n);
        eval1 = mul(mul(t, sub(z, pow(g, n - 1))), inv(sub(zn, FE_ONE)));
        for (auto& bg : boundary_groups(air, bc)) {
            Fe s = FE_ZERO;
            for (size_t k = 0; k < bg.cols.size(); k++) s = add(s, mul(bg.coeffs[k], sub(ood_cur[bg.cols[k]], bg.values[k])));
            eval1 = add(eval1, mul(s, inv(sub(z, bg.g_step))));
        }
    }
    coin.reseed(hash_elements(p.ood_trace_states.data(), 2 * w));
    Fe eval2 = FE_ZERO;
    for (size_t i = 0; i < c; i++) eval2 = add(eval2, mul(pow(z, (u128)i * n), p.ood_constraint_evals[i]));
    coin.reseed(hash_elements(p.ood_constraint_evals.data(), c));
    if (eval1 != eval2) fail("inconsistent OOD constraint evaluations");

    // 4 ---- DEEP coefficients, FRI commitments (FriVerifier::new) ----------------------------------------
    Fe dalpha = coin.draw();
    ts->deep_alpha = dalpha;
    std::vector<Fe> cct(w), ccc(c);
    { Fe cur = FE_ONE; for (auto& x : cct) { x = cur; cur = mul(cur, dalpha); } for (auto& x : ccc) { x = cur; cur = mul(cur, dalpha); } }
    std::vector<Fe> layer_alphas;
    for (size_t l = 0; l <= nlayers; l++) {
        coin.reseed(p.commitments[2 + l]);
        Fe a = coin.draw();
        if (l < nlayers) { layer_alphas.push_back(a); ts->fri_roots.push_back(p.commitments[2 + l]); ts->fri_alphas.push_back(a); }
    }
    ts->remainder_commitment = p.commitments[2 + nlayers];

    // 5 ---- proof of work, query positions -----------------------------------------------------------
    if (coin.check_leading_zeros(p.pow_nonce) < air.opt.grinding) fail("query seed proof-of-work verification failed");
    ts->pow_nonce = p.pow_nonce;
    std::vector<size_t> positions = coin.draw_integers(air.opt.num_queries, N, p.pow_nonce);
    std::sort(positions.begin(), positions.end());
    positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
    ts->positions = positions;
    if (p.num_unique_queries != positions.size()) fail("number of unique queries mismatch");
    const size_t nq = positions.size();

    // queried trace / constraint rows against their commitments
    if (p.trace_query_values.size() != nq * w * 16) fail("bad trace query values");
    if (p.constraint_query_values.size() != nq * c * 16) fail("bad constraint query values");
    std::vector<Fe> trows(nq * w), crows(nq * c);
    { Reader r(p.trace_query_values.data(), p.trace_query_values.size()); for (auto& e : trows) e = r.fe(); }
    { Reader r(p.constraint_query_values.data(), p.constraint_query_values.size()); for (auto& e : crows) e = r.fe(); }
    {
        std::vector<Digest> leaves(nq);
        for (size_t i = 0; i < nq; i++) leaves[i] = hash_elements(&trows[i * w], w);
        Digest root;
        if (!merkle_batch_root(batch_proof_from_bytes(p.trace_query_proof), positions, leaves, &root) || root != p.commitments[0])
            fail("trace query did not match the commitment");
        for (size_t i = 0; i < nq; i++) leaves[i] = hash_elements(&crows[i * c], c);
        if (!merkle_batch_root(batch_proof_from_bytes(p.constraint_query_proof), positions, leaves, &root) || root != p.commitments[1])
            fail("constraint query did not match the commitment");
    }

    // 6 ---- DEEP composition at the queried positions (verifier composer.rs) -----------------------------
    std::vector<Fe> deep(nq);
    {
        Fe gN = get_root_of_unity(ilog2(N)), zg = mul(z, g);
        for (size_t i = 0; i < nq; i++) {
            Fe x = mul(offset, pow(gN, positions[i]));
            Fe t1 = FE_ZERO, t2 = FE_ZERO, h = FE_ZERO;
            for (size_t j = 0; j < w; j++) {
                t1 = add(t1, mul(sub(trows[i * w + j], ood_cur[j]), cct[j]));
                t2 = add(t2, mul(sub(trows[i * w + j], ood_next[j]), cct[j]));
            }
            for (size_t k = 0; k < c; k++) h = add(h, mul(sub(crows[i * c + k], p.ood_constraint_evals[k]), ccc[k]));
            Fe d1 = inv(sub(x, z)), d2 = inv(sub(x, zg));
            deep[i] = add(add(mul(t1, d1), mul(t2, d2)), mul(h, d1));
        }
    }

    // 7 ---- FriVerifier::verify -----------------------------------------------------------------------------
    {
        std::vector<size_t> pos = positions;
        std::vector<Fe> evaluations = deep;
        size_t dom = N, max_deg_plus_1 = n;  // trace_poly_degree + 1
        Fe gen = get_root_of_unity(ilog2(N));
        for (size_t l = 0; l < nlayers; l++) {
            std::vector<size_t> folded = fold_positions(pos, dom, F);
            size_t rows = dom / F;
            if (p.fri_layer_values[l].size() != folded.size() * F * 16) fail("bad FRI layer values");
            std::vector<Fe> vals(folded.size() * F);
            { Reader r(p.fri_layer_values[l].data(), p.fri_layer_values[l].size()); for (auto& e : vals) e = r.fe(); }
            std::vector<Digest> leaves(folded.size());
            for (size_t i = 0; i < folded.size(); i++) leaves[i] = hash_elements(&vals[i * F], F);
            Digest root;
            if (!merkle_batch_root(batch_proof_from_bytes(p.fri_layer_paths[l]), folded, leaves, &root) || root != p.commitments[2 + l])
                fail("FRI layer query did not match the commitment");
            // get_query_values: evaluation at position q sits in row (q % rows), slot (q / rows)
            for (size_t i = 0; i < pos.size(); i++) {
                size_t q = pos[i], fi = std::find(folded.begin(), folded.end(), q % rows) - folded.begin();
                if (vals[fi * F + q / rows] != evaluations[i]) fail("invalid FRI layer folding");
            }
            std::vector<Fe> next(folded.size());
            for (size_t i = 0; i < folded.size(); i++) {
                Fe xe = mul(pow(gen, folded[i]), offset);
                next[i] = fold_row(&vals[i * F], F, xe, layer_alphas[l]);
            }
            if (max_deg_plus_1 % F) fail("FRI degree truncation");
            gen = pow(gen, F); max_deg_plus_1 /= F; dom /= F;
            pos.swap(folded); evaluations.swap(next);
        }
        if (p.fri_remainder.size() > max_deg_plus_1) fail("FRI remainder degree mismatch");
        if (hash_elements(p.fri_remainder.data(), p.fri_remainder.size()) != p.commitments[2 + nlayers]) fail("FRI remainder commitment mismatch");
        for (size_t i = 0; i < pos.size(); i++) {
            Fe x = mul(offset, pow(gen, pos[i])), r = FE_ZERO;
            for (auto& cf : p.fri_remainder) r = add(mul(r, x), cf);  // eval_horner_rev
            if (r != evaluations[i]) fail("invalid FRI remainder folding");
        }
    }
}

}  // namespace orc
