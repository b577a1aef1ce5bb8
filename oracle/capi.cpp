// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED.
// C entry points so tests/ and bench.py's cpu_baseline leg can drive the oracle through ctypes.
#include <chrono>
#include "proof.h"

using namespace orc;
namespace orc { extern Fe* g_dump_comp_trace; }

extern "C" {

typedef struct {
    uint32_t air_id, trace_width;
    uint64_t trace_len;
    uint32_t num_queries, blowup, grinding_bits, field_extension, folding, rem_max_degree, batching_constraints, batching_deep;
    const uint8_t* pub_elems; uint64_t n_pub_elems;       // PublicInputs::to_elements(), 16 B LE each
    const uint32_t* assert_cols; const uint64_t* assert_steps; const uint8_t* assert_values; uint64_t n_assertions;
    const uint8_t* params; uint64_t n_params;             // aggregation: k; mimc: round constants
} orc_air_desc;

typedef struct {
    uint8_t trace_root[32], constraint_root[32], remainder_commitment[32];
    uint8_t constraint_alpha[16], z[16], deep_alpha[16];
    uint32_t n_fri_layers, n_positions;
    uint8_t fri_roots[16][32];
    uint8_t fri_alphas[16][16];
    uint64_t pow_nonce;
    uint32_t positions[256];
    int32_t comp_degree_ok, _pad;
} orc_transcript;

static Air to_air(const orc_air_desc* d) {
    Air a;
    a.id = d->air_id; a.w = d->trace_width; a.n = d->trace_len;
    a.opt.num_queries = d->num_queries; a.opt.blowup = d->blowup; a.opt.grinding = d->grinding_bits;
    a.opt.field_extension = d->field_extension; a.opt.folding = d->folding; a.opt.rem_max_degree = d->rem_max_degree;
    a.opt.batching_constraints = d->batching_constraints; a.opt.batching_deep = d->batching_deep;
    for (uint64_t i = 0; i < d->n_pub_elems; i++) a.pub_elems.push_back(Fe(fe_from_bytes(d->pub_elems + 16 * i).v));
    for (uint64_t i = 0; i < d->n_assertions; i++)
        a.assertions.push_back({d->assert_cols[i], d->assert_steps[i], Fe(fe_from_bytes(d->assert_values + 16 * i).v)});
    for (uint64_t i = 0; i < d->n_params; i++) a.params.push_back(Fe(fe_from_bytes(d->params + 16 * i).v));
    return a;
}
static void export_ts(const Transcript& t, orc_transcript* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    memcpy(o->trace_root, t.trace_root.b, 32); memcpy(o->constraint_root, t.constraint_root.b, 32);
    memcpy(o->remainder_commitment, t.remainder_commitment.b, 32);
    fe_to_bytes(t.constraint_alpha, o->constraint_alpha); fe_to_bytes(t.z, o->z); fe_to_bytes(t.deep_alpha, o->deep_alpha);
    o->n_fri_layers = (uint32_t)t.fri_roots.size();
    for (size_t i = 0; i < t.fri_roots.size() && i < 16; i++) { memcpy(o->fri_roots[i], t.fri_roots[i].b, 32); fe_to_bytes(t.fri_alphas[i], o->fri_alphas[i]); }
    o->pow_nonce = t.pow_nonce;
    o->n_positions = (uint32_t)t.positions.size();
    for (size_t i = 0; i < t.positions.size() && i < 256; i++) o->positions[i] = (uint32_t)t.positions[i];
    o->comp_degree_ok = t.comp_degree_ok;
}
static void set_err(char* err, size_t n, const char* m) { if (err && n) { strncpy(err, m, n - 1); err[n - 1] = 0; } }

void orc_set_threads(int t) { g_threads = t < 1 ? 1 : t; }
int orc_get_threads() { return g_threads; }

void orc_fe_add(const uint8_t* a, const uint8_t* b, uint8_t* o) { fe_to_bytes(add(fe_from_bytes(a), fe_from_bytes(b)), o); }
void orc_fe_sub(const uint8_t* a, const uint8_t* b, uint8_t* o) { fe_to_bytes(sub(fe_from_bytes(a), fe_from_bytes(b)), o); }
void orc_fe_mul(const uint8_t* a, const uint8_t* b, uint8_t* o) { fe_to_bytes(mul(fe_from_bytes(a), fe_from_bytes(b)), o); }
void orc_fe_inv(const uint8_t* a, uint8_t* o) { fe_to_bytes(inv(fe_from_bytes(a)), o); }
// elementwise batch versions for bulk KATs
void orc_fe_mul_many(const uint8_t* a, const uint8_t* b, uint8_t* o, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) fe_to_bytes(mul(fe_from_bytes(a + 16 * i), fe_from_bytes(b + 16 * i)), o + 16 * i);
}
void orc_blake3(const uint8_t* in, uint64_t len, uint8_t* out) { Digest d = blake3(in, len); memcpy(out, d.b, 32); }

void orc_interpolate(uint8_t* evals, uint64_t n) { interpolate_poly((Fe*)evals, n); }
void orc_interpolate_with_offset(uint8_t* evals, uint64_t n) { interpolate_poly_with_offset((Fe*)evals, n, fe_raw(GENERATOR)); }
void orc_lde(const uint8_t* coeffs, uint64_t n, uint64_t blowup, uint8_t* out) {
    std::vector<Fe> r = evaluate_poly_with_offset((const Fe*)coeffs, n, fe_raw(GENERATOR), blowup);
    memcpy(out, r.data(), r.size() * 16);
}
void orc_merkle_root(const uint8_t* leaves, uint64_t n, uint8_t* out) {
    std::vector<Digest> l(n);
    memcpy(l.data(), leaves, n * 32);
    MerkleTree t = merkle_new(std::move(l));
    memcpy(out, t.root().b, 32);
}
// column-major trace [w][n] -> trace commitment root and (optionally) the row-major LDE [N][w]
void orc_trace_commit(const uint8_t* trace, uint64_t n, uint64_t w, uint64_t blowup, uint8_t* root_out, uint8_t* lde_out, uint8_t* polys_out) {
    std::vector<Fe> lde(n * blowup * w);
    std::vector<std::vector<Fe>> polys(w);
    parallel_for(w, [&](size_t a, size_t b) {
        for (size_t j = a; j < b; j++) {
            polys[j].assign((const Fe*)trace + j * n, (const Fe*)trace + (j + 1) * n);
            interpolate_poly(polys[j].data(), n);
            std::vector<Fe> col = evaluate_poly_with_offset(polys[j].data(), n, fe_raw(GENERATOR), blowup);
            for (size_t r = 0; r < n * blowup; r++) lde[r * w + j] = col[r];
        }
    });
    std::vector<Digest> leaves(n * blowup);
    parallel_for(n * blowup, [&](size_t a, size_t b) { for (size_t r = a; r < b; r++) leaves[r] = hash_elements(&lde[r * w], w); });
    MerkleTree t = merkle_new(std::move(leaves));
    memcpy(root_out, t.root().b, 32);
    if (lde_out) memcpy(lde_out, lde.data(), lde.size() * 16);
    if (polys_out) for (size_t j = 0; j < w; j++) memcpy(polys_out + j * n * 16, polys[j].data(), n * 16);
}

// MiMC chain trace, column-major [w][n]: col_j[0] = seeds[j], col_j[i+1] = (col_j[i] + rc[i mod L])^7
void orc_mimc_trace(const uint8_t* seeds, uint64_t w, uint64_t n, const uint8_t* rc, uint64_t L, uint8_t* out) {
    parallel_for(w, [&](size_t a, size_t b) {
        for (size_t j = a; j < b; j++) {
            Fe x = fe_from_bytes(seeds + 16 * j);
            Fe* col = (Fe*)out + j * n;
            for (uint64_t i = 0; i < n; i++) {
                col[i] = x;
                Fe t = add(x, fe_from_bytes(rc + 16 * (i % L)));
                Fe t2 = mul(t, t), t4 = mul(t2, t2), t6 = mul(t4, t2);
                x = mul(t6, t);
            }
        }
    });
}
// src/helper.rs:213-220 mimc_cipher and :222-233 mimc_hash_matrix (used for the aggregation digest)
void orc_mimc_cipher(const uint8_t* input, const uint8_t* rc, const uint8_t* z, uint8_t* out) {
    Fe inp = fe_from_bytes(input), r = fe_from_bytes(rc), zz = fe_from_bytes(z);
    for (int i = 0; i < 64; i++) { Fe a = add(add(inp, r), zz); Fe a2 = mul(a, a), a4 = mul(a2, a2), a6 = mul(a4, a2); inp = mul(a6, a); }
    fe_to_bytes(add(inp, zz), out);
}

int orc_prove(const orc_air_desc* d, const uint8_t* trace_colmajor, uint64_t force_nonce, uint8_t** proof_out,
              uint64_t* proof_len, orc_transcript* ts, double* seconds, char* err, uint64_t errlen) {
    try {
        Air air = to_air(d);
        Transcript t;
        auto t0 = std::chrono::steady_clock::now();
        Proof p = prove(air, (const Fe*)trace_colmajor, &t, force_nonce);
        std::vector<uint8_t> b = proof_to_bytes(p);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        export_ts(t, ts);
        if (proof_out) { *proof_out = (uint8_t*)malloc(b.size()); memcpy(*proof_out, b.data(), b.size()); }
        if (proof_len) *proof_len = b.size();
        return 0;
    } catch (const std::exception& e) { set_err(err, errlen, e.what()); return -1; }
}
int orc_verify(const orc_air_desc* d, const uint8_t* proof, uint64_t len, orc_transcript* ts, char* err, uint64_t errlen) {
    try {
        Air air = to_air(d);
        Transcript t;
        verify(air, proof, len, &t);
        export_ts(t, ts);
        return 0;
    } catch (const std::exception& e) { set_err(err, errlen, e.what()); return -1; }
}
void orc_free(void* p) { free(p); }
// test hook: CompositionPolyTrace (ce_blowup * trace_len elements) of the next orc_prove call; pass NULL to switch it off
void orc_dump_comp_trace(uint8_t* out) { g_dump_comp_trace = (Fe*)out; }

}  // extern "C"
