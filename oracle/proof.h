// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED.
//
// Proof object and wire format: restatement of winter-air 0.12 `proof::{Proof, Context, Commitments,
// Queries, OodFrame}`, winter-fri `FriProof` and winter-crypto `BatchMerkleProof` serialisation
// (`Proof::to_bytes()`, consumed by the reference at src/main.rs:229,425,469).  Every byte-level
// convention here is recalled from upstream, not read — the product
// (zk_stark_project_b200/csrc/proof_format.hpp) restates the same layout independently and the two are
// compared byte for byte in tests/.
#pragma once
#include "air.h"

namespace orc {

struct Transcript {  // intermediate Fiat-Shamir values, exported for stage-by-stage comparison
    Digest trace_root, constraint_root, remainder_commitment;
    Fe constraint_alpha, z, deep_alpha;
    std::vector<Digest> fri_roots;
    std::vector<Fe> fri_alphas;
    uint64_t pow_nonce = 0;
    std::vector<size_t> positions;
    int comp_degree_ok = 1;
};

struct Proof {
    // Context
    uint32_t trace_width = 0, trace_len = 0;
    Options opt;
    uint8_t num_unique_queries = 0;
    std::vector<Digest> commitments;  // trace root, constraint root, FRI layer roots, remainder commitment
    std::vector<uint8_t> trace_query_values, trace_query_proof;            // Queries (main segment)
    std::vector<uint8_t> constraint_query_values, constraint_query_proof;  // Queries
    std::vector<Fe> ood_trace_states;  // [T_0(z), T_0(zg), T_1(z), T_1(zg), ...]
    std::vector<Fe> ood_constraint_evals;
    std::vector<std::vector<uint8_t>> fri_layer_values, fri_layer_paths;
    std::vector<Fe> fri_remainder;  // reversed coefficients
    uint64_t pow_nonce = 0;
};

static inline std::vector<uint8_t> batch_proof_to_bytes(const BatchMerkleProof& p) {
    Writer w;
    w.u8(p.depth);
    w.usize(p.nodes.size());
    for (auto& v : p.nodes) { w.usize(v.size()); for (auto& d : v) w.digest(d); }
    return w.buf;
}
static inline BatchMerkleProof batch_proof_from_bytes(const std::vector<uint8_t>& b) {
    Reader r(b.data(), b.size());
    BatchMerkleProof p;
    p.depth = r.u8();
    size_t n = r.usize();
    if (n > b.size()) throw std::runtime_error("proof: bad merkle proof");
    p.nodes.resize(n);
    for (auto& v : p.nodes) { size_t k = r.usize(); if (k > 64) throw std::runtime_error("proof: bad merkle proof"); v.resize(k); for (auto& d : v) d = r.digest(); }
    if (!r.done()) throw std::runtime_error("proof: trailing bytes in merkle proof");
    return p;
}

static inline std::vector<uint8_t> proof_to_bytes(const Proof& p) {
    Writer w;
    // Context: TraceInfo
    w.u8((uint8_t)p.trace_width); w.u8(0); w.u8(0);  // main width, aux width, aux rands
    w.u8((uint8_t)ilog2(p.trace_len));
    w.u16(0);  // trace meta
    // field modulus
    w.u8(16); { u128 m = P; w.bytes(&m, 16); }
    // ProofOptions
    w.u8((uint8_t)p.opt.num_queries); w.u8((uint8_t)p.opt.blowup); w.u8((uint8_t)p.opt.grinding);
    w.u8((uint8_t)p.opt.field_extension); w.u8((uint8_t)p.opt.folding); w.u8((uint8_t)p.opt.rem_max_degree);
    w.u8((uint8_t)p.opt.batching_constraints); w.u8((uint8_t)p.opt.batching_deep);
    w.u8(1); w.u8(1);  // PartitionOptions { num_partitions: 1, hash_rate: 1 }
    w.u8(p.num_unique_queries);
    // Commitments
    w.u16((uint16_t)(p.commitments.size() * 32));
    for (auto& d : p.commitments) w.digest(d);
    // trace Queries (one segment), constraint Queries
    w.usize(p.trace_query_values.size()); w.bytes(p.trace_query_values.data(), p.trace_query_values.size());
    w.usize(p.trace_query_proof.size()); w.bytes(p.trace_query_proof.data(), p.trace_query_proof.size());
    w.usize(p.constraint_query_values.size()); w.bytes(p.constraint_query_values.data(), p.constraint_query_values.size());
    w.usize(p.constraint_query_proof.size()); w.bytes(p.constraint_query_proof.data(), p.constraint_query_proof.size());
    // OodFrame: trace states (u8 frame size + elements), lagrange kernel (empty), constraint evaluations
    w.u16((uint16_t)(1 + p.ood_trace_states.size() * 16)); w.u8(2);
    for (auto& e : p.ood_trace_states) w.fe(e);
    w.u16(0);
    w.u16((uint16_t)(p.ood_constraint_evals.size() * 16));
    for (auto& e : p.ood_constraint_evals) w.fe(e);
    // FriProof
    w.u8((uint8_t)p.fri_layer_values.size());
    for (size_t i = 0; i < p.fri_layer_values.size(); i++) {
        w.u32((uint32_t)p.fri_layer_values[i].size()); w.bytes(p.fri_layer_values[i].data(), p.fri_layer_values[i].size());
        w.u32((uint32_t)p.fri_layer_paths[i].size()); w.bytes(p.fri_layer_paths[i].data(), p.fri_layer_paths[i].size());
    }
    w.u16((uint16_t)(p.fri_remainder.size() * 16));
    for (auto& e : p.fri_remainder) w.fe(e);
    w.u8(1);  // num_partitions
    w.u64_(p.pow_nonce);
    return w.buf;
}

static inline Proof proof_from_bytes(const uint8_t* data, size_t len) {
    Reader r(data, len);
    Proof p;
    p.trace_width = r.u8();
    if (r.u8() != 0 || r.u8() != 0) throw std::runtime_error("proof: aux trace segments are not supported");
    uint8_t ln = r.u8();
    if (ln < 3 || ln > 32) throw std::runtime_error("proof: bad trace length");
    p.trace_len = (uint32_t)1 << ln;
    if (r.u16() != 0) throw std::runtime_error("proof: trace metadata is not supported");
    if (r.u8() != 16) throw std::runtime_error("proof: bad modulus length");
    { auto m = r.take(16); u128 P_ = P; if (memcmp(m.data(), &P_, 16)) throw std::runtime_error("proof: wrong field modulus"); }
    p.opt.num_queries = r.u8(); p.opt.blowup = r.u8(); p.opt.grinding = r.u8(); p.opt.field_extension = r.u8();
    p.opt.folding = r.u8(); p.opt.rem_max_degree = r.u8(); p.opt.batching_constraints = r.u8(); p.opt.batching_deep = r.u8();
    if (r.u8() != 1 || r.u8() != 1) throw std::runtime_error("proof: partitions are not supported");
    p.num_unique_queries = r.u8();
    size_t clen = r.u16();
    if (clen % 32) throw std::runtime_error("proof: bad commitments length");
    p.commitments.resize(clen / 32);
    for (auto& d : p.commitments) d = r.digest();
    { size_t k = r.usize(); p.trace_query_values = r.take(k); }
    { size_t k = r.usize(); p.trace_query_proof = r.take(k); }
    { size_t k = r.usize(); p.constraint_query_values = r.take(k); }
    { size_t k = r.usize(); p.constraint_query_proof = r.take(k); }
    size_t tl = r.u16();
    if (tl < 1 || (tl - 1) % 16) throw std::runtime_error("proof: bad OOD trace frame");
    if (r.u8() != 2) throw std::runtime_error("proof: bad OOD frame size");
    p.ood_trace_states.resize((tl - 1) / 16);
    for (auto& e : p.ood_trace_states) e = r.fe();
    if (r.u16() != 0) throw std::runtime_error("proof: lagrange kernel frames are not supported");
    size_t el = r.u16();
    if (el % 16) throw std::runtime_error("proof: bad OOD evaluations");
    p.ood_constraint_evals.resize(el / 16);
    for (auto& e : p.ood_constraint_evals) e = r.fe();
    size_t nl = r.u8();
    p.fri_layer_values.resize(nl); p.fri_layer_paths.resize(nl);
    for (size_t i = 0; i < nl; i++) {
        size_t a = r.u32(); p.fri_layer_values[i] = r.take(a);
        size_t b = r.u32(); p.fri_layer_paths[i] = r.take(b);
    }
    size_t rl = r.u16();
    if (rl % 16) throw std::runtime_error("proof: bad remainder");
    p.fri_remainder.resize(rl / 16);
    for (auto& e : p.fri_remainder) e = r.fe();
    if (r.u8() != 1) throw std::runtime_error("proof: bad partition count");
    p.pow_nonce = r.u64_();
    if (!r.done()) throw std::runtime_error("proof: trailing bytes");
    return p;
}

Proof prove(const Air& air, const Fe* trace_colmajor, Transcript* ts, uint64_t force_nonce);
void verify(const Air& air, const uint8_t* proof, size_t len, Transcript* ts);  // throws on rejection

}  // namespace orc
