// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED.
//
// CPU restatement of the Winterfell 0.12.0 building blocks the reference's prove() runs through
// (winter-math fft/polynom, winter-crypto merkle + random coin, winter-utils serialisation).
// None of these crates is vendored under /root/reference (Cargo.lock:1267-1367); each function
// names the upstream module it restates and the reference line that selects it.
#pragma once
#include <algorithm>
#include <cassert>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <thread>
#include <vector>
#include "field.h"
#include "blake3.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// threads (stands in for winter-utils "concurrent"/rayon, Cargo.toml:11)
extern int g_threads;
static inline void parallel_for(size_t n, const std::function<void(size_t, size_t)>& fn) {
    int t = g_threads;
    if (t <= 1 || n < 2) { fn(0, n); return; }
    if ((size_t)t > n) t = (int)n;
    std::vector<std::thread> th;
    size_t chunk = (n + t - 1) / t;
    for (int i = 0; i < t; i++) {
        size_t a = i * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=, &fn] { fn(a, b); });
    }
    for (auto& x : th) x.join();
}

static inline int ilog2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }

// ---------------------------------------------------------------------------------------------
// winter-math `fft`: only the mathematical contracts matter for parity
//   interpolate_poly(evals over <w_n>)            -> coefficients
//   evaluate_poly_with_offset(p, offset, blowup)  -> [p(offset * w_N^i)] natural order
// twiddles w^0 .. w^(n/2-1), cached per (n, root) and per thread (winter-math precomputes them once per domain too)
static inline const std::vector<Fe>& fft_twiddles(size_t n, Fe root) {
    static thread_local std::map<std::pair<size_t, u128>, std::vector<Fe>> cache;
    auto key = std::make_pair(n, root.v);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<Fe> tw(n / 2 ? n / 2 : 1);
    tw[0] = FE_ONE;
    for (size_t i = 1; i < n / 2; i++) tw[i] = mul(tw[i - 1], root);
    return cache.emplace(key, std::move(tw)).first->second;
}
static inline void fft_in_place(Fe* a, size_t n, Fe root) {
    // iterative radix-2 DIT, natural in / natural out
    for (size_t i = 1, j = 0; i < n; i++) {  // bit-reversal permutation
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const std::vector<Fe>& tw = fft_twiddles(n, root);
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len / 2, step = n / len;
        for (size_t s = 0; s < n; s += len)
            for (size_t k = 0; k < half; k++) {
                Fe u = a[s + k], v = mul(a[s + k + half], tw[k * step]);
                a[s + k] = add(u, v);
                a[s + k + half] = sub(u, v);
            }
    }
}
static inline void interpolate_poly(Fe* evals, size_t n) {
    Fe root = inv(get_root_of_unity(ilog2(n)));
    fft_in_place(evals, n, root);
    Fe ninv = inv(fe_raw((u128)n));
    for (size_t i = 0; i < n; i++) evals[i] = mul(evals[i], ninv);
}
static inline void interpolate_poly_with_offset(Fe* evals, size_t n, Fe offset) {
    interpolate_poly(evals, n);
    Fe oi = inv(offset), f = FE_ONE;
    for (size_t i = 0; i < n; i++) { evals[i] = mul(evals[i], f); f = mul(f, oi); }
}
static inline std::vector<Fe> evaluate_poly_with_offset(const Fe* p, size_t n, Fe offset, size_t blowup) {
    size_t N = n * blowup;
    std::vector<Fe> out(N), tmp(n);
    Fe gN = get_root_of_unity(ilog2(N)), gn = get_root_of_unity(ilog2(n));
    for (size_t k = 0; k < blowup; k++) {
        Fe s = mul(offset, pow(gN, k)), f = FE_ONE;
        for (size_t m = 0; m < n; m++) { tmp[m] = mul(p[m], f); f = mul(f, s); }
        fft_in_place(tmp.data(), n, gn);
        for (size_t i = 0; i < n; i++) out[i * blowup + k] = tmp[i];
    }
    return out;
}
// winter-math polynom::eval (Horner)
static inline Fe poly_eval(const Fe* p, size_t n, Fe x) {
    Fe r = FE_ZERO;
    for (size_t i = n; i-- > 0;) r = add(mul(r, x), p[i]);
    return r;
}
// winter-math polynom::syn_div_in_place(p, 1, a): p(x) / (x - a), remainder dropped
static inline void syn_div_in_place(Fe* p, size_t n, Fe a) {
    Fe c = FE_ZERO;
    for (size_t i = n; i-- > 0;) { Fe t = add(p[i], mul(a, c)); p[i] = c; c = t; }
}

// ---------------------------------------------------------------------------------------------
// winter-crypto hashing of field elements (Blake3_256::hash_elements, IS_CANONICAL path)
static inline Digest hash_elements(const Fe* e, size_t n) {
    return blake3((const uint8_t*)e, n * 16);  // Fe is a bare little-endian u128
}

// winter-crypto merkle::MerkleTree (selected at src/training/prover.rs:226)
struct MerkleTree {
    std::vector<Digest> leaves, nodes;  // nodes[1] = root, nodes[n/2 .. n) built from leaves
    size_t depth() const { return ilog2(leaves.size()); }
    const Digest& root() const { return nodes[1]; }
};
static inline MerkleTree merkle_new(std::vector<Digest> leaves) {
    size_t n = leaves.size();
    if (n < 2 || (n & (n - 1))) throw std::runtime_error("merkle: leaf count must be a power of two >= 2");
    MerkleTree t;
    t.nodes.resize(n);
    memset(t.nodes[0].b, 0, 32);
    Digest* L = leaves.data();
    Digest* Nn = t.nodes.data();
    parallel_for(n / 2, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; i++) Nn[n / 2 + i] = merge(L[2 * i], L[2 * i + 1]);
    });
    for (size_t lvl = n / 4; lvl >= 1; lvl >>= 1) {
        parallel_for(lvl, [&](size_t a, size_t b) {
            for (size_t i = lvl + a; i < lvl + b; i++) Nn[i] = merge(Nn[2 * i], Nn[2 * i + 1]);
        });
        if (lvl == 1) break;
    }
    t.leaves = std::move(leaves);
    return t;
}

// winter-crypto BatchMerkleProof { nodes: Vec<Vec<Digest>>, depth: u8 }  (leaves not included)
struct BatchMerkleProof {
    std::vector<std::vector<Digest>> nodes;
    uint8_t depth = 0;
};
// MerkleTree::prove_batch (restated incl. its `nodes[i].push` indexing quirk)
static inline BatchMerkleProof merkle_prove_batch(const MerkleTree& t, const std::vector<size_t>& indexes) {
    size_t n = t.leaves.size();
    std::set<size_t> have(indexes.begin(), indexes.end());
    if (have.size() != indexes.size()) throw std::runtime_error("merkle: duplicate index");
    std::set<size_t> norm;
    for (size_t i : indexes) { if (i >= n) throw std::runtime_error("merkle: index out of range"); norm.insert(i & ~(size_t)1); }
    BatchMerkleProof p;
    p.depth = (uint8_t)t.depth();
    std::vector<size_t> next;
    for (size_t idx : norm) {
        std::vector<Digest> missing;
        for (size_t i = idx; i < idx + 2; i++) if (!have.count(i)) missing.push_back(t.leaves[i]);
        p.nodes.push_back(missing);
        next.push_back((idx + n) >> 1);
    }
    for (size_t d = 1; d < t.depth(); d++) {
        std::vector<size_t> cur = next;
        next.clear();
        size_t i = 0;
        while (i < cur.size()) {
            size_t sib = cur[i] ^ 1;
            if (i + 1 < cur.size() && cur[i + 1] == sib) i++;
            else p.nodes[i].push_back(t.nodes[sib]);
            next.push_back(sib >> 1);
            i++;
        }
    }
    return p;
}
// BatchMerkleProof::get_root (verifier side); leaves given in the order of `indexes`
static inline bool merkle_batch_root(const BatchMerkleProof& p, const std::vector<size_t>& indexes,
                                     const std::vector<Digest>& leaves, Digest* root_out) {
    if (indexes.size() != leaves.size() || indexes.empty()) return false;
    size_t depth = p.depth, n = (size_t)1 << depth;
    std::map<size_t, size_t> index_map;  // leaf index -> position in `leaves`
    for (size_t i = 0; i < indexes.size(); i++) {
        if (indexes[i] >= n || index_map.count(indexes[i])) return false;
        index_map[indexes[i]] = i;
    }
    std::set<size_t> norm;
    for (size_t i : indexes) norm.insert(i & ~(size_t)1);
    if (norm.size() != p.nodes.size()) return false;
    std::map<size_t, Digest> v;
    std::vector<size_t> next, ptr;
    size_t i = 0;
    for (size_t idx : norm) {
        Digest l, r;
        auto a = index_map.find(idx), b = index_map.find(idx + 1);
        if (a != index_map.end()) {
            l = leaves[a->second];
            if (b != index_map.end()) { r = leaves[b->second]; ptr.push_back(0); }
            else { if (p.nodes[i].empty()) return false; r = p.nodes[i][0]; ptr.push_back(1); }
        } else {
            if (p.nodes[i].empty() || b == index_map.end()) return false;
            l = p.nodes[i][0]; r = leaves[b->second]; ptr.push_back(1);
        }
        size_t parent = (n + idx) >> 1;
        v[parent] = merge(l, r);
        next.push_back(parent);
        i++;
    }
    for (size_t d = 1; d < depth; d++) {
        std::vector<size_t> cur = next;
        next.clear();
        size_t k = 0;
        while (k < cur.size()) {
            size_t node = cur[k], sib = node ^ 1;
            Digest sd;
            if (k + 1 < cur.size() && cur[k + 1] == sib) { if (!v.count(sib)) return false; sd = v[sib]; k++; }
            else { if (p.nodes[k].size() <= ptr[k]) return false; sd = p.nodes[k][ptr[k]++]; }
            if (!v.count(node)) return false;
            Digest nd = v[node];
            v[node >> 1] = (node & 1) ? merge(sd, nd) : merge(nd, sd);
            next.push_back(node >> 1);
            k++;
        }
    }
    auto it = v.find(1);
    if (it == v.end()) return false;
    *root_out = it->second;
    return true;
}

// ---------------------------------------------------------------------------------------------
// winter-crypto random::DefaultRandomCoin<Blake3_256> (selected at src/training/prover.rs:227)
struct Coin {
    Digest seed;
    uint64_t counter = 0;
    static Coin create(const std::vector<Fe>& elems) { Coin c; c.seed = hash_elements(elems.data(), elems.size()); c.counter = 0; return c; }
    void reseed(const Digest& d) { seed = merge(seed, d); counter = 0; }
    Digest next() { counter++; return merge_with_int(seed, counter); }
    Fe draw() {
        for (int i = 0; i < 1000; i++) {
            Digest d = next();
            u128 v; memcpy(&v, d.b, 16);
            if (v < P) return fe_raw(v);  // BaseElement::from_random_bytes rejects >= p
        }
        throw std::runtime_error("coin: failed to draw");
    }
    uint32_t check_leading_zeros(uint64_t value) const {
        Digest d = merge_with_int(seed, value);
        uint64_t head; memcpy(&head, d.b, 8);
        return head == 0 ? 64 : (uint32_t)__builtin_ctzll(head);
    }
    std::vector<size_t> draw_integers(size_t num, size_t domain, uint64_t nonce) {
        seed = merge_with_int(seed, nonce);
        counter = 0;
        std::vector<size_t> v;
        uint64_t mask = (uint64_t)domain - 1;
        for (int i = 0; i < 1000 && v.size() < num; i++) {
            Digest d = next();
            uint64_t x; memcpy(&x, d.b, 8);
            v.push_back((size_t)(x & mask));
        }
        if (v.size() != num) throw std::runtime_error("coin: failed to draw integers");
        return v;
    }
};

// ---------------------------------------------------------------------------------------------
// winter-utils ByteWriter / ByteReader (vint64 `write_usize`)
struct Writer {
    std::vector<uint8_t> buf;
    void u8(uint8_t v) { buf.push_back(v); }
    void u16(uint16_t v) { bytes(&v, 2); }
    void u32(uint32_t v) { bytes(&v, 4); }
    void u64_(uint64_t v) { bytes(&v, 8); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; buf.insert(buf.end(), q, q + n); }
    void fe(Fe x) { bytes(&x.v, 16); }
    void digest(const Digest& d) { bytes(d.b, 32); }
    void usize(uint64_t value) {
        int zeros = value == 0 ? 64 : __builtin_clzll(value);
        int len = (zeros > 0 ? zeros - 1 : 0) / 7;
        int length = 9 - std::min(len, 8);
        if (length == 9) { u8(0); u64_(value); }
        else { uint64_t enc = ((value << 1) | 1) << (length - 1); bytes(&enc, length); }
    }
};
struct Reader {
    const uint8_t* p; size_t n, pos = 0;
    Reader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    void need(size_t k) { if (pos + k > n) throw std::runtime_error("reader: unexpected end of proof"); }
    uint8_t u8() { need(1); return p[pos++]; }
    uint16_t u16() { need(2); uint16_t v; memcpy(&v, p + pos, 2); pos += 2; return v; }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p + pos, 4); pos += 4; return v; }
    uint64_t u64_() { need(8); uint64_t v; memcpy(&v, p + pos, 8); pos += 8; return v; }
    Fe fe() { need(16); u128 v; memcpy(&v, p + pos, 16); pos += 16; if (v >= P) throw std::runtime_error("reader: non-canonical element"); return fe_raw(v); }
    Digest digest() { need(32); Digest d; memcpy(d.b, p + pos, 32); pos += 32; return d; }
    std::vector<uint8_t> take(size_t k) { need(k); std::vector<uint8_t> v(p + pos, p + pos + k); pos += k; return v; }
    uint64_t usize() {
        uint8_t first = u8();
        int length = first == 0 ? 9 : __builtin_ctz(first) + 1;
        if (length == 9) return u64_();
        uint64_t enc = first;
        for (int i = 1; i < length; i++) enc |= (uint64_t)u8() << (8 * i);
        return enc >> length;
    }
    bool done() const { return pos == n; }
};

}  // namespace orc
