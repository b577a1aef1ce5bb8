// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// BLAKE3-256, portable restatement of the public BLAKE3 specification (the reference pulls
// `blake3 1.5.4` through winter-crypto 0.12.0, Cargo.lock:195-205, 1280-1289; neither is vendored
// under /root/reference).  Used as winter-crypto `Blake3_256<Felt>`, selected by the reference at
// src/training/prover.rs:225 and src/aggregation/prover.rs:198:
//   hash_elements(xs) = blake3(raw LE bytes of xs)      (IS_CANONICAL field)
//   merge(a,b)        = blake3(a || b)                  (64 bytes)
//   merge_with_int(s,v) = blake3(s || v.to_le_bytes())  (40 bytes)
// Pinned in tests/ against the Python `blake3` module for every input shape of SURVEY Appendix C.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace orc {

struct Digest {
    uint8_t b[32];
    bool operator==(const Digest& o) const { return memcmp(b, o.b, 32) == 0; }
    bool operator!=(const Digest& o) const { return !(*this == o); }
};

namespace b3 {
static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
                               0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
// out_cv = first 8 words of the compression output
static inline void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                            uint32_t block_len, uint32_t flags, uint32_t out_cv[8]) {
    uint32_t s[16], m[16], t[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = IV[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]);  g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]); g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]); g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        if (r < 6) { for (int i = 0; i < 16; i++) t[i] = m[PERM[i]]; memcpy(m, t, 64); }
    }
    for (int i = 0; i < 8; i++) out_cv[i] = s[i] ^ s[i + 8];
}

// chaining value of one chunk (<=1024 bytes); `root` marks a single-chunk input
static inline void chunk_cv(const uint8_t* in, size_t len, uint64_t chunk_idx, bool root, uint32_t out[8]) {
    uint32_t cv[8];
    memcpy(cv, IV, 32);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint32_t block[16] = {0};
        size_t bl = (b + 1 < nblocks) ? 64 : len - 64 * b;
        memcpy(block, in + 64 * b, bl);
        uint32_t flags = 0;
        if (b == 0) flags |= CHUNK_START;
        if (b + 1 == nblocks) { flags |= CHUNK_END; if (root) flags |= ROOT; }
        compress(cv, block, chunk_idx, (uint32_t)bl, flags, cv);
    }
    memcpy(out, cv, 32);
}
static inline void parent_cv(const uint32_t l[8], const uint32_t r[8], bool root, uint32_t out[8]) {
    uint32_t block[16];
    memcpy(block, l, 32); memcpy(block + 8, r, 32);
    compress(IV, block, 0, 64, PARENT | (root ? ROOT : 0), out);
}
// subtree over chunks [c0, c0+nc) of an input with more than one chunk overall
static inline void subtree(const uint8_t* in, size_t len, uint64_t c0, size_t nc, bool root, uint32_t out[8]) {
    if (nc == 1) { chunk_cv(in, len, c0, false, out); return; }
    size_t left = 1;
    while (left * 2 <= nc - 1) left *= 2;  // largest power of two <= nc-1
    uint32_t l[8], r[8];
    subtree(in, left * 1024, c0, left, false, l);
    subtree(in + left * 1024, len - left * 1024, c0 + left, nc - left, false, r);
    parent_cv(l, r, root, out);
}
}  // namespace b3

static inline Digest blake3(const uint8_t* in, size_t len) {
    uint32_t cv[8];
    size_t nc = len <= 1024 ? 1 : (len + 1023) / 1024;
    if (nc == 1) b3::chunk_cv(in, len, 0, true, cv);
    else b3::subtree(in, len, 0, nc, true, cv);
    Digest d;
    memcpy(d.b, cv, 32);
    return d;
}
static inline Digest merge(const Digest& a, const Digest& b) {
    uint8_t buf[64];
    memcpy(buf, a.b, 32); memcpy(buf + 32, b.b, 32);
    return blake3(buf, 64);
}
static inline Digest merge_with_int(const Digest& seed, uint64_t v) {
    uint8_t buf[40];
    memcpy(buf, seed.b, 32); memcpy(buf + 32, &v, 8);
    return blake3(buf, 40);
}

}  // namespace orc
