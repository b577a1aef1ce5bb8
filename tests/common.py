"""Shared seeded inputs for the parity tests (none of the reference's thread_rng sites are used)."""
import random

import numpy as np

import zk_stark_project_b200 as Z
from zk_stark_project_b200 import field as F
from zk_stark_project_b200.training import AC, FE

P = Z.P


def splitmix(seed):
    state = seed & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    return nxt


from zk_stark_project_b200.synthetic import random_felts, synthetic_training_air  # noqa: E402,F401


def options(blowup=16, queries=40, grinding=8):
    return Z.ProofOptions(queries, blowup, grinding, Z.FieldExtension.NONE, 16, 7)


def model(rng, sigma):
    w = [[F.f64_to_signed_felt(rng.gauss(0, sigma), 1e6) for _ in range(FE)] for _ in range(AC)]
    b = [F.f64_to_signed_felt(rng.gauss(0, sigma), 1e6) for _ in range(AC)]
    return ([[v for v, _ in r] for r in w], [[s for _, s in r] for r in w], [v for v, _ in b], [s for _, s in b])


def training_prover(bs, opts, seed=0x5EED0002):
    """tests/integration_tests.rs:14-58 shaped inputs, but seeded."""
    rng = random.Random(seed)
    w, ws, b, bs_ = model(rng, 1.0)
    x = [[Z.f64_to_felt((i + j) * 0.1) for j in range(FE)] for i in range(bs)]
    xs = [[0] * FE for _ in range(bs)]
    y = [[Z.f64_to_felt(1.0) if a == i % AC else 0 for a in range(AC)] for i in range(bs)]
    return Z.TrainingUpdateProver(opts, w, b, ws, bs_, x, xs, y, Z.f64_to_felt(0.01), Z.f64_to_felt(1e6), bs, seed=seed)


def aggregation_prover(num_updates, opts, seed=0x5EED0003):
    """src/main.rs:440-459 shaped inputs, but seeded."""
    rng = random.Random(seed)
    gw, _, gb, _ = model(rng, 10000.0)
    reps = [rng.randrange(2**64) for _ in range(num_updates)]
    lw = [[[Z.f64_to_felt(r / 1e6)] * FE for _ in range(AC)] for r in reps]
    lb = [[Z.f64_to_felt(r / 1e6)] * AC for r in reps]
    return Z.GlobalUpdateProver(opts, gw, gb, lw, lb, Z.f64_to_felt(float(num_updates)), seed=seed)


def mimc_prover(width, steps, opts):
    return Z.MimcProver(opts, [j + 1 for j in range(width)], steps)


TS_FIELDS = ["trace_root", "constraint_alpha", "constraint_root", "z", "deep_alpha", "n_fri_layers", "fri_roots", "fri_alphas",
             "remainder_commitment", "pow_nonce", "n_positions", "positions"]


def transcript_diff(a, b):
    """First field where two transcripts (oracle / product ctypes structs with identical layout) differ."""
    for f in TS_FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        xb = bytes(x) if hasattr(x, "__len__") else x
        yb = bytes(y) if hasattr(y, "__len__") else y
        if xb != yb:
            return f
    return None


def chacha20_block(key, counter):
    """ChaCha20 block function (20 rounds, 64-bit block counter in words 12-13, zero nonce): the device-side mask generator of
    zkb_training_trace_device restated for the tests.  Returns 64 keystream bytes."""
    import struct
    m = 0xFFFFFFFF
    rotl = lambda v, c: ((v << c) & m) | (v >> (32 - c))
    x = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(struct.unpack("<8I", key)) + [counter & m, (counter >> 32) & m, 0, 0]
    s = list(x)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & m; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & m; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & m; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & m; x[b] = rotl(x[b] ^ x[c], 7)
    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & m for a, b in zip(x, s)])


def training_mask(key, row, col, half):
    """Mask of trace cell (row, col) as zkb_training_trace_device generates it (include/zkb200.h)."""
    blocks = (half + 7) // 8
    ks = chacha20_block(key, row * blocks + col // 8)
    return int.from_bytes(ks[8 * (col % 8):8 * (col % 8) + 8], "little")
