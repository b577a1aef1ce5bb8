#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory.

Sources of truth (none comes from the reference's own tests — it has no golden vectors):
  * field_kats.json   Python big-int arithmetic modulo p = 2^128 - 45*2^40 + 1
  * blake3_kats.json  the independent Python `blake3` module (1.0.8), inputs = official test-vector pattern i % 251,
                      lengths = every hash shape on the path plus chunk/block boundaries
  * ntt_kats.json     naive O(n^2) evaluation with Python big ints (interpolation, coset LDE with offset 3)
  * mimc_kats.json    Python restatement of src/helper.rs:213-233 (mimc_cipher / mimc_hash_matrix)
  * proof_pins.json   BLAKE3 of oracle proofs for seeded inputs — a regression pin of the oracle itself
                      ("self-generated": it pins conventions against drift, not against Winterfell)
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
P = 2**128 - 45 * 2**40 + 1


def field_kats():
    rng = random.Random(0xF128)
    edge = [0, 1, 2, 3, P - 1, P - 2, 2**64 - 1, 2**64, 2**96, 2**127, 45 * 2**40 - 1, 45 * 2**40, (P - 1) // 2]
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(P), rng.randrange(P)) for _ in range(200)]
    return [dict(a=str(a), b=str(b), mul=str(a * b % P), add=str((a + b) % P), sub=str((a - b) % P),
                 inv=str(pow(a, P - 2, P) if a else 0)) for a, b in pairs]


def blake3_kats():
    import blake3
    lens = [0, 1, 16, 32, 40, 63, 64, 65, 96, 112, 127, 128, 256, 1023, 1024, 1025, 1920, 2048, 2049, 3072, 3073, 3840, 4080, 4096, 8192 + 7]
    return [dict(len=n, digest=blake3.blake3(bytes(i % 251 for i in range(n))).hexdigest()) for n in lens]


def ntt_kats():
    rng = random.Random(0x177)
    out = []
    for n, beta in ((8, 2), (16, 4), (32, 8)):
        g = pow(3, (P - 1) // n, P)
        gN = pow(3, (P - 1) // (n * beta), P)
        coeffs = [rng.randrange(P) for _ in range(n)]
        evals = [sum(c * pow(g, i * k, P) for k, c in enumerate(coeffs)) % P for i in range(n)]
        lde = [sum(c * pow(3 * pow(gN, i, P), k, P) for k, c in enumerate(coeffs)) % P for i in range(n * beta)]
        out.append(dict(n=n, blowup=beta, coeffs=[str(x) for x in coeffs], evals=[str(x) for x in evals], lde=[str(x) for x in lde]))
    return out


def mimc_kats():
    rc = [i * 10**6 for i in range(1, 65)]

    def cipher(x, r, z):
        for _ in range(64):
            x = pow((x + r + z) % P, 7, P)
        return (x + z) % P
    w = [[42 * 10**6] * 9 for _ in range(6)]  # benches/bench_mimc.rs:41-42
    b = [10**6] * 6
    z = 0
    for i in range(6):
        for j in range(9):
            z = cipher(w[i][j], rc[j % 64], z)
        z = cipher(b[i], rc[i % 64], z)
    return dict(cipher=[dict(x=str(x), rc=str(r), z=str(zz), out=str(cipher(x, r, zz))) for x, r, zz in
                        ((1, 2, 0), (123456789, 10**6, 7), (P - 1, 64 * 10**6, P - 5))], bench_hash_matrix=str(z))


def proof_pins():
    import blake3
    from oracle import pyoracle as O
    from tests import common as T
    import zk_stark_project_b200 as Z
    O.build()
    O.set_threads(4)
    pins = {}
    for name, prover in (("mimc_w4_n64_b8", T.mimc_prover(4, 64, T.options(blowup=8))),
                         ("aggregation_16", T.aggregation_prover(16, T.options())),
                         ("training_bs1", T.training_prover(1, T.options())),
                         ("aggregation_16_reference_options", T.aggregation_prover(16, Z.ProofOptions.reference()))):
        tr = prover.build_trace()
        proof, ts, _ = O.prove(prover.describe(tr), tr.to_bytes())
        pins[name] = dict(proof_len=len(proof), proof_blake3=blake3.blake3(proof).hexdigest(), trace_root=bytes(ts.trace_root).hex(),
                          pow_nonce=int(ts.pow_nonce), n_positions=int(ts.n_positions))
    return pins


if __name__ == "__main__":
    for name, fn in (("field_kats", field_kats), ("blake3_kats", blake3_kats), ("ntt_kats", ntt_kats), ("mimc_kats", mimc_kats),
                     ("proof_pins", proof_pins)):
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
