"""CPU: the oracle against golden vectors that come from independent sources (Python big ints, the `blake3` module,
naive polynomial evaluation).  The reference itself holds no golden vectors for this path, so these pin the
oracle's mathematics; its Winterfell conventions stay "parity unpinned" (oracle/README.md)."""
import json
import os

import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
P = 2**128 - 45 * 2**40 + 1


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def test_field_kats(oracle):
    for k in load("field_kats"):
        a, b = int(k["a"]), int(k["b"])
        assert oracle.fe_op("mul", a, b) == int(k["mul"])
        assert oracle.fe_op("add", a, b) == int(k["add"])
        assert oracle.fe_op("sub", a, b) == int(k["sub"])
        assert oracle.fe_op("inv", a) == int(k["inv"])


def test_field_constants(oracle):
    # winter-math f128: generator 3, two-adicity 40, root of unity
    root = 23953097886125630542083529559205016746
    assert pow(3, (P - 1) >> 40, P) == root and pow(root, 1 << 39, P) == P - 1
    # Felt::new(u128::MAX) reduces once: src/signed.rs:3
    from zk_stark_project_b200 import field as F
    assert F.MAX == 45 * 2**40 - 2


def test_blake3_kats(oracle):
    for k in load("blake3_kats"):
        data = bytes(i % 251 for i in range(k["len"]))
        assert oracle.blake3(data).hex() == k["digest"], k["len"]


def test_blake3_official_vectors(oracle):
    # first bytes of the official BLAKE3 test vectors (input i % 251)
    assert oracle.blake3(b"").hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"
    assert oracle.blake3(bytes([0])).hex().startswith("2d3adedff11b61f14c886e35afa036736dcd87a74d27b5c1510225d0f592e213")


def test_ntt_kats(oracle):
    for k in load("ntt_kats"):
        coeffs, evals, lde = [int(x) for x in k["coeffs"]], [int(x) for x in k["evals"]], [int(x) for x in k["lde"]]
        assert oracle.interpolate(evals) == coeffs
        assert oracle.lde(coeffs, k["blowup"]) == lde
        # interpolate_poly_with_offset inverts evaluation over the coset 3*<w_n> (rows 0, beta, 2*beta, ... of the LDE)
        assert oracle.interpolate(lde[::k["blowup"]], with_offset=True) == coeffs


def test_mimc_kats(oracle):
    k = load("mimc_kats")
    for c in k["cipher"]:
        assert oracle.mimc_cipher(int(c["x"]), int(c["rc"]), int(c["z"])) == int(c["out"])
    import zk_stark_project_b200 as Z
    w = [[Z.f64_to_felt(42.0)] * 9 for _ in range(6)]
    assert Z.mimc_hash_matrix(w, [Z.f64_to_felt(1.0)] * 6, Z.get_round_constants()) == int(k["bench_hash_matrix"])


def test_merkle_against_python_blake3(oracle):
    import blake3
    leaves = [blake3.blake3(bytes([i])).digest() for i in range(16)]
    level = leaves
    while len(level) > 1:
        level = [blake3.blake3(level[i] + level[i + 1]).digest() for i in range(0, len(level), 2)]
    assert oracle.merkle_root(leaves) == level[0]


def test_proof_pins(oracle):
    """Regression pin of the oracle's own transcript conventions (self-generated; see make_golden.py)."""
    import blake3
    import zk_stark_project_b200 as Z
    from tests import common as T
    pins = load("proof_pins")
    cases = {"mimc_w4_n64_b8": T.mimc_prover(4, 64, T.options(blowup=8)), "aggregation_16": T.aggregation_prover(16, T.options()),
             "training_bs1": T.training_prover(1, T.options()),
             "aggregation_16_reference_options": T.aggregation_prover(16, Z.ProofOptions.reference())}
    for name, prover in cases.items():
        tr = prover.build_trace()
        proof, ts, _ = oracle.prove(prover.describe(tr), tr.to_bytes())
        assert len(proof) == pins[name]["proof_len"]
        assert blake3.blake3(proof).hexdigest() == pins[name]["proof_blake3"], name
        assert int(ts.pow_nonce) == pins[name]["pow_nonce"]
