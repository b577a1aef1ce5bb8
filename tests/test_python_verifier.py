"""CPU: the product-side verifier (zk_stark_project_b200.verifier, the `winterfell::verify` counterpart of the Python mirror)
against proofs from the oracle prover — two independently written implementations of the protocol must agree."""
import pytest

import zk_stark_project_b200 as Z
from tests import common as T

P = Z.P


def _proof(oracle, prover):
    trace = prover.build_trace()
    air = prover.describe(trace)
    return air, oracle.prove(air, trace.to_bytes())[0]


@pytest.mark.parametrize("make", [lambda: T.mimc_prover(4, 64, T.options(blowup=8)), lambda: T.mimc_prover(1, 256, T.options(blowup=16)),
                                  lambda: T.aggregation_prover(16, T.options()), lambda: T.aggregation_prover(3, Z.ProofOptions.reference()),
                                  lambda: T.training_prover(1, T.options()), lambda: T.training_prover(2, T.options(queries=17))])
def test_python_verifier_accepts_oracle_proofs(oracle, make):
    air, proof = _proof(oracle, make())
    assert Z.verify(proof, air)
    oracle.verify(air, proof)


def test_python_verifier_rejections(oracle):
    air, proof = _proof(oracle, T.aggregation_prover(6, T.options()))
    for off in (3, 30, 60, len(proof) // 3, len(proof) // 2, len(proof) - 40, len(proof) - 3):
        bad = bytearray(proof)
        bad[off] ^= 0x10
        with pytest.raises(Z.VerifierError):
            Z.verify(bytes(bad), air)
    with pytest.raises(Z.VerifierError):
        Z.verify(proof[:-1], air)
    with pytest.raises(Z.VerifierError, match="OOD"):
        Z.verify(proof, dict(air, assertions=[(c, s, (v + 1) % P) if c == 0 else (c, s, v) for c, s, v in air["assertions"]]))
    with pytest.raises(Z.VerifierError):
        Z.verify(proof, dict(air, pub_elems=air["pub_elems"][:-1] + [7]))
    with pytest.raises(Z.VerifierError, match="options"):
        Z.verify(proof, dict(air, options=dict(air["options"], grinding=air["options"]["grinding"] + 1)))


def test_blake3_host_matches_python_blake3():
    import blake3
    from zk_stark_project_b200 import lib
    for n in (0, 1, 40, 64, 65, 1024, 1025, 3840, 5000):
        data = bytes(i % 251 for i in range(n))
        assert lib.blake3_host(data) == blake3.blake3(data).digest()
