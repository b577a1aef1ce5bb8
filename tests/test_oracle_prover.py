"""CPU: prove -> verify closure of the oracle for the three AIRs, and the failure modes the reference's tests care about."""
import numpy as np
import pytest

import zk_stark_project_b200 as Z
from tests import common as T

P = Z.P


def _roundtrip(oracle, prover, trace=None):
    trace = trace or prover.build_trace()
    air = prover.describe(trace)
    proof, ts, _ = oracle.prove(air, trace.to_bytes())
    assert ts.comp_degree_ok == 1
    ts2 = oracle.verify(air, proof)
    assert bytes(ts2.z) == bytes(ts.z) and list(ts2.positions) == list(ts.positions)
    return air, proof, ts


@pytest.mark.parametrize("bs", [1, 2])
def test_training_proof_generation_and_verification(oracle, bs):
    """tests/integration_tests.rs:60-159 (bs in {1, 2}): build trace -> prove -> verify."""
    p = T.training_prover(bs, T.options())
    trace = p.build_trace()
    assert trace.width() == 240 and trace.length() == max(1 << (120 * bs - 1).bit_length(), 16)
    air, proof, ts = _roundtrip(oracle, p, trace)
    assert ts.n_fri_layers == (1 if bs == 1 else 2)  # 2^11 -> 2^7 (128 <= (7+1)*16 stops); 2^12 -> 2^8 -> 2^4


@pytest.mark.parametrize("updates", [1, 6, 16])
def test_aggregation_roundtrip(oracle, updates):
    p = T.aggregation_prover(updates, T.options())
    trace = p.build_trace()
    assert trace.width() == 120 and trace.length() == max(1 << (updates + 1).bit_length(), 8)
    _roundtrip(oracle, p, trace)


@pytest.mark.parametrize("width,steps,blowup", [(1, 64, 8), (4, 128, 8), (2, 256, 16)])
def test_mimc_roundtrip(oracle, width, steps, blowup):
    _roundtrip(oracle, T.mimc_prover(width, steps, T.options(blowup=blowup)))


def test_fri_layer_counts(oracle):
    """2^20 -> 4 layers (remainder 1 coeff), 512 -> 1 layer, 2^17 blowup 8 -> 3 layers."""
    p = T.aggregation_prover(16, T.options())
    _, _, ts = _roundtrip(oracle, p)
    assert ts.n_fri_layers == 1
    _, _, ts = _roundtrip(oracle, T.mimc_prover(1, 1 << 14, T.options(blowup=8)),
                          Z.TraceTable(np.frombuffer(oracle.mimc_trace([1], 1 << 14, Z.get_round_constants()), dtype=np.uint64).reshape(1, 1 << 14, 2)))
    assert ts.n_fri_layers == 3


def test_tampered_proof_rejected(oracle):
    air, proof, _ = _roundtrip(oracle, T.aggregation_prover(6, T.options()))
    for off in (40, len(proof) // 3, len(proof) // 2, len(proof) - 20):
        bad = bytearray(proof)
        bad[off] ^= 1
        with pytest.raises(RuntimeError):
            oracle.verify(air, bytes(bad))


def test_wrong_public_inputs_rejected(oracle):
    p = T.mimc_prover(2, 64, T.options(blowup=8))
    air, proof, _ = _roundtrip(oracle, p)
    bad = dict(air, assertions=[(c, s, (v + 1) % P) if i == 0 else (c, s, v) for i, (c, s, v) in enumerate(air["assertions"])])
    with pytest.raises(RuntimeError, match="OOD"):
        oracle.verify(bad, proof)
    bad = dict(air, pub_elems=[(x + 1) % P for x in air["pub_elems"]])  # different coin seed
    with pytest.raises(RuntimeError):
        oracle.verify(bad, proof)


def test_invalid_trace_does_not_verify(oracle):
    """A trace violating the transition constraints yields a proof the verifier rejects (release Winterfell does not check the trace)."""
    p = T.aggregation_prover(6, T.options())
    trace = p.build_trace()
    air = p.describe(trace)
    data = trace.data.copy()
    data[3, 2, 0] ^= np.uint64(1)  # break k*(next - cur) = update at one cell
    proof, ts, _ = oracle.prove(air, Z.TraceTable(data).to_bytes())
    assert ts.comp_degree_ok == 0
    with pytest.raises(RuntimeError):
        oracle.verify(air, proof)


def test_wrong_options_rejected(oracle):
    p = T.mimc_prover(2, 64, T.options(blowup=8))
    air, proof, _ = _roundtrip(oracle, p)
    other = dict(air, options=dict(air["options"], num_queries=41))
    with pytest.raises(RuntimeError):
        oracle.verify(other, proof)


def test_forced_nonce(oracle):
    """parity against a `concurrent` reference run needs the nonce forced; any valid nonce verifies."""
    p = T.mimc_prover(1, 64, T.options(blowup=8, grinding=4))
    trace = p.build_trace()
    air = p.describe(trace)
    _, ts, _ = oracle.prove(air, trace.to_bytes())
    # find a second valid nonce by brute force through the verifier
    proof2 = None
    for nonce in range(int(ts.pow_nonce) + 1, int(ts.pow_nonce) + 400):
        cand, _, _ = oracle.prove(air, trace.to_bytes(), force_nonce=nonce)
        try:
            oracle.verify(air, cand)
            proof2 = cand
            break
        except RuntimeError:
            continue
    assert proof2 is not None
