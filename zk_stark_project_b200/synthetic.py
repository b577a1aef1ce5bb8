"""Seeded synthetic inputs of the reference's shapes for benchmarks and parity tests.

The training prover is data-oblivious — every transition constraint is identically zero and the boundary values are
read back from the trace itself (src/training/prover.rs:245-246) — so uniform random 240-column traces
exercise exactly the same path as traces built by `TrainingUpdateProver.build_trace`."""
import numpy as np

from .field import f64_to_felt
from .options import ProofOptions
from .training import AC, FE, TrainingUpdateAir, TrainingUpdateInputs


def random_felts(count, seed, out=None):
    """`count` canonical field elements as a (count, 2) uint64 array (hi word < 2^63 keeps every value below p)."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty((count, 2), dtype=np.uint64)
    flat = out.reshape(-1)
    step = 1 << 24
    for o in range(0, flat.shape[0], step):
        k = min(step, flat.shape[0] - o)
        flat[o:o + k] = rng.integers(0, 1 << 64, size=k, dtype=np.uint64)
    out[:, 1] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    return out


def batch_size_for(n):
    """A batch size for which src/training/prover.rs:63-65 picks trace length n (n = next_pow2(120*bs).max(16))."""
    if n <= 16:
        return 1 if n == 16 else None
    bs = n // 120
    while max(1 << (120 * bs - 1).bit_length(), 16) != n:
        bs -= 1
        if bs < 1:
            return None
    return bs


def synthetic_training_air(n, opts: ProofOptions, data):
    """Training AIR description over an arbitrary (240, n, 2) trace; public batch per tests/integration_tests.rs:41-55."""
    w = data.shape[0]
    half = w // 2
    get = lambda c, r: int(data[c, r, 0]) | (int(data[c, r, 1]) << 64)
    bs = batch_size_for(n) or 1
    x = [[f64_to_felt((i + j) * 0.1) for j in range(FE)] for i in range(bs)]
    y = [[f64_to_felt(1.0) if a == i % AC else 0 for a in range(AC)] for i in range(bs)]
    pub = TrainingUpdateInputs([get(c, 0) for c in range(half)], [get(c, n - 1) for c in range(half)], n - 1, x, y,
                               f64_to_felt(0.01), f64_to_felt(1e6), bs)
    return TrainingUpdateAir(w, n, pub, opts).describe()
