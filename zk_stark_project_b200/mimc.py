"""MiMC helpers of the reference (src/helper.rs:213-233,404-406; benches/bench_mimc.rs) and the MiMC hash-chain AIR
this build defines on top of them: the reference has the round function but no MiMC AIR."""
from .field import P, f64_to_felt
from .prover import Prover
from .trace import TraceTable

MIMC_CYCLE = 64


def get_round_constants():
    """src/helper.rs:404-406: f64_to_felt(i) = i*10^6 for i = 1..=64."""
    return [f64_to_felt(float(i)) for i in range(1, 65)]


def mimc_cipher(inp, round_constant, z):
    """src/helper.rs:213-220 — 64 rounds of (x + rc + z)^7 with the same rc, then + z."""
    for _ in range(64):
        inp = pow((inp + round_constant + z) % P, 7, P)
    return (inp + z) % P


def mimc_hash_matrix(w, b, round_constants):
    """src/helper.rs:222-233."""
    z = f64_to_felt(0.0)
    for i in range(len(w)):
        for j in range(len(w[i])):
            z = mimc_cipher(w[i][j], round_constants[j % len(round_constants)], z)
        z = mimc_cipher(b[i], round_constants[i % len(round_constants)], z)
    return z


class MimcInputs:
    """Public inputs of the chain proof: per-column seed and result."""

    def __init__(self, seeds, results):
        self.seeds, self.results = list(seeds), list(results)

    def to_elements(self):
        return [int(x) % P for x in self.seeds + self.results]


class MimcAir:
    """W independent chains: next_j = (cur_j + rc[i mod 64])^7 with rc a periodic column.

    `TransitionConstraintDegree::new(7)` per column (the periodic polynomial's degree stays below the trace
    polynomials', so the constraint degree is exactly 7(n-1)): ce blowup 8, 6 composition columns.
    Assertions: col_j[0] = seed_j, col_j[n-1] = result_j.
    """

    AIR_ID = 3

    def __init__(self, trace_width, trace_len, pub_inputs, options, round_constants=None):
        self.width, self.trace_len, self.pub_inputs, self.opts = trace_width, trace_len, pub_inputs, options
        self.rc = list(round_constants) if round_constants is not None else get_round_constants()

    def get_assertions(self):
        n = self.trace_len - 1
        return [(j, 0, self.pub_inputs.seeds[j]) for j in range(self.width)] + \
               [(j, n, self.pub_inputs.results[j]) for j in range(self.width)]

    def describe(self):
        return dict(air_id=self.AIR_ID, trace_width=self.width, trace_len=self.trace_len, options=self.opts.describe(),
                    pub_elems=self.pub_inputs.to_elements(), assertions=self.get_assertions(), params=self.rc)


class MimcProver(Prover):
    def __init__(self, options, seeds, num_steps, round_constants=None):
        if num_steps < 8 or num_steps & (num_steps - 1):
            raise ValueError("number of steps must be a power of two >= 8")
        self._options, self.seeds, self.num_steps = options, [int(s) % P for s in seeds], num_steps
        self.rc = list(round_constants) if round_constants is not None else get_round_constants()

    def options(self):
        return self._options

    def build_trace(self):
        """Host-side chain (small traces); large traces are generated on the device with Context.mimc_trace."""
        cols = []
        for s in self.seeds:
            col, x = [], s
            for i in range(self.num_steps):
                col.append(x)
                x = pow((x + self.rc[i % len(self.rc)]) % P, 7, P)
            cols.append(col)
        return TraceTable.init(cols)

    def get_pub_inputs(self, trace):
        n = trace.length()
        return MimcInputs([trace.get(j, 0) for j in range(trace.width())], [trace.get(j, n - 1) for j in range(trace.width())])

    def new_air(self, trace, pub_inputs):
        return MimcAir(trace.width(), trace.length(), pub_inputs, self._options, self.rc)
