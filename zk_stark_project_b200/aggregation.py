"""FedAvg aggregation AIR and prover (src/aggregation/air.rs, src/aggregation/prover.rs)."""
import numpy as np

from .field import P, inv
from .mimc import get_round_constants, mimc_hash_matrix
from .prover import Prover
from .trace import TraceTable
from .training import AC, FE


class GlobalUpdateInputs:
    """src/aggregation/air.rs:13-31."""

    def __init__(self, global_w, global_b, new_global_w, new_global_b, k, digest, steps):
        self.global_w, self.global_b, self.new_global_w, self.new_global_b = global_w, global_b, new_global_w, new_global_b
        self.k, self.digest, self.steps = k, digest, steps

    def to_elements(self):
        """src/aggregation/air.rs:57-81 — `steps` is a plain integer here (not scaled by 1e6)."""
        e = []
        for row in self.global_w:
            e.extend(row)
        e.extend(self.global_b)
        for row in self.new_global_w:
            e.extend(row)
        e.extend(self.new_global_b)
        e += [self.k, self.digest, self.steps]
        return [int(x) % P for x in e]


class GlobalUpdateAir:
    """src/aggregation/air.rs:84-151: d = AC*FE+AC degree-1 constraints k*(next-cur) - update = 0; 2d assertions at row steps-1."""

    AIR_ID = 2

    def __init__(self, trace_width, trace_len, pub_inputs, options):
        self.width, self.trace_len, self.pub_inputs, self.opts = trace_width, trace_len, pub_inputs, options

    def get_assertions(self):
        d = AC * FE + AC
        final = [v for row in self.pub_inputs.new_global_w for v in row] + list(self.pub_inputs.new_global_b)
        last_row = self.pub_inputs.steps - 1
        return [(i, last_row, final[i]) for i in range(d)] + [(i, last_row, 0) for i in range(d, 2 * d)]

    def describe(self):
        return dict(air_id=self.AIR_ID, trace_width=self.width, trace_len=self.trace_len, options=self.opts.describe(),
                    pub_elems=self.pub_inputs.to_elements(), assertions=self.get_assertions(), params=[self.pub_inputs.k])


class GlobalUpdateProver(Prover):
    """src/aggregation/prover.rs:15-249."""

    def __init__(self, options, raw_global_w, raw_global_b, local_w, local_b, k, seed=None):
        self._options = options
        self.raw_global_w, self.raw_global_b, self.local_w, self.local_b, self.k = raw_global_w, raw_global_b, local_w, local_b, k
        steps = len(local_w) + 2
        self.trace_length = max(1 << (steps - 1).bit_length(), 8)  # src/aggregation/prover.rs:63-64
        d = AC * FE + AC
        # unseeded thread_rng in the reference (src/aggregation/prover.rs:68-72); seedable here
        rng = np.random.default_rng(seed)
        self.blinding = [int(x) for x in rng.integers(0, 1 << 64, size=d, dtype=np.uint64)]
        raw_flat = self.flatten_state(raw_global_w, raw_global_b)
        masked = [(r + m) % P for r, m in zip(raw_flat, self.blinding)]
        self.masked_global_w, self.masked_global_b = self.unflatten_state(masked, AC, FE)

    @staticmethod
    def flatten_state(w, b):
        return [v for row in w for v in row] + list(b)

    @staticmethod
    def unflatten_state(state, ac, fe):
        return [list(state[i * fe:(i + 1) * fe]) for i in range(ac)], list(state[ac * fe:])

    def options(self):
        return self._options

    def compute_iterative_trace_augmented(self):
        """src/aggregation/prover.rs:98-154."""
        d = AC * FE + AC
        rows = []
        cur = self.flatten_state(self.masked_global_w, self.masked_global_b)
        rows.append(cur + [0] * d)
        raw_flat = self.flatten_state(self.raw_global_w, self.raw_global_b)
        kinv = inv(self.k)
        for lw, lb in zip(self.local_w, self.local_b):
            l = [v for row in lw for v in row] + list(lb)
            upd = [(li - g0) % P for g0, li in zip(raw_flat, l)]
            cur = [(c + u * kinv) % P for c, u in zip(cur, upd)]
            rows.append(cur + upd)
        rows.append(cur + [0] * d)
        while len(rows) < self.trace_length:
            rows.append(list(rows[-1]))
        return rows

    def build_trace(self):
        return TraceTable.from_rows(self.compute_iterative_trace_augmented())

    def get_pub_inputs(self, trace=None):
        """src/aggregation/prover.rs:163-190."""
        steps = len(self.local_w) + 2
        d = AC * FE + AC
        final = self.compute_iterative_trace_augmented()[steps - 1][:d]
        new_w, new_b = self.unflatten_state(final, AC, FE)
        digest = mimc_hash_matrix(new_w, new_b, get_round_constants())
        return GlobalUpdateInputs(self.masked_global_w, self.masked_global_b, new_w, new_b, self.k, digest, steps)

    def new_air(self, trace, pub_inputs):
        return GlobalUpdateAir(trace.width(), trace.length(), pub_inputs, self._options)
