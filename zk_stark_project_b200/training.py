"""Training-step AIR and prover (src/training/air.rs, src/training/prover.rs)."""
import numpy as np

from . import field as F
from .field import P, f64_to_felt
from .prover import Prover
from .trace import TraceTable

AC, FE = 6, 9  # src/helper.rs:18-20


class TrainingUpdateInputs:
    """src/training/air.rs:18-35."""

    def __init__(self, initial_masked, final_masked, steps, x_batch, y_batch, learning_rate, precision, batch_size):
        self.initial_masked, self.final_masked, self.steps = list(initial_masked), list(final_masked), steps
        self.x_batch, self.y_batch = x_batch, y_batch
        self.learning_rate, self.precision, self.batch_size = learning_rate, precision, batch_size

    def to_elements(self):
        """src/training/air.rs:70-94 — note steps and batch_size go through f64_to_felt (x 1e6)."""
        v = list(self.initial_masked) + list(self.final_masked)
        v.append(f64_to_felt(float(self.steps)))
        v.append(f64_to_felt(float(self.batch_size)))
        for bx in self.x_batch:
            v.extend(bx)
        for by in self.y_batch:
            v.extend(by)
        v.append(self.learning_rate)
        v.append(self.precision)
        return [int(x) % P for x in v]


class TrainingUpdateAir:
    """src/training/air.rs:100-151: `width` degree-1 transition constraints (all identically zero because
    `current_step()` is the constant 0, src/helper.rs:141-146) and `width` single assertions."""

    AIR_ID = 1

    def __init__(self, trace_width, trace_len, pub_inputs, options):
        if len(pub_inputs.x_batch) != pub_inputs.batch_size:  # src/training/air.rs:118-121
            raise ValueError("x_batch size doesn't match batch_size in public inputs")
        if len(pub_inputs.y_batch) != pub_inputs.batch_size:
            raise ValueError("y_batch size doesn't match batch_size in public inputs")
        self.width, self.trace_len, self.pub_inputs, self.opts = trace_width, trace_len, pub_inputs, options

    def get_assertions(self):
        half, n = self.width // 2, self.trace_len - 1
        a = [(i, 0, self.pub_inputs.initial_masked[i]) for i in range(half)]
        a += [(i, n, self.pub_inputs.final_masked[i]) for i in range(half)]
        return a

    def describe(self):
        return dict(air_id=self.AIR_ID, trace_width=self.width, trace_len=self.trace_len, options=self.opts.describe(),
                    pub_elems=self.pub_inputs.to_elements(), assertions=self.get_assertions(), params=[])


# ---- src/helper.rs forward / backward pass in sign-encoded fixed point --------------------------------------------
def mse_prime(y_true, y_pred, y_pred_sign, pr):  # src/helper.rs:245-270
    ac_f = f64_to_felt(float(len(y_true)))
    res, sgn = [], []
    for i in range(len(y_true)):
        t, ts = F.subtract(y_pred[i], y_true[i], y_pred_sign[i], 0)
        t2, t2s = F.multiply(t, f64_to_felt(2.0), ts, 0)
        r, rs = F.divide(t2, ac_f, t2s, 0)
        res.append(r)
        sgn.append(rs)
    return res, sgn


def forward_propagation_layer(w, b, x, w_sign, b_sign, x_sign, pr):  # src/helper.rs:282-327
    out, out_s = [], []
    for j in range(len(b)):
        t, ts = 0, 0
        for i in range(len(x)):
            ti, tis = F.multiply(w[j][i], x[i], w_sign[j][i], x_sign[i])
            t, ts = F.add(t, ti, ts, tis)
        d, ds = F.divide(t, pr, ts, 0)
        r, rs = F.add(d, b[j], ds, b_sign[j])
        out.append(r)
        out_s.append(rs)
    return out, out_s


def backward_propagation_layer(w, b, x, err, lr, pr, w_sign, b_sign, x_sign, err_sign):  # src/helper.rs:345-401
    ac, fe = len(b), len(x)
    for i in range(ac):
        t, ts = F.divide(err[i], lr, err_sign[i], 0)
        b[i], b_sign[i] = F.subtract(b[i], t, b_sign[i], ts)
    for j in range(fe):
        for i in range(ac):
            prod, ps = F.multiply(err[i], x[j], err_sign[i], x_sign[j])
            t, ts = F.divide(prod, lr, ps, 0)
            g, gs = F.divide(t, pr, ts, 0)
            w[i][j], w_sign[i][j] = F.subtract(w[i][j], g, w_sign[i][j], gs)
    return w, b, w_sign, b_sign


def _add_mod_rows(raw, masks):
    """(raw[j] + masks[i, j]) mod p for a block of rows; raw: Python ints, masks: (rows, k) uint64 -> (rows, k, 2) uint64."""
    rows, k = masks.shape
    out = np.empty((rows, k, 2), dtype=np.uint64)
    lo_mask = (1 << 64) - 1
    for j in range(k):
        r = raw[j] % P
        rlo, rhi = np.uint64(r & lo_mask), r >> 64
        lo = masks[:, j] + rlo  # wraps mod 2^64
        carry = (lo < rlo).astype(np.uint64)
        if rhi < lo_mask - 1:
            # no 128-bit overflow and the sum stays below p (p's high word is 2^64 - 1)
            out[:, j, 0] = lo
            out[:, j, 1] = np.uint64(rhi) + carry
        else:
            for i in range(rows):  # rare: raw within 2^65 of the modulus — exact big-int path
                v = (r + int(masks[i, j])) % P
                out[i, j, 0] = v & lo_mask
                out[i, j, 1] = v >> 64
    return out


class TrainingUpdateProver(Prover):
    """src/training/prover.rs:18-301."""

    def __init__(self, options, initial_w, initial_b, w_sign, b_sign, x_batch, x_batch_sign, y_batch, learning_rate,
                 precision, batch_size, seed=None):
        # src/training/prover.rs:59-61 panics on mismatched batch sizes
        assert len(x_batch) == batch_size, "x_batch size doesn't match batch_size"
        assert len(x_batch_sign) == batch_size, "x_batch_sign size doesn't match batch_size"
        assert len(y_batch) == batch_size, "y_batch size doesn't match batch_size"
        self._options = options
        self.initial_w, self.initial_b, self.w_sign, self.b_sign = initial_w, initial_b, w_sign, b_sign
        self.x_batch, self.x_batch_sign, self.y_batch = x_batch, x_batch_sign, y_batch
        self.learning_rate, self.precision, self.batch_size = learning_rate, precision, batch_size
        ac, fe = len(initial_b), len(initial_w[0])
        state_cells = ac * fe + ac
        n = 2 * state_cells * batch_size  # src/training/prover.rs:63-65
        self.trace_length = max(1 << max(n - 1, 0).bit_length(), 16)
        # the reference draws masks from an unseeded thread_rng (src/training/prover.rs:117-121, 188-190);
        # a seed makes traces reproducible here
        self.rng = np.random.default_rng(seed)

    def options(self):
        return self._options

    def raw_states(self):
        """The distinct raw (unmasked) state rows of the trace: the initial state and one per processed sample
        (src/training/prover.rs:100-115,136-183); rows past the batch repeat the last one (:185)."""
        def flatten(w, ws, b, bs):
            raw = []
            for row, srow in zip(w, ws):
                for v, s in zip(row, srow):
                    raw += [v, s]
            for v, s in zip(b, bs):
                raw += [v, s]
            return raw

        w = [list(r) for r in self.initial_w]
        ws = [list(r) for r in self.w_sign]
        b, bs = list(self.initial_b), list(self.b_sign)
        states = [flatten(w, ws, b, bs)]
        for s in range(min(self.batch_size, self.trace_length - 1)):
            out, out_s = forward_propagation_layer(w, b, self.x_batch[s], ws, bs, self.x_batch_sign[s], self.precision)
            err, err_s = mse_prime(self.y_batch[s], out, out_s, self.precision)
            w, b, ws, bs = backward_propagation_layer(w, b, self.x_batch[s], err, self.learning_rate, self.precision, ws, bs,
                                                      self.x_batch_sign[s], err_s)
            states.append(flatten(w, ws, b, bs))
        return states

    def build_trace(self):
        """src/training/prover.rs:90-218: row = [raw + mask || mask], fresh 64-bit masks every row."""
        states = self.raw_states()
        flat_len, n = len(states[0]), self.trace_length
        masks = self.rng.integers(0, 1 << 64, size=(n, flat_len), dtype=np.uint64)
        data = np.empty((n, 2 * flat_len, 2), dtype=np.uint64)  # row-major first, transposed at the end
        data[:, flat_len:, 0] = masks
        data[:, flat_len:, 1] = 0
        for i, raw in enumerate(states[:-1]):
            data[i:i + 1, :flat_len] = _add_mod_rows(raw, masks[i:i + 1])
        last = len(states) - 1  # "If step > batch_size, just maintain the same state" (src/training/prover.rs:185)
        data[last:, :flat_len] = _add_mod_rows(states[-1], masks[last:])
        return TraceTable(np.ascontiguousarray(data.transpose(1, 0, 2)))

    def build_trace_device(self, key=None, ctx=None):
        """The same trace built on the GPU: only the distinct raw states cross PCIe; the 64-bit blinding masks are a
        ChaCha20 keystream generated on the device.  key=None (the default) keys it from OS entropy — the reference draws its
        masks from rand::thread_rng(), an OS-seeded CSPRNG (src/training/prover.rs:117-121), and the masks are all that hides the
        raw model state behind the public boundary rows.  Pass 32 key bytes only for reproducible tests."""
        from .trace import DeviceTrace
        ctx = ctx or self.context()
        states = self.raw_states()
        ptr, first, last = ctx.training_trace_device(states, self.trace_length, key)
        return DeviceTrace(ctx, ptr, 2 * len(states[0]), self.trace_length, first, last)

    def get_pub_inputs(self, trace):
        """src/training/prover.rs:235-267: the masked boundary rows are read back from the trace itself."""
        rows, half = trace.length(), trace.width() // 2
        return TrainingUpdateInputs([trace.get(c, 0) for c in range(half)], [trace.get(c, rows - 1) for c in range(half)],
                                    self.trace_length - 1, self.x_batch, self.y_batch, self.learning_rate, self.precision,
                                    self.batch_size)

    def new_air(self, trace, pub_inputs):
        return TrainingUpdateAir(trace.width(), trace.length(), pub_inputs, self._options)
