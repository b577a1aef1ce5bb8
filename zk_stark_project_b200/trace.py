"""`winterfell::TraceTable<Felt>`: column-major main trace segment (src/training/prover.rs:213, src/aggregation/prover.rs:159)."""
import numpy as np

from .field import P


class TraceTable:
    """Columns of 16-byte little-endian field elements, stored as a (width, length, 2) uint64 array (lo, hi)."""

    def __init__(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        if data.ndim != 3 or data.shape[2] != 2:
            raise ValueError("trace data must have shape (width, length, 2)")
        w, n, _ = data.shape
        # TraceTable::init / TraceInfo::new panics
        if not 0 < w <= 255:
            raise ValueError("number of trace columns must be in 1..=255")
        if n < 8 or n & (n - 1):
            raise ValueError("trace length must be a power of two and at least 8")
        self.data = data

    @classmethod
    def init(cls, columns):
        """`TraceTable::init(Vec<Vec<Felt>>)` — columns of Python ints."""
        w = len(columns)
        n = len(columns[0]) if w else 0
        arr = np.empty((w, n, 2), dtype=np.uint64)
        for j, col in enumerate(columns):
            if len(col) != n:
                raise ValueError("all columns must have the same length")
            for i, v in enumerate(col):
                v = int(v) % P
                arr[j, i, 0] = v & 0xFFFFFFFFFFFFFFFF
                arr[j, i, 1] = v >> 64
        return cls(arr)

    @classmethod
    def from_rows(cls, rows):
        """`TraceTable::init(transpose(rows))` (src/helper.rs:197-211)."""
        return cls.init([list(c) for c in zip(*rows)])

    def width(self):
        return self.data.shape[0]

    def length(self):
        return self.data.shape[1]

    def get(self, col, row):
        lo, hi = self.data[col, row]
        return int(lo) | (int(hi) << 64)

    def column(self, col):
        return [self.get(col, i) for i in range(self.length())]

    def to_bytes(self):
        return self.data.tobytes()


class DeviceTrace:
    """A main trace segment that lives in GPU memory (column-major [width][length]), produced by a device-side trace builder.
    Only the boundary rows, which `get_pub_inputs` reads back (src/training/prover.rs:245-246), are mirrored on the host."""

    def __init__(self, ctx, ptr, width, length, first_row, last_row):
        self.ctx, self.ptr, self._w, self._n = ctx, ptr, width, length
        self._rows = {0: list(first_row), length - 1: list(last_row)}

    def width(self):
        return self._w

    def length(self):
        return self._n

    def get(self, col, row):
        if row not in self._rows:
            raise IndexError("only the first and last rows of a device trace are mirrored on the host")
        return self._rows[row][col]

    def to_host(self):
        """Copy the whole trace back (tests / debugging)."""
        import numpy as np
        raw = self.ctx.download(self.ptr, self._w * self._n * 16)
        return TraceTable(np.frombuffer(raw, dtype=np.uint64).reshape(self._w, self._n, 2).copy())
