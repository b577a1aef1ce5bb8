"""`winterfell::Prover` as the reference drives it: `prover.prove(trace) -> Proof` (src/main.rs:228,424,468).

The associated types the reference picks (src/training/prover.rs:225-233) are fixed here too: Blake3_256 hashing,
Merkle vector commitments, the default random coin — all executed by libzkb200.so on the GPU.
"""
import ctypes as C
import threading

import numpy as np

from . import lib as _lib
from .trace import TraceTable


class ProverError(RuntimeError):
    """`winterfell::ProverError` counterpart."""


class Proof:
    """`winterfell::Proof`: serialized form plus the Fiat-Shamir transcript it was produced with."""

    def __init__(self, data, transcript=None, stage_times=None):
        self._bytes = data
        self.transcript = transcript
        self.stage_times = stage_times

    def to_bytes(self):
        return self._bytes

    def __len__(self):
        return len(self._bytes)


_ctx_lock = threading.Lock()
_default_ctx = {}


def default_context(device=0):
    with _ctx_lock:
        if device not in _default_ctx:
            _default_ctx[device] = _lib.Context(device)
        return _default_ctx[device]


class Prover:
    """Base class: subclasses provide `options()`, `get_pub_inputs(trace)` and `air(trace_info, pub_inputs, options)`."""

    device = 0

    def options(self):
        raise NotImplementedError

    def get_pub_inputs(self, trace):
        raise NotImplementedError

    def new_air(self, trace, pub_inputs):
        raise NotImplementedError

    def context(self):
        return default_context(self.device)

    def describe(self, trace):
        pub = self.get_pub_inputs(trace)
        return self.new_air(trace, pub).describe()

    def prove(self, trace, force_nonce=0):
        """`Prover::prove(trace)`: host trace in, proof out (the e2e path; H2D copy included)."""
        from .trace import DeviceTrace
        if isinstance(trace, DeviceTrace):  # built on the GPU: no ingest at all
            air = self.describe(trace)
            try:
                proof, ts = trace.ctx.prove_device(air, trace.ptr, force_nonce)
            except _lib.ZkbError as e:
                raise ProverError(str(e)) from e
            return Proof(proof, ts, trace.ctx.stage_times())
        if not isinstance(trace, TraceTable):
            raise TypeError("trace must be a TraceTable or a DeviceTrace")
        air = self.describe(trace)
        ctx = self.context()
        data = np.ascontiguousarray(trace.data)
        try:
            proof, ts = ctx.prove_host(air, data.ctypes.data, force_nonce)
        except _lib.ZkbError as e:
            raise ProverError(str(e)) from e
        return Proof(proof, ts, ctx.stage_times())
