"""ctypes binding for the CPU oracle (oracle/liboracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs only; the product package never imports it.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P = 2**128 - 45 * 2**40 + 1


class AirDesc(C.Structure):
    _fields_ = [
        ("air_id", C.c_uint32), ("trace_width", C.c_uint32), ("trace_len", C.c_uint64),
        ("num_queries", C.c_uint32), ("blowup", C.c_uint32), ("grinding_bits", C.c_uint32),
        ("field_extension", C.c_uint32), ("folding", C.c_uint32), ("rem_max_degree", C.c_uint32),
        ("batching_constraints", C.c_uint32), ("batching_deep", C.c_uint32),
        ("pub_elems", C.c_char_p), ("n_pub_elems", C.c_uint64),
        ("assert_cols", C.POINTER(C.c_uint32)), ("assert_steps", C.POINTER(C.c_uint64)),
        ("assert_values", C.c_char_p), ("n_assertions", C.c_uint64),
        ("params", C.c_char_p), ("n_params", C.c_uint64),
    ]


class Transcript(C.Structure):
    _fields_ = [
        ("trace_root", C.c_uint8 * 32), ("constraint_root", C.c_uint8 * 32), ("remainder_commitment", C.c_uint8 * 32),
        ("constraint_alpha", C.c_uint8 * 16), ("z", C.c_uint8 * 16), ("deep_alpha", C.c_uint8 * 16),
        ("n_fri_layers", C.c_uint32), ("n_positions", C.c_uint32),
        ("fri_roots", (C.c_uint8 * 32) * 16), ("fri_alphas", (C.c_uint8 * 16) * 16),
        ("pow_nonce", C.c_uint64), ("positions", C.c_uint32 * 256),
        ("comp_degree_ok", C.c_int32), ("_pad", C.c_int32),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_prove.restype = C.c_int
        _LIB.orc_verify.restype = C.c_int
        _LIB.orc_get_threads.restype = C.c_int
    return _LIB


def fe_bytes(x):
    return int(x % P).to_bytes(16, "little")


def fe_int(b):
    return int.from_bytes(bytes(b), "little")


def set_threads(t):
    lib().orc_set_threads(int(t))


def make_desc(air):
    """`air` is the plain description dict produced by zk_stark_project_b200.air.*.describe()."""
    d = AirDesc()
    o = air["options"]
    d.air_id, d.trace_width, d.trace_len = air["air_id"], air["trace_width"], air["trace_len"]
    d.num_queries, d.blowup, d.grinding_bits = o["num_queries"], o["blowup"], o["grinding"]
    d.field_extension, d.folding, d.rem_max_degree = o.get("field_extension", 1), o["folding"], o["rem_max_degree"]
    d.batching_constraints, d.batching_deep = o.get("batching_constraints", 1), o.get("batching_deep", 1)
    keep = []
    pub = b"".join(fe_bytes(x) for x in air["pub_elems"])
    d.pub_elems, d.n_pub_elems = pub, len(air["pub_elems"])
    na = len(air["assertions"])
    cols = (C.c_uint32 * max(na, 1))(*[a[0] for a in air["assertions"]])
    steps = (C.c_uint64 * max(na, 1))(*[a[1] for a in air["assertions"]])
    vals = b"".join(fe_bytes(a[2]) for a in air["assertions"])
    d.assert_cols, d.assert_steps, d.assert_values, d.n_assertions = cols, steps, vals, na
    par = b"".join(fe_bytes(x) for x in air.get("params", []))
    d.params, d.n_params = par, len(air.get("params", []))
    keep += [pub, cols, steps, vals, par]
    d._keep = keep
    return d


def prove(air, trace_bytes, force_nonce=0):
    """trace_bytes: column-major [w][n] 16-byte LE elements.  Returns (proof bytes, Transcript, seconds)."""
    d = make_desc(air)
    out = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    ts = Transcript()
    secs = C.c_double()
    err = C.create_string_buffer(512)
    buf = (C.c_uint8 * len(trace_bytes)).from_buffer_copy(trace_bytes) if not isinstance(trace_bytes, C.Array) else trace_bytes
    rc = lib().orc_prove(C.byref(d), buf, C.c_uint64(force_nonce), C.byref(out), C.byref(n), C.byref(ts), C.byref(secs), err, C.c_uint64(512))
    if rc != 0:
        raise RuntimeError("oracle prove failed: " + err.value.decode())
    proof = bytes(C.cast(out, C.POINTER(C.c_uint8 * n.value)).contents)
    lib().orc_free(out)
    return proof, ts, secs.value


def comp_trace(air, trace_bytes, ce_blowup):
    """CompositionPolyTrace of DefaultConstraintEvaluator::evaluate (ce_blowup * n elements, natural ce-domain order), taken
    from inside a full oracle proof.  Returns (evaluations, proof bytes, Transcript)."""
    out = C.create_string_buffer(16 * air["trace_len"] * ce_blowup)
    lib().orc_dump_comp_trace(out)
    try:
        proof, ts, _ = prove(air, trace_bytes)
    finally:
        lib().orc_dump_comp_trace(None)
    return out.raw, proof, ts


def verify(air, proof):
    d = make_desc(air)
    ts = Transcript()
    err = C.create_string_buffer(512)
    rc = lib().orc_verify(C.byref(d), proof, C.c_uint64(len(proof)), C.byref(ts), err, C.c_uint64(512))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return ts


def blake3(data):
    out = C.create_string_buffer(32)
    lib().orc_blake3(data, C.c_uint64(len(data)), out)
    return out.raw


def fe_op(name, *args):
    out = C.create_string_buffer(16)
    getattr(lib(), "orc_fe_" + name)(*[fe_bytes(a) for a in args], out)
    return fe_int(out.raw)


def interpolate(evals, with_offset=False):
    buf = C.create_string_buffer(b"".join(fe_bytes(x) for x in evals), 16 * len(evals))
    (lib().orc_interpolate_with_offset if with_offset else lib().orc_interpolate)(buf, C.c_uint64(len(evals)))
    return [fe_int(buf.raw[16 * i:16 * i + 16]) for i in range(len(evals))]


def lde(coeffs, blowup):
    n = len(coeffs)
    src = b"".join(fe_bytes(x) for x in coeffs)
    out = C.create_string_buffer(16 * n * blowup)
    lib().orc_lde(src, C.c_uint64(n), C.c_uint64(blowup), out)
    return [fe_int(out.raw[16 * i:16 * i + 16]) for i in range(n * blowup)]


def merkle_root(leaves):
    out = C.create_string_buffer(32)
    lib().orc_merkle_root(b"".join(leaves), C.c_uint64(len(leaves)), out)
    return out.raw


def trace_commit(trace_bytes, n, w, blowup, want_lde=False, want_polys=False):
    root = C.create_string_buffer(32)
    lde_buf = C.create_string_buffer(16 * n * blowup * w) if want_lde else None
    polys = C.create_string_buffer(16 * n * w) if want_polys else None
    lib().orc_trace_commit(trace_bytes, C.c_uint64(n), C.c_uint64(w), C.c_uint64(blowup), root, lde_buf, polys)
    return root.raw, (lde_buf.raw if want_lde else None), (polys.raw if want_polys else None)


def mimc_trace(seeds, n, rc):
    w = len(seeds)
    out = C.create_string_buffer(16 * n * w)
    lib().orc_mimc_trace(b"".join(fe_bytes(s) for s in seeds), C.c_uint64(w), C.c_uint64(n),
                         b"".join(fe_bytes(x) for x in rc), C.c_uint64(len(rc)), out)
    return out.raw


def mimc_cipher(x, rc, z):
    out = C.create_string_buffer(16)
    lib().orc_mimc_cipher(fe_bytes(x), fe_bytes(rc), fe_bytes(z), out)
    return fe_int(out.raw)
