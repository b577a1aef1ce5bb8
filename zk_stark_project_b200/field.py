"""Scalar f128 helpers for the host side (trace builders, public inputs).

Mirrors winter-math `f128::BaseElement` as the reference uses it (src/helper.rs:11,24-27) and the
sign-encoded fixed-point arithmetic of src/signed.rs.  Bulk arithmetic never runs here.
"""
import math

P = 2**128 - 45 * 2**40 + 1
U128_MAX = 2**128 - 1


def Felt(x):
    """`Felt::new(x)` for x < 2^128: one conditional subtraction of p (src/signed.rs:3 relies on this)."""
    x = int(x)
    if not 0 <= x <= U128_MAX:
        raise ValueError("Felt::new takes a u128")
    return x - P if x >= P else x


def f64_to_felt(x):
    """src/helper.rs:25-27: scale by 1e6, round, reinterpret as u128 (negative values saturate to 0 in Rust)."""
    return Felt(min(max(_rust_round(float(x) * 1e6), 0), U128_MAX))


def _rust_round(v):
    """f64::round: half away from zero (Python's round() is half-to-even)."""
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def inv(a):
    return pow(a, P - 2, P) if a % P else 0


# ---- src/signed.rs ------------------------------------------------------------------------------------------------
MAX = Felt(U128_MAX)  # src/signed.rs:3 — as a field element this is 45*2^40 - 2
THRESHOLD = Felt(170141183460469231731687303715884105727)


def cleanse(v, s):  # src/signed.rs:11-15
    return ((1 - s) * v + s * (MAX - v + 1)) % P


def add_generic(a, s_a, b, s_b):  # src/signed.rs:17-26
    a_c, b_c = cleanse(a, s_a), cleanse(b, s_b)
    ind = s_a * s_b % P
    c = (ind * (MAX + 1 - a_c - b_c) + (1 - ind) * (a + b)) % P
    return c, ind


def sub_generic(a, s_a, b, s_b):  # src/signed.rs:28-31
    return add_generic(a, s_a, b, (1 - s_b) % P)


def mul_generic(a, s_a, b, s_b):  # src/signed.rs:33-40
    prod = cleanse(a, s_a) * cleanse(b, s_b) % P
    sign = (s_a + s_b - s_a * s_b * 2) % P
    res = (sign * (MAX - prod + 1) + (1 - sign) * prod) % P
    return res, sign


def div_generic(a, s_a, b, s_b):  # src/signed.rs:42-48
    q = cleanse(a, s_a) * inv(cleanse(b, s_b)) % P
    sign = (s_a + s_b - s_a * s_b * 2) % P
    res = (sign * (MAX + 1 - q) + (1 - sign) * q) % P
    return res, sign


# helper.rs re-exports them with (a, b, s_a, s_b) argument order (src/helper.rs:3, src/signed.rs:53-64)
def add(a, b, s_a, s_b):
    return add_generic(a, s_a, b, s_b)


def subtract(a, b, s_a, s_b):
    return sub_generic(a, s_a, b, s_b)


def multiply(a, b, s_a, s_b):
    return mul_generic(a, s_a, b, s_b)


def divide(a, b, s_a, s_b):
    return div_generic(a, s_a, b, s_b)


def encode_signed(x):  # src/helper.rs:39-46
    x = int(x)
    if x >= 0:
        return Felt(x), 0
    return Felt((U128_MAX - (-x) + 1) & U128_MAX), 1


def f64_to_signed_felt(x, scale):  # src/helper.rs:49-52
    return encode_signed(_rust_round(float(x) * scale))
