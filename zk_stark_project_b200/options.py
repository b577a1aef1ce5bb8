"""`winterfell::ProofOptions` as the reference constructs it (src/main.rs:98-107)."""
from dataclasses import dataclass


class FieldExtension:
    NONE = 1
    QUADRATIC = 2
    CUBIC = 3


class BatchingMethod:
    LINEAR = 0
    ALGEBRAIC = 1


@dataclass(frozen=True)
class ProofOptions:
    """Positional order of `ProofOptions::new` in Winterfell 0.12 (the inline comments at src/main.rs:99-104 are wrong)."""
    num_queries: int
    blowup_factor: int
    grinding_factor: int
    field_extension: int
    fri_folding_factor: int
    fri_remainder_max_degree: int
    batching_constraints: int = BatchingMethod.ALGEBRAIC
    batching_deep: int = BatchingMethod.ALGEBRAIC

    def __post_init__(self):
        # the same panics ProofOptions::new raises
        if not 0 < self.num_queries <= 255:
            raise ValueError("number of queries must be in 1..=255")
        b = self.blowup_factor
        if b < 2 or b > 128 or b & (b - 1):
            raise ValueError("blowup factor must be a power of two in 2..=128")
        if self.grinding_factor > 32:
            raise ValueError("grinding factor cannot be greater than 32")
        if self.fri_folding_factor not in (2, 4, 8, 16):
            raise ValueError("FRI folding factor must be 2, 4, 8 or 16")
        d = self.fri_remainder_max_degree
        if d > 255 or (d + 1) & d:
            raise ValueError("FRI remainder max degree must be one less than a power of two")

    @classmethod
    def reference(cls):
        """The options hard-coded at src/main.rs:98-107 and tests/integration_tests.rs:69-75."""
        return cls(40, 16, 21, FieldExtension.NONE, 16, 7, BatchingMethod.ALGEBRAIC, BatchingMethod.ALGEBRAIC)

    def describe(self):
        return dict(num_queries=self.num_queries, blowup=self.blowup_factor, grinding=self.grinding_factor,
                    field_extension=self.field_extension, folding=self.fri_folding_factor,
                    rem_max_degree=self.fri_remainder_max_degree, batching_constraints=self.batching_constraints,
                    batching_deep=self.batching_deep)
