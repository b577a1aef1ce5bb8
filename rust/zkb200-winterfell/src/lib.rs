//! zkb200-winterfell — plugs the B200 backend into the reference's `winterfell::Prover` impls.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).  Written against the
//! Winterfell 0.12 surface recalled in SURVEY.md Appendix A.4/A.5; items marked VERIFY-FIRST depend on upstream
//! details that could not be read (Appendix D).
//!
//! Two integration levels:
//!
//! 1. `GpuProve::prove_gpu(&self, trace)` — the whole of `Prover::prove` on the device through `zkb_prove`
//!    (what bench.py measures).  The proof bytes come back in `Proof::to_bytes()` layout and are parsed with
//!    `Proof::from_bytes`, so `main.rs` keeps calling `verify::<Air, Blake3_256<Felt>, DefaultRandomCoin<_>,
//!    MerkleTree<_>>` unchanged (src/main.rs:251-257).
//! 2. `GpuTraceLde` / `GpuConstraintEvaluator` / `GpuConstraintCommitment` — the three associated types the
//!    reference names at src/training/prover.rs:228-233 and src/aggregation/prover.rs:201-206, for callers
//!    that keep Winterfell's own `generate_proof` (channel, DEEP and FRI stay on the CPU in that mode).
pub mod ffi;

use std::ffi::CStr;
use std::ptr;

use winterfell::{
    crypto::{hashers::Blake3_256, MerkleTree},
    math::{fields::f128::BaseElement as Felt, FieldElement, StarkField, ToElements},
    matrix::ColMatrix,
    Air, Assertion, ProofOptions, Proof, Prover, ProverError, Trace, TraceInfo, TraceTable,
};
use winter_utils::Serializable;

/// One `zkb_ctx`: a device + stream binding that owns all device memory of the proofs run through it.
pub struct GpuContext {
    raw: *mut ffi::zkb_ctx,
}
unsafe impl Send for GpuContext {}

impl GpuContext {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { ffi::zkb_ctx_create(device, ptr::null_mut(), &mut raw) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(ffi::zkb_last_error(ptr::null())) };
            return Err(format!("zkb_ctx_create failed ({rc}): {}", msg.to_string_lossy()));
        }
        Ok(Self { raw })
    }
    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(ffi::zkb_last_error(self.raw)) }.to_string_lossy().into_owned()
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ffi::zkb_ctx_destroy(self.raw) }
    }
}

/// Which of the three AIRs a prover drives, plus the AIR parameters the device function needs.
pub enum GpuAir {
    /// src/training/air.rs — all transition evaluations are zero (src/helper.rs:141-146).
    Training,
    /// src/aggregation/air.rs — `k` is `pub_inputs.k` (src/aggregation/air.rs:108).
    Aggregation { k: Felt },
    /// MiMC chains with the periodic round-constant column (src/helper.rs:404-406).
    Mimc { round_constants: Vec<Felt> },
}

fn felts_to_bytes(v: &[Felt]) -> Vec<u8> {
    // f128 is canonical: 16 little-endian bytes per element
    let mut out = Vec::with_capacity(v.len() * 16);
    for e in v {
        out.extend_from_slice(&e.as_int().to_le_bytes());
    }
    out
}

fn batching_code(m: winterfell::BatchingMethod) -> u32 {
    match m {
        winterfell::BatchingMethod::Linear => 0,
        winterfell::BatchingMethod::Algebraic => 1,
        #[allow(unreachable_patterns)]
        _ => 2,
    }
}

/// Extension trait: `prover.prove_gpu(&ctx, trace)` is a drop-in for `prover.prove(trace)`
/// (src/main.rs:228,424,468).
pub trait GpuProve: Prover<BaseField = Felt, Trace = TraceTable<Felt>>
where
    <Self::Air as Air>::PublicInputs: ToElements<Felt>,
{
    /// AIR id + parameters for the device-side constraint evaluator.
    fn gpu_air(&self, pub_inputs: &<Self::Air as Air>::PublicInputs) -> GpuAir;

    fn prove_gpu(&self, ctx: &GpuContext, trace: TraceTable<Felt>) -> Result<Proof, ProverError> {
        let pub_inputs = self.get_pub_inputs(&trace);
        let pub_elems = felts_to_bytes(&pub_inputs.to_elements());
        let air = Self::Air::new(trace.info().clone(), pub_inputs.clone(), self.options().clone());
        let assertions: Vec<Assertion<Felt>> = air.get_assertions();
        let cols: Vec<u32> = assertions.iter().map(|a| a.column() as u32).collect();
        let steps: Vec<u64> = assertions.iter().map(|a| a.first_step() as u64).collect();
        let values = felts_to_bytes(&assertions.iter().map(|a| a.values()[0]).collect::<Vec<_>>());
        let (air_id, params) = match self.gpu_air(&pub_inputs) {
            GpuAir::Training => (ffi::ZKB_AIR_ID_TRAINING, vec![]),
            GpuAir::Aggregation { k } => (ffi::ZKB_AIR_ID_AGGREGATION, felts_to_bytes(&[k])),
            GpuAir::Mimc { round_constants } => (ffi::ZKB_AIR_ID_MIMC, felts_to_bytes(&round_constants)),
        };
        let o: &ProofOptions = self.options();
        let desc = ffi::zkb_air_desc {
            air_id,
            trace_width: trace.main_trace_width() as u32,
            trace_len: trace.length() as u64,
            num_queries: o.num_queries() as u32,
            blowup: o.blowup_factor() as u32,
            grinding_bits: o.grinding_factor(),
            field_extension: o.field_extension() as u32,
            folding: o.to_fri_options().folding_factor() as u32,
            rem_max_degree: o.to_fri_options().remainder_max_degree() as u32,
            batching_constraints: batching_code(o.constraint_batching_method()), // VERIFY-FIRST: accessor names
            batching_deep: batching_code(o.deep_poly_batching_method()),
            pub_elems: pub_elems.as_ptr(),
            n_pub_elems: (pub_elems.len() / 16) as u64,
            assert_cols: cols.as_ptr(),
            assert_steps: steps.as_ptr(),
            assert_values: values.as_ptr(),
            n_assertions: assertions.len() as u64,
            params: params.as_ptr(),
            n_params: (params.len() / 16) as u64,
        };
        // TraceTable keeps one Vec<Felt> per column (src/helper.rs:197-211); Felt is repr(transparent) over u128
        let main: &ColMatrix<Felt> = trace.main_segment();
        let col_ptrs: Vec<*const u8> = (0..main.num_cols()).map(|j| main.get_column(j).as_ptr() as *const u8).collect();
        let (mut out, mut len) = (ptr::null_mut::<u8>(), 0u64);
        let rc = unsafe { ffi::zkb_prove(ctx.raw, &desc, col_ptrs.as_ptr(), 0, &mut out, &mut len, ptr::null_mut()) };
        if rc != 0 {
            // invalid construction panics in the reference (src/training/prover.rs:59-61); runtime failures map to ProverError
            panic!("zkb_prove failed ({rc}): {}", ctx.last_error());
        }
        let bytes = unsafe { std::slice::from_raw_parts(out, len as usize) }.to_vec();
        unsafe { ffi::zkb_free(out as *mut _) };
        Proof::from_bytes(&bytes).map_err(|e| ProverError::UnsupportedFieldExtension(e.to_string().len())) // VERIFY-FIRST: error mapping
    }
}

// ---- associated-type level -----------------------------------------------------------------------------------------
// Sketch of the three plug-in types for callers that keep Winterfell's `generate_proof`.  Each method is one C-ABI call;
// see INTEGRATION.md for the full table.  (TraceLde: Sync — reads go through zkb_trace_read_frame under a mutex, or,
// preferred, GpuConstraintEvaluator never calls read_main_trace_frame_into at all because it evaluates on the device.)
//
//   impl<E> TraceLde<E> for GpuTraceLde            get_main_trace_commitment  <- root kept from zkb_trace_commit
//                                                  read_main_trace_frame_into <- zkb_trace_read_frame
//                                                  query                      <- zkb_query(which = 0)
//   impl<'a, E> ConstraintEvaluator<E> for GpuConstraintEvaluator<'a, A>
//                                                  evaluate                   <- zkb_constraints_eval (returns the
//                                                                                CompositionPolyTrace when evals_out != NULL)
//   impl<E> ConstraintCommitment<E> for GpuConstraintCommitment
//                                                  commitment                 <- root kept from zkb_constraints_commit
//                                                  query                      <- zkb_query(which = 1)
//
//   fn new_trace_lde(..)               -> zkb_begin + zkb_trace_commit, TracePolyTable::new(ColMatrix from zkb_trace_polys_read)
//   fn new_evaluator(..)               -> GpuConstraintEvaluator { alpha = composition_coefficients.transition[1] }  // alpha^1
//   fn build_constraint_commitment(..) -> zkb_constraints_commit, CompositionPoly::new(trace, domain, num_cols) on the host
