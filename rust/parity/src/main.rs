//! Byte-parity harness: `winterfell::Prover::prove` (the reference's stock CPU path) against `zkb_prove` on the SAME trace.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).  Needs the two `impl GpuProve for ...` blocks of
//! INTEGRATION.md §2 in the reference checkout (the orphan rule keeps them out of this crate).  For every case it
//!   1. builds the reference prover exactly as tests/integration_tests.rs:14-56 does (deterministic inputs; the trace itself
//!      contains `thread_rng` masks, src/training/prover.rs:119-121, so the trace is built ONCE and handed to both provers),
//!   2. proves with Winterfell and with the B200 library,
//!   3. compares `Proof::to_bytes()`; on a mismatch prints the first differing offset,
//!   4. writes `winterfell_<case>.json` — trace columns, public-input elements, assertions, options and the Winterfell proof —
//!      which `tests/test_winterfell_fixtures.py` of the B200 repository replays against its oracle and its CUDA path.  Committing
//!      those files under tests/golden/ is what turns "parity unpinned" into "pinned".
use std::{env, fs, path::PathBuf};

use winter_utils::Serializable;
use winterfell::{
    math::{fields::f128::BaseElement as Felt, FieldElement, StarkField, ToElements},
    Air, BatchingMethod, FieldExtension, ProofOptions, Prover, Trace, TraceTable,
};
use zk_stark_project::{
    aggregation::prover::GlobalUpdateProver,
    helper::{f64_to_felt, label_to_one_hot, AC, FE},
    training::prover::TrainingUpdateProver,
};
use zkb200_winterfell::{GpuContext, GpuProve};

fn options() -> ProofOptions {
    // src/main.rs:98-107
    ProofOptions::new(40, 16, 21, FieldExtension::None, 16, 7, BatchingMethod::Algebraic, BatchingMethod::Algebraic)
}

fn hex_felt(e: Felt) -> String {
    hex::encode(e.as_int().to_le_bytes())
}

/// Everything the Python side needs to replay the case (see tests/test_winterfell_fixtures.py for the schema).
fn write_fixture<A: Air<BaseField = Felt>>(
    out: &PathBuf, name: &str, air_id: u32, trace: &TraceTable<Felt>, air: &A, pub_elems: Vec<Felt>, params: Vec<Felt>, proof: &[u8],
    gpu_proof: &[u8],
) {
    let mut cols = Vec::new();
    for c in 0..trace.width() {
        let mut bytes = Vec::with_capacity(trace.length() * 16);
        for r in 0..trace.length() {
            bytes.extend_from_slice(&trace.get(c, r).as_int().to_le_bytes());
        }
        cols.push(format!("\"{}\"", hex::encode(bytes)));
    }
    let assertions: Vec<String> = air
        .get_assertions()
        .iter()
        .map(|a| format!("[{}, {}, \"{}\"]", a.column(), a.first_step(), hex_felt(a.values()[0])))
        .collect();
    let o = air.options();
    let json = format!(
        "{{\n \"generator\": \"rust/parity (winterfell 0.12, non-concurrent)\",\n \"air_id\": {air_id},\n \"trace_width\": {},\n \
         \"trace_len\": {},\n \"options\": {{\"num_queries\": {}, \"blowup\": {}, \"grinding\": {}, \"field_extension\": 1, \
         \"folding\": {}, \"rem_max_degree\": {}, \"batching_constraints\": 1, \"batching_deep\": 1}},\n \"pub_elems\": [{}],\n \
         \"assertions\": [{}],\n \"params\": [{}],\n \"columns\": [{}],\n \"winterfell_proof\": \"{}\",\n \"identical\": {}\n}}\n",
        trace.width(),
        trace.length(),
        o.num_queries(),
        o.blowup_factor(),
        o.grinding_factor(),
        o.to_fri_options().folding_factor(),
        o.to_fri_options().remainder_max_degree(),
        pub_elems.iter().map(|e| format!("\"{}\"", hex_felt(*e))).collect::<Vec<_>>().join(", "),
        assertions.join(", "),
        params.iter().map(|e| format!("\"{}\"", hex_felt(*e))).collect::<Vec<_>>().join(", "),
        cols.join(", "),
        hex::encode(proof),
        proof == gpu_proof,
    );
    fs::write(out.join(format!("winterfell_{name}.json")), json).expect("cannot write fixture");
}

fn report(name: &str, cpu: &[u8], gpu: &[u8]) -> bool {
    if cpu == gpu {
        println!("{name}: IDENTICAL ({} bytes)", cpu.len());
        return true;
    }
    let first = cpu.iter().zip(gpu.iter()).position(|(a, b)| a != b).unwrap_or(cpu.len().min(gpu.len()));
    println!("{name}: DIFFERENT — winterfell {} bytes, zkb200 {} bytes, first difference at offset {first}", cpu.len(), gpu.len());
    println!("  (offsets: context and commitments come first; see oracle/proof.h for the layout this build assumes)");
    false
}

fn training_case(gpu: &GpuContext, out: &PathBuf, bs: usize) -> bool {
    // deterministic model instead of generate_initial_model's thread_rng (src/helper.rs:108-131)
    let w: Vec<Vec<Felt>> = (0..AC).map(|j| (0..FE).map(|i| f64_to_felt(0.01 * (1 + i + j * FE) as f64)).collect()).collect();
    let b: Vec<Felt> = (0..AC).map(|j| f64_to_felt(0.1 * (j + 1) as f64)).collect();
    let w_sign = vec![vec![Felt::ZERO; FE]; AC];
    let b_sign = vec![Felt::ZERO; AC];
    let mut x_batch = Vec::new();
    let mut x_sign = Vec::new();
    let mut y_batch = Vec::new();
    for i in 0..bs {
        x_batch.push((0..FE).map(|j| f64_to_felt((i as f64 + j as f64) * 0.1)).collect::<Vec<_>>()); // tests/integration_tests.rs:41-43
        x_sign.push(vec![Felt::ZERO; FE]);
        y_batch.push(label_to_one_hot((i % AC) as f64 + 1.0, AC, 1e6).0);
    }
    let prover = TrainingUpdateProver::new(
        options(), w, b, w_sign, b_sign, x_batch, x_sign, y_batch, f64_to_felt(0.01), f64_to_felt(1e6), bs,
    );
    let trace = prover.build_trace();
    let pub_inputs = prover.get_pub_inputs(&trace);
    let cpu = prover.prove(trace.clone()).expect("winterfell prove failed").to_bytes();
    let gpu_proof = prover.prove_gpu(gpu, trace.clone()).expect("zkb200 prove failed").to_bytes();
    let air = zk_stark_project::training::air::TrainingUpdateAir::new(trace.info().clone(), pub_inputs.clone(), options());
    write_fixture(out, &format!("training_bs{bs}"), 1, &trace, &air, pub_inputs.to_elements(), vec![], &cpu, &gpu_proof);
    report(&format!("training bs={bs}"), &cpu, &gpu_proof)
}

fn aggregation_case(gpu: &GpuContext, out: &PathBuf, clients: usize) -> bool {
    // src/main.rs:442-456 with a deterministic global model
    let g_w: Vec<Vec<Felt>> = (0..AC).map(|j| (0..FE).map(|i| f64_to_felt(100.0 * (1 + i + j) as f64)).collect()).collect();
    let g_b: Vec<Felt> = (0..AC).map(|j| f64_to_felt(10.0 * (j + 1) as f64)).collect();
    let mut local_w = Vec::new();
    let mut local_b = Vec::new();
    for c in 0..clients {
        let v = 0.5 + c as f64;
        local_w.push(vec![vec![f64_to_felt(v); FE]; AC]);
        local_b.push(vec![f64_to_felt(v); AC]);
    }
    let k = f64_to_felt(clients as f64);
    let prover = GlobalUpdateProver::new(options(), g_w, g_b, local_w, local_b, k);
    let trace = prover.build_trace();
    let pub_inputs = prover.get_pub_inputs(&trace);
    let cpu = prover.prove(trace.clone()).expect("winterfell prove failed").to_bytes();
    let gpu_proof = prover.prove_gpu(gpu, trace.clone()).expect("zkb200 prove failed").to_bytes();
    let air = zk_stark_project::aggregation::air::GlobalUpdateAir::new(trace.info().clone(), pub_inputs.clone(), options());
    write_fixture(out, &format!("aggregation_{clients}"), 2, &trace, &air, pub_inputs.to_elements(), vec![k], &cpu, &gpu_proof);
    report(&format!("aggregation clients={clients}"), &cpu, &gpu_proof)
}

fn main() {
    let out = PathBuf::from(env::args().nth(1).unwrap_or_else(|| ".".into()));
    fs::create_dir_all(&out).expect("cannot create output directory");
    let gpu = GpuContext::new(0).expect("no B200 context");
    let mut ok = true;
    for bs in [1usize, 2, 5] {
        ok &= training_case(&gpu, &out, bs);
    }
    for clients in [1usize, 16] {
        ok &= aggregation_case(&gpu, &out, clients);
    }
    println!("{}", if ok { "PARITY OK: copy winterfell_*.json into tests/golden/ of the B200 repository" } else { "PARITY FAILED" });
    std::process::exit(if ok { 0 } else { 1 });
}
