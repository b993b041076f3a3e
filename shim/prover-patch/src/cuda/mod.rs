//! `prover/src/cuda/mod.rs` — `Machine::prove_with_extensions` (prover/src/machine.rs:130-297) with the B200 backend.
//!
//! Host work is unchanged (steps 1-4: emulation output -> filled traces); everything Stwo's `SimdBackend` did is replaced by
//! calls into libnexus_b200 through the `nexus-b200` crate: tree commits, lookup-element draws on the library's channel, the
//! main component's interaction trace (LogupTraceGenerator on the device), `stwo::prover::prove`.
//! Apply with `shim/prover-patch/apply.md`; this file lives INSIDE the prover crate because it uses pub(crate) items.
use std::cell::RefCell;
use std::rc::Rc;

use nexus_b200::recorder::{AirBytecode, Allocator, ParamTable, Recorder};
use nexus_b200::{Channel, Columns, Context, Error as NbError, PcsParams};
use nexus_vm::{emulator::{InternalView, View}, trace::Trace};
use stwo::core::fields::{m31::BaseField, qm31::SecureField, FieldExpOps};
use stwo::core::pcs::PcsConfig;
use stwo::prover::backend::{simd::SimdBackend, Column};
use stwo::prover::poly::{circle::CircleEvaluation, BitReversedOrder};
use stwo::prover::ProvingError;
use stwo_constraint_framework::{FrameworkEval, ORIGINAL_TRACE_IDX, PREPROCESSED_TRACE_IDX};

use crate::components::{AllLookupElements, MachineEval, LOG_CONSTRAINT_DEGREE};
use crate::extensions::{ExtensionComponent, ExtensionsConfig};
use crate::machine::{Machine, Proof, BASE_EXTENSIONS};
use crate::trace::{program::iter_program_steps, program_trace::{ProgramTraceRef, ProgramTracesBuilder}, sidenote::SideNote, PreprocessedTraces, TracesBuilder};
use crate::traits::MachineChip;

type Evals = Vec<CircleEvaluation<SimdBackend, BaseField, BitReversedOrder>>;

/// group consecutive equal-length evaluations into device batches (commitment order is preserved)
fn upload(ctx: &Context, evals: &Evals) -> Result<Vec<Columns>, NbError> {
    let mut out = vec![];
    let mut i = 0;
    while i < evals.len() {
        let log = evals[i].domain.log_size();
        let mut j = i;
        while j < evals.len() && evals[j].domain.log_size() == log { j += 1; }
        let cpu: Vec<Vec<BaseField>> = evals[i..j].iter().map(|e| e.values.to_cpu()).collect();   // BaseColumn -> Vec<BaseField>, already bit-reversed
        let refs: Vec<&[BaseField]> = cpu.iter().map(|c| c.as_slice()).collect();
        out.push(ctx.upload_finalized(&refs, log)?);
        i = j;
    }
    Ok(out)
}

fn backend_err(e: NbError) -> ProvingError {
    match e { NbError::ConstraintsNotSatisfied => ProvingError::ConstraintsNotSatisfied, other => panic!("nexus-b200 backend failure: {other}") }
}

/// one process per GPU: this process's rank, the number of ranks, and the communicator id (`Context::comm_unique_id()` on rank 0, shipped to the
/// other ranks by the caller's transport)
pub struct Shard<'a> { pub rank: i32, pub world: i32, pub unique_id: &'a [u8] }

impl<C: MachineChip + Sync> Machine<C> {
    pub fn prove_cuda(extensions: &[ExtensionComponent], trace: &impl Trace, view: &View) -> Result<Proof, ProvingError> {
        Self::prove_cuda_impl(extensions, trace, view, None)
    }
    /// ONE proof by `shard.world` GPUs (Python mirror: nexus_zkvm_b200/machine.py:prove_sharded).  Every rank calls this with the same
    /// trace (the trace filler is sequential host code) and uploads only its `shard_range` of the main component's columns; every rank returns the
    /// same `Proof`, byte-identical to `prove_cuda`'s.
    pub fn prove_cuda_sharded(extensions: &[ExtensionComponent], trace: &impl Trace, view: &View, shard: Shard) -> Result<Proof, ProvingError> {
        Self::prove_cuda_impl(extensions, trace, view, Some(shard))
    }
    fn prove_cuda_impl(extensions: &[ExtensionComponent], trace: &impl Trace, view: &View, shard: Option<Shard>) -> Result<Proof, ProvingError> {
        // ---- steps 1-4, verbatim host work (machine.rs:135-183) ----
        let num_steps = trace.get_num_steps();
        let program_len = view.get_program_memory().program.len();
        let log_size = Self::max_log_size(&[num_steps, program_len]).max(PreprocessedTraces::MIN_LOG_SIZE);
        let extensions_config = ExtensionsConfig::from(extensions);
        let extensions_iter = BASE_EXTENSIONS.iter().chain(extensions);
        let preprocessed_trace = PreprocessedTraces::new(log_size);
        let mut prover_traces = TracesBuilder::new(log_size);
        let init_memory = [view.get_ro_initial_memory(), view.get_rw_initial_memory(), view.get_public_input()].concat();
        let program_trace_ref = ProgramTraceRef { program_memory: view.get_program_memory(), init_memory: &init_memory,
                                                  exit_code: view.get_exit_code(), public_output: view.get_public_output() };
        let program_traces = ProgramTracesBuilder::new(log_size, program_trace_ref);
        let mut side_note = SideNote::new(&program_traces, view);
        for (row_idx, step) in iter_program_steps(trace, prover_traces.num_rows()).enumerate() {
            C::fill_main_trace(&mut prover_traces, row_idx, &step, &mut side_note, &extensions_config);
        }
        let finalized_trace = prover_traces.finalize();
        let finalized_program_trace = program_traces.finalize();
        let all_log_sizes: Vec<u32> = std::iter::once(log_size).chain(extensions_iter.clone().map(|ext| ext.compute_log_size(&side_note))).collect();

        // ---- backend set-up (machine.rs:184-206) ----
        let config = PcsConfig::default();
        let ctx = Context::new(shard.as_ref().map_or(0, |s| s.rank)).map_err(backend_err)?;      // one process per GPU: device = rank
        if let Some(s) = &shard { ctx.comm_init(s.rank, s.world, s.unique_id).map_err(backend_err)?; }
        // sharded: the leading 2^log_size-row columns of a tree (the main component's) -> (this rank's shard_range of them, their total number,
        // the smaller extension columns, replicated); see Scheme::commit_sharded
        let split = |evals: &Evals| -> Result<(Option<Columns>, usize, Vec<Columns>), NbError> {
            let s = shard.as_ref().expect("sharded path");
            let n_big = evals.iter().take_while(|e| e.domain.log_size() == log_size).count();
            let (first, count) = Context::shard_range(n_big, s.world, s.rank);
            let big = if count > 0 { Some(upload(&ctx, &evals[first..first + count].to_vec())?.remove(0)) } else { None };
            Ok((big, n_big, upload(&ctx, &evals[n_big..].to_vec())?))
        };
        let max_log = all_log_sizes.iter().copied().max().unwrap_or(0).max(log_size);
        ctx.precompute_twiddles(max_log + LOG_CONSTRAINT_DEGREE + config.fri_config.log_blowup_factor).map_err(backend_err)?;
        let mut ch: Channel = ctx.channel().map_err(backend_err)?;
        for byte in view.view_associated_data().unwrap_or_default() { ch.mix_u64(byte.into()); }
        let mut scheme = ctx.scheme(PcsParams::from(config)).map_err(backend_err)?;
        scheme.set_constraint_log_degree(LOG_CONSTRAINT_DEGREE).map_err(backend_err)?;
        for ls in &all_log_sizes { ch.mix_u64(*ls as u64); }

        // ---- tree 0 (machine.rs:208-228) ----
        let extension_traces: Vec<_> = extensions_iter.clone().zip(all_log_sizes.get(1..).unwrap_or_default())
            .map(|(ext, ls)| ext.generate_component_trace(*ls, program_trace_ref, &mut side_note)).collect();
        let mut t0: Evals = preprocessed_trace.clone().into_circle_evaluation().into_iter().chain(finalized_program_trace.clone().into_circle_evaluation()).collect();
        for et in &extension_traces { t0.extend(et.to_circle_evaluation(PREPROCESSED_TRACE_IDX)); }
        let d0 = if shard.is_none() { upload(&ctx, &t0).map_err(backend_err)? } else { vec![] };
        let mut keep_alive: Vec<Columns> = vec![];                    // sharded: the uploaded shards must outlive the commits that read them
        if shard.is_none() {
            scheme.commit(&d0.iter().collect::<Vec<_>>(), &mut ch).map_err(backend_err)?;
        } else {
            let (big, n_big, small) = split(&t0).map_err(backend_err)?;
            scheme.commit_sharded(big.as_ref(), n_big, log_size, &small.iter().collect::<Vec<_>>(), &[], true, &mut ch).map_err(backend_err)?;
            keep_alive.extend(big); keep_alive.extend(small);
        }
        // ---- tree 1 (machine.rs:230-237) ----
        let mut t1: Evals = finalized_trace.clone().into_circle_evaluation();
        for et in &extension_traces { t1.extend(et.to_circle_evaluation(ORIGINAL_TRACE_IDX)); }
        let d1 = if shard.is_none() { upload(&ctx, &t1).map_err(backend_err)? } else { vec![] };
        if shard.is_none() {
            scheme.commit(&d1.iter().collect::<Vec<_>>(), &mut ch).map_err(backend_err)?;
        } else {
            let (big, n_big, small) = split(&t1).map_err(backend_err)?;
            scheme.commit_sharded(big.as_ref(), n_big, log_size, &small.iter().collect::<Vec<_>>(), &next_row_columns(), true, &mut ch).map_err(backend_err)?;
            keep_alive.extend(big); keep_alive.extend(small);
        }

        // ---- lookup elements (machine.rs:239-240): drawn from the library's channel through a `Channel` adapter (lookups.rs) ----
        let mut lookup_elements = AllLookupElements::default();
        C::draw_lookup_elements(&mut lookup_elements, &mut lookups::ChannelAdapter(&mut ch), &extensions_config);

        // ---- record the AIR now that the elements are known (they reach the evaluator as literals -> parameter slots) ----
        let alloc = Rc::new(RefCell::new(Allocator::default()));
        let params = Rc::new(RefCell::new(ParamTable::default()));
        let mut comps: Vec<Vec<u32>> = vec![];
        let main_eval = MachineEval::<C>::new(log_size, lookup_elements.clone(), extensions_config.clone());
        let main_rec = main_eval.evaluate(Recorder::new(log_size, LOG_CONSTRAINT_DEGREE, alloc.clone(), params.clone()));
        let n_main_inter = 4 * main_rec.n_logup_columns();        // the main component's secure LogUp columns, as base-field coordinate columns
        comps.push(main_rec.finish());
        for (ext, ls) in extensions_iter.clone().zip(all_log_sizes.get(1..).unwrap_or_default()) {
            comps.push(ext.record_air(alloc.clone(), params.clone(), &lookup_elements, *ls));       // added by apply.md next to to_component_prover
        }
        let n_params = params.borrow().names.len();
        let air = ctx.air(&AirBytecode::assemble(n_params as u32, comps)).map_err(backend_err)?;
        let mut table = vec![SecureField::default(); n_params];
        for (i, v) in &params.borrow().literals { table[*i as usize] = *v; }
        let shift_slots: Vec<usize> = params.borrow().names.iter().enumerate().filter(|(_, n)| n.as_str() == "cumsum_shift").map(|(i, _)| i).collect();

        // ---- interaction traces (machine.rs:242-263): main component on the device, extensions (<= 2^8 rows but RamInitFinal) on the host ----
        // (sharded: from the trace rows the sharded commits kept; the result is this rank's COLUMN shard of the 4 x n_logup_cols interaction columns)
        let (main_inter, claimed_sum) = if shard.is_none() {
            ctx.gen_interaction_trace(&air, 0, &d0.iter().collect::<Vec<_>>(), &d1.iter().collect::<Vec<_>>(), &table).map_err(backend_err)?
        } else {
            scheme.gen_interaction_trace_sharded(&air, 0, &table).map_err(backend_err)?
        };
        let mut all_claimed_sums = vec![claimed_sum];
        let mut inter_batches: Vec<Columns> = vec![main_inter];
        for (ext, et) in extensions_iter.clone().zip(extension_traces) {
            let (it, cs) = ext.generate_interaction_trace(et, &side_note, &lookup_elements);
            all_claimed_sums.push(cs);
            inter_batches.extend(upload(&ctx, &it).map_err(backend_err)?);
        }
        for ((slot, cs), ls) in shift_slots.iter().zip(&all_claimed_sums).zip(&all_log_sizes) {
            table[*slot] = *cs * SecureField::from(BaseField::from_u32_unchecked(1 << ls)).inverse();   // LogupAtRow::new: claimed_sum / 2^log_size
        }
        ch.mix_felts(&all_claimed_sums);
        if shard.is_none() {
            scheme.commit(&inter_batches.iter().collect::<Vec<_>>(), &mut ch).map_err(backend_err)?;
        } else {
            // the logup constraint reads the LAST secure column at the previous row (LogupAtRow's [-1, 0] mask): replicate its 4 coordinates
            let last4: Vec<u32> = (n_main_inter as u32 - 4..n_main_inter as u32).collect();
            let big = if inter_batches[0].n_cols() > 0 { Some(&inter_batches[0]) } else { None };
            scheme.commit_sharded(big, n_main_inter, log_size, &inter_batches[1..].iter().collect::<Vec<_>>(), &last4, false, &mut ch).map_err(backend_err)?;
        }

        // ---- stwo::prover::prove (machine.rs:286-290) ----
        // nb200_prove follows the trees it finds: row-sharded constraint rows / DEEP quotients and replicated FRI when they are sharded
        let bytes = scheme.prove(&air, &table, &mut ch).map_err(backend_err)?;
        drop(keep_alive);
        let stark_proof = postcard::from_bytes(&bytes).expect("libnexus_b200 emits postcard(StarkProof<Blake2sMerkleHasher>)");
        Ok(Proof { stark_proof, claimed_sum: all_claimed_sums, log_size: all_log_sizes })
    }
}

/// main-trace column indices (in commitment order) that constraints read at the next row: `Column::reads_next_row_mask` (column.rs:17,
/// trace/eval.rs:35) — what `commit_sharded` must replicate on every rank
fn next_row_columns() -> Vec<u32> {
    use crate::column::Column;
    let mut out = vec![];
    let mut offset = 0u32;
    for col in Column::ALL_VARIANTS {                                      // enum order = commitment order (trace/eval.rs:30-45)
        if col.reads_next_row_mask() { out.extend(offset..offset + col.size() as u32); }
        offset += col.size() as u32;
    }
    out
}

pub mod lookups;
