//! `CudaProver`: sphinx-core's `Prover` for `(BabyBearPoseidon2, LairChip)` backed by liblurkstark.
//!
//!   machine.prove::<LocalProver<_, _>>(&pk, shard, &mut challenger, opts)     // today, src/core/cli/repl.rs:196
//!   machine.prove::<CudaProver<_, _>>(&pk, shard, &mut challenger, opts)      // with this file
//!
//! The machine's static half (program image + captured chip vector) is built once per toplevel next to
//! `new_machine` (src/core/stark_machine.rs:20-30) and cached; per proof the record is flattened (flatten.rs),
//! proved (`lsk_prove`: upload, every shard, one word image) and decoded (proof.rs).
use std::marker::PhantomData;
use std::sync::OnceLock;

use p3_baby_bear::BabyBear;
use sphinx_core::stark::{MachineProof, Prover, StarkMachine, StarkProvingKey};
use sphinx_core::utils::{BabyBearPoseidon2, SphinxCoreOpts};

use crate::cuda::{ffi::*, flatten::flatten, proof::decode_machine_proof};
use crate::lair::{chipset::Chipset, execute::{Shard, ShardingConfig}, lair_chip::LairChip, toplevel::Toplevel};

pub struct CudaMachine {
    ctx: *mut LskCtx,
    prog: *mut LskProgram,
    raw: *mut LskMachine,
    chip_names: Vec<String>,
}
unsafe impl Send for CudaMachine {}
unsafe impl Sync for CudaMachine {}

impl CudaMachine {
    /// `surface`: the toplevel's `FuncE`s serialised as the "surface image" (lurk_b200/csrc/host_lair.cpp); the library's
    /// host layer expands + compiles them (Toplevel::new), computes the layouts (func_chip.rs:90-277) and runs every
    /// chip's `Air::eval` symbolically (Chip::new) -- or port those two passes and hand over their images directly.
    pub fn new(surface: &[u32], entry_func: u32, chip_names: Vec<String>, opts: &SphinxCoreOpts) -> Self {
        let (mut ctx, mut prog, mut raw) = (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
        let (mut img, mut img_len, mut blob, mut blob_len) = (std::ptr::null_mut(), 0usize, std::ptr::null_mut(), 0usize);
        unsafe {
            check(lsk_ctx_create(0, &mut ctx));
            check(lsk_host_compile_toplevel(surface.as_ptr(), surface.len(), &mut img, &mut img_len));
            check(lsk_program_load(img, img_len, &mut prog));
            check(lsk_host_chip_vector(surface.as_ptr(), surface.len(), entry_func, &mut blob, &mut blob_len));
            // FRI parameters: sphinx defaults (100 queries, env FRI_QUERIES; 16 proof-of-work bits)
            check(lsk_machine_new(ctx, prog, entry_func, blob, blob_len, fri_queries(opts), 16, &mut raw));
            lsk_words_free(img);
            lsk_words_free(blob);
        }
        Self { ctx, prog, raw, chip_names }
    }
}
impl Drop for CudaMachine {
    fn drop(&mut self) {
        unsafe { lsk_machine_free(self.raw); lsk_program_free(self.prog); lsk_ctx_destroy(self.ctx); }
    }
}
fn fri_queries(_opts: &SphinxCoreOpts) -> u32 {
    std::env::var("FRI_QUERIES").ok().and_then(|s| s.parse().ok()).unwrap_or(100)
}

/// one per toplevel; filled by the caller of `new_machine`
pub static CUDA_MACHINE: OnceLock<CudaMachine> = OnceLock::new();

pub struct CudaProver<SC, A>(PhantomData<(SC, A)>);

impl<'a, C1: Chipset<BabyBear>, C2: Chipset<BabyBear>> Prover<BabyBearPoseidon2, LairChip<'a, BabyBear, C1, C2>>
    for CudaProver<BabyBearPoseidon2, LairChip<'a, BabyBear, C1, C2>>
{
    fn prove_shards(
        _machine: &StarkMachine<BabyBearPoseidon2, LairChip<'a, BabyBear, C1, C2>>,
        _pk: &StarkProvingKey<BabyBearPoseidon2>,
        shards: Vec<Shard<'a, BabyBear>>,
        _challenger: &mut <BabyBearPoseidon2 as sphinx_core::stark::StarkGenericConfig>::Challenger,
        _opts: SphinxCoreOpts,
    ) -> MachineProof<BabyBearPoseidon2> {
        let cuda = CUDA_MACHINE.get().expect("CudaMachine::new was not called for this toplevel");
        // every shard borrows the same record (Shard::shard, execute.rs:186-216); the library re-derives the shards
        // from the shard size, so that each one is uploaded and proved without materialising per-shard copies
        let record = shards[0].queries();
        let shard_size = ShardingConfig::default().max_shard_size;      // env SHARD_SIZE, default 1 << 22
        let toplevel: &Toplevel<BabyBear, C1, C2> = todo_toplevel_of(record);
        let flat = flatten(toplevel, cuda.prog, record);
        let (mut words, mut len) = (std::ptr::null_mut(), 0usize);
        check(unsafe { lsk_prove(cuda.raw, flat.raw, shard_size, &mut words, &mut len) });
        let proof = decode_machine_proof(unsafe { std::slice::from_raw_parts(words, len) }, &cuda.chip_names);
        unsafe { lsk_proof_free(words) };
        // The library runs its own transcript (DuplexChallenger over the same width-16 permutation, observing the
        // preprocessed commitment, pc_start = 0, then every shard's main commitment and public values -- the order
        // sphinx uses); `_challenger` is therefore left untouched, as LocalProver leaves it after cloning per shard.
        proof
    }
}

/// The toplevel is not reachable from a `QueryRecord`; the reference passes it around separately (`FuncChip` borrows
/// it).  A real integration stores `&Toplevel` next to CUDA_MACHINE; left as the one explicit hole of this sketch.
fn todo_toplevel_of<'t, C1: Chipset<BabyBear>, C2: Chipset<BabyBear>>(
    _record: &crate::lair::execute::QueryRecord<BabyBear>,
) -> &'t Toplevel<BabyBear, C1, C2> {
    unimplemented!("keep a &Toplevel beside CUDA_MACHINE when new_machine is called")
}

/// `machine.verify(&vk, &proof, &mut challenger)` (repl.rs:186,198; meta.rs:928-932) on the word image, no GPU needed.
pub fn verify_words(cuda: &CudaMachine, words: &[u32]) -> Result<(), String> {
    match unsafe { lsk_verify(cuda.raw, words.as_ptr(), words.len()) } {
        LSK_OK => Ok(()),
        LSK_ERR_VERIFY => Err(last_error()),
        rc => panic!("liblurkstark error {rc}: {}", last_error()),
    }
}
