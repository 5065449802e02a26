//! `decode_machine_proof`: liblurkstark's canonical-u32 word image -> `MachineProof<BabyBearPoseidon2>`, i.e. the
//! fields of `CryptoShardProof` (src/core/cli/proofs.rs:22-35) plus the public values.  Mirror of
//! `lurk_b200/proof_format.py::decode_machine_proof`, which is tested here (tests/test_proof_format.py); the word
//! layout is DESIGN.md section 4:
//!
//!   0x4c505246 nshards { shard_index main[8] perm[8] quot[8]
//!     nchips { chip_id log_degree cum_sum[4] prep_local prep_next main_local main_next perm_local perm_next q0 q1 }
//!     npv pv* nfri root[8]* final_poly[4] pow_witness
//!     nqueries { nrounds { nmats { w row[w] }* pathlen digest[8]* }* nsteps { sibling[4] pathlen digest[8]* }* }* }*
//!
//! Field names / nesting of the sphinx types are written from the SP1 v1 lineage (sphinx-core @ 8a39b95 is not in
//! this tree): adjust the struct literals if the pinned commit differs; the traversal does not change.
use hashbrown::HashMap;
use p3_baby_bear::BabyBear;
use p3_commit::BatchOpening;            // { opened_values: Vec<Vec<Val>>, opening_proof: Vec<[Val; 8]> }
use p3_field::{extension::BinomialExtensionField, AbstractExtensionField, AbstractField};
use p3_fri::{CommitPhaseProofStep, FriProof, QueryProof, TwoAdicFriPcsProof};
use sphinx_core::stark::{AirOpenedValues, ChipOpenedValues, MachineProof, ShardCommitment, ShardOpenedValues, ShardProof};
use sphinx_core::utils::BabyBearPoseidon2;

type F = BabyBear;
type EF = BinomialExtensionField<BabyBear, 4>;
const MAGIC: u32 = 0x4c50_5246;

struct Words<'a> { w: &'a [u32], p: usize }
impl<'a> Words<'a> {
    fn u(&mut self) -> u32 { let v = self.w[self.p]; self.p += 1; v }
    /// the library writes canonical residues (`as_canonical_u32`)
    fn felt(&mut self) -> F { F::from_canonical_u32(self.u()) }
    fn ext(&mut self) -> EF { EF::from_base_fn(|_| self.felt()) }
    fn digest(&mut self) -> [F; 8] { core::array::from_fn(|_| self.felt()) }
    fn ext_vec(&mut self) -> Vec<EF> { let n = self.u() as usize; (0..n).map(|_| self.ext()).collect() }
    fn path(&mut self) -> Vec<[F; 8]> { let n = self.u() as usize; (0..n).map(|_| self.digest()).collect() }
}

/// `chip_names[i]` = `MachineAir::name()` of chip i of `build_lair_chip_vector` (src/lair/lair_chip.rs:85-95,196-212):
/// "Entrypoint[k]", "Func[name]"..., "Mem[n-wide]" x 6, "CPU" (the byte chip: sphinx insists on that name).
pub fn decode_machine_proof(words: &[u32], chip_names: &[String]) -> MachineProof<BabyBearPoseidon2> {
    let mut r = Words { w: words, p: 0 };
    assert_eq!(r.u(), MAGIC, "not a liblurkstark proof");
    let nshards = r.u() as usize;
    let mut shard_proofs = Vec::with_capacity(nshards);
    for _ in 0..nshards {
        let _shard_index = r.u();
        let commitment = ShardCommitment {
            main_commit: r.digest().into(),
            permutation_commit: r.digest().into(),
            quotient_commit: r.digest().into(),
        };
        let nchips = r.u() as usize;
        let mut chips = Vec::with_capacity(nchips);
        let mut chip_ordering = HashMap::new();
        for pos in 0..nchips {
            let chip_id = r.u() as usize;
            let log_degree = r.u() as usize;
            let cumulative_sum = r.ext();
            let preprocessed = AirOpenedValues { local: r.ext_vec(), next: r.ext_vec() };
            let main = AirOpenedValues { local: r.ext_vec(), next: r.ext_vec() };
            let permutation = AirOpenedValues { local: r.ext_vec(), next: r.ext_vec() };
            let quotient = vec![r.ext_vec(), r.ext_vec()];
            chip_ordering.insert(chip_names[chip_id].clone(), pos);
            chips.push(ChipOpenedValues { preprocessed, main, permutation, quotient, cumulative_sum, log_degree });
        }
        let npv = r.u() as usize;
        let public_values: Vec<F> = (0..npv).map(|_| r.felt()).collect();
        let nfri = r.u() as usize;
        let commit_phase_commits = (0..nfri).map(|_| r.digest().into()).collect();
        let final_poly = r.ext();
        let pow_witness = r.felt();
        let nq = r.u() as usize;
        let (mut query_proofs, mut query_openings) = (Vec::with_capacity(nq), Vec::with_capacity(nq));
        for _ in 0..nq {
            let nrounds = r.u() as usize;          // preprocessed (if any), main, permutation, quotient
            let mut rounds = Vec::with_capacity(nrounds);
            for _ in 0..nrounds {
                let nmats = r.u() as usize;
                let opened_values = (0..nmats).map(|_| { let w = r.u() as usize; (0..w).map(|_| r.felt()).collect() }).collect();
                rounds.push(BatchOpening { opened_values, opening_proof: r.path() });
            }
            let nsteps = r.u() as usize;
            let commit_phase_openings =
                (0..nsteps).map(|_| CommitPhaseProofStep { sibling_value: r.ext(), opening_proof: r.path() }).collect();
            query_openings.push(rounds);
            query_proofs.push(QueryProof { commit_phase_openings });
        }
        shard_proofs.push(ShardProof {
            commitment,
            opened_values: ShardOpenedValues { chips },
            opening_proof: TwoAdicFriPcsProof {
                fri_proof: FriProof { commit_phase_commits, query_proofs, final_poly, pow_witness },
                query_openings,
            },
            chip_ordering,
            public_values,
        });
    }
    assert_eq!(r.p, words.len(), "trailing words after the last shard proof");
    MachineProof { shard_proofs }
}
// `CryptoProof::from(machine_proof)` (src/core/cli/proofs.rs:87-131) then works unchanged: it strips the public values,
// checks that every shard carries the same ones and keeps the depth (their last DEPTH_W bytes); `!(prove)` writes the
// result with bincode to ~/.lurk/proofs/<key> (src/core/cli/repl.rs:200-203) and `!(verify)` rebuilds the 44 public
// values with `into_machine_proof` (proofs.rs:46-77).
