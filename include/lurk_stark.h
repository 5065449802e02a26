/* lurk_stark.h -- C ABI of liblurkstark, the B200-native proving hot path for Lurk.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * argumentcomputer/lurk tree).  Conventions:
 *   - return 0 on success, negative on error; never throws or aborts across the boundary;
 *     lsk_last_error() gives the message (thread-local, valid until the next call on this thread);
 *   - field elements cross the boundary as canonical u32 residues of BabyBear
 *     (from_canonical_u32 / as_canonical_u32); Montgomery form is internal;
 *   - matrices cross the boundary row-major exactly as RowMajorMatrix.values;
 *   - `const uint32_t*` arguments are HOST pointers unless the name says `_dev`;
 *   - an lsk_ctx owns one CUDA device + stream; it is stream-ordered and not re-entrant;
 *     distinct contexts are independent (one process per GPU).
 *   - there is NO CPU fallback: every compute entry point fails with LSK_ERR_CUDA when no
 *     device is usable.
 */
#ifndef LURK_STARK_H
#define LURK_STARK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSK_OK 0
#define LSK_ERR_ARG (-1)
#define LSK_ERR_CUDA (-2)
#define LSK_ERR_EXEC (-3)   /* Lair execution error (the reference panics / bails) */
#define LSK_ERR_VERIFY (-4)

typedef struct lsk_ctx lsk_ctx;
typedef struct lsk_program lsk_program;   /* a loaded Toplevel (src/lair/toplevel.rs:25-56) */
typedef struct lsk_record lsk_record;     /* a flattened QueryRecord (src/lair/execute.rs:67-75) */
typedef struct lsk_machine lsk_machine;   /* StarkMachine + pk/vk (src/core/stark_machine.rs:20-30) */
typedef struct lsk_vk lsk_vk;             /* verifying key + chip vector: all `machine.verify` needs */

const char* lsk_last_error(void);
int lsk_version(void);

/* ---- context ------------------------------------------------------------------------------- */
int lsk_ctx_create(int device, lsk_ctx** out);
void lsk_ctx_destroy(lsk_ctx* ctx);
/* cudaStream_t of the context (so callers can bracket work with their own CUDA events) */
void* lsk_ctx_stream(lsk_ctx* ctx);
int lsk_ctx_sync(lsk_ctx* ctx);

/* Per-kernel accounting (new; the reference has no profiler).  Launch and
 * algorithmic-byte counters are always on; lsk_profile_enable(ctx, 1) additionally brackets every
 * kernel family with CUDA events on the context's stream.  lsk_profile_enable resets the counters;
 * lsk_profile_read synchronises, fills up to `cap` entries per array (index = tag, name from
 * lsk_profile_tag_name) and returns the number of tags. */
int lsk_profile_enable(lsk_ctx* ctx, int on);
int lsk_profile_read(lsk_ctx* ctx, double* ms, double* bytes, uint64_t* launches, int cap);
/* arithmetic work per tag since lsk_profile_enable, where the launcher states it: for the hash kernels
 * (merkle_leaf, merkle_layer, fri_leaf) the number of width-16 Poseidon2 permutations (772 modular products each) */
int lsk_profile_read_work(lsk_ctx* ctx, double* work, int cap);
const char* lsk_profile_tag_name(int tag);

/* ---- A1: Poseidon2 hashing of content-addressed data ------------------------------------------
 * replaces PoseidonChipset::hash / execute_simple, src/core/poseidon.rs:29-39,61-63 as called from
 * ZStore::hash3/4/5, src/core/zstore.rs:305-333.  width in {24, 32, 40} (also 16: the commitment
 * permutation); `in` is n rows of `width` words, `out8` n rows of 8 words. */
int lsk_poseidon2_hash(lsk_ctx* ctx, int width, const uint32_t* in, size_t n, uint32_t* out8);
/* the same hash on the host, one preimage after another: what ZStore does while it interns data
 * (`Hasher::hash` -> `PoseidonChipset::execute_simple`, src/core/zstore.rs:228-240,305-333); no GPU involved */
int lsk_poseidon2_hash_host(int width, const uint32_t* in, size_t n, uint32_t* out8);
int lsk_poseidon2_hash_dev(lsk_ctx* ctx, int width, const uint32_t* in_dev, size_t n, uint32_t* out8_dev);
/* full permutation (p3 Poseidon2::permute, hasher() in src/poseidon/config.rs:77-95) */
int lsk_poseidon2_permute(lsk_ctx* ctx, int width, const uint32_t* in, size_t n, uint32_t* out);
/* replaces PoseidonChipset::populate_witness, src/core/poseidon.rs:65-72 (+ Poseidon2Cols::populate,
 * src/poseidon/wide/trace.rs:12-82): out row = digest[8] ++ wide witness; lsk_poseidon2_witness_cols
 * gives the row length (457 / 611 / 763). */
int lsk_poseidon2_witness_cols(int width);
/* The permutation's parameters as canonical residues (src/poseidon/constants.rs via config.rs:157-287):
 * diag[width], full_rc[8][width], part_rc[R_P]; returns R_P (21 / 30 / 38; 13 for width 16).  Host-only;
 * used by the AIR capture of the hasher chips (Poseidon2Cols::eval, src/poseidon/wide/air.rs:11-124). */
int lsk_poseidon2_constants(int width, uint32_t* diag, uint32_t* full_rc, uint32_t* part_rc, size_t part_cap);
int lsk_poseidon2_witness(lsk_ctx* ctx, int width, const uint32_t* in, size_t n, uint32_t* out);

/* ---- A2/A3: program + flattened record ------------------------------------------------------
 * `words` is the flat Toplevel image produced by the host layer (lurk_b200/lair.py::serialize;
 * format documented in DESIGN.md).  The executor is the host-side producer of the hot path's
 * input (Toplevel::execute, src/lair/execute.rs:375-783) re-designed to emit GPU-ready SoA tables
 * with resolved callee / pointer row indices. */
int lsk_program_load(const uint32_t* words, size_t n, lsk_program** out);
void lsk_program_free(lsk_program* p);
int lsk_program_num_funcs(const lsk_program* p);
int lsk_program_func_width(const lsk_program* p, uint32_t func);   /* BaseAir::width, func_chip.rs:83-87 */

int lsk_record_new(const lsk_program* p, lsk_record** out);        /* QueryRecord::new, execute.rs:263-284 */
void lsk_record_free(lsk_record* r);
int lsk_execute(const lsk_program* p, lsk_record* r, uint32_t func, const uint32_t* args, size_t nargs,
                uint32_t* out, size_t nout);                        /* Toplevel::execute, execute.rs:376-392 */
int lsk_inject_inv_queries(const lsk_program* p, lsk_record* r, uint32_t func, const uint32_t* inps,
                           const uint32_t* outs, size_t n);         /* execute.rs:299-319 */
/* Importing a QueryRecord the caller's own executor produced -- the drop-in path behind the UNCHANGED Lurk
 * evaluator: Rust keeps `toplevel.execute(...)` (src/lair/execute.rs:375-783) and hands the result over table by
 * table.  lsk_record_import_func: QueryRecord::func_queries[func] (execute.rs:26-33,67-75) in insertion (= nonce)
 * order -- args [n][in], output [n][out], provide [n][2] = (last_nonce, last_count) (src/air/builder.rs:135-148),
 * requires as CSR pairs (prev_nonce, prev_count) with n + 1 offsets; for partial funcs depth [n] and
 * depth_requires (CSR, n + 1 offsets), else NULL.  lsk_record_import_mem: mem_queries[mem_idx] (lengths
 * 2,3,4,5,6,8; execute.rs:243-257) in insertion order, pointer = index + 1.  lsk_record_import_bytes: BytesRecord
 * (src/gadgets/bytes/record.rs:14-110) as [65536][6][2] (nonce, count) per ByteInput and operation (range_u8,
 * range_u16, less_than, and, xor, or); NULL = none.  lsk_record_import_public_values: execute.rs:383-391.
 * Preimages no imported query produced (inject_inv_queries, execute.rs:299-319) go through lsk_inject_inv_queries.
 * lsk_record_resolve then replays every row against the tables to find the callee / memory row of each Call,
 * PreImg and Store -- the lookups src/lair/trace.rs:325-373 repeats per row at trace-generation time -- checking
 * that every row returns its recorded output; afterwards the record is exactly what lsk_execute builds
 * (tests/test_host_native.py holds the two equal table for table, tests/test_gpu_prover.py proof for proof). */
int lsk_record_import_func(const lsk_program* p, lsk_record* r, uint32_t func, size_t n, const uint32_t* args,
                           const uint32_t* outs, const uint32_t* provide, const uint32_t* req_off, const uint32_t* reqs,
                           const uint32_t* depth, const uint32_t* dreq_off, const uint32_t* dreqs);
int lsk_record_import_mem(lsk_record* r, uint32_t mem_idx, size_t n, const uint32_t* vals, const uint32_t* provide);
int lsk_record_import_bytes(lsk_record* r, const uint32_t* records);
int lsk_record_import_public_values(lsk_record* r, const uint32_t* pv, size_t n);
int lsk_record_resolve(const lsk_program* p, lsk_record* r);
/* the inverse of the imports (row-ordered tables: call lsk_record_finalize first).  sizes = {rows, require pairs,
 * hints, depth-require pairs}; offset arrays receive rows + 1 entries; any output pointer may be NULL */
int lsk_record_func_sizes(const lsk_record* r, uint32_t func, size_t sizes[4]);
int lsk_record_export_func(const lsk_record* r, uint32_t func, uint32_t* args, uint32_t* outs, uint32_t* provide,
                           uint32_t* req_off, uint32_t* reqs, uint32_t* hint_off, uint32_t* hints, uint32_t* depth,
                           uint32_t* dreq_off, uint32_t* dreqs);
int lsk_record_export_mem(const lsk_record* r, uint32_t mem_idx, uint32_t* vals, uint32_t* provide);
int lsk_record_export_bytes(const lsk_record* r, uint32_t* records);
size_t lsk_record_num_queries(const lsk_record* r, uint32_t func);
size_t lsk_record_num_mem(const lsk_record* r, uint32_t mem_idx);
size_t lsk_record_public_values(const lsk_record* r, uint32_t* out, size_t cap);
uint32_t lsk_record_num_shards(const lsk_record* r, uint32_t shard_size);  /* Shard::shard, execute.rs:186-216 */
/* Re-pack the per-row require / hint arrays in row order (the executor appends them in completion order).
 * Optional: afterwards a shard's rows own one contiguous range of every table, and lsk_shard_commit uploads
 * only that range instead of the whole record (what each of N GPUs would otherwise copy N times). */
int lsk_record_finalize(lsk_record* r);
/* Page-lock the record's tables in place (cudaHostRegister) so that lsk_prove's upload runs at PCIe
 * speed; idempotent; undone by lsk_record_unpin, lsk_execute (the tables grow) and lsk_record_free. */
int lsk_record_pin(lsk_record* r);
void lsk_record_unpin(lsk_record* r);
/* bytes a prove call moves host->device for this record (for end-to-end accounting) */
size_t lsk_record_device_bytes(const lsk_record* r);

/* ---- A4/A5: trace generation -------------------------------------------------------------------
 * replaces FuncChip::generate_trace (src/lair/trace.rs:72-135), MemChip::generate_trace
 * (src/lair/memory.rs:30-68), BytesChip::generate_trace / preprocessed_trace
 * (src/gadgets/bytes/trace.rs:49-101).  Output: row-major canonical, height x width, where
 * height is given by the *_height call (next power of two; >= 4 for Mem; 65536 for Bytes). */
size_t lsk_func_trace_height(const lsk_record* r, uint32_t func, uint32_t shard, uint32_t shard_size);
int lsk_func_trace(lsk_ctx* ctx, const lsk_program* p, const lsk_record* r, uint32_t func, uint32_t shard,
                   uint32_t shard_size, uint32_t* out_row_major);
size_t lsk_mem_trace_height(const lsk_record* r, uint32_t mem_idx);
int lsk_mem_trace(lsk_ctx* ctx, const lsk_record* r, uint32_t mem_idx, uint32_t* out_row_major);
int lsk_bytes_trace(lsk_ctx* ctx, const lsk_record* r, uint32_t shard, uint32_t* out_row_major /* 65536 x 13 */);
int lsk_bytes_preprocessed(lsk_ctx* ctx, uint32_t* out_row_major /* 65536 x 6 */);

/* ---- A7-A12: the STARK machine -------------------------------------------------------------------
 * lsk_machine_new = new_machine (src/core/stark_machine.rs:20-30) + machine.setup (repl.rs:175-176):
 * `blob` is the captured chip vector ('MACH' image, lurk_b200/air.py::serialize_machine; format in
 * csrc/machine.h), i.e. what sphinx's Chip::new collects by running Air::eval symbolically
 * (src/lair/lair_chip.rs:196-231).  Commits the byte table's preprocessed trace.  The machine
 * borrows `ctx` and `prog`. */
int lsk_machine_new(lsk_ctx* ctx, const lsk_program* prog, uint32_t entry_func, const uint32_t* blob, size_t n,
                    uint32_t num_queries, uint32_t pow_bits, lsk_machine** out);
void lsk_machine_free(lsk_machine* m);
int lsk_machine_prep_commit(const lsk_machine* m, uint32_t out8[8]);
/* Constraint programs of >= 16 instructions per row (above all the hasher chips behind `Air::eval` of
 * src/poseidon/wide/air.rs:11-124, 5-10 k instructions) can run as machine code: the program is emitted as CUDA source, compiled by NVRTC
 * for sm_100a (1 s for a small chip, ~20 s for a hasher chip) and kept in a cubin cache (env LSK_JIT_CACHE, default <library directory>/jit_cache).
 * Env LSK_JIT: "0" always interpret, "1" compile on a cache miss, unset: use cached cubins only.  The interpreter
 * and the compiled form produce identical proofs.
 *   lsk_jit_precompile: fill the cache for every eligible chip of a 'MACH' blob; needs NVRTC, no GPU.
 *   lsk_machine_jit_chips: how many chips of this machine run compiled programs. */
int lsk_jit_precompile(const uint32_t* blob, size_t n, uint32_t* compiled, uint32_t* cached);
int lsk_machine_jit_chips(const lsk_machine* m);
/* machine.prove::<LocalProver<_,_>>(&pk, shard, &mut challenger, opts): src/core/cli/repl.rs:194-196,
 * benches/fib.rs:154-158, tests/fib.rs:90-94.  `rec` is a HOST record; it is uploaded, sharded by
 * `shard_size` rows (Shard::shard, src/lair/execute.rs:186-216) and proved; `*words` receives a
 * malloc'd proof (canonical u32 words, layout in DESIGN.md), free with lsk_proof_free. */
int lsk_prove(lsk_machine* m, const lsk_record* rec, uint32_t shard_size, uint32_t** words, size_t* len);
void lsk_proof_free(uint32_t* words);
/* machine.verify(&vk, &proof, &mut challenger): src/core/cli/repl.rs:186,198, src/core/cli/meta.rs:928-932,
 * src/lair/lair_chip.rs:273-275, tests/fib.rs:96.  Host code (the reference verifies on the CPU too);
 * LSK_OK = accepted, LSK_ERR_VERIFY = rejected with the reason in lsk_last_error().  lsk_vk_* is the
 * same verifier without a machine / device: `blob` as for lsk_machine_new, `prep_commit` from
 * lsk_machine_prep_commit (canonical), has_prep = whether any chip has a preprocessed trace. */
int lsk_verify(const lsk_machine* m, const uint32_t* words, size_t len);
int lsk_vk_new(const uint32_t* blob, size_t n, const uint32_t prep_commit[8], int has_prep, uint32_t num_queries,
               uint32_t pow_bits, lsk_vk** out);
void lsk_vk_free(lsk_vk* vk);
int lsk_vk_verify(const lsk_vk* vk, const uint32_t* words, size_t len);
/* host<->device bytes moved by the last lsk_prove / lsk_shard_* sequence */
size_t lsk_machine_last_h2d_bytes(const lsk_machine* m);
size_t lsk_machine_last_d2h_bytes(const lsk_machine* m);
/* Two-phase form for shard-per-GPU proving (one process per GPU): every rank commits its shards,
 * the 8-word roots are all-gathered by the caller, then each rank proves its shards against the
 * transcript that has observed all roots.  Concatenating {0x4c505246, nshards, shard proofs in
 * order} gives exactly lsk_prove's output. */
int lsk_shard_commit(lsk_machine* m, const lsk_record* rec, uint32_t shard, uint32_t shard_size, uint32_t root_out[8]);
int lsk_shard_prove(lsk_machine* m, const lsk_record* rec, uint32_t shard, const uint32_t* all_roots, uint32_t nshards,
                    uint32_t** words, size_t* len);
void lsk_release_record(lsk_machine* m);   /* drops the device copy of the record + uncommitted shards */

/* ---- stage-level entry points (parity tests and roofline measurement) ---------------------------------
 * Third-party p3 pieces behind machine.prove: Radix2DitParallel::coset_lde_batch + bit_reverse_rows
 * and FieldMerkleTreeMmcs::commit.  `rows` is h x w row-major canonical over in_shift*H_h (natural
 * rows), shift_kind 0: in_shift = 1, 1: in_shift = g, 2: in_shift = g*w_2h; `out` is 2h x w
 * row-major canonical over g*H_2h, rows bit-reversed. */
int lsk_coset_lde(lsk_ctx* ctx, const uint32_t* rows, size_t h, size_t w, int shift_kind, uint32_t* out);
int lsk_merkle_root(lsk_ctx* ctx, const uint32_t* const* mats, const size_t* hs, const size_t* ws, size_t nmats,
                    uint32_t out8[8]);

/* A8 / A10 / A12, one stage at a time: the third-party pieces sphinx's prover runs between the
 * commitments, exposed so that each can be held against the oracle on its own.  `chip` indexes the machine's chip
 * vector (build_lair_chip_vector order, src/lair/lair_chip.rs:196-212); matrices are row-major canonical with
 * NATURAL row order on the trace domain; extension elements are 4 canonical words.
 *   lsk_perm_trace: sphinx `generate_permutation_trace` for one chip -- per row and batch of two interactions
 *     sum +-mult / (alpha + kind + sum_k beta^(k+1) value_k), last column = running sum; perm_rows_out is
 *     h x 4*pw (pw = lsk_machine_chip_perm_width), cum_sum_out the chip's cumulative sum.
 *   lsk_quotient: `quotient_values`: Air::eval (src/lair/lair_chip.rs:156-194) + the LogUp constraints folded with
 *     powers of `alpha` on the 2h-point quotient domain g*H_2h, divided by the vanishing polynomial; inputs are the
 *     main and permutation traces on the trace domain (extended inside), output 2h x 4 in commitment (bit-reversed) order.
 *   lsk_fri_prove: TwoAdicFriPcs's FRI on ONE vector of 2^log_n extension values (commitment order): commit phase
 *     (arity-2 folds, one Merkle commitment per layer, the transcript step on the device), final constant, proof-of-
 *     work witness, query indices and openings.  challenger_state = the DuplexChallenger's 16 sponge words with an
 *     empty input buffer.  Words: nlayers, roots[nlayers][8], final_poly[4], pow_witness, nqueries, then per query:
 *     index, per layer { sibling[4], pathlen, digests[pathlen][8] }.  Free with lsk_proof_free. */
int lsk_machine_chip_perm_width(const lsk_machine* m, uint32_t chip);
int lsk_perm_trace(lsk_machine* m, uint32_t chip, const uint32_t* main_rows, size_t h, const uint32_t chal_alpha[4],
                   const uint32_t chal_beta[4], uint32_t* perm_rows_out, uint32_t cum_sum_out[4]);
int lsk_quotient(lsk_machine* m, uint32_t chip, const uint32_t* main_rows, const uint32_t* perm_rows, size_t h,
                 const uint32_t* public_values, const uint32_t cum_sum[4], const uint32_t chal_alpha[4],
                 const uint32_t chal_beta[4], const uint32_t alpha[4], uint32_t* quot_out);
int lsk_fri_prove(lsk_ctx* ctx, const uint32_t* vals, uint32_t log_n, const uint32_t challenger_state[16], uint32_t num_queries,
                  uint32_t pow_bits, uint32_t** words, size_t* len);

/* ---- C++ host layer (csrc/host_lair.h, host_air.cpp): what the reference does in Rust before it can call the
 * prover -- Toplevel::new (expand + compile the `func!` definitions, src/lair/toplevel.rs), FuncChip layout
 * (src/lair/func_chip.rs:90-277) and Chip::new's symbolic run of every Air::eval (src/lair/lair_chip.rs:
 * 196-231).  `surface` is the FuncE list as words (format in host_lair.cpp); results are malloc'd word images
 * (free with lsk_words_free): the program image for lsk_program_load and the 'MACH' blob for lsk_machine_new /
 * lsk_vk_new.  C++ callers use the builder API of host_lair.h directly. */
int lsk_host_compile_toplevel(const uint32_t* surface, size_t n, uint32_t** words, size_t* len);
int lsk_host_chip_vector(const uint32_t* surface, size_t n, uint32_t entry_func, uint32_t** words, size_t* len);
int lsk_host_demo_toplevel(uint32_t** words, size_t* len);   /* demo_toplevel, src/lair/mod.rs:53-112 */
void lsk_words_free(uint32_t* words);

/* ---- column-sharded commit of ONE matrix over G GPUs (BASELINE north star) ---------------
 * The per-GPU pieces; lurk_b200/parallel.py::column_sharded_commit strings them together with NCCL
 * (one all-to-all turning the column partition into a row partition, one all-gather of G digests).
 * `_dev` pointers are DEVICE pointers in the library's internal layout: column-major, rows in bit-reversed
 * order, Montgomery form (what lsk_import_columns_dev produces from a canonical row-major host matrix).
 *   1. every rank extends its w/G columns:            lsk_lde_columns_dev      (N -> 2N rows)
 *   2. all-to-all: rank d receives rows [d 2N/G, (d+1) 2N/G) of every column
 *   3. every rank hashes its rows and builds its subtree: lsk_merkle_subtree_root_dev
 *   4. all-gather of the G subtree roots; top log2 G levels: lsk_merkle_join_roots (host)
 * The result equals lsk_merkle_root of the whole LDE. */
int lsk_import_columns_dev(lsk_ctx* ctx, const uint32_t* rows, size_t h, size_t w, uint32_t* cols_dev);
int lsk_lde_columns_dev(lsk_ctx* ctx, const uint32_t* in_dev, size_t in_stride, uint32_t* out_dev, size_t out_stride,
                        uint32_t logn, uint32_t ncols, int shift_kind);
int lsk_merkle_subtree_root_dev(lsk_ctx* ctx, const uint32_t* cols_dev, size_t stride, uint32_t log_rows, uint32_t w,
                                uint32_t root_out8[8]);
int lsk_merkle_join_roots(const uint32_t* roots, uint32_t count, uint32_t out8[8]);

/* ---- column-sharded PROVE of one shard over G GPUs (BASELINE configs[3]) ---------------------
 * G processes, one GPU and one lsk_ctx each, join a device group; afterwards lsk_prove / lsk_shard_commit /
 * lsk_shard_prove on those contexts, called by every rank with the SAME record and arguments, prove each shard
 * together: the throughput-bound stages (LDE by column block; Merkle leaves and layers, quotient and reduced
 * openings by row range; openings by column block) are split over the group and re-assembled with NCCL
 * all-gathers over NVLink, the transcript is replayed identically by every rank, and every rank returns the
 * same proof -- word for word the proof one GPU produces.  The reference has no such path (its prover is one
 * process, sphinx `LocalProver`, src/core/cli/repl.rs:196); this is the north star's multi-GPU partition.
 * NCCL is loaded at run time (libnccl.so.2).  G must be a power of two.
 *   rank 0: lsk_comm_unique_id(id); the host distributes the 128 bytes (any transport);
 *   every rank: lsk_ctx_comm_init(ctx, rank, G, id)  (collective), ... proves ..., lsk_ctx_comm_destroy(ctx). */
#define LSK_COMM_ID_BYTES 128
int lsk_comm_unique_id(uint8_t out[LSK_COMM_ID_BYTES]);
int lsk_ctx_comm_init(lsk_ctx* ctx, int rank, int size, const uint8_t id[LSK_COMM_ID_BYTES]);
int lsk_ctx_comm_destroy(lsk_ctx* ctx);
/* rank / size of the context's group (0 / 1 without one) and the bytes its collectives delivered to this rank
 * since the last lsk_profile_enable */
int lsk_ctx_comm_info(const lsk_ctx* ctx, int* rank, int* size, double* bytes_received);
/* the contiguous share [begin, end) of n items (columns, rows, Merkle nodes) that `rank` of `size` computes;
 * host only, no GPU needed */
int lsk_comm_partition(uint64_t n, int rank, int size, uint64_t* begin, uint64_t* end);

#ifdef __cplusplus
}
#endif
#endif
