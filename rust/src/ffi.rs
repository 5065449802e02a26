//! `extern "C"` view of `include/lurk_stark.h` (liblurkstark.so).  Return codes: 0 ok, negative error
//! (`lsk_last_error()` has the message); field elements cross as canonical `u32`.
#![allow(non_camel_case_types)]
use std::os::raw::c_char;

#[repr(C)] pub struct LskCtx { _p: [u8; 0] }
#[repr(C)] pub struct LskProgram { _p: [u8; 0] }
#[repr(C)] pub struct LskRecord { _p: [u8; 0] }
#[repr(C)] pub struct LskMachine { _p: [u8; 0] }
#[repr(C)] pub struct LskVk { _p: [u8; 0] }

pub const LSK_OK: i32 = 0;
pub const LSK_ERR_VERIFY: i32 = -4;

#[link(name = "lurkstark")]
extern "C" {
    pub fn lsk_last_error() -> *const c_char;
    pub fn lsk_ctx_create(device: i32, out: *mut *mut LskCtx) -> i32;
    pub fn lsk_ctx_destroy(ctx: *mut LskCtx);

    // Toplevel image + chip vector: produced once per toplevel by the host layer (lsk_host_compile_toplevel /
    // lsk_host_chip_vector from a serialised FuncE list, or a Rust port of the same two passes)
    pub fn lsk_host_compile_toplevel(surface: *const u32, n: usize, words: *mut *mut u32, len: *mut usize) -> i32;
    pub fn lsk_host_chip_vector(surface: *const u32, n: usize, entry_func: u32, words: *mut *mut u32, len: *mut usize) -> i32;
    pub fn lsk_words_free(words: *mut u32);
    pub fn lsk_program_load(words: *const u32, n: usize, out: *mut *mut LskProgram) -> i32;
    pub fn lsk_program_free(p: *mut LskProgram);

    // QueryRecord: either the native executor ...
    pub fn lsk_record_new(p: *const LskProgram, out: *mut *mut LskRecord) -> i32;
    pub fn lsk_record_free(r: *mut LskRecord);
    pub fn lsk_execute(p: *const LskProgram, r: *mut LskRecord, func: u32, args: *const u32, nargs: usize,
                       out: *mut u32, nout: usize) -> i32;
    pub fn lsk_inject_inv_queries(p: *const LskProgram, r: *mut LskRecord, func: u32, inps: *const u32,
                                  outs: *const u32, n: usize) -> i32;
    // ... or the reference's own, imported table by table (flatten.rs)
    pub fn lsk_record_import_func(p: *const LskProgram, r: *mut LskRecord, func: u32, n: usize, args: *const u32,
                                  outs: *const u32, provide: *const u32, req_off: *const u32, reqs: *const u32,
                                  depth: *const u32, dreq_off: *const u32, dreqs: *const u32) -> i32;
    pub fn lsk_record_import_mem(r: *mut LskRecord, mem_idx: u32, n: usize, vals: *const u32, provide: *const u32) -> i32;
    pub fn lsk_record_import_bytes(r: *mut LskRecord, records: *const u32) -> i32;
    pub fn lsk_record_import_public_values(r: *mut LskRecord, pv: *const u32, n: usize) -> i32;
    pub fn lsk_record_resolve(p: *const LskProgram, r: *mut LskRecord) -> i32;
    pub fn lsk_record_finalize(r: *mut LskRecord) -> i32;
    pub fn lsk_record_pin(r: *mut LskRecord) -> i32;

    // machine = new_machine + setup
    pub fn lsk_machine_new(ctx: *mut LskCtx, prog: *const LskProgram, entry_func: u32, blob: *const u32, n: usize,
                           num_queries: u32, pow_bits: u32, out: *mut *mut LskMachine) -> i32;
    pub fn lsk_machine_free(m: *mut LskMachine);
    pub fn lsk_machine_prep_commit(m: *const LskMachine, out8: *mut u32) -> i32;
    pub fn lsk_prove(m: *mut LskMachine, rec: *const LskRecord, shard_size: u32, words: *mut *mut u32, len: *mut usize) -> i32;
    pub fn lsk_proof_free(words: *mut u32);
    pub fn lsk_verify(m: *const LskMachine, words: *const u32, len: usize) -> i32;
    pub fn lsk_vk_new(blob: *const u32, n: usize, prep_commit: *const u32, has_prep: i32, num_queries: u32,
                      pow_bits: u32, out: *mut *mut LskVk) -> i32;
    pub fn lsk_vk_verify(vk: *const LskVk, words: *const u32, len: usize) -> i32;
    pub fn lsk_vk_free(vk: *mut LskVk);

    // one shard per GPU / several GPUs on one shard
    pub fn lsk_shard_commit(m: *mut LskMachine, rec: *const LskRecord, shard: u32, shard_size: u32, root_out: *mut u32) -> i32;
    pub fn lsk_shard_prove(m: *mut LskMachine, rec: *const LskRecord, shard: u32, all_roots: *const u32, nshards: u32,
                           words: *mut *mut u32, len: *mut usize) -> i32;
    pub fn lsk_comm_unique_id(out: *mut u8) -> i32;
    pub fn lsk_ctx_comm_init(ctx: *mut LskCtx, rank: i32, size: i32, id: *const u8) -> i32;
    pub fn lsk_ctx_comm_destroy(ctx: *mut LskCtx) -> i32;

    // ZStore batch hashing (src/core/zstore.rs:305-333)
    pub fn lsk_poseidon2_hash(ctx: *mut LskCtx, width: i32, input: *const u32, n: usize, out8: *mut u32) -> i32;
}

pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(lsk_last_error()).to_string_lossy().into_owned() }
}
/// the reference's convention for internal errors is to panic (`expect` / `assert!` throughout src/lair/trace.rs)
pub fn check(rc: i32) {
    if rc != LSK_OK {
        panic!("liblurkstark error {rc}: {}", last_error());
    }
}
