//! `flatten`: the reference's `QueryRecord<F>` (src/lair/execute.rs:26-33,67-75) -> an `lsk_record`, table by table.
//! The Lurk evaluator and `toplevel.execute` stay exactly as they are; this is one linear pass over the maps.
//! The per-call "hints" (which callee / memory row a Call, PreImg or Store refers to: the lookups
//! src/lair/trace.rs:325-373 repeats at trace-generation time) are recomputed inside the library by
//! `lsk_record_resolve`, which also checks that every imported row returns its recorded output.
use p3_baby_bear::BabyBear as F;
use p3_field::PrimeField32;

use crate::air::builder::Record;
use crate::cuda::ffi::*;
use crate::gadgets::bytes::ByteInput;
use crate::lair::execute::{mem_index_from_len, QueryRecord, QueryResult, MEM_TABLE_SIZES};
use crate::lair::{chipset::Chipset, toplevel::Toplevel, List};

pub struct FlatRecord { pub raw: *mut LskRecord }
impl Drop for FlatRecord { fn drop(&mut self) { unsafe { lsk_record_free(self.raw) } } }

fn canon(xs: &[F], out: &mut Vec<u32>) { out.extend(xs.iter().map(|x| x.as_canonical_u32())); }
fn rec2(r: &Record, out: &mut Vec<u32>) { out.push(r.nonce); out.push(r.count); }

pub fn flatten<C1: Chipset<F>, C2: Chipset<F>>(
    toplevel: &Toplevel<F, C1, C2>, prog: *const LskProgram, q: &QueryRecord<F>,
) -> FlatRecord {
    let mut raw = std::ptr::null_mut();
    check(unsafe { lsk_record_new(prog, &mut raw) });
    let rec = FlatRecord { raw };

    // func_queries[i]: IndexMap in insertion order; the index IS the nonce (execute.rs:41-43,500-561)
    for (func_idx, map) in q.func_queries.iter().enumerate() {
        if map.is_empty() { continue; }
        let func = toplevel.func_by_index(func_idx);
        let (mut args, mut outs, mut provide) = (vec![], vec![], vec![]);
        let (mut req_off, mut reqs) = (vec![0u32], vec![]);
        let (mut depth, mut dreq_off, mut dreqs) = (vec![], vec![0u32], vec![]);
        for (inp, res) in map.iter() {
            let res: &QueryResult<F> = res;
            canon(inp, &mut args);
            canon(res.expect_output(), &mut outs);
            rec2(&res.provide, &mut provide);
            res.requires.iter().for_each(|r| rec2(r, &mut reqs));
            req_off.push((reqs.len() / 2) as u32);
            if func.partial {
                depth.push(res.depth);
                res.depth_requires.iter().for_each(|r| rec2(r, &mut dreqs));
                dreq_off.push((dreqs.len() / 2) as u32);
            }
        }
        let opt = |v: &Vec<u32>| if func.partial { v.as_ptr() } else { std::ptr::null() };
        check(unsafe {
            lsk_record_import_func(prog, rec.raw, func_idx as u32, map.len(), args.as_ptr(), outs.as_ptr(),
                                   provide.as_ptr(), req_off.as_ptr(), reqs.as_ptr(), opt(&depth), opt(&dreq_off), opt(&dreqs))
        });
    }
    // mem_queries[k]: values -> QueryResult whose `provide` is the memory provide record; pointer = index + 1
    for (k, map) in q.mem_queries.iter().enumerate() {
        if map.is_empty() { continue; }
        debug_assert_eq!(mem_index_from_len(MEM_TABLE_SIZES[k]), k);
        let (mut vals, mut provide) = (vec![], vec![]);
        for (v, res) in map.iter() { canon(v, &mut vals); rec2(&res.provide, &mut provide); }
        check(unsafe { lsk_record_import_mem(rec.raw, k as u32, map.len(), vals.as_ptr(), provide.as_ptr()) });
    }
    // BytesRecord: BTreeMap<ByteInput, BytesInputRecord>; dense [65536][6][2] for the library
    if !q.bytes.is_empty() {
        let mut dense = vec![0u32; 65536 * 12];
        for input in 0..=u16::MAX {
            if let Some(r) = q.bytes.get(ByteInput::from_u16(input)) {
                for (op, rec) in r.iter_records().into_iter().enumerate() {
                    dense[input as usize * 12 + 2 * op] = rec.nonce;
                    dense[input as usize * 12 + 2 * op + 1] = rec.count;
                }
            }
        }
        check(unsafe { lsk_record_import_bytes(rec.raw, dense.as_ptr()) });
    }
    let mut pv = vec![];
    canon(q.expect_public_values(), &mut pv);
    check(unsafe { lsk_record_import_public_values(rec.raw, pv.as_ptr(), pv.len()) });
    // preimages that were injected rather than computed (inject_inv_queries, execute.rs:299-319): the ones
    // computed by an imported query are rebuilt from the tables
    for (func_idx, inv) in q.inv_func_queries.iter().enumerate() {
        let Some(inv) = inv else { continue };
        let (mut inps, mut outs, mut n) = (vec![], vec![], 0usize);
        for (out, inp) in inv.iter() {
            let out: &List<F> = out;
            if q.func_queries[func_idx].contains_key(inp) { continue; }
            canon(inp, &mut inps); canon(out, &mut outs); n += 1;
        }
        if n > 0 {
            check(unsafe { lsk_inject_inv_queries(prog, rec.raw, func_idx as u32, inps.as_ptr(), outs.as_ptr(), n) });
        }
    }
    check(unsafe { lsk_record_resolve(prog, rec.raw) });
    check(unsafe { lsk_record_pin(rec.raw) });      // page-lock: lsk_prove uploads at PCIe speed
    rec
}
