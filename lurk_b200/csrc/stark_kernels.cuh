// Device stages of the shard prover other than LDE and Merkle: LogUp permutation trace (A8),
// quotient evaluation (A6 + A10), openings (A11), FRI reduce / fold / grind / query gathers (A12).
// Everything here replaces third-party sphinx-core / p3-fri code reached from `machine.prove`
// (src/core/cli/repl.rs:196, benches/fib.rs:158); the oracle restates it in oracle/stark_prove.h.
// All matrices are column-major, Montgomery form, rows in bit-reversed order.
#pragma once
#include <memory>
#include <vector>

#include "device.cuh"
#include "field.cuh"
#include "jit.h"
#include "quotient_kernel.cuh"

namespace lsk {

struct ChipDev {  // a chip's compiled AIR, device resident (built once per machine)
  DevBuf<uint32_t> code;    // constraint program, 4 words per instruction (device opcodes, see above)
  DevBuf<uint32_t> consts;  // Montgomery
  DevBuf<uint32_t> inter;   // interaction stream: {kind_mont sign nvalues affine(mult) affine(value)*}*, sends first
  uint32_t ninstr = 0, nint = 0, n_regs = 0, width = 0, prep_width = 0, pw = 0, n_constraints = 0;
  std::shared_ptr<JitKernel> jit;   // the constraint program compiled by NVRTC (jit.cu), or null: interpret `code`
};

void init_stark_kernels();

// ---- LogUp -----------------------------------------------------------------------------------------
// main/prep: N rows (natural domain H_N, bit-reversed order).  perm_out: 4*pw columns of N rows.
// cum_sum_out: one Ext (device).  scratch: N Ext.
void perm_trace_dev(lsk_ctx* ctx, const ChipDev& chip, const uint32_t* main, size_t mstride, const uint32_t* prep,
                    size_t pstride, unsigned logn, Ext chal_alpha, const Ext* beta_pows_dev, uint32_t* perm_out,
                    size_t perm_stride, Ext* scratch, Ext* cum_sum_out);

// ---- quotient --------------------------------------------------------------------------------------
void quotient_dev(lsk_ctx* ctx, const ChipDev& chip, const QuotientArgs& a);

// ---- openings --------------------------------------------------------------------------------------
// wts[r] = barycentric weight of row r (< N) of the coset g*H_N (bit-reversed) for the point z
void bary_weights_dev(lsk_ctx* ctx, unsigned logn, Ext z, Ext* wts);
// out[mat.out_off + c * 2 + p] = sum_r wts_p[r] * m[c][r] for c < w, p < 2, for ALL matrices of one height in
// two launches (a shard has ~5 matrices per chip and most chips are small); one_point: both outputs use
// wts0 (quotient chunks).
constexpr unsigned OPEN_ROWS = 2048;
struct OpenMat {
  const uint32_t* p;
  uint64_t stride;
  uint32_t w, out_off, one_point;
};
void open_batch_dev(lsk_ctx* ctx, const std::vector<OpenMat>& mats, unsigned logn, const Ext* wts0, const Ext* wts1,
                    Ext* out);

// ---- FRI ---------------------------------------------------------------------------------------------
struct ReduceMat {  // one opened matrix of LDE height 2^lh
  const uint32_t* p;
  uint64_t stride;
  uint32_t w, npoints;
  uint32_t open_off, pad_;   // the matrix's opened values: opened[open_off + 2 c + p] (reduce_y_dev fills y from them)
  Ext y[2];    // sum_c alpha^c * opened_value_p[c]
  Ext off[2];  // alpha offset of (matrix, point)
};
// ro[r] = sum over mats, points of off * (y - sum_c alpha^c m[r][c]) / (z_p - x_r), x_r = g * w_{2^lh}^bitrev(r),
// z_0 = zeta, z_1 = zeta * w_{2^(lh-1)}
// y[p] of every matrix of the array from the opened values on the device (one CTA per matrix): the ~14 k extension
// products of a Lurk shard that the host used to do while the device waited
void reduce_y_dev(lsk_ctx* ctx, ReduceMat* mats_dev, uint32_t nmats, const Ext* opened_dev, const Ext* alpha_pows_dev);
void reduce_openings_dev(lsk_ctx* ctx, const ReduceMat* mats_dev, uint32_t nmats, unsigned lh, Ext zeta,
                         const Ext* alpha_pows_dev, Ext* ro, double algorithmic_bytes);
// out[i] = (1/2 + t) cur[2i] + (1/2 - t) cur[2i+1] (+ add[i]),  t = beta / (2 * w_{2^(logh+1)}^bitrev(i))
// beta is read from device memory (written by the device-side transcript step, merkle.cuh::FriChallenge)
void fri_fold_dev(lsk_ctx* ctx, const Ext* cur, unsigned logh, const Ext* beta_dev, const Ext* add, Ext* out);
// smallest witness w such that the duplex challenger, after observing w, samples `bits` zero bits.
// state: 16 canonical words; pending: the challenger's input buffer (n_pending < 8 canonical words)
uint32_t grind_dev(lsk_ctx* ctx, const uint32_t state[16], const uint32_t* pending, unsigned n_pending, unsigned bits);

// ---- query phase ---------------------------------------------------------------------------------------
// Query phase: every FRI query has the same shape, so the host describes ONE query as a list of
// segments and the device instantiates it for each sampled index, writing the proof words of all
// queries (literals included) in their final order.  A segment is `count` words at dst:
//   p == nullptr : the literal `lit`;
//   else         : from_mont(p[first + k * stride]), first = (((index >> shift) ^ xor1) * mul).
// In a device group (comm.cuh) every rank instantiates the same template but writes zeros for the words another
// rank supplies, and the query words are summed over the group: own_rank = QSEG_ALL writes always, QSEG_NONE
// never, r >= 0 only when the source node ((index >> shift) ^ xor1) >> own_shift equals r (distributed Merkle
// layers: the owner of a node also owns its sibling).
constexpr int32_t QSEG_ALL = -1, QSEG_NONE = -2;
struct QuerySeg {
  const uint32_t* p;
  uint64_t stride;
  uint32_t count, dst, shift, xor1, mul, lit;
  uint32_t own_shift;
  int32_t own_rank;
};
// out: nqueries x words_per_query
void query_assemble_dev(lsk_ctx* ctx, const QuerySeg* segs_dev, uint32_t nsegs, const uint32_t* indices_dev,
                        uint32_t nqueries, uint32_t words_per_query, uint32_t* out);

}  // namespace lsk
