// The shard prover: host orchestration of the device stages behind the C ABI's lsk_machine_* /
// lsk_prove entry points.
//
// Replaces (third-party, reached from the reference at the cited call sites; restated by the
// oracle in oracle/stark_prove.h):
//   StarkMachine::new / setup      src/core/stark_machine.rs:20-30, src/core/cli/repl.rs:175-176
//   StarkMachine::prove<LocalProver>  src/core/cli/repl.rs:194-196, benches/fib.rs:154-158,
//                                     src/lair/lair_chip.rs:258-275, tests/fib.rs:90-94
//   MachineAir::{included, generate_trace}  src/lair/lair_chip.rs:96-139 (trace_kernels.cu)
// Per shard: traces -> coset LDE -> Merkle commit (main), LogUp permutation traces -> LDE ->
// commit, quotient -> LDE of the two chunks -> commit, openings at zeta / zeta*w, FRI.  The
// transcript (challenger.h) runs on the host; everything that touches a trace cell runs on the
// device and nothing but commitments, opened values and query answers is copied back.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>

#include "challenger.h"
#include "comm.cuh"
#include "machine.h"
#include "merkle.cuh"
#include "ntt.cuh"
#include "stark_kernels.cuh"
#include "trace.cuh"
#include "verify.h"

namespace lsk {

constexpr uint32_t PROOF_MAGIC = 0x4c505246u;  // 'LPRF'
constexpr unsigned LOG_BLOWUP = 1;

// LSK_PHASES=1: host wall-clock per prover phase on stderr (synchronises at phase boundaries)
struct PhaseTimer {
  lsk_ctx* ctx;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(lsk_ctx* c) : ctx(c), on(std::getenv("LSK_PHASES") != nullptr) { t0 = std::chrono::steady_clock::now(); }
  void mark(const char* name) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[lsk phase] %-18s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

struct ShardState {  // phase-1 result of one shard, kept until its proof is written
  uint32_t shard = 0;
  std::vector<uint32_t> chip_ids;  // chip_ordering: height descending, stable
  std::vector<unsigned> logn;
  std::vector<DevBuf<uint32_t>> main, main_lde;
  MerkleTreeDev main_tree;
  uint32_t main_root[8];  // canonical
};

}  // namespace lsk

using namespace lsk;

struct lsk_machine {
  lsk_ctx* ctx = nullptr;
  const lsk_program* prog = nullptr;
  uint32_t entry = 0, num_pv = 0;
  unsigned num_queries = 100, pow_bits = 16;
  std::vector<ChipH> chips;
  std::vector<ChipDev> dchips;
  HostPerm16 perm;
  // setup: the only preprocessed trace is the byte table's (src/lair/lair_chip.rs:141-153)
  int prep_chip = -1;
  DevBuf<uint32_t> prep_trace, prep_lde;
  MerkleTreeDev prep_tree;
  uint32_t prep_commit[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // two-phase API state
  std::map<uint32_t, std::unique_ptr<ShardState>> pending;
  // device copies of the record: key 0xffffffff = the whole record, else the slice of that shard
  std::map<uint32_t, std::unique_ptr<DeviceRecord>> drecs;
  const lsk_record* drec_for = nullptr;
  uint32_t drec_ss = 0;
  size_t last_h2d = 0, last_d2h = 0;
};

namespace lsk {

// ---- machine construction -------------------------------------------------------------------------------
static void push_affine(std::vector<uint32_t>& o, const AffineH& a, uint32_t width, uint32_t prep_width) {
  o.push_back((uint32_t)a.terms.size());
  o.push_back(to_mont(a.cst % P));
  for (auto& t : a.terms) {
    if (t[0] > 1 || t[1] >= (t[0] ? prep_width : width)) fail(LSK_ERR_ARG, "interaction references a bad column");
    o.push_back(t[0]);
    o.push_back(t[1]);
    o.push_back(to_mont(t[2] % P));
  }
}
static void compile_chip(lsk_ctx* ctx, const ChipH& c, uint32_t num_pv, ChipDev& d) {
  d.ninstr = (uint32_t)(c.code.size() / 4);
  d.nint = (uint32_t)c.n_interactions();
  d.n_regs = c.n_regs;
  d.width = c.width;
  d.prep_width = c.prep_width;
  d.pw = (uint32_t)c.perm_width();
  d.n_constraints = c.n_constraints;
  // validate the constraint program so the kernel can run it unchecked
  uint32_t asserts = 0;
  auto check_operand = [&](uint32_t w) {
    uint32_t idx = w & 0x0fffffffu, lim;
    switch (w >> 28) {
      case 0: lim = c.n_regs; break;
      case 1: lim = (uint32_t)c.consts.size(); break;
      case 2: case 3: lim = c.width; break;
      case 4: case 5: lim = c.prep_width; break;
      case 6: lim = num_pv; break;
      case 7: lim = 3; break;
      default: lim = 0;
    }
    if (idx >= lim) fail(LSK_ERR_ARG, "constraint program of %s has an out-of-range operand", c.name.c_str());
  };
  for (size_t i = 0; i < c.code.size(); i += 4) {
    uint32_t op = c.code[i];
    if (op < COP_ADD || op > COP_ASSERT) fail(LSK_ERR_ARG, "bad constraint opcode in %s", c.name.c_str());
    if (op != COP_ASSERT && c.code[i + 1] >= c.n_regs) fail(LSK_ERR_ARG, "bad destination register in %s", c.name.c_str());
    check_operand(c.code[i + 2]);
    if (op <= COP_MUL) check_operand(c.code[i + 3]);
    asserts += op == COP_ASSERT;
  }
  if (asserts != c.n_constraints) fail(LSK_ERR_ARG, "constraint count mismatch in %s", c.name.c_str());
  std::vector<uint32_t> consts(c.consts.size());
  for (size_t i = 0; i < consts.size(); i++) consts[i] = to_mont(c.consts[i] % P);
  std::vector<uint32_t> inter;
  for (int s = 0; s < 2; s++)
    for (auto& it : (s == 0 ? c.sends : c.receives)) {
      inter.push_back(to_mont(it.kind % P));
      inter.push_back((uint32_t)s);
      inter.push_back((uint32_t)it.values.size());
      push_affine(inter, it.mult, c.width, c.prep_width);
      for (auto& v : it.values) push_affine(inter, v, c.width, c.prep_width);
    }
  // device form: fold the operand kinds of add / sub / mul into the opcode where both operands are a
  // register, a main-trace cell of the row or a constant (stark_kernels.cuh)
  std::vector<uint32_t> code = c.code;
  auto kind_of = [](uint32_t w) -> int {  // tag 0 reg, 1 const, 2 main local
    switch (w >> 28) { case 0: return 0; case 2: return 1; case 1: return 2; default: return -1; }
  };
  for (size_t i = 0; i < code.size(); i += 4) {
    uint32_t op = code[i], a = code[i + 2], b = code[i + 3];
    if (op >= COP_ADD && op <= COP_MUL) {
      int ka = kind_of(a), kb = kind_of(b);
      if (ka >= 0 && kb >= 0) {
        code[i] = 9 * (op - COP_ADD) + 3 * (uint32_t)ka + (uint32_t)kb;
        code[i + 2] = a & 0x0fffffffu;
        code[i + 3] = b & 0x0fffffffu;
      } else {
        code[i] = op == COP_ADD ? QOP_ADD : op == COP_SUB ? QOP_SUB : QOP_MUL;
      }
    } else if (op == COP_NEG) {
      code[i] = QOP_NEG;
    } else {  // COP_ASSERT
      if ((a >> 28) == 0) code[i] = QOP_ASSERT_REG, code[i + 2] = a & 0x0fffffffu;
      else code[i] = QOP_ASSERT;
    }
  }
  if (code.empty()) code.assign(4, 0), d.ninstr = 0;
  if (consts.empty()) consts.push_back(0);
  if (inter.empty()) inter.push_back(0);
  d.code.alloc(ctx, code.size()); d.code.upload(code.data(), code.size());
  d.consts.alloc(ctx, consts.size()); d.consts.upload(consts.data(), consts.size());
  d.inter.alloc(ctx, inter.size()); d.inter.upload(inter.data(), inter.size());
  // long programs (the hasher chips): machine code from NVRTC when the policy / cubin cache provides it (jit.cu)
  d.jit.reset(jit_quotient(c), jit_free);
}

static void machine_setup(lsk_machine& m) {
  lsk_ctx* ctx = m.ctx;
  for (size_t i = 0; i < m.chips.size(); i++)
    if (m.chips[i].prep_width) {
      if (m.prep_chip >= 0 || m.chips[i].kind != CHIP_BYTES || m.chips[i].prep_width != 6)
        fail(LSK_ERR_ARG, "only the byte table may have a preprocessed trace");
      m.prep_chip = (int)i;
    }
  if (m.prep_chip < 0) return;
  const size_t n = 65536;
  m.prep_trace.alloc(ctx, n * 6);
  bytes_preprocessed_dev(ctx, TraceLayout::ColMajorBitrevMont, m.prep_trace.p);
  m.prep_lde.alloc(ctx, 2 * n * 6);
  lde_dev(ctx, m.prep_trace.p, n, m.prep_lde.p, 2 * n, 16, 6, LdeShift::Trace);
  merkle_commit_dev(ctx, {MatRef{m.prep_lde.p, 2 * n, 17, 6}}, m.prep_tree);
  merkle_root_host(ctx, m.prep_tree, m.prep_commit);
}

// ---- phase 1: traces + main commitment ---------------------------------------------------------------
static bool chip_included(const lsk_machine& m, const ChipH& c, const lsk_record& rec, uint32_t shard, uint32_t ss) {
  switch (c.kind) {  // MachineAir::included, src/lair/lair_chip.rs:124-139
    case CHIP_FUNC: {
      size_t n = rec.funcs[c.index].rows();
      return (size_t)shard * ss < n;
    }
    case CHIP_MEM: case CHIP_ENTRY: return shard == 0;
    default: return true;
  }
}

// whole = false: the caller proves one shard of a sharded record (lsk_shard_commit): upload its slice only
static DeviceRecord& device_record(lsk_machine& m, const lsk_record& rec, uint32_t shard, uint32_t ss, bool whole) {
  const bool slice = !whole && record_row_ordered(rec) && lsk_record_num_shards(&rec, ss) > 1;
  const uint32_t key = slice ? shard : 0xffffffffu;
  if (m.drec_for != &rec || m.drec_ss != ss) {
    m.drecs.clear();
    m.drec_for = &rec;
    m.drec_ss = ss;
  }
  auto& slot = m.drecs[key];
  if (!slot) {
    slot = std::make_unique<DeviceRecord>();
    slot->upload(m.ctx, *m.prog, rec, slice ? shard : 0, slice ? ss : 0);
    m.last_h2d += slot->h2d_bytes;
  }
  return *slot;
}

// lde_dev of a whole matrix; in a device group (comm.cuh) each rank extends a block of the columns of a tall
// matrix and the blocks are all-gathered (columns are contiguous in the column-major LDE buffer)
static void lde_matrix(lsk_ctx* ctx, const uint32_t* in, size_t in_stride, uint32_t* out, size_t out_stride, unsigned logn,
                       uint32_t w, LdeShift kind, LdeBatch* batch) {
  if (!comm_splits(ctx, logn + 1)) return lde_dev(ctx, in, in_stride, out, out_stride, logn, w, kind, batch);
  size_t c0, c1;
  comm_range(ctx, w, &c0, &c1);
  lde_dev(ctx, in + c0 * in_stride, in_stride, out + c0 * out_stride, out_stride, logn, (uint32_t)(c1 - c0), kind, nullptr);
  comm_allgather_columns(ctx, out, out_stride, w);
}

static std::unique_ptr<ShardState> commit_shard(lsk_machine& m, const lsk_record& rec, uint32_t shard, uint32_t ss,
                                                bool whole_record) {
  lsk_ctx* ctx = m.ctx;
  PhaseTimer pt(ctx);
  DeviceRecord& drec = device_record(m, rec, shard, ss, whole_record);
  pt.mark("record upload");
  if (rec.public_values.size() != m.num_pv) fail(LSK_ERR_ARG, "record has %zu public values, machine expects %u",
                                                 rec.public_values.size(), m.num_pv);
  struct Item { uint32_t chip; unsigned logn; };
  std::vector<Item> items;
  for (uint32_t i = 0; i < m.chips.size(); i++) {
    const ChipH& c = m.chips[i];
    if (!chip_included(m, c, rec, shard, ss)) continue;
    size_t h = 1;
    switch (c.kind) {
      case CHIP_FUNC: h = lsk_func_trace_height(&rec, c.index, shard, ss); break;
      case CHIP_MEM: h = lsk_mem_trace_height(&rec, c.index); break;
      case CHIP_BYTES: h = 65536; break;
      default: h = 1;
    }
    items.push_back({i, log2_exact(h)});
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.logn > b.logn; });
  auto sd = std::make_unique<ShardState>();
  sd->shard = shard;
  std::vector<MatRef> refs;
  LdeBatch small_ldes;
  for (auto& it : items) {
    const ChipH& c = m.chips[it.chip];
    size_t n = (size_t)1 << it.logn;
    DevBuf<uint32_t> tr(ctx, n * c.width), lde(ctx, 2 * n * c.width);
    switch (c.kind) {
      case CHIP_FUNC: {
        if (c.index >= m.prog->nfuncs || m.prog->width(c.index) != c.width) fail(LSK_ERR_ARG, "func chip / program mismatch");
        size_t rows = rec.funcs[c.index].rows();
        size_t s = (size_t)shard * ss, e = std::min(rows, (size_t)(shard + 1) * ss);
        func_trace_dev(ctx, *m.prog, drec, c.index, (uint32_t)s, (uint32_t)(e - s), (uint32_t)n,
                       TraceLayout::ColMajorBitrevMont, tr.p);
      } break;
      case CHIP_MEM:
        if (c.index >= NUM_MEM_TABLES || c.width != 4 + rec.mems[c.index].len) fail(LSK_ERR_ARG, "bad memory chip");
        mem_trace_dev(ctx, drec, c.index, (uint32_t)n, TraceLayout::ColMajorBitrevMont, tr.p);
        break;
      case CHIP_BYTES:
        if (c.width != 13) fail(LSK_ERR_ARG, "bad byte chip");
        bytes_trace_dev(ctx, drec, shard == 0 && rec.bytes_used, TraceLayout::ColMajorBitrevMont, tr.p);
        break;
      default: {  // Entrypoint: one row of public values (src/lair/lair_chip.rs:112-118)
        if (c.width != m.num_pv) fail(LSK_ERR_ARG, "bad entrypoint chip");
        std::vector<uint32_t> pv(m.num_pv);
        for (uint32_t k = 0; k < m.num_pv; k++) pv[k] = to_mont(rec.public_values[k]);
        tr.upload(pv.data(), pv.size());
      }
    }
    lde_matrix(ctx, tr.p, n, lde.p, 2 * n, it.logn, c.width, LdeShift::Trace, &small_ldes);
    refs.push_back(MatRef{lde.p, 2 * n, it.logn + 1, c.width});
    sd->chip_ids.push_back(it.chip);
    sd->logn.push_back(it.logn);
    sd->main.push_back(std::move(tr));
    sd->main_lde.push_back(std::move(lde));
  }
  if (refs.empty()) fail(LSK_ERR_ARG, "shard %u has no chips", shard);
  drec.wait_all(ctx);   // tables of funcs without rows in this shard included: nothing of the upload outlives this point
  lde_flush(ctx, small_ldes);
  pt.mark("trace + main LDE");
  merkle_commit_dev(ctx, refs, sd->main_tree);
  merkle_root_host(ctx, sd->main_tree, sd->main_root);
  pt.mark("main commit");
  return sd;
}

// The proof under construction: words in a malloc'd buffer that is handed to the caller as it is (lsk_proof_free =
// free) -- a std::vector would be copied once more, and fresh pages faulted in twice, for the 2.6 MB of a Lurk shard proof.
struct ProofWords {
  uint32_t* p = nullptr;
  size_t n = 0, cap = 0;
  ProofWords() = default;
  ProofWords(const ProofWords&) = delete;
  ProofWords& operator=(const ProofWords&) = delete;
  ~ProofWords() { free(p); }
  void reserve(size_t want) {
    if (want <= cap) return;
    size_t c = std::max<size_t>(want, cap * 2 + 64);
    uint32_t* q = (uint32_t*)realloc(p, c * sizeof(uint32_t));
    if (!q) fail(LSK_ERR_ARG, "out of memory");
    p = q, cap = c;
  }
  void push_back(uint32_t v) {
    if (n == cap) reserve(n + 1);
    p[n++] = v;
  }
  void append(const uint32_t* src, size_t count) {
    reserve(n + count);
    memcpy(p + n, src, count * sizeof(uint32_t));
    n += count;
  }
  size_t size() const { return n; }
  uint32_t* release() { uint32_t* q = p; p = nullptr; n = cap = 0; return q; }
};
static void push_ext(ProofWords& w, const Ext& e) {
  for (int k = 0; k < 4; k++) w.push_back(from_mont(e.c[k]));
}

// ---- phase 2: everything after the main commitment --------------------------------------------------------
static void prove_shard(lsk_machine& m, const lsk_record& rec, ShardState& sd, Challenger ch, ProofWords& W) {
  lsk_ctx* ctx = m.ctx;
  cudaStream_t st = ctx->stream;
  const size_t nchips = sd.chip_ids.size();
  PhaseTimer pt(ctx);
  Ext chal_alpha = ch.sample_ext(), chal_beta = ch.sample_ext();

  size_t max_vals = 1, max_total = 1, max_w = 8, max_n = 1;
  for (size_t i = 0; i < nchips; i++) {
    const ChipH& c = m.chips[sd.chip_ids[i]];
    max_vals = std::max(max_vals, c.max_values());
    max_total = std::max(max_total, (size_t)c.n_constraints + c.n_perm_constraints());
    max_w = std::max({max_w, (size_t)c.width, 4 * c.perm_width(), (size_t)c.prep_width});
    max_n = std::max(max_n, (size_t)1 << sd.logn[i]);
  }
  auto upload_pows = [&](const Ext& base, size_t n, bool from_one, std::vector<Ext>& host, DevBuf<Ext>& dev) {
    host.resize(n);
    Ext p = from_one ? ext_one() : base;
    for (size_t i = 0; i < n; i++) host[i] = p, p = emul(p, base);
    dev.alloc(ctx, n);
    dev.upload(host.data(), n);
  };
  std::vector<Ext> h_beta, h_alpha, h_fri;
  DevBuf<Ext> d_beta, d_alpha, d_fri;
  upload_pows(chal_beta, max_vals, false, h_beta, d_beta);

  // ---- LogUp permutation traces, LDE, commit
  std::vector<DevBuf<uint32_t>> perm_lde(nchips);
  DevBuf<Ext> cums(ctx, nchips), scratch(ctx, max_n);
  std::vector<MatRef> refs;
  LdeBatch small_ldes;
  std::vector<DevBuf<uint32_t>> small_perms;  // natural-domain traces of queued small LDEs, alive until the flush
  // A shard of a Lurk machine has a dozen chips with a handful of rows (hash3/5, the u64 chips, ...): their
  // permutation-trace and quotient launches are ONE CTA walking a long straight-line program (0.05-0.5 ms each, 3.5 ms
  // per prove) while 147 SMs idle.  They go to the context's side stream and run beside the big chips' launches; their
  // LDEs are queued anyway (lde_flush), after the join.  Not in a device group (one stream orders the collectives).
  const bool use_side = !comm_active(ctx) && !getenv("LSK_NO_SIDE_STREAM");
  constexpr unsigned SIDE_MAX_LOG = 8;
  SideStream side(ctx);
  DevBuf<Ext> scratch_side(ctx, (size_t)1 << SIDE_MAX_LOG);
  std::vector<DevBuf<uint32_t>> side_inputs;   // buffers side-stream kernels read: released only after the join
  // the tiny chips first: the side stream waits (once) for what the main stream has queued so far, and every buffer its
  // kernels use comes from the pool as it stood then
  refs.resize(nchips);
  for (int pass = 0; pass < 2; pass++)
    for (size_t i = 0; i < nchips; i++) {
      const bool on_side = use_side && sd.logn[i] <= SIDE_MAX_LOG;
      if (on_side != (pass == 0)) continue;
      const ChipH& c = m.chips[sd.chip_ids[i]];
      const ChipDev& d = m.dchips[sd.chip_ids[i]];
      size_t n = (size_t)1 << sd.logn[i];
      uint32_t pcols = 4 * d.pw;
      DevBuf<uint32_t> perm(ctx, n * pcols);
      perm_lde[i].alloc(ctx, 2 * n * pcols);
      if (on_side) side.enter();
      perm_trace_dev(ctx, d, sd.main[i].p, n, c.prep_width ? m.prep_trace.p : nullptr, n, sd.logn[i], chal_alpha, d_beta.p,
                     perm.p, n, on_side ? scratch_side.p : scratch.p, cums.p + i);
      if (on_side) side.leave();
      lde_matrix(ctx, perm.p, n, perm_lde[i].p, 2 * n, sd.logn[i], pcols, LdeShift::Trace, &small_ldes);
      refs[i] = MatRef{perm_lde[i].p, 2 * n, sd.logn[i] + 1, pcols};
      if (sd.logn[i] <= 12) small_perms.push_back(std::move(perm));
      if (on_side) side_inputs.push_back(std::move(sd.main[i]));
      else sd.main[i].release();  // the natural-domain trace is not needed past this point
    }
  side.join();
  side_inputs.clear();
  lde_flush(ctx, small_ldes);
  small_perms.clear();
  pt.mark("perm trace + LDE");
  MerkleTreeDev perm_tree;
  merkle_commit_dev(ctx, refs, perm_tree);
  uint32_t perm_root[8];
  merkle_root_host(ctx, perm_tree, perm_root);
  for (int k = 0; k < 8; k++) ch.observe(to_mont(perm_root[k]));
  Ext alpha = ch.sample_ext();
  pt.mark("perm commit");

  // ---- quotient, LDE of the two chunks, commit
  upload_pows(alpha, max_total, true, h_alpha, d_alpha);
  DevBuf<uint32_t> d_pv(ctx, std::max<size_t>(1, m.num_pv));
  {
    std::vector<uint32_t> pv(std::max<size_t>(1, m.num_pv), 0);
    for (uint32_t k = 0; k < m.num_pv; k++) pv[k] = to_mont(rec.public_values[k]);
    d_pv.upload(pv.data(), pv.size());
  }
  std::vector<DevBuf<uint32_t>> quot_lde(2 * nchips);
  refs.clear();
  {
    DevBuf<uint32_t> qv(ctx, 2 * max_n * 4);
    LdeBatch small_ldes;
    std::vector<DevBuf<uint32_t>> small_qv;  // quotient values of the queued small LDEs
    refs.resize(2 * nchips);
    for (int pass = 0; pass < 2; pass++)
      for (size_t i = 0; i < nchips; i++) {
        const bool on_side = use_side && sd.logn[i] <= SIDE_MAX_LOG;   // tiny chips first, on the side stream (see above)
        if (on_side != (pass == 0)) continue;
        const ChipH& c = m.chips[sd.chip_ids[i]];
        const ChipDev& d = m.dchips[sd.chip_ids[i]];
        size_t n = (size_t)1 << sd.logn[i];
        uint32_t* qvp = qv.p;
        if (sd.logn[i] <= 12) {
          small_qv.emplace_back(ctx, 2 * n * 4);
          qvp = small_qv.back().p;
        }
        QuotientArgs a{};
        a.main = sd.main_lde[i].p; a.mstride = 2 * n;
        a.prep = c.prep_width ? m.prep_lde.p : nullptr; a.pstride = 2 * n;
        a.perm = perm_lde[i].p; a.permstride = 2 * n;
        a.pv = d_pv.p;
        a.apow = d_alpha.p;
        a.n_total = (uint32_t)(c.n_constraints + c.n_perm_constraints());
        a.beta_pows = d_beta.p;
        a.cum_sum = cums.p + i;
        a.chal_alpha = chal_alpha;
        a.logn = sd.logn[i];
        a.out = qvp; a.ostride = 2 * n;
        if (on_side) side.enter();
        quotient_dev(ctx, d, a);
        if (on_side) side.leave();
        for (int j = 0; j < 2; j++) {
          quot_lde[2 * i + j].alloc(ctx, 2 * n * 4);
          lde_matrix(ctx, qvp + (size_t)j * n, 2 * n, quot_lde[2 * i + j].p, 2 * n, sd.logn[i], 4,
                     j ? LdeShift::QuotChunk1 : LdeShift::QuotChunk0, &small_ldes);
          refs[2 * i + j] = MatRef{quot_lde[2 * i + j].p, 2 * n, sd.logn[i] + 1, 4};
        }
      }
    side.join();
    lde_flush(ctx, small_ldes);
  }
  pt.mark("quotient + LDE");
  MerkleTreeDev quot_tree;
  merkle_commit_dev(ctx, refs, quot_tree);
  uint32_t quot_root[8];
  merkle_root_host(ctx, quot_tree, quot_root);
  for (int k = 0; k < 8; k++) ch.observe(to_mont(quot_root[k]));
  Ext zeta = ch.sample_ext();
  pt.mark("quotient commit");

  // ---- openings at zeta and zeta * w_N
  // layout of `opened` per chip: prep | main | perm | quot0 | quot1, each as [col][point]
  struct OpenOff { size_t prep, main, perm, quot[2]; };
  std::vector<OpenOff> oo(nchips);
  // round-major (all preprocessed, all main, all permutation, all quotient values): the order the transcript observes
  // them in, so that a round's values are one contiguous range that can leave the device as soon as it is complete
  size_t nopen = 0, round_end[3];
  for (size_t i = 0; i < nchips; i++) oo[i].prep = nopen, nopen += 2 * m.chips[sd.chip_ids[i]].prep_width;
  for (size_t i = 0; i < nchips; i++) oo[i].main = nopen, nopen += 2 * m.chips[sd.chip_ids[i]].width;
  round_end[0] = nopen;
  for (size_t i = 0; i < nchips; i++) oo[i].perm = nopen, nopen += 2 * 4 * m.chips[sd.chip_ids[i]].perm_width();
  round_end[1] = nopen;
  for (size_t i = 0; i < nchips; i++) {
    oo[i].quot[0] = nopen; nopen += 8;
    oo[i].quot[1] = nopen; nopen += 8;
  }
  round_end[2] = nopen;
  DevBuf<Ext> opened(ctx, nopen);
  Ext* h_open = (Ext*)pinned_scratch(ctx, 0, (nopen + nchips) * sizeof(Ext));   // page-locked: the copies below are asynchronous
  Ext* h_cums = h_open + nopen;
  struct Events {   // (round, log height) -> its values have reached the host
    std::map<std::pair<int, unsigned>, cudaEvent_t> by_group;
    cudaEvent_t by_round[3] = {nullptr, nullptr, nullptr};   // device group: a round arrives as a whole
    cudaEvent_t last = nullptr;
    ~Events() {
      for (auto& kv : by_group) cudaEventDestroy(kv.second);
      for (cudaEvent_t e : by_round) if (e) cudaEventDestroy(e);
      if (last) cudaEventDestroy(last);
    }
  } open_done;
  const bool grouped = comm_active(ctx);
  if (grouped) opened.zero();   // every value is computed by exactly one rank and summed over the group
  {
    // one batch (two launches) per distinct height: all matrices of the chips of that height
    std::map<unsigned, std::vector<OpenMat>> by_logn;
    for (size_t i = 0; i < nchips; i++) {
      const ChipH& c = m.chips[sd.chip_ids[i]];
      const size_t n2 = (size_t)2 << sd.logn[i];
      auto& v = by_logn[sd.logn[i]];
      if (c.prep_width) v.push_back(OpenMat{m.prep_lde.p, n2, c.prep_width, (uint32_t)oo[i].prep, 0});
      v.push_back(OpenMat{sd.main_lde[i].p, n2, c.width, (uint32_t)oo[i].main, 0});
      v.push_back(OpenMat{perm_lde[i].p, n2, 4 * (uint32_t)c.perm_width(), (uint32_t)oo[i].perm, 0});
      for (int j = 0; j < 2; j++) v.push_back(OpenMat{quot_lde[2 * i + j].p, n2, 4, (uint32_t)oo[i].quot[j], 1});
    }
    // Weights of every height first (in a device group the tall ones are a collective: every rank calls), then the
    // openings round by round (preprocessed + main, permutation, quotient), tallest chips first = the order the
    // transcript observes them in.  Each (round, height) group's values are copied to page-locked memory behind its
    // kernels: the host feeds them to the transcript -- 7 000 dependent Poseidon2 permutations for a Lurk shard, ~5 ms
    // on a core -- while the device computes the next group.  In a device group every value is computed by one rank
    // (tall matrices: a block of the concatenated column list per rank; the rest: rank 0) and a round is summed over the
    // group before it leaves the device.
    std::map<unsigned, std::pair<DevBuf<Ext>, DevBuf<Ext>>> wts;
    for (auto& kv : by_logn) {
      auto& w = wts[kv.first];
      w.first.alloc(ctx, (size_t)1 << kv.first);
      w.second.alloc(ctx, (size_t)1 << kv.first);
      bary_weights_dev(ctx, kv.first, zeta, w.first.p);
      bary_weights_dev(ctx, kv.first, emul_base(zeta, two_adic_generator(kv.first)), w.second.p);
    }
    auto new_event = [&]() {
      cudaEvent_t ev = nullptr;
      LSK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      LSK_CUDA(cudaEventRecord(ev, st));
      return ev;
    };
    size_t lo = 0;
    for (int r = 0; r < 3; r++) {
      const size_t hi = round_end[r];
      for (auto kv = by_logn.rbegin(); kv != by_logn.rend(); ++kv) {
        const unsigned logn = kv->first;
        std::vector<OpenMat> part;
        for (auto& om : kv->second)
          if (om.out_off >= lo && om.out_off < hi) part.push_back(om);
        if (part.empty()) continue;
        std::vector<OpenMat> mine;
        if (!grouped) {
          mine = part;
        } else if (comm_splits(ctx, logn + 1)) {
          size_t total = 0, b, e, at = 0;
          for (auto& om : part) total += om.w;
          comm_range(ctx, total, &b, &e);
          for (auto& om : part) {
            const size_t l2 = std::max(b, at), h2 = std::min(e, at + om.w);
            if (l2 < h2) {
              OpenMat piece = om;
              piece.p += (l2 - at) * om.stride;
              piece.w = (uint32_t)(h2 - l2);
              piece.out_off += 2 * (uint32_t)(l2 - at);
              mine.push_back(piece);
            }
            at += om.w;
          }
        } else if (ctx->comm->rank == 0) {
          mine = part;
        }
        open_batch_dev(ctx, mine, logn, wts[logn].first.p, wts[logn].second.p, opened.p);
        if (grouped) continue;   // the round leaves the device as a whole, after the all-reduce below
        // the group's values: runs of adjacent matrices ([col][point] pairs, 2 w values each)
        std::sort(part.begin(), part.end(), [](const OpenMat& a, const OpenMat& b) { return a.out_off < b.out_off; });
        for (size_t i = 0; i < part.size();) {
          size_t b0 = part[i].out_off, e0 = b0 + 2 * (size_t)part[i].w;
          for (i++; i < part.size() && part[i].out_off == e0; i++) e0 += 2 * (size_t)part[i].w;
          LSK_CUDA(cudaMemcpyAsync(h_open + b0, opened.p + b0, (e0 - b0) * sizeof(Ext), cudaMemcpyDeviceToHost, st));
        }
        open_done.by_group[{r, logn}] = new_event();
      }
      if (grouped && hi > lo) {
        comm_allreduce_u32(ctx, (uint32_t*)(opened.p + lo), 4 * (hi - lo));
        LSK_CUDA(cudaMemcpyAsync(h_open + lo, opened.p + lo, (hi - lo) * sizeof(Ext), cudaMemcpyDeviceToHost, st));
        open_done.by_round[r] = new_event();
      }
      lo = hi;
    }
    LSK_CUDA(cudaMemcpyAsync(h_cums, cums.p, nchips * sizeof(Ext), cudaMemcpyDeviceToHost, st));
    open_done.last = new_event();
  }
  m.last_d2h += (nopen + nchips) * sizeof(Ext);
  pt.mark("openings");

  // rounds in the order the PCS sees them: preprocessed, main, permutation, quotient
  struct RoundMat { const uint32_t* p; size_t stride; unsigned lh; uint32_t w, npoints; size_t open_off; };
  struct Round { const MerkleTreeDev* tree; std::vector<RoundMat> mats; int wait; };   // wait: open_done event of its values
  std::vector<Round> rounds;
  {
    Round rp{&m.prep_tree, {}, 0}, rm{&sd.main_tree, {}, 0}, rperm{&perm_tree, {}, 1}, rq{&quot_tree, {}, 2};
    for (size_t i = 0; i < nchips; i++) {
      const ChipH& c = m.chips[sd.chip_ids[i]];
      size_t n2 = (size_t)2 << sd.logn[i];
      unsigned lh = sd.logn[i] + 1;
      if (c.prep_width) rp.mats.push_back({m.prep_lde.p, n2, lh, c.prep_width, 2, oo[i].prep});
      rm.mats.push_back({sd.main_lde[i].p, n2, lh, c.width, 2, oo[i].main});
      rperm.mats.push_back({perm_lde[i].p, n2, lh, 4 * (uint32_t)c.perm_width(), 2, oo[i].perm});
      for (int j = 0; j < 2; j++) rq.mats.push_back({quot_lde[2 * i + j].p, n2, lh, 4, 1, oo[i].quot[j]});
    }
    if (!rp.mats.empty()) rounds.push_back(std::move(rp));
    rounds.push_back(std::move(rm));
    rounds.push_back(std::move(rperm));
    rounds.push_back(std::move(rq));
  }
  for (auto& rd : rounds) {
    unsigned waited = ~0u;
    for (auto& mt : rd.mats) {
      if (mt.lh - 1 != waited) {   // matrices come tallest first: one wait per height
        auto ev = open_done.by_group.find({rd.wait, mt.lh - 1});
        if (ev != open_done.by_group.end()) LSK_CUDA(cudaEventSynchronize(ev->second));
        else if (open_done.by_round[rd.wait]) LSK_CUDA(cudaEventSynchronize(open_done.by_round[rd.wait]));
        waited = mt.lh - 1;
      }
      for (uint32_t p = 0; p < mt.npoints; p++)
        for (uint32_t c = 0; c < mt.w; c++) ch.observe(h_open[mt.open_off + 2 * c + p]);
    }
  }
  if (open_done.last) LSK_CUDA(cudaEventSynchronize(open_done.last));   // (the cumulative sums ride behind the last group)
  Ext fri_alpha = ch.sample_ext();

  // ---- reduced openings per LDE height
  unsigned log_max = 0;
  for (auto& rd : rounds)
    for (auto& mt : rd.mats) log_max = std::max(log_max, mt.lh);
  if (log_max <= LOG_BLOWUP) fail(LSK_ERR_ARG, "degenerate shard: every chip has a single row");
  upload_pows(fri_alpha, max_w + 1, true, h_fri, d_fri);
  std::vector<std::vector<ReduceMat>> by_height(log_max + 1);
  {
    std::vector<Ext> alpha_off(log_max + 1, ext_one());
    for (auto& rd : rounds)
      for (auto& mt : rd.mats) {
        ReduceMat rmat{};
        rmat.p = mt.p; rmat.stride = mt.stride; rmat.w = mt.w; rmat.npoints = mt.npoints;
        rmat.open_off = (uint32_t)mt.open_off;   // y[p] = sum_c alpha^c opened[open_off + 2 c + p]: reduce_y_dev
        for (uint32_t p = 0; p < mt.npoints; p++) {
          rmat.off[p] = alpha_off[mt.lh];
          alpha_off[mt.lh] = emul(alpha_off[mt.lh], h_fri[mt.w]);
        }
        by_height[mt.lh].push_back(rmat);
      }
  }
  std::vector<DevBuf<Ext>> ro(log_max + 1);
  std::vector<DevBuf<ReduceMat>> ro_desc(log_max + 1);
  for (unsigned lh = 1; lh <= log_max; lh++) {
    if (by_height[lh].empty()) continue;
    ro[lh].alloc(ctx, (size_t)1 << lh);
    ro_desc[lh].alloc(ctx, by_height[lh].size());
    ro_desc[lh].upload(by_height[lh].data(), by_height[lh].size());
    reduce_y_dev(ctx, ro_desc[lh].p, (uint32_t)by_height[lh].size(), opened.p, d_fri.p);
    double bytes = 16.0 * (double)((size_t)1 << lh);
    for (auto& rmat : by_height[lh]) bytes += 4.0 * rmat.w * (double)((size_t)1 << lh);
    reduce_openings_dev(ctx, ro_desc[lh].p, (uint32_t)by_height[lh].size(), lh, zeta, d_fri.p, ro[lh].p, bytes);
  }

  pt.mark("reduce openings");
  // ---- FRI commit phase: the transcript step (observe root, sample beta) runs on the device, so the
  // layers are queued back to back; the host reads all roots once at the end and replays them
  const unsigned nlayers = log_max - LOG_BLOWUP;
  std::vector<DevBuf<Ext>> fri_vals(nlayers + 1);
  std::vector<MerkleTreeDev> fri_trees(nlayers);
  std::vector<std::array<uint32_t, 8>> fri_roots(nlayers);
  fri_vals[0] = std::move(ro[log_max]);
  if (!ch.in.empty()) fail(LSK_ERR_EXEC, "transcript input buffer not empty before the FRI commit phase");
  DevBuf<uint32_t> d_chal(ctx, 16), d_roots(ctx, 8 * (size_t)nlayers);
  DevBuf<Ext> d_betas(ctx, nlayers);
  d_chal.upload(ch.state, 16);
  for (unsigned li = 0; li < nlayers; li++) {
    unsigned lfh = log_max - 1 - li;  // log height of the folded vector = rows of this layer's matrix
    FriChallenge fc{d_chal.p, d_betas.p + li, d_roots.p + 8 * (size_t)li};
    if (lfh <= FRI_TAIL_LOG) {  // the remaining layers in one launch
      std::vector<Ext*> vals;
      std::vector<const Ext*> add;
      for (unsigned j = 0; j <= lfh; j++) {
        if (j) fri_vals[li + j].alloc(ctx, (size_t)2 << (lfh - j));
        vals.push_back(fri_vals[li + j].p);
        if (j < lfh) add.push_back(ro[lfh - j].p);
      }
      fri_commit_tail_dev(ctx, lfh, vals.data(), &fri_trees[li], add.data(), fc);
      break;
    }
    merkle_commit_fri_layer_dev(ctx, (const uint32_t*)fri_vals[li].p, lfh, fri_trees[li], fc);
    fri_vals[li + 1].alloc(ctx, (size_t)1 << lfh);
    fri_fold_dev(ctx, fri_vals[li].p, lfh, d_betas.p + li, ro[lfh].p, fri_vals[li + 1].p);
  }
  Ext fin[2];
  std::vector<uint32_t> h_roots(8 * (size_t)nlayers);
  fri_vals[nlayers].download(fin, 2);
  d_roots.download(h_roots.data(), h_roots.size());
  LSK_CUDA(cudaStreamSynchronize(st));
  for (unsigned li = 0; li < nlayers; li++) {
    for (int k = 0; k < 8; k++) {
      ch.observe(h_roots[8 * li + k]);
      fri_roots[li][k] = from_mont(h_roots[8 * li + k]);
    }
    ch.sample_ext();  // the beta the device already used
  }
  m.last_d2h += h_roots.size() * 4 + sizeof fin;
  if (!ext_eq(fin[0], fin[1])) fail(LSK_ERR_EXEC, "FRI final polynomial is not constant (prover bug)");
  Ext final_poly = fin[0];
  ch.observe(final_poly);
  pt.mark("FRI commit phase");
  // proof of work: smallest witness, found on the device
  uint32_t pow_witness;
  {
    uint32_t state_c[16], pend_c[8];
    for (int k = 0; k < 16; k++) state_c[k] = from_mont(ch.state[k]);
    for (size_t k = 0; k < ch.in.size(); k++) pend_c[k] = from_mont(ch.in[k]);
    pow_witness = grind_dev(ctx, state_c, pend_c, (unsigned)ch.in.size(), m.pow_bits);
    if (!ch.check_witness(m.pow_bits, to_mont(pow_witness))) fail(LSK_ERR_EXEC, "device grinding returned a bad witness");
  }

  pt.mark("grind");
  // ---- query phase: one query template, instantiated on the device for every sampled index
  std::vector<QuerySeg> segs;
  uint32_t wpq = 0;  // words per query
  // in a device group rank 0 supplies the literals and everything read from replicated buffers; the
  // distributed bottom layers of a tree are supplied by the rank that owns the node (comm.cuh)
  const int32_t common = !grouped ? QSEG_ALL : (ctx->comm->rank == 0 ? QSEG_ALL : QSEG_NONE);
  auto lit = [&](uint32_t v) {
    segs.push_back(QuerySeg{nullptr, 0, 1, wpq, 0, 0, 0, v, 0, common});
    wpq += 1;
  };
  auto data = [&](const uint32_t* p, uint64_t stride, uint32_t count, uint32_t shift, uint32_t xor1, uint32_t mul) {
    if (!count) return;
    segs.push_back(QuerySeg{p, stride, count, wpq, shift, xor1, mul, 0, 0, common});
    wpq += count;
  };
  auto digest = [&](const MerkleTreeDev& t, unsigned l, uint32_t shift) {   // sibling digest of layer l
    data(t.layer(l), 1, 8, shift, 1, 8);
    if (l < t.dist_layers) {
      segs.back().own_shift = t.log_max - l - ctx->comm->log_size;
      segs.back().own_rank = ctx->comm->rank;
    }
  };
  lit((uint32_t)rounds.size());
  for (auto& rd : rounds) {
    const MerkleTreeDev& t = *rd.tree;
    const uint32_t tshift = log_max - t.log_max;  // idx = index >> tshift
    lit((uint32_t)rd.mats.size());
    for (auto& mt : rd.mats) {
      lit(mt.w);
      data(mt.p, mt.stride, mt.w, log_max - mt.lh, 0, 1);  // row r = index >> (log_max - lh), one word per column
    }
    lit(t.log_max);
    for (unsigned l = 0; l < t.log_max; l++) digest(t, l, tshift + l);
  }
  lit(nlayers);
  for (unsigned li = 0; li < nlayers; li++) {
    const MerkleTreeDev& t = fri_trees[li];
    data((const uint32_t*)fri_vals[li].p, 1, 4, li, 1, 4);  // the other element of the pair
    lit(t.log_max);
    for (unsigned l = 0; l < t.log_max; l++) digest(t, l, li + 1 + l);
  }
  std::vector<uint32_t> indices(m.num_queries);
  for (auto& ix : indices) ix = ch.sample_bits(log_max);
  // the query words (the bulk of the proof: 100 queries x ~6.5 k words for a Lurk shard) are assembled on the device and
  // land in page-locked memory while the host writes the head of the shard proof
  const size_t ngathered = (size_t)m.num_queries * wpq;
  uint32_t* gathered = (uint32_t*)pinned_scratch(ctx, 1, ngathered * 4);
  DevBuf<QuerySeg> d_segs(ctx, segs.size());
  DevBuf<uint32_t> d_idx(ctx, indices.size()), d_out(ctx, ngathered);
  d_segs.upload(segs.data(), segs.size());
  d_idx.upload(indices.data(), indices.size());
  query_assemble_dev(ctx, d_segs.p, (uint32_t)segs.size(), d_idx.p, m.num_queries, wpq, d_out.p);
  if (grouped) comm_allreduce_u32(ctx, d_out.p, ngathered);
  d_out.download(gathered, ngathered);
  m.last_d2h += ngathered * 4;

  pt.mark("queries");
  // ---- serialise (word-for-word the oracle's / verifier's layout, canonical values)
  W.reserve(W.size() + 64 + 4 * nopen + 16 * nchips + m.num_pv + 8 * (size_t)nlayers + ngathered);
  W.push_back(sd.shard);
  for (int k = 0; k < 8; k++) W.push_back(sd.main_root[k]);
  for (int k = 0; k < 8; k++) W.push_back(perm_root[k]);
  for (int k = 0; k < 8; k++) W.push_back(quot_root[k]);
  W.push_back((uint32_t)nchips);
  auto push_opened = [&](size_t off, uint32_t w, uint32_t point) {
    W.push_back(w);
    for (uint32_t c = 0; c < w; c++) push_ext(W, h_open[off + 2 * c + point]);
  };
  for (size_t i = 0; i < nchips; i++) {
    const ChipH& c = m.chips[sd.chip_ids[i]];
    W.push_back(sd.chip_ids[i]);
    W.push_back(sd.logn[i]);
    push_ext(W, h_cums[i]);
    push_opened(oo[i].prep, c.prep_width, 0);
    push_opened(oo[i].prep, c.prep_width, 1);
    push_opened(oo[i].main, c.width, 0);
    push_opened(oo[i].main, c.width, 1);
    push_opened(oo[i].perm, 4 * (uint32_t)c.perm_width(), 0);
    push_opened(oo[i].perm, 4 * (uint32_t)c.perm_width(), 1);
    push_opened(oo[i].quot[0], 4, 0);
    push_opened(oo[i].quot[1], 4, 0);
  }
  W.push_back(m.num_pv);
  for (uint32_t k = 0; k < m.num_pv; k++) W.push_back(rec.public_values[k]);
  W.push_back(nlayers);
  for (auto& r : fri_roots)
    for (int k = 0; k < 8; k++) W.push_back(r[k]);
  push_ext(W, final_poly);
  W.push_back(pow_witness);
  W.push_back(m.num_queries);
  LSK_CUDA(cudaStreamSynchronize(st));   // the query words have arrived (segs / indices stay alive until here)
  W.append(gathered, ngathered);
  pt.mark("serialise");
}

static Challenger initial_challenger(const lsk_machine& m) {
  Challenger ch(&m.perm);
  for (int k = 0; k < 8; k++) ch.observe(to_mont(m.prep_commit[k]));
  ch.observe(0);  // pc_start of LairMachineProgram, src/lair/lair_chip.rs:71-77
  return ch;
}
static void observe_shard(const lsk_machine& m, Challenger& ch, const uint32_t* root, const uint32_t* pv) {
  for (int k = 0; k < 8; k++) ch.observe(to_mont(root[k]));
  for (uint32_t k = 0; k < m.num_pv; k++) ch.observe(to_mont(pv[k]));
}

// ---- layout conversion for the stage-level entry points ---------------------------------------------------
__global__ void import_rows_kernel(const uint32_t* __restrict__ rows, size_t h, uint32_t w, unsigned logh, bool bitrev,
                                   uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= h * w) return;
  size_t r = i / w, c = i % w;
  size_t dr = bitrev ? bitrev32((uint32_t)r, logh) : r;
  out[c * h + dr] = to_mont(rows[i]);
}
__global__ void export_rows_kernel(const uint32_t* __restrict__ cols, size_t h, uint32_t w, uint32_t* __restrict__ rows,
                                   unsigned logh = 0, bool bitrev = false) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= h * w) return;
  size_t r = i / w, c = i % w;
  rows[i] = from_mont(cols[c * h + (bitrev ? bitrev32((uint32_t)r, logh) : r)]);
}
static Ext ext_from_canonical(const uint32_t* w) {
  Ext e;
  for (int k = 0; k < 4; k++) {
    if (w[k] >= P) fail(LSK_ERR_ARG, "non-canonical extension element");
    e.c[k] = to_mont(w[k]);
  }
  return e;
}
static void upload_ext_pows(lsk_ctx* ctx, const Ext& base, size_t n, bool from_one, DevBuf<Ext>& dev) {
  std::vector<Ext> host(std::max<size_t>(n, 1));
  Ext p = from_one ? ext_one() : base;
  for (auto& x : host) x = p, p = emul(p, base);
  dev.alloc(ctx, host.size());
  dev.upload(host.data(), host.size());
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));   // `host` dies with this call
}

void record_unpin(lsk_record* r) {
  for (void* p : r->pinned) cudaHostUnregister(p);
  r->pinned.clear();
}

void init_prover_kernels() {
  init_ntt_kernels();
  init_merkle_kernels();
  init_stark_kernels();
}

}  // namespace lsk

extern "C" {

int lsk_machine_new(lsk_ctx* ctx, const lsk_program* prog, uint32_t entry_func, const uint32_t* blob, size_t n,
                    uint32_t num_queries, uint32_t pow_bits, lsk_machine** out) {
  LSK_API_BEGIN
  if (!ctx || !prog || !blob || !out) fail(LSK_ERR_ARG, "null argument");
  if (num_queries == 0 || num_queries > 1024 || pow_bits > 30) fail(LSK_ERR_ARG, "bad FRI parameters");
  LSK_CUDA(cudaSetDevice(ctx->device));
  auto m = std::make_unique<lsk_machine>();
  m->ctx = ctx;
  m->prog = prog;
  m->entry = entry_func;
  m->num_queries = num_queries;
  m->pow_bits = pow_bits;
  m->chips = parse_machine_blob(blob, n, &m->num_pv);
  m->dchips.resize(m->chips.size());
  for (size_t i = 0; i < m->chips.size(); i++) compile_chip(ctx, m->chips[i], m->num_pv, m->dchips[i]);
  machine_setup(*m);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = m.release();
  return LSK_OK;
  LSK_API_END
}
void lsk_machine_free(lsk_machine* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  delete m;
}
int lsk_machine_prep_commit(const lsk_machine* m, uint32_t out8[8]) {
  LSK_API_BEGIN
  if (!m || !out8) fail(LSK_ERR_ARG, "null argument");
  memcpy(out8, m->prep_commit, 32);
  return LSK_OK;
  LSK_API_END
}

int lsk_machine_jit_chips(const lsk_machine* m) {
  if (!m) return 0;
  int n = 0;
  for (auto& d : m->dchips) n += d.jit != nullptr;
  return n;
}

int lsk_jit_precompile(const uint32_t* blob, size_t n, uint32_t* compiled, uint32_t* cached) {
  LSK_API_BEGIN
  if (!blob) fail(LSK_ERR_ARG, "null blob");
  uint32_t num_pv = 0;
  const std::vector<ChipH> chips = parse_machine_blob(blob, n, &num_pv);
  std::vector<const ChipH*> todo;
  for (const ChipH& c : chips)
    if (jit_eligible(c)) todo.push_back(&c);
  // one NVRTC compilation per chip, side by side (each takes tens of seconds)
  std::vector<std::string> how(todo.size()), err(todo.size());
  std::vector<std::thread> workers;
  for (size_t i = 0; i < todo.size(); i++)
    workers.emplace_back([&, i] {
      try {
        std::vector<char> cubin;
        jit_cubin(*todo[i], true, cubin, &how[i]);
      } catch (const std::exception& e) {
        err[i] = e.what();
      }
    });
  for (auto& w : workers) w.join();
  uint32_t nc = 0, nh = 0;
  for (size_t i = 0; i < todo.size(); i++) {
    if (!err[i].empty()) fail(LSK_ERR_CUDA, "%s", err[i].c_str());
    (how[i] == "cache" ? nh : nc)++;
  }
  if (compiled) *compiled = nc;
  if (cached) *cached = nh;
  return LSK_OK;
  LSK_API_END
}

int lsk_verify(const lsk_machine* m, const uint32_t* words, size_t len) {
  LSK_API_BEGIN
  if (!m || !words) fail(LSK_ERR_ARG, "null argument");
  VerifyingKey vk;
  vk.chips = m->chips;
  vk.num_pv = m->num_pv;
  for (int k = 0; k < 8; k++) vk.prep_commit[k] = to_mont(m->prep_commit[k]);
  vk.has_prep = m->prep_chip >= 0;
  vk.num_queries = m->num_queries;
  vk.pow_bits = m->pow_bits;
  verify_proof(vk, words, len);
  return LSK_OK;
  LSK_API_END
}

int lsk_shard_commit(lsk_machine* m, const lsk_record* rec, uint32_t shard, uint32_t shard_size, uint32_t root_out[8]) {
  LSK_API_BEGIN
  if (!m || !rec || !root_out || !shard_size) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(m->ctx->device));
  if (shard >= lsk_record_num_shards(rec, shard_size)) fail(LSK_ERR_ARG, "shard %u out of range", shard);
  auto sd = commit_shard(*m, *rec, shard, shard_size, false);
  memcpy(root_out, sd->main_root, 32);
  m->pending[shard] = std::move(sd);
  return LSK_OK;
  LSK_API_END
}
int lsk_shard_prove(lsk_machine* m, const lsk_record* rec, uint32_t shard, const uint32_t* all_roots, uint32_t nshards,
                    uint32_t** words, size_t* len) {
  LSK_API_BEGIN
  if (!m || !rec || !all_roots || !words || !len) fail(LSK_ERR_ARG, "null argument");
  LSK_CUDA(cudaSetDevice(m->ctx->device));
  auto it = m->pending.find(shard);
  if (it == m->pending.end()) fail(LSK_ERR_ARG, "shard %u was not committed", shard);
  std::unique_ptr<ShardState> sd = std::move(it->second);
  m->pending.erase(it);
  Challenger ch = initial_challenger(*m);
  for (uint32_t s = 0; s < nshards; s++) observe_shard(*m, ch, all_roots + 8 * s, rec->public_values.data());
  ProofWords W;
  prove_shard(*m, *rec, *sd, ch, W);
  *len = W.size();
  *words = W.release();
  return LSK_OK;
  LSK_API_END
}
void lsk_release_record(lsk_machine* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  m->pending.clear();
  m->drecs.clear();
  m->drec_for = nullptr;
  m->last_h2d = m->last_d2h = 0;
}

int lsk_record_pin(lsk_record* r) {
  LSK_API_BEGIN
  if (!r) fail(LSK_ERR_ARG, "null record");
  if (!r->pinned.empty()) return LSK_OK;
  auto pin = [&](std::vector<uint32_t>& v) {
    if (v.size() < 16384) return;  // small tables are not worth a registration
    LSK_CUDA(cudaHostRegister(v.data(), v.size() * 4, cudaHostRegisterDefault));
    r->pinned.push_back(v.data());
  };
  try {
    for (auto& t : r->funcs) {
      pin(t.args); pin(t.outs); pin(t.provide); pin(t.req_off); pin(t.reqs); pin(t.hint_off); pin(t.hints); pin(t.depth); pin(t.dreq_off); pin(t.dreqs);
    }
    for (auto& mt : r->mems) { pin(mt.vals); pin(mt.provide); }
    pin(r->bytes);
  } catch (...) {
    record_unpin(r);
    throw;
  }
  return LSK_OK;
  LSK_API_END
}
void lsk_record_unpin(lsk_record* r) {
  if (r) record_unpin(r);
}

int lsk_prove(lsk_machine* m, const lsk_record* rec, uint32_t shard_size, uint32_t** words, size_t* len) {
  LSK_API_BEGIN
  if (!m || !rec || !words || !len || !shard_size) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(m->ctx->device));
  uint32_t ns = lsk_record_num_shards(rec, shard_size);
  if (!ns) fail(LSK_ERR_ARG, "empty record");
  m->last_h2d = m->last_d2h = 0;
  m->drecs.clear();  // always re-upload: the record is a host buffer owned by the caller
  m->drec_for = nullptr;
  struct Cleanup {
    lsk_machine* m;
    ~Cleanup() { m->pending.clear(); m->drecs.clear(); m->drec_for = nullptr; }
  } cleanup{m};
  std::vector<std::unique_ptr<ShardState>> shards;
  Challenger ch = initial_challenger(*m);
  for (uint32_t s = 0; s < ns; s++) {
    shards.push_back(commit_shard(*m, *rec, s, shard_size, true));
    observe_shard(*m, ch, shards.back()->main_root, rec->public_values.data());
  }
  ProofWords W;
  W.push_back(PROOF_MAGIC);
  W.push_back(ns);
  for (uint32_t s = 0; s < ns; s++) {
    prove_shard(*m, *rec, *shards[s], ch, W);
    shards[s].reset();
  }
  *len = W.size();
  *words = W.release();
  return LSK_OK;
  LSK_API_END
}
void lsk_proof_free(uint32_t* words) { free(words); }
size_t lsk_machine_last_h2d_bytes(const lsk_machine* m) { return m ? m->last_h2d : 0; }
size_t lsk_machine_last_d2h_bytes(const lsk_machine* m) { return m ? m->last_d2h : 0; }

// ---- stage-level entry points (parity tests, roofline measurement) ------------------------------------------
int lsk_coset_lde(lsk_ctx* ctx, const uint32_t* rows, size_t h, size_t w, int shift_kind, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || !rows || !out || !h || !w || (h & (h - 1)) || shift_kind < 0 || shift_kind > 2) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(ctx->device));
  unsigned logh = log2_exact(h);
  DevBuf<uint32_t> d_rows(ctx, h * w), d_cols(ctx, h * w), d_lde(ctx, 2 * h * w), d_out(ctx, 2 * h * w);
  d_rows.upload(rows, h * w);
  import_rows_kernel<<<div_up(h * w, 256), 256, 0, ctx->stream>>>(d_rows.p, h, (uint32_t)w, logh, true, d_cols.p);
  lde_dev(ctx, d_cols.p, h, d_lde.p, 2 * h, logh, (uint32_t)w, (LdeShift)shift_kind);
  export_rows_kernel<<<div_up(2 * h * w, 256), 256, 0, ctx->stream>>>(d_lde.p, 2 * h, (uint32_t)w, d_out.p);
  LSK_CUDA(cudaGetLastError());
  d_out.download(out, 2 * h * w);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_merkle_root(lsk_ctx* ctx, const uint32_t* const* mats, const size_t* hs, const size_t* ws, size_t nmats,
                    uint32_t out8[8]) {
  LSK_API_BEGIN
  if (!ctx || !mats || !hs || !ws || !nmats || !out8) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(ctx->device));
  std::vector<DevBuf<uint32_t>> bufs;
  std::vector<MatRef> refs;
  for (size_t i = 0; i < nmats; i++) {
    size_t h = hs[i], w = ws[i];
    if (!h || !w || (h & (h - 1))) fail(LSK_ERR_ARG, "matrix heights must be powers of two");
    DevBuf<uint32_t> d_rows(ctx, h * w);
    bufs.emplace_back(ctx, h * w);
    d_rows.upload(mats[i], h * w);
    import_rows_kernel<<<div_up(h * w, 256), 256, 0, ctx->stream>>>(d_rows.p, h, (uint32_t)w, 0, false, bufs.back().p);
    refs.push_back(MatRef{bufs.back().p, h, log2_exact(h), (uint32_t)w});
  }
  LSK_CUDA(cudaGetLastError());
  MerkleTreeDev tree;
  merkle_commit_dev(ctx, refs, tree);
  merkle_root_host(ctx, tree, out8);
  return LSK_OK;
  LSK_API_END
}


// ---- A8 / A10 / A12 as stages: the third-party pieces behind machine.prove, one at a time, so that a
// mismatch with the oracle can be localised without bisecting a whole proof ---------------------------------------
int lsk_machine_chip_perm_width(const lsk_machine* m, uint32_t chip) {
  return (m && chip < m->chips.size()) ? (int)m->chips[chip].perm_width() : -1;
}
static const ChipH& stage_chip(const lsk_machine* m, uint32_t chip, size_t h, const uint32_t* main_rows) {
  if (!m || !main_rows || chip >= m->chips.size() || !h || (h & (h - 1)) || h > ((size_t)1 << 24)) fail(LSK_ERR_ARG, "bad argument");
  const ChipH& c = m->chips[chip];
  if (c.prep_width && h != 65536) fail(LSK_ERR_ARG, "the byte chip's traces have 65536 rows");
  return c;
}
int lsk_perm_trace(lsk_machine* m, uint32_t chip, const uint32_t* main_rows, size_t h, const uint32_t chal_alpha[4],
                   const uint32_t chal_beta[4], uint32_t* perm_rows_out, uint32_t cum_sum_out[4]) {
  LSK_API_BEGIN
  const ChipH& c = stage_chip(m, chip, h, main_rows);
  if (!chal_alpha || !chal_beta || !perm_rows_out || !cum_sum_out) fail(LSK_ERR_ARG, "null argument");
  lsk_ctx* ctx = m->ctx;
  LSK_CUDA(cudaSetDevice(ctx->device));
  const unsigned logn = log2_exact(h);
  const uint32_t pcols = 4 * (uint32_t)c.perm_width();
  DevBuf<uint32_t> d_rows(ctx, h * c.width), tr(ctx, h * c.width), perm(ctx, h * pcols), d_out(ctx, h * pcols);
  d_rows.upload(main_rows, h * c.width);
  import_rows_kernel<<<div_up(h * c.width, 256), 256, 0, ctx->stream>>>(d_rows.p, h, c.width, logn, true, tr.p);
  DevBuf<Ext> d_beta, scratch(ctx, h), cum(ctx, 1);
  upload_ext_pows(ctx, ext_from_canonical(chal_beta), c.max_values(), false, d_beta);
  perm_trace_dev(ctx, m->dchips[chip], tr.p, h, c.prep_width ? m->prep_trace.p : nullptr, h, logn, ext_from_canonical(chal_alpha),
                 d_beta.p, perm.p, h, scratch.p, cum.p);
  export_rows_kernel<<<div_up(h * pcols, 256), 256, 0, ctx->stream>>>(perm.p, h, pcols, d_out.p, logn, true);
  LSK_CUDA(cudaGetLastError());
  d_out.download(perm_rows_out, h * pcols);
  Ext hc;
  cum.download(&hc, 1);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 4; k++) cum_sum_out[k] = from_mont(hc.c[k]);
  return LSK_OK;
  LSK_API_END
}
int lsk_quotient(lsk_machine* m, uint32_t chip, const uint32_t* main_rows, const uint32_t* perm_rows, size_t h,
                 const uint32_t* public_values, const uint32_t cum_sum[4], const uint32_t chal_alpha[4],
                 const uint32_t chal_beta[4], const uint32_t alpha[4], uint32_t* quot_out) {
  LSK_API_BEGIN
  const ChipH& c = stage_chip(m, chip, h, main_rows);
  if (!perm_rows || !cum_sum || !chal_alpha || !chal_beta || !alpha || !quot_out || (m->num_pv && !public_values))
    fail(LSK_ERR_ARG, "null argument");
  lsk_ctx* ctx = m->ctx;
  LSK_CUDA(cudaSetDevice(ctx->device));
  const unsigned logn = log2_exact(h);
  const uint32_t pcols = 4 * (uint32_t)c.perm_width();
  const size_t n2 = 2 * h;
  DevBuf<uint32_t> d_rows(ctx, h * std::max<size_t>(c.width, pcols)), tr(ctx, h * c.width), pm(ctx, h * pcols);
  DevBuf<uint32_t> main_lde(ctx, n2 * c.width), perm_lde(ctx, n2 * pcols), qv(ctx, n2 * 4), d_out(ctx, n2 * 4);
  d_rows.upload(main_rows, h * c.width);
  import_rows_kernel<<<div_up(h * c.width, 256), 256, 0, ctx->stream>>>(d_rows.p, h, c.width, logn, true, tr.p);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  d_rows.upload(perm_rows, h * pcols);
  import_rows_kernel<<<div_up(h * pcols, 256), 256, 0, ctx->stream>>>(d_rows.p, h, pcols, logn, true, pm.p);
  lde_dev(ctx, tr.p, h, main_lde.p, n2, logn, c.width, LdeShift::Trace);
  lde_dev(ctx, pm.p, h, perm_lde.p, n2, logn, pcols, LdeShift::Trace);
  DevBuf<Ext> d_beta, d_alpha, d_cum(ctx, 1);
  upload_ext_pows(ctx, ext_from_canonical(chal_beta), c.max_values(), false, d_beta);
  const uint32_t n_total = (uint32_t)(c.n_constraints + c.n_perm_constraints());
  upload_ext_pows(ctx, ext_from_canonical(alpha), n_total, true, d_alpha);
  Ext hc = ext_from_canonical(cum_sum);
  d_cum.upload(&hc, 1);
  std::vector<uint32_t> pv(std::max<size_t>(1, m->num_pv), 0);
  for (uint32_t k = 0; k < m->num_pv; k++) {
    if (public_values[k] >= P) fail(LSK_ERR_ARG, "non-canonical public value");
    pv[k] = to_mont(public_values[k]);
  }
  DevBuf<uint32_t> d_pv(ctx, pv.size());
  d_pv.upload(pv.data(), pv.size());
  QuotientArgs a{};
  a.main = main_lde.p; a.mstride = n2;
  a.prep = c.prep_width ? m->prep_lde.p : nullptr; a.pstride = n2;
  a.perm = perm_lde.p; a.permstride = n2;
  a.pv = d_pv.p;
  a.apow = d_alpha.p;
  a.n_total = n_total;
  a.beta_pows = d_beta.p;
  a.cum_sum = d_cum.p;
  a.chal_alpha = ext_from_canonical(chal_alpha);
  a.logn = logn;
  a.out = qv.p; a.ostride = n2;
  quotient_dev(ctx, m->dchips[chip], a);
  export_rows_kernel<<<div_up(n2 * 4, 256), 256, 0, ctx->stream>>>(qv.p, n2, 4, d_out.p);
  LSK_CUDA(cudaGetLastError());
  d_out.download(quot_out, n2 * 4);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_fri_prove(lsk_ctx* ctx, const uint32_t* vals, uint32_t log_n, const uint32_t challenger_state[16], uint32_t num_queries,
                  uint32_t pow_bits, uint32_t** words, size_t* len) {
  LSK_API_BEGIN
  if (!ctx || !vals || !challenger_state || !words || !len || log_n < 2 || log_n > 25 || !num_queries || num_queries > 1024 ||
      pow_bits > 30)
    fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  static const HostPerm16 perm;
  Challenger ch(&perm);
  for (int k = 0; k < 16; k++) {
    if (challenger_state[k] >= P) fail(LSK_ERR_ARG, "non-canonical challenger state");
    ch.state[k] = to_mont(challenger_state[k]);
  }
  const size_t n = (size_t)1 << log_n;
  const unsigned log_max = log_n, nlayers = log_max - LOG_BLOWUP;
  std::vector<DevBuf<Ext>> fri_vals(nlayers + 1);
  std::vector<MerkleTreeDev> fri_trees(nlayers);
  {
    std::vector<Ext> h(n);
    for (size_t i = 0; i < n; i++) h[i] = ext_from_canonical(vals + 4 * i);
    fri_vals[0].alloc(ctx, n);
    fri_vals[0].upload(h.data(), n);
    LSK_CUDA(cudaStreamSynchronize(st));
  }
  DevBuf<uint32_t> d_chal(ctx, 16), d_roots(ctx, 8 * (size_t)nlayers);
  DevBuf<Ext> d_betas(ctx, nlayers);
  d_chal.upload(ch.state, 16);
  for (unsigned li = 0; li < nlayers; li++) {
    unsigned lfh = log_max - 1 - li;
    FriChallenge fc{d_chal.p, d_betas.p + li, d_roots.p + 8 * (size_t)li};
    if (lfh <= FRI_TAIL_LOG) {
      std::vector<Ext*> v;
      std::vector<const Ext*> add;
      for (unsigned j = 0; j <= lfh; j++) {
        if (j) fri_vals[li + j].alloc(ctx, (size_t)2 << (lfh - j));
        v.push_back(fri_vals[li + j].p);
        if (j < lfh) add.push_back(nullptr);
      }
      fri_commit_tail_dev(ctx, lfh, v.data(), &fri_trees[li], add.data(), fc);
      break;
    }
    merkle_commit_fri_layer_dev(ctx, (const uint32_t*)fri_vals[li].p, lfh, fri_trees[li], fc);
    fri_vals[li + 1].alloc(ctx, (size_t)1 << lfh);
    fri_fold_dev(ctx, fri_vals[li].p, lfh, d_betas.p + li, nullptr, fri_vals[li + 1].p);
  }
  Ext fin[2];
  std::vector<uint32_t> h_roots(8 * (size_t)nlayers);
  fri_vals[nlayers].download(fin, 2);
  d_roots.download(h_roots.data(), h_roots.size());
  LSK_CUDA(cudaStreamSynchronize(st));
  std::vector<uint32_t> W;
  W.push_back(nlayers);
  for (unsigned li = 0; li < nlayers; li++) {
    for (int k = 0; k < 8; k++) {
      ch.observe(h_roots[8 * li + k]);
      W.push_back(from_mont(h_roots[8 * li + k]));
    }
    ch.sample_ext();
  }
  if (!ext_eq(fin[0], fin[1])) fail(LSK_ERR_EXEC, "the folded vector is not constant: the input is not a low-degree codeword");
  for (int k = 0; k < 4; k++) W.push_back(from_mont(fin[0].c[k]));
  ch.observe(fin[0]);
  uint32_t state_c[16], pend_c[8];
  for (int k = 0; k < 16; k++) state_c[k] = from_mont(ch.state[k]);
  for (size_t k = 0; k < ch.in.size(); k++) pend_c[k] = from_mont(ch.in[k]);
  uint32_t pow_witness = grind_dev(ctx, state_c, pend_c, (unsigned)ch.in.size(), pow_bits);
  if (!ch.check_witness(pow_bits, to_mont(pow_witness))) fail(LSK_ERR_EXEC, "device grinding returned a bad witness");
  W.push_back(pow_witness);
  W.push_back(num_queries);
  // one query: index, then per layer the sibling value and the Merkle path of the pair's row
  std::vector<QuerySeg> segs;
  uint32_t wpq = 0;
  auto data = [&](const uint32_t* p, uint64_t stride, uint32_t count, uint32_t shift, uint32_t xor1, uint32_t mul) {
    segs.push_back(QuerySeg{p, stride, count, wpq, shift, xor1, mul, 0, 0, QSEG_ALL});
    wpq += count;
  };
  auto lit = [&](uint32_t v) {
    segs.push_back(QuerySeg{nullptr, 0, 1, wpq, 0, 0, 0, v, 0, QSEG_ALL});
    wpq += 1;
  };
  for (unsigned li = 0; li < nlayers; li++) {
    const MerkleTreeDev& t = fri_trees[li];
    data((const uint32_t*)fri_vals[li].p, 1, 4, li, 1, 4);
    lit(t.log_max);
    for (unsigned l = 0; l < t.log_max; l++) data(t.layer(l), 1, 8, li + 1 + l, 1, 8);
  }
  std::vector<uint32_t> indices(num_queries);
  for (auto& ix : indices) ix = ch.sample_bits(log_max);
  std::vector<uint32_t> gathered((size_t)num_queries * wpq);
  {
    DevBuf<QuerySeg> d_segs(ctx, segs.size());
    DevBuf<uint32_t> d_idx(ctx, indices.size()), d_out(ctx, gathered.size());
    d_segs.upload(segs.data(), segs.size());
    d_idx.upload(indices.data(), indices.size());
    query_assemble_dev(ctx, d_segs.p, (uint32_t)segs.size(), d_idx.p, num_queries, wpq, d_out.p);
    d_out.download(gathered.data(), gathered.size());
    LSK_CUDA(cudaStreamSynchronize(st));
  }
  for (uint32_t q = 0; q < num_queries; q++) {
    W.push_back(indices[q]);
    W.insert(W.end(), gathered.begin() + (size_t)q * wpq, gathered.begin() + (size_t)(q + 1) * wpq);
  }
  *words = (uint32_t*)malloc(W.size() * 4 + 4);
  if (!*words) fail(LSK_ERR_ARG, "out of memory");
  memcpy(*words, W.data(), W.size() * 4);
  *len = W.size();
  return LSK_OK;
  LSK_API_END
}

// ---- column-sharded commit: device-pointer pieces strung together by the host layer ----------
int lsk_import_columns_dev(lsk_ctx* ctx, const uint32_t* rows, size_t h, size_t w, uint32_t* cols_dev) {
  LSK_API_BEGIN
  if (!ctx || !rows || !cols_dev || !h || !w || (h & (h - 1))) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(ctx->device));
  DevBuf<uint32_t> d_rows(ctx, h * w);
  d_rows.upload(rows, h * w);
  import_rows_kernel<<<div_up(h * w, 256), 256, 0, ctx->stream>>>(d_rows.p, h, (uint32_t)w, log2_exact(h), true, cols_dev);
  LSK_CUDA(cudaGetLastError());
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_lde_columns_dev(lsk_ctx* ctx, const uint32_t* in_dev, size_t in_stride, uint32_t* out_dev, size_t out_stride,
                        uint32_t logn, uint32_t ncols, int shift_kind) {
  LSK_API_BEGIN
  if (!ctx || !in_dev || !out_dev || shift_kind < 0 || shift_kind > 2) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(ctx->device));
  lde_dev(ctx, in_dev, in_stride, out_dev, out_stride, logn, ncols, (LdeShift)shift_kind);
  return LSK_OK;
  LSK_API_END
}
int lsk_merkle_subtree_root_dev(lsk_ctx* ctx, const uint32_t* cols_dev, size_t stride, uint32_t log_rows, uint32_t w,
                                uint32_t root_out8[8]) {
  LSK_API_BEGIN
  if (!ctx || !cols_dev || !w || !root_out8) fail(LSK_ERR_ARG, "bad argument");
  LSK_CUDA(cudaSetDevice(ctx->device));
  MerkleTreeDev tree;
  merkle_commit_dev(ctx, {MatRef{cols_dev, stride, log_rows, w}}, tree);
  merkle_root_host(ctx, tree, root_out8);
  return LSK_OK;
  LSK_API_END
}
int lsk_merkle_join_roots(const uint32_t* roots, uint32_t count, uint32_t out8[8]) {
  LSK_API_BEGIN
  if (!roots || !out8 || !count || (count & (count - 1))) fail(LSK_ERR_ARG, "the number of subtree roots must be a power of two");
  static const HostPerm16 perm;
  std::vector<uint32_t> cur(roots, roots + 8 * (size_t)count);
  for (auto& x : cur) {
    if (x >= P) fail(LSK_ERR_ARG, "non-canonical digest");
    x = to_mont(x);
  }
  for (uint32_t n = count; n > 1; n >>= 1) {
    for (uint32_t i = 0; i < n / 2; i++) {
      uint32_t st[16];
      memcpy(st, &cur[16 * i], 64);
      perm.permute(st);
      memcpy(&cur[8 * i], st, 32);
    }
  }
  for (int k = 0; k < 8; k++) out8[k] = from_mont(cur[k]);
  return LSK_OK;
  LSK_API_END
}

}  // extern "C"
