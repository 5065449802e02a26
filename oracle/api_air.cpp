// ORACLE (test infrastructure, NOT product code): C entry points over air_check.h for ctypes.
//   orc_air_check        = debug_constraints_collecting_queries (src/air/debug.rs:161-206) on one trace
//   orc_tq_verify        = TraceQueries::verify / verify_many (src/air/debug.rs:91-116)
//   orc_air_cross_check  = the direct restatement of Air::eval against a captured 'MACH' chip vector
//                          (what sphinx's Chip::new collects, src/lair/lair_chip.rs:196-231) on random rows
#include "air_check.h"
#include "stark.h"
using namespace orc;

static thread_local std::string g_air_err;
#define AIR_TRY try {
#define AIR_CATCH } catch (std::exception& e) { g_air_err = e.what(); return -1; }

static uint32_t air_index(uint32_t kind, uint32_t index) {
  if (kind == AIR_MEM) {
    if (index >= 6) throw std::runtime_error("bad memory table");
    return (uint32_t)MEM_TABLE_SIZES[index];
  }
  return index;
}

extern "C" {
const char* orc_air_last_error() { return g_air_err.c_str(); }
void* orc_tq_new() { return new TraceQueries; }
void orc_tq_free(void* q) { delete (TraceQueries*)q; }

// kind: 0 func (index = func), 1 mem (index = table 0..5), 2 bytes, 3 entrypoint (index = entry func).
// main / prep: row-major canonical words; the trace's queries are merged into `tq`.
int orc_air_check(void* top, void* tq, uint32_t kind, uint32_t index, const uint32_t* main, size_t h, size_t w,
                  const uint32_t* prep, size_t pw, const uint32_t* pv, size_t npv) {
  AIR_TRY
  const Toplevel& tl = *(Toplevel*)top;
  std::vector<Fp> m(h * w), p, v(npv);
  for (size_t i = 0; i < h * w; i++) m[i] = Fp::from_u32(main[i]);
  if (prep) {
    p.resize(h * pw);
    for (size_t i = 0; i < h * pw; i++) p[i] = Fp::from_u32(prep[i]);
  }
  for (size_t i = 0; i < npv; i++) v[i] = Fp::from_u32(pv[i]);
  if (kind == AIR_FUNC && tl.funcs.at(index).width() != w) throw std::runtime_error("trace width != func width");
  TraceQueries local;
  air_check_trace(tl, kind, air_index(kind, index), m, h, w, prep ? &p : nullptr, pw, v, local);
  ((TraceQueries*)tq)->merge(local);
  return 0;
  AIR_CATCH
}
int orc_tq_verify(void* tq) {
  AIR_TRY
  ((TraceQueries*)tq)->verify();
  return 0;
  AIR_CATCH
}
int orc_tq_has_memoset_key(void* tq, const uint32_t* key, size_t n) {
  return ((TraceQueries*)tq)->memoset.count(std::vector<uint32_t>(key, key + n)) ? 1 : 0;
}
void orc_tq_counts(void* tq, uint64_t out[3]) {
  auto* q = (TraceQueries*)tq;
  out[0] = q->n_sends;
  out[1] = q->n_receives;
  out[2] = q->memoset.size();
}

// For chip `ci` of the captured machine: `nrows` random (local, next, preprocessed, public values, row
// selectors) tuples; the captured constraint program and interaction list must give, value for value and
// in order, what the direct restatement of Air::eval gives.  counts = {constraints, sends, receives}.
int orc_air_cross_check(void* top, const uint32_t* blob, size_t n, uint32_t ci, uint32_t nrows, uint64_t seed,
                        uint32_t counts[3]) {
  AIR_TRY
  const Toplevel& tl = *(Toplevel*)top;
  uint32_t num_pv = 0;
  std::vector<Chip> chips = parse_machine(blob, n, &num_pv);
  const Chip& c = chips.at(ci);
  uint64_t s = seed * 0x9e3779b97f4a7c15ull + 0x1234567ull + ci;
  auto rnd = [&]() {
    s += 0x9e3779b97f4a7c15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return Fp::from_u32((uint32_t)((z ^ (z >> 31)) % P));
  };
  for (uint32_t r = 0; r < nrows; r++) {
    std::vector<Fp> local(c.width), next(c.width), pl(c.prep_width + 1), pn(c.prep_width + 1), pv(num_pv + 1);
    for (auto* v : {&local, &next, &pl, &pn, &pv})
      for (auto& x : *v) x = rnd();
    DebugBuilder b;
    AirRowOut out;
    b.local = local.data(); b.next = next.data(); b.prep_local = pl.data(); b.prep_next = pn.data(); b.pv = pv.data();
    b.is_first = rnd(); b.is_last = rnd(); b.is_transition = rnd();
    b.strict = false;
    b.out = &out;
    air_eval_row(tl, c.kind, air_index(c.kind, c.index), num_pv, b);
    // the captured program on the same values
    std::vector<Fp> regs(c.n_regs), got;
    auto operand = [&](uint32_t w) -> Fp {
      uint32_t tag = w >> 28, idx = w & 0x0fffffffu;
      switch (tag) {
        case 0: return regs.at(idx);
        case 1: return c.consts.at(idx);
        case 2: return local.at(idx);
        case 3: return next.at(idx);
        case 4: return pl.at(idx);
        case 5: return pn.at(idx);
        case 6: return pv.at(idx);
        default: return idx == 0 ? b.is_first : idx == 1 ? b.is_last : b.is_transition;
      }
    };
    for (size_t i = 0; i < c.code.size(); i += 4) {
      uint32_t op = c.code[i], dst = c.code[i + 1];
      switch (op) {
        case 1: regs.at(dst) = operand(c.code[i + 2]) + operand(c.code[i + 3]); break;
        case 2: regs.at(dst) = operand(c.code[i + 2]) - operand(c.code[i + 3]); break;
        case 3: regs.at(dst) = operand(c.code[i + 2]) * operand(c.code[i + 3]); break;
        case 4: regs.at(dst) = -operand(c.code[i + 2]); break;
        case 5: got.push_back(operand(c.code[i + 2])); break;
        default: throw std::runtime_error("bad constraint opcode");
      }
    }
    auto fail = [&](const std::string& what, size_t k) {
      std::ostringstream m;
      m << "chip " << ci << " (" << c.name << "): " << what << " " << k << " differs between the captured program and Air::eval";
      throw std::runtime_error(m.str());
    };
    if (got.size() != out.constraints.size()) {
      std::ostringstream m;
      m << "chip " << ci << " (" << c.name << "): captured program has " << got.size() << " constraints, Air::eval asserts "
        << out.constraints.size();
      throw std::runtime_error(m.str());
    }
    if (got.size() != c.n_constraints) throw std::runtime_error("n_constraints of the descriptor is wrong");
    for (size_t k = 0; k < got.size(); k++)
      if (got[k] != out.constraints[k]) fail("constraint", k);
    auto gm = [&](uint32_t col) { return local.at(col); };
    auto gp = [&](uint32_t col) { return pl.at(col); };
    for (int side = 0; side < 2; side++) {
      const auto& cap = side == 0 ? c.sends : c.receives;
      const auto& dir = side == 0 ? out.sends : out.receives;
      if (cap.size() != dir.size()) {
        std::ostringstream m;
        m << "chip " << ci << " (" << c.name << "): captured " << cap.size() << (side ? " receives" : " sends") << ", Air::eval makes "
          << dir.size();
        throw std::runtime_error(m.str());
      }
      for (size_t k = 0; k < cap.size(); k++) {
        if (cap[k].kind != 1) fail("interaction kind (InteractionKind::Memory = 1, src/air/builder.rs:107-133) of message", k);
        if (cap[k].values.size() != dir[k].values.size()) fail(side ? "length of receive" : "length of send", k);
        for (size_t j = 0; j < cap[k].values.size(); j++)
          if (cap[k].values[j].eval<Fp>(gm, gp) != dir[k].values[j]) fail(side ? "receive" : "send", k);
        if (cap[k].mult.eval<Fp>(gm, gp) != dir[k].mult) fail(side ? "multiplicity of receive" : "multiplicity of send", k);
      }
    }
    counts[0] = (uint32_t)got.size();
    counts[1] = (uint32_t)c.sends.size();
    counts[2] = (uint32_t)c.receives.size();
  }
  return 0;
  AIR_CATCH
}
}  // extern "C"
