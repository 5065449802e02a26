// Lair front half on the host (see host_lair.h for the reference map).
#include "host_lair.h"

#include <algorithm>
#include <stdexcept>

namespace lurk {
namespace host {

const uint32_t MEM_TABLE_SIZES[6] = {2, 3, 4, 5, 6, 8};  // src/lair/execute.rs:243-244

[[noreturn]] static void bail(const std::string& msg) { throw std::runtime_error(msg); }

// ---- builder helpers --------------------------------------------------------------------------------------
static St mk(SK k) {
  St s;
  s.kind = k;
  return s;
}
static St bin(SK k, const std::string& tgt, const std::string& a, const std::string& b) {
  St s = mk(k);
  s.tgt = tgt; s.a = a; s.b = b;
  return s;
}
St const_(const std::string& tgt, int64_t value) {
  St s = mk(SK::Const);
  s.tgt = tgt; s.value = F(value);
  return s;
}
St array(const std::string& tgt, const std::vector<int64_t>& values) {
  St s = mk(SK::Array);
  s.tgt = tgt;
  for (int64_t v : values) s.values.push_back(F(v));
  return s;
}
St add(const std::string& t, const std::string& a, const std::string& b) { return bin(SK::Add, t, a, b); }
St sub(const std::string& t, const std::string& a, const std::string& b) { return bin(SK::Sub, t, a, b); }
St mul(const std::string& t, const std::string& a, const std::string& b) { return bin(SK::Mul, t, a, b); }
St div(const std::string& t, const std::string& a, const std::string& b) { return bin(SK::Div, t, a, b); }
St inv(const std::string& t, const std::string& a) { return bin(SK::Inv, t, a, ""); }
St not_(const std::string& t, const std::string& a) { return bin(SK::Not, t, a, ""); }
St eq(const std::string& t, const std::string& a, const std::string& b) { return bin(SK::Eq, t, a, b); }
St assert_eq(const std::string& a, const std::string& b) { return bin(SK::AssertEq, "", a, b); }
St assert_ne(const std::string& a, const std::string& b) { return bin(SK::AssertNe, "", a, b); }
St contains(const std::string& a, const std::string& b) { return bin(SK::Contains, "", a, b); }
static St call_like(SK k, const std::vector<Var>& outs, const std::string& f, const std::vector<std::string>& inps) {
  St s = mk(k);
  s.outs = outs; s.callee = f; s.names = inps;
  return s;
}
St call(const std::vector<Var>& o, const std::string& f, const std::vector<std::string>& i) { return call_like(SK::Call, o, f, i); }
St preimg(const std::vector<Var>& o, const std::string& f, const std::vector<std::string>& i) { return call_like(SK::PreImg, o, f, i); }
St extern_call(const std::vector<Var>& o, const std::string& c, const std::vector<std::string>& i) {
  return call_like(SK::ExternCall, o, c, i);
}
St store(const std::string& ptr, const std::vector<std::string>& vals) {
  St s = mk(SK::Store);
  s.tgt = ptr; s.names = vals;
  return s;
}
St load(const std::vector<Var>& vals, const std::string& ptr) {
  St s = mk(SK::Load);
  s.outs = vals; s.a = ptr;
  return s;
}
St range_u8(const std::vector<std::string>& xs) {
  St s = mk(SK::RangeU8);
  s.names = xs;
  return s;
}
St slice(const std::vector<Var>& outs, const std::vector<std::string>& inps) {
  St s = mk(SK::Slice);
  s.outs = outs;
  s.names = inps;
  return s;
}
St emit(const std::vector<std::string>& xs) {
  St s = mk(SK::Emit);
  s.names = xs;
  return s;
}
St ret(const std::vector<std::string>& vs) {
  St s = mk(SK::Return);
  s.names = vs;
  return s;
}
St if_(const std::string& x, Stmts then) {
  St s = mk(SK::If);
  s.a = x;
  s.arms.push_back(Arm{{}, std::move(then)});
  return s;
}
St if_else(const std::string& x, Stmts then, Stmts else_) {
  St s = if_(x, std::move(then));
  s.dflt = std::make_shared<Stmts>(std::move(else_));
  return s;
}
static St match_like(SK k, const std::string& x, std::vector<std::pair<std::vector<int64_t>, Stmts>> arms, bool exhaustive) {
  St s = mk(k);
  s.a = x;
  s.exhaustive = exhaustive;
  for (auto& arm : arms) {
    Arm a;
    for (int64_t key : arm.first) a.keys.push_back(F(key));
    a.body = std::move(arm.second);
    s.arms.push_back(std::move(a));
  }
  return s;
}
St match(const std::string& x, std::vector<std::pair<std::vector<int64_t>, Stmts>> arms, bool exhaustive) {
  return match_like(SK::Match, x, std::move(arms), exhaustive);
}
St match_many(const std::string& x, std::vector<std::pair<std::vector<int64_t>, Stmts>> arms, bool exhaustive) {
  return match_like(SK::MatchMany, x, std::move(arms), exhaustive);
}
FuncE func(const std::string& name, std::vector<Var> params, uint32_t output_size, Stmts body, bool invertible, bool partial) {
  FuncE f;
  f.name = name; f.input_params = std::move(params); f.output_size = output_size; f.body = std::move(body);
  f.invertible = invertible; f.partial = partial;
  return f;
}

// ---- compilation: names -> stack indices, If / Match -> Choose (src/lair/toplevel.rs) ---------------------
namespace {
struct FuncInfo { uint32_t index, in, out; bool partial; };
struct Linker {
  uint32_t var_index = 0, return_ident = 0, uniq = 0;
  std::vector<uint32_t> return_idents;
  std::map<std::string, std::vector<uint32_t>> link;
  const std::map<std::string, FuncInfo>* info = nullptr;
  const std::map<std::string, std::pair<uint32_t, ChipInfo>>* chips = nullptr;
  bool partial = false;
  std::vector<uint32_t> fresh_var(const std::string& name, uint32_t n) {
    std::vector<uint32_t> idxs(n);
    for (uint32_t i = 0; i < n; i++) idxs[i] = var_index + i;
    var_index += n;
    link[name] = idxs;
    return idxs;
  }
  const std::vector<uint32_t>& get(const std::string& name) const {
    auto it = link.find(name);
    if (it == link.end()) bail("unbound variable " + name);
    return it->second;
  }
  std::string fresh() { return "$" + std::to_string(++uniq); }
};
struct Pre { Stmts stmts; };

std::unique_ptr<Block> compile_block(const Stmts& body, const Stmts& pre, Linker& ctx);

std::unique_ptr<Block> branch_block(const Stmts& body, const Stmts& pre, Linker& ctx, bool scoped) {
  if (!scoped) return compile_block(body, pre, ctx);
  uint32_t vi = ctx.var_index;
  auto saved = ctx.link;
  auto b = compile_block(body, pre, ctx);
  ctx.var_index = vi;
  ctx.link = std::move(saved);
  return b;
}

Ctrl compile_ctrl(const St& s, const Stmts& rest, Linker& ctx) {
  Ctrl c;
  switch (s.kind) {
    case SK::Return: {
      if (!rest.empty()) bail("statements after return");
      c.kind = Ctrl::Return;
      for (auto& v : s.names)
        for (uint32_t j : ctx.get(v)) c.vars.push_back(j);
      c.ident = ctx.return_ident;
      ctx.return_idents.push_back(ctx.return_ident++);
      return c;
    }
    case SK::If: {
      // CtrlE::If (toplevel.rs:426-441): branch [0..] = else with AssertEq, default = then with AssertNe
      const Stmts& else_ = s.dflt ? *s.dflt : rest;
      const uint32_t n = (uint32_t)ctx.get(s.a).size();
      const std::string z = ctx.fresh();
      std::vector<int64_t> zeros(n, 0);
      Stmts t_pre{array(z, zeros), assert_ne(s.a, z)}, f_pre{array(z, zeros), assert_eq(s.a, z)};
      c.kind = n == 1 ? Ctrl::Choose : Ctrl::ChooseMany;
      c.vars = n == 1 ? std::vector<uint32_t>{ctx.get(s.a)[0]} : ctx.get(s.a);
      c.branches.emplace_back(std::vector<uint32_t>(n, 0u), branch_block(else_, f_pre, ctx, true));
      c.dflt = branch_block(s.arms[0].body, t_pre, ctx, false);
      return c;
    }
    case SK::Match:
    case SK::MatchMany: {
      const bool many = s.kind == SK::MatchMany;
      const Stmts* dflt = s.dflt ? s.dflt.get() : (s.exhaustive ? nullptr : &rest);
      c.kind = many ? Ctrl::ChooseMany : Ctrl::Choose;
      // fresh names are drawn in the same order as the branches' witnesses are declared
      std::vector<Stmts> pres;
      for (auto& arm : s.arms) {
        const std::string arr = ctx.fresh();
        std::vector<int64_t> keys(arm.keys.begin(), arm.keys.end());
        pres.push_back(many ? Stmts{array(arr, keys), assert_eq(s.a, arr)} : Stmts{array(arr, keys), contains(arr, s.a)});
      }
      Stmts dpre;
      if (dflt) {
        for (auto& arm : s.arms) {
          if (many) {
            const std::string arr = ctx.fresh();
            std::vector<int64_t> keys(arm.keys.begin(), arm.keys.end());
            dpre.push_back(array(arr, keys));
            dpre.push_back(assert_ne(s.a, arr));
          } else {
            for (uint32_t f : arm.keys) {
              const std::string fv = ctx.fresh();
              dpre.push_back(const_(fv, f));
              dpre.push_back(assert_ne(s.a, fv));
            }
          }
        }
      }
      c.vars = many ? ctx.get(s.a) : std::vector<uint32_t>{ctx.get(s.a)[0]};
      for (size_t i = 0; i < s.arms.size(); i++)
        c.branches.emplace_back(s.arms[i].keys, branch_block(s.arms[i].body, pres[i], ctx, true));
      if (dflt) c.dflt = branch_block(*dflt, dpre, ctx, false);
      return c;
    }
    default: bail("not a control statement");
  }
}

std::unique_ptr<Block> compile_block(const Stmts& body, const Stmts& pre, Linker& ctx) {
  Stmts stmts(pre);
  stmts.insert(stmts.end(), body.begin(), body.end());
  auto blk = std::make_unique<Block>();
  auto& ops = blk->ops;
  auto flat = [&](const std::vector<std::string>& names) {
    std::vector<uint32_t> out;
    for (auto& v : names)
      for (uint32_t j : ctx.get(v)) out.push_back(j);
    return out;
  };
  auto out_size = [](const std::vector<Var>& vs) {
    uint32_t n = 0;
    for (auto& v : vs) n += v.size;
    return n;
  };
  size_t i = 0;
  for (; i < stmts.size();) {
    const St s = stmts[i];
    bool last = false;
    switch (s.kind) {
      case SK::Const: {
        Op op{OK::Const};
        op.a = s.value;
        ops.push_back(op);
        ctx.fresh_var(s.tgt, 1);
      } break;
      case SK::Array:
        for (uint32_t f : s.values) {
          Op op{OK::Const};
          op.a = f;
          ops.push_back(op);
        }
        ctx.fresh_var(s.tgt, (uint32_t)s.values.size());
        break;
      case SK::Add: case SK::Sub: case SK::Mul: {
        const std::vector<uint32_t> xa = ctx.get(s.a), xb = ctx.get(s.b);
        if (xa.size() != xb.size()) bail("Var mismatch on arithmetic");
        for (size_t k = 0; k < xa.size(); k++) {
          Op op{s.kind == SK::Add ? OK::Add : s.kind == SK::Sub ? OK::Sub : OK::Mul};
          op.a = xa[k]; op.b = xb[k];
          ops.push_back(op);
        }
        ctx.fresh_var(s.tgt, (uint32_t)xa.size());
      } break;
      case SK::Div: {  // inv + mul (OpE::expand)
        const std::string t = ctx.fresh();
        St i1 = inv(t, s.b), m1 = mul(s.tgt, s.a, t);
        stmts[i] = i1;
        stmts.insert(stmts.begin() + i + 1, m1);
        continue;
      }
      case SK::Eq: {  // sub + not
        const std::string t = ctx.fresh();
        St s1 = sub(t, s.a, s.b), n1 = not_(s.tgt, t);
        stmts[i] = s1;
        stmts.insert(stmts.begin() + i + 1, n1);
        continue;
      }
      case SK::Inv: {
        const std::vector<uint32_t> xa = ctx.get(s.a);
        for (uint32_t u : xa) {
          Op op{OK::Inv};
          op.a = u;
          ops.push_back(op);
        }
        ctx.fresh_var(s.tgt, (uint32_t)xa.size());
      } break;
      case SK::Not: {
        Op op{OK::Not};
        op.a = ctx.get(s.a)[0];
        ops.push_back(op);
        ctx.fresh_var(s.tgt, 1);
      } break;
      case SK::AssertEq: case SK::AssertNe: {
        Op op{s.kind == SK::AssertEq ? OK::AssertEq : OK::AssertNe};
        op.xs = ctx.get(s.a); op.ys = ctx.get(s.b);
        ops.push_back(op);
      } break;
      case SK::Contains: {
        Op op{OK::Contains};
        op.xs = ctx.get(s.a); op.b = ctx.get(s.b)[0];
        ops.push_back(op);
      } break;
      case SK::Call: case SK::PreImg: {
        auto it = ctx.info->find(s.callee);
        if (it == ctx.info->end()) bail("unknown function " + s.callee);
        const FuncInfo& fi = it->second;
        if (fi.partial && !ctx.partial) bail("assertion failed: ctx.partial");  // src/lair/toplevel.rs:640-641
        Op op{s.kind == SK::Call ? OK::Call : OK::PreImg};
        op.a = fi.index;
        op.xs = flat(s.names);
        const uint32_t want_in = s.kind == SK::Call ? fi.in : fi.out, want_out = s.kind == SK::Call ? fi.out : fi.in;
        if (op.xs.size() != want_in) bail("Input mismatch on call of " + s.callee);
        if (out_size(s.outs) != want_out) bail("Output mismatch on call of " + s.callee);
        ops.push_back(op);
        for (auto& o : s.outs) ctx.fresh_var(o.name, o.size);
      } break;
      case SK::Store: {
        Op op{OK::Store};
        op.xs = flat(s.names);
        ops.push_back(op);
        ctx.fresh_var(s.tgt, 1);
      } break;
      case SK::Load: {
        Op op{OK::Load};
        op.a = out_size(s.outs);
        op.b = ctx.get(s.a)[0];
        ops.push_back(op);
        for (auto& v : s.outs) ctx.fresh_var(v.name, v.size);
      } break;
      case SK::ExternCall: {
        auto it = ctx.chips->find(s.callee);
        if (it == ctx.chips->end()) bail("unknown chip " + s.callee);
        Op op{OK::ExternCall};
        op.a = it->second.first;
        op.xs = flat(s.names);
        if (op.xs.size() != it->second.second.input_size || out_size(s.outs) != it->second.second.output_size)
          bail("extern call arity mismatch on " + s.callee);
        ops.push_back(op);
        for (auto& o : s.outs) ctx.fresh_var(o.name, o.size);
      } break;
      case SK::RangeU8: {
        Op op{OK::RangeU8};
        op.xs = flat(s.names);
        ops.push_back(op);
      } break;
      case SK::Emit:
        for (auto& v : s.names) ctx.get(v);   // must be bound; no bytecode
        break;
      case SK::Slice: {   // pure re-linking (toplevel.rs OpE::Slice)
        const std::vector<uint32_t> vals = flat(s.names);
        if (out_size(s.outs) != vals.size()) bail("Var mismatch on slice");
        size_t pos = 0;
        for (auto& o : s.outs) {
          ctx.link[o.name] = std::vector<uint32_t>(vals.begin() + pos, vals.begin() + pos + o.size);
          pos += o.size;
        }
      } break;
      default: last = true;
    }
    if (last) break;
    i++;
  }
  if (i >= stmts.size()) bail("block must end in a control statement");
  Stmts rest(stmts.begin() + i + 1, stmts.end());
  std::vector<uint32_t> saved;
  saved.swap(ctx.return_idents);
  blk->ctrl = compile_ctrl(stmts[i], rest, ctx);
  blk->return_idents = ctx.return_idents;
  if (blk->return_idents.empty()) bail("A block must have at least one return ident");
  saved.insert(saved.end(), blk->return_idents.begin(), blk->return_idents.end());
  ctx.return_idents.swap(saved);
  return blk;
}
}  // namespace

Toplevel::Toplevel(const std::vector<FuncE>& funcs, std::vector<ChipInfo> chips) : chips_(std::move(chips)) {
  std::map<std::string, std::pair<uint32_t, ChipInfo>> chip_map;
  for (uint32_t i = 0; i < chips_.size(); i++) chip_map[chips_[i].name] = {i, chips_[i]};
  std::map<std::string, FuncInfo> info;
  for (uint32_t i = 0; i < funcs.size(); i++) {
    if (info.count(funcs[i].name)) bail("duplicate function name");
    info[funcs[i].name] = FuncInfo{i, funcs[i].input_size(), funcs[i].output_size, funcs[i].partial};
  }
  for (uint32_t i = 0; i < funcs.size(); i++) {
    const FuncE& fe = funcs[i];
    Linker ctx;
    ctx.info = &info;
    ctx.chips = &chip_map;
    ctx.partial = fe.partial;
    for (auto& v : fe.input_params) ctx.fresh_var(v.name, v.size);
    auto body = compile_block(fe.body, {}, ctx);
    Func f;
    f.name = fe.name; f.invertible = fe.invertible; f.partial = fe.partial; f.index = i;
    f.input_size = fe.input_size(); f.output_size = fe.output_size;
    f.body = std::move(*body);
    funcs_.push_back(std::move(f));
    by_name_[fe.name] = i;
  }
}
const Func& Toplevel::func_by_name(const std::string& name) const {
  auto it = by_name_.find(name);
  if (it == by_name_.end()) bail("unknown function " + name);
  return funcs_[it->second];
}

// ---- column layout (src/lair/func_chip.rs:90-277) ---------------------------------------------------------
namespace {
struct LayoutState { uint32_t aux, sel, max_map; };
struct Layouter {
  const Toplevel& top;
  void op(const Op& o, std::vector<uint32_t>& deg, LayoutState& st) const {
    uint32_t op_aux = 0;
    switch (o.kind) {
      case OK::AssertNe: op_aux = (uint32_t)o.xs.size(); break;
      case OK::Contains: op_aux = (uint32_t)o.xs.size() - 1; break;
      case OK::Const: deg.push_back(0); break;
      case OK::Add: case OK::Sub: deg.push_back(std::max(deg[o.a], deg[o.b])); break;
      case OK::Mul: {
        uint32_t d = deg[o.a] + deg[o.b];
        if (d < 2) deg.push_back(d);
        else deg.push_back(1), op_aux = 1;
      } break;
      case OK::Inv:
        if (deg[o.a] == 0) deg.push_back(0);
        else deg.push_back(1), op_aux = 1;
        break;
      case OK::Not:
        if (deg[o.a] == 0) deg.push_back(0);
        else deg.push_back(1), op_aux = 2;
        break;
      case OK::Call: case OK::PreImg: {
        const Func& callee = top.func_by_index(o.a);
        uint32_t n = o.kind == OK::Call ? callee.output_size : callee.input_size;
        op_aux = n + 3;
        if (callee.partial) op_aux += DEPTH_W + DEPTH_LESS_THAN_SIZE + 3 * DEPTH_LESS_THAN_REQUIRES;
        deg.insert(deg.end(), n, 1u);
      } break;
      case OK::Store: op_aux = 4; deg.push_back(1); break;
      case OK::Load: op_aux = o.a + 3; deg.insert(deg.end(), o.a, 1u); break;
      case OK::ExternCall: {
        const ChipInfo& chip = top.chips()[o.a];
        op_aux = chip.witness_size + 3 * chip.require_size;
        // the reference extends `degrees` by aux_size here (func_chip.rs:262) while executor and trace push
        // output_size values; degrees are tracked for the outputs only
        deg.insert(deg.end(), chip.output_size, 1u);
      } break;
      case OK::RangeU8: op_aux = 3 * (((uint32_t)o.xs.size() + 1) / 2); break;
      default: break;
    }
    st.aux += op_aux;
  }
  void block(const Block& b, std::vector<uint32_t>& deg, LayoutState& st) const {
    for (auto& o : b.ops) op(o, deg, st);
    st.max_map = std::max<uint32_t>(st.max_map, (uint32_t)deg.size());
    if (b.ctrl.kind == Ctrl::Return) {
      st.sel += 1;
      return;
    }
    const size_t n = deg.size();
    const uint32_t base_aux = st.aux;
    uint32_t max_aux = st.aux;
    auto one = [&](const Block& blk) {
      st.aux = base_aux;
      block(blk, deg, st);
      deg.resize(n);
      max_aux = std::max(max_aux, st.aux);
    };
    for (auto& br : b.ctrl.branches) one(*br.second);
    if (b.ctrl.dflt) one(*b.ctrl.dflt);
    st.aux = max_aux;
  }
};
}  // namespace

LayoutSizes Toplevel::compute_layout_sizes(const Func& f) const {
  uint32_t aux = 2;
  if (f.partial) aux += DEPTH_W + 3 * ((DEPTH_W + 1) / 2);
  LayoutState st{aux, 0, f.input_size};
  std::vector<uint32_t> deg(f.input_size, 1u);
  Layouter{*this}.block(f.body, deg, st);
  LayoutSizes ls;
  ls.input = f.input_size; ls.output = f.output_size; ls.aux = st.aux; ls.sel = st.sel; ls.max_map = st.max_map;
  return ls;
}

// ---- flat program image (format: csrc/program.h) -----------------------------------------------------------
namespace {
constexpr uint32_t CTRL_RETURN = 100, CTRL_CHOOSE = 101, CTRL_CHOOSE_MANY = 102;
void ser_block(const Toplevel& top, const Block& b, std::vector<uint32_t>& deg, std::vector<uint32_t>& out) {
  const size_t start = out.size();
  out.push_back(0);  // block length in words (filled below)
  out.push_back((uint32_t)b.ops.size());
  auto put = [&](const std::vector<uint32_t>& v) { out.insert(out.end(), v.begin(), v.end()); };
  for (auto& o : b.ops) {
    const uint32_t code = (uint32_t)o.kind;
    switch (o.kind) {
      case OK::Const: out.push_back(code); out.push_back(o.a); deg.push_back(0); break;
      case OK::Add: case OK::Sub:
        out.push_back(code); out.push_back(o.a); out.push_back(o.b);
        deg.push_back(std::max(deg[o.a], deg[o.b]));
        break;
      case OK::Mul: {
        uint32_t d = deg[o.a] + deg[o.b], has_aux = d >= 2;
        out.push_back(code); out.push_back(o.a); out.push_back(o.b); out.push_back(has_aux);
        deg.push_back(has_aux ? 1 : d);
      } break;
      case OK::Inv: case OK::Not: {
        uint32_t has_aux = deg[o.a] != 0;
        out.push_back(code); out.push_back(o.a); out.push_back(has_aux);
        deg.push_back(has_aux);
      } break;
      case OK::AssertEq: case OK::AssertNe:
        out.push_back(code); out.push_back((uint32_t)o.xs.size()); put(o.xs); put(o.ys);
        break;
      case OK::Contains: out.push_back(code); out.push_back((uint32_t)o.xs.size()); put(o.xs); out.push_back(o.b); break;
      case OK::Call: case OK::PreImg: {
        const Func& callee = top.func_by_index(o.a);
        uint32_t n = o.kind == OK::Call ? callee.output_size : callee.input_size;
        out.push_back(code); out.push_back(o.a); out.push_back((uint32_t)o.xs.size()); put(o.xs);
        out.push_back(n); out.push_back(callee.partial);
        deg.insert(deg.end(), n, 1u);
      } break;
      case OK::Store: out.push_back(code); out.push_back((uint32_t)o.xs.size()); put(o.xs); deg.push_back(1); break;
      case OK::Load: out.push_back(code); out.push_back(o.a); out.push_back(o.b); deg.insert(deg.end(), o.a, 1u); break;
      case OK::ExternCall:
        out.push_back(code); out.push_back(o.a); out.push_back((uint32_t)o.xs.size()); put(o.xs);
        deg.insert(deg.end(), top.chips()[o.a].output_size, 1u);
        break;
      case OK::RangeU8: out.push_back(code); out.push_back((uint32_t)o.xs.size()); put(o.xs); break;
    }
  }
  const Ctrl& c = b.ctrl;
  if (c.kind == Ctrl::Return) {
    out.push_back(CTRL_RETURN); out.push_back(c.ident); out.push_back((uint32_t)c.vars.size()); put(c.vars);
  } else {
    const bool many = c.kind == Ctrl::ChooseMany;
    out.push_back(many ? CTRL_CHOOSE_MANY : CTRL_CHOOSE);
    out.push_back((uint32_t)c.vars.size()); put(c.vars);
    out.push_back((uint32_t)c.branches.size()); out.push_back(c.dflt ? 1u : 0u);
    const size_t n = deg.size();
    for (auto& br : c.branches) {
      out.push_back(many ? 1u : (uint32_t)br.first.size());  // choose_many: one key tuple of len(vars)
      put(br.first);
      ser_block(top, *br.second, deg, out);
      deg.resize(n);
    }
    if (c.dflt) {
      ser_block(top, *c.dflt, deg, out);
      deg.resize(n);
    }
  }
  out.push_back((uint32_t)b.return_idents.size()); put(b.return_idents);
  out[start] = (uint32_t)(out.size() - start);
}
}  // namespace

std::vector<uint32_t> Toplevel::serialize() const {
  std::vector<uint32_t> out{0x4c414952u, (uint32_t)funcs_.size(), (uint32_t)chips_.size()};  // 'LAIR'
  for (auto& c : chips_) {
    out.push_back(c.kind); out.push_back(c.input_size); out.push_back(c.output_size);
    out.push_back(c.witness_size); out.push_back(c.require_size);
  }
  const size_t table_at = out.size();
  out.resize(out.size() + funcs_.size(), 0);
  for (auto& f : funcs_) {
    out[table_at + f.index] = (uint32_t)out.size();
    LayoutSizes ls = compute_layout_sizes(f);
    out.push_back(f.input_size); out.push_back(f.output_size); out.push_back(f.partial); out.push_back(f.invertible);
    out.push_back(ls.aux); out.push_back(ls.sel); out.push_back(ls.max_map);
    std::vector<uint32_t> deg(f.input_size, 1u);
    ser_block(*this, f.body, deg, out);
  }
  return out;
}

// ---- demo_toplevel (src/lair/mod.rs:53-112) ----------------------------------------------------------------
Toplevel demo_toplevel() {
  FuncE factorial = func("factorial", {"n"}, 1, {
      const_("one", 1),
      if_("n", {sub("pred", "n", "one"), call({"m"}, "factorial", {"pred"}), mul("res", "n", "m"), ret({"res"})}),
      ret({"one"}),
  });
  FuncE fib = func("fib", {"n"}, 1, {
      const_("one", 1),
      match("n", {{{0}, {const_("zero", 0), ret({"zero"})}}, {{1}, {ret({"one"})}}}),
      sub("n_1", "n", "one"),
      call({"a"}, "fib", {"n_1"}),
      sub("n_2", "n_1", "one"),
      call({"b"}, "fib", {"n_2"}),
      add("res", "a", "b"),
      ret({"res"}),
  });
  FuncE even = func("even", {"n"}, 1, {
      const_("one", 1),
      match("n", {{{0}, {ret({"one"})}}}),
      sub("pred", "n", "one"),
      call({"res"}, "odd", {"pred"}),
      ret({"res"}),
  });
  FuncE odd = func("odd", {"n"}, 1, {
      const_("one", 1),
      match("n", {{{0}, {const_("zero", 0), ret({"zero"})}}}),
      sub("pred", "n", "one"),
      call({"res"}, "even", {"pred"}),
      ret({"res"}),
  });
  return Toplevel::new_pure({factorial, fib, even, odd});
}

// ---- surface image ---------------------------------------------------------------------------------------------
// image   := 'LSRF' nstrings { nbytes words.. }* nchips { name kind in out witness requires }* nfuncs func*
// func    := name nparams { name size }* output_size invertible partial block
// block   := nstmts stmt*
// stmt    := kind payload   (kind = SK; names are string-table indices)
namespace {
struct Reader {
  const uint32_t* w;
  size_t n, p = 0;
  std::vector<std::string> strs;
  uint32_t u() {
    if (p >= n) bail("surface image truncated");
    return w[p++];
  }
  const std::string& s() {
    uint32_t i = u();
    if (i >= strs.size()) bail("bad string index in surface image");
    return strs[i];
  }
  Stmts block() {
    uint32_t cnt = u();
    Stmts out;
    for (uint32_t i = 0; i < cnt; i++) out.push_back(stmt());
    return out;
  }
  std::vector<Var> vars() {
    uint32_t cnt = u();
    std::vector<Var> v;
    for (uint32_t i = 0; i < cnt; i++) {
      std::string name = s();
      uint32_t size = u();
      if (size == 0 || size > (1u << 16)) bail("bad variable size in surface image");
      v.emplace_back(name, size);
    }
    return v;
  }
  std::vector<std::string> names() {
    uint32_t cnt = u();
    std::vector<std::string> v;
    for (uint32_t i = 0; i < cnt; i++) v.push_back(s());
    return v;
  }
  St stmt() {
    St st;
    uint32_t k = u();
    if (k > (uint32_t)SK::Emit) bail("bad statement kind in surface image");
    st.kind = (SK)k;
    switch (st.kind) {
      case SK::Const: st.tgt = s(); st.value = u(); break;
      case SK::Array: {
        st.tgt = s();
        uint32_t cnt = u();
        for (uint32_t i = 0; i < cnt; i++) st.values.push_back(u());
      } break;
      case SK::Add: case SK::Sub: case SK::Mul: case SK::Div: case SK::Eq: st.tgt = s(); st.a = s(); st.b = s(); break;
      case SK::Inv: case SK::Not: st.tgt = s(); st.a = s(); break;
      case SK::AssertEq: case SK::AssertNe: case SK::Contains: st.a = s(); st.b = s(); break;
      case SK::Call: case SK::PreImg: case SK::ExternCall: st.outs = vars(); st.callee = s(); st.names = names(); break;
      case SK::Store: st.tgt = s(); st.names = names(); break;
      case SK::Load: st.outs = vars(); st.a = s(); break;
      case SK::RangeU8: case SK::Return: case SK::Emit: st.names = names(); break;
      case SK::Slice: st.outs = vars(); st.names = names(); break;
      case SK::If: {
        st.a = s();
        st.arms.push_back(Arm{{}, block()});
        if (u()) st.dflt = std::make_shared<Stmts>(block());
      } break;
      case SK::Match: case SK::MatchMany: {
        st.a = s();
        uint32_t na = u();
        for (uint32_t i = 0; i < na; i++) {
          Arm arm;
          uint32_t nk = u();
          for (uint32_t j = 0; j < nk; j++) arm.keys.push_back(u());
          arm.body = block();
          st.arms.push_back(std::move(arm));
        }
        if (u()) st.dflt = std::make_shared<Stmts>(block());
        st.exhaustive = u() != 0;
      } break;
    }
    return st;
  }
};
}  // namespace

std::pair<std::vector<FuncE>, std::vector<ChipInfo>> parse_surface(const uint32_t* w, size_t n) {
  Reader r{w, n};
  if (r.u() != 0x4c535246u) bail("bad surface image magic");
  uint32_t ns = r.u();
  for (uint32_t i = 0; i < ns; i++) {
    const size_t nb = r.u(), nw = (nb + 3) / 4;
    if (nw > n - r.p) bail("surface image truncated");
    r.strs.emplace_back((const char*)(w + r.p), nb);
    r.p += nw;
  }
  std::vector<ChipInfo> chips(r.u());
  for (auto& c : chips) {
    c.name = r.s(); c.kind = r.u(); c.input_size = r.u(); c.output_size = r.u(); c.witness_size = r.u(); c.require_size = r.u();
  }
  std::vector<FuncE> funcs(r.u());
  for (auto& f : funcs) {
    f.name = r.s();
    f.input_params = r.vars();
    f.output_size = r.u(); f.invertible = r.u() != 0; f.partial = r.u() != 0;
    f.body = r.block();
  }
  if (r.p != n) bail("trailing words in surface image");
  return {std::move(funcs), std::move(chips)};
}

}  // namespace host
}  // namespace lurk
