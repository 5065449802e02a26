// Run-time compilation of long constraint programs (quotient_kernel.cuh, LSK_JIT_PROGRAM).
//
// The interpreter pays ~24 machine instructions per program instruction (fetch, indexed branch, register file
// in local memory with computed indices); the hasher chips of a Lurk machine have 5-10 k instructions per row.
// For such a chip the program is emitted as CUDA source -- one statement per instruction, constants and
// alpha-power indices as literals, 128-statement __noinline__ chunks so that NVVM never sees one giant basic
// block (a single 7.5 k-statement block compiles for 5 minutes into 255 registers and runs slower than the
// chunked form) -- compiled by NVRTC for sm_100a and loaded through the driver API.  NVRTC and libcuda are
// bound with dlopen: the library links against neither.
//
// Compiling one hash chip takes ~30 s, so cubins are kept in a disk cache (LSK_JIT_CACHE, default
// <directory of liblurkstark.so>/jit_cache) keyed by a hash of the generated source; lsk_jit_precompile fills
// it without a GPU (the build step does that for the stock hasher programs).  Policy (env LSK_JIT): "0" never
// use compiled programs; "1" compile on a cache miss; unset: use what the cache holds, interpret the rest.
// The interpreter stays the reference path: both forms execute the same field operations in the same order,
// and the GPU tests compare their proofs word for word.
#include "jit.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>

#include "device.cuh"
#include "field.cuh"

namespace lsk {
namespace {

#include "build/jit_sources.inc"   // FIELD_CUH_SRC, QUOTIENT_KERNEL_CUH_SRC (Makefile)

// ---- NVRTC / driver API through dlopen ---------------------------------------------------------------------
struct Nvrtc {
  int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(void**) = nullptr;
  int (*CompileProgram)(void*, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(void*, size_t*) = nullptr;
  int (*GetProgramLog)(void*, char*) = nullptr;
  int (*GetCUBINSize)(void*, size_t*) = nullptr;
  int (*GetCUBIN)(void*, char*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  bool ok = false;
};
struct Driver {
  int (*ModuleLoadData)(void**, const void*) = nullptr;
  int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*ModuleUnload)(void*) = nullptr;
  int (*FuncSetAttribute)(void*, int, int) = nullptr;
  int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**,
                      void**) = nullptr;
  bool ok = false;
};
Nvrtc g_rtc;
Driver g_drv;
std::once_flag g_rtc_once, g_drv_once;

template <class F>
bool sym(void* h, F& f, const char* name) {
  f = reinterpret_cast<F>(dlsym(h, name));
  return f != nullptr;
}
Nvrtc& nvrtc() {
  std::call_once(g_rtc_once, [] {
    void* h = nullptr;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"})
      if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return;
    g_rtc.ok = sym(h, g_rtc.CreateProgram, "nvrtcCreateProgram") && sym(h, g_rtc.DestroyProgram, "nvrtcDestroyProgram") &&
               sym(h, g_rtc.CompileProgram, "nvrtcCompileProgram") && sym(h, g_rtc.GetProgramLogSize, "nvrtcGetProgramLogSize") &&
               sym(h, g_rtc.GetProgramLog, "nvrtcGetProgramLog") && sym(h, g_rtc.GetCUBINSize, "nvrtcGetCUBINSize") &&
               sym(h, g_rtc.GetCUBIN, "nvrtcGetCUBIN") && sym(h, g_rtc.Version, "nvrtcVersion");
  });
  return g_rtc;
}
Driver& driver() {
  std::call_once(g_drv_once, [] {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_drv.ok = sym(h, g_drv.ModuleLoadData, "cuModuleLoadData") && sym(h, g_drv.ModuleGetFunction, "cuModuleGetFunction") &&
               sym(h, g_drv.ModuleUnload, "cuModuleUnload") && sym(h, g_drv.LaunchKernel, "cuLaunchKernel") &&
               sym(h, g_drv.FuncSetAttribute, "cuFuncSetAttribute");
  });
  return g_drv;
}

// ---- source generation ---------------------------------------------------------------------------------------
// Option LSK_JIT_SMEM=1: register file of the compiled program in shared memory ([register][thread], conflict-free)
// instead of local memory.  Off by default: measured slower (hash4 chip, 2^21 rows: 13.0 ms vs 9.7 ms) -- the chunk
// functions reach it through generic pointers and it halves the occupancy of the wide chips.
constexpr uint32_t JIT_THREADS = 256;
uint32_t smem_bytes_for(const ChipH& c) {
  const char* e = getenv("LSK_JIT_SMEM");
  if (!e || e[0] != '1') return 0;
  const size_t bytes = (size_t)(c.n_regs ? c.n_regs : 1u) * JIT_THREADS * 4;
  return bytes <= (96u << 10) ? (uint32_t)bytes : 0u;   // at least two CTAs per SM
}
// threads per CTA of the compiled quotient kernel (LSK_JIT_QTHREADS overrides, for tuning; part of the source, hence of the
// cache key)
uint32_t quotient_threads(const ChipH& c) {
  if (const char* e = getenv("LSK_JIT_QTHREADS")) {
    long v = atol(e);
    if (v >= 64 && v <= 1024 && v % 32 == 0) return (uint32_t)v;
  }
  // measured on B200 (quotient kernels per prove, ms): Lurk fib 2^20 256 -> 12.5, 512 -> 16.5, 768 -> 12.4; Lurk lcs 2^22
  // 256 -> 67.7, 512 -> 83.9, 768 -> 59.6.  768 threads = one CTA per SM at 80 registers: a single instruction stream per SM
  // (256: three CTAs at different places of the megabyte-sized program; 512: one CTA, a third fewer warps).  The hasher
  // chips are the exception: their programs keep 40-70 values alive across chunks and want the 86 registers a 256-thread
  // CTA may have (hash4 chip at 2^20 rows: 7.2 ms at 256 threads, 8.4 at 768)
  return c.n_regs > 16 ? 256 : 768;
}
uint32_t perm_threads() {   // threads per CTA of the compiled permutation-row kernel (LSK_JIT_PTHREADS: tuning)
  if (const char* e = getenv("LSK_JIT_PTHREADS")) {
    long v = atol(e);
    if (v >= 64 && v <= 1024 && v % 32 == 0) return (uint32_t)v;
  }
  // measured on B200 (permutation-row kernels per prove, ms): Lurk fib 2^20 256 -> 5.36, 512 -> 4.85, 768 -> 4.91; Lurk
  // lcs 2^22 256 -> 37.2, 512 -> 28.9, 768 -> 24.7 (the same one-instruction-stream-per-SM effect as the quotient's)
  return 768;
}
// statements per __noinline__ chunk (LSK_JIT_CHUNK overrides, for tuning; part of the source, hence of the cache key)
uint32_t chunk_size() {
  if (const char* e = getenv("LSK_JIT_CHUNK")) {
    long v = atol(e);
    if (v >= 8 && v <= 4096) return (uint32_t)v;
  }
  return 128;   // measured on the hash4 chip, 2^21 rows: 32 -> 12.9 ms, 64 -> 11.5, 128 -> 10.7, 192 -> 10.9, 256 -> 14.5
}

// `local_def`: for a register operand, the index of the chunk-local temporary that holds its value, or -1
void operand(std::string& o, uint32_t w, const ChipH& c, long local_def) {
  const uint32_t idx = w & 0x0fffffffu;
  char buf[48];
  switch (w >> 28) {
    case 0:
      if (local_def >= 0) snprintf(buf, sizeof buf, "t%ld", local_def);
      else snprintf(buf, sizeof buf, "LSK_R(%u)", idx);
      break;
    case 1: snprintf(buf, sizeof buf, "%uu", to_mont(c.consts.at(idx) % P)); break;   // Montgomery literal
    case 2: snprintf(buf, sizeof buf, "LSK_ML(%u)", idx); break;
    case 3: snprintf(buf, sizeof buf, "LSK_MN(%u)", idx); break;
    case 4: snprintf(buf, sizeof buf, "LSK_PL(%u)", idx); break;
    case 5: snprintf(buf, sizeof buf, "LSK_PN(%u)", idx); break;
    case 6: snprintf(buf, sizeof buf, "LSK_PV(%u)", idx); break;
    default: snprintf(buf, sizeof buf, "LSK_SEL(%u)", idx); break;
  }
  o += buf;
}

std::string program_source(const ChipH& c) {
  const size_t ninstr = c.code.size() / 4;
  std::string body, calls;
  body.reserve(ninstr * 48);
  uint32_t ci = 0;
  char buf[192];
  static const char* const SIG =
      "(uint32_t* __restrict__ R, Ext& acc, const QuotientArgs& a, const QuotientRow& q)";
  const uint32_t CHUNK = chunk_size();
  // Liveness: the program is straight-line, so every register read has one reaching definition.  A value whose
  // last read lies in the chunk that defines it becomes a C++ temporary (a machine register); only values that
  // cross a chunk boundary go through the register file R[] in local memory.
  std::vector<long> reach(c.n_regs ? c.n_regs : 1, -1), last_use(ninstr, -1), src_a(ninstr, -1), src_b(ninstr, -1);
  for (size_t i = 0; i < ninstr; i++) {
    const uint32_t op = c.code[4 * i], dst = c.code[4 * i + 1], x = c.code[4 * i + 2], y = c.code[4 * i + 3];
    if ((x >> 28) == 0) {
      src_a[i] = reach[x & 0x0fffffffu];
      if (src_a[i] >= 0) last_use[src_a[i]] = (long)i;
    }
    if (op >= COP_ADD && op <= COP_MUL && (y >> 28) == 0) {
      src_b[i] = reach[y & 0x0fffffffu];
      if (src_b[i] >= 0) last_use[src_b[i]] = (long)i;
    }
    if (op != COP_ASSERT) reach[dst] = (long)i;
  }
  auto is_local = [&](long def) { return def >= 0 && (last_use[def] < 0 || (size_t)last_use[def] / CHUNK == (size_t)def / CHUNK); };
  bool asserts_in_chunk = false, pending = false;   // pending: an LSK_ASSERT_A term awaits its reducing partner
  // a chunk folds its constraints into four 64-bit sums (kept below P * 2^32, one Montgomery reduction per
  // coefficient when the chunk ends) instead of an extension-field multiply-add per constraint
  auto close_chunk = [&] {
    if (asserts_in_chunk) body += "  LSK_ASSERT_END;\n";
    body += "}\n";
    asserts_in_chunk = pending = false;
  };
  for (size_t i = 0; i < ninstr; i++) {
    if (i % CHUNK == 0) {
      if (i) close_chunk();
      snprintf(buf, sizeof buf, "__device__ __noinline__ void lsk_chunk_%zu", i / CHUNK);
      body += buf;
      body += SIG;
      body += " {\n";
      snprintf(buf, sizeof buf, "  lsk_chunk_%zu(R, acc, a, q);\n  __syncthreads();\n", i / CHUNK);
      calls += buf;
    }
    const uint32_t op = c.code[4 * i], dst = c.code[4 * i + 1], x = c.code[4 * i + 2], y = c.code[4 * i + 3];
    body += "  ";
    if (op == COP_ASSERT) {
      if (!asserts_in_chunk) body += "LSK_ASSERT_BEGIN;\n  ";
      asserts_in_chunk = true;
      snprintf(buf, sizeof buf, pending ? "LSK_ASSERT_B(%u, " : "LSK_ASSERT_A(%u, ", ci++);
      pending = !pending;
      body += buf;
      operand(body, x, c, is_local(src_a[i]) ? src_a[i] : -1);
      body += ");\n";
      continue;
    }
    if (is_local((long)i)) snprintf(buf, sizeof buf, "const uint32_t t%zu = ", i);
    else snprintf(buf, sizeof buf, "LSK_R(%u) = ", dst);
    body += buf;
    body += op == COP_ADD ? "fadd(" : op == COP_SUB ? "fsub(" : op == COP_MUL ? "fmul(" : "fneg(";
    operand(body, x, c, is_local(src_a[i]) ? src_a[i] : -1);
    if (op != COP_NEG) {
      body += ", ";
      operand(body, y, c, is_local(src_b[i]) ? src_b[i] : -1);
    }
    body += ");\n";
  }
  if (ninstr) close_chunk();
  body += "__device__ __forceinline__ void lsk_program";
  body += SIG;
  body += " {\n" + calls + "}\n";
  return body;
}


// ---- the chip's interactions (sends, then receives: the order of the device descriptor `inter`) as straight-line
// code.  The interpreter (eval_interaction / eval_affine in quotient_kernel.cuh) walks a descriptor word by word and
// multiplies every cell by its coefficient; here a coefficient of 1 is no product at all, -1 a subtraction, constants
// are Montgomery literals, and the LogUp batches of two are unrolled.  Same field operations on the same reduced
// values, hence the same words.  `V` supplies what a batch is used for: the permutation-trace row (perm_rows) or the
// LogUp constraints of the quotient.
void affine_source(std::string& o, const AffineH& a) {
  std::string acc;
  char buf[96];
  const uint32_t cst = a.cst % P;
  if (cst) {
    snprintf(buf, sizeof buf, "%uu", to_mont(cst));
    acc = buf;
  }
  for (auto& t : a.terms) {
    snprintf(buf, sizeof buf, t[0] ? "LSK_CP(%u)" : "LSK_CM(%u)", t[1]);
    const std::string cell = buf;
    const uint32_t k = t[2] % P;
    if (k == 0) continue;
    if (k == 1) {
      acc = acc.empty() ? cell : "fadd(" + acc + ", " + cell + ")";
    } else if (k == P - 1) {
      acc = acc.empty() ? "fneg(" + cell + ")" : "fsub(" + acc + ", " + cell + ")";
    } else {
      snprintf(buf, sizeof buf, "fmul(%s, %uu)", cell.c_str(), to_mont(k));
      acc = acc.empty() ? std::string(buf) : "fadd(" + acc + ", " + buf + ")";
    }
  }
  o += acc.empty() ? "0u" : acc;
}
std::string interactions_source(const ChipH& c) {
  std::string o;
  char buf[256];
  const size_t nint = c.n_interactions();
  snprintf(buf, sizeof buf, "#define LSK_JIT_NINT %zu\n", nint);
  o += buf;
  // one real function per interaction: NVVM's time grows much faster than linearly with the size of a basic block
  // (cf. the chunks of program_source), and both kernels share the one copy
  auto one = [&](size_t k) {
    const bool recv = k >= c.sends.size();
    const InteractionH& it = recv ? c.receives[k - c.sends.size()] : c.sends[k];
    snprintf(buf, sizeof buf, "__device__ __noinline__ LskInt lsk_int_%zu(const Cells c, const Ext alpha, const Ext* __restrict__ bp) "
                              "{\n  ExtAcc s = ext_acc_zero();\n  const uint32_t m = ", k);
    o += buf;
    std::string m;
    affine_source(m, it.mult);
    o += recv ? "fneg(" + m + ")" : m;
    o += ";\n";
    // sum_j beta^(j+1) * value_j as four 64-bit sums of raw products (ExtAcc, field.cuh): add / add_reduce alternate
    bool pending = false;
    for (size_t j = 0; j < it.values.size(); j++) {
      std::string val;
      affine_source(val, it.values[j]);
      if (val == "0u") continue;   // adds beta^(j+1) * 0
      snprintf(buf, sizeof buf, "  s.%s(ldg_ext(bp + %zu), ", pending ? "add_reduce" : "add", j);
      pending = !pending;
      o += buf;
      o += val + ");\n";
    }
    snprintf(buf, sizeof buf, "  const Ext r = eadd(eadd_base(alpha, %uu), s.value());\n", to_mont(it.kind % P));
    o += buf;
    o += "  return LskInt{r, m};\n}\n";
  };
  o += "struct LskInt { Ext r; uint32_t m; };\n";
  for (size_t k = 0; k < nint; k++) one(k);
  o += "template <class V>\n__device__ __forceinline__ void lsk_interactions(const Cells& c, const Ext& alpha, "
       "const Ext* __restrict__ bp, V& v) {\n";
  for (size_t b = 0; 2 * b < nint; b++) {
    snprintf(buf, sizeof buf, "  {\n    const LskInt i0 = lsk_int_%zu(c, alpha, bp);\n", 2 * b);
    o += buf;
    if (2 * b + 1 < nint) {
      snprintf(buf, sizeof buf, "    const LskInt i1 = lsk_int_%zu(c, alpha, bp);\n    v.pair(%zu, i0.r, i0.m, i1.r, i1.m);\n  }\n", 2 * b + 1, b);
    } else {
      snprintf(buf, sizeof buf, "    v.single(%zu, i0.r, i0.m);\n  }\n", b);
    }
    o += buf;
  }
  o += "}\n";
  return o;
}

// the translation unit NVRTC compiles: three defines and the kernel header, which includes "lsk_program.cuh"
std::string main_source(const ChipH& c) {
  uint32_t asserts = 0;
  for (size_t i = 0; i < c.code.size(); i += 4) asserts += c.code[i] == COP_ASSERT;
  char buf[256];
  snprintf(buf, sizeof buf,
           "#define LSK_JIT_PROGRAM 1\n#define LSK_JIT_REGS %u\n#define LSK_JIT_ASSERTS %u\n#define LSK_JIT_SMEM %u\n"
           "#define LSK_JIT_QTHREADS %u\n#define LSK_JIT_PTHREADS %u\n#include \"quotient_kernel.cuh\"\n",
           c.n_regs ? c.n_regs : 1u, asserts, smem_bytes_for(c) ? 1u : 0u, smem_bytes_for(c) ? 256u : quotient_threads(c),
           perm_threads());
  return buf;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char ch : s) h = (h ^ ch) * 1099511628211ull;
  return h;
}

std::string cache_dir() {
  if (const char* e = getenv("LSK_JIT_CACHE")) return e;
  Dl_info info;
  if (dladdr((void*)&cache_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t k = p.rfind('/');
    return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/jit_cache";
  }
  return "./jit_cache";
}

bool read_file(const std::string& path, std::vector<char>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  bool ok = n > 0 && fread(out.data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  return ok;
}
void write_file(const std::string& dir, const std::string& path, const std::vector<char>& data) {
  mkdir(dir.c_str(), 0755);
  std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), path.c_str());
  else remove(tmp.c_str());
}

}  // namespace

bool jit_eligible(const ChipH& c) {
  size_t min_instr = JIT_MIN_INSTR;
  if (const char* e = getenv("LSK_JIT_MIN_INSTR")) min_instr = (size_t)std::max(1L, atol(e));
  return c.code.size() / 4 >= min_instr;
}

// Cache entry = 48-byte header {magic, NVRTC version, key[2], payload length, payload checksum} + cubin.  The key is 128
// bits over everything that decides the machine code: the generated sources, the two device headers and the compile
// options; the header also names the NVRTC version that produced the entry (informational: a cubin is self-contained
// machine code, and the NVRTC a process happens to find -- torch ships its own -- says nothing about its validity).  An entry whose header, length or checksum does not match is a miss (it is
// recompiled under LSK_JIT=1, otherwise the chip is interpreted): a stale, truncated or foreign file never runs.
constexpr uint64_t JIT_CACHE_MAGIC = 0x4c534b514a495431ull;  // "LSKQJIT1"
static const char* const JIT_OPTS[] = {"--gpu-architecture=sm_100a", "--std=c++17"};
struct JitCacheHeader { uint64_t magic, nvrtc_version, key0, key1, length, checksum; };

bool jit_cubin(const ChipH& c, bool compile_on_miss, std::vector<char>& cubin, std::string* how) {
  const std::string program_hdr = program_source(c), main_src = main_source(c), inter_hdr = interactions_source(c);
  uint64_t key[2];
  for (int h = 0; h < 2; h++) {
    uint64_t k = h ? 0x9e3779b97f4a7c15ull : 1469598103934665603ull;
    for (const std::string& part : {program_hdr, inter_hdr, main_src, std::string(FIELD_CUH_SRC), std::string(QUOTIENT_KERNEL_CUH_SRC),
                                    std::string(JIT_OPTS[0]), std::string(JIT_OPTS[1])}) {
      k = fnv1a(std::to_string(part.size()) + ":", k);
      k = fnv1a(part, k);
      if (h) k = (k ^ (k >> 29)) * 0xbf58476d1ce4e5b9ull;
    }
    key[h] = k;
  }
  char name[96];
  snprintf(name, sizeof name, "/q_%016llx%016llx_sm100a.cubin", (unsigned long long)key[0], (unsigned long long)key[1]);
  const std::string dir = cache_dir(), path = dir + name;
  std::vector<char> file;
  if (read_file(path, file) && file.size() > sizeof(JitCacheHeader)) {
    JitCacheHeader hd;
    memcpy(&hd, file.data(), sizeof hd);
    const size_t n = file.size() - sizeof hd;
    if (hd.magic == JIT_CACHE_MAGIC && hd.key0 == key[0] && hd.key1 == key[1] && hd.length == n &&
        hd.checksum == fnv1a(std::string(file.data() + sizeof hd, n))) {
      cubin.assign(file.begin() + sizeof hd, file.end());
      if (how) *how = "cache";
      return true;
    }
  }
  if (getenv("LSK_JIT_DEBUG"))
    fprintf(stderr, "[lsk jit] miss for chip %s: %s (%zu bytes read)\n", c.name.c_str(), path.c_str(), file.size());
  if (!compile_on_miss) return false;
  if (!nvrtc().ok) fail(LSK_ERR_CUDA, "NVRTC (libnvrtc.so.12) could not be loaded");
  void* prog = nullptr;
  const char* hdr_src[4] = {FIELD_CUH_SRC, QUOTIENT_KERNEL_CUH_SRC, program_hdr.c_str(), inter_hdr.c_str()};
  const char* hdr_name[4] = {"field.cuh", "quotient_kernel.cuh", "lsk_program.cuh", "lsk_interactions.cuh"};
  if (nvrtc().CreateProgram(&prog, main_src.c_str(), "lsk_quotient_jit.cu", 4, hdr_src, hdr_name) != 0)
    fail(LSK_ERR_CUDA, "nvrtcCreateProgram failed");
  int rc = nvrtc().CompileProgram(prog, 2, JIT_OPTS);
  if (rc != 0) {
    size_t n = 0;
    nvrtc().GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtc().GetProgramLog(prog, &log[0]);
    nvrtc().DestroyProgram(&prog);
    fail(LSK_ERR_CUDA, "NVRTC failed on the constraint program of %s: %.400s", c.name.c_str(), log.c_str());
  }
  size_t n = 0;
  nvrtc().GetCUBINSize(prog, &n);
  cubin.resize(n);
  nvrtc().GetCUBIN(prog, cubin.data());
  nvrtc().DestroyProgram(&prog);
  {
    int vmaj = 0, vmin = 0;
    if (nvrtc().Version) nvrtc().Version(&vmaj, &vmin);
    const uint64_t rtc_version = (uint64_t)vmaj * 1000 + (uint64_t)vmin;
    JitCacheHeader hd{JIT_CACHE_MAGIC, rtc_version, key[0], key[1], cubin.size(), fnv1a(std::string(cubin.data(), cubin.size()))};
    std::vector<char> out(sizeof hd + cubin.size());
    memcpy(out.data(), &hd, sizeof hd);
    memcpy(out.data() + sizeof hd, cubin.data(), cubin.size());
    write_file(dir, path, out);
  }
  if (how) *how = "compiled";
  return true;
}

JitKernel* jit_quotient(const ChipH& c) {
  const char* env = getenv("LSK_JIT");
  if (env && env[0] == '0') return nullptr;
  if (!jit_eligible(c)) return nullptr;
  std::vector<char> cubin;
  if (!jit_cubin(c, env && env[0] == '1', cubin, nullptr)) return nullptr;
  if (!driver().ok) fail(LSK_ERR_CUDA, "libcuda.so.1 could not be loaded for a compiled constraint program");
  auto k = new JitKernel;
  k->smem_bytes = smem_bytes_for(c);
  k->qthreads = k->smem_bytes ? 256u : quotient_threads(c);
  k->pthreads = perm_threads();
  int rc = driver().ModuleLoadData(&k->module, cubin.data());
  if (rc == 0) rc = driver().ModuleGetFunction(&k->function, k->module, "lsk_quotient_jit");
  if (rc == 0) rc = driver().ModuleGetFunction(&k->perm_function, k->module, "lsk_perm_rows_jit");
  if (rc == 0 && k->smem_bytes > (48u << 10))
    rc = driver().FuncSetAttribute(k->function, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)k->smem_bytes);
  if (rc != 0) {
    delete k;
    fail(LSK_ERR_CUDA, "loading the compiled constraint program of %s failed (CUresult %d)", c.name.c_str(), rc);
  }
  return k;
}

void jit_free(JitKernel* k) {
  if (!k) return;
  if (k->module && g_drv.ok) g_drv.ModuleUnload(k->module);
  delete k;
}

void jit_launch(lsk_ctx* ctx, const JitKernel* k, const QuotientArgs& a, const QuotientConsts& qc, const uint32_t* inter,
                uint32_t nint, uint32_t pw) {
  const size_t nrows = a.row_end - a.row_begin;
  void* params[] = {(void*)&a, (void*)&qc, (void*)&inter, (void*)&nint, (void*)&pw};
  // (a two-rows-per-thread variant -- every chunk run for row A, then, instructions cached, for row B -- was measured
  // and removed: quotient 13.7 -> 17.9 ms on Lurk fib 2^20, 68.8 -> 95.8 ms on lcs 2^22)
  int rc = driver().LaunchKernel(k->function, div_up(nrows, k->qthreads), 1, 1, k->qthreads, 1, 1, k->smem_bytes, ctx->stream,
                                 params, nullptr);
  if (rc != 0) fail(LSK_ERR_CUDA, "cuLaunchKernel of a compiled constraint program failed (CUresult %d)", rc);
}


void jit_launch_perm(lsk_ctx* ctx, const JitKernel* k, const PermRowsArgs& a) {
  const size_t nrows = a.row_end - a.row_begin;
  void* params[] = {(void*)&a};
  int rc = driver().LaunchKernel(k->perm_function, div_up(nrows, k->pthreads), 1, 1, k->pthreads, 1, 1, 0, ctx->stream, params,
                                 nullptr);
  if (rc != 0) fail(LSK_ERR_CUDA, "cuLaunchKernel of a compiled permutation-trace kernel failed (CUresult %d)", rc);
}

}  // namespace lsk
