"""Experiment behind csrc/jit.cu: emit the hash4 constraint program as CUDA (CHUNK statements per __noinline__
function, env CHUNK) and compile it with NVRTC, printing the compile time; writes k.cu / k.cubin to the cwd."""
import sys, time, ctypes, re
sys.path.insert(0,'/root/repo')
from lurk_b200.programs import hash_toplevel
from lurk_b200.air import build_chip_vector
from lurk_b200.execute import FuncChip
top = hash_toplevel()
f = top.func_by_name("cons_chain")
chips = build_chip_vector(FuncChip(f, top))
name = sys.argv[1] if len(sys.argv) > 1 else "Func[hash4]"
c = [c for c in chips if c.name == name][0]
code = c.code
def operand(w):
    k, i = w >> 28, w & 0x0fffffff
    return {0: f"r{i}", 1: f"C({i})", 2: f"ML({i})", 3: f"MN({i})", 4: f"PL({i})", 5: f"PN({i})", 6: f"PV({i})", 7: f"sel{i}"}[k]
body = []
ci = 0
for i in range(len(code)//4):
    op, d, a, b = code[4*i:4*i+4]
    if op == 1: body.append(f"r{d} = fadd({operand(a)}, {operand(b)});")
    elif op == 2: body.append(f"r{d} = fsub({operand(a)}, {operand(b)});")
    elif op == 3: body.append(f"r{d} = fmul({operand(a)}, {operand(b)});")
    elif op == 4: body.append(f"r{d} = fsub(0u, {operand(a)});")
    else:
        body.append(f"ASSERT({ci}, {operand(a)});"); ci += 1
CH=int(__import__("os").environ.get("CHUNK","256"))
body=[re.sub(r"\br(\d+)\b", r"R[\1]", b) for b in body]
funcs=[]
calls=[]
for k in range(0,len(body),CH):
    funcs.append("__device__ __noinline__ void chunk%d(uint32_t* __restrict__ R, Ext& acc, const uint32_t* __restrict__ main, size_t mstride, size_t r, size_t rn, const uint32_t* __restrict__ consts, const Ext* __restrict__ apow, uint32_t n_total, uint32_t sel0, uint32_t sel1, uint32_t sel2) {\n  %s\n}\n" % (k//CH, "\n  ".join(body[k:k+CH])))
    calls.append("chunk%d(R, acc, main, mstride, r, rn, consts, apow, n_total, sel0, sel1, sel2);" % (k//CH))
regs="uint32_t R[%d];" % c.n_regs
body=calls
FUNCS="".join(funcs)
src = r'''
typedef unsigned int uint32_t; typedef unsigned long long uint64_t; typedef unsigned long size_t;
#define P 0x78000001u
#define PINV 0x88000001u
struct Ext { uint32_t c[4]; };
__device__ __forceinline__ uint32_t fmul(uint32_t a, uint32_t b) { uint64_t x = (uint64_t)a * b; uint32_t t = (uint32_t)x * PINV; uint32_t hu = __umulhi(t, P); uint32_t r = (uint32_t)(x >> 32) - hu; return min(r, r + P); }
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b) { uint32_t s = a + b; return min(s, s - P); }
__device__ __forceinline__ uint32_t fsub(uint32_t a, uint32_t b) { uint32_t s = a - b; return min(s, s + P); }
#define C(i) __ldg(consts + i)
#define ML(i) main[(size_t)(i) * mstride + r]
#define MN(i) main[(size_t)(i) * mstride + rn]
#define PL(i) 0u
#define PN(i) 0u
#define PV(i) 0u
#define ASSERT(ci, v) { const Ext ap = apow[n_total - 1 - (ci)]; uint32_t x = (v); acc.c[0] = fadd(acc.c[0], fmul(ap.c[0], x)); acc.c[1] = fadd(acc.c[1], fmul(ap.c[1], x)); acc.c[2] = fadd(acc.c[2], fmul(ap.c[2], x)); acc.c[3] = fadd(acc.c[3], fmul(ap.c[3], x)); }
FUNCS_HERE
extern "C" __global__ void __launch_bounds__(256) qk(const uint32_t* __restrict__ main, size_t mstride, const uint32_t* __restrict__ consts, const Ext* __restrict__ apow, uint32_t n_total, uint32_t* out, size_t n2, uint32_t sel0, uint32_t sel1, uint32_t sel2) {
  size_t r = (size_t)blockIdx.x * 256 + threadIdx.x; if (r >= n2) return; size_t rn = (r + 2) % n2;
  Ext acc = {0,0,0,0};
  REGS
  BODY
  for (int e = 0; e < 4; e++) out[(size_t)e * n2 + r] = acc.c[e];
}
'''.replace("FUNCS_HERE", FUNCS).replace("REGS", regs).replace("BODY", "\n  ".join(body))
open("k.cu","w").write(src)
print(name, "statements", len(body), "regs", c.n_regs, "source bytes", len(src))
nv = ctypes.CDLL("libnvrtc.so")
prog = ctypes.c_void_p()
assert nv.nvrtcCreateProgram(ctypes.byref(prog), src.encode(), b"k.cu", 0, None, None) == 0
opts = [b"--gpu-architecture=sm_100a", b"--std=c++17"] + [o.encode() for o in sys.argv[2:]]
arr = (ctypes.c_char_p * len(opts))(*opts)
t0 = time.time()
rc = nv.nvrtcCompileProgram(prog, len(opts), arr)
dt = time.time() - t0
sz = ctypes.c_size_t(); nv.nvrtcGetProgramLogSize(prog, ctypes.byref(sz)); log = ctypes.create_string_buffer(sz.value); nv.nvrtcGetProgramLog(prog, log)
print("rc", rc, "compile s", round(dt, 2), log.value.decode()[:500])
nv.nvrtcGetCUBINSize(prog, ctypes.byref(sz)); print("cubin bytes", sz.value)
cub = ctypes.create_string_buffer(sz.value); nv.nvrtcGetCUBIN(prog, cub); open("k.cubin","wb").write(cub.raw)
