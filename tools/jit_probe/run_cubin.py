"""Times a standalone cubin of the hash4 constraint program on random data (the experiment behind csrc/jit.cu:
monolithic vs chunked code generation; cubins from gen_chunked.py).  usage: run_cubin.py a.cubin [b.cubin ...]"""
import sys, ctypes, numpy as np, torch
from cuda.bindings import driver as cu
P = 2013265921
torch.cuda.init(); torch.zeros(1, device="cuda")
W, L = 655, 21
n2 = 1 << L
main = torch.randint(0, P, (W, n2), dtype=torch.int32, device="cuda")
consts = torch.randint(0, P, (4096,), dtype=torch.int32, device="cuda")
apow = torch.randint(0, P, (4096, 4), dtype=torch.int32, device="cuda")
out = torch.zeros((4, n2), dtype=torch.int32, device="cuda")
def chk(r):
    if isinstance(r, tuple):
        if int(r[0]) != 0: raise RuntimeError(str(r[0]))
        return r[1] if len(r) == 2 else r[1:]
for name in sys.argv[1:]:
    data = open(name, "rb").read()
    mod = chk(cu.cuModuleLoadData(data))
    fn = chk(cu.cuModuleGetFunction(mod, b"qk"))
    args = [ctypes.c_void_p(main.data_ptr()), ctypes.c_size_t(n2), ctypes.c_void_p(consts.data_ptr()), ctypes.c_void_p(apow.data_ptr()),
            ctypes.c_uint32(700), ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n2), ctypes.c_uint32(5), ctypes.c_uint32(7), ctypes.c_uint32(9)]
    ptrs = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
    stream = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        chk(cu.cuLaunchKernel(fn, n2 // 256, 1, 1, 256, 1, 1, 0, stream, ctypes.addressof(ptrs), 0))
        b.record(); torch.cuda.synchronize()
        print(name, "ms", round(a.elapsed_time(b), 3), "checksum", int(out.sum()), flush=True)
