// Run-time compiled constraint programs (jit.cu; kernel in quotient_kernel.cuh).
#pragma once
#include <string>
#include <vector>

#include "machine.h"
#include "quotient_kernel.cuh"

struct lsk_ctx;

namespace lsk {

constexpr size_t JIT_MIN_INSTR = 16;   // shorter programs are interpreted (the fixed per-row work dominates them); env LSK_JIT_MIN_INSTR

struct JitKernel {
  void* module = nullptr;    // CUmodule
  void* function = nullptr;  // CUfunction of lsk_quotient_jit
  void* perm_function = nullptr;  // CUfunction of lsk_perm_rows_jit (the chip's LogUp permutation-trace rows)
  uint32_t qthreads = 768;   // threads per CTA the quotient kernel was compiled for
  uint32_t pthreads = 768;   // ... and the permutation-row kernel
  uint32_t smem_bytes = 0;   // dynamic shared memory of a launch: the program's register file, or 0 (local memory)
};

bool jit_eligible(const ChipH& c);
// cubin of the chip's compiled quotient kernel: from the disk cache, else (compile_on_miss) from NVRTC, which
// also stores it.  No GPU needed.  false: not cached and not compiled.
bool jit_cubin(const ChipH& c, bool compile_on_miss, std::vector<char>& cubin, std::string* how);
// nullptr when the policy (env LSK_JIT, cache contents) says "interpret"
JitKernel* jit_quotient(const ChipH& c);
void jit_free(JitKernel* k);
void jit_launch(lsk_ctx* ctx, const JitKernel* k, const QuotientArgs& a, const QuotientConsts& qc, const uint32_t* inter,
                uint32_t nint, uint32_t pw);
void jit_launch_perm(lsk_ctx* ctx, const JitKernel* k, const PermRowsArgs& a);

}  // namespace lsk
