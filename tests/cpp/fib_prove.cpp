// End to end from C++ through the C ABI, the shape of the reference's prover test
// (src/lair/lair_chip.rs:246-276: build the toplevel, execute, prove, verify) and of benches/fib.rs:136-163.
// Needs a B200; run by tests/test_gpu_prover.py::test_cpp_end_to_end.   usage: fib_prove [n]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "host_lair.h"
#include "lurk_stark.h"

using namespace lurk::host;

#define OK(call)                                                          \
  do {                                                                    \
    int rc_ = (call);                                                     \
    if (rc_ != LSK_OK) {                                                  \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lsk_last_error());   \
      return 1;                                                           \
    }                                                                     \
  } while (0)

int main(int argc, char** argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 1000;
  Toplevel toplevel = demo_toplevel();                       // let toplevel = demo_toplevel::<F>();
  const Func& fib = toplevel.func_by_name("fib");            // let chip = FuncChip::from_name("fib", &toplevel);
  std::vector<uint32_t> image = toplevel.serialize(), blob = chip_vector_blob(toplevel, fib.index);

  lsk_ctx* ctx = nullptr;
  lsk_program* prog = nullptr;
  lsk_record* rec = nullptr;
  lsk_machine* machine = nullptr;
  OK(lsk_ctx_create(0, &ctx));
  OK(lsk_program_load(image.data(), image.size(), &prog));
  OK(lsk_record_new(prog, &rec));                             // let mut queries = QueryRecord::new(&toplevel);
  uint32_t args[1] = {n}, out[1] = {0};
  OK(lsk_execute(prog, rec, fib.index, args, 1, out, 1));     // toplevel.execute(fib, &args, &mut queries, None)
  OK(lsk_machine_new(ctx, prog, fib.index, blob.data(), blob.size(), 100, 16, &machine));   // new_machine + setup
  uint32_t* proof = nullptr;
  size_t len = 0;
  OK(lsk_prove(machine, rec, 1u << 22, &proof, &len));        // machine.prove::<LocalProver<_, _>>(&pk, shard, ..)
  OK(lsk_verify(machine, proof, len));                        // machine.verify(&vk, &proof, &mut challenger)
  proof[len / 2] ^= 1;                                        // a tampered proof must be rejected
  if (lsk_verify(machine, proof, len) != LSK_ERR_VERIFY) {
    fprintf(stderr, "tampered proof accepted\n");
    return 1;
  }
  printf("fib(%u) = %u mod p; proof of %zu words verified\n", n, out[0], len);
  lsk_proof_free(proof);
  lsk_machine_free(machine);
  lsk_record_free(rec);
  lsk_program_free(prog);
  lsk_ctx_destroy(ctx);
  return 0;
}
