// placeholder translation unit: prover stages are added incrementally
#include "device.cuh"
namespace lsk {
void init_prover_kernels() {}
}
