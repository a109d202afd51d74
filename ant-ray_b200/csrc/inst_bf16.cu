// Instantiates the same-type collective kernels for bf16_t.
#include "launch_typed.cuh"
namespace b200c {
int launch_bf16(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) { return launch_typed_impl<bf16_t>(kind, op, a, grid, s); }
}  // namespace b200c
