// Instantiates the same-type collective kernels for int8_t.
#include "launch_typed.cuh"
namespace b200c {
int launch_i8(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) { return launch_typed_impl<int8_t>(kind, op, a, grid, s); }
}  // namespace b200c
