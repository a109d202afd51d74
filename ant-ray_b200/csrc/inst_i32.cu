// Instantiates the same-type collective kernels for int32_t.
#include "launch_typed.cuh"
namespace b200c {
int launch_i32(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) { return launch_typed_impl<int32_t>(kind, op, a, grid, s); }
}  // namespace b200c
