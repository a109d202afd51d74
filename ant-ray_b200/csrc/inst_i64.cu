// Instantiates the same-type collective kernels for int64_t.
#include "launch_typed.cuh"
namespace b200c {
int launch_i64(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) { return launch_typed_impl<int64_t>(kind, op, a, grid, s); }
}  // namespace b200c
