// Instantiates the same-type collective kernels for uint32_t.
#include "launch_typed.cuh"
namespace b200c {
int launch_u32(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) { return launch_typed_impl<uint32_t>(kind, op, a, grid, s); }
}  // namespace b200c
