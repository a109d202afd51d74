// Instantiates the same-type collective kernels for float.
#include "launch_typed.cuh"
namespace b200c {
int launch_f32(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) { return launch_typed_impl<float>(kind, op, a, grid, s); }
}  // namespace b200c
