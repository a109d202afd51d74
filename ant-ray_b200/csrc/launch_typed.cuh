// Per-dtype launch tables.  Each inst_<dtype>.cu instantiates this for one storage type so the
// translation units compile in parallel.
#pragma once
#include "coll_kernels.cuh"

namespace b200c {

enum Kind { KIND_ONESHOT = 0, KIND_TWOSHOT = 1, KIND_REDUCESCATTER = 2, KIND_REDUCE = 3, KIND_LL = 4 };

template <typename T, int OP, int WT>
static int launch_kind_w(int kind, const CollArgs& a, int grid, cudaStream_t s) {
  switch (kind) {
    case KIND_ONESHOT: k_allreduce_oneshot<T, T, OP, WT><<<grid, kThreads, 0, s>>>(a); break;
    case KIND_TWOSHOT: k_allreduce_twoshot<T, T, OP, WT><<<grid, kThreads, 0, s>>>(a); break;
    case KIND_REDUCESCATTER: k_reducescatter<T, OP, WT><<<grid, kThreads, 0, s>>>(a); break;
    case KIND_REDUCE: k_reduce<T, OP, WT><<<grid, kThreads, 0, s>>>(a); break;
    default: return B200C_EINVAL;
  }
  return B200C_OK;
}

// the reducing kernels exist once per world size 2 / 4 / 8 and once for the other sizes (WT = 0)
template <typename T, int OP>
static int launch_kind(int kind, const CollArgs& a, int grid, cudaStream_t s) {
  if (kind == KIND_LL) {
    k_allreduce_ll<T, OP><<<grid, kLLThreads, 0, s>>>(a);
    return B200C_OK;
  }
  switch (a.c.world) {
    case 2: return launch_kind_w<T, OP, 2>(kind, a, grid, s);
    case 4: return launch_kind_w<T, OP, 4>(kind, a, grid, s);
    case 8: return launch_kind_w<T, OP, 8>(kind, a, grid, s);
    default: return launch_kind_w<T, OP, 0>(kind, a, grid, s);
  }
}

template <typename T>
static int launch_typed_impl(int kind, int op, const CollArgs& a, int grid, cudaStream_t s) {
  switch (op) {
    case B200C_SUM: case B200C_AVG: return launch_kind<T, B200C_SUM>(kind, a, grid, s);
    case B200C_PROD: return launch_kind<T, B200C_PROD>(kind, a, grid, s);
    case B200C_MAX: return launch_kind<T, B200C_MAX>(kind, a, grid, s);
    case B200C_MIN: return launch_kind<T, B200C_MIN>(kind, a, grid, s);
    default: return B200C_EINVAL;
  }
}

#define B200C_DECLARE_TYPED(name) int launch_##name(int kind, int op, const CollArgs& a, int grid, cudaStream_t s)
B200C_DECLARE_TYPED(i8);
B200C_DECLARE_TYPED(u8);
B200C_DECLARE_TYPED(i32);
B200C_DECLARE_TYPED(u32);
B200C_DECLARE_TYPED(i64);
B200C_DECLARE_TYPED(u64);
B200C_DECLARE_TYPED(f16);
B200C_DECLARE_TYPED(f32);
B200C_DECLARE_TYPED(f64);
B200C_DECLARE_TYPED(bf16);

}  // namespace b200c
