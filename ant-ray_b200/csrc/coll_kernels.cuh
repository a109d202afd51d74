// Collective kernels over peer-mapped arenas (hand-written for sm_100a; no NCCL on this path).
//
// Partitioning: a piece of n elements is cut into `world` rank chunks of
// `chunk` elements; each chunk is cut into per-block tiles of `tile` elements.  Block b of
// every rank works on tile b of every chunk, so cross-rank dependencies are only between
// blocks with the same index and are carried by the flag pair [block][src] in the signal pad.
// No intra-grid synchronisation exists, so blocks need not be co-resident.
//
// Staging is double-buffered by sequence parity (a.seq & 1); coll_prologue() makes the reuse
// safe for asymmetric ops as well.
#pragma once
#include "dev_common.cuh"

namespace b200c {

template <typename T>
__device__ __forceinline__ T* staging_ptr(const DevComm& c, int rank, uint32_t seq, size_t byte_off) {
  return reinterpret_cast<T*>(c.arena[rank] + c.off_staging + (size_t)(seq & 1) * c.staging_bytes + byte_off);
}
// clip [lo, hi) against n, return count
__device__ __forceinline__ size_t clip_count(size_t lo, size_t hi, size_t n) {
  if (lo >= n) return 0;
  return (hi < n ? hi : n) - lo;
}

// ---------------------------------------------------------------------------------------------
// one-shot allreduce: push the whole buffer to every peer, reduce locally.  One flag round.
// staging slot s (n_pad elements of TW) on rank j holds rank s's data.
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TW, int OP>
__global__ void __launch_bounds__(kThreads, 2) k_allreduce_oneshot(CollArgs a) {
  const DevComm& c = a.c;
  const int r = c.rank, W = c.world;
  if (!coll_prologue(a)) return;
  const size_t t0 = (size_t)blockIdx.x * a.tile;
  const size_t cnt = clip_count(t0, t0 + a.tile, a.n);
  const TI* in = static_cast<const TI*>(a.in);
  TI* out = static_cast<TI*>(a.out);
  const size_t slot_bytes = a.chunk * sizeof(TW);  // chunk == padded n for one-shot
  if (cnt) {
    for (int k = 1; k < W; k++) {
      int j = r + k; if (j >= W) j -= W;
      TW* dst = staging_ptr<TW>(c, j, a.seq, (size_t)r * slot_bytes) + t0;
      move_tile<TI, TW, false>(dst, in + t0, cnt);
    }
  }
  block_signal_all(kOffFlagA, a.seq, c);
  if (!block_wait_all(my_flags(kOffFlagA, c), a.seq, c, 1)) return;
  check_signature(a);
  if (cnt) {
    reduce_tile<TI, TW, OP>(a, staging_ptr<TW>(c, r, a.seq, 0) + t0, a.chunk, r, in + t0, nullptr, out + t0, cnt);
  }
}

// ---------------------------------------------------------------------------------------------
// two-shot allreduce.
//   A: push tile b of chunk j to rank j's staging slot [r]          (NVLink egress, stores)
//   B: reduce the W contributions of the own chunk in rank order; result -> own slot [r] + out
//   C: pull every other rank's reduced tile                          (NVLink ingress, loads)
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TW, int OP>
__global__ void __launch_bounds__(kThreads, 2) k_allreduce_twoshot(CollArgs a) {
  const DevComm& c = a.c;
  const int r = c.rank, W = c.world;
  if (!coll_prologue(a)) return;
  const size_t t0 = (size_t)blockIdx.x * a.tile;
  const size_t t1 = (t0 + a.tile < a.chunk) ? t0 + a.tile : a.chunk;
  const TI* in = static_cast<const TI*>(a.in);
  TI* out = static_cast<TI*>(a.out);
  const size_t slot_bytes = a.chunk * sizeof(TW);
  // ---- A
  for (int k = 1; k < W; k++) {
    int j = r + k; if (j >= W) j -= W;
    size_t lo = (size_t)j * a.chunk + t0;
    size_t cnt = t0 < t1 ? clip_count(lo, (size_t)j * a.chunk + t1, a.n) : 0;
    if (cnt) move_tile<TI, TW, false>(staging_ptr<TW>(c, j, a.seq, (size_t)r * slot_bytes) + t0, in + lo, cnt);
  }
  block_signal_all(kOffFlagA, a.seq, c);
  if (!block_wait_all(my_flags(kOffFlagA, c), a.seq, c, 1)) return;
  check_signature(a);
  // ---- B
  {
    size_t lo = (size_t)r * a.chunk + t0;
    size_t cnt = t0 < t1 ? clip_count(lo, (size_t)r * a.chunk + t1, a.n) : 0;
    if (cnt) {
      reduce_tile<TI, TW, OP>(a, staging_ptr<TW>(c, r, a.seq, 0) + t0, a.chunk, r, in + lo, staging_ptr<TW>(c, r, a.seq, (size_t)r * slot_bytes) + t0, out + lo, cnt);
    }
  }
  block_signal_all(kOffFlagB, a.seq, c);
  if (!block_wait_all(my_flags(kOffFlagB, c), a.seq, c, 2)) return;
  // ---- C
  for (int k = 1; k < W; k++) {
    int j = r + k; if (j >= W) j -= W;
    size_t lo = (size_t)j * a.chunk + t0;
    size_t cnt = t0 < t1 ? clip_count(lo, (size_t)j * a.chunk + t1, a.n) : 0;
    if (cnt) move_tile<TW, TI, true>(out + lo, staging_ptr<TW>(c, j, a.seq, (size_t)j * slot_bytes) + t0, cnt);
  }
}

// ---------------------------------------------------------------------------------------------
// reducescatter: in_ptrs[j] (n elements) is this rank's contribution to rank j.
// reduce (root): every non-root pushes to root; root folds; root then releases the others.
// Both are phases A+B of two-shot with full-size chunks.
// ---------------------------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(kThreads, 2) k_reducescatter(CollArgs a) {
  const DevComm& c = a.c;
  const int r = c.rank, W = c.world;
  if (!coll_prologue(a)) return;
  const size_t t0 = (size_t)blockIdx.x * a.tile;
  const size_t cnt = clip_count(t0, t0 + a.tile, a.n);
  const size_t slot_bytes = a.chunk * sizeof(T);
  if (cnt) {
    for (int k = 1; k < W; k++) {
      int j = r + k; if (j >= W) j -= W;
      copy_tile<T, false>(staging_ptr<T>(c, j, a.seq, (size_t)r * slot_bytes) + t0, static_cast<const T*>(a.in_ptrs[j]) + t0, cnt);
    }
  }
  block_signal_all(kOffFlagA, a.seq, c);
  if (!block_wait_all(my_flags(kOffFlagA, c), a.seq, c, 1)) return;
  check_signature(a);
  if (cnt) {
    reduce_tile<T, T, OP>(a, staging_ptr<T>(c, r, a.seq, 0) + t0, a.chunk, r, static_cast<const T*>(a.in_ptrs[r]) + t0, nullptr, static_cast<T*>(a.out) + t0, cnt);
  }
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads, 2) k_reduce(CollArgs a) {
  const DevComm& c = a.c;
  const int r = c.rank, root = a.root;
  if (!coll_prologue(a)) return;
  const size_t t0 = (size_t)blockIdx.x * a.tile;
  const size_t cnt = clip_count(t0, t0 + a.tile, a.n);
  const size_t slot_bytes = a.chunk * sizeof(T);
  if (r != root) {
    if (cnt) copy_tile<T, false>(staging_ptr<T>(c, root, a.seq, (size_t)r * slot_bytes) + t0, static_cast<const T*>(a.in) + t0, cnt);
    block_signal_one(kOffFlagA, a.seq, c, root);
    // wait for root's release: completion of any rank then implies every rank has arrived
    block_wait_one(my_flags(kOffFlagB, c) + root, a.seq, c, root, 2);
  } else {
    if (!block_wait_all(my_flags(kOffFlagA, c), a.seq, c, 1)) return;
    check_signature(a);
    if (cnt) {
      reduce_tile<T, T, OP>(a, staging_ptr<T>(c, r, a.seq, 0) + t0, a.chunk, r, static_cast<const T*>(a.in) + t0, nullptr, static_cast<T*>(a.out) + t0, cnt);
    }
    block_signal_all(kOffFlagB, a.seq, c);
  }
}

// ---------------------------------------------------------------------------------------------
// NVLS allreduce (SUM; f32 / bf16 / f16): the NVSwitch reduces (multimem.ld_reduce) and
// broadcasts (multimem.st).  Data must sit at the same arena offset on every rank:
//   staged   : A copies the user tensor into staging, C copies the result back;
//   symmetric: the tensor already lives in the symmetric region, A and C are flag-only.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Multimem;
template <> struct Multimem<float> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Multimem<bf16_t> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Multimem<f16_t> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
__device__ __forceinline__ void multimem_st16(void* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename TI, typename TW>
__global__ void __launch_bounds__(kThreads, 2) k_allreduce_nvls(CollArgs a) {
  const DevComm& c = a.c;
  const int r = c.rank, W = c.world;
  if (!coll_prologue(a)) return;
  constexpr int V = 16 / sizeof(TW);
  const size_t t0 = (size_t)blockIdx.x * a.tile;
  const size_t t1 = (t0 + a.tile < a.chunk) ? t0 + a.tile : a.chunk;
  const TI* in = static_cast<const TI*>(a.in);
  TI* out = static_cast<TI*>(a.out);
  // byte offset (inside the arena) of element 0 of the buffer the switch works on
  const size_t base_off = a.symmetric ? a.sym_off : c.off_staging + (size_t)(a.seq & 1) * c.staging_bytes;
  TW* mine = reinterpret_cast<TW*>(c.arena[r] + base_off);
  // ---- A: stage in (zero-pad the last vector so the switch reduces defined values)
  if (!a.symmetric && t0 < t1) {
    for (int j = 0; j < W; j++) {
      size_t lo = (size_t)j * a.chunk + t0, hi = (size_t)j * a.chunk + t1;
      size_t cnt = clip_count(lo, hi, a.n);
      if (cnt) move_tile<TI, TW, false>(mine + lo, in + lo, cnt);
      size_t end = lo + cnt, padded = (end + V - 1) / V * V;
      if (cnt && padded > end && padded <= hi) {
        TW z = Traits<TW>::from_acc((typename Traits<TW>::A)0);
        for (size_t k = end + threadIdx.x; k < padded; k += kThreads) mine[k] = z;
      }
    }
  }
  block_signal_all(kOffFlagA, a.seq, c);
  if (!block_wait_all(my_flags(kOffFlagA, c), a.seq, c, 1)) return;
  check_signature(a);
  // ---- B: in-switch reduce of the own chunk's tile, broadcast back in place
  if (t0 < t1) {
    size_t lo = (size_t)r * a.chunk + t0;
    size_t cnt = clip_count(lo, (size_t)r * a.chunk + t1, a.n);
    size_t nv = (cnt + V - 1) / V;
    char* mc = c.mc_arena + base_off + lo * sizeof(TW);
    size_t i = threadIdx.x;
    for (; i + (kUnroll - 1) * kThreads < nv; i += kUnroll * kThreads) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; u++) v[u] = Multimem<TW>::ld_reduce(mc + (i + u * kThreads) * 16);
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        if (a.has_scale) {
          Pack16<TW> p; p.u = v[u];
#pragma unroll
          for (int e = 0; e < V; e++) p.e[e] = Traits<TW>::from_acc(Traits<TW>::to_acc(p.e[e]) * a.scale);
          v[u] = p.u;
        }
        multimem_st16(mc + (i + u * kThreads) * 16, v[u]);
      }
    }
    for (; i < nv; i += kThreads) {
      uint4 v = Multimem<TW>::ld_reduce(mc + i * 16);
      if (a.has_scale) {
        Pack16<TW> p; p.u = v;
#pragma unroll
        for (int e = 0; e < V; e++) p.e[e] = Traits<TW>::from_acc(Traits<TW>::to_acc(p.e[e]) * a.scale);
        v = p.u;
      }
      multimem_st16(mc + i * 16, v);
    }
  }
  block_signal_all(kOffFlagB, a.seq, c);
  if (!block_wait_all(my_flags(kOffFlagB, c), a.seq, c, 2)) return;
  // ---- C: stage out
  if (!a.symmetric && t0 < t1) {
    for (int j = 0; j < W; j++) {
      size_t lo = (size_t)j * a.chunk + t0;
      size_t cnt = clip_count(lo, (size_t)j * a.chunk + t1, a.n);
      if (cnt) move_tile<TW, TI, true>(out + lo, mine + lo, cnt);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Pipelined staged NVLS allreduce (large plain tensors).  The CTA is split into three warp-
// specialised roles that run concurrently on sub-tiles of the CTA's slab:
//   IN  (192 threads): user tensor -> staging (every chunk's slice of sub-tile k), then raises
//                      pipeA[block][rank] = base+k+1 on every rank (own pad included);
//   RED (128 threads): waits pipeA from all W ranks, multimem.ld_reduce + multimem.st of the OWN
//                      chunk's slice of sub-tile k, then raises pipeB the same way;
//   OUT (192 threads): waits pipeB from all W ranks, staging -> user tensor.
// The local HBM copies of sub-tiles k+1 / k-1 therefore overlap the switch traffic of sub-tile k,
// instead of three back-to-back phases as in k_allreduce_nvls.  Roles talk only through the flags
// (global memory) and synchronise internally with named barriers.
// ---------------------------------------------------------------------------------------------
constexpr int kPipeIn = 192, kPipeRed = 128, kPipeOut = 192;
static_assert(kPipeIn + kPipeRed + kPipeOut == kThreads, "roles must cover the CTA");

template <typename TI, typename TW>
__global__ void __launch_bounds__(kThreads, 2) k_allreduce_nvls_pipe(CollArgs a) {
  const DevComm& c = a.c;
  const int r = c.rank, W = c.world;
  if (!coll_prologue(a)) return;
  check_signature(a);
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  constexpr int V = 16 / sizeof(TW);
  const size_t t0 = (size_t)blockIdx.x * a.tile;
  const size_t t1 = (t0 + a.tile < a.chunk) ? t0 + a.tile : a.chunk;
  if (t0 >= t1) return;
  const int K = (int)((t1 - t0 + a.sub - 1) / a.sub);
  const TI* in = static_cast<const TI*>(a.in);
  TI* out = static_cast<TI*>(a.out);
  const size_t base_off = c.off_staging + (size_t)(a.seq & 1) * c.staging_bytes;
  TW* mine = reinterpret_cast<TW*>(c.arena[r] + base_off);
  const size_t flag_idx = (size_t)blockIdx.x * 8;
  const uint32_t* myA = reinterpret_cast<const uint32_t*>(c.arena[r] + kOffPipeA) + flag_idx;
  const uint32_t* myB = reinterpret_cast<const uint32_t*>(c.arena[r] + kOffPipeB) + flag_idx;
  const int tid = threadIdx.x;

  if (tid < kPipeIn) {
    // ------------------------------------------------------------------ IN
    const int t = tid;
    for (int k = 0; k < K; k++) {
      const size_t s0 = t0 + (size_t)k * a.sub;
      const size_t s1 = (s0 + a.sub < t1) ? s0 + a.sub : t1;
      for (int j = 0; j < W; j++) {
        size_t lo = (size_t)j * a.chunk + s0, hi = (size_t)j * a.chunk + s1;
        size_t cnt = clip_count(lo, hi, a.n);
        if (cnt) move_tile<TI, TW, false>(mine + lo, in + lo, cnt, t, kPipeIn);
        size_t end = lo + cnt, padded = (end + V - 1) / V * V;
        if (cnt && padded > end && padded <= hi) {
          TW z = Traits<TW>::from_acc((typename Traits<TW>::A)0);
          for (size_t q = end + t; q < padded; q += kPipeIn) mine[q] = z;
        }
      }
      role_sync(1, kPipeIn);
      if (t < W) st_release_sys(reinterpret_cast<uint32_t*>(c.arena[t] + kOffPipeA) + flag_idx + r, a.pipe_base + k + 1);
    }
  } else if (tid < kPipeIn + kPipeRed) {
    // ------------------------------------------------------------------ RED
    const int t = tid - kPipeIn;
    for (int k = 0; k < K; k++) {
      if (t < W && !wait_flag(myA + t, a.pipe_base + k + 1, c, t, 1)) s_fail = 1;
      role_sync(2, kPipeRed);
      if (s_fail) return;
      const size_t s0 = t0 + (size_t)k * a.sub;
      const size_t s1 = (s0 + a.sub < t1) ? s0 + a.sub : t1;
      size_t lo = (size_t)r * a.chunk + s0;
      size_t cnt = clip_count(lo, (size_t)r * a.chunk + s1, a.n);
      size_t nv = (cnt + V - 1) / V;
      char* mc = c.mc_arena + base_off + lo * sizeof(TW);
      size_t i = t;
      for (; i + (size_t)(kUnroll - 1) * kPipeRed < nv; i += (size_t)kUnroll * kPipeRed) {
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) v[u] = Multimem<TW>::ld_reduce(mc + (i + (size_t)u * kPipeRed) * 16);
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
          if (a.has_scale) {
            Pack16<TW> p; p.u = v[u];
#pragma unroll
            for (int e = 0; e < V; e++) p.e[e] = Traits<TW>::from_acc(Traits<TW>::to_acc(p.e[e]) * a.scale);
            v[u] = p.u;
          }
          multimem_st16(mc + (i + (size_t)u * kPipeRed) * 16, v[u]);
        }
      }
      for (; i < nv; i += kPipeRed) {
        uint4 v = Multimem<TW>::ld_reduce(mc + i * 16);
        if (a.has_scale) {
          Pack16<TW> p; p.u = v;
#pragma unroll
          for (int e = 0; e < V; e++) p.e[e] = Traits<TW>::from_acc(Traits<TW>::to_acc(p.e[e]) * a.scale);
          v = p.u;
        }
        multimem_st16(mc + i * 16, v);
      }
      role_sync(2, kPipeRed);
      if (t < W) st_release_sys(reinterpret_cast<uint32_t*>(c.arena[t] + kOffPipeB) + flag_idx + r, a.pipe_base + k + 1);
    }
  } else {
    // ------------------------------------------------------------------ OUT
    const int t = tid - kPipeIn - kPipeRed;
    for (int k = 0; k < K; k++) {
      if (t < W && !wait_flag(myB + t, a.pipe_base + k + 1, c, t, 2)) s_fail = 1;
      role_sync(3, kPipeOut);
      if (s_fail) return;
      const size_t s0 = t0 + (size_t)k * a.sub;
      const size_t s1 = (s0 + a.sub < t1) ? s0 + a.sub : t1;
      for (int j = 0; j < W; j++) {
        size_t lo = (size_t)j * a.chunk + s0;
        size_t cnt = clip_count(lo, (size_t)j * a.chunk + s1, a.n);
        if (cnt) move_tile<TW, TI, true>(out + lo, mine + lo, cnt, t, kPipeOut);
      }
    }
  }
}

// world == 1: no peers, only the wire rounding and the scale remain (the DDP hook at N = 1).
// Grid-stride over 16-byte vectors, four loads in flight per thread; HBM-bound (read n, write n).
template <typename TI, typename TW>
__device__ __forceinline__ TI local_scale_one(TI x, const CollArgs& a) {
  using A = typename Traits<TW>::A;
  A v = Traits<TW>::to_acc(Traits<TW>::from_acc((A)Traits<TI>::to_acc(x)));
  if (a.has_scale) v = apply_scale<A>(v, a.scale, 1);
  return Traits<TI>::from_acc((typename Traits<TI>::A)Traits<TW>::to_acc(Traits<TW>::from_acc(v)));
}

template <typename TI, typename TW>
__global__ void __launch_bounds__(kThreads) k_local_scale(CollArgs a) {
  constexpr int VI = 16 / sizeof(TI);
  const TI* in = static_cast<const TI*>(a.in);
  TI* out = static_cast<TI*>(a.out);
  const size_t n = a.n;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t gtid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (aligned16(in) && aligned16(out)) {
    const size_t nv = n / VI;
    const uint4* s = reinterpret_cast<const uint4*>(in);
    uint4* d = reinterpret_cast<uint4*>(out);
    size_t i = gtid;
    for (; i + (kUnroll - 1) * stride < nv; i += kUnroll * stride) {
      Pack16<TI> p[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; u++) p[u].u = s[i + u * stride];
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
#pragma unroll
        for (int e = 0; e < VI; e++) p[u].e[e] = local_scale_one<TI, TW>(p[u].e[e], a);
        d[i + u * stride] = p[u].u;
      }
    }
    for (; i < nv; i += stride) {
      Pack16<TI> p;
      p.u = s[i];
#pragma unroll
      for (int e = 0; e < VI; e++) p.e[e] = local_scale_one<TI, TW>(p.e[e], a);
      d[i] = p.u;
    }
    for (size_t k = nv * VI + gtid; k < n; k += stride) out[k] = local_scale_one<TI, TW>(in[k], a);
  } else {
    for (size_t k = gtid; k < n; k += stride) out[k] = local_scale_one<TI, TW>(in[k], a);
  }
}

}  // namespace b200c
