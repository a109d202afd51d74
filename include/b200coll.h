/*
 * b200coll.h — C-ABI of the B200 peer-memory collective / tensor-transport library.
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json.
 * The reference has no C seam of its own on this path: its only native boundary is
 * cupy's NcclCommunicator, which takes raw integer device pointers, element counts,
 * NCCL dtype / redop enums and a raw stream pointer.  Every entry point below replaces
 * one of those call sites and keeps that call shape (plain pointers and sizes, no torch
 * or cupy types).  Enum values are the ncclDataType_t / ncclRedOp_t numbering so the
 * reference's dtype/op maps (nccl_util.py:22-87) carry over unchanged.
 *
 * All collective / p2p calls are asynchronous: they enqueue work on `stream` and return.
 * Return value: 0 on success, a negative B200C_E* code otherwise; b200c_last_error()
 * returns a thread-local human-readable message for the last failure.
 *
 * Threading: a communicator is NOT thread-safe (same contract as the reference:
 * nccl_collective_group.py:127 "we need a lock here", nccl_group.py:26 "not thread-safe").
 */
#ifndef B200COLL_H_
#define B200COLL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200C_VERSION 200 /* 0.2.0 */
#define B200C_MAX_RANKS 8 /* one NVSwitch domain */

/* ncclDataType_t numbering (reference: nccl_util.py:30-71 maps numpy/torch dtypes onto these). */
typedef enum {
  B200C_INT8 = 0,
  B200C_UINT8 = 1,
  B200C_INT32 = 2,
  B200C_UINT32 = 3,
  B200C_INT64 = 4,
  B200C_UINT64 = 5,
  B200C_FLOAT16 = 6,
  B200C_FLOAT32 = 7,
  B200C_FLOAT64 = 8,
  B200C_BFLOAT16 = 9,
  B200C_NUM_DTYPES = 10
} b200c_dtype_t;

/* ncclRedOp_t numbering.  ray.experimental.util.types.ReduceOp passes its .value raw
 * (nccl_group.py:304,325); ray.util.collective.types.ReduceOp goes through
 * NCCL_REDUCE_OP_MAP (nccl_util.py:22-27). */
typedef enum {
  B200C_SUM = 0,
  B200C_PROD = 1,
  B200C_MAX = 2,
  B200C_MIN = 3,
  B200C_AVG = 4,
  B200C_NUM_OPS = 5
} b200c_redop_t;

/* Algorithm selector for allreduce. AUTO picks by message size (thresholds in b200c_config_t). */
typedef enum {
  B200C_ALGO_AUTO = 0,
  B200C_ALGO_ONESHOT = 1, /* every rank pushes its whole buffer to every peer, reduces locally */
  B200C_ALGO_TWOSHOT = 2, /* push reduce-scatter + pull all-gather over peer memory */
  B200C_ALGO_NVLS = 3,    /* multimem.ld_reduce / multimem.st on the NVSwitch multicast object */
  B200C_ALGO_NVLS_PIPE = 4,/* same, staged copies overlapped with the switch traffic (per-round flags, software-pipelined blocks) */
  B200C_ALGO_LL = 5,       /* packed {data, flag} 8-byte stores, one NVLink hop, no fence: small messages */
  B200C_ALGO_NVLS_LANES = 6,/* staged NVLS in lanes: few switch-only CTAs + many copy-only CTAs over an L2-resident staging ring */
  B200C_ALGO_NVLS_STREAMS = 7 /* staged NVLS as a pipeline of kernels on internal streams: copy-in | 32-CTA zero-copy NVLS | copy-out per piece */
} b200c_algo_t;

typedef enum {
  B200C_OK = 0,
  B200C_EINVAL = -1,      /* bad argument (maps to ValueError / RuntimeError on the Python side) */
  B200C_ECUDA = -2,       /* CUDA runtime / driver call failed */
  B200C_ESTATE = -3,      /* communicator not ready / already destroyed */
  B200C_EUNSUPPORTED = -4,/* dtype/op/algorithm combination not available on this device */
  B200C_ETIMEOUT = -5,    /* a kernel gave up waiting for a peer flag (dead or mismatched peer) */
  B200C_EABORTED = -6,    /* b200c_comm_abort() was called while kernels were waiting */
  B200C_EMISMATCH = -7,   /* peers disagreed on op / dtype / count for the same sequence number */
  B200C_ENOMEM = -8
} b200c_status_t;

/* How a rank's arena is shared with its peers. */
typedef enum {
  B200C_SHARE_VMM_FD = 0,     /* cuMemCreate + POSIX fd (SCM_RIGHTS side channel); multicast capable */
  B200C_SHARE_LEGACY_IPC = 1  /* cudaMalloc + 64-byte cudaIpcMemHandle_t (plain bytes, object store) */
} b200c_share_mode_t;

typedef struct {
  uint32_t struct_size;        /* sizeof(b200c_config_t), for forward compatibility */
  int32_t share_mode;          /* b200c_share_mode_t */
  uint64_t staging_bytes;      /* bytes of ONE staging half (two halves are allocated) */
  uint64_t symmetric_bytes;    /* user-visible symmetric region (0 = none) */
  uint64_t p2p_slot_bytes;     /* bytes of one p2p ring slot */
  uint32_t p2p_slots;          /* ring slots per ordered (src,dst) pair */
  uint32_t max_blocks;         /* upper bound on CTAs per collective kernel (<= 2048) */
  uint64_t oneshot_max_bytes;  /* AUTO: message <= this -> one-shot */
  uint64_t nvls_min_bytes;     /* AUTO: message >= this and multicast bound -> NVLS */
  uint64_t nvls_pipe_min_bytes;/* AUTO: staged NVLS pieces >= this use the round-pipelined kernel (0 = never) */
  uint64_t timeout_ms;         /* device-side bounded spin; 0 = default (600 s; NCCL's watchdog default is of that order) */
  uint64_t granule_bytes;      /* block-cyclic granule of the large-message kernels (multiple of 16 KiB; 0 = 32 KiB) */
  uint64_t ll_max_bytes;       /* AUTO: same-type allreduce <= this goes by the LL kernel; also sizes the LL region (0 = no LL) */
  uint64_t bcast_rounds_min_bytes; /* broadcast >= this uses scatter + multicast-allgather rounds (0 = never) */
  uint32_t nvls_blocks;        /* CTAs of the zero-copy NVLS kernel (0 = max_blocks) */
  uint32_t nvls_lanes;         /* lane kernel: lanes (each 1 switch CTA + (max_blocks / lanes - 1 <= 7) copy CTAs) */
  uint64_t lane_granule_bytes; /* lane kernel: bytes of one rank chunk's granule per round (multiple of 8 KiB) */
  uint64_t nvls_lanes_min_bytes; /* AUTO: staged NVLS messages >= this use the lane kernel (0 = never) */
  uint32_t nvls_unroll;        /* reserved (ignored): round-2 experiment, 8 instead of 4 multimem vectors in flight - no gain, removed */
  uint32_t rounds_order;       /* reserved (ignored): round-2 experiment, copy-out before the switch stage - slower, removed */
  uint64_t nvls_streams_min_bytes; /* AUTO: staged NVLS messages >= this use the multi-stream pipeline (0 = never) */
  uint64_t nvls_streams_piece_bytes; /* bytes (on the wire) per pipeline piece; at least 3 pieces must fit 2 * staging_bytes */
} b200c_config_t;

typedef struct {
  int32_t device;              /* CUDA ordinal queried */
  int32_t sm_count;
  int32_t cc_major, cc_minor;
  int32_t vmm_supported;       /* CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED */
  int32_t posix_fd_supported;  /* CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED */
  int32_t multicast_supported; /* CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED */
  int32_t reserved;
  uint64_t total_mem;
} b200c_props_t;

/* Opaque bytes a rank publishes so that peers can map its arena.  With SHARE_VMM_FD the
 * `fd` must travel by SCM_RIGHTS (the Python host side does this over a Unix socket);
 * with SHARE_LEGACY_IPC `ipc` is self-contained and can go through Ray's object store
 * (north_star: "CUDA-IPC handles exchanged through the object store"). */
typedef struct {
  int32_t share_mode;
  int32_t fd;                  /* -1 unless SHARE_VMM_FD */
  uint64_t arena_bytes;
  uint64_t layout_hash;        /* peers must agree on the arena layout */
  int32_t pid;
  int32_t device_uuid_lo;      /* low 32 bits of the device UUID, diagnostics only */
  uint8_t ipc[64];             /* cudaIpcMemHandle_t when SHARE_LEGACY_IPC */
} b200c_export_t;

typedef struct b200c_comm b200c_comm_t;
typedef void* b200c_stream_t;  /* cudaStream_t / CUstream, passed as intptr like cupy's stream.ptr */

/* ---- library ---- */
int b200c_version(void);
const char* b200c_last_error(void);
const char* b200c_status_string(int status);
size_t b200c_dtype_size(int dtype);
int b200c_device_props(int device, b200c_props_t* out);
void b200c_default_config(b200c_config_t* cfg);

/* ---- communicator lifecycle ----
 * Replaces NcclCommunicator(world, uid, rank) (nccl_util.py:107-118; nccl_group.py:90) and the
 * Rendezvous around it (nccl_collective_group.py:29-118).  The rendezvous transport itself
 * (named actor / internal KV / torch store / Unix socket) lives on the host side; the library
 * only produces and consumes the bytes.
 *
 * Order: create -> export (publish) -> import x (world-1) -> [multicast: mc_create on rank 0,
 * mc_import elsewhere, mc_add_device everywhere, <barrier>, mc_bind everywhere, <barrier>]
 * -> ready.  Host-side barriers between the steps are the caller's job. */
int b200c_comm_create(int rank, int world, int device, const b200c_config_t* cfg, b200c_comm_t** out);
int b200c_comm_export(b200c_comm_t* comm, b200c_export_t* out);
int b200c_comm_import(b200c_comm_t* comm, int peer, const b200c_export_t* peer_export);
int b200c_comm_mc_create(b200c_comm_t* comm, int* fd_out);   /* rank 0 */
int b200c_comm_mc_import(b200c_comm_t* comm, int fd);        /* ranks != 0 */
int b200c_comm_mc_add_device(b200c_comm_t* comm);
int b200c_comm_mc_bind(b200c_comm_t* comm);
/* Drop the multicast mapping (called on every rank when any rank failed to bind). */
int b200c_comm_mc_disable(b200c_comm_t* comm);
int b200c_comm_ready(b200c_comm_t* comm);
/* Replaces comm.abort() (nccl_group.py:347-365): makes every kernel of this communicator that is
 * spinning on a peer flag give up; safe to call from another thread. */
int b200c_comm_abort(b200c_comm_t* comm);
int b200c_comm_destroy(b200c_comm_t* comm);
/* Non-blocking: returns B200C_OK, or the first error a kernel of this communicator recorded
 * (ETIMEOUT / EABORTED / EMISMATCH).  Reads host-pinned memory; does not synchronise. */
int b200c_comm_check(b200c_comm_t* comm);
int b200c_comm_rank(const b200c_comm_t* comm);
int b200c_comm_world(const b200c_comm_t* comm);
int b200c_comm_has_multicast(const b200c_comm_t* comm);
uint64_t b200c_comm_seq(const b200c_comm_t* comm);

/* Symmetric region: the same offset names the same logical buffer on every rank.  A tensor
 * living there takes the zero-copy paths (NVLS reads/writes it in place). */
void* b200c_comm_symmetric_base(b200c_comm_t* comm);
uint64_t b200c_comm_symmetric_bytes(const b200c_comm_t* comm);

/* Symmetric pool: allocator entry points in the shape torch.cuda.memory.CUDAPluggableAllocator expects
 * (alloc(size, device, stream) / free(ptr, size, device, stream)), carving 2 MiB-granular segments out of
 * the symmetric region of the communicator selected by b200c_pool_bind (one per process; NULL unbinds).
 * Tensors allocated from a torch.cuda.MemPool built on them are peer-mapped and multicast-bound, so
 * collectives on them are zero-copy.  First fit over a sorted free list: the same allocation sequence on
 * every rank yields the same offsets. */
int b200c_pool_bind(b200c_comm_t* comm);
void* b200c_pool_malloc(size_t size, int device, void* stream);
void b200c_pool_free(void* ptr, size_t size, int device, void* stream);

/* ---- collectives (K1-K7, K10-K13 in SURVEY.md §2d) ---- */

/* allReduce(sendptr, recvptr, count, dtype, op, stream): nccl_collective_group.py:181-188,
 * nccl_group.py:293-312.  send == recv (in place) is allowed. */
int b200c_allreduce(b200c_comm_t* comm, const void* send, void* recv, size_t count, int dtype,
                    int op, int algo, b200c_stream_t stream);

/* Fused gradient allreduce for the DDP bucket hook (K13): reads `count` elements of `dtype`
 * from every rank's bucket, moves `wire_dtype` over NVLink (FLOAT32 bucket + BFLOAT16 wire is
 * the bf16-compress case), accumulates in fp32 in rank order, multiplies by `scale`
 * (1/world for the mean) and writes `dtype` back in place.  One launch per piece; replaces
 * div_() + ncclAllReduce (+ to(bf16)/copy_() in bf16_compress_hook). */
int b200c_allreduce_scaled(b200c_comm_t* comm, const void* send, void* recv, size_t count, int dtype,
                           int wire_dtype, float scale, int algo, b200c_stream_t stream);

/* reduce(sendptr, recvptr, count, dtype, op, root, stream): nccl_collective_group.py:226-234.
 * Only `root` writes `recv`; other ranks' buffers are left untouched (gloo semantics,
 * torch_gloo_collective_group.py:170-179). */
int b200c_reduce(b200c_comm_t* comm, const void* send, void* recv, size_t count, int dtype, int op,
                 int root, b200c_stream_t stream);

/* broadcast(ptr, ptr, count, dtype, root, stream): nccl_collective_group.py:253-260. */
int b200c_broadcast(b200c_comm_t* comm, void* buf, size_t count, int dtype, int root,
                    b200c_stream_t stream);

/* allGather: nccl_collective_group.py:278-284 + the W copy_tensor calls of postprocess_fn
 * (:292-296).  `recv_ptrs` holds `world` device pointers (the caller's W output tensors, or
 * base + j*count*size for a contiguous output as in nccl_group.py:274-291); rank j's data lands
 * in recv_ptrs[j].  No flat temp buffer, no extra D2D copies. */
int b200c_allgather(b200c_comm_t* comm, const void* send, void* const* recv_ptrs, size_t count,
                    int dtype, b200c_stream_t stream);

/* reduceScatter: nccl_collective_group.py:319-326 + the W copy_tensor calls of preprocess_fn
 * (:334-337).  `send_ptrs[j]` is this rank's contribution to rank j (`count` elements each). */
int b200c_reducescatter(b200c_comm_t* comm, const void* const* send_ptrs, void* recv, size_t count,
                        int dtype, int op, b200c_stream_t stream);

/* send / recv: nccl_collective_group.py:355-363, 381-389; nccl_group.py:178-184, 217-237.
 * Sender writes the receiver's HBM (ring of slots) and raises a flag; receiver copies out. */
int b200c_send(b200c_comm_t* comm, const void* buf, size_t bytes, int peer, b200c_stream_t stream);
int b200c_recv(b200c_comm_t* comm, void* buf, size_t bytes, int peer, b200c_stream_t stream);

/* Multi-reader send: torch_tensor_accelerator_channel.py:586-590 sends the same tensor once per reader
 * ("TODO: If there are multiple readers, can replace with a broadcast").  One call delivers `bytes` to
 * every rank in `peers`; with a bound multicast object the payload leaves this GPU once (multimem.st)
 * and the NVSwitch replicates it.  Every reader receives with b200c_recv_multi(src = the sender).
 * A source rank multi-sends to ONE reader set for the lifetime of the communicator (ring positions are
 * counted per source); another reader set -> B200C_EUNSUPPORTED, use per-reader b200c_send.  With a
 * single reader, or a world of two, both calls are the pairwise b200c_send / b200c_recv. */
int b200c_send_multi(b200c_comm_t* comm, const void* buf, size_t bytes, const int* peers, int npeers,
                     b200c_stream_t stream);
int b200c_recv_multi(b200c_comm_t* comm, void* buf, size_t bytes, int src, b200c_stream_t stream);

/* barrier: nccl_collective_group.py:192-210 (an allreduce of [1] in the reference). */
int b200c_barrier(b200c_comm_t* comm, b200c_stream_t stream);

/* Profiling aid: set every flag of this rank's signal pad to `value`.  With value = 0x7fffffff every
 * wait of every later op is already satisfied, so ONE rank's kernel can run without its peers —
 * which is what Nsight Compute's kernel replay needs (it re-runs a kernel in isolation, so a kernel
 * that waits for a concurrently running peer kernel would never finish).  Results are garbage;
 * memory traffic and instruction mix are those of the real run.  The LL region is stamped with the
 * flag of the next LL op (one LL launch per call).  Never call it on a live group. */
int b200c_debug_fill_flags(b200c_comm_t* comm, uint32_t value);

/* Launch statistics (bench.py's gpu_launches claim). */
uint64_t b200c_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200COLL_H_ */
