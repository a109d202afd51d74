// Host side of the C-ABI declared in include/b200coll.h: arena creation and sharing (VMM + POSIX fd,
// or legacy CUDA IPC), NVSwitch multicast binding, algorithm selection and kernel launches.
//
// The driver API (cuMem*, cuMulticast*) is reached through cudaGetDriverEntryPoint so the library
// has no link-time dependency on libcuda and can be loaded (symbol check) on a machine without a GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <mutex>

#include "byte_kernels.cuh"
#include "launch_typed.cuh"

using namespace b200c;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define RT(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) return fail(B200C_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); \
  } while (0)

static std::atomic<uint64_t> g_launches{0};

// ------------------------------------------------------------------------------------------------
// driver entry points
// ------------------------------------------------------------------------------------------------
struct Driver {
  bool ok = false;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
};
static Driver g_drv;
static std::once_flag g_drv_once;

template <typename F>
static bool load_sym(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(p);
  return true;
}
static void load_driver() {
  Driver& d = g_drv;
  bool ok = true;
  ok &= load_sym("cuGetErrorString", &d.GetErrorString);
  ok &= load_sym("cuDeviceGet", &d.DeviceGet);
  ok &= load_sym("cuDeviceGetAttribute", &d.DeviceGetAttribute);
  ok &= load_sym("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity);
  ok &= load_sym("cuMemCreate", &d.MemCreate);
  ok &= load_sym("cuMemRelease", &d.MemRelease);
  ok &= load_sym("cuMemAddressReserve", &d.MemAddressReserve);
  ok &= load_sym("cuMemAddressFree", &d.MemAddressFree);
  ok &= load_sym("cuMemMap", &d.MemMap);
  ok &= load_sym("cuMemUnmap", &d.MemUnmap);
  ok &= load_sym("cuMemSetAccess", &d.MemSetAccess);
  ok &= load_sym("cuMemExportToShareableHandle", &d.MemExportToShareableHandle);
  ok &= load_sym("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle);
  // multicast is optional
  load_sym("cuMulticastCreate", &d.MulticastCreate);
  load_sym("cuMulticastAddDevice", &d.MulticastAddDevice);
  load_sym("cuMulticastBindMem", &d.MulticastBindMem);
  load_sym("cuMulticastUnbind", &d.MulticastUnbind);
  load_sym("cuMulticastGetGranularity", &d.MulticastGetGranularity);
  d.ok = ok;
}
static const char* cu_str(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "unknown";
}
#define DRV(call)                                                                                   \
  do {                                                                                              \
    CUresult r_ = (call);                                                                           \
    if (r_ != CUDA_SUCCESS) return fail(B200C_ECUDA, "%s failed: %d (%s)", #call, (int)r_, cu_str(r_)); \
  } while (0)

struct DeviceGuard {
  int prev = -1, dev;
  bool switched = false;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != d) { cudaSetDevice(d); switched = true; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------------
// communicator
// ------------------------------------------------------------------------------------------------
struct b200c_comm {
  int rank = 0, world = 1, device = 0;
  CUdevice cudev = 0;
  b200c_config_t cfg{};
  int sm_count = 148;
  // arena
  size_t arena_bytes = 0, gran = 0;
  size_t off_staging = 0, off_p2p = 0, off_mring = 0, off_ll = 0, ll_words = 0, off_sym = 0, sym_bytes = 0;
  uint32_t mcells = 0;
  uint64_t layout_hash = 0;
  bool vmm = true;
  CUmemGenericAllocationHandle own_handle = 0;
  CUmemGenericAllocationHandle peer_handle[kMaxRanks] = {};
  char* arena[kMaxRanks] = {};
  bool imported[kMaxRanks] = {};
  // multicast
  CUmemGenericAllocationHandle mc_handle = 0;
  bool mc_have_handle = false, mc_added = false, mc_bound = false;
  char* mc_arena = nullptr;
  // status
  Status* status_host = nullptr;
  Status* status_dev = nullptr;
  // state
  bool ready = false, destroyed = false;
  uint32_t seq = 0;
  bool bcast_mc = false;   // broadcast through one multicast store stream (wins for W > 2)
  uint32_t pipe_base = 0;  // flag epoch of the round-pipelined kernels (advanced by the round count of each op)
  uint32_t ll_seq = 0;     // LL op counter (flag value of the packed stores; region half = ll_seq & 1)
  int local_scale_ctas_per_sm = 0, tma_ctas_per_sm = 0;
  // multi-stream NVLS pipeline: internal streams (copy-in, switch, copy-out) and per-region events
  cudaStream_t ps_in = nullptr, ps_nv = nullptr, ps_out = nullptr;
  cudaEvent_t pe_start = nullptr, pe_in[8] = {}, pe_nv[8] = {}, pe_out[8] = {};
  uint32_t send_cells[kMaxRanks] = {};
  uint32_t recv_cells[kMaxRanks] = {};
  uint32_t msend_cells = 0;            // multi-reader ring of this rank: cells sent so far
  uint32_t msend_mask = 0;             // ... and its (fixed) reader set, 0 = not chosen yet
  uint32_t mrecv_cells[kMaxRanks] = {};
  DevComm dev{};
};

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// symmetric pool state (see b200c_pool_bind)
struct PoolBlock { size_t off, len; };
static std::mutex g_pool_mu;
static b200c_comm* g_pool_comm = nullptr;
static PoolBlock g_pool_free[4096];
static int g_pool_nfree = 0;
constexpr size_t kPoolGran = 2ull << 20;


extern "C" int b200c_version(void) { return B200C_VERSION; }
extern "C" const char* b200c_last_error(void) { return g_err; }
extern "C" const char* b200c_status_string(int s) {
  switch (s) {
    case B200C_OK: return "ok";
    case B200C_EINVAL: return "invalid argument";
    case B200C_ECUDA: return "CUDA error";
    case B200C_ESTATE: return "communicator not ready or destroyed";
    case B200C_EUNSUPPORTED: return "unsupported dtype/op/algorithm";
    case B200C_ETIMEOUT: return "timed out waiting for a peer";
    case B200C_EABORTED: return "communicator aborted";
    case B200C_EMISMATCH: return "peers disagree on the collective's arguments";
    case B200C_ENOMEM: return "out of memory";
    default: return "unknown status";
  }
}
extern "C" size_t b200c_dtype_size(int dtype) {
  switch (dtype) {
    case B200C_INT8: case B200C_UINT8: return 1;
    case B200C_FLOAT16: case B200C_BFLOAT16: return 2;
    case B200C_INT32: case B200C_UINT32: case B200C_FLOAT32: return 4;
    case B200C_INT64: case B200C_UINT64: case B200C_FLOAT64: return 8;
    default: return 0;
  }
}
extern "C" uint64_t b200c_launch_count(void) { return g_launches.load(); }

extern "C" void b200c_default_config(b200c_config_t* cfg) {
  memset(cfg, 0, sizeof *cfg);
  cfg->struct_size = sizeof *cfg;
  cfg->share_mode = B200C_SHARE_VMM_FD;
  cfg->staging_bytes = 256ull << 20;
  cfg->symmetric_bytes = 0;
  cfg->p2p_slot_bytes = 32ull << 10;
  cfg->p2p_slots = 1024;   // 32 MiB per ordered pair: at ~500 GB/s the ring must hold the ~60 us of data a ready/ack round trip is behind
  cfg->max_blocks = 296;
  cfg->oneshot_max_bytes = 0;  // 0 = pick by world size in b200c_comm_create
  cfg->nvls_min_bytes = (1ull << 20) + 1;
  cfg->nvls_pipe_min_bytes = 32ull << 20;  // staged NVLS pieces from 32 MiB up run round-pipelined (W=8: 540 vs 505 GB/s at 64 MiB)
  cfg->timeout_ms = 600000;  // a slow peer (data loading, first-step autotuning skew) is not a dead peer
  cfg->granule_bytes = 32ull << 10;
  cfg->ll_max_bytes = 64ull << 10;    // W=8: LL 12 us vs one-shot 13 us at 64 KiB, 15 vs 15 at 128 KiB (profiles/r02_sweep8_small.log)
  cfg->bcast_rounds_min_bytes = 4ull << 20;
  cfg->nvls_unroll = 4;
  cfg->nvls_streams_min_bytes = 768ull << 20;   // W=8, 1 GiB: 737 GB/s vs 704 (rounds kernel) and NCCL 726; 512 MiB: a tie (r02_sweep8_streams.log)
  cfg->nvls_streams_piece_bytes = 128ull << 20;
  cfg->rounds_order = 0;
  cfg->nvls_lanes = 48;
  cfg->lane_granule_bytes = 64ull << 10;
  cfg->nvls_lanes_min_bytes = 0;   // opt-in until measured on the target box
  cfg->nvls_blocks = 32;   // zero-copy NVLS saturates the switch with 32 CTAs; more only scatter the access pattern (r02_sweep8_large.log)
}

static int ensure_driver() {
  std::call_once(g_drv_once, load_driver);
  if (!g_drv.ok) return fail(B200C_ECUDA, "CUDA driver entry points unavailable (no GPU driver?)");
  return B200C_OK;
}

extern "C" int b200c_device_props(int device, b200c_props_t* out) {
  if (!out) return fail(B200C_EINVAL, "out is null");
  memset(out, 0, sizeof *out);
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return fail(B200C_ECUDA, "no CUDA device %d (count %d)", device, n);
  }
  int rc = ensure_driver();
  if (rc) return rc;
  cudaDeviceProp p;
  RT(cudaGetDeviceProperties(&p, device));
  out->device = device;
  out->sm_count = p.multiProcessorCount;
  out->cc_major = p.major;
  out->cc_minor = p.minor;
  out->total_mem = p.totalGlobalMem;
  DeviceGuard g(device);
  RT(cudaFree(0));
  CUdevice cd;
  DRV(g_drv.DeviceGet(&cd, device));
  g_drv.DeviceGetAttribute(&out->vmm_supported, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cd);
  g_drv.DeviceGetAttribute(&out->posix_fd_supported, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cd);
  g_drv.DeviceGetAttribute(&out->multicast_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd);
  return B200C_OK;
}

static CUmemAllocationProp alloc_prop(CUdevice cd) {
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof ap);
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = cd;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return ap;
}
static int map_handle(b200c_comm* c, CUmemGenericAllocationHandle h, char** out) {
  CUdeviceptr va = 0;
  DRV(g_drv.MemAddressReserve(&va, c->arena_bytes, c->gran, 0, 0));
  CUresult r = g_drv.MemMap(va, c->arena_bytes, 0, h, 0);
  if (r != CUDA_SUCCESS) { g_drv.MemAddressFree(va, c->arena_bytes); return fail(B200C_ECUDA, "cuMemMap failed: %d (%s)", (int)r, cu_str(r)); }
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof ad);
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = c->cudev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = g_drv.MemSetAccess(va, c->arena_bytes, &ad, 1);
  if (r != CUDA_SUCCESS) { g_drv.MemUnmap(va, c->arena_bytes); g_drv.MemAddressFree(va, c->arena_bytes); return fail(B200C_ECUDA, "cuMemSetAccess failed: %d (%s)", (int)r, cu_str(r)); }
  *out = reinterpret_cast<char*>(va);
  return B200C_OK;
}

extern "C" int b200c_comm_create(int rank, int world, int device, const b200c_config_t* cfg_in, b200c_comm_t** out) {
  if (!out) return fail(B200C_EINVAL, "out is null");
  if (world < 1 || world > kMaxRanks) return fail(B200C_EINVAL, "world size %d not in [1, %d] (one NVSwitch domain)", world, kMaxRanks);
  if (rank < 0 || rank >= world) return fail(B200C_EINVAL, "rank %d not in [0, %d)", rank, world);
  int rc = ensure_driver();
  if (rc) return rc;
  b200c_config_t cfg;
  b200c_default_config(&cfg);
  if (cfg_in) {
    if (cfg_in->struct_size != sizeof(b200c_config_t)) return fail(B200C_EINVAL, "config struct_size %u != %zu", cfg_in->struct_size, sizeof(b200c_config_t));
    cfg = *cfg_in;
  }
  if (cfg.max_blocks == 0 || cfg.max_blocks > (uint32_t)kMaxBlocks) return fail(B200C_EINVAL, "max_blocks %u not in [1, %d]", cfg.max_blocks, kMaxBlocks);
  if (cfg.p2p_slots == 0 || cfg.p2p_slots > (uint32_t)kMaxCells) return fail(B200C_EINVAL, "p2p_slots %u not in [1, %d]", cfg.p2p_slots, kMaxCells);
  if (cfg.p2p_slot_bytes < 512 || cfg.p2p_slot_bytes % 16) return fail(B200C_EINVAL, "p2p_slot_bytes must be a multiple of 16 and >= 512");
  if (cfg.staging_bytes < (1u << 16) || cfg.staging_bytes % 4096) return fail(B200C_EINVAL, "staging_bytes must be a multiple of 4096 and >= 64 KiB");
  if (cfg.timeout_ms == 0) cfg.timeout_ms = 600000;
  if (cfg.granule_bytes == 0) cfg.granule_bytes = 32ull << 10;
  if (cfg.granule_bytes % 16384) return fail(B200C_EINVAL, "granule_bytes must be a multiple of 16 KiB");
  if (cfg.ll_max_bytes > (1ull << 20)) return fail(B200C_EINVAL, "ll_max_bytes must be <= 1 MiB");
  if (cfg.nvls_blocks > (uint32_t)kMaxBlocks) return fail(B200C_EINVAL, "nvls_blocks %u > %d", cfg.nvls_blocks, kMaxBlocks);
  if (cfg.nvls_lanes == 0) cfg.nvls_lanes = 48;
  if (cfg.nvls_lanes > cfg.max_blocks / 2) cfg.nvls_lanes = cfg.max_blocks / 2 ? cfg.max_blocks / 2 : 1;
  if (cfg.lane_granule_bytes == 0) cfg.lane_granule_bytes = 64ull << 10;
  if (cfg.lane_granule_bytes % 8192) return fail(B200C_EINVAL, "lane_granule_bytes must be a multiple of 8 KiB");
  if (cfg.nvls_streams_piece_bytes == 0) cfg.nvls_streams_piece_bytes = 128ull << 20;
  if (cfg.nvls_streams_piece_bytes % (1u << 20)) return fail(B200C_EINVAL, "nvls_streams_piece_bytes must be a multiple of 1 MiB");
  // measured crossovers (profiles/r01_sweep_*): W=2 one-shot wins to 8 MiB; W=8 one-shot 23 us vs NVLS 28 us at 1 MiB
  if (cfg.oneshot_max_bytes == 0) cfg.oneshot_max_bytes = world <= 2 ? (8ull << 20) : (1ull << 20);

  int ndev = 0;
  RT(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B200C_EINVAL, "device %d not visible (count %d)", device, ndev);
  DeviceGuard g(device);
  RT(cudaFree(0));
  b200c_comm* c = new b200c_comm();
  c->rank = rank; c->world = world; c->device = device; c->cfg = cfg;
  c->vmm = cfg.share_mode == B200C_SHARE_VMM_FD;
  cudaDeviceProp prop;
  RT(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  DRV(g_drv.DeviceGet(&c->cudev, device));

  // layout
  c->off_staging = kPadBytes;
  c->off_p2p = c->off_staging + 2 * cfg.staging_bytes;
  size_t p2p_bytes = world > 1 ? (size_t)world * cfg.p2p_slots * cfg.p2p_slot_bytes : 0;   // one ring per source rank
  c->mcells = world > 2 ? (cfg.p2p_slots < 64 ? cfg.p2p_slots : 64) : 0;  // with one possible reader the pairwise ring is the multi-reader ring
  c->off_mring = round_up(c->off_p2p + p2p_bytes, 4096);
  size_t mring_bytes = (size_t)world * c->mcells * cfg.p2p_slot_bytes;
  c->off_ll = round_up(c->off_mring + mring_bytes, 4096);
  c->ll_words = world > 1 ? round_up((cfg.ll_max_bytes + 3) / 4, 4) : 0;   // whole 16-byte vectors
  size_t ll_bytes = 2 * (size_t)kMaxRanks * c->ll_words * 8;
  c->off_sym = round_up(c->off_ll + ll_bytes, 2ull << 20);
  size_t want = c->off_sym + cfg.symmetric_bytes;
  size_t gran = 2ull << 20;
  if (c->vmm) {
    CUmemAllocationProp ap = alloc_prop(c->cudev);
    size_t g1 = 0;
    DRV(g_drv.MemGetAllocationGranularity(&g1, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (g1 > gran) gran = g1;
    int mc = 0;
    g_drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, c->cudev);
    if (mc && g_drv.MulticastGetGranularity && world > 1) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof mp);
      mp.numDevices = world; mp.size = round_up(want, gran); mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t g2 = 0;
      if (g_drv.MulticastGetGranularity(&g2, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) == CUDA_SUCCESS && g2 > gran) gran = g2;
    }
  }
  c->gran = gran;
  c->arena_bytes = round_up(want, gran);
  c->sym_bytes = c->arena_bytes - c->off_sym;
  c->layout_hash = (uint64_t)cfg.staging_bytes * 1000003ull ^ (uint64_t)cfg.p2p_slot_bytes * 10007ull ^ (uint64_t)cfg.p2p_slots * 101ull ^
                   (uint64_t)c->arena_bytes ^ ((uint64_t)cfg.max_blocks << 48) ^ ((uint64_t)world << 56) ^
                   (uint64_t)cfg.ll_max_bytes * 7919ull ^ (uint64_t)cfg.granule_bytes * 31ull ^ ((uint64_t)cfg.nvls_blocks << 36);

  if (c->vmm) {
    CUmemAllocationProp ap = alloc_prop(c->cudev);
    CUresult r = g_drv.MemCreate(&c->own_handle, c->arena_bytes, &ap, 0);
    if (r != CUDA_SUCCESS) { size_t ab = c->arena_bytes; delete c; return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? B200C_ENOMEM : B200C_ECUDA, "cuMemCreate(%zu bytes) failed: %d (%s)", ab, (int)r, cu_str(r)); }
    rc = map_handle(c, c->own_handle, &c->arena[rank]);
    if (rc) { g_drv.MemRelease(c->own_handle); delete c; return rc; }
  } else {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, c->arena_bytes);
    if (e != cudaSuccess) { size_t ab = c->arena_bytes; delete c; return fail(e == cudaErrorMemoryAllocation ? B200C_ENOMEM : B200C_ECUDA, "cudaMalloc(%zu) failed: %s", ab, cudaGetErrorString(e)); }
    c->arena[rank] = static_cast<char*>(p);
  }
  c->imported[rank] = true;
  RT(cudaMemset(c->arena[rank], 0, kPadBytes));
  if (ll_bytes) RT(cudaMemset(c->arena[rank] + c->off_ll, 0, ll_bytes));  // LL flags start at 0 (never a valid ll_seq)
  RT(cudaHostAlloc(reinterpret_cast<void**>(&c->status_host), sizeof(Status), cudaHostAllocMapped | cudaHostAllocPortable));
  memset((void*)c->status_host, 0, sizeof(Status));
  RT(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->status_dev), (void*)c->status_host, 0));
  RT(cudaDeviceSynchronize());
  *out = c;
  return B200C_OK;
}

extern "C" int b200c_comm_export(b200c_comm_t* c, b200c_export_t* out) {
  if (!c || !out) return fail(B200C_EINVAL, "null argument");
  if (c->destroyed) return fail(B200C_ESTATE, "communicator destroyed");
  memset(out, 0, sizeof *out);
  out->share_mode = c->vmm ? B200C_SHARE_VMM_FD : B200C_SHARE_LEGACY_IPC;
  out->fd = -1;
  out->arena_bytes = c->arena_bytes;
  out->layout_hash = c->layout_hash;
  out->pid = (int32_t)getpid();
  DeviceGuard g(c->device);
  if (c->vmm) {
    int fd = -1;
    DRV(g_drv.MemExportToShareableHandle(&fd, c->own_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    out->fd = fd;
  } else {
    cudaIpcMemHandle_t h;
    RT(cudaIpcGetMemHandle(&h, c->arena[c->rank]));
    static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out->ipc, &h, 64);
  }
  return B200C_OK;
}

extern "C" int b200c_comm_import(b200c_comm_t* c, int peer, const b200c_export_t* e) {
  if (!c || !e) return fail(B200C_EINVAL, "null argument");
  if (c->destroyed) return fail(B200C_ESTATE, "communicator destroyed");
  if (peer < 0 || peer >= c->world || peer == c->rank) return fail(B200C_EINVAL, "bad peer %d", peer);
  if (c->imported[peer]) return fail(B200C_ESTATE, "peer %d already imported", peer);
  if (e->arena_bytes != c->arena_bytes || e->layout_hash != c->layout_hash)
    return fail(B200C_EMISMATCH, "peer %d arena layout differs (bytes %llu vs %zu): all ranks must use the same config", peer,
                (unsigned long long)e->arena_bytes, c->arena_bytes);
  if ((e->share_mode == B200C_SHARE_VMM_FD) != c->vmm) return fail(B200C_EMISMATCH, "peer %d uses a different share mode", peer);
  DeviceGuard g(c->device);
  if (c->vmm) {
    if (e->fd < 0) return fail(B200C_EINVAL, "peer %d export carries no fd", peer);
    DRV(g_drv.MemImportFromShareableHandle(&c->peer_handle[peer], (void*)(uintptr_t)e->fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    int rc = map_handle(c, c->peer_handle[peer], &c->arena[peer]);
    if (rc) return rc;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, e->ipc, 64);
    void* p = nullptr;
    RT(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->arena[peer] = static_cast<char*>(p);
  }
  c->imported[peer] = true;
  return B200C_OK;
}

extern "C" int b200c_comm_mc_create(b200c_comm_t* c, int* fd_out) {
  if (!c || !fd_out) return fail(B200C_EINVAL, "null argument");
  if (!c->vmm) return fail(B200C_EUNSUPPORTED, "multicast needs the VMM share mode");
  if (!g_drv.MulticastCreate) return fail(B200C_EUNSUPPORTED, "driver has no cuMulticastCreate");
  int mc = 0;
  g_drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, c->cudev);
  if (!mc) return fail(B200C_EUNSUPPORTED, "device does not support multicast");
  DeviceGuard g(c->device);
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof mp);
  mp.numDevices = c->world; mp.size = c->arena_bytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  DRV(g_drv.MulticastCreate(&c->mc_handle, &mp));
  c->mc_have_handle = true;
  int fd = -1;
  DRV(g_drv.MemExportToShareableHandle(&fd, c->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *fd_out = fd;
  return B200C_OK;
}
extern "C" int b200c_comm_mc_import(b200c_comm_t* c, int fd) {
  if (!c || fd < 0) return fail(B200C_EINVAL, "bad argument");
  if (!c->vmm) return fail(B200C_EUNSUPPORTED, "multicast needs the VMM share mode");
  DeviceGuard g(c->device);
  DRV(g_drv.MemImportFromShareableHandle(&c->mc_handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  c->mc_have_handle = true;
  return B200C_OK;
}
extern "C" int b200c_comm_mc_add_device(b200c_comm_t* c) {
  if (!c || !c->mc_have_handle) return fail(B200C_ESTATE, "no multicast handle");
  DeviceGuard g(c->device);
  DRV(g_drv.MulticastAddDevice(c->mc_handle, c->cudev));
  c->mc_added = true;
  return B200C_OK;
}
extern "C" int b200c_comm_mc_bind(b200c_comm_t* c) {
  if (!c || !c->mc_added) return fail(B200C_ESTATE, "device not added to the multicast object");
  DeviceGuard g(c->device);
  DRV(g_drv.MulticastBindMem(c->mc_handle, 0, c->own_handle, 0, c->arena_bytes, 0));
  c->mc_bound = true;
  int rc = map_handle(c, c->mc_handle, &c->mc_arena);
  if (rc) { c->mc_arena = nullptr; return rc; }
  return B200C_OK;
}

extern "C" int b200c_comm_mc_disable(b200c_comm_t* c) {
  if (!c) return fail(B200C_EINVAL, "null communicator");
  DeviceGuard g(c->device);
  if (c->mc_arena) { g_drv.MemUnmap((CUdeviceptr)c->mc_arena, c->arena_bytes); g_drv.MemAddressFree((CUdeviceptr)c->mc_arena, c->arena_bytes); c->mc_arena = nullptr; }
  if (c->ready) c->dev.mc_arena = nullptr;
  return B200C_OK;
}

extern "C" int b200c_comm_ready(b200c_comm_t* c) {
  if (!c) return fail(B200C_EINVAL, "null communicator");
  if (c->destroyed) return fail(B200C_ESTATE, "communicator destroyed");
  for (int j = 0; j < c->world; j++)
    if (!c->imported[j]) return fail(B200C_ESTATE, "peer %d not imported yet", j);
  DevComm& d = c->dev;
  memset(&d, 0, sizeof d);
  d.rank = c->rank; d.world = c->world;
  for (int j = 0; j < c->world; j++) d.arena[j] = c->arena[j];
  d.mc_arena = c->mc_arena;
  d.status = c->status_dev;
  d.timeout_ns = (unsigned long long)c->cfg.timeout_ms * 1000000ull;
  d.staging_bytes = c->cfg.staging_bytes;
  d.off_staging = c->off_staging;
  d.off_p2p = c->off_p2p;
  d.p2p_cell_bytes = c->cfg.p2p_slot_bytes;
  d.p2p_cells = (int)c->cfg.p2p_slots;
  d.off_mring = c->off_mring;
  d.mcells = (int)c->mcells;
  d.off_ll = c->off_ll;
  d.ll_words = c->ll_words;
  // a single multimem.st stream leaves the root at ~340 GB/s (measured), a unicast push at ~690 GB/s:
  // multicast pays off as soon as there is more than one receiver
  c->bcast_mc = c->world > 2;
  if (const char* e = getenv("B200COLL_BCAST_MULTICAST")) c->bcast_mc = e[0] == '1';
  c->ready = true;
  return B200C_OK;
}

extern "C" int b200c_comm_abort(b200c_comm_t* c) {
  if (!c) return fail(B200C_EINVAL, "null communicator");
  if (c->status_host) c->status_host->abort_flag = 1;
  return B200C_OK;
}

extern "C" int b200c_comm_check(b200c_comm_t* c) {
  if (!c) return fail(B200C_EINVAL, "null communicator");
  if (!c->status_host) return B200C_OK;
  int e = c->status_host->error;
  if (e == 0) return B200C_OK;
  static const char* phases[] = {"arrive", "flagA", "flagB", "p2p-ready", "p2p-ack"};
  int ph = c->status_host->err_phase;
  if (e == B200C_EMISMATCH)
    return fail(e, "%s: rank %d, op seq %u, peer %d announced signature %08x, this rank expected %08x", b200c_status_string(e), c->rank,
                c->status_host->err_seq, c->status_host->err_peer, c->status_host->err_a, c->status_host->err_b);
  return fail(e, "%s: rank %d, op seq %u, waiting on peer %d (%s)", b200c_status_string(e), c->rank, c->status_host->err_seq,
              c->status_host->err_peer, ph >= 0 && ph < 5 ? phases[ph] : "?");
}

extern "C" int b200c_comm_destroy(b200c_comm_t* c) {
  if (!c) return B200C_OK;
  if (c->destroyed) return B200C_OK;
  c->destroyed = true;
  c->ready = false;
  { std::lock_guard<std::mutex> lk(g_pool_mu); if (g_pool_comm == c) { g_pool_comm = nullptr; g_pool_nfree = 0; } }
  if (c->status_host) c->status_host->abort_flag = 1;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  cudaGetLastError();
  if (c->vmm) {
    if (c->mc_arena) { g_drv.MemUnmap((CUdeviceptr)c->mc_arena, c->arena_bytes); g_drv.MemAddressFree((CUdeviceptr)c->mc_arena, c->arena_bytes); }
    if (c->mc_bound && g_drv.MulticastUnbind) g_drv.MulticastUnbind(c->mc_handle, c->cudev, 0, c->arena_bytes);
    if (c->mc_have_handle) g_drv.MemRelease(c->mc_handle);
    for (int j = 0; j < c->world; j++) {
      if (!c->arena[j]) continue;
      g_drv.MemUnmap((CUdeviceptr)c->arena[j], c->arena_bytes);
      g_drv.MemAddressFree((CUdeviceptr)c->arena[j], c->arena_bytes);
      if (j != c->rank && c->peer_handle[j]) g_drv.MemRelease(c->peer_handle[j]);
    }
    if (c->own_handle) g_drv.MemRelease(c->own_handle);
  } else {
    for (int j = 0; j < c->world; j++) {
      if (!c->arena[j]) continue;
      if (j == c->rank) cudaFree(c->arena[j]);
      else cudaIpcCloseMemHandle(c->arena[j]);
    }
  }
  if (c->ps_in) { cudaStreamDestroy(c->ps_in); cudaStreamDestroy(c->ps_nv); cudaStreamDestroy(c->ps_out); cudaEventDestroy(c->pe_start);
    for (int i = 0; i < 8; i++) { if (c->pe_in[i]) cudaEventDestroy(c->pe_in[i]); if (c->pe_nv[i]) cudaEventDestroy(c->pe_nv[i]); if (c->pe_out[i]) cudaEventDestroy(c->pe_out[i]); } }
  if (c->status_host) cudaFreeHost((void*)c->status_host);
  c->status_host = nullptr;
  cudaGetLastError();
  delete c;
  return B200C_OK;
}

extern "C" int b200c_debug_fill_flags(b200c_comm_t* c, uint32_t value) {
  if (!c || c->destroyed) return fail(B200C_ESTATE, "communicator destroyed");
  DeviceGuard g(c->device);
  RT(cudaDeviceSynchronize());
  static_assert(kPadUsed % 4 == 0, "pad is u32 words");
  uint32_t* host = new uint32_t[kPadUsed / 4];
  for (size_t i = 0; i < kPadUsed / 4; i++) host[i] = value;
  cudaError_t e = cudaMemcpy(c->arena[c->rank], host, kPadUsed, cudaMemcpyHostToDevice);
  delete[] host;
  if (e != cudaSuccess) return fail(B200C_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
  // LL slots are matched by equality, not by >=: pre-stamp both halves with the flag of the NEXT LL op
  // (payload 0) so that exactly one LL launch can run without its peers
  size_t ll_u64 = 2 * (size_t)kMaxRanks * c->ll_words;
  if (ll_u64) {
    unsigned long long* h = new unsigned long long[ll_u64];
    for (size_t i = 0; i < ll_u64; i++) h[i] = (unsigned long long)(c->ll_seq + 1) << 32;
    e = cudaMemcpy(c->arena[c->rank] + c->off_ll, h, ll_u64 * 8, cudaMemcpyHostToDevice);
    delete[] h;
    if (e != cudaSuccess) return fail(B200C_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  return B200C_OK;
}

extern "C" int b200c_comm_rank(const b200c_comm_t* c) { return c ? c->rank : -1; }
extern "C" int b200c_comm_world(const b200c_comm_t* c) { return c ? c->world : -1; }
extern "C" int b200c_comm_has_multicast(const b200c_comm_t* c) { return c && c->mc_arena ? 1 : 0; }
extern "C" uint64_t b200c_comm_seq(const b200c_comm_t* c) { return c ? c->seq : 0; }
extern "C" void* b200c_comm_symmetric_base(b200c_comm_t* c) { return c && c->sym_bytes ? c->arena[c->rank] + c->off_sym : nullptr; }
extern "C" uint64_t b200c_comm_symmetric_bytes(const b200c_comm_t* c) { return c ? c->sym_bytes : 0; }

// ------------------------------------------------------------------------------------------------
// Symmetric pool: a torch.cuda.MemPool (CUDAPluggableAllocator) whose segments are carved out of one
// communicator's symmetric region, so ordinary torch tensors (and DDP's gradient buckets) allocated
// under `torch.cuda.use_mem_pool(pool)` are peer-mapped and multicast-bound: collectives on them take
// the zero-copy paths.  The allocator is deterministic (first fit over a sorted free list, 2 MiB
// granules): ranks that perform the same allocation sequence get the same offsets, which is what the
// symmetric paths require — and what b200c_allreduce verifies through the op signature (the offset is
// part of it), so a divergence is reported as EMISMATCH instead of reducing unrelated memory.
// ------------------------------------------------------------------------------------------------

extern "C" int b200c_pool_bind(b200c_comm_t* c) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!c) { g_pool_comm = nullptr; g_pool_nfree = 0; return B200C_OK; }
  if (c->destroyed || !c->sym_bytes) return fail(B200C_ESTATE, "communicator has no symmetric region (config.symmetric_bytes)");
  g_pool_comm = c;
  g_pool_free[0] = PoolBlock{0, c->sym_bytes / kPoolGran * kPoolGran};
  g_pool_nfree = 1;
  return B200C_OK;
}
extern "C" void* b200c_pool_malloc(size_t size, int device, void* stream) {
  (void)stream;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  b200c_comm* c = g_pool_comm;
  if (!c || c->destroyed || device != c->device || size == 0) return nullptr;
  size_t need = round_up(size, kPoolGran);
  for (int i = 0; i < g_pool_nfree; i++) {
    if (g_pool_free[i].len < need) continue;
    size_t off = g_pool_free[i].off;
    g_pool_free[i].off += need;
    g_pool_free[i].len -= need;
    if (g_pool_free[i].len == 0) { for (int j = i; j + 1 < g_pool_nfree; j++) g_pool_free[j] = g_pool_free[j + 1]; g_pool_nfree--; }
    return c->arena[c->rank] + c->off_sym + off;
  }
  return nullptr;  // torch reports the out-of-memory condition
}
extern "C" void b200c_pool_free(void* ptr, size_t size, int device, void* stream) {
  (void)device; (void)stream;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  b200c_comm* c = g_pool_comm;
  if (!c || !ptr) return;
  char* base = c->arena[c->rank] + c->off_sym;
  if ((char*)ptr < base || (char*)ptr >= base + c->sym_bytes) return;
  size_t off = (size_t)((char*)ptr - base), len = round_up(size, kPoolGran);
  int i = 0;
  while (i < g_pool_nfree && g_pool_free[i].off < off) i++;
  if (g_pool_nfree >= 4095) return;  // cannot track it: leak the block rather than corrupt the list
  for (int j = g_pool_nfree; j > i; j--) g_pool_free[j] = g_pool_free[j - 1];
  g_pool_free[i] = PoolBlock{off, len};
  g_pool_nfree++;
  if (i + 1 < g_pool_nfree && g_pool_free[i].off + g_pool_free[i].len == g_pool_free[i + 1].off) {
    g_pool_free[i].len += g_pool_free[i + 1].len;
    for (int j = i + 1; j + 1 < g_pool_nfree; j++) g_pool_free[j] = g_pool_free[j + 1];
    g_pool_nfree--;
  }
  if (i > 0 && g_pool_free[i - 1].off + g_pool_free[i - 1].len == g_pool_free[i].off) {
    g_pool_free[i - 1].len += g_pool_free[i].len;
    for (int j = i; j + 1 < g_pool_nfree; j++) g_pool_free[j] = g_pool_free[j + 1];
    g_pool_nfree--;
  }
}

// ------------------------------------------------------------------------------------------------
// launch planning
// ------------------------------------------------------------------------------------------------
static int check_ready(b200c_comm* c) {
  if (!c) return fail(B200C_EINVAL, "null communicator");
  if (c->destroyed || !c->ready) return fail(B200C_ESTATE, "communicator not ready or destroyed");
  if (c->status_host->error) return b200c_comm_check(c);
  return B200C_OK;
}
// A launch that fails after its peers may already have launched the same op leaves this rank behind:
// poison the communicator so the failure is loud on every later call instead of a silent divergence.
static int launch_check(b200c_comm* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (c && c->status_host && c->status_host->error == 0) { c->status_host->error = B200C_ECUDA; c->status_host->err_seq = c->seq + 1; }
    return fail(B200C_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  }
  g_launches.fetch_add(1);
  return B200C_OK;
}
// Block-cyclic plan for an extent of `units` elements (vector width `vec`).
//   small extents: one contiguous granule per block of at least min_tile_bytes (more blocks = lower latency);
//   large extents: fixed granule of `granule_bytes`, `max_blocks` blocks, each looping with stride grid * granule.
static void plan_tiles(size_t units, size_t elem_size, size_t vec, uint32_t max_blocks, size_t min_tile_bytes, size_t granule_bytes,
                       size_t* tile, int* grid) {
  if (units == 0) { *tile = vec; *grid = 1; return; }
  size_t bytes = units * elem_size;
  size_t nb = (bytes + min_tile_bytes - 1) / min_tile_bytes;
  if (nb < 1) nb = 1;
  if (nb > max_blocks) nb = max_blocks;
  size_t t = round_up((units + nb - 1) / nb, vec);
  if (t * elem_size > granule_bytes) t = granule_bytes / elem_size;  // granule_bytes is a multiple of 16 KiB, hence of vec
  size_t g = (units + t - 1) / t;
  *tile = t;
  *grid = (int)(g < max_blocks ? g : max_blocks);
}
// Plan for the round-pipelined kernels: at least ~4 rounds per block when the extent allows it, so that
// the software pipeline has something to overlap; granules of 8 KiB .. granule_bytes.
static void plan_rounds(size_t units, size_t elem_size, size_t vec, uint32_t max_blocks, size_t granule_bytes, size_t* tile, int* grid,
                        uint32_t* rounds) {
  size_t bytes = units * elem_size;
  size_t tb = bytes / ((size_t)max_blocks * 4) / 8192 * 8192;
  if (tb < 8192) tb = 8192;
  if (tb > granule_bytes) tb = granule_bytes;
  size_t t = tb / elem_size;
  (void)vec;
  size_t g = (units + t - 1) / t;
  if (g < 1) g = 1;
  *tile = t;
  *grid = (int)(g < max_blocks ? g : max_blocks);
  *rounds = (uint32_t)((units + (size_t)*grid * t - 1) / ((size_t)*grid * t));
}
static uint32_t make_sig(int opcode, int dtype, int op, size_t n, int root, int extra) {
  uint64_t h = 1469598103934665603ull;
  uint64_t v[6] = {(uint64_t)opcode, (uint64_t)dtype, (uint64_t)op, (uint64_t)n, (uint64_t)(root + 1), (uint64_t)extra};
  for (int i = 0; i < 6; i++) { h ^= v[i]; h *= 1099511628211ull; }
  return ((uint32_t)(h ^ (h >> 32)) & 0x7fffffffu) | 1u;  // never 0 and never the 0xFFFFFFFF wildcard
}
// The op takes sequence number seq + 1; the counter itself only advances once the launch has succeeded
// (commit_args), so a rejected call (EINVAL / EUNSUPPORTED before any launch) leaves the ranks aligned.
static void base_args(b200c_comm* c, CollArgs* a) {
  memset(a, 0, sizeof *a);
  a->c = c->dev;
  a->seq = c->seq + 1;
  a->root = -1;
}
static void commit_args(b200c_comm* c, const CollArgs& a) { c->seq = a.seq; }
static int launch_same_type(int dtype, int kind, int op, const CollArgs& a, int grid, cudaStream_t s) {
  switch (dtype) {
    case B200C_INT8: return launch_i8(kind, op, a, grid, s);
    case B200C_UINT8: return launch_u8(kind, op, a, grid, s);
    case B200C_INT32: return launch_i32(kind, op, a, grid, s);
    case B200C_UINT32: return launch_u32(kind, op, a, grid, s);
    case B200C_INT64: return launch_i64(kind, op, a, grid, s);
    case B200C_UINT64: return launch_u64(kind, op, a, grid, s);
    case B200C_FLOAT16: return launch_f16(kind, op, a, grid, s);
    case B200C_FLOAT32: return launch_f32(kind, op, a, grid, s);
    case B200C_FLOAT64: return launch_f64(kind, op, a, grid, s);
    case B200C_BFLOAT16: return launch_bf16(kind, op, a, grid, s);
    default: return fail(B200C_EUNSUPPORTED, "dtype %d", dtype);
  }
}

enum { OPC_ALLREDUCE = 1, OPC_REDUCE, OPC_BROADCAST, OPC_ALLGATHER, OPC_REDUCESCATTER, OPC_BARRIER };
constexpr size_t kMinTileBytes = 8192;

// mixed-type (bucket dtype != wire dtype) and NVLS launches live in this TU
template <typename TI, typename TW, int WT>
static void launch_mixed_w(int algo, const CollArgs& a, int grid, cudaStream_t s) {
  if (algo == B200C_ALGO_ONESHOT) k_allreduce_oneshot<TI, TW, B200C_SUM, WT><<<grid, kThreads, 0, s>>>(a);
  else k_allreduce_twoshot<TI, TW, B200C_SUM, WT><<<grid, kThreads, 0, s>>>(a);
}
template <typename TI, typename TW>
static void launch_mixed(int algo, const CollArgs& a, int grid, cudaStream_t s) {
  switch (a.c.world) {
    case 2: launch_mixed_w<TI, TW, 2>(algo, a, grid, s); break;
    case 4: launch_mixed_w<TI, TW, 4>(algo, a, grid, s); break;
    case 8: launch_mixed_w<TI, TW, 8>(algo, a, grid, s); break;
    default: launch_mixed_w<TI, TW, 0>(algo, a, grid, s); break;
  }
}
template <typename TI, typename TW>
static void launch_nvls(const CollArgs& a, int grid, cudaStream_t s, bool pipe, bool lanes = false) {
  if (lanes) k_allreduce_nvls_lanes<TI, TW><<<grid, kThreads, 0, s>>>(a);
  else if (pipe) k_allreduce_nvls_rounds<TI, TW><<<grid, kThreads, 0, s>>>(a);
  else k_allreduce_nvls<TI, TW><<<grid, kThreads, 0, s>>>(a);
}
template <typename TI, typename TW>
static int local_scale_grid(b200c_comm* c, size_t bytes) {
  // no peers to wait for, so the grid is sized for HBM: one CTA per 32 KiB, but never more CTAs than are
  // resident at once (a second, partial wave would idle part of the chip for a whole pass)
  if (c->local_scale_ctas_per_sm == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_local_scale<float, bf16_t>, kThreads, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 2; }
    c->local_scale_ctas_per_sm = nb;
  }
  size_t want = (bytes + 32767) / 32768, cap = (size_t)c->sm_count * c->local_scale_ctas_per_sm;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// ------------------------------------------------------------------------------------------------
// Multi-stream staged NVLS (the largest plain tensors).  Instead of one kernel that interleaves local copies
// with switch traffic, the message is cut into pieces and every piece runs three ordinary kernels on three
// internal streams, chained by events:
//     copy-in(i)  [user -> staging region i % R, any grid]      on ps_in
//     switch(i)   [the 32-CTA zero-copy NVLS kernel on region]  on ps_nv   (the only kernel that waits for peers)
//     copy-out(i) [region -> user]                              on ps_out
// so copy-in(i+1) and copy-out(i-1) overlap switch(i), the switch kernel keeps its few CTAs streaming, and the
// copies are plain full-speed local kernels that never spin.  A barrier op opens the pipeline (every peer has
// finished what it did before, so both staging halves are free to serve as R regions) and another one closes it
// on the caller's stream after the last copy-out (no later op of a peer can touch this rank's staging while a
// copy-out still reads it).  Region reuse: copy-in(i+R) waits for copy-out(i); every peer finished reducing
// region i before switch(i) completed (flag B).
// ------------------------------------------------------------------------------------------------
template <typename TS, typename TD, bool BYPASS>
static void launch_stage_copy(const void* src, void* dst, size_t n, int sm_count, cudaStream_t s) {
  size_t tile = kStageTileBytes / (sizeof(TS) > sizeof(TD) ? sizeof(TS) : sizeof(TD));
  size_t want = (n + tile - 1) / tile, cap = (size_t)sm_count * 4;
  int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  k_stage_copy<TS, TD, BYPASS><<<grid, kThreads, 0, s>>>(static_cast<const TS*>(src), static_cast<TD*>(dst), n);
}
static int pipeline_setup(b200c_comm* c) {
  if (c->ps_in) return B200C_OK;
  int lo = 0, hi = 0;
  RT(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  RT(cudaStreamCreateWithPriority(&c->ps_in, cudaStreamNonBlocking, lo));
  RT(cudaStreamCreateWithPriority(&c->ps_nv, cudaStreamNonBlocking, hi));   // the switch kernel's CTAs go first
  RT(cudaStreamCreateWithPriority(&c->ps_out, cudaStreamNonBlocking, lo));
  RT(cudaEventCreateWithFlags(&c->pe_start, cudaEventDisableTiming));
  for (int i = 0; i < 8; i++) {
    RT(cudaEventCreateWithFlags(&c->pe_in[i], cudaEventDisableTiming));
    RT(cudaEventCreateWithFlags(&c->pe_nv[i], cudaEventDisableTiming));
    RT(cudaEventCreateWithFlags(&c->pe_out[i], cudaEventDisableTiming));
  }
  return B200C_OK;
}
static int barrier_op(b200c_comm* c, cudaStream_t s) {
  CollArgs a;
  base_args(c, &a);
  a.sig = make_sig(OPC_BARRIER, 0, 0, 0, -1, 0);
  k_barrier<<<1, 32, 0, s>>>(a);
  int rc = launch_check(c, "barrier");
  if (rc) return rc;
  commit_args(c, a);
  return B200C_OK;
}
static int allreduce_streams(b200c_comm* c, const void* send, void* recv, size_t count, int dtype, int wire, int op, float scale,
                             int has_scale, cudaStream_t user) {
  (void)op;
  const int W = c->world;
  const size_t esz = b200c_dtype_size(dtype), wsz = b200c_dtype_size(wire), vec = 16 / wsz;
  int rc = pipeline_setup(c);
  if (rc) return rc;
  size_t piece_bytes = c->cfg.nvls_streams_piece_bytes;
  size_t area = 2 * c->cfg.staging_bytes;
  while (area / piece_bytes < 3 && piece_bytes > (1u << 20)) piece_bytes /= 2;
  int R = (int)(area / piece_bytes);
  if (R < 3) return fail(B200C_EINVAL, "staging_bytes too small for the multi-stream pipeline");
  if (R > 8) R = 8;
  const size_t piece_elems = piece_bytes / wsz / (vec * W) * (vec * W);
  // Piece sizes: the pipeline's fill (first copy-in) and drain (last copy-out) are not overlapped with anything, so
  // the first and the last piece are a quarter of the regular size.
  const size_t small = piece_elems / 4 / (vec * W) * (vec * W);
  const uint32_t cap = c->cfg.nvls_blocks && c->cfg.nvls_blocks < c->cfg.max_blocks ? c->cfg.nvls_blocks : c->cfg.max_blocks;
  // open: every peer has finished its previous op, so the whole staging area is ours to partition
  rc = barrier_op(c, user);
  if (rc) return rc;
  RT(cudaEventRecord(c->pe_start, user));
  RT(cudaStreamWaitEvent(c->ps_in, c->pe_start, 0));
  size_t e0 = 0;
  int last_reg = 0;
  for (int i = 0; e0 < count; i++) {
    const int reg = i % R;
    last_reg = reg;
    const size_t left = count - e0;
    size_t n;
    if (small && count > 2 * piece_elems) {
      if (i == 0) n = small;
      else if (left <= small) n = left;
      else if (left <= piece_elems + small) n = left - small;   // the piece before the short last one takes the remainder
      else n = piece_elems;
    } else {
      n = left < piece_elems ? left : piece_elems;
    }
    char* region = c->arena[c->rank] + c->off_staging + (size_t)reg * piece_bytes;
    const char* src = static_cast<const char*>(send) + e0 * esz;
    char* dst = static_cast<char*>(recv) + e0 * esz;
    if (i >= R) RT(cudaStreamWaitEvent(c->ps_in, c->pe_out[reg], 0));   // the region's previous piece has been copied out
    if (wire == dtype) {
      if (esz == 4) launch_stage_copy<float, float, false>(src, region, n, c->sm_count, c->ps_in);
      else launch_stage_copy<bf16_t, bf16_t, false>(src, region, n, c->sm_count, c->ps_in);   // 2-byte payload: a plain copy either way
    } else if (wire == B200C_BFLOAT16) launch_stage_copy<float, bf16_t, false>(src, region, n, c->sm_count, c->ps_in);
    else launch_stage_copy<float, f16_t, false>(src, region, n, c->sm_count, c->ps_in);
    rc = launch_check(c, "stage_in");
    if (rc) return rc;
    RT(cudaEventRecord(c->pe_in[reg], c->ps_in));
    RT(cudaStreamWaitEvent(c->ps_nv, c->pe_in[reg], 0));
    CollArgs a;
    base_args(c, &a);
    a.in = region; a.out = region;
    a.has_scale = has_scale; a.scale = scale;
    a.n = n;
    a.chunk = round_up((n + W - 1) / W, vec);
    a.symmetric = 1;
    a.sym_off = (size_t)(region - c->arena[c->rank]);
    int grid;
    plan_tiles(a.chunk, wsz, vec, cap, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    a.sig = make_sig(OPC_ALLREDUCE, dtype * 16 + wire, B200C_SUM, n, reg, B200C_ALGO_NVLS_STREAMS * 8 + i % 8);
    if (wire == B200C_FLOAT32) launch_nvls<float, float>(a, grid, c->ps_nv, false);
    else if (wire == B200C_BFLOAT16) launch_nvls<bf16_t, bf16_t>(a, grid, c->ps_nv, false);
    else launch_nvls<f16_t, f16_t>(a, grid, c->ps_nv, false);
    rc = launch_check(c, "nvls(stream pipeline)");
    if (rc) return rc;
    commit_args(c, a);
    RT(cudaEventRecord(c->pe_nv[reg], c->ps_nv));
    RT(cudaStreamWaitEvent(c->ps_out, c->pe_nv[reg], 0));
    if (wire == dtype) {
      if (esz == 4) launch_stage_copy<float, float, true>(region, dst, n, c->sm_count, c->ps_out);
      else launch_stage_copy<bf16_t, bf16_t, true>(region, dst, n, c->sm_count, c->ps_out);
    } else if (wire == B200C_BFLOAT16) launch_stage_copy<bf16_t, float, true>(region, dst, n, c->sm_count, c->ps_out);
    else launch_stage_copy<f16_t, float, true>(region, dst, n, c->sm_count, c->ps_out);
    rc = launch_check(c, "stage_out");
    if (rc) return rc;
    RT(cudaEventRecord(c->pe_out[reg], c->ps_out));
    e0 += n;
  }
  // close: the caller's stream continues after the last copy-out, and peers only move on after this rank got here
  RT(cudaStreamWaitEvent(user, c->pe_out[last_reg], 0));
  return barrier_op(c, user);
}

static int allreduce_impl(b200c_comm* c, const void* send, void* recv, size_t count, int dtype, int wire, int op, float scale,
                          int has_scale, int algo, cudaStream_t s) {
  int rc = check_ready(c);
  if (rc) return rc;
  size_t esz = b200c_dtype_size(dtype), wsz = b200c_dtype_size(wire);
  if (!esz || !wsz) return fail(B200C_EINVAL, "bad dtype %d / wire %d", dtype, wire);
  if (op < 0 || op >= B200C_NUM_OPS) return fail(B200C_EINVAL, "bad reduce op %d", op);
  if (count && (!send || !recv)) return fail(B200C_EINVAL, "null buffer");
  if (wire != dtype) {
    if (!(dtype == B200C_FLOAT32 && (wire == B200C_BFLOAT16 || wire == B200C_FLOAT16))) return fail(B200C_EUNSUPPORTED, "wire dtype %d for buffer dtype %d", wire, dtype);
    if (op != B200C_SUM && op != B200C_AVG) return fail(B200C_EUNSUPPORTED, "compressed wire supports SUM/AVG only");
  }
  if (algo < B200C_ALGO_AUTO || algo > B200C_ALGO_NVLS_STREAMS) return fail(B200C_EINVAL, "bad algo %d", algo);
  if (op == B200C_AVG) { has_scale = 1; scale = 1.f / (float)c->world; }
  if (count == 0) return B200C_OK;
  DeviceGuard g(c->device);
  const int W = c->world;
  const size_t vec = 16 / wsz;
  if (W == 1) {
    if (send == recv && !has_scale && wire == dtype) return B200C_OK;
    if (!has_scale && wire == dtype) { RT(cudaMemcpyAsync(recv, send, count * esz, cudaMemcpyDeviceToDevice, s)); return B200C_OK; }
    CollArgs a; memset(&a, 0, sizeof a);
    a.c = c->dev; a.in = send; a.out = recv; a.n = count; a.has_scale = has_scale; a.scale = scale;
    // Large, 16-byte aligned buffers stream through shared memory with the bulk copy engine (TMA); the
    // plain LSU kernel takes small buffers, unaligned views and the sub-tile tail.
    static const bool use_tma = [] { const char* e = getenv("B200COLL_LOCAL_SCALE_TMA"); return !e || e[0] != '0'; }();
    size_t tma_elems = 0;
    if (use_tma && count * esz >= (1u << 20) && ((uintptr_t)send & 15) == 0 && ((uintptr_t)recv & 15) == 0 &&
        (dtype == B200C_FLOAT32 || ((dtype == B200C_BFLOAT16 || dtype == B200C_FLOAT16) && wire == dtype))) {
      size_t ntiles = count * esz / kTmaTileBytes;
      tma_elems = ntiles * kTmaTileBytes / esz;
      if (c->tma_ctas_per_sm == 0) {
        int nb = 0;
        cudaFuncSetAttribute(k_local_scale_tma<float, bf16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
        cudaFuncSetAttribute(k_local_scale_tma<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
        cudaFuncSetAttribute(k_local_scale_tma<float, f16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
        cudaFuncSetAttribute(k_local_scale_tma<bf16_t, bf16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
        cudaFuncSetAttribute(k_local_scale_tma<f16_t, f16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_local_scale_tma<float, bf16_t>, kTmaThreads, kTmaSmemBytes) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
        c->tma_ctas_per_sm = nb;
      }
      // as many CTAs as are resident at once (measured: a smaller "balanced" grid of 384 x 5 tiles is slower, 13.7 vs
      // 12.9 us per 30 MiB bucket — what counts is how many bulk loads are in flight from the first microsecond)
      size_t cap = (size_t)c->sm_count * c->tma_ctas_per_sm;
      int tgrid = (int)(ntiles < cap ? ntiles : cap);
      switch (dtype * 16 + wire) {
        case B200C_FLOAT32 * 16 + B200C_FLOAT32: k_local_scale_tma<float, float><<<tgrid, kTmaThreads, kTmaSmemBytes, s>>>(a); break;
        case B200C_FLOAT32 * 16 + B200C_BFLOAT16: k_local_scale_tma<float, bf16_t><<<tgrid, kTmaThreads, kTmaSmemBytes, s>>>(a); break;
        case B200C_FLOAT32 * 16 + B200C_FLOAT16: k_local_scale_tma<float, f16_t><<<tgrid, kTmaThreads, kTmaSmemBytes, s>>>(a); break;
        case B200C_BFLOAT16 * 16 + B200C_BFLOAT16: k_local_scale_tma<bf16_t, bf16_t><<<tgrid, kTmaThreads, kTmaSmemBytes, s>>>(a); break;
        default: k_local_scale_tma<f16_t, f16_t><<<tgrid, kTmaThreads, kTmaSmemBytes, s>>>(a); break;
      }
      rc = launch_check(c, "local_scale_tma");
      if (rc) return rc;
      if (tma_elems == count) return B200C_OK;
      a.in = static_cast<const char*>(send) + tma_elems * esz;
      a.out = static_cast<char*>(recv) + tma_elems * esz;
      a.n = count - tma_elems;
    }
    int grid = local_scale_grid<float, bf16_t>(c, a.n * esz);
    switch (dtype * 16 + wire) {
      case B200C_FLOAT32 * 16 + B200C_FLOAT32: k_local_scale<float, float><<<grid, kThreads, 0, s>>>(a); break;
      case B200C_FLOAT32 * 16 + B200C_BFLOAT16: k_local_scale<float, bf16_t><<<grid, kThreads, 0, s>>>(a); break;
      case B200C_FLOAT32 * 16 + B200C_FLOAT16: k_local_scale<float, f16_t><<<grid, kThreads, 0, s>>>(a); break;
      case B200C_BFLOAT16 * 16 + B200C_BFLOAT16: k_local_scale<bf16_t, bf16_t><<<grid, kThreads, 0, s>>>(a); break;
      case B200C_FLOAT16 * 16 + B200C_FLOAT16: k_local_scale<f16_t, f16_t><<<grid, kThreads, 0, s>>>(a); break;
      case B200C_FLOAT64 * 16 + B200C_FLOAT64: k_local_scale<double, double><<<grid, kThreads, 0, s>>>(a); break;
      default:
        // integer AVG over one rank is the identity
        if (send != recv) RT(cudaMemcpyAsync(recv, send, count * esz, cudaMemcpyDeviceToDevice, s));
        return B200C_OK;
    }
    return launch_check(c, "local_scale");
  }

  const bool nvls_ok = c->mc_arena && (op == B200C_SUM || op == B200C_AVG) &&
                       (wire == B200C_FLOAT32 || wire == B200C_BFLOAT16 || wire == B200C_FLOAT16);
  if ((algo == B200C_ALGO_NVLS || algo == B200C_ALGO_NVLS_PIPE || algo == B200C_ALGO_NVLS_LANES || algo == B200C_ALGO_NVLS_STREAMS) && !nvls_ok) return fail(B200C_EUNSUPPORTED, "NVLS needs a bound multicast object, SUM/AVG and f32/bf16/f16");
  const bool ll_ok = wire == dtype && c->ll_words && count * esz <= c->ll_words * 4;
  if (algo == B200C_ALGO_LL && !ll_ok) return fail(B200C_EUNSUPPORTED, "LL needs wire == dtype and at most %zu bytes (ll_max_bytes)", c->ll_words * 4);
  // the zero-copy kernel is pure switch traffic (few CTAs are best); the staged kernels also do the local copies
  const uint32_t nvls_sym_cap = c->cfg.nvls_blocks && c->cfg.nvls_blocks < c->cfg.max_blocks ? c->cfg.nvls_blocks : c->cfg.max_blocks;
  const char* in = static_cast<const char*>(send);
  char* out = static_cast<char*>(recv);
  // in place, same dtype on the wire, inside the symmetric region: eligible for the zero-copy path
  bool in_sym = false;
  if (wire == dtype && send == recv && c->sym_bytes && (((uintptr_t)send) & 15) == 0 && (count * esz) % 16 == 0) {
    const char* base = c->arena[c->rank] + c->off_sym;
    in_sym = in >= base && in + count * esz <= base + c->sym_bytes;
  }
  if (!in_sym && (algo == B200C_ALGO_NVLS_STREAMS ||
                  (algo == B200C_ALGO_AUTO && nvls_ok && W >= 6 && c->cfg.nvls_streams_min_bytes && count * wsz >= c->cfg.nvls_streams_min_bytes)))
    return allreduce_streams(c, send, recv, count, dtype, wire, op, scale, has_scale, s);
  size_t done = 0;
  while (done < count) {
    size_t left = count - done;
    size_t bytes_left = left * wsz;
    int al = algo;
    bool pipe = false, lanes = false;
    if (al == B200C_ALGO_NVLS_PIPE) { al = B200C_ALGO_NVLS; pipe = true; }
    if (al == B200C_ALGO_NVLS_LANES) { al = B200C_ALGO_NVLS; lanes = true; }
    if (al == B200C_ALGO_AUTO) {
      if (ll_ok && count * esz <= c->cfg.ll_max_bytes) al = B200C_ALGO_LL;
      else if (bytes_left <= c->cfg.oneshot_max_bytes) al = B200C_ALGO_ONESHOT;
      else if (nvls_ok && W > 2 && bytes_left >= c->cfg.nvls_min_bytes && (W >= 6 || in_sym)) al = B200C_ALGO_NVLS;  // W = 4: two-shot beats STAGED NVLS (r02_sweep4_large.log)
      else al = B200C_ALGO_TWOSHOT;
    }
    CollArgs a;
    base_args(c, &a);
    a.in = in + done * esz; a.out = out + done * esz;
    a.has_scale = has_scale; a.scale = scale;
    size_t n;
    int grid;
    uint32_t rounds = 0;
    // symmetric zero-copy NVLS: buffer lives in the symmetric region at the same offset everywhere
    const bool sym = al == B200C_ALGO_NVLS && in_sym;
    if (al == B200C_ALGO_LL) {
      n = left;
      a.chunk = round_up(n, vec);
      a.tile = vec;
      a.ll_seq = c->ll_seq + 1;
      size_t nvec = (n * esz + 15) / 16;
      size_t nb = (nvec + kLLThreads - 1) / kLLThreads;
      grid = (int)(nb < 1 ? 1 : (nb > c->cfg.max_blocks ? c->cfg.max_blocks : nb));
    } else if (al == B200C_ALGO_ONESHOT) {
      size_t cap = c->cfg.staging_bytes / W / wsz / vec * vec;  // elements per slot
      n = left < cap ? left : cap;
      a.chunk = round_up(n, vec);
      plan_tiles(n, wsz, vec, c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    } else if (al == B200C_ALGO_TWOSHOT) {
      size_t cap_chunk = c->cfg.staging_bytes / W / wsz / vec * vec;
      size_t cap = cap_chunk * W;
      n = left < cap ? left : cap;
      a.chunk = round_up((n + W - 1) / W, vec);
      plan_tiles(a.chunk, wsz, vec, c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    } else {
      // symmetric buffers need no staging; pieces of at most 256 MiB keep the peers' TLB reach
      size_t cap = sym ? ((size_t)256 << 20) / wsz : c->cfg.staging_bytes / wsz / vec * vec;
      n = left < cap ? left : cap;
      a.chunk = round_up((n + W - 1) / W, vec);
      a.symmetric = sym ? 1 : 0;
      a.sym_off = sym ? (size_t)((const char*)a.in - c->arena[c->rank]) : 0;
      if (!sym && algo == B200C_ALGO_AUTO && c->cfg.nvls_lanes_min_bytes && left * wsz >= c->cfg.nvls_lanes_min_bytes) lanes = true;
      if (!sym && !lanes && algo == B200C_ALGO_AUTO && c->cfg.nvls_pipe_min_bytes && n * wsz >= c->cfg.nvls_pipe_min_bytes) pipe = true;
      if (sym) pipe = lanes = false;  // nothing to overlap: the symmetric path has no staging copies
      if (lanes) {
        // the ring is rewritten every three rounds, so the staging capacity does not bound the piece: take it all
        n = left;
        a.chunk = round_up((n + W - 1) / W, vec);
        uint32_t L = c->cfg.nvls_lanes;
        uint32_t Kc = c->cfg.max_blocks / L - 1;
        if (Kc > 7) Kc = 7;
        if (Kc < 1) Kc = 1;
        // granule: the configured size, smaller for messages that would otherwise give a lane fewer than ~4 rounds
        size_t tb = c->cfg.lane_granule_bytes, chunk_bytes = a.chunk * wsz;
        size_t want = chunk_bytes / ((size_t)L * 4) / 8192 * 8192;
        if (want < 8192) want = 8192;
        if (tb > want) tb = want;
        size_t ring = (size_t)L * kLaneSlots * W * tb;
        while (ring > c->cfg.staging_bytes && tb > 8192) { tb -= 8192; ring = (size_t)L * kLaneSlots * W * tb; }
        while (ring > c->cfg.staging_bytes && L > 1) { L--; ring = (size_t)L * kLaneSlots * W * tb; }   // a small staging area: fewer lanes
        if (ring > c->cfg.staging_bytes) return fail(B200C_EINVAL, "staging_bytes too small for the lane kernel's ring (%zu bytes)", ring);
        a.tile = tb / wsz;
        a.lane_copy = (int)Kc;
        size_t ngran = (a.chunk + a.tile - 1) / a.tile;
        uint32_t used = (uint32_t)(ngran < L ? ngran : L);   // lanes that own at least one granule
        grid = (int)(used * (1 + Kc));
        rounds = (uint32_t)((ngran + used - 1) / used);
        a.pipe_base = c->pipe_base;
      } else if (pipe) {
        plan_rounds(a.chunk, wsz, vec, c->cfg.max_blocks, c->cfg.granule_bytes, &a.tile, &grid, &rounds);
        a.pipe_base = c->pipe_base;
      } else {
        plan_tiles(a.chunk, wsz, vec, sym ? nvls_sym_cap : c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
      }
    }
    a.n = n;
    // a symmetric buffer must sit at the same arena offset on every rank: the offset is part of the signature
    a.sig = make_sig(OPC_ALLREDUCE, dtype * 16 + wire, op, n, sym ? (int)((a.sym_off >> 4) & 0x3fffffff) : -1, al * 8 + (sym ? 1 : 0) + (pipe ? 2 : 0) + (lanes ? 4 : 0));
    if (al == B200C_ALGO_NVLS) {
      if (dtype == B200C_FLOAT32 && wire == B200C_FLOAT32) launch_nvls<float, float>(a, grid, s, pipe, lanes);
      else if (dtype == B200C_BFLOAT16) launch_nvls<bf16_t, bf16_t>(a, grid, s, pipe, lanes);
      else if (dtype == B200C_FLOAT16) launch_nvls<f16_t, f16_t>(a, grid, s, pipe, lanes);
      else if (wire == B200C_BFLOAT16) launch_nvls<float, bf16_t>(a, grid, s, pipe, lanes);
      else launch_nvls<float, f16_t>(a, grid, s, pipe, lanes);
    } else if (wire != dtype) {
      if (wire == B200C_BFLOAT16) launch_mixed<float, bf16_t>(al, a, grid, s);
      else launch_mixed<float, f16_t>(al, a, grid, s);
    } else {
      int kind = al == B200C_ALGO_LL ? KIND_LL : (al == B200C_ALGO_ONESHOT ? KIND_ONESHOT : KIND_TWOSHOT);
      rc = launch_same_type(dtype, kind, op, a, grid, s);
      if (rc) return rc;
    }
    rc = launch_check(c, "allreduce");
    if (rc) return rc;
    commit_args(c, a);
    c->pipe_base += rounds;
    if (al == B200C_ALGO_LL) c->ll_seq = a.ll_seq;
    done += n;
  }
  return B200C_OK;
}

extern "C" int b200c_allreduce(b200c_comm_t* c, const void* send, void* recv, size_t count, int dtype, int op, int algo,
                               b200c_stream_t stream) {
  // AVG on integers: SUM then truncating divide by world (ncclAvg semantics); handled by apply_scale.
  return allreduce_impl(c, send, recv, count, dtype, dtype, op, 1.f, 0, algo, (cudaStream_t)stream);
}

extern "C" int b200c_allreduce_scaled(b200c_comm_t* c, const void* send, void* recv, size_t count, int dtype, int wire_dtype,
                                      float scale, int algo, b200c_stream_t stream) {
  if (dtype != B200C_FLOAT32 && dtype != B200C_BFLOAT16 && dtype != B200C_FLOAT16)
    return fail(B200C_EUNSUPPORTED, "scaled allreduce supports f32/bf16/f16 buffers, got %d", dtype);
  return allreduce_impl(c, send, recv, count, dtype, wire_dtype, B200C_SUM, scale, 1, algo, (cudaStream_t)stream);
}

extern "C" int b200c_reduce(b200c_comm_t* c, const void* send, void* recv, size_t count, int dtype, int op, int root,
                            b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  size_t esz = b200c_dtype_size(dtype);
  if (!esz) return fail(B200C_EINVAL, "bad dtype %d", dtype);
  if (op < 0 || op >= B200C_NUM_OPS) return fail(B200C_EINVAL, "bad reduce op %d", op);
  if (root < 0 || root >= c->world) return fail(B200C_EINVAL, "bad root %d", root);
  if (count == 0) return B200C_OK;
  if (!send || (c->rank == root && !recv)) return fail(B200C_EINVAL, "null buffer");
  cudaStream_t s = (cudaStream_t)stream;
  DeviceGuard g(c->device);
  if (c->world == 1) {
    if (send != recv) RT(cudaMemcpyAsync(recv, send, count * esz, cudaMemcpyDeviceToDevice, s));
    return B200C_OK;
  }
  const size_t vec = 16 / esz;
  size_t cap = c->cfg.staging_bytes / c->world / esz / vec * vec;
  size_t done = 0;
  while (done < count) {
    size_t n = count - done < cap ? count - done : cap;
    CollArgs a;
    base_args(c, &a);
    a.in = static_cast<const char*>(send) + done * esz;
    a.out = recv ? static_cast<char*>(recv) + done * esz : nullptr;
    a.n = n; a.chunk = round_up(n, vec); a.root = root;
    if (op == B200C_AVG) { a.has_scale = 1; a.scale = 1.f / c->world; }
    int grid;
    plan_tiles(n, esz, vec, c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    a.sig = make_sig(OPC_REDUCE, dtype, op, n, root, 0);
    rc = launch_same_type(dtype, KIND_REDUCE, op, a, grid, s);
    if (rc) return rc;
    rc = launch_check(c, "reduce");
    if (rc) return rc;
    commit_args(c, a);
    done += n;
  }
  return B200C_OK;
}

extern "C" int b200c_reducescatter(b200c_comm_t* c, const void* const* send_ptrs, void* recv, size_t count, int dtype, int op,
                                   b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  size_t esz = b200c_dtype_size(dtype);
  if (!esz) return fail(B200C_EINVAL, "bad dtype %d", dtype);
  if (op < 0 || op >= B200C_NUM_OPS) return fail(B200C_EINVAL, "bad reduce op %d", op);
  if (count == 0) return B200C_OK;
  if (!send_ptrs || !recv) return fail(B200C_EINVAL, "null buffer");
  for (int j = 0; j < c->world; j++) if (!send_ptrs[j]) return fail(B200C_EINVAL, "send_ptrs[%d] is null", j);
  cudaStream_t s = (cudaStream_t)stream;
  DeviceGuard g(c->device);
  if (c->world == 1) {
    if (send_ptrs[0] != recv) RT(cudaMemcpyAsync(recv, send_ptrs[0], count * esz, cudaMemcpyDeviceToDevice, s));
    return B200C_OK;
  }
  const size_t vec = 16 / esz;
  size_t cap = c->cfg.staging_bytes / c->world / esz / vec * vec;
  size_t done = 0;
  while (done < count) {
    size_t n = count - done < cap ? count - done : cap;
    CollArgs a;
    base_args(c, &a);
    for (int j = 0; j < c->world; j++) a.in_ptrs[j] = static_cast<const char*>(send_ptrs[j]) + done * esz;
    a.out = static_cast<char*>(recv) + done * esz;
    a.n = n; a.chunk = round_up(n, vec);
    if (op == B200C_AVG) { a.has_scale = 1; a.scale = 1.f / c->world; }
    int grid;
    plan_tiles(n, esz, vec, c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    a.sig = make_sig(OPC_REDUCESCATTER, dtype, op, n, -1, 0);
    rc = launch_same_type(dtype, KIND_REDUCESCATTER, op, a, grid, s);
    if (rc) return rc;
    rc = launch_check(c, "reducescatter");
    if (rc) return rc;
    commit_args(c, a);
    done += n;
  }
  return B200C_OK;
}

extern "C" int b200c_allgather(b200c_comm_t* c, const void* send, void* const* recv_ptrs, size_t count, int dtype,
                               b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  size_t esz = b200c_dtype_size(dtype);
  if (!esz) return fail(B200C_EINVAL, "bad dtype %d", dtype);
  if (count == 0) return B200C_OK;
  if (!send || !recv_ptrs) return fail(B200C_EINVAL, "null buffer");
  for (int j = 0; j < c->world; j++) if (!recv_ptrs[j]) return fail(B200C_EINVAL, "recv_ptrs[%d] is null", j);
  cudaStream_t s = (cudaStream_t)stream;
  DeviceGuard g(c->device);
  size_t bytes = count * esz;
  if (c->world == 1) {
    if (send != recv_ptrs[0]) RT(cudaMemcpyAsync(recv_ptrs[0], send, bytes, cudaMemcpyDeviceToDevice, s));
    return B200C_OK;
  }
  size_t cap = c->cfg.staging_bytes / c->world / 16 * 16;
  size_t done = 0;
  while (done < bytes) {
    size_t n = bytes - done < cap ? bytes - done : cap;
    CollArgs a;
    base_args(c, &a);
    a.in = static_cast<const char*>(send) + done;
    for (int j = 0; j < c->world; j++) a.out_ptrs[j] = static_cast<char*>(recv_ptrs[j]) + done;
    a.n = n; a.chunk = round_up(n, 16);
    int grid;
    plan_tiles(n, 1, 16, c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    a.sig = make_sig(OPC_ALLGATHER, dtype, 0, n, -1, 0);
    k_allgather<<<grid, kThreads, 0, s>>>(a);
    rc = launch_check(c, "allgather");
    if (rc) return rc;
    commit_args(c, a);
    done += n;
  }
  return B200C_OK;
}

extern "C" int b200c_broadcast(b200c_comm_t* c, void* buf, size_t count, int dtype, int root, b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  size_t esz = b200c_dtype_size(dtype);
  if (!esz) return fail(B200C_EINVAL, "bad dtype %d", dtype);
  if (root < 0 || root >= c->world) return fail(B200C_EINVAL, "bad root %d", root);
  if (count == 0 || c->world == 1) return B200C_OK;
  if (!buf) return fail(B200C_EINVAL, "null buffer");
  cudaStream_t s = (cudaStream_t)stream;
  DeviceGuard g(c->device);
  const int W = c->world;
  size_t bytes = count * esz, cap = c->cfg.staging_bytes / 16 * 16, done = 0;
  while (done < bytes) {
    size_t n = bytes - done < cap ? bytes - done : cap;
    CollArgs a;
    base_args(c, &a);
    a.in = static_cast<char*>(buf) + done; a.out = static_cast<char*>(buf) + done;
    a.n = n; a.root = root;
    int grid;
    uint32_t rounds = 0;
    // every rank takes the same decision from (n, world, multicast)
    const bool mc = c->mc_arena && c->bcast_mc;
    const bool big = c->cfg.bcast_rounds_min_bytes && n >= c->cfg.bcast_rounds_min_bytes;
    const bool mc_rounds = big && mc && W > 2 && n % 16 == 0;     // scatter + multicast allgather
    const bool uc_rounds = big && !mc_rounds && (W == 2 || !mc);  // pipelined unicast push
    if (mc_rounds && (((uintptr_t)a.in) & 15) != 0)
      return fail(B200C_EINVAL, "broadcast of >= %llu bytes needs a 16-byte aligned buffer", (unsigned long long)c->cfg.bcast_rounds_min_bytes);
    if (mc_rounds) {
      a.chunk = round_up((n + W - 1) / W, 16);
      plan_rounds(a.chunk, 1, 16, c->cfg.max_blocks, c->cfg.granule_bytes, &a.tile, &grid, &rounds);
      a.pipe_base = c->pipe_base;
      a.symmetric = 2;
    } else if (uc_rounds) {
      a.chunk = round_up(n, 16);
      plan_rounds(a.chunk, 1, 16, c->cfg.max_blocks, c->cfg.granule_bytes, &a.tile, &grid, &rounds);
      a.pipe_base = c->pipe_base;
      a.symmetric = 3;
    } else {
      a.chunk = round_up(n, 16);
      a.symmetric = (mc && n >= 65536) ? 1 : 0;  // multicast store from the root (same choice on every rank)
      plan_tiles(n, 1, 16, c->cfg.max_blocks, kMinTileBytes, c->cfg.granule_bytes, &a.tile, &grid);
    }
    const bool rounds_mode = mc_rounds || uc_rounds;
    a.sig = make_sig(OPC_BROADCAST, dtype, 0, n, root, a.symmetric);
    if (rounds_mode) k_broadcast_rounds<<<grid, kThreads, 0, s>>>(a);
    else k_broadcast<<<grid, kThreads, 0, s>>>(a);
    rc = launch_check(c, "broadcast");
    if (rc) return rc;
    commit_args(c, a);
    c->pipe_base += rounds;
    done += n;
  }
  return B200C_OK;
}

extern "C" int b200c_barrier(b200c_comm_t* c, b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (c->world == 1) return B200C_OK;
  DeviceGuard g(c->device);
  CollArgs a;
  base_args(c, &a);
  a.sig = make_sig(OPC_BARRIER, 0, 0, 0, -1, 0);
  k_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(a);
  rc = launch_check(c, "barrier");
  if (rc) return rc;
  commit_args(c, a);
  return B200C_OK;
}

static int p2p_impl(b200c_comm* c, void* buf, size_t bytes, int peer, bool is_send, cudaStream_t s) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (peer < 0 || peer >= c->world) return fail(B200C_EINVAL, "bad peer %d", peer);
  if (peer == c->rank) return fail(B200C_EINVAL, "send/recv to self (rank %d)", peer);
  if (bytes == 0) return B200C_OK;
  if (!buf) return fail(B200C_EINVAL, "null buffer");
  DeviceGuard g(c->device);
  P2PArgs a;
  memset(&a, 0, sizeof a);
  a.c = c->dev; a.buf = buf; a.bytes = bytes; a.peer = peer;
  a.off_ring = c->off_p2p; a.off_ready = kOffP2PReady; a.off_ack = kOffP2PAck; a.cells = (int)c->cfg.p2p_slots;
  size_t cb = c->cfg.p2p_slot_bytes;
  size_t ncells = (bytes + cb - 1) / cb;
  if (ncells > 0x7fffffffull) return fail(B200C_EINVAL, "message too large for the cell ring");
  uint32_t* ctr = is_send ? &c->send_cells[peer] : &c->recv_cells[peer];
  a.first_cell = *ctr; a.ncells = (uint32_t)ncells;
  // no block may wait on a cell that one of its own later iterations has to free: grid <= ring size
  uint32_t grid = (uint32_t)ncells;
  uint32_t lim = c->cfg.max_blocks < c->cfg.p2p_slots ? c->cfg.max_blocks : c->cfg.p2p_slots;
  if (is_send) {
    // A sender block publishes `batch` cells per release fence (stride grid).  Small messages keep batch 1
    // (one cell per block: lowest latency); large ones amortise the fence.  All cells of one pass over the
    // grid must fit in the ring together, or a block would wait for an ack that only a later cell of its own
    // pass triggers: grid * batch <= ring cells.
    static int env_batch = [] { const char* e = getenv("B200COLL_SEND_BATCH"); return e ? atoi(e) : 0; }();
    uint32_t batch = (uint32_t)((ncells + lim - 1) / lim);
    if (env_batch > 0) batch = (uint32_t)env_batch;
    if (batch < 1) batch = 1;
    if (batch > (uint32_t)kSendBatch) batch = kSendBatch;
    a.batch = (int)batch;
    uint32_t lim_s = c->cfg.p2p_slots / batch;
    if (lim_s < 1) lim_s = 1;
    if (lim > lim_s) lim = lim_s;
    uint32_t want = ((uint32_t)ncells + batch - 1) / batch;
    grid = want < 1 ? 1 : want;
  }
  if (grid > lim) grid = lim;
  if (is_send) k_send<<<grid, kThreads, 0, s>>>(a);
  else k_recv<<<grid, kThreads, 0, s>>>(a);
  rc = launch_check(c, is_send ? "send" : "recv");
  if (rc) return rc;
  *ctr += (uint32_t)ncells;  // the ring position only advances once the kernel is really queued
  return B200C_OK;
}
extern "C" int b200c_send(b200c_comm_t* c, const void* buf, size_t bytes, int peer, b200c_stream_t stream) {
  return p2p_impl(c, const_cast<void*>(buf), bytes, peer, true, (cudaStream_t)stream);
}
extern "C" int b200c_recv(b200c_comm_t* c, void* buf, size_t bytes, int peer, b200c_stream_t stream) {
  return p2p_impl(c, buf, bytes, peer, false, (cudaStream_t)stream);
}

extern "C" int b200c_send_multi(b200c_comm_t* c, const void* buf, size_t bytes, const int* peers, int npeers, b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (!peers || npeers < 1) return fail(B200C_EINVAL, "no readers");
  uint32_t mask = 0;
  for (int i = 0; i < npeers; i++) {
    if (peers[i] < 0 || peers[i] >= c->world) return fail(B200C_EINVAL, "bad peer %d", peers[i]);
    if (peers[i] == c->rank) return fail(B200C_EINVAL, "send to self (rank %d)", peers[i]);
    if (mask & (1u << peers[i])) return fail(B200C_EINVAL, "peer %d listed twice", peers[i]);
    mask |= 1u << peers[i];
  }
  if (npeers == 1 || c->mcells == 0) {   // a single reader is the pairwise ring (the reader calls b200c_recv)
    if (npeers != 1) return fail(B200C_ESTATE, "multi-reader ring unavailable");
    return p2p_impl(c, const_cast<void*>(buf), bytes, peers[0], true, (cudaStream_t)stream);
  }
  if (c->msend_mask && c->msend_mask != mask)
    return fail(B200C_EUNSUPPORTED, "rank %d already multi-sends to reader set 0x%x; a second reader set (0x%x) must use per-reader sends "
                "(ring positions are counted per source, so every reader has to see every message)", c->rank, c->msend_mask, mask);
  if (bytes == 0) return B200C_OK;
  if (!buf) return fail(B200C_EINVAL, "null buffer");
  DeviceGuard g(c->device);
  P2PArgs a;
  memset(&a, 0, sizeof a);
  a.c = c->dev; a.buf = const_cast<void*>(buf); a.bytes = bytes; a.peer = -1; a.reader_mask = mask;
  a.off_ring = c->off_mring; a.off_ready = kOffMReady; a.off_ack = kOffMAck; a.cells = (int)c->mcells;
  size_t cb = c->cfg.p2p_slot_bytes;
  size_t ncells = (bytes + cb - 1) / cb;
  if (ncells > 0x7fffffffull) return fail(B200C_EINVAL, "message too large for the cell ring");
  a.first_cell = c->msend_cells; a.ncells = (uint32_t)ncells;
  uint32_t grid = (uint32_t)ncells;
  uint32_t lim = c->cfg.max_blocks < c->mcells ? c->cfg.max_blocks : c->mcells;
  if (grid > lim) grid = lim;
  k_send_multi<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
  rc = launch_check(c, "send_multi");
  if (rc) return rc;
  c->msend_cells += (uint32_t)ncells;
  c->msend_mask = mask;
  return B200C_OK;
}

extern "C" int b200c_recv_multi(b200c_comm_t* c, void* buf, size_t bytes, int src, b200c_stream_t stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (src < 0 || src >= c->world || src == c->rank) return fail(B200C_EINVAL, "bad source %d", src);
  if (c->mcells == 0) return p2p_impl(c, buf, bytes, src, false, (cudaStream_t)stream);
  if (bytes == 0) return B200C_OK;
  if (!buf) return fail(B200C_EINVAL, "null buffer");
  DeviceGuard g(c->device);
  P2PArgs a;
  memset(&a, 0, sizeof a);
  a.c = c->dev; a.buf = buf; a.bytes = bytes; a.peer = src;
  a.off_ring = c->off_mring; a.off_ready = kOffMReady; a.off_ack = kOffMAck; a.cells = (int)c->mcells;
  size_t cb = c->cfg.p2p_slot_bytes;
  size_t ncells = (bytes + cb - 1) / cb;
  if (ncells > 0x7fffffffull) return fail(B200C_EINVAL, "message too large for the cell ring");
  a.first_cell = c->mrecv_cells[src]; a.ncells = (uint32_t)ncells;
  uint32_t grid = (uint32_t)ncells;
  uint32_t lim = c->cfg.max_blocks < c->mcells ? c->cfg.max_blocks : c->mcells;
  if (grid > lim) grid = lim;
  k_recv<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
  rc = launch_check(c, "recv_multi");
  if (rc) return rc;
  c->mrecv_cells[src] += (uint32_t)ncells;
  return B200C_OK;
}
