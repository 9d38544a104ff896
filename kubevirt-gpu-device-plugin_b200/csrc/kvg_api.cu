// kvg_api.cu — the C-ABI of libkvgpu.so (include/kvgpu.h): context, HBM layout, launch sequencing
// and result marshalling around the kernels in kvg_parse.cuh / kvg_scan.cuh.
//
// HBM layout owned by a context (all cudaMalloc'd once and grown geometrically, never per call):
//   text      pci.ids image(s), padded with '\n' to a tile multiple + 16 (TMA halo)
//   tables    open-addressed u64 slots  key(vendor<<16|device)<<32 | line offset
//   pool      sanitised names of the NVIDIA section, slot = line offset - section offset
//   recs      record staging (host entry points only)
//   surv      compacted survivors, Walk order
//   sort      2 x (keys,vals) ping-pong per ordering (device-id ordering, iommu-group ordering)
//   seg       distinct keys + offsets per ordering
// No CPU fallback exists anywhere below: every compute entry point fails with KVG_ECUDA when the
// CUDA runtime is unusable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/kvgpu.h"
#include "kvg_common.cuh"
#include "kvg_parse.cuh"
#include "kvg_parse_v2.cuh"
#include "kvg_scan.cuh"
#include "kvg_radix_exp.cuh"

using namespace kvg;

// ------------------------------------------------------------------------------------------------
// minimal NCCL surface, resolved with dlopen so single-GPU users need no NCCL at all
// ------------------------------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclUint32 = 3, ncclUint64 = 5 };
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string* err) {
    if (handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return false;
    }
#define KVG_SYM(field, name)                                   \
  *(void**)(&field) = dlsym(handle, name);                     \
  if (!field) {                                                \
    *err = std::string("NCCL symbol missing: ") + name;        \
    return false;                                              \
  }
    KVG_SYM(GetUniqueId, "ncclGetUniqueId");
    KVG_SYM(CommInitRank, "ncclCommInitRank");
    KVG_SYM(CommDestroy, "ncclCommDestroy");
    KVG_SYM(AllGather, "ncclAllGather");
    KVG_SYM(Broadcast, "ncclBroadcast");
    KVG_SYM(GroupStart, "ncclGroupStart");
    KVG_SYM(GroupEnd, "ncclGroupEnd");
    KVG_SYM(GetErrorString, "ncclGetErrorString");
#undef KVG_SYM
    return true;
  }
};
static NcclApi g_nccl;
static std::string g_create_error;
// Look-back state words carry a launch epoch so they never need clearing.  The epoch is
// PROCESS-global (not per context): cudaMalloc may hand a context memory that another context
// just freed, and a per-context counter would let a stale word alias a live epoch.  Newly
// allocated buffers are zero-filled as well (epoch 0 is never issued).
static std::atomic<uint32_t> g_epoch{0};
static inline uint32_t next_epoch() {
  uint32_t e = g_epoch.fetch_add(1, std::memory_order_relaxed) + 1;
  if ((e & 0x3fffffffu) == 0) e = g_epoch.fetch_add(1, std::memory_order_relaxed) + 1;
  return e;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
};

// page-locked byte buffer with the few std::vector members the name-pool mirror uses: device-to-host
// copies into it are truly asynchronous (a pageable destination is staged synchronously by the runtime)
struct PinnedBytes {
  uint8_t* p = nullptr;
  size_t len = 0, cap = 0;
  ~PinnedBytes() {
    if (p) cudaFreeHost(p);
  }
  size_t size() const { return len; }
  uint8_t* data() { return p; }
  uint8_t& operator[](size_t i) { return p[i]; }
  void clear() { len = 0; }
  bool resize(size_t n) {
    if (n > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      if (cudaMallocHost((void**)&p, n + n / 4 + 4096) != cudaSuccess) return false;
      cap = n + n / 4 + 4096;
    }
    len = n;
    return true;
  }
};

struct PinnedBlock {
  void* p;
  size_t size;
};

struct OrderBufs {  // one stable ordering (sorted (key, index) pairs + segment heads)
  DevBuf<uint2> p0, p1;           // ping-pong {key, survivor index}
  DevBuf<uint32_t> perm;          // final permutation
  DevBuf<uint32_t> seg_key, seg_off, seg_name;
  DevBuf<uint32_t> tile_heads, tile_off;
  DevBuf<uint64_t> heads_state;
};

struct kvg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::string err;
  uint64_t launches = 0;
  uint32_t epoch = 1;

  // pci.ids
  DevBuf<uint8_t> text;  // owned copy (host entry point)
  const uint8_t* d_text = nullptr;
  uint32_t text_len = 0;
  DevBuf<uint64_t> tables;
  uint32_t cap_log2 = 0;
  DevBuf<PciIdsInfo> info;
  DevBuf<uint32_t> tile_arrays;  // 3 x n_tiles
  DevBuf<uint64_t> parse_state;
  DevBuf<uint32_t> v2_state, v2_pending;  // KVG_PARSE=v2 (experimental): span states + pending counts, pending lines
  DevBuf<uint32_t> parse_ticket;
  DevBuf<uint8_t> pool;
  DevBuf<uint32_t> nv_index;  // [65536] vendor-10de device id -> name pool slot
  DevBuf<uint32_t> nv_lines;  // [65536 + 1] offsets of the lines that have a name, then their count
  DevBuf<uint32_t> sec_lines; // '\t' line starts of the NVIDIA section, then their count (general lookups)
  size_t sec_lines_cap = 0;
  DevBuf<uint64_t> type_hash;
  PinnedBytes h_pool;
  PciIdsInfo h_info{};
  bool table_ready = false;
  // kvg_pciids_load only ENQUEUES (copy + parse); the host-side publication (overflow check, name
  // pool mirror) happens in the first call that consumes the table -> the next call's host-to-device
  // copy overlaps the parse
  bool load_pending = false;
  size_t pend_len = 0, pend_stride = 0;
  uint32_t pend_cap_log2 = 0;
  uint32_t parsed_files = 0;
  int parse_grid = 0;

  // scans
  DevBuf<ScanCtrl> ctrl;
  ScanCtrl* h_ctrl = nullptr;  // pinned
  DevBuf<uint4> recs;          // staging for host entry points
  DevBuf<uint4> surv;
  DevBuf<uint4> ragged;          // tile-local survivor scratch of k_classify_ragged
  DevBuf<uint32_t> tile_count, tile_off;
  DevBuf<uint2> tile_max;
  DevBuf<uint64_t> offs_state;
  DevBuf<uint64_t> classify_state;
  OrderBufs ord_dev, ord_grp;
  DevBuf<uint32_t> tile_hist, bin_total;
  bool scatter_smem_set = false;  // dynamic shared-memory opt-in of k_radix_scatter<11> done on this device
  bool scatter_c_smem_set = false;
  size_t last_n = 0;     // records of the last enqueued scan
  size_t last_total = 0; // survivors capacity used by the last scan (sharded: all ranks)
  int last_kind = 0;     // 1 = pci, 2 = mdev
  bool last_owned = false;  // sharded scan: the orderings cover only the keys this rank owns
  // mdev dictionary
  DevBuf<uint8_t> type_raw, type_label;
  DevBuf<uint32_t> type_off, type_label_len, type_match, type_name_len;
  DevBuf<uint16_t> type_canon;
  DevBuf<uint8_t> type_names;
  uint32_t n_types = 0;
  std::vector<uint32_t> h_type_off;
  // health
  DevBuf<uint8_t> alive_prev;
  DevBuf<uint32_t> changed;
  size_t health_n = 0;
  // misc
  DevBuf<uint4> flush;
  DevBuf<uint16_t> nv_ids;
  DevBuf<uint32_t> probe_slots;
  DevBuf<uint8_t> keys_blob;
  DevBuf<uint32_t> keys_off, match_off, match_len;
  DevBuf<uint8_t> match_out;
  std::vector<PinnedBlock> pinned_free;
  void* h_stage = nullptr;
  size_t h_stage_cap = 0;
  // pipelined host entry point (kvg_scan_pci): copy streams, per-chunk events, mapped counters
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> pipe_ev;
  uint32_t* h_pipe = nullptr;  // pinned + mapped: cumulative survivor count after each chunk
  // kernel timing
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  std::vector<std::string> ev_names;
  size_t ev_used = 0;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  DevBuf<uint64_t> gather_counts;  // [nranks]
  DevBuf<uint4> local_surv;
  uint64_t* h_counts = nullptr;  // pinned [nranks]
  // peer-memory gather (CUDA IPC over NVLink)
  bool p2p = false;
  uint8_t* p2p_mine = nullptr;            // [P2PCtrl pad 4 KiB][window 0][window 1]
  size_t p2p_cap = 0;                     // survivors per region
  uint8_t* p2p_peer[P2P_MAX_RANKS] = {};  // peer-mapped bases (own entry = p2p_mine)
  unsigned long long p2p_step = 0;
  DevBuf<uint32_t> gather_base;
  DevBuf<uint32_t> p2p_err;
};
static const size_t P2P_HDR = 4096;

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
      return KVG_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

template <class T>
static int ensure(kvg_ctx* ctx, DevBuf<T>& b, size_t n) {
  if (n <= b.cap && b.p) return KVG_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  size_t cap = n + n / 4 + 64;
  cudaError_t e = cudaMalloc((void**)&b.p, cap * sizeof(T));
  if (e != cudaSuccess) {
    b.cap = 0;
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? KVG_ENOMEM : KVG_ECUDA;
  }
  b.cap = cap;
  cudaMemsetAsync(b.p, 0, cap * sizeof(T), ctx->stream);
  return KVG_OK;
}
#define ENSURE(buf, n)                      \
  do {                                      \
    int rc_ = ensure(ctx, buf, n);          \
    if (rc_) return rc_;                    \
  } while (0)

template <class T>
static void release(DevBuf<T>& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

// launch bookkeeping: count + optional CUDA-event bracket on the context stream
struct LaunchScope {
  kvg_ctx* ctx;
  size_t idx = (size_t)-1;
  LaunchScope(kvg_ctx* c, const char* name) : ctx(c) {
    ctx->launches++;
    if (ctx->timing) {
      if (ctx->ev_used == ctx->ev.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ctx->ev.push_back({a, b});
        ctx->ev_names.push_back("");
      }
      idx = ctx->ev_used++;
      ctx->ev_names[idx] = name;
      cudaEventRecord(ctx->ev[idx].first, ctx->stream);
    }
  }
  ~LaunchScope() {
    if (idx != (size_t)-1) cudaEventRecord(ctx->ev[idx].second, ctx->stream);
  }
};
// Launches carry the programmatic-stream-serialization attribute (PDL): the next kernel's CTAs are
// scheduled while this one drains and park in griddepcontrol.wait (pdl_enter() in every kernel), which
// removes most of the dependent-launch gap of the ~40-kernel scan.  KVG_PDL=0 restores plain launches.
// KVG_TRACE=1: host-clock phase marks of the pipelined entry point on stderr (diagnostics only)
struct PhaseTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  std::string line;
  PhaseTrace() {
    static const bool env = [] { const char* e = getenv("KVG_TRACE"); return e && e[0] == '1'; }();
    on = env;
    if (on) t0 = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    char buf[64];
    snprintf(buf, sizeof buf, " %s=%.0f", what, us);
    line += buf;
  }
  ~PhaseTrace() {
    if (on) fprintf(stderr, "[kvg trace us]%s\n", line.c_str());
  }
};
static PhaseTrace* g_trace = nullptr;
#define TRACE(what) do { if (g_trace) g_trace->mark(what); } while (0)

static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KVG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
template <class... KArgs, class... Args>
static void launch_kernel(kvg_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl_enabled() && !ctx->timing) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define LAUNCH(name, kernel, grid, block, smem, ...)                                   \
  do {                                                                                 \
    LaunchScope ls_(ctx, name);                                                        \
    launch_kernel(ctx, kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__);  \
  } while (0)

static int check_launch(kvg_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    return KVG_ECUDA;
  }
  return KVG_OK;
}

static void* pinned_alloc(kvg_ctx* ctx, size_t size) {
  for (size_t i = 0; i < ctx->pinned_free.size(); i++) {
    if (ctx->pinned_free[i].size >= size + 64) {
      void* p = ctx->pinned_free[i].p;
      ctx->pinned_free.erase(ctx->pinned_free.begin() + i);
      return p;
    }
  }
  void* p = nullptr;
  size_t cap = size + size / 4 + 4096;
  if (cudaMallocHost(&p, cap + 64) != cudaSuccess) return nullptr;
  // header: owning ctx + capacity, so kvg_result_free can recycle without a ctx argument
  ((uint64_t*)p)[0] = (uint64_t)(uintptr_t)ctx;
  ((uint64_t*)p)[1] = cap;
  return p;
}
static inline uint8_t* pinned_payload(void* blk) { return (uint8_t*)blk + 64; }

// grid for the TMA-pipelined classify kernels: co-resident, capped by the tile count
template <class Op, int ROWS, int STAGES>
static int classify_grid(kvg_ctx* ctx, size_t n_items, size_t* smem_out) {
  static int occ = 0;
  const size_t smem = (size_t)STAGES * KVG_BLOCK * ROWS * Op::REC_BYTES;
  if (!occ) {
    cudaFuncSetAttribute(k_classify_tma<Op, ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_classify_tma<Op, ROWS, STAGES>, KVG_BLOCK, smem);
    if (occ < 1) occ = 1;
  }
  *smem_out = smem;
  size_t tiles = (n_items + (size_t)KVG_BLOCK * ROWS - 1) / ((size_t)KVG_BLOCK * ROWS);
  size_t g = (size_t)ctx->sm_count * (size_t)occ;
  if (g > CLASSIFY_MAX_GRID) g = CLASSIFY_MAX_GRID;  // look-back batch covers 32*LB_KMAX CTAs
  if (tiles < g) g = tiles;
  return g < 1 ? 1 : (int)g;
}
template <class Op, int ROWS, int STAGES>
static int classify_ws_grid(kvg_ctx* ctx, size_t n_items, size_t* smem_out) {
  static int occ = 0;
  const size_t smem = (size_t)STAGES * KVG_BLOCK * ROWS * Op::REC_BYTES;
  if (!occ) {
    cudaFuncSetAttribute(k_classify_ws<Op, ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_classify_ws<Op, ROWS, STAGES>, WS_THREADS, smem);
    if (occ < 1) occ = 1;
  }
  *smem_out = smem;
  size_t tiles = (n_items + (size_t)KVG_BLOCK * ROWS - 1) / ((size_t)KVG_BLOCK * ROWS);
  size_t g = (size_t)ctx->sm_count * (size_t)occ;
  if (g > CLASSIFY_MAX_GRID) g = CLASSIFY_MAX_GRID;
  if (tiles < g) g = tiles;
  return g < 1 ? 1 : (int)g;
}
constexpr int PCI_ROWS = 4, PCI_STAGES = 4;    // 16 KiB stages, 64 KiB ring -> 3 CTAs / SM
constexpr int MDEV_ROWS = 2;
static int classify_variant() {  // KVG_CLASSIFY=tma selects the non-specialised kernel (A/B tests)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KVG_CLASSIFY");
    v = (e && !strcmp(e, "tma")) ? 1 : (e && !strcmp(e, "ws")) ? 0 : (e && !strcmp(e, "oneshot4")) ? 3
        : (e && !strcmp(e, "oneshot")) ? 2 : (e && !strcmp(e, "ragged")) ? 4 : 5;  // 5 = auto
  }
  return v;
}

extern "C" {

int kvg_abi_version(void) { return KVG_ABI_VERSION; }

int kvg_ctx_create(int cuda_device, kvg_ctx** out) {
  if (!out) return KVG_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_error = std::string("no usable CUDA device: ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return KVG_ECUDA;
  }
  if (cuda_device < 0 || cuda_device >= n) {
    g_create_error = "cuda_device out of range";
    return KVG_EINVAL;
  }
  kvg_ctx* ctx = new kvg_ctx();
  ctx->device = cuda_device;
  if ((e = cudaSetDevice(cuda_device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = std::string("cuda init: ") + cudaGetErrorString(e);
    delete ctx;
    return KVG_ECUDA;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, cuda_device);
  cudaFuncSetAttribute(k_pciids_parse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pciids_parse, KVG_BLOCK, P_SMEM);
  if (occ < 1) occ = 1;
  ctx->parse_grid = ctx->sm_count * occ;
  if (ensure(ctx, ctx->ctrl, 1) != KVG_OK || cudaMallocHost((void**)&ctx->h_ctrl, sizeof(ScanCtrl)) != cudaSuccess) {
    g_create_error = "control block allocation failed: " + ctx->err;
    kvg_ctx_destroy(ctx);
    return KVG_ENOMEM;
  }
  *out = ctx;
  return KVG_OK;
}

void kvg_ctx_destroy(kvg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->s_h2d) cudaStreamSynchronize(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  release(ctx->text); release(ctx->tables); release(ctx->info); release(ctx->tile_arrays);
  release(ctx->parse_state); release(ctx->v2_state); release(ctx->v2_pending); release(ctx->parse_ticket); release(ctx->pool); release(ctx->ctrl);
  release(ctx->nv_index); release(ctx->nv_lines); release(ctx->sec_lines); release(ctx->type_hash); release(ctx->ragged); release(ctx->tile_count); release(ctx->tile_off);
  release(ctx->tile_max); release(ctx->offs_state);
  release(ctx->recs); release(ctx->surv); release(ctx->classify_state); release(ctx->tile_hist); release(ctx->bin_total);
  for (OrderBufs* o : {&ctx->ord_dev, &ctx->ord_grp}) {
    release(o->p0); release(o->p1); release(o->perm); release(o->tile_heads); release(o->tile_off);
    release(o->seg_key); release(o->seg_off); release(o->seg_name); release(o->heads_state);
  }
  release(ctx->type_raw); release(ctx->type_label); release(ctx->type_off);
  release(ctx->type_label_len); release(ctx->type_match); release(ctx->type_name_len);
  release(ctx->type_canon); release(ctx->type_names); release(ctx->alive_prev);
  release(ctx->changed); release(ctx->flush); release(ctx->nv_ids); release(ctx->probe_slots);
  release(ctx->keys_blob); release(ctx->keys_off); release(ctx->match_off); release(ctx->match_len);
  release(ctx->match_out); release(ctx->gather_counts); release(ctx->local_surv);
  release(ctx->gather_base); release(ctx->p2p_err);
  for (int q = 0; q < P2P_MAX_RANKS; q++)
    if (ctx->p2p_peer[q] && ctx->p2p_peer[q] != ctx->p2p_mine) cudaIpcCloseMemHandle(ctx->p2p_peer[q]);
  if (ctx->p2p_mine) cudaFree(ctx->p2p_mine);
  for (auto& b : ctx->pinned_free) cudaFreeHost(b.p);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->h_pipe) cudaFreeHost(ctx->h_pipe);
  for (cudaEvent_t e : ctx->pipe_ev) cudaEventDestroy(e);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  for (auto& p : ctx->ev) {
    cudaEventDestroy(p.first);
    cudaEventDestroy(p.second);
  }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* kvg_last_error(kvg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
uint64_t kvg_launch_count(kvg_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* kvg_stream(kvg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void kvg_result_free(void* result) {
  if (!result) return;
  void* blk = (uint8_t*)result - 64;
  kvg_ctx* ctx = (kvg_ctx*)(uintptr_t)((uint64_t*)blk)[0];
  size_t cap = ((uint64_t*)blk)[1];
  // contexts outlive their results by contract (kvgpu.h); recycle the pinned block
  if (ctx->pinned_free.size() < 8)
    ctx->pinned_free.push_back({blk, cap});
  else
    cudaFreeHost(blk);
}

int kvg_set_kernel_timing(kvg_ctx* ctx, int enabled) {
  if (!ctx) return KVG_EINVAL;
  ctx->timing = enabled != 0;
  ctx->ev_used = 0;
  return KVG_OK;
}

int kvg_kernel_times(kvg_ctx* ctx, float* ms, char* names, size_t names_cap, int max_n) {
  if (!ctx) return KVG_EINVAL;
  CK(cudaStreamSynchronize(ctx->stream));
  int n = 0;
  size_t o = 0;
  for (size_t i = 0; i < ctx->ev_used && n < max_n; i++) {
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[i].first, ctx->ev[i].second);
    ms[n] = t;
    size_t l = ctx->ev_names[i].size() + 1;
    if (names && o + l <= names_cap) {
      memcpy(names + o, ctx->ev_names[i].c_str(), l);
      o += l;
    }
    n++;
  }
  ctx->ev_used = 0;
  return n;
}

// ================================================================================================
// pci.ids
// ================================================================================================
size_t kvg_text_pad(size_t len) { return ((len + P_TILE - 1) / P_TILE) * P_TILE + P_HALO; }

static uint32_t table_log2_for(size_t len) {
  // only the device lines under vendor 10de are inserted (~1 per 800 bytes of the shipped file); a
  // file that is denser than 1 per 96 bytes trips the overflow / crowding check and is re-parsed
  // with a table sized from its real entry count (parse_with_regrow)
  size_t want = len / 96 + 1024;
  uint32_t l = 10;
  while (((size_t)1 << l) < want) l++;
  return l;
}

// KVG_PARSE=v2: the barrier-free parse of kvg_parse_v2.cuh (experimental, see its header)
static bool parse_v2_enabled() {
  static const bool on = [] {
    const char* e = getenv("KVG_PARSE");
    return e && strcmp(e, "v2") == 0;
  }();
  return on;
}

static int parse_enqueue_v2(kvg_ctx* ctx, const uint8_t* d_text, size_t len, size_t stride, uint32_t n_files,
                            uint32_t cap_log2) {
  const uint32_t spf = (uint32_t)((len + V2_SPAN - 1) / V2_SPAN);
  const uint64_t n_spans64 = (uint64_t)spf * n_files;
  if (n_spans64 > 0x7fffffffull) {
    ctx->err = "too many spans";
    return KVG_EINVAL;
  }
  const uint32_t n_spans = (uint32_t)n_spans64;
  ctx->cap_log2 = cap_log2;
  const size_t cap = (size_t)1 << cap_log2;
  ENSURE(ctx->tables, cap * n_files);
  ENSURE(ctx->info, n_files);
  ENSURE(ctx->tile_arrays, 3 * (size_t)n_spans);
  ENSURE(ctx->v2_state, 2 * (size_t)n_spans);
  ENSURE(ctx->v2_pending, (size_t)n_spans * V2_PEND_CAP);
  ParseV2Args A;
  A.text = d_text;
  A.stride = stride;
  A.len = (uint32_t)len;
  A.n_files = n_files;
  A.spans_per_file = spf;
  A.n_spans = n_spans;
  A.tables = ctx->tables.p;
  A.cap_mask = (uint32_t)cap - 1;
  A.cap_shift = 32 - cap_log2;
  A.info = ctx->info.p;
  A.span_first_hdr = ctx->tile_arrays.p;
  A.span_first_nl = ctx->tile_arrays.p + n_spans;
  A.span_last_nl = ctx->tile_arrays.p + 2 * (size_t)n_spans;
  A.span_state = ctx->v2_state.p;
  A.pend_cnt = ctx->v2_state.p + n_spans;
  A.pending = ctx->v2_pending.p;
  CK(cudaMemsetAsync(ctx->tables.p, 0xff, cap * n_files * sizeof(uint64_t), ctx->stream));  // P_EMPTY
  CK(cudaMemsetAsync(ctx->info.p, 0, sizeof(PciIdsInfo) * n_files, ctx->stream));
  CK(cudaMemset2DAsync(ctx->info.p, sizeof(PciIdsInfo), 0xff, sizeof(uint32_t), n_files, ctx->stream));
  const unsigned grid = (n_spans + V2_WARPS - 1) / V2_WARPS;
  LAUNCH("pciids_parse", k_pciids_scan_v2, grid, V2_WARPS * 32, 0, A);
  LAUNCH("pciids_resolve", k_pciids_resolve_v2, grid, V2_WARPS * 32, 0, A);
  ParseArgs F;  // the finalize kernel reads the span summaries through the tile arrays
  memset(&F, 0, sizeof F);
  F.text = d_text;
  F.stride = stride;
  F.len = (uint32_t)len;
  F.n_files = n_files;
  F.tiles_per_file = spf;
  F.n_tiles = n_spans;
  F.info = ctx->info.p;
  F.tile_first_hdr = A.span_first_hdr;
  F.tile_first_nl = A.span_first_nl;
  F.tile_last_nl = A.span_last_nl;
  LAUNCH("pciids_finalize", k_pciids_finalize_v2, n_files, KVG_BLOCK, 0, F);
  ENSURE(ctx->nv_index, 65536);
  ENSURE(ctx->nv_lines, 65536 + 8);
  CK(cudaMemsetAsync(ctx->nv_lines.p + 65536, 0, sizeof(uint32_t), ctx->stream));
  LAUNCH("pciids_nv_index", k_nv_index, 256, 256, 0, ctx->tables.p, A.cap_mask, A.cap_shift, ctx->info.p,
         ctx->nv_index.p, ctx->nv_lines.p, ctx->nv_lines.p + 65536);
  return check_launch(ctx, "pciids parse (v2)");
}

static int parse_enqueue(kvg_ctx* ctx, const uint8_t* d_text, size_t len, size_t stride,
                         uint32_t n_files, uint32_t cap_log2) {
  if (len == 0 || len >= 0xfffffff0ull) {
    ctx->err = "pci.ids length out of range";
    return KVG_EINVAL;
  }
  if (parse_v2_enabled()) return parse_enqueue_v2(ctx, d_text, len, stride, n_files, cap_log2);
  uint32_t tpf = (uint32_t)((len + P_TILE - 1) / P_TILE);
  uint64_t n_tiles64 = (uint64_t)tpf * n_files;
  if (n_tiles64 > 0x7fffffffull) {
    ctx->err = "too many tiles";
    return KVG_EINVAL;
  }
  uint32_t n_tiles = (uint32_t)n_tiles64;
  ctx->cap_log2 = cap_log2;
  size_t cap = (size_t)1 << ctx->cap_log2;
  ENSURE(ctx->tables, cap * n_files);
  ENSURE(ctx->info, n_files);
  ENSURE(ctx->tile_arrays, 3 * (size_t)n_tiles);
  ENSURE(ctx->parse_state, n_tiles);

  ParseArgs A;
  A.text = d_text;
  A.stride = stride;
  A.len = (uint32_t)len;
  A.n_files = n_files;
  A.tiles_per_file = tpf;
  A.n_tiles = n_tiles;
  A.tables = ctx->tables.p;
  A.cap_mask = (uint32_t)cap - 1;
  A.cap_shift = 32 - ctx->cap_log2;
  A.info = ctx->info.p;
  A.tile_first_hdr = ctx->tile_arrays.p;
  A.tile_first_nl = ctx->tile_arrays.p + n_tiles;
  A.tile_last_nl = ctx->tile_arrays.p + 2 * (size_t)n_tiles;
  A.tile_state = ctx->parse_state.p;
  A.epoch = next_epoch();

  // table slots <- EMPTY, info <- {v_off = NONE, 0...}, ticket <- 0
  CK(cudaMemsetAsync(ctx->tables.p, 0xff, cap * n_files * sizeof(uint64_t), ctx->stream));  // P_EMPTY
  CK(cudaMemsetAsync(ctx->info.p, 0, sizeof(PciIdsInfo) * n_files, ctx->stream));
  CK(cudaMemset2DAsync(ctx->info.p, sizeof(PciIdsInfo), 0xff, sizeof(uint32_t), n_files, ctx->stream));
  int grid = ctx->parse_grid;
  if ((uint32_t)grid > n_tiles) grid = (int)n_tiles;
  LAUNCH("pciids_parse", k_pciids_parse, grid, KVG_BLOCK, P_SMEM, A);
  LAUNCH("pciids_finalize", k_pciids_finalize, n_files, KVG_BLOCK, 0, A);
  // flatten the table for vendor 10de: the scans' per-survivor join is then a single load
  ENSURE(ctx->nv_index, 65536);
  ENSURE(ctx->nv_lines, 65536 + 8);
  CK(cudaMemsetAsync(ctx->nv_lines.p + 65536, 0, sizeof(uint32_t), ctx->stream));
  LAUNCH("pciids_nv_index", k_nv_index, 256, 256, 0, ctx->tables.p, A.cap_mask, A.cap_shift, ctx->info.p,
         ctx->nv_index.p, ctx->nv_lines.p, ctx->nv_lines.p + 65536);
  return check_launch(ctx, "pciids parse");
}

// after the parse of image 0: sanitise the NVIDIA section into the pool and mirror it on the host
static int table_publish(kvg_ctx* ctx, const uint8_t* d_text, size_t len) {
  CK(cudaMemcpyAsync(&ctx->h_info, ctx->info.p, sizeof(PciIdsInfo), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->d_text = d_text;
  ctx->text_len = (uint32_t)len;
  ctx->h_pool.clear();
  if (ctx->h_info.v_off != P_NONE) {
    size_t sec = (size_t)ctx->h_info.sec_end - ctx->h_info.v_off;
    ENSURE(ctx->pool, sec + 16);
    CK(cudaMemsetAsync(ctx->pool.p, 0, sec + 16, ctx->stream));
    // one warp per named line (<= 65,536 lines; the shipped file has 1,931)
    LAUNCH("pciids_sanitise", k_pciids_sanitise_lines, 256, KVG_BLOCK, 0, d_text, (uint32_t)len, ctx->info.p,
           ctx->nv_lines.p, ctx->nv_lines.p + 65536, ctx->pool.p);
    int rc = check_launch(ctx, "pciids sanitise");
    if (rc) return rc;
    if (!ctx->h_pool.resize(sec + 16)) {
      ctx->err = "cudaMallocHost failed for the name pool mirror";
      return KVG_ENOMEM;
    }
    CK(cudaMemcpyAsync(ctx->h_pool.data(), ctx->pool.p, sec + 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  } else {
    ENSURE(ctx->pool, 16);
  }
  {  // candidate lines of the general (arbitrary-key) lookup: every '\t' line of the section
    size_t sec = ctx->h_info.v_off == P_NONE ? 0 : (size_t)ctx->h_info.sec_end - ctx->h_info.v_off;
    ctx->sec_lines_cap = sec / 2 + 8;  // a line needs >= 2 bytes ("\t\n")
    ENSURE(ctx->sec_lines, ctx->sec_lines_cap + 1);
    CK(cudaMemsetAsync(ctx->sec_lines.p + ctx->sec_lines_cap, 0, sizeof(uint32_t), ctx->stream));
    if (sec) {
      int grid = (int)((sec + KVG_BLOCK - 1) / KVG_BLOCK);
      if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
      LAUNCH("section_lines", k_section_lines, grid, KVG_BLOCK, 0, d_text, ctx->info.p, ctx->sec_lines.p,
             ctx->sec_lines.p + ctx->sec_lines_cap, (uint32_t)ctx->sec_lines_cap);
      int rc = check_launch(ctx, "section lines");
      if (rc) return rc;
    }
  }
  ctx->table_ready = true;
  return KVG_OK;
}

// The table is sized from the text length (pci.ids has ~1 device line per 77 bytes); an input with
// denser device lines overflows it, which K1 reports instead of spinning: re-parse with a table
// sized from the line count until the load factor is sane.
static int parse_with_regrow(kvg_ctx* ctx, const uint8_t* d_text, size_t len, size_t stride,
                             uint32_t n_files, bool enqueued = false, uint32_t cap_log2 = 0) {
  if (!enqueued) cap_log2 = table_log2_for(len);
  for (;;) {
    if (!enqueued) {
      int rc = parse_enqueue(ctx, d_text, len, stride, n_files, cap_log2);
      if (rc) return rc;
    }
    enqueued = false;
    int rc = table_publish(ctx, d_text, len);
    if (rc) return rc;
    size_t cap = (size_t)1 << cap_log2;
    bool crowded = (size_t)ctx->h_info.n_entries * 10 > cap * 5;
    if (!ctx->h_info.overflow && !crowded) return KVG_OK;
    ctx->table_ready = false;
    if (cap_log2 >= 30) {
      ctx->err = "pci.ids hash table cannot grow further";
      return KVG_ENOMEM;
    }
    size_t want = ctx->h_info.overflow ? cap * 4 : (size_t)ctx->h_info.n_entries * 4;
    while (((size_t)1 << cap_log2) < want) cap_log2++;
  }
}

// complete a pending kvg_pciids_load (see load_pending), then require a published table
static int table_needed(kvg_ctx* ctx, const char* why_missing) {
  if (ctx->load_pending) {
    ctx->load_pending = false;
    int rc = parse_with_regrow(ctx, ctx->text.p, ctx->pend_len, ctx->pend_stride, 1, true, ctx->pend_cap_log2);
    if (rc) return rc;
  }
  if (!ctx->table_ready) {
    ctx->err = why_missing;
    return KVG_ESTATE;
  }
  return KVG_OK;
}

int kvg_pciids_load(kvg_ctx* ctx, const uint8_t* text, size_t len) {
  if (!ctx || (!text && len)) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  ctx->table_ready = false;
  ctx->load_pending = false;
  if (len == 0) {  // an empty file: locateVendor fails, every lookup is "" (:382-385)
    ENSURE(ctx->info, 1);
    ENSURE(ctx->tables, 1024);
    ctx->cap_log2 = 10;
    LAUNCH("table_clear", k_fill64, 4, KVG_BLOCK, 0, ctx->tables.p, (size_t)1024, P_EMPTY);
    PciIdsInfo z;
    memset(&z, 0, sizeof z);
    z.v_off = z.sec_end = P_NONE;
    CK(cudaMemcpyAsync(ctx->info.p, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    ENSURE(ctx->text, 64);
    return table_publish(ctx, ctx->text.p, 0);
  }
  size_t padded = kvg_text_pad(len);
  ENSURE(ctx->text, padded);
  CK(cudaMemsetAsync(ctx->text.p, '\n', padded, ctx->stream));
  // the caller's buffer must stay readable until the copy has run: pinned memory is copied
  // asynchronously (the caller keeps it alive until the next consuming call, as with any cudaMemcpyAsync
  // source), pageable memory is copied synchronously by the runtime's own staging
  CK(cudaMemcpyAsync(ctx->text.p, text, len, cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t cap_log2 = table_log2_for(len);
  int rc = parse_enqueue(ctx, ctx->text.p, len, padded, 1, cap_log2);
  if (rc) return rc;
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) return parse_with_regrow(ctx, ctx->text.p, len, padded, 1, true, cap_log2);
  ctx->load_pending = true;
  ctx->pend_len = len;
  ctx->pend_stride = padded;
  ctx->pend_cap_log2 = cap_log2;
  return KVG_OK;
}

int kvg_dev_pciids_parse(kvg_ctx* ctx, const void* d_text, size_t len, size_t stride, uint32_t n_files) {
  if (!ctx || !d_text || n_files == 0 || ((uintptr_t)d_text & 15) || (stride & 15) ||
      (n_files > 1 && stride < kvg_text_pad(len)))
    return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  // first call on an image publishes (host mirror of the name pool, table sizing); later calls on
  // the same image stay fully asynchronous: parse + finalize + sanitise enqueued, no host sync
  if (ctx->load_pending) {
    int rc_t = table_needed(ctx, "");
    if (rc_t && rc_t != KVG_ESTATE) return rc_t;
  }
  if (!ctx->table_ready || ctx->d_text != d_text || ctx->text_len != len || ctx->parsed_files != n_files) {
    int rc0 = parse_with_regrow(ctx, (const uint8_t*)d_text, len, stride, n_files);
    if (rc0 == KVG_OK) ctx->parsed_files = n_files;
    return rc0;
  }
  int rc = parse_enqueue(ctx, (const uint8_t*)d_text, len, stride, n_files, ctx->cap_log2);
  if (rc) return rc;
  size_t sec = ctx->h_pool.size();
  if (sec > 16) {
    LAUNCH("pciids_sanitise", k_pciids_sanitise_lines, 256, KVG_BLOCK, 0, (const uint8_t*)d_text, (uint32_t)len,
           ctx->info.p, ctx->nv_lines.p, ctx->nv_lines.p + 65536, ctx->pool.p);
  }
  return check_launch(ctx, "pciids sanitise");
}

int kvg_pciids_info(kvg_ctx* ctx, uint32_t* vendor_off, uint32_t* section_end, uint32_t* n_entries,
                    uint32_t* n_lines) {
  if (!ctx) return KVG_EINVAL;
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load has not been called");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(&ctx->h_info, ctx->info.p, sizeof(PciIdsInfo), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (vendor_off) *vendor_off = ctx->h_info.v_off;
  if (section_end) *section_end = ctx->h_info.sec_end;
  if (n_entries) *n_entries = ctx->h_info.n_entries;
  if (n_lines) *n_lines = ctx->h_info.n_lines;
  return KVG_OK;
}

static bool canonical_key(const char* key, size_t keylen, uint32_t* val) {
  if (keylen != 4) return false;
  uint32_t v = 0;
  for (int i = 0; i < 4; i++) {
    char c = key[i];
    uint32_t d;
    if (c >= '0' && c <= '9') d = (uint32_t)(c - '0');
    else if (c >= 'a' && c <= 'f') d = (uint32_t)(c - 'a' + 10);
    else return false;
    v = v * 16 + d;
  }
  *val = v;
  return true;
}

// general path for n keys (arbitrary bytes); results land in ctx->match_len / match_out
static int lookup_general(kvg_ctx* ctx, const uint8_t* d_keys, const uint32_t* d_key_off,
                          uint32_t n_keys, uint32_t cap, uint8_t* d_out, uint32_t* d_out_len,
                          uint32_t* d_match) {
  if (n_keys == 0) return KVG_OK;
  {
    int grid = (int)((n_keys + KVG_BLOCK - 1) / KVG_BLOCK);
    LAUNCH("match_clear", k_fill32, grid, KVG_BLOCK, 0, d_match, (size_t)n_keys, P_NONE);
  }
  size_t sec = ctx->h_info.v_off == P_NONE ? 0 : (size_t)ctx->h_info.sec_end - ctx->h_info.v_off;
  if (sec) {
    int gx = (int)((ctx->sec_lines_cap + KVG_BLOCK - 1) / KVG_BLOCK);
    if (gx > 32) gx = 32;  // a few thousand candidate lines
    for (uint32_t k0 = 0; k0 < n_keys; k0 += 65535) {
      uint32_t nk = n_keys - k0 > 65535 ? 65535 : n_keys - k0;
      dim3 grid(gx, nk);
      LAUNCH("lookup_general", k_lookup_general, grid, KVG_BLOCK, 0, ctx->d_text, ctx->text_len,
             ctx->sec_lines.p, ctx->sec_lines.p + ctx->sec_lines_cap, d_keys, d_key_off + k0, d_match + k0);
    }
  }
  int grid = (int)((n_keys + 63) / 64);
  LAUNCH("sanitise_matches", k_sanitise_matches, grid, 64, 0, ctx->d_text, ctx->text_len, d_key_off,
         d_match, n_keys, d_out, cap, d_out_len);
  return check_launch(ctx, "lookup_general");
}

int kvg_name_lookup(kvg_ctx* ctx, const char* key, size_t keylen, char* out, size_t cap, size_t* outlen) {
  if (!ctx || (!key && keylen) || !outlen) return KVG_EINVAL;
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load has not been called");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  *outlen = 0;
  uint32_t v;
  if (canonical_key(key, keylen, &v)) {  // hash path
    ENSURE(ctx->probe_slots, 1);
    LAUNCH("probe_keys", k_probe_keys, 1, 32, 0, ctx->tables.p, (1u << ctx->cap_log2) - 1,
           32 - ctx->cap_log2, ctx->info.p, v, 1u, ctx->probe_slots.p);
    uint32_t slot;
    CK(cudaMemcpyAsync(&slot, ctx->probe_slots.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (slot == P_NONE) return KVG_OK;
    size_t n = ctx->h_pool[slot] | ((size_t)ctx->h_pool[slot + 1] << 8);
    if (n > cap) return KVG_ERANGE;
    memcpy(out, &ctx->h_pool[slot + 2], n);
    *outlen = n;
    return KVG_OK;
  }
  // general path: prefix semantics on the GPU
  uint32_t kcap = 4096;
  for (;;) {
    ENSURE(ctx->keys_blob, keylen + 16);
    ENSURE(ctx->keys_off, 2);
    ENSURE(ctx->match_off, 1);
    ENSURE(ctx->match_len, 1);
    ENSURE(ctx->match_out, kcap);
    uint32_t off[2] = {0, (uint32_t)keylen};
    if (keylen) CK(cudaMemcpyAsync(ctx->keys_blob.p, key, keylen, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->keys_off.p, off, sizeof off, cudaMemcpyHostToDevice, ctx->stream));
    int rc = lookup_general(ctx, ctx->keys_blob.p, ctx->keys_off.p, 1, kcap, ctx->match_out.p,
                            ctx->match_len.p, ctx->match_off.p);
    if (rc) return rc;
    uint32_t n;
    CK(cudaMemcpyAsync(&n, ctx->match_len.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n > kcap) {  // name longer than the scratch: retry with the scanner's line limit
      kcap = SCAN_TOKEN_MAX;
      continue;
    }
    if (n > cap) return KVG_ERANGE;
    if (n) CK(cudaMemcpy(out, ctx->match_out.p, n, cudaMemcpyDeviceToHost));
    *outlen = n;
    return KVG_OK;
  }
}

int kvg_name_table(kvg_ctx* ctx, uint32_t first, uint32_t count, uint32_t* out_off, uint8_t* out_bytes, size_t cap) {
  if (!ctx || !out_off || (!out_bytes && cap) || (uint64_t)first + count > 65536) return KVG_EINVAL;
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load has not been called");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  out_off[0] = 0;
  if (count == 0) return KVG_OK;
  ENSURE(ctx->probe_slots, count);
  LAUNCH("probe_keys", k_probe_keys, (count + 255) / 256, 256, 0, ctx->tables.p,
         (1u << ctx->cap_log2) - 1, 32 - ctx->cap_log2, ctx->info.p, first, count, ctx->probe_slots.p);
  std::vector<uint32_t> slots(count);
  CK(cudaMemcpyAsync(slots.data(), ctx->probe_slots.p, 4 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  size_t o = 0;
  for (uint32_t i = 0; i < count; i++) {
    if (slots[i] != P_NONE) {
      size_t n = ctx->h_pool[slots[i]] | ((size_t)ctx->h_pool[slots[i] + 1] << 8);
      if (o + n > cap) return KVG_ERANGE;
      memcpy(out_bytes + o, &ctx->h_pool[slots[i] + 2], n);
      o += n;
    }
    out_off[i + 1] = (uint32_t)o;
  }
  return KVG_OK;
}

// ================================================================================================
// scans
// ================================================================================================
}  // extern "C"
// k_compact needs a CO-RESIDENT grid (static round-robin tiles + look-back): cap it at
// occupancy x SMs of the instantiation, and at the tile count when the host knows it.
template <class Op>
static int compact_grid(kvg_ctx* ctx, size_t n_items) {
  static int occ = 0;
  if (!occ) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_compact<Op>, KVG_BLOCK, 0);
    if (occ < 1) occ = 1;
  }
  size_t tiles = (n_items + C_TILE - 1) / C_TILE;
  size_t g = (size_t)ctx->sm_count * (size_t)occ;
  if (tiles < g) g = tiles;
  return g < 1 ? 1 : (int)g;
}
extern "C" {

static int ensure_order(kvg_ctx* ctx, OrderBufs& o, size_t cap) {
  const size_t T = (cap + C_TILE - 1) / C_TILE + 1;
  ENSURE(o.p0, cap + 1); ENSURE(o.p1, cap + 1); ENSURE(o.perm, cap + 1);
  ENSURE(o.seg_key, cap + 1); ENSURE(o.seg_off, cap + 2); ENSURE(o.seg_name, cap + 1);
  ENSURE(o.tile_heads, T + 1); ENSURE(o.tile_off, T + 2);
  ENSURE(o.heads_state, T / C_TILE + 2);
  return KVG_OK;
}

}  // extern "C"
// Both stable orderings of the survivor list — ordering 0 by device id / mdev type (<= 16 bits,
// 2 passes), ordering 1 by iommu group / parent (32 bits, up to 4 passes) — share their launches:
// grid.y = 2 while both have a pass, grid.y = 1 (ordering 1 only, passed in slot 0) afterwards.
static int enqueue_orderings(kvg_ctx* ctx, size_t cap, int src0, int src1, bool owned_only = false) {
  int rc = ensure_order(ctx, ctx->ord_dev, cap);
  if (rc) return rc;
  rc = ensure_order(ctx, ctx->ord_grp, cap);
  if (rc) return rc;
  size_t T = (cap + C_TILE - 1) / C_TILE;
  if (T == 0) T = 1;
  ENSURE(ctx->tile_hist, 2 * (size_t)RADIX_MAX_DIGITS * T);
  ENSURE(ctx->bin_total, 2 * (size_t)RADIX_MAX_DIGITS);
  ScanCtrl* c = ctx->ctrl.p;
  OrderBufs* ob[2] = {&ctx->ord_dev, &ctx->ord_grp};
  uint32_t* maxk[2] = {&c->max_devkey, &c->max_group};
  // element count per ordering: all survivors, or (sharded) the survivors whose key this rank owns
  uint32_t* cnt[2] = {owned_only ? &c->n_own[0] : &c->n_surv, owned_only ? &c->n_own[1] : &c->n_surv};
  if (owned_only && cap) {
    // select the owned {key, index} pairs of each ordering into p1 (pass 0 then reads pairs)
    constexpr int TT = 128, RR = 8;
    const size_t tiles = (cap + (size_t)TT * RR - 1) / ((size_t)TT * RR);
    ENSURE(ctx->ragged, tiles * TT * RR);  // as uint2 this needs half of it
    ENSURE(ctx->tile_count, tiles + 1);
    ENSURE(ctx->tile_off, tiles + 2);
    ENSURE(ctx->tile_max, tiles + 1);
    const unsigned chunks = (unsigned)((tiles + C_TILE - 1) / C_TILE);
    ENSURE(ctx->offs_state, (size_t)chunks + 1);
    for (int ord = 0; ord < 2; ord++) {
      OwnedPairOp op;
      op.surv = ctx->surv.p;
      op.n_ptr = &c->n_surv;  // launches are sized for `cap`; tiles past the count retire at once
      op.out = (uint2*)ctx->ragged.p;
      op.field = (uint32_t)ord;
      op.nranks = (uint32_t)ctx->nranks;
      op.rank = (uint32_t)ctx->rank;
      op.local_max = 0;
      LAUNCH("own_select", (k_classify_ragged<OwnedPairOp, TT, RR>), (unsigned)tiles, TT, 0, op, ctx->tile_count.p,
             ctx->tile_max.p);
      TileOffsetsArgs2 tt;
      tt.o[0] = {ctx->tile_count.p, ctx->tile_max.p, nullptr, (uint32_t)tiles, ctx->tile_off.p, cnt[ord],
                 ctx->offs_state.p};
      tt.o[1] = tt.o[0];
      LAUNCH("tile_offsets", k_tile_offsets, chunks, KVG_BLOCK, 0, tt, c, next_epoch());
      LAUNCH("own_pack", k_pack_pairs, (unsigned)tiles, 128, 0, (const uint2*)ctx->ragged.p, ctx->tile_off.p,
             (uint32_t)(TT * RR), ob[ord]->p1.p);
    }
  }
  const uint32_t key_bits[2] = {16, 32};  // device id / mdev type: u16; iommu group / parent: u32
  const int srcs[2] = {src0, src1};
  // digit width: up to 11 bits where the scan is latency-bound (fewer passes: 19-bit groups sort in 2),
  // 8 bits for inputs large enough to be bandwidth-bound (half the shared memory, twice the CTAs/SM)
  const uint32_t max_bits = cap >= (8u << 20) ? 8 : RADIX_MAX_BITS;
  const int nsets[2] = {(int)((key_bits[0] + max_bits - 1) / max_bits),
                        (int)((key_bits[1] + max_bits - 1) / max_bits)};  // 2 and 3 (or 4) launch sets
  auto fill = [&](int ord, int p) {
    RadixArgs a;
    a.n_ptr = cnt[ord];
    a.max_key = maxk[ord];
    a.src_records = ctx->surv.p;
    a.pairs_in = (p == 0 && !owned_only) ? nullptr : ((p & 1) ? ob[ord]->p0.p : ob[ord]->p1.p);
    a.pairs_out = (p & 1) ? ob[ord]->p1.p : ob[ord]->p0.p;
    a.tile_hist = ctx->tile_hist.p + (size_t)ord * RADIX_MAX_DIGITS * T;
    a.bin_total = ctx->bin_total.p + (size_t)ord * RADIX_MAX_DIGITS;
    a.pass = p < nsets[ord] ? (uint32_t)p : 0xffu;
    a.key_bits_max = key_bits[ord];
    a.max_bits = max_bits;
    a.src = (p == 0 && !owned_only) ? srcs[ord] : SRC_PAIRS;
    return a;
  };
  if (!ctx->scatter_smem_set) {  // a function attribute is per device: remember it per context
    CK(cudaFuncSetAttribute(k_radix_scatter<RADIX_MAX_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)RadixScatterCfg<RADIX_MAX_BITS>::SMEM));
    ctx->scatter_smem_set = true;
  }
  for (int p = 0; p < nsets[1]; p++) {
    RadixArgs2 aa;
    // sets 0/1 always have work: one CTA per tile.  Set 2 exists only for keys wider than 22 bits and is
    // usually ruled out on the device: a small persistent grid makes a ruled-out pass nearly free.
    const bool both = p < nsets[0];
    const unsigned gx = p < 2 ? (unsigned)T : (unsigned)std::min<size_t>(T, (size_t)ctx->sm_count * (max_bits == 8 ? 5 : 3));
    dim3 grid(gx, both ? 2 : 1);
    if (both) {
      aa.o[0] = fill(0, p);
      aa.o[1] = fill(1, p);
    } else {
      aa.o[0] = fill(1, p);
      aa.o[1] = aa.o[0];
    }
    dim3 sgrid(KVG_BLOCK, grid.y);
    LAUNCH("radix_hist", k_radix_hist, grid, KVG_BLOCK, 0, aa);
    static const bool tilescan_warp = [] {  // KVG_TILESCAN=warp: experimental, see kvg_radix_exp.cuh
      const char* e = getenv("KVG_TILESCAN");
      return e && strcmp(e, "warp") == 0;
    }();
    if (tilescan_warp) {
      dim3 wgrid(RADIX_MAX_DIGITS / TS_WARPS, grid.y);
      LAUNCH("radix_tilescan", k_radix_tilescan_warp, wgrid, TS_WARPS * 32, 0, aa);
    } else {
      LAUNCH("radix_tilescan", k_radix_tilescan, sgrid, KVG_BLOCK, 0, aa);
    }
    static const bool scatter_c = [] {  // KVG_SCATTER=c: experimental, see kvg_radix_exp.cuh
      const char* e = getenv("KVG_SCATTER");
      return e && strcmp(e, "c") == 0;
    }();
    if (max_bits == 8) {
      LAUNCH("radix_scatter", k_radix_scatter<8>, grid, KVG_BLOCK, RadixScatterCfg<8>::SMEM, aa);
    } else if (scatter_c) {
      if (!ctx->scatter_c_smem_set) {
        CK(cudaFuncSetAttribute(k_radix_scatter_c, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)RadixScatterCfg<RADIX_MAX_BITS>::SMEM));
        ctx->scatter_c_smem_set = true;
      }
      LAUNCH("radix_scatter", k_radix_scatter_c, grid, KVG_BLOCK, RadixScatterCfg<RADIX_MAX_BITS>::SMEM, aa);
    } else {
      LAUNCH("radix_scatter", k_radix_scatter<RADIX_MAX_BITS>, grid, KVG_BLOCK, RadixScatterCfg<RADIX_MAX_BITS>::SMEM, aa);
    }
  }
  // final permutation + distinct keys of both orderings: count heads per tile, scan, emit
  OrderFinalArgs2 ff;
  TileOffsetsArgs2 tt;
  uint32_t* nseg[2] = {&c->n_dev_keys, &c->n_groups};
  for (int ord = 0; ord < 2; ord++) {
    OrderFinalArgs& a = ff.o[ord];
    a.p0 = ob[ord]->p0.p;
    a.p1 = ob[ord]->p1.p;
    a.max_key = maxk[ord];
    a.key_bits_max = key_bits[ord];
    a.max_bits = max_bits;
    a.n_ptr = cnt[ord];
    a.perm = ob[ord]->perm.p;
    a.tile_heads = ob[ord]->tile_heads.p;
    a.tile_off = ob[ord]->tile_off.p;
    a.seg_key = ob[ord]->seg_key.p;
    a.seg_off = ob[ord]->seg_off.p;
    a.n_seg = nseg[ord];
    // PCI device-id ordering: the bucket's joined name slot comes back with the keys
    a.head_surv = (ord == 0 && src0 == SRC_PCI_DEVICE) ? ctx->surv.p : nullptr;
    a.head_name = a.head_surv ? ob[ord]->seg_name.p : nullptr;
    TileOffsetsArgs& t = tt.o[ord];
    t.tile_count = ob[ord]->tile_heads.p;
    t.tile_max = nullptr;
    t.n_items_ptr = cnt[ord];
    t.n_tiles_host = 0;
    t.tile_off = ob[ord]->tile_off.p;
    t.total_out = nseg[ord];
    t.state = ob[ord]->heads_state.p;
  }
  dim3 fgrid((unsigned)T, 2);
  LAUNCH("order_count", k_order_final<false>, fgrid, KVG_BLOCK, 0, ff);
  dim3 ogrid((unsigned)((T + C_TILE - 1) / C_TILE), 2);
  LAUNCH("tile_offsets", k_tile_offsets, ogrid, KVG_BLOCK, 0, tt, c, next_epoch());
  LAUNCH("order_emit", k_order_final<true>, fgrid, KVG_BLOCK, 0, ff);
  return check_launch(ctx, "orderings");
}

static int enqueue_pci_orderings(kvg_ctx* ctx, size_t surv_cap, bool owned_only = false) {
  return enqueue_orderings(ctx, surv_cap, SRC_PCI_DEVICE, SRC_PCI_GROUP, owned_only);
}

static int enqueue_classify(kvg_ctx* ctx, const void* d_recs, size_t n, uint4* d_out) {
  const size_t pci_tiles = (n + (size_t)KVG_BLOCK * PCI_ROWS - 1) / ((size_t)KVG_BLOCK * PCI_ROWS) + 1;
  ENSURE(ctx->classify_state, 2 * pci_tiles + n / 512 + 2);  // aggregates + round prefixes / per-tile states
  PciClassifyOp op;
  op.recs = (const uint4*)d_recs;
  op.n = (uint32_t)n;
  op.out = (kvg_pci_surv*)d_out;
  op.ctrl = ctx->ctrl.p;
  op.table = ctx->tables.p;
  op.cap_mask = (1u << ctx->cap_log2) - 1;
  op.cap_shift = 32 - ctx->cap_log2;
  op.info = ctx->info.p;
  op.nv_index = ctx->nv_index.p;
  op.local_max_group = 0;
  op.local_max_dev = 0;
  size_t smem = 0;
  // auto: below ~2 M records the step is launch-bound and ONE look-back kernel beats the three
  // launches of the split form; above, the split form (no cross-CTA wait) runs at the HBM roofline
  int variant = classify_variant();
  if (variant == 5) variant = n < (2u << 20) ? 2 : 4;
  if (variant == 4) {
    constexpr int T = 128, R = 8;
    const size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
    if (tiles == 0) {  // nothing to classify: n_surv stays 0 from the control-block memset
      return KVG_OK;
    }
    ENSURE(ctx->ragged, tiles * T * R);
    ENSURE(ctx->tile_count, tiles + 1);
    ENSURE(ctx->tile_off, tiles + 2);
    ENSURE(ctx->tile_max, tiles + 1);
    const unsigned chunks = (unsigned)((tiles + C_TILE - 1) / C_TILE);
    ENSURE(ctx->offs_state, (size_t)chunks + 1);
    op.out = (kvg_pci_surv*)ctx->ragged.p;
    LAUNCH("classify_compact", (k_classify_ragged<PciClassifyOp, T, R>), (unsigned)tiles, T, 0, op,
           ctx->tile_count.p, ctx->tile_max.p);
    {
      TileOffsetsArgs2 tt;
      tt.o[0] = {ctx->tile_count.p, ctx->tile_max.p, nullptr, (uint32_t)tiles, ctx->tile_off.p,
                 &ctx->ctrl.p->n_surv, ctx->offs_state.p};
      tt.o[1] = tt.o[0];
      LAUNCH("tile_offsets", k_tile_offsets, chunks, KVG_BLOCK, 0, tt, ctx->ctrl.p, next_epoch());
    }
    LAUNCH("pack_survivors", k_pack_survivors<1>, (unsigned)tiles, 128, 0, ctx->ragged.p, ctx->tile_off.p,
           (uint32_t)(T * R), d_out);
  } else if (variant >= 2) {  // one tile per CTA with look-back
    if (variant == 2) {
      constexpr int T = 128, R = 8;
      size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
      LAUNCH("classify_compact", (k_classify_oneshot<PciClassifyOp, T, R>), (unsigned)(tiles ? tiles : 1), T, 0, op,
             ctx->classify_state.p, next_epoch());
    } else {
      constexpr int T = 128, R = 4;
      size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
      LAUNCH("classify_compact", (k_classify_oneshot<PciClassifyOp, T, R>), (unsigned)(tiles ? tiles : 1), T, 0, op,
             ctx->classify_state.p, next_epoch());
    }
  } else if (variant == 1) {
    int grid = classify_grid<PciClassifyOp, PCI_ROWS, PCI_STAGES>(ctx, n, &smem);
    LAUNCH("classify_compact", (k_classify_tma<PciClassifyOp, PCI_ROWS, PCI_STAGES>), grid, KVG_BLOCK, smem, op,
           ctx->classify_state.p, ctx->classify_state.p + pci_tiles, next_epoch());
  } else {
    int grid = classify_ws_grid<PciClassifyOp, PCI_ROWS, PCI_STAGES>(ctx, n, &smem);
    LAUNCH("classify_compact", (k_classify_ws<PciClassifyOp, PCI_ROWS, PCI_STAGES>), grid, WS_THREADS, smem, op,
           ctx->classify_state.p, ctx->classify_state.p + pci_tiles, next_epoch());
  }
  return check_launch(ctx, "classify");
}

extern "C" {

int kvg_dev_scan_pci(kvg_ctx* ctx, const void* d_recs, size_t n) {
  if (!ctx || (!d_recs && n) || n > 0xfffffff0ull || ((uintptr_t)d_recs & 15)) return KVG_EINVAL;
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load must precede a scan (the scan joins names)");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  ENSURE(ctx->surv, n + 1);
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  int rc = enqueue_classify(ctx, d_recs, n, ctx->surv.p);
  if (rc) return rc;
  rc = enqueue_pci_orderings(ctx, n);
  if (rc) return rc;
  ctx->last_n = n;
  ctx->last_total = n;
  ctx->last_kind = 1;
  ctx->last_owned = false;
  return KVG_OK;
}

int kvg_dev_scan_pci_count(kvg_ctx* ctx, uint64_t* n_survivors, uint32_t* n_dev_keys, uint32_t* n_groups) {
  if (!ctx || ctx->last_kind == 0) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (n_survivors) *n_survivors = ctx->h_ctrl->n_surv;
  if (n_dev_keys) *n_dev_keys = ctx->h_ctrl->n_dev_keys;
  if (n_groups) *n_groups = ctx->h_ctrl->n_groups;
  return KVG_OK;
}

static size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }

// result-block layout: [header][survivors, `surv_reserve` slots][orderings...][name pool]
static size_t pci_block_bytes(size_t surv_reserve, size_t S, size_t KD, size_t G, size_t SD, size_t SG,
                              size_t pool_len) {
  auto a64 = [](size_t x) { return (x + 63) & ~(size_t)63; };
  return a64(sizeof(kvg_pci_result)) + a64(surv_reserve * 16) + a64(KD * 4) + a64(KD * 2) + a64((KD + 1) * 4) +
         a64(SD * 4) + a64(KD * 4) + a64(G * 4) + a64((G + 1) * 4) + a64(SG * 4) + a64(pool_len);
}

// blk == NULL: allocate for the exact counts and copy the survivors too.  Otherwise `blk` was sized
// for the worst case by the pipelined entry point, which has already copied the survivors into it.
static int fetch_pci(kvg_ctx* ctx, kvg_pci_result** res, void* blk, size_t surv_reserve) {
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRACE("ctrl");
  const bool surv_done = blk != nullptr;
  const size_t S = ctx->h_ctrl->n_surv, KD = ctx->h_ctrl->n_dev_keys, G = ctx->h_ctrl->n_groups;
  if (!surv_done) surv_reserve = S;
  // members covered by each ordering: all survivors, or (sharded) those whose key this rank owns
  const size_t SD = ctx->last_owned ? ctx->h_ctrl->n_own[0] : S;
  const size_t SG = ctx->last_owned ? ctx->h_ctrl->n_own[1] : S;
  const size_t pool_len = ctx->h_pool.size();
  size_t o = align64(sizeof(kvg_pci_result));  // header first
  size_t o_surv = o; o += align64(surv_reserve * 16);
  size_t o_dkeys32 = o; o += align64(KD * 4);
  size_t o_dkeys = o; o += align64(KD * 2);
  size_t o_doff = o; o += align64((KD + 1) * 4);
  size_t o_dperm = o; o += align64(SD * 4);
  size_t o_dname = o; o += align64(KD * 4);
  size_t o_gkeys = o; o += align64(G * 4);
  size_t o_goff = o; o += align64((G + 1) * 4);
  size_t o_gperm = o; o += align64(SG * 4);
  size_t o_pool = o; o += align64(pool_len);
  if (!blk) blk = pinned_alloc(ctx, o);
  if (!blk) {
    ctx->err = "cudaMallocHost failed for the result block";
    return KVG_ENOMEM;
  }
  uint8_t* b = pinned_payload(blk);
  auto D2H = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    return cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  OrderBufs& od = ctx->ord_dev;
  OrderBufs& og = ctx->ord_grp;
  // every kernel has completed (the control block was just read): the group ordering's arrays may
  // ride the second copy stream next to the device ordering's
  cudaStream_t s2 = ctx->s_d2h ? ctx->s_d2h : ctx->stream;
  auto D2H2 = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    return cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyDeviceToHost, s2);
  };
  if (!surv_done) CK(D2H(o_surv, ctx->surv.p, S * 16));
  CK(D2H2(o_gperm, og.perm.p, SG * 4));
  CK(D2H(o_dperm, od.perm.p, SD * 4));
  CK(D2H2(o_goff, og.seg_off.p, (G + 1) * 4));
  CK(D2H(o_dkeys32, od.seg_key.p, KD * 4));
  CK(D2H2(o_gkeys, og.seg_key.p, G * 4));
  CK(D2H(o_doff, od.seg_off.p, (KD + 1) * 4));
  CK(D2H(o_dname, od.seg_name.p, KD * 4));
  CK(cudaStreamSynchronize(ctx->stream));
  if (s2 != ctx->stream) CK(cudaStreamSynchronize(s2));
  TRACE("d2h");
  kvg_pci_result* r = (kvg_pci_result*)b;
  memset(r, 0, sizeof *r);
  r->n_records = ctx->last_n;
  r->n_survivors = S;
  r->survivors = (const kvg_pci_surv*)(b + o_surv);
  r->n_dev_keys = (uint32_t)KD;
  uint16_t* dk = (uint16_t*)(b + o_dkeys);
  const uint32_t* dk32 = (const uint32_t*)(b + o_dkeys32);
  uint32_t* dname = (uint32_t*)(b + o_dname);
  const uint32_t* doff = (const uint32_t*)(b + o_doff);
  const uint32_t* dperm = (const uint32_t*)(b + o_dperm);
  if (SD == 0) ((uint32_t*)(b + o_doff))[0] = 0;
  if (SG == 0) ((uint32_t*)(b + o_goff))[0] = 0;
  for (size_t k = 0; k < KD; k++) dk[k] = (uint16_t)dk32[k];  // marshalling only: narrow the keys
  r->dev_keys = dk;
  r->dev_off = doff;
  r->dev_perm = dperm;
  r->dev_name_slot = dname;
  r->n_groups = (uint32_t)G;
  r->grp_keys = (const uint32_t*)(b + o_gkeys);
  r->grp_off = (const uint32_t*)(b + o_goff);
  r->grp_perm = (const uint32_t*)(b + o_gperm);
  if (pool_len) memcpy(b + o_pool, ctx->h_pool.data(), pool_len);
  r->name_pool = b + o_pool;
  r->name_pool_len = pool_len;
  *res = r;
  TRACE("marshalled");
  return KVG_OK;
}

int kvg_dev_scan_pci_fetch(kvg_ctx* ctx, kvg_pci_result** res) {
  if (!ctx || !res || ctx->last_kind != 1) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  return fetch_pci(ctx, res, nullptr, 0);
}

static int stage_h2d(kvg_ctx* ctx, const void* host, size_t bytes, void* dev) {
  if (!bytes) return KVG_OK;
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) {  // the caller's memory may move or vanish after return (cgo rule): stage it
    if (ctx->h_stage_cap < bytes) {
      if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
      ctx->h_stage = nullptr;
      ctx->h_stage_cap = 0;
      size_t cap = bytes + bytes / 4 + 4096;
      CK(cudaMallocHost(&ctx->h_stage, cap));
      ctx->h_stage_cap = cap;
    }
    memcpy(ctx->h_stage, host, bytes);
    host = ctx->h_stage;
  }
  CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return KVG_OK;
}

// Host entry point, pipelined: the snapshot crosses PCIe in chunks on a copy stream while the previous
// chunk is classified and packed (tile-independent k_classify_ragged; the offsets kernel re-scans the
// tile counts seen so far, a few thousand words); each chunk's survivors start their way back on a
// second copy stream as soon as their count is known (a 4-byte store to mapped host memory), so the
// device-to-host copy of the survivor list runs under the remaining host-to-device traffic and under
// the ordering kernels.  Only the orderings' own arrays are copied after the last kernel.
static const size_t PIPE_MIN_RECORDS = 128u << 10, PIPE_MAX_RECORDS = 16u << 20, PIPE_MAX_CHUNKS = 16;

__global__ void k_publish_count(const uint32_t* __restrict__ src, volatile uint32_t* host_dst) {
  pdl_enter();
  *host_dst = *src;
  __threadfence_system();
}

static int scan_pci_pipelined(kvg_ctx* ctx, const kvg_pci_rec* recs, size_t n, kvg_pci_result** res) {
  constexpr int T = 128, R = 8;
  constexpr size_t TILE = (size_t)T * R;
  PhaseTrace trace;
  g_trace = trace.on ? &trace : nullptr;
  struct Unset { ~Unset() { g_trace = nullptr; } } unset_;
  if (!ctx->s_h2d) {
    CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    CK(cudaMallocHost((void**)&ctx->h_pipe, sizeof(uint32_t) * PIPE_MAX_CHUNKS));
    ctx->pipe_ev.resize(2 * PIPE_MAX_CHUNKS + 1);
    for (auto& e : ctx->pipe_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  // ~512 K records (8 MiB) per chunk, whole tiles: per-chunk stream/event overheads outweigh finer overlap
  static const size_t chunk_target = [] {
    const char* e = getenv("KVG_PIPE_CHUNK_K");  // records per chunk in Ki (A/B knob)
    long v = e ? atol(e) : 0;
    return (size_t)(v > 0 ? v : 512) << 10;
  }();
  size_t n_chunks = std::min(PIPE_MAX_CHUNKS, std::max((size_t)2, (n + chunk_target - 1) / chunk_target));
  size_t chunk = ((n + n_chunks - 1) / n_chunks + TILE - 1) / TILE * TILE;
  n_chunks = (n + chunk - 1) / chunk;
  const size_t tiles = (n + TILE - 1) / TILE;
  const uint4* recs_before = ctx->recs.p;
  ENSURE(ctx->recs, n + 1);
  ENSURE(ctx->surv, n + 1);
  ENSURE(ctx->ragged, tiles * TILE);
  ENSURE(ctx->tile_count, tiles + 1);
  ENSURE(ctx->tile_off, tiles + 2);
  ENSURE(ctx->tile_max, tiles + 1);
  ENSURE(ctx->offs_state, (tiles + C_TILE - 1) / C_TILE + 1);
  // host side of the copies
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, recs) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const uint8_t* src = (const uint8_t*)recs;
  if (!pinned && ctx->h_stage_cap < n * 16) {  // cgo rule: never keep the caller's pointer -> stage it
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_cap = 0;
    size_t cap = n * 16 + n * 4 + 4096;
    CK(cudaMallocHost(&ctx->h_stage, cap));
    ctx->h_stage_cap = cap;
  }
  // ctx->recs is only ever touched by host entry points, which synchronise before they return, so the
  // copy stream may start at once — it must NOT wait behind a pending kvg_pciids_load's parse kernels.
  // Exception: a fresh allocation is zero-filled on the context stream.
  if (ctx->recs.p != recs_before) CK(cudaStreamSynchronize(ctx->stream));
  for (size_t k = 0; k < n_chunks; k++) {
    const size_t c0 = k * chunk, cn = std::min(chunk, n - c0);
    const uint8_t* from = src + c0 * 16;
    if (!pinned) {
      memcpy((uint8_t*)ctx->h_stage + c0 * 16, from, cn * 16);
      from = (const uint8_t*)ctx->h_stage + c0 * 16;
    }
    CK(cudaMemcpyAsync(ctx->recs.p + c0, from, cn * 16, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaEventRecord(ctx->pipe_ev[k], ctx->s_h2d));
  }
  TRACE("h2d_issued");
  // the table (a pending kvg_pciids_load is completed here, under the copies already in flight)
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load must precede a scan (the scan joins names)");
    if (rc_t) {
      cudaStreamSynchronize(ctx->s_h2d);
      return rc_t;
    }
  }
  TRACE("table");
  const size_t pool_len = ctx->h_pool.size();
  void* blk = pinned_alloc(ctx, pci_block_bytes(n, n, std::min<size_t>(n, 65536), n, n, n, pool_len));
  if (!blk) {
    cudaStreamSynchronize(ctx->s_h2d);
    ctx->err = "cudaMallocHost failed for the result block";
    return KVG_ENOMEM;
  }
  uint8_t* b = pinned_payload(blk);
  const size_t o_surv = (sizeof(kvg_pci_result) + 63) & ~(size_t)63;
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  PciClassifyOp op;
  op.ctrl = ctx->ctrl.p;
  op.table = ctx->tables.p;
  op.cap_mask = (1u << ctx->cap_log2) - 1;
  op.cap_shift = 32 - ctx->cap_log2;
  op.info = ctx->info.p;
  op.nv_index = ctx->nv_index.p;
  op.local_max_group = 0;
  op.local_max_dev = 0;
  for (size_t k = 0; k < n_chunks; k++) {
    const size_t c0 = k * chunk, cn = std::min(chunk, n - c0);
    const size_t t0 = c0 / TILE, tn = (cn + TILE - 1) / TILE, t1 = t0 + tn;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[k], 0));
    op.recs = ctx->recs.p + c0;
    op.n = (uint32_t)cn;
    op.out = (kvg_pci_surv*)(ctx->ragged.p + c0);
    LAUNCH("classify_compact", (k_classify_ragged<PciClassifyOp, T, R>), (unsigned)tn, T, 0, op,
           ctx->tile_count.p + t0, ctx->tile_max.p + t0);
    TileOffsetsArgs2 tt;
    tt.o[0] = {ctx->tile_count.p, ctx->tile_max.p, nullptr, (uint32_t)t1, ctx->tile_off.p, &ctx->ctrl.p->n_surv,
               ctx->offs_state.p};
    tt.o[1] = tt.o[0];
    LAUNCH("tile_offsets", k_tile_offsets, (unsigned)((t1 + C_TILE - 1) / C_TILE), KVG_BLOCK, 0, tt, ctx->ctrl.p,
           next_epoch());
    LAUNCH("pack_survivors", k_pack_survivors<1>, (unsigned)tn, 128, 0, (const uint4*)(ctx->ragged.p + c0),
           (const uint32_t*)(ctx->tile_off.p + t0), (uint32_t)TILE, ctx->surv.p);
    LAUNCH("publish_count", k_publish_count, 1, 32, 0, (const uint32_t*)&ctx->ctrl.p->n_surv,
           (volatile uint32_t*)(ctx->h_pipe + k));
    CK(cudaEventRecord(ctx->pipe_ev[PIPE_MAX_CHUNKS + k], ctx->stream));
  }
  int rc = check_launch(ctx, "classify");
  if (rc == KVG_OK) rc = enqueue_pci_orderings(ctx, n);
  if (rc) {
    cudaStreamSynchronize(ctx->stream);
    ctx->pinned_free.push_back({blk, (size_t)((uint64_t*)blk)[1]});
    return rc;
  }
  ctx->last_n = n;
  ctx->last_total = n;
  ctx->last_kind = 1;
  ctx->last_owned = false;
  TRACE("enqueued");
  // survivors go home chunk by chunk while later chunks and the orderings still run
  size_t prev = 0;
  for (size_t k = 0; k < n_chunks; k++) {
    CK(cudaEventSynchronize(ctx->pipe_ev[PIPE_MAX_CHUNKS + k]));
    const size_t cum = ctx->h_pipe[k];
    if (cum > prev)
      CK(cudaMemcpyAsync(b + o_surv + prev * 16, ctx->surv.p + prev, (cum - prev) * 16, cudaMemcpyDeviceToHost,
                         ctx->s_d2h));
    prev = cum;
    TRACE("chunk");
  }
  return fetch_pci(ctx, res, blk, n);
}

int kvg_scan_pci(kvg_ctx* ctx, const kvg_pci_rec* recs, size_t n, kvg_pci_result** res) {
  if (!ctx || !res || (!recs && n)) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  static const bool pipe = [] {
    const char* e = getenv("KVG_PIPELINE");
    return !(e && e[0] == '0');
  }();
  if (pipe && !ctx->timing && n >= PIPE_MIN_RECORDS && n <= PIPE_MAX_RECORDS)
    return scan_pci_pipelined(ctx, recs, n, res);
  ENSURE(ctx->recs, n + 1);
  int rc = stage_h2d(ctx, recs, n * sizeof(kvg_pci_rec), ctx->recs.p);
  if (rc) return rc;
  rc = kvg_dev_scan_pci(ctx, ctx->recs.p, n);
  if (rc) return rc;
  return kvg_dev_scan_pci_fetch(ctx, res);
}

// ---- health -------------------------------------------------------------------------------------
int kvg_health_reset(kvg_ctx* ctx) {
  if (!ctx) return KVG_EINVAL;
  ctx->health_n = 0;
  return KVG_OK;
}

int kvg_health_rescan(kvg_ctx* ctx, const kvg_pci_rec* recs, size_t n, kvg_health_delta** delta) {
  if (!ctx || !delta || (!recs && n) || n > 0x7fffffffull) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  ENSURE(ctx->recs, n + 1);
  ENSURE(ctx->changed, n + 1);
  ENSURE(ctx->classify_state, (n + C_TILE - 1) / C_TILE + 1);
  if (ctx->health_n != n) {
    ENSURE(ctx->alive_prev, n + 1);
    CK(cudaMemsetAsync(ctx->alive_prev.p, 0, n + 1, ctx->stream));
    ctx->health_n = n;
  }
  int rc = stage_h2d(ctx, recs, n * sizeof(kvg_pci_rec), ctx->recs.p);
  if (rc) return rc;
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, 64, ctx->stream));
  HealthOp op;
  op.recs = ctx->recs.p;
  op.n = (uint32_t)n;
  op.alive_prev = ctx->alive_prev.p;
  op.changed = ctx->changed.p;
  op.ctrl = ctx->ctrl.p;
  op.local_alive = 0;
  LAUNCH("health_diff", k_compact<HealthOp>, compact_grid<HealthOp>(ctx, n), KVG_BLOCK, 0, op,
         ctx->classify_state.p, next_epoch());
  rc = check_launch(ctx, "health");
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  size_t nc = ctx->h_ctrl->n_changed;
  size_t o_list = align64(sizeof(kvg_health_delta));
  void* blk = pinned_alloc(ctx, o_list + align64(nc * 4));
  if (!blk) return KVG_ENOMEM;
  uint8_t* b = pinned_payload(blk);
  if (nc) {
    CK(cudaMemcpyAsync(b + o_list, ctx->changed.p, nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  kvg_health_delta* d = (kvg_health_delta*)b;
  d->n_records = (uint32_t)n;
  d->n_alive = ctx->h_ctrl->n_alive;
  d->n_changed = (uint32_t)nc;
  d->changed = (const uint32_t*)(b + o_list);
  *delta = d;
  ctx->last_kind = 0;
  return KVG_OK;
}

// ---- mdev ---------------------------------------------------------------------------------------
static int load_type_dict(kvg_ctx* ctx, const kvg_type_dict* types) {
  uint32_t nt = types->n_types;
  if (nt > 65535) {
    ctx->err = "more than 65535 mdev types";
    return KVG_ERANGE;
  }
  size_t raw_len = nt ? types->off[nt] : 0;
  ENSURE(ctx->type_raw, raw_len + 16);
  ENSURE(ctx->type_label, raw_len + 16);
  ENSURE(ctx->type_off, (size_t)nt + 2);
  ENSURE(ctx->type_label_len, (size_t)nt + 1);
  ENSURE(ctx->type_canon, (size_t)nt + 1);
  ENSURE(ctx->type_match, (size_t)nt + 1);
  ENSURE(ctx->type_name_len, (size_t)nt + 1);
  ENSURE(ctx->type_hash, (size_t)nt + 1);
  ctx->n_types = nt;
  ctx->h_type_off.assign(types->off, types->off + nt + 1);
  if (nt == 0) return KVG_OK;
  if (raw_len) CK(cudaMemcpyAsync(ctx->type_raw.p, types->bytes, raw_len, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->type_off.p, types->off, 4 * ((size_t)nt + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the caller's dictionary may be freed after return
  int grid = (int)((nt + 63) / 64);
  LAUNCH("mdev_labels", k_mdev_labels, grid, 64, 0, ctx->type_raw.p, ctx->type_off.p, nt,
         ctx->type_label.p, ctx->type_label_len.p, ctx->type_hash.p);
  LAUNCH("mdev_canon", k_mdev_canon, grid, 64, 0, ctx->type_label.p, ctx->type_off.p,
         ctx->type_label_len.p, ctx->type_hash.p, nt, ctx->type_canon.p);
  return check_launch(ctx, "mdev labels");
}

}  // extern "C"

// label-keyed lookups need (offset,len) pairs rather than a prefix-offset array: a tiny kernel
// compacts the labels into a contiguous key blob + offsets for k_lookup_general
__global__ void k_pack_labels(const uint8_t* __restrict__ label, const uint32_t* __restrict__ raw_off,
                              const uint32_t* __restrict__ label_len, uint32_t n_types,
                              uint8_t* __restrict__ blob, uint32_t* __restrict__ blob_off) {
  pdl_enter();
  // single thread per type after a serial prefix by thread 0 (n_types <= 65535, tiny)
  __shared__ uint32_t total;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t o = 0;
    for (uint32_t k = 0; k < n_types; k++) {
      blob_off[k] = o;
      o += label_len[k];
    }
    blob_off[n_types] = o;
    total = o;
  }
  __syncthreads();
  (void)total;
  for (uint32_t k = threadIdx.x; k < n_types; k += blockDim.x) {
    uint32_t o = blob_off[k];
    for (uint32_t t = 0; t < label_len[k]; t++) blob[o + t] = label[raw_off[k] + t];
  }
}

extern "C" {

int kvg_dev_scan_mdev(kvg_ctx* ctx, const void* d_recs, size_t n, const kvg_type_dict* types) {
  if (!ctx || !types || (!d_recs && n) || n > 0xfffffff0ull || ((uintptr_t)d_recs & 15)) return KVG_EINVAL;
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load must precede a scan (the scan joins names)");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  int rc = load_type_dict(ctx, types);
  if (rc) return rc;
  const uint32_t nt = ctx->n_types;
  // resource-name join for every label: getDeviceName(label) (:152) — exact prefix semantics
  const uint32_t NAME_CAP = 256;
  if (nt) {
    ENSURE(ctx->keys_blob, (size_t)ctx->h_type_off[nt] + 16);
    ENSURE(ctx->keys_off, (size_t)nt + 2);
    ENSURE(ctx->type_names, (size_t)nt * NAME_CAP);
    LAUNCH("pack_labels", k_pack_labels, 1, KVG_BLOCK, 0, ctx->type_label.p, ctx->type_off.p,
           ctx->type_label_len.p, nt, ctx->keys_blob.p, ctx->keys_off.p);
    rc = lookup_general(ctx, ctx->keys_blob.p, ctx->keys_off.p, nt, NAME_CAP, ctx->type_names.p,
                        ctx->type_name_len.p, ctx->type_match.p);
    if (rc) return rc;
  }
  ENSURE(ctx->surv, 2 * (n + 1));
  const size_t mdev_tiles = (n + (size_t)KVG_BLOCK * MDEV_ROWS - 1) / ((size_t)KVG_BLOCK * MDEV_ROWS) + 1;
  ENSURE(ctx->classify_state, 2 * mdev_tiles);
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  MdevClassifyOp op;
  op.recs = (const uint4*)d_recs;
  op.n = (uint32_t)n;
  op.out = ctx->surv.p;
  op.ctrl = ctx->ctrl.p;
  op.type_canon = ctx->type_canon.p;
  op.n_types = nt;
  op.local_max_parent = 0;
  op.local_max_type = 0;
  {
    constexpr int T = 128, R = 4;  // 512 x 32-byte records = 16 KiB per tile
    const size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
    if (tiles) {
      ENSURE(ctx->ragged, 2 * tiles * T * R);
      ENSURE(ctx->tile_count, tiles + 1);
      ENSURE(ctx->tile_off, tiles + 2);
      ENSURE(ctx->tile_max, tiles + 1);
      const unsigned chunks = (unsigned)((tiles + C_TILE - 1) / C_TILE);
      ENSURE(ctx->offs_state, (size_t)chunks + 1);
      uint4* dense = op.out;
      op.out = ctx->ragged.p;
      LAUNCH("mdev_classify_compact", (k_classify_ragged<MdevClassifyOp, T, R>), (unsigned)tiles, T, 0, op,
             ctx->tile_count.p, ctx->tile_max.p);
      {
        TileOffsetsArgs2 tt;
        tt.o[0] = {ctx->tile_count.p, ctx->tile_max.p, nullptr, (uint32_t)tiles, ctx->tile_off.p,
                   &ctx->ctrl.p->n_surv, ctx->offs_state.p};
        tt.o[1] = tt.o[0];
        LAUNCH("tile_offsets", k_tile_offsets, chunks, KVG_BLOCK, 0, tt, ctx->ctrl.p, next_epoch());
      }
      LAUNCH("pack_survivors", k_pack_survivors<2>, (unsigned)tiles, 128, 0, ctx->ragged.p, ctx->tile_off.p,
             (uint32_t)(T * R), dense);
    }
  }
  rc = check_launch(ctx, "mdev classify");
  if (rc) return rc;
  rc = enqueue_orderings(ctx, n, SRC_MDEV_TYPE, SRC_MDEV_PARENT);
  if (rc) return rc;
  ctx->last_n = n;
  ctx->last_total = n;
  ctx->last_kind = 2;
  return KVG_OK;
}

int kvg_dev_scan_mdev_fetch(kvg_ctx* ctx, kvg_mdev_result** res) {
  if (!ctx || !res || ctx->last_kind != 2) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t S = ctx->h_ctrl->n_surv, KT = ctx->h_ctrl->n_dev_keys, P = ctx->h_ctrl->n_groups;
  const uint32_t nt = ctx->n_types;
  const size_t raw_len = nt ? ctx->h_type_off[nt] : 0;
  const uint32_t NAME_CAP = 256;
  size_t o = align64(sizeof(kvg_mdev_result));
  size_t o_surv = o; o += align64(S * 32);
  size_t o_tk32 = o; o += align64(KT * 4);
  size_t o_tk = o; o += align64(KT * 2);
  size_t o_toff = o; o += align64((KT + 1) * 4);
  size_t o_tperm = o; o += align64(S * 4);
  size_t o_lraw = o; o += align64(raw_len + 16);
  size_t o_llen = o; o += align64(((size_t)nt + 1) * 4);
  size_t o_loff = o; o += align64(((size_t)nt + 1) * 4);
  size_t o_lbytes = o; o += align64(raw_len + 16);
  size_t o_canon = o; o += align64(((size_t)nt + 1) * 2);
  size_t o_nraw = o; o += align64((size_t)nt * NAME_CAP + 16);
  size_t o_nlen = o; o += align64(((size_t)nt + 1) * 4);
  size_t o_noff = o; o += align64(((size_t)nt + 1) * 4);
  size_t o_nbytes = o; o += align64((size_t)nt * NAME_CAP + 16);
  size_t o_pk = o; o += align64(P * 4);
  size_t o_poff = o; o += align64((P + 1) * 4);
  size_t o_pperm = o; o += align64(S * 4);
  void* blk = pinned_alloc(ctx, o);
  if (!blk) return KVG_ENOMEM;
  uint8_t* b = pinned_payload(blk);
  auto D2H = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    return cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  OrderBufs& ot = ctx->ord_dev;
  OrderBufs& op = ctx->ord_grp;
  CK(D2H(o_surv, ctx->surv.p, S * 32));
  CK(D2H(o_tk32, ot.seg_key.p, KT * 4));
  CK(D2H(o_toff, ot.seg_off.p, (KT + 1) * 4));
  CK(D2H(o_tperm, ot.perm.p, S * 4));
  CK(D2H(o_pk, op.seg_key.p, P * 4));
  CK(D2H(o_poff, op.seg_off.p, (P + 1) * 4));
  CK(D2H(o_pperm, op.perm.p, S * 4));
  if (nt) {
    CK(D2H(o_lraw, ctx->type_label.p, raw_len));
    CK(D2H(o_llen, ctx->type_label_len.p, (size_t)nt * 4));
    CK(D2H(o_canon, ctx->type_canon.p, (size_t)nt * 2));
    CK(D2H(o_nraw, ctx->type_names.p, (size_t)nt * NAME_CAP));
    CK(D2H(o_nlen, ctx->type_name_len.p, (size_t)nt * 4));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  kvg_mdev_result* r = (kvg_mdev_result*)b;
  memset(r, 0, sizeof *r);
  if (S == 0) ((uint32_t*)(b + o_toff))[0] = 0, ((uint32_t*)(b + o_poff))[0] = 0;
  r->n_records = ctx->last_n;
  r->n_survivors = S;
  r->survivors = (const kvg_mdev_surv*)(b + o_surv);
  r->n_type_keys = (uint32_t)KT;
  uint16_t* tk = (uint16_t*)(b + o_tk);
  for (size_t k = 0; k < KT; k++) tk[k] = (uint16_t)((const uint32_t*)(b + o_tk32))[k];
  r->type_keys = tk;
  r->type_off = (const uint32_t*)(b + o_toff);
  r->type_perm = (const uint32_t*)(b + o_tperm);
  // repack labels / names contiguously (marshalling of GPU-produced bytes)
  r->n_types = nt;
  uint32_t* loff = (uint32_t*)(b + o_loff);
  uint32_t* noff = (uint32_t*)(b + o_noff);
  const uint32_t* llen = (const uint32_t*)(b + o_llen);
  const uint32_t* nlen = (const uint32_t*)(b + o_nlen);
  size_t lo = 0, no = 0;
  for (uint32_t k = 0; k < nt; k++) {
    loff[k] = (uint32_t)lo;
    memcpy(b + o_lbytes + lo, b + o_lraw + ctx->h_type_off[k], llen[k]);
    lo += llen[k];
    noff[k] = (uint32_t)no;
    uint32_t nl = nlen[k] > NAME_CAP ? NAME_CAP : nlen[k];
    memcpy(b + o_nbytes + no, b + o_nraw + (size_t)k * NAME_CAP, nl);
    no += nl;
  }
  loff[nt] = (uint32_t)lo;
  noff[nt] = (uint32_t)no;
  r->label_off = loff;
  r->label_bytes = b + o_lbytes;
  r->type_canon = (const uint16_t*)(b + o_canon);
  r->type_name_off = noff;
  r->type_name_bytes = b + o_nbytes;
  r->n_parents = (uint32_t)P;
  r->par_keys = (const uint32_t*)(b + o_pk);
  r->par_off = (const uint32_t*)(b + o_poff);
  r->par_perm = (const uint32_t*)(b + o_pperm);
  *res = r;
  return KVG_OK;
}

int kvg_scan_mdev(kvg_ctx* ctx, const kvg_mdev_rec* recs, size_t n, const kvg_type_dict* types,
                  kvg_mdev_result** res) {
  if (!ctx || !res || !types || (!recs && n)) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  ENSURE(ctx->recs, 2 * (n + 1));
  int rc = stage_h2d(ctx, recs, n * sizeof(kvg_mdev_rec), ctx->recs.p);
  if (rc) return rc;
  rc = kvg_dev_scan_mdev(ctx, ctx->recs.p, n, types);
  if (rc) return rc;
  return kvg_dev_scan_mdev_fetch(ctx, res);
}

// ---- generators, flush --------------------------------------------------------------------------
int kvg_dev_gen_pci(kvg_ctx* ctx, void* d_recs, uint64_t first, size_t n, const uint16_t* nv_ids,
                    uint32_t n_nv_ids, uint32_t group_bits) {
  if (!ctx || (!d_recs && n) || n > 0xfffffff0ull) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  ENSURE(ctx->nv_ids, (size_t)n_nv_ids + 1);
  if (n_nv_ids) {
    CK(cudaMemcpyAsync(ctx->nv_ids.p, nv_ids, 2 * (size_t)n_nv_ids, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (n == 0) return KVG_OK;
  LAUNCH("gen_pci", k_gen_pci, ctx->sm_count * 8, KVG_BLOCK, 0, (uint4*)d_recs, first, (uint32_t)n,
         ctx->nv_ids.p, n_nv_ids, group_bits);
  return check_launch(ctx, "gen_pci");
}
int kvg_dev_gen_mdev(kvg_ctx* ctx, void* d_recs, uint64_t first, size_t n) {
  if (!ctx || (!d_recs && n) || n > 0xfffffff0ull) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return KVG_OK;
  LAUNCH("gen_mdev", k_gen_mdev, ctx->sm_count * 8, KVG_BLOCK, 0, (uint4*)d_recs, first, (uint32_t)n);
  return check_launch(ctx, "gen_mdev");
}
// diagnostic: decomposed classify kernel (see k_debug_classify); returns device ms via *ms_out
// host-side view of the device's radix plan (same __host__ __device__ function): lets CPU-only tests pin
// the pass structure the kernels will choose for a given largest key
int kvg_debug_radix_plan(uint32_t max_key, uint32_t key_bits_max, uint32_t max_bits, uint32_t* npass,
                         uint32_t* shifts4, uint32_t* bits4) {
  if (!npass || !shifts4 || !bits4 || key_bits_max == 0 || key_bits_max > 32 || max_bits == 0 || max_bits > 16)
    return KVG_EINVAL;
  *npass = radix_plan(max_key, key_bits_max, 0, max_bits).npass;
  for (uint32_t p = 0; p < 4; p++) {
    RadixPlan r = radix_plan(max_key, key_bits_max, p, max_bits);
    shifts4[p] = r.shift;
    bits4[p] = r.bits;
  }
  return KVG_OK;
}

int kvg_dev_debug_classify(kvg_ctx* ctx, const void* d_recs, size_t n, int mode, int rows, float* ms_out) {
  if (!ctx || !d_recs || !ms_out || n == 0 || n > 0xfffffff0ull) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  ENSURE(ctx->surv, n + 1);
  ENSURE(ctx->probe_slots, 4);
  ENSURE(ctx->nv_index, 65536);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  CK(cudaMemsetAsync(ctx->probe_slots.p, 0, 4, ctx->stream));
  cudaEventRecord(a, ctx->stream);
  if (rows == 4) {
    size_t tiles = (n + 511) / 512;
    k_debug_classify<128, 4><<<(unsigned)tiles, 128, 0, ctx->stream>>>((const uint4*)d_recs, (uint32_t)n, ctx->surv.p, ctx->nv_index.p, ctx->probe_slots.p, mode);
  } else if (rows == 16) {
    size_t tiles = (n + 2047) / 2048;
    k_debug_classify<128, 16><<<(unsigned)tiles, 128, 0, ctx->stream>>>((const uint4*)d_recs, (uint32_t)n, ctx->surv.p, ctx->nv_index.p, ctx->probe_slots.p, mode);
  } else {
    size_t tiles = (n + 1023) / 1024;
    k_debug_classify<128, 8><<<(unsigned)tiles, 128, 0, ctx->stream>>>((const uint4*)d_recs, (uint32_t)n, ctx->surv.p, ctx->nv_index.p, ctx->probe_slots.p, mode);
  }
  cudaEventRecord(b, ctx->stream);
  CK(cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(ms_out, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return check_launch(ctx, "debug classify");
}

int kvg_dev_flush_l2(kvg_ctx* ctx) {
  if (!ctx) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const size_t n16 = (size_t)(192u << 20) / 16;  // 192 MiB > 126 MB L2
  ENSURE(ctx->flush, n16);
  kvg::k_fill<<<ctx->sm_count * 8, KVG_BLOCK, 0, ctx->stream>>>(ctx->flush.p, n16, (uint32_t)ctx->launches);
  return check_launch(ctx, "flush");
}

// ================================================================================================
// multi-GPU: range-sharded records, one allgatherv of survivors (BASELINE.json config 4)
// ================================================================================================
int kvg_comm_unique_id(void* out128) {
  if (!out128) return KVG_EINVAL;
  std::string err;
  if (!g_nccl.load(&err)) {
    g_create_error = err;
    return KVG_ENCCL;
  }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return KVG_ENCCL;
  memcpy(out128, &id, KVG_UNIQUE_ID_BYTES);
  return KVG_OK;
}

int kvg_comm_init(kvg_ctx* ctx, int rank, int nranks, const void* unique_id128) {
  if (!ctx || !unique_id128 || nranks < 1 || rank < 0 || rank >= nranks) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (!g_nccl.load(&ctx->err)) return KVG_ENCCL;
  ncclUniqueId id;
  memcpy(&id, unique_id128, KVG_UNIQUE_ID_BYTES);
  ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
  if (r != 0) {
    ctx->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
    return KVG_ENCCL;
  }
  ctx->rank = rank;
  ctx->nranks = nranks;
  ENSURE(ctx->gather_counts, (size_t)nranks + 1);
  if (!ctx->h_counts) CK(cudaMallocHost((void**)&ctx->h_counts, sizeof(uint64_t) * ((size_t)nranks + 1)));
  return KVG_OK;
}

// ---- peer-memory gather: set-up ------------------------------------------------------------------
// kvg_comm_p2p_export: allocate this rank's gather windows for shards of up to cap_local records and
// return the 64-byte CUDA IPC handle the other ranks need.  kvg_comm_p2p_import: open every rank's
// handle (all_handles = nranks x 64 bytes, rank order).  After both succeeded kvg_dev_scan_pci_sharded
// uses the peer-memory path (no NCCL, no host synchronisation).  Any failure leaves the NCCL path.
int kvg_comm_p2p_export(kvg_ctx* ctx, int rank, int nranks, size_t cap_local, void* handle_out64) {
  if (!ctx || !handle_out64 || nranks < 1 || nranks > P2P_MAX_RANKS || rank < 0 || rank >= nranks || !cap_local)
    return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  static_assert(sizeof(P2PCtrl) <= P2P_HDR, "control block fits its pad");
  if (ctx->p2p_mine) {
    ctx->err = "peer windows already exported";
    return KVG_ESTATE;
  }
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->p2p_cap = cap_local;
  const size_t bytes = P2P_HDR + 2 * (size_t)nranks * cap_local * 16;
  CK(cudaMalloc((void**)&ctx->p2p_mine, bytes));
  CK(cudaMemset(ctx->p2p_mine, 0, P2P_HDR));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->p2p_mine));
  memcpy(handle_out64, &h, 64);
  return KVG_OK;
}

int kvg_comm_p2p_import(kvg_ctx* ctx, const void* all_handles) {
  if (!ctx || !all_handles || !ctx->p2p_mine) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  for (int q = 0; q < ctx->nranks; q++) {
    if (q == ctx->rank) {
      ctx->p2p_peer[q] = ctx->p2p_mine;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)all_handles + 64 * (size_t)q, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      ctx->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(q) + "): " + cudaGetErrorString(e);
      cudaGetLastError();
      return KVG_ECUDA;
    }
    ctx->p2p_peer[q] = (uint8_t*)p;
  }
  ENSURE(ctx->gather_base, P2P_MAX_RANKS + 2);
  ENSURE(ctx->p2p_err, 4);
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->p2p_step = 0;
  return KVG_OK;
}

// switch the sharded scan between the peer-memory path (on != 0; needs a successful import on EVERY
// rank — the caller agrees on that collectively) and the NCCL path
int kvg_comm_p2p_enable(kvg_ctx* ctx, int on) {
  if (!ctx) return KVG_EINVAL;
  if (on && (!ctx->p2p_mine || !ctx->gather_base.p)) {
    ctx->err = "peer windows are not imported";
    return KVG_ESTATE;
  }
  ctx->p2p = on != 0;
  return KVG_OK;
}

int kvg_comm_destroy(kvg_ctx* ctx) {
  if (!ctx) return KVG_EINVAL;
  if (ctx->comm) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->nranks = 1;
  ctx->rank = 0;
  return KVG_OK;
}

}  // extern "C"

// Sharded scan over peer memory: classify -> offsets -> [pack fused with the all-gather] -> signal;
// wait for all regions -> dense copy -> ack; key-partitioned orderings.  Fully asynchronous.
static int scan_sharded_p2p(kvg_ctx* ctx, const void* d_recs, size_t n_local) {
  const int P = ctx->nranks;
  if (n_local > ctx->p2p_cap) {
    ctx->err = "shard larger than the exported peer window";
    return KVG_ERANGE;
  }
  const size_t cap_total = (size_t)P * ctx->p2p_cap;
  ENSURE(ctx->surv, cap_total + 1);
  constexpr int T = 128, R = 8;
  const size_t tiles = (n_local + (size_t)T * R - 1) / ((size_t)T * R);
  ENSURE(ctx->ragged, (tiles ? tiles : 1) * T * R);
  ENSURE(ctx->tile_count, tiles + 1);
  ENSURE(ctx->tile_off, tiles + 2);
  ENSURE(ctx->tile_max, tiles + 1);
  const unsigned chunks = (unsigned)((tiles + C_TILE - 1) / C_TILE);
  ENSURE(ctx->offs_state, (size_t)chunks + 2);
  const unsigned long long step = ++ctx->p2p_step;
  const uint32_t w = (uint32_t)(step & 1);
  P2PPeers peers;
  memset(&peers, 0, sizeof peers);
  for (int q = 0; q < P; q++) {
    peers.ctrl[q] = (P2PCtrl*)ctx->p2p_peer[q];
    peers.win[q] = (uint4*)(ctx->p2p_peer[q] + P2P_HDR) + (size_t)w * cap_total;
  }
  P2PCtrl* mine = (P2PCtrl*)ctx->p2p_mine;
  const uint4* my_window = (const uint4*)(ctx->p2p_mine + P2P_HDR) + (size_t)w * cap_total;

  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  CK(cudaMemsetAsync(ctx->p2p_err.p, 0, sizeof(uint32_t), ctx->stream));
  if (step > 2) LAUNCH("p2p_wait_acks", k_p2p_wait_acks, 1, 32, 0, mine, (uint32_t)P, step - 2, ctx->p2p_err.p);
  if (tiles) {
    PciClassifyOp op;
    op.recs = (const uint4*)d_recs;
    op.n = (uint32_t)n_local;
    op.out = (kvg_pci_surv*)ctx->ragged.p;
    op.ctrl = ctx->ctrl.p;
    op.table = ctx->tables.p;
    op.cap_mask = (1u << ctx->cap_log2) - 1;
    op.cap_shift = 32 - ctx->cap_log2;
    op.info = ctx->info.p;
    op.nv_index = ctx->nv_index.p;
    op.local_max_group = 0;
    op.local_max_dev = 0;
    LAUNCH("classify_compact", (k_classify_ragged<PciClassifyOp, T, R>), (unsigned)tiles, T, 0, op,
           ctx->tile_count.p, ctx->tile_max.p);
    TileOffsetsArgs2 tt;
    tt.o[0] = {ctx->tile_count.p, ctx->tile_max.p, nullptr, (uint32_t)tiles, ctx->tile_off.p,
               &ctx->ctrl.p->n_own[0], ctx->offs_state.p};  // n_own[0] doubles as "local count" here
    tt.o[1] = tt.o[0];
    LAUNCH("tile_offsets", k_tile_offsets, chunks, KVG_BLOCK, 0, tt, ctx->ctrl.p, next_epoch());
    // the pack IS the all-gather: every survivor goes straight into every peer's window
    LAUNCH("pack_to_peers", k_pack_to_peers, (unsigned)tiles, 128, 0, (const uint4*)ctx->ragged.p,
           ctx->tile_off.p, (uint32_t)(T * R), peers, (uint32_t)P, (size_t)ctx->rank * ctx->p2p_cap);
  }
  LAUNCH("p2p_signal", k_p2p_signal, 1, 32, 0, peers, (uint32_t)P, (uint32_t)ctx->rank, w, step,
         &ctx->ctrl.p->n_own[0]);
  LAUNCH("p2p_wait_gather", k_p2p_wait_gather, 1, 32, 0, mine, (uint32_t)P, w, step, ctx->gather_base.p,
         ctx->ctrl.p, ctx->p2p_err.p);
  dim3 cgrid((unsigned)std::max(1, ctx->sm_count * 2 / P), (unsigned)P);
  LAUNCH("p2p_copy_regions", k_p2p_copy_regions, cgrid, KVG_BLOCK, 0, my_window, ctx->p2p_cap,
         ctx->gather_base.p, ctx->surv.p);
  LAUNCH("p2p_ack", k_p2p_ack, 1, 32, 0, peers, (uint32_t)P, (uint32_t)ctx->rank, step);
  int rc = check_launch(ctx, "p2p gather");
  if (rc) return rc;
  // n_own[0] was used as scratch for the local count: the ownership select rewrites it.  The key
  // maxima reduced by k_tile_offsets (local shard, then the owned pairs) bound the radix passes.
  rc = enqueue_pci_orderings(ctx, cap_total, /*owned_only=*/P > 1);
  if (rc) return rc;
  ctx->last_n = n_local;
  ctx->last_total = cap_total;
  ctx->last_kind = 1;
  ctx->last_owned = P > 1;
  return KVG_OK;
}

// after the gather: n_surv <- total, maxima already all-reduced by construction (each rank
// recomputes them from the gathered list)
__global__ void k_gathered_maxima(const kvg_pci_surv* __restrict__ s, uint32_t n, ScanCtrl* ctrl) {
  pdl_enter();
  uint32_t mg = 0, md = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    mg = max(mg, s[i].iommu_group);
    md = max(md, (uint32_t)s[i].device);
  }
  mg = warp_max(mg);
  md = warp_max(md);
  if (lane_id() == 0) {
    if (mg) atomicMax(&ctrl->max_group, mg);
    if (md) atomicMax(&ctrl->max_devkey, md);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->n_surv = n;
}

extern "C" {

int kvg_dev_scan_pci_sharded(kvg_ctx* ctx, const void* d_recs, size_t n_local) {
  if (!ctx || (!d_recs && n_local) || n_local > 0xfffffff0ull) return KVG_EINVAL;
  if (!ctx->comm && !ctx->p2p) {
    ctx->err = "kvg_comm_init / kvg_comm_p2p_import has not been called";
    return KVG_ESTATE;
  }
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load must precede a scan");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  const int P = ctx->nranks;
  if (ctx->p2p) return scan_sharded_p2p(ctx, d_recs, n_local);
  ENSURE(ctx->local_surv, n_local + 1);
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  int rc = enqueue_classify(ctx, d_recs, n_local, ctx->local_surv.p);
  if (rc) return rc;
  // counts: every rank learns every shard's survivor count (8 bytes per rank)
  {
    // widen the device-side u32 count to u64 in place of a dedicated kernel: copy via host pinned
    CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_counts[P] = ctx->h_ctrl->n_surv;
    CK(cudaMemcpyAsync(ctx->gather_counts.p + P, &ctx->h_counts[P], 8, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = g_nccl.AllGather(ctx->gather_counts.p + P, ctx->gather_counts.p, 1, ncclUint64,
                                      ctx->comm, ctx->stream);
    if (r != 0) {
      ctx->err = std::string("ncclAllGather(counts): ") + g_nccl.GetErrorString(r);
      return KVG_ENCCL;
    }
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->gather_counts.p, 8 * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  size_t total = 0;
  std::vector<size_t> displ((size_t)P);
  for (int r = 0; r < P; r++) {
    displ[(size_t)r] = total;
    total += ctx->h_counts[r];
  }
  if (total > 0xfffffff0ull) {
    ctx->err = "gathered survivor list exceeds 2^32 entries";
    return KVG_ERANGE;
  }
  ENSURE(ctx->surv, total + 1);
  // allgatherv = one grouped broadcast per root; rank order == Walk order of the shards
  {
    ncclResult_t r = g_nccl.GroupStart();
    for (int root = 0; root < P && r == 0; root++) {
      size_t cnt = ctx->h_counts[root];
      if (cnt == 0) continue;
      r = g_nccl.Broadcast(root == ctx->rank ? (const void*)ctx->local_surv.p : nullptr,
                           ctx->surv.p + displ[(size_t)root], cnt * 16, ncclUint8, root, ctx->comm,
                           ctx->stream);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != 0 || r2 != 0) {
      ctx->err = std::string("ncclBroadcast group: ") + g_nccl.GetErrorString(r ? r : r2);
      return KVG_ENCCL;
    }
  }
  // control block for the ordering phase: zero, then n_surv <- total (known on the host; a 4-byte
  // copy from pinned memory).  The key maxima come from the ownership select (k_tile_offsets).
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  ctx->h_counts[P] = total;  // pinned; low 32 bits are the value (little endian)
  CK(cudaMemcpyAsync(&ctx->ctrl.p->n_surv, &ctx->h_counts[P], sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (P == 1) {  // single rank: no ownership select runs, so reduce the maxima here
    LAUNCH("gathered_maxima", k_gathered_maxima, ctx->sm_count * 4, KVG_BLOCK, 0,
           (const kvg_pci_surv*)ctx->surv.p, (uint32_t)total, ctx->ctrl.p);
  }
  // bucketing partitioned by key: this rank orders only the keys with key % nranks == rank
  rc = enqueue_pci_orderings(ctx, total, /*owned_only=*/P > 1);
  if (rc) return rc;
  ctx->last_n = n_local;
  ctx->last_total = total;
  ctx->last_kind = 1;
  ctx->last_owned = P > 1;
  return KVG_OK;
}

}  // extern "C"
