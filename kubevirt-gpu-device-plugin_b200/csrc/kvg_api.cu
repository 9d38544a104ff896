// kvg_api.cu — the C-ABI of libkvgpu.so (include/kvgpu.h): context, HBM layout, launch sequencing
// and result marshalling around the kernels in kvg_parse.cuh / kvg_scan.cuh.
//
// HBM layout owned by a context (all cudaMalloc'd once and grown geometrically, never per call):
//   text      pci.ids image(s), padded with '\n' to a tile multiple + 16 (TMA halo)
//   dev_off   per image 65,536 x u32: line offset of the first "\t<id>" line under a 10de header
//   nv_index  65,536 x u32: device id -> name pool slot (what the scans join against)
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
#include "kvg_parse_k1.cuh"
#include "kvg_scan.cuh"
#include "kvg_order.cuh"
#include "kvg_shard.cuh"

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
  DevBuf<uint32_t> dev_off;   // [n_files][65536]
  DevBuf<PciIdsInfo> info;
  DevBuf<uint4> span_sum;     // K1: one 16-byte summary per 4 KiB span
  int k1_grid = 0;            // co-resident CTAs of k_pciids_scan on this device
  bool sec_lines_ready = false;  // candidate lines of the general lookup collected for the current table
  DevBuf<uint8_t> pool;
  DevBuf<uint32_t> nv_index;  // [65536] vendor-10de device id -> name pool slot
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
  uint32_t parsed_files = 0;

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
  DevBuf<uint32_t> tile_hist;     // K4: [2 orderings][<= 2048 digits][T tiles]
  DevBuf<uint32_t> bin_total;     // [2 orderings][2048] digit totals
  bool health_smem_set = false;
  int final_ctas_per_sm = 0;      // occupancy of k_order_final on this device (bounds its grid)
  bool scatter_smem_set = false;  // dynamic shared-memory opt-in of k_order_scatter<11> done on this device
  size_t last_n = 0;     // records of the last enqueued scan
  size_t last_total = 0; // survivors capacity used by the last scan (sharded: all ranks)
  int last_kind = 0;     // 1 = pci, 2 = mdev
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
  uint32_t health_seq = 0;
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
  // multi-GPU (kvg_shard.cuh): receive window [ShardCtrl pad 4 KiB][parity][ordering][source][cap records]
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  bool p2p = false;                      // peer windows imported on every rank: exchange over NVLink stores
  uint8_t* win_mine = nullptr;           // my window allocation (IPC-exported in p2p mode)
  size_t win_cap = 0;                    // PCI records per region (an mdev record takes two)
  uint8_t* win_peer[SH_MAX_RANKS] = {};  // peer-mapped bases (own entry = win_mine)
  unsigned long long shard_step = 0;
  DevBuf<uint32_t> shard_cnt;            // [2][P] totals, 2 tickets, 1 error word (64-word header)
  DevBuf<uint64_t> shard_state;          // [2][P][T] chained-scan words of k_shard_send
  DevBuf<uint32_t> send_words;           // [tiles][4 * cw4] published tile counts of k_classify_send
  uint32_t send_epoch = 0;               // 1 .. 2^21 - 2, 0 = clear send_words first
  int send_cw4 = 0;
  uint32_t* shard_err = nullptr;         // that error word (device)
  DevBuf<uint4> owned0, owned1;          // dense owned lists of ordering 0 / 1
  DevBuf<uint4> gathered;                // NCCL mode: the all-gathered survivor list
  DevBuf<uint64_t> gather_counts;        // NCCL mode: [nranks + 1]
  uint64_t* h_counts = nullptr;          // pinned [nranks + 1]
  int last_units = 1;                    // 16-byte units per record of the last sharded scan
};
static const size_t SH_HDR = 4096;

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
  release(ctx->text); release(ctx->dev_off); release(ctx->info); release(ctx->span_sum); release(ctx->pool); release(ctx->ctrl);
  release(ctx->nv_index); release(ctx->sec_lines); release(ctx->type_hash); release(ctx->ragged); release(ctx->tile_count); release(ctx->tile_off);
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
  release(ctx->match_out); release(ctx->gather_counts); release(ctx->gathered);
  release(ctx->shard_cnt); release(ctx->shard_state); release(ctx->send_words); release(ctx->owned0); release(ctx->owned1);
  for (int q = 0; q < SH_MAX_RANKS; q++)
    if (ctx->win_peer[q] && ctx->win_peer[q] != ctx->win_mine) cudaIpcCloseMemHandle(ctx->win_peer[q]);
  if (ctx->win_mine) cudaFree(ctx->win_mine);
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

// K1 (kvg_parse_k1.cuh) for n_files images: prep -> scan -> resolve + finalize -> names (image 0).
// Everything is enqueued; nothing here waits for the device.
static int parse_enqueue(kvg_ctx* ctx, const uint8_t* d_text, size_t len, size_t stride, uint32_t n_files) {
  if (len == 0 || len >= 0xfffffff0ull) {
    ctx->err = "pci.ids length out of range";
    return KVG_EINVAL;
  }
  const uint32_t spf = (uint32_t)((len + K1_SPAN - 1) / K1_SPAN);
  const uint64_t n_spans64 = (uint64_t)spf * n_files;
  if (n_spans64 > 0x7fffffffull) {
    ctx->err = "too many spans";
    return KVG_EINVAL;
  }
  const uint32_t n_spans = (uint32_t)n_spans64;
  const size_t pool_bytes = (len + 16 + 15) & ~(size_t)15;
  {
    // the table is self-cleaning (k_pciids_names resets what it reads): only a fresh allocation is filled
    const uint32_t* before = ctx->dev_off.p;
    ENSURE(ctx->dev_off, (size_t)K1_IDS * n_files);
    if (ctx->dev_off.p != before) CK(cudaMemsetAsync(ctx->dev_off.p, 0xff, ctx->dev_off.cap * sizeof(uint32_t), ctx->stream));
  }
  ENSURE(ctx->info, n_files);  // zero-filled when fresh: the accumulators in PciIdsInfo::pad start at "nothing seen"
  ENSURE(ctx->span_sum, n_spans);
  ENSURE(ctx->nv_index, K1_IDS);
  ENSURE(ctx->pool, pool_bytes);
  if (!ctx->k1_grid) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pciids_scan, K1_WARPS * 32, K1_SMEM);
    ctx->k1_grid = ctx->sm_count * (occ < 1 ? 1 : occ);
  }
  K1Args A;
  A.text = d_text;
  A.stride = stride;
  A.len = (uint32_t)len;
  A.n_files = n_files;
  A.spans_per_file = spf;
  A.n_spans = n_spans;
  A.dev_off = ctx->dev_off.p;
  A.info = ctx->info.p;
  A.span_sum = ctx->span_sum.p;
  A.pool = (uint4*)ctx->pool.p;
  A.pool16 = (uint32_t)(pool_bytes / 16);
  {
    unsigned grid = (n_spans + K1_WARPS - 1) / K1_WARPS;
    if (grid > (unsigned)ctx->k1_grid) grid = (unsigned)ctx->k1_grid;
    LAUNCH("pciids_parse", k_pciids_scan, grid, K1_WARPS * 32, K1_SMEM, A);
  }
  LAUNCH("pciids_resolve", k_pciids_resolve_finalize, n_files + (n_spans + K1_RWARPS - 1) / K1_RWARPS, KVG_BLOCK,
         K1_RWARPS * K1_STAGE, A);
  LAUNCH("pciids_names", k_pciids_names, K1_IDS / 32 / KVG_WARPS, KVG_BLOCK, 0, ctx->dev_off.p, n_files, d_text,
         (uint32_t)len, ctx->info.p, ctx->nv_index.p, ctx->pool.p);
  ctx->sec_lines_ready = false;
  return check_launch(ctx, "pciids parse");
}

// after the parse of image 0: mirror the section facts and the name pool on the host
static int table_publish(kvg_ctx* ctx, const uint8_t* d_text, size_t len) {
  CK(cudaMemcpyAsync(&ctx->h_info, ctx->info.p, sizeof(PciIdsInfo), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->d_text = d_text;
  ctx->text_len = (uint32_t)len;
  ctx->h_pool.clear();
  ctx->sec_lines_ready = false;
  if (ctx->h_info.v_off != P_NONE) {
    size_t sec = (size_t)ctx->h_info.sec_end - ctx->h_info.v_off;
    if (!ctx->h_pool.resize(sec + 16)) {
      ctx->err = "cudaMallocHost failed for the name pool mirror";
      return KVG_ENOMEM;
    }
    CK(cudaMemcpyAsync(ctx->h_pool.data(), ctx->pool.p, sec + 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->table_ready = true;
  return KVG_OK;
}

// candidate lines of the general (arbitrary-key) lookup: every '\t' line of the section, collected on
// the first general lookup after a table load (most loads are never followed by one)
static int ensure_section_lines(kvg_ctx* ctx) {
  if (ctx->sec_lines_ready) return KVG_OK;
  size_t sec = ctx->h_info.v_off == P_NONE ? 0 : (size_t)ctx->h_info.sec_end - ctx->h_info.v_off;
  ctx->sec_lines_cap = sec / 2 + 8;  // a line needs >= 2 bytes ("\t\n")
  ENSURE(ctx->sec_lines, ctx->sec_lines_cap + 1);
  CK(cudaMemsetAsync(ctx->sec_lines.p + ctx->sec_lines_cap, 0, sizeof(uint32_t), ctx->stream));
  if (sec) {
    int grid = (int)((sec + KVG_BLOCK - 1) / KVG_BLOCK);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    LAUNCH("section_lines", k_section_lines, grid, KVG_BLOCK, 0, ctx->d_text, ctx->info.p, ctx->sec_lines.p,
           ctx->sec_lines.p + ctx->sec_lines_cap, (uint32_t)ctx->sec_lines_cap);
    int rc = check_launch(ctx, "section lines");
    if (rc) return rc;
  }
  ctx->sec_lines_ready = true;
  return KVG_OK;
}

static int parse_and_publish(kvg_ctx* ctx, const uint8_t* d_text, size_t len, size_t stride, uint32_t n_files,
                             bool enqueued = false) {
  if (!enqueued) {
    int rc = parse_enqueue(ctx, d_text, len, stride, n_files);
    if (rc) return rc;
  }
  return table_publish(ctx, d_text, len);
}

// complete a pending kvg_pciids_load (see load_pending), then require a published table
static int table_needed(kvg_ctx* ctx, const char* why_missing) {
  if (ctx->load_pending) {
    ctx->load_pending = false;
    int rc = parse_and_publish(ctx, ctx->text.p, ctx->pend_len, ctx->pend_stride, 1, true);
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
    ENSURE(ctx->nv_index, K1_IDS);
    ENSURE(ctx->pool, 16);
    ENSURE(ctx->text, 64);
    CK(cudaMemsetAsync(ctx->nv_index.p, 0xff, K1_IDS * sizeof(uint32_t), ctx->stream));  // no id has a name
    CK(cudaMemsetAsync(ctx->pool.p, 0, 16, ctx->stream));
    PciIdsInfo z;
    memset(&z, 0, sizeof z);
    z.v_off = z.sec_end = P_NONE;
    CK(cudaMemcpyAsync(ctx->info.p, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    return table_publish(ctx, ctx->text.p, 0);
  }
  size_t padded = kvg_text_pad(len);
  ENSURE(ctx->text, padded);
  CK(cudaMemsetAsync(ctx->text.p, '\n', padded, ctx->stream));
  // the caller's buffer must stay readable until the copy has run: pinned memory is copied
  // asynchronously (the caller keeps it alive until the next consuming call, as with any cudaMemcpyAsync
  // source), pageable memory is copied synchronously by the runtime's own staging
  CK(cudaMemcpyAsync(ctx->text.p, text, len, cudaMemcpyHostToDevice, ctx->stream));
  int rc = parse_enqueue(ctx, ctx->text.p, len, padded, 1);
  if (rc) return rc;
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) return parse_and_publish(ctx, ctx->text.p, len, padded, 1, true);
  ctx->load_pending = true;
  ctx->pend_len = len;
  ctx->pend_stride = padded;
  return KVG_OK;
}

int kvg_dev_pciids_parse(kvg_ctx* ctx, const void* d_text, size_t len, size_t stride, uint32_t n_files) {
  if (!ctx || !d_text || n_files == 0 || ((uintptr_t)d_text & 15) || (stride & 15) ||
      (n_files > 1 && stride < kvg_text_pad(len)))
    return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  // first call on an image publishes (host mirror of the section facts and the name pool); later calls on
  // the same image stay fully asynchronous: the whole K1 chain enqueued, no host sync
  if (ctx->load_pending) {
    int rc_t = table_needed(ctx, "");
    if (rc_t && rc_t != KVG_ESTATE) return rc_t;
  }
  if (!ctx->table_ready || ctx->d_text != d_text || ctx->text_len != len || ctx->parsed_files != n_files) {
    int rc0 = parse_and_publish(ctx, (const uint8_t*)d_text, len, stride, n_files);
    if (rc0 == KVG_OK) ctx->parsed_files = n_files;
    return rc0;
  }
  const bool lines_ready = ctx->sec_lines_ready;
  int rc = parse_enqueue(ctx, (const uint8_t*)d_text, len, stride, n_files);
  ctx->sec_lines_ready = lines_ready;  // same image: a candidate list collected for it stays valid
  return rc;
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
    int rc_s = ensure_section_lines(ctx);
    if (rc_s) return rc_s;
  }
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
  if (canonical_key(key, keylen, &v)) {  // table path: the id's slot, then the host mirror of the name pool
    uint32_t slot;
    CK(cudaMemcpyAsync(&slot, ctx->nv_index.p + v, 4, cudaMemcpyDeviceToHost, ctx->stream));
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
  std::vector<uint32_t> slots(count);
  CK(cudaMemcpyAsync(slots.data(), ctx->nv_index.p + first, 4 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
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
  ENSURE(o.heads_state, T + 2);
  return KVG_OK;
}

}  // extern "C"

// Both stable orderings of the survivor list — ordering 0 by device id / mdev type (<= 16 bits, 2 passes),
// ordering 1 by iommu group / parent (32 bits, up to 4 passes) — share their launches: grid.y = 2 while
// both have a pass, grid.y = 1 (ordering 1 only, passed in slot 0) afterwards.  The host always enqueues
// the launch sets a 32-bit key could need; a set whose pass does not exist for the keys at hand returns at
// once (small persistent grid).
struct OrdInput {
  const void* recs[2];     // record list each ordering reads its pass-0 keys from
  uint32_t* cnt[2];        // its length (device)
  int src[2];              // SRC_*
  uint32_t* max_key[2];    // its largest key (device): decides the digit plan
  const uint4* head_surv;  // PCI device-id ordering: records whose name slot the segment heads publish (or NULL)
};
// cap: the most elements an ordering can have (buffers; the grids of the kernels that need one CTA per tile).
// expect: how many it is expected to have — the launch shapes (digit width, grids of the tile loops, final
// variant) are chosen for it; every kernel is correct for any count up to cap.  The two differ in the
// sharded scan, whose owned lists can hold P shards but normally hold about one.
static int enqueue_orderings(kvg_ctx* ctx, size_t cap, const OrdInput& in, size_t expect = 0) {
  if (expect == 0 || expect > cap) expect = cap;
  int rc = ensure_order(ctx, ctx->ord_dev, cap);
  if (rc) return rc;
  rc = ensure_order(ctx, ctx->ord_grp, cap);
  if (rc) return rc;
  size_t T = (cap + C_TILE - 1) / C_TILE;
  if (T == 0) T = 1;
  const size_t Te = std::max<size_t>(1, (expect + C_TILE - 1) / C_TILE);
  ENSURE(ctx->tile_hist, 2 * (size_t)RADIX_MAX_DIGITS * T);
  ENSURE(ctx->bin_total, 2 * (size_t)RADIX_MAX_DIGITS);
  ScanCtrl* c = ctx->ctrl.p;
  OrderBufs* ob[2] = {&ctx->ord_dev, &ctx->ord_grp};
  uint32_t* const* maxk = in.max_key;
  uint32_t* const* cnt = in.cnt;
  const uint32_t key_bits[2] = {16, 32};  // device id / mdev type: u16; iommu group / parent: u32
  // digit width: up to 11 bits where the scan is latency-bound (fewer passes: 19-bit groups sort in 2),
  // 8 bits for inputs large enough to be bandwidth-bound (half the shared memory, more CTAs/SM)
  const bool big = expect >= (8u << 20);
  const uint32_t max_bits = big ? 8 : RADIX_MAX_BITS;
  const int nsets[2] = {(int)((key_bits[0] + max_bits - 1) / max_bits),
                        (int)((key_bits[1] + max_bits - 1) / max_bits)};  // 2 and 3 (or 4) launch sets
  auto fill = [&](int ord, int p) {
    OrdArgs a;
    a.n_ptr = cnt[ord];
    a.max_key = maxk[ord];
    a.src_records = in.recs[ord];
    a.pairs_in = p == 0 ? nullptr : ((p & 1) ? ob[ord]->p0.p : ob[ord]->p1.p);
    a.pairs_out = (p & 1) ? ob[ord]->p1.p : ob[ord]->p0.p;
    a.tile_hist = ctx->tile_hist.p + (size_t)ord * RADIX_MAX_DIGITS * T;
    a.bin_total = ctx->bin_total.p + (size_t)ord * RADIX_MAX_DIGITS;
    a.pass = p < nsets[ord] ? (uint32_t)p : 0xffu;
    a.key_bits_max = key_bits[ord];
    a.max_bits = max_bits;
    a.src = p == 0 ? in.src[ord] : SRC_PAIRS;
    return a;
  };
  if (!ctx->scatter_smem_set) {  // a function attribute is per device: remember it per context
    CK(cudaFuncSetAttribute(k_order_scatter<RADIX_MAX_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)OrdScatterCfg<RADIX_MAX_BITS>::SMEM));
    ctx->scatter_smem_set = true;
  }
  for (int p = 0; p < nsets[1]; p++) {
    OrdArgs2 aa;
    // sets 0/1 always have work: one CTA per tile.  Later sets exist only for wide keys and are usually
    // ruled out on the device: a small persistent grid makes a ruled-out pass nearly free.
    const bool both = p < nsets[0];
    const unsigned gx = p < 2 ? (unsigned)Te : (unsigned)std::min<size_t>(Te, (size_t)ctx->sm_count * (big ? 5 : 3));
    dim3 grid(gx, both ? 2 : 1);
    if (both) {
      aa.o[0] = fill(0, p);
      aa.o[1] = fill(1, p);
    } else {
      aa.o[0] = fill(1, p);
      aa.o[1] = aa.o[0];
    }
    LAUNCH("order_hist", k_order_hist, grid, KVG_BLOCK, 0, aa);
    if (Te > 2048) {
      dim3 sgrid(1u << max_bits, grid.y);
      LAUNCH("order_tilescan", k_order_tilescan_long, sgrid, KVG_BLOCK, 0, aa);
    } else {
      dim3 sgrid((1u << max_bits) / TS_WARPS, grid.y);
      LAUNCH("order_tilescan", k_order_tilescan, sgrid, TS_WARPS * 32, 0, aa);
    }
    if (big)
      LAUNCH("order_scatter", k_order_scatter<8>, grid, KVG_BLOCK, OrdScatterCfg<8>::SMEM, aa);
    else
      LAUNCH("order_scatter", k_order_scatter<RADIX_MAX_BITS>, grid, KVG_BLOCK, OrdScatterCfg<RADIX_MAX_BITS>::SMEM, aa);
  }
  // final permutation + distinct keys of both orderings
  OrdFinalArgs2 ff;
  TileOffsetsArgs2 tt;
  uint32_t* nseg[2] = {&c->n_dev_keys, &c->n_groups};
  for (int ord = 0; ord < 2; ord++) {
    OrdFinalArgs& a = ff.o[ord];
    a.p0 = ob[ord]->p0.p;
    a.p1 = ob[ord]->p1.p;
    a.max_key = maxk[ord];
    a.key_bits_max = key_bits[ord];
    a.max_bits = max_bits;
    a.n_ptr = cnt[ord];
    a.perm = ob[ord]->perm.p;
    a.state = ob[ord]->heads_state.p;
    a.tile_heads = ob[ord]->tile_heads.p;
    a.tile_off = ob[ord]->tile_off.p;
    a.seg_key = ob[ord]->seg_key.p;
    a.seg_off = ob[ord]->seg_off.p;
    a.n_seg = nseg[ord];
    // PCI device-id ordering: the bucket's joined name slot comes back with the keys
    a.head_surv = ord == 0 ? in.head_surv : nullptr;
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
  // the final kernels loop over the tiles: grids for the EXPECTED length (every tile of a longer list is still
  // visited); the chained-scan form must also fit the GPU at once (half of it: two orderings share the launch)
  dim3 fgrid((unsigned)Te, 2);
  if (expect < (2u << 20)) {  // latency-bound: one launch (chained scan of the head counts)
    if (!ctx->final_ctas_per_sm) {
      int nb = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_order_final, KVG_BLOCK, 0));
      ctx->final_ctas_per_sm = std::max(1, nb);
    }
    fgrid.x = (unsigned)std::min<size_t>(Te, (size_t)std::max(1, ctx->sm_count) * ctx->final_ctas_per_sm / 2);
    LAUNCH("order_final", k_order_final, fgrid, KVG_BLOCK, 0, ff, next_epoch());
  } else {                 // bandwidth-bound: no CTA waits for another
    LAUNCH("order_count", k_order_heads<false>, fgrid, KVG_BLOCK, 0, ff);
    dim3 ogrid((unsigned)((T + C_TILE - 1) / C_TILE), 2);
    LAUNCH("tile_offsets", k_tile_offsets, ogrid, KVG_BLOCK, 0, tt, c, next_epoch());
    LAUNCH("order_emit", k_order_heads<true>, fgrid, KVG_BLOCK, 0, ff);
  }
  return check_launch(ctx, "orderings");
}

static int enqueue_pci_orderings(kvg_ctx* ctx, size_t surv_cap) {
  ScanCtrl* c = ctx->ctrl.p;
  OrdInput in = {{ctx->surv.p, ctx->surv.p}, {&c->n_surv, &c->n_surv}, {SRC_PCI_DEVICE, SRC_PCI_GROUP},
                 {&c->max_devkey, &c->max_group}, ctx->surv.p};
  return enqueue_orderings(ctx, surv_cap, in);
}

// K3 into d_out (dense, Walk order).  Below ~2 M records the step is launch-bound and ONE look-back kernel
// beats the three launches of the split form; above, the split form (no cross-CTA wait) runs at the HBM
// roofline.
static int enqueue_classify(kvg_ctx* ctx, const void* d_recs, size_t n, uint4* d_out) {
  PciClassifyOp op;
  op.recs = (const uint4*)d_recs;
  op.n = (uint32_t)n;
  op.out = (kvg_pci_surv*)d_out;
  op.ctrl = ctx->ctrl.p;
  op.nv_index = ctx->nv_index.p;
  op.local_max_group = 0;
  op.local_max_dev = 0;
  constexpr int T = 128, R = 8;
  const size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
  if (tiles == 0) return KVG_OK;  // nothing to classify: n_surv stays 0 from the control-block memset
  if (n < (2u << 20)) {
    ENSURE(ctx->classify_state, tiles + 2);
    LAUNCH("classify_compact", (k_classify_oneshot<PciClassifyOp, T, R>), (unsigned)tiles, T, 0, op,
           ctx->classify_state.p, next_epoch());
  } else {
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
    LAUNCH("pack_survivors", k_pack_survivors<1>, (unsigned)tiles, 128, 0, (const uint4*)ctx->ragged.p,
           (const uint32_t*)ctx->tile_off.p, (uint32_t)(T * R), d_out);
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
  const size_t SD = S, SG = S;
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
  if (n && n <= HEALTH_SMALL_MAX && !ctx->timing) {
    // poll-loop sizes: one kernel reads the snapshot in place (mapped pinned memory) and writes the transition
    // list and the counters straight into the host-visible result block; one synchronisation
    if (ctx->health_n != n) {
      ENSURE(ctx->alive_prev, n + 1);
      CK(cudaMemsetAsync(ctx->alive_prev.p, 0, n + 1, ctx->stream));
      ctx->health_n = n;
    }
    static const bool zero_copy = [] {  // KVG_HEALTH_ZEROCOPY=0: DMA copy first, device-resident reads (A/B)
      const char* e = getenv("KVG_HEALTH_ZEROCOPY");
      return !(e && e[0] == '0');
    }();
    if (!ctx->health_smem_set) {
      CK(cudaFuncSetAttribute(k_health_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEALTH_SMALL_SMEM));
      ctx->health_smem_set = true;
    }
    cudaPointerAttributes attr;
    const void* dev_view = nullptr;
    if (cudaPointerGetAttributes(&attr, recs) == cudaSuccess && attr.type == cudaMemoryTypeHost) dev_view = attr.devicePointer;
    cudaGetLastError();
    if (!zero_copy) {  // one DMA copy (pageable memory is staged in pinned memory first), then device-resident reads
      ENSURE(ctx->recs, n + 1);
      int rc_c = stage_h2d(ctx, recs, n * sizeof(kvg_pci_rec), ctx->recs.p);
      if (rc_c) return rc_c;
      dev_view = ctx->recs.p;
    } else if (!dev_view) {  // pageable caller memory (cgo rule: never keep the pointer): stage it in pinned memory
      if (ctx->h_stage_cap < n * 16) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_cap = 0;
        CK(cudaMallocHost(&ctx->h_stage, n * 16 + 4096));
        ctx->h_stage_cap = n * 16 + 4096;
      }
      memcpy(ctx->h_stage, recs, n * 16);
      dev_view = ctx->h_stage;
    }
    const size_t o_list = align64(sizeof(kvg_health_delta)) + 64;  // [header][2 counters, padded][list]
    void* blk = pinned_alloc(ctx, o_list + align64(n * 4));
    if (!blk) return KVG_ENOMEM;
    uint8_t* b = pinned_payload(blk);
    uint32_t* hdr = (uint32_t*)(b + align64(sizeof(kvg_health_delta)));
    const uint32_t seq = ++ctx->health_seq ? ctx->health_seq : ++ctx->health_seq;  // never 0
    ((volatile uint32_t*)hdr)[2] = 0;
    LAUNCH("health_diff", k_health_small, 1, HEALTH_SMALL_THREADS, HEALTH_SMALL_SMEM, (const uint4*)dev_view, (uint32_t)n,
           ctx->alive_prev.p, (uint32_t*)(b + o_list), hdr, seq);
    int rc = check_launch(ctx, "health");
    if (rc == KVG_OK) {
      // the kernel's last store is a flag in this (mapped, pinned) block: poll it instead of paying a driver
      // synchronisation per tick; a launch that never completes falls back to the stream after ~1 s
      const auto t0 = std::chrono::steady_clock::now();
      uint32_t spins = 0;
      while (((volatile uint32_t*)hdr)[2] != seq) {
        if ((++spins & 0xffffu) == 0 &&
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 1.0) {
          if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || ((volatile uint32_t*)hdr)[2] != seq) {
            ctx->err = "health re-scan failed";
            rc = KVG_ECUDA;
          }
          break;
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (rc) {
      ctx->pinned_free.push_back({blk, (size_t)((uint64_t*)blk)[1]});
      return rc;
    }
    kvg_health_delta* d = (kvg_health_delta*)b;
    d->n_records = (uint32_t)n;
    d->n_alive = hdr[0];
    d->n_changed = hdr[1];
    d->changed = (const uint32_t*)(b + o_list);
    *delta = d;
    ctx->last_kind = 0;
    return KVG_OK;
  }
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

// K5 up to the dense survivor list (ctx->surv, 2 x 16 bytes per mdev): type dictionary (labels, canonical
// ids, the resource-name join of every label) + classification + stable compaction
// fused: fill *fused with the classify operator and stop in front of the classify kernels (the sharded scan
// classifies and sends in one kernel)
static int mdev_classify(kvg_ctx* ctx, const void* d_recs, size_t n, const kvg_type_dict* types,
                         MdevClassifyOp* fused = nullptr) {
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
  if (fused) {
    *fused = op;
    return check_launch(ctx, "mdev dictionary");
  }
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
      LAUNCH("pack_survivors", k_pack_survivors<2>, (unsigned)tiles, 128, 0, (const uint4*)ctx->ragged.p,
             (const uint32_t*)ctx->tile_off.p, (uint32_t)(T * R), dense);
    }
  }
  return check_launch(ctx, "mdev classify");
}

int kvg_dev_scan_mdev(kvg_ctx* ctx, const void* d_recs, size_t n, const kvg_type_dict* types) {
  if (!ctx || !types || (!d_recs && n) || n > 0xfffffff0ull || ((uintptr_t)d_recs & 15)) return KVG_EINVAL;
  int rc = mdev_classify(ctx, d_recs, n, types);
  if (rc) return rc;
  {
    ScanCtrl* c = ctx->ctrl.p;
    OrdInput in = {{ctx->surv.p, ctx->surv.p}, {&c->n_surv, &c->n_surv}, {SRC_MDEV_TYPE, SRC_MDEV_PARENT},
                   {&c->max_devkey, &c->max_group}, nullptr};
    rc = enqueue_orderings(ctx, n, in);
  }
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


// ---- peer windows: set-up -------------------------------------------------------------------------
// kvg_comm_p2p_export: allocate this rank's receive window for shards of up to cap_local PCI records
// (cap_local / 2 mdev records) and return the 64-byte CUDA IPC handle the other ranks need.
// kvg_comm_p2p_import: open every rank's handle (all_handles = nranks x 64 bytes, rank order).  After both
// succeeded on EVERY rank (kvg_comm_p2p_enable) the sharded scans exchange over NVLink stores; otherwise the
// NCCL path (kvg_comm_init) allocates a private window of the same shape on first use.
static size_t window_bytes(int nranks, size_t cap_local) {
  return SH_HDR + 2 * 2 * (size_t)nranks * cap_local * 16;
}
int kvg_comm_p2p_export(kvg_ctx* ctx, int rank, int nranks, size_t cap_local, void* handle_out64) {
  if (!ctx || !handle_out64 || nranks < 1 || nranks > SH_MAX_RANKS || rank < 0 || rank >= nranks || !cap_local)
    return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  static_assert(sizeof(ShardCtrl) <= SH_HDR, "control block fits its pad");
  if (ctx->win_mine) {
    ctx->err = "receive window already allocated";
    return KVG_ESTATE;
  }
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->win_cap = cap_local;
  CK(cudaMalloc((void**)&ctx->win_mine, window_bytes(nranks, cap_local)));
  CK(cudaMemset(ctx->win_mine, 0, SH_HDR));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->win_mine));
  memcpy(handle_out64, &h, 64);
  return KVG_OK;
}

int kvg_comm_p2p_import(kvg_ctx* ctx, const void* all_handles) {
  if (!ctx || !all_handles || !ctx->win_mine) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  for (int q = 0; q < ctx->nranks; q++) {
    if (q == ctx->rank) {
      ctx->win_peer[q] = ctx->win_mine;
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
    ctx->win_peer[q] = (uint8_t*)p;
  }
  ctx->shard_step = 0;
  return KVG_OK;
}

// collective decision: enable only when import succeeded on every rank
int kvg_comm_p2p_enable(kvg_ctx* ctx, int on) {
  if (!ctx) return KVG_EINVAL;
  if (on) {
    if (!ctx->win_mine) {
      ctx->err = "peer windows are not exported";
      return KVG_ESTATE;
    }
    for (int q = 0; q < ctx->nranks; q++)
      if (!ctx->win_peer[q]) {
        ctx->err = "peer windows are not imported";
        return KVG_ESTATE;
      }
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

// NCCL mode: every rank learns every shard's survivor count, then one allgatherv (NCCL has none: a grouped
// broadcast per root) of U x 16-byte records; rank order == Walk order.  Returns the gathered total.
static int nccl_allgatherv(kvg_ctx* ctx, const uint4* local, int U, size_t* total_out) {
  const int P = ctx->nranks;
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->h_counts[P] = ctx->h_ctrl->n_surv;
  CK(cudaMemcpyAsync(ctx->gather_counts.p + P, &ctx->h_counts[P], 8, cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = g_nccl.AllGather(ctx->gather_counts.p + P, ctx->gather_counts.p, 1, ncclUint64, ctx->comm,
                                    ctx->stream);
  if (r != 0) {
    ctx->err = std::string("ncclAllGather(counts): ") + g_nccl.GetErrorString(r);
    return KVG_ENCCL;
  }
  CK(cudaMemcpyAsync(ctx->h_counts, ctx->gather_counts.p, 8 * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  size_t total = 0;
  std::vector<size_t> displ((size_t)P);
  for (int q = 0; q < P; q++) {
    displ[(size_t)q] = total;
    total += ctx->h_counts[q];
  }
  if (total > 0x7ffffff0ull) {
    ctx->err = "gathered survivor list exceeds 2^31 entries";
    return KVG_ERANGE;
  }
  ENSURE(ctx->gathered, (total + 1) * (size_t)U);
  r = g_nccl.GroupStart();
  for (int root = 0; root < P && r == 0; root++) {
    size_t cnt = ctx->h_counts[root];
    if (cnt == 0) continue;
    r = g_nccl.Broadcast(root == ctx->rank ? (const void*)local : nullptr,
                         ctx->gathered.p + displ[(size_t)root] * U, cnt * 16 * U, ncclUint8, root, ctx->comm,
                         ctx->stream);
  }
  ncclResult_t r2 = g_nccl.GroupEnd();
  if (r != 0 || r2 != 0) {
    ctx->err = std::string("ncclBroadcast group: ") + g_nccl.GetErrorString(r ? r : r2);
    return KVG_ENCCL;
  }
  *total_out = total;
  return KVG_OK;
}

// The exchange step behind a dense local survivor list of <= n_cap records (U x 16 bytes each) whose length
// lives in ctrl->n_surv: multisplit by owner -> windows -> owned lists (ctx->owned0 / owned1, lengths in
// ctrl->n_own[], largest keys in ctrl->max_devkey / max_group).  Returns the capacity of an owned list.
// fused (peer transport, latency-bound shard): *fused is the classify operator of the shard — the records are
// classified and sent by ONE kernel (k_classify_send) and `local` is the dense list it also writes.
constexpr size_t FUSED_SEND_MAX = 2u << 20;  // records per shard up to which the T^2 prefix of k_classify_send is free
template <int U, class Op, int ROWS>
static int enqueue_exchange(kvg_ctx* ctx, const uint4* local, size_t n_cap, size_t* owned_cap_out, Op* fused = nullptr) {
  const int P = ctx->nranks;
  const bool peer = ctx->p2p;
  if (!peer && !ctx->comm) {
    ctx->err = "kvg_comm_init / kvg_comm_p2p_import has not been called";
    return KVG_ESTATE;
  }
  if (fused && !peer) return KVG_ESTATE;  // callers fuse only over the peer transport
  ScanCtrl* c = ctx->ctrl.p;
  const uint4* list = local;
  size_t list_cap = n_cap;
  uint32_t* n_ptr = &c->n_surv;
  if (!peer) {  // NCCL: all-gather, then the same kernels in local mode on the gathered list
    size_t total = 0;
    int rc = nccl_allgatherv(ctx, local, U, &total);
    if (rc) return rc;
    list = ctx->gathered.p;
    list_cap = total;
    ctx->h_counts[P] = total;  // pinned; the low 32 bits are the value (little endian)
    CK(cudaMemcpyAsync(&c->n_gathered, &ctx->h_counts[P], sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    n_ptr = &c->n_gathered;
    if (!ctx->win_mine || n_cap * U > ctx->win_cap) {  // private window of the shape a peer window has, grown on demand
      CK(cudaStreamSynchronize(ctx->stream));
      if (ctx->win_mine) cudaFree(ctx->win_mine);
      ctx->win_mine = nullptr;
      size_t cap = std::max<size_t>(ctx->win_cap, n_cap * U);
      cap += cap / 4 + 64;
      CK(cudaMalloc((void**)&ctx->win_mine, window_bytes(P, cap)));
      ctx->win_cap = cap;
      CK(cudaMemsetAsync(ctx->win_mine, 0, SH_HDR, ctx->stream));
      ctx->win_peer[ctx->rank] = ctx->win_mine;
    }
  }
  if (n_cap * U > ctx->win_cap) {
    ctx->err = "shard larger than the receive window (kvg_comm_p2p_export cap_local)";
    return KVG_ERANGE;
  }
  // an owned list holds at most what the window regions of one ordering hold
  const size_t owned_cap = (size_t)P * ctx->win_cap / U;
  ENSURE(ctx->owned0, (owned_cap + 1) * U);
  ENSURE(ctx->owned1, (owned_cap + 1) * U);
  const size_t T = (list_cap + C_TILE - 1) / C_TILE + 1;
  // shard_cnt: [0, 2P) totals, [32, 34) self-resetting tickets, [34] error word; a fresh allocation is
  // zero-filled, the tickets return to zero by themselves, the error word is sticky.  shard_state: the
  // chained-scan words of the send kernel (epoch-tagged: never cleared)
  ENSURE(ctx->shard_cnt, 64);
  ENSURE(ctx->shard_state, 2 * (size_t)P * T);
  uint32_t* totals = ctx->shard_cnt.p;
  uint32_t* tickets = ctx->shard_cnt.p + 32;
  uint32_t* err = ctx->shard_cnt.p + 34;
  ctx->shard_err = err;
  const unsigned long long step = ++ctx->shard_step;
  ShardArgs A;
  A.list = list;
  A.n_ptr = n_ptr;
  A.state = ctx->shard_state.p;
  A.totals = totals;
  A.ticket = tickets;
  A.T = (uint32_t)T;
  A.P = (uint32_t)P;
  A.Pm = shard_magic((uint32_t)P);
  A.me = (uint32_t)ctx->rank;
  A.only = peer ? SH_ALL : (uint32_t)ctx->rank;
  A.n_src = peer ? (uint32_t)P : 1u;
  A.src = peer ? (uint32_t)ctx->rank : 0u;
  A.region_cap = peer ? ctx->win_cap / U : (size_t)P * ctx->win_cap / U;
  A.parity = (uint32_t)(step & 1);
  A.step = step;
#ifdef KVG_EXP
  {
    const char* e = getenv("KVG_SHARD_EXP");
    A.exp = e ? (uint32_t)atoi(e) : 0u;
  }
#endif
  ShardPeers peers;
  memset(&peers, 0, sizeof peers);
  for (int q = 0; q < P; q++) {
    uint8_t* base = ctx->win_peer[q] ? ctx->win_peer[q] : ctx->win_mine;  // local mode: only my own entry is used
    peers.ctrl[q] = (ShardCtrl*)base;
    peers.win[q] = (uint4*)(base + SH_HDR);
  }
  const ShardCtrl* mine = (const ShardCtrl*)ctx->win_mine;
  if (fused) {
    constexpr int TH = 128;
    const size_t tiles = std::max<size_t>(1, (n_cap + (size_t)TH * ROWS - 1) / ((size_t)TH * ROWS));
    const int C = 1 + 2 * P;
    const int cw4 = C <= 8 ? 2 : C <= 16 ? 4 : C <= 20 ? 5 : 9;
    ENSURE(ctx->send_words, tiles * 4 * (size_t)cw4);
    if (ctx->send_epoch == 0 || ctx->send_cw4 != cw4) {  // the epoch field wrapped (or the row shape changed): start over
      CK(cudaMemsetAsync(ctx->send_words.p, 0, ctx->send_words.cap * sizeof(uint32_t), ctx->stream));
      ctx->send_epoch = 0;
      ctx->send_cw4 = cw4;
    }
    const uint32_t ep = ++ctx->send_epoch;
    if (ctx->send_epoch == (1u << (32 - CS_COUNT_BITS)) - 1) ctx->send_epoch = 0;
    switch (cw4) {
      case 2: LAUNCH("classify_send", (k_classify_send<Op, TH, ROWS, 2>), (unsigned)tiles, TH, 0, *fused, A, peers, mine, err, ctx->send_words.p, ep); break;
      case 4: LAUNCH("classify_send", (k_classify_send<Op, TH, ROWS, 4>), (unsigned)tiles, TH, 0, *fused, A, peers, mine, err, ctx->send_words.p, ep); break;
      case 5: LAUNCH("classify_send", (k_classify_send<Op, TH, ROWS, 5>), (unsigned)tiles, TH, 0, *fused, A, peers, mine, err, ctx->send_words.p, ep); break;
      default: LAUNCH("classify_send", (k_classify_send<Op, TH, ROWS, 9>), (unsigned)tiles, TH, 0, *fused, A, peers, mine, err, ctx->send_words.p, ep); break;
    }
  } else {
    LAUNCH("shard_send", k_shard_send<U>, (unsigned)T, KVG_BLOCK, 0, A, peers, mine, err, next_epoch());
  }
  GatherArgs G;
  G.window = (const uint4*)(ctx->win_mine + SH_HDR);
  G.owned[0] = ctx->owned0.p;
  G.owned[1] = ctx->owned1.p;
  G.n_own = &c->n_own[0];
  G.max_key = &c->own_max[0];
  dim3 ggrid((unsigned)std::max(1, ctx->sm_count), 2);
  LAUNCH("shard_gather", k_shard_gather<U>, ggrid, KVG_BLOCK, 0, A, G, peers, mine, err, 3u);
  *owned_cap_out = owned_cap;
  return check_launch(ctx, "shard exchange");
}

// orderings of the two owned lists
static int enqueue_owned_orderings(kvg_ctx* ctx, size_t owned_cap, size_t n_local, int src0, int src1, bool names) {
  ScanCtrl* c = ctx->ctrl.p;
  OrdInput in = {{ctx->owned0.p, ctx->owned1.p}, {&c->n_own[0], &c->n_own[1]}, {src0, src1},
                 {&c->own_max[0], &c->own_max[1]}, names ? ctx->owned0.p : nullptr};
  // keys spread over the owners evenly (key % P): an owned list is about as long as the local survivor list
  return enqueue_orderings(ctx, owned_cap, in, std::max<size_t>(n_local, C_TILE));
}

extern "C" {

int kvg_dev_scan_pci_sharded(kvg_ctx* ctx, const void* d_recs, size_t n_local) {
  if (!ctx || (!d_recs && n_local) || n_local > 0x7ffffff0ull || ((uintptr_t)d_recs & 15)) return KVG_EINVAL;
  {
    int rc_t = table_needed(ctx, "kvg_pciids_load must precede a scan");
    if (rc_t) return rc_t;
  }
  CK(cudaSetDevice(ctx->device));
  ENSURE(ctx->surv, n_local + 1);
  CK(cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(ScanCtrl), ctx->stream));
  size_t owned_cap = 0;
  int rc;
  bool fuse = ctx->p2p && n_local < FUSED_SEND_MAX;  // classify + send in one kernel
#ifdef KVG_EXP
  if (const char* e = getenv("KVG_SHARD_EXP")) fuse = fuse && !(atoi(e) & 8);
#endif
  if (fuse) {
    PciClassifyOp op;
    op.recs = (const uint4*)d_recs;
    op.n = (uint32_t)n_local;
    op.out = (kvg_pci_surv*)ctx->surv.p;
    op.ctrl = ctx->ctrl.p;
    op.nv_index = ctx->nv_index.p;
    op.local_max_group = 0;
    op.local_max_dev = 0;
    rc = enqueue_exchange<1, PciClassifyOp, 8>(ctx, ctx->surv.p, n_local, &owned_cap, &op);
  } else {
    rc = enqueue_classify(ctx, d_recs, n_local, ctx->surv.p);
    if (rc) return rc;
    rc = enqueue_exchange<1, PciClassifyOp, 8>(ctx, ctx->surv.p, n_local, &owned_cap);
  }
  if (rc) return rc;
  rc = enqueue_owned_orderings(ctx, owned_cap, n_local, SRC_PCI_DEVICE, SRC_PCI_GROUP, true);
  if (rc) return rc;
  ctx->last_n = n_local;
  ctx->last_total = owned_cap;
  ctx->last_kind = 3;
  ctx->last_units = 1;
  return KVG_OK;
}

}  // extern "C"

// ---- sharded fetches ---------------------------------------------------------------------------------
namespace {
struct BlockLayout {
  size_t o = 0;
  size_t take(size_t bytes) {
    size_t at = o;
    o += (bytes + 63) & ~(size_t)63;
    return at;
  }
};
// the type dictionary part of an mdev result: layout, device-to-host copies, host-side repack
struct DictPart {
  size_t o_lraw, o_llen, o_loff, o_lbytes, o_canon, o_nraw, o_nlen, o_noff, o_nbytes;
  static constexpr uint32_t NAME_CAP = 256;
  void layout(BlockLayout& L, uint32_t nt, size_t raw_len) {
    o_lraw = L.take(raw_len + 16);
    o_llen = L.take(((size_t)nt + 1) * 4);
    o_loff = L.take(((size_t)nt + 1) * 4);
    o_lbytes = L.take(raw_len + 16);
    o_canon = L.take(((size_t)nt + 1) * 2);
    o_nraw = L.take((size_t)nt * NAME_CAP + 16);
    o_nlen = L.take(((size_t)nt + 1) * 4);
    o_noff = L.take(((size_t)nt + 1) * 4);
    o_nbytes = L.take((size_t)nt * NAME_CAP + 16);
  }
};
}  // namespace

static int dict_copy(kvg_ctx* ctx, uint8_t* b, const DictPart& D) {
  const uint32_t nt = ctx->n_types;
  if (!nt) return KVG_OK;
  const size_t raw_len = ctx->h_type_off[nt];
  CK(cudaMemcpyAsync(b + D.o_lraw, ctx->type_label.p, raw_len, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(b + D.o_llen, ctx->type_label_len.p, (size_t)nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(b + D.o_canon, ctx->type_canon.p, (size_t)nt * 2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(b + D.o_nraw, ctx->type_names.p, (size_t)nt * DictPart::NAME_CAP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(b + D.o_nlen, ctx->type_name_len.p, (size_t)nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return KVG_OK;
}
// repack labels / names contiguously (marshalling of GPU-produced bytes)
static void dict_finish(kvg_ctx* ctx, uint8_t* b, const DictPart& D, uint32_t* n_types, const uint32_t** label_off,
                        const uint8_t** label_bytes, const uint16_t** type_canon, const uint32_t** name_off,
                        const uint8_t** name_bytes) {
  const uint32_t nt = ctx->n_types;
  uint32_t* loff = (uint32_t*)(b + D.o_loff);
  uint32_t* noff = (uint32_t*)(b + D.o_noff);
  const uint32_t* llen = (const uint32_t*)(b + D.o_llen);
  const uint32_t* nlen = (const uint32_t*)(b + D.o_nlen);
  size_t lo = 0, no = 0;
  for (uint32_t k = 0; k < nt; k++) {
    loff[k] = (uint32_t)lo;
    memcpy(b + D.o_lbytes + lo, b + D.o_lraw + ctx->h_type_off[k], llen[k]);
    lo += llen[k];
    noff[k] = (uint32_t)no;
    uint32_t nl = nlen[k] > DictPart::NAME_CAP ? DictPart::NAME_CAP : nlen[k];
    memcpy(b + D.o_nbytes + no, b + D.o_nraw + (size_t)k * DictPart::NAME_CAP, nl);
    no += nl;
  }
  loff[nt] = (uint32_t)lo;
  noff[nt] = (uint32_t)no;
  *n_types = nt;
  *label_off = loff;
  *label_bytes = b + D.o_lbytes;
  *type_canon = (const uint16_t*)(b + D.o_canon);
  *name_off = noff;
  *name_bytes = b + D.o_nbytes;
}

// control block + the exchange's error word; a spin that timed out means a peer never delivered
static int shard_ctrl_fetch(kvg_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&ctx->h_ctrl->reserved2[0], ctx->shard_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->h_ctrl->reserved2[0]) {
    ctx->err = "sharded scan: a peer did not deliver its regions (or acknowledge a window) within the time limit";
    return KVG_ENCCL;
  }
  return KVG_OK;
}

extern "C" {

int kvg_dev_scan_pci_shard_fetch(kvg_ctx* ctx, kvg_pci_shard_result** res) {
  if (!ctx || !res || ctx->last_kind != 3) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  int rc = shard_ctrl_fetch(ctx);
  if (rc) return rc;
  const size_t S = ctx->h_ctrl->n_surv, SD = ctx->h_ctrl->n_own[0], SG = ctx->h_ctrl->n_own[1];
  const size_t KD = ctx->h_ctrl->n_dev_keys, G = ctx->h_ctrl->n_groups;
  const size_t pool_len = ctx->h_pool.size();
  BlockLayout L;
  L.take(sizeof(kvg_pci_shard_result));
  const size_t o_local = L.take(S * 16), o_dm = L.take(SD * 16), o_dk32 = L.take(KD * 4), o_dk = L.take(KD * 2);
  const size_t o_doff = L.take((KD + 1) * 4), o_dperm = L.take(SD * 4), o_dname = L.take(KD * 4);
  const size_t o_gm = L.take(SG * 16), o_gk = L.take(G * 4), o_goff = L.take((G + 1) * 4), o_gperm = L.take(SG * 4);
  const size_t o_pool = L.take(pool_len);
  void* blk = pinned_alloc(ctx, L.o);
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
  CK(D2H(o_local, ctx->surv.p, S * 16));
  CK(D2H(o_dm, ctx->owned0.p, SD * 16));
  CK(D2H(o_dk32, od.seg_key.p, KD * 4));
  CK(D2H(o_doff, od.seg_off.p, (KD + 1) * 4));
  CK(D2H(o_dperm, od.perm.p, SD * 4));
  CK(D2H(o_dname, od.seg_name.p, KD * 4));
  CK(D2H(o_gm, ctx->owned1.p, SG * 16));
  CK(D2H(o_gk, og.seg_key.p, G * 4));
  CK(D2H(o_goff, og.seg_off.p, (G + 1) * 4));
  CK(D2H(o_gperm, og.perm.p, SG * 4));
  CK(cudaStreamSynchronize(ctx->stream));
  kvg_pci_shard_result* r = (kvg_pci_shard_result*)b;
  memset(r, 0, sizeof *r);
  if (SD == 0) ((uint32_t*)(b + o_doff))[0] = 0;
  if (SG == 0) ((uint32_t*)(b + o_goff))[0] = 0;
  uint16_t* dk = (uint16_t*)(b + o_dk);
  for (size_t k = 0; k < KD; k++) dk[k] = (uint16_t)((const uint32_t*)(b + o_dk32))[k];  // marshalling: narrow
  r->n_records = ctx->last_n;
  r->n_local = S;
  r->local = (const kvg_pci_surv*)(b + o_local);
  r->n_dev_members = SD;
  r->dev_members = (const kvg_pci_surv*)(b + o_dm);
  r->n_dev_keys = (uint32_t)KD;
  r->dev_keys = dk;
  r->dev_off = (const uint32_t*)(b + o_doff);
  r->dev_perm = (const uint32_t*)(b + o_dperm);
  r->dev_name_slot = (const uint32_t*)(b + o_dname);
  r->n_grp_members = SG;
  r->grp_members = (const kvg_pci_surv*)(b + o_gm);
  r->n_groups = (uint32_t)G;
  r->grp_keys = (const uint32_t*)(b + o_gk);
  r->grp_off = (const uint32_t*)(b + o_goff);
  r->grp_perm = (const uint32_t*)(b + o_gperm);
  if (pool_len) memcpy(b + o_pool, ctx->h_pool.data(), pool_len);
  r->name_pool = b + o_pool;
  r->name_pool_len = pool_len;
  *res = r;
  return KVG_OK;
}

int kvg_dev_scan_mdev_sharded(kvg_ctx* ctx, const void* d_recs, size_t n_local, const kvg_type_dict* types) {
  if (!ctx || !types || (!d_recs && n_local) || n_local > 0x7ffffff0ull || ((uintptr_t)d_recs & 15)) return KVG_EINVAL;
  size_t owned_cap = 0;
  int rc;
  if (ctx->p2p && n_local < FUSED_SEND_MAX) {  // classify + send in one kernel
    MdevClassifyOp op;
    rc = mdev_classify(ctx, d_recs, n_local, types, &op);
    if (rc) return rc;
    rc = enqueue_exchange<2, MdevClassifyOp, 4>(ctx, ctx->surv.p, n_local, &owned_cap, &op);
  } else {
    rc = mdev_classify(ctx, d_recs, n_local, types);
    if (rc) return rc;
    rc = enqueue_exchange<2, MdevClassifyOp, 4>(ctx, ctx->surv.p, n_local, &owned_cap);
  }
  if (rc) return rc;
  rc = enqueue_owned_orderings(ctx, owned_cap, n_local, SRC_MDEV_TYPE, SRC_MDEV_PARENT, false);
  if (rc) return rc;
  ctx->last_n = n_local;
  ctx->last_total = owned_cap;
  ctx->last_kind = 4;
  ctx->last_units = 2;
  return KVG_OK;
}

int kvg_dev_scan_mdev_shard_fetch(kvg_ctx* ctx, kvg_mdev_shard_result** res) {
  if (!ctx || !res || ctx->last_kind != 4) return KVG_EINVAL;
  CK(cudaSetDevice(ctx->device));
  int rc = shard_ctrl_fetch(ctx);
  if (rc) return rc;
  const size_t S = ctx->h_ctrl->n_surv, ST = ctx->h_ctrl->n_own[0], SP = ctx->h_ctrl->n_own[1];
  const size_t KT = ctx->h_ctrl->n_dev_keys, NP = ctx->h_ctrl->n_groups;
  const uint32_t nt = ctx->n_types;
  BlockLayout L;
  L.take(sizeof(kvg_mdev_shard_result));
  const size_t o_local = L.take(S * 32), o_tm = L.take(ST * 32), o_tk32 = L.take(KT * 4), o_tk = L.take(KT * 2);
  const size_t o_toff = L.take((KT + 1) * 4), o_tperm = L.take(ST * 4);
  const size_t o_pm = L.take(SP * 32), o_pk = L.take(NP * 4), o_poff = L.take((NP + 1) * 4), o_pperm = L.take(SP * 4);
  DictPart D;
  D.layout(L, nt, nt ? ctx->h_type_off[nt] : 0);
  void* blk = pinned_alloc(ctx, L.o);
  if (!blk) return KVG_ENOMEM;
  uint8_t* b = pinned_payload(blk);
  auto D2H = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    return cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  OrderBufs& ot = ctx->ord_dev;
  OrderBufs& op = ctx->ord_grp;
  CK(D2H(o_local, ctx->surv.p, S * 32));
  CK(D2H(o_tm, ctx->owned0.p, ST * 32));
  CK(D2H(o_tk32, ot.seg_key.p, KT * 4));
  CK(D2H(o_toff, ot.seg_off.p, (KT + 1) * 4));
  CK(D2H(o_tperm, ot.perm.p, ST * 4));
  CK(D2H(o_pm, ctx->owned1.p, SP * 32));
  CK(D2H(o_pk, op.seg_key.p, NP * 4));
  CK(D2H(o_poff, op.seg_off.p, (NP + 1) * 4));
  CK(D2H(o_pperm, op.perm.p, SP * 4));
  rc = dict_copy(ctx, b, D);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  kvg_mdev_shard_result* r = (kvg_mdev_shard_result*)b;
  memset(r, 0, sizeof *r);
  if (ST == 0) ((uint32_t*)(b + o_toff))[0] = 0;
  if (SP == 0) ((uint32_t*)(b + o_poff))[0] = 0;
  uint16_t* tk = (uint16_t*)(b + o_tk);
  for (size_t k = 0; k < KT; k++) tk[k] = (uint16_t)((const uint32_t*)(b + o_tk32))[k];
  r->n_records = ctx->last_n;
  r->n_local = S;
  r->local = (const kvg_mdev_surv*)(b + o_local);
  r->n_type_members = ST;
  r->type_members = (const kvg_mdev_surv*)(b + o_tm);
  r->n_type_keys = (uint32_t)KT;
  r->type_keys = tk;
  r->type_off = (const uint32_t*)(b + o_toff);
  r->type_perm = (const uint32_t*)(b + o_tperm);
  r->n_par_members = SP;
  r->par_members = (const kvg_mdev_surv*)(b + o_pm);
  r->n_parents = (uint32_t)NP;
  r->par_keys = (const uint32_t*)(b + o_pk);
  r->par_off = (const uint32_t*)(b + o_poff);
  r->par_perm = (const uint32_t*)(b + o_pperm);
  dict_finish(ctx, b, D, &r->n_types, &r->label_off, &r->label_bytes, &r->type_canon, &r->type_name_off,
              &r->type_name_bytes);
  *res = r;
  return KVG_OK;
}

}  // extern "C"
