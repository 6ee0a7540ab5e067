// C-ABI entry points that are not kernels: library/device queries, error string, staging copies,
// the dtype dispatcher of rnb_contract_grouped and the NCCL wrappers (dlopen, no link dependency).
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.h"

namespace rnb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void clear_error() { g_err[0] = 0; }

static std::atomic<int> g_env_gen{0};
static std::mutex g_env_mutex;

void reload_env() { g_env_gen.fetch_add(1); }

const char *EnvKnob::get() {
  // RNB_ENV_NOCACHE=1 (set before the first call; the test-suite does, its cases flip knobs mid-process): read every time
  static const bool nocache = [] { const char *e = getenv("RNB_ENV_NOCACHE"); return e && atoi(e) != 0; }();
  const int gen = g_env_gen.load(std::memory_order_acquire);
  std::lock_guard<std::mutex> lock(g_env_mutex);
  if (nocache || gen_ != gen) {
    const char *e = getenv(name_);
    set_ = e != nullptr;
    if (e) {
      strncpy(val_, e, sizeof(val_) - 1);
      val_[sizeof(val_) - 1] = 0;
    }
    gen_ = gen;
  }
  return set_ ? val_ : nullptr;
}

static int g_sm_count[64];
static int g_smem_optin[64];
static std::once_flag g_dev_once[64];

static void fill_dev_cache(int dev) {
  if (dev < 0 || dev >= 64) return;
  std::call_once(g_dev_once[dev], [dev] {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count[dev] = v > 0 ? v : 1;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_smem_optin[dev] = v;
  });
}
int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  fill_dev_cache(dev);
  return (dev >= 0 && dev < 64) ? g_sm_count[dev] : 1;
}
int max_smem_optin() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  fill_dev_cache(dev);
  return (dev >= 0 && dev < 64) ? g_smem_optin[dev] : 0;
}

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

// implemented in the kernel translation units
int contract_f64(const rnb_contract_desc *d, cudaStream_t stream);
int contract_tf32(const rnb_contract_desc *d, cudaStream_t stream);
int contract_tf32_workspace(const rnb_contract_desc *d, size_t *bytes);
int contract_f64_workspace(const rnb_contract_desc *d, size_t *bytes);
// skinny.cu: SIMT routes for scalar-like results and tiny problems (0 = use the tensor cores)
int simt_route(const rnb_contract_desc *d);
int contract_simt_workspace(const rnb_contract_desc *d, size_t *bytes);
int contract_simt(const rnb_contract_desc *d, cudaStream_t stream);
int contract_nd(const rnb_contract_nd_desc *d, cudaStream_t stream);
int contract_nd_batch(const rnb_contract_nd_desc *descs, int32_t n, cudaStream_t stream);
int contract_tf32_planes(const rnb_contract_desc *d, rnb_operand_planes *a, rnb_operand_planes *b);

}  // namespace rnb

using namespace rnb;

extern "C" {

int rnb_abi_version(void) { return RNB_ABI_VERSION; }

int rnb_reload_env(void) {
  reload_env();
  return RNB_OK;
}

const char *rnb_last_error_string(void) { return g_err; }

int rnb_get_device_props(int device, rnb_device_props *out) {
  clear_error();
  RNB_REQUIRE(out != nullptr, "out is NULL");
  cudaDeviceProp prop;
  RNB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  memset(out, 0, sizeof(*out));
  out->sm_count = prop.multiProcessorCount;
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->max_smem_per_block = (int32_t)prop.sharedMemPerBlockOptin;
  out->total_mem_bytes = (int64_t)prop.totalGlobalMem;
  out->l2_bytes_mb = (int32_t)(prop.l2CacheSize >> 20);
  return RNB_OK;
}

static int validate_contract(const rnb_contract_desc *d) {
  RNB_REQUIRE(d != nullptr, "descriptor is NULL");
  RNB_REQUIRE(d->dtype >= RNB_F32 && d->dtype <= RNB_C128, "unknown dtype %d", d->dtype);
  RNB_REQUIRE(d->m > 0 && d->n > 0 && d->k > 0, "m, n, k must be positive (got %lld, %lld, %lld)", (long long)d->m,
              (long long)d->n, (long long)d->k);
  RNB_REQUIRE(d->n_out >= 0 && d->n_pairs >= 1, "n_out must be >= 0 and n_pairs >= 1");
  RNB_REQUIRE(d->a && d->b && d->c, "operand pointers must not be NULL");
  RNB_REQUIRE(d->pair_a && d->pair_b, "pair tables must not be NULL");
  RNB_REQUIRE(d->a_major == RNB_MAJOR_K || d->a_major == RNB_MAJOR_MN, "bad a_major");
  RNB_REQUIRE(d->b_major == RNB_MAJOR_K || d->b_major == RNB_MAJOR_MN, "bad b_major");
  RNB_REQUIRE(d->a_nblocks >= 1 && d->b_nblocks >= 1 && d->c_nblocks >= 1, "block counts must be >= 1");
  RNB_REQUIRE(d->a_ld >= (d->a_major == RNB_MAJOR_K ? d->k : d->m), "a_ld too small");
  RNB_REQUIRE(d->b_ld >= (d->b_major == RNB_MAJOR_K ? d->k : d->n), "b_ld too small");
  RNB_REQUIRE(d->c_ld >= d->n, "c_ld too small");
  if (d->c_peers) {
    RNB_REQUIRE(d->dtype == RNB_F64 || d->dtype == RNB_C128, "fused peer reduction is implemented for float64/complex128 only");
    RNB_REQUIRE(d->n_peers >= 1 && d->n_peers <= RNB_MAX_PEERS && d->blocks_per_peer >= 1, "bad n_peers / blocks_per_peer");
    RNB_REQUIRE(!d->accumulate, "accumulate and c_peers are mutually exclusive (peer mode always adds)");
  }
  return RNB_OK;
}

int rnb_contract_workspace_bytes(const rnb_contract_desc *d, size_t *bytes) {
  clear_error();
  RNB_REQUIRE(bytes != nullptr, "bytes is NULL");
  int rc = validate_contract(d);
  if (rc != RNB_OK) return rc;
  *bytes = 0;
  if (simt_route(d)) return contract_simt_workspace(d, bytes);
  if (d->dtype == RNB_F32 || d->dtype == RNB_C64) return contract_tf32_workspace(d, bytes);
  return contract_f64_workspace(d, bytes);
}

int rnb_contract_planes(const rnb_contract_desc *d, rnb_operand_planes *a, rnb_operand_planes *b) {
  clear_error();
  RNB_REQUIRE(a != nullptr && b != nullptr, "a / b is NULL");
  int rc = validate_contract(d);
  if (rc != RNB_OK) return rc;
  memset(a, 0, sizeof(*a));
  memset(b, 0, sizeof(*b));
  if (simt_route(d) || !(d->dtype == RNB_F32 || d->dtype == RNB_C64) || d->a_major != RNB_MAJOR_K || d->b_major != RNB_MAJOR_K) return RNB_OK;
  return contract_tf32_planes(d, a, b);
}

int rnb_permute_split(const rnb_permute_desc *desc, const rnb_operand_planes *planes, void *workspace, int64_t rows, int64_t k,
                      int32_t nblocks, rnb_stream_t stream) {
  clear_error();
  RNB_REQUIRE(desc != nullptr && planes != nullptr && workspace != nullptr, "NULL argument");
  RNB_REQUIRE(planes->mode >= RNB_SPLIT_PLANES && planes->mode <= RNB_SPLIT_C64_3M, "planes->mode must be a fused split mode");
  RNB_REQUIRE(rows > 0 && k > 0 && nblocks >= 1, "rows, k, nblocks must be positive");
  RNB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  SplitTarget tgt;
  tgt.mode = planes->mode;
  tgt.hi = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + planes->hi_offset);
  tgt.lo = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + planes->lo_offset);
  tgt.rows = rows;
  tgt.k = k;
  tgt.nblocks = nblocks;
  rnb_permute_desc dd = *desc;
  dd.dst = tgt.hi;   // alignment checks of the destination side run against the plane base
  return permute_impl(&dd, &tgt, reinterpret_cast<cudaStream_t>(stream));
}

int rnb_contract_route(const rnb_contract_desc *d, int32_t *route) {
  clear_error();
  RNB_REQUIRE(route != nullptr, "route is NULL");
  int rc = validate_contract(d);
  if (rc != RNB_OK) return rc;
  *route = simt_route(d);
  return RNB_OK;
}

int rnb_contract_grouped(const rnb_contract_desc *d, rnb_stream_t stream) {
  clear_error();
  int rc = validate_contract(d);
  if (rc != RNB_OK) return rc;
  if (d->n_out == 0) return RNB_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (simt_route(d)) return contract_simt(d, s);
  if (d->dtype == RNB_F64 || d->dtype == RNB_C128) return contract_f64(d, s);
  return contract_tf32(d, s);
}

int rnb_contract_nd(const rnb_contract_nd_desc *d, rnb_stream_t stream) {
  clear_error();
  RNB_REQUIRE(d != nullptr, "descriptor is NULL");
  return contract_nd(d, reinterpret_cast<cudaStream_t>(stream));
}

int rnb_contract_nd_batch(const rnb_contract_nd_desc *descs, int32_t n, rnb_stream_t stream) {
  clear_error();
  return contract_nd_batch(descs, n, reinterpret_cast<cudaStream_t>(stream));
}

int rnb_memcpy_h2d_async(void *dst, const void *src, size_t bytes, rnb_stream_t stream) {
  clear_error();
  RNB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}
int rnb_memcpy_d2h_async(void *dst, const void *src, size_t bytes, rnb_stream_t stream) {
  clear_error();
  RNB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}
int rnb_memcpy_d2d_async(void *dst, const void *src, size_t bytes, rnb_stream_t stream) {
  clear_error();
  // cudaMemcpyDefault: either pointer may be a peer-mapped (CUDA IPC) buffer of another GPU of the box
  RNB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}
int rnb_stream_synchronize(rnb_stream_t stream) {
  clear_error();
  RNB_CHECK_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}

// ---- peer memory ---------------------------------------------------------------------------------
int rnb_device_malloc(size_t bytes, void **device_ptr) {
  clear_error();
  RNB_REQUIRE(device_ptr != nullptr, "device_ptr is NULL");
  RNB_CHECK_CUDA(cudaMalloc(device_ptr, bytes));
  return RNB_OK;
}
int rnb_device_free(void *device_ptr) {
  clear_error();
  RNB_CHECK_CUDA(cudaFree(device_ptr));
  return RNB_OK;
}
int rnb_memset_async(void *device_ptr, int value, size_t bytes, rnb_stream_t stream) {
  clear_error();
  RNB_CHECK_CUDA(cudaMemsetAsync(device_ptr, value, bytes, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}
int rnb_ipc_get_handle(void *device_ptr, char handle_host[RNB_IPC_HANDLE_BYTES]) {
  clear_error();
  static_assert(sizeof(cudaIpcMemHandle_t) == RNB_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  RNB_CHECK_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
  memcpy(handle_host, &h, sizeof(h));
  return RNB_OK;
}
int rnb_ipc_open_handle(const char handle_host[RNB_IPC_HANDLE_BYTES], void **device_ptr) {
  clear_error();
  RNB_REQUIRE(device_ptr != nullptr, "device_ptr is NULL");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  RNB_CHECK_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RNB_OK;
}
int rnb_ipc_close_handle(void *device_ptr) {
  clear_error();
  RNB_CHECK_CUDA(cudaIpcCloseMemHandle(device_ptr));
  return RNB_OK;
}

// ---- NCCL through dlopen -----------------------------------------------------------------------
struct Id128 {
  char internal[128];
};
namespace {
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*ReduceScatter)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static char g_nccl_why[256] = "missing symbols";

static bool nccl_load() {
  std::call_once(g_nccl_once, [] {
    const char *names[] = {getenv("RNB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n) continue;
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
      if (const char *e = dlerror()) snprintf(g_nccl_why, sizeof(g_nccl_why), "%s", e);   // dlerror() clears itself: read it once
    }
    if (!g_nccl.handle) return;
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.handle, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.handle, "ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.handle, "ncclCommDestroy"));
    g_nccl.ReduceScatter = reinterpret_cast<decltype(g_nccl.ReduceScatter)>(dlsym(g_nccl.handle, "ncclReduceScatter"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(g_nccl.handle, "ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(g_nccl.handle, "ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.handle, "ncclGetErrorString"));
  });
  if (!g_nccl.handle || !g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.ReduceScatter || !g_nccl.AllReduce) {
    set_error("NCCL is not loadable (tried $RNB_NCCL_LIB, libnccl.so.2, libnccl.so): %s", g_nccl_why);
    return false;
  }
  return true;
}

#define RNB_CHECK_NCCL(expr)                                                                  \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != 0) {                                                                            \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
      return RNB_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)

// ncclDataType_t: ncclFloat32 = 7, ncclFloat64 = 8; ncclRedOp_t: ncclSum = 0
static int nccl_type(int dtype, int *lanes) {
  *lanes = (dtype == RNB_C64 || dtype == RNB_C128) ? 2 : 1;
  return (dtype == RNB_F32 || dtype == RNB_C64) ? 7 : 8;
}

int rnb_nccl_get_unique_id(char id_host[RNB_NCCL_UNIQUE_ID_BYTES]) {
  clear_error();
  if (!nccl_load()) return RNB_ERR_NCCL;
  RNB_CHECK_NCCL(g_nccl.GetUniqueId(id_host));
  return RNB_OK;
}
int rnb_nccl_comm_init_rank(void **comm, int n_ranks, const char id_host[RNB_NCCL_UNIQUE_ID_BYTES], int rank) {
  clear_error();
  if (!nccl_load()) return RNB_ERR_NCCL;
  Id128 id;
  memcpy(id.internal, id_host, 128);
  RNB_CHECK_NCCL(g_nccl.CommInitRank(comm, n_ranks, id, rank));
  return RNB_OK;
}
int rnb_nccl_comm_destroy(void *comm) {
  clear_error();
  if (!nccl_load()) return RNB_ERR_NCCL;
  if (g_nccl.CommDestroy) RNB_CHECK_NCCL(g_nccl.CommDestroy(comm));
  return RNB_OK;
}
int rnb_reduce_scatter_sum(const void *send, void *recv, int64_t recv_count, int32_t dtype, void *comm,
                           rnb_stream_t stream) {
  clear_error();
  if (!nccl_load()) return RNB_ERR_NCCL;
  RNB_REQUIRE(dtype >= RNB_F32 && dtype <= RNB_C128, "unknown dtype");
  int lanes;
  int t = nccl_type(dtype, &lanes);
  RNB_CHECK_NCCL(g_nccl.ReduceScatter(send, recv, (size_t)recv_count * lanes, t, 0, comm, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}
int rnb_all_reduce_sum(const void *send, void *recv, int64_t count, int32_t dtype, void *comm, rnb_stream_t stream) {
  clear_error();
  if (!nccl_load()) return RNB_ERR_NCCL;
  RNB_REQUIRE(dtype >= RNB_F32 && dtype <= RNB_C128, "unknown dtype");
  int lanes;
  int t = nccl_type(dtype, &lanes);
  RNB_CHECK_NCCL(g_nccl.AllReduce(send, recv, (size_t)count * lanes, t, 0, comm, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}
int rnb_all_gather(const void *send, void *recv, int64_t send_count, int32_t dtype, void *comm, rnb_stream_t stream) {
  clear_error();
  if (!nccl_load()) return RNB_ERR_NCCL;
  RNB_REQUIRE(dtype >= RNB_F32 && dtype <= RNB_C128, "unknown dtype");
  RNB_REQUIRE(g_nccl.AllGather != nullptr, "ncclAllGather is missing from the loaded NCCL");
  int lanes;
  int t = nccl_type(dtype, &lanes);
  RNB_CHECK_NCCL(g_nccl.AllGather(send, recv, (size_t)send_count * lanes, t, comm, reinterpret_cast<cudaStream_t>(stream)));
  return RNB_OK;
}

}  // extern "C"
