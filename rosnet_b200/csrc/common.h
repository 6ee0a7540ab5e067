// Shared helpers for the rosnet_b200 CUDA library (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/rosnet_b200.h"

namespace rnb {

// thread-local message of the last failing call (returned by rnb_last_error_string)
void set_error(const char *fmt, ...);
void clear_error();

int sm_count();               // SMs of the current device (cached per device)
int max_smem_optin();         // opt-in dynamic shared memory per block

#define RNB_CHECK_CUDA(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      rnb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RNB_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define RNB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      rnb::set_error(__VA_ARGS__);    \
      return RNB_ERR_INVALID;         \
    }                                 \
  } while (0)

// Run-time knobs come from environment variables (RNB_*).  getenv() walks the whole environment, so every knob is read
// once and cached; rnb_reload_env() (tests, experiments that change os.environ mid-process) invalidates the cache.
struct EnvKnob {
  explicit EnvKnob(const char *name) : name_(name) {}
  const char *get();   // NULL when unset
 private:
  const char *name_;
  int gen_ = -1;
  bool set_ = false;
  char val_[128];
};
void reload_env();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: one flag per (call site, device), so a process
// that drives several GPUs opts in on each of them.
struct DeviceOnce {
  bool done[64] = {};
  bool scratch = false;
  bool &flag() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
      scratch = false;
      return scratch;
    }
    return done[dev];
  }
};

// Fused matricize + split (permute.cu): the store phase of the tiled permute kernel writes the hi / lo TF32 planes the
// single-precision GEMM consumes instead of the matricized operand itself (see rnb_permute_split).
struct SplitTarget {
  int mode;            // RNB_SPLIT_*
  float *hi, *lo;      // plane bases
  int64_t rows, k;     // rows and contracted extent of one matricized block (elements of the dtype)
  int32_t nblocks;
};
int permute_impl(const rnb_permute_desc *desc, const SplitTarget *split, cudaStream_t stream);

inline int dtype_bytes(int dtype) {
  switch (dtype) {
    case RNB_F32: return 4;
    case RNB_F64: return 8;
    case RNB_C64: return 8;
    case RNB_C128: return 16;
    default: return 0;
  }
}

// cuTensorMapEncodeTiled fetched through the runtime (no link-time dependency on libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled();

// ---- device-side PTX wrappers (mbarrier + TMA) ----------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Bounded wait: a protocol error in a warp-specialised kernel shows up as an mbarrier that never flips, i.e. a kernel that
// spins forever and takes the GPU with it.  This variant traps (the launch fails with an error) once a single wait has
// lasted ~5 s of SM clocks; no legitimate wait in these kernels is longer than a few hundred microseconds.
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_guard(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 0x3FFu) == 0 && clock64() - t0 > 10000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *tmap, const void *smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// round-to-nearest TF32 (10-bit mantissa) of an fp32 value, kept in fp32 storage
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- split-K second pass (shared by the FP64 and the TF32 grouped GEMMs) -----------------------
// When a launch has few output tiles and a long contracted extent, every tile's (pair, k) range is
// cut into `splits` pieces computed by different CTAs; each writes its partial bm x bn tile
// (row-major, scalars of type T) to the workspace at slot = tile * splits + s.  This kernel adds
// the splits of every output element in the FIXED order s = 0, 1, ... (deterministic) and writes /
// accumulates into C.  One CTA per output tile, 16-byte vectors; m, n, c_ld, c_block_stride, bn in
// scalars of T; bn is a power of two.
template <typename T>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const T *__restrict__ partial, T *__restrict__ c,
                                                            const int32_t *__restrict__ out_index, int64_t m, int64_t n, int64_t c_ld,
                                                            int64_t c_block_stride, int tiles_m, int tiles_n, int panel, int bm, int bn,
                                                            int splits, int accumulate, int c_vec_ok) {
  constexpr int VEC = 16 / sizeof(T);
  const int tile = blockIdx.x;
  const int tiles_per_block = tiles_m * tiles_n;
  const int g = tile / tiles_per_block;
  const int rem = tile - g * tiles_per_block;
  const int panel_tiles = panel * tiles_m;
  const int pn = rem / panel_tiles;
  const int within = rem - pn * panel_tiles;
  const int w = min(panel, tiles_n - pn * panel);
  const int tm = within / w;
  const int tn = pn * panel + (within - tm * w);
  const int64_t tile_elems = (int64_t)bm * bn;
  const T *src0 = partial + (int64_t)tile * splits * tile_elems;
  const int64_t out_blk = out_index ? out_index[g] : g;
  T *cb = c + out_blk * c_block_stride;
  const int vec_per_row = bn / VEC;
  const int shift = 31 - __clz(vec_per_row);
  for (int e = threadIdx.x; e < bm * vec_per_row; e += blockDim.x) {
    const int row_l = e >> shift;
    const int col_l = (e & (vec_per_row - 1)) * VEC;
    const int64_t row = (int64_t)tm * bm + row_l;
    const int64_t col = (int64_t)tn * bn + col_l;
    if (row >= m || col >= n) continue;
    const T *src = src0 + (int64_t)row_l * bn + col_l;
    __align__(16) T v[VEC];
    {
      const uint4 q = *reinterpret_cast<const uint4 *>(src);
      *reinterpret_cast<uint4 *>(v) = q;
    }
    for (int s = 1; s < splits; ++s) {
      __align__(16) T u[VEC];
      *reinterpret_cast<uint4 *>(u) = *reinterpret_cast<const uint4 *>(src + (int64_t)s * tile_elems);
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] += u[j];
    }
    T *dst = cb + row * c_ld + col;
    if (c_vec_ok && col + VEC <= n) {
      if (accumulate) {
        __align__(16) T o[VEC];
        *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<const uint4 *>(dst);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] += o[j];
      }
      *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(v);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (col + j < n) {
          T x = v[j];
          if (accumulate) x += dst[j];
          dst[j] = x;
        }
      }
    }
  }
}
#endif

// split-K decision shared by both GEMMs: 1 = off.  Units = tiles * splits run as waves of `sms`
// persistent CTAs; pick the smallest split count whose last wave is fullest (wave quantisation:
// 64 tiles x 2 splits fills 86 % of 148 SMs, x 9 fills 97 %) and only when it beats the unsplit
// launch by more than 10 %.  A split costs a 128 KB partial tile written and read back (~6 us of one
// SM's HBM share), so every split must keep `min_iters_many` k-iterations (~5 % overhead); when at
// most half the SMs would have a tile at all, `min_iters_few` suffices.
inline int choose_splits(int64_t total_tiles, int64_t iters_per_tile, int min_iters_few, int min_iters_many, int sms) {
  static EnvKnob knob_splitk("RNB_SPLITK");
  if (const char *e = knob_splitk.get()) {
    const int v = atoi(e);
    if (v >= 1) return (int)(v < iters_per_tile ? v : (iters_per_tile > 0 ? iters_per_tile : 1));
    if (v == 0) return 1;
  }
  if (total_tiles <= 0 || total_tiles >= 4 * (int64_t)sms) return 1;
  const int min_iters = total_tiles * 2 <= sms ? min_iters_few : min_iters_many;
  int64_t cap = iters_per_tile / min_iters;
  if (cap > 4 * (int64_t)sms / total_tiles + 1) cap = 4 * (int64_t)sms / total_tiles + 1;
  auto eff = [&](int64_t s) {
    const int64_t units = total_tiles * s;
    const int64_t waves = (units + sms - 1) / sms;
    return (double)units / (double)(waves * sms);
  };
  int best = 1;
  double best_eff = eff(1) * 1.10;
  for (int64_t s = 2; s <= cap; ++s) {
    const double e = eff(s);
    if (e > best_eff + 1e-9) {
      best_eff = e;
      best = (int)s;
    }
  }
  return best;
}

}  // namespace rnb
