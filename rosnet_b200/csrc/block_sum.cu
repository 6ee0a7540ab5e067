// Stand-alone block sum: dst = (accumulate ? dst : 0) + sum_p src[p].
// Replaces `res += np.tensordot(a, b, axes)` accumulation (rosnet/array/compss/task/tensordot.py:45-49)
// when partial results already exist as separate buffers (partials produced on other GPUs).
// HBM-bound: algorithmic bytes = (n_src + 1 [+1 if accumulate]) * count * elem_bytes.
#include "common.h"

namespace rnb {
namespace {
constexpr int kMaxSrc = 16;
struct SumParams {
  const void *src[kMaxSrc];
  void *dst;
  int64_t n;  // lanes of the real type
  int32_t n_src, accumulate;
};

template <typename T, typename V, int L>  // V = 16-byte vector of L lanes of T
__global__ void __launch_bounds__(256) block_sum_kernel(const __grid_constant__ SumParams p) {
  const int64_t nvec = p.n / L;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    T acc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) acc[l] = T(0);
    if (p.accumulate) {
      V v = reinterpret_cast<const V *>(p.dst)[i];
      const T *q = reinterpret_cast<const T *>(&v);
#pragma unroll
      for (int l = 0; l < L; ++l) acc[l] = q[l];
    }
    for (int s = 0; s < p.n_src; ++s) {
      V v = __ldg(reinterpret_cast<const V *>(p.src[s]) + i);
      const T *q = reinterpret_cast<const T *>(&v);
#pragma unroll
      for (int l = 0; l < L; ++l) acc[l] += q[l];
    }
    V o;
    T *q = reinterpret_cast<T *>(&o);
#pragma unroll
    for (int l = 0; l < L; ++l) q[l] = acc[l];
    reinterpret_cast<V *>(p.dst)[i] = o;
  }
  // tail lanes
  if (blockIdx.x == 0 && threadIdx.x < (p.n % L)) {
    const int64_t i = nvec * L + threadIdx.x;
    T acc = p.accumulate ? reinterpret_cast<const T *>(p.dst)[i] : T(0);
    for (int s = 0; s < p.n_src; ++s) acc += reinterpret_cast<const T *>(p.src[s])[i];
    reinterpret_cast<T *>(p.dst)[i] = acc;
  }
}
}  // namespace
}  // namespace rnb

using namespace rnb;

extern "C" int rnb_block_sum(void *dst, const void *const *src_host_ptrs, int32_t n_src, int64_t count, int32_t dtype,
                             int32_t accumulate, rnb_stream_t stream_) {
  clear_error();
  RNB_REQUIRE(dst && src_host_ptrs, "NULL pointer");
  RNB_REQUIRE(n_src >= 1, "n_src must be >= 1");
  RNB_REQUIRE(dtype >= RNB_F32 && dtype <= RNB_C128, "unknown dtype");
  RNB_REQUIRE(count >= 0, "negative count");
  if (count == 0) return RNB_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool f32 = (dtype == RNB_F32 || dtype == RNB_C64);
  const int64_t lanes = count * ((dtype == RNB_C64 || dtype == RNB_C128) ? 2 : 1);
  RNB_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dst must be 16-byte aligned");
  for (int s = 0; s < n_src; ++s)
    RNB_REQUIRE(src_host_ptrs[s] && (reinterpret_cast<uintptr_t>(src_host_ptrs[s]) & 15) == 0, "src[%d] must be non-NULL and 16-byte aligned", s);
  int done = 0;
  int acc = accumulate;
  while (done < n_src) {  // more than kMaxSrc partials: several passes, accumulating
    SumParams p;
    p.dst = dst;
    p.n = lanes;
    p.n_src = (n_src - done > kMaxSrc) ? kMaxSrc : (n_src - done);
    p.accumulate = acc;
    for (int s = 0; s < p.n_src; ++s) p.src[s] = src_host_ptrs[done + s];
    const int L = f32 ? 4 : 2;
    int64_t blocks = (lanes / L + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (f32) block_sum_kernel<float, float4, 4><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else block_sum_kernel<double, double2, 2><<<(unsigned)blocks, 256, 0, stream>>>(p);
    RNB_CHECK_CUDA(cudaGetLastError());
    done += p.n_src;
    acc = 1;
  }
  return RNB_OK;
}
