// SIMT paths of rnb_contract_grouped for block contractions the tensor cores cannot help with.
//
// Tensor-network contraction paths (rosnet's opt_einsum/cotengra callers) end with
// steps whose result is a scalar or a handful of numbers while the contracted extent is huge (the
// final inner product of two 2^20..2^25-element tensors), and start with hundreds of steps that are a
// few hundred flops.  A 128x256 tensor-core tile would do >10^4 times the useful work for the first
// kind (one CTA walking the whole k: measured 2.1 s for k = 2^25) and pays three launches plus a
// 200 KB-smem prologue for the second.
//
//   skinny : m <= 4 and n <= 4, any k      -> HBM-bound reduction over the flattened (pair, k) axis,
//            split over many CTAs; per-CTA partials go to the workspace, a second tiny launch adds
//            them in a FIXED order (deterministic) and writes C.  Algorithmic bytes per launch =
//            n_out * n_pairs * k * (m + n) * elem_bytes.
//   small  : few MACs in total            -> one thread per output element, one launch.
//
// Both accumulate in float64 (complex128 for complex inputs) whatever the operand dtype, so they are
// at least as accurate as the BLAS call of the reference (rosnet/array/block/__init__.py:281-283).
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.h"

namespace rnb {
namespace {

constexpr int kSkinnyThreads = 256;
constexpr int kSkinnyUnroll = 4;

struct SimtParams {
  const void *a, *b;
  void *c;
  int64_t a_rs, a_ks, a_bs;  // row / k / block strides of a, elements
  int64_t b_rs, b_ks, b_bs;
  int64_t c_ld, c_bs;
  int64_t k;                 // per pair
  int64_t k_total;           // n_pairs * k
  int64_t k_per_split;
  const int32_t *pair_a, *pair_b, *out_index;
  int32_t m, n, n_out, n_pairs, splits, accumulate;
  void *partial;             // [n_out][splits][MT*NT] wide elements
};

template <typename T> struct Wide;
template <> struct Wide<float>   { typedef double type; };
template <> struct Wide<double>  { typedef double type; };
template <> struct Wide<float2>  { typedef double2 type; };
template <> struct Wide<double2> { typedef double2 type; };

__device__ __forceinline__ double wzero(double *) { return 0.0; }
__device__ __forceinline__ double2 wzero(double2 *) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ void wfma(double &acc, float a, float b) { acc = fma((double)a, (double)b, acc); }
__device__ __forceinline__ void wfma(double &acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void wfma(double2 &acc, float2 a, float2 b) {
  const double ar = a.x, ai = a.y, br = b.x, bi = b.y;
  acc.x = fma(ar, br, acc.x); acc.x = fma(-ai, bi, acc.x);
  acc.y = fma(ar, bi, acc.y); acc.y = fma(ai, br, acc.y);
}
__device__ __forceinline__ void wfma(double2 &acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double wadd(double a, double b) { return a + b; }
__device__ __forceinline__ double2 wadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double wshfl(double v, int off) { return __shfl_down_sync(0xffffffffu, v, off); }
__device__ __forceinline__ double2 wshfl(double2 v, int off) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, off), __shfl_down_sync(0xffffffffu, v.y, off));
}
__device__ __forceinline__ float narrow(double v, float *) { return (float)v; }
__device__ __forceinline__ double narrow(double v, double *) { return v; }
__device__ __forceinline__ float2 narrow(double2 v, float2 *) { return make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ double2 narrow(double2 v, double2 *) { return v; }
__device__ __forceinline__ double widen(float v) { return (double)v; }
__device__ __forceinline__ double widen(double v) { return v; }
__device__ __forceinline__ double2 widen(float2 v) { return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ double2 widen(double2 v) { return v; }

// ---- skinny: CTA (g, s) reduces the k range [s*k_per_split, (s+1)*k_per_split) of output block g ----
template <typename T, int MT, int NT>
__global__ void __launch_bounds__(kSkinnyThreads) skinny_kernel(const SimtParams p) {
  typedef typename Wide<T>::type W;
  const int g = blockIdx.x / p.splits;
  const int s = blockIdx.x - g * p.splits;
  const T *__restrict__ A = static_cast<const T *>(p.a);
  const T *__restrict__ B = static_cast<const T *>(p.b);
  W acc[MT][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = wzero((W *)nullptr);

  const int64_t f0 = (int64_t)s * p.k_per_split;
  int64_t f1 = f0 + p.k_per_split;
  if (f1 > p.k_total) f1 = p.k_total;
  int64_t f = f0;
  while (f < f1) {
    const int pr = (int)(f / p.k);
    const int64_t klo = f - (int64_t)pr * p.k;
    int64_t khi = klo + (f1 - f);
    if (khi > p.k) khi = p.k;
    const T *Ab = A + (int64_t)p.pair_a[(int64_t)g * p.n_pairs + pr] * p.a_bs;
    const T *Bb = B + (int64_t)p.pair_b[(int64_t)g * p.n_pairs + pr] * p.b_bs;
    for (int64_t k0 = klo + threadIdx.x; k0 < khi; k0 += (int64_t)kSkinnyThreads * kSkinnyUnroll) {
      T av[kSkinnyUnroll][MT], bv[kSkinnyUnroll][NT];
#pragma unroll
      for (int u = 0; u < kSkinnyUnroll; ++u) {
        const int64_t kk = k0 + (int64_t)u * kSkinnyThreads;
        const bool ok = kk < khi;
#pragma unroll
        for (int i = 0; i < MT; ++i)
          if (ok && i < p.m) av[u][i] = Ab[(int64_t)i * p.a_rs + kk * p.a_ks];
#pragma unroll
        for (int j = 0; j < NT; ++j)
          if (ok && j < p.n) bv[u][j] = Bb[(int64_t)j * p.b_rs + kk * p.b_ks];
      }
#pragma unroll
      for (int u = 0; u < kSkinnyUnroll; ++u) {
        const int64_t kk = k0 + (int64_t)u * kSkinnyThreads;
        if (kk < khi) {
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
              if (i < p.m && j < p.n) wfma(acc[i][j], av[u][i], bv[u][j]);
        }
      }
    }
    f += khi - klo;
  }

  // fixed-order CTA reduction: warp shuffles, then the 8 warp leaders through shared memory
  __shared__ W red[kSkinnyThreads / 32][MT * NT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      W v = acc[i][j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v = wadd(v, wshfl(v, off));
      if (lane == 0) red[warp][i * NT + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < MT * NT) {
    W v = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < kSkinnyThreads / 32; ++w) v = wadd(v, red[w][threadIdx.x]);
    static_cast<W *>(p.partial)[((int64_t)g * p.splits + s) * (MT * NT) + threadIdx.x] = v;
  }
}

template <typename T, int MT, int NT>
__global__ void __launch_bounds__(128) skinny_finish_kernel(const SimtParams p) {
  typedef typename Wide<T>::type W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.n_out * (MT * NT)) return;
  const int g = (int)(idx / (MT * NT));
  const int e = (int)(idx - (int64_t)g * (MT * NT));
  const int i = e / NT, j = e - i * NT;
  if (i >= p.m || j >= p.n) return;
  const W *src = static_cast<const W *>(p.partial) + (int64_t)g * p.splits * (MT * NT) + e;
  W v = src[0];
  for (int s = 1; s < p.splits; ++s) v = wadd(v, src[(int64_t)s * (MT * NT)]);
  const int out_blk = p.out_index ? p.out_index[g] : g;
  T *dst = static_cast<T *>(p.c) + (int64_t)out_blk * p.c_bs + (int64_t)i * p.c_ld + j;
  if (p.accumulate) v = wadd(v, widen(*dst));
  *dst = narrow(v, (T *)nullptr);
}

// ---- small: one thread per output element ----
template <typename T>
__global__ void __launch_bounds__(128) small_kernel(const SimtParams p) {
  typedef typename Wide<T>::type W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_block = (int64_t)p.m * p.n;
  if (idx >= (int64_t)p.n_out * per_block) return;
  const int g = (int)(idx / per_block);
  const int64_t e = idx - (int64_t)g * per_block;
  const int64_t i = e / p.n, j = e - i * p.n;
  const T *__restrict__ A = static_cast<const T *>(p.a);
  const T *__restrict__ B = static_cast<const T *>(p.b);
  W v = wzero((W *)nullptr);
  for (int pr = 0; pr < p.n_pairs; ++pr) {
    const T *Ab = A + (int64_t)p.pair_a[(int64_t)g * p.n_pairs + pr] * p.a_bs + i * p.a_rs;
    const T *Bb = B + (int64_t)p.pair_b[(int64_t)g * p.n_pairs + pr] * p.b_bs + j * p.b_rs;
    for (int64_t kk = 0; kk < p.k; ++kk) wfma(v, Ab[kk * p.a_ks], Bb[kk * p.b_ks]);
  }
  const int out_blk = p.out_index ? p.out_index[g] : g;
  T *dst = static_cast<T *>(p.c) + (int64_t)out_blk * p.c_bs + i * p.c_ld + j;
  if (p.accumulate) v = wadd(v, widen(*dst));
  *dst = narrow(v, (T *)nullptr);
}

// ---- small, n-d strided: one thread per output element, operands addressed through their own block axes ----
// The tiny steps of a tensor-network path (hundreds per slice) used to cost three launches each: two permutes that
// matricize the operands and the small kernel above.  Here the kernel walks the operands where they lie: the free axes of
// a (in output order), the free axes of b and the contracted axes come as (extent, stride) lists; the k offsets of both
// operands are tabulated once per CTA in shared memory, every thread decodes its own row / column offset.
constexpr int kNdMaxK = 512;
struct NdParams {
  const void *a, *b;
  void *c;
  int64_t a_bs, b_bs, c_bs;
  const int32_t *pair_a, *pair_b, *out_index;
  int32_t m, n, k, n_out, n_pairs, accumulate;
  int32_t rank_m, rank_n, rank_k;
  int32_t ext_m[RNB_MAX_RANK], ext_n[RNB_MAX_RANK], ext_k[RNB_MAX_RANK];
  int32_t sa_m[RNB_MAX_RANK], sb_n[RNB_MAX_RANK], sa_k[RNB_MAX_RANK], sb_k[RNB_MAX_RANK];
};

__device__ __forceinline__ int nd_offset(int idx, int rank, const int32_t *ext, const int32_t *stride) {
  int off = 0;
  for (int d = rank - 1; d >= 0; --d) {   // last axis fastest
    const int e = ext[d];
    const int q = idx / e;
    off += (idx - q * e) * stride[d];
    idx = q;
  }
  return off;
}

template <typename T>
__global__ void __launch_bounds__(128) small_nd_kernel(const __grid_constant__ NdParams p) {
  typedef typename Wide<T>::type W;
  __shared__ int32_t ka[kNdMaxK], kb[kNdMaxK];
  for (int t = threadIdx.x; t < p.k; t += blockDim.x) {
    ka[t] = nd_offset(t, p.rank_k, p.ext_k, p.sa_k);
    kb[t] = nd_offset(t, p.rank_k, p.ext_k, p.sb_k);
  }
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_block = (int64_t)p.m * p.n;
  if (idx >= (int64_t)p.n_out * per_block) return;
  const int g = (int)(idx / per_block);
  const int e = (int)(idx - (int64_t)g * per_block);
  const int i = e / p.n, j = e - i * p.n;
  const int oa = nd_offset(i, p.rank_m, p.ext_m, p.sa_m), ob = nd_offset(j, p.rank_n, p.ext_n, p.sb_n);
  const T *__restrict__ A = static_cast<const T *>(p.a);
  const T *__restrict__ B = static_cast<const T *>(p.b);
  W v = wzero((W *)nullptr);
  for (int pr = 0; pr < p.n_pairs; ++pr) {
    const T *Ab = A + (int64_t)p.pair_a[(int64_t)g * p.n_pairs + pr] * p.a_bs + oa;
    const T *Bb = B + (int64_t)p.pair_b[(int64_t)g * p.n_pairs + pr] * p.b_bs + ob;
    for (int kk = 0; kk < p.k; ++kk) wfma(v, Ab[ka[kk]], Bb[kb[kk]]);
  }
  const int out_blk = p.out_index ? p.out_index[g] : g;
  T *dst = static_cast<T *>(p.c) + (int64_t)out_blk * p.c_bs + e;
  if (p.accumulate) v = wadd(v, widen(*dst));
  *dst = narrow(v, (T *)nullptr);
}

// the same with every offset tabulated by the caller (rnb_contract_nd_desc::tab_*): no index arithmetic, tiny parameter block
struct NdTabParams {
  const void *a, *b;
  void *c;
  int64_t a_bs, b_bs, c_bs;
  const int32_t *pair_a, *pair_b, *out_index;
  const int32_t *tab_m, *tab_n, *tab_ka, *tab_kb;
  int32_t m, n, k, n_out, n_pairs, accumulate;
};

template <typename T>
__global__ void __launch_bounds__(128) small_tab_kernel(const NdTabParams p) {
  typedef typename Wide<T>::type W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_block = (int64_t)p.m * p.n;
  if (idx >= (int64_t)p.n_out * per_block) return;
  const int g = (int)(idx / per_block);
  const int e = (int)(idx - (int64_t)g * per_block);
  const int i = e / p.n, j = e - i * p.n;
  const int oa = __ldg(p.tab_m + i), ob = __ldg(p.tab_n + j);
  const T *__restrict__ A = static_cast<const T *>(p.a);
  const T *__restrict__ B = static_cast<const T *>(p.b);
  W v = wzero((W *)nullptr);
  for (int pr = 0; pr < p.n_pairs; ++pr) {
    const T *Ab = A + (int64_t)p.pair_a[(int64_t)g * p.n_pairs + pr] * p.a_bs + oa;
    const T *Bb = B + (int64_t)p.pair_b[(int64_t)g * p.n_pairs + pr] * p.b_bs + ob;
    for (int kk = 0; kk < p.k; ++kk) wfma(v, Ab[__ldg(p.tab_ka + kk)], Bb[__ldg(p.tab_kb + kk)]);
  }
  const int out_blk = p.out_index ? p.out_index[g] : g;
  T *dst = static_cast<T *>(p.c) + (int64_t)out_blk * p.c_bs + e;
  if (p.accumulate) v = wadd(v, widen(*dst));
  *dst = narrow(v, (T *)nullptr);
}

template <typename T>
int launch_small_tab(const NdTabParams &p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_out * p.m * p.n;
  small_tab_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

// ---- many independent small contractions in ONE launch (rnb_contract_nd_batch) ----
// The first ~230 pairwise steps of a tensor-network slice are a few microseconds of work each and mostly independent of one
// another; one launch per step makes the path launch-latency bound.  The tasks of one dependency level travel in the kernel's
// parameter block (up to 32 KB on sm_100a), every CTA finds its task by bisection over the tasks' first-CTA numbers.
struct BatchTask {
  const void *a, *b;
  void *c;
  const int32_t *tab;        // tab_m [m], tab_n [n], tab_ka [k], tab_kb [k] back to back
  const int32_t *pair_a, *pair_b;
  int32_t m, n, k, n_out, n_pairs;
  int32_t a_bs, b_bs, c_bs;  // block strides, elements
  int32_t cta0;              // first CTA of this task
  int32_t accumulate;
};
constexpr int kBatchMax = 340;   // 340 * 88 B + header < 32 764 B of kernel parameters
struct BatchParams {
  int32_t n_tasks, pad_;
  BatchTask t[kBatchMax];
};
static_assert(sizeof(BatchParams) <= 32000, "kernel parameter block too large");

template <typename T>
__global__ void __launch_bounds__(128) small_batch_kernel(const __grid_constant__ BatchParams p) {
  typedef typename Wide<T>::type W;
  int lo = 0, hi = p.n_tasks - 1;          // last task whose cta0 <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (p.t[mid].cta0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const BatchTask &q = p.t[lo];
  const int64_t idx = (int64_t)((int)blockIdx.x - q.cta0) * blockDim.x + threadIdx.x;
  const int64_t per_block = (int64_t)q.m * q.n;
  if (idx >= (int64_t)q.n_out * per_block) return;
  const int g = (int)(idx / per_block);
  const int e = (int)(idx - (int64_t)g * per_block);
  const int i = e / q.n, j = e - i * q.n;
  const int32_t *tab_n = q.tab + q.m, *tab_ka = tab_n + q.n, *tab_kb = tab_ka + q.k;
  const int oa = __ldg(q.tab + i), ob = __ldg(tab_n + j);
  const T *__restrict__ A = static_cast<const T *>(q.a);
  const T *__restrict__ B = static_cast<const T *>(q.b);
  W v = wzero((W *)nullptr);
  for (int pr = 0; pr < q.n_pairs; ++pr) {
    const T *Ab = A + (int64_t)q.pair_a[(int64_t)g * q.n_pairs + pr] * q.a_bs + oa;
    const T *Bb = B + (int64_t)q.pair_b[(int64_t)g * q.n_pairs + pr] * q.b_bs + ob;
    for (int kk = 0; kk < q.k; ++kk) wfma(v, Ab[__ldg(tab_ka + kk)], Bb[__ldg(tab_kb + kk)]);
  }
  T *dst = static_cast<T *>(q.c) + (int64_t)g * q.c_bs + e;
  if (q.accumulate) v = wadd(v, widen(*dst));
  *dst = narrow(v, (T *)nullptr);
}

template <typename T>
int launch_small_nd(const NdParams &p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_out * p.m * p.n;
  small_nd_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

int pad124(int64_t x) { return x <= 1 ? 1 : (x <= 2 ? 2 : 4); }

int skinny_splits(const rnb_contract_desc *d) {
  const int64_t k_total = (int64_t)d->n_pairs * d->k;
  // >= 8192 k per CTA (32 per thread) unless that leaves SMs idle; at most ~4 CTAs per SM overall
  int64_t s = (k_total + 8191) / 8192;
  const int64_t cap = std::max<int64_t>(1, (int64_t)sm_count() * 4 / std::max<int64_t>(d->n_out, 1));
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return (int)s;
}

void fill_params(const rnb_contract_desc *d, SimtParams &p) {
  p.a = d->a; p.b = d->b; p.c = d->c;
  p.a_rs = d->a_major == RNB_MAJOR_K ? d->a_ld : 1;
  p.a_ks = d->a_major == RNB_MAJOR_K ? 1 : d->a_ld;
  p.b_rs = d->b_major == RNB_MAJOR_K ? d->b_ld : 1;
  p.b_ks = d->b_major == RNB_MAJOR_K ? 1 : d->b_ld;
  p.a_bs = d->a_block_stride; p.b_bs = d->b_block_stride;
  p.c_ld = d->c_ld; p.c_bs = d->c_block_stride;
  p.k = d->k; p.k_total = (int64_t)d->n_pairs * d->k;
  p.pair_a = d->pair_a; p.pair_b = d->pair_b; p.out_index = d->out_index;
  p.m = (int32_t)d->m; p.n = (int32_t)d->n; p.n_out = d->n_out; p.n_pairs = d->n_pairs;
  p.accumulate = d->accumulate;
  p.splits = 1; p.k_per_split = p.k_total; p.partial = nullptr;
}

template <typename T, int MT, int NT>
int launch_skinny(const SimtParams &p, cudaStream_t stream) {
  skinny_kernel<T, MT, NT><<<(unsigned)((int64_t)p.n_out * p.splits), kSkinnyThreads, 0, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  const int64_t total = (int64_t)p.n_out * MT * NT;
  skinny_finish_kernel<T, MT, NT><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

template <typename T, int MT>
int launch_skinny_n(const SimtParams &p, int nt, cudaStream_t stream) {
  if (nt == 1) return launch_skinny<T, MT, 1>(p, stream);
  if (nt == 2) return launch_skinny<T, MT, 2>(p, stream);
  return launch_skinny<T, MT, 4>(p, stream);
}

template <typename T>
int launch_skinny_mn(const SimtParams &p, int mt, int nt, cudaStream_t stream) {
  if (mt == 1) return launch_skinny_n<T, 1>(p, nt, stream);
  if (mt == 2) return launch_skinny_n<T, 2>(p, nt, stream);
  return launch_skinny_n<T, 4>(p, nt, stream);
}

template <typename T>
int launch_small(const SimtParams &p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_out * p.m * p.n;
  small_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

// ---- apply: one operand tiny, the other one large and left in its n-d layout -------------------------------------
// Contraction trees that keep their intermediates narrow (tn/path.py: search_path) turn a circuit into a sequence of "gate
// applications": C[i][j] = sum_k S[i][k] G[j][k] with S a 2x2 .. 16x16 block and G a tensor of 2^16 .. 2^28 elements whose
// contracted axes sit anywhere in its layout.  Arithmetic intensity ~2 flop/B: the step is worth exactly one read of G and
// one write of C.  The tensor-core route spends a matricize + split pass on G (read 1x, write 3x), a 128x256-tile GEMM that
// is >90 % padding and a combine pass.  Here one thread owns one free index j of G: it gathers the k values G[j][k] through
// a k-offset table, multiplies by the S tile held in shared memory (warp-uniform broadcast reads) and writes the `ns`
// outputs; consecutive threads own consecutive j, so for every i the warp writes one contiguous run.  Every thread decodes
// its own j over the (merged) free axes of G - shifts and masks when all extents are powers of two, as in circuits.
// HBM-bound: algorithmic bytes = elem_bytes * (n_big * k + n_big * ns) per launch.
constexpr int kApplyThreads = 256;
constexpr int kApplyMaxK = 64;
constexpr int kApplyMaxTile = 1024;    // n_pairs * NS * k elements of the small operand held in shared memory
struct ApplyParams {
  const void *s, *g;
  void *c;
  int64_t s_bs, g_bs, c_bs;
  const int32_t *pair_s, *pair_g, *out_index;
  int64_t c_stride_small, c_stride_big;     // output element strides of the small / big free index
  int64_t n_big;                            // free extent of the big operand
  int32_t ns, k, n_out, n_pairs, accumulate, pow2;
  int32_t rank_s, rank_k, rank_g;
  int32_t ext_s[RNB_MAX_RANK], str_s[RNB_MAX_RANK];
  int32_t ext_k[RNB_MAX_RANK], sk_s[RNB_MAX_RANK], sk_g[RNB_MAX_RANK];
  int32_t ext_g[RNB_MAX_RANK], str_g[RNB_MAX_RANK], shift_g[RNB_MAX_RANK];   // free axes of the big operand (last fastest); log2(ext) when pow2
};

// Accumulator of the apply kernel: the operand's own precision (k <= 64 products per output - float32 accumulation is what the
// reference's BLAS call does too; the wide accumulators of the other SIMT routes cost F2F conversions on the XU pipe here).
template <typename T> struct Native { typedef T type; };
__device__ __forceinline__ float wzero(float *) { return 0.f; }
__device__ __forceinline__ float2 wzero(float2 *) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ void wfma(float &acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ __forceinline__ void wfma(float2 &acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ float wadd(float a, float b) { return a + b; }
__device__ __forceinline__ float2 wadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float same(float v) { return v; }
__device__ __forceinline__ double same(double v) { return v; }
__device__ __forceinline__ float2 same(float2 v) { return v; }
__device__ __forceinline__ double2 same(double2 v) { return v; }

// NS: rows of the small operand padded to 2 / 4 / 8 / 16; J: free indices of the big operand per thread (j, j + 256, ...):
// the S tile value read from shared memory is used J times and J * k loads are in flight per thread.
template <typename T, int NS, int J>
__global__ void __launch_bounds__(kApplyThreads) apply_nd_kernel(const __grid_constant__ ApplyParams p) {
  typedef T W;
  __shared__ W tile[kApplyMaxTile];          // [pair][k][NS], rows >= ns are zero
  __shared__ int32_t kg[kApplyMaxK];
  const int g = blockIdx.y;
  const T *__restrict__ S = static_cast<const T *>(p.s);
  const T *__restrict__ G = static_cast<const T *>(p.g);
  for (int t = threadIdx.x; t < p.n_pairs * NS * p.k; t += kApplyThreads) {
    const int pr = t / (NS * p.k), r = t - pr * NS * p.k;
    const int kk = r / NS, i = r - kk * NS;
    W v = wzero((W *)nullptr);
    if (i < p.ns)
      v = S[(int64_t)p.pair_s[(int64_t)g * p.n_pairs + pr] * p.s_bs + nd_offset(i, p.rank_s, p.ext_s, p.str_s) +
            nd_offset(kk, p.rank_k, p.ext_k, p.sk_s)];
    tile[t] = v;
  }
  for (int t = threadIdx.x; t < p.k; t += kApplyThreads) kg[t] = nd_offset(t, p.rank_k, p.ext_k, p.sk_g);
  __syncthreads();
  const int out_blk = p.out_index ? p.out_index[g] : g;
  T *__restrict__ C = static_cast<T *>(p.c) + (int64_t)out_blk * p.c_bs;
  const bool vec = p.c_stride_small == 1 && (NS * sizeof(T)) % 16 == 0 && p.ns == NS && !p.accumulate &&
                   (p.c_stride_big * sizeof(T)) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
  // persistent CTAs (the prologue above is a chain of dependent global loads: ~2 us, amortised over many units of 256 * J indices)
  const int64_t n_units = (p.n_big + kApplyThreads * J - 1) / (kApplyThreads * J);
  for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int64_t j0 = unit * (kApplyThreads * J) + threadIdx.x;
    int64_t og[J];
#pragma unroll
    for (int u = 0; u < J; ++u) {
      int64_t j = j0 + u * kApplyThreads;
      if (j >= p.n_big) j = p.n_big - 1;       // clamped: loads stay in bounds, the store is skipped
      uint32_t idx = (uint32_t)j;
      int64_t off = 0;
      if (p.pow2) {
        for (int d = p.rank_g - 1; d >= 0; --d) {
          off += (int64_t)(idx & (uint32_t)(p.ext_g[d] - 1)) * p.str_g[d];
          idx >>= p.shift_g[d];
        }
      } else {
        for (int d = p.rank_g - 1; d >= 0; --d) {
          const uint32_t e = (uint32_t)p.ext_g[d], q = idx / e;
          off += (int64_t)(idx - q * e) * p.str_g[d];
          idx = q;
        }
      }
      og[u] = off;
    }
    W acc[J][NS];
#pragma unroll
    for (int u = 0; u < J; ++u)
#pragma unroll
      for (int i = 0; i < NS; ++i) acc[u][i] = wzero((W *)nullptr);
    for (int pr = 0; pr < p.n_pairs; ++pr) {
      const T *Gb = G + (int64_t)p.pair_g[(int64_t)g * p.n_pairs + pr] * p.g_bs;
      const W *tl = tile + pr * NS * p.k;
      // four k values at a time: their 4 * J loads are all issued before the first product needs one (a load per trip
      // of a run-time k loop left ONE load in flight per thread: ncu had 62 % of the stalls on the scoreboard of that load)
      for (int kk0 = 0; kk0 < p.k; kk0 += 4) {
        W gv[4][J];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ko = kg[min(kk0 + q, p.k - 1)];
#pragma unroll
          for (int u = 0; u < J; ++u) gv[q][u] = Gb[og[u] + ko];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (kk0 + q < p.k) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              const W t = tl[(kk0 + q) * NS + i];
#pragma unroll
              for (int u = 0; u < J; ++u) wfma(acc[u][i], t, gv[q][u]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < J; ++u) {
      const int64_t j = j0 + u * kApplyThreads;
      if (j >= p.n_big) continue;
      T *dst = C + j * p.c_stride_big;
      if (vec) {
        constexpr int NV = (NS * sizeof(T)) / 16 > 0 ? (NS * sizeof(T)) / 16 : 1;
#pragma unroll
        for (int v = 0; v < NV; ++v) reinterpret_cast<uint4 *>(dst)[v] = reinterpret_cast<const uint4 *>(&acc[u][0])[v];
      } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          if (i < p.ns) {
            W v = acc[u][i];
            T *d = dst + i * p.c_stride_small;
            if (p.accumulate) v = wadd(v, *d);
            *d = v;
          }
        }
      }
    }
  }
}

template <typename T, int NS, int J>
int launch_apply_nsj(const ApplyParams &p, cudaStream_t stream) {
  const int64_t per_unit = (int64_t)kApplyThreads * J;
  int64_t gx = (p.n_big + per_unit - 1) / per_unit;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)p.n_out);
  apply_nd_kernel<T, NS, J><<<grid, kApplyThreads, 0, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

template <typename T>
int launch_apply(const ApplyParams &p, cudaStream_t stream) {
  // J = 2 indices per thread once there are >= 16 units per CTA at J = 2 (RNB_APPLY_J overrides: experiments)
  int jj = p.n_big >= (int64_t)sm_count() * 8 * kApplyThreads * 2 * 4 ? 2 : 1;
  if (const char *e = ([] { static EnvKnob kn("RNB_APPLY_J"); return kn.get(); }())) jj = atoi(e);
  if (p.ns <= 2) return jj >= 4 ? launch_apply_nsj<T, 2, 4>(p, stream) : jj == 2 ? launch_apply_nsj<T, 2, 2>(p, stream) : launch_apply_nsj<T, 2, 1>(p, stream);
  if (p.ns <= 4) return jj >= 4 ? launch_apply_nsj<T, 4, 4>(p, stream) : jj == 2 ? launch_apply_nsj<T, 4, 2>(p, stream) : launch_apply_nsj<T, 4, 1>(p, stream);
  if (p.ns <= 8) return jj >= 2 ? launch_apply_nsj<T, 8, 2>(p, stream) : launch_apply_nsj<T, 8, 1>(p, stream);
  return launch_apply_nsj<T, 16, 1>(p, stream);
}

// Whether rnb_contract_nd takes the apply kernel for (m, n, k, n_out, n_pairs); small_is_a: the tiny operand is a.
bool apply_eligible(int64_t m, int64_t n, int64_t k, int64_t n_out, int64_t n_pairs, bool *small_is_a) {
  const int64_t ns = m <= n ? m : n;
  if (ns > 16 || k > kApplyMaxK || n_out > 65535) return false;
  const int64_t nsp = ns <= 2 ? 2 : (ns <= 4 ? 4 : (ns <= 8 ? 8 : 16));
  if (n_pairs * nsp * k > kApplyMaxTile) return false;
  *small_is_a = m <= n;
  return true;
}

// Free axes of the big operand (row-major, last fastest) as the kernel decodes them; false if an extent does not fit.
bool apply_free_axes(int rank, const int64_t *ext, const int64_t *stride, ApplyParams &p) {
  p.rank_g = rank;
  p.pow2 = 1;
  for (int d = 0; d < rank; ++d) {
    p.ext_g[d] = (int32_t)ext[d];
    p.str_g[d] = (int32_t)stride[d];
    int sh = 0;
    while ((1ll << sh) < ext[d]) ++sh;
    p.shift_g[d] = sh;
    if ((1ll << sh) != ext[d]) p.pow2 = 0;
  }
  return true;
}

}  // namespace

// 0: tensor-core path, 1: skinny, 2: small
int simt_route(const rnb_contract_desc *d) {
  if (d->c_peers) return 0;
  if (const char *e = ([] { static EnvKnob k("RNB_SIMT"); return k.get(); }())) {
    if (!strcmp(e, "0")) return 0;
  }
  const int64_t k_total = (int64_t)d->n_pairs * d->k;
  const int64_t outs = (int64_t)d->n_out * d->m * d->n;
  // small: every thread walks <= 512 k and the whole launch is <= 2^22 multiply-adds
  if (k_total <= 512 && outs <= (1 << 20) && outs * k_total <= (1ll << 22)) return 2;
  if (d->m <= 4 && d->n <= 4 && (int64_t)d->n_out * 16 < (1ll << 31)) return 1;
  return 0;
}

int contract_simt_workspace(const rnb_contract_desc *d, size_t *bytes) {
  *bytes = 0;
  if (simt_route(d) == 1) {
    const int mt = pad124(d->m), nt = pad124(d->n);
    *bytes = (size_t)d->n_out * skinny_splits(d) * mt * nt * 16;
  }
  return RNB_OK;
}

int contract_nd(const rnb_contract_nd_desc *d, cudaStream_t stream) {
  NdParams p;
  memset(&p, 0, sizeof(p));
  int64_t m = 1, n = 1, k = 1;
  RNB_REQUIRE(d->rank_m >= 0 && d->rank_m <= RNB_MAX_RANK && d->rank_n >= 0 && d->rank_n <= RNB_MAX_RANK && d->rank_k >= 0 &&
                  d->rank_k <= RNB_MAX_RANK, "axis counts must be within [0, RNB_MAX_RANK]");
  auto fits = [](int64_t v) { return v >= -(1ll << 30) && v <= (1ll << 30); };
  for (int i = 0; i < d->rank_m; ++i) {
    RNB_REQUIRE(d->ext_m[i] >= 1 && fits(d->ext_m[i]) && fits(d->a_stride_m[i]), "free axis %d of a out of range", i);
    p.ext_m[i] = (int32_t)d->ext_m[i]; p.sa_m[i] = (int32_t)d->a_stride_m[i]; m *= d->ext_m[i];
  }
  for (int i = 0; i < d->rank_n; ++i) {
    RNB_REQUIRE(d->ext_n[i] >= 1 && fits(d->ext_n[i]) && fits(d->b_stride_n[i]), "free axis %d of b out of range", i);
    p.ext_n[i] = (int32_t)d->ext_n[i]; p.sb_n[i] = (int32_t)d->b_stride_n[i]; n *= d->ext_n[i];
  }
  for (int i = 0; i < d->rank_k; ++i) {
    RNB_REQUIRE(d->ext_k[i] >= 1 && fits(d->ext_k[i]) && fits(d->a_stride_k[i]) && fits(d->b_stride_k[i]), "contracted axis %d out of range", i);
    p.ext_k[i] = (int32_t)d->ext_k[i]; p.sa_k[i] = (int32_t)d->a_stride_k[i]; p.sb_k[i] = (int32_t)d->b_stride_k[i]; k *= d->ext_k[i];
  }
  RNB_REQUIRE(d->n_out >= 0 && d->n_pairs >= 1 && d->a && d->b && d->c && d->pair_a && d->pair_b, "bad descriptor");
  if (d->n_out == 0) return RNB_OK;
  // one operand tiny, the other one large: the streaming apply kernel (no tables, one read of the big operand, one write of C)
  {
    bool small_is_a = false;
    const bool has_tabs = d->tab_m || d->tab_n || d->tab_ka || d->tab_kb;
    const char *e_apply = ([] { static EnvKnob kn("RNB_APPLY"); return kn.get(); }());
    const bool allowed = !(e_apply && !strcmp(e_apply, "0"));
    // (what the one-thread-per-element kernel below is good at stays there: <= 2^20 outputs and <= 2^22 multiply-adds)
    const int64_t outs = m * n * (int64_t)d->n_out, k_total = k * (int64_t)d->n_pairs;
    const bool small = k_total <= kNdMaxK && outs <= (1ll << 20) && outs * k_total <= (1ll << 22);
    if (allowed && !has_tabs && !small && apply_eligible(m, n, k, d->n_out, d->n_pairs, &small_is_a)) {
      ApplyParams q;
      memset(&q, 0, sizeof(q));
      const int64_t n_big = small_is_a ? n : m;
      const bool axes_ok = small_is_a ? apply_free_axes(d->rank_n, d->ext_n, d->b_stride_n, q)
                                      : apply_free_axes(d->rank_m, d->ext_m, d->a_stride_m, q);
      if (axes_ok && n_big < (1ll << 31)) {
        q.n_big = n_big;
        q.s = small_is_a ? d->a : d->b; q.g = small_is_a ? d->b : d->a; q.c = d->c;
        q.s_bs = small_is_a ? d->a_block_stride : d->b_block_stride;
        q.g_bs = small_is_a ? d->b_block_stride : d->a_block_stride;
        q.c_bs = d->c_block_stride;
        q.pair_s = small_is_a ? d->pair_a : d->pair_b; q.pair_g = small_is_a ? d->pair_b : d->pair_a; q.out_index = d->out_index;
        q.ns = (int32_t)(small_is_a ? m : n); q.k = (int32_t)k; q.n_out = d->n_out; q.n_pairs = d->n_pairs; q.accumulate = d->accumulate;
        q.c_stride_small = small_is_a ? n : 1;       // C is row-major [m][n]
        q.c_stride_big = small_is_a ? 1 : n;
        q.rank_s = small_is_a ? d->rank_m : d->rank_n; q.rank_k = d->rank_k;
        for (int i = 0; i < q.rank_s; ++i) {
          q.ext_s[i] = small_is_a ? p.ext_m[i] : p.ext_n[i];
          q.str_s[i] = small_is_a ? p.sa_m[i] : p.sb_n[i];
        }
        for (int i = 0; i < q.rank_k; ++i) {
          q.ext_k[i] = p.ext_k[i];
          q.sk_s[i] = small_is_a ? p.sa_k[i] : p.sb_k[i];
          q.sk_g[i] = small_is_a ? p.sb_k[i] : p.sa_k[i];
        }
        switch (d->dtype) {
          case RNB_F32: return launch_apply<float>(q, stream);
          case RNB_F64: return launch_apply<double>(q, stream);
          case RNB_C64: return launch_apply<float2>(q, stream);
          case RNB_C128: return launch_apply<double2>(q, stream);
          default: set_error("unknown dtype %d", d->dtype); return RNB_ERR_INVALID;
        }
      }
    }
  }
  RNB_REQUIRE(k <= kNdMaxK, "rnb_contract_nd: contracted extent %lld exceeds %d (use rnb_contract_grouped)", (long long)k, kNdMaxK);
  RNB_REQUIRE(m * n * (int64_t)d->n_out <= (1ll << 24), "rnb_contract_nd: too many output elements for the one-thread-per-element kernel");
  p.a = d->a; p.b = d->b; p.c = d->c;
  p.a_bs = d->a_block_stride; p.b_bs = d->b_block_stride; p.c_bs = d->c_block_stride;
  p.pair_a = d->pair_a; p.pair_b = d->pair_b; p.out_index = d->out_index;
  p.m = (int32_t)m; p.n = (int32_t)n; p.k = (int32_t)k; p.n_out = d->n_out; p.n_pairs = d->n_pairs; p.accumulate = d->accumulate;
  p.rank_m = d->rank_m; p.rank_n = d->rank_n; p.rank_k = d->rank_k;
  if (d->tab_m || d->tab_n || d->tab_ka || d->tab_kb) {
    RNB_REQUIRE(d->tab_m && d->tab_n && d->tab_ka && d->tab_kb, "rnb_contract_nd: pass all four offset tables or none");
    NdTabParams q;
    q.a = p.a; q.b = p.b; q.c = p.c; q.a_bs = p.a_bs; q.b_bs = p.b_bs; q.c_bs = p.c_bs;
    q.pair_a = p.pair_a; q.pair_b = p.pair_b; q.out_index = p.out_index;
    q.tab_m = d->tab_m; q.tab_n = d->tab_n; q.tab_ka = d->tab_ka; q.tab_kb = d->tab_kb;
    q.m = p.m; q.n = p.n; q.k = p.k; q.n_out = p.n_out; q.n_pairs = p.n_pairs; q.accumulate = p.accumulate;
    switch (d->dtype) {
      case RNB_F32: return launch_small_tab<float>(q, stream);
      case RNB_F64: return launch_small_tab<double>(q, stream);
      case RNB_C64: return launch_small_tab<float2>(q, stream);
      case RNB_C128: return launch_small_tab<double2>(q, stream);
      default: set_error("unknown dtype %d", d->dtype); return RNB_ERR_INVALID;
    }
  }
  switch (d->dtype) {
    case RNB_F32: return launch_small_nd<float>(p, stream);
    case RNB_F64: return launch_small_nd<double>(p, stream);
    case RNB_C64: return launch_small_nd<float2>(p, stream);
    case RNB_C128: return launch_small_nd<double2>(p, stream);
    default: set_error("unknown dtype %d", d->dtype); return RNB_ERR_INVALID;
  }
}

int contract_nd_batch(const rnb_contract_nd_desc *descs, int32_t n, cudaStream_t stream) {
  RNB_REQUIRE(descs != nullptr && n >= 0, "bad arguments");
  static BatchParams bp;     // 32 KB: kept off the stack; the mutex serialises concurrent callers
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int done = 0;
  while (done < n) {
    const int dtype = descs[done].dtype;
    int count = 0;
    int64_t ctas = 0;
    while (done + count < n && count < kBatchMax && descs[done + count].dtype == dtype) {
      const rnb_contract_nd_desc &d = descs[done + count];
      RNB_REQUIRE(d.tab_m && d.tab_n && d.tab_ka && d.tab_kb, "rnb_contract_nd_batch needs the offset tables");
      int64_t m = 1, nn = 1, k = 1;
      for (int i = 0; i < d.rank_m; ++i) m *= d.ext_m[i];
      for (int i = 0; i < d.rank_n; ++i) nn *= d.ext_n[i];
      for (int i = 0; i < d.rank_k; ++i) k *= d.ext_k[i];
      RNB_REQUIRE(d.tab_n == d.tab_m + m && d.tab_ka == d.tab_n + nn && d.tab_kb == d.tab_ka + k,
                  "rnb_contract_nd_batch: the four offset tables of a task must lie back to back (m, n, k, k entries)");
      RNB_REQUIRE(d.out_index == nullptr, "rnb_contract_nd_batch: out_index is not supported");
      RNB_REQUIRE(d.a && d.b && d.c && d.pair_a && d.pair_b && d.n_out >= 1 && d.n_pairs >= 1, "bad descriptor in batch");
      RNB_REQUIRE(d.a_block_stride < (1ll << 31) && d.b_block_stride < (1ll << 31) && d.c_block_stride < (1ll << 31), "block stride too large");
      const int64_t total = (int64_t)d.n_out * m * nn;
      RNB_REQUIRE(total <= (1ll << 24) && k <= kNdMaxK, "task too large for the small-contraction kernel");
      if (ctas + (total + 127) / 128 > (1 << 30)) break;
      BatchTask &t = bp.t[count];
      t.a = d.a; t.b = d.b; t.c = d.c; t.tab = d.tab_m; t.pair_a = d.pair_a; t.pair_b = d.pair_b;
      t.m = (int32_t)m; t.n = (int32_t)nn; t.k = (int32_t)k; t.n_out = d.n_out; t.n_pairs = d.n_pairs;
      t.a_bs = (int32_t)d.a_block_stride; t.b_bs = (int32_t)d.b_block_stride; t.c_bs = (int32_t)d.c_block_stride;
      t.cta0 = (int32_t)ctas; t.accumulate = d.accumulate;
      ctas += (total + 127) / 128;
      ++count;
    }
    RNB_REQUIRE(count > 0, "empty batch");
    bp.n_tasks = count;
    switch (dtype) {
      case RNB_F32: small_batch_kernel<float><<<(unsigned)ctas, 128, 0, stream>>>(bp); break;
      case RNB_F64: small_batch_kernel<double><<<(unsigned)ctas, 128, 0, stream>>>(bp); break;
      case RNB_C64: small_batch_kernel<float2><<<(unsigned)ctas, 128, 0, stream>>>(bp); break;
      case RNB_C128: small_batch_kernel<double2><<<(unsigned)ctas, 128, 0, stream>>>(bp); break;
      default: set_error("unknown dtype %d", dtype); return RNB_ERR_INVALID;
    }
    RNB_CHECK_CUDA(cudaGetLastError());
    done += count;
  }
  return RNB_OK;
}

int contract_simt(const rnb_contract_desc *d, cudaStream_t stream) {
  const int route = simt_route(d);
  SimtParams p;
  fill_params(d, p);
  if (route == 2) {
    switch (d->dtype) {
      case RNB_F32: return launch_small<float>(p, stream);
      case RNB_F64: return launch_small<double>(p, stream);
      case RNB_C64: return launch_small<float2>(p, stream);
      default: return launch_small<double2>(p, stream);
    }
  }
  const int mt = pad124(d->m), nt = pad124(d->n);
  p.splits = skinny_splits(d);
  p.k_per_split = (p.k_total + p.splits - 1) / p.splits;
  const size_t need = (size_t)d->n_out * p.splits * mt * nt * 16;
  if (!d->workspace || d->workspace_bytes < need) {
    set_error("workspace of %zu bytes required (rnb_contract_workspace_bytes), got %zu", need, d->workspace_bytes);
    return RNB_ERR_WORKSPACE;
  }
  RNB_REQUIRE((reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0, "workspace must be 16-byte aligned");
  p.partial = d->workspace;
  switch (d->dtype) {
    case RNB_F32: return launch_skinny_mn<float>(p, mt, nt, stream);
    case RNB_F64: return launch_skinny_mn<double>(p, mt, nt, stream);
    case RNB_C64: return launch_skinny_mn<float2>(p, mt, nt, stream);
    default: return launch_skinny_mn<double2>(p, mt, nt, stream);
  }
}

}  // namespace rnb
