// Grouped block GEMM for float64 / complex128 with the k-block sum fused in-kernel.
//
// Replaces, for every output block of np.tensordot(BlockArray, BlockArray, axes), the reference's
//     sum(np.tensordot(ai, bi, axes) for ai, bi in zip(a, b))        (rosnet/array/block/__init__.py:281-283)
// i.e. n_pairs BLAS dgemm/zgemm calls plus n_pairs full-block additions, with ONE accumulator
// chain held in registers across the whole pair list.
//
// sm_100a has no tcgen05 kind for FP64 (ptxas rejects .kind::f64); FP64 tensor math is
// mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4.  Design:
//   * persistent CTAs (one per SM), static tile schedule over (output block, tile_m, tile_n);
//   * one producer warp feeds a 4-stage TMA -> shared memory ring guarded by mbarriers;
//     8 consumer warps issue DMMA with register accumulators;
//   * operands are consumed in place in either storage order.  The tensor maps split the
//     matrix into (4 k) x rows slabs [K-major] or (32-byte row group) x k slabs [MN-major] by giving the TMA a
//     non-monotonic stride order, so that the box lands in shared memory already in
//     "fragment order": every warp-wide fragment load is one contiguous 256 B (real) or 512 B
//     (complex) run -> bank-conflict free without swizzling or padding;
//   * complex128 is computed on the interleaved data directly as 4 real DMMA per tile product
//     (4M): re += ar*br, re += (-ai)*bi, im += ar*bi, im += ai*br; or as 3 (3M, RNB_CPLX_3M):
//     T1 += ar*br, T2 += ai*bi, T3 += (ar+ai)*(br+bi), re = T1-T2, im = T3-T1-T2 (3/4 of the DMMAs);
//   * epilogue writes the m x n output block row-major straight from the accumulators
//     (16 B per lane, 64 B contiguous per fragment row).
#include "common.h"

namespace rnb {

namespace {

constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;
// 2 consumer warpgroups + 1 producer warpgroup (only its first lane works); the producer group
// hands its registers to the consumers with setmaxnreg, as register accumulators need them.
constexpr int kThreads = (kConsumerWarps + 4) * 32;

// MODE 0: float64;  1: complex128 as 4 real DMMA per tile product (4M);  2: complex128 as 3 (3M):
//   T1 += ar.br, T2 += ai.bi, T3 += (ar+ai).(br+bi);  re = T1 - T2, im = T3 - T1 - T2
// 3M keeps three accumulator sets, so its warp tile is 32x16 (96 accumulator registers).
template <int MODE>
struct Cfg {
  static constexpr bool CPLX = MODE != 0;
  static constexpr int BM = 128;
  static constexpr int BN = MODE == 0 ? 128 : (MODE == 1 ? 64 : 32);
  static constexpr int BK = CPLX ? 8 : 16;  // elements of the dtype per stage
  static constexpr int NACC = MODE == 0 ? 2 : (MODE == 1 ? 4 : 6);
  static constexpr int WARPS_M = 4, WARPS_N = 2;
  static constexpr int WM = BM / WARPS_M;  // 32
  static constexpr int WN = BN / WARPS_N;  // 64 (real) / 32 (complex)
  static constexpr int MT = WM / 8, NT = WN / 8;
  static constexpr int ELEM = CPLX ? 16 : 8;
  static constexpr int GS = CPLX ? 2 : 4;  // rows per MN-major slab group (32 bytes)
  static constexpr int A_STAGE = BM * BK * ELEM;
  static constexpr int B_STAGE = BN * BK * ELEM;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int SMEM = kStages * STAGE + 1024;
};

struct GemmParams {
  int64_t m, n, k;
  int64_t c_ld, c_block_stride;
  void *c;
  const int32_t *pair_a;
  const int32_t *pair_b;
  const int32_t *out_index;
  int32_t n_out, n_pairs;
  int32_t tiles_m, tiles_n;
  int32_t k_stages;
  int32_t accumulate;
  int32_t c_vec_ok;
  int32_t n_peers, blocks_per_peer;  // fused exchange: reduce into the owning rank's buffer
  double *peers[RNB_MAX_PEERS];
  // split-K (launches with fewer tiles than SMs and a long contracted extent): unit = tile * splits + s
  // covers iterations [s * iters_per_split, ...) of the flattened (pair, k-stage) space and writes its
  // partial BM x BN tile to `partial`; splitk_reduce_kernel adds the splits in a fixed order.
  int32_t splits, iters_per_split;
  double *partial;
  int32_t peer_bulk;  // fused exchange: interior warp tiles are pushed with cp.reduce.async.bulk (else scalar red)
};

// per-warp staging for the bulk peer reduction: 8 rows x 512 B
constexpr int kPeerStageBytes = 8 * 512;
constexpr int kPeerStageTotal = kConsumerWarps * kPeerStageBytes;

__device__ __forceinline__ void bulk_reduce_add_f64(void *gmem_dst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// element index inside one operand stage, see file header.  The MN-major slab groups GS rows
// (32 bytes: 4 doubles or 2 complex) so that the 16 lanes of a half warp (8 of a quarter warp for
// 16-byte elements) read one contiguous 128-byte run.
template <int MAJOR, int ROWS, int BK, int GS>
__device__ __forceinline__ int frag_index(int row, int k) {
  if (MAJOR == RNB_MAJOR_K) return (((k >> 2) * ROWS + row) << 2) + (k & 3);
  return ((row / GS) * BK + k) * GS + (row % GS);
}

template <int MODE, int AMAJ, int BMAJ>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_f64_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
  using C = Cfg<MODE>;
  constexpr bool CPLX = C::CPLX;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kStages * C::STAGE);
  uint64_t *empty_bar = full_bar + kStages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int tiles_per_block = p.tiles_m * p.tiles_n;
  const int total_tiles = p.n_out * tiles_per_block * p.splits;  // units
  const int iters_per_tile = p.n_pairs * p.k_stages;

  if (warp >= kConsumerWarps) {
    // ===== TMA producer warpgroup =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kConsumerWarps && lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
      uint32_t it = 0;
      for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x) {
        const int tile = unit / p.splits;
        const int sp = unit - tile * p.splits;
        const int g = tile / tiles_per_block;
        const int rem = tile - g * tiles_per_block;
        const int tm = rem / p.tiles_n;
        const int tn = rem - tm * p.tiles_n;
        const int m0 = tm * C::BM, n0 = tn * C::BN;
        const int it0 = sp * p.iters_per_split;
        const int it1 = min(it0 + p.iters_per_split, iters_per_tile);
        int pr = it0 / p.k_stages;
        int ks = it0 - pr * p.k_stages;
        int blk_a = p.pair_a[(int64_t)g * p.n_pairs + pr];
        int blk_b = p.pair_b[(int64_t)g * p.n_pairs + pr];
        for (int iter = it0; iter < it1; ++iter, ++it) {
          const int stage = it % kStages;
          const uint32_t phase = (it / kStages) & 1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE);
          unsigned char *sa = smem + stage * C::STAGE;
          unsigned char *sb = sa + C::A_STAGE;
          const int k0 = ks * C::BK;
          if (AMAJ == RNB_MAJOR_K)
            tma_load_4d(sa, &map_a, &full_bar[stage], 0, m0, k0 >> 2, blk_a);
          else
            tma_load_4d(sa, &map_a, &full_bar[stage], 0, k0, m0 / C::GS, blk_a);
          if (BMAJ == RNB_MAJOR_K)
            tma_load_4d(sb, &map_b, &full_bar[stage], 0, n0, k0 >> 2, blk_b);
          else
            tma_load_4d(sb, &map_b, &full_bar[stage], 0, k0, n0 / C::GS, blk_b);
          if (++ks == p.k_stages && iter + 1 < it1) {
            ks = 0;
            ++pr;
            blk_a = p.pair_a[(int64_t)g * p.n_pairs + pr];
            blk_b = p.pair_b[(int64_t)g * p.n_pairs + pr];
          }
        }
      }
    }
    return;
  }

  // ===== DMMA consumers =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int wm = warp % C::WARPS_M;
  const int wn = warp / C::WARPS_M;
  const int r = lane >> 2;   // fragment row (A) / column (B)
  const int kk = lane & 3;   // fragment k

  uint32_t it = 0;
  for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x) {
    const int tile = unit / p.splits;
    const int sp = unit - tile * p.splits;
    const int g = tile / tiles_per_block;
    const int rem = tile - g * tiles_per_block;
    const int tm = rem / p.tiles_n;
    const int tn = rem - tm * p.tiles_n;
    const int it0 = sp * p.iters_per_split;
    const int it1 = min(it0 + p.iters_per_split, iters_per_tile);

    double acc[C::MT][C::NT][C::NACC];
#pragma unroll
    for (int i = 0; i < C::MT; ++i)
#pragma unroll
      for (int j = 0; j < C::NT; ++j)
#pragma unroll
        for (int e = 0; e < C::NACC; ++e) acc[i][j][e] = 0.0;

    for (int iter = it0; iter < it1; ++iter, ++it) {
      const int stage = it % kStages;
      const uint32_t phase = (it / kStages) & 1;
      mbar_wait(&full_bar[stage], phase);
      const unsigned char *sa = smem + stage * C::STAGE;
      const unsigned char *sb = sa + C::A_STAGE;
      if (!CPLX) {
        const double *A = reinterpret_cast<const double *>(sa);
        const double *B = reinterpret_cast<const double *>(sb);
#pragma unroll
        for (int s = 0; s < C::BK / 4; ++s) {
          double a[C::MT], b[C::NT];
#pragma unroll
          for (int i = 0; i < C::MT; ++i) a[i] = A[frag_index<AMAJ, C::BM, C::BK, C::GS>(wm * C::WM + 8 * i + r, 4 * s + kk)];
#pragma unroll
          for (int j = 0; j < C::NT; ++j) b[j] = B[frag_index<BMAJ, C::BN, C::BK, C::GS>(wn * C::WN + 8 * j + r, 4 * s + kk)];
#pragma unroll
          for (int i = 0; i < C::MT; ++i)
#pragma unroll
            for (int j = 0; j < C::NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      } else if (MODE == 2) {
        const double2 *A = reinterpret_cast<const double2 *>(sa);
        const double2 *B = reinterpret_cast<const double2 *>(sb);
#pragma unroll
        for (int s = 0; s < C::BK / 4; ++s) {
          double2 a[C::MT], b[C::NT];
          double as[C::MT], bs[C::NT];
#pragma unroll
          for (int i = 0; i < C::MT; ++i) {
            a[i] = A[frag_index<AMAJ, C::BM, C::BK, C::GS>(wm * C::WM + 8 * i + r, 4 * s + kk)];
            as[i] = a[i].x + a[i].y;
          }
#pragma unroll
          for (int j = 0; j < C::NT; ++j) {
            b[j] = B[frag_index<BMAJ, C::BN, C::BK, C::GS>(wn * C::WN + 8 * j + r, 4 * s + kk)];
            bs[j] = b[j].x + b[j].y;
          }
#pragma unroll
          for (int i = 0; i < C::MT; ++i)
#pragma unroll
            for (int j = 0; j < C::NT; ++j) {
              dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);  // T1 += ar*br
              dmma884(acc[i][j][2], acc[i][j][3], a[i].y, b[j].y);  // T2 += ai*bi
              dmma884(acc[i][j][4], acc[i][j][5], as[i], bs[j]);    // T3 += (ar+ai)*(br+bi)
            }
        }
      } else {
        const double2 *A = reinterpret_cast<const double2 *>(sa);
        const double2 *B = reinterpret_cast<const double2 *>(sb);
#pragma unroll
        for (int s = 0; s < C::BK / 4; ++s) {
          double2 a[C::MT], b[C::NT];
#pragma unroll
          for (int i = 0; i < C::MT; ++i) a[i] = A[frag_index<AMAJ, C::BM, C::BK, C::GS>(wm * C::WM + 8 * i + r, 4 * s + kk)];
#pragma unroll
          for (int j = 0; j < C::NT; ++j) b[j] = B[frag_index<BMAJ, C::BN, C::BK, C::GS>(wn * C::WN + 8 * j + r, 4 * s + kk)];
#pragma unroll
          for (int i = 0; i < C::MT; ++i)
#pragma unroll
            for (int j = 0; j < C::NT; ++j) {
              dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);  // re += ar*br
              dmma884(acc[i][j][2], acc[i][j][3], a[i].x, b[j].y);  // im += ar*bi
            }
#pragma unroll
          for (int i = 0; i < C::MT; ++i) {
            const double nai = -a[i].y;
#pragma unroll
            for (int j = 0; j < C::NT; ++j) {
              dmma884(acc[i][j][0], acc[i][j][1], nai, b[j].y);     // re -= ai*bi
              dmma884(acc[i][j][2], acc[i][j][3], a[i].y, b[j].x);  // im += ai*br
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }

    // ---- epilogue: accumulators -> row-major output block ----
    const bool split = p.splits > 1;
    const int out_blk = split ? 0 : (p.out_index ? p.out_index[g] : g);
    // split-K: the same epilogue writes the tile, in tile-local coordinates, into the partial buffer
    const int64_t row_base = (split ? 0 : (int64_t)tm * C::BM) + wm * C::WM + r;
    const int64_t col_base = (split ? 0 : (int64_t)tn * C::BN) + wn * C::WN + 2 * kk;
    const int64_t lim_m = split ? min((int64_t)C::BM, p.m - (int64_t)tm * C::BM) : p.m;
    const int64_t lim_n = split ? min((int64_t)C::BN, p.n - (int64_t)tn * C::BN) : p.n;
    const int64_t ld = split ? C::BN : p.c_ld;
    const bool accumulate = !split && p.accumulate;
    const bool vec_ok = split || p.c_vec_ok;
    if (p.n_peers > 0) {
      // ---- fused exchange: add this rank's partial tile into the owner's block over peer memory ----
      const int owner = out_blk / p.blocks_per_peer;
      const int local = out_blk - owner * p.blocks_per_peer;
      constexpr int F = CPLX ? 2 : 1;  // doubles per element
      double *Cb = p.peers[owner] + (int64_t)local * p.c_block_stride * F;
      const int64_t wrow0 = (int64_t)tm * C::BM + wm * C::WM;
      const int64_t wcol0 = (int64_t)tn * C::BN + wn * C::WN;
      if (p.peer_bulk && wrow0 + C::WM <= p.m && wcol0 + C::WN <= p.n) {
        // Interior warp tile: stage 8 rows at a time in shared memory and push every row (WN elements,
        // contiguous in the owner's block) with ONE bulk reduction over NVLink instead of 2 * WN scalar
        // 8-byte red operations (which is what made the fused path lose to NCCL on 8 GPUs).
        double *stg = reinterpret_cast<double *>(smem + kStages * C::STAGE + 1024 + warp * kPeerStageBytes);
        constexpr int ROW_D = C::WN * F;  // doubles per staged row
#pragma unroll
        for (int i = 0; i < C::MT; ++i) {
#pragma unroll
          for (int j = 0; j < C::NT; ++j) {
            double *d = stg + r * ROW_D + (8 * j + 2 * kk) * F;
            if (!CPLX) {
              *reinterpret_cast<double2 *>(d) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else if (MODE == 2) {
              *reinterpret_cast<double2 *>(d) = make_double2(acc[i][j][0] - acc[i][j][2], acc[i][j][4] - acc[i][j][0] - acc[i][j][2]);
              *reinterpret_cast<double2 *>(d + 2) = make_double2(acc[i][j][1] - acc[i][j][3], acc[i][j][5] - acc[i][j][1] - acc[i][j][3]);
            } else {
              *reinterpret_cast<double2 *>(d) = make_double2(acc[i][j][0], acc[i][j][2]);
              *reinterpret_cast<double2 *>(d + 2) = make_double2(acc[i][j][1], acc[i][j][3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane < 8) {
            const int64_t row = wrow0 + 8 * i + lane;
            bulk_reduce_add_f64(Cb + (row * p.c_ld + wcol0) * F, stg + lane * ROW_D, ROW_D * 8);
            bulk_commit_group();
            bulk_wait_group_read<0>();   // the staging rows are rewritten by the next i
          }
          __syncwarp();
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < C::MT; ++i) {
        const int64_t row = row_base + 8 * i;
        if (row >= p.m) continue;
#pragma unroll
        for (int j = 0; j < C::NT; ++j) {
          const int64_t col = col_base + 8 * j;
          if (col >= p.n) continue;
          double *dst = Cb + (row * p.c_ld + col) * F;
          const bool two = col + 1 < p.n;
          if (!CPLX) {
            atomicAdd_system(dst, acc[i][j][0]);
            if (two) atomicAdd_system(dst + 1, acc[i][j][1]);
          } else {
            double re0, im0, re1, im1;
            if (MODE == 2) {
              re0 = acc[i][j][0] - acc[i][j][2]; im0 = acc[i][j][4] - acc[i][j][0] - acc[i][j][2];
              re1 = acc[i][j][1] - acc[i][j][3]; im1 = acc[i][j][5] - acc[i][j][1] - acc[i][j][3];
            } else {
              re0 = acc[i][j][0]; im0 = acc[i][j][2]; re1 = acc[i][j][1]; im1 = acc[i][j][3];
            }
            atomicAdd_system(dst, re0);
            atomicAdd_system(dst + 1, im0);
            if (two) {
              atomicAdd_system(dst + 2, re1);
              atomicAdd_system(dst + 3, im1);
            }
          }
        }
      }
      continue;
    }
    if (!CPLX) {
      double *Cb = split ? p.partial + (int64_t)unit * (C::BM * C::BN)
                         : reinterpret_cast<double *>(p.c) + (int64_t)out_blk * p.c_block_stride;
#pragma unroll
      for (int i = 0; i < C::MT; ++i) {
        const int64_t row = row_base + 8 * i;
        if (row >= lim_m) continue;
#pragma unroll
        for (int j = 0; j < C::NT; ++j) {
          const int64_t col = col_base + 8 * j;
          if (col >= lim_n) continue;
          double *dst = Cb + row * ld + col;
          double v0 = acc[i][j][0], v1 = acc[i][j][1];
          if (col + 1 < lim_n) {
            if (vec_ok) {
              double2 *d2 = reinterpret_cast<double2 *>(dst);
              if (accumulate) {
                double2 old = *d2;
                v0 += old.x;
                v1 += old.y;
              }
              *d2 = make_double2(v0, v1);
            } else {
              if (accumulate) {
                v0 += dst[0];
                v1 += dst[1];
              }
              dst[0] = v0;
              dst[1] = v1;
            }
          } else {
            if (accumulate) v0 += dst[0];
            dst[0] = v0;
          }
        }
      }
    } else {
      double2 *Cb = split ? reinterpret_cast<double2 *>(p.partial) + (int64_t)unit * (C::BM * C::BN)
                          : reinterpret_cast<double2 *>(p.c) + (int64_t)out_blk * p.c_block_stride;
#pragma unroll
      for (int i = 0; i < C::MT; ++i) {
        const int64_t row = row_base + 8 * i;
        if (row >= lim_m) continue;
#pragma unroll
        for (int j = 0; j < C::NT; ++j) {
          const int64_t col = col_base + 8 * j;
          if (col >= lim_n) continue;
          double2 *dst = Cb + row * ld + col;
          double2 v0, v1;
          if (MODE == 2) {
            v0 = make_double2(acc[i][j][0] - acc[i][j][2], acc[i][j][4] - acc[i][j][0] - acc[i][j][2]);
            v1 = make_double2(acc[i][j][1] - acc[i][j][3], acc[i][j][5] - acc[i][j][1] - acc[i][j][3]);
          } else {
            v0 = make_double2(acc[i][j][0], acc[i][j][2]);
            v1 = make_double2(acc[i][j][1], acc[i][j][3]);
          }
          if (accumulate) {
            double2 o = dst[0];
            v0.x += o.x;
            v0.y += o.y;
          }
          dst[0] = v0;
          if (col + 1 < lim_n) {
            if (accumulate) {
              double2 o = dst[1];
              v1.x += o.x;
              v1.y += o.y;
            }
            dst[1] = v1;
          }
        }
      }
    }
  }
  if (p.n_peers > 0 && p.peer_bulk && lane < 8) bulk_wait_group<0>();   // all pushed rows have landed before the CTA exits
}

// Tensor map that makes a TMA box land in "fragment order" (file header).  Units: doubles.
int make_operand_map(CUtensorMap *map, const void *base, bool cplx, int major, int64_t rows, int64_t k, int64_t ld,
                     int64_t block_stride, int nblocks, int tile_rows, int bk, const char *name) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return RNB_ERR_CUDA;
  }
  const int f = cplx ? 2 : 1;  // doubles per element
  const int64_t ebytes = 8 * f;
  RNB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "%s: base pointer must be 16-byte aligned", name);
  RNB_REQUIRE((ld * ebytes) % 16 == 0, "%s: leading dimension %lld gives a row pitch that is not a multiple of 16 bytes", name,
              (long long)ld);
  RNB_REQUIRE(nblocks == 1 || (block_stride * ebytes) % 16 == 0, "%s: block stride must be a multiple of 16 bytes", name);
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint64_t bstride = (nblocks == 1) ? (cuuint64_t)16 : (cuuint64_t)(block_stride * ebytes);
  if (major == RNB_MAJOR_K) {
    RNB_REQUIRE(k % 4 == 0 && k >= 4, "%s: K-major operand needs k %% 4 == 0 (k=%lld); matricize it into a padded workspace", name,
                (long long)k);
    dims[0] = 4 * f;  dims[1] = rows;             dims[2] = k / 4;       dims[3] = nblocks;
    strides[0] = ld * ebytes;  strides[1] = 4 * ebytes;  strides[2] = bstride;
    box[0] = 4 * f;   box[1] = tile_rows;         box[2] = bk / 4;       box[3] = 1;
  } else {
    const int gs = cplx ? 2 : 4;  // rows per slab group: 32 bytes
    RNB_REQUIRE(rows % gs == 0 && rows >= gs, "%s: MN-major operand needs rows %% %d == 0 (rows=%lld); matricize it", name, gs,
                (long long)rows);
    dims[0] = gs * f;  dims[1] = k;               dims[2] = rows / gs;   dims[3] = nblocks;
    strides[0] = ld * ebytes;  strides[1] = gs * ebytes;  strides[2] = bstride;
    box[0] = gs * f;   box[1] = bk;               box[2] = tile_rows / gs; box[3] = 1;
  }
  CUresult res = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu box %u,%u,%u,%u)",
              name, (int)res, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
              (unsigned long long)strides[2], box[0], box[1], box[2], box[3]);
    return RNB_ERR_CUDA;
  }
  return RNB_OK;
}

template <int MODE, int AMAJ, int BMAJ>
int launch(const CUtensorMap &ma, const CUtensorMap &mb, const GemmParams &p, int grid, cudaStream_t stream) {
  using C = Cfg<MODE>;
  auto kern = gemm_f64_kernel<MODE, AMAJ, BMAJ>;
  static DeviceOnce once;  // per template instance and device
  bool &attr_done = once.flag();
  if (!attr_done) {
    RNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + kPeerStageTotal));
    attr_done = true;
  }
  kern<<<grid, kThreads, C::SMEM + (p.peer_bulk ? kPeerStageTotal : 0), stream>>>(ma, mb, p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

template <int MODE>
int dispatch_major(int amaj, int bmaj, const CUtensorMap &ma, const CUtensorMap &mb, const GemmParams &p, int grid,
                   cudaStream_t stream) {
  if (amaj == RNB_MAJOR_K && bmaj == RNB_MAJOR_K) return launch<MODE, RNB_MAJOR_K, RNB_MAJOR_K>(ma, mb, p, grid, stream);
  if (amaj == RNB_MAJOR_K && bmaj == RNB_MAJOR_MN) return launch<MODE, RNB_MAJOR_K, RNB_MAJOR_MN>(ma, mb, p, grid, stream);
  if (amaj == RNB_MAJOR_MN && bmaj == RNB_MAJOR_K) return launch<MODE, RNB_MAJOR_MN, RNB_MAJOR_K>(ma, mb, p, grid, stream);
  return launch<MODE, RNB_MAJOR_MN, RNB_MAJOR_MN>(ma, mb, p, grid, stream);
}

}  // namespace

namespace {
// complex128 arithmetic of a call: 1 = 4M, 2 = 3M.  Unlike the tcgen05 path (csrc/gemm_tf32.cu: use_3m) the three products of
// 3M accumulate in REGISTERS here and are combined in the epilogue, so 3M costs no workspace pass: RNB_CPLX_AUTO takes it
// whenever there is a contracted extent worth a tile loop (measured: C1 complex128 36.0 -> 44.0 TFLOP/s, QFT-30 13.6 -> 11.3 ms,
// 2^14 Schmidt matrix 980 -> 791 ms, every GEMM of the QFT path from k = 64 up; error 3e-15 vs 2.4e-15 for 4M).
int f64_mode(const rnb_contract_desc *d) {
  if (d->dtype != RNB_C128) return 0;
  if (d->complex_mode == RNB_CPLX_3M) return 2;
  if (d->complex_mode == RNB_CPLX_AUTO && (int64_t)d->n_pairs * d->k >= 64) return 2;
  return 1;
}
struct SplitPlan {
  int splits, iters_per_split;
  int64_t tiles;
  size_t bytes;
};
// >= 8 k-stages (about 20 us of DMMA) per split; never with the fused peer exchange
SplitPlan plan_split(const rnb_contract_desc *d) {
  const bool cplx = d->dtype == RNB_C128;
  const int mode = f64_mode(d);
  const int BM = 128, BN = mode == 0 ? 128 : (mode == 1 ? 64 : 32), BK = cplx ? 8 : 16;
  SplitPlan sp;
  sp.tiles = (int64_t)d->n_out * ((d->m + BM - 1) / BM) * ((d->n + BN - 1) / BN);
  const int64_t iters = (int64_t)d->n_pairs * ((d->k + BK - 1) / BK);
  sp.splits = d->c_peers ? 1 : choose_splits(sp.tiles, iters, 8, 48, sm_count());
  sp.iters_per_split = (int)iters;
  if (sp.splits > 1) {
    const int64_t ips = (iters + sp.splits - 1) / sp.splits;
    sp.iters_per_split = (int)ips;
    sp.splits = (int)((iters + ips - 1) / ips);
  }
  sp.bytes = sp.splits > 1 ? (size_t)sp.tiles * sp.splits * BM * BN * (cplx ? 16 : 8) : 0;
  return sp;
}
}  // namespace

int contract_f64_workspace(const rnb_contract_desc *d, size_t *bytes) {
  *bytes = plan_split(d).bytes;
  return RNB_OK;
}

// Entry used by rnb_contract_grouped for RNB_F64 / RNB_C128.
int contract_f64(const rnb_contract_desc *d, cudaStream_t stream) {
  const bool cplx = d->dtype == RNB_C128;
  const int mode = f64_mode(d);
  const int BM = 128, BN = mode == 0 ? 128 : (mode == 1 ? 64 : 32), BK = cplx ? 8 : 16;
  CUtensorMap ma, mb;
  int rc = make_operand_map(&ma, d->a, cplx, d->a_major, d->m, d->k, d->a_ld, d->a_block_stride, d->a_nblocks, BM, BK, "a");
  if (rc != RNB_OK) return rc;
  rc = make_operand_map(&mb, d->b, cplx, d->b_major, d->n, d->k, d->b_ld, d->b_block_stride, d->b_nblocks, BN, BK, "b");
  if (rc != RNB_OK) return rc;
  GemmParams p;
  p.m = d->m; p.n = d->n; p.k = d->k;
  p.c = d->c; p.c_ld = d->c_ld; p.c_block_stride = d->c_block_stride;
  p.pair_a = d->pair_a; p.pair_b = d->pair_b; p.out_index = d->out_index;
  p.n_out = d->n_out; p.n_pairs = d->n_pairs;
  p.tiles_m = (int)((d->m + BM - 1) / BM);
  p.tiles_n = (int)((d->n + BN - 1) / BN);
  p.k_stages = (int)((d->k + BK - 1) / BK);
  p.accumulate = d->accumulate;
  p.n_peers = 0;
  p.blocks_per_peer = 1;
  p.peer_bulk = 0;
  if (d->c_peers) {
    p.n_peers = d->n_peers;
    p.blocks_per_peer = d->blocks_per_peer;
    for (int i = 0; i < d->n_peers; ++i) {
      RNB_REQUIRE(d->c_peers[i] != nullptr, "c_peers[%d] is NULL", i);
      p.peers[i] = static_cast<double *>(d->c_peers[i]);
    }
    RNB_REQUIRE((int64_t)d->n_out <= (int64_t)d->n_peers * d->blocks_per_peer || d->out_index, "output blocks exceed n_peers * blocks_per_peer");
    // bulk reductions need 16-byte aligned rows in every owner's buffer
    bool ok = ((d->c_ld * (cplx ? 16 : 8)) % 16 == 0) && ((d->c_block_stride * (cplx ? 16 : 8)) % 16 == 0);
    for (int i = 0; i < d->n_peers; ++i) ok = ok && ((reinterpret_cast<uintptr_t>(d->c_peers[i]) & 15) == 0);
    static EnvKnob knob_peer_bulk("RNB_PEER_BULK");
    if (const char *e = knob_peer_bulk.get()) ok = ok && atoi(e) != 0;
    p.peer_bulk = ok ? 1 : 0;
  }
  const int esz = cplx ? 16 : 8;
  p.c_vec_ok = ((reinterpret_cast<uintptr_t>(d->c) & 15) == 0) && ((d->c_ld * esz) % 16 == 0) &&
               ((d->c_block_stride * esz) % 16 == 0);
  const SplitPlan sp = plan_split(d);
  p.splits = sp.splits;
  p.iters_per_split = sp.iters_per_split;
  p.partial = nullptr;
  if (sp.splits > 1) {
    if (!d->workspace || d->workspace_bytes < sp.bytes) {
      set_error("workspace of %zu bytes required (rnb_contract_workspace_bytes), got %zu", sp.bytes, d->workspace_bytes);
      return RNB_ERR_WORKSPACE;
    }
    RNB_REQUIRE((reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0, "workspace must be 16-byte aligned");
    p.partial = static_cast<double *>(d->workspace);
  }
  const int64_t total = (int64_t)p.n_out * p.tiles_m * p.tiles_n * p.splits;
  if (total <= 0) return RNB_OK;
  RNB_REQUIRE(total < (1ll << 31), "too many tiles");
  int grid = sm_count();
  if (total < grid) grid = (int)total;
  if (mode == 2) rc = dispatch_major<2>(d->a_major, d->b_major, ma, mb, p, grid, stream);
  else if (mode == 1) rc = dispatch_major<1>(d->a_major, d->b_major, ma, mb, p, grid, stream);
  else rc = dispatch_major<0>(d->a_major, d->b_major, ma, mb, p, grid, stream);
  if (rc != RNB_OK || p.splits == 1) return rc;
  // second pass: add the splits in a fixed order; complex elements are pairs of doubles
  const int f = cplx ? 2 : 1;
  const int64_t n_tiles = (int64_t)p.n_out * p.tiles_m * p.tiles_n;
  splitk_reduce_kernel<double><<<(unsigned)n_tiles, 256, 0, stream>>>(p.partial, static_cast<double *>(p.c), p.out_index, p.m, p.n * f,
                                                                      p.c_ld * f, p.c_block_stride * f, p.tiles_m, p.tiles_n, p.tiles_n, BM,
                                                                      BN * f, p.splits, p.accumulate, p.c_vec_ok);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

}  // namespace rnb
