// Grouped block GEMM for float32 / complex64: error-compensated 3xTF32 on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM), k-block sum fused in-kernel.
//
// Replaces sum(np.tensordot(ai, bi, axes) ...) of rosnet/array/block/__init__.py:281-283 for the
// single-precision dtypes (BLAS sgemm/cgemm in the reference).
//
// Numerics.  x = hi + lo with hi = tf32(x), lo = tf32(x - hi); a.b ~= lo_a.hi_b + hi_a.lo_b + hi_a.hi_b
// (three tensor-core products, fp32 accumulation) recovers ~fp32 accuracy: measured/emulated
// relative Frobenius error ~1e-7 against a complex128 truth, the reference's own cgemm gives
// 2.6e-7, the parity bar is 1e-5.  A single TF32 product (RNB_PREC_TF32X1) gives 3e-4 and fails it.
//
// Complex.  A complex64 matrix viewed as float is [M][2K] with (re, im) interleaved along k.  The
// split pre-pass writes B as the real [2N][2K] matrix
//     row 2n   = ( b_re, -b_im ) per k      -> column 2n   of C' = Re(C)
//     row 2n+1 = ( b_im,  b_re ) per k      -> column 2n+1 of C' = Im(C)
// so ONE real GEMM C'[M][2N] = A'[M][2K] . B'^T yields the interleaved complex64 result directly:
// the 4M decomposition (8.M.N.K real flops) with no extra passes over C.
//
// Accumulation.  Measured on B200: the tensor core adds into the fp32 TMEM accumulator with
// truncation, a bias that grows LINEARLY with the number of accumulation steps (3xTF32 over
// k = 4096 gave 2.9e-5 relative error, ~1.9e-8 per step, where numpy's sgemm gives 3.8e-7).  The
// accumulation is therefore chunked: the MMA warp accumulates `chunk` k-blocks (default 2 = 64
// floats of k, 24 steps; measured 4.8e-7 at k=4096, numpy sgemm 3.8e-7) into one of two TMEM buffers starting from zero, and the epilogue warps
// drain each finished chunk into fp32 REGISTER accumulators with round-to-nearest adds while the
// MMA warp fills the other buffer.  The long sum over k (and over the contracted block pairs) thus
// happens in registers; TMEM only ever holds a short chain.
//
// Kernel (one CTA per SM, persistent, warp specialised, 384 threads):
//   warp 0      TMA producer: 128B-swizzled [rows][32 floats] boxes of the hi and lo planes of A and B
//   warp 1      MMA issuer: one thread issues 12 tcgen05.mma (M=128, N=256, K=8) per 32-wide k block
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 chunk accumulators)
//   warps 4-11  epilogue (2 warpgroups, 128 columns each): tcgen05.ld -> += registers; at the end
//               of a tile the 128x256 register tile is written row-major to global memory
//   setmaxnreg moves the register budget from the first warpgroup to the two epilogue warpgroups;
//   mbarrier pipelines: smem full/empty (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue).
#include <cuda_bf16.h>
#include <string.h>

#include "common.h"

namespace rnb {
namespace {

constexpr int TBM = 128;
constexpr int TBN = 256;
// Pipeline shape: KB k-values per stage, NS stages.  <32, 2>: 128-byte fp32 rows (SWIZZLE_128B), two 96 KB stages - only ONE
// stage (96 KB) can be in flight while the other is consumed.  <16, 4>: 64-byte rows (SWIZZLE_64B), four 48 KB stages - three
// stages (144 KB) in flight for the same shared memory: what the faster arithmetic needs to keep the tensor pipe fed.
template <int KB_, int NS_>
struct Pipe {
  static constexpr int KB = KB_, NS = NS_;
  static constexpr int A_PLANE = TBM * KB * 4;
  static constexpr int B_PLANE = TBN * KB * 4;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  // products == 2 (TF32 hi.hi + two BF16 cross products): the lo plane's bytes hold two bf16 tiles (bf16(hi), bf16(lo))
  static constexpr int A16_PLANE = TBM * KB * 2;
  static constexpr int B16_PLANE = TBN * KB * 2;
  static constexpr int SMEM = NS * STAGE + 1024 /*align*/ + 256 /*barriers*/;
  // fp32 rows: KB * 4 bytes, bf16 rows: KB * 2 bytes -> UMMA layout type (2 = 128B, 4 = 64B, 6 = 32B) and 8-row atom size
  static constexpr uint64_t LAYOUT32 = KB == 32 ? 2 : 4, SBO32 = KB * 4 * 8;
  static constexpr uint64_t LAYOUT16 = KB == 32 ? 4 : 6, SBO16 = KB * 2 * 8;
};
constexpr int T_THREADS = 384;
constexpr int T_EPI_WARPS = 8;

struct Tf32Params {
  int64_t m, n;  // rows of A', columns of C' (floats)
  int64_t c_ld, c_block_stride;  // floats
  float *c;
  const int32_t *pair_a;
  const int32_t *pair_b;
  const int32_t *out_index;
  int32_t n_out, n_pairs;
  int32_t tiles_m, tiles_n;
  int32_t k_blocks;
  int32_t accumulate;
  int32_t products;  // 3 = 3xTF32, 1 = single TF32
  int32_t interleave_terms;  // 3xTF32 on the CTA-pair kernel: the round-1 MMA order (per k step: two cross terms, then hi.hi); experiments
  int32_t c_vec_ok;
  int32_t chunk;     // k-blocks accumulated in TMEM before promotion to registers
  int32_t panel;     // rasterization: n-tiles per column panel (tiles walk m-fastest inside a panel of this width)
  int32_t splits;    // > 1: split-K, every unit = (tile, split) writes a partial tile to `partial`
  int32_t iters_per_split;
  float *partial;    // [total_tiles * splits][TBM][TBN]
  uint64_t hint_a, hint_b;   // L2 eviction-priority policies of the A / B tile loads of the CTA-pair kernel (0: none)
  int32_t gmul;      // 1, or 3 for complex64 3M: output group g = 3 * (output block) + j computes product j of that block
                     // (j = 0: re.re, 1: im.im, 2: (re+im).(re+im)) from operand blocks 3 * pair_x[block][pr] + j
};

// A unit of work: one output tile, or (split-K) one k-range of it.
struct Unit {
  int g, tm, tn;   // output block, tile coordinates
  int it0, it1;    // range of the flattened (pair, k-block) iteration space
  int slot;        // index of the partial tile in the split-K workspace (tile * splits + split)
};
__device__ __forceinline__ Unit decode_unit(const Tf32Params &p, int unit, int tiles_per_block, int iters_per_tile) {
  Unit u;
  // split-major unit order: the CTAs of one wave work on the same k-range of many tiles, so the A and
  // B panels of that range are shared through L2 (tile-major order re-read B from HBM once per tile row:
  // ncu showed 29 GB of DRAM reads for a 10 GB problem)
  const int n_tiles = p.n_out * tiles_per_block;
  const int s = unit / n_tiles;
  const int tile = unit - s * n_tiles;
  u.slot = tile * p.splits + s;
  u.g = tile / tiles_per_block;
  const int rem = tile - u.g * tiles_per_block;
  // column panels of `panel` n-tiles; inside a panel n runs fastest over the panel width, then m.
  // A wave of CTAs then covers (148 / width) x width tiles instead of 1 x 148, so every B tile is
  // shared by many CTAs through L2 (with n-fastest order a 64 x 256-tile problem streamed all of B
  // from HBM for every m-row: measured 177 vs 234 TFLOP/s).
  const int panel_tiles = p.panel * p.tiles_m;
  const int pn = rem / panel_tiles;
  const int within = rem - pn * panel_tiles;
  const int w = min(p.panel, p.tiles_n - pn * p.panel);
  u.tm = within / w;
  u.tn = pn * p.panel + (within - u.tm * w);
  u.it0 = s * p.iters_per_split;
  u.it1 = min(u.it0 + p.iters_per_split, iters_per_tile);
  return u;
}

// ---- tcgen05 wrappers ----
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t smem_addr, uint64_t layout, uint64_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address / 16
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;         // stride byte offset: next 8-row atom
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= layout << 61;                             // swizzle mode of the rows
  return d;
}

template <int KB, int NS>
__global__ void __launch_bounds__(T_THREADS, 1)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_a16, const __grid_constant__ CUtensorMap map_b16, const Tf32Params p) {
  using P = Pipe<KB, NS>;
  constexpr int TSTAGES = NS, TBK = KB, A_PLANE = P::A_PLANE, B_PLANE = P::B_PLANE, T_STAGE = P::STAGE;
  constexpr int A16_PLANE = P::A16_PLANE, B16_PLANE = P::B16_PLANE;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + TSTAGES * T_STAGE);
  uint64_t *empty_bar = full_bar + TSTAGES;
  uint64_t *tmem_full = empty_bar + TSTAGES;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], T_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_block = p.tiles_m * p.tiles_n;
  const int total_tiles = p.n_out * tiles_per_block * p.splits;  // units
  const int iters_per_tile = p.n_pairs * p.k_blocks;
  const uint32_t stage_bytes = (p.products >= 2) ? T_STAGE : (A_PLANE + B_PLANE);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
      if (p.products == 2) {
        tma_prefetch_desc(&map_a16);
        tma_prefetch_desc(&map_b16);
      }
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const Unit u = decode_unit(p, tile, tiles_per_block, iters_per_tile);
        const int tm = u.tm, tn = u.tn;
        int pr = u.it0 / p.k_blocks;
        int kb = u.it0 - pr * p.k_blocks;
        const int gsrc = u.g / p.gmul, gj = u.g - gsrc * p.gmul;
        int blk_a = p.pair_a[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
        int blk_b = p.pair_b[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
        for (int iter = u.it0; iter < u.it1; ++iter, ++it) {
          const int stage = it % TSTAGES;
          const uint32_t phase = (it / TSTAGES) & 1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          unsigned char *base = smem + stage * T_STAGE;
          tma_load_4d(base, &map_a, &full_bar[stage], kb * TBK, tm * TBM, blk_a, 0);
          tma_load_4d(base + 2 * A_PLANE, &map_b, &full_bar[stage], kb * TBK, tn * TBN, blk_b, 0);
          if (p.products == 3) {
            tma_load_4d(base + A_PLANE, &map_a, &full_bar[stage], kb * TBK, tm * TBM, blk_a, 1);
            tma_load_4d(base + 2 * A_PLANE + B_PLANE, &map_b, &full_bar[stage], kb * TBK, tn * TBN, blk_b, 1);
          } else if (p.products == 2) {
            tma_load_4d(base + A_PLANE, &map_a16, &full_bar[stage], kb * TBK, tm * TBM, blk_a, 0);                           // bf16(hi)
            tma_load_4d(base + A_PLANE + A16_PLANE, &map_a16, &full_bar[stage], kb * TBK, tm * TBM, blk_a, 1);               // bf16(lo)
            tma_load_4d(base + 2 * A_PLANE + B_PLANE, &map_b16, &full_bar[stage], kb * TBK, tn * TBN, blk_b, 0);
            tma_load_4d(base + 2 * A_PLANE + B_PLANE + B16_PLANE, &map_b16, &full_bar[stage], kb * TBK, tn * TBN, blk_b, 1);
          }
          if (++kb == p.k_blocks && iter + 1 < u.it1) {
            kb = 0;
            ++pr;
            blk_a = p.pair_a[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
            blk_b = p.pair_b[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, both K-major, N=256, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
      uint32_t it = 0, cc = 0;  // smem stage counter, chunk counter
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const Unit u = decode_unit(p, tile, tiles_per_block, iters_per_tile);
        for (int iter0 = u.it0; iter0 < u.it1; iter0 += p.chunk, ++cc) {
          const uint32_t buf = cc & 1;
          mbar_wait(&tmem_empty[buf], ((cc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * TBN;
          const int iter1 = min(iter0 + p.chunk, u.it1);
          for (int iter = iter0; iter < iter1; ++iter, ++it) {
            const int stage = it % TSTAGES;
            const uint32_t phase = (it / TSTAGES) & 1;
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + stage * T_STAGE);
            if (p.products == 2) {
              // a.b ~= bf16(lo_a).bf16(hi_b) + bf16(hi_a).bf16(lo_b) + tf32(hi_a).tf32(hi_b): the two cross terms only need
              // ~8 bits (they are 2^-11 of the product), so they run as BF16 MMAs at twice the TF32 rate (small terms first)
              const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
              const uint32_t a_hb = sbase + A_PLANE, a_lb = a_hb + A16_PLANE;
              const uint32_t b_hb = sbase + 2 * A_PLANE + B_PLANE, b_lb = b_hb + B16_PLANE;
#pragma unroll
              for (int kk = 0; kk < TBK / 16; ++kk) {
                umma_bf16(d_tmem, umma_desc_k(a_lb + kk * 32, P::LAYOUT16, P::SBO16), umma_desc_k(b_hb + kk * 32, P::LAYOUT16, P::SBO16), idesc16, (iter > iter0 || kk > 0) ? 1u : 0u);
                umma_bf16(d_tmem, umma_desc_k(a_hb + kk * 32, P::LAYOUT16, P::SBO16), umma_desc_k(b_lb + kk * 32, P::LAYOUT16, P::SBO16), idesc16, 1u);
              }
#pragma unroll
              for (int kk = 0; kk < TBK / 8; ++kk)
                umma_tf32(d_tmem, umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32), umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32), idesc, 1u);
              umma_commit(&empty_bar[stage]);
              continue;
            }
#pragma unroll
            for (int kk = 0; kk < TBK / 8; ++kk) {
              const uint64_t ah = umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32);
              const uint64_t al = umma_desc_k(sbase + A_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
              const uint64_t bh = umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
              const uint64_t bl = umma_desc_k(sbase + 2 * A_PLANE + B_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
              const uint32_t first = (iter > iter0 || kk > 0) ? 1u : 0u;
              if (p.products == 3) {
                umma_tf32(d_tmem, al, bh, idesc, first);  // small terms first
                umma_tf32(d_tmem, ah, bl, idesc, 1u);
                umma_tf32(d_tmem, ah, bh, idesc, 1u);
              } else {
                umma_tf32(d_tmem, ah, bh, idesc, first);
              }
            }
            umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          }
          umma_commit(&tmem_full[buf]);  // chunk accumulator complete
        }
      }
    }
  }
  } else {
    // ===== epilogue: drain chunk accumulators TMEM -> registers (+=), then registers -> global =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // column half: 0 -> [0,128), 1 -> [128,256)
    uint32_t cc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const Unit u = decode_unit(p, tile, tiles_per_block, iters_per_tile);
      const int g = u.g, tm = u.tm, tn = u.tn;
      const int64_t col0 = (int64_t)tn * TBN + half * 128;
      const bool have_cols = col0 < p.n;  // warp-uniform
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int iter0 = u.it0; iter0 < u.it1; iter0 += p.chunk, ++cc) {
        const uint32_t buf = cc & 1;
        mbar_wait(&tmem_full[buf], (cc >> 1) & 1);
        tc_fence_after();
        if (have_cols) {
          const uint32_t taddr = tmem_base + buf * TBN + half * 128 + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t v[16];
            tmem_ld16(taddr + c * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
      // ---- write the register tile ----
      const int64_t row = (int64_t)tm * TBM + q * 32 + lane;
      if (p.splits > 1) {
        // split-K: the partial tile goes to the workspace; splitk_reduce_kernel adds the splits in a fixed order
        if (have_cols && row < p.m) {
          float4 *d4 = reinterpret_cast<float4 *>(p.partial + (int64_t)u.slot * (TBM * TBN) + (q * 32 + lane) * TBN + half * 128);
#pragma unroll
          for (int c = 0; c < 32; ++c) d4[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
        continue;
      }
      const int out_blk = p.out_index ? p.out_index[g] : g;
      if (have_cols && row < p.m) {
        float *dst = p.c + (int64_t)out_blk * p.c_block_stride + row * p.c_ld + col0;
        if (p.c_vec_ok && col0 + 128 <= p.n) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float4 o = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            float4 *d4 = reinterpret_cast<float4 *>(dst) + c;
            if (p.accumulate) {
              const float4 old = *d4;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *d4 = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 128; ++j) {
            if (col0 + j < p.n) {
              float x = acc[j];
              if (p.accumulate) x += dst[j];
              dst[j] = x;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// v2 kernel: CTA pairs (tcgen05 cta_group::2) + coalesced epilogue.
//
// Why pairs.  The 1-CTA kernel above is bound by SHARED-MEMORY bandwidth, not by the tensor pipe: per 32-float k-block it
// writes 96 KB (TMA) and reads 12 x (4 KB of A + 8 KB of B) = 144 KB (UMMA operand fetch) = 240 KB in the 1536 cycles the
// twelve M128 N256 K8 TF32 MMAs take at full rate: 156 B/cycle against the 128 B/cycle an SM's shared memory delivers,
// i.e. at most 82 % tensor-pipe activity (ncu measured 75 %; the TF32 + 2xBF16 mode needs 187 B/cycle -> 68 %, measured
// 61 %).  With cta_group::2 one instruction computes a 256 x 256 tile on two SMs: each CTA stages ITS 128 rows of A and
// only HALF of the B tile (128 of the 256 rows), the hardware feeds both tensor cores from both shared memories:
// 64 KB written + 12 x (4 + 4) KB read = 160 KB per 1536 cycles = 104 B/cycle.  A stage shrinks from 96 to 64 KB, so
// three stages fit where two did.
//
// Protocol (CTA rank 0 of the cluster = leader):
//   - both CTAs run a TMA producer for their own shared memory; every load signals the LEADER's `full` barrier
//     (cp.async.bulk.tensor ... cta_group::2 with the barrier address mapped into CTA 0), which the leader arms with
//     expect_tx(2 x stage bytes);
//   - only the leader's MMA thread issues tcgen05.mma.cta_group::2; tcgen05.commit ... multicast::cluster (mask 0b11)
//     releases the `empty` barrier of the stage and raises `tmem_full` in BOTH CTAs;
//   - every CTA's epilogue warps drain their own 128 TMEM lanes; the leader's `tmem_empty` barrier counts the
//     epilogue warps of both CTAs (the follower's arrive through shared::cluster).
//
// Epilogue.  A lane owns one row of the register tile; storing it directly writes 32 rows x 16 bytes per instruction
// (measured ~2 TB/s on the 1-2 GiB results of low-k tensor-network steps).  Here every warp transposes 32 x 32 chunks
// through a swizzled 4 KB staging tile, so one store instruction writes 4 rows x 128 contiguous bytes.  TMEM is drained
// with two 32-column loads in flight per wait.
template <int CG>
struct Pipe2 {
  static constexpr int KB = 32;
  static constexpr int NS = CG == 1 ? 2 : 3;
  static constexpr int A_PLANE = TBM * KB * 4;        // this CTA's 128 rows of A' (one plane)
  static constexpr int B_ROWS = TBN / CG;             // rows of the B' tile staged by this CTA
  static constexpr int B_PLANE = B_ROWS * KB * 4;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int A16_PLANE = TBM * KB * 2;
  static constexpr int B16_PLANE = B_ROWS * KB * 2;
  static constexpr int STG_WARP = 32 * 32 * 4;        // epilogue staging tile per warp
  static constexpr int SMEM = NS * STAGE + T_EPI_WARPS * STG_WARP + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint64_t LAYOUT32 = 2, SBO32 = KB * 4 * 8;
  static constexpr uint64_t LAYOUT16 = 4, SBO16 = KB * 2 * 8;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");   // default .release.cta: the explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR per arrive
}
// TMA tile load issued by either CTA of a pair: data into OUR shared memory, completion bytes on a barrier that may live in the peer
__device__ __forceinline__ void tma_load_4d_pair(void *smem_dst, const CUtensorMap *tmap, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy encodings: 0x12F0... evict_first, 0x14F0... evict_last)
__device__ __forceinline__ void tma_load_4d_pair_hint(void *smem_dst, const CUtensorMap *tmap, uint32_t bar_cluster_addr, int c0, int c1,
                                                      int c2, int c3, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t *slot_in_smem, uint32_t ncols) {
  if constexpr (CG == 1) {
    tmem_alloc(slot_in_smem, ncols);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1) {
    tmem_dealloc(taddr, ncols);
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  }
}
template <int CG, bool BF16>
__device__ __forceinline__ void umma_cg(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    if constexpr (BF16) umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
    else umma_tf32(d_tmem, adesc, bdesc, idesc, accumulate);
  } else if constexpr (BF16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// MMA completion -> mbarrier arrive; for a pair the arrive is multicast to the same barrier of both CTAs
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t *bar) {
  if constexpr (CG == 1) {
    umma_commit(bar);
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One warp's 32 x 128 register tile (lane = row, acc = 128 consecutive columns) -> dst[row * ld + col], bounded by
// rows_ok x cols_ok, through the warp's swizzled 32 x 32 staging tile: every store instruction writes 4 rows x 128 bytes.
__device__ __forceinline__ void store_warp_tile(const float (&acc)[128], float *stg, int lane, float *dst, int64_t ld, int rows_ok,
                                                int cols_ok, bool vec_ok, bool accumulate) {
  float4 *s4 = reinterpret_cast<float4 *>(stg);
#pragma unroll
  for (int cgp = 0; cgp < 4; ++cgp) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      s4[lane * 8 + (j ^ (lane & 7))] = make_float4(acc[cgp * 32 + 4 * j], acc[cgp * 32 + 4 * j + 1], acc[cgp * 32 + 4 * j + 2], acc[cgp * 32 + 4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + (lane >> 3);
      const int ch = lane & 7;
      float4 v = s4[r * 8 + (ch ^ (r & 7))];
      const int col = cgp * 32 + ch * 4;
      if (r < rows_ok && col < cols_ok) {
        float *d = dst + (int64_t)r * ld + col;
        if (vec_ok && col + 4 <= cols_ok) {
          if (accumulate) {
            const float4 old = *reinterpret_cast<const float4 *>(d);
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          *reinterpret_cast<float4 *>(d) = v;
        } else {
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (col + t < cols_ok) d[t] = accumulate ? d[t] + e[t] : e[t];
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int CG>
__global__ void __launch_bounds__(T_THREADS, 1)
    gemm_tf32_v2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_a16, const __grid_constant__ CUtensorMap map_b16, const Tf32Params p) {
  using P = Pipe2<CG>;
  constexpr int NS = P::NS, TBK = P::KB, A_PLANE = P::A_PLANE, B_PLANE = P::B_PLANE, STAGE = P::STAGE;
  constexpr int A16_PLANE = P::A16_PLANE, B16_PLANE = P::B16_PLANE;
  constexpr int TILE_M = TBM * CG;   // rows of the tile a CTA group computes
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float *stg_all = reinterpret_cast<float *>(smem + NS * STAGE);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + NS * STAGE + T_EPI_WARPS * P::STG_WARP);
  uint64_t *empty_bar = full_bar + NS;
  uint64_t *tmem_full = empty_bar + NS;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int group = (CG == 2) ? (int)cluster_id_x() : (int)blockIdx.x;
  const int n_groups = (CG == 2) ? (int)cluster_count_x() : (int)gridDim.x;
  const bool leader = cta_rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], T_EPI_WARPS * CG);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_cg<CG>(tmem_slot, 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_block = p.tiles_m * p.tiles_n;
  const int total_units = p.n_out * tiles_per_block * p.splits;
  const int iters_per_tile = p.n_pairs * p.k_blocks;
  const uint32_t stage_bytes = (p.products >= 2) ? STAGE : (A_PLANE + B_PLANE);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===== TMA producer (both CTAs of a pair; completion bytes land on the leader's barrier) =====
      if (lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (p.products == 2) {
          tma_prefetch_desc(&map_a16);
          tma_prefetch_desc(&map_b16);
        }
        uint32_t it = 0;
        for (int unit = group; unit < total_units; unit += n_groups) {
          const Unit u = decode_unit(p, unit, tiles_per_block, iters_per_tile);
          const int a_row = u.tm * TILE_M + (int)cta_rank * TBM;
          const int b_row = u.tn * TBN + (int)cta_rank * P::B_ROWS;
          int pr = u.it0 / p.k_blocks;
          int kb = u.it0 - pr * p.k_blocks;
          const int gsrc = u.g / p.gmul, gj = u.g - gsrc * p.gmul;
          int blk_a = p.pair_a[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
          int blk_b = p.pair_b[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
          for (int iter = u.it0; iter < u.it1; ++iter, ++it) {
            const int stage = it % NS;
            const uint32_t phase = (it / NS) & 1;
            mbar_wait_guard(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], stage_bytes * CG);
            unsigned char *base = smem + stage * STAGE;
            const int kc = kb * TBK;
            if constexpr (CG == 1) {
              uint64_t *fb = &full_bar[stage];
              tma_load_4d(base, &map_a, fb, kc, a_row, blk_a, 0);
              tma_load_4d(base + 2 * A_PLANE, &map_b, fb, kc, b_row, blk_b, 0);
              if (p.products == 3) {
                tma_load_4d(base + A_PLANE, &map_a, fb, kc, a_row, blk_a, 1);
                tma_load_4d(base + 2 * A_PLANE + B_PLANE, &map_b, fb, kc, b_row, blk_b, 1);
              } else if (p.products == 2) {
                tma_load_4d(base + A_PLANE, &map_a16, fb, kc, a_row, blk_a, 0);
                tma_load_4d(base + A_PLANE + A16_PLANE, &map_a16, fb, kc, a_row, blk_a, 1);
                tma_load_4d(base + 2 * A_PLANE + B_PLANE, &map_b16, fb, kc, b_row, blk_b, 0);
                tma_load_4d(base + 2 * A_PLANE + B_PLANE + B16_PLANE, &map_b16, fb, kc, b_row, blk_b, 1);
              }
            } else {
              const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
              if (p.products == 3 && p.hint_a) {
                // B tiles of a column panel are re-used by every wave that sweeps the panel, A tiles only inside their wave:
                // keep B in L2 (evict_last), let A go first
                tma_load_4d_pair_hint(base, &map_a, fb, kc, a_row, blk_a, 0, p.hint_a);
                tma_load_4d_pair_hint(base + 2 * A_PLANE, &map_b, fb, kc, b_row, blk_b, 0, p.hint_b);
                tma_load_4d_pair_hint(base + A_PLANE, &map_a, fb, kc, a_row, blk_a, 1, p.hint_a);
                tma_load_4d_pair_hint(base + 2 * A_PLANE + B_PLANE, &map_b, fb, kc, b_row, blk_b, 1, p.hint_b);
                goto loaded;
              }
              tma_load_4d_pair(base, &map_a, fb, kc, a_row, blk_a, 0);
              tma_load_4d_pair(base + 2 * A_PLANE, &map_b, fb, kc, b_row, blk_b, 0);
              if (p.products == 3) {
                tma_load_4d_pair(base + A_PLANE, &map_a, fb, kc, a_row, blk_a, 1);
                tma_load_4d_pair(base + 2 * A_PLANE + B_PLANE, &map_b, fb, kc, b_row, blk_b, 1);
              } else if (p.products == 2) {
                tma_load_4d_pair(base + A_PLANE, &map_a16, fb, kc, a_row, blk_a, 0);
                tma_load_4d_pair(base + A_PLANE + A16_PLANE, &map_a16, fb, kc, a_row, blk_a, 1);
                tma_load_4d_pair(base + 2 * A_PLANE + B_PLANE, &map_b16, fb, kc, b_row, blk_b, 0);
                tma_load_4d_pair(base + 2 * A_PLANE + B_PLANE + B16_PLANE, &map_b16, fb, kc, b_row, blk_b, 1);
              }
            }
          loaded:
            if (++kb == p.k_blocks && iter + 1 < u.it1) {
              kb = 0;
              ++pr;
              blk_a = p.pair_a[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
              blk_b = p.pair_b[(int64_t)gsrc * p.n_pairs + pr] * p.gmul + gj;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer (one thread of the leader CTA) =====
      if (lane == 0 && leader) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        uint32_t it = 0, cc = 0;
        for (int unit = group; unit < total_units; unit += n_groups) {
          const Unit u = decode_unit(p, unit, tiles_per_block, iters_per_tile);
          for (int iter0 = u.it0; iter0 < u.it1; iter0 += p.chunk, ++cc) {
            const uint32_t buf = cc & 1;
            mbar_wait_guard(&tmem_empty[buf], ((cc >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * TBN;
            const int iter1 = min(iter0 + p.chunk, u.it1);
            for (int iter = iter0; iter < iter1; ++iter, ++it) {
              const int stage = it % NS;
              const uint32_t phase = (it / NS) & 1;
              mbar_wait_guard(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sbase = smem_u32(smem + stage * STAGE);
              if (p.products == 2) {
                const uint32_t a_hb = sbase + A_PLANE, a_lb = a_hb + A16_PLANE;
                const uint32_t b_hb = sbase + 2 * A_PLANE + B_PLANE, b_lb = b_hb + B16_PLANE;
#pragma unroll
                for (int kk = 0; kk < TBK / 16; ++kk) {
                  umma_cg<CG, true>(d_tmem, umma_desc_k(a_lb + kk * 32, P::LAYOUT16, P::SBO16), umma_desc_k(b_hb + kk * 32, P::LAYOUT16, P::SBO16),
                                    idesc16, (iter > iter0 || kk > 0) ? 1u : 0u);
                  umma_cg<CG, true>(d_tmem, umma_desc_k(a_hb + kk * 32, P::LAYOUT16, P::SBO16), umma_desc_k(b_lb + kk * 32, P::LAYOUT16, P::SBO16),
                                    idesc16, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < TBK / 8; ++kk)
                  umma_cg<CG, false>(d_tmem, umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32),
                                     umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32), idesc, 1u);
              } else if (p.products == 3 && p.interleave_terms) {
#pragma unroll
                for (int kk = 0; kk < TBK / 8; ++kk) {
                  const uint64_t ah = umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32);
                  const uint64_t al = umma_desc_k(sbase + A_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
                  const uint64_t bh = umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
                  const uint64_t bl = umma_desc_k(sbase + 2 * A_PLANE + B_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
                  umma_cg<CG, false>(d_tmem, al, bh, idesc, (iter > iter0 || kk > 0) ? 1u : 0u);
                  umma_cg<CG, false>(d_tmem, ah, bl, idesc, 1u);
                  umma_cg<CG, false>(d_tmem, ah, bh, idesc, 1u);
                }
              } else if (p.products == 3) {
                // The tensor core adds into the fp32 TMEM accumulator with TRUNCATION: every MMA of a chain costs about half an ulp of
                // the accumulator, biased towards zero - whatever the size of what it adds.  So the 8 cross-term MMAs of this k-block
                // go first and the 4 hi.hi MMAs last: in the first k-block of a chunk the cross terms land on an accumulator 2^-11
                // the size (their truncation is negligible), 16 instead of 22 full-size steps per 2-block chunk.  Measured on a
                // Sycamore amplitude against complex128 (profiles/r2/searched_tree_accuracy_chunk.txt): the error is linear in the
                // number of full-size steps, 1.6e-6 + 1.6e-7 per step.
#pragma unroll
                for (int kk = 0; kk < TBK / 8; ++kk) {
                  const uint64_t ah = umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32);
                  const uint64_t al = umma_desc_k(sbase + A_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
                  const uint64_t bh = umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
                  const uint64_t bl = umma_desc_k(sbase + 2 * A_PLANE + B_PLANE + kk * 32, P::LAYOUT32, P::SBO32);
                  umma_cg<CG, false>(d_tmem, al, bh, idesc, (iter > iter0 || kk > 0) ? 1u : 0u);
                  umma_cg<CG, false>(d_tmem, ah, bl, idesc, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < TBK / 8; ++kk)
                  umma_cg<CG, false>(d_tmem, umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32),
                                     umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32), idesc, 1u);
              } else {
#pragma unroll
                for (int kk = 0; kk < TBK / 8; ++kk)
                  umma_cg<CG, false>(d_tmem, umma_desc_k(sbase + kk * 32, P::LAYOUT32, P::SBO32),
                                     umma_desc_k(sbase + 2 * A_PLANE + kk * 32, P::LAYOUT32, P::SBO32), idesc, (iter > iter0 || kk > 0) ? 1u : 0u);
              }
              umma_commit_cg<CG>(&empty_bar[stage]);  // shared-memory slot free (in both CTAs) once these MMAs retire
            }
            umma_commit_cg<CG>(&tmem_full[buf]);  // chunk accumulator complete (in both CTAs)
          }
        }
      }
    }
  } else {
    // ===== epilogue: drain chunk accumulators TMEM -> registers (+=), then registers -> global (coalesced) =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // column half: 0 -> [0,128), 1 -> [128,256)
    float *stg = stg_all + (warp - 4) * (P::STG_WARP / 4);
    uint32_t empty_addr[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) empty_addr[b] = (CG == 2) ? mapa_shared(smem_u32(&tmem_empty[b]), 0) : smem_u32(&tmem_empty[b]);
    uint32_t cc = 0;
    for (int unit = group; unit < total_units; unit += n_groups) {
      const Unit u = decode_unit(p, unit, tiles_per_block, iters_per_tile);
      const int g = u.g, tn = u.tn;
      const int tm128 = u.tm * CG + (int)cta_rank;   // this CTA's 128-row tile
      const int64_t col0 = (int64_t)tn * TBN + half * 128;
      const int64_t row0 = (int64_t)tm128 * TBM + q * 32;
      const bool have = col0 < p.n && row0 < p.m;  // warp-uniform
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int iter0 = u.it0; iter0 < u.it1; iter0 += p.chunk, ++cc) {
        const uint32_t buf = cc & 1;
        mbar_wait_guard(&tmem_full[buf], (cc >> 1) & 1);
        tc_fence_after();
        if (have) {
          const uint32_t taddr = tmem_base + buf * TBN + half * 128 + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v0[32], v1[32];
            tmem_ld32_nowait(taddr + c * 64, v0);
            tmem_ld32_nowait(taddr + c * 64 + 32, v1);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 64 + j] += __uint_as_float(v0[j]);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 64 + 32 + j] += __uint_as_float(v1[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(empty_addr[buf]);
          else mbar_arrive(&tmem_empty[buf]);
        }
      }
      if (!have) continue;
      if (p.splits > 1) {
        // split-K: the partial 128 x 256 tile goes to the workspace slot the second pass expects (its own tile enumeration
        // over 128-row tiles); splitk_reduce_kernel adds the splits in a fixed order
        const int tiles_m128 = (int)((p.m + TBM - 1) / TBM);
        const int pn = tn / p.panel;
        const int w = min(p.panel, p.tiles_n - pn * p.panel);
        const int64_t tile128 = (int64_t)g * tiles_m128 * p.tiles_n + (int64_t)pn * p.panel * tiles_m128 + (int64_t)tm128 * w + (tn - pn * p.panel);
        const int s = unit / (p.n_out * tiles_per_block);
        float *dst = p.partial + (tile128 * p.splits + s) * (int64_t)(TBM * TBN) + (int64_t)(q * 32) * TBN + half * 128;
        store_warp_tile(acc, stg, lane, dst, TBN, 32, 128, true, false);
        continue;
      }
      const int out_blk = p.out_index ? p.out_index[g] : g;
      float *dst = p.c + (int64_t)out_blk * p.c_block_stride + row0 * p.c_ld + col0;
      const int rows_ok = (int)min((int64_t)32, p.m - row0);
      const int cols_ok = (int)min((int64_t)128, p.n - col0);
      store_warp_tile(acc, stg, lane, dst, p.c_ld, rows_ok, cols_ok, p.c_vec_ok != 0, p.accumulate != 0);
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

// ---- split pre-pass: x -> (tf32(x), tf32(x - tf32(x))) -------------------------------------------

// rows of floats (A operand of either dtype, B operand of float32): elementwise
__global__ void __launch_bounds__(256) split_rows_kernel(const float *__restrict__ in, float *__restrict__ hi, float *__restrict__ lo,
                                                         int64_t total, int rows, int kf, int64_t ld_in, int64_t blk_stride_in,
                                                         int64_t ld_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kf;
    const int c = (int)(i - r * kf);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float x = in[blk * blk_stride_in + (int64_t)rr * ld_in + c];
    const float h = to_tf32(x);
    hi[r * ld_out + c] = h;
    lo[r * ld_out + c] = to_tf32(x - h);
  }
}

// complex64 B operand [N][K] -> real [2N][2K] with the sign/swap arrangement of the file header
__global__ void __launch_bounds__(256) split_b_complex_kernel(const float2 *__restrict__ in, float *__restrict__ hi,
                                                              float *__restrict__ lo, int64_t total, int rows, int k,
                                                              int64_t ld_in, int64_t blk_stride_in, int64_t ld_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;  // global complex row (block * rows + n)
    const int c = (int)(i - r * k);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float2 z = in[blk * blk_stride_in + (int64_t)rr * ld_in + c];
    const float rh = to_tf32(z.x), ih = to_tf32(z.y);
    const float rl = to_tf32(z.x - rh), il = to_tf32(z.y - ih);
    float *h0 = hi + (2 * r) * ld_out + 2 * c;
    float *h1 = hi + (2 * r + 1) * ld_out + 2 * c;
    float *l0 = lo + (2 * r) * ld_out + 2 * c;
    float *l1 = lo + (2 * r + 1) * ld_out + 2 * c;
    *reinterpret_cast<float2 *>(h0) = make_float2(rh, -ih);
    *reinterpret_cast<float2 *>(h1) = make_float2(ih, rh);
    *reinterpret_cast<float2 *>(l0) = make_float2(rl, -il);
    *reinterpret_cast<float2 *>(l1) = make_float2(il, rl);
  }
}


// 128-bit variants (row length, pitches and bases multiples of 16 bytes): 4 floats / 2 complex per thread
__global__ void __launch_bounds__(256) split_rows_vec4_kernel(const float4 *__restrict__ in, float4 *__restrict__ hi, float4 *__restrict__ lo,
                                                              int64_t total_vec, int rows, int kv, int64_t ldv_in, int64_t blk_stride_v,
                                                              int64_t ldv_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kv;
    const int c = (int)(i - r * kv);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float4 x = __ldg(in + blk * blk_stride_v + (int64_t)rr * ldv_in + c);
    float4 h, l;
    h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
    l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
    hi[r * ldv_out + c] = h;
    lo[r * ldv_out + c] = l;
  }
}

__global__ void __launch_bounds__(256) split_b_complex_vec_kernel(const float4 *__restrict__ in, float4 *__restrict__ hi,
                                                                  float4 *__restrict__ lo, int64_t total_vec, int rows, int kv,
                                                                  int64_t ldv_in, int64_t blk_stride_v, int64_t ldv_out) {
  // one thread: two consecutive complex k of one row n  ->  4 floats in each of rows 2n and 2n+1
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kv;
    const int c = (int)(i - r * kv);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float4 z = __ldg(in + blk * blk_stride_v + (int64_t)rr * ldv_in + c);  // (re0, im0, re1, im1)
    const float r0h = to_tf32(z.x), i0h = to_tf32(z.y), r1h = to_tf32(z.z), i1h = to_tf32(z.w);
    const float r0l = to_tf32(z.x - r0h), i0l = to_tf32(z.y - i0h), r1l = to_tf32(z.z - r1h), i1l = to_tf32(z.w - i1h);
    const int64_t o0 = (2 * r) * ldv_out + c, o1 = (2 * r + 1) * ldv_out + c;
    hi[o0] = make_float4(r0h, -i0h, r1h, -i1h);
    hi[o1] = make_float4(i0h, r0h, i1h, r1h);
    lo[o0] = make_float4(r0l, -i0l, r1l, -i1l);
    lo[o1] = make_float4(i0l, r0l, i1l, r1l);
  }
}

// ---- split pre-pass of the TF32 + 2 x BF16 mode: x -> tf32(x) [fp32], bf16(tf32(x)), bf16(x - tf32(x)) ----
__global__ void __launch_bounds__(256) split16_rows_kernel(const float *__restrict__ in, float *__restrict__ hi, __nv_bfloat16 *__restrict__ hb,
                                                           __nv_bfloat16 *__restrict__ lb, int64_t total, int rows, int kf, int64_t ld_in,
                                                           int64_t blk_stride_in, int64_t ld_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kf;
    const int c = (int)(i - r * kf);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float x = in[blk * blk_stride_in + (int64_t)rr * ld_in + c];
    const float h = to_tf32(x);
    hi[r * ld_out + c] = h;
    hb[r * ld_out + c] = __float2bfloat16_rn(h);
    lb[r * ld_out + c] = __float2bfloat16_rn(x - h);
  }
}

__global__ void __launch_bounds__(256) split16_b_complex_kernel(const float2 *__restrict__ in, float *__restrict__ hi, __nv_bfloat16 *__restrict__ hb,
                                                                __nv_bfloat16 *__restrict__ lb, int64_t total, int rows, int k, int64_t ld_in,
                                                                int64_t blk_stride_in, int64_t ld_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;  // global complex row (block * rows + n)
    const int c = (int)(i - r * k);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float2 z = in[blk * blk_stride_in + (int64_t)rr * ld_in + c];
    const float rh = to_tf32(z.x), ih = to_tf32(z.y);
    const float rl = z.x - rh, il = z.y - ih;
    const int64_t o0 = (2 * r) * ld_out + 2 * c, o1 = (2 * r + 1) * ld_out + 2 * c;
    *reinterpret_cast<float2 *>(hi + o0) = make_float2(rh, -ih);
    *reinterpret_cast<float2 *>(hi + o1) = make_float2(ih, rh);
    *reinterpret_cast<__nv_bfloat162 *>(hb + o0) = __floats2bfloat162_rn(rh, -ih);
    *reinterpret_cast<__nv_bfloat162 *>(hb + o1) = __floats2bfloat162_rn(ih, rh);
    *reinterpret_cast<__nv_bfloat162 *>(lb + o0) = __floats2bfloat162_rn(rl, -il);
    *reinterpret_cast<__nv_bfloat162 *>(lb + o1) = __floats2bfloat162_rn(il, rl);
  }
}

// 128-bit variants of the two kernels above (same alignment conditions as split_rows_vec4_kernel)
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<const uint32_t *>(&lo);
  r.y = *reinterpret_cast<const uint32_t *>(&hi);
  return r;
}
__global__ void __launch_bounds__(256) split16_rows_vec4_kernel(const float4 *__restrict__ in, float4 *__restrict__ hi, uint2 *__restrict__ hb,
                                                                uint2 *__restrict__ lb, int64_t total_vec, int rows, int kv, int64_t ldv_in,
                                                                int64_t blk_stride_v, int64_t ldv_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kv;
    const int c = (int)(i - r * kv);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float4 x = __ldg(in + blk * blk_stride_v + (int64_t)rr * ldv_in + c);
    float4 h;
    h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
    hi[r * ldv_out + c] = h;
    hb[r * ldv_out + c] = pack_bf16x4(h.x, h.y, h.z, h.w);
    lb[r * ldv_out + c] = pack_bf16x4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
  }
}
__global__ void __launch_bounds__(256) split16_b_complex_vec_kernel(const float4 *__restrict__ in, float4 *__restrict__ hi, uint2 *__restrict__ hb,
                                                                    uint2 *__restrict__ lb, int64_t total_vec, int rows, int kv, int64_t ldv_in,
                                                                    int64_t blk_stride_v, int64_t ldv_out) {
  // one thread: two consecutive complex k of one row n  ->  4 values in each of rows 2n and 2n+1 of every plane
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kv;
    const int c = (int)(i - r * kv);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float4 z = __ldg(in + blk * blk_stride_v + (int64_t)rr * ldv_in + c);  // (re0, im0, re1, im1)
    const float r0h = to_tf32(z.x), i0h = to_tf32(z.y), r1h = to_tf32(z.z), i1h = to_tf32(z.w);
    const float r0l = z.x - r0h, i0l = z.y - i0h, r1l = z.z - r1h, i1l = z.w - i1h;
    const int64_t o0 = (2 * r) * ldv_out + c, o1 = (2 * r + 1) * ldv_out + c;
    hi[o0] = make_float4(r0h, -i0h, r1h, -i1h);
    hi[o1] = make_float4(i0h, r0h, i1h, r1h);
    hb[o0] = pack_bf16x4(r0h, -i0h, r1h, -i1h);
    hb[o1] = pack_bf16x4(i0h, r0h, i1h, r1h);
    lb[o0] = pack_bf16x4(r0l, -i0l, r1l, -i1l);
    lb[o1] = pack_bf16x4(i0l, r0l, i1l, r1l);
  }
}

// ---- complex64 3M: three REAL grouped GEMMs per complex product ---------------------------------------------------
// (a_r + i a_i)(b_r + i b_i):  T1 = a_r.b_r,  T2 = a_i.b_i,  T3 = (a_r + a_i).(b_r + b_i);  Re = T1 - T2,  Im = T3 - T1 - T2
// - 3 real multiply-adds per complex one where the interleaved-k arrangement of the file header (4M) spends 4, i.e. 9 instead
// of 12 TF32 tensor products.  The split pre-pass writes every operand block as THREE real K-major blocks (re, im, re + im;
// block 3 b + j, hi and lo planes), the real grouped kernel above runs 3 x n_out output groups (Tf32Params::gmul = 3: group
// 3 g + j multiplies the j-th real blocks of output block g's pair list, k-block sum fused as always) into a float workspace,
// and one combine pass writes the interleaved complex64 result.  The workspace costs 32 extra bytes of HBM traffic per output
// element, which the saved quarter of the tensor work repays from k ~ 700 on (make_layout: use_3m).
__global__ void __launch_bounds__(256) split3m_vec_kernel(const float4 *__restrict__ in, float4 *__restrict__ hi, float4 *__restrict__ lo,
                                                          int64_t total_quads, int rows, int kq, int64_t ldv_in, int64_t blk_stride_v,
                                                          int64_t ldv_out) {
  // one thread: four consecutive complex k of one row (two 16-byte loads) -> one float4 in each of the six planes
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kq;
    const int c = (int)(i - r * kq);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float4 *src = in + blk * blk_stride_v + (int64_t)rr * ldv_in + 2 * c;
    const float4 z0 = __ldg(src), z1 = __ldg(src + 1);   // (re0, im0, re1, im1), (re2, im2, re3, im3)
    const float re[4] = {z0.x, z0.z, z1.x, z1.z}, im[4] = {z0.y, z0.w, z1.y, z1.w};
    float h[3][4], l[3][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float v[3] = {re[t], im[t], re[t] + im[t]};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        h[j][t] = to_tf32(v[j]);
        l[j][t] = to_tf32(v[j] - h[j][t]);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int64_t o = ((3 * blk + j) * rows + rr) * ldv_out + c;
      hi[o] = make_float4(h[j][0], h[j][1], h[j][2], h[j][3]);
      lo[o] = make_float4(l[j][0], l[j][1], l[j][2], l[j][3]);
    }
  }
}

__global__ void __launch_bounds__(256) split3m_kernel(const float2 *__restrict__ in, float *__restrict__ hi, float *__restrict__ lo,
                                                      int64_t total, int rows, int k, int64_t ld_in, int64_t blk_stride_in, int64_t ld_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;
    const int c = (int)(i - r * k);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    const float2 z = in[blk * blk_stride_in + (int64_t)rr * ld_in + c];
    const float v[3] = {z.x, z.y, z.x + z.y};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int64_t o = ((3 * blk + j) * rows + rr) * ld_out + c;
      const float h = to_tf32(v[j]);
      hi[o] = h;
      lo[o] = to_tf32(v[j] - h);
    }
  }
}

// 3M planes of the TF32 + 2 x BF16 arithmetic: tf32(x) [fp32], bf16(tf32(x)), bf16(x - tf32(x)) for x = re, im, re + im
__global__ void __launch_bounds__(256) split16_3m_kernel(const float2 *__restrict__ in, float *__restrict__ hi, __nv_bfloat16 *__restrict__ hb,
                                                         __nv_bfloat16 *__restrict__ lb, int64_t total_pairs, int rows, int kh, int64_t ld_in,
                                                         int64_t blk_stride_in, int64_t ld_out, int two) {
  // one thread: two consecutive complex k of one row (two != 0: k even, 16-byte aligned operand, kh = k / 2) or a single one (kh = k)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_pairs; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kh;
    const int c = (int)(i - r * kh);
    const int64_t blk = r / rows;
    const int rr = (int)(r - blk * rows);
    if (two) {
      const float4 z = __ldg(reinterpret_cast<const float4 *>(in + blk * blk_stride_in + (int64_t)rr * ld_in + 2 * c));
      const float v[3][2] = {{z.x, z.z}, {z.y, z.w}, {z.x + z.y, z.z + z.w}};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int64_t o = ((3 * blk + j) * rows + rr) * ld_out + 2 * c;
        const float h0 = to_tf32(v[j][0]), h1 = to_tf32(v[j][1]);
        *reinterpret_cast<float2 *>(hi + o) = make_float2(h0, h1);
        *reinterpret_cast<__nv_bfloat162 *>(hb + o) = __floats2bfloat162_rn(h0, h1);
        *reinterpret_cast<__nv_bfloat162 *>(lb + o) = __floats2bfloat162_rn(v[j][0] - h0, v[j][1] - h1);
      }
    } else {
      const float2 z = in[blk * blk_stride_in + (int64_t)rr * ld_in + c];
      const float v[3] = {z.x, z.y, z.x + z.y};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int64_t o = ((3 * blk + j) * rows + rr) * ld_out + c;
        const float h = to_tf32(v[j]);
        hi[o] = h;
        hb[o] = __float2bfloat16_rn(h);
        lb[o] = __float2bfloat16_rn(v[j] - h);
      }
    }
  }
}

// T [3 n_out][m][t_ld] floats -> C complex64 (c_ld, c_block_stride in complex elements)
__global__ void __launch_bounds__(256) combine3m_kernel(const float *__restrict__ t, float2 *__restrict__ c, const int32_t *__restrict__ out_index,
                                                        int n_out, int64_t m, int64_t n, int64_t t_ld, int64_t c_ld, int64_t c_block_stride,
                                                        int accumulate, int vec) {
  const int64_t blk_t = m * t_ld;
  if (vec) {   // four consecutive columns per thread: three 16-byte loads, two 16-byte stores
    const int64_t nq = n >> 2, total = (int64_t)n_out * m * nq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t gr = i / nq;
      const int64_t col = (i - gr * nq) << 2;
      const int64_t g = gr / m, row = gr - g * m;
      const float *t0 = t + (3 * g) * blk_t + row * t_ld + col;
      const float4 t1 = *reinterpret_cast<const float4 *>(t0), t2 = *reinterpret_cast<const float4 *>(t0 + blk_t),
                   t3 = *reinterpret_cast<const float4 *>(t0 + 2 * blk_t);
      const int64_t ob = out_index ? out_index[g] : g;
      float4 *dst = reinterpret_cast<float4 *>(c + ob * c_block_stride + row * c_ld + col);
      float4 o0 = make_float4(t1.x - t2.x, t3.x - t1.x - t2.x, t1.y - t2.y, t3.y - t1.y - t2.y);
      float4 o1 = make_float4(t1.z - t2.z, t3.z - t1.z - t2.z, t1.w - t2.w, t3.w - t1.w - t2.w);
      if (accumulate) {
        const float4 p0 = dst[0], p1 = dst[1];
        o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
        o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
      }
      dst[0] = o0;
      dst[1] = o1;
    }
    return;
  }
  const int64_t total = (int64_t)n_out * m * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gr = i / n;
    const int64_t col = i - gr * n;
    const int64_t g = gr / m, row = gr - g * m;
    const float *t0 = t + (3 * g) * blk_t + row * t_ld + col;
    const float t1 = t0[0], t2 = t0[blk_t], t3 = t0[2 * blk_t];
    const int64_t ob = out_index ? out_index[g] : g;
    float2 *dst = c + ob * c_block_stride + row * c_ld + col;
    float2 o = make_float2(t1 - t2, t3 - t1 - t2);
    if (accumulate) {
      o.x += dst->x;
      o.y += dst->y;
    }
    *dst = o;
  }
}

struct Layout {
  int f;             // floats per element
  int64_t kf;        // K' (floats)
  int64_t ldk;       // plane row pitch, floats (multiple of 4)
  int64_t nf;        // N' (floats / real rows of B')
  int64_t a_plane, b_plane;  // floats per plane
  size_t off_ah, off_al, off_bh, off_bl, off_partial, total;
  int tiles_m, tiles_n, k_blocks, chunk;
  int splits, iters_per_split;
  int kb;            // k-values per pipeline stage: 32 (two 96 KB stages) or 16 (four 48 KB stages)
  int v2;            // 1: gemm_tf32_v2_kernel (coalesced epilogue, CTA pairs), 0: the first-generation kernel
  int cg;            // CTAs per tile (tcgen05 cta_group): 2 = 256-row tiles on CTA pairs
  int tiles_m128;    // 128-row tiles per output block (the split-K second pass enumerates these)
  int gmul;          // 3: complex64 as three real GEMMs (3M), 1 otherwise
  int64_t t_ld;      // 3M: row pitch (floats) of the product workspace T [3 n_out][m][t_ld]
  size_t off_t;
};

// complex64 3M pays 32 extra bytes of HBM traffic per output element (the product workspace is written, read back and
// combined) to save a quarter of the tensor work, 2 k real flops per output element: worth it from k ~ 700.
bool bf16x2_requested(const rnb_contract_desc *d) { return d->precision == RNB_PREC_TF32_BF16X2; }

bool use_3m(const rnb_contract_desc *d) {
  if (d->dtype != RNB_C64) return false;
  if (d->complex_mode == RNB_CPLX_3M) return true;
  if (d->complex_mode != RNB_CPLX_AUTO) return false;
  return (int64_t)d->n_pairs * d->k >= 1024 && d->m >= 128 && d->n >= 128;
}

Layout make_layout(const rnb_contract_desc *d) {
  Layout L;
  L.gmul = use_3m(d) ? 3 : 1;
  L.f = (d->dtype == RNB_C64 && L.gmul == 1) ? 2 : 1;
  L.kf = d->k * L.f;
  L.ldk = (L.kf + 7) & ~(int64_t)7;   // rows of the fp32 planes and of the bf16 planes start 16-byte aligned
  L.nf = d->n * L.f;
  L.a_plane = (int64_t)d->a_nblocks * L.gmul * d->m * L.ldk;
  L.b_plane = (int64_t)d->b_nblocks * L.gmul * L.nf * L.ldk;
  L.t_ld = (d->n + 3) & ~(int64_t)3;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  L.off_ah = off; off = align(off + (size_t)L.a_plane * 4);
  L.off_al = off; off = align(off + (size_t)L.a_plane * 4);
  L.off_bh = off; off = align(off + (size_t)L.b_plane * 4);
  L.off_bl = off; off = align(off + (size_t)L.b_plane * 4);
  L.tiles_m128 = (int)((d->m + TBM - 1) / TBM);
  L.tiles_n = (int)((L.nf + TBN - 1) / TBN);
  L.kb = 32;
  static EnvKnob knob_kb("RNB_TF32_KB"), knob_kernel("RNB_TF32_KERNEL"), knob_cg("RNB_TF32_CG");
  if (const char *e = knob_kb.get()) {
    if (atoi(e) == 16) L.kb = 16;
  }
  L.v2 = 1;
  if (const char *e = knob_kernel.get()) {
    if (!strcmp(e, "v1")) L.v2 = 0;
  }
  if (L.kb == 16) L.v2 = 0;   // the 16-k pipeline experiment only exists in the first-generation kernel
  // CTA pairs (256-row tiles) unless the block has a single 128-row tile or pairing would leave > ~12 % of the rows empty
  L.cg = 1;
  if (L.v2 && d->m > TBM) {
    const int64_t padded = (d->m + 2 * TBM - 1) / (2 * TBM) * (2 * TBM);
    if (padded == (int64_t)L.tiles_m128 * TBM || d->m >= 8 * TBM) L.cg = 2;
  }
  if (const char *e = knob_cg.get()) {
    const int v = atoi(e);
    if (L.v2 && (v == 1 || (v == 2 && d->m > TBM))) L.cg = v;
  }
  L.tiles_m = (int)((d->m + TBM * L.cg - 1) / (TBM * L.cg));
  L.k_blocks = (int)((L.kf + L.kb - 1) / L.kb);
  L.chunk = 64 / L.kb;   // 64 k-values (24 accumulation steps of 3xTF32) per TMEM chain
  static EnvKnob knob_chunk("RNB_TF32_CHUNK");
  if (const char *e = knob_chunk.get()) {
    const int v = atoi(e);
    if (v >= 1) L.chunk = v;
  }
  // split-K when the launch has too few tiles to occupy the SMs (>= 16 k-blocks = 20 us of MMA per split)
  const int64_t tiles = (int64_t)d->n_out * L.gmul * L.tiles_m * L.tiles_n;
  const int64_t iters = (int64_t)d->n_pairs * L.k_blocks;
  L.splits = choose_splits(tiles, iters, 16 * (32 / L.kb), 96 * (32 / L.kb), sm_count() / L.cg);
  L.iters_per_split = (int)iters;
  if (L.splits > 1) {
    int64_t ips = (iters + L.splits - 1) / L.splits;
    ips = (ips + L.chunk - 1) / L.chunk * L.chunk;   // chunk boundaries stay aligned to the tile's iteration 0
    L.iters_per_split = (int)ips;
    L.splits = (int)((iters + ips - 1) / ips);       // every split non-empty
  }
  L.off_partial = off;
  if (L.splits > 1) off = align(off + (size_t)d->n_out * L.gmul * L.tiles_m128 * L.tiles_n * L.splits * TBM * TBN * 4);
  L.off_t = off;
  if (L.gmul == 3) off = align(off + (size_t)3 * d->n_out * d->m * L.t_ld * 4);
  L.total = off;
  return L;
}

int make_plane_map(CUtensorMap *map, const void *hi_plane, int64_t kf, int64_t rows, int64_t ldk, int nblocks, int64_t plane_bytes,
                   int box_rows, const char *name, int kb, bool bf16 = false) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return RNB_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)kf, (cuuint64_t)rows, (cuuint64_t)nblocks, 2};
  const int eb = bf16 ? 2 : 4;
  cuuint64_t strides[3] = {(cuuint64_t)(ldk * eb), (cuuint64_t)(rows * ldk * eb), (cuuint64_t)plane_bytes};
  cuuint32_t box[4] = {(cuuint32_t)kb, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // the swizzle span equals the row: kb k-values of fp32 (128 / 64 bytes) or of bf16 (64 / 32 bytes)
  const int row_bytes = kb * eb;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult res = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(hi_plane), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s planes) failed with CUresult %d", name, (int)res);
    return RNB_ERR_CUDA;
  }
  return RNB_OK;
}

}  // namespace

int contract_tf32_workspace(const rnb_contract_desc *d, size_t *bytes) {
  *bytes = make_layout(d).total;
  return RNB_OK;
}

// where rnb_permute_split has to put the planes of this contraction (see include/rosnet_b200.h)
int contract_tf32_planes(const rnb_contract_desc *d, rnb_operand_planes *a, rnb_operand_planes *b) {
  const Layout L = make_layout(d);
  memset(a, 0, sizeof(*a));
  memset(b, 0, sizeof(*b));
  a->hi_offset = L.off_ah; a->lo_offset = L.off_al;
  b->hi_offset = L.off_bh; b->lo_offset = L.off_bl;
  // the fused store writes at the matricized operand's own pitch: the plane pitch must equal it (no padding columns)
  if (d->precision == RNB_PREC_TF32_BF16X2 || L.ldk != L.kf) return RNB_OK;   // modes stay RNB_SPLIT_NONE
  if (L.gmul == 3) {
    a->mode = b->mode = RNB_SPLIT_C64_3M;
  } else {
    a->mode = RNB_SPLIT_PLANES;
    b->mode = d->dtype == RNB_C64 ? RNB_SPLIT_C64_B4M : RNB_SPLIT_PLANES;
  }
  return RNB_OK;
}

int contract_tf32(const rnb_contract_desc *d, cudaStream_t stream) {
  if (d->a_major != RNB_MAJOR_K || d->b_major != RNB_MAJOR_K) {
    set_error("float32/complex64 contraction takes K-major operands only (matricize MN-major operands first)");
    return RNB_ERR_UNSUPPORTED;
  }
  const Layout L = make_layout(d);
  const bool three_m = L.gmul == 3;
  if (d->flags & (RNB_FLAG_A_PRESPLIT | RNB_FLAG_B_PRESPLIT)) {
    RNB_REQUIRE(!bf16x2_requested(d) && L.ldk == L.kf, "pre-split operands need the 3xTF32 arithmetic and k %% 8 == 0 in floats (rnb_contract_planes)");
  }
  if (!d->workspace || d->workspace_bytes < L.total) {
    set_error("workspace of %zu bytes required (rnb_contract_workspace_bytes), got %zu", L.total, d->workspace_bytes);
    return RNB_ERR_WORKSPACE;
  }
  RNB_REQUIRE((reinterpret_cast<uintptr_t>(d->workspace) & 255) == 0, "workspace must be 256-byte aligned");
  unsigned char *ws = static_cast<unsigned char *>(d->workspace);
  float *ah = reinterpret_cast<float *>(ws + L.off_ah), *al = reinterpret_cast<float *>(ws + L.off_al);
  float *bh = reinterpret_cast<float *>(ws + L.off_bh), *bl = reinterpret_cast<float *>(ws + L.off_bl);
  const int cap = sm_count() * 16;

  const bool bf16x2 = d->precision == RNB_PREC_TF32_BF16X2;
  if (bf16x2 && L.gmul == 3) {
    // ---- 3M planes of the TF32 + 2 x BF16 mode ----
    struct Op16 { const void *p; int64_t ld, stride; int nblocks; int64_t rows; float *hi; __nv_bfloat16 *b16; int64_t plane; };
    const Op16 ops16[2] = {{d->a, d->a_ld, d->a_block_stride, d->a_nblocks, d->m, ah, reinterpret_cast<__nv_bfloat16 *>(al), L.a_plane},
                           {d->b, d->b_ld, d->b_block_stride, d->b_nblocks, d->n, bh, reinterpret_cast<__nv_bfloat16 *>(bl), L.b_plane}};
    for (const Op16 &o : ops16) {
      const bool pairs = (d->k % 2 == 0) && (o.ld % 2 == 0) && (o.stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(o.p) & 15) == 0);
      const int kh = pairs ? (int)(d->k / 2) : (int)d->k;
      const int64_t total = (int64_t)o.nblocks * o.rows * kh;
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_3m_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float2 *>(o.p), o.hi, o.b16, o.b16 + o.plane, total, (int)o.rows, kh,
                                                              o.ld, o.stride, L.ldk, pairs ? 1 : 0);
      RNB_CHECK_CUDA(cudaGetLastError());
    }
  } else if (bf16x2) {
    // ---- split pre-pass of the TF32 + 2 x BF16 mode: the "lo" plane's bytes hold the two bf16 planes ----
    __nv_bfloat16 *a16 = reinterpret_cast<__nv_bfloat16 *>(al), *b16 = reinterpret_cast<__nv_bfloat16 *>(bl);
    auto vok = [&](const void *base, int64_t ld_f, int64_t stride_f) {
      return (L.kf % 4 == 0) && (ld_f % 4 == 0) && (stride_f % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    };
    if (vok(d->a, d->a_ld * L.f, d->a_block_stride * L.f)) {
      const int64_t total = (int64_t)d->a_nblocks * d->m * (L.kf / 4);
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_rows_vec4_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(d->a), reinterpret_cast<float4 *>(ah),
                                                                     reinterpret_cast<uint2 *>(a16), reinterpret_cast<uint2 *>(a16 + L.a_plane), total,
                                                                     (int)d->m, (int)(L.kf / 4), d->a_ld * L.f / 4, d->a_block_stride * L.f / 4,
                                                                     L.ldk / 4);
      RNB_CHECK_CUDA(cudaGetLastError());
    } else {
      const int64_t total = (int64_t)d->a_nblocks * d->m * L.kf;
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float *>(d->a), ah, a16, a16 + L.a_plane, total, (int)d->m,
                                                                (int)L.kf, d->a_ld * L.f, d->a_block_stride * L.f, L.ldk);
      RNB_CHECK_CUDA(cudaGetLastError());
    }
    if (d->dtype == RNB_C64 && vok(d->b, d->b_ld * 2, d->b_block_stride * 2)) {
      const int64_t total = (int64_t)d->b_nblocks * d->n * (d->k / 2);
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_b_complex_vec_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(d->b), reinterpret_cast<float4 *>(bh),
                                                                         reinterpret_cast<uint2 *>(b16), reinterpret_cast<uint2 *>(b16 + L.b_plane),
                                                                         total, (int)d->n, (int)(d->k / 2), d->b_ld / 2, d->b_block_stride / 2,
                                                                         L.ldk / 4);
      RNB_CHECK_CUDA(cudaGetLastError());
    } else if (d->dtype == RNB_C64) {
      const int64_t total = (int64_t)d->b_nblocks * d->n * d->k;
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_b_complex_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float2 *>(d->b), bh, b16, b16 + L.b_plane, total,
                                                                     (int)d->n, (int)d->k, d->b_ld, d->b_block_stride, L.ldk);
      RNB_CHECK_CUDA(cudaGetLastError());
    } else if (vok(d->b, d->b_ld, d->b_block_stride)) {
      const int64_t total = (int64_t)d->b_nblocks * d->n * (L.kf / 4);
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_rows_vec4_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(d->b), reinterpret_cast<float4 *>(bh),
                                                                     reinterpret_cast<uint2 *>(b16), reinterpret_cast<uint2 *>(b16 + L.b_plane), total,
                                                                     (int)d->n, (int)(L.kf / 4), d->b_ld / 4, d->b_block_stride / 4, L.ldk / 4);
      RNB_CHECK_CUDA(cudaGetLastError());
    } else {
      const int64_t total = (int64_t)d->b_nblocks * d->n * L.kf;
      int64_t blocks = (total + 255) / 256;
      if (blocks > cap) blocks = cap;
      split16_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float *>(d->b), bh, b16, b16 + L.b_plane, total, (int)d->n,
                                                                (int)L.kf, d->b_ld, d->b_block_stride, L.ldk);
      RNB_CHECK_CUDA(cudaGetLastError());
    }
  }
  // ---- split pre-pass ----
  auto vec_ok = [&](const void *base, int64_t ld_f, int64_t stride_f) {
    return (L.kf % 4 == 0) && (ld_f % 4 == 0) && (stride_f % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  };
  if (three_m && bf16x2) {
    // planes written above
  } else if (three_m) {
    // ---- 3M: every complex block -> three real blocks (re, im, re + im), hi and lo planes ----
    struct Op { const void *p; int64_t ld, stride; int nblocks; int64_t rows; float *hi, *lo; };
    const Op ops[2] = {{d->a, d->a_ld, d->a_block_stride, d->a_nblocks, d->m, ah, al}, {d->b, d->b_ld, d->b_block_stride, d->b_nblocks, d->n, bh, bl}};
    for (int oi = 0; oi < 2; ++oi) {
      const Op &o = ops[oi];
      if (d->flags & (oi == 0 ? RNB_FLAG_A_PRESPLIT : RNB_FLAG_B_PRESPLIT)) continue;   // planes written by rnb_permute_split
      const bool v = (d->k % 4 == 0) && (o.ld % 2 == 0) && (o.stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(o.p) & 15) == 0);
      if (v) {
        const int64_t total = (int64_t)o.nblocks * o.rows * (d->k / 4);
        int64_t blocks = (total + 255) / 256;
        if (blocks > cap) blocks = cap;
        split3m_vec_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(o.p), reinterpret_cast<float4 *>(o.hi),
                                                                 reinterpret_cast<float4 *>(o.lo), total, (int)o.rows, (int)(d->k / 4), o.ld / 2,
                                                                 o.stride / 2, L.ldk / 4);
      } else {
        const int64_t total = (int64_t)o.nblocks * o.rows * d->k;
        int64_t blocks = (total + 255) / 256;
        if (blocks > cap) blocks = cap;
        split3m_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float2 *>(o.p), o.hi, o.lo, total, (int)o.rows, (int)d->k, o.ld,
                                                             o.stride, L.ldk);
      }
      RNB_CHECK_CUDA(cudaGetLastError());
    }
  } else if (bf16x2 || (d->flags & RNB_FLAG_A_PRESPLIT)) {
    // planes already written (above, or by rnb_permute_split)
  } else if (vec_ok(d->a, d->a_ld * L.f, d->a_block_stride * L.f)) {
    const int64_t total = (int64_t)d->a_nblocks * d->m * (L.kf / 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_rows_vec4_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(d->a), reinterpret_cast<float4 *>(ah),
                                                                 reinterpret_cast<float4 *>(al), total, (int)d->m, (int)(L.kf / 4),
                                                                 d->a_ld * L.f / 4, d->a_block_stride * L.f / 4, L.ldk / 4);
    RNB_CHECK_CUDA(cudaGetLastError());
  } else {
    const int64_t total = (int64_t)d->a_nblocks * d->m * L.kf;
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float *>(d->a), ah, al, total, (int)d->m, (int)L.kf,
                                                            d->a_ld * L.f, d->a_block_stride * L.f, L.ldk);
    RNB_CHECK_CUDA(cudaGetLastError());
  }
  if (bf16x2 || three_m || (d->flags & RNB_FLAG_B_PRESPLIT)) {
  } else if (d->dtype == RNB_C64 && vec_ok(d->b, d->b_ld * 2, d->b_block_stride * 2)) {
    const int64_t total = (int64_t)d->b_nblocks * d->n * (d->k / 2);
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_b_complex_vec_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(d->b), reinterpret_cast<float4 *>(bh),
                                                                     reinterpret_cast<float4 *>(bl), total, (int)d->n, (int)(d->k / 2),
                                                                     d->b_ld / 2, d->b_block_stride / 2, L.ldk / 4);
    RNB_CHECK_CUDA(cudaGetLastError());
  } else if (d->dtype == RNB_C64) {
    const int64_t total = (int64_t)d->b_nblocks * d->n * d->k;
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_b_complex_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float2 *>(d->b), bh, bl, total, (int)d->n, (int)d->k,
                                                                 d->b_ld, d->b_block_stride, L.ldk);
    RNB_CHECK_CUDA(cudaGetLastError());
  } else if (vec_ok(d->b, d->b_ld, d->b_block_stride)) {
    const int64_t total = (int64_t)d->b_nblocks * d->n * (L.kf / 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_rows_vec4_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float4 *>(d->b), reinterpret_cast<float4 *>(bh),
                                                                 reinterpret_cast<float4 *>(bl), total, (int)d->n, (int)(L.kf / 4),
                                                                 d->b_ld / 4, d->b_block_stride / 4, L.ldk / 4);
    RNB_CHECK_CUDA(cudaGetLastError());
  } else {
    const int64_t total = (int64_t)d->b_nblocks * d->n * L.kf;
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float *>(d->b), bh, bl, total, (int)d->n, (int)L.kf,
                                                            d->b_ld, d->b_block_stride, L.ldk);
    RNB_CHECK_CUDA(cudaGetLastError());
  }

  // ---- grouped GEMM ----
  CUtensorMap ma, mb, ma16, mb16;
  int rc = make_plane_map(&ma, ah, L.kf, d->m, L.ldk, d->a_nblocks * L.gmul, (int64_t)(L.off_al - L.off_ah), TBM, "a", L.kb);
  if (rc != RNB_OK) return rc;
  rc = make_plane_map(&mb, bh, L.kf, L.nf, L.ldk, d->b_nblocks * L.gmul, (int64_t)(L.off_bl - L.off_bh), TBN / L.cg, "b", L.kb);
  if (rc != RNB_OK) return rc;
  ma16 = ma;
  mb16 = mb;
  if (bf16x2) {
    rc = make_plane_map(&ma16, al, L.kf, d->m, L.ldk, d->a_nblocks * L.gmul, L.a_plane * 2, TBM, "a (bf16)", L.kb, true);
    if (rc != RNB_OK) return rc;
    rc = make_plane_map(&mb16, bl, L.kf, L.nf, L.ldk, d->b_nblocks * L.gmul, L.b_plane * 2, TBN / L.cg, "b (bf16)", L.kb, true);
    if (rc != RNB_OK) return rc;
  }
  Tf32Params p;
  p.m = d->m;
  p.n = L.nf;
  p.c = static_cast<float *>(d->c);
  p.c_ld = d->c_ld * L.f;
  p.c_block_stride = d->c_block_stride * L.f;
  p.pair_a = d->pair_a; p.pair_b = d->pair_b; p.out_index = d->out_index;
  p.n_out = d->n_out * L.gmul; p.n_pairs = d->n_pairs;
  p.gmul = L.gmul;
  p.tiles_m = L.tiles_m;
  p.tiles_n = L.tiles_n;
  p.k_blocks = L.k_blocks;
  p.accumulate = d->accumulate;
  p.products = (d->precision == RNB_PREC_TF32X1) ? 1 : (bf16x2 ? 2 : 3);
  {
    static EnvKnob knob_order("RNB_TF32_ORDER");
    const char *e = knob_order.get();
    p.interleave_terms = (e && !strcmp(e, "interleaved")) ? 1 : 0;
  }
  p.c_vec_ok = ((reinterpret_cast<uintptr_t>(d->c) & 15) == 0) && ((p.c_ld * 4) % 16 == 0) && ((p.c_block_stride * 4) % 16 == 0);
  if (three_m) {   // the real GEMMs write the product workspace; combine3m_kernel writes (or accumulates into) C
    p.c = reinterpret_cast<float *>(ws + L.off_t);
    p.c_ld = L.t_ld;
    p.c_block_stride = d->m * L.t_ld;
    p.out_index = nullptr;
    p.accumulate = 0;
    p.c_vec_ok = 1;
  }
  p.chunk = L.chunk;
  p.splits = L.splits;
  p.iters_per_split = L.iters_per_split;
  p.partial = reinterpret_cast<float *>(ws + L.off_partial);
  p.hint_a = p.hint_b = 0;
  static EnvKnob knob_hint("RNB_TF32_L2HINT");
  if (const char *e = knob_hint.get()) {
    const int v = atoi(e);
    if (v == 1) { p.hint_a = 0x12F0000000000000ull; p.hint_b = 0x14F0000000000000ull; }   // A evict_first, B evict_last
    if (v == 2) { p.hint_a = 0x1000000000000000ull; p.hint_b = 0x14F0000000000000ull; }   // A normal, B evict_last
    if (v == 3) { p.hint_a = 0x12F0000000000000ull; p.hint_b = 0x1000000000000000ull; }   // A evict_first, B normal
  }
  p.panel = p.tiles_n > 12 ? 8 : p.tiles_n;
  static EnvKnob knob_panel("RNB_TF32_PANEL");
  if (const char *e = knob_panel.get()) {
    const int v = atoi(e);
    if (v >= 1) p.panel = v < p.tiles_n ? v : p.tiles_n;
  }
  const int64_t total = (int64_t)p.n_out * p.tiles_m * p.tiles_n * p.splits;
  if (total <= 0) return RNB_OK;
  RNB_REQUIRE(total < (1ll << 31), "too many tiles");
  int grid = sm_count() / L.cg;
  if (total < grid) grid = (int)total;
  int dev = 0;
  RNB_CHECK_CUDA(cudaGetDevice(&dev));
  static bool attr_done[4][64];   // the opt-in is a per-device attribute
  auto opt_in = [&](int which, const void *fn, int bytes) -> cudaError_t {
    if (dev >= 0 && dev < 64 && attr_done[which][dev]) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) attr_done[which][dev] = true;
    return e;
  };
  if (L.v2 && L.cg == 2) {
    RNB_CHECK_CUDA(opt_in(3, reinterpret_cast<const void *>(gemm_tf32_v2_kernel<2>), Pipe2<2>::SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid * 2);
    cfg.blockDim = dim3(T_THREADS);
    cfg.dynamicSmemBytes = Pipe2<2>::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_v2_kernel<2>, ma, mb, ma16, mb16, p));
  } else if (L.v2) {
    RNB_CHECK_CUDA(opt_in(2, reinterpret_cast<const void *>(gemm_tf32_v2_kernel<1>), Pipe2<1>::SMEM));
    gemm_tf32_v2_kernel<1><<<grid, T_THREADS, Pipe2<1>::SMEM, stream>>>(ma, mb, ma16, mb16, p);
  } else if (L.kb == 16) {
    RNB_CHECK_CUDA(opt_in(1, reinterpret_cast<const void *>(gemm_tf32_kernel<16, 4>), Pipe<16, 4>::SMEM));
    gemm_tf32_kernel<16, 4><<<grid, T_THREADS, Pipe<16, 4>::SMEM, stream>>>(ma, mb, ma16, mb16, p);
  } else {
    RNB_CHECK_CUDA(opt_in(0, reinterpret_cast<const void *>(gemm_tf32_kernel<32, 2>), Pipe<32, 2>::SMEM));
    gemm_tf32_kernel<32, 2><<<grid, T_THREADS, Pipe<32, 2>::SMEM, stream>>>(ma, mb, ma16, mb16, p);
  }
  RNB_CHECK_CUDA(cudaGetLastError());
  if (p.splits > 1) {
    const int64_t n_tiles = (int64_t)p.n_out * L.tiles_m128 * p.tiles_n;
    splitk_reduce_kernel<float><<<(unsigned)n_tiles, 256, 0, stream>>>(p.partial, p.c, p.out_index, p.m, p.n, p.c_ld, p.c_block_stride,
                                                                       L.tiles_m128, p.tiles_n, p.panel, TBM, TBN, p.splits, p.accumulate, p.c_vec_ok);
    RNB_CHECK_CUDA(cudaGetLastError());
  }
  if (three_m) {
    const bool cv = (d->n % 4 == 0) && ((reinterpret_cast<uintptr_t>(d->c) & 15) == 0) && (d->c_ld % 2 == 0) && (d->c_block_stride % 2 == 0);
    const int64_t work = (int64_t)d->n_out * d->m * (cv ? d->n / 4 : d->n);
    int64_t blocks = (work + 255) / 256;
    if (blocks > cap) blocks = cap;
    combine3m_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p.c, static_cast<float2 *>(d->c), d->out_index, d->n_out, d->m, d->n, L.t_ld, d->c_ld,
                                                           d->c_block_stride, d->accumulate, cv ? 1 : 0);
    RNB_CHECK_CUDA(cudaGetLastError());
  }
  return RNB_OK;
}

}  // namespace rnb
