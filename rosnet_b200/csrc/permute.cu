// Strided n-d copy ("permute / matricize"): moves blocks into GEMM layout, transposes BlockArrays,
// converts between dense and block-major storage.
//
// Replaces `a.transpose(newaxes).reshape(M, K)` inside numpy.tensordot (called per block pair at
// rosnet/array/block/__init__.py:283) and the per-block transpose of
// rosnet/array/block/__init__.py:272-273.  Pure data movement, bit-exact, HBM-bound:
// algorithmic bytes = 2 * elem_bytes * prod(extent).
//
// Design (persistent CTAs, shared-memory tiles, both sides of HBM coalesced):
//   host  : drop extent-1 axes, merge axes that are adjacent in both source and destination,
//           pick a tile that spans >= 512 B contiguous along the source-fastest axes AND along the
//           destination-fastest axes (they may be many tiny axes, as in tensor-network tensors);
//   load  : rows of the tile (contiguous source runs) are staged into shared memory either with the
//           TMA bulk-copy engine (cp.async.bulk -> SASS UBLKCP, one mbarrier per buffer) when the
//           runs are 16-byte aligned multiples, or with 4/8/16-byte cp.async (LDGSTS);
//           rows are padded to an odd multiple of 16 B so transposed reads spread over the banks;
//   store : threads walk the tile in destination order through two small look-up tables
//           (position-in-tile and destination offset, each split low/high so they stay tiny)
//           and write 128-bit vectors whenever the destination run allows;
//   the next tile's loads are issued before the current tile's stores (double buffering).
#include <algorithm>
#include <vector>

#include "common.h"

namespace rnb {
namespace {

constexpr int MAXD = 32;        // merged dims handled by the tiled kernels (= RNB_MAX_RANK)
constexpr int kThreadsP = 256;
constexpr int kMaxLut = 4096;

struct HighEnt {     // per-row (high group) entry, read as a warp-wide broadcast
  uint32_t off;      // element offset relative to the tile origin (source or destination)
  uint32_t pos;      // byte position of the row start inside the padded shared-memory tile
  uint16_t i0, i1;   // tile-local index along partial dims 0 / 1
};
struct IdxEnt {      // low-group partial-dim indices, only read on edge tiles
  uint16_t i0, i1;
};

struct TileDim {
  int32_t t;            // tile extent
  int32_t slot;         // partial slot (-1 none)
  int64_t ss, ds;       // strides in elements
  int32_t smem_stride;  // bytes
  int32_t pad_;
};

struct PermParams {
  const unsigned char *src;
  unsigned char *dst;
  int32_t es;
  // tile dims and their grouping (indices into td[], inner -> outer)
  int32_t n_td;
  TileDim td[MAXD];
  int32_t n_slow, n_shigh, n_dlow, n_dhigh;
  int8_t g_slow[MAXD], g_shigh[MAXD], g_dlow[MAXD], g_dhigh[MAXD];
  int32_t s_low, s_high, d_low, d_high;  // group volumes (elements / rows)
  int32_t unit_elems;                    // elements per load unit (load granularity / es)
  int32_t units_per_row, units_per_row_magic;
  int32_t vs;                            // elements per store vector
  int32_t dunits_per_row, dunits_per_row_magic;
  int32_t row_bytes, row_bytes_padded, tile_bytes;
  int32_t load_mode;                     // 0 = cp.async, 1 = TMA bulk rows
  int32_t s_low_contig, d_low_contig;    // low-group offsets are simply the running index
  // tile-grid odometer (dims with more than one tile, inner -> outer)
  int32_t n_od;
  int32_t od_ntiles[MAXD];
  int32_t od_t[MAXD];
  int32_t od_slot[MAXD];
  int64_t od_extent[MAXD];
  int64_t od_sstep[MAXD], od_dstep[MAXD];  // t*stride, elements
  int64_t total_tiles;
  int32_t tiles_per_cta;
  int32_t ctrl_offset;                     // byte offset of the Ctrl block inside dynamic shared memory
  int32_t slot_t[2];                       // tile extent of partial dim in each slot
  int32_t slot_inner_elems[2];             // (bulk mode) elements per index step if slot dim is the outermost low-group dim, else 0
  // fused split epilogue (MODE != 0): destination = the hi / lo planes of csrc/gemm_tf32.cu
  float *split_hi, *split_lo;
  int64_t split_k, split_rk;               // contracted extent; rows * k of one matricized block (elements)
  int32_t split_k_pow2, split_multi;       // k is a power of two; more than one block
};

struct Ctrl {
  int64_t src_base, dst_base;  // elements
  int32_t rem0, rem1;          // remaining extent along the partial dims (>= t when the tile is full)
  int32_t edge;
  int32_t pad_;
};

// magic == 0 encodes a divisor of 1 (magic_for): no 32-bit magic number represents it
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t magic) { return magic ? __umulhi(x, magic) : x; }

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
  else if (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int ES>
struct ElemT;
template <>
struct ElemT<4> { typedef uint32_t type; };
template <>
struct ElemT<8> { typedef uint2 type; };
template <>
struct ElemT<16> { typedef uint4 type; };

// decode `e` over the dims of a group (inner -> outer) and accumulate what the LUTs need
__device__ __forceinline__ void decode_group(const PermParams &p, const int8_t *grp, int n, uint32_t e, int64_t &soff,
                                             int64_t &doff, uint32_t &pos, uint32_t &i0, uint32_t &i1) {
  soff = 0; doff = 0; pos = 0; i0 = 0; i1 = 0;
  for (int j = 0; j < n; ++j) {
    const TileDim &d = p.td[grp[j]];
    const uint32_t digit = e % (uint32_t)d.t;
    e /= (uint32_t)d.t;
    soff += (int64_t)digit * d.ss;
    doff += (int64_t)digit * d.ds;
    pos += digit * (uint32_t)d.smem_stride;
    if (d.slot == 0) i0 = digit;
    if (d.slot == 1) i1 = digit;
  }
}

// MODE (fused split epilogue, rnb_permute_split): 0 plain copy; 1 hi/lo planes at the destination's own offsets (float32
// operands, complex64 A operand of the 4M arrangement: the interleaved float view); 2 complex64 B operand of the 4M
// arrangement (rows 2n = (re, -im), 2n+1 = (im, re)); 3 complex64 operand of the 3M arrangement (three real blocks
// re / im / re+im per block).  Modes 1-3 need 16-byte store vectors (VS * ES == 16).
template <int ES, int VS, int MODE = 0>
__global__ void __launch_bounds__(kThreadsP) permute_tiled_kernel(const __grid_constant__ PermParams p) {
  typedef typename ElemT<ES>::type elem_t;
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char *buf0 = sm;
  unsigned char *buf1 = sm + p.tile_bytes;
  // compact look-up tables (tile-invariant)
  HighEnt *srcHigh = reinterpret_cast<HighEnt *>(sm + 2 * p.tile_bytes);   // [s_high]
  HighEnt *dstHigh = srcHigh + p.s_high;                                      // [d_high]
  uint32_t *srcLowOff = reinterpret_cast<uint32_t *>(dstHigh + p.d_high);     // [units_per_row]
  uint32_t *dstLowOff = srcLowOff + p.units_per_row;                          // [dunits_per_row]
  IdxEnt *srcLowIdx = reinterpret_cast<IdxEnt *>(dstLowOff + p.dunits_per_row);  // [units_per_row]
  IdxEnt *dstLowIdx = srcLowIdx + p.units_per_row;                            // [dunits_per_row]
  uint16_t *dstLowPos = reinterpret_cast<uint16_t *>(dstLowIdx + p.dunits_per_row);  // [d_low]
  Ctrl *ctrl = reinterpret_cast<Ctrl *>(sm + p.ctrl_offset);
  uint64_t *bars = reinterpret_cast<uint64_t *>(ctrl + 2);
  int32_t *digits = reinterpret_cast<int32_t *>(bars + 2);

  const int tid = threadIdx.x;
  const int64_t first_tile = (int64_t)blockIdx.x * p.tiles_per_cta;
  int64_t n_my = p.total_tiles - first_tile;
  if (n_my > p.tiles_per_cta) n_my = p.tiles_per_cta;
  if (n_my <= 0) return;

  // ---- build the look-up tables ----
  for (int e = tid; e < p.units_per_row; e += kThreadsP) {
    int64_t so, dn; uint32_t pos, i0, i1;
    decode_group(p, p.g_slow, p.n_slow, (uint32_t)e * p.unit_elems, so, dn, pos, i0, i1);
    srcLowOff[e] = (uint32_t)so; srcLowIdx[e].i0 = (uint16_t)i0; srcLowIdx[e].i1 = (uint16_t)i1;
  }
  for (int e = tid; e < p.s_high; e += kThreadsP) {
    int64_t so, dn; uint32_t pos, i0, i1;
    decode_group(p, p.g_shigh, p.n_shigh, (uint32_t)e, so, dn, pos, i0, i1);
    srcHigh[e].off = (uint32_t)so; srcHigh[e].pos = pos; srcHigh[e].i0 = (uint16_t)i0; srcHigh[e].i1 = (uint16_t)i1;
  }
  for (int e = tid; e < p.d_low; e += kThreadsP) {
    int64_t so, dn; uint32_t pos, i0, i1;
    decode_group(p, p.g_dlow, p.n_dlow, (uint32_t)e, so, dn, pos, i0, i1);
    dstLowPos[e] = (uint16_t)pos;
    if (e % VS == 0) {
      dstLowOff[e / VS] = (uint32_t)dn; dstLowIdx[e / VS].i0 = (uint16_t)i0; dstLowIdx[e / VS].i1 = (uint16_t)i1;
    }
  }
  for (int e = tid; e < p.d_high; e += kThreadsP) {
    int64_t so, dn; uint32_t pos, i0, i1;
    decode_group(p, p.g_dhigh, p.n_dhigh, (uint32_t)e, so, dn, pos, i0, i1);
    dstHigh[e].off = (uint32_t)dn; dstHigh[e].pos = pos; dstHigh[e].i0 = (uint16_t)i0; dstHigh[e].i1 = (uint16_t)i1;
  }

  // ---- tile odometer, owned by thread 0 ----
  auto write_ctrl = [&](Ctrl *c) {
    int64_t sb = 0, db = 0;
    int32_t rem0 = 0x7fffffff, rem1 = 0x7fffffff;
    for (int j = 0; j < p.n_od; ++j) {
      const int dg = digits[j];
      sb += (int64_t)dg * p.od_sstep[j];
      db += (int64_t)dg * p.od_dstep[j];
      if (p.od_slot[j] >= 0) {
        int64_t rem = p.od_extent[j] - (int64_t)dg * p.od_t[j];
        int32_t r32 = rem > 0x7fffffff ? 0x7fffffff : (int32_t)rem;
        if (p.od_slot[j] == 0) rem0 = r32; else rem1 = r32;
      }
    }
    c->src_base = sb; c->dst_base = db; c->rem0 = rem0; c->rem1 = rem1;
    c->edge = (rem0 < p.slot_t[0]) || (rem1 < p.slot_t[1]);
  };
  if (tid == 0) {
    int64_t tix = first_tile;
    for (int j = 0; j < p.n_od; ++j) {
      digits[j] = (int32_t)(tix % p.od_ntiles[j]);
      tix /= p.od_ntiles[j];
    }
    write_ctrl(&ctrl[0]);
    if (p.load_mode == 1) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_mbar_init();
    }
  }
  __syncthreads();

  const uint32_t total_units = (uint32_t)p.units_per_row * (uint32_t)p.s_high;
  const uint32_t total_dunits = (uint32_t)p.dunits_per_row * (uint32_t)p.d_high;
  const int G = p.unit_elems * ES;

  auto issue_loads = [&](const Ctrl &c, unsigned char *buf, uint64_t *bar) {
    const unsigned char *sbase = p.src + c.src_base * ES;
    if (p.load_mode == 1) {
      // TMA bulk copies: one per row, issued by warp 0
      if (tid < 32) {
        fence_proxy_async();  // generic-proxy reads of this buffer (previous tile) precede the async-proxy writes
        uint32_t row_bytes = p.row_bytes;
        if (c.edge) {
          if (p.slot_inner_elems[0] && c.rem0 < p.slot_t[0]) row_bytes = (uint32_t)c.rem0 * p.slot_inner_elems[0] * ES;
          if (p.slot_inner_elems[1] && c.rem1 < p.slot_t[1]) row_bytes = (uint32_t)c.rem1 * p.slot_inner_elems[1] * ES;
        }
        uint32_t nrows = 0;
        for (int rrow = tid; rrow < p.s_high; rrow += 32) {
          const HighEnt h = srcHigh[rrow];
          if (!c.edge || (h.i0 < c.rem0 && h.i1 < c.rem1)) ++nrows;
        }
        nrows = __reduce_add_sync(0xffffffffu, nrows);
        if (tid == 0) mbar_arrive_expect_tx(bar, nrows * row_bytes);
        __syncwarp();
        for (int rrow = tid; rrow < p.s_high; rrow += 32) {
          const HighEnt h = srcHigh[rrow];
          if (!c.edge || (h.i0 < c.rem0 && h.i1 < c.rem1))
            bulk_load(buf + h.pos, sbase + (size_t)h.off * ES, row_bytes, bar);
        }
      }
      return;
    }
    for (uint32_t u = tid; u < total_units; u += kThreadsP) {
      const uint32_t qh = fast_div(u, (uint32_t)p.units_per_row_magic);
      const uint32_t ul = u - qh * p.units_per_row;
      const HighEnt hi = srcHigh[qh];
      if (c.edge) {
        const IdxEnt li = srcLowIdx[ul];
        if (!((uint32_t)li.i0 + hi.i0 < (uint32_t)c.rem0 && (uint32_t)li.i1 + hi.i1 < (uint32_t)c.rem1)) continue;
      }
      const uint32_t lo_off = p.s_low_contig ? ul * (uint32_t)p.unit_elems : srcLowOff[ul];
      const unsigned char *g = sbase + ((size_t)lo_off + hi.off) * ES;
      unsigned char *s = buf + hi.pos + (size_t)ul * G;
      if (G == 16) cp_async<16>(s, g);
      else if (G == 8) cp_async<8>(s, g);
      else cp_async<4>(s, g);
    }
    cp_async_commit();
  };

  issue_loads(ctrl[0], buf0, &bars[0]);

  for (int64_t i = 0; i < n_my; ++i) {
    const int cur = (int)(i & 1);
    const bool has_next = (i + 1 < n_my);
    if (tid == 0 && has_next) {
      for (int j = 0; j < p.n_od; ++j) {  // odometer increment
        if (++digits[j] < p.od_ntiles[j]) break;
        digits[j] = 0;
      }
      write_ctrl(&ctrl[cur ^ 1]);
    }
    __syncthreads();  // ctrl visible; everyone is done reading buffer cur^1 (stores of tile i-1)
    if (has_next) {
      issue_loads(ctrl[cur ^ 1], cur ? buf0 : buf1, &bars[cur ^ 1]);
    }
    if (p.load_mode == 1) {
      mbar_wait(&bars[cur], (uint32_t)((i >> 1) & 1));
    } else {
      if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
    }

    // ---- store phase: destination order, vectorised ----
    const Ctrl c = ctrl[cur];
    const unsigned char *buf = cur ? buf1 : buf0;
    unsigned char *dbase = p.dst + c.dst_base * ES;
    for (uint32_t w = tid; w < total_dunits; w += kThreadsP) {
      const uint32_t qh = fast_div(w, (uint32_t)p.dunits_per_row_magic);
      const uint32_t ul = w - qh * p.dunits_per_row;
      const HighEnt hi = dstHigh[qh];
      if (c.edge) {
        const IdxEnt li = dstLowIdx[ul];
        if (!((uint32_t)li.i0 + hi.i0 < (uint32_t)c.rem0 && (uint32_t)li.i1 + hi.i1 < (uint32_t)c.rem1)) continue;
      }
      const uint32_t lo_off = p.d_low_contig ? ul * VS : dstLowOff[ul];
      elem_t v[VS];
#pragma unroll
      for (int e = 0; e < VS; ++e) v[e] = *reinterpret_cast<const elem_t *>(buf + hi.pos + dstLowPos[ul * VS + e]);
      if (MODE != 0) {
        const int64_t eoff = c.dst_base + (int64_t)hi.off + lo_off;   // element offset inside the matricized operand
        const float *x = reinterpret_cast<const float *>(&v[0]);     // the 16-byte vector as 4 floats
        float h[4], l[4];
        if (MODE == 3) {
          // (re0, im0, re1, im1) -> planes re, im, re + im of the 3M arrangement, two k-values each
          const int64_t blk = p.split_multi ? eoff / p.split_rk : 0;
          const int64_t base = eoff + 2 * blk * p.split_rk;
          const float val[3][2] = {{x[0], x[2]}, {x[1], x[3]}, {x[0] + x[1], x[2] + x[3]}};
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float h0 = to_tf32(val[j][0]), h1 = to_tf32(val[j][1]);
            *reinterpret_cast<float2 *>(p.split_hi + base + j * p.split_rk) = make_float2(h0, h1);
            *reinterpret_cast<float2 *>(p.split_lo + base + j * p.split_rk) = make_float2(to_tf32(val[j][0] - h0), to_tf32(val[j][1] - h1));
          }
          continue;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h[e] = to_tf32(x[e]);
          l[e] = to_tf32(x[e] - h[e]);
        }
        if (MODE == 1) {
          const int64_t off = eoff * (ES / 4);
          *reinterpret_cast<float4 *>(p.split_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4 *>(p.split_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
          const int64_t kk = p.split_k_pow2 ? (eoff & (p.split_k - 1)) : (eoff % p.split_k);
          const int64_t off = 4 * eoff - 2 * kk;      // row 2n of the real [2N][2K] matrix; row 2n+1 follows 2k floats later
          *reinterpret_cast<float4 *>(p.split_hi + off) = make_float4(h[0], -h[1], h[2], -h[3]);
          *reinterpret_cast<float4 *>(p.split_hi + off + 2 * p.split_k) = make_float4(h[1], h[0], h[3], h[2]);
          *reinterpret_cast<float4 *>(p.split_lo + off) = make_float4(l[0], -l[1], l[2], -l[3]);
          *reinterpret_cast<float4 *>(p.split_lo + off + 2 * p.split_k) = make_float4(l[1], l[0], l[3], l[2]);
        }
        continue;
      }
      unsigned char *g = dbase + ((size_t)hi.off + lo_off) * ES;
      if (ES * VS == 16) {
        uint4 o;
        if (ES == 16) {
          o = *reinterpret_cast<uint4 *>(&v[0]);
        } else if (ES == 8) {
          const uint2 *q = reinterpret_cast<const uint2 *>(&v[0]);
          o = make_uint4(q[0].x, q[0].y, q[VS > 1 ? 1 : 0].x, q[VS > 1 ? 1 : 0].y);
        } else {
          const uint32_t *q = reinterpret_cast<const uint32_t *>(&v[0]);
          o = make_uint4(q[0], q[VS > 1 ? 1 : 0], q[VS > 2 ? 2 : 0], q[VS > 3 ? 3 : 0]);
        }
        *reinterpret_cast<uint4 *>(g) = o;
      } else if (ES * VS == 8) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(&v[0]);
        *reinterpret_cast<uint2 *>(g) = make_uint2(q[0], q[1]);
      } else {
        *reinterpret_cast<uint32_t *>(g) = *reinterpret_cast<const uint32_t *>(&v[0]);
      }
    }
  }
}


// =================================================================================================
// 8-byte elements, true transposition (source-fastest axis i != destination-fastest axis o):
// warp-specialised pipeline + 2x2 register micro-transpose.
//   producer warp : per tile, contiguous source chunks -> dense shared-memory tile with TMA bulk copies
//                   (cp.async.bulk, completion on a full[] mbarrier), ring of `stages` buffers
//   8 consumer warps: each lane takes a 2(i) x 2(o) micro tile: two 128-bit shared loads (rows o, o+1;
//                   the 8 lanes of a quarter warp read one contiguous 128 B run -> conflict free, no
//                   padding), swaps the halves in registers and issues two 128-bit global stores
//                   (columns i, i+1); an empty[] mbarrier hands the buffer back. No CTA-wide barrier.
// =================================================================================================
constexpr int kMicroConsumers = 8;
constexpr int kMicroThreads = (kMicroConsumers + 1) * 32;
constexpr int kMaxStages = 4;

struct MicroParams {
  const unsigned char *src;
  unsigned char *dst;
  int32_t n_td;
  TileDim td[MAXD];  // tile dims, source order inner -> outer; smem_stride = dense byte stride
  int32_t n_dig;     // micro-unit enumeration, lane-fastest digit first
  int8_t dig_td[MAXD + 4];
  int32_t dig_radix[MAXD + 4];
  int32_t dig_mul[MAXD + 4];
  int32_t n_units;   // micro tiles per tile
  int32_t so_bytes;  // shared-memory byte stride of one step along o
  int64_t di_elems;  // destination stride of one step along i
  int32_t chunk_bytes, n_chunks;
  int32_t n_cdims;
  int8_t cdim_td[MAXD];
  int32_t chunk_slot, chunk_slot_inner_bytes;
  int32_t tile_bytes, stages;
  int32_t lut_offset, ctrl_offset;
  int32_t n_od;
  int32_t od_ntiles[MAXD];
  int32_t od_t[MAXD];
  int32_t od_slot[MAXD];
  int64_t od_extent[MAXD];
  int64_t od_sstep[MAXD], od_dstep[MAXD];
  int64_t total_tiles;
  int32_t tiles_per_cta;
  int32_t slot_t[2];
};

struct ChunkEnt {
  uint32_t off;  // source element offset of the chunk relative to the tile origin
  uint16_t i0, i1;
};

__global__ void __launch_bounds__(kMicroThreads) permute_micro8_kernel(const __grid_constant__ MicroParams p) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint2 *unitLut = reinterpret_cast<uint2 *>(sm + p.lut_offset);            // {smem byte pos, dst element offset}
  IdxEnt *unitIdx = reinterpret_cast<IdxEnt *>(unitLut + p.n_units);
  ChunkEnt *chunkLut = reinterpret_cast<ChunkEnt *>(unitIdx + p.n_units);
  Ctrl *ctrl = reinterpret_cast<Ctrl *>(sm + p.ctrl_offset);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(ctrl + kMaxStages);
  uint64_t *empty_bar = full_bar + kMaxStages;
  int32_t *digits = reinterpret_cast<int32_t *>(empty_bar + kMaxStages);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t first_tile = (int64_t)blockIdx.x * p.tiles_per_cta;
  int64_t n_my = p.total_tiles - first_tile;
  if (n_my > p.tiles_per_cta) n_my = p.tiles_per_cta;
  if (n_my <= 0) return;

  // ---- tile-invariant tables ----
  for (int w = tid; w < p.n_units; w += kMicroThreads) {
    int idx[MAXD];
#pragma unroll
    for (int j = 0; j < MAXD; ++j) idx[j] = 0;
    uint32_t e = (uint32_t)w;
    for (int j = 0; j < p.n_dig; ++j) {
      const uint32_t digit = e % (uint32_t)p.dig_radix[j];
      e /= (uint32_t)p.dig_radix[j];
      idx[p.dig_td[j]] += (int)digit * p.dig_mul[j];
    }
    uint32_t spos = 0, i0 = 0, i1 = 0;
    int64_t doff = 0;
    for (int j = 0; j < p.n_td; ++j) {
      spos += (uint32_t)idx[j] * (uint32_t)p.td[j].smem_stride;
      doff += (int64_t)idx[j] * p.td[j].ds;
      if (p.td[j].slot == 0) i0 = idx[j];
      if (p.td[j].slot == 1) i1 = idx[j];
    }
    unitLut[w] = make_uint2(spos, (uint32_t)doff);
    unitIdx[w].i0 = (uint16_t)i0;
    unitIdx[w].i1 = (uint16_t)i1;
  }
  for (int c = tid; c < p.n_chunks; c += kMicroThreads) {
    uint32_t e = (uint32_t)c, i0 = 0, i1 = 0;
    int64_t so = 0;
    for (int j = 0; j < p.n_cdims; ++j) {
      const TileDim &d = p.td[p.cdim_td[j]];
      const uint32_t digit = e % (uint32_t)d.t;
      e /= (uint32_t)d.t;
      so += (int64_t)digit * d.ss;
      if (d.slot == 0) i0 = digit;
      if (d.slot == 1) i1 = digit;
    }
    chunkLut[c].off = (uint32_t)so;
    chunkLut[c].i0 = (uint16_t)i0;
    chunkLut[c].i1 = (uint16_t)i1;
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kMicroConsumers);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kMicroConsumers) {
    // ===== producer warp =====
    if (lane == 0) {
      int64_t tix = first_tile;
      for (int j = 0; j < p.n_od; ++j) {
        digits[j] = (int32_t)(tix % p.od_ntiles[j]);
        tix /= p.od_ntiles[j];
      }
    }
    for (int64_t i = 0; i < n_my; ++i) {
      const int stage = (int)(i % p.stages);
      const uint32_t phase = (uint32_t)((i / p.stages) & 1);
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (lane == 0) {
        if (i > 0) {
          for (int j = 0; j < p.n_od; ++j) {
            if (++digits[j] < p.od_ntiles[j]) break;
            digits[j] = 0;
          }
        }
        int64_t sb = 0, db = 0;
        int32_t rem0 = 0x7fffffff, rem1 = 0x7fffffff;
        for (int j = 0; j < p.n_od; ++j) {
          const int dg = digits[j];
          sb += (int64_t)dg * p.od_sstep[j];
          db += (int64_t)dg * p.od_dstep[j];
          if (p.od_slot[j] >= 0) {
            const int64_t rem = p.od_extent[j] - (int64_t)dg * p.od_t[j];
            const int32_t r32 = rem > 0x7fffffff ? 0x7fffffff : (int32_t)rem;
            if (p.od_slot[j] == 0) rem0 = r32; else rem1 = r32;
          }
        }
        Ctrl *c = &ctrl[stage];
        c->src_base = sb; c->dst_base = db; c->rem0 = rem0; c->rem1 = rem1;
        c->edge = (rem0 < p.slot_t[0]) || (rem1 < p.slot_t[1]);
      }
      __syncwarp();
      const Ctrl c = ctrl[stage];
      fence_proxy_async();
      uint32_t bytes = (uint32_t)p.chunk_bytes;
      if (c.edge && p.chunk_slot >= 0) {
        const int32_t rem = p.chunk_slot == 0 ? c.rem0 : c.rem1;
        if (rem < p.slot_t[p.chunk_slot]) bytes = (uint32_t)rem * (uint32_t)p.chunk_slot_inner_bytes;
      }
      uint32_t nvalid = 0;
      for (int k = lane; k < p.n_chunks; k += 32) {
        const ChunkEnt e = chunkLut[k];
        if (!c.edge || (e.i0 < c.rem0 && e.i1 < c.rem1)) ++nvalid;
      }
      nvalid = __reduce_add_sync(0xffffffffu, nvalid);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], nvalid * bytes);
      __syncwarp();
      unsigned char *buf = sm + (size_t)stage * p.tile_bytes;
      const unsigned char *sbase = p.src + c.src_base * 8;
      for (int k = lane; k < p.n_chunks; k += 32) {
        const ChunkEnt e = chunkLut[k];
        if (!c.edge || (e.i0 < c.rem0 && e.i1 < c.rem1))
          bulk_load(buf + (size_t)k * p.chunk_bytes, sbase + (size_t)e.off * 8, bytes, &full_bar[stage]);
      }
    }
    return;
  }

  // ===== consumers =====
  const int ctid = tid;  // 0 .. 255
  for (int64_t i = 0; i < n_my; ++i) {
    const int stage = (int)(i % p.stages);
    const uint32_t phase = (uint32_t)((i / p.stages) & 1);
    mbar_wait(&full_bar[stage], phase);
    const Ctrl c = ctrl[stage];
    const unsigned char *buf = sm + (size_t)stage * p.tile_bytes;
    unsigned char *dbase = p.dst + c.dst_base * 8;
    for (int w = ctid; w < p.n_units; w += kMicroConsumers * 32) {
      const uint2 e = unitLut[w];
      if (c.edge) {
        const IdxEnt li = unitIdx[w];
        if (!((int)li.i0 < c.rem0 && (int)li.i1 < c.rem1)) continue;
      }
      const uint4 v0 = *reinterpret_cast<const uint4 *>(buf + e.x);               // (i, o), (i+1, o)
      const uint4 v1 = *reinterpret_cast<const uint4 *>(buf + e.x + p.so_bytes);  // (i, o+1), (i+1, o+1)
      unsigned char *g = dbase + (size_t)e.y * 8;
      *reinterpret_cast<uint4 *>(g) = make_uint4(v0.x, v0.y, v1.x, v1.y);                          // column i
      *reinterpret_cast<uint4 *>(g + p.di_elems * 8) = make_uint4(v0.z, v0.w, v1.z, v1.w);          // column i+1
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
  }
}


// =================================================================================================
// 8- and 16-byte elements, true transposition, tile-multiple extents: the data moves through the
// TMA in both directions, nothing through the LSU's global path.
//   i = source-fastest axis, o = destination-fastest axis; "run" = o followed by the axes that
//   continue it contiguously in the destination.  One tile = (128 B along i) x (RB <= 128 run
//   positions), 16 KB.
//   load  : ONE cp.async.bulk.tensor per tile, box [u][o][i] -> shared memory rows of 128 B whose
//           16-byte chunks are XOR-swizzled with the row number (CU_TENSOR_MAP_SWIZZLE_128B);
//   move  : 8 consumer warps: LDS.128 down the columns of the swizzled tile (2x2 register
//           micro-transpose for 8-byte elements), STS.128 along the rows of the output tile
//           [run/16][i][16 run positions], which is swizzled the same way.  The lane mapping below
//           makes both accesses bank-conflict free;
//   store : ONE cp.async.bulk.tensor (shared -> global) per tile, issued by one thread: full-line
//           writes independent of the thread mapping; bulk groups gate re-use of the 2 out stages.
// Producer warp + full/empty mbarriers as in the GEMMs; 3 input stages, 2 CTAs per SM.
// =================================================================================================
constexpr int kT8TileBytes = 16384;
constexpr int kT8MaxIn = 6;            // input stages (run-time choice, <= this)
constexpr int kT8Consumers = 8;
constexpr int kT8Threads = (kT8Consumers + 1) * 32;
constexpr int kT8MaxOd = 8;
inline int t8_smem(int in_stages, int out_stages) { return (in_stages + out_stages) * kT8TileBytes + 1024 + 512; }

struct Tma8Params {
  int32_t n_od;                     // tile-grid dims, fastest first
  int32_t od_ntiles[kT8MaxOd];
  int32_t od_prefix[kT8MaxOd];      // product of the tile counts of the faster dims
  int8_t ld_dim[kT8MaxOd], st_dim[kT8MaxOd];    // tensor-map coordinate each dim feeds (load / store map)
  int32_t ld_mul[kT8MaxOd], st_mul[kT8MaxOd];   // coordinate units per tile step
  int32_t total_tiles;
  int32_t tiles_per_cta;
  int32_t run_groups;               // RB / 16 (8-byte elements) or RB / 8 (16-byte elements)
  int32_t tile_bytes;               // 128 * RB
  int32_t in_stages;
  int32_t interleave;               // 1: CTA b takes tiles b, b + grid, ... (all CTAs sweep one compact window); 0: a contiguous range
};

// Tensor-map coordinates of a tile, computed by a whole warp: lane d owns tile-grid dim d (one
// divide + one modulo), the five load and five store coordinates are warp-wide integer reductions.
// (A single thread walking the dims with div/mod and coordinate arrays in local memory cost more
// than the tile's 16 KB of traffic: measured 4.4 vs 5.8 TB/s on an 8-dim transpose.)
struct T8Lane {
  int nt, pre, lmul, smul, ldim, sdim;
};
__device__ __forceinline__ T8Lane t8_lane_setup(const Tma8Params &p, int lane) {
  T8Lane L;
  const bool act = lane < p.n_od;
  const int d = act ? lane : 0;
  L.nt = act ? p.od_ntiles[d] : 1;
  L.pre = act ? p.od_prefix[d] : 1;
  L.lmul = act ? p.ld_mul[d] : 0;
  L.smul = act ? p.st_mul[d] : 0;
  L.ldim = act ? p.ld_dim[d] : -1;
  L.sdim = act ? p.st_dim[d] : -1;
  return L;
}
__device__ __forceinline__ void t8_coords(const T8Lane &L, int tile, int (&lc)[5], int (&sc)[5]) {
  const int digit = (tile / L.pre) % L.nt;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    lc[j] = __reduce_add_sync(0xffffffffu, L.ldim == j ? digit * L.lmul : 0);
    sc[j] = __reduce_add_sync(0xffffffffu, L.sdim == j ? digit * L.smul : 0);
  }
}

template <int ES, int OUTS>
__global__ void __launch_bounds__(kT8Threads) permute_tma_kernel(const __grid_constant__ CUtensorMap map_in,
                                                                  const __grid_constant__ CUtensorMap map_out,
                                                                  const __grid_constant__ Tma8Params p) {
  extern __shared__ unsigned char sm_raw[];
  unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  const int kT8InStages = p.in_stages;
  unsigned char *in_buf = sm;                                     // [in_stages][16 KB]
  unsigned char *out_buf = sm + kT8InStages * kT8TileBytes;        // [OUTS][16 KB]
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(out_buf + OUTS * kT8TileBytes);
  uint64_t *empty_bar = full_bar + kT8MaxIn;
  int *st_coords = reinterpret_cast<int *>(empty_bar + kT8MaxIn);   // [in_stages][8]: store coordinates of the tile in each stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int first, step, n_my;
  if (p.interleave) {
    first = blockIdx.x; step = gridDim.x;
    n_my = (p.total_tiles - first + step - 1) / step;
  } else {
    first = blockIdx.x * p.tiles_per_cta; step = 1;
    n_my = p.total_tiles - first;
    if (n_my > p.tiles_per_cta) n_my = p.tiles_per_cta;
  }
  if (n_my <= 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT8InStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kT8Consumers);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kT8Consumers) {
    // ===== producer warp: all lanes compute coordinates, lane 0 talks to the TMA =====
    if (lane == 0) tma_prefetch_desc(&map_in);
    const T8Lane L = t8_lane_setup(p, lane);
    for (int t = 0; t < n_my; ++t) {
      const int s = t % kT8InStages;
      int lc[5], sc[5];
      t8_coords(L, first + t * step, lc, sc);
      mbar_wait(&empty_bar[s], (uint32_t)(((t / kT8InStages) & 1) ^ 1));
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) st_coords[s * 8 + j] = sc[j];   // published by the mbarrier arrive below
        fence_proxy_async();
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.tile_bytes);
        tma_load_5d(in_buf + s * kT8TileBytes, &map_in, &full_bar[s], lc[0], lc[1], lc[2], lc[3], lc[4]);
      }
      __syncwarp();
    }
    return;
  }

  // ===== consumers =====
  if (threadIdx.x == 0) tma_prefetch_desc(&map_out);
  // lane -> (ip, op): a quarter warp holds op_lo = 0..3 and the two column pairs {a, a ^ 3}
  const int op = (lane & 3) | ((lane >> 1) & 4);
  const int xb = (lane >> 2) & 1, yb = (lane >> 4) & 1;
  const int ip = yb ? (xb ? 2 : 1) : (xb ? 3 : 0);
  const int n_iter = 2 * p.run_groups;   // warp-iterations per tile (8-byte elements)
  for (int t = 0; t < n_my; ++t) {
    const int s = t % kT8InStages;
    const int os = t % OUTS;
    // the bulk store that last read out stage `os` (OUTS tiles ago) must have finished reading it
    if (threadIdx.x == 0) bulk_wait_group_read<OUTS - 1>();
    asm volatile("bar.sync 1, %0;" ::"n"(kT8Consumers * 32) : "memory");
    mbar_wait(&full_bar[s], (uint32_t)((t / kT8InStages) & 1));
    const unsigned char *tin = in_buf + s * kT8TileBytes;
    unsigned char *tout = out_buf + os * kT8TileBytes;
    if (ES == 8) {
      // warp-iteration q: i-group ig (8 columns) x run group rg (16 run positions)
      for (int q = warp; q < n_iter; q += kT8Consumers) {
        const int ig = q & 1, rg = q >> 1;
        const int r = 16 * rg + 2 * op;                  // run position (= input row) of the even element
        const int chunk = 4 * ig + ip;                   // 16-byte column of the input row: i = 2 * chunk (+1)
        const uint4 v0 = *reinterpret_cast<const uint4 *>(tin + r * 128 + ((chunk ^ (r & 7)) << 4));              // (i, r) (i+1, r)
        const uint4 v1 = *reinterpret_cast<const uint4 *>(tin + (r + 1) * 128 + ((chunk ^ ((r + 1) & 7)) << 4));  // (i, r+1) (i+1, r+1)
        const int i0 = 2 * chunk;                        // output row inside run group rg
        unsigned char *orow = tout + (rg * 16 + i0) * 128;
        *reinterpret_cast<uint4 *>(orow + ((op ^ (i0 & 7)) << 4)) = make_uint4(v0.x, v0.y, v1.x, v1.y);
        *reinterpret_cast<uint4 *>(orow + 128 + ((op ^ ((i0 + 1) & 7)) << 4)) = make_uint4(v0.z, v0.w, v1.z, v1.w);
      }
    } else if (ES == 4) {
      // 4-byte elements: 4(i) x 4(run) register micro-transpose per lane.  Lane bits (p, u0, u1 | w0, w1):
      // i-quad ip = 2*u1 + u0 + 4*w0, run-quad op = 2*(u1 ^ flip) + p + 4*w1; the two passes (flip) cover the
      // 8 x 8 quads of a run group.  Inside a quarter warp (p, u0, u1) both the swizzled loads
      // (chunk ip ^ (4p + r)) and the swizzled stores (chunk op ^ (4*u0 + c)) hit 8 distinct 16-byte banks.
      const int pb = lane & 1, u0 = (lane >> 1) & 1, u1 = (lane >> 2) & 1, w0 = (lane >> 3) & 1, w1 = (lane >> 4) & 1;
      const int ip = 2 * u1 + u0 + 4 * w0;
      for (int q = warp; q < 2 * p.run_groups; q += kT8Consumers) {
        const int rg = q >> 1, flip = q & 1;
        const int op = 2 * (u1 ^ flip) + pb + 4 * w1;
        const int r0 = 32 * rg + 4 * op;
        uint4 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = r0 + r;
          v[r] = *reinterpret_cast<const uint4 *>(tin + row * 128 + ((ip ^ (row & 7)) << 4));
        }
        unsigned char *obase = tout + (rg * 32 + 4 * ip) * 128;
        *reinterpret_cast<uint4 *>(obase + ((op ^ ((4 * ip) & 7)) << 4)) = make_uint4(v[0].x, v[1].x, v[2].x, v[3].x);
        *reinterpret_cast<uint4 *>(obase + 128 + ((op ^ ((4 * ip + 1) & 7)) << 4)) = make_uint4(v[0].y, v[1].y, v[2].y, v[3].y);
        *reinterpret_cast<uint4 *>(obase + 256 + ((op ^ ((4 * ip + 2) & 7)) << 4)) = make_uint4(v[0].z, v[1].z, v[2].z, v[3].z);
        *reinterpret_cast<uint4 *>(obase + 384 + ((op ^ ((4 * ip + 3) & 7)) << 4)) = make_uint4(v[0].w, v[1].w, v[2].w, v[3].w);
      }
    } else {
      // 16-byte elements: input rows = run positions, 8 elements per row; output rows [run/8][i], 8 run positions each
      const int rl = lane & 7, il = lane >> 3;          // quarter warp: one i, 8 consecutive run positions
      for (int q = warp; q < 2 * p.run_groups; q += kT8Consumers) {
        const int rg = q >> 1;
        const int i_l = il + 4 * (q & 1);
        const int r = 8 * rg + rl;
        const uint4 v = *reinterpret_cast<const uint4 *>(tin + r * 128 + ((i_l ^ (r & 7)) << 4));
        *reinterpret_cast<uint4 *>(tout + (rg * 8 + i_l) * 128 + ((rl ^ (i_l & 7)) << 4)) = v;
      }
    }
    int sc[5];
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < 5; ++j) sc[j] = st_coords[s * 8 + j];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp is done reading the input stage (and its coordinates)
    fence_proxy_async();                          // generic-proxy writes of the out tile -> visible to the TMA
    asm volatile("bar.sync 1, %0;" ::"n"(kT8Consumers * 32) : "memory");
    if (threadIdx.x == 0) {
      tma_store_5d(&map_out, tout, sc[0], sc[1], sc[2], sc[3], sc[4]);
      bulk_commit_group();
    }
  }
  if (threadIdx.x == 0) bulk_wait_group<0>();
}

// =================================================================================================
// Copy-like moves (the fastest axis is the same on both sides: blockify / unblockify of BlockArrays,
// operands whose inner run survives the permutation): TMA load -> TMA store through shared memory,
// no thread ever touches the data.  Tile = R rows x up to 2 KB of the shared inner run, both tensor
// maps describe it as (inner 64-bit words, the row axis, rest...), so the shared-memory image of the
// load box IS the store box.  One warp per CTA: all lanes compute tile coordinates, lane 0 keeps
// kCpLook loads in flight and turns every landed tile into a bulk store.
// =================================================================================================
constexpr int kCpStages = 6;
constexpr int kCpLook = 3;
constexpr int kCpTileBytes = 16384;
constexpr int kCpSmem = kCpStages * kCpTileBytes + 1024 + 256;

__global__ void __launch_bounds__(32) permute_tmacopy_kernel(const __grid_constant__ CUtensorMap map_in,
                                                             const __grid_constant__ CUtensorMap map_out,
                                                             const __grid_constant__ Tma8Params p) {
  extern __shared__ unsigned char sm_raw[];
  unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sm + kCpStages * kCpTileBytes);
  const int lane = threadIdx.x;
  const int first = blockIdx.x, step = gridDim.x;
  const int n_my = (p.total_tiles - first + step - 1) / step;
  if (n_my <= 0) return;
  if (lane == 0) {
    for (int s = 0; s < kCpStages; ++s) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_in);
    tma_prefetch_desc(&map_out);
  }
  __syncwarp();
  const T8Lane L = t8_lane_setup(p, lane);
  // software pipeline: iteration t issues the load of tile t and the store of tile t - kCpLook
  for (int t = 0; t < n_my + kCpLook; ++t) {
    if (t < n_my) {
      int lc[5], sc[5];
      t8_coords(L, first + t * step, lc, sc);
      if (lane == 0) {
        const int s = t % kCpStages;
        // stage s was last read by the store of tile t - kCpStages: at most (kCpStages - kCpLook - 1) newer stores may be pending
        bulk_wait_group_read<kCpStages - kCpLook - 1>();
        fence_proxy_async();
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.tile_bytes);
        tma_load_5d(sm + s * kCpTileBytes, &map_in, &full_bar[s], lc[0], lc[1], lc[2], lc[3], lc[4]);
      }
    }
    const int u = t - kCpLook;
    if (u >= 0) {
      int lc[5], sc[5];
      t8_coords(L, first + u * step, lc, sc);
      if (lane == 0) {
        const int s = u % kCpStages;
        mbar_wait(&full_bar[s], (uint32_t)((u / kCpStages) & 1));
        tma_store_5d(&map_out, sm + s * kCpTileBytes, sc[0], sc[1], sc[2], sc[3], sc[4]);
        bulk_commit_group();
      }
    }
    __syncwarp();
  }
  if (lane == 0) bulk_wait_group<0>();
}

// ---- slow but fully general fallback (merged rank > MAXD or offsets beyond 32 bits) ----
struct GenericParams {
  const unsigned char *src;
  unsigned char *dst;
  int32_t es, nd;
  int64_t total;
  int64_t extent[RNB_MAX_RANK], ss[RNB_MAX_RANK], ds[RNB_MAX_RANK];  // ordered by destination stride, inner first
};
__global__ void permute_generic_kernel(const __grid_constant__ GenericParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, so = 0, dn = 0;
    for (int d = 0; d < p.nd; ++d) {
      const int64_t digit = rem % p.extent[d];
      rem /= p.extent[d];
      so += digit * p.ss[d];
      dn += digit * p.ds[d];
    }
    const unsigned char *s = p.src + so * p.es;
    unsigned char *t = p.dst + dn * p.es;
    if (p.es == 16) *reinterpret_cast<uint4 *>(t) = *reinterpret_cast<const uint4 *>(s);
    else if (p.es == 8) *reinterpret_cast<uint2 *>(t) = *reinterpret_cast<const uint2 *>(s);
    else *reinterpret_cast<uint32_t *>(t) = *reinterpret_cast<const uint32_t *>(s);
  }
}

struct Dim {
  int64_t extent, ss, ds;
};

int64_t pick_tile(int64_t need, int64_t extent) {
  if (extent <= need) return extent;
  for (int64_t c = need; c <= need + need / 2 && c <= extent; ++c)
    if (extent % c == 0) return c;
  return need;
}

uint32_t magic_for(uint32_t d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) / d) + 1); }

int launch_generic(const std::vector<Dim> &dims, int es, const void *src, void *dst, cudaStream_t stream) {
  GenericParams g;
  g.src = static_cast<const unsigned char *>(src);
  g.dst = static_cast<unsigned char *>(dst);
  g.es = es;
  std::vector<Dim> byd(dims);
  std::sort(byd.begin(), byd.end(), [](const Dim &a, const Dim &b) { return a.ds < b.ds; });
  g.nd = (int)byd.size();
  g.total = 1;
  for (int i = 0; i < g.nd; ++i) {
    g.extent[i] = byd[i].extent; g.ss[i] = byd[i].ss; g.ds[i] = byd[i].ds;
    g.total *= byd[i].extent;
  }
  int64_t blocks = (g.total + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  permute_generic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}


// Returns 0 = launched, 1 = not applicable (caller falls through to the other kernels), < 0 = -error code.
int try_launch_tma(const std::vector<Dim> &m, int es, const void *src, void *dst, cudaStream_t stream) {
  if (es != 4 && es != 8 && es != 16) return 1;
  const int nd = (int)m.size();
  if (nd < 2 || nd > kT8MaxOd) return 1;
  int di = -1, dq = -1;
  for (int d = 0; d < nd; ++d) {
    if (m[d].ss == 1) di = d;
    if (m[d].ds == 1) dq = d;
  }
  if (di < 0 || dq < 0 || di == dq) return 1;
  const int64_t TI = 128 / es;   // elements in one 128-byte swizzle row (tile extent along i, and along o inside an output row)
  const int64_t OL = TI;
  const int64_t Ei = m[di].extent, Eo = m[dq].extent;
  if (Ei % TI || Eo % OL) return 1;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return 1;
  for (int d = 0; d < nd; ++d) {
    if (m[d].extent >= (1ll << 31)) return 1;
    if (d != di && (m[d].ss * es) % 16) return 1;
    if (d != dq && (m[d].ds * es) % 16) return 1;
    if (m[d].ss * es >= (1ll << 40) || m[d].ds * es >= (1ll << 40)) return 1;
  }
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return 1;

  // ---- tile: TI along i, TO along o, BU along the axis that continues o in the destination ----
  int64_t TO = OL;
  for (int64_t c = 128; c >= OL; c -= OL)
    if (Eo % c == 0) { TO = c; break; }
  int du = -1;
  int64_t BU = 1;
  if (TO == Eo && TO < 128) {
    for (int d = 0; d < nd; ++d)
      if (d != di && d != dq && m[d].ds == Eo) du = d;
    if (du >= 0) {
      for (int64_t c = 128 / TO; c >= 1; --c)
        if (m[du].extent % c == 0) { BU = c; break; }
      if (BU == 1) du = -1;
    }
  }
  const int64_t RB = TO * BU;

  struct MapDim { int64_t extent, stride; };   // stride in elements
  int ld_of[MAXD], st_of[MAXD];
  int64_t ld_mul[MAXD], st_mul[MAXD], tile_ext[MAXD];
  for (int d = 0; d < MAXD; ++d) { ld_of[d] = st_of[d] = -1; ld_mul[d] = st_mul[d] = 0; tile_ext[d] = 1; }
  tile_ext[di] = TI; tile_ext[dq] = TO;
  if (du >= 0) tile_ext[du] = BU;

  // ---- load map: i | o | u | the remaining dims merged by SOURCE contiguity ----
  std::vector<MapDim> ld;
  ld.push_back({Ei, 1});               ld_of[di] = 0; ld_mul[di] = es == 16 ? 2 * TI : TI;   // 16-byte elements: coordinate 0 counts 64-bit words
  ld.push_back({Eo, m[dq].ss});        ld_of[dq] = 1; ld_mul[dq] = TO;
  if (du >= 0) { ld.push_back({m[du].extent, m[du].ss}); ld_of[du] = 2; ld_mul[du] = BU; }
  {
    std::vector<int> rest;
    for (int d = 0; d < nd; ++d) if (d != di && d != dq && d != du) rest.push_back(d);
    std::sort(rest.begin(), rest.end(), [&](int a, int b) { return m[a].ss < m[b].ss; });   // inner -> outer
    bool fresh = true;
    for (int d : rest) {
      MapDim &last = ld.back();
      if (!fresh && last.stride * last.extent == m[d].ss && last.extent * m[d].extent < (1ll << 31)) {
        ld_of[d] = (int)ld.size() - 1;
        ld_mul[d] = m[d].ss / last.stride;
        last.extent *= m[d].extent;
      } else {
        ld.push_back({m[d].extent, m[d].ss});
        ld_of[d] = (int)ld.size() - 1;
        ld_mul[d] = 1;
        fresh = false;
      }
    }
  }
  // ---- store map: 16 bytes x OL along o | i | run / OL | the remaining dims merged by DESTINATION contiguity ----
  std::vector<int> by_ds;
  for (int d = 0; d < nd; ++d) if (d != di && d != dq) by_ds.push_back(d);
  std::sort(by_ds.begin(), by_ds.end(), [&](int a, int b) { return m[a].ds < m[b].ds; });
  std::vector<MapDim> st;
  st.push_back({OL, 1});
  st.push_back({Ei, m[di].ds});        st_of[di] = 1; st_mul[di] = TI;
  int64_t run = Eo;                    st_of[dq] = 2; st_mul[dq] = TO / OL;
  size_t k3 = 0;
  for (; k3 < by_ds.size(); ++k3) {
    const int d = by_ds[k3];
    if (run != m[d].ds || run * m[d].extent >= (1ll << 31)) break;
    st_of[d] = 2;
    st_mul[d] = m[d].ds * tile_ext[d] / OL;
    run *= m[d].extent;
  }
  if (du >= 0 && st_of[du] != 2) return 1;
  st.push_back({run / OL, OL});
  {
    bool fresh = true;
    for (; k3 < by_ds.size(); ++k3) {
      const int d = by_ds[k3];
      MapDim &last = st.back();
      if (!fresh && last.stride * last.extent == m[d].ds && last.extent * m[d].extent < (1ll << 31)) {
        st_of[d] = (int)st.size() - 1;
        st_mul[d] = m[d].ds / last.stride;
        last.extent *= m[d].extent;
      } else {
        st.push_back({m[d].extent, m[d].ds});
        st_of[d] = (int)st.size() - 1;
        st_mul[d] = 1;
        fresh = false;
      }
    }
  }
  if (ld.size() > 5 || st.size() > 5) return 1;

  auto encode = [&](CUtensorMap *map, const void *base, const std::vector<MapDim> &dims, const cuuint32_t *box) {
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
    cuuint64_t gstr[4] = {16, 16, 16, 16};
    cuuint32_t bx[5] = {1, 1, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (size_t j = 0; j < dims.size(); ++j) {
      gdim[j] = (cuuint64_t)dims[j].extent;
      if (j > 0) gstr[j - 1] = (cuuint64_t)(dims[j].stride * es);
      bx[j] = box[j];
    }
    if (es == 16) {   // 16-byte elements are described as pairs of 64-bit words (no tensor-map type is that wide)
      gdim[0] *= 2;
      bx[0] *= 2;
    }
    return enc(map, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void *>(base), gdim, gstr, bx, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap map_in, map_out;
  cuuint32_t box_in[5] = {(cuuint32_t)TI, (cuuint32_t)TO, (cuuint32_t)(du >= 0 ? BU : 1), 1, 1};
  cuuint32_t box_out[5] = {(cuuint32_t)OL, (cuuint32_t)TI, (cuuint32_t)(RB / OL), 1, 1};
  if (encode(&map_in, src, ld, box_in) != CUDA_SUCCESS) return 1;
  if (encode(&map_out, dst, st, box_out) != CUDA_SUCCESS) return 1;

  Tma8Params p;
  memset(&p, 0, sizeof(p));
  // tile grid, fastest first: i tiles, o tiles, then the other dims by ascending destination stride
  std::vector<int> od;
  od.push_back(di);
  od.push_back(dq);
  for (int d : by_ds) od.push_back(d);
  int64_t total = 1;
  for (int d : od) {
    const int64_t nt = m[d].extent / tile_ext[d];
    if (nt == 1) continue;
    if (ld_mul[d] >= (1ll << 31) || st_mul[d] >= (1ll << 31)) return 1;
    const int j = p.n_od++;
    p.od_ntiles[j] = (int32_t)nt;
    p.od_prefix[j] = (int32_t)total;
    p.ld_dim[j] = (int8_t)ld_of[d]; p.ld_mul[j] = (int32_t)ld_mul[d];
    p.st_dim[j] = (int8_t)st_of[d]; p.st_mul[j] = (int32_t)st_mul[d];
    total *= nt;
    if (total >= (1ll << 31)) return 1;
  }
  for (int j = p.n_od; j < kT8MaxOd; ++j) { p.od_ntiles[j] = 1; p.od_prefix[j] = 1; p.ld_dim[j] = p.st_dim[j] = -1; }
  p.total_tiles = (int32_t)total;
  p.run_groups = (int32_t)(RB / OL);
  p.tile_bytes = (int32_t)(128 * RB);
  int in_stages = 2, out_stages = 2, ctas = 2, interleave = 1;   // measured best of the sweep in profiles/r1/permute_probe_tma_knobs2.txt
  if (const char *e = ([] { static EnvKnob k("RNB_TMA_IN"); return k.get(); }())) { const int v = atoi(e); if (v >= 1 && v <= kT8MaxIn) in_stages = v; }
  if (const char *e = ([] { static EnvKnob k("RNB_TMA_OUT"); return k.get(); }())) { const int v = atoi(e); if (v >= 2 && v <= 4) out_stages = v; }
  if (const char *e = ([] { static EnvKnob k("RNB_TMA_CTAS"); return k.get(); }())) { const int v = atoi(e); if (v >= 1 && v <= 4) ctas = v; }
  if (const char *e = ([] { static EnvKnob k("RNB_TMA_INTERLEAVE"); return k.get(); }())) interleave = atoi(e) != 0;
  const int smem = t8_smem(in_stages, out_stages);
  int per_sm = max_smem_optin() / (smem + 1024);
  if (per_sm < 1) return 1;
  if (per_sm > ctas) per_sm = ctas;
  p.in_stages = in_stages;
  p.interleave = interleave;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > total) grid = total;
  p.tiles_per_cta = (int32_t)((total + grid - 1) / grid);
  if (!interleave) grid = (total + p.tiles_per_cta - 1) / p.tiles_per_cta;
  static DeviceOnce attr_done[3][3];
  bool &done = attr_done[es == 4 ? 0 : (es == 8 ? 1 : 2)][out_stages - 2].flag();
  auto launch = [&](auto kern) -> int {
    if (!done) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()) != cudaSuccess) return 1;
      done = true;
    }
    kern<<<(unsigned)grid, kT8Threads, smem, stream>>>(map_in, map_out, p);
    return 0;
  };
  int lrc;
  if (es == 4) lrc = out_stages == 2 ? launch(permute_tma_kernel<4, 2>) : (out_stages == 3 ? launch(permute_tma_kernel<4, 3>) : launch(permute_tma_kernel<4, 4>));
  else if (es == 8) lrc = out_stages == 2 ? launch(permute_tma_kernel<8, 2>) : (out_stages == 3 ? launch(permute_tma_kernel<8, 3>) : launch(permute_tma_kernel<8, 4>));
  else lrc = out_stages == 2 ? launch(permute_tma_kernel<16, 2>) : (out_stages == 3 ? launch(permute_tma_kernel<16, 3>) : launch(permute_tma_kernel<16, 4>));
  if (lrc) return 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("permute_tma_kernel launch failed: %s", cudaGetErrorString(e));
    return -RNB_ERR_CUDA;
  }
  return 0;
}


// Copy-like form of try_launch_tma (same return convention).
int try_launch_tmacopy(const std::vector<Dim> &m, int es, const void *src, void *dst, cudaStream_t stream) {
  const int nd = (int)m.size();
  if (nd < 2 || nd > kT8MaxOd) return 1;
  int di = -1;
  for (int d = 0; d < nd; ++d)
    if (m[d].ss == 1 && m[d].ds == 1) di = d;
  if (di < 0) return 1;
  const int64_t run_bytes = m[di].extent * es;
  if (run_bytes < 256 || run_bytes % 16) return 1;          // short inner runs: the LUT-tiled kernel coalesces better
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return 1;
  for (int d = 0; d < nd; ++d) {
    if (m[d].extent >= (1ll << 31)) return 1;
    if (d != di && ((m[d].ss * es) % 16 || (m[d].ds * es) % 16)) return 1;
    if (m[d].ss * es >= (1ll << 40) || m[d].ds * es >= (1ll << 40)) return 1;
  }
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return 1;
  // inner box: 64-bit words, an even count (16 bytes) dividing the run, at most 256 words (2 KB)
  const int64_t words = run_bytes / 8;
  int64_t IB = 2;
  for (int64_t c = 256; c >= 2; c -= 2)
    if (words % c == 0) { IB = c; break; }
  if (IB * 8 < 256) return 1;
  // row axis: the remaining axis with the smallest destination stride
  std::vector<int> by_ds;
  for (int d = 0; d < nd; ++d) if (d != di) by_ds.push_back(d);
  std::sort(by_ds.begin(), by_ds.end(), [&](int a, int b) { return m[a].ds < m[b].ds; });
  const int dr = by_ds[0];
  int64_t R = 1;
  for (int64_t c = std::min<int64_t>(256, kCpTileBytes / (IB * 8)); c >= 1; --c)
    if (m[dr].extent % c == 0) { R = c; break; }

  struct MapDim { int64_t extent, stride; };   // stride in elements of the tensor (bytes = stride * es)
  int ld_of[MAXD], st_of[MAXD];
  int64_t ld_mul[MAXD], st_mul[MAXD];
  for (int d = 0; d < MAXD; ++d) { ld_of[d] = st_of[d] = -1; ld_mul[d] = st_mul[d] = 0; }
  auto build = [&](bool load, std::vector<MapDim> &out, int *of, int64_t *mul) {
    out.push_back({words, 0});                     // dim 0: 64-bit words of the inner run (stride implicit)
    of[di] = 0; mul[di] = IB;
    out.push_back({m[dr].extent, load ? m[dr].ss : m[dr].ds});
    of[dr] = 1; mul[dr] = R;
    std::vector<int> rest;
    for (int d = 0; d < nd; ++d) if (d != di && d != dr) rest.push_back(d);
    std::sort(rest.begin(), rest.end(), [&](int a, int b) { return load ? m[a].ss < m[b].ss : m[a].ds < m[b].ds; });
    bool fresh = true;
    for (int d : rest) {
      const int64_t st = load ? m[d].ss : m[d].ds;
      MapDim &last = out.back();
      if (!fresh && last.stride * last.extent == st && last.extent * m[d].extent < (1ll << 31)) {
        of[d] = (int)out.size() - 1;
        mul[d] = st / last.stride;
        last.extent *= m[d].extent;
      } else {
        out.push_back({m[d].extent, st});
        of[d] = (int)out.size() - 1;
        mul[d] = 1;
        fresh = false;
      }
    }
  };
  std::vector<MapDim> ld, st;
  build(true, ld, ld_of, ld_mul);
  build(false, st, st_of, st_mul);
  if (ld.size() > 5 || st.size() > 5) return 1;
  auto encode = [&](CUtensorMap *map, const void *base, const std::vector<MapDim> &dims) {
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
    cuuint64_t gstr[4] = {16, 16, 16, 16};
    cuuint32_t bx[5] = {(cuuint32_t)IB, (cuuint32_t)R, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (size_t j = 0; j < dims.size(); ++j) {
      gdim[j] = (cuuint64_t)dims[j].extent;
      if (j > 0) gstr[j - 1] = (cuuint64_t)(dims[j].stride * es);
    }
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void *>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap map_in, map_out;
  if (encode(&map_in, src, ld) != CUDA_SUCCESS) return 1;
  if (encode(&map_out, dst, st) != CUDA_SUCCESS) return 1;

  Tma8Params p;
  memset(&p, 0, sizeof(p));
  std::vector<int> od;
  od.push_back(di);
  for (int d : by_ds) od.push_back(d);
  int64_t total = 1;
  for (int d : od) {
    const int64_t nt = d == di ? words / IB : (d == dr ? m[d].extent / R : m[d].extent);
    if (nt == 1) continue;
    if (ld_mul[d] >= (1ll << 31) || st_mul[d] >= (1ll << 31)) return 1;
    const int j = p.n_od++;
    p.od_ntiles[j] = (int32_t)nt;
    p.od_prefix[j] = (int32_t)total;
    p.ld_dim[j] = (int8_t)ld_of[d]; p.ld_mul[j] = (int32_t)ld_mul[d];
    p.st_dim[j] = (int8_t)st_of[d]; p.st_mul[j] = (int32_t)st_mul[d];
    total *= nt;
    if (total >= (1ll << 31)) return 1;
  }
  for (int j = p.n_od; j < kT8MaxOd; ++j) { p.od_ntiles[j] = 1; p.od_prefix[j] = 1; p.ld_dim[j] = p.st_dim[j] = -1; }
  p.total_tiles = (int32_t)total;
  p.tile_bytes = (int32_t)(IB * 8 * R);
  int per_sm = max_smem_optin() / (kCpSmem + 1024);
  if (per_sm < 1) return 1;
  if (per_sm > 2) per_sm = 2;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > total) grid = total;
  static DeviceOnce once;
  bool &attr_done = once.flag();
  if (!attr_done) {
    if (cudaFuncSetAttribute(permute_tmacopy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCpSmem) != cudaSuccess) return 1;
    attr_done = true;
  }
  permute_tmacopy_kernel<<<(unsigned)grid, 32, kCpSmem, stream>>>(map_in, map_out, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("permute_tmacopy_kernel launch failed: %s", cudaGetErrorString(e));
    return -RNB_ERR_CUDA;
  }
  return 0;
}

template <int ES, int VS, int MODE = 0>
int launch_tiled(const PermParams &p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = permute_tiled_kernel<ES, VS, MODE>;
  static DeviceOnce once;
  bool &attr_done = once.flag();
  if (!attr_done) {
    RNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
    attr_done = true;
  }
  kern<<<grid, kThreadsP, smem, stream>>>(p);
  RNB_CHECK_CUDA(cudaGetLastError());
  return RNB_OK;
}

}  // namespace
}  // namespace rnb

namespace rnb {

int permute_impl(const rnb_permute_desc *desc, const SplitTarget *split, cudaStream_t stream) {
  RNB_REQUIRE(desc != nullptr, "descriptor is NULL");
  const int es = desc->elem_bytes;
  RNB_REQUIRE(es == 4 || es == 8 || es == 16, "elem_bytes must be 4, 8 or 16 (got %d)", es);
  RNB_REQUIRE(desc->rank > 0 && desc->rank <= RNB_MAX_RANK, "rank must be in 1..%d (got %d)", RNB_MAX_RANK, desc->rank);
  RNB_REQUIRE(desc->src && desc->dst, "src/dst must not be NULL");
  RNB_REQUIRE((reinterpret_cast<uintptr_t>(desc->src) % es) == 0 && (reinterpret_cast<uintptr_t>(desc->dst) % es) == 0,
              "src/dst must be aligned to the element size");

  std::vector<Dim> dims;
  for (int i = 0; i < desc->rank; ++i) {
    RNB_REQUIRE(desc->extent[i] >= 0, "negative extent");
    if (desc->extent[i] == 0) return RNB_OK;
    RNB_REQUIRE(desc->src_stride[i] >= 0 && desc->dst_stride[i] >= 0, "negative strides are not supported");
    if (desc->extent[i] > 1) dims.push_back({desc->extent[i], desc->src_stride[i], desc->dst_stride[i]});
  }
  if (dims.empty()) dims.push_back({1, 1, 1});
  for (const Dim &d : dims) RNB_REQUIRE(d.extent == 1 || (d.ss > 0 && d.ds > 0), "zero stride on an axis with extent > 1");

  // source order (outer -> inner), then merge axes adjacent in both layouts
  std::stable_sort(dims.begin(), dims.end(), [](const Dim &a, const Dim &b) { return a.ss > b.ss; });
  std::vector<Dim> m;
  for (const Dim &d : dims) {
    if (!m.empty()) {
      Dim &o = m.back();  // o is outer, d is inner
      if (o.ss == d.ss * d.extent && o.ds == d.ds * d.extent) {
        o.extent *= d.extent; o.ss = d.ss; o.ds = d.ds;
        continue;
      }
    }
    m.push_back(d);
  }
  const int nd = (int)m.size();

  const char *force_generic = ([] { static EnvKnob k("RNB_PERMUTE_KERNEL"); return k.get(); }());
  const bool use_generic = nd > MAXD || (force_generic && !strcmp(force_generic, "generic"));
  if (use_generic && split) {
    set_error("rnb_permute_split: the layout needs the generic kernel, which has no split epilogue");
    return RNB_ERR_UNSUPPORTED;
  }
  if (use_generic) return launch_generic(m, es, desc->src, desc->dst, stream);


  // ---- 8 / 16-byte elements, transposition with tile-multiple extents: all-TMA kernel ----
  if (!split) {
    const char *force = ([] { static EnvKnob k("RNB_PERMUTE_KERNEL"); return k.get(); }());
    const bool allowed = !(force && (!strcmp(force, "tiled") || !strcmp(force, "micro")));
    int rc = allowed ? try_launch_tma(m, es, desc->src, desc->dst, stream) : 1;
    if (rc <= 0) return rc == 0 ? RNB_OK : -rc;
    rc = allowed && nd >= 2 ? try_launch_tmacopy(m, es, desc->src, desc->dst, stream) : 1;
    if (rc <= 0) return rc == 0 ? RNB_OK : -rc;
  }

  // ---- tile selection ----
  std::vector<int64_t> t(nd, 1);
  int64_t target = (es == 16) ? 32 : 64;
  int64_t target_out = target;
  if (const char *e = ([] { static EnvKnob k("RNB_PERMUTE_TARGET_IN"); return k.get(); }())) { const int v = atoi(e); if (v >= 2) target = v; }
  if (const char *e = ([] { static EnvKnob k("RNB_PERMUTE_TARGET_OUT"); return k.get(); }())) { const int v = atoi(e); if (v >= 2) target_out = v; }
  std::vector<int> by_dst(nd);
  for (int i = 0; i < nd; ++i) by_dst[i] = i;
  std::stable_sort(by_dst.begin(), by_dst.end(), [&](int a, int b) { return m[a].ds < m[b].ds; });
  {
    int64_t cur = 1;
    for (int d = nd - 1; d >= 0 && cur < target; --d) {
      t[d] = pick_tile((target + cur - 1) / cur, m[d].extent);
      cur *= t[d];
    }
    cur = 1;
    for (int k = 0; k < nd && cur < target_out; ++k) {
      const int d = by_dst[k];
      t[d] = std::max<int64_t>(t[d], pick_tile((target_out + cur - 1) / cur, m[d].extent));
      cur *= t[d];
    }
  }
  auto volume = [&]() { int64_t v = 1; for (int d = 0; d < nd; ++d) v *= t[d]; return v; };
  int64_t vmax = (40 * 1024) / es;
  if (const char *e = ([] { static EnvKnob k("RNB_PERMUTE_VMAX_KB"); return k.get(); }())) { const int v = atoi(e); if (v >= 1) vmax = (int64_t)v * 1024 / es; }
  const int64_t vmin = std::min<int64_t>(vmax / 2, 2048);
  // shrink if the two passes multiplied up
  for (int guard = 0; volume() > vmax && guard < 64; ++guard) {
    int best = -1;
    for (int d = 0; d < nd; ++d)
      if (t[d] > 1 && (best < 0 || t[d] > t[best])) best = d;
    if (best < 0) break;
    t[best] = (t[best] + 1) / 2;
  }
  // grow along source-inner axes when both sets overlap (copy-like): bigger tiles, fewer of them
  for (int d = nd - 1; d >= 0 && volume() < vmin; --d) {
    if (t[d] >= m[d].extent) continue;
    const int64_t v = volume();
    int64_t want = t[d] * ((vmin + v - 1) / v);
    if (want > 512) want = 512;
    if (want >= m[d].extent) want = m[d].extent;
    else {
      int64_t c = want;
      while (c > t[d] && m[d].extent % c != 0) --c;
      if (c > t[d]) want = c; else want = (want / t[d]) * t[d];
    }
    if (want > t[d] && (v / t[d]) * want <= vmax) t[d] = want;
  }
  // at most two partially covered dims whose extent is not a multiple of the tile
  int slot_of[MAXD];
  int nslots = 0;
  for (int d = nd - 1; d >= 0; --d) {
    slot_of[d] = -1;
    if (t[d] < m[d].extent && m[d].extent % t[d] != 0) {
      if (nslots < 2) slot_of[d] = nslots++;
      else {
        int64_t c = t[d];
        while (c > 1 && m[d].extent % c != 0) --c;
        t[d] = c;
      }
    }
  }


  // ---- 8-byte elements, true transposition: warp-specialised micro-transpose kernel ----
  {
    const int di = nd - 1;         // source-fastest axis
    const int dq = by_dst[0];      // destination-fastest axis
    const char *force_old = ([] { static EnvKnob k("RNB_PERMUTE_KERNEL"); return k.get(); }());
    auto even_strides = [&]() {
      if ((reinterpret_cast<uintptr_t>(desc->src) & 15) || (reinterpret_cast<uintptr_t>(desc->dst) & 15)) return false;
      for (int d = 0; d < nd; ++d) {
        if (d != di && (m[d].ss & 1)) return false;
        if (d != dq && (m[d].ds & 1)) return false;
      }
      return true;
    };
    if (!split && es == 8 && nd >= 2 && di != dq && m[di].ss == 1 && m[dq].ds == 1 && t[di] % 2 == 0 && t[dq] % 2 == 0 &&
        m[di].extent % 2 == 0 && m[dq].extent % 2 == 0 && t[di] >= 16 && t[dq] >= 8 && even_strides() && !(force_old && !strcmp(force_old, "tiled"))) {
      MicroParams q;
      memset(&q, 0, sizeof(q));
      q.src = static_cast<const unsigned char *>(desc->src);
      q.dst = static_cast<unsigned char *>(desc->dst);
      int td_of[MAXD];
      int64_t dense = 8;
      int64_t max_soff = 0, max_doff = 0;
      for (int d = nd - 1; d >= 0; --d) {
        td_of[d] = -1;
        if (t[d] > 1) {
          td_of[d] = q.n_td;
          TileDim &e = q.td[q.n_td++];
          e.t = (int32_t)t[d]; e.ss = m[d].ss; e.ds = m[d].ds; e.slot = slot_of[d];
          e.smem_stride = (int32_t)dense;
          dense *= t[d];
          if (slot_of[d] >= 0) q.slot_t[slot_of[d]] = (int32_t)t[d];
          max_soff += (t[d] - 1) * m[d].ss;
          max_doff += (t[d] - 1) * m[d].ds;
        }
      }
      const int64_t V = dense / 8;
      q.tile_bytes = (int32_t)((dense + 127) & ~(int64_t)127);
      q.n_units = (int32_t)(V / 4);
      q.so_bytes = q.td[td_of[dq]].smem_stride;
      q.di_elems = m[di].ds;
      // lane-fastest digits: 8 i-pairs, then up to 32/8 o-pairs, then the rest
      auto largest_divisor_le = [](int64_t x, int64_t cap) { int64_t c = cap < x ? cap : x; while (c > 1 && x % c) --c; return c < 1 ? 1 : c; };
      const int64_t ip = t[di] / 2, op = t[dq] / 2;
      const int64_t i_lo = largest_divisor_le(ip, 8);
      const int64_t o_lo = largest_divisor_le(op, 32 / i_lo);
      auto push = [&](int tdi, int64_t radix, int64_t mul) {
        if (radix <= 1) return;
        q.dig_td[q.n_dig] = (int8_t)tdi; q.dig_radix[q.n_dig] = (int32_t)radix; q.dig_mul[q.n_dig] = (int32_t)mul; ++q.n_dig;
      };
      // lane-fastest: 8 i pairs (one contiguous 128 B shared-memory run per quarter warp), then
      // o pairs (contiguous destination bytes), then the remaining digits in source order
      push(td_of[di], i_lo, 2);
      push(td_of[dq], o_lo, 2);
      push(td_of[di], ip / i_lo, 2 * i_lo);
      push(td_of[dq], op / o_lo, 2 * o_lo);
      for (int d = nd - 1; d >= 0; --d)
        if (d != di && d != dq && t[d] > 1) push(td_of[d], t[d], 1);
      // contiguous source chunks
      int64_t run = 1;
      int d = nd - 1;
      q.chunk_slot = -1;
      for (; d >= 0; --d) {
        if (t[d] == 1 && m[d].extent > 1) break;       // tile covers a single index: run ends
        if (t[d] == 1) continue;
        if (m[d].ss != run) break;
        run *= t[d];
        if (t[d] != m[d].extent) {
          if (slot_of[d] >= 0) { q.chunk_slot = slot_of[d]; q.chunk_slot_inner_bytes = (int32_t)(run / t[d] * 8); }
          --d;
          break;
        }
      }
      q.chunk_bytes = (int32_t)(run * 8);
      q.n_chunks = (int32_t)(V / run);
      for (; d >= 0; --d)
        if (t[d] > 1) q.cdim_td[q.n_cdims++] = (int8_t)td_of[d];
      q.stages = 3;
      if (const char *e = ([] { static EnvKnob k("RNB_PERMUTE_STAGES"); return k.get(); }())) { const int v = atoi(e); if (v >= 1 && v <= kMaxStages) q.stages = v; }
      // odometer
      q.total_tiles = 1;
      const char *od_env = ([] { static EnvKnob k("RNB_PERMUTE_OD"); return k.get(); }());
      const bool od_dst = !(od_env && !strcmp(od_env, "src"));
      for (int kk = 0; kk < nd; ++kk) {
        const int dd = od_dst ? by_dst[kk] : nd - 1 - kk;
        const int64_t nt = (m[dd].extent + t[dd] - 1) / t[dd];
        if (nt > 1) {
          const int j = q.n_od++;
          q.od_ntiles[j] = (int32_t)nt; q.od_t[j] = (int32_t)t[dd]; q.od_slot[j] = slot_of[dd]; q.od_extent[j] = m[dd].extent;
          q.od_sstep[j] = t[dd] * m[dd].ss; q.od_dstep[j] = t[dd] * m[dd].ds;
        }
        q.total_tiles *= nt;
      }
      q.lut_offset = q.stages * q.tile_bytes;
      size_t lut = (size_t)q.n_units * (sizeof(uint2) + sizeof(IdxEnt)) + (size_t)q.n_chunks * sizeof(ChunkEnt);
      lut = (lut + 15) & ~(size_t)15;
      q.ctrl_offset = (int32_t)(q.lut_offset + lut);
      const size_t smem = (size_t)q.ctrl_offset + kMaxStages * sizeof(Ctrl) + 2 * kMaxStages * sizeof(uint64_t) + MAXD * sizeof(int32_t) + 16;
      const bool fits = V >= 4 && q.n_units <= 4096 && (q.chunk_bytes % 16) == 0 && (int)smem <= max_smem_optin() &&
                        max_soff < (1ll << 32) && max_doff < (1ll << 32) &&
                        (q.chunk_slot < 0 || (q.chunk_slot_inner_bytes % 16) == 0 || true);
      if (fits) {
        int per_sm = (int)((size_t)max_smem_optin() / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        int64_t grid = (int64_t)sm_count() * per_sm;
        if (grid > q.total_tiles) grid = q.total_tiles;
        q.tiles_per_cta = (int32_t)((q.total_tiles + grid - 1) / grid);
        grid = (q.total_tiles + q.tiles_per_cta - 1) / q.tiles_per_cta;
        static DeviceOnce once;
        bool &attr_done = once.flag();
        if (!attr_done) {
          RNB_CHECK_CUDA(cudaFuncSetAttribute(permute_micro8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
          attr_done = true;
        }
        permute_micro8_kernel<<<(unsigned)grid, kMicroThreads, smem, stream>>>(q);
        RNB_CHECK_CUDA(cudaGetLastError());
        return RNB_OK;
      }
    }
  }

  PermParams p;
  memset(&p, 0, sizeof(p));
  p.src = static_cast<const unsigned char *>(desc->src);
  p.dst = static_cast<unsigned char *>(desc->dst);
  p.es = es;
  p.slot_t[0] = p.slot_t[1] = 0;

  // tile dims (t > 1), remembered in source order inner -> outer
  int td_of[MAXD];
  std::vector<int> src_inner_order;
  for (int d = nd - 1; d >= 0; --d) {
    td_of[d] = -1;
    if (t[d] > 1) {
      td_of[d] = p.n_td;
      TileDim &q = p.td[p.n_td++];
      q.t = (int32_t)t[d]; q.ss = m[d].ss; q.ds = m[d].ds; q.slot = slot_of[d];
      if (slot_of[d] >= 0) p.slot_t[slot_of[d]] = (int32_t)t[d];
      src_inner_order.push_back(d);
    }
  }
  if (p.n_td == 0) {  // a single element per tile cannot happen unless the tensor is 1 element
    td_of[nd - 1] = 0;
    TileDim &q = p.td[p.n_td++];
    q.t = 1; q.ss = m[nd - 1].ss; q.ds = m[nd - 1].ds; q.slot = -1;
    src_inner_order.push_back(nd - 1);
  }
  // source groups
  int64_t prod = 1;
  size_t k = 0;
  for (; k < src_inner_order.size(); ++k) {
    const int d = src_inner_order[k];
    const int64_t tt = std::max<int64_t>(t[d], 1);
    if (k > 0 && prod * tt > 512) break;
    p.td[td_of[d]].smem_stride = (int32_t)(prod * es);
    p.g_slow[p.n_slow++] = (int8_t)td_of[d];
    prod *= tt;
  }
  p.s_low = (int32_t)prod;
  p.row_bytes = p.s_low * es;
  p.row_bytes_padded = ((p.row_bytes + 15) & ~15) | 16;
  prod = 1;
  for (; k < src_inner_order.size(); ++k) {
    const int d = src_inner_order[k];
    p.td[td_of[d]].smem_stride = (int32_t)(prod * p.row_bytes_padded);
    p.g_shigh[p.n_shigh++] = (int8_t)td_of[d];
    prod *= t[d];
  }
  p.s_high = (int32_t)prod;
  p.tile_bytes = (p.s_high * p.row_bytes_padded + 127) & ~127;
  // destination groups
  std::vector<int> dst_inner_order;
  for (int kk = 0; kk < nd; ++kk)
    if (td_of[by_dst[kk]] >= 0) dst_inner_order.push_back(by_dst[kk]);
  prod = 1;
  k = 0;
  for (; k < dst_inner_order.size(); ++k) {
    const int d = dst_inner_order[k];
    const int64_t tt = std::max<int64_t>(t[d], 1);
    if (k > 0 && prod * tt > 512) break;
    p.g_dlow[p.n_dlow++] = (int8_t)td_of[d];
    prod *= tt;
  }
  p.d_low = (int32_t)prod;
  prod = 1;
  for (; k < dst_inner_order.size(); ++k) {
    const int d = dst_inner_order[k];
    p.g_dhigh[p.n_dhigh++] = (int8_t)td_of[d];
    prod *= t[d];
  }
  p.d_high = (int32_t)prod;

  // ---- load granularity ----
  const int di_src = src_inner_order[0];  // innermost source tile dim
  auto aligned_all = [&](int64_t bytes, bool src_side) {
    const uintptr_t base = reinterpret_cast<uintptr_t>(src_side ? desc->src : desc->dst);
    if (base % bytes) return false;
    for (int d = 0; d < nd; ++d) {
      const int64_t s = src_side ? m[d].ss : m[d].ds;
      const bool inner = src_side ? (d == di_src) : (d == dst_inner_order[0]);
      if (inner) continue;
      if (m[d].extent > 1 && (s * es) % bytes) return false;
    }
    return true;
  };
  int G = es;
  for (int cand = 16; cand > es; cand >>= 1) {
    const int ue = cand / es;
    if (m[di_src].ss == 1 && t[di_src] % ue == 0 && m[di_src].extent % ue == 0 && aligned_all(cand, true)) {
      G = cand;
      break;
    }
  }
  p.unit_elems = G / es;
  p.units_per_row = p.s_low / p.unit_elems;
  p.units_per_row_magic = (int32_t)magic_for((uint32_t)p.units_per_row);
  // ---- store vector ----
  const int di_dst = dst_inner_order[0];
  int vs = 1;
  for (int cand = 16 / es; cand > 1; cand >>= 1) {
    if (m[di_dst].ds == 1 && t[di_dst] % cand == 0 && m[di_dst].extent % cand == 0 && aligned_all((int64_t)cand * es, false)) {
      vs = cand;
      break;
    }
  }
  p.vs = vs;
  p.dunits_per_row = p.d_low / vs;
  p.dunits_per_row_magic = (int32_t)magic_for((uint32_t)p.dunits_per_row);

  // ---- TMA bulk rows: the whole low group must be one contiguous, 16-byte aligned source run ----
  bool bulk = (p.row_bytes % 16 == 0) && aligned_all(16, true) && m[di_src].ss == 1;
  {
    int64_t run = 1;
    for (int j = 0; j < p.n_slow && bulk; ++j) {
      const int d = src_inner_order[j];
      if (m[d].ss != run) bulk = false;
      if (j + 1 < p.n_slow && t[d] != m[d].extent) bulk = false;  // inner dims of the run must be fully covered
      if (slot_of[d] >= 0) {
        if (j + 1 != p.n_slow) bulk = false;
        else {
          p.slot_inner_elems[slot_of[d]] = (int32_t)run;
          if ((run * es) % 16) bulk = false;  // every possible shortened row must stay a multiple of 16 B
        }
      }
      run *= t[d];
    }
  }
  const char *force = ([] { static EnvKnob k("RNB_PERMUTE_LOAD"); return k.get(); }());
  if (force && !strcmp(force, "cpasync")) bulk = false;
  if (!bulk) p.slot_inner_elems[0] = p.slot_inner_elems[1] = 0;
  p.load_mode = bulk ? 1 : 0;

  // ---- odometer over tiles ----
  p.total_tiles = 1;
  for (int d = nd - 1; d >= 0; --d) {
    const int64_t nt = (m[d].extent + t[d] - 1) / t[d];
    if (nt > 1) {
      const int j = p.n_od++;
      RNB_REQUIRE(nt < (1ll << 31), "too many tiles along one axis");
      p.od_ntiles[j] = (int32_t)nt;
      p.od_t[j] = (int32_t)t[d];
      p.od_slot[j] = slot_of[d];
      p.od_extent[j] = m[d].extent;
      p.od_sstep[j] = t[d] * m[d].ss;
      p.od_dstep[j] = t[d] * m[d].ds;
    }
    p.total_tiles *= nt;
  }
  // a partial dim with a single tile never appears in the odometer: extent % t != 0 implies nt > 1. ok.

  // low-group offsets that are just the running index need no table
  {
    int64_t run = 1;
    bool contig = true;
    for (int j = 0; j < p.n_slow; ++j) {
      const TileDim &q = p.td[p.g_slow[j]];
      if (q.ss != run) contig = false;
      run *= q.t;
    }
    p.s_low_contig = contig;
    run = 1;
    contig = true;
    for (int j = 0; j < p.n_dlow; ++j) {
      const TileDim &q = p.td[p.g_dlow[j]];
      if (q.ds != run) contig = false;
      run *= q.t;
    }
    p.d_low_contig = contig;
  }
  // in-tile offsets are 32-bit, positions inside the low group 16-bit
  int64_t max_soff = 0, max_doff = 0;
  for (int j = 0; j < p.n_td; ++j) {
    max_soff += (int64_t)(p.td[j].t - 1) * p.td[j].ss;
    max_doff += (int64_t)(p.td[j].t - 1) * p.td[j].ds;
  }
  size_t lut_bytes = (size_t)(p.s_high + p.d_high) * sizeof(HighEnt) +
                     (size_t)(p.units_per_row + p.dunits_per_row) * (sizeof(uint32_t) + sizeof(IdxEnt)) +
                     (size_t)p.d_low * sizeof(uint16_t);
  lut_bytes = (lut_bytes + 15) & ~(size_t)15;
  p.ctrl_offset = (int32_t)(2 * (size_t)p.tile_bytes + lut_bytes);
  const size_t smem = (size_t)p.ctrl_offset + 2 * sizeof(Ctrl) + 2 * sizeof(uint64_t) + MAXD * sizeof(int32_t) + 16;
  if (p.s_low > kMaxLut || p.s_high > kMaxLut || p.d_low > kMaxLut || p.d_high > kMaxLut || (int)smem > max_smem_optin() ||
      max_soff >= (1ll << 32) || max_doff >= (1ll << 32) || (int64_t)p.row_bytes_padded >= 65536) {
    if (split) {
      set_error("rnb_permute_split: the layout does not fit the tiled kernel");
      return RNB_ERR_UNSUPPORTED;
    }
    return launch_generic(m, es, desc->src, desc->dst, stream);
  }

  int per_sm = (int)((size_t)max_smem_optin() / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  p.tiles_per_cta = (int32_t)((p.total_tiles + grid - 1) / grid);
  grid = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;

  if (split) {
    if (vs * es != 16) {
      set_error("rnb_permute_split: the destination run does not allow 16-byte store vectors");
      return RNB_ERR_UNSUPPORTED;
    }
    p.split_hi = split->hi;
    p.split_lo = split->lo;
    p.split_k = split->k;
    p.split_rk = split->rows * split->k;
    p.split_k_pow2 = (split->k & (split->k - 1)) == 0;
    p.split_multi = split->nblocks > 1;
    if (es == 4 && split->mode == RNB_SPLIT_PLANES) return launch_tiled<4, 4, 1>(p, (int)grid, smem, stream);
    if (es == 8 && split->mode == RNB_SPLIT_PLANES) return launch_tiled<8, 2, 1>(p, (int)grid, smem, stream);
    if (es == 8 && split->mode == RNB_SPLIT_C64_B4M) return launch_tiled<8, 2, 2>(p, (int)grid, smem, stream);
    if (es == 8 && split->mode == RNB_SPLIT_C64_3M) return launch_tiled<8, 2, 3>(p, (int)grid, smem, stream);
    set_error("rnb_permute_split: mode %d does not go with %d-byte elements", split->mode, es);
    return RNB_ERR_INVALID;
  }
  if (es == 4) {
    if (vs == 4) return launch_tiled<4, 4>(p, (int)grid, smem, stream);
    if (vs == 2) return launch_tiled<4, 2>(p, (int)grid, smem, stream);
    return launch_tiled<4, 1>(p, (int)grid, smem, stream);
  }
  if (es == 8) {
    if (vs == 2) return launch_tiled<8, 2>(p, (int)grid, smem, stream);
    return launch_tiled<8, 1>(p, (int)grid, smem, stream);
  }
  return launch_tiled<16, 1>(p, (int)grid, smem, stream);
}

}  // namespace rnb

using namespace rnb;

extern "C" int rnb_permute(const rnb_permute_desc *desc, rnb_stream_t stream_) {
  clear_error();
  return permute_impl(desc, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}
