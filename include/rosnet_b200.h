/*
 * rosnet_b200.h - C ABI of the B200-native blocked-tensor contraction library.
 *
 * The reference (bsc-quantic/rosnet v0.3.0) is pure Python and defines NO C-ABI / FFI for this
 * path: its boundary is the NumPy protocol pair __array_function__ / __array__ on BlockArray
 * (rosnet/core/mixin.py:17-41, rosnet/array/block/__init__.py:194-196) plus by-name module
 * attributes (rosnet/__init__.py:3-5).  The entry points below are what a reference-side binding
 * for the hot path binds instead of its NumPy calls; each one names the reference step it
 * replaces.  INTEGRATION.md shows the ctypes stub a rosnet maintainer would add.
 *
 * Conventions
 *  - every function returns an int status (RNB_OK == 0) and never throws; the message for the
 *    last non-zero status on the calling thread is rnb_last_error_string();
 *  - all data pointers are DEVICE pointers owned by the caller unless a parameter says "host";
 *  - every compute call takes a cudaStream_t (as void*) and is asynchronous w.r.t. the host;
 *  - extents / strides / leading dimensions are in ELEMENTS of the operand dtype
 *    (a complex element is one element);
 *  - there is no CPU fallback: without a CUDA device every compute call returns RNB_ERR_CUDA.
 */
#ifndef ROSNET_B200_H
#define ROSNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNB_ABI_VERSION 1
#define RNB_MAX_RANK 32

typedef void *rnb_stream_t; /* cudaStream_t */

/* status codes */
enum {
  RNB_OK = 0,
  RNB_ERR_INVALID = 1,     /* bad descriptor (rank, dtype, alignment, null pointer ...) */
  RNB_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed; see rnb_last_error_string() */
  RNB_ERR_UNSUPPORTED = 3, /* valid request the library does not implement */
  RNB_ERR_WORKSPACE = 4,   /* workspace missing or too small */
  RNB_ERR_NCCL = 5         /* NCCL not loadable or a NCCL call failed */
};

/* element types: what numpy.result_type gives for the reference's blocks */
enum { RNB_F32 = 0, RNB_F64 = 1, RNB_C64 = 2, RNB_C128 = 3 };

/* operand storage order of one block seen as a matrix for the GEMM
 *   RNB_MAJOR_K : [rows][K], K contiguous   (a with contracted axes last / b with them last)
 *   RNB_MAJOR_MN: [K][rows], rows contiguous (a with contracted axes first / b with them first) */
enum { RNB_MAJOR_K = 0, RNB_MAJOR_MN = 1 };

/* arithmetic selection
 *   RNB_PREC_DEFAULT: f64/c128 -> FP64 DMMA (mma.sync m8n8k4);  f32/c64 -> error-compensated
 *                     3xTF32 on tcgen05 with fp32 TMEM accumulators
 *   RNB_PREC_TF32X1 : f32/c64 only, single TF32 product (does NOT meet the 1e-5 parity bar;
 *                     for roofline experiments)
 *   RNB_PREC_TF32_BF16X2: f32/c64 only, opt-in: the hi.hi product in TF32, the two cross products
 *                     (each 2^-11 of the result) as BF16 MMAs at twice the TF32 rate: 2 instead of
 *                     3 TF32-equivalents of tensor work, ~1e-6 relative error */
enum { RNB_PREC_DEFAULT = 0, RNB_PREC_TF32X1 = 1, RNB_PREC_TF32_BF16X2 = 2 };

/* complex product decomposition into real GEMMs: 4M (4 real products) or 3M (3 products:
 * T1 = re.re, T2 = im.im, T3 = (re+im).(re+im); Re = T1 - T2, Im = T3 - T1 - T2).
 *   RNB_CPLX_AUTO: the library picks per call - complex128: 3M when n_pairs*k >= 64 (the three
 *                  products stay in registers, no extra pass); complex64: 3M when the summed
 *                  contracted extent n_pairs*k >= 1024 and m, n >= 128 (below that the extra pass
 *                  over the product workspace costs more than the saved quarter of tensor work). */
enum { RNB_CPLX_4M = 0, RNB_CPLX_3M = 1, RNB_CPLX_AUTO = 2 };

/* ---- library / device ------------------------------------------------------------------ */

int rnb_abi_version(void);
/* Experiment knobs are read from RNB_* environment variables once and cached; call this after
 * changing the environment of a running process to make the library read them again. */
int rnb_reload_env(void);
/* message for the last failing call on this thread ("" if none). Never NULL. */
const char *rnb_last_error_string(void);

typedef struct {
  int32_t sm_count;
  int32_t cc_major, cc_minor;
  int32_t max_smem_per_block; /* opt-in bytes */
  int64_t total_mem_bytes;
  int32_t l2_bytes_mb;
  int32_t reserved;
} rnb_device_props;
int rnb_get_device_props(int device, rnb_device_props *out);

/* ---- (i) permute / matricize -------------------------------------------------------------
 * Strided n-d copy dst[sum_i idx_i*dst_stride_i] = src[sum_i idx_i*src_stride_i] over
 * extent[0..rank).  Replaces `a.transpose(newaxes).reshape(M, K)` inside numpy.tensordot
 * (the per-pair call at rosnet/array/block/__init__.py:283) and the per-block
 * autoray.do("transpose") of rosnet/array/block/__init__.py:272-273; batched over all blocks of
 * a BlockArray by making the block index the leading axis.  Pure data movement: bit-exact.
 * HBM-bound; algorithmic bytes = 2 * elem_bytes * prod(extent). */
typedef struct {
  int32_t elem_bytes; /* 4, 8 or 16 */
  int32_t rank;       /* 0 < rank <= RNB_MAX_RANK */
  int64_t extent[RNB_MAX_RANK];
  int64_t src_stride[RNB_MAX_RANK]; /* elements */
  int64_t dst_stride[RNB_MAX_RANK]; /* elements */
  const void *src;
  void *dst;
} rnb_permute_desc;
int rnb_permute(const rnb_permute_desc *desc, rnb_stream_t stream);

/* ---- (i-b) matricize fused with the split pre-pass of the single-precision GEMM ---------------
 * float32 / complex64 operands are not multiplied as they are: rnb_contract_grouped first writes
 * them as "hi" and "lo" TF32 planes (3xTF32 error compensation).  For an operand that is NOT in GEMM
 * layout that used to be two passes over HBM - rnb_permute into a K-major workspace, then the split.
 * rnb_permute_split does both: `desc` describes the matricize copy exactly as for rnb_permute (desc->dst
 * is ignored), the kernel's store phase converts every element and writes the planes of the
 * contraction's workspace where rnb_contract_planes says they live.  The contraction is then run with
 * RNB_FLAG_A_PRESPLIT / RNB_FLAG_B_PRESPLIT set.  Returns RNB_ERR_UNSUPPORTED (nothing launched) for
 * layouts the fused kernel cannot take (no 16-byte destination runs, more than 32 merged axes): use
 * rnb_permute and an unflagged contraction then. */
enum { RNB_SPLIT_NONE = 0,
       RNB_SPLIT_PLANES = 1,   /* float32 operand, or complex64 `a` of the 4M arrangement: same offsets as the matricized operand */
       RNB_SPLIT_C64_B4M = 2,  /* complex64 `b` of the 4M arrangement: real [2n][2k] rows (re,-im) / (im,re) */
       RNB_SPLIT_C64_3M = 3 }; /* complex64 operand of the 3M arrangement: three real blocks re / im / re+im per block */
typedef struct {
  int32_t mode;        /* RNB_SPLIT_* for operand a and for operand b; RNB_SPLIT_NONE: this contraction cannot take pre-split operands */
  int32_t reserved;
  size_t hi_offset;    /* byte offsets of the hi / lo planes inside the contraction workspace */
  size_t lo_offset;
} rnb_operand_planes;

/* ---- (ii) grouped block contraction with fused k-block sum --------------------------------
 * For every output block g in [0, n_out):
 *     C[out_index[g]] (+)= sum_{p < n_pairs} A[pair_a[g*n_pairs+p]] (m x k) . B[pair_b[...]] (k x n)
 * with the whole pair list accumulated inside the kernel (one accumulator chain; no per-pair
 * round trip, no separate add pass).  Replaces the Python double loop over output blocks and
 * `sum(np.tensordot(ai, bi, axes) for ai, bi in zip(a, b))`
 * (rosnet/array/block/__init__.py:307-334 and :281-283), i.e. also the COMPSs task bodies
 * rosnet/array/compss/task/tensordot.py:17-21 (sequential) and :45-49 (commutative).
 * Block j of an operand starts at base + j*block_stride elements; row r of a block at
 * + r*ld.  C blocks are row-major m x n (a's free axes then b's free axes).
 * Algorithmic flops = n_out*n_pairs * (2 | 8 for complex) * m*n*k. */
enum { RNB_FLAG_A_PRESPLIT = 1, RNB_FLAG_B_PRESPLIT = 2 };
typedef struct {
  int32_t dtype;        /* RNB_F32 .. RNB_C128 (both operands and the result) */
  int32_t precision;    /* RNB_PREC_* */
  int32_t complex_mode; /* RNB_CPLX_* (ignored for real dtypes) */
  int32_t accumulate;   /* 0: C = sum, 1: C += sum (MaybeArray / commutative semantics) */
  int64_t m, n, k;      /* per block-pair GEMM extents */
  const void *a;
  int64_t a_ld, a_block_stride;
  int32_t a_major, a_nblocks;
  const void *b;
  int64_t b_ld, b_block_stride;
  int32_t b_major, b_nblocks;
  void *c;
  int64_t c_ld, c_block_stride;
  int32_t c_nblocks;
  int32_t n_out;            /* number of output blocks computed by this call */
  int32_t n_pairs;          /* contracted block pairs per output block (>= 1) */
  int32_t flags;            /* RNB_FLAG_*: operands whose split planes rnb_permute_split already wrote into `workspace` */
  const int32_t *pair_a;    /* device, [n_out*n_pairs] block numbers into a */
  const int32_t *pair_b;    /* device, [n_out*n_pairs] block numbers into b */
  const int32_t *out_index; /* device, [n_out] block numbers into c; NULL = 0..n_out-1 */
  void *workspace;          /* device scratch, rnb_contract_workspace_bytes() bytes (may be NULL if 0) */
  size_t workspace_bytes;
  /* Fused partial-sum exchange (float64/complex128 only; NULL / 0 = off).  When the contracted block
   * index is split across GPUs every rank holds a partial of EVERY output block.  With c_peers set, the
   * GEMM epilogue does not store to `c`: output block g is ADDED over NVLink peer memory (one
   * cp.reduce.async.bulk ... add.f64 per 512-byte tile row; system-scope red.add.f64 for ragged edge
   * tiles) into rank (g / blocks_per_peer)'s buffer c_peers[g / blocks_per_peer] at block
   * (g % blocks_per_peer).  The owners' buffers must be zeroed before and all ranks synchronised after. */
  void *const *c_peers;     /* host array of n_peers device pointers (peer-mapped, see rnb_ipc_*) */
  int32_t n_peers;
  int32_t blocks_per_peer;
} rnb_contract_desc;
int rnb_contract_workspace_bytes(const rnb_contract_desc *desc, size_t *bytes);
/* Where the split planes of `a` and `b` live inside the workspace of THIS descriptor and which
 * RNB_SPLIT_* mode writes them (float32 / complex64 on the tensor-core route; mode RNB_SPLIT_NONE otherwise). */
int rnb_contract_planes(const rnb_contract_desc *desc, rnb_operand_planes *a, rnb_operand_planes *b);
/* `workspace`: the contraction's workspace; `planes`: this operand's entry from rnb_contract_planes; `rows`, `k`,
 * `nblocks`: the operand's matricized block (m or n rows, k columns, elements of the dtype) and block count. */
int rnb_permute_split(const rnb_permute_desc *desc, const rnb_operand_planes *planes, void *workspace,
                      int64_t rows, int64_t k, int32_t nblocks, rnb_stream_t stream);
int rnb_contract_grouped(const rnb_contract_desc *desc, rnb_stream_t stream);
/* ---- (ii-b) small contractions on operands left in their own layout --------------------------
 * The first few hundred steps of a tensor-network path are contractions of a few thousand
 * multiply-adds whose operands are NOT in GEMM layout; matricizing them costs two permute launches
 * per step (numpy.tensordot's transpose + reshape, rosnet/array/block/__init__.py:281-283 ->
 * numpy/_core/numeric.py tensordot).  rnb_contract_nd takes the block axes as they are:
 *   C[g][i][j] (+)= sum over pairs p, k of  A[pair_a[g][p]][i, k] * B[pair_b[g][p]][j, k]
 * where i runs row-major over a's free axes (ext_m, a_stride_m; in output order), j over b's free
 * axes (ext_n, b_stride_n) and k over the contracted axes (ext_k with a's and b's strides), all
 * strides in elements; C blocks are written row-major [prod ext_m][prod ext_n].  One launch.  Two kernels:
 *   small : one thread per output element, float64 accumulation.  Limits: prod(ext_k) <= 512,
 *           n_out * prod(ext_m) * prod(ext_n) <= 2^24 output elements.
 *   apply : ONE operand tiny, the other one large ("a gate applied to a state tensor": the steps of a
 *           contraction tree that keeps its intermediates narrow, tn/path.py search_path): taken when no offset
 *           tables are passed, min(m, n) <= 16, prod(ext_k) <= 64, n_pairs * pad(min(m, n)) * k <= 1024 (pad: next
 *           of 2, 4, 8, 16), n_out <= 65535, the large operand < 2^31 elements with strides < 2^30, and the
 *           contraction is beyond the small kernel's range (> 2^20 outputs or > 2^22 multiply-adds).  One thread per
 *           free index of the large operand: k strided loads, the tiny operand from shared memory, min(m, n) coalesced
 *           stores; accumulation in the operand's own precision (k <= 64 terms).  HBM-bound: algorithmic bytes =
 *           elem_bytes * n_out * (n_pairs * (m*k + n*k) + m*n), i.e. the large operand read once and C written once -
 *           where rnb_contract_grouped would add a matricize + split pass over the large operand (1 read, 3 writes),
 *           a GEMM that is > 90 % tile padding and a combine pass.  RNB_APPLY=0 disables it. */
typedef struct {
  int32_t dtype;       /* RNB_F32 .. RNB_C128 */
  int32_t accumulate;  /* 0: C = sum, 1: C += sum */
  int32_t rank_m, rank_n, rank_k;
  int32_t n_out, n_pairs, reserved;
  int64_t ext_m[RNB_MAX_RANK], a_stride_m[RNB_MAX_RANK];
  int64_t ext_n[RNB_MAX_RANK], b_stride_n[RNB_MAX_RANK];
  int64_t ext_k[RNB_MAX_RANK], a_stride_k[RNB_MAX_RANK], b_stride_k[RNB_MAX_RANK];
  const void *a;
  const void *b;
  void *c;
  int64_t a_block_stride, b_block_stride, c_block_stride; /* elements */
  const int32_t *pair_a; /* device, [n_out][n_pairs] */
  const int32_t *pair_b;
  const int32_t *out_index; /* device, [n_out] or NULL */
  /* Optional precomputed offset tables (device int32; all four or none): offset of a's element for every free index
   * i (prod ext_m entries), of b's for every j (prod ext_n), and of a's / b's for every contracted index (prod ext_k
   * each).  A caller that issues the same contraction many times (every slice of a tensor network) builds them once;
   * the kernel then does no index arithmetic at all. */
  const int32_t *tab_m, *tab_n, *tab_ka, *tab_kb;
} rnb_contract_nd_desc;
int rnb_contract_nd(const rnb_contract_nd_desc *desc, rnb_stream_t stream);
/* `n` INDEPENDENT small contractions (no task reads what another task of the same call writes) in as few launches as
 * possible: the tasks travel in the kernel parameter block (340 per launch), so a tensor-network path whose first
 * few hundred steps are microseconds of work each pays one launch per dependency level instead of one per step.
 * Every descriptor must carry the four offset tables, back to back in ONE device array (tab_m, then tab_n, tab_ka,
 * tab_kb), and out_index == NULL. */
int rnb_contract_nd_batch(const rnb_contract_nd_desc *descs, int32_t n, rnb_stream_t stream);

/* Which kernel family rnb_contract_grouped uses for this descriptor (for profiling / reporting):
 *   RNB_ROUTE_TENSOR: the grouped tensor-core GEMM (DMMA or 3xTF32 tcgen05), split-K when few tiles;
 *   RNB_ROUTE_SKINNY: m, n <= 4 - HBM-bound SIMT reduction over the contracted extent (the final
 *                     inner products of a tensor-network path), float64 accumulation;
 *   RNB_ROUTE_SMALL : a few thousand multiply-adds - one SIMT launch, no pre-pass. */
enum { RNB_ROUTE_TENSOR = 0, RNB_ROUTE_SKINNY = 1, RNB_ROUTE_SMALL = 2 };
int rnb_contract_route(const rnb_contract_desc *desc, int32_t *route);

/* ---- block sum (stand-alone form of the reduction) -----------------------------------------
 * dst[i] = (accumulate ? dst[i] : 0) + sum_{p<n_src} src[p][i] over count elements.
 * Replaces `res += ...` of rosnet/array/compss/task/tensordot.py:45-49 when partial results
 * already exist as separate buffers (e.g. partials received from other GPUs).  HBM-bound. */
int rnb_block_sum(void *dst, const void *const *src_host_ptrs, int32_t n_src, int64_t count,
                  int32_t dtype, int32_t accumulate, rnb_stream_t stream);

/* ---- host <-> device staging (H2D ingest / to_numpy D2H of rosnet/array/block/__init__.py:203-205) */
int rnb_memcpy_h2d_async(void *dst_device, const void *src_host, size_t bytes, rnb_stream_t stream);
int rnb_memcpy_d2h_async(void *dst_host, const void *src_device, size_t bytes, rnb_stream_t stream);
/* device -> device on one GPU or between GPUs of the box (either pointer may be an rnb_ipc_open_handle mapping) */
int rnb_memcpy_d2d_async(void *dst_device, const void *src_device, size_t bytes, rnb_stream_t stream);
int rnb_stream_synchronize(rnb_stream_t stream);

/* ---- peer memory for the fused exchange ------------------------------------------------------------
 * Plain cudaMalloc'ed buffers that can be exported to the other ranks of the box with CUDA IPC, so a
 * kernel on one GPU can reduce into a buffer owned by another process over NVLink / NVSwitch. */
#define RNB_MAX_PEERS 8
#define RNB_IPC_HANDLE_BYTES 64
int rnb_device_malloc(size_t bytes, void **device_ptr);
int rnb_device_free(void *device_ptr);
int rnb_memset_async(void *device_ptr, int value, size_t bytes, rnb_stream_t stream);
int rnb_ipc_get_handle(void *device_ptr, char handle_host[RNB_IPC_HANDLE_BYTES]);
int rnb_ipc_open_handle(const char handle_host[RNB_IPC_HANDLE_BYTES], void **device_ptr);
int rnb_ipc_close_handle(void *device_ptr);

/* ---- (iii) partial-sum exchange when the contracted block index is split across GPUs ------
 * Thin wrappers over NCCL (loaded with dlopen at first use; no link-time dependency).
 * rnb_reduce_scatter_sum: every rank passes n_ranks*recv_count elements; rank r receives the
 * sum over ranks of chunk r.  Complex dtypes are reduced as 2x real lanes. */
#define RNB_NCCL_UNIQUE_ID_BYTES 128
int rnb_nccl_get_unique_id(char id_host[RNB_NCCL_UNIQUE_ID_BYTES]);
int rnb_nccl_comm_init_rank(void **comm, int n_ranks, const char id_host[RNB_NCCL_UNIQUE_ID_BYTES], int rank);
int rnb_nccl_comm_destroy(void *comm);
int rnb_reduce_scatter_sum(const void *send, void *recv, int64_t recv_count, int32_t dtype,
                           void *comm, rnb_stream_t stream);
int rnb_all_reduce_sum(const void *send, void *recv, int64_t count, int32_t dtype, void *comm,
                       rnb_stream_t stream);
/* rnb_all_gather: every rank passes send_count elements and receives n_ranks*send_count, rank order
 * (assembling the output blocks the ranks own into the whole BlockArray on every rank). */
int rnb_all_gather(const void *send, void *recv, int64_t send_count, int32_t dtype, void *comm,
                   rnb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ROSNET_B200_H */
