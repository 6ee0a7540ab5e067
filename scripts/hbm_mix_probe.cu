// What does HBM deliver for the byte mix of the fused matricize + split pass (read 8 B, write 24 B per complex64 element)?
// Streaming kernels without any permutation: copy (1R:1W), write-only, and 1R:3W with plain / streaming / bulk (TMA) stores.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/hbm_mix_probe scripts/hbm_mix_probe.cu && gpurun_out/hbm_mix_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) copy_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) out[i] = __ldg(in + i);
}

__global__ void __launch_bounds__(256) write_kernel(float4 *__restrict__ out, int64_t nvec) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}

template <int STORE>
__device__ __forceinline__ void st4(float4 *p, float4 v) {
  if constexpr (STORE == 0) *p = v;
  else if constexpr (STORE == 1) __stcs(p, v);
  else asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 4 complex per thread -> one float4 in each of the 6 planes (re, im, re+im) x (hi, lo)
template <int STORE>
__global__ void __launch_bounds__(256) split3_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int64_t nquad, int64_t plane_vec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquad; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 z0 = __ldg(in + 2 * i), z1 = __ldg(in + 2 * i + 1);
    const float re[4] = {z0.x, z0.z, z1.x, z1.z}, im[4] = {z0.y, z0.w, z1.y, z1.w};
    float h[3][4], l[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = re[j] + im[j];
      h[0][j] = tf32(re[j]); l[0][j] = tf32(re[j] - h[0][j]);
      h[1][j] = tf32(im[j]); l[1][j] = tf32(im[j] - h[1][j]);
      h[2][j] = tf32(s);     l[2][j] = tf32(s - h[2][j]);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      st4<STORE>(out + (2 * p) * plane_vec + i, make_float4(h[p][0], h[p][1], h[p][2], h[p][3]));
      st4<STORE>(out + (2 * p + 1) * plane_vec + i, make_float4(l[p][0], l[p][1], l[p][2], l[p][3]));
    }
  }
}

// the same mix with the stores issued as bulk copies shared -> global: every warp stages 6 x 512 B and one lane issues 6 cp.async.bulk
__global__ void __launch_bounds__(256) split3_bulk_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int64_t nquad, int64_t plane_vec) {
  __shared__ __align__(128) float4 stage[8][2][6][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 8;
  int buf = 0;
  for (int64_t base = ((int64_t)blockIdx.x * 8 + warp) * 32; base < nquad; base += warps * 32, buf ^= 1) {
    const int64_t i = base + lane;
    const float4 z0 = __ldg(in + 2 * i), z1 = __ldg(in + 2 * i + 1);
    const float re[4] = {z0.x, z0.z, z1.x, z1.z}, im[4] = {z0.y, z0.w, z1.y, z1.w};
    // the buffer written two iterations ago must have been read by its bulk copies
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = p == 0 ? re[j] : p == 1 ? im[j] : re[j] + im[j];
        h[j] = tf32(x);
        l[j] = tf32(x - h[j]);
      }
      stage[warp][buf][2 * p][lane] = make_float4(h[0], h[1], h[2], h[3]);
      stage[warp][buf][2 * p + 1][lane] = make_float4(l[0], l[1], l[2], l[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[warp][buf][q][0]);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(out + q * plane_vec + base), "r"(s) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_ms(F f, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

int main() {
  const int64_t n_complex = (int64_t)1 << 27;       // 1 GiB of complex64, 3 GiB of planes
  const int64_t in_bytes = n_complex * 8, out_bytes = n_complex * 24;
  float4 *in, *out;
  CK(cudaMalloc(&in, in_bytes));
  CK(cudaMalloc(&out, out_bytes));
  CK(cudaMemset(in, 0x3c, in_bytes));
  CK(cudaMemset(out, 0, out_bytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int64_t nvec_in = in_bytes / 16, nquad = n_complex / 4, plane_vec = nquad;
  const int iters = 20;
  for (int per_sm : {4, 8, 16}) {
    const int grid = sms * per_sm;
    float ms = time_ms([&] { copy_kernel<<<grid, 256>>>(in, out, nvec_in); }, iters);
    printf("copy 1R:1W        grid=%5d  %.3f ms  %7.0f GB/s\n", grid, ms, 2.0 * in_bytes / ms / 1e6);
    ms = time_ms([&] { write_kernel<<<grid, 256>>>(out, out_bytes / 16); }, iters);
    printf("write only        grid=%5d  %.3f ms  %7.0f GB/s\n", grid, ms, 1.0 * out_bytes / ms / 1e6);
    ms = time_ms([&] { split3_kernel<0><<<grid, 256>>>(in, out, nquad, plane_vec); }, iters);
    printf("split 1R:3W st    grid=%5d  %.3f ms  %7.0f GB/s\n", grid, ms, 1.0 * (in_bytes + out_bytes) / ms / 1e6);
    ms = time_ms([&] { split3_kernel<1><<<grid, 256>>>(in, out, nquad, plane_vec); }, iters);
    printf("split 1R:3W st.cs grid=%5d  %.3f ms  %7.0f GB/s\n", grid, ms, 1.0 * (in_bytes + out_bytes) / ms / 1e6);
    ms = time_ms([&] { split3_kernel<2><<<grid, 256>>>(in, out, nquad, plane_vec); }, iters);
    printf("split 1R:3W noall grid=%5d  %.3f ms  %7.0f GB/s\n", grid, ms, 1.0 * (in_bytes + out_bytes) / ms / 1e6);
    ms = time_ms([&] { split3_bulk_kernel<<<grid, 256>>>(in, out, nquad, plane_vec); }, iters);
    printf("split 1R:3W bulk  grid=%5d  %.3f ms  %7.0f GB/s\n", grid, ms, 1.0 * (in_bytes + out_bytes) / ms / 1e6);
  }
  CK(cudaDeviceSynchronize());
  // cudaMemcpy d2d as the vendor's own copy
  float ms = time_ms([&] { cudaMemcpyAsync(out, in, in_bytes, cudaMemcpyDeviceToDevice); }, iters);
  printf("cudaMemcpy d2d 1 GiB: %.3f ms  %7.0f GB/s\n", ms, 2.0 * in_bytes / ms / 1e6);
  ms = time_ms([&] { cudaMemsetAsync(out, 0, out_bytes); }, iters);
  printf("cudaMemset 3 GiB: %.3f ms  %7.0f GB/s\n", ms, 1.0 * out_bytes / ms / 1e6);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  return 0;
}
