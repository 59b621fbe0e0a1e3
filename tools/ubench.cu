// ubench.cu -- sm_100a micro-benchmarks that size the kNN-2 epilogue budget:
//   * tcgen05.ld (TMEM -> registers) throughput per SM for 4 / 8 / 16 reading warps
//   * integer pipe throughput: IMAD, 2-input min, 3-input min (DPX), LEA-style shift-add, mixes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench ubench.cu
// Prints one JSON object.  Cycle counts are SM clocks (clock64), max over the 148 CTAs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../sfm_opencv_b200/csrc/ptx.cuh"
using namespace sfm;

__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, int wait_every, long long* cyc,
                                                     uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_ptr), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr);
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    tmem_ld_x32(base + lane_addr + ((i * 32 + (warp >> 2) * 64) & 511 & ~31), r);
    if ((i % wait_every) == wait_every - 1) tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 32; k += 8) acc ^= r[k];
  }
  tmem_ld_wait();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(base, 512);
  }
}

// kind: 7 fused add+min (VIADDMNMX); 0 imad, 1 min2, 2 min3, 3 shift-add (lea), 4 imad+min3 (1:0.5), 5 iadd3, 6 fadd-ish fma
template <int KIND>
__global__ void __launch_bounds__(512, 1) alu_kernel(int iters, int a0, int b0, long long* cyc,
                                                    int* sink) {
  int x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 7 + k * a0;
  int a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  int y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) y[k] = threadIdx.x * 13 + k * b0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (KIND == 0) x[k] = x[k] * a + b;
        if (KIND == 1) x[k] = min(x[k] ^ a, b + k);
        if (KIND == 2) x[k] = __vimin3_s32(x[k], a + k, b ^ x[(k + 1) & 7]);
        if (KIND == 3) x[k] = (x[k] << 9) + a;
        if (KIND == 4) {
          const int key = x[k] * -512 + a;
          if (k & 1) x[k] = __vimin3_s32(x[k], key, x[k - 1]); else x[k] = key + b;
        }
        if (KIND == 5) x[k] = x[k] + a + b;
        if (KIND == 6) x[k] = __float_as_int(__int_as_float(x[k]) * 1.0001f + 0.5f);
        if (KIND == 7) x[k] = min(x[k], y[k] + a);
      }
      a += 3; b -= 5;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x[k];
  if (s == 0x7fffffff) sink[0] = s;
}

static long long maxcyc(long long* d, int n) {
  static long long h[1024];
  cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  long long m = 0;
  for (int i = 0; i < n; ++i) m = h[i] > m ? h[i] : m;
  return m;
}

int main() {
  int n_sms = 0;
  cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, 0);
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaMalloc(&sink, 64);
  printf("{\"n_sms\": %d", n_sms);
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    for (int we : {1, 4}) {
      ldtm_kernel<<<n_sms, warps * 32, 0>>>(iters, we, cyc, sink);
      ldtm_kernel<<<n_sms, warps * 32, 0>>>(iters, we, cyc, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf(", \"error\": \"ldtm\"}\n"); return 1; }
      const double c = (double)maxcyc(cyc, n_sms);
      printf(", \"ldtm_B_per_clk_sm_w%d_wait%d\": %.1f", warps, we,
             (double)iters * warps * 4096.0 / c);
    }
  }
  const char* names[] = {"imad", "min2", "min3", "shladd", "imad_min3_mix", "iadd3", "ffma", "viaddmnmx"};
  for (int warps : {4, 8, 16}) {
    for (int kind = 0; kind < 8; ++kind) {
      const int it = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        switch (kind) {
          case 0: alu_kernel<0><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 1: alu_kernel<1><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 2: alu_kernel<2><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 3: alu_kernel<3><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 4: alu_kernel<4><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 5: alu_kernel<5><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 6: alu_kernel<6><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
          case 7: alu_kernel<7><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
        }
      }
      if (cudaDeviceSynchronize() != cudaSuccess) { printf(", \"error\": \"alu\"}\n"); return 1; }
      const double c = (double)maxcyc(cyc, n_sms);
      // source-level ops per thread: it * 4 * 8
      printf(", \"%s_srcops_per_clk_sm_w%d\": %.1f", names[kind], warps,
             (double)it * 32.0 * warps * 32.0 / c);
    }
  }
  printf("}\n");
  return 0;
}
