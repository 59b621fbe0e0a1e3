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

// Integer pipe throughput with volatile PTX (ptxas keeps every op):
// 0 imad  1 min2  2 min3  3 add+min (VIADDMNMX)  4 xor (LOP3)  5 iadd3  6 ffma
// 7 imad + 0.5 min3 (top-1 key stream)  8 imad-imm (x*-256+y)  9 setp+selp  10 max2+min3 (merge)
template <int KIND>
__global__ void __launch_bounds__(512, 1) alu_kernel(int iters, int a0, int b0, long long* cyc,
                                                    int* sink) {
  int x[8], y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 7 + k * a0; y[k] = threadIdx.x * 13 + k * b0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (KIND == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
        if (KIND == 1) {
          if (u & 1) asm volatile("min.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));
          else asm volatile("max.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[(k + 3) & 7]));
        }
        if (KIND == 2) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(x[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
        if (KIND == 3) asm volatile("{.reg .s32 t; add.s32 t, %1, %2; min.s32 %0, t, %0;}" : "+r"(x[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
        if (KIND == 4) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));
        if (KIND == 5) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(x[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
        if (KIND == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
        if (KIND == 7) {
          int t;
          asm volatile("mad.lo.s32 %0, %1, -256, %2;" : "=r"(t) : "r"(y[k]), "r"(x[k >> 1]));
          if (k & 1) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(x[k >> 1]) : "r"(t), "r"(x[4 + (k >> 1)]));
          else x[4 + (k >> 1)] = t;
        }
        if (KIND == 8) asm volatile("mad.lo.s32 %0, %0, -256, %1;" : "+r"(x[k]) : "r"(y[k]));
        if (KIND == 9) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %2, %0, p;}" : "+r"(x[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
        if (KIND == 11) { unsigned b; asm volatile("{.reg .pred p; setp.lt.s32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "=r"(b) : "r"(x[k]), "r"(y[k])); x[k] += b; }
        if (KIND == 12) { unsigned b; asm volatile("{.reg .pred p, q; setp.lt.s32 p, %1, %2; vote.sync.any.pred q, p, 0xffffffff; selp.u32 %0, 1, 0, q;}" : "=r"(b) : "r"(x[k]), "r"(y[k])); x[k] += b; }
        if (KIND == 13) { unsigned b; asm volatile("redux.sync.or.b32 %0, %1, 0xffffffff;" : "=r"(b) : "r"(x[k])); x[k] += b; }
        if (KIND == 10) asm volatile("{.reg .s32 t, u; max.s32 t, %0, %2; min.s32 %0, %0, %2; min.s32 u, t, %1; min.s32 %1, u, %3;}" : "+r"(x[k]), "+r"(y[k]) : "r"(y[(k + 1) & 7]), "r"(y[(k + 2) & 7]));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x[k] ^ y[k];
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
  const char* names[] = {"imad", "min2", "min3", "addmin", "xor", "iadd3", "ffma", "imad_halfmin3",
                         "imad_imm", "setp_selp", "merge_top2", "ballot", "vote_any", "redux_or"};
  for (int warps : {4, 8, 16}) {
    for (int kind = 0; kind < 14; ++kind) {
      const int it = 2000;
      for (int rep = 0; rep < 2; ++rep) {
#define RUN(K) case K: alu_kernel<K><<<n_sms, warps * 32>>>(it, 3, 5, cyc, (int*)sink); break;
        switch (kind) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) }
      }
      if (cudaDeviceSynchronize() != cudaSuccess) { printf(", \"error\": \"alu\"}\n"); return 1; }
      const double c = (double)maxcyc(cyc, n_sms);
      // asm blocks per thread: it * 4 * 8
      printf(", \"%s_blocks_per_clk_sm_w%d\": %.1f", names[kind], warps,
             (double)it * 32.0 * warps * 32.0 / c);
    }
  }
  printf("}\n");
  return 0;
}
