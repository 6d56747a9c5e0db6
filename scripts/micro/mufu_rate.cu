// Micro-probe: per-SM throughput of the exponential flavours the softmax could use (B200, sm_100a):
//   ex2.approx.ftz.f32 (MUFU.EX2), ex2.approx.f16x2, ex2.approx.ftz.bf16x2 (two results per instruction if the MUFU
//   pipe executes packed halves in one pass), and the packed-FMA polynomial of csrc/attn.cu.
// Each warp runs 8 independent dependency chains; 16 warps per SM, one CTA per SM. Prints results per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, long long* clk, float seed) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (i + 1) * 0.01f - threadIdx.x * 1e-4f; h[i] = 0x3c003800u + i + threadIdx.x; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 3) {
        // scalar FMA-pipe filler: is the FMA pipe free while MUFU runs? (ex2 + 4 dependent FFMA per chain)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        a[i] = fmaf(a[i], 0.999f, 0.001f); a[i] = fmaf(a[i], 0.999f, 0.001f);
        a[i] = fmaf(a[i], 0.999f, 0.001f); a[i] = fmaf(a[i], 0.999f, 0.001f);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int results_per_instr, int extra) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 8);
  probe<MODE><<<148, 512>>>(out, clk, 1.0f);
  probe<MODE><<<148, 512>>>(out, clk, 1.0f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double instr = double(ITERS) * 8 * 512;   // thread-level instructions per SM
  printf("%-28s %8.0f clk  %6.2f instr/clk/SM  %6.2f results/clk/SM  (%s)\n", name, avg, instr / avg,
         instr * results_per_instr / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1, 0);
  run<1>("ex2.approx.f16x2", 2, 0);
  run<2>("ex2.approx.ftz.bf16x2", 2, 0);
  run<3>("ex2.f32 + 4 FFMA", 1, 4);
  return 0;
}
