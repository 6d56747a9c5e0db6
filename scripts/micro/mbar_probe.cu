// Micro-probe: raw mbarrier state word across phases, and latency of try_wait vs a plain ld.shared peek.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(unsigned long long* out) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile uint64_t* vb = &bar;
    out[0] = *vb;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    out[1] = *vb;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    out[2] = *vb;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    out[3] = *vb;
    // latency: try_wait on an already completed phase (parity 0 completed: current phase is 3 -> parity 1; wait for parity 0)
    uint32_t ok;
    long long t0 = clock64();
    for (int i = 0; i < 16; ++i) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (!ok) out[15] = 1;
    }
    long long t1 = clock64();
    out[4] = (t1 - t0) / 16;
    t0 = clock64();
    unsigned long long acc = 0;
    for (int i = 0; i < 16; ++i) { acc += (*vb) >> 63; asm volatile("" ::: "memory"); }
    t1 = clock64();
    out[5] = (t1 - t0) / 16; out[6] = acc;
    t0 = clock64();
    for (int i = 0; i < 16; ++i) {
      asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    t1 = clock64();
    out[7] = (t1 - t0) / 16;
    t0 = clock64();
    for (int i = 0; i < 16; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t1 = clock64();
    out[8] = (t1 - t0) / 16;
  }
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
  probe<<<1, 32>>>(d);
  unsigned long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < 4; ++i) printf("state after %d arrivals: 0x%016llx\n", i, h[i]);
  printf("try_wait (complete) %llu clk; ld.shared peek %llu clk (acc %llu); test_wait %llu clk; tcgen05.fence::after %llu clk; flag %llu\n", h[4], h[5], h[6], h[7], h[8], h[15]);
  return 0;
}
