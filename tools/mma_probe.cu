// Probe: cycles per tcgen05.mma (cta_group::1, SS operands, K-major) on sm_100a as a function of kind, M, N and the
// shared-memory swizzle mode of the operand tiles; what the MMA-bound roles of lstm_tc / tc_conv pay per instruction.
// One CTA per SM issues `reps` MMAs back to back from one thread, commits to an mbarrier and waits. Diagnostic only.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>   // 0: f16, 1: tf32
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// SW: 128 | 64 | 32 (bytes of a swizzle row); kstep: descriptor advance between consecutive MMAs (0: same slab)
template <int KIND>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int M, int N, int SW, int reps, int nacc, int kwalk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t lt = SW == 128 ? 2u : SW == 64 ? 4u : 6u;
    const uint64_t hi = ((uint64_t)((8 * SW) >> 4) << 32) | (1ull << 46) | ((uint64_t)lt << 61);
    const uint32_t a_addr = base, b_addr = base + 48 * 1024;
    const uint64_t da = hi | (uint64_t)(((a_addr & 0x3FFFFu) >> 4) | (1u << 16));
    const uint64_t db = hi | (uint64_t)(((b_addr & 0x3FFFFu) >> 4) | (1u << 16));
    const uint32_t idesc = KIND == 0 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24))
                                     : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24));
    // k slabs inside a swizzle row: 32 bytes each; SW128 has 4, SW64 2, SW32 1 (then the next slab is the next 8-row-group block)
    const int slabs = SW / 32;
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      // 8 MMAs per iteration, everything but the descriptors' low words hoisted: the issuing thread must not be the limit
      const uint32_t d0 = tmem, d1 = tmem + (uint32_t)((nacc > 1 ? 1 : 0) * N);
      const uint64_t a0 = da, a1 = da + (kwalk && slabs > 1 ? 2u : 0u), a2 = da + (kwalk && slabs > 2 ? 4u : 0u), a3 = da + (kwalk && slabs > 2 ? 6u : 0u);
      const uint64_t b0 = db, b1 = db + (kwalk && slabs > 1 ? 2u : 0u), b2 = db + (kwalk && slabs > 2 ? 4u : 0u), b3 = db + (kwalk && slabs > 2 ? 6u : 0u);
      mma<KIND>(d0, a0, b0, idesc, 0u);
      mma<KIND>(d1, a1, b1, idesc, 0u);
      for (int i = 0; i < reps / 8; ++i) {
        mma<KIND>(d0, a0, b0, idesc, 1u);
        mma<KIND>(d1, a1, b1, idesc, 1u);
        mma<KIND>(d0, a2, b2, idesc, 1u);
        mma<KIND>(d1, a3, b3, idesc, 1u);
        mma<KIND>(d0, a0, b0, idesc, 1u);
        mma<KIND>(d1, a1, b1, idesc, 1u);
        mma<KIND>(d0, a2, b2, idesc, 1u);
        mma<KIND>(d1, a3, b3, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      const long long t1 = clock64();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(&bar)), "r"((uint32_t)rep)
                     : "memory");
      const long long t2 = clock64();
      if (rep == 1 && blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int reps = 512;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("%5s %4s %4s %4s %5s %6s | %10s %10s\n", "kind", "M", "N", "SW", "nacc", "kwalk", "issue/mma", "cycles/mma");
  for (int kind = 0; kind < 2; ++kind)
    for (int M : {128, 64})
      for (int N : {32, 64, 128, 256})
        for (int SW : {128, 32})
          for (int nacc : {1, 2})
            for (int kwalk : {0, 1}) {
              if (nacc * N > 512) continue;
              if (M == 64 && (nacc == 2 || kwalk == 0)) continue;
              if (kind == 0) probe<0><<<148, 128, 100 * 1024>>>(out, M, N, SW, reps, nacc, kwalk);
              else probe<1><<<148, 128, 100 * 1024>>>(out, M, N, SW, reps, nacc, kwalk);
              long long h[2] = {0, 0};
              cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
              if (e != cudaSuccess) {
                printf("error: %s\n", cudaGetErrorString(e));
                return 1;
              }
              printf("%5s %4d %4d %4d %5d %6d | %10.1f %10.1f\n", kind == 0 ? "f16" : "tf32", M, N, SW, nacc, kwalk, (double)h[0] / reps,
                     (double)h[1] / reps);
            }
  return 0;
}
