// Probe: can 8 clusters of 16 CTAs (512 threads, ~160 KB smem) be co-resident on a B200, and what does one
// all-to-all exchange of a recurrent state through distributed shared memory + a cluster barrier cost per step?
// (Skeleton of an LSTM recurrence whose h exchange stays inside a cluster.) Diagnostic only.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int CL = 16, THREADS = 512, ITEMS = 8, UNITS = 32, H = 512;

__global__ void __launch_bounds__(THREADS, 1) probe(float* out, int T, int mode) {
  extern __shared__ __align__(16) float smem[];
  float* hbuf = smem;   // [2][ITEMS][H]
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  for (int i = threadIdx.x; i < 2 * ITEMS * H; i += THREADS) hbuf[i] = 0.f;
  cluster.sync();
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    float* cur = hbuf + (t & 1) * ITEMS * H;
    float* nxt = hbuf + ((t + 1) & 1) * ITEMS * H;
    // "compute": read the whole state
    for (int i = threadIdx.x; i < ITEMS * H; i += THREADS) acc += cur[i];
    // publish this CTA's 32 units x 8 items = 256 floats to every CTA of the cluster
    if (mode == 0) {
      // every thread stores one float to 8 peers (512 threads x 8 = 256 floats x 16 peers)
      const int v = threadIdx.x & 255;
      const int item = v >> 5, unit = v & 31;
      const int half = threadIdx.x >> 8;
      for (int p = half * 8; p < half * 8 + 8; ++p) {
        float* dst = cluster.map_shared_rank(nxt, p);
        dst[item * H + rank * UNITS + unit] = acc * 1e-9f + (float)t;
      }
    } else {
      // 16-byte stores: 64 float4 per peer, 16 peers -> 1024 stores over 512 threads
      for (int s = threadIdx.x; s < 64 * CL; s += THREADS) {
        const int p = s >> 6, q = s & 63;
        const int item = q >> 3, u4 = q & 7;
        float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(nxt, p) + item * H + rank * UNITS) + u4;
        *dst = make_float4(acc * 1e-9f, (float)t, 0.f, 1.f);
      }
    }
    cluster.sync();
  }
  out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * THREADS * 4);
  const size_t smem = 160 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(128);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nclusters = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, probe, &cfg);
  printf("max active clusters of %d CTAs: %d (%s)\n", CL, nclusters, cudaGetErrorString(e));
  for (int mode = 0; mode < 2; ++mode) {
    const int T = 3000;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    e = cudaLaunchKernelEx(&cfg, probe, out, 10, mode);
    cudaEventRecord(a);
    e = cudaLaunchKernelEx(&cfg, probe, out, T, mode);
    cudaEventRecord(b);
    cudaError_t e2 = cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("mode %d: launch %s, sync %s, %.3f ms for %d steps = %.3f us per step\n", mode, cudaGetErrorString(e),
           cudaGetErrorString(e2), ms, T, ms * 1e3 / T);
  }
  return 0;
}
