// Probe: fp32 FMA issue rates on sm_100a with register operands (scalar FFMA vs packed FFMA2), the LSTM recurrence's
// inner-loop shape: acc[r][i] += w[r][k] * h[i][k] with w resident in registers. Diagnostic only.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) probe(float* out, const float* in, int iters) {
  // 4 rows x 8 k-pairs of weights (64 regs), 4 items
  float2 w[4][8];
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 8; ++k) w[r][k] = make_float2(in[(r * 8 + k) * 2 + threadIdx.x % 7], in[(r * 8 + k) * 2 + 1 + threadIdx.x % 5]);
  float2 acc[4][4];
  for (int r = 0; r < 4; ++r)
    for (int i = 0; i < 4; ++i) acc[r][i] = make_float2(0.f, 0.f);
  float2 h[4][8];
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 8; ++k) h[i][k] = make_float2(in[64 + i * 8 + k + threadIdx.x % 3], in[100 + i * 8 + k]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          // MODE 0/1: row innermost (h operand repeats); MODE 2: item innermost (w operand repeats);
          // MODE 3: k innermost (accumulator chain, nothing repeats)
          const int k = MODE == 3 ? rr * 2 + (ii & 1) : kk;
          const int i = MODE == 2 ? rr : (MODE == 3 ? (kk & 3) : ii);
          const int r = MODE == 2 ? ii : (MODE == 3 ? (ii >> 1) + 2 * (kk >> 2) : rr);
          if (MODE == 0) {
            acc[r][i].x = fmaf(w[r][k].x, h[i][k].x, acc[r][i].x);
            acc[r][i].y = fmaf(w[r][k].y, h[i][k].y, acc[r][i].y);
          } else {
            unsigned long long a, b, c;
            a = *reinterpret_cast<unsigned long long*>(&w[r][k]);
            b = *reinterpret_cast<unsigned long long*>(&h[i][k]);
            c = *reinterpret_cast<unsigned long long*>(&acc[r][i]);
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
            acc[r][i] = *reinterpret_cast<float2*>(&c);
          }
        }
    // perturb h slightly so nothing is hoisted
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i][0].x += 1e-9f;
  }
  float s = 0.f;
  for (int r = 0; r < 4; ++r)
    for (int i = 0; i < 4; ++i) s += acc[r][i].x + acc[r][i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int MAXT>
void run(const char* name) {
  const int threads = MAXT;
  float *out, *in;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  probe<MODE, MAXT><<<148, threads>>>(out, in, 100);
  cudaEventRecord(a);
  probe<MODE, MAXT><<<148, threads>>>(out, in, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double fma = 148.0 * threads * iters * 8 * 4 * 4 * 2;
  printf("%-8s threads/SM %4d: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, threads, ms, 2 * fma / ms / 1e9,
         fma / 148 / (ms * 1e-3 * 1.9e9));
  cudaFree(out);
  cudaFree(in);
}

int main() {
#define PROBE_ALL(T)           \
  run<0, T>("FFMA");           \
  run<1, T>("FFMA2");          \
  run<2, T>("FFMA2-i");        \
  run<3, T>("FFMA2-k");
  PROBE_ALL(128)
  PROBE_ALL(256)
  PROBE_ALL(384)
  PROBE_ALL(512)
  return 0;
}
