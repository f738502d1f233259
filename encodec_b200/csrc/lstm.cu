// SLSTM recurrence (reference modules/lstm.py:12-28 -> nn.LSTM(512, 512, 2), gates i,f,g,o, zero state).
//
// The input projection W_ih x_t + b_ih + b_hh of ALL time steps is one GEMM (conv_gemm.cu, 1 tap);
// this file holds the part that is sequential in time: gates_t = pre_t + W_hh h_{t-1}, the cell update
// and the skip connection y = h + x (lstm.py:25-26).
//
// One persistent cooperative kernel per LSTM layer: 128 CTAs, CTA c owns hidden units 4c..4c+3 (16 gate
// rows of W_hh, 32 KB, resident in shared memory for the whole sequence). Per step every CTA pulls
// h_{t-1} (batch tile x 512) from L2 into shared memory, computes its 64 x 16 gate block with K split over
// 4 warp pairs, reduces, applies the cell non-linearity with one (batch, unit) pair per thread, publishes
// its 4 columns of h_t and joins a grid-wide barrier (one atomic counter).
#include <cooperative_groups.h>

#include "common.cuh"

namespace ecb {
namespace {

constexpr int LH = 512;             // hidden size
constexpr int L_UNITS = 4;          // hidden units per CTA
constexpr int L_CTAS = LH / L_UNITS;  // 128
constexpr int L_BT = 64;            // batch tile
constexpr int L_HLD = LH + 4;       // padded h row in smem (floats)
constexpr int L_THREADS = 256;

struct LstmParams {
  const float* pre;    // [B][T][4H]
  const float* wp;     // packed W_hh: [CTA][k4=128][gate=4][unit=4][4]
  const float* skip;   // [B][T][H] or nullptr
  float* out;          // [B][T][H]
  float* hbuf;         // [2][B][H] recurrent state, hbuf[0] zeroed by the host
  unsigned int* bar;   // grid barrier counter, zeroed by the host
  long long skip_stride, out_stride;  // floats between items
  int B, T, out_elu;
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(L_THREADS, 1)
lstm_recurrent_kernel(const LstmParams p) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                              // [128][4][4][4] = 8192 floats
  float* hs = ws + 8192;                         // [64][516]
  float* part = hs + L_BT * L_HLD;               // [4 ksplit][4 i][64][4] = 4096 floats
  float* cs = part + 4096;                       // [n_btiles][256] cell state

  const int tid = threadIdx.x;
  const int cta = blockIdx.x;
  const int n_bt = (p.B + L_BT - 1) / L_BT;

  {  // resident W_hh slice
    const float4* src = reinterpret_cast<const float4*>(p.wp + (long long)cta * 8192);
    float4* dst = reinterpret_cast<float4*>(ws);
    for (int i = tid; i < 2048; i += L_THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < n_bt * 256; i += L_THREADS) cs[i] = 0.f;
  }
  __syncthreads();

  // GEMM-phase role: K split ks, thread slot t64 = (w2, g, u): rows b = 32*w2 + g + 8*i, unit u
  const int ks = tid >> 6;
  const int t64 = tid & 63;
  const int w2 = t64 >> 5;
  const int g = (t64 & 31) >> 2;
  const int u = t64 & 3;
  // cell-phase role: one (row, unit) pair per thread: i_c = tid / 64 selects which of the slot's 4 rows
  const int i_c = tid >> 6;
  const int b_c = 32 * w2 + g + 8 * i_c;          // row inside the batch tile
  const int unit = cta * L_UNITS + u;

  unsigned int bar_target = 0;
  for (int t = 0; t < p.T; ++t) {
    const float* hprev = p.hbuf + (long long)(t & 1) * p.B * LH;
    float* hnext = p.hbuf + (long long)((t + 1) & 1) * p.B * LH;
    for (int bt = 0; bt < n_bt; ++bt) {
      const int b0 = bt * L_BT;
      // prefetch this thread's pre-gates (read-only input, independent of the recurrence)
      const int bg = b0 + b_c;
      float pg[4] = {0.f, 0.f, 0.f, 0.f};
      float skipv = 0.f;
      if (bg < p.B) {
        const float* pr = p.pre + ((long long)bg * p.T + t) * (4 * LH) + unit;
#pragma unroll
        for (int q = 0; q < 4; ++q) pg[q] = __ldg(pr + q * LH);
        if (p.skip) skipv = __ldg(p.skip + (long long)bg * p.skip_stride + (long long)t * LH + unit);
      }
      // h_{t-1} tile -> shared (L2 only: other CTAs wrote it during the previous step)
      for (int f = tid; f < L_BT * (LH / 4); f += L_THREADS) {
        const int r = f >> 7;
        const int c4 = f & 127;
        float* dst = hs + r * L_HLD + c4 * 4;
        if (b0 + r < p.B) {
          const float* src = hprev + (long long)(b0 + r) * LH + c4 * 4;
          const unsigned int sa = (unsigned int)__cvta_generic_to_shared(dst);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();

      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
      const float* hrow = hs + (32 * w2 + g) * L_HLD;
#pragma unroll 4
      for (int k4 = ks * 32; k4 < ks * 32 + 32; ++k4) {
        float4 hv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(hrow + (8 * i) * L_HLD + k4 * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = *reinterpret_cast<const float4*>(ws + ((k4 * 4 + q) * 4 + u) * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[i][q] = fmaf(hv[i].x, wv[q].x, acc[i][q]);
            acc[i][q] = fmaf(hv[i].y, wv[q].y, acc[i][q]);
            acc[i][q] = fmaf(hv[i].z, wv[q].z, acc[i][q]);
            acc[i][q] = fmaf(hv[i].w, wv[q].w, acc[i][q]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + ((ks * 4 + i) * 64 + t64) * 4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();

      // cell update for (row b_c, unit u)
      float4 s = *reinterpret_cast<const float4*>(part + ((0 * 4 + i_c) * 64 + t64) * 4);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(part + ((k * 4 + i_c) * 64 + t64) * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      if (bg < p.B) {
        const float gi = sigmoidf_acc(pg[0] + s.x);
        const float gf = sigmoidf_acc(pg[1] + s.y);
        const float gg = tanhf(pg[2] + s.z);
        const float go = sigmoidf_acc(pg[3] + s.w);
        const float c_new = gf * cs[bt * 256 + tid] + gi * gg;
        cs[bt * 256 + tid] = c_new;
        const float h_new = go * tanhf(c_new);
        hnext[(long long)bg * LH + unit] = h_new;
        float y = h_new;
        if (p.skip) y += skipv;
        if (p.out_elu) y = elu1(y);
        p.out[(long long)bg * p.out_stride + (long long)t * LH + unit] = y;
      }
      // `part` and `hs` are rewritten only after the next tile's barriers
    }
    bar_target += L_CTAS;
    grid_barrier(p.bar, bar_target);
  }
}

// W_hh [4H][H] -> [CTA][k4][gate][unit][4]
__global__ void pack_whh_kernel(const float* __restrict__ w, float* __restrict__ out) {
  const int n = L_CTAS * 8192;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i & 3;
    const int uu = (i >> 2) & 3;
    const int q = (i >> 4) & 3;
    const int k4 = (i >> 6) & 127;
    const int cta = i >> 13;
    out[i] = w[((long long)q * LH + cta * L_UNITS + uu) * LH + k4 * 4 + e];
  }
}

size_t lstm_smem_bytes(int batch) {
  const int n_bt = (batch + L_BT - 1) / L_BT;
  return sizeof(float) * (size_t)(8192 + L_BT * L_HLD + 4096 + n_bt * 256);
}

}  // namespace

int lstm_recurrent_workspace_floats(int batch) { return 2 * batch * LH + 64; }

int launch_pack_lstm_whh(const float* w_hh, float* packed, int H, cudaStream_t s) {
  ECB_REQUIRE(H == LH, "lstm: hidden size %d unsupported (only %d)", H, LH);
  pack_whh_kernel<<<512, 256, 0, s>>>(w_hh, packed);
  ECB_LAUNCHED();
  return 0;
}

int launch_lstm_recurrent(const float* pre, const float* w_hh_packed, const float* skip, long long skip_item_stride,
                          float* out, long long out_item_stride, int batch, int T, int H, int out_elu,
                          float* workspace, cudaStream_t s) {
  ECB_REQUIRE(H == LH, "lstm: hidden size %d unsupported (only %d)", H, LH);
  ECB_REQUIRE(batch > 0 && T > 0, "lstm: bad batch %d / T %d", batch, T);
  const size_t smem = lstm_smem_bytes(batch);
  ECB_REQUIRE(smem <= 227 * 1024, "lstm: batch %d needs %zu bytes of shared memory; split the batch", batch, smem);
  static bool attr_set = false;
  if (!attr_set) {
    ECB_CUDA(cudaFuncSetAttribute(lstm_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  LstmParams p;
  p.pre = pre;
  p.wp = w_hh_packed;
  p.skip = skip;
  p.out = out;
  p.skip_stride = skip_item_stride ? skip_item_stride : (long long)T * LH;
  p.out_stride = out_item_stride ? out_item_stride : (long long)T * LH;
  p.hbuf = workspace;
  p.bar = reinterpret_cast<unsigned int*>(workspace + 2 * (size_t)batch * LH);
  p.B = batch;
  p.T = T;
  p.out_elu = out_elu;
  // zero h_{-1} and the barrier counter
  ECB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)batch * LH, s));
  ECB_CUDA(cudaMemsetAsync(p.bar, 0, 64 * sizeof(float), s));
  const double bt = (double)batch * T;
  ProfScope prof(PROF_LSTM_REC, s, 2.0 * bt * 4 * H * H, 4.0 * (bt * 4 * H + bt * H * (skip ? 2 : 1) + 4.0 * H * H));
  void* args[] = {(void*)&p};
  ECB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_recurrent_kernel, dim3(L_CTAS), dim3(L_THREADS), args, smem, s));
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
