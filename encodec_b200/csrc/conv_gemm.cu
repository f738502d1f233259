// Generic channels-last implicit-GEMM 1-D convolution, fp32 FFMA (CUDA cores).
//
// Replaces, for every SConv1d / SConvTranspose1d with >= 16 input channels:
//   F.pad(mode='reflect') + nn.Conv1d      (reference modules/conv.py:202-221, :116-121)
//   nn.ConvTranspose1d + unpad1d           (reference modules/conv.py:241-263, :156-161)
//   nn.ELU before the conv                 (reference modules/seanet.py:43,126,138,207,226)
//   the residual add of SEANetResnetBlock  (reference modules/seanet.py:63-64), as a second K-source
//   nn.LSTM's input projection             (reference modules/lstm.py:20,24), as a 1-tap conv
//
// GEMM view per item: out[M][N] = A[M][Kt] * W[Kt][N] + bias, with A[m][(j, ci)] =
// f(in[reflect(m*stride + j - pad_left)][ci]) gathered on the fly (never materialised), Kt = taps*C.
// A stride-s, k=2s transposed convolution is the same GEMM with 2 taps, zero padding and N = s*C_out:
// row q of the result is the s output samples q*s .. q*s+s-1, and the trim of conv.py:252-262 is the
// element window [out_lo, out_hi) of the row-major result.
//
// Tiling: BM x BN output tile per CTA, BK = 16, register tile TM x TN per thread, double-buffered smem.
#include "common.cuh"

namespace ecb {

namespace {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
struct TileCfg {
  static constexpr int TX = BN / TN;  // threads along N
  static constexpr int TY = BM / TM;  // threads along M
  static constexpr int THREADS = TX * TY;
  static constexpr int AS_LD = BM + 4;
  static constexpr int A_F4 = BM * BK / 4;  // float4 loads per A chunk
  static constexpr int B_F4 = BK * BN / 4;
  static constexpr int A_PER_THREAD = (A_F4 + THREADS - 1) / THREADS;
  static constexpr int B_PER_THREAD = (B_F4 + THREADS - 1) / THREADS;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(TileCfg<BM, BN, TM, TN>::THREADS, 2)
conv_gemm_kernel(const ConvParams p) {
  using Cfg = TileCfg<BM, BN, TM, TN>;
  constexpr int THREADS = Cfg::THREADS;
  constexpr int AS_LD = Cfg::AS_LD;
  static_assert(TM == 8 || TM == 4, "TM");
  static_assert(TN == 8 || TN == 4, "TN");

  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % Cfg::TX;
  const int ty = tid / Cfg::TX;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int item = blockIdx.z;

  const float* __restrict__ in0 = p.s0.ptr + (long long)item * p.s0.item_stride;
  const float* __restrict__ in1 = p.s1.taps ? p.s1.ptr + (long long)item * p.s1.item_stride : nullptr;
  const float* __restrict__ w = p.w;

  const int cpt0 = p.s0.C / BK;  // chunks per tap, source 0
  const int nch0 = p.s0.taps * cpt0;
  const int nch1 = p.s1.taps ? p.s1.C / BK : 0;
  const int nch = nch0 + nch1;

  float4 a_reg[Cfg::A_PER_THREAD];
  float4 b_reg[Cfg::B_PER_THREAD];

  auto load_chunk = [&](int q) {
    // ---- A: BM rows x 16 channels of the (virtual) im2col matrix
    const bool second = q >= nch0;
    int j, ci0;
    if (!second) {
      j = q / cpt0;
      ci0 = (q - j * cpt0) * BK;
    } else {
      j = 0;
      ci0 = (q - nch0) * BK;
    }
#pragma unroll
    for (int i = 0; i < Cfg::A_PER_THREAD; ++i) {
      const int f = tid + i * THREADS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (Cfg::A_F4 % THREADS == 0 || f < Cfg::A_F4) {
        const int m = f >> 2;
        const int c4 = f & 3;
        const int mg = m0 + m;
        if (mg < p.M) {
          if (!second) {
            int r = mg * p.stride + j - p.pad_left;
            bool ok = true;
            if (p.pad_zero) {
              ok = (r >= 0) && (r < p.s0.T);
            } else {
              r = reflect_index(r, p.s0.T_ref);
              ok = r < p.s0.T;
            }
            if (ok) {
              v = __ldg(reinterpret_cast<const float4*>(in0 + (long long)r * p.s0.C + ci0 + c4 * 4));
              if (p.s0.elu) {
                v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w);
              }
            }
          } else {
            v = __ldg(reinterpret_cast<const float4*>(in1 + (long long)mg * p.s1.C + ci0 + c4 * 4));
            if (p.s1.elu) {
              v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w);
            }
          }
        }
      }
      a_reg[i] = v;
    }
    // ---- B: 16 x BN slab of the packed weights
    const long long krow0 = (long long)q * BK;
#pragma unroll
    for (int i = 0; i < Cfg::B_PER_THREAD; ++i) {
      const int f = tid + i * THREADS;
      if (Cfg::B_F4 % THREADS == 0 || f < Cfg::B_F4) {
        const int k = f / (BN / 4);
        const int n4 = f - k * (BN / 4);
        b_reg[i] = __ldg(reinterpret_cast<const float4*>(w + (krow0 + k) * p.N + n0 + n4 * 4));
      }
    }
  };

  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int i = 0; i < Cfg::A_PER_THREAD; ++i) {
      const int f = tid + i * THREADS;
      if (Cfg::A_F4 % THREADS == 0 || f < Cfg::A_F4) {
        const int m = f >> 2;
        const int c4 = f & 3;
        As[buf][c4 * 4 + 0][m] = a_reg[i].x;
        As[buf][c4 * 4 + 1][m] = a_reg[i].y;
        As[buf][c4 * 4 + 2][m] = a_reg[i].z;
        As[buf][c4 * 4 + 3][m] = a_reg[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < Cfg::B_PER_THREAD; ++i) {
      const int f = tid + i * THREADS;
      if (Cfg::B_F4 % THREADS == 0 || f < Cfg::B_F4) {
        const int k = f / (BN / 4);
        const int n4 = f - k * (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][k][n4 * 4]) = b_reg[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // rows owned by this thread: groups of 4, group g at ty*4 + g*(BM/ (TM/4)) ; same for columns
  constexpr int MG = TM / 4;
  constexpr int NG = TN / 4;
  constexpr int M_GSTRIDE = BM / MG;
  constexpr int N_GSTRIDE = BN / NG;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();

  for (int q = 0; q < nch; ++q) {
    const int buf = q & 1;
    if (q + 1 < nch) load_chunk(q + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < MG; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + g * M_GSTRIDE]);
        a[g * 4 + 0] = t.x; a[g * 4 + 1] = t.y; a[g * 4 + 2] = t.z; a[g * 4 + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + g * N_GSTRIDE]);
        b[g * 4 + 0] = t.x; b[g * 4 + 1] = t.y; b[g * 4 + 2] = t.z; b[g * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (q + 1 < nch) {
      store_chunk(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: bias, optional ELU, windowed store, optional GroupNorm partial statistics
  float* __restrict__ outp = p.out + (long long)item * p.out_item_stride;
  float lsum = 0.f, lsq = 0.f;
#pragma unroll
  for (int gi = 0; gi < MG; ++gi) {
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int i = gi * 4 + ii;
      const int mg = m0 + ty * 4 + gi * M_GSTRIDE + ii;
      if (mg >= p.M) continue;
#pragma unroll
      for (int gj = 0; gj < NG; ++gj) {
        const int n = n0 + tx * 4 + gj * N_GSTRIDE;
        float4 v = make_float4(acc[i][gj * 4 + 0], acc[i][gj * 4 + 1], acc[i][gj * 4 + 2], acc[i][gj * 4 + 3]);
        if (p.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (p.stats) {
          lsum += (v.x + v.y) + (v.z + v.w);
          lsq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        if (p.out_elu) {
          v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w);
        }
        const long long e = (long long)mg * p.N + n;
        if (e >= p.out_lo && e < p.out_hi) *reinterpret_cast<float4*>(outp + (e - p.out_lo)) = v;
      }
    }
  }
  if (p.stats) {
    __shared__ float red[2][THREADS / 32];
    lsum = warp_sum(lsum);
    lsq = warp_sum(lsq);
    if ((tid & 31) == 0) {
      red[0][tid >> 5] = lsum;
      red[1][tid >> 5] = lsq;
    }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0, sq = 0.0;
      for (int i = 0; i < THREADS / 32; ++i) {
        s += (double)red[0][i];
        sq += (double)red[1][i];
      }
      const long long slot = ((long long)item * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x) * 2;
      p.stats[slot] = s;
      p.stats[slot + 1] = sq;
    }
  }
}

struct Pick {
  int bm, bn;
};

Pick pick_tile(const ConvParams& p) {
  if (p.N % 128 == 0) return {128, 128};
  if (p.N % 64 == 0) return {128, 64};
  if (p.N % 32 == 0) return {256, 32};
  return {256, 16};
}

}  // namespace

int conv_gemm_stat_slots(const ConvParams& p) {
  Pick t = pick_tile(p);
  return (int)(cdiv(p.M, t.bm) * (p.N / t.bn));
}

int launch_conv_gemm(const ConvParams& p, cudaStream_t stream) {
  ECB_REQUIRE(p.N % 16 == 0 && p.N > 0, "conv_gemm: N=%d must be a positive multiple of 16", p.N);
  ECB_REQUIRE(p.s0.C % 16 == 0 && p.s0.taps > 0, "conv_gemm: C0=%d must be a multiple of 16", p.s0.C);
  ECB_REQUIRE(p.s1.taps == 0 || (p.s1.taps == 1 && p.s1.C % 16 == 0), "conv_gemm: bad second source");
  ECB_REQUIRE(p.M > 0 && p.n_items > 0 && p.n_items <= 65535, "conv_gemm: bad M=%d / items=%d", p.M, p.n_items);
  ECB_REQUIRE(p.out_lo % 4 == 0 && p.out_hi % 4 == 0, "conv_gemm: output window must be float4 aligned");
  if (!p.pad_zero) {
    const long long last = (long long)(p.M - 1) * p.stride + p.s0.taps - 1 - p.pad_left;
    ECB_REQUIRE(p.pad_left < p.s0.T_ref && last - (p.s0.T_ref - 1) < p.s0.T_ref && p.s0.T_ref >= p.s0.T,
                "conv_gemm: reflection length %d too short for T=%d", p.s0.T_ref, p.s0.T);
  }
  Pick t = pick_tile(p);
  dim3 grid((unsigned)cdiv(p.M, t.bm), (unsigned)(p.N / t.bn), (unsigned)p.n_items);
  const double kt = (double)p.s0.taps * p.s0.C + (double)p.s1.taps * p.s1.C;
  const double rows = (double)p.M * p.n_items;
  ProfScope prof(PROF_CONV_GEMM, stream, 2.0 * rows * p.N * kt,
                 4.0 * (rows * p.stride * p.s0.C + rows * p.s1.taps * p.s1.C + kt * p.N +
                        (double)(p.out_hi - p.out_lo) * p.n_items));
  if (t.bn == 128) {
    conv_gemm_kernel<128, 128, 8, 8><<<grid, TileCfg<128, 128, 8, 8>::THREADS, 0, stream>>>(p);
  } else if (t.bn == 64) {
    conv_gemm_kernel<128, 64, 8, 4><<<grid, TileCfg<128, 64, 8, 4>::THREADS, 0, stream>>>(p);
  } else if (t.bn == 32) {
    conv_gemm_kernel<256, 32, 8, 4><<<grid, TileCfg<256, 32, 8, 4>::THREADS, 0, stream>>>(p);
  } else {
    conv_gemm_kernel<256, 16, 4, 4><<<grid, TileCfg<256, 16, 4, 4>::THREADS, 0, stream>>>(p);
  }
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
