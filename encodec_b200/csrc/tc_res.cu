// One kernel per SEANetResnetBlock at the 32-channel level (reference modules/seanet.py:37-64):
//     Y = ELU( shortcut(X) + block3( ELU( block1( ELU(X) ) ) ) )          (the trailing ELU belongs to the next module)
// with block1 = SConv1d(32 -> 16, k3), block3 = SConv1d(16 -> 32, k1), shortcut = SConv1d(32 -> 32, k1).
//
// The two-kernel form (tc_conv.cu: block1, then block3 + shortcut as one GEMM) moves 7.5 tensors of [T][32] through HBM
// per block (X and ELU(X) written by the producer, ELU(X) read, H written and read, X read, Y written); this level holds
// a third of the codec's activation bytes. Fused, the block reads X once and writes Y once:
//   * warp 0 TMA-loads the raw X tile ([128 + 8 rows x 32 ch], halo-padded input, so the 3 taps are row shifts of it),
//   * warps 2-9 turn it into the tensor-core operands in shared memory: x_lo, e = ELU(x), e_lo (split-operand TF32),
//   * warp 1 issues GEMM 1 (3 taps x 4 K steps, [W1_hi | W1_lo] and the a_lo correction) into TMEM,
//   * warps 10-13 read it back, add the bias, apply ELU and write H / H_lo to shared memory as the A operand of GEMM 2,
//   * warp 1 issues GEMM 2: H * W3 + X * Ws (the shortcut reads the SAME X tile, shifted by the conv's left padding),
//   * warps 14-17 read the result, add the biases, apply ELU and store Y (plus its reflected halo rows).
// Both weight matrices (28 KB as hi/lo K-major tiles) stay resident in shared memory for the whole kernel; raw X tiles sit
// in a 3-deep TMA ring, the processed operands and both TMEM accumulators are double-buffered, so loads run ahead and
// GEMM 1 of tile i+1 runs while tile i is between its two GEMMs.
// The element-wise work (three ELUs per element, ~16 instructions each) is what bounds it: ~2250 issue cycles per tile.
// SPLIT = 3: split-operand TF32 (fp32-accurate; the encoder). SPLIT = 1: one TF32 pass (the decoder's default scheme): X
// arrives TF32-rounded from its producer, e = rn_tf32(ELU(x)) and H = rn_tf32(ELU(.)) are rounded when they are written,
// the *_lo operands, their MMAs and the correction accumulators do not exist.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ecb {
namespace {
using namespace tc;

constexpr int RC = 32;                    // channels of the block (and padded hidden width of the H operand)
constexpr int RH = 16;                    // real hidden width (block.1 output channels)
constexpr int RM = 128;                   // output rows per tile
constexpr int RROWS = RM + 8;             // X tile rows (halo for the 3 taps)
constexpr int XT = RROWS * RC * 4;        // 17408 B
constexpr int HT = RM * RC * 4;           // 16384 B
constexpr int W1CH = 2 * RH * RC * 4;     // one block.1 weight chunk: hi tile [16 x 32] + lo tile = 4096 B
constexpr int WCCH = 2 * RC * RC * 4;     // one [block.3 ; shortcut] chunk: hi tile [32 x 32] + lo tile = 8192 B
constexpr int NRAW = 3;                   // raw X ring (TMA runs this many tiles ahead of GEMM 2)
constexpr int P_STAGE = 3 * XT;           // processed stage: x_lo | e | e_lo
constexpr int R_THREADS = 576;   // 18 warps: producer, MMA issuer, 8 transform, 4 mid-epilogue, 4 final-epilogue
constexpr int OFF_P = NRAW * XT;                    // 2 processed stages
constexpr int OFF_H = OFF_P + 2 * P_STAGE;          // H | H_lo
constexpr int OFF_W = OFF_H + 2 * HT;               // 3 chunks of W1, 2 chunks of Wcat
constexpr int OFF_STG = OFF_W + 3 * W1CH + 2 * WCCH;   // 4 final-epilogue warps x [32 rows x 64 B]
constexpr int OFF_BAR = OFF_STG + 4 * 2048;
constexpr int R_SMEM = OFF_BAR + 256 + 1024;
static_assert(R_SMEM <= 232448, "tc_res: shared memory budget");

struct ResArgs {
  const float* b1;       // [32] (hidden bias, zero-padded)
  const float* bcat;     // [32] b3 + bs
  float* out;            // (item 0, row 0) of Y
  long long out_item_stride;
  int M, n_items, tiles_m, total_tiles;
  int row_base;          // row coordinate of output row 0's first tap in the X map
  int pad_left;          // left padding of the k3 conv (2 causal, 1 otherwise): row shift of the shortcut's X rows
  int halo;              // reflected rows to write around Y
  int round_out;         // store TF32-rounded Y (its consumer is a single-pass TF32 conv)
};

template <int SPLIT>
__global__ void __launch_bounds__(R_THREADS, 1)
tc_res_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1h,
              const __grid_constant__ CUtensorMap map_w1l, const __grid_constant__ CUtensorMap map_wch,
              const __grid_constant__ CUtensorMap map_wcl, const ResArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sg = smem_raw + (sb - smem_u32(smem_raw));
  const uint32_t bar = sb + OFF_BAR;
  auto xr_full = [&](int r) { return bar + 8u * r; };             // TMA landed raw X in ring slot r
  auto xr_empty = [&](int r) { return bar + 8u * (3 + r); };      // GEMM 2 of the tile has read it
  auto p_ready = [&](int s) { return bar + 8u * (6 + s); };       // transform wrote x_lo / e / e_lo of stage s
  auto p_empty = [&](int s) { return bar + 8u * (8 + s); };       // both GEMMs of the tile have read stage s
  auto a1_full = [&](int b) { return bar + 8u * (10 + b); };      // GEMM 1 accumulator complete
  auto a1_empty = [&](int b) { return bar + 8u * (12 + b); };
  auto a2_full = [&](int b) { return bar + 8u * (14 + b); };      // GEMM 2 accumulator complete
  auto a2_empty = [&](int b) { return bar + 8u * (16 + b); };
  const uint32_t h_full = bar + 8u * 18;                          // H / H_lo written
  const uint32_t h_empty = bar + 8u * 19;                         // GEMM 2 has read them
  const uint32_t w_full = bar + 8u * 20;                          // weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sg + OFF_BAR + 8 * 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int r = 0; r < NRAW; ++r) {
      mbar_init(xr_full(r), 1);
      mbar_init(xr_empty(r), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(p_ready(s), 8);
      mbar_init(p_empty(s), 1);
      mbar_init(a1_full(s), 1);
      mbar_init(a1_empty(s), 4);
      mbar_init(a2_full(s), 1);
      mbar_init(a2_empty(s), 4);
    }
    mbar_init(h_full, 4);
    mbar_init(h_empty, 1);
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // H / H_lo: the hidden layer has 16 real channels; the upper 16 columns of the 32-wide operand stay zero for good
  for (int q = threadIdx.x; q < 2 * HT / 16; q += R_THREADS) reinterpret_cast<float4*>(sg + OFF_H)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: GEMM 1 accumulators [main 16 | corr 16] at 0 and 32, GEMM 2 accumulators [main 32 | corr 32] at 64 and 128
  const int n_my = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      mbar_expect_tx(w_full, SPLIT == 3 ? 3 * W1CH + 2 * WCCH : (3 * W1CH + 2 * WCCH) / 2);
      for (int c = 0; c < 3; ++c) {
        tma_load_2d(sb + OFF_W + c * W1CH, &map_w1h, w_full, c * RC, 0);
        if (SPLIT == 3) tma_load_2d(sb + OFF_W + c * W1CH + W1CH / 2, &map_w1l, w_full, c * RC, 0);
      }
      for (int c = 0; c < 2; ++c) {
        tma_load_2d(sb + OFF_W + 3 * W1CH + c * WCCH, &map_wch, w_full, c * RC, 0);
        if (SPLIT == 3) tma_load_2d(sb + OFF_W + 3 * W1CH + c * WCCH + WCCH / 2, &map_wcl, w_full, c * RC, 0);
      }
    }
    __syncwarp();
    for (int i = 0; i < n_my; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int mt = tile % p.tiles_m;
      const int item = tile / p.tiles_m;
      const int r = i % NRAW;
      mbar_wait(xr_empty(r), (((uint32_t)(i / NRAW)) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(xr_full(r), XT);
        tma_load_3d(sb + r * XT, &map_x, xr_full(r), 0, mt * RM + p.row_base, item);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc16 = umma_idesc_tf32(RM, RH);
    constexpr uint32_t idesc32 = umma_idesc_tf32(RM, RC);
    constexpr uint32_t idesc64 = umma_idesc_tf32(RM, 2 * RC);
    constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);
    auto mk = [](uint32_t addr) { return ((uint64_t)DESC_HI << 32) | (uint64_t)(((addr & 0x3FFFFu) >> 4) | (1u << 16)); };
    mbar_wait(w_full, 0);
    auto gemm1 = [&](int i) {
      const int s = i & 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(s * 32);
        const uint32_t e_hi = sb + OFF_P + s * P_STAGE + XT, e_lo = e_hi + XT;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint64_t da = mk(e_hi + j * 128), dl = mk(e_lo + j * 128), db = mk(sb + OFF_W + j * W1CH);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (SPLIT == 3) {
              tcgen05_mma_tf32(d, da + 2u * k, db + 2u * k, idesc32, (j > 0 || k > 0) ? 1u : 0u);   // [main | corr] (+)= e * [W1_hi | W1_lo]
              tcgen05_mma_tf32(d + RH, dl + 2u * k, db + 2u * k, idesc16, 1u);                      // corr += e_lo * W1_hi
            } else {
              tcgen05_mma_tf32(d, da + 2u * k, db + 2u * k, idesc16, (j > 0 || k > 0) ? 1u : 0u);   // main (+)= e * W1_hi
            }
          }
        }
        tcgen05_commit(a1_full(s));
      }
      __syncwarp();
    };
    auto gemm2 = [&](int i) {
      const int s = i & 1;
      const int r = i % NRAW;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(64 + s * 64);
        const uint32_t x_hi = sb + r * XT + p.pad_left * 128, x_lo = sb + OFF_P + s * P_STAGE + p.pad_left * 128;
        const uint64_t dh = mk(sb + OFF_H), dhl = mk(sb + OFF_H + HT), db0 = mk(sb + OFF_W + 3 * W1CH);
        const uint64_t dx = mk(x_hi), dxl = mk(x_lo), db1 = mk(sb + OFF_W + 3 * W1CH + WCCH);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tcgen05_mma_tf32(d, dh + 2u * k, db0 + 2u * k, SPLIT == 3 ? idesc64 : idesc32, k > 0 ? 1u : 0u);   // H * [W3_hi | W3_lo]
          if (SPLIT == 3) tcgen05_mma_tf32(d + RC, dhl + 2u * k, db0 + 2u * k, idesc32, 1u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tcgen05_mma_tf32(d, dx + 2u * k, db1 + 2u * k, SPLIT == 3 ? idesc64 : idesc32, 1u);                 // + X * [Ws_hi | Ws_lo]
          if (SPLIT == 3) tcgen05_mma_tf32(d + RC, dxl + 2u * k, db1 + 2u * k, idesc32, 1u);
        }
        tcgen05_commit(a2_full(s));
        tcgen05_commit(h_empty);
        tcgen05_commit(p_empty(s));
        tcgen05_commit(xr_empty(r));
      }
      __syncwarp();
    };
    // Two queues, served in whichever order their inputs arrive: GEMM 2 of tile i (needs the mid epilogue's H) and GEMM 1
    // of the next tile (needs the transform). A fixed order would make the transform of tile i+2 wait for GEMM 2 of tile
    // i+1 behind GEMM 1 of tile i+2, i.e. serialise transform and tensor work.
    int i1 = 0, i2 = 0;
    while (i2 < n_my) {
      bool go2 = false, go1 = false;
      if (lane == 0) {
        if (i2 < i1) go2 = mbar_test(h_full, (uint32_t)i2 & 1u) && mbar_test(a2_empty(i2 & 1), (((uint32_t)i2 >> 1) & 1u) ^ 1u);
        if (!go2 && i1 < n_my && i1 < i2 + 2)
          go1 = mbar_test(p_ready(i1 & 1), ((uint32_t)i1 >> 1) & 1u) && mbar_test(a1_empty(i1 & 1), (((uint32_t)i1 >> 1) & 1u) ^ 1u);
      }
      const int sel = __shfl_sync(0xffffffffu, go2 ? 2 : (go1 ? 1 : 0), 0);
      if (sel == 2) gemm2(i2++);
      else if (sel == 1) gemm1(i1++);
    }
  } else if (warp < 10) {
    // ================================ transform: x -> x_lo, e = ELU(x), e_lo ================================
    // eight warps: the transform is the longest stage of the per-tile chain (ELU + two operand splits per element)
    const int tt = threadIdx.x - 64;   // 0..255
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      const int r = i % NRAW;
      mbar_wait(xr_full(r), ((uint32_t)(i / NRAW)) & 1u);
      mbar_wait(p_empty(s), (((uint32_t)i >> 1) & 1u) ^ 1u);
      const float4* xr = reinterpret_cast<const float4*>(sg + r * XT);
      float4* xl = reinterpret_cast<float4*>(sg + OFF_P + s * P_STAGE);
      float4* eh = reinterpret_cast<float4*>(sg + OFF_P + s * P_STAGE + XT);
      float4* el = reinterpret_cast<float4*>(sg + OFF_P + s * P_STAGE + 2 * XT);
      auto split4 = [](const float4& a) {
        return make_float4(rn_tf32(a.x - trunc_tf32(a.x)), rn_tf32(a.y - trunc_tf32(a.y)), rn_tf32(a.z - trunc_tf32(a.z)),
                           rn_tf32(a.w - trunc_tf32(a.w)));
      };
      // three float4 per step: their 12 ELUs run interleaved (elu_vec) instead of as 12 serial dependency chains
#pragma unroll 1
      for (int q0 = tt; q0 < XT / 16; q0 += 3 * 256) {
        float4 v[3];
        float e[12];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int q = q0 + u * 256;
          v[u] = q < XT / 16 ? xr[q] : make_float4(0.f, 0.f, 0.f, 0.f);
          e[u * 4 + 0] = v[u].x; e[u * 4 + 1] = v[u].y; e[u * 4 + 2] = v[u].z; e[u * 4 + 3] = v[u].w;
        }
        elu_any<SPLIT == 1, 12>(e);   // SPLIT == 1: every ELU of this kernel feeds a TF32-rounded operand / output
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int q = q0 + u * 256;
          if (q < XT / 16) {
            const float4 ev = make_float4(e[u * 4 + 0], e[u * 4 + 1], e[u * 4 + 2], e[u * 4 + 3]);
            if (SPLIT == 3) {
              xl[q] = split4(v[u]);
              eh[q] = ev;
              el[q] = split4(ev);
            } else {
              eh[q] = make_float4(rn_tf32(ev.x), rn_tf32(ev.y), rn_tf32(ev.z), rn_tf32(ev.w));
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(s));
    }
  } else if (warp < 14) {
    // ================================ mid epilogue: GEMM 1 -> H = ELU(. + b1) as the A operand of GEMM 2 ================================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;   // tile row of this thread (= TMEM lane)
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      mbar_wait(a1_full(s), ((uint32_t)i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t vm[16], vc[16];
      tcgen05_ld16(lane_base + (uint32_t)(s * 32), vm);
      if (SPLIT == 3) tcgen05_ld16(lane_base + (uint32_t)(s * 32 + RH), vc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (SPLIT != 3) {
#pragma unroll
        for (int g = 0; g < 16; ++g) vc[g] = 0u;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_empty(s));
      if (i > 0) mbar_wait(h_empty, ((uint32_t)(i - 1)) & 1u);   // GEMM 2 of the previous tile has read H
      float4* hh = reinterpret_cast<float4*>(sg + OFF_H + r * 128);
      float4* hl = reinterpret_cast<float4*>(sg + OFF_H + HT + r * 128);
      float hv[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {     // 16 real hidden channels = the first four 16-byte chunks of the row
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1) + g);
        hv[g * 4 + 0] = __uint_as_float(vm[g * 4 + 0]) + __uint_as_float(vc[g * 4 + 0]) + bb.x;
        hv[g * 4 + 1] = __uint_as_float(vm[g * 4 + 1]) + __uint_as_float(vc[g * 4 + 1]) + bb.y;
        hv[g * 4 + 2] = __uint_as_float(vm[g * 4 + 2]) + __uint_as_float(vc[g * 4 + 2]) + bb.z;
        hv[g * 4 + 3] = __uint_as_float(vm[g * 4 + 3]) + __uint_as_float(vc[g * 4 + 3]) + bb.w;
      }
      elu_any<SPLIT == 1, 16>(hv);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 h = make_float4(hv[g * 4 + 0], hv[g * 4 + 1], hv[g * 4 + 2], hv[g * 4 + 3]);
        if (SPLIT == 3) {
          hh[g ^ (r & 7)] = h;
          hl[g ^ (r & 7)] = make_float4(rn_tf32(h.x - trunc_tf32(h.x)), rn_tf32(h.y - trunc_tf32(h.y)),
                                        rn_tf32(h.z - trunc_tf32(h.z)), rn_tf32(h.w - trunc_tf32(h.w)));
        } else {
          hh[g ^ (r & 7)] = make_float4(rn_tf32(h.x), rn_tf32(h.y), rn_tf32(h.z), rn_tf32(h.w));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(h_full);
    }
  } else {
    // ================================ final epilogue: GEMM 2 -> Y = ELU(. + b3 + bs) -> global ================================
    const int quad = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint8_t* slot = sg + OFF_STG + (warp - 14) * 2048;   // [32 rows x 64 B]
    for (int i = 0; i < n_my; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int mt = tile % p.tiles_m;
      const int item = tile / p.tiles_m;
      const int s = i & 1;
      mbar_wait(a2_full(s), ((uint32_t)i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t vm[32], vc[32];
      tcgen05_ld32(lane_base + (uint32_t)(64 + s * 64), vm);
      if (SPLIT == 3) tcgen05_ld32(lane_base + (uint32_t)(64 + s * 64 + RC), vc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (SPLIT != 3) {
#pragma unroll
        for (int g = 0; g < 32; ++g) vc[g] = 0u;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_empty(s));
      const int m_warp = mt * RM + quad * 32;
      const bool mirrors = p.halo > 0 && (m_warp <= p.halo || m_warp + 31 >= p.M - 1 - p.halo);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {       // 16 columns at a time
        float4* s0 = reinterpret_cast<float4*>(slot + lane * 64);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = hf * 16 + g * 4;
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bcat + c));
          s0[g ^ ((lane >> 1) & 3)] = make_float4(__uint_as_float(vm[c + 0]) + __uint_as_float(vc[c + 0]) + bb.x,
                                                  __uint_as_float(vm[c + 1]) + __uint_as_float(vc[c + 1]) + bb.y,
                                                  __uint_as_float(vm[c + 2]) + __uint_as_float(vc[c + 2]) + bb.z,
                                                  __uint_as_float(vm[c + 3]) + __uint_as_float(vc[c + 3]) + bb.w);
        }
        __syncwarp();
        // copy-out: 4 lanes per 64-byte row segment, ELU on the way, plus the reflected halo rows (conv.py:80-97)
        const int c16 = lane & 3;
        float* base = p.out + (long long)item * p.out_item_stride + hf * 16 + c16 * 4;
        float yv[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rr = (lane >> 2) + 8 * q;
          const float4 v = *reinterpret_cast<const float4*>(slot + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) << 4));
          yv[q * 4 + 0] = v.x; yv[q * 4 + 1] = v.y; yv[q * 4 + 2] = v.z; yv[q * 4 + 3] = v.w;
        }
        elu_any<SPLIT == 1, 16>(yv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rr = (lane >> 2) + 8 * q;
          const int m = m_warp + rr;
          if (m >= p.M) break;
          const float4 y = p.round_out ? make_float4(rn_tf32(yv[q * 4 + 0]), rn_tf32(yv[q * 4 + 1]), rn_tf32(yv[q * 4 + 2]), rn_tf32(yv[q * 4 + 3]))
                                       : make_float4(yv[q * 4 + 0], yv[q * 4 + 1], yv[q * 4 + 2], yv[q * 4 + 3]);
          *reinterpret_cast<float4*>(base + (long long)m * RC) = y;
          if (mirrors) {
            if (m >= 1 && m <= p.halo) *reinterpret_cast<float4*>(base - (long long)m * RC) = y;
            if (m <= p.M - 2 && m >= p.M - 1 - p.halo) *reinterpret_cast<float4*>(base + (2LL * (p.M - 1) - m) * RC) = y;
          }
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace

int launch_tc_res32(const TcResParams& p, cudaStream_t stream) {
  ECB_REQUIRE(p.split == 1 || p.split == 3, "tc_res: split must be 1 or 3");
  ECB_REQUIRE(p.x && p.out && p.w1_hi && p.w1_lo && p.wc_hi && p.wc_lo && p.b1 && p.bcat, "tc_res: null argument");
  ECB_REQUIRE(p.M > ACT_HALO && p.n_items > 0, "tc_res: bad M=%lld / items=%d", p.M, p.n_items);
  ECB_REQUIRE(p.pad_left >= 0 && p.pad_left <= 2 && p.x_first <= -p.pad_left, "tc_res: bad padding");
  CUtensorMap mx, w1h, w1l, wch, wcl;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)RC, (cuuint64_t)p.x_rows, (cuuint64_t)p.n_items};
    const cuuint64_t strides[2] = {(cuuint64_t)RC * 4, (cuuint64_t)p.x_item_stride * 4};
    const cuuint32_t box[3] = {RC, RROWS, 1};
    if (make_tensor_map(&mx, p.x, 3, dims, strides, box)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)3 * RC, (cuuint64_t)RC};
    const cuuint64_t strides[1] = {(cuuint64_t)3 * RC * 4};
    const cuuint32_t box[2] = {RC, RH};    // the 16 real hidden rows (rows 16..31 of the padded matrix are zero)
    if (make_tensor_map(&w1h, p.w1_hi, 2, dims, strides, box) || make_tensor_map(&w1l, p.w1_lo, 2, dims, strides, box)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)2 * RC, (cuuint64_t)RC};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * RC * 4};
    const cuuint32_t box[2] = {RC, RC};
    if (make_tensor_map(&wch, p.wc_hi, 2, dims, strides, box) || make_tensor_map(&wcl, p.wc_lo, 2, dims, strides, box)) return 1;
  }
  ResArgs a;
  a.b1 = p.b1;
  a.bcat = p.bcat;
  a.out = p.out;
  a.out_item_stride = p.out_item_stride;
  a.M = (int)p.M;
  a.n_items = p.n_items;
  a.tiles_m = (int)cdiv(p.M, RM);
  const long long total = (long long)a.tiles_m * p.n_items;
  ECB_REQUIRE(total < (1LL << 31), "tc_res: too many tiles");
  a.total_tiles = (int)total;
  a.row_base = (int)(-p.pad_left - p.x_first);
  a.pad_left = p.pad_left;
  a.halo = p.halo;
  a.round_out = p.round_out;
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    ECB_CUDA(cudaFuncSetAttribute(tc_res_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    ECB_CUDA(cudaFuncSetAttribute(tc_res_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    attr_set.mark();
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  const double rows = (double)p.M * p.n_items;
  ProfScope prof(PROF_TC_RES, stream, 2.0 * rows * (3 * RC * RH + RH * RC + RC * RC), 4.0 * rows * RC * 2);
  if (p.split == 3) tc_res_kernel<3><<<grid, R_THREADS, R_SMEM, stream>>>(mx, w1h, w1l, wch, wcl, a);
  else tc_res_kernel<1><<<grid, R_THREADS, R_SMEM, stream>>>(mx, w1h, w1l, wch, wcl, a);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
