// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (tensor-core path of the SEANet stacks).
//
// GEMM view (same as conv_gemm.cu): per item, out[M][N] = A[M][Ktot] * W + bias, where row m of A is the
// window of a channels-last activation [T][C] that starts at sample m*stride - pad_left and spans `taps`
// samples (reference modules/conv.py:202-221 SConv1d, :241-263 SConvTranspose1d as a 2-tap GEMM over frames,
// modules/lstm.py:24 input projection as a 1-tap GEMM), optionally followed by the columns of a second
// 1-tap source (the 1x1 shortcut of SEANetResnetBlock, modules/seanet.py:63-64, fused into the block's last
// conv).
//
// What is different from the CUDA-core kernel:
//   * activations live in HALO-PADDED buffers: HALO rows before and after each item hold the reflected
//     samples (conv.py:80-97 pad1d), written by the producer. Every A tile is then a plain box: one TMA
//     load (cp.async.bulk.tensor, SWIZZLE_128B) of [128 rows x 32 channels] per K chunk, straight into the
//     K-major canonical layout tcgen05.mma reads. A strided k = 2s conv addresses its input through the folded
//     view [T/s][s*C], so its window is two consecutive folded rows. Zero padding (transposed convs) is TMA
//     out-of-bounds fill.
//   * products run on the 5th-generation tensor cores: tcgen05.mma.cta_group::1, M = 128, N = BN, fp32 accumulators in
//     TMEM, issued by one elected thread; kind::tf32 (K = 8 per instruction) or kind::f16 (K = 16).
//   * three operand schemes (template parameter SPLIT, see Cfg): 2 = fp16 PAIR operands, the default of the fp32-accurate
//     layers (a = a1 + 2^-11 a2, w = w1 + 2^-11 w2; a1 w1 -> main accumulator, a1 w2 + a2 w1 -> correction accumulator);
//     3 = split TF32 (a = a_hi + a_lo with a_hi = the truncation the tensor core applies itself, a_lo = rn_tf32(a - a_hi);
//     the same three products at K = 8; the fp32-range alternative, ECB_F16_PAIR=0); 1 = one TF32 pass (operands rounded
//     to TF32 by their producers: the weight-norm decoder). The dropped lo x lo term is zero-mean and below 2^-22 relative.
//   * persistent, warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (owns TMEM), NT = 2 / 4 / 6 transform warps
//     (operand remainder or fp16 pair of every staged element; GroupNorm + ELU of the input when it is applied on load),
//     then 8 epilogue warps. Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
//     The epilogue reads TMEM (tcgen05.ld 32x32b), adds the bias, optionally applies ELU, stages 32-row blocks in swizzled
//     shared memory and writes them with coalesced 16-byte stores; an output may be written raw, ELU'd or both (a residual
//     block consumes x through its shortcut and ELU(x) through its first conv), its reflected halo rows are written
//     directly, and GroupNorm partial statistics are taken on the way.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ecb {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;                      // fp32 per K chunk = one 128-byte swizzle row
constexpr int TC_THREADS = 384;              // 12 warps: TMA, MMA, 2 x transform, 8 x epilogue (NT = 2; NT = 4: 14 warps)
constexpr int STAGING_BYTES = 8 * 2048;     // 8 epilogue warps x [32 rows x 64 B]

constexpr int A_ROWS = BM + 8;               // an A tile carries up to 8 extra rows: the taps of a conv are row shifts of it
constexpr int A_TILE = A_ROWS * BK * 4;      // 17 KB

constexpr int A16_TILE = A_ROWS * BK * 2;    // fp16 image of an A tile: rows of 64 bytes (SWIZZLE_64B)

// fp16 pair operands: how many operand tiles held a value that the saturating conversion clipped to +-65504 (a model whose
// activations leave the fp16 range must run with ECB_F16_PAIR=0). Read and cleared through ecb_f16_saturation_count.
__device__ unsigned int g_f16_sat_tiles = 0;

// SPLIT = 3: split-operand TF32 (a, a_lo; w_hi, w_lo). SPLIT = 1: one TF32 pass. SPLIT = 2: fp16 PAIR operands (the scheme of
// lstm_tc.cu): a = a1 + 2^-11 a2, w = w1 + 2^-11 w2 with fp16 a1, a2, w1, w2 -- the same three products and the same dropped
// 2^-22 term as SPLIT = 3 (fp16 carries TF32's 11 significand bits; its products are exact in the fp32 accumulator), but one
// tcgen05.mma.kind::f16 covers K = 16 instead of 8 and reads half the operand bytes: half the tensor time of the wide layers.
// fp16 spans 6e-5 .. 65504 at full precision, below that its subnormals (the remainder a2 keeps the absolute error at
// 2^-35); values beyond +-65504 saturate (satfinite conversion) -- see TcConvParams::split.
template <int BN, int SPLIT>
struct Cfg {
  static constexpr int B_BYTES = (SPLIT == 2) ? BN * BK * 2 : BN * BK * 4;
  static constexpr int A_STAGE = (SPLIT == 3) ? 2 * A_TILE : (SPLIT == 2 ? A_TILE + 2 * A16_TILE : A_TILE);   // a (+ a_lo | + a1, a2)
  static constexpr int B_STAGE = (SPLIT != 1) ? 2 * B_BYTES : B_BYTES;   // w_hi (+ w_lo, directly behind it) | w1, w2
  static constexpr int BUDGET = 227 * 1024 - STAGING_BYTES - 1024 - 512;
  // A ring: 3 stages with wide tiles, 4 otherwise; the B ring takes what is left (at most 8 stages)
  static constexpr int A_STAGES = (BN >= 128) ? 3 : 4;
  static constexpr int SB0 = (BUDGET - A_STAGES * A_STAGE) / B_STAGE;
  static constexpr int B_STAGES = SB0 > 8 ? 8 : SB0;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + B_STAGES * B_STAGE + STAGING_BYTES + 1024 + 512;
  static constexpr int TMEM_COLS = (SPLIT != 1) ? 4 * BN : 2 * BN;   // two main (+ two correction) accumulators
  static_assert(B_STAGES >= 2, "pipeline too shallow");
  static_assert(TMEM_COLS <= 512, "TMEM overflow");
};

struct TcArgs {
  const float* bias;        // [N] or nullptr
  float* out_raw;           // (item 0, row 0) of the raw output, or nullptr
  float* out_elu;           // same for the ELU output
  long long out_item_stride;
  int N, M, n_items;
  int tiles_m, tiles_n, total_tiles;
  int n_cb0, n_cb1;         // 32-channel column blocks of source 0 (of its folded row) / source 1: one A tile each
  int shifts;               // row shifts per source-0 tile (taps of a stride-1 conv, 2 folded rows of a strided one)
  int a_rows;               // rows per A tile box: BM (+ 8 when shifts > 1)
  int row_base0, row_base1; // row coordinate of output row 0's window start in each source's map
  int round_out, halo;
  int w_resident;           // 1: all weight chunks of the layer fit the B ring and the CTA keeps one N tile: load them once
  int lo_tma;               // 1: the A_lo tile is loaded by TMA (map_a0lo) instead of being computed by the transform warps
  int cell_on;              // 1: LSTM cell epilogue (cell) instead of the output stores
  TcCell cell;
  double* stats;            // [item][tiles_m][tiles_n][8 warps][2] or nullptr
  int group;                // K chunks per main-accumulator group (SPLIT == 1: all of them)
  const float* norm_mr;     // GroupNorm of source 0 on load (TcConvParams::norm_mr), or nullptr
  const float* norm_gamma;
  const float* norm_beta;
  int norm_elu;
  int C0;                   // channels of source 0 (before folding)
};

using namespace tc;

// Copy one staged [32 rows x 64 B] half block out of shared memory with coalesced 16-byte stores (four lanes per
// 64-byte row segment, 8 rows per warp instruction): raw output as is, ELU output through elu1. The output modes are
// template parameters so that the loop body is nothing but LDS -> (ELU) -> STG.
template <bool RAW, bool ELU, bool ROUND, bool STATS>
__device__ __forceinline__ void copy_out_rows(const uint8_t* slot, int lane, float* praw, float* pelu, long long row_step,
                                              int rows_left, float& st_sum, float& st_sq) {
  const int c16 = lane & 3;
  const int rr0 = lane >> 2;
  praw += (long long)rr0 * (row_step >> 3);
  pelu += (long long)rr0 * (row_step >> 3);
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = rr0 + 8 * i;
    v[i] = *reinterpret_cast<const float4*>(slot + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) << 4));
  }
  if (STATS) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (rr0 + 8 * i < rows_left) {
        st_sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        st_sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
  }
  if (RAW) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (rr0 + 8 * i < rows_left)
        *reinterpret_cast<float4*>(praw + i * row_step) =
            ROUND ? make_float4(rn_tf32(v[i].x), rn_tf32(v[i].y), rn_tf32(v[i].z), rn_tf32(v[i].w)) : v[i];
  }
  if (ELU) {
    // the 16 ELUs of this lane run interleaved (elu_vec): one elu1 is a chain of ~14 dependent operations
    float e[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[i * 4 + 0] = v[i].x; e[i * 4 + 1] = v[i].y; e[i * 4 + 2] = v[i].z; e[i * 4 + 3] = v[i].w;
    }
    elu_any<ROUND, 16>(e);   // ROUND: the stored value is TF32-rounded (single-pass consumers): the fast exp is exact enough
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (rr0 + 8 * i < rows_left) {
        float4 w = make_float4(e[i * 4 + 0], e[i * 4 + 1], e[i * 4 + 2], e[i * 4 + 3]);
        if (ROUND) w = make_float4(rn_tf32(w.x), rn_tf32(w.y), rn_tf32(w.z), rn_tf32(w.w));
        *reinterpret_cast<float4*>(pelu + i * row_step) = w;
      }
  }
}

// NT = transform warps (2 | 4 | 6; eight were measured slower on the 48 kHz model: 61 vs 56 ms of narrow convs per config-3 step). The narrow split-operand layers are bound by the transform role (remainder of every staged
// element, and affine + ELU when GroupNorm is applied on load): they run with four; 448 threads leave 146 registers each, which
// only the BN <= 64 instances fit.
template <int BN, int SPLIT, int NT>
__global__ void __launch_bounds__(TC_THREADS + 32 * (NT - 2), 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
               const __grid_constant__ CUtensorMap map_a0lo, const TcArgs p) {
  using C = Cfg<BN, SPLIT>;
  constexpr int SA = C::A_STAGES, SB = C::B_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: A ring [SA][a | a_lo] | B ring [SB][w_hi | w_lo] | epilogue staging | barriers
  const uint32_t a_ring = smem_base;
  const uint32_t b_ring = smem_base + SA * C::A_STAGE;
  constexpr int STAGING_OFF = SA * C::A_STAGE + SB * C::B_STAGE;
  const uint32_t bar_base = smem_base + STAGING_OFF + STAGING_BYTES;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };                 // TMA landed the A tile
  auto aready_bar = [&](int s) { return bar_base + 8u * (SA + s); };         // transform wrote a_lo
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * SA + s); };     // MMAs have read the A tile
  auto bfull_bar = [&](int s) { return bar_base + 8u * (3 * SA + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (3 * SA + SB + s); };
  auto mainf_bar = [&](int b) { return bar_base + 8u * (3 * SA + 2 * SB + b); };       // accumulator b holds a finished K group
  auto maine_bar = [&](int b) { return bar_base + 8u * (3 * SA + 2 * SB + 2 + b); };   // ... has been drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + STAGING_OFF + STAGING_BYTES + 8 * (3 * SA + 2 * SB + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // K chunk sequence of a tile: for every source-0 column block its `shifts` row shifts, then the source-1 blocks.
  // Chunk c reads A tile (group) ag at row shift j and weight columns [k0, k0 + 32).
  const int nch0 = p.n_cb0 * p.shifts;
  const int nch = nch0 + p.n_cb1;
  const int n_groups_a = p.n_cb0 + p.n_cb1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aready_bar(s), NT);   // one arrive per transform warp
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(mainf_bar(b), 1);
      mbar_init(maine_bar(b), 8);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns of accumulator buffer b: SPLIT == 3: [main | correction] = 2 BN columns, SPLIT == 1: BN columns
  constexpr int ACC_COLS = (SPLIT != 1) ? 2 * BN : BN;
  auto main_col = [&](int b) { return (uint32_t)(b * ACC_COLS); };

  if (warp == 0) {
    // ================================ TMA producer ================================
    // Every A tile ([a_rows x 32 channels]) is loaded ONCE per output tile; the taps of the conv read it at different
    // row offsets. Weight chunks stream through their own ring.
    {
      uint32_t ia = 0, ib = 0;
      const uint32_t a_bytes = (uint32_t)p.a_rows * BK * 4;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.tiles_n;
        const int mt_all = tile / p.tiles_n;
        const int mt = mt_all % p.tiles_m;
        const int item = mt_all / p.tiles_m;
        for (int ag = 0; ag < n_groups_a; ++ag, ++ia) {
          const int sa = (int)(ia % SA);
          mbar_wait(aempty_bar(sa), ((ia / SA) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(afull_bar(sa), (SPLIT == 3 && p.lo_tma) ? 2 * a_bytes : a_bytes);
            if (ag < p.n_cb0) {
              tma_load_3d(a_ring + sa * C::A_STAGE, &map_a0, afull_bar(sa), ag * BK, mt * BM + p.row_base0, item);
              if (SPLIT == 3 && p.lo_tma)   // the producer of a0 also wrote its TF32 remainder: no transform needed
                tma_load_3d(a_ring + sa * C::A_STAGE + A_TILE, &map_a0lo, afull_bar(sa), ag * BK, mt * BM + p.row_base0, item);
            } else
              tma_load_3d(a_ring + sa * C::A_STAGE, &map_a1, afull_bar(sa), (ag - p.n_cb0) * BK, mt * BM + p.row_base1, item);
          }
          __syncwarp();
          const int nj = ag < p.n_cb0 ? p.shifts : 1;
          for (int j = 0; j < nj; ++j, ++ib) {
            if (p.w_resident && ib >= (uint32_t)nch) continue;   // the weights of this CTA's only N tile are already resident
            const int sb = (int)(ib % SB);
            mbar_wait(bempty_bar(sb), ((ib / SB) & 1u) ^ 1u);
            const int k0 = ag < p.n_cb0 ? (j * p.n_cb0 + ag) * BK : (nch0 + (ag - p.n_cb0)) * BK;
            const uint32_t dst = b_ring + sb * C::B_STAGE;
            if (elect_one()) {
              mbar_expect_tx(bfull_bar(sb), C::B_STAGE);
              tma_load_2d(dst, &map_bhi, bfull_bar(sb), k0, nt * BN);
              if (SPLIT != 1) tma_load_2d(dst + C::B_BYTES, &map_blo, bfull_bar(sb), k0, nt * BN);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The tensor core truncates when it adds into the fp32 accumulator, a bias of ~1.7e-8 per accumulation step.
    // The large a*w_hi products therefore go to a MAIN accumulator and the small correction products (2^-11 of the
    // result: their truncation is irrelevant) to a CORRECTION accumulator in the adjacent TMEM columns; the epilogue
    // drains both every `group` K chunks and re-accumulates in registers (round-to-nearest), alternating between two
    // TMEM buffers. Per K step: D[main | corr] (+)= a * [w_hi | w_lo] (one MMA, N = 2 BN: w_lo's tile follows w_hi's
    // in shared memory) and D[corr] += a_lo * w_hi. The whole warp stays converged and one elected lane issues.
    {
      // kind::f16 instruction descriptor: D fp32, A / B fp16 (format 0), K-major
      constexpr uint32_t idesc = SPLIT == 2 ? ((1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24)) : umma_idesc_tf32(BM, BN);
      constexpr uint32_t idesc2 = SPLIT == 2 ? ((1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24))
                                             : umma_idesc_tf32(BM, SPLIT == 3 ? 2 * BN : BN);
      // descriptors: constant high word (SBO, version, swizzle mode) + start address >> 4 in the low word. fp16 tiles have rows of
      // 64 bytes: SWIZZLE_64B (layout type 4), 8-row groups 512 bytes apart
      constexpr uint32_t DESC_HI = SPLIT == 2 ? (32u | (1u << 14) | (4u << 29)) : (64u | (1u << 14) | (2u << 29));
      auto mk_desc = [](uint32_t addr, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)(((addr & 0x3FFFFu) >> 4) | (1u << 16)); };
      uint32_t ia = 0, ib = 0, gcount = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int c = 0;         // chunk index inside the tile
        int in_group = 0;  // chunks accumulated into the current accumulator
        for (int ag = 0; ag < n_groups_a; ++ag, ++ia) {
          const int sa = (int)(ia % SA);
          mbar_wait(afull_bar(sa), (ia / SA) & 1u);
          if (SPLIT != 1) mbar_wait(aready_bar(sa), (ia / SA) & 1u);
          const uint32_t a_addr = a_ring + sa * C::A_STAGE;
          const int nj = ag < p.n_cb0 ? p.shifts : 1;
          for (int j = 0; j < nj; ++j, ++ib, ++c) {
            if (in_group == 0) mbar_wait(maine_bar((int)(gcount & 1u)), ((gcount >> 1) & 1u) ^ 1u);   // accumulator drained
            const uint32_t d_main = tmem_base + main_col((int)(gcount & 1u));
            const int sb = p.w_resident ? c : (int)(ib % SB);
            mbar_wait(bfull_bar(sb), p.w_resident ? 0u : ((ib / SB) & 1u));   // resident: phase 0 completed once, for good
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool close_group = (in_group + 1 == p.group) || (c + 1 == nch);
            if (elect_one()) {
              // tap j of the conv = the A tile read j rows further down: the start moves by j * 128 bytes. The 128-byte
              // swizzle is a function of the absolute shared-memory address (verified on B200: with the descriptor's
              // base-offset field left 0 every row shift reads the rows TMA wrote), so nothing else changes.
              const uint64_t db0 = mk_desc(b_ring + sb * C::B_STAGE, DESC_HI);
              if (SPLIT == 2) {
                // fp16 pair: rows of 64 bytes (a tap = 64 bytes further), K = 16 per instruction: two steps per 32-channel chunk.
                // D[main | corr] (+)= a1 [w1 | w2], D[corr] += a2 w1; the epilogue adds corr 2^-11.
                const uint32_t a_sh = a_addr + A_TILE + (uint32_t)j * 64u;
                const uint64_t da1 = mk_desc(a_sh, DESC_HI);
                const uint64_t da2 = mk_desc(a_sh + A16_TILE, DESC_HI);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  tcgen05_mma_f16(d_main, da1 + 2u * k, db0 + 2u * k, idesc2, (in_group > 0 || k > 0) ? 1u : 0u);
                  tcgen05_mma_f16(d_main + BN, da2 + 2u * k, db0 + 2u * k, idesc, 1u);
                }
              } else {
              const uint32_t a_sh = a_addr + (uint32_t)j * 128u;
              const uint64_t da0 = mk_desc(a_sh, DESC_HI);
              const uint64_t dal0 = mk_desc(a_sh + A_TILE, DESC_HI);
#pragma unroll
              for (int k = 0; k < BK / 8; ++k) {
                tcgen05_mma_tf32(d_main, da0 + 2u * k, db0 + 2u * k, idesc2, (in_group > 0 || k > 0) ? 1u : 0u);
                if (SPLIT == 3) tcgen05_mma_tf32(d_main + BN, dal0 + 2u * k, db0 + 2u * k, idesc, 1u);
              }
              }
              if (!p.w_resident) tcgen05_commit(bempty_bar(sb));               // weight stage free once these MMAs have read it
              if (close_group) tcgen05_commit(mainf_bar((int)(gcount & 1u)));  // K group complete -> epilogue
              if (j + 1 == nj) tcgen05_commit(aempty_bar(sa));                 // A tile free once all its taps are done
            }
            __syncwarp();
            ++in_group;
            if (close_group) {
              ++gcount;
              in_group = 0;
            }
          }
        }
      }
    }
  } else if (warp < 2 + NT) {
    // ================================ transform: a_lo = rn_tf32(a - trunc_tf32(a)) ================================
    if (SPLIT != 1) {
      constexpr int NTT = 32 * NT;      // transform threads
      const int tt = threadIdx.x - 64;  // 0..NTT-1
      const int n4 = p.a_rows * (BK / 4);   // float4 per A tile (a multiple of 64)
      uint32_t ia = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        float nm_mean = 0.f, nm_rstd = 0.f;
        if (p.norm_mr) {
          const int item = (tile / p.tiles_n) / p.tiles_m;
          nm_mean = __ldg(p.norm_mr + 2 * item);
          nm_rstd = __ldg(p.norm_mr + 2 * item + 1);
        }
        for (int ag = 0; ag < n_groups_a; ++ag, ++ia) {
          const int sa = (int)(ia % SA);
          mbar_wait(afull_bar(sa), (ia / SA) & 1u);
          const float4* a = reinterpret_cast<const float4*>(smem_gen + sa * C::A_STAGE);
          float4* alo = reinterpret_cast<float4*>(smem_gen + sa * C::A_STAGE + A_TILE);
          if (p.lo_tma) {   // a_lo arrived by TMA together with a
            __syncwarp();
            if (lane == 0) mbar_arrive(aready_bar(sa));
            continue;
          }
          // SPLIT == 2: the fp16 pair of one float4 (four channels of row i >> 3): a1 = fp16(v) (saturating), a2 = fp16((v - a1) 2^11),
          // stored at the SWIZZLE_64B position of logical 8-byte slot lc8 of the 64-byte row
          uint8_t* a16 = smem_gen + sa * C::A_STAGE + A_TILE;
          __half2 hmax = __float2half2_rn(0.f);   // largest |a1| this thread wrote to the tile: 65504 = the conversion saturated
          auto store_pair = [&](int i, const float4& v) {
            const int r = i >> 3;
            const int lc8 = (i & 7) ^ (r & 7);
            uint32_t h01, h23, l01, l23;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(v.y), "f"(v.x));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(v.w), "f"(v.z));
            const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
            const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l01) : "f"((v.y - f01.y) * 2048.f), "f"((v.x - f01.x) * 2048.f));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l23) : "f"((v.w - f23.y) * 2048.f), "f"((v.z - f23.x) * 2048.f));
            const uint32_t off = (uint32_t)r * 64u + ((uint32_t)((lc8 >> 1) ^ ((r >> 1) & 3)) << 4) + ((uint32_t)(lc8 & 1) << 3);
            hmax = __hmax2(hmax, __hmax2(__habs2(*reinterpret_cast<const __half2*>(&h01)), __habs2(*reinterpret_cast<const __half2*>(&h23))));
            *reinterpret_cast<uint2*>(a16 + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(a16 + A16_TILE + off) = make_uint2(l01, l23);
          };
          if (p.norm_mr && ag < p.n_cb0) {
            // GroupNorm on load: a <- act(((a - mean) * rstd) * gamma + beta) in place (the arithmetic of gn_apply, misc.cu),
            // then the remainder. Thread tt always meets the same logical 16-byte chunk of a row: its float4s are 8 or 16 rows
            // apart (NTT threads x 16 B = NTT / 8 rows of 128 B) and the 128-byte swizzle depends on the row only through row & 7.
            const float mean = nm_mean, rstd = nm_rstd;
            const int lc = (tt & 7) ^ ((tt >> 3) & 7);
            const int c0 = (ag * BK) % p.C0 + lc * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.norm_gamma + c0));
            const float4 be = __ldg(reinterpret_cast<const float4*>(p.norm_beta + c0));
            float4* aw = reinterpret_cast<float4*>(smem_gen + sa * C::A_STAGE);
            for (int i0 = tt; i0 < n4; i0 += 4 * NTT) {
              float e[16];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = i0 + NTT * u;
                const float4 v = i < n4 ? aw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                e[4 * u + 0] = (v.x - mean) * rstd * g.x + be.x;
                e[4 * u + 1] = (v.y - mean) * rstd * g.y + be.y;
                e[4 * u + 2] = (v.z - mean) * rstd * g.z + be.z;
                e[4 * u + 3] = (v.w - mean) * rstd * g.w + be.w;
              }
              if (p.norm_elu) elu_vec<16>(e);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = i0 + NTT * u;
                if (i < n4 && SPLIT == 2) {
                  store_pair(i, make_float4(e[4 * u + 0], e[4 * u + 1], e[4 * u + 2], e[4 * u + 3]));
                } else if (i < n4) {
                  aw[i] = make_float4(e[4 * u + 0], e[4 * u + 1], e[4 * u + 2], e[4 * u + 3]);
                  alo[i] = make_float4(rn_tf32(e[4 * u + 0] - trunc_tf32(e[4 * u + 0])), rn_tf32(e[4 * u + 1] - trunc_tf32(e[4 * u + 1])),
                                       rn_tf32(e[4 * u + 2] - trunc_tf32(e[4 * u + 2])), rn_tf32(e[4 * u + 3] - trunc_tf32(e[4 * u + 3])));
                }
              }
            }
            if (SPLIT == 2 && __hge(__hmax(__low2half(hmax), __high2half(hmax)), __ushort_as_half((unsigned short)0x7BFF)))
              atomicAdd(&g_f16_sat_tiles, 1u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(aready_bar(sa));
            continue;
          }
          if (SPLIT == 2) {
            // thread tt meets the same swizzled slot of every row it visits (its rows are NTT / 8 = 8 or 16 apart, the swizzles
            // depend on row & 7): the store address just advances by a constant
            const int r0 = tt >> 3;
            const int lc8 = (tt & 7) ^ (r0 & 7);
            uint8_t* d = a16 + r0 * 64 + ((((lc8 >> 1) ^ ((r0 >> 1) & 3)) << 4) + ((lc8 & 1) << 3));
#pragma unroll 4
            for (int i = tt; i < n4; i += NTT, d += (NTT / 8) * 64) {
              const float4 v = a[i];
              uint32_t h01, h23, l01, l23;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(v.y), "f"(v.x));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(v.w), "f"(v.z));
              const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
              const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l01) : "f"((v.y - f01.y) * 2048.f), "f"((v.x - f01.x) * 2048.f));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l23) : "f"((v.w - f23.y) * 2048.f), "f"((v.z - f23.x) * 2048.f));
              *reinterpret_cast<uint2*>(d) = make_uint2(h01, h23);
              *reinterpret_cast<uint2*>(d + A16_TILE) = make_uint2(l01, l23);
              hmax = __hmax2(hmax, __hmax2(__habs2(*reinterpret_cast<const __half2*>(&h01)), __habs2(*reinterpret_cast<const __half2*>(&h23))));
            }
            if (__hge(__hmax(__low2half(hmax), __high2half(hmax)), __ushort_as_half((unsigned short)0x7BFF))) atomicAdd(&g_f16_sat_tiles, 1u);
          } else {
#pragma unroll 4
          for (int i = tt; i < n4; i += NTT) {
            const float4 v = a[i];
            float4 r;
            r.x = rn_tf32(v.x - trunc_tf32(v.x));
            r.y = rn_tf32(v.y - trunc_tf32(v.y));
            r.z = rn_tf32(v.z - trunc_tf32(v.z));
            r.w = rn_tf32(v.w - trunc_tf32(v.w));
            alo[i] = r;   // same offset => same swizzled position
          }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(aready_bar(sa));
        }
      }
    }
  } else {
    // ================================ epilogue: TMEM -> registers -> staging slot -> coalesced global stores ================================
    // Eight warps: warp w reads TMEM lane quadrant w % 4 (hardware rule) and, of every 32-column block, the 16-column
    // half (w - 4) / 4. The element-wise work (bias, ELU, optional TF32 rounding) is ~20 instructions per output
    // element, which is what sizes this warp group.
    constexpr int EW0 = 2 + NT;   // first epilogue warp; the eight of them cover every (lane quadrant, column half) pair once
    const int quad = warp & 3;
    const int half = (warp - EW0) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 16);
    uint8_t* slot_gen = smem_gen + STAGING_OFF + (warp - EW0) * 2048;   // this warp's [32 rows x 64 B] staging slot
    uint32_t gcount = 0;
    const int out_mode = (p.out_raw ? 1 : 0) | (p.out_elu ? 2 : 0) | (p.round_out ? 4 : 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n;
      const int mt_all = tile / p.tiles_n;
      const int mt = mt_all % p.tiles_m;
      const int item = mt_all / p.tiles_m;
      const int m_warp = mt * BM + quad * 32;       // first output row of this warp
      // does this warp hold rows whose reflected copies (conv.py:80-97) must be written too?
      const bool mirrors = p.halo > 0 && (m_warp <= p.halo || m_warp + 31 >= p.M - 1 - p.halo);
      float st_sum = 0.f, st_sq = 0.f;   // GroupNorm partial statistics of this warp's part of the tile (p.stats only)

      // One 16-column half block: the bias-added values go to this warp's staging slot in a chunk-swizzled layout, then
      // the warp copies the slot out with coalesced 16-byte stores (four lanes per 64-byte row segment): the raw
      // output as is, the ELU output through elu1 in the same pass, plus the reflected halo rows where they exist.
      auto finish_block = [&](const float* o, int cc) {
        const int n0 = nt * BN + cc + half * 16;
        __syncwarp();   // the previous block's copy-out is done with the slot
        float4* s0 = reinterpret_cast<float4*>(slot_gen + lane * 64);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 v = make_float4(o[g * 4 + 0], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
          if (p.bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + g * 4));
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          }
          s0[g ^ ((lane >> 1) & 3)] = v;
        }
        __syncwarp();
        const long long base = (long long)item * p.out_item_stride + (long long)m_warp * p.N + n0 + (lane & 3) * 4;
        const int rows_left = p.M - m_warp;
        if (!mirrors) {
          const long long step = 8LL * p.N;
          float* praw = p.out_raw + base;   // only dereferenced when the mode says so
          float* pelu = p.out_elu + base;
          switch (out_mode) {
            case 1:
              if (p.stats) copy_out_rows<true, false, false, true>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq);
              else copy_out_rows<true, false, false, false>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq);
              break;
            case 2: copy_out_rows<false, true, false, false>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq); break;
            case 3: copy_out_rows<true, true, false, false>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq); break;
            case 5: copy_out_rows<true, false, true, false>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq); break;
            case 6: copy_out_rows<false, true, true, false>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq); break;
            default: copy_out_rows<true, true, true, false>(slot_gen, lane, praw, pelu, step, rows_left, st_sum, st_sq); break;
          }
        } else {
          // first / last rows of an item: also write the reflected copies (conv.py:80-97); rare, generic loop
          const int c16 = lane & 3;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int rr = (lane >> 2) + 8 * i;
            const int m = m_warp + rr;
            if (rr >= rows_left) break;
            const float4 v = *reinterpret_cast<const float4*>(slot_gen + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) << 4));
            long long d1 = 0, d2 = 0;
            if (m >= 1 && m <= p.halo) d1 = -2LL * m * p.N;
            if (m <= p.M - 2 && m >= p.M - 1 - p.halo) d2 = 2LL * (p.M - 1 - m) * p.N;
            const long long off = base + (long long)rr * p.N;
            if (p.out_raw) {
              const float4 w = p.round_out ? make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w)) : v;
              *reinterpret_cast<float4*>(p.out_raw + off) = w;
              if (d1) *reinterpret_cast<float4*>(p.out_raw + off + d1) = w;
              if (d2) *reinterpret_cast<float4*>(p.out_raw + off + d2) = w;
            }
            if (p.out_elu) {
              float4 w = make_float4(elu1(v.x), elu1(v.y), elu1(v.z), elu1(v.w));
              if (p.round_out) w = make_float4(rn_tf32(w.x), rn_tf32(w.y), rn_tf32(w.z), rn_tf32(w.w));
              *reinterpret_cast<float4*>(p.out_elu + off) = w;
              if (d1) *reinterpret_cast<float4*>(p.out_elu + off + d1) = w;
              if (d2) *reinterpret_cast<float4*>(p.out_elu + off + d2) = w;
            }
          }
        }
      };

      // LSTM cell epilogue (step-wise recurrence): this thread's 16 columns are the four gates of four units of item
      // m_warp + lane; everything else of the step happens here, in registers (TcCell, common.cuh).
      auto finish_cell = [&](const float* o, int cc) {
        const int n0 = nt * BN + cc + half * 16;
        const int u0 = (n0 >> 4) << 2;
        const int m = m_warp + lane;
        if (m >= p.M) return;
        const TcCell& q = p.cell;
        const float* pr = q.pre + (long long)m * q.pre_stride + u0;
        const float4 pi = __ldg(reinterpret_cast<const float4*>(pr));
        const float4 pf = __ldg(reinterpret_cast<const float4*>(pr + q.H));
        const float4 pg = __ldg(reinterpret_cast<const float4*>(pr + 2 * q.H));
        const float4 po = __ldg(reinterpret_cast<const float4*>(pr + 3 * q.H));
        float4 cv = *reinterpret_cast<const float4*>(q.c + (long long)m * q.H + u0);
        float4 hv;
        auto sig = [](float x) { return 1.f / (1.f + expf(-x)); };
#define ECB_CELL(f, j)                                                                                         \
  {                                                                                                            \
    const float gi = sig(pi.f + o[0 + j]), gf = sig(pf.f + o[4 + j]), gg = tanhf(pg.f + o[8 + j]), go = sig(po.f + o[12 + j]); \
    cv.f = gf * cv.f + gi * gg;                                                                                \
    hv.f = go * tanhf(cv.f);                                                                                   \
  }
        ECB_CELL(x, 0) ECB_CELL(y, 1) ECB_CELL(z, 2) ECB_CELL(w, 3)
#undef ECB_CELL
        *reinterpret_cast<float4*>(q.c + (long long)m * q.H + u0) = cv;
        *reinterpret_cast<float4*>(q.h_out + (long long)m * q.H + u0) = hv;
        if (q.h_lo_out)
          *reinterpret_cast<float4*>(q.h_lo_out + (long long)m * q.H + u0) =
              make_float4(rn_tf32(hv.x - trunc_tf32(hv.x)), rn_tf32(hv.y - trunc_tf32(hv.y)), rn_tf32(hv.z - trunc_tf32(hv.z)),
                          rn_tf32(hv.w - trunc_tf32(hv.w)));
        float4 y = hv;
        if (q.skip) {
          const float4 sv = __ldg(reinterpret_cast<const float4*>(q.skip + (long long)m * q.skip_stride + u0));
          y.x += sv.x; y.y += sv.y; y.z += sv.z; y.w += sv.w;
        }
        if (q.out_elu) y = make_float4(elu1(y.x), elu1(y.y), elu1(y.z), elu1(y.w));
        *reinterpret_cast<float4*>(q.out + (long long)m * q.out_stride + u0) = y;
      };

      // Where the finished fp32 tile is read from: TMEM columns src_col (+ the correction columns BN further when
      // they still have to be added), this warp's 16 columns of every 32-column block.
      const int n_groups = (nch + p.group - 1) / p.group;
      const int mb_last = (int)((gcount + (uint32_t)n_groups - 1u) & 1u);
      const uint32_t src_col = main_col(mb_last);
      bool add_corr = (SPLIT != 1);
      if (SPLIT != 1 && n_groups > 1) {
        // several K groups: re-accumulate main + correction in registers (round-to-nearest), park the sum back in the
        // last group's TMEM buffer and stream it out from there
        float acc[BN / 2];
        for (int g = 0; g < n_groups; ++g, ++gcount) {
          const int mb = (int)(gcount & 1u);
          mbar_wait(mainf_bar(mb), (gcount >> 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int cc = 0; cc < BN; cc += 32) {
            uint32_t v[16], w[16];
            tcgen05_ld16(lane_base + main_col(mb) + (uint32_t)cc, v);
            tcgen05_ld16(lane_base + main_col(mb) + (uint32_t)(BN + cc), w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = SPLIT == 2 ? fmaf(__uint_as_float(w[i]), 1.f / 2048.f, __uint_as_float(v[i]))
                                         : __uint_as_float(v[i]) + __uint_as_float(w[i]);
              acc[cc / 2 + i] = g == 0 ? t : acc[cc / 2 + i] + t;
            }
          }
          if (g + 1 < n_groups) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(maine_bar(mb));
          }
        }
#pragma unroll
        for (int cc = 0; cc < BN; cc += 32) {
          uint32_t v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(acc[cc / 2 + i]);
          tcgen05_st16(lane_base + src_col + (uint32_t)cc, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        add_corr = false;
      } else {
        mbar_wait(mainf_bar(mb_last), (gcount >> 1) & 1u);
        ++gcount;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        uint32_t v[16];
        float o[16];
        tcgen05_ld16(lane_base + src_col + (uint32_t)cc, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(v[i]);
        if (add_corr) {
          tcgen05_ld16(lane_base + src_col + (uint32_t)(BN + cc), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = SPLIT == 2 ? fmaf(__uint_as_float(v[i]), 1.f / 2048.f, o[i]) : o[i] + __uint_as_float(v[i]);
        }
        if (cc + 32 >= BN) {   // accumulator fully read: hand it back before the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(maine_bar(mb_last));
        }
        if (p.cell_on) finish_cell(o, cc);
        else finish_block(o, cc);
      }
      if (p.stats) {
        // deterministic two-stage reduction: per-warp partials here, summed per item in gn_apply (misc.cu)
        double ds = (double)st_sum, dq = (double)st_sq;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          ds += __shfl_xor_sync(0xffffffffu, ds, off);
          dq += __shfl_xor_sync(0xffffffffu, dq, off);
        }
        if (lane == 0) {
          double* sp = p.stats + (((long long)item * p.tiles_m + mt) * p.tiles_n + nt) * 16 + (warp - EW0) * 2;
          sp[0] = ds;
          sp[1] = dq;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// w [Ktot][N_src] (the CUDA-core packing, N contiguous) -> hi = rn_tf32(w), lo = rn_tf32(w - hi), both [N_pad][K_pad]
// (K contiguous: the K-major B operand); rows n >= N_src and columns k >= Ktot are zero.
__global__ void split_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int K,
                                     int N, int K_pad, int N_pad) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? w[(long long)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N_pad && k < K_pad) {
      const float v = tile[threadIdx.x][i];
      const float h = rn_tf32(v);
      hi[(long long)n * K_pad + k] = h;
      lo[(long long)n * K_pad + k] = rn_tf32(v - h);
    }
  }
}

// the fp16 pair of the same weights: h1 = fp16(w), h2 = fp16((w - h1) 2^11), both [N_pad][K_pad] halves, K contiguous
__global__ void split_weights_f16_kernel(const float* __restrict__ w, __half* __restrict__ h1, __half* __restrict__ h2, int K, int N,
                                         int K_pad, int N_pad) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? w[(long long)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N_pad && k < K_pad) {
      const float v = tile[threadIdx.x][i];
      const __half a = __float2half_rn(v);
      h1[(long long)n * K_pad + k] = a;
      h2[(long long)n * K_pad + k] = __float2half_rn((v - __half2float(a)) * 2048.f);
    }
  }
}

template <int BN, int SPLIT, int NT = 2>
int launch_one(const CUtensorMap* maps, const TcArgs& a, int grid, cudaStream_t stream) {
  using C = Cfg<BN, SPLIT>;
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    ECB_CUDA(cudaFuncSetAttribute(tc_conv_kernel<BN, SPLIT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set.mark();
  }
  TcArgs b = a;
  {
    // small layers: every weight chunk fits the B ring, and with a single N tile each CTA needs the same ones for all
    // of its tiles -> load them once (saves the TMA refill and the L2 reads of the weights per output tile)
    static int off = -1;
    if (off < 0) {
      const char* e = getenv("ECB_TC_WRES");   // diagnostic: ECB_TC_WRES=0 reloads the weights for every tile
      off = (e && e[0] == '0') ? 1 : 0;
    }
    const int nch = a.n_cb0 * a.shifts + a.n_cb1;
    b.w_resident = (!off && a.tiles_n == 1 && nch <= C::B_STAGES) ? 1 : 0;
  }
  tc_conv_kernel<BN, SPLIT, NT><<<grid, TC_THREADS + 32 * (NT - 2), C::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], b);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

int make_tensor_map(CUtensorMap* map, const float* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  ECB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECB_REQUIRE(r == CUDA_SUCCESS,
              "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, base %p, dims %llu x %llu x %llu, strides %llu, %llu)",
              (int)r, rank, (const void*)base, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0],
              (unsigned long long)(rank > 2 ? strides_bytes[1] : 0));
  return 0;
}

// same for fp16 tensors (lstm_tc.cu: recurrent state / weight slices), SWIZZLE_128B
int make_tensor_map_f16(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  ECB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed with CUresult %d (rank %d, base %p)", (int)r, rank, base);
  return 0;
}

// SM count of the current device (cached per device: a process may drive several GPUs)
int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}


int launch_split_weights(const float* w, float* hi, float* lo, int K, int N, int K_pad, int N_pad, cudaStream_t s) {
  dim3 grid((unsigned)cdiv(K_pad, 32), (unsigned)cdiv(N_pad, 32));
  split_weights_kernel<<<grid, dim3(32, 8), 0, s>>>(w, hi, lo, K, N, K_pad, N_pad);
  ECB_LAUNCHED();
  return 0;
}

// number of operand tiles (since the last reset) in which the fp16 pair conversion saturated; synchronises the device
long long tc_f16_saturation_count(int reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_f16_sat_tiles, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned int z = 0;
    cudaMemcpyToSymbol(g_f16_sat_tiles, &z, sizeof(z));
  }
  return (long long)v;
}

int launch_split_weights_f16(const float* w, void* h1, void* h2, int K, int N, int K_pad, int N_pad, cudaStream_t s) {
  dim3 grid((unsigned)cdiv(K_pad, 32), (unsigned)cdiv(N_pad, 32));
  split_weights_f16_kernel<<<grid, dim3(32, 8), 0, s>>>(w, reinterpret_cast<__half*>(h1), reinterpret_cast<__half*>(h2), K, N, K_pad, N_pad);
  ECB_LAUNCHED();
  return 0;
}

int tc_pick_bn(int N, int split, int bn_max) {
  int bn = split != 1 ? (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32))
                      : (N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32)));
  while (bn_max > 0 && bn > bn_max && bn > 32) bn /= 2;
  return bn;
}

int tc_stat_slots(const TcConvParams& p) {
  return (int)(cdiv(p.M, BM) * (p.N / tc_pick_bn(p.N, p.split, p.bn_max)) * 8);
}

int launch_tc_conv(const TcConvParams& p, cudaStream_t stream) {
  ECB_REQUIRE(p.split == 1 || p.split == 2 || p.split == 3, "tc_conv: split must be 1, 2 or 3");
  ECB_REQUIRE(p.N % 32 == 0 && p.N > 0, "tc_conv: N=%d must be a multiple of 32", p.N);
  ECB_REQUIRE(p.C0 % 32 == 0 && p.taps >= 1 && p.stride >= 1, "tc_conv: C0=%d must be a multiple of 32", p.C0);
  ECB_REQUIRE(p.a1 == nullptr || p.C1 % 32 == 0, "tc_conv: C1=%d must be a multiple of 32", p.C1);
  ECB_REQUIRE(p.M > 0 && p.n_items > 0, "tc_conv: bad M=%lld / items=%d", p.M, p.n_items);
  ECB_REQUIRE(p.out_raw || p.out_elu || p.cell, "tc_conv: no output");
  ECB_REQUIRE(!p.cell || (p.n_items == 1 && p.N == 4 * p.cell->H && !p.stats && p.halo == 0), "tc_conv: bad LSTM cell epilogue setup");
  ECB_REQUIRE(p.halo == 0 || p.M > p.halo, "tc_conv: %lld rows are too few for a %d-row reflected halo", p.M, p.halo);
  ECB_REQUIRE(!p.stats || (p.out_raw && !p.out_elu && p.halo == 0 && !p.round_out), "tc_conv: statistics need a plain raw output");
  const int bn = tc_pick_bn(p.N, p.split, p.bn_max);
  const int s = p.stride;
  const int ktot = p.taps * p.C0 + (p.a1 ? p.C1 : 0);
  ECB_REQUIRE(p.taps % s == 0, "tc_conv: kernel size %d must be a multiple of the stride %d", p.taps, s);
  CUtensorMap maps[5];
  TcArgs a;
  ECB_REQUIRE(!p.a0_lo || (p.split == 3 && !p.a1), "tc_conv: a0_lo needs split == 3 and a single source");
  ECB_REQUIRE(!p.norm_mr || (p.split != 1 && !p.a0_lo && p.norm_gamma && p.norm_beta),
              "tc_conv: normalise-on-load needs split operands, no a0_lo, gamma and beta");
  ECB_REQUIRE(p.split != 2 || !p.cell, "tc_conv: the LSTM cell epilogue runs with TF32 operands");
  a.norm_mr = p.norm_mr;
  a.norm_gamma = p.norm_gamma;
  a.norm_beta = p.norm_beta;
  a.norm_elu = p.norm_elu;
  a.C0 = p.C0;
  a.lo_tma = p.a0_lo ? 1 : 0;
  // the taps of the conv are row shifts of one [a_rows x 32] tile per 32-channel block of the (folded) input row
  a.shifts = p.taps / s;
  a.n_cb0 = s * p.C0 / BK;
  a.n_cb1 = p.a1 ? p.C1 / BK : 0;
  a.a_rows = a.shifts > 1 ? A_ROWS : BM;
  ECB_REQUIRE(a.shifts - 1 <= A_ROWS - BM, "tc_conv: %d taps exceed the %d-row tile halo", a.shifts, A_ROWS - BM);
  {
    // folded view of source 0: rows of s*C0 floats; the fold is aligned so that output row 0's window starts a folded
    // row: first mapped sample b0 = smallest sample >= a0_first congruent to -pad_left (mod s)
    long long d = (-(long long)p.pad_left - p.a0_first) % s;
    if (d < 0) d += s;
    const long long b0 = p.a0_first + d;
    const long long rows = (p.a0_first + p.a0_rows - b0) / s;
    ECB_REQUIRE(rows > 0, "tc_conv: empty source");
    a.row_base0 = (int)((-(long long)p.pad_left - b0) / s);   // exact; may be negative (TMA zero fill)
    const cuuint64_t dims[3] = {(cuuint64_t)s * p.C0, (cuuint64_t)rows, (cuuint64_t)p.n_items};
    const cuuint64_t strides[2] = {(cuuint64_t)s * p.C0 * 4, (cuuint64_t)p.a0_item_stride * 4};
    const cuuint32_t box[3] = {BK, (cuuint32_t)a.a_rows, 1};
    if (make_tensor_map(&maps[0], p.a0 + d * p.C0, 3, dims, strides, box)) return 1;
    if (p.a0_lo) {
      if (make_tensor_map(&maps[4], p.a0_lo + d * p.C0, 3, dims, strides, box)) return 1;
    } else {
      maps[4] = maps[0];
    }
  }
  if (p.a1) {
    const cuuint64_t dims[3] = {(cuuint64_t)p.C1, (cuuint64_t)p.a1_rows, (cuuint64_t)p.n_items};
    const cuuint64_t strides[2] = {(cuuint64_t)p.C1 * 4, (cuuint64_t)p.a1_item_stride * 4};
    const cuuint32_t box[3] = {BK, (cuuint32_t)a.a_rows, 1};
    if (make_tensor_map(&maps[1], p.a1, 3, dims, strides, box)) return 1;
  } else {
    maps[1] = maps[0];
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)p.N};
    const cuuint64_t strides[1] = {(cuuint64_t)ktot * 4};
    const cuuint32_t box[2] = {BK, (cuuint32_t)bn};
    if (p.split == 2) {   // fp16 pair [N][Ktot] halves, rows of the box = 64 bytes: SWIZZLE_64B
      const cuuint64_t hstrides[1] = {(cuuint64_t)ktot * 2};
      if (make_tensor_map_f16(&maps[2], p.w_hi, 2, dims, hstrides, box, 64)) return 1;
      if (make_tensor_map_f16(&maps[3], p.w_lo, 2, dims, hstrides, box, 64)) return 1;
    } else {
      if (make_tensor_map(&maps[2], p.w_hi, 2, dims, strides, box)) return 1;
      if (make_tensor_map(&maps[3], p.split == 3 ? p.w_lo : p.w_hi, 2, dims, strides, box)) return 1;
    }
  }
  a.bias = p.bias;
  a.out_raw = p.out_raw;
  a.out_elu = p.out_elu;
  a.out_item_stride = p.out_item_stride;
  a.N = p.N;
  a.M = (int)p.M;
  a.n_items = p.n_items;
  a.tiles_m = (int)cdiv(p.M, BM);
  a.tiles_n = p.N / bn;
  const long long total = (long long)a.tiles_m * a.tiles_n * p.n_items;
  ECB_REQUIRE(total < (1LL << 31), "tc_conv: too many tiles");
  a.total_tiles = (int)total;
  a.row_base1 = 0;
  a.round_out = p.round_out;
  a.halo = p.halo;
  a.stats = p.stats;
  a.cell_on = p.cell ? 1 : 0;
  if (p.cell) a.cell = *p.cell;
  {
    const int nch = a.n_cb0 * a.shifts + a.n_cb1;
    // <= 24 truncating accumulation steps per group (a chunk is 4 TF32 or 2 fp16 K steps)
    a.group = (p.split == 3 && nch > 6) ? 4 : ((p.split == 2 && nch > 12) ? 8 : nch);
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  const double rows = (double)p.M * p.n_items;
  // narrow layers (<= 64 channels on every side) are HBM-bound, the others tensor-bound
  const bool narrow = p.C0 <= 64 && p.N <= 64 && (!p.a1 || p.C1 <= 64);
  ProfScope prof(narrow ? PROF_TC_CONV_NARROW : PROF_TC_CONV_WIDE, stream, 2.0 * rows * p.N * ktot,
                 4.0 * (rows * s * p.C0 + (p.a1 ? rows * p.C1 : 0.0) + (double)ktot * p.N +
                        rows * p.N * ((p.out_raw ? 1 : 0) + (p.out_elu ? 1 : 0))));
#define ECB_TC_CASE(BN_, SP_) \
  if (bn == BN_ && p.split == SP_) return launch_one<BN_, SP_>(maps, a, grid, stream);
  {
    // four transform warps for the narrow split-operand layers (ECB_TC_NT=2 keeps two: diagnostic)
    static int nt4 = -1;
    if (nt4 < 0) {
      const char* e = getenv("ECB_TC_NT");
      nt4 = (e && e[0] == '2') ? 0 : ((e && e[0] == '3') ? 3 : 1);   // 3: four warps only for BN <= 64
    }
    if (nt4 && p.split == 3 && !p.a0_lo && !p.cell) {
      if (bn == 32) return launch_one<32, 3, 4>(maps, a, grid, stream);
      if (bn == 64) return launch_one<64, 3, 4>(maps, a, grid, stream);
    }
    if (nt4 && p.split == 2) {
      if (bn == 32) return launch_one<32, 2, 4>(maps, a, grid, stream);
      if (bn == 64) return launch_one<64, 2, 4>(maps, a, grid, stream);
      // BN = 128 with four transform warps: 448 threads leave 128 registers (44 bytes of spills), still 20-30 % faster on the
      // 1- and 2-tap layers, which are bound by the transform (down256 0.87 -> 0.68 ms, lstm.proj 0.78 -> 0.55 ms); the TF32
      // split instance gets slower with the same change and keeps two
      // ... and six are another 3-8 % on those layers (lstm.proj 0.56 -> 0.51 ms); ECB_TC_NT=3 keeps two at BN = 128
      if (bn == 128 && nt4 == 1) return launch_one<128, 2, 6>(maps, a, grid, stream);
    }
  }
  ECB_TC_CASE(32, 2)
  ECB_TC_CASE(64, 2)
  ECB_TC_CASE(128, 2)
  ECB_TC_CASE(32, 3)
  ECB_TC_CASE(64, 3)
  ECB_TC_CASE(128, 3)
  ECB_TC_CASE(32, 1)
  ECB_TC_CASE(64, 1)
  ECB_TC_CASE(128, 1)
  ECB_TC_CASE(256, 1)
#undef ECB_TC_CASE
  set_error("tc_conv: no kernel for BN=%d split=%d", bn, p.split);
  return 1;
}

}  // namespace ecb
