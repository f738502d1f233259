// Fused residual vector quantiser on the 5th-generation tensor cores (sm_100a).
//
// Replaces the Python loop of ResidualVectorQuantization.forward/encode (reference
// quantization/core_vq.py:385-432) and, per layer, EuclideanCodebook.quantize + dequantize
// (core_vq.py:178-202): dist = (|x|^2 - 2 x.E^T) + |E|^2 in that association order, arg-min with the
// lowest index on ties (== the reference's first arg-max of -dist), gather, residual -= q, out += q.
//
// One CTA carries a tile of 128 frames through ALL n_q layers:
//   * the residual tile R [128 x 128] lives in shared memory as the K-major SWIZZLE_128B A operand of
//     tcgen05.mma.kind::tf32 (4 chunk tiles of [128 rows x 32 dims]), together with its TF32 remainder R_lo;
//   * each codebook streams from L2 through a TMA ring of [128 entries x 32 dims] (hi, lo) tiles -- the split
//     E = E_hi + E_lo is made once at load time -- and x.E^T is computed to fp32 accuracy as
//     R*[E_hi | E_lo] (one MMA, N = 256: main | correction accumulator) + R_lo*E_hi (correction);
//   * 8 epilogue warps read the [128 x 128-entry] distance block from TMEM (double-buffered, so the scan of
//     block b overlaps the MMAs of block b+1), form (|x|^2 - 2 dot) + |e|^2 in fp32 and keep the running
//     arg-min (strict <, entries ascending => lowest index on ties);
//   * at the end of a layer the same warps gather the chosen entries (exact fp32 rows of the codebook), update
//     R / R_lo / |x|^2 in place and accumulate the quantised output in registers, in layer order, so that
//     `quantized` is bit-identical to decode(codes).
// Nothing but the codes (and the final quantised frames) goes back to HBM.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ecb {
namespace {
using namespace tc;

constexpr int RD = 128;          // frame dimension (= K)
constexpr int FT = 128;          // frames per CTA tile (= UMMA M)
constexpr int EB = 128;          // codebook entries per block (= UMMA N of the correction MMA)
constexpr int KC = 32;           // dims per K chunk (one 128-byte swizzle row)
constexpr int NKC = RD / KC;     // 4
constexpr int A_TILE = FT * KC * 4;          // 16 KB
constexpr int B_TILE = EB * KC * 4;          // 16 KB
constexpr int STAGE_BYTES = 2 * B_TILE;      // E_hi tile followed by E_lo tile
constexpr int STAGES = 2;
constexpr int THREADS = 320;                 // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue / update
constexpr int ACC_COLS = 2 * EB;             // [main | correction]
constexpr int TMEM_COLS = 2 * ACC_COLS;      // double-buffered: 512
constexpr int SMEM_BYTES = 2 * NKC * A_TILE + STAGES * STAGE_BYTES + 8192 + 1024;
// SPLIT = 2 (fp16 pair operands, the scheme of tc_conv.cu: R = R1 + 2^-11 R2, E = E1 + 2^-11 E2 in fp16, kind::f16 MMAs with
// K = 16): the fp32 residual stays where it is (exact update arithmetic, re-rank operand) but is no MMA operand any more; R1 / R2
// are two [128 rows x 64 dims] half tiles each (SWIZZLE_128B rows of 128 bytes) in the place of R_lo, a codebook stage is
// [E1 | E2] x 64 dims (the same 32 KB), so a 128-entry block is 2 stages and 8 K steps instead of 4 and 16.
constexpr int H_TILE = FT * 64 * 2;          // 16 KB: [128 rows x 64 halves]
__device__ unsigned int g_rvq_f16_sat = 0;   // residual tiles in which the fp16 conversion saturated (see ecb_f16_saturation_count)

struct RvqTcArgs {
  const float* frames;     // [n][128]
  const float* codebooks;  // [n_q][bins][128] exact fp32 (gather source)
  const float* e2;         // [n_q][bins]
  long long* codes;        // [n_q][n]
  float* quantized;        // [n][128] or nullptr
  float* stack;            // [n_q][n][128] or nullptr
  long long n;
  int n_q, bins, n_tiles;
};

template <int SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
rvq_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const RvqTcArgs p) {
  constexpr int KCH = SPLIT == 2 ? 2 : NKC;   // codebook stages (K chunks) per entry block
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: R_hi [4][16 KB] | R_lo [4][16 KB] | B ring | misc
  const uint32_t r_hi = smem_base;
  const uint32_t r_lo = smem_base + NKC * A_TILE;
  const uint32_t b_ring = smem_base + 2 * NKC * A_TILE;
  uint8_t* misc = smem_gen + 2 * NKC * A_TILE + STAGES * STAGE_BYTES;
  float* x2s = reinterpret_cast<float*>(misc);                 // [2][128] partial |x|^2 of the two dim halves
  float* cand_d = x2s + 2 * FT;                                // [4][128] best / runner-up of the two entry halves
  int* cand_i = reinterpret_cast<int*>(cand_d + 4 * FT);       // [4][128]
  float* ex_d = reinterpret_cast<float*>(cand_i + 4 * FT);     // [2][128] exact distances of re-ranked candidates
  const uint32_t bar_base = smem_base + 2 * NKC * A_TILE + STAGES * STAGE_BYTES + 7168;   // after the 6 KB of arrays above
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t rready_bar = bar_base + 8u * (2 * STAGES + 4);   // residual tile (re)written for the next layer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 7168 + 8 * (2 * STAGES + 5));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blocks = p.bins / EB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 8);     // one arrive per epilogue warp
    }
    mbar_init(rready_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer: codebook tiles, independent of the residual ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int layer = 0; layer < p.n_q; ++layer) {
          for (int blk = 0; blk < n_blocks; ++blk) {
            for (int kc = 0; kc < KCH; ++kc, ++it) {
              const int s = (int)(it % STAGES);
              mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
              const uint32_t dst = b_ring + s * STAGE_BYTES;
              mbar_expect_tx(full_bar(s), STAGE_BYTES);
              const int row = layer * p.bins + blk * EB;
              tma_load_2d(dst, &map_hi, full_bar(s), kc * (SPLIT == 2 ? 64 : KC), row);
              tma_load_2d(dst + B_TILE, &map_lo, full_bar(s), kc * (SPLIT == 2 ? 64 : KC), row);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc2 = SPLIT == 2 ? ((1u << 4) | ((uint32_t)((2 * EB) >> 3) << 17) | ((uint32_t)(FT >> 4) << 24)) : umma_idesc_tf32(FT, 2 * EB);
      constexpr uint32_t idesc1 = SPLIT == 2 ? ((1u << 4) | ((uint32_t)(EB >> 3) << 17) | ((uint32_t)(FT >> 4) << 24)) : umma_idesc_tf32(FT, EB);
      uint32_t it = 0, bc = 0, lc = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int layer = 0; layer < p.n_q; ++layer, ++lc) {
          mbar_wait(rready_bar, lc & 1u);              // R / R_lo of this layer are in place
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int blk = 0; blk < n_blocks; ++blk, ++bc) {
            const int ab = (int)(bc & 1u);
            mbar_wait(acce_bar(ab), ((bc >> 1) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_main = tmem_base + (uint32_t)(ab * ACC_COLS);
            for (int kc = 0; kc < KCH; ++kc, ++it) {
              const int s = (int)(it % STAGES);
              mbar_wait(full_bar(s), (it / STAGES) & 1u);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t b_addr = b_ring + s * STAGE_BYTES;
              if (SPLIT == 2) {
                // [main | corr] (+)= R1 * [E1 | E2], corr += R2 * E1 (the epilogue adds corr 2^-11): four K steps of 16 per 64 dims
                const uint32_t a1_addr = r_lo + kc * H_TILE;
                const uint32_t a2_addr = r_lo + 2 * H_TILE + kc * H_TILE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = umma_desc_sw128(a1_addr + k * 32);
                  const uint64_t dal = umma_desc_sw128(a2_addr + k * 32);
                  const uint64_t db = umma_desc_sw128(b_addr + k * 32);
                  tcgen05_mma_f16(d_main, da, db, idesc2, (kc > 0 || k > 0) ? 1u : 0u);
                  tcgen05_mma_f16(d_main + EB, dal, db, idesc1, 1u);
                }
              } else {
              const uint32_t a_addr = r_hi + kc * A_TILE;
              const uint32_t alo_addr = r_lo + kc * A_TILE;
#pragma unroll
              for (int k = 0; k < KC / 8; ++k) {
                const uint64_t da = umma_desc_sw128(a_addr + k * 32);
                const uint64_t dal = umma_desc_sw128(alo_addr + k * 32);
                const uint64_t db = umma_desc_sw128(b_addr + k * 32);
                tcgen05_mma_tf32(d_main, da, db, idesc2, (kc > 0 || k > 0) ? 1u : 0u);   // [main | corr] (+)= R * [E_hi | E_lo]
                tcgen05_mma_tf32(d_main + EB, dal, db, idesc1, 1u);                       // corr += R_lo * E_hi
              }
              }
              tcgen05_commit(empty_bar(s));
            }
            tcgen05_commit(accf_bar(ab));
          }
        }
      }
    }
  } else {
    // ================================ epilogue / update warps ================================
    const int et = threadIdx.x - 64;            // 0..255
    const int quad = warp & 3;                  // TMEM lane quadrant (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;           // which 64 of a block's 128 entries (scan) / which 64 dims (update)
    const int f = quad * 32 + lane;             // frame row inside the tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t bc = 0;
    // element (row f, dim d) of the residual: chunk tile d / 32, row f, 16-byte chunk ((d % 32) / 4) ^ (f & 7)
    auto r_off = [&](int d4) { return (d4 >> 3) * A_TILE + f * 128 + (((d4 & 7) ^ (f & 7)) << 4); };
    // the second operand image of float4 d4 of this row: SPLIT == 3 the TF32 remainder at the same offset in R_lo; SPLIT == 2 the
    // fp16 pair: tile d4 / 16 of R1 (and of R2, two tiles further), row f, 16-byte chunk ((d4 % 16) / 2) ^ (f & 7), 8-byte half d4 & 1
    __half2 hmax = __float2half2_rn(0.f);
    auto store_operand = [&](int d4, const float4& v) {
      if (SPLIT == 2) {
        uint32_t h01, h23, l01, l23;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(v.w), "f"(v.z));
        const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
        const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l01) : "f"((v.y - f01.y) * 2048.f), "f"((v.x - f01.x) * 2048.f));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l23) : "f"((v.w - f23.y) * 2048.f), "f"((v.z - f23.x) * 2048.f));
        hmax = __hmax2(hmax, __hmax2(__habs2(*reinterpret_cast<const __half2*>(&h01)), __habs2(*reinterpret_cast<const __half2*>(&h23))));
        const int off = NKC * A_TILE + (d4 >> 4) * H_TILE + f * 128 + ((((d4 & 15) >> 1) ^ (f & 7)) << 4) + ((d4 & 1) << 3);
        *reinterpret_cast<uint2*>(smem_gen + off) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(smem_gen + off + 2 * H_TILE) = make_uint2(l01, l23);
      } else {
        *reinterpret_cast<float4*>(smem_gen + NKC * A_TILE + r_off(d4)) =
            make_float4(rn_tf32(v.x - trunc_tf32(v.x)), rn_tf32(v.y - trunc_tf32(v.y)), rn_tf32(v.z - trunc_tf32(v.z)),
                        rn_tf32(v.w - trunc_tf32(v.w)));
      }
    };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long fg = (long long)tile * FT + f;
      const bool live = fg < p.n;
      // ---- load this thread's half row of the frame tile: R, R_lo, partial |x|^2
      float qacc[64];
      float x2p = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int d4 = half * 16 + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) v = __ldg(reinterpret_cast<const float4*>(p.frames + fg * RD) + d4);
        *reinterpret_cast<float4*>(smem_gen + r_off(d4)) = v;
        store_operand(d4, v);
        x2p += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        qacc[j * 4 + 0] = 0.f; qacc[j * 4 + 1] = 0.f; qacc[j * 4 + 2] = 0.f; qacc[j * 4 + 3] = 0.f;
      }
      x2s[half * FT + f] = x2p;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");            // all 8 warps: tile written, |x|^2 halves visible
      if (lane == 0) mbar_arrive(rready_bar);

      for (int layer = 0; layer < p.n_q; ++layer) {
        const float x2 = x2s[f] + x2s[FT + f];
        const float* e2 = p.e2 + (long long)layer * p.bins;
        float best_d = INFINITY, sec_d = INFINITY;   // best and runner-up of this thread's entries
        int best_i = 0, sec_i = 0;
        for (int blk = 0; blk < n_blocks; ++blk, ++bc) {
          const int ab = (int)(bc & 1u);
          mbar_wait(accf_bar(ab), (bc >> 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t col0 = (uint32_t)(ab * ACC_COLS + half * 64);
#pragma unroll 1
          for (int c = 0; c < 64; c += 16) {
            uint32_t vm[16], vc[16];
            tcgen05_ld16(lane_base + col0 + (uint32_t)c, vm);
            tcgen05_ld16(lane_base + col0 + (uint32_t)(EB + c), vc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int j0 = blk * EB + half * 64 + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float dot = SPLIT == 2 ? fmaf(__uint_as_float(vc[i]), 1.f / 2048.f, __uint_as_float(vm[i]))
                                           : __uint_as_float(vm[i]) + __uint_as_float(vc[i]);
              const float t = __fsub_rn(x2, 2.f * dot);                 // core_vq.py:183-187 association order
              const float d = __fadd_rn(t, __ldg(e2 + j0 + i));
              if (d < sec_d) {
                const bool nb = d < best_d;        // strict: entries ascend, so ties keep the lowest index in front
                sec_d = nb ? best_d : d;
                sec_i = nb ? best_i : j0 + i;
                best_d = nb ? d : best_d;
                best_i = nb ? j0 + i : best_i;
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acce_bar(ab));
        }
        // ---- arg-min across the two entry halves (lowest index on ties). The tensor-core distances carry ~1e-6
        // relative error after the |x|^2 - 2 x.e + |e|^2 cancellation; when the two leading candidates are closer than
        // 1e-4 relative they are re-ranked with an exact fp32 evaluation (sequential fmaf over the 128 dims, the
        // arithmetic of the CUDA-core kernel), so near-ties resolve the way an fp32 implementation resolves them.
        cand_d[(half * 2 + 0) * FT + f] = best_d;
        cand_i[(half * 2 + 0) * FT + f] = best_i;
        cand_d[(half * 2 + 1) * FT + f] = sec_d;
        cand_i[(half * 2 + 1) * FT + f] = sec_i;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float td[4];
        int ti[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          td[c] = cand_d[c * FT + f];
          ti[c] = cand_i[c * FT + f];
        }
        // top two of the four candidates, ordered by (distance, index)
        float d0 = INFINITY, d1 = INFINITY;
        int i0 = 0x7fffffff, i1 = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool lt1 = td[c] < d1 || (td[c] == d1 && ti[c] < i1);
          if (lt1) {
            const bool lt0 = td[c] < d0 || (td[c] == d0 && ti[c] < i0);
            d1 = lt0 ? d0 : td[c];
            i1 = lt0 ? i0 : ti[c];
            d0 = lt0 ? td[c] : d0;
            i0 = lt0 ? ti[c] : i0;
          }
        }
        const bool rerank = (d1 - d0) <= 1e-4f * fabsf(d0) && i1 != 0x7fffffff && i1 < p.bins;
        if (rerank) {
          // thread `half` evaluates candidate `half` exactly
          const int ci = half == 0 ? i0 : i1;
          const float4* e4 = reinterpret_cast<const float4*>(p.codebooks + ((long long)layer * p.bins + ci) * RD);
          float dot = 0.f;
#pragma unroll 8
          for (int d4 = 0; d4 < 32; ++d4) {
            const float4 rv = *reinterpret_cast<const float4*>(smem_gen + r_off(d4));
            const float4 ev = __ldg(e4 + d4);
            dot = fmaf(rv.x, ev.x, dot);
            dot = fmaf(rv.y, ev.y, dot);
            dot = fmaf(rv.z, ev.z, dot);
            dot = fmaf(rv.w, ev.w, dot);
          }
          ex_d[half * FT + f] = __fadd_rn(__fsub_rn(x2, 2.f * dot), __ldg(e2 + ci));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        int code = i0;
        if (rerank) {
          const float e0 = ex_d[f], e1 = ex_d[FT + f];
          code = (e1 < e0 || (e1 == e0 && i1 < i0)) ? i1 : i0;
        }
        if (half == 0 && live) p.codes[(long long)layer * p.n + fg] = (long long)code;
        const float4* q4 = reinterpret_cast<const float4*>(p.codebooks + ((long long)layer * p.bins + code) * RD) + half * 16;
        float x2n = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int d4 = half * 16 + j;
          const float4 q = __ldg(q4 + j);
          float4 v = *reinterpret_cast<float4*>(smem_gen + r_off(d4));
          v.x -= q.x; v.y -= q.y; v.z -= q.z; v.w -= q.w;                                   // core_vq.py:402
          *reinterpret_cast<float4*>(smem_gen + r_off(d4)) = v;
          store_operand(d4, v);
          qacc[j * 4 + 0] += q.x; qacc[j * 4 + 1] += q.y; qacc[j * 4 + 2] += q.z; qacc[j * 4 + 3] += q.w;   // core_vq.py:404
          if (p.stack && live)
            *(reinterpret_cast<float4*>(p.stack + ((long long)layer * p.n + fg) * RD) + d4) = q;
          x2n += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone has read the old |x|^2 / candidates
        x2s[half * FT + f] = x2n;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (layer + 1 < p.n_q && lane == 0) mbar_arrive(rready_bar);
      }
      if (SPLIT == 2 && __hge(__hmax(__low2half(hmax), __high2half(hmax)), __ushort_as_half((unsigned short)0x7BFF))) {
        atomicAdd(&g_rvq_f16_sat, 1u);
        hmax = __float2half2_rn(0.f);
      }
      if (p.quantized && live) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          *(reinterpret_cast<float4*>(p.quantized + fg * RD) + half * 16 + j) =
              make_float4(qacc[j * 4 + 0], qacc[j * 4 + 1], qacc[j * 4 + 2], qacc[j * 4 + 3]);
      }
      (void)et;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// codebooks [rows][128] -> hi = rn_tf32(e), lo = rn_tf32(e - hi)
__global__ void split_codebook_kernel(const float* __restrict__ cb, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = cb[i];
    const float h = rn_tf32(v);
    hi[i] = h;
    lo[i] = rn_tf32(v - h);
  }
}

// codebooks [rows][128] -> fp16 pair: e1 = fp16(e), e2 = fp16((e - e1) 2^11), both [rows][128] halves
__global__ void split_codebook_f16_kernel(const float* __restrict__ cb, __half* __restrict__ e1, __half* __restrict__ e2, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = cb[i];
    const __half h = __float2half_rn(v);
    e1[i] = h;
    e2[i] = __float2half_rn((v - __half2float(h)) * 2048.f);
  }
}

}  // namespace

long long rvq_f16_saturation_count(int reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_rvq_f16_sat, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned int z = 0;
    cudaMemcpyToSymbol(g_rvq_f16_sat, &z, sizeof(z));
  }
  return (long long)v;
}

// the two half arrays of numel elements each fit one float array of numel elements
int launch_rvq_split_f16(const float* codebooks, void* e1, void* e2, long long numel, cudaStream_t s) {
  split_codebook_f16_kernel<<<1024, 256, 0, s>>>(codebooks, reinterpret_cast<__half*>(e1), reinterpret_cast<__half*>(e2), numel);
  ECB_LAUNCHED();
  return 0;
}

int launch_rvq_split(const float* codebooks, float* hi, float* lo, long long numel, cudaStream_t s) {
  split_codebook_kernel<<<1024, 256, 0, s>>>(codebooks, hi, lo, numel);
  ECB_LAUNCHED();
  return 0;
}

// f16_pair != 0: cb_hi / cb_lo point at the half arrays of launch_rvq_split_f16 and the distances run on fp16 pair operands
int launch_rvq_encode_tc(const float* frames, long long n, const float* codebooks, const float* cb_hi, const float* cb_lo,
                         const float* e2, int n_q_total, int n_q, int bins, long long* codes, float* quantized, float* stack,
                         cudaStream_t s, int f16_pair) {
  ECB_REQUIRE(n > 0 && n_q > 0 && n_q <= n_q_total, "rvq: empty input (n=%lld, n_q=%d)", n, n_q);
  ECB_REQUIRE(bins % EB == 0, "rvq: bins=%d must be a multiple of %d", bins, EB);
  CUtensorMap mh, ml;
  const cuuint64_t dims[2] = {(cuuint64_t)RD, (cuuint64_t)n_q_total * bins};
  if (f16_pair) {
    const cuuint64_t strides[1] = {(cuuint64_t)RD * 2};
    const cuuint32_t box[2] = {64, EB};
    if (make_tensor_map_f16(&mh, cb_hi, 2, dims, strides, box)) return 1;
    if (make_tensor_map_f16(&ml, cb_lo, 2, dims, strides, box)) return 1;
  } else {
    const cuuint64_t strides[1] = {(cuuint64_t)RD * 4};
    const cuuint32_t box[2] = {KC, EB};
    if (make_tensor_map(&mh, cb_hi, 2, dims, strides, box)) return 1;
    if (make_tensor_map(&ml, cb_lo, 2, dims, strides, box)) return 1;
  }
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    ECB_CUDA(cudaFuncSetAttribute(rvq_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ECB_CUDA(cudaFuncSetAttribute(rvq_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set.mark();
  }
  RvqTcArgs a;
  a.frames = frames;
  a.codebooks = codebooks;
  a.e2 = e2;
  a.codes = codes;
  a.quantized = quantized;
  a.stack = stack;
  a.n = n;
  a.n_q = n_q;
  a.bins = bins;
  a.n_tiles = (int)cdiv(n, FT);
  const int grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  ProfScope prof(PROF_RVQ, s, 2.0 * (double)n * n_q * bins * RD,
                 4.0 * ((double)n * RD * (quantized ? 2 : 1) + (double)n_q * bins * RD) + 8.0 * (double)n * n_q);
  if (f16_pair) rvq_tc_kernel<2><<<grid, THREADS, SMEM_BYTES, s>>>(mh, ml, a);
  else rvq_tc_kernel<3><<<grid, THREADS, SMEM_BYTES, s>>>(mh, ml, a);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
