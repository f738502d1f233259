// SLSTM recurrence (reference modules/lstm.py:12-28 -> nn.LSTM(H, H, 2), gates i,f,g,o, zero state; H = 512 for the EnCodec
// models, 1024 for the fork's 5-ratio 10 Hz models -- described below for H = 512, the 1024 variant keeps the per-CTA
// weight slice at 32 K floats by owning 8 units of all items: 128 unit blocks x 1 batch group, K split over 32 lanes).
//
// The input projection W_ih x_t + b_ih + b_hh of ALL time steps is one tensor-core GEMM (tc_conv.cu, 1 tap);
// this file holds the part that is sequential in time: gates_t = pre_t + W_hh h_{t-1}, the cell update
// and the skip connection y = h + x (lstm.py:25-26). It is latency-bound (T dependent steps), so the design
// goal is to keep every step short and to hide the inter-CTA exchange of h behind independent work:
//
//   * one persistent cooperative kernel per LSTM layer, 128 CTAs = 32 unit blocks x 4 batch quarters. CTA
//     (ub, q) owns hidden units 16 ub .. 16 ub + 15 (64 gate rows of W_hh) for the items of batch quarter q.
//   * its 64 x 512 slice of W_hh lives in REGISTERS for the whole sequence (128 per thread: thread (rg, ks)
//     holds 4 rows x 32 k); per step only h_{t-1} moves. fp32 FMA (packed fma.rn.f32x2), fp32 accumulate: no precision trade.
//   * a quarter is processed as independent SUB-GROUPS of 4 items, and the CTA is warp-specialised so that the
//     L2 round trips of one sub-group (poll the arrival counter, fetch h_{t-1}, publish h_t, fence) overlap the
//     FFMA work of the others:
//       warps 0-7   compute: gates = W_hh_slice . h_{t-1} for one sub-group, K split over 16 lanes, butterfly
//                   reduction, result to shared memory (setmaxnreg raises their budget for the resident weights);
//       warps 8-11  cell teams (team = sub-group index mod 4): add the pre-gates, apply the cell non-linearity,
//                   store h_t (L2) and the layer output, fence, release-increment the sub-group's counter;
//       warps 12-15 loaders (two stages each of an 8-deep shared-memory ring): poll the counter of the sub-group
//                   the stage is for and cp.async its h_{t-1} in.
//     mbarriers connect the three roles; nobody ever waits on a whole-grid barrier.
#include "common.cuh"

namespace ecb {
namespace {

constexpr int L_CTAS = 128;
constexpr int L_SB = 4;               // items per sub-group
constexpr int L_THREADS = 512;        // 4 warpgroups: 2 x compute, cell teams, loader
constexpr int L_STAGES = 8;           // h ring depth
constexpr int L_LOADERS = 4;          // loader warps; loader k fills the stages of running indices n = k (mod 4)
constexpr int L_TEAMS = 4;

// Geometry as a function of the hidden size: every compute thread holds 4 gate rows x 32 k of W_hh (128 registers).
template <int LH>
struct Geo {
  static constexpr int KS = LH / 32;              // lanes a dot product is split over (16 | 32)
  static constexpr int RGS = 8 * (32 / KS);       // row groups of 4 gate rows per CTA (16 | 8)
  static constexpr int UNITS = RGS;               // hidden units per CTA: 4 gates x UNITS rows = 4 x RGS
  static constexpr int UQ = UNITS / 4;            // unit quads per gate
  static constexpr int UB = LH / UNITS;           // unit blocks (32 | 128)
  static constexpr int NQ = L_CTAS / UB;          // batch groups (4 | 1)
  static constexpr int HLD = LH;                  // h row in smem (floats): unpadded, so that lanes reading different items
                                                  // at the same k offset still hit distinct banks (see the slot permutation)
  static constexpr int GROW = 4 * UNITS;          // gate values per item in the gs exchange
  static constexpr int PAIRS = L_SB * UNITS / 32; // (item, unit) pairs per cell-team lane (2 | 1)
};

struct LstmParams {
  const float* pre;    // [B][T][4H]
  const float* w_hh;   // [4H][H] (reference layout, rows i,f,g,o)
  const float* skip;   // item b at skip + b*skip_stride, [T][H], or nullptr
  float* out;          // item b at out + b*out_stride, [T][H]
  float* hbuf;         // [2][B_pad][H] recurrent state exchange (L2)
  unsigned int* cnt;   // one arrival counter per sub-group, zeroed by the host
  long long skip_stride, out_stride;
  int B, T, out_elu;
  int q_items;         // items per batch quarter (multiple of L_SB)
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

template <int LH>
__global__ void __launch_bounds__(L_THREADS, 1)
lstm_recurrent_kernel(const LstmParams p) {
  using G = Geo<LH>;
  constexpr int L_HLD = G::HLD, L_UNITS = G::UNITS, L_UB = G::UB, L_NQ = G::NQ, GROW = G::GROW;
  extern __shared__ __align__(16) float smem[];
  float* hs = smem;                                   // [L_STAGES][L_SB][L_HLD]  h_{t-1} ring
  float* gs = hs + L_STAGES * L_SB * L_HLD;           // [L_TEAMS][L_SB][GROW]    reduced gate pre-activations
  float* cs = gs + L_TEAMS * L_SB * GROW;             // [q_items][UNITS]         cell state of this CTA's (items, units)
  uint64_t* bars = reinterpret_cast<uint64_t*>(cs + p.q_items * L_UNITS);   // 8-byte aligned: all counts above are even
  const uint32_t bar0 = smem_addr(bars);
  auto h_full = [&](int s) { return bar0 + 8u * s; };
  auto h_empty = [&](int s) { return bar0 + 8u * (L_STAGES + s); };
  auto g_full = [&](int k) { return bar0 + 8u * (2 * L_STAGES + k); };
  auto g_empty = [&](int k) { return bar0 + 8u * (2 * L_STAGES + L_TEAMS + k); };

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int ub = blockIdx.x % L_UB;
  const int q = blockIdx.x / L_UB;
  const int b_first = q * p.q_items;
  const int n_items = max(0, min(p.q_items, p.B - b_first));
  const int n_sub = (n_items + L_SB - 1) / L_SB;
  if (n_sub == 0) return;                    // empty quarter: nobody waits for this CTA

  if (tid == 0) {
    for (int s = 0; s < L_STAGES; ++s) {
      mbar_init(h_full(s), 32);              // one (async) arrive per loader lane
      mbar_init(h_empty(s), 8);              // one arrive per compute warp
    }
    for (int k = 0; k < L_TEAMS; ++k) {
      mbar_init(g_full(k), 8);
      mbar_init(g_empty(k), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < p.q_items * L_UNITS; i += L_THREADS) cs[i] = 0.f;
  __syncthreads();
  const long long B_pad = (long long)L_NQ * p.q_items;
  unsigned int* cnt_q = p.cnt + q * (p.q_items / L_SB);

  // register budget: the compute warpgroups hold the weights, the helper warpgroups give registers up
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;" ::: "memory");
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
  }

  if (warp < 8) {
    // ================================ compute warps ================================
    // row group rg (4 gate rows: gate rg >> 2, units 4 (rg & 3) .. + 3), k split ks (k = (16 j + ks) 4 + e)
    const int ks = lane & (G::KS - 1);
    const int rg = warp * (32 / G::KS) + lane / G::KS;
    const int gate = rg / G::UQ;
    const int uq = rg % G::UQ;
    // weights as packed fp32 pairs (k, k+1): the dot products run on fma.rn.f32x2 (two FMAs per issue slot), which
    // keeps even-k and odd-k partial sums in the two halves of a 64-bit accumulator
    // Slot permutation: accumulator slot (rs, is) of this lane works on row rs ^ pr and item is ^ pi with (pr, pi) taken
    // from the lane index. The transposing butterfly below then needs no selects: at exchange distance m every lane
    // keeps its lower slots and sends its upper slots, because the partner's upper slots hold exactly the logical
    // values of this lane's lower slots.
    const int pr = (lane >> 2) & 3, pi = lane & 3;
    unsigned long long w2[4][8][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* wr = p.w_hh + ((long long)gate * LH + ub * L_UNITS + uq * 4 + (r ^ pr)) * LH;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(wr + (j * G::KS + ks) * 4));
        w2[r][j][0] = v.x;
        w2[r][j][1] = v.y;
      }
    }
    uint32_t n = 0;                 // running (t, sg) index -> h ring stage and phase
    uint32_t team_use[L_TEAMS] = {0, 0, 0, 0};
    for (int t = 0; t < p.T; ++t) {
      for (int sg = 0; sg < n_sub; ++sg, ++n) {
        const int st = (int)(n % L_STAGES);
        mbar_wait(h_full(st), (n / L_STAGES) & 1u);
        const float* hsg = hs + st * (L_SB * L_HLD);
        unsigned long long acc2[4][L_SB];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int i = 0; i < L_SB; ++i) acc2[r][i] = 0ull;
#pragma unroll
        for (int i = 0; i < L_SB; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const ulonglong2 hv = *reinterpret_cast<const ulonglong2*>(hsg + (i ^ pi) * L_HLD + (j * G::KS + ks) * 4);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][i]) : "l"(w2[r][j][0]), "l"(hv.x));
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][i]) : "l"(w2[r][j][1]), "l"(hv.y));
            }
          }
        }
        float acc[4][L_SB];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int i = 0; i < L_SB; ++i)
            acc[r][i] = __uint_as_float((unsigned int)(acc2[r][i] & 0xffffffffull)) + __uint_as_float((unsigned int)(acc2[r][i] >> 32));
        __syncwarp();
        if (lane == 0) mbar_arrive(h_empty(st));   // this warp is done reading the ring slot
        // butterfly over 16 k splits: after 4 exchange steps slot 0 of lane l holds the sum over those 16 lanes of the
        // logical value index l & 15 (= row * 4 + item) -- 15 shuffles instead of 64; with 32 k splits (H = 1024) one
        // more exchange adds the two halves
        float v16[16];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int i = 0; i < L_SB; ++i) v16[r * 4 + i] = acc[r][i];
#pragma unroll
        for (int step = 0; step < 4; ++step) {
          const int m = 8 >> step;
#pragma unroll
          for (int k = 0; k < m; ++k) v16[k] += __shfl_xor_sync(0xffffffffu, v16[k + m], m);
        }
        if (G::KS == 32) v16[0] += __shfl_xor_sync(0xffffffffu, v16[0], 16);
        const int team = sg % L_TEAMS;
        mbar_wait(g_empty(team), (team_use[team] & 1u) ^ 1u);   // the team has consumed its previous gates
        // lane holds value index vi = lane & 15: row r = vi >> 2 of the row group, item i = vi & 3
        const int vi = lane & 15;
        if (G::KS == 16 || lane < 16)
          gs[team * (L_SB * GROW) + (vi & 3) * GROW + gate * L_UNITS + uq * 4 + (vi >> 2)] = v16[0];
        __syncwarp();
        if (lane == 0) mbar_arrive(g_full(team));
        ++team_use[team];
      }
    }
  } else if (warp < 12) {
    // ================================ cell teams ================================
    const int team = warp - 8;
    uint32_t use = 0;
    for (int t = 0; t < p.T; ++t) {
      float* hnext = p.hbuf + (long long)((t + 1) & 1) * B_pad * LH;
      for (int sg = team; sg < n_sub; sg += L_TEAMS, ++use) {
        const int b0 = b_first + sg * L_SB;
        // this lane's two (item, unit) pairs; their pre-gates / skip inputs do not depend on the recurrence
        float pg[G::PAIRS][4], skipv[G::PAIRS];
        bool valid[G::PAIRS];
#pragma unroll
        for (int e = 0; e < G::PAIRS; ++e) {
          const int pi = lane + 32 * e;
          const int bg = b0 + pi / L_UNITS;
          const int unit = ub * L_UNITS + pi % L_UNITS;
          valid[e] = bg < p.B;
          skipv[e] = 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) pg[e][g] = 0.f;
          if (valid[e]) {
            const float* pr = p.pre + ((long long)bg * p.T + t) * (4 * LH) + unit;
#pragma unroll
            for (int g = 0; g < 4; ++g) pg[e][g] = __ldg(pr + g * LH);
            if (p.skip) skipv[e] = __ldg(p.skip + (long long)bg * p.skip_stride + (long long)t * LH + unit);
          }
        }
        mbar_wait(g_full(team), use & 1u);
        const float* g4 = gs + team * (L_SB * GROW);
#pragma unroll
        for (int e = 0; e < G::PAIRS; ++e) {
          const int pi = lane + 32 * e;
          const int ci = pi / L_UNITS, cu = pi % L_UNITS;
          const int bg = b0 + ci;
          const int unit = ub * L_UNITS + cu;
          const float gi = sigmoidf_acc(pg[e][0] + g4[ci * GROW + 0 * L_UNITS + cu]);
          const float gf = sigmoidf_acc(pg[e][1] + g4[ci * GROW + 1 * L_UNITS + cu]);
          const float gg = tanhf(pg[e][2] + g4[ci * GROW + 2 * L_UNITS + cu]);
          const float go = sigmoidf_acc(pg[e][3] + g4[ci * GROW + 3 * L_UNITS + cu]);
          if (valid[e]) {
            float* cptr = cs + (sg * L_SB + ci) * L_UNITS + cu;
            const float c_new = gf * (*cptr) + gi * gg;
            *cptr = c_new;
            const float h_new = go * tanhf(c_new);
            asm volatile("st.global.cg.f32 [%0], %1;" ::"l"(hnext + (long long)bg * LH + unit), "f"(h_new) : "memory");
            float y = h_new;
            if (p.skip) y += skipv[e];
            if (p.out_elu) y = elu1(y);
            p.out[(long long)bg * p.out_stride + (long long)t * LH + unit] = y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(g_empty(team));   // gates consumed: the compute warps may refill the slot
        __threadfence();                              // this lane's h_t stores are visible device-wide ...
        __syncwarp();
        if (lane == 0 && t + 1 < p.T)                 // ... before the sub-group's arrival counter moves
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cnt_q + sg) : "memory");   // ordered by the fences above
      }
    }
  } else {
    // ================================ loaders ================================
    // Four loader warps, loader k owns ring stage k, i.e. the running (t, sg) indices n = k (mod 4): each poll of an
    // arrival counter is a blocking L2 round trip, so four of them have to be in flight to keep the compute warps fed.
    const uint32_t k = (uint32_t)(warp - 12);
    const uint32_t total = (uint32_t)p.T * (uint32_t)n_sub;
    for (uint32_t n = k; n < total; n += L_LOADERS) {
      const int t = (int)(n / (uint32_t)n_sub);
      const int sg = (int)(n - (uint32_t)t * (uint32_t)n_sub);
      const float* hprev = p.hbuf + (long long)(t & 1) * B_pad * LH;
      const int st = (int)(n % L_STAGES);
      mbar_wait(h_empty(st), ((n / L_STAGES) & 1u) ^ 1u);
      float* dst = hs + st * (L_SB * L_HLD);
      const int b0 = b_first + sg * L_SB;
      if (t == 0) {
        for (int f = lane; f < L_SB * (LH / 4); f += 32)
          *reinterpret_cast<float4*>(dst + (f / (LH / 4)) * L_HLD + (f % (LH / 4)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_arrive(h_full(st));
      } else {
        if (lane == 0) {
          // all unit blocks have published h_{t-1} of this sub-group once the counter reaches UB * t
          const unsigned int target = (unsigned int)L_UB * (unsigned int)t;
          unsigned int v, spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt_q + sg) : "memory");
            if (++spins > (1u << 28)) __trap();   // a lost arrival must not hang the device
          } while (v < target);
        }
        __syncwarp();
        for (int f = lane; f < L_SB * (LH / 4); f += 32) {
          const float* src = hprev + (long long)(b0 + f / (LH / 4)) * LH + (f % (LH / 4)) * 4;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst + (f / (LH / 4)) * L_HLD + (f % (LH / 4)) * 4)),
                       "l"(src)
                       : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(h_full(st)) : "memory");
      }
    }
  }
}

template <int LH>
size_t lstm_smem_bytes(int q_items) {
  using G = Geo<LH>;
  return sizeof(float) * (size_t)(L_STAGES * L_SB * G::HLD + L_TEAMS * L_SB * G::GROW + q_items * G::UNITS) +
         8 * (2 * L_STAGES + 2 * L_TEAMS);
}

int group_items(int batch, int nq) {
  const int per = (batch + nq - 1) / nq;
  return (per + L_SB - 1) / L_SB * L_SB;
}

template <int LH>
int launch_impl(const float* pre, const float* w_hh, const float* skip, long long skip_item_stride, float* out,
                long long out_item_stride, int batch, int T, int out_elu, float* workspace, cudaStream_t s) {
  using G = Geo<LH>;
  const int qi = group_items(batch, G::NQ);
  const size_t smem = lstm_smem_bytes<LH>(qi);
  ECB_REQUIRE(smem <= 200 * 1024, "lstm: batch %d needs %zu bytes of shared memory; split the batch", batch, smem);
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    ECB_CUDA(cudaFuncSetAttribute(lstm_recurrent_kernel<LH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.mark();
  }
  LstmParams p;
  p.pre = pre;
  p.w_hh = w_hh;
  p.skip = skip;
  p.out = out;
  p.skip_stride = skip_item_stride ? skip_item_stride : (long long)T * LH;
  p.out_stride = out_item_stride ? out_item_stride : (long long)T * LH;
  p.hbuf = workspace;
  p.cnt = reinterpret_cast<unsigned int*>(workspace + 2LL * G::NQ * qi * LH);
  p.B = batch;
  p.T = T;
  p.out_elu = out_elu;
  p.q_items = qi;
  ECB_CUDA(cudaMemsetAsync(p.cnt, 0, sizeof(unsigned int) * (size_t)(G::NQ * qi / L_SB + 64), s));
  const double bt = (double)batch * T;
  ProfScope prof(PROF_LSTM_REC, s, 2.0 * bt * 4 * LH * LH, 4.0 * (bt * 4 * LH + bt * LH * (skip ? 2 : 1) + 4.0 * LH * LH));
  void* args[] = {(void*)&p};
  ECB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_recurrent_kernel<LH>, dim3(L_CTAS), dim3(L_THREADS), args, smem, s));
  ECB_LAUNCHED();
  return 0;
}

// ---- step-wise recurrence for large batches ------------------------------------------------------------------------
// With hundreds of items per launch the recurrence is FMA-throughput bound on the CUDA cores (67 ns per item step); the
// recurrent product h_{t-1} W_hh^T of ALL items is then a well-shaped GEMM for the tensor-core conv kernel (tc_conv.cu,
// split-operand TF32, fp32-accurate), launched once per time step (from a CUDA graph, codec.cu) and followed by this
// element-wise cell update: gates = pre_t + rec (i, f, g, o), c = f c + i g, h = o tanh(c), y = h (+ skip) (ELU).
__global__ void __launch_bounds__(256)
lstm_cell_kernel(const float* __restrict__ pre, long long pre_item_stride, const float* __restrict__ rec,
                 float* __restrict__ c, float* __restrict__ h_out, float* __restrict__ h_lo_out, const float* __restrict__ skip,
                 long long skip_stride,
                 float* __restrict__ out, long long out_stride, int B, int H, int first, int out_elu) {
  const long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (item, unit quad)
  const int q4 = H / 4;
  if (i4 >= (long long)B * q4) return;
  const int b = (int)(i4 / q4), u = (int)(i4 % q4) * 4;
  float4 g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g[k] = __ldg(reinterpret_cast<const float4*>(pre + (long long)b * pre_item_stride + k * H + u));
    if (!first) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(rec + (long long)b * 4 * H + k * H + u));
      g[k].x += r.x; g[k].y += r.y; g[k].z += r.z; g[k].w += r.w;
    }
  }
  float4 cv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(c + (long long)b * H + u);
  float4 hv;
#define ECB_CELL(f)                                                                 \
  {                                                                                 \
    const float gi = sigmoidf_acc(g[0].f), gf = sigmoidf_acc(g[1].f), gg = tanhf(g[2].f), go = sigmoidf_acc(g[3].f); \
    cv.f = gf * cv.f + gi * gg;                                                     \
    hv.f = go * tanhf(cv.f);                                                        \
  }
  ECB_CELL(x) ECB_CELL(y) ECB_CELL(z) ECB_CELL(w)
#undef ECB_CELL
  *reinterpret_cast<float4*>(c + (long long)b * H + u) = cv;
  *reinterpret_cast<float4*>(h_out + (long long)b * H + u) = hv;
  if (h_lo_out) {   // TF32 remainder of h: the next step's GEMM takes both operand halves by TMA (tc_conv a0_lo)
    auto lo = [](float v) {
      return __uint_as_float((__float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u)) + 0x1000u) & 0xffffe000u);
    };
    *reinterpret_cast<float4*>(h_lo_out + (long long)b * H + u) = make_float4(lo(hv.x), lo(hv.y), lo(hv.z), lo(hv.w));
  }
  float4 y = hv;
  if (skip) {
    const float4 sv = __ldg(reinterpret_cast<const float4*>(skip + (long long)b * skip_stride + u));
    y.x += sv.x; y.y += sv.y; y.z += sv.z; y.w += sv.w;
  }
  if (out_elu) y = make_float4(elu1(y.x), elu1(y.y), elu1(y.z), elu1(y.w));
  *reinterpret_cast<float4*>(out + (long long)b * out_stride + u) = y;
}

}  // namespace

int launch_lstm_cell(const float* pre_t, long long pre_item_stride, const float* rec, float* c, float* h_out, float* h_lo_out,
                     const float* skip_t,
                     long long skip_item_stride, float* out_t, long long out_item_stride, int batch, int H, int first, int out_elu,
                     cudaStream_t s) {
  const long long n4 = (long long)batch * (H / 4);
  ProfScope prof(PROF_LSTM_REC, s, 0.0, 4.0 * (double)batch * H * (first ? 6 : 12));
  lstm_cell_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, s>>>(pre_t, pre_item_stride, rec, c, h_out, h_lo_out, skip_t, skip_item_stride, out_t,
                                                          out_item_stride, batch, H, first, out_elu);
  ECB_LAUNCHED();
  return 0;
}

// 2 x B_pad x H state + one counter per sub-group
int lstm_recurrent_workspace_floats(int batch, int H) {
  const int nq = H == 1024 ? Geo<1024>::NQ : Geo<512>::NQ;
  const int qi = group_items(batch, nq);
  const int persistent = 2 * nq * qi * H + nq * qi / L_SB + 64;
  const int stepwise = 9 * H * batch + 64;   // rec [B][4H], h [2][B][H], c [B][H], h_lo [2][B][H]
  int need = persistent > stepwise ? persistent : stepwise;
  if (lstm_tc_supported(batch, H) && lstm_tc_workspace_floats(batch) > need) need = lstm_tc_workspace_floats(batch);
  return need;
}

int launch_lstm_recurrent(const float* pre, const float* w_hh, const float* skip, long long skip_item_stride, float* out,
                          long long out_item_stride, int batch, int T, int H, int out_elu, float* workspace,
                          cudaStream_t s) {
  ECB_REQUIRE(H == 512 || H == 1024, "lstm: hidden size %d unsupported (512 or 1024)", H);
  ECB_REQUIRE(batch > 0 && T > 0, "lstm: bad batch %d / T %d", batch, T);
  if (H == 512) return launch_impl<512>(pre, w_hh, skip, skip_item_stride, out, out_item_stride, batch, T, out_elu, workspace, s);
  return launch_impl<1024>(pre, w_hh, skip, skip_item_stride, out, out_item_stride, batch, T, out_elu, workspace, s);
}

}  // namespace ecb
