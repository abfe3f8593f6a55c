// Fused spatial self / reference attention for sm_100a (flash-style, tcgen05 + TMEM + TMA).
//
//   O[f, h] = softmax( Q[f,h] . [K_own[f,h] ; K_bank[h]]^T * scale ) . [V_own[f,h] ; V_bank[h]]
//
// One launch covers every frame of a UNet call: frames before `first_bank_frame` (the unconditional CFG branch)
// attend to their own N tokens only, the others to 2N keys (own tokens followed by the ReferenceNet bank, whose K/V
// were projected once per video and are read in place: no `bank.repeat(F)` / `torch.cat` copy, no discarded
// reference-attention for the unconditional half, no CPU-mask scatter).
//
// Replaces diffusers Attention/AttnProcessor2_0 -> F.scaled_dot_product_attention as driven by the patched block
// forward of ReferenceAttentionControl (reference src/models/mutual_self_attention.py:147-186; plain blocks
// src/models/attention.py:323-330, write mode :137-146; PoseGuider's self-attention src/models/pose_guider.py:86-89).
//
// Three kernel generations share this file (dispatch in ap_attention_f16; measurements in profiles/):
//   attention_kernel   (v1)  d <= 192 or <= 128 tokens: the structure described next
//   attention3_kernel  (v3)  d <= 128: two query tiles per CTA, P kept in TMEM
//   attention5_kernel  (v5)  d <= 64 (the 64x64 level, 88 % of the FLOPs): one query tile per CTA, two CTAs per SM
// v1: persistent CTAs over (frame, head, 128-query tile) work units; 6 warps:
//   warp 0     TMA producer: Q tile once per unit, K/V tiles through an mbarrier ring (own rows, then bank rows)
//   warp 1     MMA issuer:   S = Q.K^T (accumulator double-buffered in TMEM), O += P.V (P from smem, V MN-major)
//   warps 2-5  softmax + epilogue: thread = query row; S row held in registers; base-2 online softmax with lazy
//              (thresholded) rescaling of the TMEM-resident O; P written to 128B-swizzled smem as the next A operand.
// Head dim d is zero-padded to DPAD in {64,128,192} by the QKV projection (padded weight rows), so every operand slab
// is a clean 64-column / 128-byte swizzle atom.
#include <stdio.h>
#include <stdlib.h>

#include "ap_host.h"
#include "ap_ptx.cuh"

namespace ap {

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  int n_frames, tokens, heads, head_dim;
  int bank_tokens;        // 0 = no bank
  int first_bank_frame;   // frames >= this attend to the bank as well
  int frames_per_bank;    // bank index = (frame - first_bank_frame) / frames_per_bank
  int m_tiles;            // ceil(tokens / 128)
  int num_units;          // n_frames * heads * m_tiles
  float scale_log2;       // softmax scale * log2(e)
  __half* out;
  long long ldo;
};

// Row maximum with the 3-input FMNMX3 of sm_100 (half the instructions of a 2-input reduction).
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
template <int N>
__device__ __forceinline__ float row_max(const uint32_t* sr) {
  static_assert(N % 32 == 0, "row length");
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    m[i] = fmax3(__uint_as_float(sr[i]), __uint_as_float(sr[8 + i]), __uint_as_float(sr[16 + i]));
#pragma unroll
  for (int i = 24; i + 16 <= N; i += 16)
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmax3(m[j], __uint_as_float(sr[i + j]), __uint_as_float(sr[i + 8 + j]));
  if ((N - 24) % 16 != 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], __uint_as_float(sr[N - 8 + j]));
  }
  return fmax3(fmax3(m[0], m[1], m[2]), fmax3(m[3], m[4], m[5]), fmaxf(m[6], m[7]));
}


template <int DPAD, int BN>
struct AttnCfg {
  static constexpr int SLABS = DPAD / 64;
  static constexpr int Q_BYTES = 128 * DPAD * 2;
  static constexpr int K_BYTES = BN * DPAD * 2;
  static constexpr int P_BYTES = 128 * BN * 2;
  static constexpr int KV_STAGE = 2 * K_BYTES;
  static constexpr int BUDGET = 227 * 1024 - 2048 - Q_BYTES - P_BYTES;
  static constexpr int STAGES_RAW = BUDGET / KV_STAGE;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int SMEM_BYTES = Q_BYTES + P_BYTES + STAGES * KV_STAGE + 1024 + 256;
  static constexpr uint32_t TMEM_S0 = 0, TMEM_S1 = 128, TMEM_O = 256;
  static_assert(STAGES >= 2, "need at least a double-buffered K/V ring");
  static_assert(DPAD % 64 == 0 && DPAD <= 256 && (BN == 64 || BN == 128), "tile config");
};

constexpr float kRescaleThreshold = 8.0f;  // log2 units: stale row maxima keep P <= 2^8 (fine for fp16 P, fp32 O)

template <int DPAD, int BN>
__global__ void __launch_bounds__(192, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmBK,
                 const __grid_constant__ CUtensorMap tmBV, const AttnParams p) {
  using Cfg = AttnCfg<DPAD, BN>;
  constexpr int ST = Cfg::STAGES;
  griddep_launch_dependents();   // PDL (ap_host.h::launch_pdl)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_p = smem_q + Cfg::Q_BYTES;
  uint8_t* smem_kv = smem_p + Cfg::P_BYTES;  // stage s: K at s*KV_STAGE, V at s*KV_STAGE + K_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + ST * Cfg::KV_STAGE);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;            // [ST]
  uint64_t* v_full = k_full + ST;         // [ST]
  uint64_t* kv_empty = v_full + ST;       // [ST]
  uint64_t* s_full = kv_empty + ST;       // [2]
  uint64_t* s_empty = s_full + 2;         // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();   // PDL: Q/K/V come from the previous kernel

  const int own_tiles = (p.tokens + BN - 1) / BN;
  const int bank_tiles = (p.bank_tokens + BN - 1) / BN;
  const int units_per_frame = p.heads * p.m_tiles;

  // unit -> (frame, head, m_tile); frames with a bank (the heavy ones) are scheduled first
  auto decode = [&](int unit, int& frame, int& head, int& m_tile, int& T) {
    const int fk = unit / units_per_frame;
    const int rem = unit % units_per_frame;
    frame = (fk + p.first_bank_frame) % p.n_frames;
    head = rem / p.m_tiles;
    m_tile = rem % p.m_tiles;
    const bool has_bank = p.bank_tokens > 0 && frame >= p.first_bank_frame;
    T = own_tiles + (has_bank ? bank_tiles : 0);
  };

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t g = 0;   // global K/V tile counter
      uint32_t uc = 0;  // unit counter
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
        int frame, head, m_tile, T;
        decode(unit, frame, head, m_tile, T);
        mbar_wait(q_empty, (uc & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SLABS; ++s)
          tma_load_2d(&tmQ, q_full, smem_q + s * (128 * 128), head * DPAD + s * 64,
                      frame * p.tokens + m_tile * 128);
        const int bank_idx = (frame - p.first_bank_frame) / p.frames_per_bank;
        for (int j = 0; j < T; ++j, ++g) {
          const int stage = g % ST;
          const uint32_t ph = (g / ST) & 1;
          mbar_wait(&kv_empty[stage], ph ^ 1);
          const bool own = j < own_tiles;
          const CUtensorMap* mk = own ? &tmK : &tmBK;
          const CUtensorMap* mv = own ? &tmV : &tmBV;
          const int row = own ? frame * p.tokens + j * BN : bank_idx * p.bank_tokens + (j - own_tiles) * BN;
          uint8_t* kd = smem_kv + stage * Cfg::KV_STAGE;
          uint8_t* vd = kd + Cfg::K_BYTES;
          mbar_arrive_expect_tx(&k_full[stage], Cfg::K_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(mk, &k_full[stage], kd + s * (BN * 128), head * DPAD + s * 64, row);
          mbar_arrive_expect_tx(&v_full[stage], Cfg::K_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(mv, &v_full[stage], vd + s * (BN * 128), head * DPAD + s * 64, row);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, DPAD, 0, 1);  // B (= V) is MN-major
      uint32_t g = 0, uc = 0;
      auto issue_qk = [&](uint32_t gi, bool last_of_unit) {
        const int sbuf = gi & 1;
        const int stage = gi % ST;
        mbar_wait(&s_empty[sbuf], ((gi >> 1) & 1) ^ 1);
        mbar_wait(&k_full[stage], (gi / ST) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (sbuf ? Cfg::TMEM_S1 : Cfg::TMEM_S0);
        const uint32_t qa = smem_u32(smem_q);
        const uint32_t ka = smem_u32(smem_kv + stage * Cfg::KV_STAGE);
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk) {
          const uint64_t da = umma_desc_k_sw128(qa + (kk / 4) * (128 * 128) + (kk % 4) * 32);
          const uint64_t db = umma_desc_k_sw128(ka + (kk / 4) * (BN * 128) + (kk % 4) * 32);
          umma_f16_ss(d_tmem, da, db, idesc_qk, kk != 0);
        }
        umma_commit(&s_full[sbuf]);
        if (last_of_unit) umma_commit(q_empty);
      };
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
        int frame, head, m_tile, T;
        decode(unit, frame, head, m_tile, T);
        mbar_wait(q_full, uc & 1);
        for (int j = 0; j < T; ++j, ++g) {
          if (j == 0) issue_qk(g, T == 1);
          if (j + 1 < T) issue_qk(g + 1, j + 2 == T);
          const int stage = g % ST;
          mbar_wait(p_full, g & 1);
          mbar_wait(&v_full[stage], (g / ST) & 1);
          tc_fence_after();
          const uint32_t pa = smem_u32(smem_p);
          const uint32_t va = smem_u32(smem_kv + stage * Cfg::KV_STAGE + Cfg::K_BYTES);
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            const uint64_t da = umma_desc_k_sw128(pa + (kk / 4) * (128 * 128) + (kk % 4) * 32);
            const uint64_t db = umma_desc_mn_sw128(va + kk * (16 * 128), BN * 128);
            umma_f16_ss(tmem_base + Cfg::TMEM_O, da, db, idesc_pv, (j | kk) != 0);
          }
          umma_commit(&kv_empty[stage]);
          umma_commit(pv_done);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int lane_group = warp & 3;
    const int row = lane_group * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(lane_group * 32) << 16;
    uint32_t g = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int frame, head, m_tile, T;
      decode(unit, frame, head, m_tile, T);
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j, ++g) {
        const int sbuf = g & 1;
        mbar_wait(&s_full[sbuf], (g >> 1) & 1);
        tc_fence_after();
        uint32_t sr[BN];
        const uint32_t t_s = tmem_base + (sbuf ? Cfg::TMEM_S1 : Cfg::TMEM_S0) + t_lane;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tmem_ld_32x32b_x32(t_s + c * 32, sr + c * 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sbuf]);

        // valid keys in this tile
        const bool own = j < own_tiles;
        const int jj = own ? j : j - own_tiles;
        const int ntok = own ? p.tokens : p.bank_tokens;
        const int valid = min(BN, ntok - jj * BN);
        if (valid != BN) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= valid) sr[i] = __float_as_uint(-INFINITY);
        }
        // row max with 8 independent chains (a single 128-long fmax chain costs ~4 clk per element)
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sr[i]);
#pragma unroll
        for (int i = 8; i < BN; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(sr[i]));
        float tmax = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                           fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        tmax *= p.scale_log2;
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_run = tmax;
        } else if (tmax > m_run + kRescaleThreshold) {
          alpha = fast_exp2(m_run - tmax);
          m_run = tmax;
          need = true;
        }
        const bool warp_need = __any_sync(0xffffffffu, need);
        l_run *= alpha;

        // P = exp2(S*c - m), packed to fp16 pairs; row sum in fp32
        uint32_t pk[BN / 2];
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < BN; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2, -m_run));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2, -m_run));
          ls[(i >> 1) & 3] += p0 + p1;
          const __half2 h = __floats2half2_rn(p0, p1);
          pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h);
        }
        l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);

        // previous P.V must have completed before P is overwritten / O is rescaled
        if (g > 0) mbar_wait(pv_done, (g - 1) & 1);
        if (warp_need) {
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DPAD / 32; ++c) {
            uint32_t orr[32];
            tmem_ld_32x32b_x32(tmem_base + Cfg::TMEM_O + t_lane + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
            tmem_st_32x32b_x32(tmem_base + Cfg::TMEM_O + t_lane + c * 32, orr);
          }
          tmem_st_wait();
        }
        // P -> smem (K-major, 128B swizzle: 16B chunk index XOR (row & 7))
#pragma unroll
        for (int c = 0; c < BN / 8; ++c) {
          const int atom = c >> 3, cc = c & 7;
          uint4 val = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(smem_p + atom * (128 * 128) + row * 128 + ((cc ^ (row & 7)) << 4)) = val;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ------------------------------------------------------------------ epilogue: O / l -> fp16 -> global
      mbar_wait(pv_done, (g - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const int q_idx = m_tile * 128 + row;
      const bool row_ok = q_idx < p.tokens;
      __half* dst = p.out + ((long long)frame * p.tokens + q_idx) * p.ldo + head * p.head_dim;
#pragma unroll 1
      for (int c = 0; c < DPAD / 32; ++c) {
        if (c * 32 >= p.head_dim) break;
        uint32_t orr[32];
        tmem_ld_32x32b_x32(tmem_base + Cfg::TMEM_O + t_lane + c * 32, orr);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c * 32 + q * 8 < p.head_dim) {
              __half2 o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = __floats2half2_rn(__uint_as_float(orr[q * 8 + 2 * i]) * inv_l,
                                         __uint_as_float(orr[q * 8 + 2 * i + 1]) * inv_l);
              *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = *reinterpret_cast<uint4*>(o);
            }
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ============================================================================================================
// v3: two 128-query tiles per CTA, ping-pong (10 warps: TMA, MMA, 4 softmax warps per tile; one S buffer per tile at
// TMEM columns [0,128) / [128,256), O_A / O_B behind them, K/V ring shared by both tiles), and P never touches shared
// memory: the softmax warps write the fp16 probabilities back into tensor memory over the first BN/2 columns of their own S tile (tcgen05.st) and the P.V product
// reads its A operand from TMEM (tcgen05.mma "TS" form). Per 128x128 tile this removes 32 KB of st.shared plus 32 KB of
// tensor-core smem reads (of ~144 KB total), which was the co-limiter next to the MUFU exp throughput, and frees 64 KB
// of shared memory for a deeper K/V ring (BN = 128 also for DPAD = 128).
// Ordering relies on (a) tcgen05.mma executing in issue order, (b) tcgen05.commit tracking ALL earlier MMAs of the
// issuing thread: when s_full[t] of tile j fires, P.V of tile j-1 has completed, so S/P columns and O are quiescent.
// ============================================================================================================
template <int DPAD, int BN>
struct Attn3Cfg {
  static constexpr int SLABS = DPAD / 64;
  static constexpr int QT_BYTES = 128 * DPAD * 2;
  static constexpr int Q_BYTES = 2 * QT_BYTES;
  static constexpr int K_BYTES = BN * DPAD * 2;
  static constexpr int KV_STAGE = 2 * K_BYTES;
  static constexpr int BUDGET = 227 * 1024 - 2048 - Q_BYTES;
  static constexpr int STAGES_RAW = BUDGET / KV_STAGE;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * KV_STAGE + 1024 + 256;
  static constexpr uint32_t TMEM_S = 0 /* + t*128 */, TMEM_O = 256 /* + t*DPAD */;
  static_assert(STAGES >= 2, "need at least a double-buffered K/V ring");
  static_assert(DPAD % 64 == 0 && DPAD <= 128 && (BN == 64 || BN == 128), "tile config");
};

template <int DPAD, int BN>
__global__ void __launch_bounds__(320, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmBK,
                  const __grid_constant__ CUtensorMap tmBV, const AttnParams p) {
  using Cfg = Attn3Cfg<DPAD, BN>;
  constexpr int ST = Cfg::STAGES;
  griddep_launch_dependents();   // PDL (ap_host.h::launch_pdl)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem_q + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + ST * Cfg::KV_STAGE);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;            // [ST]
  uint64_t* v_full = k_full + ST;         // [ST]
  uint64_t* kv_empty = v_full + ST;       // [ST]
  uint64_t* s_full = kv_empty + ST;       // [2] per query tile
  uint64_t* p_full = s_full + 2;          // [2]
  uint64_t* o_done = p_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();   // PDL: Q/K/V come from the previous kernel

  const int own_tiles = (p.tokens + BN - 1) / BN;
  const int bank_tiles = (p.bank_tokens + BN - 1) / BN;
  const int m_pairs = (p.tokens + 255) / 256;
  const int units_per_frame = p.heads * m_pairs;
  const int num_units = p.n_frames * units_per_frame;

  auto decode = [&](int unit, int& frame, int& head, int& m_pair, int& T) {
    const int fk = unit / units_per_frame;
    const int rem = unit % units_per_frame;
    frame = (fk + p.first_bank_frame) % p.n_frames;
    head = rem / m_pairs;
    m_pair = rem % m_pairs;
    const bool has_bank = p.bank_tokens > 0 && frame >= p.first_bank_frame;
    T = own_tiles + (has_bank ? bank_tiles : 0);
  };

  if (warp < 2) {
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------------------ TMA producer
      uint32_t g = 0, uc = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++uc) {
        int frame, head, m_pair, T;
        decode(unit, frame, head, m_pair, T);
        mbar_wait(q_empty, (uc & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(&tmQ, q_full, smem_q + t * Cfg::QT_BYTES + s * (128 * 128), head * DPAD + s * 64,
                        frame * p.tokens + m_pair * 256 + t * 128);
        const int bank_idx = (frame - p.first_bank_frame) / p.frames_per_bank;
        for (int j = 0; j < T; ++j, ++g) {
          const int stage = g % ST;
          const uint32_t ph = (g / ST) & 1;
          mbar_wait(&kv_empty[stage], ph ^ 1);
          const bool own = j < own_tiles;
          const CUtensorMap* mk = own ? &tmK : &tmBK;
          const CUtensorMap* mv = own ? &tmV : &tmBV;
          const int row = own ? frame * p.tokens + j * BN : bank_idx * p.bank_tokens + (j - own_tiles) * BN;
          uint8_t* kd = smem_kv + stage * Cfg::KV_STAGE;
          uint8_t* vd = kd + Cfg::K_BYTES;
          mbar_arrive_expect_tx(&k_full[stage], Cfg::K_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(mk, &k_full[stage], kd + s * (BN * 128), head * DPAD + s * 64, row);
          mbar_arrive_expect_tx(&v_full[stage], Cfg::K_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(mv, &v_full[stage], vd + s * (BN * 128), head * DPAD + s * 64, row);
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------------ MMA issuer (warp-uniform loop, one
      // elected lane issues: keeps descriptor / stage arithmetic in uniform registers)
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, BN, 0, 0);
      // the zero padding of the head dimension is not multiplied: Q.K^T stops after ceil(d/16) k-steps and P.V only
      // produces the first 16*ceil(d/16) output columns (d = 80 in a 128-wide slot: 5 of 8 k-steps, N = 80)
      const int k_steps = (p.head_dim + 15) >> 4;
      const uint32_t idesc_pv = umma_idesc_f16(128, (uint32_t)k_steps * 16, 0, 1);
      uint32_t g = 0, uc = 0;
      auto issue_qk = [&](int t, uint32_t gi, bool release_q) {
        const int stage = gi % ST;
        mbar_wait(&k_full[stage], (gi / ST) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + Cfg::TMEM_S + t * 128;
        const uint32_t qa = smem_u32(smem_q + t * Cfg::QT_BYTES);
        const uint32_t ka = smem_u32(smem_kv + stage * Cfg::KV_STAGE);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < DPAD / 16; ++kk) {
            if (kk < k_steps) {
              const uint64_t da = umma_desc_k_sw128(qa + (kk / 4) * (128 * 128) + (kk % 4) * 32);
              const uint64_t db = umma_desc_k_sw128(ka + (kk / 4) * (BN * 128) + (kk % 4) * 32);
              umma_f16_ss(d_tmem, da, db, idesc_qk, kk != 0);
            }
          }
          umma_commit(&s_full[t]);
          if (release_q) umma_commit(q_empty);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, uint32_t gi, int j, bool last) {
        const int stage = gi % ST;
        mbar_wait(&p_full[t], gi & 1);
        mbar_wait(&v_full[stage], (gi / ST) & 1);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + Cfg::TMEM_S + t * 128;   // P aliases the first BN/2 columns of S
        const uint32_t va = smem_u32(smem_kv + stage * Cfg::KV_STAGE + Cfg::K_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            const uint64_t db = umma_desc_mn_sw128(va + kk * (16 * 128), BN * 128);
            umma_f16_ts(tmem_base + Cfg::TMEM_O + t * DPAD, a_tmem + kk * 8, db, idesc_pv, (j | kk) != 0);
          }
          if (t == 1) umma_commit(&kv_empty[stage]);
          if (last) umma_commit(&o_done[t]);
        }
        __syncwarp();
      };
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++uc) {
        int frame, head, m_pair, T;
        decode(unit, frame, head, m_pair, T);
        mbar_wait(q_full, uc & 1);
        issue_qk(0, g, false);
        issue_qk(1, g, T == 1);
        for (int j = 0; j < T; ++j, ++g) {
          issue_pv(0, g, j, j + 1 == T);
          if (j + 1 < T) issue_qk(0, g + 1, false);
          issue_pv(1, g, j, j + 1 == T);
          if (j + 1 < T) issue_qk(1, g + 1, j + 2 == T);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------------------- softmax + epilogue, tile t
    const int t = (warp - 2) >> 2;
    const int lane_group = warp & 3;
    const int row = lane_group * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(lane_group * 32) << 16;
    const uint32_t t_s = tmem_base + Cfg::TMEM_S + t * 128 + t_lane;
    const uint32_t t_o = tmem_base + Cfg::TMEM_O + t * DPAD + t_lane;
    uint32_t g = 0, uc = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++uc) {
      int frame, head, m_pair, T;
      decode(unit, frame, head, m_pair, T);
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j, ++g) {
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        uint32_t sr[BN];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tmem_ld_32x32b_x32(t_s + c * 32, sr + c * 32);
        tmem_ld_wait();

        const bool own = j < own_tiles;
        const int jj = own ? j : j - own_tiles;
        const int ntok = own ? p.tokens : p.bank_tokens;
        const int valid = min(BN, ntok - jj * BN);
        if (valid != BN) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= valid) sr[i] = __float_as_uint(-INFINITY);
        }
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sr[i]);
#pragma unroll
        for (int i = 8; i < BN; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(sr[i]));
        float tmax = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                           fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        tmax *= p.scale_log2;
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_run = tmax;
        } else if (tmax > m_run + kRescaleThreshold) {
          alpha = fast_exp2(m_run - tmax);
          m_run = tmax;
          need = true;
        }
        const bool warp_need = __any_sync(0xffffffffu, need);
        l_run *= alpha;
        if (warp_need) {   // rare: P.V of the previous tile has completed (see header), O is quiescent
#pragma unroll 1
          for (int c = 0; c < DPAD / 32; ++c) {
            uint32_t orr[32];
            tmem_ld_32x32b_x32(t_o + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
            tmem_st_32x32b_x32(t_o + c * 32, orr);
          }
        }

        uint32_t pk[BN / 2];
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < BN; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2, -m_run));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2, -m_run));
          ls[(i >> 1) & 3] += p0 + p1;
          const __half2 h = __floats2half2_rn(p0, p1);
          pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h);
        }
        l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        // P (fp16 pairs) -> TMEM, over the first BN/2 columns of this tile's S
#pragma unroll
        for (int c = 0; c < BN / 64; ++c) tmem_st_32x32b_x32(t_s + c * 32, pk + c * 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ------------------------------------------------------------------ epilogue
      mbar_wait(&o_done[t], uc & 1);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const int q_idx = m_pair * 256 + t * 128 + row;
      const bool row_ok = q_idx < p.tokens;
      __half* dst = p.out + ((long long)frame * p.tokens + q_idx) * p.ldo + head * p.head_dim;
#pragma unroll 1
      for (int c = 0; c < DPAD / 32; ++c) {
        if (c * 32 >= p.head_dim) break;
        uint32_t orr[32];
        tmem_ld_32x32b_x32(t_o + c * 32, orr);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c * 32 + q * 8 < p.head_dim) {
              __half2 o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = __floats2half2_rn(__uint_as_float(orr[q * 8 + 2 * i]) * inv_l,
                                         __uint_as_float(orr[q * 8 + 2 * i + 1]) * inv_l);
              *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = *reinterpret_cast<uint4*>(o);
            }
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int DPAD, int BN>
static int launch_attention3(const CUtensorMap* maps, const AttnParams& p, cudaStream_t stream) {
  using Cfg = Attn3Cfg<DPAD, BN>;
  static bool attr_set = false;
  if (!attr_set) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(attention3_kernel<DPAD, BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int units = p.n_frames * p.heads * ((p.tokens + 255) / 256);
  const int grid = units < num_sms() ? units : num_sms();
  AP_LAUNCH((attention3_kernel<DPAD, BN>), grid, 320, Cfg::SMEM_BYTES, stream, maps[0], maps[1], maps[2],
                                                                                         maps[3], maps[4], p);
  AP_CHECK_CUDA(cudaGetLastError());
  return AP_OK;
}

// ============================================================================================================
// v5: ONE 128-query tile per CTA, TWO CTAs per SM. S is double-buffered in TMEM (Q.K^T of key tile j+1 is issued before
// the softmax of tile j finishes, so the softmax warps are decoupled from the softmax -> P.V -> Q.K^T chain that costs
// v3 ~900 idle cycles per tile), P goes back into its own S buffer (tcgen05.st) and P.V reads it from TMEM (TS form).
// Key tiles of 96 keep the allocation at 256 TMEM columns (S0 [0,96), S1 [96,192), O [192,256)) and ~90 KB of shared
// memory, so two CTAs (2 x 4 softmax warps = two per SM sub-partition) co-reside and hide each other's latencies.
// Only for DPAD = 64 (d <= 64: the 64x64-resolution layers, 88 % of the attention FLOPs).
// ============================================================================================================
struct Attn5Cfg {
  static constexpr int DPAD = 64, BN = 96;
  static constexpr int Q_BYTES = 128 * 128;            // 128 rows x 64 fp16
  static constexpr int K_BYTES = BN * 128;
  static constexpr int KV_STAGE = 2 * K_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * KV_STAGE + 1024 + 256;
  static constexpr uint32_t TMEM_S0 = 0, TMEM_S1 = 96, TMEM_O = 192, TMEM_COLS = 256;
  static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs must fit one SM");
};

__global__ void __launch_bounds__(192, 2)
attention5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmBK,
                  const __grid_constant__ CUtensorMap tmBV, const AttnParams p) {
  using Cfg = Attn5Cfg;
  constexpr int ST = Cfg::STAGES, BN = Cfg::BN, DPAD = Cfg::DPAD;
  griddep_launch_dependents();   // PDL (ap_host.h::launch_pdl)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem_q + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + ST * Cfg::KV_STAGE);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;            // [ST]
  uint64_t* v_full = k_full + ST;         // [ST]
  uint64_t* kv_empty = v_full + ST;       // [ST]
  uint64_t* s_full = kv_empty + ST;       // [2] per S buffer
  uint64_t* p_full = s_full + 2;          // [2] per S buffer
  uint64_t* pv_done = p_full + 2;         // every P.V (rescale path only)
  uint64_t* o_done = pv_done + 1;         // last P.V of a unit
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 4);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();   // PDL: Q/K/V come from the previous kernel

  const int own_tiles = (p.tokens + BN - 1) / BN;
  const int bank_tiles = (p.bank_tokens + BN - 1) / BN;
  const int units_per_frame = p.heads * p.m_tiles;

  auto decode = [&](int unit, int& frame, int& head, int& m_tile, int& T) {
    const int fk = unit / units_per_frame;
    const int rem = unit % units_per_frame;
    frame = (fk + p.first_bank_frame) % p.n_frames;
    head = rem / p.m_tiles;
    m_tile = rem % p.m_tiles;
    const bool has_bank = p.bank_tokens > 0 && frame >= p.first_bank_frame;
    T = own_tiles + (has_bank ? bank_tiles : 0);
  };

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA producer
    uint32_t g = 0, uc = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
      int frame, head, m_tile, T;
      decode(unit, frame, head, m_tile, T);
      mbar_wait(q_empty, (uc & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
        tma_load_2d(&tmQ, q_full, smem_q, head * DPAD, frame * p.tokens + m_tile * 128);
      }
      __syncwarp();
      const int bank_idx = (frame - p.first_bank_frame) / p.frames_per_bank;
      for (int j = 0; j < T; ++j, ++g) {
        const int stage = g % ST;
        mbar_wait(&kv_empty[stage], ((g / ST) & 1) ^ 1);
        if (elect_one()) {
          const bool own = j < own_tiles;
          const CUtensorMap* mk = own ? &tmK : &tmBK;
          const CUtensorMap* mv = own ? &tmV : &tmBV;
          const int row = own ? frame * p.tokens + j * BN : bank_idx * p.bank_tokens + (j - own_tiles) * BN;
          uint8_t* kd = smem_kv + stage * Cfg::KV_STAGE;
          mbar_arrive_expect_tx(&k_full[stage], Cfg::K_BYTES);
          tma_load_2d(mk, &k_full[stage], kd, head * DPAD, row);
          mbar_arrive_expect_tx(&v_full[stage], Cfg::K_BYTES);
          tma_load_2d(mv, &v_full[stage], kd + Cfg::K_BYTES, head * DPAD, row);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer (warp-uniform, elected lane)
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, BN, 0, 0);
    // d = 40 in a 64-wide slot: the zero padding is not multiplied (3 of 4 k-steps for Q.K^T, N = 48 for P.V)
    const int k_steps = (p.head_dim + 15) >> 4;
    const uint32_t idesc_pv = umma_idesc_f16(128, (uint32_t)k_steps * 16, 0, 1);
    const uint32_t qa = smem_u32(smem_q);
    uint32_t g = 0, uc = 0;
    auto issue_qk = [&](uint32_t gi, bool last_of_unit) {
      const int stage = gi % ST;
      mbar_wait(&k_full[stage], (gi / ST) & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ((gi & 1) ? Cfg::TMEM_S1 : Cfg::TMEM_S0);
      const uint32_t ka = smem_u32(smem_kv + stage * Cfg::KV_STAGE);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk)
          if (kk < k_steps)
            umma_f16_ss(d_tmem, umma_desc_k_sw128(qa + kk * 32), umma_desc_k_sw128(ka + kk * 32), idesc_qk, kk != 0);
        umma_commit(&s_full[gi & 1]);
        if (last_of_unit) umma_commit(q_empty);
      }
      __syncwarp();
    };
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
      int frame, head, m_tile, T;
      decode(unit, frame, head, m_tile, T);
      mbar_wait(q_full, uc & 1);
      for (int j = 0; j < T; ++j, ++g) {
        // S buffer (g+1)&1 last held P(g-1), consumed by P.V(g-1) which was issued earlier (tcgen05.mma runs in order)
        if (j == 0) issue_qk(g, T == 1);
        if (j + 1 < T) issue_qk(g + 1, j + 2 == T);
        const int stage = g % ST;
        mbar_wait(&p_full[g & 1], (g >> 1) & 1);
        mbar_wait(&v_full[stage], (g / ST) & 1);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + ((g & 1) ? Cfg::TMEM_S1 : Cfg::TMEM_S0);   // P aliases its S buffer
        const uint32_t va = smem_u32(smem_kv + stage * Cfg::KV_STAGE + Cfg::K_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_f16_ts(tmem_base + Cfg::TMEM_O, a_tmem + kk * 8, umma_desc_mn_sw128(va + kk * (16 * 128), BN * 128),
                        idesc_pv, (j | kk) != 0);
          umma_commit(&kv_empty[stage]);
          umma_commit(pv_done);
          if (j + 1 == T) umma_commit(o_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int lane_group = warp & 3;
    const int row = lane_group * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(lane_group * 32) << 16;
    const uint32_t t_o = tmem_base + Cfg::TMEM_O + t_lane;
    uint32_t g = 0, uc = 0;
    uint32_t sr[BN];
    // (Round 2: delaying the softmax warps of every second CTA of an SM by 400..1400 clocks, and halving the row-maximum
    // instructions with FMNMX3, both left the launch at 2.22 ms; eight softmax warps per CTA (two per TMEM lane quarter, each
    // owning half of a key tile's columns, row maxima exchanged through shared memory) measured 2.52 ms. The variants were
    // removed after the measurements: profiles/r02_summary.md.)
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
      int frame, head, m_tile, T;
      decode(unit, frame, head, m_tile, T);
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j, ++g) {
        const uint32_t t_s = tmem_base + ((g & 1) ? Cfg::TMEM_S1 : Cfg::TMEM_S0) + t_lane;
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tmem_ld_32x32b_x32(t_s + c * 32, sr + c * 32);
        tmem_ld_wait();
        const bool own = j < own_tiles;
        const int jj = own ? j : j - own_tiles;
        const int ntok = own ? p.tokens : p.bank_tokens;
        const int valid = min(BN, ntok - jj * BN);
        if (valid != BN) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= valid) sr[i] = __float_as_uint(-INFINITY);
        }
        float tmax = row_max<BN>(sr) * p.scale_log2;
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_run = tmax;
        } else if (tmax > m_run + kRescaleThreshold) {
          alpha = fast_exp2(m_run - tmax);
          m_run = tmax;
          need = true;
        }
        const bool warp_need = __any_sync(0xffffffffu, need);
        l_run *= alpha;
        if (warp_need) {   // rare: O may only be touched once P.V of the previous key tile has completed
          mbar_wait(pv_done, (g - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DPAD / 32; ++c) {
            uint32_t orr[32];
            tmem_ld_32x32b_x32(t_o + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
            tmem_st_32x32b_x32(t_o + c * 32, orr);
          }
        }
        uint32_t pk[BN / 2];
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < BN; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2, -m_run));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2, -m_run));
          ls[(i >> 1) & 3] += p0 + p1;
          const __half2 h = __floats2half2_rn(p0, p1);
          pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h);
        }
        l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        // P (48 packed columns) -> TMEM over the first half of this S buffer
        tmem_st_32x32b_x32(t_s, pk);
        tmem_st_32x32b_x16(t_s + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1]);
      }
      // ------------------------------------------------------------------ epilogue
      mbar_wait(o_done, uc & 1);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const int q_idx = m_tile * 128 + row;
      const bool row_ok = q_idx < p.tokens;
      __half* dst = p.out + ((long long)frame * p.tokens + q_idx) * p.ldo + head * p.head_dim;
#pragma unroll 1
      for (int c = 0; c < DPAD / 32; ++c) {
        if (c * 32 >= p.head_dim) break;
        uint32_t orr[32];
        tmem_ld_32x32b_x32(t_o + c * 32, orr);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c * 32 + q * 8 < p.head_dim) {
              __half2 o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = __floats2half2_rn(__uint_as_float(orr[q * 8 + 2 * i]) * inv_l,
                                         __uint_as_float(orr[q * 8 + 2 * i + 1]) * inv_l);
              *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = *reinterpret_cast<uint4*>(o);
            }
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

static int launch_attention5(const CUtensorMap* maps, const AttnParams& p, cudaStream_t stream) {
  using Cfg = Attn5Cfg;
  static bool attr_set = false;
  if (!attr_set) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(attention5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int max_ctas = 2 * num_sms();
  const int grid = p.num_units < max_ctas ? p.num_units : max_ctas;
  AP_LAUNCH((attention5_kernel), grid, 192, Cfg::SMEM_BYTES, stream, maps[0], maps[1], maps[2], maps[3], maps[4], p);
  AP_CHECK_CUDA(cudaGetLastError());
  return AP_OK;
}

template <int DPAD, int BN>
static int launch_attention(const CUtensorMap* maps, const AttnParams& p, cudaStream_t stream) {
  using Cfg = AttnCfg<DPAD, BN>;
  static bool attr_set = false;
  if (!attr_set) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<DPAD, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = p.num_units < num_sms() ? p.num_units : num_sms();
  AP_LAUNCH((attention_kernel<DPAD, BN>), grid, 192, Cfg::SMEM_BYTES, stream, maps[0], maps[1], maps[2], maps[3], maps[4], p);
  AP_CHECK_CUDA(cudaGetLastError());
  return AP_OK;
}

}  // namespace ap

using namespace ap;

extern "C" int ap_attention_f16(const void* q, const void* k, const void* v, long long ld_qkv, const void* bank_k,
                                const void* bank_v, long long ld_bank, int bank_tokens, int n_banks, int n_frames,
                                int tokens, int heads, int head_dim, int dpad, int first_bank_frame,
                                int frames_per_bank, float scale, void* out, long long ldo, void* stream) {
  AP_REQUIRE(q && k && v && out, "attention: null pointer");
  AP_REQUIRE(dpad == 64 || dpad == 128 || dpad == 192, "attention: dpad must be 64/128/192 (got %d)", dpad);
  AP_REQUIRE(head_dim % 8 == 0 && head_dim <= dpad, "attention: head_dim %d must be a multiple of 8 and <= dpad", head_dim);
  AP_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be multiples of 8 elements");
  AP_REQUIRE(n_frames > 0 && tokens > 0 && heads > 0, "attention: bad shape");
  const bool has_bank = bank_k != nullptr && bank_tokens > 0;
  AP_REQUIRE(!has_bank || (bank_v && frames_per_bank > 0 && first_bank_frame >= 0 && n_banks > 0),
             "attention: bad bank arguments");
  // v3 (two query tiles per CTA, P kept in TMEM) for dpad <= 128; v1 for dpad = 192 / tiny token counts.
  // AP_ATTENTION_V1=1 / AP_ATTENTION_V2=1 select the older variants (A/B timing).
  static const bool force_v1 = (getenv("AP_ATTENTION_V1") != nullptr);
  // kernel generations (measurements: profiles/r01_ncu_full_top_kernels.md):
  //   v5  d <= 64 and more than one query tile: one query tile per CTA, two CTAs per SM, S double-buffered, P in TMEM
  //   v3  d <= 128: two query tiles per CTA, P in TMEM
  //   v1  d <= 192 or at most 128 tokens: one query tile per CTA, S double-buffered, P through shared memory
  // AP_ATTENTION_V5=0 / AP_ATTENTION_V1=1 fall back to the older generation for A/B timing.
  static const int v5_env = getenv("AP_ATTENTION_V5") ? atoi(getenv("AP_ATTENTION_V5")) : -1;
  const bool two_tiles = !force_v1 && dpad <= 128 && tokens > 128;
  const bool use_v5 = two_tiles && v5_env != 0 && dpad == 64;
  const bool use_v3 = two_tiles && !use_v5;
  const int bn = use_v5 ? 96 : (use_v3 ? 128 : (dpad == 192 ? 64 : 128));

  AttnParams p{};
  p.n_frames = n_frames;
  p.tokens = tokens;
  p.heads = heads;
  p.head_dim = head_dim;
  p.bank_tokens = has_bank ? bank_tokens : 0;
  p.first_bank_frame = has_bank ? first_bank_frame : 0;
  p.frames_per_bank = has_bank ? frames_per_bank : 1;
  p.m_tiles = (tokens + 127) / 128;
  p.num_units = n_frames * heads * p.m_tiles;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = (__half*)out;
  p.ldo = ldo;
  if (has_bank) {
    const int max_bank = (n_frames - 1 - first_bank_frame) / frames_per_bank;
    AP_REQUIRE(first_bank_frame >= n_frames || max_bank < n_banks, "attention: bank index out of range");
  }

  CUtensorMap maps[5];
  const uint64_t cols = (uint64_t)heads * dpad;
  auto mk = [&](CUtensorMap* tm, const void* base, uint64_t rows, long long ld, uint32_t box_rows) -> int {
    const uint64_t dims[2] = {cols, rows};
    const uint64_t strides[1] = {(uint64_t)ld * 2};
    const uint32_t box[2] = {64, box_rows};
    return encode_tmap(tm, base, 2, dims, strides, box, true);
  };
  int rc;
  const uint64_t rows = (uint64_t)n_frames * tokens;
  if ((rc = mk(&maps[0], q, rows, ld_qkv, 128))) return rc;
  if ((rc = mk(&maps[1], k, rows, ld_qkv, bn))) return rc;
  if ((rc = mk(&maps[2], v, rows, ld_qkv, bn))) return rc;
  if (has_bank) {
    const uint64_t brows = (uint64_t)n_banks * bank_tokens;
    if ((rc = mk(&maps[3], bank_k, brows, ld_bank, bn))) return rc;
    if ((rc = mk(&maps[4], bank_v, brows, ld_bank, bn))) return rc;
  } else {
    maps[3] = maps[1];
    maps[4] = maps[2];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (use_v5) return launch_attention5(maps, p, st);
  if (use_v3) return dpad == 64 ? launch_attention3<64, 128>(maps, p, st) : launch_attention3<128, 128>(maps, p, st);
  if (dpad == 64) return launch_attention<64, 128>(maps, p, st);
  if (dpad == 128) return launch_attention<128, 128>(maps, p, st);
  return launch_attention<192, 64>(maps, p, st);
}
