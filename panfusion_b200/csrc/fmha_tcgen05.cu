// Flash-attention forward on tcgen05: S = Q K^T and O += P V on the 5th-gen tensor cores (accumulators in
// TMEM), TMA-fed K/V ring, online softmax by one warpgroup (thread <-> query row), optional additive fp32 bias
// shared by all heads (the EPPA correspondence bias).
//
// Replaces: xformers.ops.memory_efficient_attention(q, k, v, attn_bias) at models/modules/transformer.py:71
// (EPPA, head dim 32, dense bias repeated per head at :68 — here never repeated) and the diffusers AttnProcessor
// bmm-softmax-bmm inside Transformer2DModel (MVGenModel.py:104,116,185,190,227,241; head dim 64, self and text
// cross attention).
//
// Tile: 128 queries x 64 keys. S is double-buffered in TMEM so Q K^T of tile j+1 runs under the softmax of tile j;
// P goes registers -> swizzled smem (K-major A operand), P V lands in a TMEM scratch tile that the softmax
// threads fold into their fp32 register accumulator with the online-softmax rescale.
// 192 threads: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..5 softmax / epilogue.
#include <stdlib.h>

#include "pf_common.cuh"

namespace pf {

constexpr int FA_BLOCK_M = 128;
constexpr int FA_BLOCK_N = 64;
constexpr int FA_THREADS = 192;
// Measured on B200 (scripts/fmha_micro.py, 16x5x4096x4096 d64): MUFU-only + cvt packing 668 TFLOP/s; moving every
// 4th exponential to the FMA pipe and the bf16 packing to the ALU: 615 TFLOP/s (issue slots, not the XU pipe, are
// the binding constraint at 3 CTAs/SM). Both tricks are kept switchable for the next round's re-tuning.
#ifndef PF_FA_POLY_EXP
#define PF_FA_POLY_EXP 0
#endif
#ifndef PF_FA_ALU_PACK
#define PF_FA_ALU_PACK 0
#endif
#ifndef PF_FA_POLY_DEFAULT
#define PF_FA_POLY_DEFAULT 0
#endif
constexpr bool FA_POLY_EXP = PF_FA_POLY_EXP != 0;
constexpr bool FA_ALU_PACK = PF_FA_ALU_PACK != 0;

struct FmhaParams {
  int B, H, Lq, Lk;
  float scale_log2;  // softmax scale * log2(e)
  void* out;         // [B, Lq, out_ld] 16-bit, head h at columns [h*D, (h+1)*D)
  int out_ld;
  const float* bias;  // [bias_batches, Lq, bias_ld] or null
  long long bias_bstride;
  int bias_ld;
  // optional: per (128-query x 64-key) tile flag, 1 = every bias entry of the tile equals -1 (no correspondence):
  // the tile's loads are replaced by the constant. [bias_batches, ceil(Lq/128), ceil(Lk/64)] bytes.
  const uint8_t* bias_flags;
  long long flags_bstride;
  int flags_ld;
  // tile-packed bias (pf_bias_tile_pack): tile_off[bias batch][ceil(Lq/128)][ceil(Lk/64)] = index of the 128x64 tile in
  // `bias` (then a [n_live][128][64] store) or -1 for an all -1 tile; replaces bias_ld / bias_flags addressing
  const int* tile_off;
};

// ex2.approx.ftz: one MUFU op (exp2f() adds denormal / range fix-up instructions we do not need: inputs are <= 0)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax of 2^f on [-0.5, 0.5], max rel. error 7.5e-5 — far
// below the 16-bit rounding of P). The MUFU (XU) pipe is the busiest unit of this kernel (ncu: 64 %), the FMA pipe
// the idlest (19 %): evaluating every 4th probability here moves a quarter of the exponentials off the bottleneck.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;   // 1.5 * 2^23: round-to-nearest integer lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517164f, 0.24261112f);
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992807f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// exp2 of two values at once on the FMA / ALU pipes (same Cody-Waite split and cubic as poly_exp2): 2 FMNMX + 2 FADD2 +
// 4 FFMA2 + 2 LEA = 10 issue slots for two exponentials and NO MUFU cycles (an ex2 occupies the quarter-rate XU pipe for 8
// cycles per warp). Used for a fixed fraction of the key pairs of every tile (POLY_PAIRS of 8) to balance XU against issue.
__device__ __forceinline__ void poly_exp2_pair(float x0, float x1, float& y0, float& y1) {
  const f32x2 x = pack_f2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const f32x2 magic = pack_f2(12582912.0f, 12582912.0f);
  const f32x2 t = add_f2(x, magic);
  const f32x2 r = add_f2(t, pack_f2(-12582912.0f, -12582912.0f));
  const f32x2 f = fma_f2(r, pack_f2(-1.0f, -1.0f), x);
  f32x2 p = fma_f2(f, pack_f2(0.05517164f, 0.05517164f), pack_f2(0.24261112f, 0.24261112f));
  p = fma_f2(p, f, pack_f2(0.69326099f, 0.69326099f));
  p = fma_f2(p, f, pack_f2(0.99992807f, 0.99992807f));
  float p0, p1, t0, t1;
  unpack_f2(p, p0, p1);
  unpack_f2(t, t0, t1);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// two non-negative fp32 -> packed bf16x2 by integer rounding (half-up; ties differ from RNE by 1 ulp with
// probability 2^-16). cvt.rn.bf16x2.f32 issues on the XU pipe next to the exponentials; this stays on the ALU.
__device__ __forceinline__ uint32_t pack_bf16_alu(float a, float b) {
  const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
  return __byte_perm(ua, ub, 0x7632);
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack_prob(float a, float b) {
  if constexpr (BF16) return pack_bf16_alu(a, b);
  else return pack2<false>(a, b);
}

template <int D>
__host__ __device__ constexpr int fa_stages() { return D == 64 ? 2 : 4; }
template <int D>
__host__ __device__ constexpr int fmha_smem_bytes() {
  return FA_BLOCK_M * D * 2 + fa_stages<D>() * 2 * FA_BLOCK_N * D * 2 + FA_BLOCK_M * FA_BLOCK_N * 2 + 256;
}

// O stays in TMEM for the whole KV loop (P V accumulates in place). The running maximum used for scaling is only
// raised when a row's new maximum exceeds it by more than 2^8 ("lazy rescale"): softmax is shift-invariant, so any
// reference value gives the same result as long as exp2 stays in range, and the O / l rescale (TMEM load-scale-store)
// becomes a rare event instead of per-tile work. Without a per-thread fp32 O accumulator the kernel fits three CTAs
// per SM (112 registers, 64 KB smem, 128 TMEM columns each): three softmax warps per scheduler hide the
// exp / TMEM / barrier latencies of one another.
template <int D, bool BF16, bool HAS_BIAS, int POLY_PAIRS>
__global__ void __launch_bounds__(FA_THREADS, 3)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  static_assert(D == 32 || D == 64, "head dim 32 (EPPA) or 64 (SD-2 UNet)");
  constexpr int STAGES = fa_stages<D>();
  constexpr int Q_BYTES = FA_BLOCK_M * D * 2;
  constexpr int KV_BYTES = FA_BLOCK_N * D * 2;  // one of K or V
  constexpr int P_BYTES = FA_BLOCK_M * FA_BLOCK_N * 2;
  constexpr uint32_t SW_LAYOUT = (D == 64) ? 2u : 4u;           // 128B / 64B swizzle
  constexpr uint32_t SW_ATOM_BYTES = (D == 64) ? 1024u : 512u;  // 8 rows of D*2 bytes
  constexpr int TMEM_COLS = 128;
  constexpr uint32_t TM_S = 0, TM_O = 64;
  constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Q_BYTES;                   // stage s: K at s*2*KV_BYTES, V right after
  uint8_t* sP = sKV + STAGES * 2 * KV_BYTES;     // [128][64] 16-bit, SW128 K-major
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + STAGES;
  uint64_t* s_full = kv_empty + STAGES;
  uint64_t* s_free = s_full + 1;   // 128 arrivals: S has been read out of TMEM
  uint64_t* p_full = s_free + 1;   // 128 arrivals: P is in shared memory (and O was rescaled if needed)
  uint64_t* o_done = p_full + 1;   // P V of the tile has completed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H;  // heads fastest: CTAs sharing a bias tile run together (L2 reuse)
  const int qt = blockIdx.x / p.H;
  const int b = blockIdx.y;
  const int q0 = qt * FA_BLOCK_M;
  const int n_tiles = (p.Lk + FA_BLOCK_N - 1) / FA_BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // Q/K/V (and the bias flags) written by the predecessor are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * KV_BYTES);
        uint8_t* dst = sKV + s * 2 * KV_BYTES;
        tma_load_4d(dst, &tmK, &kv_full[s], 0, h, j * FA_BLOCK_N, b);
        tma_load_4d(dst + KV_BYTES, &tmV, &kv_full[s], 0, h, j * FA_BLOCK_N, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(BF16 ? 1 : 0, FA_BLOCK_M, FA_BLOCK_N, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(BF16 ? 1 : 0, FA_BLOCK_M, D, 0, 1);  // V is MN-major
      const uint64_t qdesc = make_smem_desc(smem_u32(sQ), 16, SW_ATOM_BYTES, SW_LAYOUT);
      const uint64_t pdesc = make_smem_desc(smem_u32(sP), 16, 1024, 2);
      auto issue_qk = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&kv_full[s], (j / STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = make_smem_desc(smem_u32(sKV + s * 2 * KV_BYTES), 16, SW_ATOM_BYTES, SW_LAYOUT);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base + TM_S, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        // S_j has been copied to registers: Q K^T of the next tile may overwrite it while the softmax runs
        mbar_wait(s_free, j & 1);
        tc_fence_after();
        if (j + 1 < n_tiles) issue_qk(j + 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int s = j % STAGES;
        // V tile [64 keys][D]: MN-major B operand; 8-key groups are SW_ATOM_BYTES apart (SBO); one atom along N
        const uint64_t vdesc =
            make_smem_desc(smem_u32(sKV + s * 2 * KV_BYTES + KV_BYTES), SW_ATOM_BYTES, SW_ATOM_BYTES, SW_LAYOUT);
#pragma unroll
        for (int k = 0; k < FA_BLOCK_N / 16; ++k) {
          // P: +32 B per 16 keys inside the 128 B swizzle row; V: +16 key rows = 2 swizzle atoms
          umma_f16(tmem_base + TM_O, pdesc + 2 * k, vdesc + uint64_t((2 * SW_ATOM_BYTES * k) >> 4), idesc_pv,
                   (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(o_done);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    float m_used = -INFINITY, l_run = 0.f;
    const float* bias_row = nullptr;
    const uint8_t* flag_row = nullptr;
    const int* off_row = nullptr;
    if constexpr (HAS_BIAS) {
      const int qq = q < p.Lq ? q : p.Lq - 1;
      if (p.tile_off) {
        off_row = p.tile_off + (long long)b * p.flags_bstride + (long long)qt * p.flags_ld;
        bias_row = p.bias + (quarter * (FA_BLOCK_N / 4) * 32 + lane) * 4;  // lane-interleaved tile; + tile index * 128 * 64
      } else {
        bias_row = p.bias + (long long)b * p.bias_bstride + (long long)qq * p.bias_ld;
        if (p.bias_flags) flag_row = p.bias_flags + (long long)b * p.flags_bstride + (long long)qt * p.flags_ld;
      }
    }

    for (int j = 0; j < n_tiles; ++j) {
      const int k0 = j * FA_BLOCK_N;
      // this tile's entry of the packed-bias index (fetching it one iteration ahead measured slower: one more live register)
      int toff = 0;
      if constexpr (HAS_BIAS) {
        if (off_row) toff = off_row[j];
      }
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float sv[FA_BLOCK_N];
      {
        uint32_t raw[32];
        tmem_ld32(lane_addr + TM_S, raw);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) sv[e] = __uint_as_float(raw[e]);
        tmem_ld32(lane_addr + TM_S + 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) sv[32 + e] = __uint_as_float(raw[e]);
      }
      tc_fence_before();
      mbar_arrive(s_free);
      // Logits in log2 units are sv * sc + shift. Without bias, and on a tile whose bias is the constant -1, the raw scores
      // stay in sv and (sc, shift) = (scale*log2e, 0 or -log2e) are folded into the exp2 argument below (one FFMA per
      // element); a live bias tile is added here, t = s*scale*log2e + bias*log2e, and continues with (1, 0).
      float sc = p.scale_log2, shift = 0.f;
      if constexpr (HAS_BIAS) {
        const bool constant_tile = off_row ? toff < 0 : (flag_row != nullptr && flag_row[j] != 0);
        if (constant_tile) {
          shift = -LOG2E;
        } else if (off_row || k0 + FA_BLOCK_N <= p.Lk) {
          // A packed tile is lane-interleaved (pf_bias_tile_pack): 16-byte piece e of this warp's 32 rows is 512 contiguous
          // bytes, one fully coalesced request (a row-major tile costs 32 sectors per request: measured 269 -> 205 us at
          // the C2 level-32 direction-1 shape). Packed tiles are always full 64-wide (zero beyond Lk: masked below).
          const float4* b4 = off_row ? reinterpret_cast<const float4*>(bias_row + (long long)toff * (FA_BLOCK_M * FA_BLOCK_N))
                                     : reinterpret_cast<const float4*>(bias_row + k0);
          const int st4 = off_row ? 32 : 1;
#pragma unroll
          for (int e = 0; e < FA_BLOCK_N / 4; ++e) {
            const float4 t = __ldg(b4 + e * st4);
            sv[4 * e + 0] = fmaf(t.x, LOG2E, sv[4 * e + 0] * p.scale_log2);
            sv[4 * e + 1] = fmaf(t.y, LOG2E, sv[4 * e + 1] * p.scale_log2);
            sv[4 * e + 2] = fmaf(t.z, LOG2E, sv[4 * e + 2] * p.scale_log2);
            sv[4 * e + 3] = fmaf(t.w, LOG2E, sv[4 * e + 3] * p.scale_log2);
          }
          sc = 1.0f;
        } else {  // ragged last key tile of a dense table
#pragma unroll
          for (int e = 0; e < FA_BLOCK_N; ++e) {
            const float bv = (k0 + e < p.Lk) ? __ldg(bias_row + k0 + e) : 0.f;
            sv[e] = fmaf(bv, LOG2E, sv[e] * p.scale_log2);
          }
          sc = 1.0f;
        }
      }
      if (k0 + FA_BLOCK_N > p.Lk) {
#pragma unroll
        for (int e = 0; e < FA_BLOCK_N; ++e)
          if (k0 + e >= p.Lk) sv[e] = -INFINITY;
      }
      // tile maximum: 8 independent chains
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = sv[i];
#pragma unroll
      for (int e = 8; e < FA_BLOCK_N; ++e) mxs[e & 7] = fmaxf(mxs[e & 7], sv[e]);
      const float mx_raw = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                                 fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      const float mx_tile = HAS_BIAS ? fmaf(mx_raw, sc, shift) : sc * mx_raw;
      // lazy rescale: raise the reference only when this row would otherwise exceed 2^8
      const bool need = mx_tile > m_used + RESCALE_THRESHOLD;
      bool o_waited = false;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx_tile : m_used;
        const float alpha = fast_exp2(m_used - m_new);  // 1 for rows that keep their reference, 0 on the first tile
        if (j > 0) {
          mbar_wait(o_done, (j - 1) & 1);  // P V of tile j-1 has landed in O
          tc_fence_after();
          o_waited = true;
          uint32_t raw[16];  // 16-column pieces: this path runs with the 64 logits of the tile live in registers
#pragma unroll 1
          for (int c = 0; c < D; c += 16) {
            tmem_ld16(lane_addr + TM_O + c, raw);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) raw[e] = __float_as_uint(__uint_as_float(raw[e]) * alpha);
            tmem_st16(lane_addr + TM_O + c, raw);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        m_used = m_new;
      }
      // exp2(s * scale - m) for two neighbouring keys per FFMA2, row sums on FADD2 (packed fp32x2, sm_100): the softmax
      // warps share their issue slots with the MUFU pipe that bounds this kernel, so every FMA-pipe instruction saved counts
      f32x2 ps2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ps2[i] = pack_f2(0.f, 0.f);
      const f32x2 sc2 = pack_f2(sc, sc), nm2 = HAS_BIAS ? pack_f2(shift - m_used, shift - m_used) : pack_f2(-m_used, -m_used);
      uint32_t pk[FA_BLOCK_N / 2];
#pragma unroll
      for (int e = 0; e < FA_BLOCK_N; e += 2) {
        float a0, a1;
        unpack_f2(fma_f2(pack_f2(sv[e], sv[e + 1]), sc2, nm2), a0, a1);
        float p0, p1;
        if (((e >> 1) & 7) < POLY_PAIRS) {  // compile-time after unrolling: this pair's exponentials skip the MUFU
          poly_exp2_pair(a0, a1, p0, p1);
        } else {
          p0 = fast_exp2(a0);
          p1 = (FA_POLY_EXP && ((e >> 1) & 1)) ? poly_exp2(a1) : fast_exp2(a1);
        }
        ps2[(e >> 1) & 3] = add_f2(ps2[(e >> 1) & 3], pack_f2(p0, p1));
        pk[e >> 1] = FA_ALU_PACK ? pack_prob<BF16>(p0, p1) : pack2<BF16>(p0, p1);
      }
      {
        float lo, hi;
        unpack_f2(add_f2(add_f2(ps2[0], ps2[1]), add_f2(ps2[2], ps2[3])), lo, hi);
        l_run += lo + hi;
      }
      if (j > 0 && !o_waited) {
        mbar_wait(o_done, (j - 1) & 1);  // the tensor core is done reading sP (tile j-1)
        tc_fence_after();
      }
      // P row -> swizzled smem: 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
      {
        uint8_t* prow = sP + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 v = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // normalise and store
    mbar_wait(o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    {
      const float inv = 1.0f / l_run;
      uint16_t* orow = static_cast<uint16_t*>(p.out) + ((long long)b * p.Lq + q) * p.out_ld + h * D;
      uint32_t raw[32];
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        tmem_ld32(lane_addr + TM_O + c, raw);
        tmem_ld_wait();
        if (q < p.Lq) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 v = make_uint4(pack2<BF16>(__uint_as_float(raw[e]) * inv, __uint_as_float(raw[e + 1]) * inv),
                                 pack2<BF16>(__uint_as_float(raw[e + 2]) * inv, __uint_as_float(raw[e + 3]) * inv),
                                 pack2<BF16>(__uint_as_float(raw[e + 4]) * inv, __uint_as_float(raw[e + 5]) * inv),
                                 pack2<BF16>(__uint_as_float(raw[e + 6]) * inv, __uint_as_float(raw[e + 7]) * inv));
            *reinterpret_cast<uint4*>(orow + c + e) = v;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int make_qkv_tmap(CUtensorMap* tm, int dtype, const void* ptr, int B, int H, int L, int D, int ld,
                         long long bstride, int box_rows) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)D * 2, (uint64_t)ld * 2, (uint64_t)bstride * 2};
  uint32_t box[4] = {(uint32_t)D, 1, (uint32_t)box_rows, 1};
  return make_tmap(tm, dtype, 4, ptr, dims, str, box, D == 64 ? 128 : 64);
}

template <int D, bool BF16, bool HAS_BIAS, int POLY_PAIRS>
static int launch_fmha(const pf_fmha_args* a, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_qkv_tmap(&tmQ, a->dtype, a->q, a->B, a->H, a->Lq, D, a->q_ld, a->q_bstride, FA_BLOCK_M))) return rc;
  if ((rc = make_qkv_tmap(&tmK, a->dtype, a->k, a->B, a->H, a->Lk, D, a->k_ld, a->k_bstride, FA_BLOCK_N))) return rc;
  if ((rc = make_qkv_tmap(&tmV, a->dtype, a->v, a->B, a->H, a->Lk, D, a->v_ld, a->v_bstride, FA_BLOCK_N))) return rc;
  FmhaParams p;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = a->out; p.out_ld = a->out_ld;
  p.bias = a->bias; p.bias_bstride = a->bias_bstride; p.bias_ld = a->bias_ld;
  p.bias_flags = a->bias_flags; p.flags_bstride = a->flags_bstride; p.flags_ld = a->flags_ld;
  p.tile_off = a->bias_tile_off;
  auto kern = fmha_fwd_kernel<D, BF16, HAS_BIAS, POLY_PAIRS>;
  constexpr int SMEM = fmha_smem_bytes<D>();
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM),
                    "cudaFuncSetAttribute(fmha)");
    if (rc) return rc;
    attr_set = true;
  }
  dim3 grid(((a->Lq + FA_BLOCK_M - 1) / FA_BLOCK_M) * a->H, a->B);
  if ((rc = check_cuda(launch_pdl(kern, grid, dim3(FA_THREADS), SMEM, st, tmQ, tmK, tmV, p), "launch(fmha)"))) return rc;
  PF_CHECK_LAUNCH("fmha_fwd_kernel");
  return PF_OK;
}

}  // namespace pf

extern "C" int pf_fmha_fwd(const pf_fmha_args* a, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(a != nullptr, "pf_fmha_fwd: null args");
  PF_CHECK_ARG(a->dtype == PF_BF16 || a->dtype == PF_F16, "pf_fmha_fwd: dtype must be PF_F16 or PF_BF16");
  PF_CHECK_ARG(a->q && a->k && a->v && a->out, "pf_fmha_fwd: null operand");
  PF_CHECK_ARG(a->head_dim == 32 || a->head_dim == 64, "pf_fmha_fwd: head_dim %d unsupported (32 or 64)", a->head_dim);
  PF_CHECK_ARG(a->B > 0 && a->B <= 65535 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "pf_fmha_fwd: empty shape");
  PF_CHECK_ARG(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->out_ld % 8 == 0,
               "pf_fmha_fwd: leading dims must be multiples of 8 elements");
  PF_CHECK_ARG(((uintptr_t)a->q & 15) == 0 && ((uintptr_t)a->k & 15) == 0 && ((uintptr_t)a->v & 15) == 0 &&
                   ((uintptr_t)a->out & 15) == 0,
               "pf_fmha_fwd: operands must be 16-byte aligned");
  PF_CHECK_ARG(!a->bias || a->bias_tile_off || (a->bias_ld % 4 == 0 && ((uintptr_t)a->bias & 15) == 0 && a->bias_ld >= a->Lk),
               "pf_fmha_fwd: bias must be 16-byte aligned with bias_ld %% 4 == 0");
  PF_CHECK_ARG(!a->bias_tile_off || (a->bias && ((uintptr_t)a->bias & 15) == 0 && !a->bias_flags && a->flags_ld > 0),
               "pf_fmha_fwd: a tile-packed bias needs the packed store in `bias`, flags_ld = tiles per row, no bias_flags");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bf = a->dtype == PF_BF16;
  const bool hb = a->bias != nullptr;
  // PF_FA_POLY = how many of every 8 key pairs evaluate their exponentials on the FMA pipe (0 = all on the MUFU); measured
  // default below. Only the head-dim-64 bf16 kernels — the ones that matter for the step — carry the variants.
  static const int poly = [] {
    const char* e = getenv("PF_FA_POLY");
    return e ? atoi(e) : PF_FA_POLY_DEFAULT;
  }();
  if (a->head_dim == 64) {
    if (bf && !hb) {
      switch (poly) {
        case 2: return launch_fmha<64, true, false, 2>(a, st);
        case 3: return launch_fmha<64, true, false, 3>(a, st);
        case 4: return launch_fmha<64, true, false, 4>(a, st);
        default: return launch_fmha<64, true, false, 0>(a, st);
      }
    }
    if (bf) return launch_fmha<64, true, true, 0>(a, st);
    return hb ? launch_fmha<64, false, true, 0>(a, st) : launch_fmha<64, false, false, 0>(a, st);
  }
  if (bf) {
    if (hb) {
      switch (poly) {
        case 2: return launch_fmha<32, true, true, 2>(a, st);
        case 3: return launch_fmha<32, true, true, 3>(a, st);
        default: return launch_fmha<32, true, true, 0>(a, st);
      }
    }
    return launch_fmha<32, true, false, 0>(a, st);
  }
  return hb ? launch_fmha<32, false, true, 0>(a, st) : launch_fmha<32, false, false, 0>(a, st);
}
