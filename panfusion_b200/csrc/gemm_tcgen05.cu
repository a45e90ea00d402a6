// Tap-GEMM on tcgen05: TMA (SWIZZLE_128B) -> shared -> tcgen05.mma (UMMA 128 x BLOCK_N x 16, fp32 in TMEM)
// -> tcgen05.ld epilogue. One kernel serves nn.Linear and every 1x1 / 3x3 convolution of the UNet walk
// (reference call sites: models/pano/MVGenModel.py:85-295 through diffusers ResnetBlock2D / Transformer2DModel,
//  models/modules/transformer.py:57-74,8-35). A convolution is a sum of `num_taps` GEMMs whose A operand is the
// same channels-last image shifted by a constant row offset (zero-haloed "padded-flat" layout), so the im2col
// matrix is never formed: each K-slab is one plain 2-D TMA box.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..5 = epilogue (thread <-> accumulator row, TMEM lane quarter = warp_id % 4).
#include <stdlib.h>

#include "gemm_common.cuh"

namespace pf {

constexpr int GEMM_THREADS = 192;

__host__ __device__ constexpr int gemm_tmem_cols(int block_n) {
  return block_n <= 32 ? 32 : block_n <= 64 ? 64 : block_n <= 128 ? 128 : block_n <= 256 ? 256 : 512;
}
__host__ __device__ constexpr int gemm_smem_bytes(int block_n, int stages) {
  return stages * gemm_stage_bytes(block_n) + 128 /*barriers*/ + 2 * block_n * 4 /*bias row + LayerNorm column sums*/;
}

// EPI_TMA: plain row map + 16-bit output. The output tile is staged in shared memory as BLOCK_N/32 sub-tiles of
// [128 rows][32 cols] (64-byte rows, SWIZZLE_64B) and written with TMA tile stores (fully coalesced, rows >= M
// clipped by the hardware); a 16-bit residual tile is TMA-prefetched into the same staging buffer at kernel start,
// so it arrives under the main loop instead of as per-thread scattered loads in the epilogue.
// PAIR: launched as clusters of 2 CTAs that issue ONE tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own
// 128 A rows and only HALF of the B (weight) tile, so the operand bytes pulled from L2 per FLOP drop by ~28 % — the big
// convolutions are L2->SM bandwidth bound at 128 x 160 tiles.
template <int BLOCK_N, int STAGES, bool BF16, bool EPI_TMA, bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const GemmKernelParams p) {
  constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * GEMM_BLOCK_K * 2;  // per CTA
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGING_BYTES = EPI_TMA ? GEMM_BLOCK_M * BLOCK_N * 2 : 0;
  constexpr int SUB_BYTES = GEMM_BLOCK_M * 64;  // one [128][32] 16-bit sub-tile
  constexpr int TMEM_COLS = gemm_tmem_cols(BLOCK_N);
  static_assert(!(PAIR && EPI_TMA), "the CTA-pair variant uses the direct epilogue");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA M=128 needs N%16==0, N<=256");
  static_assert(!EPI_TMA || BLOCK_N % 32 == 0, "staged epilogue works on 32-column sub-tiles");

  // 1024-byte alignment (128 B swizzle atoms) is requested on the declaration; verified once, never padded for
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* res_full_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(staging + STAGING_BYTES + 128);  // [BLOCK_N]
  float* s_cs = s_bias + BLOCK_N;                                            // [BLOCK_N] LayerNorm column sums

  pdl_launch_dependents();  // the next kernel of the stream may start its prologue while this one runs
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / BLOCK_N;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs of the pair)
  const int tile = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int n_tile = tile % n_tiles;  // n fastest: concurrent CTAs share the A tile through L2
  const int m_tile = tile / n_tiles;
  const int m0 = m_tile * (PAIR ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M) + int(rank) * GEMM_BLOCK_M;
  const int n0 = n_tile * BLOCK_N;
  // split-K: this CTA owns K-slabs [kb_begin, kb_end)
  const int split = (!PAIR && p.k_splits > 1) ? int(blockIdx.y) : 0;
  const int kb_begin = (p.k_splits > 1) ? (int)((long long)split * p.num_kb / p.k_splits) : 0;
  const int kb_end = (p.k_splits > 1) ? (int)((long long)(split + 1) * p.num_kb / p.k_splits) : p.num_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(res_full_bar, 1);
    if constexpr (EPI_TMA) tma_prefetch_desc(&tmC);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // barrier inits + TMEM allocation visible to the peer CTA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above touched only on-chip state; operands of the predecessor are visible from here on

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      if constexpr (EPI_TMA) {
        if (p.residual) {
          mbar_expect_tx(res_full_bar, STAGING_BYTES);
#pragma unroll
          for (int sub = 0; sub < BLOCK_N / 32; ++sub)
            tma_load_2d(staging + sub * SUB_BYTES, &tmR, res_full_bar, n0 + sub * 32, m0);
        }
      }
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int s = (kb - kb_begin) % STAGES;
        const uint32_t ph = ((kb - kb_begin) / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = kb / p.kb_per_tap;
        const int kk = kb - tap * p.kb_per_tap;
        uint8_t* sa = smem + s * STAGE_BYTES;
        if constexpr (PAIR) {
          // the leader arms its barrier with the bytes of BOTH CTAs; each CTA loads its A rows and its half of B
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
          tma_load_2d_2sm(sa, &tmA, &full_bar[s], kk * GEMM_BLOCK_K, m0 + p.tap_off[tap]);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, &full_bar[s], kb * GEMM_BLOCK_K, n0 + int(rank) * (BLOCK_N / 2));
        } else {
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[s], kk * GEMM_BLOCK_K, m0 + p.tap_off[tap]);
          tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * GEMM_BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if constexpr (PAIR) {
      if (lane == 0 && rank == 0) {
        constexpr uint32_t idesc2 = make_idesc_f16(BF16 ? 1 : 0, 2 * GEMM_BLOCK_M, BLOCK_N, 0, 0);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int s = kb % STAGES;
          mbar_wait(&full_bar[s], (kb / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
          const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_f16_2sm(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[s]);  // frees the slot in BOTH CTAs
        }
        umma_commit_2sm(tmem_full_bar);  // accumulators (128 rows in each CTA's TMEM) complete
      }
    } else if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BF16 ? 1 : 0, GEMM_BLOCK_M, BLOCK_N, 0, 0);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int s = (kb - kb_begin) % STAGES;
        const uint32_t ph = ((kb - kb_begin) / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
          // +32 B per UMMA_K step inside the 128 B swizzle row => +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ------------------------------ epilogue -----------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int m = m0 + row;
    bool valid = m < p.M;
    long long orow = m;
    int group = 0;
    if (p.map_mode == 1) {
      const int hw = p.Hm * p.Wm;
      const int img = m / hw;
      const int r = m - img * hw;
      const int i = r / p.Wm;
      const int j = r - i * p.Wm;
      valid = valid && i >= p.i0 && i < p.i0 + p.Hout && j >= p.j0 && j < p.j0 + p.Wout;
      orow = ((long long)img * p.Hout * p.osy + (i - p.i0) * p.osy + p.oa) * (p.Wout * p.osx) + (j - p.j0) * p.osx + p.ob;
      group = img;
    } else if (p.rowbias) {
      group = m / p.rows_per_group;
    }
    const uint32_t taddr_row = tmem_base + (uint32_t(q * 32) << 16);
    // bias row of this column tile -> shared memory while the main loop is still running
    if (p.bias) {
      for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) s_bias[i] = __ldg(p.bias + n0 + i);
    }
    float ln_a = 1.f, ln_b = 0.f;
    if (p.ln_stats) {
      for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) s_cs[i] = __ldg(p.ln_colsum + n0 + i);
      if (m < p.M) ln_row_coeffs(p, m, ln_a, ln_b);
    }
    named_bar_sync(1, 128);

    if (p.k_splits > 1) {
      // split-K partial: raw fp32 accumulators, M-space rows (the reduce kernel applies row map + epilogue)
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      float* wrow = p.ws + ((long long)split * p.M + m) * p.N + n0;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr_row + c, v);
        tmem_ld_wait();
        if (m < p.M) {
          float4* dst = reinterpret_cast<float4*>(wrow + c);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                 __uint_as_float(v[4 * e + 3]));
        }
      }
    } else if constexpr (EPI_TMA) {
      if (!valid) group = 0;  // rows past M are computed (and clipped by the TMA store): keep their table reads in bounds
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      if (p.residual) mbar_wait(res_full_bar, 0);
      const uint32_t sw = uint32_t((row >> 1) & 3);  // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
      // row statistics in the persistent kernel's order: even 16-column chunks -> slot 0, odd chunks -> slot 1
      float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr_row + c, v);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(v[e]);
        if (p.ln_stats) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = fmaf(o[e], ln_a, s_cs[c + e] * ln_b);
        }
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] += s_bias[c + e];
        }
        if (p.rowbias) {
          const float* rb = p.rowbias + (long long)group * p.rowbias_ld + n0 + c;
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] += __ldg(rb + e);
        }
        if (p.act == PF_ACT_SILU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = silu_f(o[e]);
        } else if (p.act == PF_ACT_GELU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = gelu_erf_f(o[e]);
        }
        uint8_t* srow = staging + (c >> 5) * SUB_BYTES + row * 64;
        const uint32_t q0 = uint32_t((c >> 4) & 1) * 2;  // first 16-byte chunk of this 16-column group in the 64 B row
        uint4* s0 = reinterpret_cast<uint4*>(srow + (((q0 + 0) ^ sw) << 4));
        uint4* s1 = reinterpret_cast<uint4*>(srow + (((q0 + 1) ^ sw) << 4));
        if (p.residual) {
          const uint4 r0 = *s0, r1 = *s1;
          float2 f;
          f = unpack2<BF16>(r0.x); o[0] += f.x; o[1] += f.y;
          f = unpack2<BF16>(r0.y); o[2] += f.x; o[3] += f.y;
          f = unpack2<BF16>(r0.z); o[4] += f.x; o[5] += f.y;
          f = unpack2<BF16>(r0.w); o[6] += f.x; o[7] += f.y;
          f = unpack2<BF16>(r1.x); o[8] += f.x; o[9] += f.y;
          f = unpack2<BF16>(r1.y); o[10] += f.x; o[11] += f.y;
          f = unpack2<BF16>(r1.z); o[12] += f.x; o[13] += f.y;
          f = unpack2<BF16>(r1.w); o[14] += f.x; o[15] += f.y;
        }
        if (p.row_stats) {
          const bool odd = (c >> 4) & 1;
          float ss = odd ? st_s1 : st_s0, qq = odd ? st_q1 : st_q0;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            ss += o[e];
            qq = fmaf(o[e], o[e], qq);
          }
          if (odd) {
            st_s1 = ss;
            st_q1 = qq;
          } else {
            st_s0 = ss;
            st_q0 = qq;
          }
        }
        *s0 = make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]), pack2<BF16>(o[4], o[5]),
                         pack2<BF16>(o[6], o[7]));
        *s1 = make_uint4(pack2<BF16>(o[8], o[9]), pack2<BF16>(o[10], o[11]), pack2<BF16>(o[12], o[13]),
                         pack2<BF16>(o[14], o[15]));
      }
      if (p.row_stats && m < p.M) {
        float2* st = reinterpret_cast<float2*>(p.row_stats) + (long long)m * p.stat_slots + n_tile * 2;
        st[0] = make_float2(st_s0, st_q0);
        st[1] = make_float2(st_s1, st_q1);
      }
      fence_proxy_async_smem();           // generic-proxy writes -> visible to the TMA store
      named_bar_sync(1, 128);             // the four epilogue warps only
      if (warp == 2 && lane == 0) {
#pragma unroll
        for (int sub = 0; sub < BLOCK_N / 32; ++sub) tma_store_2d(&tmC, staging + sub * SUB_BYTES, n0 + sub * 32, m0);
        tma_store_commit();
        tma_store_wait_read();
      }
    } else if (p.act == PF_ACT_GEGLU) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      constexpr int HALF = BLOCK_N / 2;
      const int on0 = n_tile * HALF;
      if constexpr (HALF % 16 == 0) {
#pragma unroll 1
        for (int c = 0; c < HALF; c += 16) {
          uint32_t va[16], vg[16];
          tmem_ld16(taddr_row + c, va);
          tmem_ld16(taddr_row + HALF + c, vg);
          tmem_ld_wait();
          if (valid) {
            float o[16], ba[16], bg[16];
            if (p.bias) {  // HALF and c are multiples of 16: 128-bit shared-memory loads
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                *reinterpret_cast<float4*>(ba + 4 * e) = *reinterpret_cast<const float4*>(s_bias + c + 4 * e);
                *reinterpret_cast<float4*>(bg + 4 * e) = *reinterpret_cast<const float4*>(s_bias + HALF + c + 4 * e);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) ba[e] = bg[e] = 0.f;
            }
            if (p.ln_stats) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                va[e] = __float_as_uint(fmaf(__uint_as_float(va[e]), ln_a, s_cs[c + e] * ln_b));
                vg[e] = __float_as_uint(fmaf(__uint_as_float(vg[e]), ln_a, s_cs[HALF + c + e] * ln_b));
              }
            }
#pragma unroll
            for (int e = 0; e < 16; e += 2)  // two columns per FFMA2 / FMUL2 / FADD2 (bit-identical to gelu_erf_f)
              geglu_pair(__uint_as_float(va[e]), __uint_as_float(va[e + 1]), __uint_as_float(vg[e]),
                         __uint_as_float(vg[e + 1]), ba[e], ba[e + 1], bg[e], bg[e + 1], o[e], o[e + 1]);
            if (p.out_f32) {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.out_ld + on0 + c);
#pragma unroll
              for (int e = 0; e < 4; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + orow * p.out_ld + on0 + c);
              dst[0] = make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]), pack2<BF16>(o[4], o[5]),
                                  pack2<BF16>(o[6], o[7]));
              dst[1] = make_uint4(pack2<BF16>(o[8], o[9]), pack2<BF16>(o[10], o[11]), pack2<BF16>(o[12], o[13]),
                                  pack2<BF16>(o[14], o[15]));
            }
          }
        }
      }
    } else {
      // Direct path (convolutions: halo-dropping row map; fp32 outputs). Global operands of the epilogue are fetched
      // ahead of use — the 16-bit residual through a 4-deep register ring started BEFORE the accumulator wait, the
      // per-image row bias one chunk ahead — so their latency hides under the main loop / the previous chunk.
      constexpr int NCH = BLOCK_N / 16;
      const bool res16 = p.residual != nullptr && !p.res_f32 && valid;
      const uint4* rsrc = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + orow * p.res_ld + n0);
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      // residual: chunks c and c+1 in flight (rolled loop, explicit double buffer)
      uint4 ra0 = z4, ra1 = z4, rb0 = z4, rb1 = z4;
      if (res16) {
        ra0 = __ldg(rsrc);
        ra1 = __ldg(rsrc + 1);
        if (NCH > 1) {
          rb0 = __ldg(rsrc + 2);
          rb1 = __ldg(rsrc + 3);
        }
      }
      const float* rb_base = p.rowbias ? p.rowbias + (long long)(valid ? group : 0) * p.rowbias_ld + n0 : nullptr;
      float4 rbn[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) rbn[e] = rb_base ? __ldg(reinterpret_cast<const float4*>(rb_base) + e) : make_float4(0, 0, 0, 0);
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = ci * 16;
        uint32_t v[16];
        tmem_ld16(taddr_row + c, v);
        float4 rbc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) rbc[e] = rbn[e];
        if (rb_base && ci + 1 < NCH) {
#pragma unroll
          for (int e = 0; e < 4; ++e) rbn[e] = __ldg(reinterpret_cast<const float4*>(rb_base + c + 16) + e);
        }
        const uint4 r0 = ra0, r1 = ra1;
        ra0 = rb0;
        ra1 = rb1;
        if (res16 && ci + 2 < NCH) {
          rb0 = __ldg(rsrc + 2 * (ci + 2));
          rb1 = __ldg(rsrc + 2 * (ci + 2) + 1);
        }
        tmem_ld_wait();
        if (valid) {
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(v[e]);
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] += s_bias[c + e];
          }
          if (rb_base) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              o[4 * e] += rbc[e].x;
              o[4 * e + 1] += rbc[e].y;
              o[4 * e + 2] += rbc[e].z;
              o[4 * e + 3] += rbc[e].w;
            }
          }
          if (p.act == PF_ACT_SILU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = silu_f(o[e]);
          } else if (p.act == PF_ACT_GELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = gelu_erf_f(o[e]);
          }
          if (p.residual) {
            if (p.res_f32) {
              const float4* r4 =
                  reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + orow * p.res_ld + n0 + c);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float4 t = r4[e];
                o[4 * e] += t.x;
                o[4 * e + 1] += t.y;
                o[4 * e + 2] += t.z;
                o[4 * e + 3] += t.w;
              }
            } else {
              float2 f;
              f = unpack2<BF16>(r0.x); o[0] += f.x; o[1] += f.y;
              f = unpack2<BF16>(r0.y); o[2] += f.x; o[3] += f.y;
              f = unpack2<BF16>(r0.z); o[4] += f.x; o[5] += f.y;
              f = unpack2<BF16>(r0.w); o[6] += f.x; o[7] += f.y;
              f = unpack2<BF16>(r1.x); o[8] += f.x; o[9] += f.y;
              f = unpack2<BF16>(r1.y); o[10] += f.x; o[11] += f.y;
              f = unpack2<BF16>(r1.z); o[12] += f.x; o[13] += f.y;
              f = unpack2<BF16>(r1.w); o[14] += f.x; o[15] += f.y;
            }
          }
          if (p.out_f32) {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.out_ld + n0 + c);
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + orow * p.out_ld + n0 + c);
            dst[0] = make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]), pack2<BF16>(o[4], o[5]),
                                pack2<BF16>(o[6], o[7]));
            dst[1] = make_uint4(pack2<BF16>(o[8], o[9]), pack2<BF16>(o[10], o[11]), pack2<BF16>(o[12], o[13]),
                                pack2<BF16>(o[14], o[15]));
          }
        }
      }
    }
    tc_fence_before();
  }

  if constexpr (PAIR) cluster_sync_all();  // the peer may still be reading this CTA's smem / TMEM through the pair MMA
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// split-K reduce: fixed-order sum of the k_splits partials, then the same epilogue as the in-kernel one
// (bias, per-image row bias, activation, residual, halo-dropping row map). thread <-> (M-space row, 8 columns)
template <bool BF16>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmKernelParams p) {
  const int vecs = p.N / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * vecs) return;
  const int m = int(idx / vecs), c = int(idx % vecs) * 8;
  bool valid = true;
  long long orow = m;
  int group = 0;
  if (p.map_mode == 1) {
    const int hw = p.Hm * p.Wm;
    const int img = m / hw;
    const int r = m - img * hw;
    const int i = r / p.Wm;
    const int j = r - i * p.Wm;
    valid = i >= p.i0 && i < p.i0 + p.Hout && j >= p.j0 && j < p.j0 + p.Wout;
    orow = ((long long)img * p.Hout * p.osy + (i - p.i0) * p.osy + p.oa) * (p.Wout * p.osx) + (j - p.j0) * p.osx + p.ob;
    group = img;
  } else if (p.rowbias) {
    group = m / p.rows_per_group;
  }
  if (!valid) return;
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  for (int sp = 0; sp < p.k_splits; ++sp) {
    const float4* w4 = reinterpret_cast<const float4*>(p.ws + ((long long)sp * p.M + m) * p.N + c);
    const float4 a = __ldcs(w4), b = __ldcs(w4 + 1);
    o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
    o[4] += b.x; o[5] += b.y; o[6] += b.z; o[7] += b.w;
  }
  if (p.bias) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] += __ldg(p.bias + c + e);
  }
  if (p.rowbias) {
    const float* rb = p.rowbias + (long long)group * p.rowbias_ld + c;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] += __ldg(rb + e);
  }
  if (p.act == PF_ACT_SILU) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = silu_f(o[e]);
  } else if (p.act == PF_ACT_GELU) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = gelu_erf_f(o[e]);
  }
  if (p.residual) {
    if (p.res_f32) {
      const float4* r4 = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + orow * p.res_ld + c);
      const float4 a = r4[0], b = r4[1];
      o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
      o[4] += b.x; o[5] += b.y; o[6] += b.z; o[7] += b.w;
    } else {
      const uint4 t = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + orow * p.res_ld + c);
      float2 f;
      f = unpack2<BF16>(t.x); o[0] += f.x; o[1] += f.y;
      f = unpack2<BF16>(t.y); o[2] += f.x; o[3] += f.y;
      f = unpack2<BF16>(t.z); o[4] += f.x; o[5] += f.y;
      f = unpack2<BF16>(t.w); o[6] += f.x; o[7] += f.y;
    }
  }
  if (p.out_f32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.out_ld + c);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    *reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + orow * p.out_ld + c) =
        make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]), pack2<BF16>(o[4], o[5]), pack2<BF16>(o[6], o[7]));
  }
}

template <int BLOCK_N, int STAGES, bool EPI_TMA>
static int launch_gemm(const pf_gemm_args* a, const GemmKernelParams& kp, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmC, tmR;
  {
    uint64_t dims[2] = {(uint64_t)a->Kc, (uint64_t)a->a_rows};
    uint64_t str[1] = {(uint64_t)a->a_ld * 2};
    uint32_t box[2] = {GEMM_BLOCK_K, GEMM_BLOCK_M};
    int rc = make_tmap(&tmA, a->dtype, 2, a->A, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->Kc * a->num_taps, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->b_ld * 2};
    uint32_t box[2] = {GEMM_BLOCK_K, (uint32_t)BLOCK_N};
    int rc = make_tmap(&tmB, a->dtype, 2, a->B, dims, str, box, 128);
    if (rc) return rc;
  }
  tmC = tmA;
  tmR = tmA;
  if constexpr (EPI_TMA) {
    uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
    uint32_t box[2] = {32, GEMM_BLOCK_M};
    uint64_t str[1] = {(uint64_t)a->out_ld * 2};
    int rc = make_tmap(&tmC, a->dtype, 2, a->out, dims, str, box, 64);
    if (rc) return rc;
    if (a->residual) {
      uint64_t rstr[1] = {(uint64_t)a->res_ld * 2};
      rc = make_tmap(&tmR, a->dtype, 2, a->residual, dims, rstr, box, 64);
      if (rc) return rc;
    }
  }
  constexpr int SMEM = gemm_smem_bytes(BLOCK_N, STAGES) + (EPI_TMA ? GEMM_BLOCK_M * BLOCK_N * 2 : 0);
  const int m_tiles = (a->M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const dim3 grid(m_tiles * (a->N / BLOCK_N), kp.k_splits > 1 ? kp.k_splits : 1);
  if (a->dtype == PF_BF16) {
    auto kern = gemm_taps_kernel<BLOCK_N, STAGES, true, EPI_TMA, false>;
    static bool attr_set = false;
    if (!attr_set) {
      int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM),
                          "cudaFuncSetAttribute(gemm)");
      if (rc) return rc;
      attr_set = true;
    }
    if (int rc = check_cuda(launch_pdl(kern, grid, dim3(GEMM_THREADS), SMEM, st, tmA, tmB, tmC, tmR, kp), "launch(gemm)")) return rc;
  } else {
    auto kern = gemm_taps_kernel<BLOCK_N, STAGES, false, EPI_TMA, false>;
    static bool attr_set = false;
    if (!attr_set) {
      int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM),
                          "cudaFuncSetAttribute(gemm)");
      if (rc) return rc;
      attr_set = true;
    }
    if (int rc = check_cuda(launch_pdl(kern, grid, dim3(GEMM_THREADS), SMEM, st, tmA, tmB, tmC, tmR, kp), "launch(gemm)")) return rc;
  }
  PF_CHECK_LAUNCH("gemm_taps_kernel");
  return PF_OK;
}

template <int BLOCK_N, int STAGES>
static int launch_gemm_pair(const pf_gemm_args* a, const GemmKernelParams& kp, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)a->Kc, (uint64_t)a->a_rows};
    uint64_t str[1] = {(uint64_t)a->a_ld * 2};
    uint32_t box[2] = {GEMM_BLOCK_K, GEMM_BLOCK_M};
    if ((rc = make_tmap(&tmA, a->dtype, 2, a->A, dims, str, box, 128))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->Kc * a->num_taps, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->b_ld * 2};
    uint32_t box[2] = {GEMM_BLOCK_K, (uint32_t)BLOCK_N / 2};  // each CTA of the pair stages half of the weight tile
    if ((rc = make_tmap(&tmB, a->dtype, 2, a->B, dims, str, box, 128))) return rc;
  }
  constexpr int SMEM = STAGES * (GEMM_BLOCK_M * GEMM_BLOCK_K * 2 + (BLOCK_N / 2) * GEMM_BLOCK_K * 2) + 128 + 2 * BLOCK_N * 4;
  const int m_pairs = (a->M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * m_pairs * (a->N / BLOCK_N));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (a->dtype == PF_BF16) {
    auto kern = gemm_taps_kernel<BLOCK_N, STAGES, true, false, true>;
    static bool attr_set = false;
    if (!attr_set) {
      if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM), "gemm pair attr"))) return rc;
      attr_set = true;
    }
    if ((rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA, tmA, kp), "cudaLaunchKernelEx(gemm pair)"))) return rc;
  } else {
    auto kern = gemm_taps_kernel<BLOCK_N, STAGES, false, false, true>;
    static bool attr_set = false;
    if (!attr_set) {
      if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM), "gemm pair attr"))) return rc;
      attr_set = true;
    }
    if ((rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA, tmA, kp), "cudaLaunchKernelEx(gemm pair)"))) return rc;
  }
  PF_CHECK_LAUNCH("gemm_taps_kernel(pair)");
  return PF_OK;
}

}  // namespace pf

namespace pf {
// tile width pf_gemm_taps runs with: the caller's / tuner's request, else the heuristic; the staged (TMA-store) epilogue
// that carries the fused-LayerNorm statistics has no 256-wide variant
static int resolve_block_n(const pf_gemm_args* a) {
  int bn = (a->block_n & 0xffff) ? (a->block_n & 0xffff) : pf_gemm_pick_block_n(a->N, a->act);
  // a statistics PRODUCER always uses the width the heuristic derives from N: the slot partition of the row sums (and so
  // their fp32 rounding) must not depend on M-specific tuning, or a sharded rank would not reproduce the full batch
  if (a->row_stats_out) bn = pf_gemm_pick_block_n(a->N, a->act);
  if ((a->row_stats_out || a->ln_stats) && a->act != PF_ACT_GEGLU && bn == 256) bn = 128;
  return bn;
}
}  // namespace pf

extern "C" int pf_gemm_row_stats_slots(const pf_gemm_args* a) {
  if (!a || a->N <= 0) return 0;
  pf_gemm_args producer = *a;  // the question is about a PRODUCER, whether or not the caller has set row_stats_out yet
  static float dummy;
  producer.row_stats_out = &dummy;
  const int bn = pf::resolve_block_n(&producer);
  return bn > 0 && a->N % bn == 0 ? 2 * (a->N / bn) : 0;
}

extern "C" int pf_gemm_pick_block_n(int N, int act) {
  // GEGLU tiles are epilogue-bound (one erf-GELU per output, K as short as 5 slabs): the 256-wide tile gives each
  // epilogue warp-group 4+4 balanced 16-column chunks (160: 3+2) and halves the per-tile fixed cost — measured 18-23 %
  // faster than 160 on every FF1 shape of the step (scripts/geglu_sweep.sh)
  if (act == PF_ACT_GEGLU && N % 256 == 0) return 256;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

extern "C" int pf_gemm_splitk_plan(const pf_gemm_args* a) {
  if (!a || a->act == PF_ACT_GEGLU || a->M <= 0 || a->N <= 0 || a->Kc <= 0) return 1;
  const int bn = pf_gemm_pick_block_n(a->N, a->act);
  if (!bn) return 1;
  const long long tiles = (long long)((a->M + 127) / 128) * (a->N / bn);
  const int num_kb = a->Kc / 64 * a->num_taps;
  // Only the skinny deep-K problems: a sharded rank's 8x8 / 16x16-level convolutions have <= 60 output tiles and 90-360
  // K-slabs, i.e. a handful of SMs each streaming megabytes of weights at the per-SM L2 rate (60-120 us per launch).
  // Shapes that already cover half the machine, or short K, lose more to the partial-sum traffic than they gain.
  if (tiles > 74 || num_kb < 64) return 1;
  int s = (int)((148 + tiles - 1) / tiles);           // aim for ~1 CTA per SM
  const int max_by_k = num_kb / 16;                   // >= 16 K-slabs per split
  if (s > max_by_k) s = max_by_k;
  if (s > 16) s = 16;
  return s < 2 ? 1 : s;
}

extern "C" int pf_gemm_taps(const pf_gemm_args* a, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(a != nullptr, "pf_gemm_taps: null args");
  PF_CHECK_ARG(a->dtype == PF_BF16 || a->dtype == PF_F16, "pf_gemm_taps: dtype must be PF_F16 or PF_BF16");
  PF_CHECK_ARG(a->A && a->B && a->out, "pf_gemm_taps: null operand");
  PF_CHECK_ARG(a->M > 0 && a->N > 0 && a->Kc > 0, "pf_gemm_taps: empty problem M=%d N=%d Kc=%d", a->M, a->N, a->Kc);
  PF_CHECK_ARG(a->Kc % GEMM_BLOCK_K == 0, "pf_gemm_taps: Kc=%d must be a multiple of 64", a->Kc);
  PF_CHECK_ARG(a->num_taps >= 1 && a->num_taps <= PF_MAX_TAPS, "pf_gemm_taps: num_taps=%d out of range", a->num_taps);
  PF_CHECK_ARG(a->a_ld % 8 == 0 && a->b_ld % 8 == 0 && a->a_ld >= a->Kc && a->b_ld >= a->Kc * a->num_taps,
               "pf_gemm_taps: bad leading dims a_ld=%d b_ld=%d", a->a_ld, a->b_ld);
  PF_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
               "pf_gemm_taps: operands must be 16-byte aligned");
  PF_CHECK_ARG(a->out_dtype == PF_F32 || a->out_dtype == a->dtype, "pf_gemm_taps: out_dtype must be f32 or dtype");
  PF_CHECK_ARG(!a->residual || a->res_dtype == PF_F32 || a->res_dtype == a->dtype,
               "pf_gemm_taps: res_dtype must be f32 or dtype");
  PF_CHECK_ARG(a->act >= PF_ACT_NONE && a->act <= PF_ACT_GEGLU, "pf_gemm_taps: unknown act %d", a->act);
  // block_n: low 16 bits = tile width (0 = auto); bits 16.. = schedule override (0 auto, 1 one-tile-per-CTA,
  // 2 persistent, 3 CTA pair) — written by scripts/tune_gemm.py into gemm_tuning.json, never needed by callers
  const int sched_req = a->block_n >> 16;
  int bn = pf::resolve_block_n(a);
  PF_CHECK_ARG(bn == 64 || bn == 128 || bn == 160 || bn == 256, "pf_gemm_taps: unsupported block_n %d (N=%d)", bn, a->N);
  PF_CHECK_ARG(a->N % bn == 0, "pf_gemm_taps: N=%d not a multiple of block_n=%d", a->N, bn);
  const int n_out = a->act == PF_ACT_GEGLU ? a->N / 2 : a->N;
  PF_CHECK_ARG(a->out_ld % 8 == 0 && a->out_ld >= n_out, "pf_gemm_taps: bad out_ld %d", a->out_ld);
  PF_CHECK_ARG(!a->residual || (a->res_ld % 8 == 0 && a->res_ld >= n_out), "pf_gemm_taps: bad res_ld %d", a->res_ld);
  PF_CHECK_ARG(!(a->act == PF_ACT_GEGLU && (a->residual || a->rowbias)),
               "pf_gemm_taps: GEGLU epilogue takes no residual/rowbias");
  if (a->map_mode == 1) {
    PF_CHECK_ARG(a->Hm > 0 && a->Wm > 0 && a->Hout > 0 && a->Wout > 0 && a->M % (a->Hm * a->Wm) == 0,
                 "pf_gemm_taps: bad image map Hm=%d Wm=%d M=%d", a->Hm, a->Wm, a->M);
    PF_CHECK_ARG(a->out_sy >= 0 && a->out_sx >= 0 && a->out_a >= 0 && a->out_b >= 0 &&
                     a->out_a < (a->out_sy > 0 ? a->out_sy : 1) && a->out_b < (a->out_sx > 0 ? a->out_sx : 1),
                 "pf_gemm_taps: bad output scatter (%d,%d) phase (%d,%d)", a->out_sy, a->out_sx, a->out_a, a->out_b);
    PF_CHECK_ARG((a->out_sy <= 1 && a->out_sx <= 1) || !a->residual, "pf_gemm_taps: the scattered output map takes no residual");
  } else {
    PF_CHECK_ARG(a->map_mode == 0, "pf_gemm_taps: unknown map_mode %d", a->map_mode);
    PF_CHECK_ARG(!a->rowbias || a->rows_per_group > 0, "pf_gemm_taps: rowbias needs rows_per_group");
  }

  GemmKernelParams kp;
  kp.M = a->M;
  kp.N = a->N;
  kp.kb_per_tap = a->Kc / GEMM_BLOCK_K;
  kp.num_kb = kp.kb_per_tap * a->num_taps;
  for (int t = 0; t < PF_MAX_TAPS; ++t) kp.tap_off[t] = t < a->num_taps ? a->tap_off[t] : 0;
  kp.out = a->out;
  kp.out_ld = a->out_ld;
  kp.out_f32 = a->out_dtype == PF_F32;
  kp.bias = a->bias;
  kp.rowbias = a->rowbias;
  kp.rowbias_ld = a->rowbias_ld;
  kp.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  kp.residual = a->residual;
  kp.res_ld = a->res_ld;
  kp.res_f32 = a->res_dtype == PF_F32;
  kp.act = a->act;
  kp.map_mode = a->map_mode;
  kp.Hm = a->Hm;
  kp.Wm = a->Wm;
  kp.i0 = a->i0;
  kp.j0 = a->j0;
  kp.Hout = a->Hout;
  kp.Wout = a->Wout;
  kp.osy = a->out_sy > 0 ? a->out_sy : 1;
  kp.osx = a->out_sx > 0 ? a->out_sx : 1;
  kp.oa = a->out_a;
  kp.ob = a->out_b;
  kp.k_splits = a->k_splits > 1 ? a->k_splits : 1;
  kp.ws = a->splitk_ws;
  kp.row_stats = a->row_stats_out;
  kp.stat_slots = 2 * (a->N / bn);
  kp.ln_stats = a->ln_stats;
  kp.ln_slots = a->ln_slots;
  kp.ln_colsum = a->ln_colsum;
  kp.ln_inv_k = 1.0f / float((long long)a->Kc * a->num_taps);
  kp.ln_eps = a->ln_eps;
  PF_CHECK_ARG(kp.k_splits == 1 || (a->splitk_ws && a->act != PF_ACT_GEGLU && kp.k_splits <= kp.num_kb),
               "pf_gemm_taps: split-K needs a workspace, no GEGLU and k_splits <= K-slabs");
  const bool fused_ln = a->row_stats_out || a->ln_stats;
  if (fused_ln) {
    PF_CHECK_ARG(kp.k_splits == 1, "pf_gemm_taps: fused LayerNorm does not combine with split-K");
    PF_CHECK_ARG(!a->ln_stats || (a->ln_colsum && a->ln_slots > 0 && a->ln_slots % 2 == 0 && a->ln_eps > 0.f &&
                                  (reinterpret_cast<uintptr_t>(a->ln_stats) & 15) == 0),
                 "pf_gemm_taps: ln_stats needs ln_colsum, ln_slots and ln_eps");
    const bool plain16 = a->map_mode == 0 && a->out_dtype == a->dtype && (!a->residual || a->res_dtype == a->dtype) &&
                         (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0;
    PF_CHECK_ARG(a->act == PF_ACT_GEGLU ? !a->row_stats_out : plain16,
                 "pf_gemm_taps: fused LayerNorm needs the plain row map with 16-bit output (consumer: or GEGLU)");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kp.k_splits > 1) {
    int rc;
    switch (bn) {
      case 64: rc = launch_gemm<64, 4, false>(a, kp, st); break;
      case 128: rc = launch_gemm<128, 3, false>(a, kp, st); break;
      case 160: rc = launch_gemm<160, 3, false>(a, kp, st); break;
      default: rc = launch_gemm<256, 2, false>(a, kp, st); break;
    }
    if (rc) return rc;
    const long long total = (long long)a->M * (a->N / 8);
    if (a->dtype == PF_BF16) splitk_reduce_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(kp);
    else splitk_reduce_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(kp);
    PF_CHECK_LAUNCH("splitk_reduce_kernel");
    return PF_OK;
  }
  // staged TMA-store epilogue: plain row map, 16-bit output, 16-bit (or no) residual, no GEGLU
  const bool epi_tma = a->map_mode == 0 && a->out_dtype == a->dtype && a->act != PF_ACT_GEGLU &&
                       (!a->residual || a->res_dtype == a->dtype) && bn != 256 &&
                       (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0;
  // default: persistent warp-specialised kernel (gemm_persist.cu); PF_GEMM_LEGACY=1 keeps the one-tile-per-CTA kernel
  // below reachable for A/B debugging only
  // Scheduling choice (measured on B200, scripts/gemm_micro.py): linear layers (short K, staged TMA-store epilogue)
  // run best as ONE persistent CTA per SM with double-buffered accumulators; convolutions (long K, direct stores) run
  // best as two co-resident one-tile CTAs per SM, whose two TMA producers keep more K-slabs in flight.
  // PF_GEMM_SCHED=persistent|tile forces one of them (debugging / A-B timing only).
  static const char* force = getenv("PF_GEMM_SCHED");
  const int m_tiles_ = (a->M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  bool persistent = (epi_tma || a->act == PF_ACT_GEGLU) && (long long)m_tiles_ * (a->N / bn) >= 2 * 148;
  if (force) persistent = force[0] == 'p';
  if (sched_req) persistent = sched_req == 2;
  if (persistent && (bn != 256 || !epi_tma)) return launch_gemm_persistent(a, kp, bn, epi_tma && bn != 256, st);
  // CTA pairs (cta_group::2, M = 256) for the long-K direct-epilogue GEMMs = the 3x3 convolutions
  static const char* pair_env = getenv("PF_GEMM_PAIR");
  bool pair = !epi_tma && a->act != PF_ACT_GEGLU && a->num_taps * a->Kc >= 1024 && (bn == 160 || bn == 256) &&
              (long long)m_tiles_ * (a->N / bn) >= 148;
  if (pair_env) pair = pair && pair_env[0] != '0';
  if (sched_req) pair = sched_req == 3 && !epi_tma && a->act != PF_ACT_GEGLU && (bn == 160 || bn == 256);
  if (pair) {
    if (bn == 160) return launch_gemm_pair<160, 4>(a, kp, st);
    return launch_gemm_pair<256, 3>(a, kp, st);
  }
  if (epi_tma) {
    switch (bn) {
      case 64: return launch_gemm<64, 4, true>(a, kp, st);
      case 128: return launch_gemm<128, 2, true>(a, kp, st);
      case 160: return launch_gemm<160, 2, true>(a, kp, st);
    }
  }
  switch (bn) {
    case 64: return launch_gemm<64, 4, false>(a, kp, st);
    case 128: return launch_gemm<128, 3, false>(a, kp, st);
    case 160: return launch_gemm<160, 3, false>(a, kp, st);
    case 256: return launch_gemm<256, 2, false>(a, kp, st);
  }
  return PF_ERR_UNSUPPORTED;
}
