// Persistent warp-specialised tap-GEMM (tcgen05 / TMEM / TMA), one CTA per SM looping over output tiles.
//
//   warp 0      TMA producer: A/B K-slabs into a STAGES-deep smem ring that runs AHEAD across tile boundaries,
//               plus (linear layers) the 16-bit residual tile of the next output tile into a staging buffer
//   warp 1      TMEM owner + single-thread tcgen05.mma issuer; TWO accumulator stages in TMEM, so the MMAs of tile
//               i+1 overlap the epilogue of tile i
//   warps 2..9  epilogue (2 warps per TMEM lane quarter, alternating 16-column chunks): TMEM -> registers -> bias /
//               per-image temb row / SiLU / GELU / GEGLU / residual -> either (EPI_TMA) a swizzled smem staging tile
//               written back with TMA tile stores, or (convolutions) direct stores through the halo-dropping row map.
//
// Same contract as pf_gemm_taps in include/panfusion_b200.h; replaces the one-tile-per-CTA kernel for every
// nn.Linear and 1x1 / 3x3 convolution of the UNet walk (models/pano/MVGenModel.py:85-295).
#include "gemm_common.cuh"

namespace pf {

constexpr int PG_THREADS = 320;
constexpr int PG_EPI_THREADS = 256;

__host__ __device__ constexpr int pg_acc_stride(int block_n) { return block_n <= 64 ? 64 : block_n <= 128 ? 128 : 256; }
__host__ __device__ constexpr int pg_smem_bytes(int block_n, int stages, bool epi_tma) {
  return stages * gemm_stage_bytes(block_n) + (epi_tma ? 2 * GEMM_BLOCK_M * block_n * 2 : 0) + 256 /*barriers*/ +
         4 * block_n * 4 /*bias rows + LayerNorm column sums*/;
}

template <int BLOCK_N, int STAGES, bool BF16, bool EPI_TMA>
__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const GemmKernelParams p) {
  constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGING_BYTES = EPI_TMA ? GEMM_BLOCK_M * BLOCK_N * 2 : 0;  // one of two buffers
  constexpr int SUB_BYTES = GEMM_BLOCK_M * 64;                              // [128][32] 16-bit sub-tile, SWIZZLE_64B
  constexpr int ACC_STRIDE = pg_acc_stride(BLOCK_N);
  constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  constexpr int NCH = BLOCK_N / 16;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "tile width");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* res_full = tmem_empty + 2;        // [2]
  uint64_t* stg_empty = res_full + 2;         // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(stg_empty + 2);
  float* s_bias = reinterpret_cast<float*>(staging + 2 * STAGING_BYTES + 256);  // [2][BLOCK_N]
  float* s_cs = s_bias + 2 * BLOCK_N;                                            // [2][BLOCK_N] LayerNorm column sums

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.N / BLOCK_N;
  const int m_tiles = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI_TMA) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PG_EPI_THREADS);
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_empty[i], 1);
    }
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
  pdl_wait();  // on-chip prologue done; the predecessor's results are visible from here on

  if (warp == 0) {
    // ---------------------------------- TMA producer ----------------------------------
    if (lane == 0) {
      int g = 0;  // global K-slab counter: the ring never drains between tiles
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int m0 = (tile / n_tiles) * GEMM_BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
        if constexpr (EPI_TMA) {
          if (p.residual) {
            const int buf = it & 1;
            mbar_wait(&stg_empty[buf], ((it >> 1) & 1) ^ 1);  // staging[buf] released by the store of tile it-2
            mbar_expect_tx(&res_full[buf], STAGING_BYTES);
#pragma unroll
            for (int sub = 0; sub < BLOCK_N / 32; ++sub)
              tma_load_2d(staging + buf * STAGING_BYTES + sub * SUB_BYTES, &tmR, &res_full[buf], n0 + sub * 32, m0);
          }
        }
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          const int tap = kb / p.kb_per_tap;
          const int kk = kb - tap * p.kb_per_tap;
          uint8_t* sa = smem + s * STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full_bar[s], kk * GEMM_BLOCK_K, m0 + p.tap_off[tap]);
          tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * GEMM_BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------- MMA issuer -------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BF16 ? 1 : 0, GEMM_BLOCK_M, BLOCK_N, 0, 0);
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);  // epilogue drained this accumulator (tile it-2)
        tc_fence_after();
        const uint32_t td = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&full_bar[s], (g / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
          const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_f16(td, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------------------------- epilogue (8 warps) ------------------------------
    const int et = threadIdx.x - 64;        // 0..255
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;       // which of the two warps of the quarter: even / odd 16-column chunks
    const int row = q * 32 + lane;
    const uint32_t sw = uint32_t((row >> 1) & 3);
    // bias row of the first tile
    if (p.bias && blockIdx.x < num_tiles) {
      const int n0 = (blockIdx.x % n_tiles) * BLOCK_N;
      if (et < BLOCK_N) s_bias[et] = __ldg(p.bias + n0 + et);
    }
    if (p.ln_stats && blockIdx.x < num_tiles) {
      const int n0 = (blockIdx.x % n_tiles) * BLOCK_N;
      if (et < BLOCK_N) s_cs[et] = __ldg(p.ln_colsum + n0 + et);
    }
    named_bar_sync(1, PG_EPI_THREADS);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % n_tiles;
      const int m0 = (tile / n_tiles) * GEMM_BLOCK_M, n0 = n_tile * BLOCK_N;
      const int acc = it & 1, buf = it & 1;
      const float* bias_s = s_bias + (it & 1) * BLOCK_N;
      // prefetch the next tile's bias entry (published at the end of this iteration)
      float bias_next = 0.f;
      const int next_tile = tile + gridDim.x;
      if (p.bias && next_tile < num_tiles && et < BLOCK_N)
        bias_next = __ldg(p.bias + (next_tile % n_tiles) * BLOCK_N + et);
      const float* cs_s = s_cs + (it & 1) * BLOCK_N;
      float cs_next = 0.f;
      if (p.ln_stats && next_tile < num_tiles && et < BLOCK_N)
        cs_next = __ldg(p.ln_colsum + (next_tile % n_tiles) * BLOCK_N + et);

      const int m = m0 + row;
      bool valid = m < p.M;
      long long orow = m;
      int group = 0;
      float ln_a = 1.f, ln_b = 0.f;
      if (p.ln_stats && valid) ln_row_coeffs(p, m, ln_a, ln_b);
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
      if (!valid) group = 0;
      const uint32_t taddr_row = tmem_base + acc * ACC_STRIDE + (uint32_t(q * 32) << 16);
      const float* rb_base = p.rowbias ? p.rowbias + (long long)group * p.rowbias_ld + n0 : nullptr;

      if constexpr (EPI_TMA) {
        uint8_t* stg = staging + buf * STAGING_BYTES;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        tc_fence_after();
        // every earlier TMA store has finished reading shared memory: staging[buf] (store of tile it-2) may be
        // rewritten, and staging[buf^1] (store of tile it-1) is handed to the producer for the next residual tile
        if (et == 0) {
          tma_store_wait_read();
          if (it >= 1) mbar_arrive(&stg_empty[buf ^ 1]);
        }
        named_bar_sync(2, PG_EPI_THREADS);
        if (p.residual) mbar_wait(&res_full[buf], (it >> 1) & 1);
        float st_s = 0.f, st_q = 0.f;
#pragma unroll 1
        for (int ci = half; ci < NCH; ci += 2) {
          const int c = ci * 16;
          uint32_t v[16];
          tmem_ld16(taddr_row + c, v);
          tmem_ld_wait();
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(v[e]);
          if (p.ln_stats) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = fmaf(o[e], ln_a, cs_s[c + e] * ln_b);
          }
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] += bias_s[c + e];
          }
          if (rb_base) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] += __ldg(rb_base + c + e);
          }
          if (p.act == PF_ACT_SILU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = silu_f(o[e]);
          } else if (p.act == PF_ACT_GELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = gelu_erf_f(o[e]);
          }
          uint8_t* srow = stg + (c >> 5) * SUB_BYTES + row * 64;
          const uint32_t q0 = uint32_t((c >> 4) & 1) * 2;
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
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              st_s += o[e];
              st_q = fmaf(o[e], o[e], st_q);
            }
          }
          *s0 = make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]), pack2<BF16>(o[4], o[5]),
                           pack2<BF16>(o[6], o[7]));
          *s1 = make_uint4(pack2<BF16>(o[8], o[9]), pack2<BF16>(o[10], o[11]), pack2<BF16>(o[12], o[13]),
                           pack2<BF16>(o[14], o[15]));
        }
        if (p.row_stats && valid)
          reinterpret_cast<float2*>(p.row_stats)[(long long)m * p.stat_slots + n_tile * 2 + half] = make_float2(st_s, st_q);
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);  // accumulator stage free for the MMAs of tile it+2
        if (p.bias && et < BLOCK_N) s_bias[((it + 1) & 1) * BLOCK_N + et] = bias_next;
        if (p.ln_stats && et < BLOCK_N) s_cs[((it + 1) & 1) * BLOCK_N + et] = cs_next;
        fence_proxy_async_smem();
        named_bar_sync(1, PG_EPI_THREADS);
        if (et == 0) {
#pragma unroll
          for (int sub = 0; sub < BLOCK_N / 32; ++sub) tma_store_2d(&tmC, stg + sub * SUB_BYTES, n0 + sub * 32, m0);
          tma_store_commit();
        }
      } else {
        // direct stores (convolutions / fp32 outputs / GEGLU): residual + per-image row bias fetched ahead of use
        const bool geglu = p.act == PF_ACT_GEGLU;
        constexpr int HALF_N = BLOCK_N / 2;
        const bool res16 = p.residual != nullptr && !p.res_f32 && valid;
        const uint4* rsrc =
            reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + orow * p.res_ld + n0);
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        uint4 ra0 = z4, ra1 = z4;
        if (res16 && half < NCH) {
          ra0 = __ldg(rsrc + 2 * half);
          ra1 = __ldg(rsrc + 2 * half + 1);
        }
        float4 rbn[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          rbn[e] = rb_base ? __ldg(reinterpret_cast<const float4*>(rb_base + half * 16) + e) : make_float4(0, 0, 0, 0);
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        tc_fence_after();
        if (geglu) {
          const int on0 = n_tile * HALF_N;
          if constexpr (HALF_N % 16 == 0) {
#pragma unroll 1
            for (int c = half * 16; c < HALF_N; c += 32) {
              uint32_t va[16], vg[16];
              tmem_ld16(taddr_row + c, va);
              tmem_ld16(taddr_row + HALF_N + c, vg);
              tmem_ld_wait();
              if (valid) {
                float o[16], ba[16], bg[16];
                if (p.bias) {  // HALF_N and c are multiples of 16: 128-bit shared-memory loads
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    *reinterpret_cast<float4*>(ba + 4 * e) = *reinterpret_cast<const float4*>(bias_s + c + 4 * e);
                    *reinterpret_cast<float4*>(bg + 4 * e) = *reinterpret_cast<const float4*>(bias_s + HALF_N + c + 4 * e);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 16; ++e) ba[e] = bg[e] = 0.f;
                }
                if (p.ln_stats) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) {
                    va[e] = __float_as_uint(fmaf(__uint_as_float(va[e]), ln_a, cs_s[c + e] * ln_b));
                    vg[e] = __float_as_uint(fmaf(__uint_as_float(vg[e]), ln_a, cs_s[HALF_N + c + e] * ln_b));
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
#pragma unroll 1
          for (int ci = half; ci < NCH; ci += 2) {
            const int c = ci * 16;
            uint32_t v[16];
            tmem_ld16(taddr_row + c, v);
            float4 rbc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) rbc[e] = rbn[e];
            const uint4 r0 = ra0, r1 = ra1;
            if (ci + 2 < NCH) {
              if (rb_base) {
#pragma unroll
                for (int e = 0; e < 4; ++e) rbn[e] = __ldg(reinterpret_cast<const float4*>(rb_base + c + 32) + e);
              }
              if (res16) {
                ra0 = __ldg(rsrc + 2 * (ci + 2));
                ra1 = __ldg(rsrc + 2 * (ci + 2) + 1);
              }
            }
            tmem_ld_wait();
            if (valid) {
              float o[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(v[e]);
              if (p.bias) {
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] += bias_s[c + e];
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
                  const float4* r4 = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) +
                                                                     orow * p.res_ld + n0 + c);
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
        mbar_arrive(&tmem_empty[acc]);
        if (p.bias && et < BLOCK_N) s_bias[((it + 1) & 1) * BLOCK_N + et] = bias_next;
        if (p.ln_stats && et < BLOCK_N) s_cs[((it + 1) & 1) * BLOCK_N + et] = cs_next;
        named_bar_sync(1, PG_EPI_THREADS);  // publishes the next bias row; keeps the two bias buffers in step
      }
    }
    if constexpr (EPI_TMA) {
      if (et == 0) tma_store_wait_read();  // shared memory must outlive the last TMA store's reads
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int STAGES, bool EPI_TMA>
static int launch_persist(const pf_gemm_args* a, const GemmKernelParams& kp, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmC, tmR;
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
    uint32_t box[2] = {GEMM_BLOCK_K, (uint32_t)BLOCK_N};
    if ((rc = make_tmap(&tmB, a->dtype, 2, a->B, dims, str, box, 128))) return rc;
  }
  tmC = tmA;
  tmR = tmA;
  if constexpr (EPI_TMA) {
    uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
    uint32_t box[2] = {32, GEMM_BLOCK_M};
    uint64_t str[1] = {(uint64_t)a->out_ld * 2};
    if ((rc = make_tmap(&tmC, a->dtype, 2, a->out, dims, str, box, 64))) return rc;
    if (a->residual) {
      uint64_t rstr[1] = {(uint64_t)a->res_ld * 2};
      if ((rc = make_tmap(&tmR, a->dtype, 2, a->residual, dims, rstr, box, 64))) return rc;
    }
  }
  constexpr int SMEM = pg_smem_bytes(BLOCK_N, STAGES, EPI_TMA);
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  const int m_tiles = (a->M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tiles = m_tiles * (a->N / BLOCK_N);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (a->dtype == PF_BF16) {
    auto kern = gemm_persist_kernel<BLOCK_N, STAGES, true, EPI_TMA>;
    static bool attr_set = false;
    if (!attr_set) {
      if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM),
                           "cudaFuncSetAttribute(gemm_persist)")))
        return rc;
      attr_set = true;
    }
    if ((rc = check_cuda(launch_pdl(kern, dim3(grid), dim3(PG_THREADS), SMEM, st, tmA, tmB, tmC, tmR, kp), "launch(gemm_persist)"))) return rc;
  } else {
    auto kern = gemm_persist_kernel<BLOCK_N, STAGES, false, EPI_TMA>;
    static bool attr_set = false;
    if (!attr_set) {
      if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM),
                           "cudaFuncSetAttribute(gemm_persist)")))
        return rc;
      attr_set = true;
    }
    if ((rc = check_cuda(launch_pdl(kern, dim3(grid), dim3(PG_THREADS), SMEM, st, tmA, tmB, tmC, tmR, kp), "launch(gemm_persist)"))) return rc;
  }
  PF_CHECK_LAUNCH("gemm_persist_kernel");
  return PF_OK;
}

int launch_gemm_persistent(const pf_gemm_args* a, const GemmKernelParams& kp, int bn, bool epi_tma, cudaStream_t st) {
  if (epi_tma) {
    switch (bn) {
      case 64: return launch_persist<64, 6, true>(a, kp, st);
      case 128: return launch_persist<128, 5, true>(a, kp, st);
      case 160: return launch_persist<160, 4, true>(a, kp, st);
    }
  }
  switch (bn) {
    case 64: return launch_persist<64, 8, false>(a, kp, st);
    case 128: return launch_persist<128, 6, false>(a, kp, st);
    case 160: return launch_persist<160, 6, false>(a, kp, st);
    case 256: return launch_persist<256, 4, false>(a, kp, st);
  }
  set_error("pf_gemm_taps: unsupported block_n %d", bn);
  return PF_ERR_UNSUPPORTED;
}

}  // namespace pf
