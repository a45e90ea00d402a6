// Shared declarations of the tap-GEMM kernels (gemm_persist.cu, gemm_tcgen05.cu).
#pragma once
#include "pf_common.cuh"

namespace pf {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 x 16-bit = 128 B = one swizzle row

struct GemmKernelParams {
  int M, N, num_kb, kb_per_tap;
  int tap_off[PF_MAX_TAPS];
  void* out;
  int out_ld;
  int out_f32;
  const float* bias;
  const float* rowbias;
  int rowbias_ld;
  int rows_per_group;
  const void* residual;
  int res_ld;
  int res_f32;
  int act;
  int map_mode, Hm, Wm, i0, j0, Hout, Wout;
  int osy, osx, oa, ob;  // output scatter of map_mode 1: row ((img*Hout + i-i0)*osy + oa), column ((j-j0)*osx + ob)
  int k_splits;     // > 1: blockIdx.y = split index, raw fp32 partials go to ws
  float* ws;        // [k_splits][M][N]
  // fused LayerNorm (see pf_gemm_args): producer side / consumer side
  float* row_stats;        // [M][stat_slots][2] or null
  int stat_slots;          // 2 * n_tiles
  const float* ln_stats;   // [M][ln_slots][2] or null
  int ln_slots;
  const float* ln_colsum;  // [N]
  float ln_inv_k, ln_eps;
};

// mean / rstd of one row from the producer's partial sums, as the two epilogue coefficients of the LayerNorm fold:
// acc <- acc * a + colsum[n] * b with a = rstd, b = -mean * rstd
__device__ __forceinline__ void ln_row_coeffs(const GemmKernelParams& p, long long m, float& a, float& b) {
  // slots come in pairs (two per column tile): 16-byte loads, four of them in flight, summed in slot order
  const float4* st = reinterpret_cast<const float4*>(p.ln_stats) + m * (p.ln_slots >> 1);
  const int pairs = p.ln_slots >> 1;
  float s = 0.f, q = 0.f;
  for (int i0 = 0; i0 < pairs; i0 += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = i0 + u < pairs ? __ldg(st + i0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s += v[u].x;
      q += v[u].y;
      s += v[u].z;
      q += v[u].w;
    }
  }
  const float mean = s * p.ln_inv_k;
  const float var = fmaxf(q * p.ln_inv_k - mean * mean, 0.f);
  a = rsqrtf(var + p.ln_eps);
  b = -mean * a;
}

__host__ __device__ constexpr int gemm_stage_bytes(int block_n) {
  return GEMM_BLOCK_M * GEMM_BLOCK_K * 2 + block_n * GEMM_BLOCK_K * 2;
}

int launch_gemm_persistent(const pf_gemm_args* a, const GemmKernelParams& kp, int bn, bool epi_tma, cudaStream_t st);

}  // namespace pf
