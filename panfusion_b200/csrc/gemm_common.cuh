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
  int k_splits;     // > 1: blockIdx.y = split index, raw fp32 partials go to ws
  float* ws;        // [k_splits][M][N]
};

__host__ __device__ constexpr int gemm_stage_bytes(int block_n) {
  return GEMM_BLOCK_M * GEMM_BLOCK_K * 2 + block_n * GEMM_BLOCK_K * 2;
}

int launch_gemm_persistent(const pf_gemm_args* a, const GemmKernelParams& kp, int bn, bool epi_tma, cudaStream_t st);

}  // namespace pf
