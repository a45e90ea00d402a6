// EPPA geometry tables: correspondence bias (both directions) and spherical positional encodings, built directly
// from the camera records — no one-hot tensors, no grid_sample over thousands of channels, no host round trip.
//
// Reference: models/pano/utils.py:10-84 (get_masks) builds, per camera, two dense stacks of one-hot images
// (67 MB + 268 MB at the 32x32 / 32x64 level), warps them with p2e / e2p, unions, blurs and normalises them
// (9.97 s on 8 CPU cores per call, twice per step). The same numbers follow from the 4 bilinear taps of each pixel:
//   a(e,p) = weight of equirect pixel e in the e2p sample of perspective pixel p      (utils.py:38-41)
//   b(p,e) = weight of perspective pixel p in the masked p2e sample of equirect pixel e (utils.py:34-37)
//   A  = clamp(a + b, 0, 1);  Bm = clamp(b + A, 0, 1)                                  (utils.py:52-60)
//   5x5 gaussian (sigma 1): A rows over the perspective image, replicate border; Bm rows over the equirect image,
//   circular in W / replicate in H (utils.py:63-68); per-row max normalisation, *2-1   (utils.py:69-76)
// Kernel 1 writes the tap tables, kernels 2/3 expand one query row per CTA in shared memory and write the bias
// tables directly in the layout the attention kernel reads (modules.py:46,53):
//   bias1[g][e][vl*P + p]  (query = pano pixel e, keys = all views of group g)
//   bias2[g][vl*P + p][e]  (query = view pixel, keys = pano pixels)
// PE tables: models/pano/utils.py:87-106 (get_coords) + models/modules/transformer.py:185-201 (SphericalPE).
#include "sphere_grid.cuh"

namespace pf {

// ---- tap tables -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eppa_taps_kernel(const double* __restrict__ cams_e2p, const double* __restrict__ cams_p2e, int V, int ph, int pw,
                 int eh, int ew, int* __restrict__ a_idx, float* __restrict__ a_w, int* __restrict__ b_idx,
                 float* __restrict__ b_w) {
  const int P = ph * pw, E = eh * ew;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  Taps t;
  if (idx < P) {
    float px, py;
    e2p_grid(cams_e2p + (size_t)v * PF_CAM_DOUBLES, idx / pw, idx % pw, ph, pw, eh, ew, px, py);
    make_taps(px, py, eh, ew, 0, true, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a_idx[((size_t)v * P + idx) * 4 + k] = t.idx[k];
      a_w[((size_t)v * P + idx) * 4 + k] = t.w[k];
    }
  }
  if (idx < E) {
    float px, py;
    bool m;
    p2e_grid(cams_p2e + (size_t)v * PF_CAM_DOUBLES, idx / ew, idx % ew, eh, ew, ph, pw, px, py, m);
    make_taps(px, py, ph, pw, 0, m, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      b_idx[((size_t)v * E + idx) * 4 + k] = t.idx[k];
      b_w[((size_t)v * E + idx) * 4 + k] = t.w[k];
    }
  }
}

struct Blur5 {
  float k[5];
};

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  return r;
}

// ---- bias for direction 1: CTA <-> (equirect query pixel e, view v); image over perspective pixels ----
__global__ void __launch_bounds__(256)
eppa_bias1_kernel(const int* __restrict__ a_idx, const float* __restrict__ a_w, const int* __restrict__ b_idx,
                  const float* __restrict__ b_w, int ph, int pw, int E, int m, Blur5 bk, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int P = ph * pw;
  float* img = sm;
  float* tmp = sm + P;
  __shared__ float red[8];
  const int e = blockIdx.x, v = blockIdx.y;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float s = 0.f;
    const int4 id = __ldg(reinterpret_cast<const int4*>(a_idx + ((size_t)v * P + p) * 4));
    const float4 w = __ldg(reinterpret_cast<const float4*>(a_w + ((size_t)v * P + p) * 4));
    if (id.x == e) s += w.x;
    if (id.y == e) s += w.y;
    if (id.z == e) s += w.z;
    if (id.w == e) s += w.w;
    img[p] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t o = ((size_t)v * E + e) * 4;
    for (int k = 0; k < 4; ++k) {
      const int p = b_idx[o + k];
      if (p >= 0) img[p] += b_w[o + k];
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) img[p] = fminf(fmaxf(img[p], 0.f), 1.f);
  __syncthreads();
  // horizontal then vertical 5-tap pass, replicate border (kornia gaussian_blur2d separable)
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int y = p / pw, x = p % pw;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) s += bk.k[k] * img[y * pw + min(max(x + k - 2, 0), pw - 1)];
    tmp[p] = s;
  }
  __syncthreads();
  float mx = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int y = p / pw, x = p % pw;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) s += bk.k[k] * tmp[min(max(y + k - 2, 0), ph - 1) * pw + x];
    img[p] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_max(mx, red);
  if (mx == 0.f) mx = 1.f;
  const int g = v / m, vl = v % m;
  float* orow = out + ((size_t)g * E + e) * ((size_t)m * P) + (size_t)vl * P;
  for (int p = threadIdx.x; p < P; p += blockDim.x) orow[p] = img[p] / mx * 2.f - 1.f;
}

// ---- bias for direction 2: CTA <-> (perspective query pixel p, view v); image over equirect pixels ----
__global__ void __launch_bounds__(256)
eppa_bias2_kernel(const int* __restrict__ a_idx, const float* __restrict__ a_w, const int* __restrict__ b_idx,
                  const float* __restrict__ b_w, int P, int eh, int ew, int m, Blur5 bk, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int E = eh * ew;
  float* img = sm;
  float* tmp = sm + E;
  __shared__ float red[8];
  const int p = blockIdx.x, v = blockIdx.y;
  // b(p, e) for every e
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float s = 0.f;
    const int4 id = __ldg(reinterpret_cast<const int4*>(b_idx + ((size_t)v * E + e) * 4));
    const float4 w = __ldg(reinterpret_cast<const float4*>(b_w + ((size_t)v * E + e) * 4));
    if (id.x == p) s += w.x;
    if (id.y == p) s += w.y;
    if (id.z == p) s += w.z;
    if (id.w == p) s += w.w;
    img[e] = s;
    tmp[e] = 0.f;  // a(e, p)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t o = ((size_t)v * P + p) * 4;
    for (int k = 0; k < 4; ++k) {
      const int e = a_idx[o + k];
      if (e >= 0) tmp[e] += a_w[o + k];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float b = img[e];
    const float A = fminf(fmaxf(tmp[e] + b, 0.f), 1.f);  // updated pers mask entry (utils.py:55-56)
    img[e] = fminf(fmaxf(b + A, 0.f), 1.f);              // utils.py:59-60
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int y = e / ew, x = e % ew;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      int xx = x + k - 2;
      xx = xx < 0 ? xx + ew : (xx >= ew ? xx - ew : xx);  // pad_pano(2): circular in W
      s += bk.k[k] * img[y * ew + xx];
    }
    tmp[e] = s;
  }
  __syncthreads();
  float mx = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int y = e / ew, x = e % ew;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) s += bk.k[k] * tmp[min(max(y + k - 2, 0), eh - 1) * ew + x];
    img[e] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_max(mx, red);
  if (mx == 0.f) mx = 1.f;
  const int g = v / m, vl = v % m;
  float* orow = out + ((size_t)g * m * P + (size_t)vl * P + p) * (size_t)E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) orow[e] = img[e] / mx * 2.f - 1.f;
}

// ---- positional encodings ----------------------------------------------------------------------------
// thread <-> (token, frequency k): writes sin(lon f), sin(lat f), cos(lon f), cos(lat f) at channels k, N+k, 2N+k, 3N+k
__global__ void __launch_bounds__(256)
eppa_pe_kernel(const double* __restrict__ cams_e2p, int V, int ph, int pw, int eh, int ew,
               const float* __restrict__ freq, int nf, float* __restrict__ pers_pe, float* __restrict__ equi_pe) {
  const int P = ph * pw, E = eh * ew;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = ((long long)V * P + E) * nf;
  if (idx >= total) return;
  const int k = int(idx % nf);
  const long long tok = idx / nf;
  float lon, lat;
  float* dst;
  if (tok < (long long)V * P) {
    const int v = int(tok / P), pp = int(tok % P);
    float px, py;
    double dlon, dlat;
    e2p_grid(cams_e2p + (size_t)v * PF_CAM_DOUBLES, pp / pw, pp % pw, ph, pw, eh, ew, px, py, &dlon, &dlat);
    lon = float(dlon);
    lat = float(dlat);  // down-positive for perspective pixels (e2p.py:35)
    dst = pers_pe + (size_t)tok * 4 * nf;
  } else {
    const int e = int(tok - (long long)V * P);
    lon = float(np_linspace(-M_PI, M_PI, ew, e % ew));          // models/pano/utils.py:92
    lat = float(np_linspace(M_PI / 2, -M_PI / 2, eh, e / ew));   // up-positive for the panorama
    dst = equi_pe + (size_t)e * 4 * nf;
  }
  const float f = __ldg(freq + k);
  const float al = __fmul_rn(lon, f), at = __fmul_rn(lat, f);
  dst[k] = sinf(al);
  dst[nf + k] = sinf(at);
  dst[2 * nf + k] = cosf(al);
  dst[3 * nf + k] = cosf(at);
}

// flags[g][qt][kt] = 1 iff the 128 x 64 bias tile is entirely -1 (row of "no correspondence" entries)
__global__ void __launch_bounds__(256)
bias_tile_flags_kernel(const float* __restrict__ bias, int Lq, int Lk, int ld, long long bstride,
                       uint8_t* __restrict__ flags) {
  const int kt = blockIdx.x, qt = blockIdx.y, g = blockIdx.z;
  const float* base = bias + (long long)g * bstride;
  int ok = 1;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = qt * 128 + i / 64, c = kt * 64 + i % 64;
    if (r < Lq && c < Lk && __ldg(base + (long long)r * ld + c) != -1.0f) ok = 0;
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0) flags[((size_t)g * gridDim.y + qt) * gridDim.x + kt] = (uint8_t)ok;
}

// ---- tile-packed bias: only the 128 x 64 tiles that are NOT entirely -1 are kept (~15 % of an EPPA table) ----
// single-block exclusive scan of the "tile is live" bits -> tile_off[t] = index of tile t in the packed store, -1 for constant
// tiles; n_live[0] = number of live tiles. T <= a few 10^5 tiles: one block walks them in chunks of 1024.
__global__ void __launch_bounds__(1024) bias_tile_scan_kernel(const uint8_t* __restrict__ flags, int T, int* __restrict__ tile_off,
                                                              int* __restrict__ n_live) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    const int live = (t < T && flags[t] == 0) ? 1 : 0;
    int v = live;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) >= o) v += n;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += n;
      }
      s_warp[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = s_base + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0) + v - live;
    if (t < T) tile_off[t] = live ? before : -1;
    __syncthreads();
    if (threadIdx.x == 0) s_base += s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_live[0] = s_base;
}

// packed[off][128][64] <- dense tile (qt, kt) of batch g; positions outside [Lq, Lk] are written as 0 (the attention kernel
// masks keys >= Lk and never stores rows >= Lq)
__global__ void __launch_bounds__(256)
bias_tile_pack_kernel(const float* __restrict__ bias, int Lq, int Lk, int ld, long long bstride,
                      const int* __restrict__ tile_off, float* __restrict__ packed) {
  const int kt = blockIdx.x, qt = blockIdx.y, g = blockIdx.z;
  const int off = tile_off[((size_t)g * gridDim.y + qt) * gridDim.x + kt];
  if (off < 0) return;
  const float* base = bias + (long long)g * bstride;
  float* dst = packed + (size_t)off * (128 * 64);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int rl = i / 64, cl = i % 64;
    const int r = qt * 128 + rl, c = kt * 64 + cl;
    // lane-interleaved: the attention kernel's thread = query row (warp rl / 32, lane rl % 32) reads 16-byte piece cl / 4;
    // piece e of one warp is 32 lanes x 16 B = 512 contiguous bytes
    const int o = (((rl >> 5) * 16 + (cl >> 2)) * 32 + (rl & 31)) * 4 + (cl & 3);
    dst[o] = (r < Lq && c < Lk) ? __ldg(base + (long long)r * ld + c) : 0.f;
  }
}

}  // namespace pf

extern "C" int pf_bias_tile_scan(const uint8_t* flags, int num_tiles, int* tile_off, int* n_live, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(flags && tile_off && n_live && num_tiles > 0, "pf_bias_tile_scan: bad arguments");
  bias_tile_scan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(flags, num_tiles, tile_off, n_live);
  PF_CHECK_LAUNCH("bias_tile_scan_kernel");
  return PF_OK;
}

extern "C" int pf_bias_tile_pack(const float* bias, int G, int Lq, int Lk, int bias_ld, int64_t bias_bstride,
                                 const int* tile_off, float* packed, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(bias && tile_off && packed && G > 0 && Lq > 0 && Lk > 0 && bias_ld >= Lk, "pf_bias_tile_pack: bad arguments");
  dim3 grid((Lk + 63) / 64, (Lq + 127) / 128, G);
  PF_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "pf_bias_tile_pack: too many tiles");
  bias_tile_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(bias, Lq, Lk, bias_ld, bias_bstride, tile_off, packed);
  PF_CHECK_LAUNCH("bias_tile_pack_kernel");
  return PF_OK;
}

extern "C" int pf_bias_tile_flags(const float* bias, int G, int Lq, int Lk, int bias_ld, int64_t bias_bstride,
                                  uint8_t* flags, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(bias && flags && G > 0 && Lq > 0 && Lk > 0 && bias_ld >= Lk, "pf_bias_tile_flags: bad arguments");
  dim3 grid((Lk + 63) / 64, (Lq + 127) / 128, G);
  PF_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "pf_bias_tile_flags: too many tiles");
  bias_tile_flags_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(bias, Lq, Lk, bias_ld, bias_bstride, flags);
  PF_CHECK_LAUNCH("bias_tile_flags_kernel");
  return PF_OK;
}

extern "C" int pf_eppa_tables(const double* cams_e2p, const double* cams_p2e, int V, int m, int ph, int pw, int eh,
                              int ew, const float* blur5, int* ws_idx, float* ws_w, float* bias1, float* bias2,
                              void* stream) {
  using namespace pf;
  PF_CHECK_ARG(cams_e2p && cams_p2e && blur5 && ws_idx && ws_w && bias1 && bias2, "pf_eppa_tables: null pointer");
  PF_CHECK_ARG(V > 0 && m > 0 && V % m == 0 && ph > 0 && pw > 0 && eh > 0 && ew > 0 && V <= 65535,
               "pf_eppa_tables: bad shape V=%d m=%d", V, m);
  const int P = ph * pw, E = eh * ew;
  PF_CHECK_ARG((size_t)2 * (P > E ? P : E) * sizeof(float) <= 200 * 1024, "pf_eppa_tables: level %dx%d / %dx%d too large",
               ph, pw, eh, ew);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* a_idx = ws_idx;
  int* b_idx = ws_idx + (size_t)V * P * 4;
  float* a_w = ws_w;
  float* b_w = ws_w + (size_t)V * P * 4;
  {
    const int mx = P > E ? P : E;
    dim3 grid((mx + 255) / 256, V);
    eppa_taps_kernel<<<grid, 256, 0, st>>>(cams_e2p, cams_p2e, V, ph, pw, eh, ew, a_idx, a_w, b_idx, b_w);
    PF_CHECK_LAUNCH("eppa_taps_kernel");
  }
  Blur5 bk;
  for (int i = 0; i < 5; ++i) bk.k[i] = blur5[i];
  int rc;
  {
    const size_t smem = 2 * (size_t)P * sizeof(float);
    if ((rc = check_cuda(cudaFuncSetAttribute(eppa_bias1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "bias1 attr"))) return rc;
    eppa_bias1_kernel<<<dim3(E, V), 256, smem, st>>>(a_idx, a_w, b_idx, b_w, ph, pw, E, m, bk, bias1);
    PF_CHECK_LAUNCH("eppa_bias1_kernel");
  }
  {
    const size_t smem = 2 * (size_t)E * sizeof(float);
    if ((rc = check_cuda(cudaFuncSetAttribute(eppa_bias2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "bias2 attr"))) return rc;
    eppa_bias2_kernel<<<dim3(P, V), 256, smem, st>>>(a_idx, a_w, b_idx, b_w, P, eh, ew, m, bk, bias2);
    PF_CHECK_LAUNCH("eppa_bias2_kernel");
  }
  return PF_OK;
}

extern "C" int pf_eppa_pe(const double* cams_e2p, int V, int ph, int pw, int eh, int ew, const float* freq_bands,
                          int n_freqs, float* pers_pe, float* equi_pe, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(cams_e2p && freq_bands && pers_pe && equi_pe, "pf_eppa_pe: null pointer");
  PF_CHECK_ARG(V > 0 && ph > 0 && pw > 0 && eh > 0 && ew > 0 && n_freqs > 0, "pf_eppa_pe: bad shape");
  const long long total = ((long long)V * ph * pw + (long long)eh * ew) * n_freqs;
  pf::eppa_pe_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cams_e2p, V, ph, pw, eh, ew, freq_bands, n_freqs, pers_pe, equi_pe);
  PF_CHECK_LAUNCH("eppa_pe_kernel");
  return PF_OK;
}
