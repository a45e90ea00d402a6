// Spherical projection grids shared by the resampling kernels (resample.cu) and the EPPA table builders
// (eppa_tables.cu). fp64 grid math following external/Perspective_and_Equirectangular/e2p.py:9-51 and
// p2e.py:9-49, fp32 sampling arithmetic following kornia.remap + ATen grid_sampler (align_corners=True, zeros).
#pragma once
#include <math.h>

#include "pf_common.cuh"

namespace pf {

struct Taps {
  int idx[4];   // linear source index (y*W+x) or -1 when out of bounds / masked
  float w[4];
};

// kornia.geometry.transform.remap -> normalize_pixel_coordinates (factor = 2/(size-1), fp32) followed by
// ATen grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1), all in fp32, no FMA contraction.
__device__ __forceinline__ float roundtrip_coord(float pix, int size) {
  const float factor = __fdiv_rn(2.0f, fmaxf(float(size - 1), 1e-14f));
  const float g = __fsub_rn(__fmul_rn(factor, pix), 1.0f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), float(size - 1));
}

// ATen CPU GridSamplerKernel bilinear: w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n;
// nw = s*e, ne = s*w, sw = n*e, se = n*w; out-of-range taps are dropped (padding_mode='zeros').
__device__ __forceinline__ void make_taps(float px, float py, int H, int W, int mode, bool live, Taps& t) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    t.idx[k] = -1;
    t.w[k] = 0.f;
  }
  if (!live) return;
  const float x = roundtrip_coord(px, W);
  const float y = roundtrip_coord(py, H);
  if (mode == 1) {  // nearest: round half to even
    const float xr = nearbyintf(x), yr = nearbyintf(y);
    if (xr >= 0.f && xr <= float(W - 1) && yr >= 0.f && yr <= float(H - 1)) {
      t.idx[0] = int(yr) * W + int(xr);
      t.w[0] = 1.f;
    }
    return;
  }
  const float xw = floorf(x), yn = floorf(y);
  const float w = __fsub_rn(x, xw), e = __fsub_rn(1.0f, w);
  const float n = __fsub_rn(y, yn), s = __fsub_rn(1.0f, n);
  // NaN coordinates compare false everywhere -> all taps dropped (matches the masked gathers)
  const bool x0 = xw >= 0.f && xw <= float(W - 1);
  const bool x1 = (xw + 1.f) >= 0.f && (xw + 1.f) <= float(W - 1);
  const bool y0 = yn >= 0.f && yn <= float(H - 1);
  const bool y1 = (yn + 1.f) >= 0.f && (yn + 1.f) <= float(H - 1);
  const int ix = int(xw), iy = int(yn);
  if (x0 && y0) { t.idx[0] = iy * W + ix;           t.w[0] = __fmul_rn(s, e); }
  if (x1 && y0) { t.idx[1] = iy * W + ix + 1;       t.w[1] = __fmul_rn(s, w); }
  if (x0 && y1) { t.idx[2] = (iy + 1) * W + ix;     t.w[2] = __fmul_rn(n, e); }
  if (x1 && y1) { t.idx[3] = (iy + 1) * W + ix + 1; t.w[3] = __fmul_rn(n, w); }
}

// numpy.linspace(start, stop, n)[i]: i*step + start with step = (stop-start)/(n-1); last element pinned to stop.
__device__ __forceinline__ double np_linspace(double start, double stop, int n, int i) {
  if (n == 1) return start;
  if (i == n - 1) return stop;
  const double step = (stop - start) / double(n - 1);
  return double(i) * step + start;
}

__device__ __forceinline__ void rot3(const double* R, double x, double y, double z, double& ox, double& oy,
                                     double& oz) {
  ox = R[0] * x + R[1] * y + R[2] * z;
  oy = R[3] * x + R[4] * y + R[5] * z;
  oz = R[6] * x + R[7] * y + R[8] * z;
}

// e2p.py:9-51 (map_pers_coords_to_equi + map_pers_pix_to_equi): pixel coords in the equirect image
__device__ __forceinline__ void e2p_grid(const double* cam, int r, int c, int h, int w, int He, int We, float& px,
                                         float& py, double* lon_out = nullptr, double* lat_out = nullptr) {
  const double w_len = cam[18], h_len = cam[19];
  const double ym = np_linspace(-w_len, w_len, w, c);
  const double zm = -np_linspace(-h_len, h_len, h, r);
  const double D = sqrt(1.0 + ym * ym + zm * zm);
  const double vx = 1.0 / D, vy = ym / D, vz = zm / D;
  double ax, ay, az, bx, by, bz;
  rot3(cam, vx, vy, vz, ax, ay, az);
  rot3(cam + 9, ax, ay, az, bx, by, bz);
  double lat = asin(bz);
  const double lon = atan2(by, bx);
  lat = -lat;
  if (lon_out) *lon_out = lon;
  if (lat_out) *lat_out = lat;
  const double cx = double(We - 1) / 2.0, cy = double(He - 1) / 2.0;
  const double lon_d = lon / M_PI * 180.0, lat_d = lat / M_PI * 180.0;
  px = float(lon_d / 180.0 * cx + cx);
  py = float(lat_d / 90.0 * cy + cy);
}

// p2e.py:9-49 (map_equi_pix_to_pers): pixel coords in the perspective image + validity mask
__device__ __forceinline__ void p2e_grid(const double* cam, int i, int j, int He, int We, int hp, int wp, float& px,
                                         float& py, bool& mask) {
  const double w_len = cam[18], h_len = cam[19];
  const double xd = np_linspace(-180.0, 180.0, We, j);
  const double yd = np_linspace(90.0, -90.0, He, i);
  const double xr = xd * (M_PI / 180.0), yr = yd * (M_PI / 180.0);
  const double vx = cos(xr) * cos(yr), vy = sin(xr) * cos(yr), vz = sin(yr);
  double ax, ay, az, bx, by, bz;
  rot3(cam + 9, vx, vy, vz, ax, ay, az);  // inv(R2) first (p2e.py:32)
  rot3(cam, ax, ay, az, bx, by, bz);      // then inv(R1) (p2e.py:33)
  const bool front = bx > 0.0;
  const double u = by / bx, v = bz / bx;
  const bool inside = (-w_len < u) && (u < w_len) && (-h_len < v) && (v < h_len);
  const double lon_map = inside ? (u + w_len) / 2.0 / w_len * double(wp) : 0.0;
  const double lat_map = inside ? (-v + h_len) / 2.0 / h_len * double(hp) : 0.0;
  px = float(lon_map);
  py = float(lat_map);
  mask = inside && front;
}


}  // namespace pf
