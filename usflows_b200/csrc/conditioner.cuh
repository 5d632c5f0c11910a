// Row-wise glue of the reference's own MLP-style conditioner, `networks.ConvNet` with 1-D in_dims (networks.py:287-307,
// 379-389): GatedMLP (networks.py:222-245) `x + val * sigmoid(gate)` and LayerNormVector (networks.py:205-219) between
// the contractions.  The contractions themselves run on usf_linear; this kernel does everything between two of them in
// ONE pass over the row (HBM-bound):
//
//   v = gated ? xres[r, j] + o[r, j] * sigmoid(o[r, n + j]) : o[r, j]          o = [val | gate] of the preceding Linear
//   v = pre_relu ? max(v, 0) : v                                               (ConvNet2D: GatedConv -> ReLU -> LayerNorm)
//   v = gamma ? (v - mean_r) / sqrt(var_r + eps) * gamma[j] + beta[j] : v        (biased variance, as nn.LayerNorm)
//   y_f32 <- v;   act planes <- relu ? max(v, 0) : v;   raw planes <- v            (every output optional)
//
// G lanes per row (a whole warp for wide rows, 8 lanes for 32-wide ones); the row is staged in shared memory between the
// statistics passes and the store pass (8 x 32/G rows x n floats of dynamic shared memory per block).  y_f32 may alias xres (the residual stream is updated in place:
// a row is read completely before any of it is written).
#pragma once
#include "elementwise.cuh"
#include "radial.cuh"

namespace usf {

constexpr int GN_THREADS = 256;

template <int G>
__device__ __forceinline__ float group_sum(float v) {      // sum over an aligned group of G lanes
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// G lanes per row (32 / G rows per warp): narrow rows (n = 32 channels of an image conditioner) keep every lane busy.
template <bool VEC, int G>
__global__ void __launch_bounds__(GN_THREADS)
gate_norm_kernel(const float* __restrict__ o, long long ldo, const float* xres, long long ldx, long long rows,
                 int n, int gated, int pre_relu, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* y_f32, long long ldy, OutPlanes act, int act_on, int act_relu, OutPlanes raw, int raw_on) {
  extern __shared__ __align__(16) float gn_smem[];
  constexpr int RPW = 32 / G;                                   // rows per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / G, gl = lane % G;
  const int wpb = GN_THREADS / 32;
  float* row = gn_smem + ((size_t)warp * RPW + sub) * n;
  for (long long r0 = ((long long)blockIdx.x * wpb + warp) * RPW; r0 < rows; r0 += (long long)gridDim.x * wpb * RPW) {
    const long long r = r0 + sub;
    const bool valid = r < rows;
    const float* orow = o + r * ldo;
    const float* xr = gated ? xres + r * ldx : nullptr;
    float sum = 0.f;
    if (valid) {
      if (VEC) {
        for (int j = gl * 4; j < n; j += G * 4) {
          float4 v = __ldcs(reinterpret_cast<const float4*>(orow + j));
          if (gated) {
            const float4 g = __ldcs(reinterpret_cast<const float4*>(orow + n + j));
            const float4 x = *reinterpret_cast<const float4*>(xr + j);
            v.x = x.x + v.x * (1.f / (1.f + expf(-g.x)));
            v.y = x.y + v.y * (1.f / (1.f + expf(-g.y)));
            v.z = x.z + v.z * (1.f / (1.f + expf(-g.z)));
            v.w = x.w + v.w * (1.f / (1.f + expf(-g.w)));
          }
          if (pre_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4*>(row + j) = v;
          sum += (v.x + v.y) + (v.z + v.w);
        }
      } else {
        for (int j = gl; j < n; j += G) {
          float v = orow[j];
          if (gated) v = xr[j] + v * (1.f / (1.f + expf(-orow[n + j])));
          if (pre_relu) v = fmaxf(v, 0.f);
          row[j] = v;
          sum += v;
        }
      }
    }
    __syncwarp();
    float mean = 0.f, rstd = 1.f;
    if (gamma) {
      mean = group_sum<G>(sum) / (float)n;
      float sq = 0.f;
      if (valid) {
        for (int j = gl; j < n; j += G) {
          const float t = row[j] - mean;
          sq = fmaf(t, t, sq);
        }
      }
      rstd = 1.f / sqrtf(group_sum<G>(sq) / (float)n + eps);
    }
    if (valid) {
      if (VEC) {
        for (int j = gl * 4; j < n; j += G * 4) {
          const float4 t = *reinterpret_cast<const float4*>(row + j);
          float v[4] = {t.x, t.y, t.z, t.w};
          if (gamma) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + j));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + j));
            v[0] = (v[0] - mean) * rstd * g.x + b.x;
            v[1] = (v[1] - mean) * rstd * g.y + b.y;
            v[2] = (v[2] - mean) * rstd * g.z + b.z;
            v[3] = (v[3] - mean) * rstd * g.w + b.w;
          }
          if (y_f32) *reinterpret_cast<float4*>(y_f32 + r * ldy + j) = make_float4(v[0], v[1], v[2], v[3]);
          if (raw_on) store_planes4(raw, r, j, v);
          if (act_on) {
            if (act_relu) {
#pragma unroll
              for (int t2 = 0; t2 < 4; ++t2) v[t2] = fmaxf(v[t2], 0.f);
            }
            store_planes4(act, r, j, v);
          }
        }
      } else {
        for (int j = gl; j < n; j += G) {
          float v = row[j];
          if (gamma) v = (v - mean) * rstd * __ldg(gamma + j) + __ldg(beta + j);
          if (y_f32) y_f32[r * ldy + j] = v;
          if (raw_on) store_planes1(raw, r, j, v);
          if (act_on) store_planes1(act, r, j, act_relu ? fmaxf(v, 0.f) : v);
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace usf
