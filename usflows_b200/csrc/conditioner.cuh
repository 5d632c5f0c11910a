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
// One warp per row; the row is staged in shared memory between the statistics passes and the store pass
// (8 rows x n floats of dynamic shared memory per block).  y_f32 may alias xres (the residual stream is updated in place:
// a row is read completely before any of it is written).
#pragma once
#include "elementwise.cuh"
#include "radial.cuh"

namespace usf {

constexpr int GN_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(GN_THREADS)
gate_norm_kernel(const float* __restrict__ o, long long ldo, const float* xres, long long ldx, long long rows,
                 int n, int gated, int pre_relu, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* y_f32, long long ldy, OutPlanes act, int act_on, int act_relu, OutPlanes raw, int raw_on) {
  extern __shared__ __align__(16) float gn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = GN_THREADS / 32;
  float* row = gn_smem + (size_t)warp * n;
  for (long long r = (long long)blockIdx.x * wpb + warp; r < rows; r += (long long)gridDim.x * wpb) {
    const float* orow = o + r * ldo;
    const float* xr = gated ? xres + r * ldx : nullptr;
    float sum = 0.f;
    if (VEC) {
      for (int j = lane * 4; j < n; j += 128) {
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
      for (int j = lane; j < n; j += 32) {
        float v = orow[j];
        if (gated) v = xr[j] + v * (1.f / (1.f + expf(-orow[n + j])));
        if (pre_relu) v = fmaxf(v, 0.f);
        row[j] = v;
        sum += v;
      }
    }
    __syncwarp();
    float mean = 0.f, rstd = 1.f;
    if (gamma) {
      mean = warp_sum(sum) / (float)n;
      float sq = 0.f;
      for (int j = lane; j < n; j += 32) {
        const float t = row[j] - mean;
        sq = fmaf(t, t, sq);
      }
      rstd = 1.f / sqrtf(warp_sum(sq) / (float)n + eps);
    }
    if (VEC) {
      for (int j = lane * 4; j < n; j += 128) {
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
      for (int j = lane; j < n; j += 32) {
        float v = row[j];
        if (gamma) v = (v - mean) * rstd * __ldg(gamma + j) + __ldg(beta + j);
        if (y_f32) y_f32[r * ldy + j] = v;
        if (raw_on) store_planes1(raw, r, j, v);
        if (act_on) store_planes1(act, r, j, act_relu ? fmaxf(v, 0.f) : v);
      }
    }
    __syncwarp();
  }
}

}  // namespace usf
