// Kernels of the maximum-likelihood training step (reference: Flow.fit, src/usflows/flows.py:195-207, whose backward
// pass is torch autograd through F.linear / torch.inverse / tril / triu, transforms.py:1264-1293).
//
// The batch-side contractions of the step (forward, dX = dY . W, dW = dY^T . X) run on the tcgen05 engines
// (gemm_tc2.cuh, dW with split-K); this file holds what sits between them:
//   planes_glue_kernel        one pass over a [rows, n] matrix held as fp16 split planes: ReLU-backward mask from the saved
//                             activation, sign, the operand copy for the next dX product, the TRANSPOSED planes the dW
//                             product reads (K-major along the batch rows), and the column sums (bias gradients)
//   base_backward_kernel      d(-log p)/dz of the Laplace / Normal base density as operand planes (+ transposed planes),
//                             column sums for the gradients of loc / scale
//   mat_prep_kernel           weight-side: fp32 matrix -> (transposed / row- or column-gathered / scaled) fp32 copy and / or
//                             fp16 split operand planes
//   tri_*_batched_kernel      inverses of a stack of lower-triangular matrices by recursive block doubling: 64x64 diagonal
//                             blocks by substitution, then X21 = -X22 (T21 X11) level by level (two launches per level
//                             for the WHOLE stack, instead of the per-matrix panel sweep of prep.cuh)
//   tri_mask_kernel           gradient masks of LUTransform (transforms.py:1209-1213): strict-lower / upper part, scale,
//                             optional coef / diag(U) term of the log-determinant
#pragma once
#include "common.cuh"

namespace usf {

// ---------------------------------------------------------------------------------------------------------------------
// planes glue
// ---------------------------------------------------------------------------------------------------------------------
struct GlueArgs {
  const __half* h;        // input planes [rows, n] (value = h + l 2^-11)
  const __half* l;
  long long ld;
  const float* src_f32;   // alternative input: fp32 [rows, n] (pitch ld_src); used when h == nullptr
  long long ld_src;
  long long rows;
  int n;
  const __half* mask_h;   // optional: saved post-ReLU activation (hi plane); the value is zeroed where it is <= 0
  long long ld_mask;
  float sign;             // the value is multiplied by this (+1 / -1: exact on both planes)
  __half* out_h;          // optional: (masked, signed) planes [rows, n]; may alias the input
  __half* out_l;
  long long ld_out;
  __half* t_h;            // optional: transposed planes [n, rows]
  __half* t_l;
  long long ld_t;
  float* colsum;          // optional: colsum[j] += sum_r value[r, j]           (atomic)
  const float* mul;       // optional with colsum2: fp32 [rows, n]
  long long ld_mul;
  float* colsum2;         // optional: colsum2[j] += sum_r value[r, j] * mul[r, j]   (atomic)
  int* overflow_flag;
};

constexpr int GL_TILE = 64;   // 64 x 64 tile per CTA, 256 threads; a thread owns TWO 8-column vectors of the tile:
                              // (row = tid / 8 + 32 * i, columns (tid % 8) * 8 .. + 7), i = 0, 1
// The tile is staged for the transposed store with the 16-byte chunk index XOR-swizzled by the row group (row / 8):
// row-major vector writes and column reads (8 rows of one column per thread) are both bank-conflict free, no padding.
__device__ __forceinline__ int gl_off(int row, int col) {
  return row * GL_TILE + ((((col >> 3) ^ (row >> 3)) & 7) << 3) + (col & 7);
}
// transposed store of the staged tile: thread -> (column cc = tid / 8 + 32 * i of the tile, rows (tid % 8) * 8 .. + 7)
__device__ __forceinline__ void gl_store_transposed(const __half* sh, const __half* sl, int tid, long long r0, int c0,
                                                    long long rows, int n, __half* th, __half* tl, long long ldt) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int cc = (tid >> 3) + 32 * i;
    const int rv = (tid & 7) * 8;
    const int c = c0 + cc;
    const long long r = r0 + rv;
    if (c < n && r < rows) {
      __align__(16) __half hv[8];
      __align__(16) __half lv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { hv[e] = sh[gl_off(rv + e, cc)]; lv[e] = sl[gl_off(rv + e, cc)]; }
      if (r + 7 < rows) {                        // (ld_t % 8 == 0 and r % 8 == 0: 16-byte aligned)
        *reinterpret_cast<uint4*>(th + (long long)c * ldt + r) = *reinterpret_cast<const uint4*>(hv);
        *reinterpret_cast<uint4*>(tl + (long long)c * ldt + r) = *reinterpret_cast<const uint4*>(lv);
      } else {
        for (int e = 0; e < 8 && r + e < rows; ++e) {
          th[(long long)c * ldt + r + e] = hv[e];
          tl[(long long)c * ldt + r + e] = lv[e];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
planes_glue_kernel(const GlueArgs a) {
  __shared__ __align__(16) __half sh[GL_TILE * GL_TILE];
  __shared__ __align__(16) __half sl[GL_TILE * GL_TILE];
  __shared__ float scol[GL_TILE], scol2[GL_TILE];
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.y * GL_TILE;
  const int c0 = blockIdx.x * GL_TILE;
  if (tid < GL_TILE) { scol[tid] = 0.f; scol2[tid] = 0.f; }
  __syncthreads();
  const int cv = (tid & 7) * 8;
  float cs[8], cs2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { cs[e] = 0.f; cs2[e] = 0.f; }
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int rr = (tid >> 3) + 32 * i;
    const long long r = r0 + rr;
    const int c = c0 + cv;
    __align__(16) __half hv[8];
    __align__(16) __half lv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { hv[e] = __float2half_rn(0.f); lv[e] = __float2half_rn(0.f); }
    if (r < a.rows && c < a.n) {                       // n % 8 == 0: a vector is inside or outside as a whole
      if (a.h) {
        *reinterpret_cast<uint4*>(hv) = *reinterpret_cast<const uint4*>(a.h + r * a.ld + c);
        *reinterpret_cast<uint4*>(lv) = *reinterpret_cast<const uint4*>(a.l + r * a.ld + c);
      } else {
        const float4 f0 = *reinterpret_cast<const float4*>(a.src_f32 + r * a.ld_src + c);
        const float4 f1 = *reinterpret_cast<const float4*>(a.src_f32 + r * a.ld_src + c + 4);
        const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f16_split(fv[e], hv[e], lv[e]);
          bad = bad || !(fabsf(fv[e]) <= F16_GUARD);
        }
      }
      if (a.mask_h) {
        __align__(16) __half mv[8];
        *reinterpret_cast<uint4*>(mv) = *reinterpret_cast<const uint4*>(a.mask_h + r * a.ld_mask + c);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!(__half2float(mv[e]) > 0.f)) { hv[e] = __float2half_rn(0.f); lv[e] = __float2half_rn(0.f); }
      }
      if (a.sign < 0.f) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { hv[e] = __hneg(hv[e]); lv[e] = __hneg(lv[e]); }
      }
      if (a.out_h) {
        *reinterpret_cast<uint4*>(a.out_h + r * a.ld_out + c) = *reinterpret_cast<const uint4*>(hv);
        *reinterpret_cast<uint4*>(a.out_l + r * a.ld_out + c) = *reinterpret_cast<const uint4*>(lv);
      }
      if (a.colsum || a.colsum2) {
        float mulv[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (a.colsum2) {
          const float4 m0 = *reinterpret_cast<const float4*>(a.mul + r * a.ld_mul + c);
          const float4 m1 = *reinterpret_cast<const float4*>(a.mul + r * a.ld_mul + c + 4);
          mulv[0] = m0.x; mulv[1] = m0.y; mulv[2] = m0.z; mulv[3] = m0.w;
          mulv[4] = m1.x; mulv[5] = m1.y; mulv[6] = m1.z; mulv[7] = m1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = f16_join(hv[e], lv[e]);
          cs[e] += v;
          cs2[e] = fmaf(v, mulv[e], cs2[e]);
        }
      }
    }
    if (a.t_h) {
      *reinterpret_cast<uint4*>(&sh[gl_off(rr, cv)]) = *reinterpret_cast<const uint4*>(hv);
      *reinterpret_cast<uint4*>(&sl[gl_off(rr, cv)]) = *reinterpret_cast<const uint4*>(lv);
    }
  }
  if (bad && a.overflow_flag) *a.overflow_flag = 1;
  if (a.colsum || a.colsum2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (a.colsum) atomicAdd(&scol[cv + e], cs[e]);
      if (a.colsum2) atomicAdd(&scol2[cv + e], cs2[e]);
    }
  }
  __syncthreads();
  if ((a.colsum || a.colsum2) && tid < GL_TILE && c0 + tid < a.n) {
    if (a.colsum) atomicAdd(a.colsum + c0 + tid, scol[tid]);
    if (a.colsum2) atomicAdd(a.colsum2 + c0 + tid, scol2[tid]);
  }
  if (a.t_h) gl_store_transposed(sh, sl, tid, r0, c0, a.rows, a.n, a.t_h, a.t_l, a.ld_t);
}

inline int launch_planes_glue(const GlueArgs& a, cudaStream_t st) {
  if (a.rows == 0 || a.n == 0) return USF_OK;
  USF_REQUIRE(a.n % 8 == 0, "planes glue: the width must be a multiple of 8");
  USF_REQUIRE(a.h || a.src_f32, "planes glue: no input");
  bool ok = true;
  if (a.h) ok = ok && aligned16(a.h) && aligned16(a.l) && a.ld % 8 == 0;
  else ok = ok && aligned16(a.src_f32) && a.ld_src % 4 == 0;
  if (a.mask_h) ok = ok && aligned16(a.mask_h) && a.ld_mask % 8 == 0;
  if (a.out_h) ok = ok && aligned16(a.out_h) && aligned16(a.out_l) && a.ld_out % 8 == 0;
  if (a.t_h) ok = ok && aligned16(a.t_h) && aligned16(a.t_l) && a.ld_t % 8 == 0;
  if (a.colsum2) ok = ok && a.mul && aligned16(a.mul) && a.ld_mul % 4 == 0;
  USF_REQUIRE(ok, "planes glue: 16-byte aligned planes and pitches that are multiples of 16 bytes");
  dim3 grid((a.n + GL_TILE - 1) / GL_TILE, (unsigned)((a.rows + GL_TILE - 1) / GL_TILE));
  planes_glue_kernel<<<grid, 256, 0, st>>>(a);
  USF_CUDA_OK(cudaGetLastError());
  return USF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// base density backward:  g = d(-log p(z)) / dz  per element (un-normalised: the caller applies 1 / global batch at the end)
//   Laplace: g = sign(z - mu) / s              d/dmu = -g      d/ds = 1/s - |z - mu| / s^2
//   Normal:  g = (z - mu) / s^2                d/dmu = -g      d/ds = 1/s - (z - mu)^2 / s^3
// (distributions.py:199-238 through torch.distributions; the forward value comes from usf_base_logprob)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
base_backward_kernel(const float* __restrict__ z, long long ldz, long long rows, int d, const float* __restrict__ loc,
                     const float* __restrict__ scale, int kind, __half* gh, __half* gl, long long ldg, __half* th, __half* tl,
                     long long ldt, float* __restrict__ dloc, float* __restrict__ dscale) {
  __shared__ __align__(16) __half sh[GL_TILE * GL_TILE];
  __shared__ __align__(16) __half sl[GL_TILE * GL_TILE];
  __shared__ float s_dl[GL_TILE], s_ds[GL_TILE];
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.y * GL_TILE;
  const int c0 = blockIdx.x * GL_TILE;
  if (tid < GL_TILE) { s_dl[tid] = 0.f; s_ds[tid] = 0.f; }
  __syncthreads();
  const int cv = (tid & 7) * 8;
  float a_dl[8], a_ds[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { a_dl[e] = 0.f; a_ds[e] = 0.f; }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int rr = (tid >> 3) + 32 * i;
    const long long r = r0 + rr;
    const int c = c0 + cv;
    __align__(16) __half hv[8];
    __align__(16) __half lv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { hv[e] = __float2half_rn(0.f); lv[e] = __float2half_rn(0.f); }
    if (r < rows && c < d) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float zz = z[r * ldz + c + e] - __ldg(loc + c + e);
        const float s = __ldg(scale + c + e);
        float g, ds;
        if (kind == USF_BASE_LAPLACE) {
          g = (zz > 0.f ? 1.f : zz < 0.f ? -1.f : 0.f) / s;
          ds = 1.f / s - fabsf(zz) / (s * s);
        } else {
          g = zz / (s * s);
          ds = 1.f / s - zz * zz / (s * s * s);
        }
        f16_split(g, hv[e], lv[e]);
        a_dl[e] -= g;
        a_ds[e] += ds;
      }
      *reinterpret_cast<uint4*>(gh + r * ldg + c) = *reinterpret_cast<const uint4*>(hv);
      *reinterpret_cast<uint4*>(gl + r * ldg + c) = *reinterpret_cast<const uint4*>(lv);
    }
    if (th) {
      *reinterpret_cast<uint4*>(&sh[gl_off(rr, cv)]) = *reinterpret_cast<const uint4*>(hv);
      *reinterpret_cast<uint4*>(&sl[gl_off(rr, cv)]) = *reinterpret_cast<const uint4*>(lv);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&s_dl[cv + e], a_dl[e]);
    atomicAdd(&s_ds[cv + e], a_ds[e]);
  }
  __syncthreads();
  if (tid < GL_TILE && c0 + tid < d) {
    if (dloc) atomicAdd(dloc + c0 + tid, s_dl[tid]);
    if (dscale) atomicAdd(dscale + c0 + tid, s_ds[tid]);
  }
  if (th) gl_store_transposed(sh, sl, tid, r0, c0, rows, d, th, tl, ldt);
}

// ---------------------------------------------------------------------------------------------------------------------
// weight-side matrix preparation:  out[i, j] = scale * src[ri(i), cj(j)]   (src read transposed when `transpose`)
//   ri / cj: optional gather indices over the rows / columns of the (possibly transposed) source
// ---------------------------------------------------------------------------------------------------------------------
// 32 x 32 output tile per CTA (256 threads = 32 x 8), staged in shared memory so that the source is read along its
// contiguous dimension (whichever of rows / columns that is under `transpose`) and every output -- the fp32 copy, the
// operand planes and the TRANSPOSED operand planes -- is written along its own contiguous dimension.
__global__ void __launch_bounds__(256)
mat_prep_kernel(const float* __restrict__ src, long long ld_src, int rows, int cols, int transpose,
                const int* __restrict__ row_idx, const int* __restrict__ col_idx, float scale, float* out_f32, long long ld_f32,
                __half* out_h, __half* out_l, long long ld_16, __half* outT_h, __half* outT_l, long long ld_T, int* overflow_flag) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  // load: logical element (r, c) = scale * S[ri(r), cj(c)], S = src or src^T
  if (!transpose) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k, c = c0 + tx;
      if (r < rows && c < cols) {
        const int sr = row_idx ? __ldg(row_idx + r) : r, sc = col_idx ? __ldg(col_idx + c) : c;
        tile[ty + 8 * k][tx] = scale * src[(long long)sr * ld_src + sc];
      }
    }
  } else {       // S[a, b] = src[b, a]: walk the logical ROW index with tx (contiguous in src when no gather is given)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + tx, c = c0 + ty + 8 * k;
      if (r < rows && c < cols) {
        const int sr = row_idx ? __ldg(row_idx + r) : r, sc = col_idx ? __ldg(col_idx + c) : c;
        tile[tx][ty + 8 * k] = scale * src[(long long)sc * ld_src + sr];
      }
    }
  }
  __syncthreads();
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < rows && c < cols) {
      const float v = tile[ty + 8 * k][tx];
      if (out_f32) out_f32[(long long)r * ld_f32 + c] = v;
      if (out_h) {
        __half h, l;
        f16_split(v, h, l);
        out_h[(long long)r * ld_16 + c] = h;
        out_l[(long long)r * ld_16 + c] = l;
        bad = bad || !(fabsf(v) <= F16_GUARD);
      }
    }
  }
  if (outT_h) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, r = r0 + tx;     // transposed planes: row c, column r
      if (r < rows && c < cols) {
        const float v = tile[tx][ty + 8 * k];
        __half h, l;
        f16_split(v, h, l);
        outT_h[(long long)c * ld_T + r] = h;
        outT_l[(long long)c * ld_T + r] = l;
        bad = bad || !(fabsf(v) <= F16_GUARD);
      }
    }
  }
  if (bad && overflow_flag) *overflow_flag = 1;
}

// ---- small vector / matrix products of the bias path (c = -W^-1 b and its derivative) -------------------------------
// out[i] = alpha * sum_k W[r(i), k] v[k]      (one warp per output row)
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ W, long long ld, int n_rows, int K, const int* __restrict__ row_idx,
              const float* __restrict__ v, float alpha, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const float* w = W + (long long)(row_idx ? __ldg(row_idx + warp) : warp) * ld;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(w[k], __ldg(v + k), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[warp] = alpha * s;
}
// out[j] += alpha * sum_i v[i] W[i, j]        (thread = column, blockIdx.y = chunk of 32 rows: coalesced row reads, the
// chunks meet in `out` through atomics)
constexpr int CC_ROWS = 32;
__global__ void __launch_bounds__(256)
colcomb_kernel(const float* __restrict__ W, long long ld, int rows, int cols, const float* __restrict__ v, float alpha,
               float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  const int i0 = blockIdx.y * CC_ROWS, i1 = min(rows, i0 + CC_ROWS);
  float s = 0.f;
#pragma unroll 8
  for (int i = i0; i < i1; ++i) s = fmaf(__ldg(v + i), W[(long long)i * ld + j], s);
  atomicAdd(out + j, alpha * s);
}
// A[i, j] += alpha * u[i] v[j]
__global__ void __launch_bounds__(256)
rank1_kernel(float* __restrict__ A, long long ld, int rows, int cols, const float* __restrict__ u, const float* __restrict__ v,
             float alpha) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    A[(long long)r * ld + c] = fmaf(alpha * __ldg(u + r), __ldg(v + c), A[(long long)r * ld + c]);
  }
}

// out = part(src) * scale (+ coef / diag_src[i, i] on the diagonal):  mode 0 = strict lower, 1 = upper incl. diagonal
__global__ void __launch_bounds__(256)
tri_mask_kernel(const float* __restrict__ src, long long ld_src, int d, int mode, float scale, const float* __restrict__ diag_src,
                long long ld_diag, float coef, float* __restrict__ out, long long ld_out) {
  const long long total = (long long)d * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d), c = (int)(i - (long long)r * d);
    const bool keep = mode == 0 ? (c < r) : (c >= r);
    float v = keep ? scale * src[(long long)r * ld_src + c] : 0.f;
    if (r == c && diag_src && mode == 1) v += coef / diag_src[(long long)r * ld_diag + r];
    out[(long long)r * ld_out + c] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// batched lower-triangular inverse by recursive block doubling
//   T, X: stacks of n_mats matrices [d, d] (row-major, pitch ld, matrix stride `ms` floats); bit m of unit_mask = matrix m
//   has an implicit unit diagonal.  X must be zero-filled before the first kernel (the launcher does).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TB = 64;

__global__ void __launch_bounds__(TB)
tri_diag_batched_kernel(const float* __restrict__ T, float* __restrict__ X, int d, long long ld, long long ms, unsigned unit_mask) {
  __shared__ float sT[TB][TB + 1];
  __shared__ float sX[TB][TB + 1];
  const int b = blockIdx.x, m = blockIdx.y, j = threadIdx.x, base = b * TB;
  const float* Tm = T + (long long)m * ms;
  float* Xm = X + (long long)m * ms;
  const bool unit = (unit_mask >> m) & 1u;
  for (int r = 0; r < TB; ++r) {
    const int gr = base + r, gc = base + j;
    float v = (r == j) ? 1.f : 0.f;
    if (gr < d && gc < d && j <= r) {
      v = Tm[(long long)gr * ld + gc];
      if (r == j && unit) v = 1.f;
    }
    sT[r][j] = v;
  }
  __syncthreads();
  for (int i = 0; i < TB; ++i) {        // thread j: column j of the block inverse by forward substitution
    float s = (i == j) ? 1.f : 0.f;
    if (i > j)
      for (int k = j; k < i; ++k) s = fmaf(-sT[i][k], sX[k][j], s);
    sX[i][j] = (i < j) ? 0.f : s / sT[i][i];
  }
  __syncthreads();
  for (int r = 0; r < TB; ++r) {
    const int gr = base + r, gc = base + j;
    if (gr < d && gc < d) Xm[(long long)gr * ld + gc] = sX[r][j];
  }
}

// One level of the doubling (block size s, a multiple of 64): for every pair (lo = [r0, r0 + s), hi = [r0 + s, r0 + 2 s) n [0, d))
//   phase 0:  Tmp[hi, lo] = T[hi, lo] . X[lo, lo]          phase 1:  X[hi, lo] = - X[hi, hi] . Tmp[hi, lo]
// grid = (tiles over the s columns of lo, tiles over the rows of hi, pair * n_mats + matrix)
__global__ void __launch_bounds__(256)
tri_combine_batched_kernel(const float* __restrict__ T, float* X, float* Tmp, int d, long long ld, long long ms, int s, int phase,
                           int n_mats) {
  __shared__ float sA[TB][TB + 4];   // A tile stored transposed: sA[k][i]
  __shared__ float sB[TB][TB + 4];   // B tile: sB[k][j]
  const int m = blockIdx.z % n_mats, pair = blockIdx.z / n_mats;
  const int r0 = pair * 2 * s;
  const int hi0 = r0 + s;
  if (hi0 >= d) return;
  const int hi1 = min(r0 + 2 * s, d);
  const int row0 = hi0 + blockIdx.y * TB, col0 = r0 + blockIdx.x * TB;
  if (row0 >= hi1) return;
  const float* A = (phase == 0 ? T : X) + (long long)m * ms;      // phase 0: T[hi, lo] ; phase 1: X[hi, hi]
  const float* B = (phase == 0 ? X : Tmp) + (long long)m * ms;    // phase 0: X[lo, lo] ; phase 1: Tmp[hi, lo]
  float* O = (phase == 0 ? Tmp : X) + (long long)m * ms;
  // contraction range: phase 0 over the lo block (X11 is lower triangular: k >= col0), phase 1 over hi (X22 lower: k <= row)
  const int k_begin = phase == 0 ? col0 : hi0;
  const int k_end = phase == 0 ? hi0 : min(row0 + TB, hi1);
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += TB) {
    for (int e = tid; e < TB * TB; e += 256) {
      const int r = e / TB, c = e % TB;
      const int gi = row0 + r, gk = k0 + c;
      sA[c][r] = (gi < hi1 && gk < k_end) ? A[(long long)gi * ld + gk] : 0.f;
      const int gk2 = k0 + r, gj = col0 + c;
      sB[r][c] = (gk2 < k_end && gj < hi0) ? B[(long long)gk2 * ld + gj] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < TB; ++k) {
      float a[4], b[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
    __syncthreads();
  }
  const float sgn = phase == 0 ? 1.f : -1.f;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gi = row0 + ty * 4 + p, gj = col0 + tx * 4 + q;
      if (gi < hi1 && gj < hi0) O[(long long)gi * ld + gj] = sgn * acc[p][q];
    }
}

inline int launch_tri_inverse_batched(const float* T, float* X, float* Tmp, int d, long long ld, long long ms, int n_mats,
                                      unsigned unit_mask, cudaStream_t st) {
  USF_REQUIRE(T && X && Tmp && d > 0 && n_mats > 0 && n_mats <= 32 && T != X && X != Tmp, "bad input");
  USF_REQUIRE(ms >= (long long)d * ld || n_mats == 1, "matrix stride smaller than a matrix");
  USF_CUDA_OK(cudaMemsetAsync(X, 0, sizeof(float) * ((size_t)(n_mats - 1) * ms + (size_t)(d - 1) * ld + d), st));
  const int nb = (d + TB - 1) / TB;
  tri_diag_batched_kernel<<<dim3(nb, n_mats), TB, 0, st>>>(T, X, d, ld, ms, unit_mask);
  for (int s = TB; s < d; s *= 2) {
    const int pairs = (d + 2 * s - 1) / (2 * s);
    const dim3 grid(s / TB, s / TB, pairs * n_mats);
    tri_combine_batched_kernel<<<grid, 256, 0, st>>>(T, X, Tmp, d, ld, ms, s, 0, n_mats);
    tri_combine_batched_kernel<<<grid, 256, 0, st>>>(T, X, Tmp, d, ld, ms, s, 1, n_mats);
  }
  USF_CUDA_OK(cudaGetLastError());
  return USF_OK;
}

}  // namespace usf
