// Weight-preparation kernels: run once per weight version (the reference redoes this work on every call,
// transforms.py:1264-1293, 1457-1476).  fp32 throughout; reductions accumulate in fp64.
#pragma once
#include "common.cuh"

namespace usf {

// L = tril(L_raw,-1) + I ; U = triu(U_raw) (optionally transposed)
__global__ void __launch_bounds__(256)
lu_assemble_kernel(const float* __restrict__ L_raw, const float* __restrict__ U_raw, int d, long long ld_raw,
                   float* __restrict__ L, float* __restrict__ U, long long ld_out, int transpose_u) {
  const long long total = (long long)d * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d), c = (int)(i % d);
    if (L) L[(long long)r * ld_out + c] = c < r ? L_raw[(long long)r * ld_raw + c] : (c == r ? 1.f : 0.f);
    if (U) {
      const float u = c >= r ? U_raw[(long long)r * ld_raw + c] : 0.f;
      if (transpose_u) U[(long long)c * ld_out + r] = u;
      else U[(long long)r * ld_out + c] = u;
    }
  }
}

// out[0] = sum log|v[i*stride]| (fp64 accumulation), out[1] = #zeros ; single block
__global__ void __launch_bounds__(1024)
logabs_kernel(const float* __restrict__ v, long long n, long long stride, float* __restrict__ out) {
  __shared__ double ssum[32];
  __shared__ int szero[32];
  double acc = 0.0;
  int zeros = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i * stride];
    if (x == 0.f) ++zeros;
    acc += (double)logf(fabsf(x));
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
  }
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = acc; szero[threadIdx.x >> 5] = zeros; }
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? ssum[threadIdx.x] : 0.0;
    zeros = threadIdx.x < (blockDim.x >> 5) ? szero[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    }
    if (threadIdx.x == 0) { out[0] = (float)acc; out[1] = (float)zeros; }
  }
}

__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int rows, int cols, long long ld_in, float* __restrict__ out, long long ld_out) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = by + i, c = bx + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(long long)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int r = bx + i, c = by + tx;  // out is [cols, rows]
    if (r < cols && c < rows) out[(long long)r * ld_out + c] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256)
scale_rows_cols_kernel(const float* __restrict__ in, int rows, int cols, long long ld_in, const float* __restrict__ rowf,
                       const float* __restrict__ colf, float* __restrict__ out, long long ld_out) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    float v = in[(long long)r * ld_in + c];
    if (rowf) v *= rowf[r];
    if (colf) v *= colf[c];
    out[(long long)r * ld_out + c] = v;
  }
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ in, long long rows, int cols, long long ld_in, float* __restrict__ hi,
                  float* __restrict__ lo, long long ld_out) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = in[r * ld_in + c];
    const float h = tf32_round(v);
    hi[r * ld_out + c] = h;
    if (lo) lo[r * ld_out + c] = tf32_round(v - h);
  }
}

__global__ void __launch_bounds__(256)
to_bf16_kernel(const float* __restrict__ in, long long rows, int cols, long long ld_in, __nv_bfloat16* __restrict__ out,
               long long ld_out) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[r * ld_out + c] = __float2bfloat16_rn(in[r * ld_in + c]);
  }
}

__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ in, long long rows, int cols, long long ld_in, __half* __restrict__ hi,
                 __half* __restrict__ lo, long long ld_out, int* __restrict__ overflow_flag) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = in[r * ld_in + c];
    __half h, l;
    f16_split(v, h, l);
    hi[r * ld_out + c] = h;
    lo[r * ld_out + c] = l;
    if (!(fabsf(v) <= F16_GUARD) && overflow_flag) *overflow_flag = 1;
  }
}

__global__ void __launch_bounds__(256)
softplus_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = in[i];
    out[i] = x > 20.f ? x : log1pf(expf(x));  // torch softplus (beta=1, threshold=20)
  }
}

// ---- Householder: W <- W - (2 / v.v) (W v) v^T -------------------------------------------------------
// step 1: work[r] = (W[r,:] . v) * 2 / (v.v)   (one warp per row)
__global__ void __launch_bounds__(256)
householder_matvec_kernel(const float* __restrict__ W, int d, long long ld, const float* __restrict__ v, float* __restrict__ work) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= d) return;
  float dot = 0.f, vv = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float x = v[j];
    dot = fmaf(W[(long long)r * ld + j], x, dot);
    vv = fmaf(x, x, vv);
  }
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    vv += __shfl_xor_sync(0xffffffffu, vv, o);
  }
  if (lane == 0) work[r] = 2.f * dot / vv;
}
// step 2: W[r,c] -= work[r] * v[c]
__global__ void __launch_bounds__(256)
householder_rank1_kernel(float* __restrict__ W, int d, long long ld, const float* __restrict__ v, const float* __restrict__ work) {
  const long long total = (long long)d * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d), c = (int)(i % d);
    W[(long long)r * ld + c] = fmaf(-work[r], v[c], W[(long long)r * ld + c]);
  }
}

// ---- triangular inverse (lower triangular, optional unit diagonal) -----------------------------------
// Blocked forward substitution with NB = 64:  X[I,J] = Tinv_II (delta_IJ I - sum_{J<=K<I} T[I,K] X[K,J]).
// Kernel 1 inverts the 64x64 diagonal blocks in shared memory (one CTA per block);
// kernel 2 gives every CTA one 64-column panel J of X and sweeps block rows I = J .. nb-1.
// An upper-triangular inverse is taken as the transpose of the lower-triangular inverse of T^T (host side).
constexpr int TI_NB = 64;

// Dinv[b] (64x64, row-major, ld 64) = inverse of the b-th diagonal block of T; blocks past d are identity-padded
__global__ void __launch_bounds__(TI_NB)
tri_diag_inverse_kernel(const float* __restrict__ T, int d, long long ldt, int unit_diag, float* __restrict__ Dinv) {
  __shared__ float sT[TI_NB][TI_NB + 1];
  __shared__ float sX[TI_NB][TI_NB + 1];
  const int b = blockIdx.x, j = threadIdx.x, base = b * TI_NB;
  for (int r = 0; r < TI_NB; ++r) {
    const int gr = base + r, gc = base + j;
    float v = (r == j) ? 1.f : 0.f;
    if (gr < d && gc < d && j <= r) {
      v = T[(long long)gr * ldt + gc];
      if (r == j && unit_diag) v = 1.f;
    }
    sT[r][j] = v;
  }
  __syncthreads();
  // thread j solves T x = e_j (column j of the inverse), forward substitution
  for (int i = 0; i < TI_NB; ++i) {
    float s = (i == j) ? 1.f : 0.f;
    if (i > j) {
      for (int k = j; k < i; ++k) s = fmaf(-sT[i][k], sX[k][j], s);
    }
    sX[i][j] = (i < j) ? 0.f : s / sT[i][i];
  }
  __syncthreads();
  for (int r = 0; r < TI_NB; ++r) Dinv[((long long)b * TI_NB + r) * TI_NB + j] = sX[r][j];
}

// 256 threads, each a 4x4 micro tile of the 64x64 accumulator
__global__ void __launch_bounds__(256)
tri_panel_sweep_kernel(const float* __restrict__ T, int d, long long ldt, const float* __restrict__ Dinv,
                       float* X, long long ldx) {
  __shared__ float sA[TI_NB][TI_NB + 4];   // T[I,K] tile, stored transposed: sA[k][i]
  __shared__ float sB[TI_NB][TI_NB + 4];   // X[K,J] tile: sB[k][j]
  const int nb = (d + TI_NB - 1) / TI_NB;
  const int J = blockIdx.x;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;

  for (int I = J; I < nb; ++I) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int K = J; K < I; ++K) {
      for (int e = tid; e < TI_NB * TI_NB; e += 256) {
        const int r = e / TI_NB, c = e % TI_NB;
        const int gi = I * TI_NB + r, gk = K * TI_NB + c;
        sA[c][r] = (gi < d && gk < d) ? T[(long long)gi * ldt + gk] : 0.f;
        const int gk2 = K * TI_NB + r, gj = J * TI_NB + c;
        sB[r][c] = (gk2 < d && gj < d) ? X[(long long)gk2 * ldx + gj] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < TI_NB; ++k) {
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
    // R = delta_IJ I - acc  -> sB ; Dinv_I -> sA (transposed) ; X[I,J] = Dinv_I . R
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = ty * 4 + p, c = tx * 4 + q;
        sB[r][c] = ((I == J && r == c) ? 1.f : 0.f) - acc[p][q];
      }
    for (int e = tid; e < TI_NB * TI_NB; e += 256) {
      const int r = e / TI_NB, c = e % TI_NB;
      sA[c][r] = Dinv[((long long)I * TI_NB + r) * TI_NB + c];
    }
    __syncthreads();
    float out[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) out[a][b] = 0.f;
#pragma unroll 8
    for (int k = 0; k < TI_NB; ++k) {
      float a[4], b[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) out[p][q] = fmaf(a[p], b[q], out[p][q]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = I * TI_NB + ty * 4 + p, gj = J * TI_NB + tx * 4 + q;
        if (gi < d && gj < d) X[(long long)gi * ldx + gj] = out[p][q];
      }
    __threadfence_block();
    __syncthreads();  // X[I,J] visible to the whole CTA before it is re-read as a K block
  }
  // blocks above the diagonal panel start are zero
  for (int I = 0; I < J; ++I)
    for (int e = tid; e < TI_NB * TI_NB; e += 256) {
      const int gi = I * TI_NB + e / TI_NB, gj = J * TI_NB + e % TI_NB;
      if (gi < d && gj < d) X[(long long)gi * ldx + gj] = 0.f;
    }
}

// ---- fp64 product for weight composition: C[M,N] = A[M,K] . B[K,N], all row-major double -------------
// Merging neighbouring affine layers (W2 . W1, W . diag, W . c) happens once per weight version; doing it in
// fp64 keeps the merged operator closer to the exact product than the reference's chained fp32 layers.
// 64x64x16 tiles, 256 threads, 4x4 micro tile.
__global__ void __launch_bounds__(256)
matmul_f64_kernel(const double* __restrict__ A, long long lda, const double* __restrict__ B, long long ldb,
                  double* __restrict__ C, long long ldc, int M, int N, int K) {
  __shared__ double sA[16][64 + 2];
  __shared__ double sB[16][64 + 2];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = tid; e < 64 * 16; e += 256) {
      const int r = e / 16, c = e % 16;         // A tile: 64 rows x 16 k
      const int gm = m0 + r, gk = k0 + c;
      sA[c][r] = (gm < M && gk < K) ? A[(long long)gm * lda + gk] : 0.0;
      const int kr = e / 64, nc = e % 64;       // B tile: 16 k x 64 cols
      const int gk2 = k0 + kr, gn = n0 + nc;
      sB[kr][nc] = (gk2 < K && gn < N) ? B[(long long)gk2 * ldb + gn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) C[(long long)gm * ldc + gn] = acc[i][j];
    }
}

// The same product on the fp64 tensor-core path: mma.sync.m8n8k4.f64 (tcgen05 has no fp64 kind; DMMA is the fp64 MMA that
// sm_100a offers).  The SIMT kernel above reaches about 4.5 TFLOP/s; composing the nine 3072 x 3072 operators of the C5
// stack (58 GFLOP each) took 13 ms a product, most of the 100 ms a weight update cost the inference path.
// BM x BN x 16 tiles; warps hold WM x WN; A is staged k-major so that both fragments are read with stride LD = B* + 4
// doubles (LD mod 16 = 4: the 16 lanes of a half warp hit 16 distinct 8-byte bank pairs); next tile prefetched into
// registers while the current one is multiplied.
template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
matmul_f64_mma_kernel(const double* __restrict__ A, long long lda, const double* __restrict__ B, long long ldb,
                      double* __restrict__ C, long long ldc, int M, int N, int K, int tri) {
  constexpr int BK = 16, T = (BM / WM) * (BN / WN) * 32, LDA = BM + 4, LDB = BN + 4;
  constexpr int NA = BM * BK / T, NB = BN * BK / T, MT = WM / 8, NT = WN / 8;
  static_assert(BM * BK % T == 0 && BN * BK % T == 0 && LDA % 16 == 4 && LDB % 16 == 4, "tile shape");
  __shared__ double sA[BK * LDA];
  __shared__ double sB[BK * LDB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp / (BN / WN)) * WM, wn = (warp % (BN / WN)) * WN;
  // lower x upper: a tile's k range grows with min(row, column), so the heaviest tiles sit at the bottom right -- walk the
  // grid backwards there, heaviest first (the ncu capture of the forward order: 49% DMMA-active over the launch against
  // 73% while a CTA is resident, the tail of the light-first order); upper x lower is heaviest at the top left already
  const int by = tri == USF_TRI_LOWER_UPPER ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int bx = tri == USF_TRI_LOWER_UPPER ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int m0 = by * BM, n0 = bx * BN;
  double acc[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[NA], rb[NB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {              // A tile: BM rows x 16 k, 16 consecutive threads read one 128-byte row piece
      const int e = tid + i * T, r = e >> 4, c = e & 15;
      const int gm = m0 + r, gk = k0 + c;
      ra[i] = (gm < M && gk < K) ? __ldg(A + (long long)gm * lda + gk) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {              // B tile: 16 k x BN columns
      const int e = tid + i * T, kr = e / BN, nc = e % BN;
      const int gk = k0 + kr, gn = n0 + nc;
      rb[i] = (gk < K && gn < N) ? __ldg(B + (long long)gk * ldb + gn) : 0.0;
    }
  };
  // triangular factors (stored dense, zeros included): a tile of lower x upper only has terms k < min(row, column) + 1,
  // one of upper x lower only k >= max(row, column) -- a third of the work on average
  int kb = 0, ke = K;
  if (tri == USF_TRI_LOWER_UPPER) ke = min(K, min(m0 + BM, n0 + BN));
  else if (tri == USF_TRI_UPPER_LOWER) kb = (max(m0, n0) / BK) * BK;
  fetch(kb);
  for (int k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int i = 0; i < NA; ++i) { const int e = tid + i * T; sA[(e & 15) * LDA + (e >> 4)] = ra[i]; }
#pragma unroll
    for (int i = 0; i < NB; ++i) { const int e = tid + i * T; sB[(e / BN) * LDB + (e % BN)] = rb[i]; }
    __syncthreads();
    if (k0 + BK < ke) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = sA[(kk + t) * LDA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = sB[(kk + t) * LDB + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int gm = m0 + wm + i * 8 + g, gn = n0 + wn + j * 8 + 2 * t;
      if (gm < M) {
        if (gn < N) C[(long long)gm * ldc + gn] = acc[i][j][0];
        if (gn + 1 < N) C[(long long)gm * ldc + gn + 1] = acc[i][j][1];
      }
    }
}


}  // namespace usf
