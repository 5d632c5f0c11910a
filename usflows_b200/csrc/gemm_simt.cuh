// fp32 CUDA-core contraction: out = epilogue(A[M,K] . W[N,K]^T), any shape / alignment.
// Used for tiny or unaligned problems (d = 2 ... ), for weight preparation with transposed operands,
// and as the fp32 cross-check of the tcgen05 engines.  64x64x16 tiles, 256 threads, 4x4 per thread,
// register-staged double buffering.
#pragma once
#include "common.cuh"

namespace usf {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <bool TRANS_W>
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ A_lo, long long lda,
                 const float* __restrict__ W, const float* __restrict__ W_lo, long long ldw,
                 long long M, int N, int K, Epilogue ep) {
  __shared__ float sA[2][SG_BK][SG_BM + 4];
  __shared__ float sW[2][SG_BK][SG_BN + 4];

  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each a 4x4 micro tile

  // loader mapping: A tile 64 rows x 16 k -> each thread 4 consecutive k of one row
  const int la_row = tid / 4, la_k = (tid % 4) * 4;
  // W tile: [N,K] layout -> same mapping; [K,N] layout -> each thread 4 consecutive n of one k
  const int lw_row = TRANS_W ? tid / 16 : tid / 4;        // k index (trans) or n index
  const int lw_col = TRANS_W ? (tid % 16) * 4 : (tid % 4) * 4;

  float ra[4], rw[4];
  auto load_tiles = [&](int k0) {
    const long long gm = m0 + la_row;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gk = k0 + la_k + i;
      float v = 0.f;
      if (gm < M && gk < K) {
        v = A[gm * lda + gk];
        if (A_lo) v += A_lo[gm * lda + gk];
      }
      ra[i] = v;
    }
    if (TRANS_W) {
      const int gk = k0 + lw_row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gn = n0 + lw_col + i;
        float v = 0.f;
        if (gn < N && gk < K) {
          v = W[(long long)gk * ldw + gn];
          if (W_lo) v += W_lo[(long long)gk * ldw + gn];
        }
        rw[i] = v;
      }
    } else {
      const int gn = n0 + lw_row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gk = k0 + lw_col + i;
        float v = 0.f;
        if (gn < N && gk < K) {
          v = W[(long long)gn * ldw + gk];
          if (W_lo) v += W_lo[(long long)gn * ldw + gk];
        }
        rw[i] = v;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sA[buf][la_k + i][la_row] = ra[i];
    if (TRANS_W) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sW[buf][lw_row][lw_col + i] = rw[i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) sW[buf][lw_col + i][lw_row] = rw[i];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + SG_BK - 1) / SG_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles((kb + 1) * SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[4], w[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(&sW[buf][k][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kb + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m < M) epi_row_chunk<4>(ep, acc[i], m, n0 + tx * 4, N);
  }
}

// Narrow problems (N, K <= 16: the 16 x 16 1x1-convolution affine layers of the MNIST image flow, transforms.py:904-962,
// over N*H*W channels-last rows): HBM-bound -- 64 B in, 64 B out per row -- and the 64 x 64 tiles above move them at
// 1 TB/s.  Here a thread owns a row: the zero-padded 16 x 16 weight sits in shared memory and is read as broadcast 128-bit
// loads (every lane the same address), the row and its 16 sums stay in registers (~60 registers: full occupancy, 2 KB of
// distinct loads in flight per warp -- a first version with four threads per row and the weight in registers kept only
// 512 B per warp in flight and stopped at 1.5 TB/s).  Same summation order over k as the tiled kernel.
constexpr int NR_K = 16;
__global__ void __launch_bounds__(256)
gemm_rows_narrow_kernel(const float* __restrict__ A, const float* __restrict__ A_lo, long long lda,
                        const float* __restrict__ W, const float* __restrict__ W_lo, long long ldw,
                        long long M, int N, int K, int vec_a, Epilogue ep) {
  __shared__ __align__(16) float sW[NR_K][NR_K];
  {
    const int n = threadIdx.x >> 4, k = threadIdx.x & 15;
    float v = 0.f;
    if (n < N && k < K) {
      v = W[(long long)n * ldw + k];
      if (W_lo) v += W_lo[(long long)n * ldw + k];
    }
    sW[n][k] = v;
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += stride) {
    float a[NR_K];
    const float* ar = A + r * lda;
    if (vec_a) {                          // K % 4 == 0, 16-byte aligned rows
#pragma unroll
      for (int k = 0; k < NR_K; k += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
          v = *reinterpret_cast<const float4*>(ar + k);
          if (A_lo) {
            const float4 l = *reinterpret_cast<const float4*>(A_lo + r * lda + k);
            v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
          }
        }
        a[k] = v.x; a[k + 1] = v.y; a[k + 2] = v.z; a[k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NR_K; ++k) {
        float v = 0.f;
        if (k < K) {
          v = ar[k];
          if (A_lo) v += A_lo[r * lda + k];
        }
        a[k] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < NR_K; j += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < NR_K; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&sW[j + i][k]);
          acc[i] = fmaf(a[k], w.x, acc[i]);
          acc[i] = fmaf(a[k + 1], w.y, acc[i]);
          acc[i] = fmaf(a[k + 2], w.z, acc[i]);
          acc[i] = fmaf(a[k + 3], w.w, acc[i]);
        }
      if (j < N) epi_row_chunk<4>(ep, acc, r, j, N);
    }
  }
}

inline int launch_gemm_simt(const usf_linear_args* a, const Epilogue& ep, cudaStream_t st) {
  if (a->M == 0 || a->N == 0) return USF_OK;
  if (!a->trans_w && a->N <= 16 && a->K <= NR_K && a->M >= 4096) {
    const bool vec_a = a->K % 4 == 0 && a->lda % 4 == 0 && aligned16(a->a) && (!a->a_lo || aligned16(a->a_lo));
    const long long blocks = (a->M + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
    gemm_rows_narrow_kernel<<<grid, 256, 0, st>>>((const float*)a->a, (const float*)a->a_lo, a->lda, (const float*)a->w,
                                                  (const float*)a->w_lo, a->ldw, a->M, a->N, a->K, vec_a ? 1 : 0, ep);
    USF_CUDA_OK(cudaGetLastError());
    return USF_OK;
  }
  USF_REQUIRE((a->M + SG_BM - 1) / SG_BM <= 0x7fffffffLL && (a->N + SG_BN - 1) / SG_BN <= 65535,
              "SIMT engine: problem too large for one launch");
  dim3 grid((unsigned)((a->M + SG_BM - 1) / SG_BM), (a->N + SG_BN - 1) / SG_BN);
  if (a->trans_w)
    gemm_simt_kernel<true><<<grid, SG_THREADS, 0, st>>>(
        (const float*)a->a, (const float*)a->a_lo, a->lda, (const float*)a->w, (const float*)a->w_lo, a->ldw,
        a->M, a->N, a->K, ep);
  else
    gemm_simt_kernel<false><<<grid, SG_THREADS, 0, st>>>(
        (const float*)a->a, (const float*)a->a_lo, a->lda, (const float*)a->w, (const float*)a->w_lo, a->ldw,
        a->M, a->N, a->K, ep);
  USF_CUDA_OK(cudaGetLastError());
  return USF_OK;
}

}  // namespace usf
