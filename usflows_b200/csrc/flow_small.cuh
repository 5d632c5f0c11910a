// Whole-flow kernel for tiny event sizes (d <= 8, conditioner width <= 64): one launch evaluates the complete layer
// stack of `Flow.log_prob` / `Flow.sample` (reference flows.py:234-245 / 258-265) for every row.
//
// At d = 2, H = 32 (BASELINE config C1) a layer is far below one MMA tile and the per-layer route costs ~100 launches
// and an HBM round trip of the activations per layer.  Here a thread owns R rows: the row and the conditioner's hidden
// activations live in registers, all weights of the stack (C1: 50 KB) live in shared memory and are read as broadcast
// 128-bit loads shared by the R rows, so the kernel is bound by the FP32 FMA issue rate (C1: 11.1 k FMA per row), not
// by HBM (12 B per row).
//
// Program = a list of ops over x[D] (d padded to D in {2, 4, 8}; padding rows/columns are identity / zero):
//   AFF       x <- W x + c                                  (dense affine layers incl. folded scale / permutation)
//   COUPLING  x <- x + W_last relu( ... relu(W_0 x + b_0) ...) + b_last
//             with the mask folded into W_0 (input side) and sign * (1 - mask) into W_last / b_last
//             (exact: multiplications by 0, 1, -1), hidden widths padded to H with zero rows / columns.
#pragma once
#include "common.cuh"

namespace usf {

constexpr int FS_OP_AFF = 0, FS_OP_COUPLING = 1;
constexpr int FS_THREADS = 128;

// y[i] = b[i] + sum_k W[i][k] v[k]  for NOUT outputs and NIN inputs, weights row-major in shared memory (row pitch NIN,
// NIN a multiple of 2), for R rows at once
template <int NOUT, int NIN, int R, bool RELU>
__device__ __forceinline__ void fs_linear(const float* __restrict__ w, const float* __restrict__ b, const float (&in)[R][NIN],
                                          float (&out)[R][NOUT]) {
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    float acc[R];
    const float bi = b[i];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = bi;
    if (NIN % 4 == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(w + i * NIN);
#pragma unroll
      for (int k = 0; k < NIN / 4; ++k) {
        const float4 t = w4[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r] = fmaf(t.x, in[r][4 * k], acc[r]);
          acc[r] = fmaf(t.y, in[r][4 * k + 1], acc[r]);
          acc[r] = fmaf(t.z, in[r][4 * k + 2], acc[r]);
          acc[r] = fmaf(t.w, in[r][4 * k + 3], acc[r]);
        }
      }
    } else {
      const float2* w2 = reinterpret_cast<const float2*>(w + i * NIN);
#pragma unroll
      for (int k = 0; k < NIN / 2; ++k) {
        const float2 t = w2[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r] = fmaf(t.x, in[r][2 * k], acc[r]);
          acc[r] = fmaf(t.y, in[r][2 * k + 1], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r][i] = RELU ? fmaxf(acc[r], 0.f) : acc[r];
  }
}

// prog[2*op] = type | (number of H x H middle layers << 8), prog[2*op+1] = offset (floats) of the op's weights in blob
template <int D, int H, int R>
__global__ void __launch_bounds__(FS_THREADS)
flow_small_kernel(const float* __restrict__ x, long long ldx, long long rows, int d, const int* __restrict__ prog, int n_ops,
                  const float* __restrict__ blob, int blob_floats, float* __restrict__ out, long long ldo) {
  extern __shared__ float4 fs_smem4[];
  float* sw = reinterpret_cast<float*>(fs_smem4);
  for (int i = threadIdx.x; i < blob_floats / 4; i += blockDim.x)
    fs_smem4[i] = __ldg(reinterpret_cast<const float4*>(blob) + i);
  __syncthreads();

  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long groups = (rows + R - 1) / R;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    float v[R][D];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = g * R + r;
#pragma unroll
      for (int j = 0; j < D; ++j) v[r][j] = (row < rows && j < d) ? __ldg(x + row * ldx + j) : 0.f;
    }
    for (int op = 0; op < n_ops; ++op) {
      const int code = prog[2 * op];
      const float* w = sw + prog[2 * op + 1];
      if ((code & 0xff) == FS_OP_AFF) {
        float y[R][D];
        fs_linear<D, D, R, false>(w, w + D * D, v, y);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int j = 0; j < D; ++j) v[r][j] = y[r][j];
      } else {
        const int n_mid = code >> 8;
        float ha[R][H], hb[R][H];
        fs_linear<H, D, R, true>(w, w + H * D, v, ha);
        w += H * D + H;
        for (int l = 0; l < n_mid; ++l) {
          fs_linear<H, H, R, true>(w, w + H * H, ha, hb);
          w += H * H + H;
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < H; ++j) ha[r][j] = hb[r][j];
        }
        float t[R][D];
        fs_linear<D, H, R, false>(w, w + D * H, ha, t);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int j = 0; j < D; ++j) v[r][j] += t[r][j];
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = g * R + r;
      if (row < rows) {
#pragma unroll
        for (int j = 0; j < D; ++j)
          if (j < d) out[row * ldo + j] = v[r][j];
      }
    }
  }
}

template <int D, int H, int R>
int launch_flow_small_cfg(const float* x, long long ldx, long long rows, int d, const int* prog, int n_ops, const float* blob,
                          int blob_floats, float* out, long long ldo, cudaStream_t st) {
  auto kern = flow_small_kernel<D, H, R>;
  const size_t smem = (size_t)blob_floats * sizeof(float);
  static size_t attr_bytes_dev[MAX_DEVICES] = {0};
  size_t& attr_bytes = attr_bytes_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > attr_bytes) {
    USF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  const long long groups = (rows + R - 1) / R;
  long long blocks = (groups + FS_THREADS - 1) / FS_THREADS;
  const long long cap = (long long)num_sms() * 8;     // grid-stride beyond a few waves: the weight copy is per block
  if (blocks > cap) blocks = cap;
  kern<<<(int)blocks, FS_THREADS, smem, st>>>(x, ldx, rows, d, prog, n_ops, blob, blob_floats, out, ldo);
  USF_CUDA_OK(cudaGetLastError());
  return USF_OK;
}

inline int launch_flow_small(const float* x, long long ldx, long long rows, int d, const int* prog, int n_ops,
                             const float* blob, int blob_floats, int D, int H, float* out, long long ldo, cudaStream_t st) {
  USF_REQUIRE(blob_floats % 4 == 0 && aligned16(blob), "weight blob must be 16-byte aligned and a multiple of 4 floats");
  USF_REQUIRE((size_t)blob_floats * 4 <= 200 * 1024, "weight blob does not fit shared memory");
  USF_REQUIRE(d >= 1 && d <= D, "d exceeds the padded width");
  if (rows == 0) return USF_OK;
#define USF_FS_CASE(DD, HH, RR) \
  if (D == DD && H == HH) return launch_flow_small_cfg<DD, HH, RR>(x, ldx, rows, d, prog, n_ops, blob, blob_floats, out, ldo, st);
  USF_FS_CASE(2, 32, 2)
  USF_FS_CASE(4, 32, 2)
  USF_FS_CASE(8, 32, 1)
  USF_FS_CASE(2, 64, 1)
  USF_FS_CASE(4, 64, 1)
  USF_FS_CASE(8, 64, 1)
#undef USF_FS_CASE
  return fail(USF_ERR_INVALID, "usf_flow_small: unsupported (D, H); built: D in {2,4,8}, H in {32,64}%s%s");
}

}  // namespace usf
