// Shared helpers for the usflows_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/usflows_b200.h"

namespace usf {

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local message, returned through usf_last_error())
// ---------------------------------------------------------------------------------------------
extern thread_local char g_err[512];

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define USF_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      snprintf(usf::g_err, sizeof(usf::g_err), "%s failed: %s (%s:%d)", #expr,                    \
               cudaGetErrorString(_e), __FILE__, __LINE__);                                       \
      return USF_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define USF_REQUIRE(cond, msg)                                                                    \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      snprintf(usf::g_err, sizeof(usf::g_err), "invalid argument: %s [%s] (%s:%d)", msg, #cond,   \
               __FILE__, __LINE__);                                                               \
      return USF_ERR_INVALID;                                                                     \
    }                                                                                             \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device
// Function attributes (the opt-in dynamic shared memory size) are PER DEVICE: a process that drives several GPUs (the
// row-shard driver usflows_b200.parallel) has to set them once on each.  Index of the current device for the
// per-device "already set" tables of the launchers.
constexpr int MAX_DEVICES = 64;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 0;
  return dev;
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: the kernels of one evaluation form a chain on one stream; with the attribute set, the
// next kernel's CTAs are scheduled while the current kernel drains (launch latency and the next kernel's prologue --
// barrier init, TMEM allocation, tensor-map prefetch -- hide under the tail) and block in `griddep_wait()` until the
// previous grid has completed and its memory is visible.  Every kernel launched this way executes griddep_wait() before
// its first access to global memory.
// ---------------------------------------------------------------------------------------------
extern int g_use_pdl;
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// fused epilogue shared by the tcgen05 and the SIMT contraction kernels
//   v = acc (+ bias[n]) ; relu ; v = resid[m,n] + sign * v ; v *= colscale[n] ; v -= postsub[n]
//   then written to any of: fp32 plane, tf32 hi/lo split planes, bf16 plane.
// ---------------------------------------------------------------------------------------------
struct Epilogue {
  const float* bias;
  const float* resid_hi;
  const float* resid_lo;
  const float* colscale;
  const float* postsub;
  float* out_f32;
  float* out_hi;
  float* out_lo;
  __nv_bfloat16* out_bf16;
  long long ldr, ld_f32, ld_split, ld_bf16;
  float resid_sign;
  int relu;
  int vec_ok;  // every pointer 16-byte aligned and every ld a multiple of 4 (8 for bf16 / fp16)
  // fp16 split planes (engine TC_3XF16):  x = h16 + l16 * 2^-11
  const __half* resid_h16;
  const __half* resid_l16;
  __half* out_h16;
  __half* out_l16;
  long long ldr_16, ld_16;
  int* overflow_flag;  // set to 1 when a value written to the fp16 planes leaves the fp16 range
  int fast_store;      // pair kernel: aligned planes, N % 8 == 0: staged, coalesced 16-byte stores (else generic path)
  int async_store;     // pair kernel: fp16-split planes only -> staged boxes leave through TMA tensor stores
  int atomic_out;      // pair kernel, split-K: partial tiles are ADDED to out_f32 (red.global.add.v4.f32)
};

constexpr float F16_LO_SCALE = 2048.f;          // 2^11: the low plane is stored scaled up so it stays normal
constexpr float F16_LO_UNSCALE = 1.f / 2048.f;
constexpr float F16_GUARD = 65000.f;            // |x| above this cannot be represented in the high plane

// x -> (hi, lo') with x ~= hi + lo' * 2^-11, both fp16
__device__ __forceinline__ void f16_split(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * F16_LO_SCALE);
}
__device__ __forceinline__ float f16_join(__half hi, __half lo) {
  return fmaf(__half2float(lo), F16_LO_UNSCALE, __half2float(hi));
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float epi_value(const Epilogue& ep, float v, long long m, int n) {
  if (ep.bias) v += __ldg(ep.bias + n);
  if (ep.relu) v = fmaxf(v, 0.f);
  if (ep.resid_hi) {
    float r = ep.resid_hi[m * ep.ldr + n];
    if (ep.resid_lo) r += ep.resid_lo[m * ep.ldr + n];
    v = fmaf(ep.resid_sign, v, r);
  }
  if (ep.resid_h16) v = fmaf(ep.resid_sign, v, f16_join(ep.resid_h16[m * ep.ldr_16 + n], ep.resid_l16[m * ep.ldr_16 + n]));
  if (ep.colscale) v *= __ldg(ep.colscale + n);
  if (ep.postsub) v -= __ldg(ep.postsub + n);
  return v;
}

__device__ __forceinline__ void epi_store1(const Epilogue& ep, float v, long long m, int n) {
  if (ep.out_f32) ep.out_f32[m * ep.ld_f32 + n] = v;
  if (ep.out_hi) {
    float hi = tf32_round(v);
    ep.out_hi[m * ep.ld_split + n] = hi;
    ep.out_lo[m * ep.ld_split + n] = tf32_round(v - hi);
  }
  if (ep.out_bf16) ep.out_bf16[m * ep.ld_bf16 + n] = __float2bfloat16_rn(v);
  if (ep.out_h16) {
    __half h, l;
    f16_split(v, h, l);
    ep.out_h16[m * ep.ld_16 + n] = h;
    ep.out_l16[m * ep.ld_16 + n] = l;
    if (!(fabsf(v) <= F16_GUARD) && ep.overflow_flag) *ep.overflow_flag = 1;
  }
}

// arithmetic part of the epilogue on four consecutive columns n..n+3 of row m (n % 4 == 0, all in range, ep.vec_ok)
__device__ __forceinline__ void epi_math4(const Epilogue& ep, float (&v)[4], long long m, int n) {
  if (ep.bias) {
    float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.relu) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (ep.resid_hi) {
    float4 r = *reinterpret_cast<const float4*>(ep.resid_hi + m * ep.ldr + n);
    if (ep.resid_lo) {
      float4 l = *reinterpret_cast<const float4*>(ep.resid_lo + m * ep.ldr + n);
      r.x += l.x; r.y += l.y; r.z += l.z; r.w += l.w;
    }
    v[0] = fmaf(ep.resid_sign, v[0], r.x); v[1] = fmaf(ep.resid_sign, v[1], r.y);
    v[2] = fmaf(ep.resid_sign, v[2], r.z); v[3] = fmaf(ep.resid_sign, v[3], r.w);
  }
  if (ep.resid_h16) {
    const uint2 rh = *reinterpret_cast<const uint2*>(ep.resid_h16 + m * ep.ldr_16 + n);
    const uint2 rl = *reinterpret_cast<const uint2*>(ep.resid_l16 + m * ep.ldr_16 + n);
    const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
    const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 hf = __half22float2(h2[i]), lf = __half22float2(l2[i]);
      v[2 * i] = fmaf(ep.resid_sign, v[2 * i], fmaf(lf.x, F16_LO_UNSCALE, hf.x));
      v[2 * i + 1] = fmaf(ep.resid_sign, v[2 * i + 1], fmaf(lf.y, F16_LO_UNSCALE, hf.y));
    }
  }
  if (ep.colscale) {
    float4 s = __ldg(reinterpret_cast<const float4*>(ep.colscale + n));
    v[0] *= s.x; v[1] *= s.y; v[2] *= s.z; v[3] *= s.w;
  }
  if (ep.postsub) {
    float4 s = __ldg(reinterpret_cast<const float4*>(ep.postsub + n));
    v[0] -= s.x; v[1] -= s.y; v[2] -= s.z; v[3] -= s.w;
  }
}

// epi_math4 followed by the stores to every requested output plane
__device__ __forceinline__ void epi_apply4(const Epilogue& ep, float (&v)[4], long long m, int n) {
  epi_math4(ep, v, m, n);
  if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + m * ep.ld_f32 + n) = make_float4(v[0], v[1], v[2], v[3]);
  if (ep.out_hi) {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i] = tf32_round(v[i]); l[i] = tf32_round(v[i] - h[i]); }
    *reinterpret_cast<float4*>(ep.out_hi + m * ep.ld_split + n) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(ep.out_lo + m * ep.ld_split + n) = make_float4(l[0], l[1], l[2], l[3]);
  }
  if (ep.out_bf16) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(ep.out_bf16 + m * ep.ld_bf16 + n) = u;
  }
  if (ep.out_h16) {
    const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn((v[0] - f0.x) * F16_LO_SCALE, (v[1] - f0.y) * F16_LO_SCALE);
    const __half2 l1 = __floats2half2_rn((v[2] - f1.x) * F16_LO_SCALE, (v[3] - f1.y) * F16_LO_SCALE);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&h0); uh.y = *reinterpret_cast<const uint32_t*>(&h1);
    ul.x = *reinterpret_cast<const uint32_t*>(&l0); ul.y = *reinterpret_cast<const uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(ep.out_h16 + m * ep.ld_16 + n) = uh;
    *reinterpret_cast<uint2*>(ep.out_l16 + m * ep.ld_16 + n) = ul;
    const bool bad = !(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))) <= F16_GUARD);
    if (bad && ep.overflow_flag) *ep.overflow_flag = 1;
  }
}

// NC consecutive columns starting at n0 of row m, with range checks on N
template <int NC>
__device__ __forceinline__ void epi_row_chunk(const Epilogue& ep, float (&v)[NC], long long m, int n0, int N) {
  if (ep.vec_ok) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      int n = n0 + j;
      if (n + 3 < N) {
        float t[4] = {v[j], v[j + 1], v[j + 2], v[j + 3]};
        epi_apply4(ep, t, m, n);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (n + i < N) epi_store1(ep, epi_value(ep, v[j + i], m, n + i), m, n + i);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (n0 + j < N) epi_store1(ep, epi_value(ep, v[j], m, n0 + j), m, n0 + j);
  }
}

int make_epilogue(const usf_linear_args* a, Epilogue* ep);  // validates pointers / strides

}  // namespace usf
