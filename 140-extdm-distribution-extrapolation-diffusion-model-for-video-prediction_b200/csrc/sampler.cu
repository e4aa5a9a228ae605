// DDIM(eta) update with dynamic thresholding, fp32, arithmetic order identical to the reference
// (model/BaseDM_adaptor/Diffusion.py:231-255); the per-sample 0.9-quantile of |x_start| reproduces
// torch.quantile (linear interpolation, rank computed in fp32) through an exact radix select of the two
// neighbouring order statistics instead of a full sort.
#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {

__device__ __forceinline__ float x_start_of(float img, float pred, float c_recip, float c_recipm1) {
  // two roundings then a subtract, never an FMA: matches `a * x_t - b * noise` evaluated op by op
  return __fsub_rn(__fmul_rn(c_recip, img), __fmul_rn(c_recipm1, pred));
}

// One CTA per sample.  |x| >= 0 so the IEEE bit pattern orders like an unsigned integer.
__global__ void __launch_bounds__(1024) ddim_threshold_kernel(const float* __restrict__ img,
                                                              const float* __restrict__ pred, float c_recip,
                                                              float c_recipm1, float q, float* __restrict__ s_out,
                                                              int n) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_k;
  __shared__ unsigned int s_cnt_le, s_min_gt;
  const int b = blockIdx.x;
  const float* xi = img + static_cast<long long>(b) * n;
  const float* xp = pred + static_cast<long long>(b) * n;

  const float pos = q * static_cast<float>(n - 1);       // fp32 rank, as ATen computes it in the input dtype
  const float lo_f = floorf(pos);
  const unsigned int k_lo = static_cast<unsigned int>(lo_f);
  const unsigned int k_hi = static_cast<unsigned int>(ceilf(pos));
  const float wgt = pos - lo_f;

  if (threadIdx.x == 0) { s_prefix = 0u; s_k = k_lo; }
  __syncthreads();
  // 4 passes of 8 bits, most significant first
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int u = __float_as_uint(fabsf(x_start_of(xi[i], xp[i], c_recip, c_recipm1)));
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int k = s_k, acc = 0u;
      int bin = 0;
      for (; bin < 256; ++bin) {
        if (acc + hist[bin] > k) break;
        acc += hist[bin];
      }
      s_k = k - acc;
      s_prefix = prefix | (static_cast<unsigned int>(bin) << shift);
    }
    __syncthreads();
  }
  const unsigned int v_lo = s_prefix;                     // exact k_lo-th order statistic (bit pattern)
  if (threadIdx.x == 0) { s_cnt_le = 0u; s_min_gt = 0xFFFFFFFFu; }
  __syncthreads();
  unsigned int cnt = 0u, mn = 0xFFFFFFFFu;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned int u = __float_as_uint(fabsf(x_start_of(xi[i], xp[i], c_recip, c_recipm1)));
    if (u <= v_lo) ++cnt;
    else mn = min(mn, u);
  }
  atomicAdd(&s_cnt_le, cnt);
  atomicMin(&s_min_gt, mn);
  __syncthreads();
  if (threadIdx.x == 0) {
    const float a = __uint_as_float(v_lo);
    const float bb = (k_hi == k_lo || s_cnt_le >= k_hi + 1u) ? a : __uint_as_float(s_min_gt);
    // at::lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float diff = __fsub_rn(bb, a);
    const float r = wgt < 0.5f ? __fadd_rn(a, __fmul_rn(wgt, diff))
                               : __fsub_rn(bb, __fmul_rn(diff, __fsub_rn(1.0f, wgt)));
    s_out[b] = fmaxf(r, 1.0f);                            // s.clamp_(min=1.)
  }
}

__global__ void __launch_bounds__(256) ddim_update_kernel(const float* __restrict__ img,
                                                          const float* __restrict__ pred,
                                                          const float* __restrict__ noise,
                                                          const float* __restrict__ s, float c_recip, float c_recipm1,
                                                          float sqrt_alpha_next, float c, float sigma,
                                                          float* __restrict__ img_out, float* __restrict__ xs_out,
                                                          int n4_per_sample, long long total4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float sb = __ldg(s + i / n4_per_sample);
    const float4 x = reinterpret_cast<const float4*>(img)[i];
    const float4 e = reinterpret_cast<const float4*>(pred)[i];
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (noise) z = reinterpret_cast<const float4*>(noise)[i];
    float xv[4] = {x.x, x.y, x.z, x.w}, ev[4] = {e.x, e.y, e.z, e.w}, zv[4] = {z.x, z.y, z.z, z.w}, o[4], xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = x_start_of(xv[j], ev[j], c_recip, c_recipm1);
      v = __fdiv_rn(fminf(fmaxf(v, -sb), sb), sb);
      xs[j] = v;
      float r = __fadd_rn(__fmul_rn(v, sqrt_alpha_next), __fmul_rn(c, ev[j]));
      if (noise) r = __fadd_rn(r, __fmul_rn(sigma, zv[j]));
      o[j] = r;
    }
    reinterpret_cast<float4*>(img_out)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (xs_out) reinterpret_cast<float4*>(xs_out)[i] = make_float4(xs[0], xs[1], xs[2], xs[3]);
  }
}

}  // namespace extdm

using namespace extdm;

extern "C" int extdm_ddim_threshold(const float* img, const float* pred, float c_recip, float c_recipm1, float q,
                                    float* s, int B, int n, void* stream) {
  if (n < 2 || B < 1) {
    extdm_set_error("ddim_threshold: need n >= 2", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  ddim_threshold_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(img, pred, c_recip, c_recipm1, q, s, n);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_ddim_update(const float* img, const float* pred, const float* noise, const float* s,
                                 float c_recip, float c_recipm1, float sqrt_alpha_next, float c, float sigma,
                                 float* img_out, float* x_start_out, int B, int n, void* stream) {
  if (n % 4) {
    extdm_set_error("ddim_update: elements per sample must be a multiple of 4", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  const long long total4 = static_cast<long long>(B) * n / 4;
  long long g = (total4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  ddim_update_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, pred, noise, s, c_recip, c_recipm1, sqrt_alpha_next, c, sigma, img_out, x_start_out, n / 4, total4);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}
