// LFAE flow-warp + occlusion-blend kernels (Generator.deform_input / apply_optical, model/LFAE/generator.py:63-93).
// The bilinear resize of flow / occlusion (F.interpolate, align_corners=False), the grid_sample
// (bilinear, zeros padding, align_corners=True) and the blend are fused; index math follows ATen
// (UpSampleBilinear2d / GridSampler) term by term in fp32 so neighbour indices and weights are bit-exact.
#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {

__device__ __forceinline__ void resize_src(int dst, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  float s = __fsub_rn(__fmul_rn(scale, static_cast<float>(dst) + 0.5f), 0.5f);
  if (s < 0.f) s = 0.f;
  i0 = static_cast<int>(s);
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = __fsub_rn(s, static_cast<float>(i0));
  l0 = __fsub_rn(1.0f, l1);
}

// bilinear sample of a single-channel fp32 plane with precomputed taps (ATen order: h0*(w0*a + w1*b) + h1*(w0*c + w1*d))
__device__ __forceinline__ float resize_tap(const float* __restrict__ plane, int stride_y, int stride_x, int y0, int y1,
                                            int x0, int x1, float ly0, float ly1, float lx0, float lx1) {
  const float a = __ldg(plane + y0 * stride_y + x0 * stride_x), b = __ldg(plane + y0 * stride_y + x1 * stride_x);
  const float c = __ldg(plane + y1 * stride_y + x0 * stride_x), d = __ldg(plane + y1 * stride_y + x1 * stride_x);
  const float top = __fadd_rn(__fmul_rn(lx0, a), __fmul_rn(lx1, b));
  const float bot = __fadd_rn(__fmul_rn(lx0, c), __fmul_rn(lx1, d));
  return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

struct WarpTaps {
  int x0, y0;               // north-west corner
  float nw, ne, sw, se;     // ATen grid_sampler_2d bilinear weights
};

// flow value (gx, gy) in [-1,1] -> taps on an (H, W) image, align_corners=True
__device__ __forceinline__ WarpTaps grid_taps(float gx, float gy, int H, int W) {
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), static_cast<float>(W - 1));
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), static_cast<float>(H - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  WarpTaps t;
  t.x0 = static_cast<int>(fx);
  t.y0 = static_cast<int>(fy);
  const float ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy);   // (x1-ix), (y1-iy)
  const float wx = __fsub_rn(ix, fx), wy = __fsub_rn(iy, fy);
  t.nw = __fmul_rn(ex, ey);
  t.ne = __fmul_rn(wx, ey);
  t.sw = __fmul_rn(ex, wy);
  t.se = __fmul_rn(wx, wy);
  return t;
}

// flow at (Y, X) of the (H, W) target resolution, resized on the fly from (h, w); flow is (h, w, 2) interleaved
__device__ __forceinline__ void flow_occ_at(const float* __restrict__ flow, const float* __restrict__ occ, int h, int w,
                                            int H, int W, int Y, int X, float& gx, float& gy, float& oc) {
  if (h == H && w == W) {
    gx = __ldg(flow + (Y * w + X) * 2);
    gy = __ldg(flow + (Y * w + X) * 2 + 1);
    oc = occ ? __ldg(occ + Y * w + X) : 1.0f;
    return;
  }
  const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  resize_src(Y, sy, h, y0, y1, ly0, ly1);
  resize_src(X, sx, w, x0, x1, lx0, lx1);
  gx = resize_tap(flow, w * 2, 2, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
  gy = resize_tap(flow + 1, w * 2, 2, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
  oc = occ ? resize_tap(occ, w, 1, y0, y1, x0, x1, ly0, ly1, lx0, lx1) : 1.0f;
}

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

__global__ void __launch_bounds__(256) warp_blend_cl_kernel(const __nv_bfloat16* __restrict__ skip,
                                                            const __nv_bfloat16* __restrict__ prev,
                                                            const float* __restrict__ flow,
                                                            const float* __restrict__ occ,
                                                            __nv_bfloat16* __restrict__ out, long long F, int rep,
                                                            int H, int W, int C, int h, int w, int up2) {
  const int vecs = C / 8;
  const long long total = F * H * W * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int v = r % vecs; r /= vecs;
    const int X = r % W; r /= W;
    const int Y = r % H; r /= H;
    const long long f = r;
    float gx, gy, oc;
    flow_occ_at(flow + f * h * w * 2, occ ? occ + f * h * w : nullptr, h, w, H, W, Y, X, gx, gy, oc);
    const WarpTaps t = grid_taps(gx, gy, H, W);
    const __nv_bfloat16* sf = skip + (f / rep) * static_cast<long long>(H) * W * C + v * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float tv[8];
    const bool x0ok = t.x0 >= 0 && t.x0 < W, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool y0ok = t.y0 >= 0 && t.y0 < H, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    if (y0ok && x0ok) {
      ld8(sf + (static_cast<long long>(t.y0) * W + t.x0) * C, tv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += tv[j] * t.nw;
    }
    if (y0ok && x1ok) {
      ld8(sf + (static_cast<long long>(t.y0) * W + t.x0 + 1) * C, tv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += tv[j] * t.ne;
    }
    if (y1ok && x0ok) {
      ld8(sf + (static_cast<long long>(t.y0 + 1) * W + t.x0) * C, tv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += tv[j] * t.sw;
    }
    if (y1ok && x1ok) {
      ld8(sf + (static_cast<long long>(t.y0 + 1) * W + t.x0 + 1) * C, tv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += tv[j] * t.se;
    }
    if (occ) {
      if (prev) {
        ld8(prev + ((f * H + Y) * W + X) * static_cast<long long>(C) + v * 8, tv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = acc[j] * oc + tv[j] * (1.0f - oc);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= oc;
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
    if (!up2) {
      *reinterpret_cast<uint4*>(out + ((f * H + Y) * W + X) * static_cast<long long>(C) + v * 8) = o;
    } else {
      __nv_bfloat16* ob = out + ((f * 2 * H + 2 * Y) * (2 * W) + 2 * X) * static_cast<long long>(C) + v * 8;
      *reinterpret_cast<uint4*>(ob) = o;
      *reinterpret_cast<uint4*>(ob + C) = o;
      *reinterpret_cast<uint4*>(ob + 2ll * W * C) = o;
      *reinterpret_cast<uint4*>(ob + 2ll * W * C + C) = o;
    }
  }
}

__device__ __forceinline__ float sample_plane(const float* __restrict__ p, const WarpTaps& t, int H, int W) {
  const bool x0ok = t.x0 >= 0 && t.x0 < W, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool y0ok = t.y0 >= 0 && t.y0 < H, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  float acc = 0.f;
  if (y0ok && x0ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + t.y0 * W + t.x0), t.nw));
  if (y0ok && x1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + t.y0 * W + t.x0 + 1), t.ne));
  if (y1ok && x0ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + (t.y0 + 1) * W + t.x0), t.sw));
  if (y1ok && x1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + (t.y0 + 1) * W + t.x0 + 1), t.se));
  return acc;
}

__global__ void __launch_bounds__(256) warp_image_kernel(const float* __restrict__ src, const float* __restrict__ dec,
                                                         int dec_stride, const float* __restrict__ flow,
                                                         const float* __restrict__ occ, float* __restrict__ pred,
                                                         float* __restrict__ deformed, long long F, int rep, int H,
                                                         int W, int h, int w) {
  const long long total = F * H * W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = i % W;
    const int Y = (i / W) % H;
    const long long f = i / (static_cast<long long>(H) * W);
    float gx, gy, oc;
    flow_occ_at(flow + f * h * w * 2, occ ? occ + f * h * w : nullptr, h, w, H, W, Y, X, gx, gy, oc);
    const WarpTaps t = grid_taps(gx, gy, H, W);
    const float* sp = src + (f / rep) * 3ll * H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = sample_plane(sp + static_cast<long long>(c) * H * W, t, H, W);
      const long long o = (f * 3 + c) * static_cast<long long>(H) * W + static_cast<long long>(Y) * W + X;
      if (deformed) deformed[o] = d;
      if (pred) {
        float pv = d;
        if (occ) {
          const float dv = __ldg(dec + i * dec_stride + c);
          pv = __fadd_rn(__fmul_rn(d, oc), __fmul_rn(dv, __fsub_rn(1.0f, oc)));
        }
        pred[o] = pv;
      }
    }
  }
}

// Test/diagnostic view of the index math: per output pixel the resized flow (gx, gy), occlusion, the north-west
// tap (x0, y0) and the four bilinear weights -- what "warp indexing bit-exact in fp32" is checked on.
__global__ void __launch_bounds__(256) warp_taps_kernel(const float* __restrict__ flow, const float* __restrict__ occ,
                                                        int* __restrict__ xy, float* __restrict__ wts,
                                                        float* __restrict__ gflow, long long F, int H, int W, int h,
                                                        int w) {
  const long long total = F * H * W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = i % W;
    const int Y = (i / W) % H;
    const long long f = i / (static_cast<long long>(H) * W);
    float gx, gy, oc;
    flow_occ_at(flow + f * h * w * 2, occ ? occ + f * h * w : nullptr, h, w, H, W, Y, X, gx, gy, oc);
    const WarpTaps t = grid_taps(gx, gy, H, W);
    xy[i * 2] = t.x0; xy[i * 2 + 1] = t.y0;
    wts[i * 4] = t.nw; wts[i * 4 + 1] = t.ne; wts[i * 4 + 2] = t.sw; wts[i * 4 + 3] = t.se;
    gflow[i * 3] = gx; gflow[i * 3 + 1] = gy; gflow[i * 3 + 2] = oc;
  }
}

}  // namespace extdm

using namespace extdm;

static inline int grid_of(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<int>(g);
}

extern "C" int extdm_warp_blend_cl(const void* skip, const void* prev, const float* flow, const float* occ, void* out,
                                   long long F, long long Fs, int H, int W, int C, int h, int w, int up2,
                                   void* stream) {
  if (C % 8 || Fs < 1 || F % Fs) {
    extdm_set_error("warp_blend_cl: C % 8 == 0 and F % Fs == 0 required", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  warp_blend_cl_kernel<<<grid_of(F * H * W * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(skip), reinterpret_cast<const __nv_bfloat16*>(prev), flow, occ,
      reinterpret_cast<__nv_bfloat16*>(out), F, static_cast<int>(F / Fs), H, W, C, h, w, up2);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_warp_image(const float* src, const float* dec, int dec_stride, const float* flow,
                                const float* occ, float* prediction, float* deformed, long long F, long long Fs, int H,
                                int W, int h, int w, void* stream) {
  if (Fs < 1 || F % Fs || (prediction && occ && !dec)) {
    extdm_set_error("warp_image: F % Fs == 0 and a decoder output are required", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  warp_image_kernel<<<grid_of(F * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dec, dec_stride, flow, occ, prediction, deformed, F, static_cast<int>(F / Fs), H, W, h, w);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_warp_taps(const float* flow, const float* occ, int* xy, float* weights, float* gflow, long long F,
                               int H, int W, int h, int w, void* stream) {
  if (!flow || !xy || !weights || !gflow) {
    extdm_set_error("warp_taps: null operand", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  warp_taps_kernel<<<grid_of(F * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(flow, occ, xy, weights,
                                                                                            gflow, F, H, W, h, w);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}
