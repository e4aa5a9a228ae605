// TrajWarp kernels of the BAIR ('u12') Unet3D variant:
//   reference model/BaseDM_adaptor/DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_u12.py
//     :719-728  ScaledDotProductAttention   softmax(q k^T / sqrt(dk)) v
//     :730-803  MultiHeadAttentionOp        (the four Linear+ReLU run on the tcgen05 GEMM, conv_gemm.cu)
//     :804-827  TrajWarp                    MaxPool3d((1,2,2)) of the noisy-frame features, cross attention of the
//                                           future-frame queries over the conditioning-frame keys, 1x1 fuser
// plus frame-range variants of the channels-last resize / pool used around it.
#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {

__device__ __forceinline__ void tj_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tj_ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void tj_ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void tj_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ float tj_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Flash-style multi-head cross attention, head dim 32.  q: (B, Lq, ldq) bf16, k / v: (B, Lk, ldk); head h uses
// columns [h*32, h*32+32).  One CTA = 64 query rows of one (batch, head); 4 warps x 16 rows; keys stream through
// shared memory in 64-row blocks (cp.async double buffer) with an online softmax in the exp2 domain.
constexpr int kCaDh = 32;
constexpr int kCaPitch = kCaDh + 8;     // bf16 elements per smem row (80 B: conflict-free ldmatrix)

__global__ void __launch_bounds__(128) cross_attention_kernel(const __nv_bfloat16* __restrict__ q,
                                                              const __nv_bfloat16* __restrict__ k,
                                                              const __nv_bfloat16* __restrict__ v,
                                                              __nv_bfloat16* __restrict__ out, int Lq, int Lk, int ldq,
                                                              int ldk, int ldo, float scale_log2e) {
  __shared__ __align__(16) __nv_bfloat16 s_k[2][64][kCaPitch];
  __shared__ __align__(16) __nv_bfloat16 s_v[2][64][kCaPitch];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int lrow = lane & 15, lcol = (lane >> 4) * 8;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const __nv_bfloat16* qb = q + (static_cast<long long>(b) * Lq + q0 + warp * 16) * ldq + h * kCaDh;
  const __nv_bfloat16* kb = k + static_cast<long long>(b) * Lk * ldk + h * kCaDh;
  const __nv_bfloat16* vb = v + static_cast<long long>(b) * Lk * ldk + h * kCaDh;

  auto load_block = [&](int blk, int buf) {
    // 64 rows x 32 dims x {K, V} = 2 x 256 chunks of 16 B over 128 threads
    for (int i = tid; i < 512; i += 128) {
      const int which = i >> 8, r = (i & 255) >> 2, c = (i & 3) * 8;
      const __nv_bfloat16* src = (which ? vb : kb) + static_cast<long long>(blk * 64 + r) * ldk + c;
      tj_cp_async16(which ? &s_v[buf][r][c] : &s_k[buf][r][c], src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_block(0, 0);

  // Q fragments straight from global (row g / g+8 of this warp's 16 rows)
  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const __nv_bfloat16* p0 = qb + static_cast<long long>(g) * ldq + ks * 16 + tg * 2;
    const __nv_bfloat16* p1 = p0 + 8ll * ldq;
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(p0);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(p1);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
  }
  float o[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
  float m0 = -3.0e38f, m1 = -3.0e38f, l0 = 0.f, l1 = 0.f;

  const int nblk = Lk / 64;
  for (int blk = 0; blk < nblk; ++blk) {
    const int buf = blk & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                       // block `blk` landed; everyone is done with buf^1
    if (blk + 1 < nblk) load_block(blk + 1, buf ^ 1);
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      // K rows np*16..+15 (n) x dims 0..31 (k): two ldmatrix.x4 = B fragments of 2 n-tiles x 2 k-steps
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t kf[4];
        tj_ldsm_x4(kf, &s_k[buf][np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
        tj_mma16816(s[2 * np], qa[ks], kf[0], kf[1]);
        tj_mma16816(s[2 * np + 1], qa[ks], kf[2], kf[3]);
      }
    }
    float bm0 = -3.0e38f, bm1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[nt][j] *= scale_log2e;
      bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
      bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
    const float c0 = tj_exp2(m0 - n0), c1 = tj_exp2(m1 - n1);
    m0 = n0; m1 = n1;
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = tj_exp2(s[nt][0] - m0);
      s[nt][1] = tj_exp2(s[nt][1] - m0);
      s[nt][2] = tj_exp2(s[nt][2] - m1);
      s[nt][3] = tj_exp2(s[nt][3] - m1);
      r0 += s[nt][0] + s[nt][1];
      r1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + r0;                                     // per-thread partial row sums (quad-reduced at the end)
    l1 = l1 * c1 + r1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * ps][0], s[2 * ps][1]);
      a[1] = pack_bf16(s[2 * ps][2], s[2 * ps][3]);
      a[2] = pack_bf16(s[2 * ps + 1][0], s[2 * ps + 1][1]);
      a[3] = pack_bf16(s[2 * ps + 1][2], s[2 * ps + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        uint32_t vf[4];
        tj_ldsm_x4_trans(vf, &s_v[buf][ps * 16 + lrow][dp * 16 + lcol]);
        tj_mma16816(o[2 * dp], a, vf[0], vf[1]);
        tj_mma16816(o[2 * dp + 1], a, vf[2], vf[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = out + (static_cast<long long>(b) * Lq + q0 + warp * 16) * ldo + h * kCaDh;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(g) * ldo + dt * 8 + tg * 2) =
        pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
    *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(g + 8) * ldo + dt * 8 + tg * 2) =
        pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------ frame-range helpers
__device__ __forceinline__ void tj_load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void tj_store8(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// MaxPool3d((1,2,2)) on channels-last frames; frame f of group gi lives at x + gi*x_group_stride + fi*H*W*C.
__global__ void __launch_bounds__(256) maxpool2_frames_kernel(const __nv_bfloat16* __restrict__ x,
                                                              __nv_bfloat16* __restrict__ y, int groups, int fpg,
                                                              long long xgs, long long ygs, int H, int W, int C) {
  const int vecs = C / 8, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(groups) * fpg * Ho * Wo * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int vv = r % vecs; r /= vecs;
    const int xo = r % Wo; r /= Wo;
    const int yo = r % Ho; r /= Ho;
    const int fi = r % fpg;
    const long long gi = r / fpg;
    const __nv_bfloat16* p = x + gi * xgs + ((static_cast<long long>(fi) * H + 2 * yo) * W + 2 * xo) * C + vv * 8;
    float a[8], b[8], c[8], d[8], o[8];
    tj_load8(p, a);
    tj_load8(p + C, b);
    tj_load8(p + static_cast<long long>(W) * C, c);
    tj_load8(p + static_cast<long long>(W) * C + C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j]));
    tj_store8(y + gi * ygs + ((static_cast<long long>(fi) * Ho + yo) * Wo + xo) * C + vv * 8, o);
  }
}

__device__ __forceinline__ void tj_bilinear_src(int dst, float scale, int in, int& i0, int& i1, float& l) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = static_cast<int>(s);
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l = s - static_cast<float>(i0);
}

// F.interpolate(bilinear, align_corners=False) of channels-last frames, frame groups as above.
__global__ void __launch_bounds__(256) bilinear_resize_frames_kernel(const __nv_bfloat16* __restrict__ x,
                                                                     __nv_bfloat16* __restrict__ y, int groups,
                                                                     int fpg, long long xgs, long long ygs, int h,
                                                                     int w, int H, int W, int C) {
  const int vecs = C / 8;
  const float sy = static_cast<float>(h) / H, sx = static_cast<float>(w) / W;
  // one (output pixel, 16-byte channel vector) per iteration; 32-bit index arithmetic (the launcher bounds the total):
  // the 64-bit divisions of the first version made this pass ALU bound at 1.6 TB/s
  const unsigned total = static_cast<unsigned>(groups) * fpg * H * W * vecs;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i;
    const int vv = r % vecs; r /= vecs;
    const int X = r % W; r /= W;
    const int Y = r % H; r /= H;
    const int fi = r % fpg;
    const long long gi = r / fpg;
    int y0, y1, x0, x1;
    float ly, lx;
    tj_bilinear_src(Y, sy, h, y0, y1, ly);
    tj_bilinear_src(X, sx, w, x0, x1, lx);
    const __nv_bfloat16* xf = x + gi * xgs + static_cast<long long>(fi) * h * w * C + vv * 8;
    float a[8], b[8], c[8], d[8], o[8];
    tj_load8(xf + (static_cast<long long>(y0) * w + x0) * C, a);
    tj_load8(xf + (static_cast<long long>(y0) * w + x1) * C, b);
    tj_load8(xf + (static_cast<long long>(y1) * w + x0) * C, c);
    tj_load8(xf + (static_cast<long long>(y1) * w + x1) * C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
    tj_store8(y + gi * ygs + ((static_cast<long long>(fi) * H + Y) * W + X) * C + vv * 8, o);
  }
}

// Companion of the polyphase init_conv (extdm_b200/composite.py: compose_upsampled): the 7x7 convolution of the x2
// bilinearly up-sampled TrajWarp features is evaluated on the low-resolution tensor, which needs (a) that tensor
// replicate-padded by 2 (the up-sampling's border clamping) and (b) the border rows / columns of the up-sampled tensor
// (what the zero padding of the convolution replaces outside the image is a repetition of them).
// fp: (F, h + 4, w + 4, C) with the interior written; top / bottom: (F, W + 6, C) = up[0 | H-1, clamp(x' - 3)];
// left / right: (F, H, C) = up[y, 0 | W-1].  One block per frame.
__global__ void __launch_bounds__(256) upsample2_border_kernel(__nv_bfloat16* __restrict__ fp, __nv_bfloat16* __restrict__ top,
                                                               __nv_bfloat16* __restrict__ bottom,
                                                               __nv_bfloat16* __restrict__ left,
                                                               __nv_bfloat16* __restrict__ right, int h, int w, int C) {
  const int vecs = C / 8, hp = h + 4, wp = w + 4, H = 2 * h, W = 2 * w;
  __nv_bfloat16* f = fp + static_cast<long long>(blockIdx.x) * hp * wp * C;
  // (a) replicate padding: every cell outside the interior copies its clamped interior cell
  for (int i = threadIdx.x; i < hp * wp * vecs; i += blockDim.x) {
    const int vv = i % vecs, x = (i / vecs) % wp, y = i / (vecs * wp);
    const int ys = min(max(y, 2), h + 1), xs = min(max(x, 2), w + 1);
    if (ys != y || xs != x)
      *reinterpret_cast<uint4*>(f + (static_cast<long long>(y) * wp + x) * C + vv * 8) =
          *reinterpret_cast<const uint4*>(f + (static_cast<long long>(ys) * wp + xs) * C + vv * 8);
  }
  // (b) border rows / columns of the up-sampled tensor, with the arithmetic of bilinear_resize_frames_kernel
  const int n_tb = W + 6;
  for (int i = threadIdx.x; i < (2 * n_tb + 2 * H) * vecs; i += blockDim.x) {
    const int vv = i % vecs;
    int r = i / vecs;
    __nv_bfloat16* dst;
    int ya, yb, xa, xb;
    float ly, lx;
    if (r < 2 * n_tb) {                                        // a row: Y = 0 or H - 1, X = clamp(x' - 3)
      const bool bot = r >= n_tb;
      if (bot) r -= n_tb;
      tj_bilinear_src(bot ? H - 1 : 0, 0.5f, h, ya, yb, ly);
      tj_bilinear_src(min(max(r - 3, 0), W - 1), 0.5f, w, xa, xb, lx);
      dst = (bot ? bottom : top) + (static_cast<long long>(blockIdx.x) * n_tb + r) * C;
    } else {                                                   // a column: X = 0 or W - 1
      r -= 2 * n_tb;
      const bool rgt = r >= H;
      if (rgt) r -= H;
      tj_bilinear_src(r, 0.5f, h, ya, yb, ly);
      tj_bilinear_src(rgt ? W - 1 : 0, 0.5f, w, xa, xb, lx);
      dst = (rgt ? right : left) + (static_cast<long long>(blockIdx.x) * H + r) * C;
    }
    const __nv_bfloat16* src = f + vv * 8;                     // interior cell (y, x) sits at (y + 2, x + 2)
    float a[8], b[8], c[8], d[8], o[8];
    tj_load8(src + (static_cast<long long>(ya + 2) * wp + xa + 2) * C, a);
    tj_load8(src + (static_cast<long long>(ya + 2) * wp + xb + 2) * C, b);
    tj_load8(src + (static_cast<long long>(yb + 2) * wp + xa + 2) * C, c);
    tj_load8(src + (static_cast<long long>(yb + 2) * wp + xb + 2) * C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
    tj_store8(dst + vv * 8, o);
  }
}

static inline int tj_grid(long long total, int threads) {
  long long gsz = (total + threads - 1) / threads;
  if (gsz < 1) gsz = 1;
  if (gsz > 148 * 16) gsz = 148 * 16;
  return static_cast<int>(gsz);
}

}  // namespace extdm

using namespace extdm;

extern "C" int extdm_cross_attention(const void* q, const void* k, const void* v, void* out, int B, int heads, int dh,
                                     int Lq, int Lk, int ldq, int ldk, int ldo, void* stream) {
  if (dh != kCaDh || Lq % 64 || Lk % 64 || Lq < 64 || Lk < 64 || ldq % 8 || ldk % 8 || ldo % 2 || heads < 1) {
    extdm_set_error("cross_attention: head dim 32, Lq and Lk multiples of 64, 16-byte aligned rows required", __FILE__,
                    __LINE__);
    return EXTDM_ERR_ARG;
  }
  dim3 grid(Lq / 64, heads, B);
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  cross_attention_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), Lq, Lk, ldq, ldk, ldo,
      scale_log2e);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_maxpool2_frames_cl(const void* x, void* y, int groups, int frames_per_group,
                                        long long x_group_stride, long long y_group_stride, int H, int W, int C,
                                        void* stream) {
  if (C % 8 || H % 2 || W % 2 || x_group_stride % 8 || y_group_stride % 8) {
    extdm_set_error("maxpool2_frames_cl: C % 8, even H and W, 16-byte aligned group strides", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  const long long total = static_cast<long long>(groups) * frames_per_group * (H / 2) * (W / 2) * (C / 8);
  maxpool2_frames_kernel<<<tj_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), groups, frames_per_group,
      x_group_stride, y_group_stride, H, W, C);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_bilinear_resize_frames_cl(const void* x, void* y, int groups, int frames_per_group,
                                               long long x_group_stride, long long y_group_stride, int h, int w, int H,
                                               int W, int C, void* stream) {
  if (C % 8 || x_group_stride % 8 || y_group_stride % 8) {
    extdm_set_error("bilinear_resize_frames_cl: C % 8, 16-byte aligned group strides", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  const long long total = static_cast<long long>(groups) * frames_per_group * H * W * (C / 8);
  if (total >= (1ll << 31)) {
    extdm_set_error("bilinear_resize_frames_cl: more than 2^31 output vectors", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  bilinear_resize_frames_kernel<<<tj_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), groups, frames_per_group,
      x_group_stride, y_group_stride, h, w, H, W, C);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_upsample2_border(void* fpad, void* top, void* bottom, void* left, void* right, long long F, int h, int w,
                                      int C, void* stream) {
  if (!fpad || !top || !bottom || !left || !right || F < 1 || h < 2 || w < 2 || C % 8) {
    extdm_set_error("upsample2_border: bad arguments", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  upsample2_border_kernel<<<static_cast<int>(F), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(fpad), reinterpret_cast<__nv_bfloat16*>(top), reinterpret_cast<__nv_bfloat16*>(bottom),
      reinterpret_cast<__nv_bfloat16*>(left), reinterpret_cast<__nv_bfloat16*>(right), h, w, C);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}
