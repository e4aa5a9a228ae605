// Fused small-sequence attention for the ExtDM UNet (sm_100a).
//   * 3-D shifted-window attention: 32- or 64-token windows, 8 heads, dh 16/32 -- window gather with the
//     cyclic roll, zero padding of T, region mask (-100), relative-position bias, rotary, softmax, PV and
//     the scatter back are one kernel (reference: ~25 ATen ops + 2 bmm per layer).
//   * temporal attention: one pixel's T <= 32 frames, 8 heads, rotary + T5-bucket relative bias.
// One CTA = one window / one pixel, one warp per head.  The tiles (<= 64x64x32) are far below a tcgen05
// tile (M = 128 rows per CTA), so the contractions run on warp-level mma.sync m16n8k16 bf16 with the
// scores kept in registers (flash-style: S accumulators are re-used as the A operand of P*V).
#include <stdlib.h>

#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Per-warp attention for one head.  tile: NTOK rows of `pitch` bf16, row = [q | k | v] each heads*DH wide.
// add(i, j) returns the additive logit term (bias + mask) for query i / key j.
// Output O (NTOK x DH, bf16) overwrites this head's q slice.
template <int NTOK, int DH, typename AddFn>
__device__ __forceinline__ void warp_head_attention(__nv_bfloat16* tile, int pitch, int hid, int head, int lane,
                                                    AddFn add) {
  constexpr int NT = NTOK / 8;      // score n-tiles
  constexpr int KS = DH / 16;       // k-steps of Q K^T
  constexpr int DT = DH / 8;        // output n-tiles
  constexpr int PS = NTOK / 16;     // k-steps of P V
  const int g = lane >> 2, tg = lane & 3;
  __nv_bfloat16* Q = tile + head * DH;
  const __nv_bfloat16* K = tile + hid + head * DH;
  const __nv_bfloat16* V = tile + 2 * hid + head * DH;
  const unsigned short* Vs = reinterpret_cast<const unsigned short*>(V);

#pragma unroll 1
  for (int mt = 0; mt < NTOK / 16; ++mt) {
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(Q + r0 * pitch + ks * 16 + tg * 2);
      a[1] = *reinterpret_cast<const uint32_t*>(Q + r1 * pitch + ks * 16 + tg * 2);
      a[2] = *reinterpret_cast<const uint32_t*>(Q + r0 * pitch + ks * 16 + 8 + tg * 2);
      a[3] = *reinterpret_cast<const uint32_t*>(Q + r1 * pitch + ks * 16 + 8 + tg * 2);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const __nv_bfloat16* kr = K + (nt * 8 + g) * pitch + ks * 16 + tg * 2;
        mma_bf16_16816(s[nt], a, *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
      }
    }
    // logits + bias/mask, row max
    float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c0 = nt * 8 + tg * 2;
      s[nt][0] += add(r0, c0);
      s[nt][1] += add(r0, c0 + 1);
      s[nt][2] += add(r1, c0);
      s[nt][3] += add(r1, c0 + 1);
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = __expf(s[nt][0] - m0);
      s[nt][1] = __expf(s[nt][1] - m0);
      s[nt][2] = __expf(s[nt][2] - m1);
      s[nt][3] = __expf(s[nt][3] - m1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;

    float o[DT][4];
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
    for (int ps = 0; ps < PS; ++ps) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * ps][0], s[2 * ps][1]);
      a[1] = pack_bf16(s[2 * ps][2], s[2 * ps][3]);
      a[2] = pack_bf16(s[2 * ps + 1][0], s[2 * ps + 1][1]);
      a[3] = pack_bf16(s[2 * ps + 1][2], s[2 * ps + 1][3]);
      const int k0 = ps * 16 + tg * 2;
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        const int col = dt * 8 + g;
        const uint32_t b0 = static_cast<uint32_t>(Vs[k0 * pitch + col]) |
                            (static_cast<uint32_t>(Vs[(k0 + 1) * pitch + col]) << 16);
        const uint32_t b1 = static_cast<uint32_t>(Vs[(k0 + 8) * pitch + col]) |
                            (static_cast<uint32_t>(Vs[(k0 + 9) * pitch + col]) << 16);
        mma_bf16_16816(o[dt], a, b0, b1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      *reinterpret_cast<uint32_t*>(Q + r0 * pitch + dt * 8 + tg * 2) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
      *reinterpret_cast<uint32_t*>(Q + r1 * pitch + dt * 8 + tg * 2) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
  }
}

// Load one token row (3*hid bf16) into smem applying q-scale and rotary (interleaved pairs) to q and k.
// src == nullptr -> zeros (padding token).  Called by a group of threads: vector index vi in [0, 3*hid/8).
template <int DH>
__device__ __forceinline__ void load_token_vec(const __nv_bfloat16* src, __nv_bfloat16* dst, int vi, int hid, int pos,
                                               const float* __restrict__ rcos, const float* __restrict__ rsin,
                                               float qscale) {
  uint4 raw = make_uint4(0, 0, 0, 0);
  if (src) raw = *reinterpret_cast<const uint4*>(src + vi * 8);
  const int ch = vi * 8;
  if (src && ch < 2 * hid) {
    const float sc = ch < hid ? qscale : 1.0f;
    const int d0 = (ch % hid) % DH;                 // first of 8 dims inside the head
    uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = unpack_bf16(w[j]);
      const int pr = (d0 >> 1) + j;
      const float c = __ldg(rcos + pos * (DH / 2) + pr), s = __ldg(rsin + pos * (DH / 2) + pr);
      const float x0 = v.x * sc, x1 = v.y * sc;
      w[j] = pack_bf16(x0 * c - x1 * s, x1 * c + x0 * s);
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
  }
  *reinterpret_cast<uint4*>(dst + vi * 8) = raw;
}

// ------------------------------------------------------------------------------------------------ window attention
template <int NTOK, int DH>
__global__ void __launch_bounds__(256) window_attention_kernel(
    const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, const float* __restrict__ bias_table,
    const float* __restrict__ rcos, const float* __restrict__ rsin, int T, int H, int W, int heads, int wd, int wh,
    int ww, int sd, int sh, int sw, int Dp) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int hid = heads * DH;
  const int pitch = 3 * hid + 8;
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_att);
  const int tbl_n = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1);
  float* s_tbl = reinterpret_cast<float*>(smem_att + static_cast<size_t>(NTOK) * pitch * 2);   // [heads][tbl_n]
  int* s_lin = reinterpret_cast<int*>(s_tbl + heads * tbl_n);                                   // [NTOK]
  int* s_reg = s_lin + NTOK;                                                                    // [NTOK]
  long long* s_src = reinterpret_cast<long long*>(s_reg + NTOK);                                // [NTOK] (-1 = pad)

  const int nWw = W / ww, nWh = H / wh, nWd = Dp / wd;
  int widx = blockIdx.x;
  const int iw = widx % nWw; widx /= nWw;
  const int ih = widx % nWh; widx /= nWh;
  const int id = widx % nWd; widx /= nWd;
  const int b = widx;
  const bool shifted = (sd | sh | sw) != 0;

  for (int i = threadIdx.x; i < heads * tbl_n; i += blockDim.x) {
    const int h = i / tbl_n, r = i % tbl_n;
    s_tbl[i] = bias_table[r * heads + h];
  }
  if (threadIdx.x < NTOK) {
    const int n = threadIdx.x;
    const int w_ = n % ww, h_ = (n / ww) % wh, d_ = n / (ww * wh);
    s_lin[n] = (d_ * (2 * wh - 1) + h_) * (2 * ww - 1) + w_;
    const int zd = id * wd + d_, zh = ih * wh + h_, zw = iw * ww + w_;      // coordinates in the rolled volume
    int rd = 0, rh = 0, rw = 0;                                             // compute_mask region ids
    if (sd) rd = zd < Dp - wd ? 0 : (zd < Dp - sd ? 1 : 2);
    if (sh) rh = zh < H - wh ? 0 : (zh < H - sh ? 1 : 2);
    if (sw) rw = zw < W - ww ? 0 : (zw < W - sw ? 1 : 2);
    s_reg[n] = (rd * 3 + rh) * 3 + rw;
    const int od = (zd + sd) % Dp, oh = (zh + sh) % H, ow = (zw + sw) % W;  // roll(-shift): z[i] = x[(i+s) % n]
    s_src[n] = od < T ? ((static_cast<long long>(b) * T + od) * H + oh) * W + ow : -1;
  }
  __syncthreads();

  const int vpr = 3 * hid / 8;                    // 16-byte vectors per token row
  const float qscale = rsqrtf(static_cast<float>(DH));
  for (int i = threadIdx.x; i < NTOK * vpr; i += blockDim.x) {
    const int n = i / vpr, vi = i % vpr;
    const long long src = s_src[n];
    load_token_vec<DH>(src >= 0 ? qkv + src * 3 * hid : nullptr, tile + n * pitch, vi, hid, n, rcos, rsin, qscale);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int head = warp; head < heads; head += blockDim.x >> 5) {
    const float* tb = s_tbl + head * tbl_n;
    const int c0 = ((wd - 1) * (2 * wh - 1) + (wh - 1)) * (2 * ww - 1) + (ww - 1);
    warp_head_attention<NTOK, DH>(tile, pitch, hid, head, lane, [&](int i, int j) {
      float v = tb[s_lin[i] - s_lin[j] + c0];
      if (shifted && s_reg[i] != s_reg[j]) v += -100.0f;
      return v;
    });
  }
  __syncthreads();

  const int vpo = hid / 8;
  for (int i = threadIdx.x; i < NTOK * vpo; i += blockDim.x) {
    const int n = i / vpo, vi = i % vpo;
    const long long dst = s_src[n];
    if (dst >= 0)
      *reinterpret_cast<uint4*>(out + dst * hid + vi * 8) = *reinterpret_cast<const uint4*>(tile + n * pitch + vi * 8);
  }
}

// ------------------------------------------------------------------------------------------------ temporal attention
template <int NTOK, int DH>
__global__ void __launch_bounds__(256) temporal_attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ out,
                                                                 const float* __restrict__ rel_bias,
                                                                 const float* __restrict__ rcos,
                                                                 const float* __restrict__ rsin, int T, int HW,
                                                                 int heads) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int hid = heads * DH;
  const int pitch = 3 * hid + 8;
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_att);
  float* s_bias = reinterpret_cast<float*>(smem_att + static_cast<size_t>(NTOK) * pitch * 2);   // [heads][2T-1]
  const int b = blockIdx.x / HW, p = blockIdx.x % HW;
  for (int i = threadIdx.x; i < heads * (2 * T - 1); i += blockDim.x) s_bias[i] = rel_bias[i];
  const int vpr = 3 * hid / 8;
  const float qscale = rsqrtf(static_cast<float>(DH));
  for (int i = threadIdx.x; i < NTOK * vpr; i += blockDim.x) {
    const int n = i / vpr, vi = i % vpr;
    const __nv_bfloat16* src = n < T ? qkv + ((static_cast<long long>(b) * T + n) * HW + p) * 3 * hid : nullptr;
    load_token_vec<DH>(src, tile + n * pitch, vi, hid, n, rcos, rsin, qscale);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int head = warp; head < heads; head += blockDim.x >> 5) {
    const float* hb = s_bias + head * (2 * T - 1) + (T - 1);
    warp_head_attention<NTOK, DH>(tile, pitch, hid, head, lane, [&](int i, int j) {
      return (j < T && i < T) ? hb[j - i] : (j < T ? 0.f : -1.0e30f);
    });
  }
  __syncthreads();
  const int vpo = hid / 8;
  for (int i = threadIdx.x; i < T * vpo; i += blockDim.x) {
    const int n = i / vpo, vi = i % vpo;
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * T + n) * HW + p) * hid + vi * 8) =
        *reinterpret_cast<const uint4*>(tile + n * pitch + vi * 8);
  }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
    return EXTDM_ERR_CUDA;
  }
  return EXTDM_OK;
}

}  // namespace extdm

using namespace extdm;

namespace extdm {
// attn_core32.cu: tcgen05 edition for (2,4,4) windows, 8 heads x 32 (lane-masked compact scores)
bool attn_core32_supported(int heads, int dh, int wd, int wh, int ww, int H, int W);
int attn_core32_launch(const void* qkv, void* out, const float* bias_table, const float* rope_cos, const float* rope_sin,
                       int B, int T, int H, int W, int sd, int sh, int sw, cudaStream_t st);
}  // namespace extdm

extern "C" int extdm_window_attention(const void* qkv, void* out, const float* bias_table, const float* rope_cos,
                                      const float* rope_sin, int B, int T, int H, int W, int heads, int dh, int wd,
                                      int wh, int ww, int sd, int sh, int sw, void* stream) {
  static const bool legacy32 = getenv("EXTDM_WINATT_LEGACY") != nullptr;      // A/B: the mma.sync kernel below
  // ahead of the mma.sync kernel only with several tiles per SM (measured on B200, batch 32: 90 vs 105 us at the 16 x 16
  // level, 37 vs 30 / 22 vs 15 us at 8 x 8 / 4 x 4 where a CTA gets one or two tiles and their four head groups run one
  // after the other); EXTDM_WINATT_TC=1 forces it (tests)
  static const bool force32 = getenv("EXTDM_WINATT_TC") != nullptr;
  if (!legacy32 && extdm::attn_core32_supported(heads, dh, wd, wh, ww, H, W) &&
      (force32 || static_cast<long long>(B) * ((T + 1) / 2) * (H / 4) * (W / 4) >= 4ll * 3 * extdm::device_sm_count()))
    return extdm::attn_core32_launch(qkv, out, bias_table, rope_cos, rope_sin, B, T, H, W, sd, sh, sw,
                                     static_cast<cudaStream_t>(stream));
  const int ntok = wd * wh * ww;
  if (H % wh || W % ww || heads < 1 || heads > 8 || !((ntok == 64 && dh == 16) || (ntok == 32 && dh == 32) ||
                                                       (ntok == 64 && dh == 32) || (ntok == 32 && dh == 16))) {
    extdm_set_error("window_attention: supported windows have 32/64 tokens, dh 16/32, H,W multiples of the window",
                    __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  const int Dp = (T + wd - 1) / wd * wd;
  const int hid = heads * dh;
  const int tbl_n = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1);
  const size_t smem = static_cast<size_t>(ntok) * (3 * hid + 8) * 2 + static_cast<size_t>(heads) * tbl_n * 4 +
                      static_cast<size_t>(ntok) * (4 + 4 + 8);
  const int grid = B * (Dp / wd) * (H / wh) * (W / ww);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LAUNCH_WIN(N, D)                                                                                            \
  do {                                                                                                              \
    static SmemConfigured configured;                                                                               \
    if (!configured.covers(smem)) {                                                                                   \
      int rc = set_smem(window_attention_kernel<N, D>, smem);                                                       \
      if (rc) return rc;                                                                                            \
      configured.set(smem);                                                                                         \
    }                                                                                                               \
    window_attention_kernel<N, D><<<grid, 256, smem, st>>>(                                                         \
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), bias_table, rope_cos,   \
        rope_sin, T, H, W, heads, wd, wh, ww, sd, sh, sw, Dp);                                                      \
  } while (0)
  if (ntok == 64 && dh == 16) LAUNCH_WIN(64, 16);
  else if (ntok == 32 && dh == 32) LAUNCH_WIN(32, 32);
  else if (ntok == 64 && dh == 32) LAUNCH_WIN(64, 32);
  else LAUNCH_WIN(32, 16);
#undef LAUNCH_WIN
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

extern "C" int extdm_temporal_attention(const void* qkv, void* out, const float* rel_bias, const float* rope_cos,
                                        const float* rope_sin, int B, int T, int HW, int heads, int dh,
                                        void* stream) {
  if (T < 1 || T > 32 || (dh != 16 && dh != 32) || heads < 1 || heads > 8) {
    extdm_set_error("temporal_attention: T <= 32, dh 16/32, heads <= 8", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  const int ntok = T <= 16 ? 16 : 32;
  const int hid = heads * dh;
  const size_t smem = static_cast<size_t>(ntok) * (3 * hid + 8) * 2 + static_cast<size_t>(heads) * (2 * T - 1) * 4;
  const int grid = B * HW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LAUNCH_TMP(N, D)                                                                                            \
  do {                                                                                                              \
    static SmemConfigured configured;                                                                               \
    if (!configured.covers(smem)) {                                                                                   \
      int rc = set_smem(temporal_attention_kernel<N, D>, smem);                                                     \
      if (rc) return rc;                                                                                            \
      configured.set(smem);                                                                                         \
    }                                                                                                               \
    temporal_attention_kernel<N, D><<<grid, 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),           \
                                                             reinterpret_cast<__nv_bfloat16*>(out), rel_bias,       \
                                                             rope_cos, rope_sin, T, HW, heads);                     \
  } while (0)
  if (ntok == 16 && dh == 16) LAUNCH_TMP(16, 16);
  else if (ntok == 16 && dh == 32) LAUNCH_TMP(16, 32);
  else if (ntok == 32 && dh == 16) LAUNCH_TMP(32, 16);
  else LAUNCH_TMP(32, 32);
#undef LAUNCH_TMP
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}
