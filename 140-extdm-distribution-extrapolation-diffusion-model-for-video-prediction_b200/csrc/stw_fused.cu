// Fully fused shifted-window attention layer for the high-resolution UNet levels (C = 64 / 128):
//     y = x + proj( WindowAttention3D( chanLN(x) ) )          Residual(PreNorm(STWAttentionLayer))
// reference: model/BaseDM_adaptor/DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:139-159 (PreNorm/LayerNorm),
// :409-497 (WindowAttention3D), :499-560 (STWAttentionLayer: pad / roll / partition / mask / reverse).
//
// Un-fused, one layer moves x, LN(x), qkv (6x the size of x), attn-out and y through HBM (~21x |x| of traffic);
// fused, it reads x once and writes y once.  Persistent CTAs (one per SM) loop over windows; the next window's
// tokens are prefetched with cp.async while the current one is computed.  One warp per head:
//   * the head's Wq/Wk/Wv tiles live in REGISTERS as mma B-fragments for the whole kernel (loop-invariant),
//   * Q/K/V projections, rotary, QK^T, bias (+ shift mask), softmax and PV stay in registers
//     (mma.sync m16n8k16 bf16, accumulator -> operand fragment re-use, ldmatrix for every smem operand),
//   * the relative-position bias is expanded once per CTA into a [head][query][key] bf16 matrix (pre-multiplied by
//     log2 e) that ldmatrix delivers directly in accumulator layout: it initialises the score accumulators,
//   * the -100 shift mask is a 64-bit "different region" word per query row built with 8 ballots per window.
// (64-token x 64..128-channel tiles are below a tcgen05 tile of M = 128 rows per CTA and the score work is
// MUFU/issue bound, not MMA bound -- 537 M exponentials per level-0 launch -- so the warp-level tensor path is used.)
#include <stdlib.h>

#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16 pair (packed) -> two floats, one ALU op each
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

struct StwParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* gamma;
  const __nv_bfloat16* wqkv;   // [3*HID][C]
  const __nv_bfloat16* wproj;  // [C][HID]
  const float* proj_bias;
  const float* bias_table;     // [tbl_n][heads]
  const float* rcos;
  const float* rsin;           // [NTOK][DH/2]
  const float* ln_w;           // temporal mode: nn.LayerNorm weight / bias applied after the channel LayerNorm
  const float* ln_b;
  // optional fused producer (16-warp kernel): the layer input is x = silu(h*a[b,c] + d[b,c]) + res, i.e. the last
  // GroupNorm + SiLU + residual of the preceding ResnetBlock applied on load (h = p.x, the block2 conv output)
  const __nv_bfloat16* pre_res;
  const float* pre_ad;         // (B, 2, C): a then d
  int B, T, H, W, sd, sh, sw, Dp, n_windows;
  float eps;
};

constexpr float kLog2e = 1.4426950408889634f;

template <int NTOK, int DH, int C>
struct StwSmem {
  static constexpr int HEADS = 8;
  static constexpr int HID = HEADS * DH;
  static constexpr int XP = C + 8;        // pitch (bf16) of token x channel tiles (16 B skew: conflict-free ldmatrix)
  static constexpr int HP = HID + 8;      // pitch of token x hidden tiles
  static constexpr int BP = NTOK + 8;     // pitch of the bias matrix rows
  static constexpr size_t wproj = 0;                                          // [C][HP]
  static constexpr size_t bias = wproj + size_t(C) * HP * 2;                  // [HEADS][NTOK][BP] bf16
  static constexpr size_t raw = bias + size_t(HEADS) * NTOK * BP * 2;         // 2 x [NTOK][XP] raw tokens (cp.async)
  static constexpr size_t xn = raw + 2 * size_t(NTOK) * XP * 2;               // [NTOK][XP] normalised tokens / staging
  static constexpr size_t v = xn + size_t(NTOK) * XP * 2;                     // [NTOK][HP]
  static constexpr size_t o = v + size_t(NTOK) * HP * 2;                      // [NTOK][HP]
  static constexpr size_t rope = o + size_t(NTOK) * HP * 2;                   // cos, sin [NTOK][DH/2] fp32
  static constexpr size_t misc = rope + 2 * size_t(NTOK) * (DH / 2) * 4;      // gamma[C], pbias[C]
  static constexpr size_t emask = misc + 2 * size_t(C) * 4;                   // [2][8] u32 region-membership words
  static constexpr size_t total = emask + 2 * 8 * 4;
};

// TEMPORAL = false: a "window" is a (WD,4,4) block of the rolled volume (STWAttentionLayer).
// TEMPORAL = true : a "window" is the T (<= NTOK) frames of one pixel -- the full temporal attention layer
//   Residual(PreNorm(EinopsToAndFrom('b c t h w -> b (h w) t c', AttentionLayer))), ...cross_multi.py:253-328:
//   z = chanLN(x); u = LayerNorm(z); y = x + z + to_out(attn(u)), rotary over the frame index, T5 relative-position
//   bias (heads, 2T-1) in p.bias_table, no projection bias; key slots >= T are masked out.
template <int NTOK, int DH, int C, bool TEMPORAL>
__global__ void __launch_bounds__(256, (TEMPORAL && DH == 16) ? 2 : 1) stw_fused_kernel(const __grid_constant__ StwParams p) {
  using L = StwSmem<NTOK, DH, C>;
  constexpr int HEADS = 8, HID = L::HID, XP = L::XP, HP = L::HP, BP = L::BP;
  constexpr int WD = NTOK / 16;          // window = (WD, 4, 4)
  constexpr int MT = NTOK / 16;          // query / token m-tiles
  constexpr int DT = DH / 8;             // n-tiles over the head dim
  constexpr int KS = DH / 16;            // k-steps of Q K^T
  constexpr int NT = NTOK / 8;           // key n-tiles
  constexpr int CK = C / 16;             // k-steps of the projections from C
  constexpr int TPT = 256 / NTOK;        // threads per token in the LayerNorm / copy phases
  constexpr int CPT = C / TPT;           // channels per thread there
  static_assert(CPT % 8 == 0, "LayerNorm slice must be a multiple of 8 channels");
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* s_wproj = reinterpret_cast<__nv_bfloat16*>(sm + L::wproj);
  __nv_bfloat16* s_bias = reinterpret_cast<__nv_bfloat16*>(sm + L::bias);
  __nv_bfloat16* s_raw = reinterpret_cast<__nv_bfloat16*>(sm + L::raw);
  __nv_bfloat16* s_xn = reinterpret_cast<__nv_bfloat16*>(sm + L::xn);
  __nv_bfloat16* s_v = reinterpret_cast<__nv_bfloat16*>(sm + L::v);
  __nv_bfloat16* s_o = reinterpret_cast<__nv_bfloat16*>(sm + L::o);
  float* s_cos = reinterpret_cast<float*>(sm + L::rope);
  float* s_sin = s_cos + NTOK * (DH / 2);
  float* s_gamma = reinterpret_cast<float*>(sm + L::misc);
  float* s_pbias = s_gamma + C;
  uint32_t* s_E = reinterpret_cast<uint32_t*>(sm + L::emask);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const bool shifted = (p.sd | p.sh | p.sw) != 0;
  const int nWw = p.W / 4, nWh = p.H / 4, nWd = p.Dp / WD;
  // ldmatrix row/column offsets of this lane inside a 16x16 tile
  const int lrow = lane & 15, lcol = (lane >> 4) * 8;

  // ---- one-time staging: proj weights, tables, the expanded bias matrix, this warp's Wq/Wk/Wv fragments
  for (int i = tid; i < C * (HID / 8); i += 256) {
    const int r = i / (HID / 8), c8 = i % (HID / 8);
    *reinterpret_cast<uint4*>(s_wproj + r * HP + c8 * 8) = *reinterpret_cast<const uint4*>(p.wproj + r * HID + c8 * 8);
  }
  for (int i = tid; i < NTOK * (DH / 2); i += 256) { s_cos[i] = p.rcos[i]; s_sin[i] = p.rsin[i]; }
  for (int i = tid; i < C; i += 256) { s_gamma[i] = p.gamma[i]; s_pbias[i] = p.proj_bias ? p.proj_bias[i] : 0.f; }
  for (int i = tid; i < HEADS * NTOK * NTOK; i += 256) {
    const int j = i % NTOK, q = (i / NTOK) % NTOK, h = i / (NTOK * NTOK);
    float bv;
    if (TEMPORAL) {
      bv = j >= p.T ? -1.0e30f : (q < p.T ? p.bias_table[h * (2 * p.T - 1) + (j - q + p.T - 1)] * kLog2e : 0.f);
    } else {
      const int rel = (((q >> 4) - (j >> 4) + WD - 1) * 7 + (((q >> 2) & 3) - ((j >> 2) & 3) + 3)) * 7 +
                      ((q & 3) - (j & 3) + 3);
      bv = p.bias_table[rel * HEADS + h] * kLog2e;
    }
    s_bias[(h * NTOK + q) * BP + j] = __float2bfloat16(bv);
  }
  const int head = warp;
  uint32_t wq[CK][DT][2], wk[CK][DT][2], wv[CK][DT][2];
#pragma unroll
  for (int ks = 0; ks < CK; ++ks)
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int row = head * DH + dt * 8 + g, col = ks * 16 + tg * 2;
      const __nv_bfloat16* bq = p.wqkv + static_cast<long long>(row) * C + col;
      const __nv_bfloat16* bk = p.wqkv + static_cast<long long>(HID + row) * C + col;
      const __nv_bfloat16* bv = p.wqkv + static_cast<long long>(2 * HID + row) * C + col;
      wq[ks][dt][0] = *reinterpret_cast<const uint32_t*>(bq);
      wq[ks][dt][1] = *reinterpret_cast<const uint32_t*>(bq + 8);
      wk[ks][dt][0] = *reinterpret_cast<const uint32_t*>(bk);
      wk[ks][dt][1] = *reinterpret_cast<const uint32_t*>(bk + 8);
      wv[ks][dt][0] = *reinterpret_cast<const uint32_t*>(bv);
      wv[ks][dt][1] = *reinterpret_cast<const uint32_t*>(bv + 8);
    }

  // window bookkeeping (all arithmetic, no shared state): source pixel of token n of window widx, or -1 (T padding)
  struct Win { int b, id, ih, iw; };
  auto decode = [&](int widx) {
    Win w;
    if (TEMPORAL) {                       // widx = (b, pixel): b in w.b, pixel in w.iw
      w.iw = widx % (p.H * p.W);
      w.b = widx / (p.H * p.W);
      w.id = w.ih = 0;
      return w;
    }
    w.iw = widx % nWw; widx /= nWw;
    w.ih = widx % nWh; widx /= nWh;
    w.id = widx % nWd;
    w.b = widx / nWd;
    return w;
  };
  auto src_pixel = [&](const Win& w, int n) -> int {
    if (TEMPORAL) return n < p.T ? (w.b * p.T + n) * (p.H * p.W) + w.iw : -1;
    int od = w.id * WD + (n >> 4) + p.sd, oh = w.ih * 4 + ((n >> 2) & 3) + p.sh, ow = w.iw * 4 + (n & 3) + p.sw;
    if (od >= p.Dp) od -= p.Dp;
    if (oh >= p.H) oh -= p.H;
    if (ow >= p.W) ow -= p.W;
    return od < p.T ? ((w.b * p.T + od) * p.H + oh) * p.W + ow : -1;
  };
  // mask region code of token n: one bit per shifted dim, set for the wrapped part of the last window slab
  auto region_code = [&](const Win& w, int n) -> int {
    int c = 0;
    if (p.sd && w.id == nWd - 1 && (n >> 4) >= WD - p.sd) c |= 1;
    if (p.sh && w.ih == nWh - 1 && ((n >> 2) & 3) >= 4 - p.sh) c |= 2;
    if (p.sw && w.iw == nWw - 1 && (n & 3) >= 4 - p.sw) c |= 4;
    return c;
  };
  auto prefetch_window = [&](const Win& w, int buf) {
    __nv_bfloat16* dst = s_raw + buf * NTOK * XP;
    for (int i = tid; i < NTOK * (C / 8); i += 256) {
      const int n = i / (C / 8), c8 = i % (C / 8);
      const int s = src_pixel(w, n);
      cp_async16(dst + n * XP + c8 * 8, p.x + (s >= 0 ? static_cast<long long>(s) * C + c8 * 8 : 0), s >= 0 ? 16 : 0);
    }
    cp_async_commit();
  };

  int widx = blockIdx.x;
  int buf = 0;
  if (widx < p.n_windows) prefetch_window(decode(widx), 0);

  const float qscale = rsqrtf(static_cast<float>(DH)) * kLog2e;
  constexpr float kMask = -100.0f * kLog2e;

  for (; widx < p.n_windows; widx += gridDim.x) {
    const Win win = decode(widx);
    const int nxt = widx + gridDim.x;
    cp_async_wait<0>();
    __syncthreads();                                       // S1: raw[buf] landed; previous window fully retired
    if (nxt < p.n_windows) prefetch_window(decode(nxt), buf ^ 1);
    const __nv_bfloat16* raw = s_raw + buf * NTOK * XP;
    const bool has_mask = !TEMPORAL && shifted && ((p.sd && win.id == nWd - 1) || (p.sh && win.ih == nWh - 1) ||
                                      (p.sw && win.iw == nWw - 1));

    // ---- region-membership words: E[c] bit j = (code_j == c); query row i masks the keys in ~E[code_i]
    if (has_mask && warp < NTOK / 32) {
      const int code = region_code(win, warp * 32 + lane);
      uint32_t mine = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t bal = __ballot_sync(0xffffffffu, code == c);
        if (lane == c) mine = bal;
      }
      if (lane < 8) s_E[warp * 8 + lane] = mine;
    }

    // ---- channel LayerNorm (biased variance, gamma only); TPT threads per token; padding tokens stay exactly zero
    {
      const int n = tid / TPT, part = tid % TPT;
      float v[CPT];
#pragma unroll
      for (int k = 0; k < CPT; k += 8) {
        const uint4 t = *reinterpret_cast<const uint4*>(raw + n * XP + part * CPT + k);
        const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
        v[k] = a.x; v[k + 1] = a.y; v[k + 2] = b.x; v[k + 3] = b.y;
        v[k + 4] = c.x; v[k + 5] = c.y; v[k + 6] = d.x; v[k + 7] = d.y;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) sum += v[j];
#pragma unroll
      for (int o = 1; o < TPT; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) { const float d = v[j] - mean; sq += d * d; }
#pragma unroll
      for (int o = 1; o < TPT; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = src_pixel(win, n) >= 0 ? rsqrtf(sq * (1.0f / C) + p.eps) : 0.f;
      if (TEMPORAL) {
        // z = chanLN(x); residual stream becomes x + z (written back over the raw tile); u = LayerNorm(z)*w + b
        float z[CPT];
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          z[j] = (v[j] - mean) * rstd * s_gamma[part * CPT + j];
          s2 += z[j];
        }
#pragma unroll
        for (int o = 1; o < TPT; o <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const float mean2 = s2 * (1.0f / C);
        float q2 = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) { const float d = z[j] - mean2; q2 += d * d; }
#pragma unroll
        for (int o = 1; o < TPT; o <<= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
        const float rstd2 = src_pixel(win, n) >= 0 ? rsqrtf(q2 * (1.0f / C) + p.eps) : 0.f;
        __nv_bfloat16* rw = s_raw + buf * NTOK * XP + n * XP + part * CPT;
#pragma unroll
        for (int k = 0; k < CPT; k += 8) {
          uint32_t pu[4], pr[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = part * CPT + k + 2 * j;
            const float u0 = src_pixel(win, n) >= 0 ? (z[k + 2 * j] - mean2) * rstd2 * __ldg(p.ln_w + c) + __ldg(p.ln_b + c) : 0.f;
            const float u1 = src_pixel(win, n) >= 0
                                 ? (z[k + 2 * j + 1] - mean2) * rstd2 * __ldg(p.ln_w + c + 1) + __ldg(p.ln_b + c + 1)
                                 : 0.f;
            pu[j] = pack_bf16(u0, u1);
            pr[j] = pack_bf16(v[k + 2 * j] + z[k + 2 * j], v[k + 2 * j + 1] + z[k + 2 * j + 1]);
          }
          *reinterpret_cast<uint4*>(s_xn + n * XP + part * CPT + k) = make_uint4(pu[0], pu[1], pu[2], pu[3]);
          *reinterpret_cast<uint4*>(rw + k) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < CPT; k += 8) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = part * CPT + k + 2 * j;
            pk[j] = pack_bf16((v[k + 2 * j] - mean) * rstd * s_gamma[c], (v[k + 2 * j + 1] - mean) * rstd * s_gamma[c + 1]);
          }
          *reinterpret_cast<uint4*>(s_xn + n * XP + part * CPT + k) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    __syncthreads();                                       // S2: s_xn, s_E ready

    // ---- per-head projections (Wq/Wk/Wv fragments from registers): K -> B fragments, V -> smem, Q -> A fragments
    uint32_t kfrag[NT][KS][2];
    uint32_t qa[MT][KS][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float aq[DT][4], ak[DT][4], av[DT][4];
#pragma unroll
      for (int dt = 0; dt < DT; ++dt)
#pragma unroll
        for (int j = 0; j < 4; ++j) { aq[dt][j] = 0.f; ak[dt][j] = 0.f; av[dt][j] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < CK; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, s_xn + (mt * 16 + lrow) * XP + ks * 16 + lcol);
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
          mma16816(aq[dt], a, wq[ks][dt][0], wq[ks][dt][1]);
          mma16816(ak[dt], a, wk[ks][dt][0], wk[ks][dt][1]);
          mma16816(av[dt], a, wv[ks][dt][0], wv[ks][dt][1]);
        }
      }
      // rotary on Q and K (pair = the two accumulator columns a thread owns)
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        const int pr = dt * 4 + tg;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int tok = mt * 16 + hf * 8 + g;
          const float c = s_cos[tok * (DH / 2) + pr], s = s_sin[tok * (DH / 2) + pr];
          const float k0 = ak[dt][hf * 2], k1 = ak[dt][hf * 2 + 1];
          // B fragment of key n-tile (2*mt + hf): k-step dt/2, register dt%2
          kfrag[2 * mt + hf][dt / 2][dt % 2] = pack_bf16(k0 * c - k1 * s, k1 * c + k0 * s);
          const float q0 = aq[dt][hf * 2] * qscale, q1 = aq[dt][hf * 2 + 1] * qscale;
          // A fragment of k-step dt/2: registers (dt%2)*2 + {0: row g, 1: row g+8}
          qa[mt][dt / 2][(dt % 2) * 2 + hf] = pack_bf16(q0 * c - q1 * s, q1 * c + q0 * s);
          *reinterpret_cast<uint32_t*>(s_v + tok * HP + head * DH + dt * 8 + tg * 2) =
              pack_bf16(av[dt][hf * 2], av[dt][hf * 2 + 1]);
        }
      }
    }
    __syncwarp();                                          // this head's V columns are read back by this warp only

    const __nv_bfloat16* bias_h = s_bias + head * NTOK * BP;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      // scores start from the bias matrix (ldmatrix delivers it in accumulator layout), then += Q K^T
      float s[NT][4];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bb[4];
        ldsm_x4(bb, bias_h + (mt * 16 + lrow) * BP + np * 16 + lcol);
        s[2 * np][0] = bf_lo(bb[0]); s[2 * np][1] = bf_hi(bb[0]);
        s[2 * np][2] = bf_lo(bb[1]); s[2 * np][3] = bf_hi(bb[1]);
        s[2 * np + 1][0] = bf_lo(bb[2]); s[2 * np + 1][1] = bf_hi(bb[2]);
        s[2 * np + 1][2] = bf_lo(bb[3]); s[2 * np + 1][3] = bf_hi(bb[3]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma16816(s[nt], qa[mt][ks], kfrag[nt][ks][0], kfrag[nt][ks][1]);
      if (has_mask) {
        const int c0 = region_code(win, r0), c1 = region_code(win, r1);
        uint32_t e0[NTOK / 32], e1[NTOK / 32];
#pragma unroll
        for (int w = 0; w < NTOK / 32; ++w) { e0[w] = ~s_E[w * 8 + c0]; e1[w] = ~s_E[w * 8 + c1]; }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int bit = (nt * 8) % 32 + tg * 2;
          const uint32_t m0 = e0[nt / 4] >> bit, m1 = e1[nt / 4] >> bit;
          if (m0 & 1) s[nt][0] += kMask;
          if (m0 & 2) s[nt][1] += kMask;
          if (m1 & 1) s[nt][2] += kMask;
          if (m1 & 2) s[nt][3] += kMask;
        }
      }
      float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
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
        s[nt][0] = fast_exp2(s[nt][0] - m0);
        s[nt][1] = fast_exp2(s[nt][1] - m0);
        s[nt][2] = fast_exp2(s[nt][2] - m1);
        s[nt][3] = fast_exp2(s[nt][3] - m1);
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
      for (int ps = 0; ps < NTOK / 16; ++ps) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * ps][0], s[2 * ps][1]);
        a[1] = pack_bf16(s[2 * ps][2], s[2 * ps][3]);
        a[2] = pack_bf16(s[2 * ps + 1][0], s[2 * ps + 1][1]);
        a[3] = pack_bf16(s[2 * ps + 1][2], s[2 * ps + 1][3]);
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {
          // V^T fragments: tokens ps*16..+15 (k) x head dims dp*16..+15 (n), transposed on the fly by ldmatrix
          uint32_t vb[4];
          ldsm_x4_trans(vb, s_v + (ps * 16 + lrow) * HP + head * DH + dp * 16 + lcol);
          mma16816(o[2 * dp], a, vb[0], vb[1]);
          mma16816(o[2 * dp + 1], a, vb[2], vb[3]);
        }
      }
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        *reinterpret_cast<uint32_t*>(s_o + r0 * HP + head * DH + dt * 8 + tg * 2) =
            pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        *reinterpret_cast<uint32_t*>(s_o + r1 * HP + head * DH + dt * 8 + tg * 2) =
            pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
      }
    }
    __syncthreads();                                       // S3: all heads' outputs in s_o; s_xn is dead

    // ---- output projection + bias + residual -> s_xn (staging), then coalesced stores
    {
      constexpr int NPW = C / 64;                          // n-tiles per warp (C/8 n-tiles over 8 warps)
      float acc[MT][NPW][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int q = 0; q < NPW; ++q) { acc[mt][q][0] = acc[mt][q][1] = acc[mt][q][2] = acc[mt][q][3] = 0.f; }
#pragma unroll
      for (int kp = 0; kp < HID / 32; ++kp) {
        // B fragments of this warp's n-tiles for two k-steps (32 hidden columns) per ldmatrix.x4
        uint32_t b[NPW][4];
#pragma unroll
        for (int q = 0; q < NPW; ++q)
          ldsm_x4(b[q], s_wproj + ((warp * NPW + q) * 8 + (lane & 7)) * HP + kp * 32 + (lane >> 3) * 8);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t a[4];
            ldsm_x4(a, s_o + (mt * 16 + lrow) * HP + (kp * 2 + kk) * 16 + lcol);
#pragma unroll
            for (int q = 0; q < NPW; ++q) mma16816(acc[mt][q], a, b[q][kk * 2], b[q][kk * 2 + 1]);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
          const int col = (warp * NPW + q) * 8 + tg * 2;
          const float b0 = s_pbias[col], b1 = s_pbias[col + 1];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int tok = mt * 16 + hf * 8 + g;
            const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t*>(raw + tok * XP + col));
            *reinterpret_cast<uint32_t*>(s_xn + tok * XP + col) =
                pack_bf16(acc[mt][q][hf * 2] + b0 + r.x, acc[mt][q][hf * 2 + 1] + b1 + r.y);
          }
        }
    }
    __syncthreads();                                       // S4: projected tile staged
    for (int i = tid; i < NTOK * (C / 8); i += 256) {
      const int n = i / (C / 8), c8 = i % (C / 8);
      const int d = src_pixel(win, n);
      if (d >= 0)
        *reinterpret_cast<uint4*>(p.y + static_cast<long long>(d) * C + c8 * 8) =
            *reinterpret_cast<const uint4*>(s_xn + n * XP + c8 * 8);
    }
    buf ^= 1;
  }
  cp_async_wait<0>();
}

template <int NTOK, int DH, int C, bool TEMPORAL = false>
static int launch_stw(const StwParams& p, cudaStream_t st) {
  using L = StwSmem<NTOK, DH, C>;
  constexpr size_t smem = L::total;
  static SmemConfigured configured;
  const int sms = device_sm_count();
  if (!configured.covers(smem)) {
    cudaError_t e = cudaFuncSetAttribute(stw_fused_kernel<NTOK, DH, C, TEMPORAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem);
  }
  const int resident = sms * ((TEMPORAL && DH == 16) ? 2 : 1);
  const int grid = p.n_windows < resident ? p.n_windows : resident;
  stw_fused_kernel<NTOK, DH, C, TEMPORAL><<<grid, 256, smem, st>>>(p);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}


// ------------------------------------------------------------------------------------------------ 16-warp variant
// Same layer, 512 threads: TWO warps per head, each owning half of the query m-tiles of the window.  ncu on the
// 8-warp kernel showed two warps per scheduler issuing one dependent instruction per ~7 cycles ("wait" / "selected"
// stalls, 21% issue utilisation); doubling the warps per SM doubles the latency hiding.  To fit 128 registers per
// thread the Wq/Wk/Wv tiles go back to shared memory (ldmatrix per use) and K is exchanged between the two warps of
// a head through shared memory (one 64-thread named barrier per head).  C = 64 only (231 KB of shared memory).
template <int NTOK, int DH, int C>
struct Stw16Smem {
  static constexpr int HEADS = 8;
  static constexpr int HID = HEADS * DH;
  static constexpr int XP = C + 8, HP = HID + 8, BP = NTOK + 8;
  static constexpr size_t wproj = 0;                                          // [C][HP]
  static constexpr size_t wqkv = wproj + size_t(C) * HP * 2;                  // [3*HID][XP]
  static constexpr size_t bias = wqkv + size_t(3 * HID) * XP * 2;             // [HEADS][NTOK][BP]
  static constexpr size_t raw = bias + size_t(HEADS) * NTOK * BP * 2;         // 2 x [NTOK][XP]
  static constexpr size_t xn = raw + 2 * size_t(NTOK) * XP * 2;               // [NTOK][XP]
  static constexpr size_t k = xn + size_t(NTOK) * XP * 2;                     // [NTOK][HP]
  static constexpr size_t v = k + size_t(NTOK) * HP * 2;                      // [NTOK][HP]
  static constexpr size_t o = v + size_t(NTOK) * HP * 2;                      // [NTOK][HP]
  static constexpr size_t rope = o + size_t(NTOK) * HP * 2;
  static constexpr size_t misc = rope + 2 * size_t(NTOK) * (DH / 2) * 4;
  static constexpr size_t emask = misc + 2 * size_t(C) * 4;
  static constexpr size_t total = emask + 2 * 8 * 4;
};

template <int NTOK, int DH, int C>
__global__ void __launch_bounds__(512, 1) stw_fused16_kernel(const __grid_constant__ StwParams p) {
  using L = Stw16Smem<NTOK, DH, C>;
  constexpr int HEADS = 8, HID = L::HID, XP = L::XP, HP = L::HP, BP = L::BP;
  constexpr int NTH = 512;
  constexpr int WD = NTOK / 16;
  constexpr int MT = NTOK / 16, MH = MT / 2;   // m-tiles per window / per warp
  constexpr int DT = DH / 8, KS = DH / 16, NT = NTOK / 8, CK = C / 16;
  constexpr int TPT = NTH / NTOK, CPT = C / TPT;
  static_assert(MT % 2 == 0 && CPT == 8 || CPT == 4, "unsupported window / channel combination");
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* s_wproj = reinterpret_cast<__nv_bfloat16*>(sm + L::wproj);
  __nv_bfloat16* s_wqkv = reinterpret_cast<__nv_bfloat16*>(sm + L::wqkv);
  __nv_bfloat16* s_bias = reinterpret_cast<__nv_bfloat16*>(sm + L::bias);
  __nv_bfloat16* s_raw = reinterpret_cast<__nv_bfloat16*>(sm + L::raw);
  __nv_bfloat16* s_xn = reinterpret_cast<__nv_bfloat16*>(sm + L::xn);
  __nv_bfloat16* s_k = reinterpret_cast<__nv_bfloat16*>(sm + L::k);
  __nv_bfloat16* s_v = reinterpret_cast<__nv_bfloat16*>(sm + L::v);
  __nv_bfloat16* s_o = reinterpret_cast<__nv_bfloat16*>(sm + L::o);
  float* s_cos = reinterpret_cast<float*>(sm + L::rope);
  float* s_sin = s_cos + NTOK * (DH / 2);
  float* s_gamma = reinterpret_cast<float*>(sm + L::misc);
  float* s_pbias = s_gamma + C;
  uint32_t* s_E = reinterpret_cast<uint32_t*>(sm + L::emask);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int head = warp & 7, half = warp >> 3;
  const bool shifted = (p.sd | p.sh | p.sw) != 0;
  const int nWw = p.W / 4, nWh = p.H / 4, nWd = p.Dp / WD;
  const int lrow = lane & 15, lcol = (lane >> 4) * 8;

  for (int i = tid; i < 3 * HID * (C / 8); i += NTH) {
    const int r = i / (C / 8), c8 = i % (C / 8);
    *reinterpret_cast<uint4*>(s_wqkv + r * XP + c8 * 8) = *reinterpret_cast<const uint4*>(p.wqkv + r * C + c8 * 8);
  }
  for (int i = tid; i < C * (HID / 8); i += NTH) {
    const int r = i / (HID / 8), c8 = i % (HID / 8);
    *reinterpret_cast<uint4*>(s_wproj + r * HP + c8 * 8) = *reinterpret_cast<const uint4*>(p.wproj + r * HID + c8 * 8);
  }
  for (int i = tid; i < NTOK * (DH / 2); i += NTH) { s_cos[i] = p.rcos[i]; s_sin[i] = p.rsin[i]; }
  for (int i = tid; i < C; i += NTH) { s_gamma[i] = p.gamma[i]; s_pbias[i] = p.proj_bias ? p.proj_bias[i] : 0.f; }
  for (int i = tid; i < HEADS * NTOK * NTOK; i += NTH) {
    const int j = i % NTOK, q = (i / NTOK) % NTOK, h = i / (NTOK * NTOK);
    const int rel = (((q >> 4) - (j >> 4) + WD - 1) * 7 + (((q >> 2) & 3) - ((j >> 2) & 3) + 3)) * 7 +
                    ((q & 3) - (j & 3) + 3);
    s_bias[(h * NTOK + q) * BP + j] = __float2bfloat16(p.bias_table[rel * HEADS + h] * kLog2e);
  }

  struct Win { int b, id, ih, iw; };
  auto decode = [&](int widx) {
    Win w;
    w.iw = widx % nWw; widx /= nWw;
    w.ih = widx % nWh; widx /= nWh;
    w.id = widx % nWd;
    w.b = widx / nWd;
    return w;
  };
  auto src_pixel = [&](const Win& w, int n) -> int {
    int od = w.id * WD + (n >> 4) + p.sd, oh = w.ih * 4 + ((n >> 2) & 3) + p.sh, ow = w.iw * 4 + (n & 3) + p.sw;
    if (od >= p.Dp) od -= p.Dp;
    if (oh >= p.H) oh -= p.H;
    if (ow >= p.W) ow -= p.W;
    return od < p.T ? ((w.b * p.T + od) * p.H + oh) * p.W + ow : -1;
  };
  auto region_code = [&](const Win& w, int n) -> int {
    int c = 0;
    if (p.sd && w.id == nWd - 1 && (n >> 4) >= WD - p.sd) c |= 1;
    if (p.sh && w.ih == nWh - 1 && ((n >> 2) & 3) >= 4 - p.sh) c |= 2;
    if (p.sw && w.iw == nWw - 1 && (n & 3) >= 4 - p.sw) c |= 4;
    return c;
  };
  auto prefetch_window = [&](const Win& w, int buf) {
    __nv_bfloat16* dst = s_raw + buf * NTOK * XP;
    for (int i = tid; i < NTOK * (C / 8); i += NTH) {
      const int n = i / (C / 8), c8 = i % (C / 8);
      const int s = src_pixel(w, n);
      cp_async16(dst + n * XP + c8 * 8, p.x + (s >= 0 ? static_cast<long long>(s) * C + c8 * 8 : 0), s >= 0 ? 16 : 0);
    }
    cp_async_commit();
  };

  // fused-producer mode: this thread's 16-byte slice of the residual branch, fetched one window ahead
  const bool pre = p.pre_res != nullptr;
  auto load_res = [&](const Win& w) -> uint4 {
    const int s = src_pixel(w, tid / TPT);
    if (!pre || s < 0 || CPT != 8) return make_uint4(0u, 0u, 0u, 0u);
    return *reinterpret_cast<const uint4*>(p.pre_res + static_cast<long long>(s) * C + (tid % TPT) * CPT);
  };

  int widx = blockIdx.x;
  int buf = 0;
  uint4 res_next = make_uint4(0u, 0u, 0u, 0u);
  if (widx < p.n_windows) {
    prefetch_window(decode(widx), 0);
    res_next = load_res(decode(widx));
  }
  const float qscale = rsqrtf(static_cast<float>(DH)) * kLog2e;
  constexpr float kMask = -100.0f * kLog2e;

  for (; widx < p.n_windows; widx += gridDim.x) {
    const Win win = decode(widx);
    const int nxt = widx + gridDim.x;
    cp_async_wait<0>();
    __syncthreads();                                       // S1
    const uint4 res_cur = res_next;
    if (nxt < p.n_windows) {
      prefetch_window(decode(nxt), buf ^ 1);
      res_next = load_res(decode(nxt));
    }
    __nv_bfloat16* raw = s_raw + buf * NTOK * XP;
    const bool has_mask = shifted && ((p.sd && win.id == nWd - 1) || (p.sh && win.ih == nWh - 1) ||
                                      (p.sw && win.iw == nWw - 1));
    if (has_mask && warp < NTOK / 32) {
      const int code = region_code(win, warp * 32 + lane);
      uint32_t mine = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t bal = __ballot_sync(0xffffffffu, code == c);
        if (lane == c) mine = bal;
      }
      if (lane < 8) s_E[warp * 8 + lane] = mine;
    }
    // ---- channel LayerNorm: TPT threads per token, CPT channels each
    {
      const int n = tid / TPT, part = tid % TPT;
      float v[CPT];
      if constexpr (CPT == 8) {
        const uint4 t = *reinterpret_cast<const uint4*>(raw + n * XP + part * CPT);
        const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
        if (pre) {
          // x = silu(GroupNorm(h)) + res, rounded to bf16 exactly as the stand-alone groupnorm_apply stores it; the
          // tile keeps x (the layer's residual); T-padding tokens stay zero
          const bool live = src_pixel(win, n) >= 0;
          const float* ad = p.pre_ad + static_cast<long long>(win.b) * 2 * C + part * CPT;
          const uint32_t rw[4] = {res_cur.x, res_cur.y, res_cur.z, res_cur.w};
          uint32_t xb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 rr = unpack_bf16(rw[j]);
            const float x0 = live ? silu_fast(v[2 * j] * __ldg(ad + 2 * j) + __ldg(ad + C + 2 * j)) + rr.x : 0.f;
            const float x1 = live ? silu_fast(v[2 * j + 1] * __ldg(ad + 2 * j + 1) + __ldg(ad + C + 2 * j + 1)) + rr.y : 0.f;
            xb[j] = pack_bf16(x0, x1);
            const float2 back = unpack_bf16(xb[j]);
            v[2 * j] = back.x;
            v[2 * j + 1] = back.y;
          }
          *reinterpret_cast<uint4*>(raw + n * XP + part * CPT) = make_uint4(xb[0], xb[1], xb[2], xb[3]);
        }
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(raw + n * XP + part * CPT);
        const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) sum += v[j];
#pragma unroll
      for (int o = 1; o < TPT; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) { const float d = v[j] - mean; sq += d * d; }
#pragma unroll
      for (int o = 1; o < TPT; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = src_pixel(win, n) >= 0 ? rsqrtf(sq * (1.0f / C) + p.eps) : 0.f;
      uint32_t pk[CPT / 2];
#pragma unroll
      for (int j = 0; j < CPT / 2; ++j) {
        const int c = part * CPT + 2 * j;
        pk[j] = pack_bf16((v[2 * j] - mean) * rstd * s_gamma[c], (v[2 * j + 1] - mean) * rstd * s_gamma[c + 1]);
      }
      if constexpr (CPT == 8)
        *reinterpret_cast<uint4*>(s_xn + n * XP + part * CPT) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      else
        *reinterpret_cast<uint2*>(s_xn + n * XP + part * CPT) = make_uint2(pk[0], pk[1]);
    }
    __syncthreads();                                       // S2: s_xn, s_E ready

    // ---- projections of this warp's m-tiles: Q -> registers, K (rotated) and V -> shared memory
    uint32_t qa[MH][KS][4];
    const __nv_bfloat16* wq = s_wqkv + (head * DH) * XP;
    const __nv_bfloat16* wk = s_wqkv + (HID + head * DH) * XP;
    const __nv_bfloat16* wv = s_wqkv + (2 * HID + head * DH) * XP;
#pragma unroll
    for (int mi = 0; mi < MH; ++mi) {
      const int mt = half * MH + mi;
      float aq[DT][4], ak[DT][4], av[DT][4];
#pragma unroll
      for (int dt = 0; dt < DT; ++dt)
#pragma unroll
        for (int j = 0; j < 4; ++j) { aq[dt][j] = 0.f; ak[dt][j] = 0.f; av[dt][j] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < CK; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, s_xn + (mt * 16 + lrow) * XP + ks * 16 + lcol);
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {
          // weight rows dp*16..+15 (n) x channels ks*16..+15 (k): B fragments of two n-tiles
          uint32_t bq[4], bk[4], bv[4];
          const int wrow = dp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, wcol = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(bq, wq + wrow * XP + wcol);
          ldsm_x4(bk, wk + wrow * XP + wcol);
          ldsm_x4(bv, wv + wrow * XP + wcol);
          mma16816(aq[2 * dp], a, bq[0], bq[1]);
          mma16816(aq[2 * dp + 1], a, bq[2], bq[3]);
          mma16816(ak[2 * dp], a, bk[0], bk[1]);
          mma16816(ak[2 * dp + 1], a, bk[2], bk[3]);
          mma16816(av[2 * dp], a, bv[0], bv[1]);
          mma16816(av[2 * dp + 1], a, bv[2], bv[3]);
        }
      }
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        const int pr = dt * 4 + tg;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int tok = mt * 16 + hf * 8 + g;
          const float c = s_cos[tok * (DH / 2) + pr], s = s_sin[tok * (DH / 2) + pr];
          const float k0 = ak[dt][hf * 2], k1 = ak[dt][hf * 2 + 1];
          *reinterpret_cast<uint32_t*>(s_k + tok * HP + head * DH + dt * 8 + tg * 2) =
              pack_bf16(k0 * c - k1 * s, k1 * c + k0 * s);
          const float q0 = aq[dt][hf * 2] * qscale, q1 = aq[dt][hf * 2 + 1] * qscale;
          qa[mi][dt / 2][(dt % 2) * 2 + hf] = pack_bf16(q0 * c - q1 * s, q1 * c + q0 * s);
          *reinterpret_cast<uint32_t*>(s_v + tok * HP + head * DH + dt * 8 + tg * 2) =
              pack_bf16(av[dt][hf * 2], av[dt][hf * 2 + 1]);
        }
      }
    }
    // the two warps of this head exchange K / V halves
    asm volatile("bar.sync %0, 64;" ::"r"(1 + head) : "memory");
    uint32_t kfrag[NT][KS][2];
#pragma unroll
    for (int np = 0; np < NT / 2; ++np)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t kf[4];
        ldsm_x4(kf, s_k + (np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * HP + head * DH + ks * 16 +
                        ((lane >> 3) & 1) * 8);
        kfrag[2 * np][ks][0] = kf[0]; kfrag[2 * np][ks][1] = kf[1];
        kfrag[2 * np + 1][ks][0] = kf[2]; kfrag[2 * np + 1][ks][1] = kf[3];
      }

    const __nv_bfloat16* bias_h = s_bias + head * NTOK * BP;
#pragma unroll
    for (int mi = 0; mi < MH; ++mi) {
      const int mt = half * MH + mi;
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      float s[NT][4];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bb[4];
        ldsm_x4(bb, bias_h + (mt * 16 + lrow) * BP + np * 16 + lcol);
        s[2 * np][0] = bf_lo(bb[0]); s[2 * np][1] = bf_hi(bb[0]);
        s[2 * np][2] = bf_lo(bb[1]); s[2 * np][3] = bf_hi(bb[1]);
        s[2 * np + 1][0] = bf_lo(bb[2]); s[2 * np + 1][1] = bf_hi(bb[2]);
        s[2 * np + 1][2] = bf_lo(bb[3]); s[2 * np + 1][3] = bf_hi(bb[3]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma16816(s[nt], qa[mi][ks], kfrag[nt][ks][0], kfrag[nt][ks][1]);
      if (has_mask) {
        const int c0 = region_code(win, r0), c1 = region_code(win, r1);
        uint32_t e0[NTOK / 32], e1[NTOK / 32];
#pragma unroll
        for (int w = 0; w < NTOK / 32; ++w) { e0[w] = ~s_E[w * 8 + c0]; e1[w] = ~s_E[w * 8 + c1]; }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int bit = (nt * 8) % 32 + tg * 2;
          const uint32_t m0 = e0[nt / 4] >> bit, m1 = e1[nt / 4] >> bit;
          if (m0 & 1) s[nt][0] += kMask;
          if (m0 & 2) s[nt][1] += kMask;
          if (m1 & 1) s[nt][2] += kMask;
          if (m1 & 2) s[nt][3] += kMask;
        }
      }
      float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
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
        s[nt][0] = fast_exp2(s[nt][0] - m0);
        s[nt][1] = fast_exp2(s[nt][1] - m0);
        s[nt][2] = fast_exp2(s[nt][2] - m1);
        s[nt][3] = fast_exp2(s[nt][3] - m1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float inv0 = __frcp_rn(l0), inv1 = __frcp_rn(l1);
      float o[DT][4];
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
      for (int ps = 0; ps < NTOK / 16; ++ps) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * ps][0], s[2 * ps][1]);
        a[1] = pack_bf16(s[2 * ps][2], s[2 * ps][3]);
        a[2] = pack_bf16(s[2 * ps + 1][0], s[2 * ps + 1][1]);
        a[3] = pack_bf16(s[2 * ps + 1][2], s[2 * ps + 1][3]);
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {
          uint32_t vb[4];
          ldsm_x4_trans(vb, s_v + (ps * 16 + lrow) * HP + head * DH + dp * 16 + lcol);
          mma16816(o[2 * dp], a, vb[0], vb[1]);
          mma16816(o[2 * dp + 1], a, vb[2], vb[3]);
        }
      }
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        *reinterpret_cast<uint32_t*>(s_o + r0 * HP + head * DH + dt * 8 + tg * 2) =
            pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        *reinterpret_cast<uint32_t*>(s_o + r1 * HP + head * DH + dt * 8 + tg * 2) =
            pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
      }
    }
    __syncthreads();                                       // S3: s_o complete; s_xn dead

    // ---- output projection: warp = (n-tile = warp & 7 [C = 64: 8 n-tiles], m-tiles of its half)
    {
      static_assert(C == 64, "16-warp variant: C = 64");
      const int ntile = warp & 7;
      float acc[MH][4];
#pragma unroll
      for (int mi = 0; mi < MH; ++mi) acc[mi][0] = acc[mi][1] = acc[mi][2] = acc[mi][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < HID / 32; ++kp) {
        uint32_t b[4];
        ldsm_x4(b, s_wproj + (ntile * 8 + (lane & 7)) * HP + kp * 32 + (lane >> 3) * 8);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int mi = 0; mi < MH; ++mi) {
            uint32_t a[4];
            ldsm_x4(a, s_o + ((half * MH + mi) * 16 + lrow) * HP + (kp * 2 + kk) * 16 + lcol);
            mma16816(acc[mi], a, b[kk * 2], b[kk * 2 + 1]);
          }
      }
#pragma unroll
      for (int mi = 0; mi < MH; ++mi) {
        const int col = ntile * 8 + tg * 2;
        const float b0 = s_pbias[col], b1 = s_pbias[col + 1];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int tok = (half * MH + mi) * 16 + hf * 8 + g;
          const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t*>(raw + tok * XP + col));
          *reinterpret_cast<uint32_t*>(s_xn + tok * XP + col) =
              pack_bf16(acc[mi][hf * 2] + b0 + r.x, acc[mi][hf * 2 + 1] + b1 + r.y);
        }
      }
    }
    __syncthreads();                                       // S4
    for (int i = tid; i < NTOK * (C / 8); i += NTH) {
      const int n = i / (C / 8), c8 = i % (C / 8);
      const int d = src_pixel(win, n);
      if (d >= 0)
        *reinterpret_cast<uint4*>(p.y + static_cast<long long>(d) * C + c8 * 8) =
            *reinterpret_cast<const uint4*>(s_xn + n * XP + c8 * 8);
    }
    buf ^= 1;
  }
  cp_async_wait<0>();
}

template <int NTOK, int DH, int C>
static int launch_stw16(const StwParams& p, cudaStream_t st) {
  using L = Stw16Smem<NTOK, DH, C>;
  constexpr size_t smem = L::total;
  static SmemConfigured configured;
  const int sms = device_sm_count();
  if (!configured.covers(smem)) {
    cudaError_t e = cudaFuncSetAttribute(stw_fused16_kernel<NTOK, DH, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem);
  }
  const int grid = p.n_windows < sms ? p.n_windows : sms;
  stw_fused16_kernel<NTOK, DH, C><<<grid, 512, smem, st>>>(p);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

}  // namespace extdm

using namespace extdm;

// tcgen05 edition of the C = 64, (4,4,4)-window layer (stw_tc.cu)
int extdm_stw_tc_launch(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                        const float* proj_bias, const float* bias_table, const float* rope_cos, const float* rope_sin,
                        int B, int T, int H, int W, int sd, int sh, int sw, float eps, void* stream);

// tcgen05 edition of the dim_head-32 layers, C = 64 (attn_ws32.cu): (2,4,4) windows and temporal sequences; warp-specialised
// (drain / softmax / issue / load roles), softmax output kept in tensor memory
int extdm_stw_ws32_launch(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                          const float* proj_bias, const float* bias_table, const float* rope_cos, const float* rope_sin,
                          int B, int T, int H, int W, int C, int sd, int sh, int sw, float eps, void* stream);
int extdm_temporal_ws32_launch(const void* x, void* y, const float* gamma, const float* ln_w, const float* ln_b,
                               const void* wqkv, const void* wout, const float* rel_bias, const float* rope_cos,
                               const float* rope_sin, int B, int T, int HW, int C, float eps, void* stream);

extern "C" int extdm_stw_fused_supported(int C, int heads, int dh, int wd, int wh, int ww) {
  const int ntok = wd * wh * ww;
  if (heads != 8 || wh != 4 || ww != 4) return 0;
  if (ntok == 64 && dh == 16 && (C == 64 || C == 128)) return 1;
  if (ntok == 32 && dh == 32 && C == 64) return 1;
  return 0;
}

static int stw_fused_impl(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                          const float* proj_bias, const float* bias_table, const float* rope_cos,
                          const float* rope_sin, const void* pre_res, const float* pre_ad, int B, int T, int H, int W,
                          int C, int heads, int dh, int wd, int wh, int ww, int sd, int sh, int sw, float eps,
                          void* stream) {
  if (!extdm_stw_fused_supported(C, heads, dh, wd, wh, ww) || H % wh || W % ww || sd < 0 || sd >= wd || sh < 0 ||
      sh >= wh || sw < 0 || sw >= ww) {
    extdm_set_error("stw_fused: unsupported (C, heads, dh, window, shift) combination", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  StwParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.gamma = gamma;
  p.wqkv = reinterpret_cast<const __nv_bfloat16*>(wqkv);
  p.wproj = reinterpret_cast<const __nv_bfloat16*>(wproj);
  p.proj_bias = proj_bias;
  p.bias_table = bias_table;
  p.rcos = rope_cos;
  p.rsin = rope_sin;
  p.ln_w = p.ln_b = nullptr;
  p.pre_res = reinterpret_cast<const __nv_bfloat16*>(pre_res);
  p.pre_ad = pre_ad;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.sd = sd; p.sh = sh; p.sw = sw;
  p.Dp = (T + wd - 1) / wd * wd;
  p.n_windows = B * (p.Dp / wd) * (H / wh) * (W / ww);
  p.eps = eps;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ntok = wd * wh * ww;
  // Three implementations of the C = 64 / 64-token layer, all parity-tested (tests/test_kernels_gpu.py runs each):
  //   default       projections on tcgen05 (stw_tc.cu), two windows per M = 128 tile; level-0 launch on B200: see
  //                 DESIGN.md section 5 (needs H/4 and W/4 powers of two, else the 16-warp kernel runs)
  //   EXTDM_STW16   all-mma.sync, 16 warps            716 us per level-0 launch (legacy-HMMA-pipe bound)
  //   EXTDM_STW8    all-mma.sync, 8 warps             716 us
  static const bool use8 = getenv("EXTDM_STW8") != nullptr, use16 = getenv("EXTDM_STW16") != nullptr;
  const int nww_ = W / 4, nwh_ = H / 4;
  const bool use_tc = !use8 && !use16 && !(nww_ & (nww_ - 1)) && !(nwh_ & (nwh_ - 1));
  if (pre_res) {
    if (!(ntok == 64 && C == 64)) {
      extdm_set_error("stw_fused_pre: the fused GroupNorm producer exists for the 16-warp C = 64 kernel only", __FILE__,
                      __LINE__);
      return EXTDM_ERR_ARG;
    }
    return launch_stw16<64, 16, 64>(p, st);
  }
  if (ntok == 64 && C == 64) {
    if (use8) return launch_stw<64, 16, 64>(p, st);
    if (use_tc)
      return extdm_stw_tc_launch(x, y, gamma, wqkv, wproj, proj_bias, bias_table, rope_cos, rope_sin, B, T, H, W, sd,
                                 sh, sw, eps, stream);
    return launch_stw16<64, 16, 64>(p, st);
  }
  if (ntok == 64 && C == 128) return launch_stw<64, 16, 128>(p, st);
  // (2,4,4) windows, 8 heads x 32: every product on tcgen05 (attn_ws32.cu).  B200, BAIR level 0 (393 k tokens): 243 us vs
  // 345 us for the mma.sync kernel below, which stays as the fallback for geometries the tcgen05 kernel does not take
  // (H/4 or W/4 not a power of two) and as A/B partner behind EXTDM_ATTN32_LEGACY=1.
  static const bool legacy32 = getenv("EXTDM_ATTN32_LEGACY") != nullptr;
  if (!legacy32) {
    const int rc = extdm_stw_ws32_launch(x, y, gamma, wqkv, wproj, proj_bias, bias_table, rope_cos, rope_sin, B, T, H, W,
                                         C, sd, sh, sw, eps, stream);
    if (rc != -1) return rc;
  }
  return launch_stw<32, 32, 64>(p, st);       // (2,4,4) windows: the 16-warp layout would need 233 KB of shared memory
}

extern "C" int extdm_stw_fused(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                               const float* proj_bias, const float* bias_table, const float* rope_cos,
                               const float* rope_sin, int B, int T, int H, int W, int C, int heads, int dh, int wd,
                               int wh, int ww, int sd, int sh, int sw, float eps, void* stream) {
  return stw_fused_impl(x, y, gamma, wqkv, wproj, proj_bias, bias_table, rope_cos, rope_sin, nullptr, nullptr, B, T, H, W,
                        C, heads, dh, wd, wh, ww, sd, sh, sw, eps, stream);
}

extern "C" int extdm_stw_fused_pre_supported(int C, int heads, int dh, int wd, int wh, int ww) {
  return heads == 8 && dh == 16 && C == 64 && wd == 4 && wh == 4 && ww == 4;
}

extern "C" int extdm_stw_fused_pre(const void* h, const void* res, const float* ad, void* y, const float* gamma,
                                   const void* wqkv, const void* wproj, const float* proj_bias, const float* bias_table,
                                   const float* rope_cos, const float* rope_sin, int B, int T, int H, int W, int C,
                                   int heads, int dh, int wd, int wh, int ww, int sd, int sh, int sw, float eps,
                                   void* stream) {
  if (!res || !ad) {
    extdm_set_error("stw_fused_pre: residual and affine table required", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  return stw_fused_impl(h, y, gamma, wqkv, wproj, proj_bias, bias_table, rope_cos, rope_sin, res, ad, B, T, H, W, C,
                        heads, dh, wd, wh, ww, sd, sh, sw, eps, stream);
}

extern "C" int extdm_temporal_fused_supported(int C, int heads, int dh, int T) {
  return heads == 8 && (dh == 16 || dh == 32) && C == 64 && T >= 1 && T <= 32;
}

extern "C" int extdm_temporal_fused(const void* x, void* y, const float* gamma, const float* ln_w, const float* ln_b,
                                    const void* wqkv, const void* wout, const float* rel_bias, const float* rope_cos,
                                    const float* rope_sin, int B, int T, int HW, int C, int heads, int dh, float eps,
                                    void* stream) {
  if (!extdm_temporal_fused_supported(C, heads, dh, T)) {
    extdm_set_error("temporal_fused: supported for C = 64, 8 heads x 16 or 32, T <= 32", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  StwParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.gamma = gamma;
  p.wqkv = reinterpret_cast<const __nv_bfloat16*>(wqkv);
  p.wproj = reinterpret_cast<const __nv_bfloat16*>(wout);
  p.proj_bias = nullptr;
  p.bias_table = rel_bias;
  p.rcos = rope_cos;
  p.rsin = rope_sin;
  p.ln_w = ln_w;
  p.ln_b = ln_b;
  p.pre_res = nullptr;
  p.pre_ad = nullptr;
  p.B = B; p.T = T; p.H = HW; p.W = 1;
  p.sd = p.sh = p.sw = 0;
  p.Dp = 32;
  p.n_windows = B * HW;
  p.eps = eps;
  // dim_head 32 (u12 / base / ada_u22): tcgen05 kernel (attn_ws32.cu); the mma.sync one behind EXTDM_ATTN32_LEGACY
  if (dh == 32) {
    static const bool legacy32 = getenv("EXTDM_ATTN32_LEGACY") != nullptr;
    if (!legacy32) {
      const int rc = extdm_temporal_ws32_launch(x, y, gamma, ln_w, ln_b, wqkv, wout, rel_bias, rope_cos, rope_sin, B, T,
                                                HW, C, eps, stream);
      if (rc != -1) return rc;
    }
    return launch_stw<32, 32, 64, true>(p, static_cast<cudaStream_t>(stream));
  }
  return launch_stw<32, 16, 64, true>(p, static_cast<cudaStream_t>(stream));
}
