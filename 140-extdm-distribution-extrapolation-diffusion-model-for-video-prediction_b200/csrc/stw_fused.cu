// Fully fused shifted-window attention layer for the high-resolution UNet levels (C = 64 / 128):
//     y = x + proj( WindowAttention3D( chanLN(x) ) )          Residual(PreNorm(STWAttentionLayer))
// reference: model/BaseDM_adaptor/DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:139-159 (PreNorm/LayerNorm),
// :409-497 (WindowAttention3D), :499-560 (STWAttentionLayer: pad / roll / partition / mask / reverse).
//
// Un-fused, one layer moves x, LN(x), qkv (6x the size of x), attn-out and y through HBM (~21x |x| of traffic);
// fused, it reads x once and writes y once.  Persistent CTAs (one per SM) keep Wqkv / Wproj / the bias table in
// shared memory and loop over windows; the next window's tokens are prefetched with cp.async while the current
// one is computed.  One warp per head: K/V/Q projections, rotary, QK^T, mask + bias, softmax and PV stay in
// registers (mma.sync m16n8k16 bf16, accumulator->operand fragment re-use); only V, the head outputs and the
// projected tile pass through shared memory.  (64-token x 64..128-channel tiles are below a tcgen05 tile of
// M = 128 rows per CTA; the window is the natural unit, so the warp-level tensor path is used here.)
#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
  int B, T, H, W, wd, wh, ww, sd, sh, sw, Dp, n_windows;
  float eps;
};

template <int NTOK, int DH, int C>
struct StwSmem {
  static constexpr int HEADS = 8;
  static constexpr int HID = HEADS * DH;
  static constexpr int XP = C + 8;        // pitch (bf16) of token x channel tiles
  static constexpr int HP = HID + 8;      // pitch of token x hidden tiles
  static constexpr size_t wqkv = 0;
  static constexpr size_t wproj = wqkv + size_t(3 * HID) * XP * 2;
  static constexpr bool DB = (C <= 64);   // double-buffered token prefetch when it fits in 227 KB
  static constexpr size_t raw = wproj + size_t(C) * HP * 2;            // (DB ? 2 : 1) x [NTOK][C] raw tokens (cp.async)
  static constexpr size_t xn = raw + (DB ? 2 : 1) * size_t(NTOK) * C * 2;   // [NTOK][XP] normalised tokens / staging
  static constexpr size_t v = xn + size_t(NTOK) * XP * 2;              // [NTOK][HP]
  static constexpr size_t o = v + size_t(NTOK) * HP * 2;               // [NTOK][HP]
  static constexpr size_t rope = o + size_t(NTOK) * HP * 2;            // cos, sin [NTOK][DH/2] fp32
  static constexpr size_t misc = rope + 2 * size_t(NTOK) * (DH / 2) * 4;   // gamma[C], pbias[C]
  static constexpr size_t idx = misc + 2 * size_t(C) * 4;              // 2 x { src[NTOK] (int), lin[NTOK], reg[NTOK] }
  static constexpr size_t tbl = idx + 2 * 3 * size_t(NTOK) * 4;        // [HEADS][tbl_n] fp32 (dynamic length)
};

template <int NTOK, int DH, int C>
__global__ void __launch_bounds__(256, 1) stw_fused_kernel(const __grid_constant__ StwParams p) {
  using L = StwSmem<NTOK, DH, C>;
  constexpr int HEADS = 8, HID = L::HID, XP = L::XP, HP = L::HP;
  constexpr int MT = NTOK / 16;          // query / token m-tiles
  constexpr int DT = DH / 8;             // n-tiles over the head dim
  constexpr int KS = DH / 16;            // k-steps of Q K^T
  constexpr int NT = NTOK / 8;           // key n-tiles
  constexpr int CK = C / 16;             // k-steps of the projections from C
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* s_wqkv = reinterpret_cast<__nv_bfloat16*>(sm + L::wqkv);
  __nv_bfloat16* s_wproj = reinterpret_cast<__nv_bfloat16*>(sm + L::wproj);
  __nv_bfloat16* s_raw = reinterpret_cast<__nv_bfloat16*>(sm + L::raw);
  __nv_bfloat16* s_xn = reinterpret_cast<__nv_bfloat16*>(sm + L::xn);
  __nv_bfloat16* s_v = reinterpret_cast<__nv_bfloat16*>(sm + L::v);
  __nv_bfloat16* s_o = reinterpret_cast<__nv_bfloat16*>(sm + L::o);
  float* s_cos = reinterpret_cast<float*>(sm + L::rope);
  float* s_sin = s_cos + NTOK * (DH / 2);
  float* s_gamma = reinterpret_cast<float*>(sm + L::misc);
  float* s_pbias = s_gamma + C;
  int* s_idx = reinterpret_cast<int*>(sm + L::idx);
  float* s_tbl = reinterpret_cast<float*>(sm + L::tbl);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int tbl_n = (2 * p.wd - 1) * (2 * p.wh - 1) * (2 * p.ww - 1);
  const bool shifted = (p.sd | p.sh | p.sw) != 0;
  const int nWw = p.W / p.ww, nWh = p.H / p.wh, nWd = p.Dp / p.wd;

  // ---- one-time staging of the weights and tables
  for (int i = tid; i < 3 * HID * (C / 8); i += 256) {
    const int r = i / (C / 8), c8 = i % (C / 8);
    *reinterpret_cast<uint4*>(s_wqkv + r * XP + c8 * 8) = *reinterpret_cast<const uint4*>(p.wqkv + r * C + c8 * 8);
  }
  for (int i = tid; i < C * (HID / 8); i += 256) {
    const int r = i / (HID / 8), c8 = i % (HID / 8);
    *reinterpret_cast<uint4*>(s_wproj + r * HP + c8 * 8) = *reinterpret_cast<const uint4*>(p.wproj + r * HID + c8 * 8);
  }
  for (int i = tid; i < HEADS * tbl_n; i += 256) s_tbl[i] = p.bias_table[(i % tbl_n) * HEADS + i / tbl_n];
  for (int i = tid; i < NTOK * (DH / 2); i += 256) { s_cos[i] = p.rcos[i]; s_sin[i] = p.rsin[i]; }
  for (int i = tid; i < C; i += 256) { s_gamma[i] = p.gamma[i]; s_pbias[i] = p.proj_bias[i]; }

  // per-window token bookkeeping (source pixel, bias-table linear index, mask region) into buffer `buf`
  auto index_window = [&](int widx, int buf) {
    if (tid < NTOK) {
      int* src = s_idx + buf * 3 * NTOK;
      int* lin = src + NTOK;
      int* reg = lin + NTOK;
      int w_ = widx;
      const int iw = w_ % nWw; w_ /= nWw;
      const int ih = w_ % nWh; w_ /= nWh;
      const int id = w_ % nWd; w_ /= nWd;
      const int b = w_;
      const int n = tid;
      const int tw = n % p.ww, th = (n / p.ww) % p.wh, td = n / (p.ww * p.wh);
      lin[n] = (td * (2 * p.wh - 1) + th) * (2 * p.ww - 1) + tw;
      const int zd = id * p.wd + td, zh = ih * p.wh + th, zw = iw * p.ww + tw;
      int rd = 0, rh = 0, rw = 0;
      if (p.sd) rd = zd < p.Dp - p.wd ? 0 : (zd < p.Dp - p.sd ? 1 : 2);
      if (p.sh) rh = zh < p.H - p.wh ? 0 : (zh < p.H - p.sh ? 1 : 2);
      if (p.sw) rw = zw < p.W - p.ww ? 0 : (zw < p.W - p.sw ? 1 : 2);
      reg[n] = (rd * 3 + rh) * 3 + rw;
      const int od = (zd + p.sd) % p.Dp, oh = (zh + p.sh) % p.H, ow = (zw + p.sw) % p.W;
      src[n] = od < p.T ? ((b * p.T + od) * p.H + oh) * p.W + ow : -1;
    }
  };
  auto prefetch_window = [&](int buf) {
    const int* src = s_idx + buf * 3 * NTOK;
    __nv_bfloat16* dst = s_raw + buf * NTOK * C;
    for (int i = tid; i < NTOK * (C / 8); i += 256) {
      const int n = i / (C / 8), c8 = i % (C / 8);
      const int s = src[n];
      cp_async16(dst + n * C + c8 * 8, p.x + (s >= 0 ? static_cast<long long>(s) * C + c8 * 8 : 0), s >= 0 ? 16 : 0);
    }
    cp_async_commit();
  };

  int widx = blockIdx.x;
  int buf = 0;
  if (widx < p.n_windows) index_window(widx, 0);
  __syncthreads();
  if (widx < p.n_windows) prefetch_window(0);

  const float qscale = rsqrtf(static_cast<float>(DH));
  const int c0_tbl = ((p.wd - 1) * (2 * p.wh - 1) + (p.wh - 1)) * (2 * p.ww - 1) + (p.ww - 1);

  constexpr bool DB = L::DB;
  for (; widx < p.n_windows; widx += gridDim.x) {
    const int nxt = widx + gridDim.x;
    if (DB && nxt < p.n_windows) index_window(nxt, buf ^ 1);
    cp_async_wait<0>();
    __syncthreads();                                       // raw[buf] landed; idx[buf^1] visible
    if (DB && nxt < p.n_windows) prefetch_window(buf ^ 1);
    const int* s_src = s_idx + buf * 3 * NTOK;
    const int* s_lin = s_src + NTOK;
    const int* s_reg = s_lin + NTOK;
    const __nv_bfloat16* raw = s_raw + buf * NTOK * C;

    // ---- channel LayerNorm (biased variance, gamma only); padding tokens stay exactly zero
    for (int n = warp; n < NTOK; n += 8) {
      constexpr int V = C / 32;
      float v[V];
      if constexpr (V == 2) {
        const float2 t = unpack_bf16(*reinterpret_cast<const uint32_t*>(raw + n * C + lane * 2));
        v[0] = t.x; v[1] = t.y;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(raw + n * C + lane * 4);
        const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) sum += v[j];
      const float mean = warp_sum(sum) * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) { const float d = v[j] - mean; sq += d * d; }
      const float rstd = s_src[n] >= 0 ? rsqrtf(warp_sum(sq) * (1.0f / C) + p.eps) : 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = (v[j] - mean) * rstd * s_gamma[lane * V + j];
      if constexpr (V == 2) {
        *reinterpret_cast<uint32_t*>(s_xn + n * XP + lane * 2) = pack_bf16(v[0], v[1]);
      } else {
        uint2 t;
        t.x = pack_bf16(v[0], v[1]); t.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(s_xn + n * XP + lane * 4) = t;
      }
    }
    __syncthreads();

    // ---- per-head projections: K, V for every token (registers), then Q per m-tile
    const int head = warp;
    const __nv_bfloat16* wq = s_wqkv + (head * DH) * XP;
    const __nv_bfloat16* wk = s_wqkv + (HID + head * DH) * XP;
    const __nv_bfloat16* wv = s_wqkv + (2 * HID + head * DH) * XP;
    uint32_t kfrag[NT][KS][2];
    {
      float ak[MT][DT][4], av[MT][DT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int dt = 0; dt < DT; ++dt)
#pragma unroll
          for (int j = 0; j < 4; ++j) { ak[mt][dt][j] = 0.f; av[mt][dt][j] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < CK; ++ks) {
        uint32_t bk[DT][2], bv[DT][2];
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
          const __nv_bfloat16* rk = wk + (dt * 8 + g) * XP + ks * 16 + tg * 2;
          const __nv_bfloat16* rv = wv + (dt * 8 + g) * XP + ks * 16 + tg * 2;
          bk[dt][0] = *reinterpret_cast<const uint32_t*>(rk);
          bk[dt][1] = *reinterpret_cast<const uint32_t*>(rk + 8);
          bv[dt][0] = *reinterpret_cast<const uint32_t*>(rv);
          bv[dt][1] = *reinterpret_cast<const uint32_t*>(rv + 8);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a[4];
          const __nv_bfloat16* xr = s_xn + (mt * 16 + g) * XP + ks * 16 + tg * 2;
          a[0] = *reinterpret_cast<const uint32_t*>(xr);
          a[1] = *reinterpret_cast<const uint32_t*>(xr + 8 * XP);
          a[2] = *reinterpret_cast<const uint32_t*>(xr + 8);
          a[3] = *reinterpret_cast<const uint32_t*>(xr + 8 * XP + 8);
#pragma unroll
          for (int dt = 0; dt < DT; ++dt) {
            mma16816(ak[mt][dt], a, bk[dt][0], bk[dt][1]);
            mma16816(av[mt][dt], a, bv[dt][0], bv[dt][1]);
          }
        }
      }
      // rotary on K (pair = the two accumulator columns a thread owns), K -> B fragments, V -> smem
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
          const int pr = dt * 4 + tg;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int tok = mt * 16 + hf * 8 + g;
            const float c = s_cos[tok * (DH / 2) + pr], s = s_sin[tok * (DH / 2) + pr];
            const float x0 = ak[mt][dt][hf * 2], x1 = ak[mt][dt][hf * 2 + 1];
            // B fragment of key n-tile (2*mt + hf): k-step dt/2, register dt%2
            kfrag[2 * mt + hf][dt / 2][dt % 2] = pack_bf16(x0 * c - x1 * s, x1 * c + x0 * s);
            *reinterpret_cast<uint32_t*>(s_v + tok * HP + head * DH + dt * 8 + tg * 2) =
                pack_bf16(av[mt][dt][hf * 2], av[mt][dt][hf * 2 + 1]);
          }
        }
      }
    }
    __syncwarp();
    const float* tb = s_tbl + head * tbl_n;
    const unsigned short* Vs = reinterpret_cast<const unsigned short*>(s_v + head * DH);
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      float aq[DT][4];
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) { aq[dt][0] = aq[dt][1] = aq[dt][2] = aq[dt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < CK; ++ks) {
        uint32_t a[4];
        const __nv_bfloat16* xr = s_xn + (mt * 16 + g) * XP + ks * 16 + tg * 2;
        a[0] = *reinterpret_cast<const uint32_t*>(xr);
        a[1] = *reinterpret_cast<const uint32_t*>(xr + 8 * XP);
        a[2] = *reinterpret_cast<const uint32_t*>(xr + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(xr + 8 * XP + 8);
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
          const __nv_bfloat16* rq = wq + (dt * 8 + g) * XP + ks * 16 + tg * 2;
          mma16816(aq[dt], a, *reinterpret_cast<const uint32_t*>(rq), *reinterpret_cast<const uint32_t*>(rq + 8));
        }
      }
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      uint32_t qa[KS][4];
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        const int pr = dt * 4 + tg;
        const float c0 = s_cos[r0 * (DH / 2) + pr], s0 = s_sin[r0 * (DH / 2) + pr];
        const float c1 = s_cos[r1 * (DH / 2) + pr], s1 = s_sin[r1 * (DH / 2) + pr];
        const float x0 = aq[dt][0] * qscale, x1 = aq[dt][1] * qscale, y0 = aq[dt][2] * qscale, y1 = aq[dt][3] * qscale;
        // A fragment of k-step dt/2: registers (dt%2)*2 + {0: row g, 1: row g+8}
        qa[dt / 2][(dt % 2) * 2 + 0] = pack_bf16(x0 * c0 - x1 * s0, x1 * c0 + x0 * s0);
        qa[dt / 2][(dt % 2) * 2 + 1] = pack_bf16(y0 * c1 - y1 * s1, y1 * c1 + y0 * s1);
      }
      float s[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma16816(s[nt], qa[ks], kfrag[nt][ks][0], kfrag[nt][ks][1]);
      }
      const int li0 = s_lin[r0] + c0_tbl, li1 = s_lin[r1] + c0_tbl, rg0 = s_reg[r0], rg1 = s_reg[r1];
      float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + tg * 2 + e;
          const int lj = s_lin[j], rj = s_reg[j];
          float b0 = tb[li0 - lj], b1 = tb[li1 - lj];
          if (shifted) {
            if (rg0 != rj) b0 += -100.0f;
            if (rg1 != rj) b1 += -100.0f;
          }
          s[nt][e] += b0;
          s[nt][2 + e] += b1;
        }
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
      for (int ps = 0; ps < NTOK / 16; ++ps) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * ps][0], s[2 * ps][1]);
        a[1] = pack_bf16(s[2 * ps][2], s[2 * ps][3]);
        a[2] = pack_bf16(s[2 * ps + 1][0], s[2 * ps + 1][1]);
        a[3] = pack_bf16(s[2 * ps + 1][2], s[2 * ps + 1][3]);
        const int k0 = ps * 16 + tg * 2;
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
          const int col = dt * 8 + g;
          const uint32_t b0 = static_cast<uint32_t>(Vs[k0 * HP + col]) |
                              (static_cast<uint32_t>(Vs[(k0 + 1) * HP + col]) << 16);
          const uint32_t b1 = static_cast<uint32_t>(Vs[(k0 + 8) * HP + col]) |
                              (static_cast<uint32_t>(Vs[(k0 + 9) * HP + col]) << 16);
          mma16816(o[dt], a, b0, b1);
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
    __syncthreads();                                       // all heads' outputs in s_o; s_xn is dead

    // ---- output projection + bias + residual -> s_xn (staging), then coalesced stores
    {
      constexpr int NPW = C / 64;                          // n-tiles per warp (C/8 n-tiles over 8 warps)
      float acc[MT][NPW][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int q = 0; q < NPW; ++q) { acc[mt][q][0] = acc[mt][q][1] = acc[mt][q][2] = acc[mt][q][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < HID / 16; ++ks) {
        uint32_t b[NPW][2];
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
          const __nv_bfloat16* rw = s_wproj + ((warp * NPW + q) * 8 + g) * HP + ks * 16 + tg * 2;
          b[q][0] = *reinterpret_cast<const uint32_t*>(rw);
          b[q][1] = *reinterpret_cast<const uint32_t*>(rw + 8);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a[4];
          const __nv_bfloat16* orow = s_o + (mt * 16 + g) * HP + ks * 16 + tg * 2;
          a[0] = *reinterpret_cast<const uint32_t*>(orow);
          a[1] = *reinterpret_cast<const uint32_t*>(orow + 8 * HP);
          a[2] = *reinterpret_cast<const uint32_t*>(orow + 8);
          a[3] = *reinterpret_cast<const uint32_t*>(orow + 8 * HP + 8);
#pragma unroll
          for (int q = 0; q < NPW; ++q) mma16816(acc[mt][q], a, b[q][0], b[q][1]);
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
            const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t*>(raw + tok * C + col));
            *reinterpret_cast<uint32_t*>(s_xn + tok * XP + col) =
                pack_bf16(acc[mt][q][hf * 2] + b0 + r.x, acc[mt][q][hf * 2 + 1] + b1 + r.y);
          }
        }
    }
    __syncthreads();
    for (int i = tid; i < NTOK * (C / 8); i += 256) {
      const int n = i / (C / 8), c8 = i % (C / 8);
      const int d = s_src[n];
      if (d >= 0)
        *reinterpret_cast<uint4*>(p.y + static_cast<long long>(d) * C + c8 * 8) =
            *reinterpret_cast<const uint4*>(s_xn + n * XP + c8 * 8);
    }
    __syncthreads();                                       // s_xn / raw / idx of this window are dead from here on
    if (DB) {
      buf ^= 1;
    } else if (nxt < p.n_windows) {
      index_window(nxt, 0);
      __syncthreads();
      prefetch_window(0);
    }
  }
  cp_async_wait<0>();
}

template <int NTOK, int DH, int C>
static int launch_stw(const StwParams& p, int tbl_n, cudaStream_t st) {
  using L = StwSmem<NTOK, DH, C>;
  const size_t smem = L::tbl + static_cast<size_t>(8) * tbl_n * 4;
  static size_t configured = 0;
  static int sms = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(stw_fused_kernel<NTOK, DH, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured = smem;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = p.n_windows < sms ? p.n_windows : sms;
  stw_fused_kernel<NTOK, DH, C><<<grid, 256, smem, st>>>(p);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

}  // namespace extdm

using namespace extdm;

extern "C" int extdm_stw_fused_supported(int C, int heads, int dh, int wd, int wh, int ww) {
  const int ntok = wd * wh * ww;
  if (heads != 8) return 0;
  if (ntok == 64 && dh == 16 && (C == 64 || C == 128)) return 1;
  if (ntok == 32 && dh == 32 && C == 64) return 1;
  return 0;
}

extern "C" int extdm_stw_fused(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                               const float* proj_bias, const float* bias_table, const float* rope_cos,
                               const float* rope_sin, int B, int T, int H, int W, int C, int heads, int dh, int wd,
                               int wh, int ww, int sd, int sh, int sw, float eps, void* stream) {
  if (!extdm_stw_fused_supported(C, heads, dh, wd, wh, ww) || H % wh || W % ww) {
    extdm_set_error("stw_fused: unsupported (C, heads, dh, window) combination", __FILE__, __LINE__);
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
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.wd = wd; p.wh = wh; p.ww = ww; p.sd = sd; p.sh = sh; p.sw = sw;
  p.Dp = (T + wd - 1) / wd * wd;
  p.n_windows = B * (p.Dp / wd) * (H / wh) * (W / ww);
  p.eps = eps;
  const int tbl_n = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ntok = wd * wh * ww;
  if (ntok == 64 && C == 64) return launch_stw<64, 16, 64>(p, tbl_n, st);
  if (ntok == 64 && C == 128) return launch_stw<64, 16, 128>(p, tbl_n, st);
  return launch_stw<32, 32, 64>(p, tbl_n, st);
}
