// Attention layers of the dim_head-32 UNet variants (u12 = BAIR, base = SMMNIST, ada_u22) with EVERY product on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM): 8 heads x 32, hidden 256.
//
//   window mode    y = x + proj( WindowAttention3D( chanLN(x) ) ) + b        Residual(PreNorm(STWAttentionLayer)), (2,4,4) windows
//                  reference: model/BaseDM_adaptor/DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:139-159, :409-560
//   temporal mode  y = x + chanLN(x) + to_out( Attention( LayerNorm(chanLN(x)) ) )   Residual(PreNorm(EinopsToAndFrom(AttentionLayer)))
//                  reference: ...cross_multi.py:253-328 (double residual: App. B.4 of SURVEY.md)
//
// One M = 128 tile = 4 windows of 32 tokens, or 4 / 8 pixel sequences of T <= 32 / <= 16 frames.  The 8 heads are
// processed in 4 groups of 2 (hidden slice of 64), all of it on UMMA:
//   QKV_g  [128 x 192] = LN(x)[128 x C] . Wqkv_g^T            A = normalised tokens (K-major, SW128), B = weight slice
//   S_h    [128 x 128] = Q_h[128 x 32] . K_h^T                 block-diagonal: only the 32 x 32 (16 x 16) blocks of a
//                                                              row's own window / sequence are read back
//   O_h    [128 x 32]  = P_h[128 x 128] . V_h                  P = softmax(S + bias [+ mask]) in bf16, zero off the diagonal
//                                                              blocks (written once), V stored transposed (K-major B)
//   D      [128 x C]  += O_g[128 x 64] . Wproj[:, g]^T         accumulated over the 4 groups in TMEM
// Weight slices (Wqkv_g | Wproj_g, 32 KB for C = 64) stream through a two-stage cp.async ring from L2; the next tile's
// tokens are prefetched into the A tile as soon as the last QKV product of the current tile has retired.
// TMEM map (512 columns): [0,192) QKV_g, re-used as S_0 [0,128) after the drain; S_1 [192,320); O_g [320,384); D [384,384+C).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {
namespace {

constexpr float kL2e = 1.4426950408889634f;
constexpr int NTH = 512, HEADS = 8, DH = 32, HID = 256;

__device__ __forceinline__ int sw128(int r, int j) { return (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int C>
struct Lay {
  static constexpr int KB = C / 64;                          // 64-channel k-blocks of the layer width
  static constexpr int WQ = KB * 192 * 128;                  // Wqkv_g: KB x [192][64] SW128
  static constexpr int WP = C * 128;                         // Wproj_g: [C][64] SW128
  static constexpr int STAGE = WQ + WP;
  static constexpr int NSTAGE = 2;
  static constexpr int stage = 0;
  static constexpr int a = stage + NSTAGE * STAGE;           // KB x [128][64] SW128: raw tokens, then LN output in place
  static constexpr int q = a + KB * 16384;                   // [128][64] SW128: Q of the group, later its O
  static constexpr int k = q + 16384;                        // [128][64] SW128
  static constexpr int vt = k + 16384;                       // V^T: 2 k-blocks x [64][64] SW128
  static constexpr int p = vt + 16384;                       // 2 heads x 2 k-blocks x [128][64] SW128
  static constexpr int bias = p + 2 * 32768;                 // [8][TP][TP] bf16 * log2(e), chunk-swizzled
  static constexpr int rope = bias + 16384;                  // cos, sin [16 pairs][32 positions] fp32 (transposed: lanes = positions)
  static constexpr int vec = rope + 4096;                    // gamma, ln_w, ln_b, proj bias: 4 x C fp32
  static constexpr int stat = vec + 16 * C;                  // [128] (mean, rstd) of the channel LayerNorm
  static constexpr int ml = stat + 1024;                     // [2 heads][128 rows][2 halves] (max, sum)
  static constexpr int bars = ml + 4096;                     // 4 mbarriers + TMEM slot
  static constexpr int total = bars + 64;
};

// EXTDM_ATTN32_PROF=1: per-phase cycle counts of CTA 0 (thread 0's view, waits included), printed by the launcher
__device__ unsigned long long g_prof32[16];

struct P32 {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* gamma;
  const float* ln_w;           // temporal mode: nn.LayerNorm applied after the channel LayerNorm
  const float* ln_b;
  const __nv_bfloat16* wqkv;   // [768][C]
  const __nv_bfloat16* wproj;  // [C][256]
  const float* proj_bias;      // window mode
  const float* bias_table;     // window: [147][8]; temporal: [8][2T-1]
  const float* rcos;
  const float* rsin;           // [32][16]
  int B, T, H, W;              // temporal: H = pixels per frame, W = 1
  int sd, sh, sw, Dp, n_units, n_tiles, lw, lh;
  float eps;
  int prof;
};

// TP = tokens per attention unit in the tile (32: a (2,4,4) window or a sequence of 17..32 frames; 16: T <= 16)
template <int C, int TP, bool TEMPORAL>
__global__ void __launch_bounds__(NTH, 1) attn_tc32_kernel(const __grid_constant__ P32 p) {
  using L = Lay<C>;
  constexpr int KB = L::KB, CH = C / 8;                    // 16-byte chunks per token row
  constexpr int NU = 128 / TP;                             // attention units per tile
  constexpr int NC = TP / 2;                               // score columns per softmax thread
  constexpr uint32_t S0 = 0, S1 = 192, OG = 320, DO = 384;
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = sm + L::a;
  uint8_t* s_q = sm + L::q;
  uint8_t* s_k = sm + L::k;
  uint8_t* s_vt = sm + L::vt;
  uint8_t* s_p = sm + L::p;
  uint8_t* s_bias = sm + L::bias;
  float* s_cos = reinterpret_cast<float*>(sm + L::rope);
  float* s_sin = s_cos + 32 * 16;
  float* s_gamma = reinterpret_cast<float*>(sm + L::vec);
  float* s_lnw = s_gamma + C;
  float* s_lnb = s_lnw + C;
  float* s_pbias = s_lnb + C;
  float2* s_stat = reinterpret_cast<float2*>(sm + L::stat);
  float2* s_ml = reinterpret_cast<float2*>(sm + L::ml);
  uint64_t* bar_qkv = reinterpret_cast<uint64_t*>(sm + L::bars);
  uint64_t* bar_s = bar_qkv + 1;
  uint64_t* bar_o = bar_qkv + 2;
  uint64_t* bar_d = bar_qkv + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qkv + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nWw = TEMPORAL ? 1 : p.W / 4, nWh = TEMPORAL ? 1 : p.H / 4, nWd = TEMPORAL ? 1 : p.Dp / 2;
  const bool shifted = !TEMPORAL && (p.sd | p.sh | p.sw) != 0;

  // ---- geometry: tile row r = unit slot (r / TP), token n = r % TP
  struct Unit { int b, id, ih, iw; };                       // b < 0: no such unit (tail)
  auto decode = [&](int u) {
    Unit w;
    if (u >= p.n_units) { w.b = -1; w.id = w.ih = w.iw = 0; return w; }
    if (TEMPORAL) {                                         // u = (b, pixel)
      w.b = u / p.H;
      w.iw = u - w.b * p.H;
      w.id = w.ih = 0;
      return w;
    }
    w.iw = u & (nWw - 1); u >>= p.lw;
    w.ih = u & (nWh - 1); u >>= p.lh;
    w.b = u / nWd;
    w.id = u - w.b * nWd;
    return w;
  };
  auto src_pixel = [&](const Unit& w, int n) -> int {
    if (w.b < 0) return -1;
    if (TEMPORAL) return n < p.T ? (w.b * p.T + n) * p.H + w.iw : -1;
    int od = w.id * 2 + (n >> 4) + p.sd, oh = w.ih * 4 + ((n >> 2) & 3) + p.sh, ow = w.iw * 4 + (n & 3) + p.sw;
    if (od >= p.Dp) od -= p.Dp;
    if (oh >= p.H) oh -= p.H;
    if (ow >= p.W) ow -= p.W;
    return od < p.T ? ((w.b * p.T + od) * p.H + oh) * p.W + ow : -1;
  };
  auto row_pixel = [&](int tile, int r) -> int { return src_pixel(decode(tile * NU + r / TP), r % TP); };
  auto region_code = [&](const Unit& w, int n) -> int {
    int c = 0;
    if (p.sd && w.id == nWd - 1 && (n >> 4) >= 2 - p.sd) c |= 1;
    if (p.sh && w.ih == nWh - 1 && ((n >> 2) & 3) >= 4 - p.sh) c |= 2;
    if (p.sw && w.iw == nWw - 1 && (n & 3) >= 4 - p.sw) c |= 4;
    return c;
  };
  auto unit_masked = [&](const Unit& w) -> bool {
    return shifted && w.b >= 0 && ((p.sd && w.id == nWd - 1) || (p.sh && w.ih == nWh - 1) || (p.sw && w.iw == nWw - 1));
  };

  // ---- async loaders
  auto load_stage = [&](int g, int buf) {                   // Wqkv rows {q,k,v} x [64g, 64g+64) and Wproj[:, 64g : 64g+64]
    uint8_t* st = sm + L::stage + buf * L::STAGE;
    for (int i = tid; i < 192 * CH; i += NTH) {
      const int r = i / CH, j = i % CH;
      const int grow = (r >> 6) * HID + g * 64 + (r & 63);
      cp_async16(st + (j >> 3) * (192 * 128) + sw128(r, j & 7), p.wqkv + static_cast<size_t>(grow) * C + j * 8, 16);
    }
    for (int i = tid; i < C * 8; i += NTH) {
      const int r = i >> 3, j = i & 7;
      cp_async16(st + L::WQ + sw128(r, j), p.wproj + static_cast<size_t>(r) * HID + g * 64 + j * 8, 16);
    }
    cp_commit();
  };
  auto prefetch_tokens = [&](int tile) {                    // raw bf16 tokens of a tile -> A tile (LayerNorm runs in place)
    for (int i = tid; i < 128 * CH; i += NTH) {
      const int n = i / CH, j = i % CH;
      const int s = row_pixel(tile, n);
      cp_async16(s_a + (j >> 3) * 16384 + sw128(n, j & 7), p.x + (s >= 0 ? static_cast<long long>(s) * C + j * 8 : 0),
                 s >= 0 ? 16 : 0);
    }
    cp_commit();
  };

  // ---- one-time staging
  int tile = blockIdx.x;
  load_stage(0, 0);
  prefetch_tokens(tile);
  for (int i = tid; i < 32 * 16; i += NTH) {                // global [pos][pair] -> shared [pair][pos]
    s_cos[(i & 15) * 32 + (i >> 4)] = p.rcos[i];
    s_sin[(i & 15) * 32 + (i >> 4)] = p.rsin[i];
  }
  for (int i = tid; i < C; i += NTH) {
    s_gamma[i] = p.gamma[i];
    s_lnw[i] = TEMPORAL ? p.ln_w[i] : 1.f;
    s_lnb[i] = TEMPORAL ? p.ln_b[i] : 0.f;
    s_pbias[i] = p.proj_bias ? p.proj_bias[i] : 0.f;
  }
  {
    // additive score bias, expanded to [head][query][key], times log2(e); 16-byte chunks XOR-swizzled against the
    // row-strided reads of the softmax threads
    constexpr int CPR = TP / 8, RPL = 8 / CPR;               // chunks per row, rows per 128 bytes
    for (int idx = tid; idx < HEADS * TP * TP; idx += NTH) {
      const int h = idx / (TP * TP), i = (idx / TP) % TP, j = idx % TP;
      float v;
      if (TEMPORAL) {
        v = (i < p.T && j < p.T) ? p.bias_table[h * (2 * p.T - 1) + (j - i) + p.T - 1] : 0.f;
      } else {
        const int e = ((i >> 4) - (j >> 4) + 1) * 49 + (((i >> 2) & 3) - ((j >> 2) & 3) + 3) * 7 + ((i & 3) - (j & 3) + 3);
        v = p.bias_table[e * HEADS + h];
      }
      const int cs = (j >> 3) ^ ((i / RPL) & (CPR - 1));
      reinterpret_cast<__nv_bfloat16*>(s_bias)[(h * TP + i) * TP + cs * 8 + (j & 7)] = __float2bfloat16(v * kL2e);
    }
  }
  for (int i = tid; i < 2 * 32768 / 16; i += NTH) reinterpret_cast<uint4*>(s_p)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar_qkv, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_d, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const float qscale = 0.17677669529663687f * kL2e;         // dh^-1/2 * log2(e)
  constexpr float kMask = -100.0f * kL2e;
  constexpr uint32_t idesc_qkv = umma_idesc_bf16(128, 192), idesc_s = umma_idesc_bf16(128, 128),
                     idesc_pv = umma_idesc_bf16(128, 32), idesc_d = umma_idesc_bf16(128, C);
  const int dq = warp & 3, cg = warp >> 2;                  // TMEM lane quarter of this warp, column group 0..3
  const int row = dq * 32 + lane;                           // tile row owned in every TMEM drain (thread = token)
  const uint32_t tlane = tmem_u + (static_cast<uint32_t>(dq * 32) << 16);

  uint32_t gc = 0;                                          // groups processed by this CTA (mbarrier phases)
  const bool prof = p.prof && blockIdx.x == 0 && tid == 0;
  long long tk = 0;
  auto tick = [&](int slot) {
    if (prof) {
      const long long now = clock64();
      g_prof32[slot] += static_cast<unsigned long long>(now - tk);
      tk = now;
    }
  };
  for (; tile < p.n_tiles; tile += gridDim.x) {
    const int next_tile = tile + gridDim.x;
    if (prof) { tk = clock64(); g_prof32[15] += 1; }
    // ---- channel LayerNorm of the tile, in place: 4 threads per token, C/4 channels each
    cp_wait<0>();
    __syncthreads();
    {
      constexpr int CPT = C / 4, CHT = CPT / 8;
      const int n = tid >> 2, part = tid & 3;
      const bool valid = row_pixel(tile, n) >= 0;
      float v[CPT];
#pragma unroll
      for (int h2 = 0; h2 < CHT; ++h2) {
        const int j = part * CHT + h2;
        const uint4 t = *reinterpret_cast<const uint4*>(s_a + (j >> 3) * 16384 + sw128(n, j & 7));
        const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
        v[h2 * 8] = a.x; v[h2 * 8 + 1] = a.y; v[h2 * 8 + 2] = b.x; v[h2 * 8 + 3] = b.y;
        v[h2 * 8 + 4] = c.x; v[h2 * 8 + 5] = c.y; v[h2 * 8 + 6] = d.x; v[h2 * 8 + 7] = d.y;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) sum += v[j];
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) { const float dd = v[j] - mean; sq += dd * dd; }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      const float rstd = valid ? rsqrtf(sq * (1.0f / C) + p.eps) : 0.f;
      if (part == 0) s_stat[n] = make_float2(mean, rstd);
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = (v[j] - mean) * rstd * s_gamma[part * CPT + j];
      if (TEMPORAL) {                                       // u = LayerNorm(z) * w + b
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) s2 += v[j];
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        const float mean2 = s2 * (1.0f / C);
        float q2 = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) { const float dd = v[j] - mean2; q2 += dd * dd; }
        q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
        q2 += __shfl_xor_sync(0xffffffffu, q2, 2);
        const float rstd2 = rsqrtf(q2 * (1.0f / C) + p.eps);
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          v[j] = valid ? (v[j] - mean2) * rstd2 * s_lnw[part * CPT + j] + s_lnb[part * CPT + j] : 0.f;
      }
#pragma unroll
      for (int h2 = 0; h2 < CHT; ++h2) {
        const int j = part * CHT + h2;
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(v[h2 * 8 + 2 * e], v[h2 * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(s_a + (j >> 3) * 16384 + sw128(n, j & 7)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }

    tick(0);                                                // token wait + LayerNorm
    for (int g = 0; g < 4; ++g, ++gc) {
      const int buf = gc & 1;
      uint8_t* st = sm + L::stage + buf * L::STAGE;
      if (gc > 0) {                                         // out-proj of the previous group retired: its O tile (s_q) and
        mbar_wait(bar_d, (gc - 1) & 1);                     // its weight stage are free
        tc_fence_after();
      }
      const bool more = g < 3 || next_tile < p.n_tiles;
      if (more) {
        load_stage((g + 1) & 3, buf ^ 1);
        cp_wait<1>();                                       // everything but the stage just issued has landed
      } else {
        cp_wait<0>();
      }
      fence_proxy_async();                                  // A tile / weight stage -> visible to the tensor-core proxy
      tc_fence_before();
      __syncthreads();
      tick(1);                                              // previous out-proj + weight stage wait

      // ---- QKV_g = LN(x) . Wqkv_g^T
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t da = umma_desc_sw128(smem_u32(s_a) + kb * 16384);
            const uint64_t db = umma_desc_sw128(smem_u32(st) + kb * (192 * 128));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_u, da + 2 * ks, db + 2 * ks, idesc_qkv, (kb | ks) ? 1u : 0u);
          }
          umma_commit(bar_qkv);
        }
        __syncwarp();
      }
      mbar_wait(bar_qkv, gc & 1);
      tc_fence_after();
      tick(2);                                              // QKV product
      if (g == 3 && next_tile < p.n_tiles) prefetch_tokens(next_tile);     // the A tile is free: next tile's raw tokens

      // ---- drain: thread = token, 48 of the 192 columns; q-scale + rotary; Q, K row-major, V transposed
      {
        const int pos = row % TP;
        uint32_t rr[3][16];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) tmem_ld16(tlane + cg * 48 + ch * 16, rr[ch]);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const int col0 = cg * 48 + ch * 16;
          const int region = col0 >> 6, hh = (col0 >> 5) & 1, d0 = col0 & 31;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(rr[ch][j]);
          if (region < 2) {
            const float sc = region == 0 ? qscale : 1.0f;
#pragma unroll
            for (int pr = 0; pr < 8; ++pr) {
              const float cs = s_cos[((d0 >> 1) + pr) * 32 + pos], sn = s_sin[((d0 >> 1) + pr) * 32 + pos];
              const float x0 = f[2 * pr] * sc, x1 = f[2 * pr + 1] * sc;
              f[2 * pr] = x0 * cs - x1 * sn;
              f[2 * pr + 1] = x1 * cs + x0 * sn;
            }
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
            uint8_t* dst = region == 0 ? s_q : s_k;
            const int c0 = (hh * 32 + d0) >> 3;
            *reinterpret_cast<uint4*>(dst + sw128(row, c0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + sw128(row, c0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            uint8_t* dst = s_vt + (row >> 6) * 8192 + (row & 7) * 2;
            const int kc = (row & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              *reinterpret_cast<__nv_bfloat16*>(dst + sw128(hh * 32 + d0 + j, kc)) = __float2bfloat16(f[j]);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tick(3);                                              // QKV drain

      // ---- S_h = Q_h . K_h^T for the group's two heads
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dqd = umma_desc_sw128(smem_u32(s_q)), dkd = umma_desc_sw128(smem_u32(s_k));
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tmem_u + (hh ? S1 : S0), dqd + 4 * hh + 2 * ks, dkd + 4 * hh + 2 * ks, idesc_s, ks ? 1u : 0u);
          umma_commit(bar_s);
        }
        __syncwarp();
      }
      mbar_wait(bar_s, gc & 1);
      tc_fence_after();
      tick(4);                                              // score product

      // ---- softmax: warp = (lane quarter, head of the group, column half); thread = query row
      {
        const int hh = cg & 1, half = cg >> 1;
        const int head = g * 2 + hh;
        const uint32_t sb = tlane + (hh ? S1 : S0) + dq * 32;
        float s[NC];
        if (TP == 32) {
          uint32_t rr[16];
          tmem_ld16(sb + half * 16, rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; ++j) s[j] = __uint_as_float(rr[j % 16]);
        } else {
          uint32_t ra[16], rb[16];
          tmem_ld16(sb + half * 8, ra);
          tmem_ld16(sb + 16 + half * 8, rb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; ++j) s[j] = __uint_as_float(lane < 16 ? ra[j % 16] : rb[j % 16]);
        }
        const int i = lane & (TP - 1), j0 = half * NC;
        {
          constexpr int CPR = TP / 8, RPL = 8 / CPR;
          const uint8_t* brow = s_bias + (head * TP + i) * TP * 2;
#pragma unroll
          for (int c = 0; c < NC / 8; ++c) {
            const int cs = ((j0 >> 3) + c) ^ ((i / RPL) & (CPR - 1));
            const uint4 t = *reinterpret_cast<const uint4*>(brow + cs * 16);
            const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c2 = unpack_bf16(t.z), d = unpack_bf16(t.w);
            s[c * 8] += a.x; s[c * 8 + 1] += a.y; s[c * 8 + 2] += b.x; s[c * 8 + 3] += b.y;
            s[c * 8 + 4] += c2.x; s[c * 8 + 5] += c2.y; s[c * 8 + 6] += d.x; s[c * 8 + 7] += d.y;
          }
        }
        if (TEMPORAL) {
#pragma unroll
          for (int j = 0; j < NC; ++j)
            if (j0 + j >= p.T) s[j] = -1.0e30f;             // padded frames are not keys
        } else {
          const Unit w = decode(tile * NU + dq);            // TP == 32: the warp's 32 lanes are one window
          if (unit_masked(w)) {                             // -100 where the Swin region ids of query and key differ
            const int code = region_code(w, lane);
            uint32_t same = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t bal = __ballot_sync(0xffffffffu, code == c);
              if (code == c) same = bal;
            }
#pragma unroll
            for (int j = 0; j < NC; ++j)
              if (!((same >> (j0 + j)) & 1u)) s[j] += kMask;
          }
        }
        float m = s[0];
#pragma unroll
        for (int j = 1; j < NC; ++j) m = fmaxf(m, s[j]);
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j) { s[j] = ex2(s[j] - m); l += s[j]; }
        s_ml[(hh * 128 + row) * 2 + half] = make_float2(m, l);
        pair_barrier(1 + (warp & 7));                       // the two column halves of a (quarter, head) exchange (max, sum)
        const float2 o = s_ml[(hh * 128 + row) * 2 + (half ^ 1)];
        const float mm = fmaxf(m, o.x);
        const float f1 = ex2(m - mm);
        const float f = __fdividef(f1, l * f1 + o.y * ex2(o.x - mm));
        const int kr = (TP == 32 ? dq * 32 : (dq * 2 + (lane >> 4)) * 16) + j0;   // tile row of this thread's first key
        uint8_t* pd = s_p + hh * 32768;
#pragma unroll
        for (int c = 0; c < NC / 8; ++c) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(s[c * 8 + 2 * e] * f, s[c * 8 + 2 * e + 1] * f);
          const int chunk = (kr >> 3) + c;
          *reinterpret_cast<uint4*>(pd + (chunk >> 3) * 16384 + sw128(row, chunk & 7)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tick(5);                                              // softmax

      // ---- O_h = P_h . V_h
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t dp = umma_desc_sw128(smem_u32(s_p) + hh * 32768 + kb * 16384);
              const uint64_t dv = umma_desc_sw128(smem_u32(s_vt) + kb * 8192 + hh * 4096);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tmem_u + OG + hh * 32, dp + 2 * ks, dv + 2 * ks, idesc_pv, (kb | ks) ? 1u : 0u);
            }
          umma_commit(bar_o);
        }
        __syncwarp();
      }
      mbar_wait(bar_o, gc & 1);
      tc_fence_after();
      tick(6);                                              // PV product

      // ---- O_g -> bf16 A tile of the output projection (over the group's Q slots)
      {
        uint32_t rr[16];
        tmem_ld16(tlane + OG + cg * 16, rr);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(rr[2 * j]), __uint_as_float(rr[2 * j + 1]));
        *reinterpret_cast<uint4*>(s_q + sw128(row, cg * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(s_q + sw128(row, cg * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tick(7);                                              // O drain

      // ---- D (+)= O_g . Wproj[:, 64g : 64g+64]^T
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_sw128(smem_u32(s_q)), db = umma_desc_sw128(smem_u32(st) + L::WQ);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_u + DO, da + 2 * ks, db + 2 * ks, idesc_d, (g | ks) ? 1u : 0u);
          umma_commit(bar_d);
        }
        __syncwarp();
      }
    }

    // ---- epilogue: D + bias + residual (+ chanLN(x) in temporal mode) -> y
    mbar_wait(bar_d, (gc - 1) & 1);
    tc_fence_after();
    {
      constexpr int CW = C / 4;                             // columns per warp
      const int d = row_pixel(tile, row);
      const float2 stt = s_stat[row];
#pragma unroll
      for (int c16 = 0; c16 < CW / 16; ++c16) {
        const int c0 = cg * CW + c16 * 16;
        uint32_t rr[16];
        tmem_ld16(tlane + DO + c0, rr);
        tmem_ld_wait();
        if (d >= 0) {
          const uint4* xp = reinterpret_cast<const uint4*>(p.x + static_cast<long long>(d) * C + c0);
          const uint4 r0v = xp[0], r1v = xp[1];
          const uint32_t rw[8] = {r0v.x, r0v.y, r0v.z, r0v.w, r1v.x, r1v.y, r1v.z, r1v.w};
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 rv = unpack_bf16(rw[j]);
            float o0 = __uint_as_float(rr[2 * j]) + s_pbias[c0 + 2 * j] + rv.x;
            float o1 = __uint_as_float(rr[2 * j + 1]) + s_pbias[c0 + 2 * j + 1] + rv.y;
            if (TEMPORAL) {
              o0 += (rv.x - stt.x) * stt.y * s_gamma[c0 + 2 * j];
              o1 += (rv.y - stt.x) * stt.y * s_gamma[c0 + 2 * j + 1];
            }
            pk[j] = pack_bf16(o0, o1);
          }
          __nv_bfloat16* yp = p.y + static_cast<long long>(d) * C + c0;
          *reinterpret_cast<uint4*>(yp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(yp + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
    tc_fence_before();                                      // TMEM reads done before the next tile's products overwrite it
    tick(8);                                                // last out-proj wait + epilogue
  }
  cp_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_u, 512);
  }
}

template <int C, int TP, bool TEMPORAL>
int launch32(const P32& p, cudaStream_t st) {
  constexpr int smem = Lay<C>::total + 1024;
  static SmemConfigured configured;
  if (!configured.covers(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc32_kernel<C, TP, TEMPORAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem);
  }
  const int sms = device_sm_count();
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  static const bool prof_env = getenv("EXTDM_ATTN32_PROF") != nullptr;
  bool prof = prof_env;
  if (prof) {                                               // the read-back synchronises: not during graph capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) prof = false;
  }
  P32 q = p;
  q.prof = prof ? 1 : 0;
  if (prof) {
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_prof32, z, sizeof(z));
  }
  attn_tc32_kernel<C, TP, TEMPORAL><<<grid, NTH, smem, st>>>(q);
  if (prof) {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_prof32, sizeof(h));
    const double n = h[15] ? static_cast<double>(h[15]) : 1.0;
    fprintf(stderr, "[attn32 prof] temporal=%d TP=%d tiles/CTA=%.0f cycles/tile: ln %.0f | per tile (4 groups): stage_wait %.0f qkv_mma %.0f "
            "drain %.0f s_mma %.0f softmax %.0f pv_mma %.0f o_drain %.0f | epilogue %.0f\n", (int)TEMPORAL, TP, n, h[0] / n, h[1] / n,
            h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n);
  }
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

}  // namespace
}  // namespace extdm

using namespace extdm;

// (2,4,4)-window attention layer, 8 heads x 32, C = 64, on tcgen05.  Called by extdm_stw_fused.  Returns -1 when the
// geometry is not supported (the caller falls back to the mma.sync kernel).
int extdm_stw_tc32_launch(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                          const float* proj_bias, const float* bias_table, const float* rope_cos, const float* rope_sin,
                          int B, int T, int H, int W, int C, int sd, int sh, int sw, float eps, void* stream) {
  const int nww = W / 4, nwh = H / 4;
  if (C != 64 || W % 4 || H % 4 || (nww & (nww - 1)) || (nwh & (nwh - 1))) return -1;
  P32 p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.gamma = gamma;
  p.ln_w = p.ln_b = nullptr;
  p.wqkv = reinterpret_cast<const __nv_bfloat16*>(wqkv);
  p.wproj = reinterpret_cast<const __nv_bfloat16*>(wproj);
  p.proj_bias = proj_bias;
  p.bias_table = bias_table;
  p.rcos = rope_cos;
  p.rsin = rope_sin;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.sd = sd; p.sh = sh; p.sw = sw;
  p.Dp = (T + 1) / 2 * 2;
  p.n_units = B * (p.Dp / 2) * nwh * nww;
  p.n_tiles = (p.n_units + 3) / 4;
  p.lw = 0; p.lh = 0;
  while ((1 << p.lw) < nww) ++p.lw;
  while ((1 << p.lh) < nwh) ++p.lh;
  p.eps = eps;
  return launch32<64, 32, false>(p, static_cast<cudaStream_t>(stream));
}

// Temporal attention layer, 8 heads x 32, C = 64, T <= 32, on tcgen05.  Called by extdm_temporal_fused.
int extdm_temporal_tc32_launch(const void* x, void* y, const float* gamma, const float* ln_w, const float* ln_b,
                               const void* wqkv, const void* wout, const float* rel_bias, const float* rope_cos,
                               const float* rope_sin, int B, int T, int HW, int C, float eps, void* stream) {
  if (C != 64 || T < 1 || T > 32) return -1;
  P32 p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.gamma = gamma;
  p.ln_w = ln_w;
  p.ln_b = ln_b;
  p.wqkv = reinterpret_cast<const __nv_bfloat16*>(wqkv);
  p.wproj = reinterpret_cast<const __nv_bfloat16*>(wout);
  p.proj_bias = nullptr;
  p.bias_table = rel_bias;
  p.rcos = rope_cos;
  p.rsin = rope_sin;
  p.B = B; p.T = T; p.H = HW; p.W = 1;
  p.sd = p.sh = p.sw = 0;
  p.Dp = T;
  p.n_units = B * HW;
  p.eps = eps;
  if (T <= 16) {
    p.n_tiles = (p.n_units + 7) / 8;
    return launch32<64, 16, true>(p, static_cast<cudaStream_t>(stream));
  }
  p.n_tiles = (p.n_units + 3) / 4;
  return launch32<64, 32, true>(p, static_cast<cudaStream_t>(stream));
}
