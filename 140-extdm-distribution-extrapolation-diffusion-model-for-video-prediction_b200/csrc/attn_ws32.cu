// Attention layers of the dim_head-32 UNet variants (u12 = BAIR, base = SMMNIST, ada_u22), C = 64, 8 heads x 32 --
// every product on the 5th-generation tensor cores (tcgen05.mma, accumulators and the softmax output in tensor memory),
// warp-specialised so that the tensor pipe, the TMEM read port and the CUDA cores work on different head groups at once.
//
//   window mode    y = x + proj( WindowAttention3D( chanLN(x) ) ) + b        Residual(PreNorm(STWAttentionLayer)), (2,4,4) windows
//                  reference: model/BaseDM_adaptor/DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:139-159, :409-560
//   temporal mode  y = x + chanLN(x) + to_out( Attention( LayerNorm(chanLN(x)) ) )   Residual(PreNorm(EinopsToAndFrom(AttentionLayer)))
//                  reference: ...cross_multi.py:253-328 (double residual: App. B.4 of SURVEY.md)
//
// One M = 128 tile = 4 windows of 32 tokens, or 4 / 8 pixel sequences of T <= 32 / <= 16 frames; heads in 4 groups of 2:
//   QKV_g [128 x 192] = LN(x)[128 x 64] . Wqkv_g^T       SS-form MMA (A: normalised tokens, B: weight slice, both smem)
//   S_h   [128 x 128] = Q_h[128 x 32] . K_h^T             SS; only a row's own 32 (16) block-diagonal columns are read back
//   O_h   [128 x 32]  = P_h[128 x 128] . V_h              TS: P = softmax(S + bias [+ mask]) stays in TMEM as bf16 pairs,
//                                                         written over the head's S columns (zero off the diagonal block)
//   D     [128 x 64] += O_g[128 x 64] . Wproj[:, g]^T     SS, accumulated over the 4 groups
//
// Roles (18 warps):
//   warps 0-7   "drain"    TMEM -> registers -> shared operands: QKV_g (q-scale + rotary; Q, K row-major, V transposed),
//                          O_g (bf16 A tile of the output projection)
//   warps 8-15  "softmax"  scores -> P (thread = query row x head), plus everything with slack: weight / token cp.async,
//                          the channel LayerNorm of the NEXT tile, the epilogue (D + bias + residual -> y)
//   warp 16     "issue"    one elected lane issues every tcgen05.mma and commits the mbarriers the other roles wait on
//   warp 17     "load"     one elected lane streams the weight slices by TMA (Wproj once, Wqkv_g through a two-slot ring)
// so the QKV drain of group g+1 (TMEM-read bound, 64 B/clk) runs under the softmax of group g.
// TMEM map (512 columns): QKV [0,192) | head A: S [192,320), P over [192,256), O over [256,288) | head B: S [320,448),
// P over [320,384), O over [384,416) | D [448,512).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {
namespace {

constexpr float kL2e = 1.4426950408889634f;
constexpr int C = 64, HEADS = 8, HID = 256;
constexpr int NTH = 576, ND = 256, NS = 256;               // threads: all (18 warps) / drain group / softmax group
constexpr uint32_t T_QKV = 0, T_SA = 192, T_SB = 320, T_D = 448;

__device__ __forceinline__ int sw128(int r, int j) { return (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void group_barrier(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct Lay {
  static constexpr int wq = 0;                               // 2 x [192][64] SW128: Wqkv_g ring
  static constexpr int wp = wq + 2 * 24576;                  // 4 x [64][64] SW128: Wproj[:, 64g : 64g+64], resident
  static constexpr int a = wp + 4 * 8192;                    // 2 x [128][64] SW128: raw tokens -> LN output, ping-pong by tile
  static constexpr int q = a + 2 * 16384;                    // [128][64] SW128: Q of the group
  static constexpr int k = q + 16384;                        // [128][64] SW128
  static constexpr int vt = k + 16384;                       // 2 x V^T: 2 k-blocks x [64][64] SW128, ping-pong by group
  static constexpr int o = vt + 2 * 16384;                   // [128][64] SW128: O_g (A of the output projection)
  static constexpr int bias = o + 16384;                     // [8][TP][TP] bf16 * log2(e), chunk-swizzled
  static constexpr int rope = bias + 16384;                  // (cos, sin) float2 [16 pairs][32 positions]
  static constexpr int vec = rope + 4096;                    // gamma, ln_w, ln_b, proj bias: 4 x 64 fp32
  static constexpr int stat = vec + 1024;                    // 2 x [128] (mean, rstd), ping-pong by tile
  static constexpr int bars = stat + 2048;                   // mbarriers + TMEM slot
  static constexpr int total = bars + 128;
};
enum Bar { B_QKV_DONE = 0, B_QKV_FREE, B_QK_READY, B_S_DONE, B_P_READY, B_O_DONE, B_O_READY, B_D_DONE, B_D_FREE,
           B_A_READY, B_WQ_READY /* two: one per ring slot */, B_WQ_READY1, B_WP_READY, B_COUNT };

// EXTDM_ATTN32_PROF=1: cycle counts of CTA 0, one thread per role, printed by the launcher
//   [0..3] drain: wait QKV | drain QKV | wait O | drain O      [4..7] softmax: wait S | softmax | LN | epilogue
//   [8..11] issue: wait QK_READY | wait for the early QKV's inputs | wait P_READY | wait O_READY      [15] groups
__device__ unsigned long long g_prof_ws[24];

struct P32 {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* gamma;
  const float* ln_w;           // temporal mode: nn.LayerNorm applied after the channel LayerNorm
  const float* ln_b;
  const __nv_bfloat16* wqkv;   // [768][64]
  const __nv_bfloat16* wproj;  // [64][256]
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
template <int TP, bool TEMPORAL, bool PROF>
__global__ void __launch_bounds__(NTH, 1) attn_ws32_kernel(const __grid_constant__ P32 p,
                                                           const __grid_constant__ CUtensorMap map_wq,
                                                           const __grid_constant__ CUtensorMap map_wp) {
  using L = Lay;
  constexpr int NU = 128 / TP;                             // attention units per tile
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = sm + L::q;
  uint8_t* s_k = sm + L::k;
  uint8_t* s_o = sm + L::o;
  uint8_t* s_bias = sm + L::bias;
  float2* s_rope = reinterpret_cast<float2*>(sm + L::rope);   // (cos, sin) [16 pairs][32 positions]
  float* s_gamma = reinterpret_cast<float*>(sm + L::vec);
  float* s_lnw = s_gamma + C;
  float* s_lnb = s_lnw + C;
  float* s_pbias = s_lnb + C;
  float2* s_stat = reinterpret_cast<float2*>(sm + L::stat);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

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

  // ---- one-time staging (all threads): resident Wproj, rope / bias tables, vectors; barriers; TMEM
  for (int i = tid; i < 32 * 16; i += NTH) {                // global [pos][pair] -> shared [pair][pos]
    s_rope[(i & 15) * 32 + (i >> 4)] = make_float2(p.rcos[i], p.rsin[i]);
  }
  for (int i = tid; i < C; i += NTH) {
    s_gamma[i] = p.gamma[i];
    s_lnw[i] = TEMPORAL ? p.ln_w[i] : 1.f;
    s_lnb[i] = TEMPORAL ? p.ln_b[i] : 0.f;
    s_pbias[i] = p.proj_bias ? p.proj_bias[i] : 0.f;
  }
  {
    // additive score bias expanded to [head][query][key], times log2(e); 16-byte chunks XOR-swizzled against the
    // row-strided reads of the softmax threads
    constexpr int CPR = TP / 8, RPL = 8 / CPR;
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
  if (tid == 0) {
    mbar_init(bars + B_QKV_DONE, 1);
    mbar_init(bars + B_QKV_FREE, ND);
    mbar_init(bars + B_QK_READY, ND);
    mbar_init(bars + B_S_DONE, 1);
    mbar_init(bars + B_P_READY, NS);
    mbar_init(bars + B_O_DONE, 1);
    mbar_init(bars + B_O_READY, NS);
    mbar_init(bars + B_D_DONE, 1);
    mbar_init(bars + B_D_FREE, ND);
    mbar_init(bars + B_A_READY, ND);
    mbar_init(bars + B_WQ_READY, 1);                        // expect_tx arrival of the thread that issues the TMA loads
    mbar_init(bars + B_WQ_READY1, 1);
    mbar_init(bars + B_WP_READY, 1);
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

  const bool prof = PROF && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8 || warp == 16);
  long long tk = (PROF && prof) ? clock64() : 0;
  auto tick = [&](int slot) {
    if (PROF && prof) {
      const long long now = clock64();
      g_prof_ws[slot] += static_cast<unsigned long long>(now - tk);
      tk = now;
    }
  };
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = (p.n_tiles - first + stride - 1) / stride;        // >= 1: the grid never exceeds n_tiles
  const int n_groups = my_tiles * 4;

  auto load_tokens = [&](int it, int t256) {                // raw bf16 tokens of this CTA's it-th tile -> A buffer it & 1:
    uint8_t* dst = sm + L::a + (it & 1) * 16384;            // two threads per token row, four 16-byte chunks each
    const int tile = first + it * stride;
    const int n = t256 >> 1, part = t256 & 1;
    const int s = row_pixel(tile, n);
    const __nv_bfloat16* src = p.x + (s >= 0 ? static_cast<long long>(s) * C : 0) + part * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) cp_async16(dst + sw128(n, part * 4 + j), src + j * 8, s >= 0 ? 16 : 0);
  };
  auto layernorm = [&](int it, int t256) {                          // in place on A buffer it & 1: 2 threads per token, 32 channels each
      uint8_t* buf = sm + L::a + (it & 1) * 16384;
      const int tile = first + it * stride;
      const int n = t256 >> 1, part = t256 & 1;
      const bool valid = row_pixel(tile, n) >= 0;
      float v[32];
#pragma unroll
      for (int h2 = 0; h2 < 4; ++h2) {
        const uint4 t = *reinterpret_cast<const uint4*>(buf + sw128(n, part * 4 + h2));
        const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
        v[h2 * 8] = a.x; v[h2 * 8 + 1] = a.y; v[h2 * 8 + 2] = b.x; v[h2 * 8 + 3] = b.y;
        v[h2 * 8 + 4] = c.x; v[h2 * 8 + 5] = c.y; v[h2 * 8 + 6] = d.x; v[h2 * 8 + 7] = d.y;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float dd = v[j] - mean; sq += dd * dd; }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      const float rstd = valid ? rsqrtf(sq * (1.0f / C) + p.eps) : 0.f;
      if (part == 0) s_stat[(it & 1) * 128 + n] = make_float2(mean, rstd);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * s_gamma[part * 32 + j];
      if (TEMPORAL) {                                       // u = LayerNorm(z) * w + b
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) s2 += v[j];
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        const float mean2 = s2 * (1.0f / C);
        float q2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float dd = v[j] - mean2; q2 += dd * dd; }
        q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
        const float rstd2 = rsqrtf(q2 * (1.0f / C) + p.eps);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = valid ? (v[j] - mean2) * rstd2 * s_lnw[part * 32 + j] + s_lnb[part * 32 + j] : 0.f;
      }
#pragma unroll
      for (int h2 = 0; h2 < 4; ++h2) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(v[h2 * 8 + 2 * e], v[h2 * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(buf + sw128(n, part * 4 + h2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    };
  if (warp == 16) {
    // =========================================================================================== issue
    if (elect_one()) {
      constexpr uint32_t idesc_qkv = umma_idesc_bf16(128, 192), idesc_s = umma_idesc_bf16(128, 128),
                         idesc_pv = umma_idesc_bf16(128, 32), idesc_d = umma_idesc_bf16(128, 64);
      auto issue_qkv = [&](int G, bool wait_free) {         // waits: weight slice landed, QKV columns free, A tile ready
        const int it = G >> 2;
        mbar_wait(bars + B_WQ_READY + (G & 1), (G >> 1) & 1);
        if (wait_free && G > 0) mbar_wait(bars + B_QKV_FREE, (G - 1) & 1);
        if ((G & 3) == 0) mbar_wait(bars + B_A_READY, it & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(sm + L::a + (it & 1) * 16384));
        const uint64_t db = umma_desc_sw128(smem_u32(sm + L::wq + (G & 1) * 24576));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_u + T_QKV, da + 2 * ks, db + 2 * ks, idesc_qkv, ks ? 1u : 0u);
        umma_commit(bars + B_QKV_DONE);
      };
      auto issue_s = [&](int G) {                           // S_h = Q_h . K_h^T for the group's two heads
        mbar_wait(bars + B_QK_READY, G & 1);
        tc_fence_after();
        const uint64_t dqd = umma_desc_sw128(smem_u32(s_q)), dkd = umma_desc_sw128(smem_u32(s_k));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tmem_u + (hh ? T_SB : T_SA), dqd + 4 * hh + 2 * ks, dkd + 4 * hh + 2 * ks, idesc_s, ks ? 1u : 0u);
        umma_commit(bars + B_S_DONE);
      };
      // Issue order (software pipeline over the groups):  QKV(0) S(0) QKV(1) | PV(G) S(G+1) QKV(G+2) OUT(G) | ...
      // so that a group's scores are being computed while the previous group's O is drained and projected, and a group's
      // QKV accumulators are drained (other warps) under the previous group's softmax.
      issue_qkv(0, true);
      issue_s(0);
      if (n_groups > 1) issue_qkv(1, true);
      tick(9);
      for (int G = 0; G < n_groups; ++G) {
        const int g = G & 3, it = G >> 2;
        // ---- O_h = P_h . V_h, P from tensor memory; O lands in the QKV columns, whose current tenant (group G+1's
        // accumulators) must be in registers by now
        mbar_wait(bars + B_P_READY, G & 1);
        if (G + 1 < n_groups) mbar_wait(bars + B_QKV_FREE, (G + 1) & 1);
        tc_fence_after();
        tick(10);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t sbase = tmem_u + (hh ? T_SB : T_SA);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t dv = umma_desc_sw128(smem_u32(sm + L::vt + (G & 1) * 16384) + kb * 8192 + hh * 4096);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_ts(tmem_u + T_QKV + hh * 32, sbase + kb * 32 + ks * 8, dv + 2 * ks, idesc_pv, (kb | ks) ? 1u : 0u);
          }
        }
        umma_commit(bars + B_O_DONE);
        // ---- next group's scores over the S / P columns, once the PV product has consumed P
        if (G + 1 < n_groups) {
          mbar_wait(bars + B_O_DONE, G & 1);
          issue_s(G + 1);
        }
        tick(8);
        // ---- O of this group is in shared memory: the QKV columns are free again, the output projection can run
        mbar_wait(bars + B_O_READY, G & 1);
        tick(11);
        if (G + 2 < n_groups) issue_qkv(G + 2, false);
        tick(9);
        if (G == 0) mbar_wait(bars + B_WP_READY, 0);
        if (g == 0 && it > 0) mbar_wait(bars + B_D_FREE, (it - 1) & 1);      // previous tile's epilogue has read D
        tc_fence_after();
        {
          const uint64_t da = umma_desc_sw128(smem_u32(s_o)), db = umma_desc_sw128(smem_u32(sm + L::wp + g * 8192));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_u + T_D, da + 2 * ks, db + 2 * ks, idesc_d, (g | ks) ? 1u : 0u);
          if (g == 3) umma_commit(bars + B_D_DONE);
        }
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // =========================================================================================== load (TMA)
    if (elect_one()) {
      mbar_expect_tx(bars + B_WP_READY, 32768);             // resident Wproj: 4 hidden slices of [64][64]
#pragma unroll
      for (int g = 0; g < 4; ++g) tma_load_2d(&map_wp, sm + L::wp + g * 8192, bars + B_WP_READY, g * 64, 0);
      for (int G = 0; G < n_groups; ++G) {                  // Wqkv rows {q,k,v} x [64g, 64g+64) -> ring slot G & 1
        if (G >= 2) mbar_wait(bars + B_QKV_DONE, (G - 2) & 1);      // the slot's previous user, QKV(G-2), has retired
        uint8_t* st = sm + L::wq + (G & 1) * 24576;
        uint64_t* bar = bars + B_WQ_READY + (G & 1);
        mbar_expect_tx(bar, 24576);
#pragma unroll
        for (int r = 0; r < 3; ++r) tma_load_2d(&map_wq, st + r * 8192, bar, 0, r * HID + (G & 3) * 64);
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // =========================================================================================== drain
    const int dq = warp & 3, dw = warp >> 2;                // TMEM lane quarter, column half
    const int row = dq * 32 + lane, pos = row % TP;
    const uint32_t tlane = tmem_u + (static_cast<uint32_t>(dq * 32) << 16);
    const float qscale = 0.17677669529663687f * kL2e;       // dh^-1/2 * log2(e)
    auto epilogue = [&](int it) {                           // D + bias + residual (+ chanLN(x) in temporal mode) -> y
      const int tile = first + it * stride;
      const int d = row_pixel(tile, row);
      const float2 stt = s_stat[(it & 1) * 128 + row];
      uint4 xr[4];
      if (d >= 0) {                                         // residual row slice: in flight while the out-proj retires
        const uint4* xp = reinterpret_cast<const uint4*>(p.x + static_cast<long long>(d) * C + dw * 32);
#pragma unroll
        for (int e = 0; e < 4; ++e) xr[e] = xp[e];
      }
      mbar_wait(bars + B_D_DONE, it & 1);
      tc_fence_after();
      uint32_t ra[16], rb[16];
      tmem_ld16(tlane + T_D + dw * 32, ra);
      tmem_ld16(tlane + T_D + dw * 32 + 16, rb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bars + B_D_FREE);
      if (d >= 0) {
        __nv_bfloat16* yp = p.y + static_cast<long long>(d) * C + dw * 32;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t rw[4] = {xr[e].x, xr[e].y, xr[e].z, xr[e].w};
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c0 = dw * 32 + e * 8 + 2 * j;
            const float2 rv = unpack_bf16(rw[j]);
            const int ri = (e & 1) * 8 + 2 * j;
            float o0 = __uint_as_float(e < 2 ? ra[ri] : rb[ri]) + s_pbias[c0] + rv.x;
            float o1 = __uint_as_float(e < 2 ? ra[ri + 1] : rb[ri + 1]) + s_pbias[c0 + 1] + rv.y;
            if (TEMPORAL) {
              o0 += (rv.x - stt.x) * stt.y * s_gamma[c0];
              o1 += (rv.y - stt.x) * stt.y * s_gamma[c0 + 1];
            }
            pk[j] = pack_bf16(o0, o1);
          }
          *reinterpret_cast<uint4*>(yp + e * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    };

    // ---- prologue: tokens of the first tile, its LayerNorm
    load_tokens(0, tid);
    cp_commit();
    cp_wait_all();
    group_barrier(1, ND);
    layernorm(0, tid);
    fence_proxy_async();
    mbar_arrive(bars + B_A_READY);
    for (int G = 0; G < n_groups; ++G) {
      // ---- QKV_g: thread = token, 96 of the 192 columns in two passes of 48
      mbar_wait(bars + B_QKV_DONE, G & 1);
      tc_fence_after();
      tick(0);
      uint8_t* vt = sm + L::vt + (G & 1) * 16384;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        uint32_t rr[3][16];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) tmem_ld16(tlane + T_QKV + dw * 96 + pass * 48 + ch * 16, rr[ch]);
        tmem_ld_wait();
        if (pass == 1) {                                    // every QKV column of this warp is in registers
          tc_fence_before();
          mbar_arrive(bars + B_QKV_FREE);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const int col0 = dw * 96 + pass * 48 + ch * 16;   // compile-time after unrolling except for dw (warp-uniform)
          const int region = col0 >> 6, hh = (col0 >> 5) & 1, d0 = col0 & 31;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(rr[ch][j]);
          if (region < 2) {
            const float sc = region == 0 ? qscale : 1.0f;
#pragma unroll
            for (int pr = 0; pr < 8; ++pr) {
              const float2 rp = s_rope[((d0 >> 1) + pr) * 32 + pos];
              const float cs = rp.x, sn = rp.y;
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
            uint8_t* dst = vt + (row >> 6) * 8192 + (row & 7) * 2;
            const int kc = (row & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              *reinterpret_cast<__nv_bfloat16*>(dst + sw128(hh * 32 + d0 + j, kc)) = __float2bfloat16(f[j]);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bars + B_QK_READY);
      tick(1);
      // ---- per-tile work of this group, spread over the tile's groups:
      //   g == 0  request the next tile's tokens (their buffer's last reader, QKV(G-1), retired before QKV(G))
      //   g == 1  the previous tile's epilogue (its last out-proj retired a drain ago)
      //   g == 2  the next tile's channel LayerNorm, in place on the landed tokens
      const int g = G & 3, it = G >> 2;
      if (g == 0 && it + 1 < my_tiles) { load_tokens(it + 1, tid); cp_commit(); tick(7); }
      if (g == 0 && it >= 1) {                              // the previous tile's residual rows -> L1, for the epilogue a group later
        const int d = row_pixel(first + (it - 1) * stride, row);
        if (d >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.x + static_cast<long long>(d) * C + dw * 32));
      }
      if (g == 1 && it >= 1) { epilogue(it - 1); tick(12); }
      if (g == 2 && it + 1 < my_tiles) {                    // the softmax chain is the critical one: LayerNorm runs here
        cp_wait_all();
        group_barrier(1, ND);                               // every thread's copies landed; epilogue(it-1) has read its stats
        layernorm(it + 1, tid);
        fence_proxy_async();
        mbar_arrive(bars + B_A_READY);
        tick(6);
      }
    }
    epilogue(my_tiles - 1);
  } else {
    // =========================================================================================== softmax (+ loads, LN, epilogue)
    const int sw_ = warp - 8;
    const int dq = sw_ & 3, hh = sw_ >> 2;                  // TMEM lane quarter, head of the group
    const int row = dq * 32 + lane;
    const int stid = tid - 256;                             // 0..255 within the group
    const uint32_t tlane = tmem_u + (static_cast<uint32_t>(dq * 32) << 16);
    constexpr float kMask = -100.0f * kL2e;
    uint32_t same = 0xffffffffu;                            // keys of this row's window that share its Swin region id
    auto drain_o = [&](int G) {                             // O of head hh -> bf16 -> chunks [4 hh, 4 hh + 4) of the O tile
      mbar_wait(bars + B_O_DONE, G & 1);
      tc_fence_after();
      tick(2);
      uint32_t ra[16], rb[16];
      const uint32_t ob = tlane + T_QKV + hh * 32;
      tmem_ld16(ob, ra);
      tmem_ld16(ob + 16, rb);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
      *reinterpret_cast<uint4*>(s_o + sw128(row, hh * 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(s_o + sw128(row, hh * 4 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
      *reinterpret_cast<uint4*>(s_o + sw128(row, hh * 4 + 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(s_o + sw128(row, hh * 4 + 3)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bars + B_O_READY);
      tick(3);
    };

    for (int G = 0; G < n_groups; ++G) {
      const int g = G & 3, it = G >> 2;
      const int tile = first + it * stride;
      mbar_wait(bars + B_S_DONE, G & 1);                    // also: every product issued before S(G) has retired, i.e.
      tc_fence_after();                                     // QKV(G) (ring slot G & 1 is free) and the previous tile's QKV
      tick(4);

      // ---- softmax of (row, head 2g + hh): scores of the row's own unit
      {
        const int head = g * 2 + hh;
        const uint32_t sb = tlane + (hh ? T_SB : T_SA);
        float s[TP];
        if (TP == 32) {
          uint32_t ra[16], rb[16];
          tmem_ld16(sb + dq * 32, ra);
          tmem_ld16(sb + dq * 32 + 16, rb);
          tmem_ld_wait();
          tick(13);
#pragma unroll
          for (int j = 0; j < 16; ++j) { s[j % TP] = __uint_as_float(ra[j]); s[(16 + j) % TP] = __uint_as_float(rb[j]); }
        } else {
          uint32_t ra[16], rb[16];
          tmem_ld16(sb + dq * 32, ra);
          tmem_ld16(sb + dq * 32 + 16, rb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j % TP] = __uint_as_float(lane < 16 ? ra[j] : rb[j]);
        }
        const int i = lane & (TP - 1);
        {
          constexpr int CPR = TP / 8, RPL = 8 / CPR;
          const uint8_t* brow = s_bias + (head * TP + i) * TP * 2;
#pragma unroll
          for (int c = 0; c < CPR; ++c) {
            const int cs = c ^ ((i / RPL) & (CPR - 1));
            const uint4 t = *reinterpret_cast<const uint4*>(brow + cs * 16);
            const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c2 = unpack_bf16(t.z), d = unpack_bf16(t.w);
            s[c * 8] += a.x; s[c * 8 + 1] += a.y; s[c * 8 + 2] += b.x; s[c * 8 + 3] += b.y;
            s[c * 8 + 4] += c2.x; s[c * 8 + 5] += c2.y; s[c * 8 + 6] += d.x; s[c * 8 + 7] += d.y;
          }
        }
        if (TEMPORAL) {
#pragma unroll
          for (int j = 0; j < TP; ++j)
            if (j >= p.T) s[j] = -1.0e30f;                  // padded frames are not keys
        } else {
          if (g == 0) {                                     // per tile: TP == 32, the warp's 32 lanes are one window
            const Unit w = decode(tile * NU + dq);
            same = 0xffffffffu;
            if (unit_masked(w)) {                           // -100 where the Swin region ids of query and key differ
              const int code = region_code(w, lane);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t bal = __ballot_sync(0xffffffffu, code == c);
                if (code == c) same = bal;
              }
            }
          }
          if (same != 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < TP; ++j)
              if (!((same >> j) & 1u)) s[j] += kMask;
          }
        }
        float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int j = 4; j < TP; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
        const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < TP; ++j) { s[j] = ex2(s[j] - m); l4[j & 3] += s[j]; }
        const float f = __fdividef(1.0f, (l4[0] + l4[1]) + (l4[2] + l4[3]));
        // P (bf16 pairs, K = 128 keys -> 64 columns) over the head's S columns: the row's own block, zeros elsewhere
        uint32_t pw[16];
        if (TP == 32) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pw[e] = pack_bf16(s[(2 * e) % TP] * f, s[(2 * e + 1) % TP] * f);
        } else {
          const bool hi = lane >= 16;                       // unit 2 dq (+1): words [0,8) or [8,16) of the 16-word block
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t v = pack_bf16(s[(2 * e) % TP] * f, s[(2 * e + 1) % TP] * f);
            pw[e] = hi ? 0u : v;
            pw[8 + e] = hi ? v : 0u;
          }
        }
        tick(14);
        {                                                   // (one x64 store with aliased zero registers measured slower)
          const uint32_t z[16] = {};
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b == dq) tmem_st16(sb + b * 16, pw);
            else tmem_st16(sb + b * 16, z);
          }
        }
        tmem_st_wait();
        tc_fence_before();
      }
      mbar_arrive(bars + B_P_READY);
      tick(5);
      // ---- O of this head, as soon as its PV product retires: the output projection and the next group's scores wait on it
      drain_o(G);
    }
    if (PROF && prof) g_prof_ws[15] = static_cast<unsigned long long>(n_groups);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_u, 512);
  }
}

template <int TP, bool TEMPORAL>
int launch_ws(const P32& p, cudaStream_t st) {
  constexpr int smem = Lay::total + 1024;
  static SmemConfigured configured;
  if (!configured.covers(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attn_ws32_kernel<TP, TEMPORAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_ws32_kernel<TP, TEMPORAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
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
    unsigned long long z[24] = {};
    cudaMemcpyToSymbol(g_prof_ws, z, sizeof(z));
  }
  CUtensorMap map_wq, map_wp;
  int rc = extdm_encode_matrix_map(&map_wq, p.wqkv, 3 * HID, C, 64);
  if (rc) return rc;
  rc = extdm_encode_matrix_map(&map_wp, p.wproj, C, HID, 64);
  if (rc) return rc;
  if (prof) attn_ws32_kernel<TP, TEMPORAL, true><<<grid, NTH, smem, st>>>(q, map_wq, map_wp);
  else attn_ws32_kernel<TP, TEMPORAL, false><<<grid, NTH, smem, st>>>(q, map_wq, map_wp);
  if (prof) {
    unsigned long long h[24];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_prof_ws, sizeof(h));
    const double n = h[15] ? static_cast<double>(h[15]) : 1.0;
    fprintf(stderr, "[attn32 ws prof] temporal=%d TP=%d groups/CTA=%.0f cycles/group  drain: wait_qkv %.0f drain_qkv %.0f wait_o %.0f "
            "(softmax grp) drain_o %.0f | tok_issue %.0f epilogue %.0f | softmax: ln %.0f wait_s %.0f tmem_ld %.0f math %.0f p_store %.0f | issue: wait_qk %.0f early_qkv %.0f wait_p %.0f "
            "wait_o_ready %.0f\n", (int)TEMPORAL, TP, n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[7] / n, h[12] / n, h[6] / n, h[4] / n, h[13] / n, h[14] / n, h[5] / n,
            h[8] / n, h[9] / n, h[10] / n, h[11] / n);
  }
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

}  // namespace
}  // namespace extdm

using namespace extdm;

// (2,4,4)-window attention layer, 8 heads x 32, C = 64.  Called by extdm_stw_fused.  Returns -1 when the geometry is not
// supported (the caller falls back).
int extdm_stw_ws32_launch(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                          const float* proj_bias, const float* bias_table, const float* rope_cos, const float* rope_sin,
                          int B, int T, int H, int W, int Cc, int sd, int sh, int sw, float eps, void* stream) {
  const int nww = W / 4, nwh = H / 4;
  if (Cc != 64 || W % 4 || H % 4 || (nww & (nww - 1)) || (nwh & (nwh - 1))) return -1;
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
  return launch_ws<32, false>(p, static_cast<cudaStream_t>(stream));
}

// Temporal attention layer, 8 heads x 32, C = 64, T <= 32.  Called by extdm_temporal_fused.
int extdm_temporal_ws32_launch(const void* x, void* y, const float* gamma, const float* ln_w, const float* ln_b,
                               const void* wqkv, const void* wout, const float* rel_bias, const float* rope_cos,
                               const float* rope_sin, int B, int T, int HW, int Cc, float eps, void* stream) {
  if (Cc != 64 || T < 1 || T > 32) return -1;
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
    return launch_ws<16, true>(p, static_cast<cudaStream_t>(stream));
  }
  p.n_tiles = (p.n_units + 3) / 4;
  return launch_ws<32, true>(p, static_cast<cudaStream_t>(stream));
}
