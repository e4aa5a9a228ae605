// Shifted-window attention layer, tcgen05 edition (C = 64, window (4,4,4) = 64 tokens, 8 heads x 16):
//     y = x + proj( WindowAttention3D( chanLN(x) ) )          Residual(PreNorm(STWAttentionLayer))
// reference: model/BaseDM_adaptor/DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:139-159, :409-560.
//
// ncu on the all-mma.sync kernel (stw_fused.cu) showed the legacy HMMA path saturated: on B200 mma.sync.m16n8k16
// issues about once per 27 cycles per scheduler (~165 TFLOP/s chip-wide, 1/13 of tcgen05), and 2/3 of the layer's
// matrix work are the two dense projections.  Here those run on the 5th-gen tensor cores:
//   * QKV projection  D[128 x 384] = LN(x)[128 x 64] . Wqkv^T   -- 8 tcgen05.mma (M=128, N=256 + N=128, K=16 x 4),
//     A = the window's normalised tokens written by the LayerNorm threads straight into a 128B-swizzled K-major
//     tile, B = Wqkv resident in shared memory, accumulators in TMEM, drained with tcgen05.ld (thread = token),
//     rotary + q-scale applied on the way to the bf16 Q/K/V tiles;
//   * output projection D[128 x 64] = O[128 x 128] . Wproj^T    -- 8 tcgen05.mma (N = 64, 2 k-blocks x 4),
//     A = the attention output written by the attention warps in swizzled K-major layout, epilogue (bias + residual)
//     from TMEM straight to global memory.
// Only rows 0..63 of the M = 128 tiles are real (one window per iteration); the other accumulator lanes are ignored.
// The 64 x 64 x 16 per-head score / PV products (block-diagonal, below any tcgen05 tile) stay on mma.sync with
// ldmatrix operands, bias-initialised accumulators and the ballot shift mask, two warps per head.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {
namespace {

__device__ __forceinline__ void t_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void t_ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void t_ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void t_cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ float t_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float t_bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float t_bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

constexpr float kL2e = 1.4426950408889634f;
constexpr int NTOK = 64, DH = 16, C = 64, HEADS = 8, HID = 128;
constexpr int NTH = 512;

// byte offset of 16-byte chunk j (0..7) of row r inside a K-major 128B-swizzled tile of 64 bf16 per row
__device__ __forceinline__ int sw128(int r, int j) { return (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4); }
// byte offset of 16-byte chunk j (0..15) of row r in a [rows][128] bf16 tile with the low 3 chunk bits XOR-swizzled
__device__ __forceinline__ int swrow256(int r, int j) { return r * 256 + ((j ^ (r & 7)) << 4); }

struct Smem {
  static constexpr int wqkv = 0;                       // [384][64] SW128                         49152
  static constexpr int wp = wqkv + 49152;              // 2 x [64][64] SW128                      16384
  static constexpr int a = wp + 16384;                 // [128][64] SW128 (LayerNorm output)      16384
  static constexpr int qo = a + 16384;                 // 2 x [128][64] SW128: Q, then O in place 32768
  static constexpr int k = qo + 32768;                 // 2 x [128][64] SW128                     32768
  static constexpr int v = k + 32768;                  //                                         32768
  static constexpr int raw = v + 32768;                // 2 x [128][64] bf16 raw tokens, chunk-swizzled 32768
  static constexpr int tbl = raw + 32768;              // [8][344] words: bias pairs (e-1, e) * log2(e) 11008
  static constexpr int rope = tbl + 11008;             // cos, sin [64][8] fp32                    4096
  static constexpr int misc = rope + 4096;             // gamma[64], pbias[64] fp32                 512
  static constexpr int emask = misc + 512;             // [2 windows][2 halves][8] u32              128
  static constexpr int bars = emask + 128;             // 2 mbarriers + tmem slot                    32
  static constexpr int total = bars + 32;
};
constexpr int XPR = 64;                                // raw pitch (bf16); 16-byte chunk j of token n sits at chunk j ^ (n & 7)
constexpr int TBLP = 344;                              // bias-table pitch per head (343 entries), one 32-bit word each:
                                                       // low half = entry e - 1, high half = entry e (adjacent keys)
__device__ __forceinline__ int raw_off(int n, int chunk) { return n * XPR + ((chunk ^ (n & 7)) << 3); }
__device__ __forceinline__ uint32_t t_lds32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// EXTDM_STW_PROF=1: per-phase cycle counts of CTA 0 (thread 0's view, barrier waits included), printed by the launcher
__device__ unsigned long long g_stw_prof[8];

struct TcParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* gamma;
  const __nv_bfloat16* wqkv;
  const __nv_bfloat16* wproj;
  const float* proj_bias;
  const float* bias_table;
  const float* rcos;
  const float* rsin;
  int B, T, H, W, sd, sh, sw, Dp, n_windows;
  int lw, lh;                  // log2 of the window counts along W and H (powers of two: checked by the launcher)
  float eps;
  int prof;
};

// byte offset of hidden-dim chunk j (0..15, 8 bf16 each) of token row r in the two-k-block swizzled [128][128] tile
__device__ __forceinline__ int sw2(int r, int j) { return (j >> 3) * 16384 + sw128(r, j & 7); }

// Two windows (128 tokens) per iteration: rows 0..63 of every M = 128 tile are window 2p, rows 64..127 window 2p+1.
__global__ void __launch_bounds__(NTH, 1) stw_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~uintptr_t(1023));
  uint8_t* s_wqkv = sm + Smem::wqkv;
  uint8_t* s_wp = sm + Smem::wp;
  uint8_t* s_a = sm + Smem::a;
  uint8_t* s_qo = sm + Smem::qo;
  uint8_t* s_k = sm + Smem::k;
  uint8_t* s_v = sm + Smem::v;
  __nv_bfloat16* s_raw = reinterpret_cast<__nv_bfloat16*>(sm + Smem::raw);
  float* s_cos = reinterpret_cast<float*>(sm + Smem::rope);
  float* s_sin = s_cos + NTOK * (DH / 2);
  float* s_gamma = reinterpret_cast<float*>(sm + Smem::misc);
  float* s_pbias = s_gamma + C;
  uint32_t* s_E = reinterpret_cast<uint32_t*>(sm + Smem::emask);
  uint64_t* bar_qkv = reinterpret_cast<uint64_t*>(sm + Smem::bars);
  uint64_t* bar_o = bar_qkv + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const bool shifted = (p.sd | p.sh | p.sw) != 0;
  const int nWw = p.W / 4, nWh = p.H / 4, nWd = p.Dp / 4;
  const int lrow = lane & 15, lhi = lane >> 4;

  // ---- one-time staging
  for (int i = tid; i < 384 * 8; i += NTH) {           // Wqkv [384][64] -> swizzled B tile
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(s_wqkv + sw128(r, j)) = *reinterpret_cast<const uint4*>(p.wqkv + r * C + j * 8);
  }
  for (int i = tid; i < 64 * 16; i += NTH) {           // Wproj [64][128] -> two swizzled k-blocks
    const int r = i >> 4, j = i & 15;
    *reinterpret_cast<uint4*>(s_wp + (j >> 3) * 8192 + sw128(r, j & 7)) =
        *reinterpret_cast<const uint4*>(p.wproj + r * HID + j * 8);
  }
  for (int i = tid; i < NTOK * (DH / 2); i += NTH) {    // global [pos][pair] -> shared [pair][pos]: lanes (= tokens) read
    s_cos[(i & 7) * NTOK + (i >> 3)] = p.rcos[i];       // consecutive words (the [pos][pair] layout was an 8-way bank conflict)
    s_sin[(i & 7) * NTOK + (i >> 3)] = p.rsin[i];
  }
  for (int i = tid; i < C; i += NTH) { s_gamma[i] = p.gamma[i]; s_pbias[i] = p.proj_bias ? p.proj_bias[i] : 0.f; }
  for (int i = tid; i < HEADS * TBLP; i += NTH) {
    const int h = i / TBLP, e = i % TBLP;
    const float hi = e < 343 ? p.bias_table[e * HEADS + h] * kL2e : 0.f;
    const float lo = e >= 1 ? p.bias_table[(e - 1) * HEADS + h] * kL2e : 0.f;
    reinterpret_cast<uint32_t*>(sm + Smem::tbl)[i] = pack_bf16(lo, hi);
  }
  if (tid == 0) {
    mbar_init(bar_qkv, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();                                  // weight tiles were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);

  struct Win { int b, id, ih, iw; };                    // b < 0: no such window (odd tail)
  auto decode = [&](int widx) {
    Win w;
    if (widx >= p.n_windows) { w.b = -1; w.id = w.ih = w.iw = 0; return w; }
    w.iw = widx & (nWw - 1); widx >>= p.lw;
    w.ih = widx & (nWh - 1); widx >>= p.lh;
    w.b = widx / nWd;
    w.id = widx - w.b * nWd;
    return w;
  };
  auto src_pixel = [&](const Win& w, int n) -> int {
    if (w.b < 0) return -1;
    int od = w.id * 4 + (n >> 4) + p.sd, oh = w.ih * 4 + ((n >> 2) & 3) + p.sh, ow = w.iw * 4 + (n & 3) + p.sw;
    if (od >= p.Dp) od -= p.Dp;
    if (oh >= p.H) oh -= p.H;
    if (ow >= p.W) ow -= p.W;
    return od < p.T ? ((w.b * p.T + od) * p.H + oh) * p.W + ow : -1;
  };
  auto region_code = [&](const Win& w, int n) -> int {
    int c = 0;
    if (p.sd && w.id == nWd - 1 && (n >> 4) >= 4 - p.sd) c |= 1;
    if (p.sh && w.ih == nWh - 1 && ((n >> 2) & 3) >= 4 - p.sh) c |= 2;
    if (p.sw && w.iw == nWw - 1 && (n & 3) >= 4 - p.sw) c |= 4;
    return c;
  };
  auto win_masked = [&](const Win& w) -> bool {
    return shifted && w.b >= 0 && ((p.sd && w.id == nWd - 1) || (p.sh && w.ih == nWh - 1) || (p.sw && w.iw == nWw - 1));
  };
  auto prefetch_pair = [&](int pair, int buf) {
    __nv_bfloat16* dst = s_raw + buf * 128 * XPR;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * NTH;                     // 128 tokens x 8 chunks: k = 0 -> window 2p, k = 1 -> 2p+1
      const int n = i >> 3, c8 = i & 7;
      const Win w = decode(2 * pair + k);
      const int s = src_pixel(w, n & 63);
      t_cp_async16(dst + raw_off(n, c8), p.x + (s >= 0 ? static_cast<long long>(s) * C + c8 * 8 : 0), s >= 0 ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int n_pairs = (p.n_windows + 1) / 2;
  const float qscale = 0.25f * kL2e;                    // dh^-1/2 * log2(e)
  constexpr float kMask = -100.0f * kL2e;
  constexpr uint32_t idesc256 = umma_idesc_bf16(128, 256), idesc128 = umma_idesc_bf16(128, 128),
                     idesc64 = umma_idesc_bf16(128, 64);
  // attention role of this warp: one (window, head)
  const int awin = warp >> 3, head = warp & 7;
  const int hchunk = (head & 3) * 2, hkb = head >> 2;   // this head's 16 hidden dims = chunks hchunk, hchunk+1 of k-block hkb
  // relative-position table offsets (lin(i) - lin(j) + 171): per-thread parts, the rest are compile-time immediates
  const int ltg = (tg >> 1) * 7 + (tg & 1) * 2;
  const int pt0 = (g >> 2) * 7 + (g & 3) - ltg + 171, pt1 = pt0 + 14;
  const uint32_t tb0 = smem_u32(sm + Smem::tbl) + static_cast<uint32_t>(head * TBLP + pt0) * 4u;
  const uint32_t tb1 = tb0 + static_cast<uint32_t>(pt1 - pt0) * 4u;
  // drain / epilogue role: TMEM lane quarter warp & 3 (thread = token), column group warp >> 2
  const int dq = warp & 3, de = warp >> 2;
  const int dtok = dq * 32 + lane;                      // 0..127 (window dtok >> 6, position dtok & 63)
  const uint32_t tlane = tmem_u + (static_cast<uint32_t>(dq * 32) << 16);

  // channel LayerNorm of one pair: 4 threads per token, two 16-byte chunks each -> swizzled A tile
  auto layernorm_pair = [&](int pr, const __nv_bfloat16* raw) {
    const int n = tid >> 2, part = tid & 3;
    const Win w = decode(2 * pr + (n >> 6));
    float v[16];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const uint4 t = *reinterpret_cast<const uint4*>(raw + raw_off(n, part * 2 + h2));
      const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
      v[h2 * 8] = a.x; v[h2 * 8 + 1] = a.y; v[h2 * 8 + 2] = b.x; v[h2 * 8 + 3] = b.y;
      v[h2 * 8 + 4] = c.x; v[h2 * 8 + 5] = c.y; v[h2 * 8 + 6] = d.x; v[h2 * 8 + 7] = d.y;
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) sum += v[j];
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float mean = sum * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float dd = v[j] - mean; sq += dd * dd; }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    const float rstd = src_pixel(w, n & 63) >= 0 ? rsqrtf(sq * (1.0f / C) + p.eps) : 0.f;
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = part * 16 + h2 * 8 + 2 * j;
        pk[j] = pack_bf16((v[h2 * 8 + 2 * j] - mean) * rstd * s_gamma[cc],
                          (v[h2 * 8 + 2 * j + 1] - mean) * rstd * s_gamma[cc + 1]);
      }
      *reinterpret_cast<uint4*>(s_a + sw128(n, part * 2 + h2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  };
  // QKV projection of the pair whose normalised tokens sit in s_a: D[:, 0:256] (Q | K) and D[:, 256:384] (V)
  auto issue_qkv = [&]() {
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_sw128(smem_u32(s_a));
        const uint64_t db0 = umma_desc_sw128(smem_u32(s_wqkv));
        const uint64_t db1 = umma_desc_sw128(smem_u32(s_wqkv) + 256 * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_u, da + 2 * k, db0 + 2 * k, idesc256, k > 0 ? 1u : 0u);
          umma_bf16(tmem_u + 256, da + 2 * k, db1 + 2 * k, idesc128, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_qkv);
      }
      __syncwarp();
    }
  };
  // one 16-column accumulator chunk (one head of Q, K or V) -> rotary / q-scale -> bf16 operand tile
  auto emit_chunk = [&](const uint32_t (&rr)[16], int region, int hd) {
    const int pos = dtok & 63;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(rr[j]);
    if (region < 2) {
      const float sc = region == 0 ? qscale : 1.0f;
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {
        const float cs = s_cos[pr * NTOK + pos], sn = s_sin[pr * NTOK + pos];
        const float x0 = f[2 * pr] * sc, x1 = f[2 * pr + 1] * sc;
        f[2 * pr] = x0 * cs - x1 * sn;
        f[2 * pr + 1] = x1 * cs + x0 * sn;
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
    uint8_t* dst = region == 0 ? s_qo : (region == 1 ? s_k : s_v);
    *reinterpret_cast<uint4*>(dst + sw2(dtok, hd * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(dst + sw2(dtok, hd * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  };
  auto drain_region = [&](int region) {                 // this warp's two heads (2 de, 2 de + 1) of Q, K or V
    uint32_t ra[16], rb[16];
    tmem_ld16(tlane + region * 128 + de * 32, ra);
    tmem_ld16(tlane + region * 128 + de * 32 + 16, rb);
    tmem_ld_wait();
    emit_chunk(ra, region, 2 * de);
    emit_chunk(rb, region, 2 * de + 1);
  };
  // output-projection epilogue of a finished pair: TMEM columns 384.. + bias + residual -> global
  auto epilogue = [&](int d) {
    uint32_t rr[16];
    tmem_ld16(tlane + 384 + de * 16, rr);
    tmem_ld_wait();
    if (d >= 0) {
      const uint4* xp = reinterpret_cast<const uint4*>(p.x + static_cast<long long>(d) * C + de * 16);
      const uint4 r0v = xp[0], r1v = xp[1];
      const uint32_t rw[8] = {r0v.x, r0v.y, r0v.z, r0v.w, r1v.x, r1v.y, r1v.z, r1v.w};
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 rv = unpack_bf16(rw[j]);
        pk[j] = pack_bf16(__uint_as_float(rr[2 * j]) + s_pbias[de * 16 + 2 * j] + rv.x,
                          __uint_as_float(rr[2 * j + 1]) + s_pbias[de * 16 + 2 * j + 1] + rv.y);
      }
      __nv_bfloat16* yp = p.y + static_cast<long long>(d) * C + de * 16;
      *reinterpret_cast<uint4*>(yp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(yp + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  };

  // Software pipeline over this CTA's pairs (it = 0, 1, ...):
  //   phase A(it): drain K, V of pair it | LayerNorm of pair it+1 | wait out-proj(it-1) | drain Q | epilogue(it-1)
  //   barrier S3 ; issue QKV MMA(it+1) ; prefetch raw tokens of pair it+2
  //   phase B(it): attention(it) ; barrier S4 ; issue out-proj MMA(it)
  // so both tcgen05 products and the global loads run under the attention of a neighbouring pair.
  int pair = blockIdx.x;
  uint32_t it = 0;
  int d_prev = -1;
  if (pair < n_pairs) {
    prefetch_pair(pair, 0);
    if (pair + static_cast<int>(gridDim.x) < n_pairs) prefetch_pair(pair + gridDim.x, 1);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    layernorm_pair(pair, s_raw);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    issue_qkv();
  }
  for (; pair < n_pairs; pair += gridDim.x, ++it) {
    const Win w0 = decode(2 * pair), w1 = decode(2 * pair + 1);
    const int nxt = pair + gridDim.x;
    long long tk[8];
    const bool prof = p.prof && blockIdx.x == 0 && tid == 0;
    if (prof) tk[0] = clock64();
    if (warp < 4) {                                     // region-membership words of window warp>>1, token half warp&1
      const Win& w = (warp >> 1) ? w1 : w0;
      if (win_masked(w)) {
        const int code = region_code(w, (warp & 1) * 32 + lane);
        uint32_t mine = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t bal = __ballot_sync(0xffffffffu, code == c);
          if (lane == c) mine = bal;
        }
        if (lane < 8) s_E[warp * 8 + lane] = mine;
      }
    }
    const int d_cur = src_pixel((dtok >> 6) ? w1 : w0, dtok & 63);
    mbar_wait(bar_qkv, it & 1);
    tc_fence_after();
    if (prof) tk[1] = clock64();
    drain_region(1);
    drain_region(2);
    if (nxt < n_pairs) layernorm_pair(nxt, s_raw + ((it + 1) & 1) * 128 * XPR);
    if (prof) tk[2] = clock64();
    if (it > 0) {
      mbar_wait(bar_o, (it - 1) & 1);                   // out-proj(it-1) done: its A operand (the Q / O slots) is free
      tc_fence_after();
    }
    if (prof) tk[3] = clock64();
    drain_region(0);
    if (it > 0) epilogue(d_prev);
    d_prev = d_cur;
    fence_proxy_async();                                // next pair's A tile -> visible to the tensor-core proxy
    tc_fence_before();
    __syncthreads();                                    // S3: Q / K / V tiles complete, TMEM drained, s_a written
    if (prof) tk[4] = clock64();
    if (nxt < n_pairs) {
      issue_qkv();
      if (nxt + static_cast<int>(gridDim.x) < n_pairs) prefetch_pair(nxt + gridDim.x, it & 1);
    }

    // ---- attention core (mma.sync): warp = (window awin, head); all four query m-tiles; O overwrites Q in place
    {
      const Win& aw = awin ? w1 : w0;
      const bool has_mask = win_masked(aw);
      const int rb = awin * 64;                         // first row of this window in the tiles
      uint32_t kfrag[8][2];
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4];
        const int r = rb + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
        t_ldsm_x4(kf, s_k + hkb * 16384 + sw128(r, hchunk + ((lane >> 3) & 1)));
        kfrag[2 * np][0] = kf[0]; kfrag[2 * np][1] = kf[1];
        kfrag[2 * np + 1][0] = kf[2]; kfrag[2 * np + 1][1] = kf[3];
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;        // window-local query rows
        uint32_t qa[4];
        t_ldsm_x4(qa, s_qo + hkb * 16384 + sw128(rb + mt * 16 + lrow, hchunk + lhi));
        // scores start from the relative-position bias: tbl[lin(i) - lin(j) + 171], adjacent keys = adjacent entries
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int kk = mt * 49 - (nt >> 1) * 49 - (nt & 1) * 14;      // compile-time part of the table index
          const uint32_t w0b = t_lds32(tb0 + kk * 4), w1b = t_lds32(tb1 + kk * 4);
          s[nt][0] = t_bf_hi(w0b);                      // key column 2*tg     -> entry e
          s[nt][1] = t_bf_lo(w0b);                      // key column 2*tg + 1 -> entry e - 1
          s[nt][2] = t_bf_hi(w1b);
          s[nt][3] = t_bf_lo(w1b);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) t_mma16816(s[nt], qa, kfrag[nt][0], kfrag[nt][1]);
        if (has_mask) {
          const int c0 = region_code(aw, r0), c1 = region_code(aw, r1);
          const uint32_t* E = s_E + awin * 16;
          const uint32_t e0[2] = {~E[c0], ~E[8 + c0]}, e1[2] = {~E[c1], ~E[8 + c1]};
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
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
        for (int nt = 0; nt < 8; ++nt) {
          m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
          m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          s[nt][0] = t_exp2(s[nt][0] - m0);
          s[nt][1] = t_exp2(s[nt][1] - m0);
          s[nt][2] = t_exp2(s[nt][2] - m1);
          s[nt][3] = t_exp2(s[nt][3] - m1);
          l0 += s[nt][0] + s[nt][1];
          l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = __frcp_rn(l0), inv1 = __frcp_rn(l1);
        float o[2][4];
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          uint32_t a[4];
          a[0] = pack_bf16(s[2 * ps][0], s[2 * ps][1]);
          a[1] = pack_bf16(s[2 * ps][2], s[2 * ps][3]);
          a[2] = pack_bf16(s[2 * ps + 1][0], s[2 * ps + 1][1]);
          a[3] = pack_bf16(s[2 * ps + 1][2], s[2 * ps + 1][3]);
          uint32_t vb[4];
          t_ldsm_x4_trans(vb, s_v + hkb * 16384 + sw128(rb + ps * 16 + lrow, hchunk + lhi));
          t_mma16816(o[0], a, vb[0], vb[1]);
          t_mma16816(o[1], a, vb[2], vb[3]);
        }
        // head output over this head's Q slots (this warp has consumed them): the A tile of the output projection
        __syncwarp();
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
          *reinterpret_cast<uint32_t*>(s_qo + hkb * 16384 + sw128(rb + r0, hchunk + dt) + tg * 4) =
              pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
          *reinterpret_cast<uint32_t*>(s_qo + hkb * 16384 + sw128(rb + r1, hchunk + dt) + tg * 4) =
              pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
        }
      }
    }
    if (prof) tk[5] = clock64();
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // raw tokens of the pair after next: visible after S4
    fence_proxy_async();
    __syncthreads();                                    // S4: O tile complete
    if (prof) tk[6] = clock64();

    // ---- output projection on tcgen05: D_o[128 x 64] at TMEM column 384; consumed in the next iteration's phase A
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = umma_desc_sw128(smem_u32(s_qo) + kb * 16384);
          const uint64_t db = umma_desc_sw128(smem_u32(s_wp) + kb * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_u + 384, da + 2 * k, db + 2 * k, idesc64, (kb | k) ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    if (prof) {
      tk[7] = clock64();
      for (int i = 0; i < 7; ++i) g_stw_prof[i] += static_cast<unsigned long long>(tk[i + 1] - tk[i]);
      g_stw_prof[7] += 1;
    }
  }
  if (it > 0) {                                         // epilogue of this CTA's last pair
    mbar_wait(bar_o, (it - 1) & 1);
    tc_fence_after();
    epilogue(d_prev);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace extdm

using namespace extdm;

// Launch the tcgen05 window-attention layer (C = 64, (4,4,4) windows, 8 heads x 16).  Called by extdm_stw_fused.
int extdm_stw_tc_launch(const void* x, void* y, const float* gamma, const void* wqkv, const void* wproj,
                        const float* proj_bias, const float* bias_table, const float* rope_cos, const float* rope_sin,
                        int B, int T, int H, int W, int sd, int sh, int sw, float eps, void* stream) {
  TcParams p;
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
  p.sd = sd; p.sh = sh; p.sw = sw;
  p.Dp = (T + 3) / 4 * 4;
  p.n_windows = B * (p.Dp / 4) * (H / 4) * (W / 4);
  p.eps = eps;
  const int nww = W / 4, nwh = H / 4;
  if ((nww & (nww - 1)) || (nwh & (nwh - 1)) || W % 4 || H % 4) {
    extdm_set_error("stw_tc: H/4 and W/4 must be powers of two", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  p.lw = 0; p.lh = 0;
  while ((1 << p.lw) < nww) ++p.lw;
  while ((1 << p.lh) < nwh) ++p.lh;
  constexpr int smem = Smem::total + 1024;
  static SmemConfigured configured;
  const int sms = device_sm_count();
  if (!configured.covers(smem)) {
    cudaError_t e = cudaFuncSetAttribute(stw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem);
  }
  const int n_pairs = (p.n_windows + 1) / 2;
  const int grid = n_pairs < sms ? n_pairs : sms;
  static const bool prof_env = getenv("EXTDM_STW_PROF") != nullptr;
  bool prof = prof_env;
  if (prof) {                                           // the profile read-back synchronises: illegal during graph capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cs);
    if (cs != cudaStreamCaptureStatusNone) prof = false;
  }
  p.prof = prof ? 1 : 0;
  if (prof) {
    unsigned long long z[8] = {};
    cudaMemcpyToSymbol(g_stw_prof, z, sizeof(z));
  }
  stw_tc_kernel<<<grid, NTH, smem, static_cast<cudaStream_t>(stream)>>>(p);
  if (prof) {
    unsigned long long h[8];
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    cudaMemcpyFromSymbol(h, g_stw_prof, sizeof(h));
    const double n = h[7] ? static_cast<double>(h[7]) : 1.0;
    fprintf(stderr, "[stw_tc prof] B=%d T=%d H=%d W=%d shift=%d pairs/CTA=%.0f cycles/pair: masks+qkv_wait %.0f drainKV+ln %.0f "
            "proj_wait %.0f drainQ+epilogue+S3 %.0f attention(+issue qkv) %.0f S4 %.0f issue_proj %.0f\n", B, T, H, W, sd | sh | sw, n, h[0] / n, h[1] / n,
            h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n);
  }
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}
