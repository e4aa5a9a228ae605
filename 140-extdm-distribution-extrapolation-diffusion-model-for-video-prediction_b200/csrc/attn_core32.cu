// Window-attention core of the dim_head-32 UNet variants at the levels whose layer is NOT fused end to end (C = 128 / 256 /
// 512: attn_ws32.cu holds the C = 64 layers): qkv (rows, 768) bf16 -> out (rows, 256) bf16 for (2,4,4) windows, 8 heads x 32,
//   out = softmax(rot(q * dh^-1/2) . rot(k)^T + relative-position bias [+ Swin shift mask]) . v
// reference: model/BaseDM_adaptor/DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:462-497 (WindowAttention3D.forward),
// :522-560 (shift / partition / mask).  Both products run on tcgen05 with the scores kept COMPACT in tensor memory:
//
// One M = 128 tile = 4 windows of 32 tokens; heads in 4 groups of 2.  The score matrix of a tile is block diagonal (a
// row only attends to its own window), so S_h is computed by one N = 32 MMA per window whose disable-output-lane operand
// enables only that window's 32 rows -- all four windows write the SAME 32 columns, S_h is [128 x 32] instead of
// [128 x 128].  P = softmax(S) goes back as bf16 pairs over the first 16 of those columns and is the A operand (TS form)
// of O_h = P_h . V_h[window], again one lane-masked MMA per window.  S / P and O each have two buffers (by group parity),
// so S(G+1) is computed while group G is in its softmax and O(G-1) is drained after softmax(G).
//
// Roles (17 warps): warps 0-7 "prep" (raw q | k | v rows of the group by cp.async into a staging buffer one group ahead;
// q-scale + rotary; Q, K row-major and V transposed into 128B-swizzled operand tiles, double buffered), warps 8-15
// "softmax" (thread = query row x head; also drains O to global memory), warp 16 "issue".
// TMEM (256 columns): S / P buffers [0,64) [64,128), O buffers [128,192) [192,256); head hh of the group at + 32 hh.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/extdm_b200.h"

namespace extdm {
namespace {

constexpr float kL2e = 1.4426950408889634f;
constexpr int HEADS = 8, HID = 256, DH = 32, TP = 32, NU = 4;
constexpr int NTH = 544, NP = 256, NS = 256;               // threads: all (17 warps) / prep group / softmax group
constexpr uint32_t T_S = 0, T_O = 128;
constexpr int STG_PITCH = 400;                             // bytes per staged row: 3 x 64 bf16 = 384, padded against bank conflicts

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

struct Lay {
  static constexpr int q = 0;                                // 2 x [128][64] SW128: Q of the group (two heads), by group parity
  static constexpr int k = q + 2 * 16384;
  static constexpr int vt = k + 2 * 16384;                   // 2 x V^T: 2 key blocks x [64][64] SW128
  static constexpr int stage = vt + 2 * 16384;               // 2 x [128][STG_PITCH]: raw q | k | v of the group
  static constexpr int bias = stage + 2 * 128 * STG_PITCH;   // [8][32][32] bf16 * log2(e), chunk-swizzled
  static constexpr int rope = bias + 16384;                  // (cos, sin) float2 [16 pairs][32 positions]
  static constexpr int bars = rope + 4096;
  static constexpr int total = bars + 256;
};
enum Bar { B_QK_READY = 0, B_S_DONE = 2, B_P_READY = 4, B_O_DONE = 6, B_O_FREE = 8, B_COUNT = 10 };   // two of each

// -DEXTDM_CORE32_PROF: cycle counters of CTA 0 (one thread per role), printed by the launcher after a synchronise
#ifdef EXTDM_CORE32_PROF
__device__ unsigned long long g_core_prof[16];
#define CORE_TICK(slot)                                                        \
  do {                                                                         \
    if (prof_thread) {                                                         \
      const long long now_ = clock64();                                        \
      g_core_prof[slot] += static_cast<unsigned long long>(now_ - tk_);        \
      tk_ = now_;                                                              \
    }                                                                          \
  } while (0)
#else
#define CORE_TICK(slot) do {} while (0)
#endif

struct PC {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  const float* bias_table;     // [147][8]
  const float* rcos;
  const float* rsin;           // [32][16]
  int B, T, H, W;
  int sd, sh, sw, Dp, n_units, n_tiles, lw, lh;
};

__global__ void __launch_bounds__(NTH, 1) attn_core32_kernel(const __grid_constant__ PC p) {
  using L = Lay;
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~uintptr_t(1023));
  uint8_t* s_bias = sm + L::bias;
  float2* s_rope = reinterpret_cast<float2*>(sm + L::rope);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nWw = p.W / 4, nWh = p.H / 4, nWd = p.Dp / 2;
  const bool shifted = (p.sd | p.sh | p.sw) != 0;

  // ---- geometry: tile row r = window slot (r / 32), token n = r % 32 (same conventions as attn_ws32.cu)
  struct Unit { int b, id, ih, iw; };                       // b < 0: no such window (tail)
  auto decode = [&](int u) {
    Unit w;
    if (u >= p.n_units) { w.b = -1; w.id = w.ih = w.iw = 0; return w; }
    w.iw = u & (nWw - 1); u >>= p.lw;
    w.ih = u & (nWh - 1); u >>= p.lh;
    w.b = u / nWd;
    w.id = u - w.b * nWd;
    return w;
  };
  auto src_pixel = [&](const Unit& w, int n) -> int {
    if (w.b < 0) return -1;
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

  // ---- one-time staging: rope / bias tables, barriers, tensor memory
  for (int i = tid; i < 32 * 16; i += NTH) s_rope[(i & 15) * 32 + (i >> 4)] = make_float2(p.rcos[i], p.rsin[i]);
  {
    constexpr int CPR = TP / 8, RPL = 8 / CPR;
    for (int idx = tid; idx < HEADS * TP * TP; idx += NTH) {
      const int h = idx / (TP * TP), i = (idx / TP) % TP, j = idx % TP;
      const int e = ((i >> 4) - (j >> 4) + 1) * 49 + (((i >> 2) & 3) - ((j >> 2) & 3) + 3) * 7 + ((i & 3) - (j & 3) + 3);
      const float v = p.bias_table[e * HEADS + h];
      const int cs = (j >> 3) ^ ((i / RPL) & (CPR - 1));
      reinterpret_cast<__nv_bfloat16*>(s_bias)[(h * TP + i) * TP + cs * 8 + (j & 7)] = __float2bfloat16(v * kL2e);
    }
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + B_QK_READY + b, NP);
      mbar_init(bars + B_S_DONE + b, 1);
      mbar_init(bars + B_P_READY + b, NS);
      mbar_init(bars + B_O_DONE + b, 1);
      mbar_init(bars + B_O_FREE + b, NS);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, *tmem_slot, 0);

#ifdef EXTDM_CORE32_PROF
  const bool prof_thread = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8 || warp == 16);
  long long tk_ = clock64();
#endif
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = (p.n_tiles - first + stride - 1) / stride;        // >= 1: the grid never exceeds n_tiles
  const int n_groups = my_tiles * 4;

  if (warp == 16) {
    // =========================================================================================== issue
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, TP), idesc_pv = umma_idesc_bf16(128, DH);
      auto issue_s = [&](int G) {                           // S_h[window rows] = Q_h . K_h[window]^T
        const int b = G & 1;
        mbar_wait(bars + B_QK_READY + b, (G >> 1) & 1);
        if (G >= 2) mbar_wait(bars + B_O_DONE + b, ((G - 2) >> 1) & 1);   // PV(G-2) has consumed this buffer's P
        tc_fence_after();
        CORE_TICK(8);
        const uint64_t dq = umma_desc_sw128(smem_u32(sm + L::q + b * 16384));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t sd = tmem_u + T_S + b * 64 + hh * 32;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const uint64_t dk = umma_desc_sw128(smem_u32(sm + L::k + b * 16384) + u * TP * 128);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16_lanes(sd, dq + 4 * hh + 2 * ks, dk + 4 * hh + 2 * ks, idesc_s, ks ? 1u : 0u, u == 0 ? 0u : ~0u,
                              u == 1 ? 0u : ~0u, u == 2 ? 0u : ~0u, u == 3 ? 0u : ~0u);
          }
        }
        umma_commit(bars + B_S_DONE + b);
        CORE_TICK(9);
      };
      auto issue_pv = [&](int G) {                          // O_h[window rows] = P_h . V_h[window], P from tensor memory
        const int b = G & 1;
        mbar_wait(bars + B_P_READY + b, (G >> 1) & 1);
        if (G >= 2) mbar_wait(bars + B_O_FREE + b, ((G - 2) >> 1) & 1);   // O(G-2) has been read out of this buffer
        tc_fence_after();
        CORE_TICK(7);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t pa = tmem_u + T_S + b * 64 + hh * 32;
          const uint32_t od = tmem_u + T_O + b * 64 + hh * 32;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            // V^T: rows = (head, d), columns = the tile's 128 keys in two 64-key blocks; window u = keys [32 u, 32 u + 32)
            const uint64_t dv = umma_desc_sw128(smem_u32(sm + L::vt + b * 16384) + (u >> 1) * 8192 + hh * 4096) + 4 * (u & 1);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16_ts_lanes(od, pa + ks * 8, dv + 2 * ks, idesc_pv, ks ? 1u : 0u, u == 0 ? 0u : ~0u, u == 1 ? 0u : ~0u,
                                 u == 2 ? 0u : ~0u, u == 3 ? 0u : ~0u);
          }
        }
        umma_commit(bars + B_O_DONE + b);
        CORE_TICK(9);
      };
      // Scores run one group AHEAD of the softmax: S(G+1) is issued as soon as its operands are stored, i.e. while the
      // softmax warps are still on group G, and PV(G) the moment softmax(G) is done -- the softmax warps never wait for
      // the tensor pipe (a group's 32 small MMAs cost ~58 cycles each whatever their N: tools/ubench/mma_rate.cu)
      issue_s(0);
      for (int G = 0; G < n_groups; ++G) {
        if (G + 1 < n_groups) issue_s(G + 1);
        issue_pv(G);
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // =========================================================================================== prep
    const int row = tid >> 1, hh = tid & 1, pos = row & 31;
    const float qscale = 0.17677669529663687f * kL2e;       // dh^-1/2 * log2(e)
    int px = -1;
    auto request = [&](int G) {                             // q | k | v slices of (row, head hh) of group G
      const int g = G & 3;
      if (g == 0) px = row_pixel(first + (G >> 2) * stride, row);
      uint8_t* dst = sm + L::stage + (G & 1) * (128 * STG_PITCH) + row * STG_PITCH + hh * 64;
      const __nv_bfloat16* src = p.qkv + (px >= 0 ? static_cast<long long>(px) * (3 * HID) : 0) + g * 64 + hh * 32;
      const int nb = px >= 0 ? 16 : 0;
#pragma unroll
      for (int part = 0; part < 3; ++part)
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16(dst + part * 128 + j * 16, src + part * HID + j * 8, nb);
    };
    request(0);
    cp_commit();
    for (int G = 0; G < n_groups; ++G) {
      const int b = G & 1;
      if (G + 1 < n_groups) request(G + 1);                 // (px of the next tile replaces this tile's: not needed below)
      cp_commit();
      cp_wait<1>();                                         // this thread's copies of group G have landed (it reads only those)
      CORE_TICK(0);
      if (G >= 2) {                                         // operand tiles of parity b: Q / K read by S(G-2), V^T by PV(G-2)
        mbar_wait(bars + B_S_DONE + b, ((G - 2) >> 1) & 1);
        mbar_wait(bars + B_O_DONE + b, ((G - 2) >> 1) & 1);
      }
      CORE_TICK(1);
      const uint8_t* stg = sm + L::stage + b * (128 * STG_PITCH) + row * STG_PITCH + hh * 64;
#pragma unroll
      for (int part = 0; part < 2; ++part) {              // q, k: scale + rotary (interleaved pairs), row-major tiles
        uint8_t* dst = sm + (part == 0 ? L::q : L::k) + b * 16384;
        const float sc = part == 0 ? qscale : 1.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 t = *reinterpret_cast<const uint4*>(stg + part * 128 + c * 16);
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 v = unpack_bf16(w[j]);
            const float2 rp = s_rope[(c * 4 + j) * 32 + pos];
            const float x0 = v.x * sc, x1 = v.y * sc;
            o[j] = pack_bf16(x0 * rp.x - x1 * rp.y, x1 * rp.x + x0 * rp.y);
          }
          *reinterpret_cast<uint4*>(dst + sw128(row, hh * 4 + c)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      {                                                     // v: transposed, [head, d][key]
        uint8_t* dst = sm + L::vt + b * 16384 + (row >> 6) * 8192 + (row & 7) * 2;
        const int kc = (row & 63) >> 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 t = *reinterpret_cast<const uint4*>(stg + 256 + c * 16);
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = hh * 32 + c * 8 + 2 * j;
            *reinterpret_cast<uint16_t*>(dst + sw128(d, kc)) = static_cast<uint16_t>(w[j] & 0xffffu);
            *reinterpret_cast<uint16_t*>(dst + sw128(d + 1, kc)) = static_cast<uint16_t>(w[j] >> 16);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bars + B_QK_READY + b);
      CORE_TICK(2);
    }
    cp_wait<0>();
  } else {
    // =========================================================================================== softmax (+ O drain)
    const int sw_ = warp - 8;
    const int dq = sw_ & 3, hh = sw_ >> 2;                  // TMEM lane quarter, head of the group
    const int row = dq * 32 + lane;
    const uint32_t tlane = tmem_u + (static_cast<uint32_t>(dq * 32) << 16);
    constexpr float kMask = -100.0f * kL2e;
    uint32_t same = 0xffffffffu;                            // keys of this row's window that share its Swin region id
    int px_cur = -1, px_prev = -1;
    auto drain_o = [&](int G) {                             // O of (row, head) -> out[pixel][head * 32 ..]
      const int b = G & 1;
      mbar_wait(bars + B_O_DONE + b, (G >> 1) & 1);
      tc_fence_after();
      CORE_TICK(5);
      uint32_t ra[16], rb[16];
      const uint32_t ob = tlane + T_O + b * 64 + hh * 32;
      tmem_ld16(ob, ra);
      tmem_ld16(ob + 16, rb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bars + B_O_FREE + b);
      // drain_o(G) runs after softmax(G+1): when G is the last group of its tile, px_cur already belongs to the next tile
      const int px = ((G & 3) == 3 && G + 1 < n_groups) ? px_prev : px_cur;
      if (px >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long long>(px) * HID + ((G & 3) * 2 + hh) * DH);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
        dst[2] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[3] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      CORE_TICK(6);
    };

    for (int G = 0; G < n_groups; ++G) {
      const int b = G & 1, g = G & 3, it = G >> 2;
      const int tile = first + it * stride;
      if (g == 0) {                                         // per tile: output pixel of this row, Swin mask word
        px_prev = px_cur;
        px_cur = row_pixel(tile, row);
        const Unit w = decode(tile * NU + dq);
        same = 0xffffffffu;
        if (unit_masked(w)) {                               // -100 where the Swin region ids of query and key differ
          const int code = region_code(w, lane);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t bal = __ballot_sync(0xffffffffu, code == c);
            if (code == c) same = bal;
          }
        }
      }
      mbar_wait(bars + B_S_DONE + b, (G >> 1) & 1);
      tc_fence_after();
      CORE_TICK(3);
      {
        const int head = g * 2 + hh;
        const uint32_t sb = tlane + T_S + b * 64 + hh * 32;
        float s[TP];
        {
          uint32_t ra[16], rb[16];
          tmem_ld16(sb, ra);
          tmem_ld16(sb + 16, rb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) { s[j] = __uint_as_float(ra[j]); s[16 + j] = __uint_as_float(rb[j]); }
        }
        const int i = lane;
        {
          constexpr int CPR = TP / 8, RPL = 8 / CPR;
          const uint8_t* brow = s_bias + (head * TP + i) * TP * 2;
#pragma unroll
          for (int c = 0; c < CPR; ++c) {
            const int cs = c ^ ((i / RPL) & (CPR - 1));
            const uint4 t = *reinterpret_cast<const uint4*>(brow + cs * 16);
            const float2 a = unpack_bf16(t.x), bb = unpack_bf16(t.y), c2 = unpack_bf16(t.z), d = unpack_bf16(t.w);
            s[c * 8] += a.x; s[c * 8 + 1] += a.y; s[c * 8 + 2] += bb.x; s[c * 8 + 3] += bb.y;
            s[c * 8 + 4] += c2.x; s[c * 8 + 5] += c2.y; s[c * 8 + 6] += d.x; s[c * 8 + 7] += d.y;
          }
        }
        if (same != 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < TP; ++j)
            if (!((same >> j) & 1u)) s[j] += kMask;
        }
        float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int j = 4; j < TP; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
        const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < TP; ++j) { s[j] = ex2(s[j] - m); l4[j & 3] += s[j]; }
        const float f = __fdividef(1.0f, (l4[0] + l4[1]) + (l4[2] + l4[3]));
        uint32_t pw[16];                                    // P: bf16 pairs of the row's 32 keys over the first 16 S columns
#pragma unroll
        for (int e = 0; e < 16; ++e) pw[e] = pack_bf16(s[2 * e] * f, s[2 * e + 1] * f);
        tmem_st16(sb, pw);
        tmem_st_wait();
        tc_fence_before();
      }
      mbar_arrive(bars + B_P_READY + b);
      CORE_TICK(4);
      if (G >= 1) drain_o(G - 1);                           // its PV product retired while this group was normalised
    }
    drain_o(n_groups - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_u, 256);
  }
}

}  // namespace

bool attn_core32_supported(int heads, int dh, int wd, int wh, int ww, int H, int W) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  return heads == HEADS && dh == DH && wd == 2 && wh == 4 && ww == 4 && H % 4 == 0 && W % 4 == 0 && pow2(H / 4) && pow2(W / 4);
}

int attn_core32_launch(const void* qkv, void* out, const float* bias_table, const float* rope_cos, const float* rope_sin,
                       int B, int T, int H, int W, int sd, int sh, int sw, cudaStream_t st) {
  PC p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.bias_table = bias_table;
  p.rcos = rope_cos;
  p.rsin = rope_sin;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.sd = sd; p.sh = sh; p.sw = sw;
  p.Dp = (T + 1) / 2 * 2;
  p.n_units = B * (p.Dp / 2) * (H / 4) * (W / 4);
  p.n_tiles = (p.n_units + NU - 1) / NU;
  p.lw = 0;
  while ((1 << p.lw) < W / 4) ++p.lw;
  p.lh = 0;
  while ((1 << p.lh) < H / 4) ++p.lh;
  constexpr int smem = Lay::total + 1024;
  static SmemConfigured configured;
  if (!configured.covers(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attn_core32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem);
  }
  const int sms = device_sm_count();
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
#ifdef EXTDM_CORE32_PROF
  unsigned long long z[16] = {};
  cudaMemcpyToSymbol(g_core_prof, z, sizeof(z));
#endif
  attn_core32_kernel<<<grid, NTH, smem, st>>>(p);
#ifdef EXTDM_CORE32_PROF
  {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_core_prof, sizeof(h));
    const double n = 4.0 * ((p.n_tiles + grid - 1) / grid);
    fprintf(stderr, "[attn_core32 prof] tiles=%d grid=%d groups/CTA=%.0f cycles/group | prep: cp_wait %.0f bar_wait %.0f work %.0f | "
            "softmax: wait_s %.0f softmax %.0f wait_o %.0f drain_o %.0f | issue: wait_p %.0f wait_qk %.0f issue %.0f\n", p.n_tiles, grid, n,
            h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n);
  }
#endif
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

}  // namespace extdm
