// Implicit-GEMM convolution / linear layer for sm_100a:
//   TMA (tiled, 128B swizzle, OOB = zero padding)  ->  smem ring  ->  tcgen05.mma (bf16, fp32 accumulators in TMEM)
//   ->  tcgen05.ld epilogue (bias, residual, per-sample column affine, activation)  ->  global.
// One CTA = one 128-row x BN-column output tile.  Warp 0: TMA producer.  Warp 1: TMEM allocator + MMA issuer.
// Warps 2..5: epilogue (warp w owns TMEM lanes 32*(w%4) .. +31 = tile rows).
// See include/extdm_b200.h (ExtdmGemm) for the operand description and the reference call sites replaced.
#include "common.cuh"
#include "../../include/extdm_b200.h"

#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace extdm {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 B = one swizzle row
constexpr int kATileBytes = kTileM * kBlockK * 2; // 16 KB
constexpr int kGemmThreads = 192;

struct GemmDev {
  int nk0, nk1, ntaps;
  int box[4], start[4], count[4], ntile[4];
  signed char tap[64][4];
  int n;
  int n_tiles_n, total_tiles;
  void* out;
  int out_fp32;
  long long out_base, out_stride[4];
  int col_group;
  long long col_group_stride;
  const float* bias;
  const void* res;
  int res_fp32;
  long long res_base, res_stride[4];
  const float* col_scale;
  const float* col_shift;
  int act;
  float* gn_part;
  // halo kernel (k x k 'same' convolutions whose 128-row tile is bh full-width rows of one frame)
  int kh, kw, stages, a_ext_bytes;
  int n_phase;    // >= 1: independent products over the same A tiles (ExtdmGemm.n_phase), phase = tile index digit
  long long phase_out[4];
  int tf32;       // operands are fp32, product runs as kind::tf32 (ExtdmGemm.tf32)
  int prefetch;   // L2 prefetch of the tile two ahead (helps the multi-block shapes, measured per shape)
  int dbg;   // profiling experiments only (EXTDM_GEMM_DBG): 1 = no epilogue stores, 2 = no MMA issue, 4 = no TMA loads,
             // 8 = no tcgen05.ld, 16 = no GroupNorm partial reduction, 32 = plain arrives instead of tcgen05.commit,
             // 64 = invert the L2-prefetch decision
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return silu(v);
  if (act == 3) return 1.0f / (1.0f + __expf(-v));
  return v;
}

struct TileCoord {
  int c1, c2, c3, c4, n0, m_tile, phase;
};
// tile index digits, fastest first: n-tile, phase, D1 .. D4 tile -- neighbouring CTAs share the A tile
__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int tile, int BN) {
  TileCoord t;
  const int nt = tile % p.n_tiles_n;
  int tidx = tile / p.n_tiles_n;
  if (p.n_phase > 1) {
    t.phase = tidx % p.n_phase;
    tidx /= p.n_phase;
  } else {
    t.phase = 0;
  }
  t.m_tile = tidx;
  const int t0 = tidx % p.ntile[0]; tidx /= p.ntile[0];
  const int t1 = tidx % p.ntile[1]; tidx /= p.ntile[1];
  const int t2 = tidx % p.ntile[2]; tidx /= p.ntile[2];
  t.c1 = p.start[0] + t0 * p.box[0];
  t.c2 = p.start[1] + t1 * p.box[1];
  t.c3 = p.start[2] + t2 * p.box[2];
  t.c4 = p.start[3] + tidx * p.box[3];
  t.n0 = nt * BN;
  return t;
}

// Epilogue warps (4 x 32 threads; warp w owns TMEM lanes 32*(w%4)..+31 = tile rows): drain the accumulator of each
// tile this CTA owns, apply bias / residual / affine / activation, store, optionally emit GroupNorm partials.
// SIMPLE: bias (+ activation) -> bf16 rows, n % 16 == 0, no residual / column affine / column groups -- the
// convolution + GroupNorm-statistics case, kept free of the general path's code (the epilogue is instruction-fetch
// sensitive: with one CTA per SM only these four warps hide each other's latency).
template <int BN, bool GN, bool SIMPLE>
__device__ __forceinline__ void epilogue_loop(const GemmDev& p, uint64_t* acc_full, uint64_t* acc_empty, float* s_gn,
                                              uint32_t tmem_base, int warp, int lane) {
  float* s_bias = s_gn + 128;             // [BN] bias of the current n-tile (zeros when there is no bias)
  int staged_n0 = -1;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  const int q = warp & 3;                 // TMEM lane quarter this warp may access
  const int m = q * 32 + lane;            // tile row
  int r = m;
  const int i1 = r % p.box[0]; r /= p.box[0];
  const int i2 = r % p.box[1]; r /= p.box[1];
  const int i3 = r % p.box[2]; r /= p.box[2];
  const int i4 = r;
  constexpr int kChunks = (BN + 15) / 16;
  int local = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
    const int as = local & 1;
    const uint32_t aphase = (local >> 1) & 1;
    const TileCoord tc = decode_tile(p, tile, BN);
    const int n0 = tc.n0;
    const int g1 = tc.c1 + i1, g2 = tc.c2 + i2, g3 = tc.c3 + i3, g4 = tc.c4 + i4;
    const bool row_ok = g1 < p.start[0] + p.count[0] && g2 < p.start[1] + p.count[1] &&
                        g3 < p.start[2] + p.count[2] && g4 < p.start[3] + p.count[3];
    const long long orow = p.out_base + p.phase_out[tc.phase] + g1 * p.out_stride[0] + g2 * p.out_stride[1] +
                           g3 * p.out_stride[2] + g4 * p.out_stride[3];
    const long long rrow = p.res_base + p.phase_out[tc.phase] + g1 * p.res_stride[0] + g2 * p.res_stride[1] +
                           g3 * p.res_stride[2] + g4 * p.res_stride[3];
    // The accumulator is drained in groups of 4 chunks (64 columns): the group body is unrolled (static register
    // indices for the double-buffered tcgen05.ld and the prefetched residual), the group loop is not (code size).
    // residual prefetch: bf16 residuals 64 columns (8 x 16 B) at a time, fp32 residuals 32 columns
    uint4 rpre[8];
    const bool res_vec = !SIMPLE && p.res != nullptr && row_ok && p.col_group >= p.n && (n0 + BN <= p.n);
    if (n0 != staged_n0) {                // uniform over the 128 epilogue threads
      if (staged_n0 >= 0) asm volatile("bar.sync 2, 128;" ::: "memory");        // previous tile's readers are done
#pragma unroll
      for (int j = m; j < BN; j += 128) s_bias[j] = (p.bias && n0 + j < p.n) ? __ldg(p.bias + n0 + j) : 0.f;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      staged_n0 = n0;
    }
    auto prefetch_res = [&](int ch) {
      if (!res_vec) return;
      if (p.res_fp32) {
        const float* rp = reinterpret_cast<const float*>(p.res) + rrow + n0 + ch * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (ch * 16 + j * 4 < BN) rpre[j] = *reinterpret_cast<const uint4*>(rp + j * 4);
      } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + rrow + n0 + ch * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (ch * 16 + j * 8 < BN) rpre[j] = *reinterpret_cast<const uint4*>(rp + j * 8);
      }
    };
    prefetch_res(0);

    mbar_wait(&acc_full[as], aphase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + as * kAccCols + (static_cast<uint32_t>(q * 32) << 16);
    float* sg = s_gn + as * 64;
    uint32_t raw[2][16];
    tmem_ld16(taddr, raw[0]);
    constexpr int kGroups = (kChunks + 3) / 4;
    constexpr int kCpg = BN >= 64 ? BN / 8 : 8;          // GroupNorm(8) group width in columns
    constexpr int kNV = 2 * (64 / kCpg);                 // statistics per 64-column group (sums, sums of squares)
#pragma unroll 1
    for (int grp = 0; grp < kGroups; ++grp) {
      float gst[kNV];
      if (GN) {
#pragma unroll
        for (int j = 0; j < kNV; ++j) gst[j] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < kChunks) {
          const int ch = grp * 4 + c;
          tmem_ld_wait();
          if (ch + 1 < kChunks) tmem_ld16(taddr + (ch + 1) * 16, raw[(c + 1) & 1]);
          const int nb = n0 + ch * 16;
          if (row_ok && nb < p.n) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[c & 1][j]);
            const long long coff =
                SIMPLE ? nb : static_cast<long long>(nb / p.col_group) * p.col_group_stride + (nb % p.col_group);
            const bool full = SIMPLE || (nb + 16 <= p.n);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(s_bias + ch * 16 + j);
              v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
            }
            if (!SIMPLE && p.res) {
              if (res_vec) {
                if (p.res_fp32) {
                  const int j0 = (c & 1) * 4;
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint4 t = rpre[j0 + j];
                    v[j * 4] += __uint_as_float(t.x); v[j * 4 + 1] += __uint_as_float(t.y);
                    v[j * 4 + 2] += __uint_as_float(t.z); v[j * 4 + 3] += __uint_as_float(t.w);
                  }
                } else {
                  const int j0 = c * 2;
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const uint4 t = rpre[j0 + j];
                    const float2 a = unpack_bf16(t.x), b2 = unpack_bf16(t.y), c2 = unpack_bf16(t.z),
                                 d2 = unpack_bf16(t.w);
                    v[j * 8] += a.x; v[j * 8 + 1] += a.y; v[j * 8 + 2] += b2.x; v[j * 8 + 3] += b2.y;
                    v[j * 8 + 4] += c2.x; v[j * 8 + 5] += c2.y; v[j * 8 + 6] += d2.x; v[j * 8 + 7] += d2.y;
                  }
                }
              } else if (p.res_fp32) {
                const float* rp = reinterpret_cast<const float*>(p.res) + rrow + coff;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (nb + j < p.n) v[j] += rp[j];
              } else {
                const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + rrow + coff;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (nb + j < p.n) v[j] += __bfloat162float(rp[j]);
              }
            }
            if (!SIMPLE && p.col_scale) {
              const float* cs = p.col_scale + static_cast<long long>(g4) * p.n + nb;
              const float* cb = p.col_shift + static_cast<long long>(g4) * p.n + nb;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (full || nb + j < p.n) v[j] = v[j] * __ldg(cs + j) + __ldg(cb + j);
            }
            if (p.act) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
            }
            if (!SIMPLE && p.out_fp32) {
              if (p.tf32 == 2) {                        // the consumer is another tf32 product: round, do not truncate
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
              }
              float* op = reinterpret_cast<float*>(p.out) + orow + coff;
              if (full) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (nb + j < p.n) op[j] = v[j];
              }
            } else {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow + coff;
              if (full) {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                if (GN) {
                  // statistics of the values as stored (bf16-rounded)
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float2 f = unpack_bf16(pk[j]);
                    const int g = (c * 16 + 2 * j) / kCpg;           // GroupNorm group within this 64-column group
                    gst[g] += f.x + f.y;
                    gst[kNV / 2 + g] += f.x * f.x + f.y * f.y;
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (nb + j < p.n) op[j] = __float2bfloat16(v[j]);
              }
            }
          }
          // next residual group (issued after this chunk's use of rpre)
          if (!SIMPLE) {
            if (p.res_fp32) {
              if ((c & 1) == 1 && ch + 1 < kChunks) prefetch_res(ch + 1);
            } else {
              if (c == 3 && ch + 1 < kChunks) prefetch_res(ch + 1);
            }
          }
        }
      }
      if (GN) {
        // fixed-order butterfly over the warp: kNV per-thread values -> one total per value, held by the lanes
        // whose high bits spell the value index (MSB first); lanes with zero low bits publish it.
        int idx = 0, off = 16;
#pragma unroll
        for (int half = kNV / 2; half >= 1; half >>= 1, off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send = up ? gst[i] : gst[i + half];
            const float keep = up ? gst[i + half] : gst[i];
            gst[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
          idx = idx * 2 + (up ? 1 : 0);
        }
        const int low_mask = 2 * off - 1;                 // offsets not consumed by the transposing steps
        for (; off >= 1; off >>= 1) gst[0] += __shfl_xor_sync(0xffffffffu, gst[0], off);
        if ((lane & low_mask) == 0) {
          const int stat = idx / (kNV / 2), gl = idx % (kNV / 2);       // 0 = sum, 1 = sum of squares
          sg[q * 16 + stat * 8 + grp * (kNV / 2) + gl] = gst[0];
        }
      }
    }
    // accumulator drained (all tcgen05.ld completed): hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[as]);

    if (GN) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (m < 16)
        p.gn_part[static_cast<long long>(tc.m_tile) * 16 + m] = (sg[m] + sg[16 + m]) + (sg[32 + m] + sg[48 + m]);
    }
  }
}


// Multi-warp SIMPLE epilogue for the one-CTA-per-SM halo kernel: EPW warps share each TMEM lane quarter and take the
// 16-column chunks round-robin (ncu: with a single epilogue warp per scheduler the ~900 dependent instructions per
// tile issue at one per ~5 cycles and the epilogue, not the tensor pipe or L2, bounds the kernel).
// bias (+ activation) -> bf16 rows; optional GroupNorm(8) partial sums of the stored values.
template <int BN, bool GN, int EPW, int NACC = 2>
__device__ __forceinline__ void epilogue_simple(const GemmDev& p, uint64_t* acc_full, uint64_t* acc_empty, float* s_gn,
                                                uint32_t tmem_base, int warp, int lane) {
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  constexpr int kChunks = BN / 16;
  // The EPW warps of a TMEM lane quarter form two groups: group 0 drains the even tiles of this CTA (accumulator
  // stage 0), group 1 the odd ones (stage 1), so the per-tile latency chain of one group (barrier wake-up, tcgen05.ld,
  // stores) overlaps the other group's tile.  Within a group the warps take the 16-column chunks round-robin.
  constexpr int kGroups = EPW >= NACC ? NACC : (EPW >= 2 ? 2 : 1);   // NACC accumulator stages -> NACC tile groups
  constexpr int kWpg = EPW / kGroups;      // warps per (quarter, group)
  constexpr int kGThreads = 128 * kWpg;    // threads per group
  const int q = warp & 3;                  // TMEM lane quarter
  const int e = (warp - 2) >> 2;           // which of the EPW warps of this quarter
  const int grp = e / kWpg, eg = e % kWpg;
  float* s_bias = s_gn + grp * 256;        // [BN] per group
  const int gt = (eg * 4 + ((warp - 2) & 3)) * 32 + lane;     // thread index inside the group
  const int m = q * 32 + lane;             // tile row
  int r = m;
  const int i1 = r % p.box[0]; r /= p.box[0];
  const int i2 = r % p.box[1]; r /= p.box[1];
  const int i3 = r % p.box[2]; r /= p.box[2];
  const int i4 = r;
  int staged_n0 = -1;
  int local = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
    if (kGroups > 1 && (local % kGroups) != grp) continue;
    const int as = local % NACC;
    const uint32_t aphase = (local / NACC) & 1;
    const TileCoord tc = decode_tile(p, tile, BN);
    const int n0 = tc.n0;
    const int g1 = tc.c1 + i1, g2 = tc.c2 + i2, g3 = tc.c3 + i3, g4 = tc.c4 + i4;
    const bool row_ok = g1 < p.start[0] + p.count[0] && g2 < p.start[1] + p.count[1] &&
                        g3 < p.start[2] + p.count[2] && g4 < p.start[3] + p.count[3];
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + p.out_base + p.phase_out[tc.phase] +
                          g1 * p.out_stride[0] + g2 * p.out_stride[1] + g3 * p.out_stride[2] + g4 * p.out_stride[3] + n0;
    if (n0 != staged_n0) {                 // uniform over the group's threads; named barrier 2 + grp
      if (staged_n0 >= 0) asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGThreads) : "memory");
      for (int j = gt; j < BN; j += kGThreads) s_bias[j] = (p.bias && n0 + j < p.n) ? __ldg(p.bias + n0 + j) : 0.f;
      asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGThreads) : "memory");
      staged_n0 = n0;
    }
    mbar_wait(&acc_full[as], aphase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + as * kAccCols + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int ch = eg; ch < kChunks; ch += kWpg) {
      uint32_t raw[16];
      if (p.dbg & 8) {
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = 0u;
      } else {
        tmem_ld16(taddr + ch * 16, raw);
        tmem_ld_wait();
      }
      float gst[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok && n0 + ch * 16 < p.n) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + ch * 16 + j);
          v[j] = __uint_as_float(raw[j]) + bb.x; v[j + 1] = __uint_as_float(raw[j + 1]) + bb.y;
          v[j + 2] = __uint_as_float(raw[j + 2]) + bb.z; v[j + 3] = __uint_as_float(raw[j + 3]) + bb.w;
        }
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
        if (!(p.dbg & 1)) {
          *reinterpret_cast<uint4*>(orow + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(orow + ch * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if (GN) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = unpack_bf16(pk[j]);
            gst[j >> 2] += f.x + f.y;
            gst[2 + (j >> 2)] += f.x * f.x + f.y * f.y;
          }
        }
      }
      if (GN && !(p.dbg & 16)) {
        // 4 values per thread -> warp totals (fixed-order butterfly); lanes 0, 8, 16, 24 hold value (lane >> 3) of
        // {sum lo8, sum hi8, sq lo8, sq hi8}.  Every warp publishes its own partial record slots straight to global
        // memory -- record (tile, lane quarter[, chunk parity for BN = 256]) of 16 floats = 8 group sums + 8 group sums
        // of squares -- so the epilogue warps never synchronise with each other; groupnorm_apply folds the records.
        {
          const bool up = (lane & 16) != 0;
          const float s0 = up ? gst[0] : gst[2], k0 = up ? gst[2] : gst[0];
          const float s1 = up ? gst[1] : gst[3], k1 = up ? gst[3] : gst[1];
          gst[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
          gst[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
        }
        {
          const bool up = (lane & 8) != 0;
          const float s0 = up ? gst[0] : gst[1], k0 = up ? gst[1] : gst[0];
          gst[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
        }
        gst[0] += __shfl_xor_sync(0xffffffffu, gst[0], 4);
        gst[0] += __shfl_xor_sync(0xffffffffu, gst[0], 2);
        gst[0] += __shfl_xor_sync(0xffffffffu, gst[0], 1);
        // Channels per GroupNorm group over ALL p.n output channels (the tile may be one of several n-tiles): a 16-column
        // chunk holds two groups (cpg 8), one group (16) or a 1/2, 1/4 slice of a group (32, 64).  Slices of one group go
        // to different records (krec per (m-tile, quarter)) so that no two warps ever add into the same float; every
        // float of every record is written by exactly one (n-tile, warp), groupnorm_apply folds the records.
        const int cpg = p.n >> 3;
        const int krec = cpg >= 32 ? (cpg >> 4) : 1;
        const int gcol = n0 + ch * 16;                 // first output channel of this chunk
        float* rec = p.gn_part + ((static_cast<long long>(tc.m_tile) * 4 + q) * krec + ((gcol >> 4) & (krec - 1))) * 16;
        const int stat = lane >> 4;                    // lanes 0 / 8 -> sums, 16 / 24 -> sums of squares
        if (cpg == 8) {
          if ((lane & 7) == 0) rec[stat * 8 + (gcol >> 3) + ((lane >> 3) & 1)] = gst[0];
        } else {
          const float tot = gst[0] + __shfl_xor_sync(0xffffffffu, gst[0], 8);       // lo8 + hi8 of this chunk
          if ((lane & 15) == 0) rec[stat * 8 + gcol / cpg] = tot;
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[as]);
  }
}

// Persistent CTAs: each loops over output tiles (tile = blockIdx.x + i*gridDim.x, n-tile fastest).  The TMA
// producer runs ahead through the smem ring across tile boundaries; the MMA issuer alternates between two TMEM
// accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
// GN: the epilogue also emits per-tile GroupNorm partial sums (8 groups over the BN == n columns) of the
// bf16-rounded output -- the statistics pass of Block.forward's GroupNorm costs no extra read of the tensor.
// SIMPLE launches use the multi-warp epilogue: EPW warps per TMEM lane quarter (the epilogue, not the tensor pipe,
// paces the short-K shapes), so the block is 64 + 128*EPW threads.
template <int BN, bool SIMPLE>
struct PlainCfg {
  static constexpr int kEpw = SIMPLE ? (BN == 256 ? 4 : 2) : 1;
  static constexpr int kThreads = 64 + 128 * kEpw;
};

template <int BN, int STAGES, bool GN, bool SIMPLE, bool TF32 = false>
__global__ void __launch_bounds__(PlainCfg<BN, SIMPLE>::kThreads, BN == 256 ? 1 : 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ GemmDev p) {
  constexpr int kBTileBytes = BN * kBlockK * 2;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_gn = reinterpret_cast<float*>(tmem_slot + 4);          // [2][4][16]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = p.nk0 + p.nk1;
  const int total_k = p.ntaps * nk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    if (p.nk1 > 0) tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], SIMPLE ? 4 * (PlainCfg<BN, SIMPLE>::kEpw / 2) : 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer (warp-uniform control flow, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile, BN);
        // optional L2 prefetch of the (un-shifted) A tile this CTA will need two tiles from now (GemmDev.prefetch)
        if (p.prefetch && tc.n0 == 0 && tile + 2 * gridDim.x < p.total_tiles) {
          const TileCoord tf = decode_tile(p, tile + 2 * gridDim.x, BN);
          const int o3 = p.tap[p.ntaps / 2][2];
          if (elect_one()) {
            for (int kc = 0; kc < nk; ++kc) {
              if (kc < p.nk0)
                tma_prefetch_5d(&map_a0, kc * kBlockK, tf.c1, tf.c2, tf.c3 + o3, tf.c4);
              else
                tma_prefetch_5d(&map_a1, (kc - p.nk0) * kBlockK, tf.c1, tf.c2, tf.c3 + o3, tf.c4);
            }
          }
        }
        const int tap0 = tc.phase * p.ntaps;      // phase: own taps, own weight rows
        const int wrow = tc.n0 + tc.phase * p.n;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int o1 = p.tap[tap0 + tap][0], o2 = p.tap[tap0 + tap][1], o3 = p.tap[tap0 + tap][2];
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (p.dbg & 4) {
              if (elect_one()) mbar_arrive(&full_bar[stage]);
            } else if (elect_one()) {
              uint8_t* sa = smem + stage * kStageBytes;
              uint8_t* sb = sa + kATileBytes;
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              if (kc < p.nk0)
                tma_load_5d(&map_a0, sa, &full_bar[stage], kc * kBlockK, tc.c1 + o1, tc.c2 + o2, tc.c3 + o3, tc.c4);
              else
                tma_load_5d(&map_a1, sa, &full_bar[stage], (kc - p.nk0) * kBlockK, tc.c1 + o1, tc.c2 + o2, tc.c3 + o3,
                            tc.c4);
              tma_load_2d(&map_b, sb, &full_bar[stage], (tap * nk + kc) * kBlockK, wrow);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (warp-uniform control flow, one elected lane issues)
    {
      constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(kTileM, BN < 16 ? 16 : BN) : umma_idesc_bf16(kTileM, BN < 16 ? 16 : BN);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);            // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + as * kAccCols;
        for (int it = 0; it < total_k; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const uint64_t da = umma_desc_sw128(sa);
            const uint64_t db = umma_desc_sw128(sa + kATileBytes);
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                // advance 16 bf16 = 32 B along K inside the 128B swizzle row: +2 in (addr >> 4) units
                if (TF32) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&acc_full[as]);
      }
    }
    __syncwarp();
  } else {
    if constexpr (SIMPLE)
      epilogue_simple<BN, GN, PlainCfg<BN, SIMPLE>::kEpw>(p, acc_full, acc_empty, s_gn, tmem_base, warp, lane);
    else
      epilogue_loop<BN, GN, false>(p, acc_full, acc_empty, s_gn, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------ halo variant
// k x k convolution with A re-use across the taps of one kernel column.  The plain kernel above re-reads the
// 128-row A tile from L2 once per tap (9x / 49x); ncu showed the level-0 convolutions pinned at the L2 throughput
// cap.  Here one pipeline stage is (kernel column kx, 64-channel block): TMA loads ONE box of bh + kh - 1 image rows
// shifted by kx; the kh taps of that column are the same smem tile read at row offsets ky*bw (multiples of 8 rows,
// so the 128B-swizzle phase is preserved and the UMMA descriptor just advances by ky*bw*128 bytes).
// RESB: the whole weight matrix (<= ~150 KB) is loaded once per CTA and stays resident, so a stage carries A only.
constexpr int kHaloEpw = 4;                                   // epilogue warps per TMEM lane quarter
constexpr int kHaloThreads = 64 + 128 * kHaloEpw;

template <int BN, bool GN, bool RESB, bool SIMPLE>
__global__ void __launch_bounds__(SIMPLE ? kHaloThreads : kGemmThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ GemmDev p) {
  constexpr int kBTileBytes = BN * kBlockK * 2;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  // accumulator stages: the accumulator hand-off (commit -> epilogue wake-up -> drain -> arrive -> MMA wake-up) is a
  // ~2.5 us latency chain per tile; four stages with one epilogue group each keep four tiles in flight
  constexpr int kAcc = SIMPLE ? (BN <= 64 ? 4 : (BN <= 128 ? 4 : 2)) : 2;
  constexpr uint32_t kTmemCols = kAcc * kAccCols;
  static_assert(kTmemCols <= 512, "TMEM budget");
  constexpr int kMaxStages = 8;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nk = p.nk0 + p.nk1;
  const int ntaps = p.kh * p.kw;
  const int resb_bytes = RESB ? ntaps * nk * kBTileBytes : 0;
  const int stage_bytes = p.a_ext_bytes + (RESB ? 0 : p.kh * kBTileBytes);
  uint8_t* s_resb = smem;
  uint8_t* s_stage = smem + resb_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stage + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* acc_full = empty_bar + kMaxStages;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* resb_bar = acc_empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resb_bar + 2);      // keeps s_gn / s_bias 16-byte aligned
  float* s_gn = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    if (p.nk1 > 0) tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], SIMPLE ? 4 * (kHaloEpw / kAcc) : 4);
    }
    mbar_init(resb_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer (warp-uniform control flow, one elected lane issues)
    {
      if (RESB) {
        if (elect_one()) {
          mbar_expect_tx(resb_bar, resb_bytes);
          for (int t = 0; t < ntaps * nk; ++t)
            tma_load_2d(&map_b, s_resb + t * kBTileBytes, resb_bar, t * kBlockK, 0);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile, BN);
        // optional L2 prefetch of the rows this CTA will need two tiles from now (GemmDev.prefetch)
        if (p.prefetch && tile + 2 * gridDim.x < p.total_tiles) {
          const TileCoord tf = decode_tile(p, tile + 2 * gridDim.x, BN);
          if (elect_one()) {
            for (int kc = 0; kc < nk; ++kc) {
              if (kc < p.nk0)
                tma_prefetch_5d(&map_a0, kc * kBlockK, tf.c1, tf.c2 - p.kh / 2, tf.c3, tf.c4);
              else
                tma_prefetch_5d(&map_a1, (kc - p.nk0) * kBlockK, tf.c1, tf.c2 - p.kh / 2, tf.c3, tf.c4);
            }
          }
        }
        for (int kx = 0; kx < p.kw; ++kx) {
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (p.dbg & 4) {
              if (elect_one()) mbar_arrive(&full_bar[stage]);
            } else if (elect_one()) {
              uint8_t* sa = s_stage + stage * stage_bytes;
              mbar_expect_tx(&full_bar[stage], stage_bytes);
              const int x0 = tc.c1 + kx - p.kw / 2, y0 = tc.c2 - p.kh / 2;
              if (kc < p.nk0)
                tma_load_5d(&map_a0, sa, &full_bar[stage], kc * kBlockK, x0, y0, tc.c3, tc.c4);
              else
                tma_load_5d(&map_a1, sa, &full_bar[stage], (kc - p.nk0) * kBlockK, x0, y0, tc.c3, tc.c4);
              if (!RESB) {
                uint8_t* sb = sa + p.a_ext_bytes;
                for (int ky = 0; ky < p.kh; ++ky)
                  tma_load_2d(&map_b, sb + ky * kBTileBytes, &full_bar[stage],
                              ((ky * p.kw + kx) * nk + kc) * kBlockK, tc.n0);
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (warp-uniform control flow, one elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN < 16 ? 16 : BN);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const int row_shift = p.box[0] * 128;            // bytes between the A views of consecutive kernel rows
      if (RESB) {
        mbar_wait(resb_bar, 0);
        tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int as = local % kAcc;
        const uint32_t aphase = (local / kAcc) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + as * kAccCols;
        uint32_t accumulate = 0;
        for (int kx = 0; kx < p.kw; ++kx) {
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = smem_u32(s_stage + stage * stage_bytes);
              uint32_t acc = accumulate;
              for (int ky = 0; ky < ((p.dbg & 2) ? 0 : p.kh); ++ky) {
                const uint32_t sb = RESB ? smem_u32(s_resb + ((ky * p.kw + kx) * nk + kc) * kBTileBytes)
                                         : sa + p.a_ext_bytes + ky * kBTileBytes;
                const uint64_t da = umma_desc_sw128(sa + ky * row_shift);
                const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, acc);
                  acc = 1;
                }
              }
              if (p.dbg & 32) mbar_arrive(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            }
            accumulate = 1;
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) { if (p.dbg & 32) mbar_arrive(&acc_full[as]); else umma_commit(&acc_full[as]); }
      }
    }
    __syncwarp();
  } else {
    if constexpr (SIMPLE)
      epilogue_simple<BN, GN, kHaloEpw, kAcc>(p, acc_full, acc_empty, s_gn, tmem_base, warp, lane);
    else
      epilogue_loop<BN, GN, false>(p, acc_full, acc_empty, s_gn, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapL2promotion promo) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    extdm_set_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return EXTDM_ERR_DRIVER;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu %llu %llu)", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull);
    extdm_set_error(buf, __FILE__, __LINE__);
    return EXTDM_ERR_DRIVER;
  }
  return EXTDM_OK;
}

static int encode_a(CUtensorMap* map, const void* base, int channels, const long long* dim, const long long* stride,
                    const int* box) {
  cuuint64_t dims[5] = {(cuuint64_t)channels, (cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2],
                        (cuuint64_t)dim[3]};
  cuuint64_t strides[4] = {(cuuint64_t)stride[0] * 2, (cuuint64_t)stride[1] * 2, (cuuint64_t)stride[2] * 2,
                           (cuuint64_t)stride[3] * 2};
  cuuint32_t bx[5] = {(cuuint32_t)kBlockK, (cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2],
                      (cuuint32_t)box[3]};
  return encode_map(map, base, 5, dims, strides, bx, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

static int sm_count() { return device_sm_count(); }

constexpr int kSmemBudget = 232448 - 4608;       // 227 KB minus alignment slack, barriers, bias and GN scratch

template <int BN, bool GN, bool RESB, bool SIMPLE>
static int launch_halo(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb, GemmDev& dev, int m_tiles,
                       int smem_bytes, cudaStream_t stream) {
  static SmemConfigured configured;
  if (!configured.covers(smem_bytes)) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, GN, RESB, SIMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem_bytes);
  }
  dev.n_tiles_n = (dev.n + BN - 1) / BN;
  dev.total_tiles = m_tiles * dev.n_tiles_n;
  const int resident = sm_count();
  const int grid = dev.total_tiles < resident ? dev.total_tiles : resident;
  conv_halo_kernel<BN, GN, RESB, SIMPLE><<<grid, SIMPLE ? kHaloThreads : kGemmThreads, smem_bytes, stream>>>(ma0, ma1, mb, dev);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

template <int BN, int STAGES, bool GN, bool SIMPLE, bool TF32 = false>
static int launch(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb, GemmDev& dev, int m_tiles,
                  cudaStream_t stream) {
  constexpr int kStageBytes = kATileBytes + BN * kBlockK * 2;
  constexpr int smem_bytes = STAGES * kStageBytes + (2 * STAGES + 4) * 8 + 16 + 4096 + 1024;
  static SmemConfigured configured;
  if (!configured.covers(smem_bytes)) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, GN, SIMPLE, TF32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured.set(smem_bytes);
  }
  dev.n_tiles_n = (dev.n + BN - 1) / BN;
  dev.total_tiles = m_tiles * dev.n_tiles_n;
  const int resident = sm_count() * (BN == 256 ? 1 : 2);
  const int grid = dev.total_tiles < resident ? dev.total_tiles : resident;
  conv_gemm_kernel<BN, STAGES, GN, SIMPLE, TF32><<<grid, PlainCfg<BN, SIMPLE>::kThreads, smem_bytes, stream>>>(ma0, ma1, mb,
                                                                                                     dev);
  EXTDM_CHECK_LAUNCH();
  return EXTDM_OK;
}

}  // namespace extdm

extern "C" int extdm_conv_gemm(const ExtdmGemm* g, void* stream_) {
  using namespace extdm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!g || !g->a0 || !g->w || !g->out) {
    extdm_set_error("extdm_conv_gemm: null operand", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  const int n_phase = g->n_phase > 1 ? g->n_phase : 1;
  if (n_phase > 4 || n_phase * g->ntaps > 64 || (n_phase > 1 && (g->gn_partials || g->tf32 || g->w_rows != n_phase * g->n))) {
    extdm_set_error("extdm_conv_gemm: n_phase <= 4, n_phase*ntaps <= 64, w_rows == n_phase*n, no gn_partials / tf32", __FILE__,
                    __LINE__);
    return EXTDM_ERR_ARG;
  }
  if (g->a0_channels <= 0 || g->a0_channels % kBlockK || g->a1_channels % kBlockK || g->ntaps < 1 || g->ntaps > 64 ||
      g->n < 1 || g->box[0] * g->box[1] * g->box[2] * g->box[3] != kTileM || g->col_group < 1) {
    extdm_set_error("extdm_conv_gemm: channels must be multiples of 64, 1..64 taps, box product 128", __FILE__,
                    __LINE__);
    return EXTDM_ERR_ARG;
  }
  GemmDev dev;
  memset(&dev, 0, sizeof dev);
  dev.nk0 = g->a0_channels / kBlockK;
  dev.nk1 = g->a1 ? g->a1_channels / kBlockK : 0;
  dev.ntaps = g->ntaps;
  int m_tiles = 1;
  for (int i = 0; i < 4; ++i) {
    if (g->box[i] < 1 || g->box[i] > 256 || g->count[i] < 1) {
      extdm_set_error("extdm_conv_gemm: bad box/count", __FILE__, __LINE__);
      return EXTDM_ERR_ARG;
    }
    dev.box[i] = g->box[i];
    dev.start[i] = g->start[i];
    dev.count[i] = g->count[i];
    dev.ntile[i] = (g->count[i] + g->box[i] - 1) / g->box[i];
    m_tiles *= dev.ntile[i];
    dev.out_stride[i] = g->out_stride[i];
    dev.res_stride[i] = g->res_stride[i];
  }
  memcpy(dev.tap, g->tap, sizeof dev.tap);
  dev.n = g->n;
  dev.out = g->out;
  dev.out_fp32 = g->out_fp32;
  dev.out_base = g->out_base;
  dev.col_group = g->col_group;
  dev.col_group_stride = g->col_group_stride;
  dev.bias = g->bias;
  dev.res = g->res;
  dev.res_fp32 = g->res_fp32;
  dev.res_base = g->res_base;
  dev.col_scale = g->col_scale;
  dev.col_shift = g->col_shift;
  dev.act = g->act;
  dev.gn_part = g->gn_partials;
  dev.tf32 = g->tf32;
  dev.n_phase = n_phase;
  for (int i = 0; i < 4; ++i) dev.phase_out[i] = (n_phase > 1 && i < n_phase) ? g->phase_out_offset[i] : 0;
  m_tiles *= n_phase;                              // a phase is one more digit of the tile index
  static const int dbg_flags = getenv("EXTDM_GEMM_DBG") ? atoi(getenv("EXTDM_GEMM_DBG")) : 0;
  dev.dbg = dbg_flags;
  // measured on B200 (profiles/kernel_table_r1.md): the prefetch pays when a tile streams two or more 64-channel blocks
  // (7x7 init_conv 9.9 -> 9.0 ms per round, two-source 3x3 5.8 -> 5.0 ms), not for single-block shapes
  dev.prefetch = ((dev.nk0 + dev.nk1 >= 2) ? 1 : 0) ^ ((dbg_flags & 64) ? 1 : 0);
  if ((g->col_scale == nullptr) != (g->col_shift == nullptr)) {
    extdm_set_error("extdm_conv_gemm: col_scale and col_shift go together", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }

  int bn = g->block_n;
  if (bn == 0) {
    bn = g->n <= 16 ? 16 : (g->n <= 64 ? 64 : (g->n <= 128 || g->n % 256 ? 128 : 256));
    // Small-M shapes (the 8 x 8 and 4 x 4 UNet levels, small batches): with few 128-row tiles a wide n-tile leaves most
    // SMs idle and the launch is latency bound (profiles/kernel_table_r2.md: rows = 15360, n = 256 at 47 TFLOP/s).
    // Narrower n-tiles multiply the CTA count; the A tile is re-read from L2 once per n-tile, which is free at these sizes.
    static const bool no_split = getenv("EXTDM_GEMM_NO_NSPLIT") != nullptr;
    // Measured on B200 (gpurun_out/configs_r2[e-h].md): halving while the narrower tiling still fits the resident CTA slots
    // (2 per SM) gains 3-5 % on BAIR / UCF / Cityscapes, but the K >= 4096 convolutions of the 512-channel SMMNIST level
    // lose 8 % with tiles narrower than one wave allows (BN <= 128 tiles are shared-memory bound).
    if (!no_split && !g->tf32) {
      const long long ktot_ = static_cast<long long>(g->ntaps) * (dev.nk0 + dev.nk1) * kBlockK;
      const long long slots = (ktot_ >= 4096 ? 1ll : 2ll) * sm_count();
      while (bn > 64 && g->n % (bn / 2) == 0 && static_cast<long long>(m_tiles) * (g->n / (bn / 2)) <= slots) bn /= 2;
    }
  }
  if (n_phase > 1 && g->n % bn) {
    extdm_set_error("extdm_conv_gemm: phases need block_n | n", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  if (bn != 16 && bn != 64 && bn != 128 && bn != 256) {
    extdm_set_error("extdm_conv_gemm: block_n must be 16, 64, 128 or 256", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }

  // ---- halo variant: a full k x k tap grid over tiles made of whole-width rows of a single frame
  int kk = g->ntaps == 9 ? 3 : (g->ntaps == 25 ? 5 : (g->ntaps == 49 ? 7 : 0));
  if (kk) {
    for (int t = 0; t < g->ntaps && kk; ++t)
      if (g->tap[t][0] != t % kk - kk / 2 || g->tap[t][1] != t / kk - kk / 2 || g->tap[t][2] != 0) kk = 0;
  }
  const bool simple = !g->res && !g->col_scale && !g->out_fp32 && g->col_group >= g->n && g->n % 16 == 0;
  static const bool halo_off = getenv("EXTDM_NO_HALO") != nullptr;
  static const bool halo7_off = getenv("EXTDM_NO_HALO7") != nullptr;
  // measured on B200: the halo kernel wins on every 3x3 / 7x7 shape it supports
  // (level-0 3x3: 96 vs 115 us; two sources: 160 vs 171 us; 7x7: 0.92 vs 1.12 ms)
  static const bool halo_all = getenv("EXTDM_HALO_ALL") != nullptr;
  bool halo = kk && n_phase == 1 && (simple || kk == 7) && !halo_off && !(kk == 7 && halo7_off) &&
              (halo_all || kk == 3 || dev.nk0 + dev.nk1 >= 2) && g->box[2] == 1 && g->box[3] == 1 && g->box[0] % 8 == 0 &&
              (bn == 64 || bn == 128) && !g->tf32;
  int ebox[4] = {g->box[0], g->box[1] + kk - 1, 1, 1};
  bool resb = false;
  int halo_smem = 0;
  if (halo) {
    const int nk = dev.nk0 + dev.nk1;
    const int a_ext = ebox[1] * ebox[0] * 128;
    const int btile = bn * kBlockK * 2;
    const long long resb_bytes = static_cast<long long>(g->ntaps) * nk * btile;
    int stages;
    if (g->n <= bn && resb_bytes + 3ll * a_ext <= kSmemBudget) {
      resb = true;
      stages = static_cast<int>((kSmemBudget - resb_bytes) / a_ext);
    } else {
      stages = kSmemBudget / (a_ext + kk * btile);
    }
    const long long per_kx = a_ext + (resb ? 0 : kk * btile);
    if (stages > 8) stages = 8;
    if (stages < 2) {
      halo = false;
    } else {
      dev.kh = dev.kw = kk;
      dev.stages = stages;
      dev.a_ext_bytes = a_ext;
      halo_smem = static_cast<int>((resb ? resb_bytes : 0) + static_cast<long long>(stages) * per_kx) +
                  (2 * 8 + 10) * 8 + 16 + 4096 + 1024;
    }
  }

  CUtensorMap ma0, ma1, mb;
  int rc = encode_a(&ma0, g->a0, g->a0_channels, g->a0_dim, g->a0_stride, halo ? ebox : g->box);
  if (rc) return rc;
  if (dev.nk1 > 0) {
    rc = encode_a(&ma1, g->a1, g->a1_channels, g->a1_dim, g->a1_stride, halo ? ebox : g->box);
    if (rc) return rc;
  } else {
    ma1 = ma0;
  }
  const long long ktot = static_cast<long long>(g->ntaps) * (dev.nk0 + dev.nk1) * kBlockK;
  cuuint64_t wd[2] = {(cuuint64_t)ktot, (cuuint64_t)g->w_rows};
  cuuint64_t ws[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t wb[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
  rc = encode_map(&mb, g->w, 2, wd, ws, wb, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc) return rc;

  if (g->gn_partials && ((g->n != 64 && g->n != 128 && g->n != 256 && g->n != 512) || g->n % bn || bn < 64 || !simple ||
                         g->box[3] != 1)) {
    extdm_set_error("extdm_conv_gemm: gn_partials needs n in {64,128,256,512}, block_n | n, a bias-only bf16 epilogue, box[3] == 1",
                    __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }
  if (halo) {
    const bool gn = g->gn_partials != nullptr;
#define HALO(BN_, GN_, RB_)                                                                             \
  return simple ? launch_halo<BN_, GN_, RB_, true>(ma0, ma1, mb, dev, m_tiles, halo_smem, stream)      \
                : launch_halo<BN_, GN_, RB_, false>(ma0, ma1, mb, dev, m_tiles, halo_smem, stream)
    if (bn == 64) {
      if (gn) { if (resb) HALO(64, true, true); else HALO(64, true, false); }
      else { if (resb) HALO(64, false, true); else HALO(64, false, false); }
    } else {
      if (gn) { if (resb) HALO(128, true, true); else HALO(128, true, false); }
      else { if (resb) HALO(128, false, true); else HALO(128, false, false); }
    }
#undef HALO
  }
  if (g->gn_partials) {
    // per-tile GroupNorm partials: one n-tile covering all channels, 8 groups, bf16 dense rows, tile within a sample
#define PLAIN(BN_, ST_, GN_)                                                                  \
  return simple ? launch<BN_, ST_, GN_, true>(ma0, ma1, mb, dev, m_tiles, stream)            \
                : launch<BN_, ST_, GN_, false>(ma0, ma1, mb, dev, m_tiles, stream)
    switch (bn) {
      case 64: PLAIN(64, 4, true);
      case 128: PLAIN(128, 3, true);
      default: PLAIN(256, 4, true);
    }
  }
  if (g->tf32) {                                        // fp32 operands (tf32 product), fp32 output: general epilogue
    if (g->gn_partials || !g->out_fp32 || g->res || g->col_scale) {
      extdm_set_error("extdm_conv_gemm: tf32 mode supports a bias / activation fp32 epilogue only", __FILE__, __LINE__);
      return EXTDM_ERR_ARG;
    }
    switch (bn) {
      case 16: return launch<16, 5, false, false, true>(ma0, ma1, mb, dev, m_tiles, stream);
      case 64: return launch<64, 4, false, false, true>(ma0, ma1, mb, dev, m_tiles, stream);
      case 128: return launch<128, 3, false, false, true>(ma0, ma1, mb, dev, m_tiles, stream);
      default: return launch<256, 4, false, false, true>(ma0, ma1, mb, dev, m_tiles, stream);
    }
  }
  switch (bn) {
    case 16: PLAIN(16, 5, false);
    case 64: PLAIN(64, 4, false);
    case 128: PLAIN(128, 3, false);
    default: PLAIN(256, 4, false);
  }
#undef PLAIN
}

// bf16 row-major matrix [rows][cols] -> 2-D tiled map with 128-byte swizzle, box = [box_rows][64 columns]; used by the
// attention kernels of the other translation units (declared in common.cuh)
int extdm_encode_matrix_map(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
  using namespace extdm;
  cuuint64_t d[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t st[1] = {(cuuint64_t)cols * 2};
  cuuint32_t bx[2] = {64u, (cuuint32_t)box_rows};
  return encode_map(map, base, 2, d, st, bx, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}
