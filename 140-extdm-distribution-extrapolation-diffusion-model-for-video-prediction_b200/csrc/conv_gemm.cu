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
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return silu(v);
  if (act == 3) return 1.0f / (1.0f + __expf(-v));
  return v;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, BN == 256 ? 1 : 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ GemmDev p) {
  constexpr int kBTileBytes = BN * kBlockK * 2;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  int tidx = blockIdx.x;
  int t0 = tidx % p.ntile[0]; tidx /= p.ntile[0];
  int t1 = tidx % p.ntile[1]; tidx /= p.ntile[1];
  int t2 = tidx % p.ntile[2]; tidx /= p.ntile[2];
  int t3 = tidx;
  const int c1 = p.start[0] + t0 * p.box[0];
  const int c2 = p.start[1] + t1 * p.box[1];
  const int c3 = p.start[2] + t2 * p.box[2];
  const int c4 = p.start[3] + t3 * p.box[3];
  const int n0 = blockIdx.y * BN;
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
    mbar_init(acc_bar, 1);
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
    // =========================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int o1 = p.tap[tap][0], o2 = p.tap[tap][1], o3 = p.tap[tap][2];
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kATileBytes;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          if (kc < p.nk0)
            tma_load_5d(&map_a0, sa, &full_bar[stage], kc * kBlockK, c1 + o1, c2 + o2, c3 + o3, c4);
          else
            tma_load_5d(&map_a1, sa, &full_bar[stage], (kc - p.nk0) * kBlockK, c1 + o1, c2 + o2, c3 + o3, c4);
          tma_load_2d(&map_b, sb, &full_bar[stage], (tap * nk + kc) * kBlockK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN < 16 ? 16 : BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total_k; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sa + kATileBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128B swizzle row: +2 in (addr >> 4) units
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_bar);
    }
  } else {
    // =========================== epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // tile row
    int r = m;
    const int i1 = r % p.box[0]; r /= p.box[0];
    const int i2 = r % p.box[1]; r /= p.box[1];
    const int i3 = r % p.box[2]; r /= p.box[2];
    const int i4 = r;
    const int g1 = c1 + i1, g2 = c2 + i2, g3 = c3 + i3, g4 = c4 + i4;
    const bool row_ok = g1 < p.start[0] + p.count[0] && g2 < p.start[1] + p.count[1] &&
                        g3 < p.start[2] + p.count[2] && g4 < p.start[3] + p.count[3];
    const long long orow = p.out_base + g1 * p.out_stride[0] + g2 * p.out_stride[1] + g3 * p.out_stride[2] +
                           g4 * p.out_stride[3];
    const long long rrow = p.res_base + g1 * p.res_stride[0] + g2 * p.res_stride[1] + g3 * p.res_stride[2] +
                           g4 * p.res_stride[3];
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    constexpr int kChunks = (BN + 15) / 16;
#pragma unroll 1
    for (int ch = 0; ch < kChunks; ++ch) {
      uint32_t raw[16];
      tmem_ld16(taddr + ch * 16, raw);
      tmem_ld_wait();
      const int nb = n0 + ch * 16;
      if (!row_ok || nb >= p.n) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
      const long long coff = static_cast<long long>(nb / p.col_group) * p.col_group_stride + (nb % p.col_group);
      const bool full = (nb + 16 <= p.n);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full || nb + j < p.n) v[j] += __ldg(p.bias + nb + j);
      }
      if (p.res) {
        if (p.res_fp32) {
          const float* rp = reinterpret_cast<const float*>(p.res) + rrow + coff;
          if (full) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 t = *reinterpret_cast<const float4*>(rp + j);
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          } else {
            for (int j = 0; j < 16 && nb + j < p.n; ++j) v[j] += rp[j];
          }
        } else {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + rrow + coff;
          if (full) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              uint4 t = *reinterpret_cast<const uint4*>(rp + j);
              float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), c = unpack_bf16(t.z), d = unpack_bf16(t.w);
              v[j] += a.x; v[j + 1] += a.y; v[j + 2] += b.x; v[j + 3] += b.y;
              v[j + 4] += c.x; v[j + 5] += c.y; v[j + 6] += d.x; v[j + 7] += d.y;
            }
          } else {
            for (int j = 0; j < 16 && nb + j < p.n; ++j) v[j] += __bfloat162float(rp[j]);
          }
        }
      }
      if (p.col_scale) {
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
      if (p.out_fp32) {
        float* op = reinterpret_cast<float*>(p.out) + orow + coff;
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 16 && nb + j < p.n; ++j) op[j] = v[j];
        }
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow + coff;
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 t;
            t.x = pack_bf16(v[j], v[j + 1]);
            t.y = pack_bf16(v[j + 2], v[j + 3]);
            t.z = pack_bf16(v[j + 4], v[j + 5]);
            t.w = pack_bf16(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(op + j) = t;
          }
        } else {
          for (int j = 0; j < 16 && nb + j < p.n; ++j) op[j] = __float2bfloat16(v[j]);
        }
      }
    }
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

template <int BN, int STAGES>
static int launch(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb, const GemmDev& dev,
                  int m_tiles, cudaStream_t stream) {
  constexpr int kStageBytes = kATileBytes + BN * kBlockK * 2;
  constexpr int smem_bytes = STAGES * kStageBytes + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) {
      extdm_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return EXTDM_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(m_tiles, (dev.n + BN - 1) / BN);
  conv_gemm_kernel<BN, STAGES><<<grid, kGemmThreads, smem_bytes, stream>>>(ma0, ma1, mb, dev);
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
  if ((g->col_scale == nullptr) != (g->col_shift == nullptr)) {
    extdm_set_error("extdm_conv_gemm: col_scale and col_shift go together", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }

  int bn = g->block_n;
  if (bn == 0) bn = g->n <= 16 ? 16 : (g->n <= 64 ? 64 : (g->n <= 128 || g->n % 256 ? 128 : 256));
  if (bn != 16 && bn != 64 && bn != 128 && bn != 256) {
    extdm_set_error("extdm_conv_gemm: block_n must be 16, 64, 128 or 256", __FILE__, __LINE__);
    return EXTDM_ERR_ARG;
  }

  CUtensorMap ma0, ma1, mb;
  int rc = encode_a(&ma0, g->a0, g->a0_channels, g->a0_dim, g->a0_stride, g->box);
  if (rc) return rc;
  if (dev.nk1 > 0) {
    rc = encode_a(&ma1, g->a1, g->a1_channels, g->a1_dim, g->a1_stride, g->box);
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

  switch (bn) {
    case 16: return launch<16, 5>(ma0, ma1, mb, dev, m_tiles, stream);
    case 64: return launch<64, 4>(ma0, ma1, mb, dev, m_tiles, stream);
    case 128: return launch<128, 3>(ma0, ma1, mb, dev, m_tiles, stream);
    default: return launch<256, 4>(ma0, ma1, mb, dev, m_tiles, stream);
  }
}
