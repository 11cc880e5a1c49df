// Persistent, warp-specialised implicit-GEMM engine for sm_100a.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer     (one elected thread, tcgen05.mma kind::f16, 3-term split-bf16)
//   warps 2..5  : epilogue       (tcgen05.ld TMEM -> regs -> bias/act/aux -> 128B-swizzled smem panels ->
//                                  TMA tile store; rows scattered through a rowmap, transposed / NCHW outputs and
//                                  launches with two outputs store straight from registers instead)
//
// One CTA per SM loops over output tiles (128 output positions x bn channels). The fp32
// accumulator lives in TMEM in 2..8 buffers (512 columns / tile width) so the epilogue of tile i overlaps the
// main loops of the following tiles — short-K layers (one or two K iterations per tile) are otherwise bound by
// the load -> MMA -> epilogue latency chain of a tile, not by bandwidth. Used for every Linear/Conv on the path (see include/fgt_b200.h).
#include "common.h"
#include "ptx.cuh"

namespace fgt {

constexpr int kMaxAMaps = 18;  // 2 segments x up to 9 stride phases
constexpr int kMaxTaps = 64;   // up to 7x7 (=49) taps
constexpr int kBK = 64;        // K elements per pipeline stage (128 bytes of bf16)
constexpr int kAPlaneBytes = 128 * 128;  // one A plane per stage: 128 rows x 128 B

struct TapDesc {
  int8_t phase;  // which stride-phase view
  int8_t dx, dy, dz;
};
struct SegDesc {
  int map_base;     // first tensor map of this segment (+ phase)
  int c_base;       // channel coordinate for group 0
  int c_per_group;  // channel advance per group
  int chunks;       // 64-channel chunks per tap
};

struct GemmParams {
  CUtensorMap a_maps[kMaxAMaps];
  CUtensorMap b_map;
  CUtensorMap out_map;  // epilogue TMA store: (C, X, Y, Z, plane) bf16 for a split output, (C, X, Y, Z) fp32
  CUtensorMap aux_map;  // epilogue TMA load of the fp32 aux operand (residual / gate), same geometry as an fp32 out_map
  TapDesc taps[kMaxTaps];
  SegDesc segs[2];
  int num_taps, num_segs, num_maps, k_iters;
  int tiles_x, tiles_y, tiles_z, n_tiles, total_tiles;
  int bn, bn_p2, stages;
  int n_acc;    // accumulator buffers in TMEM (power of two, n_acc * bn_p2 <= 512)
  int planes;   // 2: split-bf16 operands, 3 MMAs per K step; 1: single-plane fp16 operands, 1 MMA per K step
  int out_half; // the 16-bit output is ONE plane of fp16 (input of a 1-term layer) instead of split-bf16
  int tma_out;  // epilogue stores through shared memory + TMA (single output, unit channel stride, no rowmap)
  int aux_tma;    // the aux tile is fetched into the staging panels by TMA (fp32 output launches) instead of per-row loads
  int stg_slots;  // staging ring of the TMA epilogue: 1, 2 or 4 slots (a slot = the panels of one 64-channel chunk)
  int stg_slot_bytes;
  int box_w, box_h, a_rows;
  int N, cout_per_group;
  int out_w, out_h;
  int linear;  // 1: rows on x only (rowmap / lin_batch addressing)
  int lin_batch;
  int vec_ok;  // 16-byte vector stores allowed
  long long os_z, os_y, os_x, os_c;
  const int* rowmap;
  float* out_f32;
  __nv_bfloat16* out_hi;
  long long out_plane;
  const float* aux;
  const float* aux2;
  int aux_mode;
  const float* bias;
  int act;
  float alpha;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case FGT_ACT_LEAKY02: return v > 0.f ? v : 0.2f * v;
    case FGT_ACT_RELU: return fmaxf(v, 0.f);
    case FGT_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case FGT_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// Epilogue parameters hoisted into registers once per CTA.
struct EpiArgs {
  float* out_f32;
  __nv_bfloat16* out_hi;
  const float* aux;
  const float* aux2;
  long long out_plane, os_c;
  float alpha;
  int act, aux_mode, vec_ok, N, out_half;
};

// kExt = 1 adds the rarely used epilogue modes (pre-activation add, add+ReLU, GRU update, slope-0.01
// LeakyReLU); keeping them out of the common instantiation keeps it at ~150 registers with no spills.
template <int NC, int kExt>
__device__ __forceinline__ void epi_chunk(const EpiArgs& e, const uint32_t (&raw)[NC], const float* sb, long long off,
                                          int col0) {
  float v[NC];
#pragma unroll
  for (int q = 0; q < NC / 4; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
    v[4 * q] = fmaf(__uint_as_float(raw[4 * q]), e.alpha, b.x);
    v[4 * q + 1] = fmaf(__uint_as_float(raw[4 * q + 1]), e.alpha, b.y);
    v[4 * q + 2] = fmaf(__uint_as_float(raw[4 * q + 2]), e.alpha, b.z);
    v[4 * q + 3] = fmaf(__uint_as_float(raw[4 * q + 3]), e.alpha, b.w);
  }
  const bool vec = e.vec_ok && col0 + NC <= e.N;
  if (kExt && e.aux_mode == FGT_AUX_ADD_PRE) {  // residual added BEFORE the activation (LAFC edge head)
    if (vec) {
      const float4* ap = reinterpret_cast<const float4*>(e.aux + off + col0);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) {
        const float4 a = __ldg(ap + q);
        v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
      }
    } else {
      for (int j = 0; j < NC; ++j)
        if (col0 + j < e.N) v[j] += __ldg(e.aux + off + static_cast<long long>(col0 + j) * e.os_c);
    }
  }
  if (e.act == FGT_ACT_LEAKY02) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  } else if (kExt && e.act == FGT_ACT_LEAKY001) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
  } else if (e.act == FGT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (e.act == FGT_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
  } else if (e.act == FGT_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = tanhf(v[j]);
  }
  if (kExt && e.aux_mode == FGT_AUX_GRU_ZR) {  // fused z | r gates: two C-channel outputs (host guarantees vec)
    const int half = e.N >> 1;
    if (col0 < half) {
      float4* op = reinterpret_cast<float4*>(e.out_f32 + off + col0);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
      const long long o = off + (col0 - half);
      const float4* ap = reinterpret_cast<const float4*>(e.aux + o);
      uint4* hp = reinterpret_cast<uint4*>(e.out_hi + o);
      uint4* lp = reinterpret_cast<uint4*>(e.out_hi + e.out_plane + o);
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) {
        const float4 a0 = __ldg(ap + 2 * q), a1 = __ldg(ap + 2 * q + 1);
        uint32_t hw[4], lw[4];
        split_bf16x2(v[8 * q] * a0.x, v[8 * q + 1] * a0.y, hw[0], lw[0]);
        split_bf16x2(v[8 * q + 2] * a0.z, v[8 * q + 3] * a0.w, hw[1], lw[1]);
        split_bf16x2(v[8 * q + 4] * a1.x, v[8 * q + 5] * a1.y, hw[2], lw[2]);
        split_bf16x2(v[8 * q + 6] * a1.z, v[8 * q + 7] * a1.w, hw[3], lw[3]);
        hp[q] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        lp[q] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
    return;
  }
  if (vec) {
    const long long o = off + col0;
    if (e.aux_mode == FGT_AUX_ADD) {
      const float4* ap = reinterpret_cast<const float4*>(e.aux + o);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) {
        const float4 a = __ldg(ap + q);
        v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
      }
    } else if (e.aux_mode == FGT_AUX_MUL) {
      const float4* ap = reinterpret_cast<const float4*>(e.aux + o);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) {
        const float4 a = __ldg(ap + q);
        v[4 * q] *= a.x; v[4 * q + 1] *= a.y; v[4 * q + 2] *= a.z; v[4 * q + 3] *= a.w;
      }
    } else if (kExt && e.aux_mode == FGT_AUX_ADD_RELU) {
      const float4* ap = reinterpret_cast<const float4*>(e.aux + o);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) {
        const float4 a = __ldg(ap + q);
        v[4 * q] = fmaxf(v[4 * q] + a.x, 0.f); v[4 * q + 1] = fmaxf(v[4 * q + 1] + a.y, 0.f);
        v[4 * q + 2] = fmaxf(v[4 * q + 2] + a.z, 0.f); v[4 * q + 3] = fmaxf(v[4 * q + 3] + a.w, 0.f);
      }
    } else if (kExt && e.aux_mode == FGT_AUX_GRU) {  // h' = (1 - z) * h + z * q   (aux = h, aux2 = z, v = q)
      const float4* hp4 = reinterpret_cast<const float4*>(e.aux + o);
      const float4* zp4 = reinterpret_cast<const float4*>(e.aux2 + o);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) {
        const float4 h = __ldg(hp4 + q);
        const float4 z = __ldg(zp4 + q);
        v[4 * q] = (1.f - z.x) * h.x + z.x * v[4 * q];
        v[4 * q + 1] = (1.f - z.y) * h.y + z.y * v[4 * q + 1];
        v[4 * q + 2] = (1.f - z.z) * h.z + z.z * v[4 * q + 2];
        v[4 * q + 3] = (1.f - z.w) * h.w + z.w * v[4 * q + 3];
      }
    }
    if (e.out_f32) {
      float4* op = reinterpret_cast<float4*>(e.out_f32 + o);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    if (e.out_hi && e.out_half) {
      uint4* hp = reinterpret_cast<uint4*>(e.out_hi + o);
#pragma unroll
      for (int q = 0; q < NC / 8; ++q)
        hp[q] = make_uint4(pack_f16x2(v[8 * q], v[8 * q + 1]), pack_f16x2(v[8 * q + 2], v[8 * q + 3]),
                           pack_f16x2(v[8 * q + 4], v[8 * q + 5]), pack_f16x2(v[8 * q + 6], v[8 * q + 7]));
    } else if (e.out_hi) {
      uint4* hp = reinterpret_cast<uint4*>(e.out_hi + o);
      uint4* lp = reinterpret_cast<uint4*>(e.out_hi + e.out_plane + o);
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) split_bf16x2(v[8 * q + 2 * t], v[8 * q + 2 * t + 1], hw[t], lw[t]);
        hp[q] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        lp[q] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  } else {
    // generic path: strided channel stores (transposed / NCHW outputs) and N tails
    for (int j = 0; j < NC; ++j) {
      if (col0 + j < e.N) {
        const long long o = off + static_cast<long long>(col0 + j) * e.os_c;
        float x = v[j];
        if (e.aux_mode == FGT_AUX_ADD) x += __ldg(e.aux + o);
        else if (e.aux_mode == FGT_AUX_MUL) x *= __ldg(e.aux + o);
        else if (kExt && e.aux_mode == FGT_AUX_ADD_RELU) x = fmaxf(x + __ldg(e.aux + o), 0.f);
        else if (kExt && e.aux_mode == FGT_AUX_GRU) {
          const float z = __ldg(e.aux2 + o);
          x = (1.f - z) * __ldg(e.aux + o) + z * x;
        }
        if (e.out_f32) e.out_f32[o] = x;
        if (e.out_hi && e.out_half) {
          reinterpret_cast<__half*>(e.out_hi)[o] = __float2half_rn(x);
        } else if (e.out_hi) {
          __nv_bfloat16 h, l;
          split_bf16(x, h, l);
          e.out_hi[o] = h;
          e.out_hi[e.out_plane + o] = l;
        }
      }
    }
  }
}

// Register-only part of the epilogue for the TMA-store path (unit channel stride guaranteed): bias, activation and
// the aux operand for 32 consecutive channels of one output row; channels >= N (N tail) skip their aux loads.
template <int kExt>
__device__ __forceinline__ void epi_math(const EpiArgs& e, const uint32_t (&raw)[32], const float* sb, long long off,
                                         int col0, bool valid, float (&v)[32], uint32_t aux_smem = 0u, uint32_t sw = 0u) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
    v[4 * q] = fmaf(__uint_as_float(raw[4 * q]), e.alpha, b.x);
    v[4 * q + 1] = fmaf(__uint_as_float(raw[4 * q + 1]), e.alpha, b.y);
    v[4 * q + 2] = fmaf(__uint_as_float(raw[4 * q + 2]), e.alpha, b.z);
    v[4 * q + 3] = fmaf(__uint_as_float(raw[4 * q + 3]), e.alpha, b.w);
  }
  const long long o = off + col0;
  const bool aux_ok = valid && e.aux_mode != FGT_AUX_NONE;
  float4 a[8];
  if (aux_ok && aux_smem) {  // the tile's aux values were fetched by TMA into this thread's swizzled panel row
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = ld_shared_f4(aux_smem + ((static_cast<uint32_t>(q) ^ sw) << 4));
  } else if (aux_ok) {
    const float4* ap = reinterpret_cast<const float4*>(e.aux + o);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      a[q] = (col0 + 4 * q + 4 <= e.N) ? __ldg(ap + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (kExt && e.aux_mode == FGT_AUX_ADD_PRE) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[4 * q] += a[q].x; v[4 * q + 1] += a[q].y; v[4 * q + 2] += a[q].z; v[4 * q + 3] += a[q].w;
    }
  }
  if (e.act == FGT_ACT_LEAKY02) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  } else if (kExt && e.act == FGT_ACT_LEAKY001) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
  } else if (e.act == FGT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (e.act == FGT_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
  } else if (e.act == FGT_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
  if (e.aux_mode == FGT_AUX_ADD) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[4 * q] += a[q].x; v[4 * q + 1] += a[q].y; v[4 * q + 2] += a[q].z; v[4 * q + 3] += a[q].w;
    }
  } else if (e.aux_mode == FGT_AUX_MUL) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[4 * q] *= a[q].x; v[4 * q + 1] *= a[q].y; v[4 * q + 2] *= a[q].z; v[4 * q + 3] *= a[q].w;
    }
  } else if (kExt && e.aux_mode == FGT_AUX_ADD_RELU) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[4 * q] = fmaxf(v[4 * q] + a[q].x, 0.f); v[4 * q + 1] = fmaxf(v[4 * q + 1] + a[q].y, 0.f);
      v[4 * q + 2] = fmaxf(v[4 * q + 2] + a[q].z, 0.f); v[4 * q + 3] = fmaxf(v[4 * q + 3] + a[q].w, 0.f);
    }
  } else if (kExt && e.aux_mode == FGT_AUX_GRU) {  // h' = (1 - z) * h + z * q   (aux = h, aux2 = z, v = q)
    if (valid) {
      const float4* zp = reinterpret_cast<const float4*>(e.aux2 + o);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 z = (col0 + 4 * q + 4 <= e.N) ? __ldg(zp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * q] = (1.f - z.x) * a[q].x + z.x * v[4 * q];
        v[4 * q + 1] = (1.f - z.y) * a[q].y + z.y * v[4 * q + 1];
        v[4 * q + 2] = (1.f - z.z) * a[q].z + z.z * v[4 * q + 2];
        v[4 * q + 3] = (1.f - z.w) * a[q].w + z.w * v[4 * q + 3];
      }
    }
  }
}

// Debugging aid (fgt_debug_gemm_trace): clock64 timeline of one CTA, [role 0 producer | 1 MMA | 2 epilogue][local tile
// 0..63][4 slots]. Only the kTrace instantiation reads these; the production instantiations are unchanged by it.
__device__ long long* g_gemm_trace = nullptr;
__device__ int g_gemm_trace_cta = 0;

template <int kExt, int kTrace = 0>
__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  long long* trace = nullptr;
  if (kTrace) trace = (static_cast<int>(blockIdx.x) == g_gemm_trace_cta) ? g_gemm_trace : nullptr;
#define FGT_GTRACE(role, lt, slot) \
  if (kTrace && trace && (lt) < 64) trace[((role) * 64 + (lt)) * 4 + (slot)] = clock64()
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t b_plane_bytes = static_cast<uint32_t>(p.bn) * 128u;
  const uint32_t nplanes = static_cast<uint32_t>(p.planes);
  const uint32_t stage_bytes = nplanes * (kAPlaneBytes + b_plane_bytes);
  const uint32_t stg_ring = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;  // output panels (ring of chunks)
  const uint32_t bar_base = stg_ring + static_cast<uint32_t>(p.tma_out ? p.stg_slots * p.stg_slot_bytes : 0);
  // barrier layout: full[stages], empty[stages], acc_full[n_acc], acc_empty[n_acc], tmem slot
  const int n_acc = p.n_acc;
  const int acc_shift = 31 - __clz(n_acc);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + n_acc + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 2 * n_acc);
  // per-tile bias slice staged in smem (2 buffers x 256 floats), after the 256-byte barrier block
  float* sbias_base = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 256);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < n_acc; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 4);
    }
    fence_mbar_init();
    for (int i = 0; i < p.num_maps; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map);
    if (p.tma_out) tma_prefetch_desc(&p.out_map);
    if (p.aux_tma) {
      tma_prefetch_desc(&p.aux_map);
      mbar_init(bar_base + 8u * 30, 1);  // aux tile landed (one load in flight at a time)
      fence_mbar_init();
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, static_cast<uint32_t>(n_acc * p.bn_p2));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));
  // PDL: the prologue above overlapped the previous kernel's tail; nothing before this line touches global data
  pdl_launch_dependents();
  pdl_wait();

  const int tiles_per_z = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = nplanes * (static_cast<uint32_t>(p.a_rows) * 128u + b_plane_bytes);
      int plt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++plt) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int z = mt / tiles_per_z;
        const int rem = mt - z * tiles_per_z;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int x0 = tx * p.box_w, y0 = ty * p.box_h;
        const int n0 = nt * p.bn;
        const int group = n0 / p.cout_per_group;
        int kidx = 0;
        FGT_GTRACE(0, plt, 0);
        for (int t = 0; t < p.num_taps; ++t) {
          const TapDesc tap = p.taps[t];
          for (int sgi = 0; sgi < p.num_segs; ++sgi) {
            const SegDesc sg = p.segs[sgi];
            const CUtensorMap* amap = &p.a_maps[sg.map_base + tap.phase];
            const int c0 = sg.c_base + group * sg.c_per_group;
            for (int ch = 0; ch < sg.chunks; ++ch, ++kidx) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              if (kTrace && kidx == 0) { FGT_GTRACE(0, plt, 1); }
              const uint32_t sa = smem_base + stage * stage_bytes;
              const uint32_t sb = sa + nplanes * kAPlaneBytes;
              const uint32_t fb = full_bar(stage);
              mbar_expect_tx(fb, tx_bytes);
              const int cc = c0 + ch * kBK;
              tma_load_5d(sa, amap, fb, cc, x0 + tap.dx, y0 + tap.dy, z + tap.dz, 0);
              tma_load_3d(sb, &p.b_map, fb, kidx * kBK, n0, 0);
              if (nplanes == 2) {
                tma_load_5d(sa + kAPlaneBytes, amap, fb, cc, x0 + tap.dx, y0 + tap.dy, z + tap.dz, 1);
                tma_load_3d(sb + b_plane_bytes, &p.b_map, fb, kidx * kBK, n0, 1);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        FGT_GTRACE(0, plt, 2);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = p.planes == 1 ? umma_idesc_f16(128, p.bn) : umma_idesc_bf16(128, p.bn);
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & (n_acc - 1);
        const uint32_t aph = (lt >> acc_shift) & 1u;
        FGT_GTRACE(1, lt, 0);
        mbar_wait(acce_bar(buf), aph ^ 1u);
        tc_fence_after();
        FGT_GTRACE(1, lt, 1);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.bn_p2);
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (kTrace && it == 0) { FGT_GTRACE(1, lt, 2); }
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + nplanes * kAPlaneBytes;
          const uint64_t a_hi = umma_desc_sw128(sa);
          const uint64_t b_hi = umma_desc_sw128(sb);
          if (nplanes == 2) {
            const uint64_t a_lo = umma_desc_sw128(sa + kAPlaneBytes);
            const uint64_t b_lo = umma_desc_sw128(sb + b_plane_bytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t ko = static_cast<uint64_t>(k * 2);  // 32 bytes >> 4
              umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, (it | k) != 0);
              umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t ko = static_cast<uint64_t>(k * 2);
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, (it | k) != 0);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(accf_bar(buf));
        FGT_GTRACE(1, lt, 3);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int r = quarter * 32 + lane;
    EpiArgs ep;
    ep.out_f32 = p.out_f32; ep.out_hi = p.out_hi; ep.aux = p.aux; ep.aux2 = p.aux2; ep.out_plane = p.out_plane; ep.os_c = p.os_c;
    ep.alpha = p.alpha; ep.act = p.act; ep.aux_mode = p.aux_mode; ep.vec_ok = p.vec_ok; ep.N = p.N;
    ep.out_half = p.out_half;
    const int e_bn = p.bn, e_bn_p2 = p.bn_p2, e_N = p.N;
    int lt = 0;
    uint32_t stg_it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & (n_acc - 1);
      const uint32_t aph = (lt >> acc_shift) & 1u;
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int z = mt / tiles_per_z;
      const int rem = mt - z * tiles_per_z;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int n0 = nt * p.bn;

      // output position of this thread's accumulator row
      bool valid = r < p.a_rows;
      long long off = 0;
      if (p.linear) {
        int row = tx * p.box_w + r;
        valid = valid && row < p.out_w;
        if (valid && p.rowmap) {
          row = p.rowmap[row];
          valid = row >= 0;
        }
        if (valid) {
          if (p.lin_batch > 0) {
            const int zb = row / p.lin_batch;
            off = zb * p.os_z + static_cast<long long>(row - zb * p.lin_batch) * p.os_x;
          } else {
            off = static_cast<long long>(row) * p.os_x;
          }
        }
      } else {
        const int ry = r / p.box_w;
        const int oy = ty * p.box_h + ry;
        const int ox = tx * p.box_w + (r - ry * p.box_w);
        valid = valid && oy < p.out_h && ox < p.out_w;
        off = z * p.os_z + oy * p.os_y + ox * p.os_x;
      }

      // stage this tile's bias slice (zeros beyond N / without bias); 128 epilogue threads cooperate
      float* sbias = sbias_base + (lt & 1) * p.bn;  // two bn-float slices (sized exactly: 3 stages + staging must fit)
      for (int c = threadIdx.x - 64; c < p.bn; c += 128)
        sbias[c] = (p.bias && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;\n" ::: "memory");

      if (kTrace && warp == 2 && lane == 0) { FGT_GTRACE(2, lt, 0); }
      mbar_wait(accf_bar(buf), aph);
      tc_fence_after();
      if (kTrace && warp == 2 && lane == 0) { FGT_GTRACE(2, lt, 1); }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(buf * e_bn_p2);
      if (p.tma_out) {
        // 64 output channels at a time: each thread writes its row into two 128-row x 128-byte panels laid out
        // with the 128B swizzle (16-byte chunk j of row r at j ^ (r & 7): conflict-free), then one thread hands
        // the panels to the TMA engine, which writes full 128-byte lines and clips the M / N tails.
        const bool issuer = threadIdx.x == 64;
        const int x0 = tx * p.box_w, y0 = p.linear ? 0 : ty * p.box_h, zz = p.linear ? 0 : z;
        const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        const uint32_t aux_bar = bar_base + 8u * 30;
        for (int c0 = 0; c0 < e_bn && n0 + c0 < e_N; c0 += 64, ++stg_it) {
          // ring of staging slots: this chunk's slot is free once all but the (slots - 1) most recent store groups
          // have finished reading shared memory
          const uint32_t stg_base = stg_ring + static_cast<uint32_t>((stg_it & (p.stg_slots - 1)) * p.stg_slot_bytes);
          if (issuer) {
            if (p.stg_slots == 1) bulk_wait_read<0>();
            else if (p.stg_slots == 2) bulk_wait_read<1>();
            else bulk_wait_read<3>();
            if (p.aux_tma) {  // fetch this chunk's aux tile (two 32-column fp32 panels) into the slot it will be stored from
              const bool two = n0 + c0 + 32 < e_N;
              mbar_expect_tx(aux_bar, static_cast<uint32_t>(p.a_rows) * 128u * (two ? 2u : 1u));
              tma_load_4d(stg_base, &p.aux_map, aux_bar, n0 + c0, x0, y0, zz);
              if (two) tma_load_4d(stg_base + kAPlaneBytes, &p.aux_map, aux_bar, n0 + c0 + 32, x0, y0, zz);
            }
          }
          if (p.aux_tma) {
            mbar_wait(aux_bar, stg_it & 1u);  // implies the slot was free: the loads were issued after the read-wait
          } else {
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t raw[32];
            tmem_ld32(t_row + c0 + hf * 32, raw);
            tmem_ld_wait();
            float v[32];
            epi_math<kExt>(ep, raw, sbias + c0 + hf * 32, off, n0 + c0 + hf * 32, valid, v,
                           p.aux_tma ? stg_base + hf * kAPlaneBytes + row_off : 0u, sw);
            if (ep.out_hi && ep.out_half) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t a = stg_base + row_off + (((static_cast<uint32_t>(hf * 4 + q)) ^ sw) << 4);
                st_shared_v4(a, pack_f16x2(v[8 * q], v[8 * q + 1]), pack_f16x2(v[8 * q + 2], v[8 * q + 3]),
                             pack_f16x2(v[8 * q + 4], v[8 * q + 5]), pack_f16x2(v[8 * q + 6], v[8 * q + 7]));
              }
            } else if (ep.out_hi) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) split_bf16x2(v[8 * q + 2 * t], v[8 * q + 2 * t + 1], hw[t], lw[t]);
                const uint32_t a = stg_base + row_off + (((static_cast<uint32_t>(hf * 4 + q)) ^ sw) << 4);
                st_shared_v4(a, hw[0], hw[1], hw[2], hw[3]);
                st_shared_v4(a + kAPlaneBytes, lw[0], lw[1], lw[2], lw[3]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const uint32_t a = stg_base + hf * kAPlaneBytes + row_off + ((static_cast<uint32_t>(q) ^ sw) << 4);
                st_shared_v4(a, __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                             __float_as_uint(v[4 * q + 3]));
              }
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
          if (issuer) {
            if (ep.out_hi) {
              tma_store_5d(&p.out_map, stg_base, n0 + c0, x0, y0, zz, 0);
              if (!ep.out_half) tma_store_5d(&p.out_map, stg_base + kAPlaneBytes, n0 + c0, x0, y0, zz, 1);
            } else {
              tma_store_4d(&p.out_map, stg_base, n0 + c0, x0, y0, zz);
              if (n0 + c0 + 32 < e_N) tma_store_4d(&p.out_map, stg_base + kAPlaneBytes, n0 + c0 + 32, x0, y0, zz);
            }
            bulk_commit();
          }
        }
      } else if ((e_bn & 31) == 0) {
        for (int c0 = 0; c0 < e_bn; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(t_row + c0, raw);
          tmem_ld_wait();
          if (valid && n0 + c0 < e_N) epi_chunk<32, kExt>(ep, raw, sbias + c0, off, n0 + c0);
        }
      } else {
        for (int c0 = 0; c0 < e_bn; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(t_row + c0, raw);
          tmem_ld_wait();
          if (valid && n0 + c0 < e_N) epi_chunk<16, kExt>(ep, raw, sbias + c0, off, n0 + c0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(buf));
      if (kTrace && warp == 2 && lane == 0) { FGT_GTRACE(2, lt, 2); }
    }
  }
#undef FGT_GTRACE

  if (p.tma_out && threadIdx.x == 64) bulk_wait0();  // all tile stores complete before the CTA's smem goes away
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(n_acc * p.bn_p2));
  }
}

static int next_pow2_ge32(int v) {
  int r = 32;
  while (r < v) r <<= 1;
  return r;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

static bool g_gemm_trace_on = false;  // debugging aid, see fgt_debug_gemm_trace()
static bool g_gemm_direct_epilogue = false;  // debugging aid: force the register-store epilogue (A/B comparisons)
static bool g_gemm_direct_aux = false;       // debugging aid: aux operand by per-row global loads even on the TMA path

int gemm_tc_launch(const FgtGemmDesc& d, cudaStream_t stream) {
  FGT_REQUIRE(d.num_segs >= 1 && d.num_segs <= 2, FGT_ERR_ARG, "gemm_tc: num_segs=%d", d.num_segs);
  FGT_REQUIRE(d.kx >= 1 && d.ky >= 1 && d.kz >= 1 && d.kx * d.ky * d.kz <= kMaxTaps, FGT_ERR_ARG,
              "gemm_tc: taps %dx%dx%d unsupported", d.kx, d.ky, d.kz);
  FGT_REQUIRE(d.stride >= 1 && d.stride <= 3 && d.dil >= 1, FGT_ERR_ARG, "gemm_tc: stride=%d dil=%d",
              d.stride, d.dil);
  FGT_REQUIRE(d.bn >= 16 && d.bn <= 256 && d.bn % 16 == 0, FGT_ERR_ARG, "gemm_tc: bn=%d", d.bn);
  FGT_REQUIRE(d.groups >= 1 && d.N % d.groups == 0, FGT_ERR_ARG, "gemm_tc: N=%d groups=%d", d.N, d.groups);
  const int cpg_out = d.N / d.groups;
  FGT_REQUIRE(d.groups == 1 || cpg_out % d.bn == 0, FGT_ERR_ARG,
              "gemm_tc: bn=%d must divide N/groups=%d", d.bn, cpg_out);
  const int a_rows = d.box_w * d.box_h;
  FGT_REQUIRE(d.box_w >= 1 && d.box_h >= 1 && a_rows <= 128 && a_rows % 8 == 0 && d.box_w <= 256,
              FGT_ERR_ARG, "gemm_tc: box %dx%d", d.box_w, d.box_h);
  FGT_REQUIRE(d.out_f32 || d.out_hi, FGT_ERR_ARG, "gemm_tc: no output");
  FGT_REQUIRE(d.aux_mode == FGT_AUX_NONE || d.aux, FGT_ERR_ARG, "gemm_tc: aux_mode without aux");
  FGT_REQUIRE(d.aux_mode != FGT_AUX_GRU || d.aux2, FGT_ERR_ARG, "gemm_tc: FGT_AUX_GRU needs aux2");
  FGT_REQUIRE((reinterpret_cast<uintptr_t>(d.w_hi) & 15) == 0 && d.k_pad % 64 == 0, FGT_ERR_ARG,
              "gemm_tc: weights misaligned / k_pad=%d", d.k_pad);

  GemmParams p;
  memset(&p, 0, sizeof(p));
  FGT_REQUIRE(d.terms == 0 || d.terms == 1 || d.terms == 3, FGT_ERR_ARG, "gemm_tc: terms=%d (1 or 3)", d.terms);
  const bool one_term = d.terms == 1;  // operands are single-plane fp16 tensors
  const int s = d.stride;
  const int phases = s * s;
  FGT_REQUIRE(d.num_segs * phases <= kMaxAMaps, FGT_ERR_ARG, "gemm_tc: too many A maps");

  // ---- A tensor maps: one per (segment, stride phase); dims (C, X', Y', Z, plane)
  int chunks_total = 0;
  for (int sgi = 0; sgi < d.num_segs; ++sgi) {
    const FgtASeg& sg = d.seg[sgi];
    FGT_REQUIRE((reinterpret_cast<uintptr_t>(sg.hi) & 15) == 0, FGT_ERR_ARG, "gemm_tc: A base misaligned");
    FGT_REQUIRE(sg.sx % 8 == 0 && (sg.DY == 1 || sg.sy % 8 == 0) && (sg.DZ == 1 || sg.sz % 8 == 0) &&
                    (one_term || sg.plane % 8 == 0),
                FGT_ERR_ARG, "gemm_tc: A strides must be multiples of 8 elements (sx=%lld sy=%lld sz=%lld)",
                sg.sx, sg.sy, sg.sz);
    FGT_REQUIRE(sg.c_count >= 1, FGT_ERR_ARG, "gemm_tc: c_count");
    for (int py = 0; py < s; ++py)
      for (int px = 0; px < s; ++px) {
        const uint64_t dx = static_cast<uint64_t>((sg.DX - px + s - 1) / s);
        const uint64_t dy = static_cast<uint64_t>((sg.DY - py + s - 1) / s);
        uint64_t dims[5] = {static_cast<uint64_t>(sg.C), dx > 0 ? dx : 1, dy > 0 ? dy : 1,
                            static_cast<uint64_t>(sg.DZ), one_term ? 1u : 2u};
        uint64_t strides[4] = {static_cast<uint64_t>(sg.sx) * s * 2, static_cast<uint64_t>(sg.sy) * s * 2,
                               static_cast<uint64_t>(sg.sz) * 2, static_cast<uint64_t>(sg.plane) * 2};
        if (sg.DY == 1) strides[1] = strides[0] * dims[1];
        if (sg.DZ == 1) strides[2] = strides[1] * dims[2];
        if (one_term) strides[3] = strides[2] * dims[3];  // single fp16 plane: the plane dimension has extent 1
        uint32_t box[5] = {kBK, static_cast<uint32_t>(d.box_w), static_cast<uint32_t>(d.box_h), 1, 1};
        const __nv_bfloat16* base =
            reinterpret_cast<const __nv_bfloat16*>(sg.hi) + py * sg.sy + px * sg.sx;
        int rc = encode_map_bf16(&p.a_maps[sgi * phases + py * s + px], base, 5, dims, strides, box);
        if (rc) return rc;
      }
    p.segs[sgi].map_base = sgi * phases;
    p.segs[sgi].c_base = sg.c_base;
    p.segs[sgi].c_per_group = sg.c_per_group;
    p.segs[sgi].chunks = (sg.c_count + kBK - 1) / kBK;
    chunks_total += p.segs[sgi].chunks;
  }
  // ---- taps (z-major, then y, then x: matches packing.py)
  int nt = 0;
  for (int kz = 0; kz < d.kz; ++kz)
    for (int ky = 0; ky < d.ky; ++ky)
      for (int kx = 0; kx < d.kx; ++kx) {
        const int ox = kx * d.dil - d.pad_x, oy = ky * d.dil - d.pad_y;
        const int qx = floordiv(ox, s), qy = floordiv(oy, s);
        TapDesc& t = p.taps[nt++];
        t.phase = static_cast<int8_t>((oy - qy * s) * s + (ox - qx * s));
        FGT_REQUIRE(qx >= -128 && qx <= 127 && qy >= -128 && qy <= 127, FGT_ERR_ARG, "gemm_tc: tap offset");
        t.dx = static_cast<int8_t>(qx);
        t.dy = static_cast<int8_t>(qy);
        t.dz = static_cast<int8_t>(kz - d.pad_z);
      }
  p.num_taps = nt;
  p.num_segs = d.num_segs;
  p.num_maps = d.num_segs * phases;
  p.k_iters = nt * chunks_total;
  FGT_REQUIRE(p.k_iters * kBK == d.k_pad, FGT_ERR_ARG, "gemm_tc: k_pad=%d but schedule needs %d", d.k_pad,
              p.k_iters * kBK);
  // ---- B map: (K, N, plane)
  {
    uint64_t dims[3] = {static_cast<uint64_t>(d.k_pad), static_cast<uint64_t>(d.N), one_term ? 1u : 2u};
    uint64_t strides[2] = {static_cast<uint64_t>(d.k_pad) * 2, static_cast<uint64_t>(d.w_plane) * 2};
    if (one_term) strides[1] = strides[0] * dims[1];
    uint32_t box[3] = {kBK, static_cast<uint32_t>(d.bn), 1};
    int rc = encode_map_bf16(&p.b_map, d.w_hi, 3, dims, strides, box);
    if (rc) return rc;
  }
  p.linear = (d.out_h == 1 && d.out_z == 1 && d.box_h == 1) ? 1 : 0;
  FGT_REQUIRE(p.linear || (!d.rowmap && d.lin_batch == 0), FGT_ERR_ARG, "gemm_tc: rowmap/lin_batch need linear mode");
  p.tiles_x = (d.out_w + d.box_w - 1) / d.box_w;
  p.tiles_y = (d.out_h + d.box_h - 1) / d.box_h;
  p.tiles_z = d.out_z;
  p.n_tiles = (d.N + d.bn - 1) / d.bn;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_z * p.n_tiles;
  p.bn = d.bn;
  p.bn_p2 = next_pow2_ge32(d.bn);
  p.n_acc = 512 / p.bn_p2 < 8 ? 512 / p.bn_p2 : 8;
  p.box_w = d.box_w;
  p.box_h = d.box_h;
  p.a_rows = a_rows;
  p.N = d.N;
  p.cout_per_group = cpg_out;
  p.out_w = d.out_w;
  p.out_h = d.out_h;
  p.lin_batch = d.lin_batch;
  p.os_z = d.os_z; p.os_y = d.os_y; p.os_x = d.os_x; p.os_c = d.os_c;
  p.rowmap = d.rowmap;
  p.out_f32 = d.out_f32;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(d.out_hi);
  p.out_plane = d.out_plane;
  p.aux = d.aux;
  p.aux2 = d.aux2;
  p.aux_mode = d.aux_mode;
  p.bias = d.bias;
  p.act = d.act;
  p.alpha = d.alpha;
  // 16-byte vector path: unit channel stride and every address component a multiple of 8 elements
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_ok = (d.os_c == 1 && d.os_x % 8 == 0 && d.os_y % 8 == 0 && d.os_z % 8 == 0 && (d.out_half || d.out_plane % 8 == 0) &&
              al(d.out_f32) && al(d.out_hi) && al(d.aux) && al(d.aux2))
                 ? 1
                 : 0;

  // ---- numerics: 3-term split-bf16 (default) or single-plane fp16 operands (1 MMA per K step)
  p.planes = one_term ? 1 : 2;
  p.out_half = d.out_half ? 1 : 0;
  FGT_REQUIRE(!p.out_half || d.out_hi, FGT_ERR_ARG, "gemm_tc: out_half without a 16-bit output");
  // ---- epilogue through shared memory + TMA tile stores: one output, channels contiguous, plain row order
  const bool one_out = (d.out_f32 != nullptr) != (d.out_hi != nullptr);
  p.tma_out = (one_out && p.vec_ok && !d.rowmap && d.lin_batch == 0 && d.bn % 64 == 0 && d.N % 4 == 0 &&
               d.aux_mode != FGT_AUX_GRU_ZR && !g_gemm_direct_epilogue)
                  ? 1
                  : 0;
  if (p.tma_out) {
    const bool lin = p.linear != 0;
    uint64_t dims[5] = {static_cast<uint64_t>(d.N), static_cast<uint64_t>(d.out_w), static_cast<uint64_t>(d.out_h),
                        static_cast<uint64_t>(d.out_z), p.out_half ? 1u : 2u};
    const uint64_t es = d.out_hi ? 2 : 4;
    uint64_t str[4] = {static_cast<uint64_t>(d.os_x) * es, static_cast<uint64_t>(d.os_y) * es,
                       static_cast<uint64_t>(d.os_z) * es, static_cast<uint64_t>(d.out_plane) * es};
    if (lin || d.out_h == 1) str[1] = str[0] * dims[1];
    if (lin || d.out_z == 1) str[2] = str[1] * dims[2];
    if (p.out_half) str[3] = str[2] * dims[3];
    uint32_t box[5] = {d.out_hi ? 64u : 32u, static_cast<uint32_t>(d.box_w), static_cast<uint32_t>(d.box_h), 1, 1};
    int rc = d.out_hi ? encode_map_bf16(&p.out_map, d.out_hi, 5, dims, str, box)
                      : encode_map_f32(&p.out_map, d.out_f32, 4, dims, str, box);
    if (rc) return rc;
    // fp32 output + an aux operand that is a plain fp32 tensor addressed like the output (residual add, gate multiply):
    // fetch the aux tile by TMA as well — per-row 16-byte loads of a row-per-thread epilogue touch 32 lines per request
    p.aux_tma = (!d.out_hi && d.aux && !g_gemm_direct_aux &&
                 (d.aux_mode == FGT_AUX_ADD || d.aux_mode == FGT_AUX_MUL || d.aux_mode == FGT_AUX_ADD_PRE ||
                  d.aux_mode == FGT_AUX_ADD_RELU))
                    ? 1
                    : 0;
    if (p.aux_tma) {
      rc = encode_map_f32(&p.aux_map, d.aux, 4, dims, str, box);
      if (rc) return rc;
    }
  }

  const uint32_t stage_bytes = static_cast<uint32_t>(p.planes) * (kAPlaneBytes + static_cast<uint32_t>(d.bn) * 128u);
  // staging ring of the TMA epilogue: a slot holds one 64-channel chunk (two 16 KB panels; one for an fp16 output).
  // More slots let the stores of one chunk drain while the next is written — what short-K layers need; long-K
  // layers keep the shared memory for pipeline stages instead.
  p.stg_slot_bytes = static_cast<int>(p.out_half ? kAPlaneBytes : 2 * kAPlaneBytes);
  // alignment slack + barriers + two bias slices of bn floats. Sized exactly: a 128-wide split-bf16 tile with the
  // TMA epilogue needs 3 x 64 KB stages + 32 KB of panels + this block within the 227 KB limit.
  const uint32_t base_fixed = 1024u /*align*/ + 256u /*barriers*/ + 2u * static_cast<uint32_t>(d.bn) * 4u /*bias*/;
  int stages = 0;
  p.stg_slots = 0;
  if (p.tma_out) {
    for (int slots = 4; slots >= 1 && p.stg_slots == 0; slots >>= 1) {
      const uint32_t fixed = base_fixed + static_cast<uint32_t>(slots * p.stg_slot_bytes);
      const int st = static_cast<int>((227u * 1024u - fixed) / stage_bytes);
      const int want = slots == 1 ? 2 : 4;  // extra slots only while at least 4 pipeline stages remain
      if (st >= want) {
        p.stg_slots = slots;
        stages = st;
      }
    }
    if (p.stg_slots == 0) p.tma_out = 0;  // tile too large to also hold the output panels: store from registers
  }
  if (!p.tma_out) stages = static_cast<int>((227u * 1024u - base_fixed) / stage_bytes);
  if (stages > 6) stages = 6;
  FGT_REQUIRE(stages >= 2, FGT_ERR_ARG, "gemm_tc: tile too large for shared memory");
  p.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + base_fixed +
                      static_cast<size_t>(p.tma_out ? p.stg_slots * p.stg_slot_bytes : 0);

  static bool attr_set = false;
  if (!attr_set) {
    FGT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FGT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FGT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const bool ext = d.aux_mode == FGT_AUX_ADD_PRE || d.aux_mode == FGT_AUX_ADD_RELU || d.aux_mode == FGT_AUX_GRU ||
                   d.aux_mode == FGT_AUX_GRU_ZR || d.act == FGT_ACT_LEAKY001;
  if (d.aux_mode == FGT_AUX_GRU_ZR)
    FGT_REQUIRE(d.out_f32 && d.out_hi && p.vec_ok && d.N % 64 == 0 && (d.N / 2) % d.bn == 0 && d.groups == 1, FGT_ERR_ARG,
                "gemm_tc: FGT_AUX_GRU_ZR needs both outputs, aligned unit-stride channels, N=%d a multiple of 64 and "
                "bn=%d dividing N/2", d.N, d.bn);
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  FGT_REQUIRE(grid >= 1, FGT_ERR_ARG, "gemm_tc: empty problem");
  if (ext) launch_k(gemm_tc_kernel<1>, dim3(grid), dim3(192), smem, stream, p);
  else if (g_gemm_trace_on) launch_k(gemm_tc_kernel<0, 1>, dim3(grid), dim3(192), smem, stream, p);
  else launch_k(gemm_tc_kernel<0>, dim3(grid), dim3(192), smem, stream, p);
  FGT_CUDA(cudaGetLastError());
  return FGT_OK;
}

}  // namespace fgt

extern "C" int fgt_gemm_tc(const FgtGemmDesc* desc, fgt_stream_t stream) {
  if (!desc) return fgt::set_err(FGT_ERR_ARG, "fgt_gemm_tc: null desc");
  return fgt::gemm_tc_launch(*desc, reinterpret_cast<cudaStream_t>(stream));
}

// Debugging aid (not part of the product interface, like fgt_debug_flash_trace): while buf != NULL, common-epilogue
// launches use the traced instantiation and CTA `cta` writes its clock64 timeline into buf
// ([3 roles][64 local tiles][4 slots] int64, device memory): producer {tile start, first stage slot free, last load
// issued}, MMA {tile start, accumulator buffer free, first operands landed, last MMA committed}, epilogue {tile start,
// accumulator complete, stores issued}.
extern "C" int fgt_debug_gemm_direct_epilogue(int on) {
  fgt::g_gemm_direct_epilogue = (on & 1) != 0;
  fgt::g_gemm_direct_aux = (on & 2) != 0;
  return FGT_OK;
}

extern "C" int fgt_debug_gemm_trace(long long* buf, int cta) {
  FGT_CUDA(cudaMemcpyToSymbol(fgt::g_gemm_trace, &buf, sizeof(buf)));
  FGT_CUDA(cudaMemcpyToSymbol(fgt::g_gemm_trace_cta, &cta, sizeof(cta)));
  fgt::g_gemm_trace_on = buf != nullptr;
  return FGT_OK;
}
