// Persistent, warp-specialised implicit-GEMM engine for sm_100a.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer     (one elected thread, tcgen05.mma kind::f16, 3-term split-bf16)
//   warps 2..5  : epilogue       (tcgen05.ld TMEM -> regs -> bias/act/aux -> global)
//
// One CTA per SM loops over output tiles (128 output positions x bn channels). The fp32
// accumulator lives in TMEM and is double-buffered so the epilogue of tile i overlaps the
// main loop of tile i+1. Used for every Linear/Conv on the path (see include/fgt_b200.h).
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
  TapDesc taps[kMaxTaps];
  SegDesc segs[2];
  int num_taps, num_segs, num_maps, k_iters;
  int tiles_x, tiles_y, tiles_z, n_tiles, total_tiles;
  int bn, bn_p2, stages;
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
  int act, aux_mode, vec_ok, N;
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
    if (e.out_hi) {
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
        if (e.out_hi) {
          __nv_bfloat16 h, l;
          split_bf16(x, h, l);
          e.out_hi[o] = h;
          e.out_hi[e.out_plane + o] = l;
        }
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
  const uint32_t stage_bytes = 2u * kAPlaneBytes + 2u * b_plane_bytes;
  const uint32_t bar_base = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;
  // barrier layout: full[stages], empty[stages], acc_full[2], acc_empty[2], tmem slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  // per-tile bias slice staged in smem (2 buffers x 256 floats), after the 256-byte barrier block
  float* sbias_base = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 256);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 4);
    }
    fence_mbar_init();
    for (int i = 0; i < p.num_maps; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2u * p.bn_p2);
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
      const uint32_t tx_bytes = 2u * static_cast<uint32_t>(p.a_rows) * 128u + 2u * b_plane_bytes;
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
              const uint32_t sb = sa + 2u * kAPlaneBytes;
              const uint32_t fb = full_bar(stage);
              mbar_expect_tx(fb, tx_bytes);
              const int cc = c0 + ch * kBK;
              tma_load_5d(sa, amap, fb, cc, x0 + tap.dx, y0 + tap.dy, z + tap.dz, 0);
              tma_load_5d(sa + kAPlaneBytes, amap, fb, cc, x0 + tap.dx, y0 + tap.dy, z + tap.dz, 1);
              tma_load_3d(sb, &p.b_map, fb, kidx * kBK, n0, 0);
              tma_load_3d(sb + b_plane_bytes, &p.b_map, fb, kidx * kBK, n0, 1);
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
      const uint32_t idesc = umma_idesc_bf16(128, p.bn);
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const uint32_t aph = (lt >> 1) & 1u;
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
          const uint32_t sb = sa + 2u * kAPlaneBytes;
          const uint64_t a_hi = umma_desc_sw128(sa);
          const uint64_t a_lo = umma_desc_sw128(sa + kAPlaneBytes);
          const uint64_t b_hi = umma_desc_sw128(sb);
          const uint64_t b_lo = umma_desc_sw128(sb + b_plane_bytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ko = static_cast<uint64_t>(k * 2);  // 32 bytes >> 4
            umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, (it | k) != 0);
            umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
            umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
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
    const int e_bn = p.bn, e_bn_p2 = p.bn_p2, e_N = p.N;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t aph = (lt >> 1) & 1u;
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
      float* sbias = sbias_base + buf * 256;
      for (int c = threadIdx.x - 64; c < p.bn; c += 128)
        sbias[c] = (p.bias && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;\n" ::: "memory");

      if (kTrace && warp == 2 && lane == 0) { FGT_GTRACE(2, lt, 0); }
      mbar_wait(accf_bar(buf), aph);
      tc_fence_after();
      if (kTrace && warp == 2 && lane == 0) { FGT_GTRACE(2, lt, 1); }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(buf * e_bn_p2);
      if ((e_bn & 31) == 0) {
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2u * p.bn_p2);
  }
}

static int next_pow2_ge32(int v) {
  int r = 32;
  while (r < v) r <<= 1;
  return r;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

static bool g_gemm_trace_on = false;  // debugging aid, see fgt_debug_gemm_trace()

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
  const int s = d.stride;
  const int phases = s * s;
  FGT_REQUIRE(d.num_segs * phases <= kMaxAMaps, FGT_ERR_ARG, "gemm_tc: too many A maps");

  // ---- A tensor maps: one per (segment, stride phase); dims (C, X', Y', Z, plane)
  int chunks_total = 0;
  for (int sgi = 0; sgi < d.num_segs; ++sgi) {
    const FgtASeg& sg = d.seg[sgi];
    FGT_REQUIRE((reinterpret_cast<uintptr_t>(sg.hi) & 15) == 0, FGT_ERR_ARG, "gemm_tc: A base misaligned");
    FGT_REQUIRE(sg.sx % 8 == 0 && (sg.DY == 1 || sg.sy % 8 == 0) && (sg.DZ == 1 || sg.sz % 8 == 0) &&
                    sg.plane % 8 == 0,
                FGT_ERR_ARG, "gemm_tc: A strides must be multiples of 8 elements (sx=%lld sy=%lld sz=%lld)",
                sg.sx, sg.sy, sg.sz);
    FGT_REQUIRE(sg.c_count >= 1, FGT_ERR_ARG, "gemm_tc: c_count");
    for (int py = 0; py < s; ++py)
      for (int px = 0; px < s; ++px) {
        const uint64_t dx = static_cast<uint64_t>((sg.DX - px + s - 1) / s);
        const uint64_t dy = static_cast<uint64_t>((sg.DY - py + s - 1) / s);
        uint64_t dims[5] = {static_cast<uint64_t>(sg.C), dx > 0 ? dx : 1, dy > 0 ? dy : 1,
                            static_cast<uint64_t>(sg.DZ), 2};
        uint64_t strides[4] = {static_cast<uint64_t>(sg.sx) * s * 2, static_cast<uint64_t>(sg.sy) * s * 2,
                               static_cast<uint64_t>(sg.sz) * 2, static_cast<uint64_t>(sg.plane) * 2};
        if (sg.DY == 1) strides[1] = strides[0] * dims[1];
        if (sg.DZ == 1) strides[2] = strides[1] * dims[2];
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
    uint64_t dims[3] = {static_cast<uint64_t>(d.k_pad), static_cast<uint64_t>(d.N), 2};
    uint64_t strides[2] = {static_cast<uint64_t>(d.k_pad) * 2, static_cast<uint64_t>(d.w_plane) * 2};
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
  p.vec_ok = (d.os_c == 1 && d.os_x % 8 == 0 && d.os_y % 8 == 0 && d.os_z % 8 == 0 && d.out_plane % 8 == 0 &&
              al(d.out_f32) && al(d.out_hi) && al(d.aux) && al(d.aux2))
                 ? 1
                 : 0;

  const uint32_t stage_bytes = 2u * kAPlaneBytes + 2u * static_cast<uint32_t>(d.bn) * 128u;
  int stages = static_cast<int>((227u * 1024u - 1024u - 4096u) / stage_bytes);
  if (stages > 6) stages = 6;
  FGT_REQUIRE(stages >= 2, FGT_ERR_ARG, "gemm_tc: tile too large for shared memory");
  p.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*bias*/;

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
extern "C" int fgt_debug_gemm_trace(long long* buf, int cta) {
  FGT_CUDA(cudaMemcpyToSymbol(fgt::g_gemm_trace, &buf, sizeof(buf)));
  FGT_CUDA(cudaMemcpyToSymbol(fgt::g_gemm_trace_cta, &cta, sizeof(cta)));
  fgt::g_gemm_trace_on = buf != nullptr;
  return FGT_OK;
}
