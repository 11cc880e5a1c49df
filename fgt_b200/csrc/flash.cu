// Fused softmax(QK^T * scale) V on tcgen05 for FGT's two attentions (never materialises scores).
//
// One CTA = 128 query rows x one head (head dim 128). Keys are consumed in tiles of 64:
//   warp 0     : TMA producer (Q once; K and V^T tiles through two 2-deep rings)
//   warp 1     : MMA issuer   (S_j = Q K_j^T into TMEM; O_j = P_j V_j into TMEM; 3-term split-bf16)
//   warps 2..5 : softmax      (thread = query row: online max/sum in fp32, P written to smem as
//                              split-bf16 in the 128B-swizzled K-major layout)
// S is double-buffered in TMEM (S_{j+1} runs while the softmax warps work on tile j); the O
// accumulator stays in TMEM across key tiles and is rescaled in place only when the running max
// outgrows the reference max by more than 2^8 (lazy rescaling), so it is read back exactly once.
//
// mode DENSE    (TMHSA, attention_base.py:93-99): every key tile of the (batch=zone) sequence.
// mode WINDOWED (SWMHSA, attention_flow.py:130-164): a 128-row query tile is two 64-token windows;
//   its key list is [global tiles..., own window tile 2i, own window tile 2i+1] with a block-diagonal
//   mask on the two local tiles. Global (pooled) tokens live after the local rows of each frame.
#include "common.h"
#include "ptx.cuh"

namespace fgt {

constexpr int kQBlk = 128 * 128;      // one (plane, k-chunk) block of Q: 128 rows x 128 B
constexpr int kKBlk = 64 * 128;       // one (plane, k-chunk) block of K: 64 keys x 128 B
constexpr int kVBlk = 128 * 128;      // one plane of V^T: 128 dims x 64 keys (128 B)
constexpr int kPBlk = 128 * 128;      // one plane of P: 128 rows x 64 keys
constexpr int kKStage = 4 * kKBlk;    // 32 KB
constexpr int kVStage = 2 * kVBlk;    // 32 KB
constexpr int kSmemQ = 0;
constexpr int kSmemK = 4 * kQBlk;                 // 64 KB
constexpr int kSmemV = kSmemK + 2 * kKStage;      // +64 KB
constexpr int kSmemP = kSmemV + 2 * kVStage;      // +64 KB
constexpr int kSmemBar = kSmemP + 2 * kPBlk;      // +32 KB = 224 KB
constexpr int kFlashSmem = kSmemBar + 256 + 1024;

struct FlashParams {
  CUtensorMap q_map, k_map, v_map;
  int Lq, Lk, heads;
  int mode;  // 0 dense, 1 windowed
  int n_glob_tiles, glob_start, glob_count;
  float scale_log2;
  __nv_bfloat16* out_hi;
  long long out_plane, out_batch_stride;
  int out_ld;
  long long* trace;  // optional per-tile clock64 trace of CTA (0,0,0): [role 0..2][tile][8]
};

__global__ void __launch_bounds__(192, 1) flash_kernel(const __grid_constant__ FlashParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;

  const uint32_t bar = sbase + kSmemBar;
  const uint32_t q_full = bar;
  auto k_full = [&](int s) { return bar + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar + 8u * (3 + s); };
  auto v_full = [&](int s) { return bar + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bar + 8u * (7 + s); };
  auto s_full = [&](int s) { return bar + 8u * (9 + s); };
  auto s_empty = [&](int s) { return bar + 8u * (11 + s); };
  const uint32_t p_full = bar + 8u * 13;   // softmax wrote P_j (and finished any O rescale)
  const uint32_t p_empty = bar + 8u * 14;  // PV_j retired: P buffer free, O accumulator up to date
  const uint32_t tmem_slot = bar + 8u * 15;

  const int n_tiles = (p.mode == 0) ? (p.Lk + 63) / 64 : p.n_glob_tiles + 2;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define FGT_TRACE(role, tile, slot) \
  if (tr) p.trace[((role) * 64 + ((tile) & 63)) * 8 + (slot)] = clock64()

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(s_empty(s), 128);
    }
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    tma_prefetch_desc(&p.v_map);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));
  // TMEM columns: S buffers at [0,64) and [64,128); the O accumulator at [128,256)

  auto key_row0 = [&](int j) -> int {
    if (p.mode == 0) return j * 64;
    if (j < p.n_glob_tiles) return p.glob_start + j * 64;
    return (2 * qt + (j - p.n_glob_tiles)) * 64;
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 4 * kQBlk);
      for (int pl = 0; pl < 2; ++pl)
        for (int kc = 0; kc < 2; ++kc)
          tma_load_4d(sbase + kSmemQ + (pl * 2 + kc) * kQBlk, &p.q_map, q_full, head * 128 + kc * 64, qt * 128,
                      batch, pl);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        const int r0 = key_row0(j);
        mbar_wait(k_empty(s), ph ^ 1u);
        FGT_TRACE(0, j, 0);
        mbar_expect_tx(k_full(s), kKStage);
        for (int pl = 0; pl < 2; ++pl)
          for (int kc = 0; kc < 2; ++kc)
            tma_load_4d(sbase + kSmemK + s * kKStage + (pl * 2 + kc) * kKBlk, &p.k_map, k_full(s),
                        head * 128 + kc * 64, r0, batch, pl);
        mbar_wait(v_empty(s), ph ^ 1u);
        FGT_TRACE(0, j, 1);
        mbar_expect_tx(v_full(s), kVStage);
        for (int pl = 0; pl < 2; ++pl)
          tma_load_4d(sbase + kSmemV + s * kVStage + pl * kVBlk, &p.v_map, v_full(s), r0, head * 128, batch, pl);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 64);
      const uint32_t idesc_o = umma_idesc_bf16(128, 128);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        mbar_wait(k_full(s), ph);
        FGT_TRACE(1, j, 0);
        mbar_wait(s_empty(s), ph ^ 1u);
        FGT_TRACE(1, j, 1);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(s * 64);
        uint32_t acc = 0;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint64_t a_hi = umma_desc_sw128(sbase + kSmemQ + (0 * 2 + kc) * kQBlk);
          const uint64_t a_lo = umma_desc_sw128(sbase + kSmemQ + (1 * 2 + kc) * kQBlk);
          const uint64_t b_hi = umma_desc_sw128(sbase + kSmemK + s * kKStage + (0 * 2 + kc) * kKBlk);
          const uint64_t b_lo = umma_desc_sw128(sbase + kSmemK + s * kKStage + (1 * 2 + kc) * kKBlk);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = static_cast<uint64_t>(k * 2);
            umma_bf16(d, a_lo + ko, b_hi + ko, idesc_s, acc);
            acc = 1;
            umma_bf16(d, a_hi + ko, b_lo + ko, idesc_s, 1u);
            umma_bf16(d, a_hi + ko, b_hi + ko, idesc_s, 1u);
          }
        }
        umma_commit(k_empty(s));
        umma_commit(s_full(s));
        FGT_TRACE(1, j, 2);
      };
      auto issue_pv = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        mbar_wait(v_full(s), ph);
        FGT_TRACE(1, j, 3);
        mbar_wait(p_full, j & 1u);
        FGT_TRACE(1, j, 4);
        tc_fence_after();
        const uint32_t d = tmem_base + 128u;
        const uint64_t a_hi = umma_desc_sw128(sbase + kSmemP);
        const uint64_t a_lo = umma_desc_sw128(sbase + kSmemP + kPBlk);
        const uint64_t b_hi = umma_desc_sw128(sbase + kSmemV + s * kVStage);
        const uint64_t b_lo = umma_desc_sw128(sbase + kSmemV + s * kVStage + kVBlk);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ko = static_cast<uint64_t>(k * 2);
          umma_bf16(d, a_lo + ko, b_hi + ko, idesc_o, (j | k) != 0);  // O accumulates across key tiles
          umma_bf16(d, a_hi + ko, b_lo + ko, idesc_o, 1u);
          umma_bf16(d, a_hi + ko, b_hi + ko, idesc_o, 1u);
        }
        umma_commit(v_empty(s));
        umma_commit(p_empty);
        FGT_TRACE(1, j, 5);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        issue_pv(j);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax / output warps
    // The O accumulator stays in TMEM. Exponentials are taken against a reference max m_ref that is
    // only raised (and O / l rescaled in place) when the running max outgrows it by more than 2^8 —
    // softmax is shift-invariant, so the result is exact; P values stay <= 256.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float m_ref = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const bool trs = tr && threadIdx.x == 64;
      if (trs) p.trace[(2 * 64 + (j & 63)) * 8 + 0] = clock64();
      mbar_wait(s_full(s), (j >> 1) & 1u);
      if (trs) p.trace[(2 * 64 + (j & 63)) * 8 + 1] = clock64();
      tc_fence_after();
      float sv[64];
      {
        uint32_t raw[32];
        tmem_ld32(lane_base + static_cast<uint32_t>(s * 64), raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[i] = __uint_as_float(raw[i]);
        tmem_ld32(lane_base + static_cast<uint32_t>(s * 64 + 32), raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[32 + i] = __uint_as_float(raw[i]);
      }
      tc_fence_before();
      mbar_arrive(s_empty(s));
      if (trs) p.trace[(2 * 64 + (j & 63)) * 8 + 2] = clock64();

      int n_valid = 64;  // columns [0, n_valid) are live for this row
      if (p.mode == 0) {
        n_valid = min(64, p.Lk - j * 64);
      } else if (j < p.n_glob_tiles) {
        n_valid = min(64, p.glob_count - j * 64);
      } else {
        n_valid = ((r >> 6) == (j - p.n_glob_tiles)) ? 64 : 0;
      }
      if (n_valid != 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) sv[i] = (i < n_valid) ? sv[i] : -INFINITY;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        mx0 = fmaxf(mx0, sv[i]);
        mx1 = fmaxf(mx1, sv[i + 1]);
      }
      const float m_tile = fmaxf(mx0, mx1) * p.scale_log2;  // scale > 0
      const bool grow = m_tile > m_ref + 8.f;               // also true for the first finite maximum
      const float alpha = grow ? ex2_approx(m_ref - m_tile) : 1.f;  // m_ref = -inf -> 0
      if (grow) m_ref = m_tile;
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        sv[i] = ex2_approx(fmaf(sv[i], p.scale_log2, neg_m));
        sv[i + 1] = ex2_approx(fmaf(sv[i + 1], p.scale_log2, neg_m));
        sv[i + 2] = ex2_approx(fmaf(sv[i + 2], p.scale_log2, neg_m));
        sv[i + 3] = ex2_approx(fmaf(sv[i + 3], p.scale_log2, neg_m));
        ps0 += sv[i]; ps1 += sv[i + 1]; ps2 += sv[i + 2]; ps3 += sv[i + 3];
      }
      l_run = l_run * alpha + ((ps0 + ps1) + (ps2 + ps3));

      // PV_{j-1} must have retired before P is overwritten or O is rescaled
      if (trs) p.trace[(2 * 64 + (j & 63)) * 8 + 3] = clock64();
      mbar_wait(p_empty, (j & 1u) ^ 1u);
      if (trs) p.trace[(2 * 64 + (j & 63)) * 8 + 4] = clock64();
      if (j > 0 && __any_sync(0xffffffffu, grow)) {  // tcgen05.ld/st are warp-collective
        tc_fence_after();
        const float a = grow ? alpha : 1.f;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(lane_base + 128u + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * a);
          tmem_st32(lane_base + 128u + c0, raw);
        }
        tmem_st_wait();
      }
      {
        const uint32_t row_hi = sbase + kSmemP + static_cast<uint32_t>(r) * 128u;
        const uint32_t row_lo = row_hi + kPBlk;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_bf16x2(sv[c * 8 + 2 * q], sv[c * 8 + 2 * q + 1], hw[q], lw[q]);
          const uint32_t off = static_cast<uint32_t>((c ^ (r & 7)) * 16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(row_hi + off), "r"(hw[0]), "r"(hw[1]),
                       "r"(hw[2]), "r"(hw[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(row_lo + off), "r"(lw[0]), "r"(lw[1]),
                       "r"(lw[2]), "r"(lw[3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (trs) p.trace[(2 * 64 + (j & 63)) * 8 + 5] = clock64();
    }
    // last PV retired -> read the accumulator once, normalise, store split-bf16
    mbar_wait(p_empty, (n_tiles - 1) & 1u);
    tc_fence_after();
    const int qrow = qt * 128 + r;
    const float inv = 1.f / l_run;
    __nv_bfloat16* oh = p.out_hi + batch * p.out_batch_stride + static_cast<long long>(qrow) * p.out_ld + head * 128;
    __nv_bfloat16* ol = oh + p.out_plane;
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t raw[32];
      tmem_ld32(lane_base + 128u + c0, raw);
      tmem_ld_wait();
      if (qrow < p.Lq) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            split_bf16x2(__uint_as_float(raw[c * 8 + 2 * q]) * inv, __uint_as_float(raw[c * 8 + 2 * q + 1]) * inv, hw[q],
                         lw[q]);
          reinterpret_cast<uint4*>(oh + c0)[c] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          reinterpret_cast<uint4*>(ol + c0)[c] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static long long* g_flash_trace = nullptr;  // debugging aid, see fgt_debug_flash_trace()

int attention_launch(const FgtAttnDesc& d, cudaStream_t stream) {
  FGT_REQUIRE(d.head_dim == 128, FGT_ERR_ARG, "attention: head_dim=%d (only 128 supported)", d.head_dim);
  FGT_REQUIRE(d.batches >= 1 && d.heads >= 1 && d.Lq >= 1 && d.Lk >= 1, FGT_ERR_ARG, "attention: empty problem");
  FGT_REQUIRE(d.q_ld % 8 == 0 && d.k_ld % 8 == 0 && d.vt_ld % 8 == 0 && d.out_ld % 8 == 0, FGT_ERR_ARG,
              "attention: leading dimensions must be multiples of 8 elements");
  FGT_REQUIRE(d.q_batch_stride % 8 == 0 && d.k_batch_stride % 8 == 0 && d.vt_batch_stride % 8 == 0 &&
                  d.out_batch_stride % 8 == 0 && d.out_plane % 8 == 0,
              FGT_ERR_ARG, "attention: batch strides must be multiples of 8 elements");
  FGT_REQUIRE(d.mode == 0 || d.mode == 1, FGT_ERR_ARG, "attention: mode=%d", d.mode);
  FlashParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t hd = static_cast<uint64_t>(d.heads) * 128;
  {
    uint64_t dims[4] = {hd, static_cast<uint64_t>(d.Lq), static_cast<uint64_t>(d.batches), 2};
    uint64_t str[3] = {static_cast<uint64_t>(d.q_ld) * 2, static_cast<uint64_t>(d.q_batch_stride) * 2,
                       static_cast<uint64_t>(d.q_plane) * 2};
    if (d.batches == 1) str[1] = str[0] * dims[1];
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = encode_map_bf16(&p.q_map, d.q_hi, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {hd, static_cast<uint64_t>(d.Lk_rows), static_cast<uint64_t>(d.batches), 2};
    uint64_t str[3] = {static_cast<uint64_t>(d.k_ld) * 2, static_cast<uint64_t>(d.k_batch_stride) * 2,
                       static_cast<uint64_t>(d.k_plane) * 2};
    if (d.batches == 1) str[1] = str[0] * dims[1];
    uint32_t box[4] = {64, 64, 1, 1};
    int rc = encode_map_bf16(&p.k_map, d.k_hi, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(d.Lk_rows), hd, static_cast<uint64_t>(d.batches), 2};
    uint64_t str[3] = {static_cast<uint64_t>(d.vt_ld) * 2, static_cast<uint64_t>(d.vt_batch_stride) * 2,
                       static_cast<uint64_t>(d.vt_plane) * 2};
    if (d.batches == 1) str[1] = str[0] * dims[1];
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = encode_map_bf16(&p.v_map, d.vt_hi, 4, dims, str, box);
    if (rc) return rc;
  }
  p.Lq = d.Lq;
  p.Lk = d.Lk;
  p.heads = d.heads;
  p.mode = d.mode;
  if (d.mode == 1) {
    FGT_REQUIRE(d.glob_count >= 1 && d.glob_start % 64 == 0 && d.Lq % 128 == 0, FGT_ERR_ARG,
                "attention: windowed mode needs glob_count>=1, glob_start%%64==0, Lq%%128==0");
    p.n_glob_tiles = (d.glob_count + 63) / 64;
    p.glob_start = d.glob_start;
    p.glob_count = d.glob_count;
  }
  p.scale_log2 = d.scale * 1.4426950408889634f;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(d.out_hi);
  p.out_plane = d.out_plane;
  p.out_batch_stride = d.out_batch_stride;
  p.out_ld = d.out_ld;
  p.trace = g_flash_trace;
  FGT_REQUIRE((reinterpret_cast<uintptr_t>(d.out_hi) & 15) == 0, FGT_ERR_ARG, "attention: output misaligned");

  static bool attr_set = false;
  if (!attr_set) {
    FGT_CUDA(cudaFuncSetAttribute(flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlashSmem));
    attr_set = true;
  }
  dim3 grid((d.Lq + 127) / 128, d.heads, d.batches);
  flash_kernel<<<grid, 192, kFlashSmem, stream>>>(p);
  FGT_CUDA(cudaGetLastError());
  return FGT_OK;
}

}  // namespace fgt

extern "C" int fgt_attention(const FgtAttnDesc* desc, fgt_stream_t stream) {
  if (!desc) return fgt::set_err(FGT_ERR_ARG, "fgt_attention: null desc");
  return fgt::attention_launch(*desc, reinterpret_cast<cudaStream_t>(stream));
}

// Debug aid (not part of the data path): when given a device buffer of 3*64*8 int64, the next
// fgt_attention launches record clock64 timestamps of CTA (0,0,0) per role / key tile into it.
extern "C" int fgt_debug_flash_trace(long long* device_buf) {
  fgt::g_flash_trace = device_buf;
  return FGT_OK;
}
