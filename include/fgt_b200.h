/* fgt_b200 — C-ABI of the B200-native (sm_100a) kernels behind FGT's flow-guided-attention
 * inference path. Plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 * name ends in _host. All work is enqueued on the stream passed by the caller; the library never
 * allocates or frees device memory and never synchronises unless stated.
 *
 * Every entry point returns 0 on success or a negative FGT_ERR_* code; fgt_last_error() gives the
 * message (thread-local). There is no CPU fallback: without an sm_100 device the calls fail.
 *
 * Data carried between tensor-core kernels uses the "split-bf16" format: a logical fp32 tensor is
 * stored as two bf16 planes, hi = bf16(v) and lo = bf16(v - hi), the lo plane `plane` elements
 * after the hi plane. Three bf16 tcgen05 MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM)
 * reproduce the fp32 product to ~2^-16 relative, which is what lets the path meet the 1e-3
 * end-to-end tolerance against the fp32 reference.
 *
 * Reference interfaces replaced (all under /root/reference, see INTEGRATION.md for the bindings):
 *   fgt_gemm_tc      : every nn.Linear / nn.Conv2d / nn.Conv3d on the inference path —
 *                      FGT/models/model.py:33-50,206-215,95-96; transformer_base/attention_flow.py:
 *                      33-36,44-48,51-54; attention_base.py:38-41; ffn_base.py:39-46;
 *                      LAFC/models/lafc.py:23-79; RAFT/extractor.py:118-192, RAFT/update.py:6-136,
 *                      RAFT/corr.py:52-60 (all-pairs correlation matmul).
 *   fgt_attention    : Attention.forward, FGT/models/transformer_base/attention_base.py:16-22 and
 *                      attention_flow.py:16-22 (softmax(QK^T/sqrt(d))V), with the window/zone
 *                      structure of TMHSA (attention_base.py:93-99) and SWMHSA (attention_flow.py:
 *                      130-164) expressed as key-tile lists instead of permuted copies.
 */
#ifndef FGT_B200_H_
#define FGT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* fgt_stream_t; /* == cudaStream_t */

enum {
  FGT_OK = 0,
  FGT_ERR_ARG = -1,     /* bad shape / alignment / unsupported geometry */
  FGT_ERR_CUDA = -2,    /* CUDA runtime or driver call failed */
  FGT_ERR_DEVICE = -3   /* no sm_100 device / driver entry point missing */
};

enum { FGT_ACT_NONE = 0, FGT_ACT_LEAKY02 = 1, FGT_ACT_RELU = 2, FGT_ACT_SIGMOID = 3, FGT_ACT_TANH = 4 };
enum { FGT_AUX_NONE = 0, FGT_AUX_ADD = 1, FGT_AUX_MUL = 2 };

int fgt_version(void);
const char* fgt_last_error(void);

/* ------------------------------------------------------------------------------------------
 * fgt_gemm_tc: implicit-GEMM convolution / linear layer on tcgen05 tensor cores.
 *
 *   D[m, n] = act(alpha * sum_{tap, seg, c} A_seg[c, x(m)+dx(tap), y(m)+dy(tap), z(m)+dz(tap)]
 *                                          * W[n, tap, seg, c]  + bias[n])   (then aux add / mul)
 *
 * A operands are split-bf16 tensors with channels contiguous (NHWC / token-major). Up to two
 * segments are concatenated along the channel axis (the reference's torch.cat before a conv, and
 * its grouped skip-concat, FGT/models/model.py:57-65). A linear layer is the 1x1x1 case with
 * DX = rows. Strided convolutions are decomposed into `stride*stride` phase views of the input so
 * that every TMA box has unit element stride; zero padding is TMA out-of-bounds fill.
 * W is pre-packed (fgt_b200/packing.py) as split-bf16 [N, k_pad] in K-iteration order, each
 * (tap, segment) block padded to a multiple of 64 channels with zeros.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* hi;      /* bf16 hi plane */
  long long plane;     /* element offset hi -> lo plane */
  int C, DX, DY, DZ;   /* logical extents: channels, x, y, z(images / frames) */
  long long sx, sy, sz;/* element strides of x, y, z (channel stride is 1) */
  int c_base;          /* first channel of the slice consumed (group 0) */
  int c_per_group;     /* channel advance per conv group (0 if groups == 1) */
  int c_count;         /* channels consumed per group from this segment */
} FgtASeg;

typedef struct {
  int num_segs;
  FgtASeg seg[2];
  int kx, ky, kz;            /* filter taps along x, y, z */
  int stride, dil;           /* spatial stride / dilation (x and y) */
  int pad_x, pad_y, pad_z;   /* zero padding (low side; high side implied by out extents) */
  const void* w_hi;          /* packed weights, split-bf16 [N, k_pad] */
  long long w_plane;
  int N;                     /* output channels (all groups) */
  int k_pad;                 /* packed K = kz*ky*kx * sum_seg(ceil(c_count/64)*64) */
  int groups;
  int out_w, out_h, out_z;   /* output extents along x, y, z */
  int box_w, box_h;          /* output tile = box_w x box_h positions (<=128, multiple of 8) */
  int bn;                    /* output-channel tile (multiple of 16, <= 256, divides N/groups if groups>1) */
  /* epilogue */
  const float* bias;         /* [N] or NULL */
  float alpha;               /* scale applied to the accumulator before bias */
  int act;                   /* FGT_ACT_* */
  const float* aux;          /* fp32 tensor addressed like the output, or NULL */
  int aux_mode;              /* FGT_AUX_* : out = aux + v  |  out = aux * v */
  float* out_f32;            /* fp32 output or NULL */
  void* out_hi;              /* split-bf16 output (hi plane) or NULL */
  long long out_plane;
  long long os_z, os_y, os_x, os_c; /* output element strides for (z, y, x, channel) */
  const int* rowmap;         /* linear mode only (out_h == out_z == 1): row -> output row, <0 = drop */
  int lin_batch;             /* linear mode: if >0, z = row / lin_batch, x = row % lin_batch */
} FgtGemmDesc;

int fgt_gemm_tc(const FgtGemmDesc* desc, fgt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FGT_B200_H_ */
