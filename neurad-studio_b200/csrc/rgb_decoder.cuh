// rgb_decoder.cuh -- NeuRADModel.rgb_decoder (models/neurad.py:201-216, model_components/cnns.py:19-46) in eval mode:
//   Conv2d(in->32, 1x1) + ReLU -> 2 x BasicBlock(32, 7x7, BN) -> ConvTranspose2d(32->32, k = s = 3) -> 2 x BasicBlock
//   -> Conv2d(32->3, 1x1) -> Sigmoid,            feature image [B,H,W,in] (row-major rays) -> rgb [B,3H,3W,3].
//
// 97 % of the work is the eight 7x7 convolutions (50 176 MAC per pixel each).  They run as implicit GEMMs on the
// tcgen05 tensor cores: M = 128 consecutive pixels of one image row, N = 32 output channels, K = 49 taps x 32 input
// channels, fp32 accumulators in TMEM.  BatchNorm is folded into the conv weights/bias when the parameters are set.
//
// fp32-level accuracy from bf16 tensor-core inputs: every activation and weight is split into two bf16 numbers
// (hi = bf16(v), lo = bf16(v - hi)) and each k-step issues three MMAs  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  into the
// same accumulator; the dropped a_lo*w_lo term is ~2^-18 relative.  This costs 1.5x the tensor time of a single TF32
// pass (bf16 runs at twice the TF32 rate) and needs exactly the bytes of fp32 storage.
//
// Activations between layers therefore live in HBM already split ("ACT" layout): per pixel 128 B = 8 chunks of 8 bf16,
// chunk c < 4: hi of channels 8c..8c+7, chunk 4+c: lo.  A conv CTA copies a (4+6) x (128+6) pixel window of it into
// shared memory as 8 planes [chunk][row][pixel][16 B]; in that layout the A operand of tap (dy,dx) for output row r is
// the SAME planes read from a shifted start address ((r+dy)*PW + dx)*16 B -- the canonical K-major no-swizzle UMMA
// layout with SBO = 128 B (8-pixel groups are contiguous) and LBO = the plane stride -- so the im2col matrix is never
// materialised.  The folded weights of one tap COLUMN dx ({hi,lo} x 7 taps x 32x32 bf16 = 28 KB, pre-arranged in the
// UMMA B layout by dec_fold_conv_kernel) are double-buffered in shared memory and streamed from L2 while the tensor
// core works on the previous column.
//
// Shared-memory bandwidth, not the tensor pipe, bounds an SS-mode MMA this narrow (128 x 32 x 16: 4 KB of A per 65 k
// MAC), so the loop is arranged to read each A block once for ALL the output rows it feeds: input row i at shift dx
// contributes to output row r through tap dy = i - r, for up to four r at once.  With the four accumulators side by
// side in TMEM ([D0|D1|D2|D3], 32 columns each) and the tap tiles stored in descending dy, that is ONE MMA with N = 128
// (N = 32..96 at the window's top and bottom rows): 420 MMAs per 4-row tile instead of 1176.  Measured, the operand
// fetch sustains ~64 B/clk, which makes shared-memory bytes per MAC the bound of this kernel (profiles/).  The accumulators start from the folded bias (written with tcgen05.st), every MMA accumulates.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_mlp.cuh"

namespace dec {

constexpr int kC = 32;           // hidden channels (rgb_hidden_dim)
constexpr int kUp = 3;           // rgb_upsample_factor
constexpr int kK7 = 7, kPad = 3; // BasicBlock kernel_size / padding
constexpr int kTaps = kK7 * kK7;
constexpr int kStrip = 128;      // pixels per MMA (M)
constexpr int kPW = kStrip + 2 * kPad;
constexpr int kTH = 4;           // output rows per tile
constexpr int kIR = kTH + 2 * kPad;
constexpr int kPlaneBytes = (kIR * kPW * 16 + 127) / 128 * 128;  // TMA destinations are 128-byte aligned
constexpr int kActBytes = 8 * kPlaneBytes;
constexpr int kWTileBytes = kC * kC * 2;                 // one 32x32 bf16 B tile
constexpr int kWRowBytes = kK7 * 2 * kWTileBytes;        // one tap column: {hi, lo} x 7 taps (dy descending)
constexpr int kConvThreads = 256;
constexpr int kTmemCols = 128;                           // kTH accumulators x 32 columns, power of two
static_assert(kTH * kC <= kTmemCols, "accumulators do not fit the TMEM allocation");

struct ConvSmem {
  alignas(128) unsigned char act[kActBytes];
  alignas(128) unsigned char w[2][kWRowBytes];
  float bias[kC];
  float out_w[3 * kC];
  float out_b[4];
  alignas(8) uint64_t bar[2];      // MMA completion per weight buffer
  alignas(8) uint64_t bar_w[2];    // TMA kernel: weight column landed in w[b]
  alignas(8) uint64_t bar_win;     // TMA kernel: input window landed
  alignas(8) uint64_t bar_done;    // TMA kernel: every MMA of the tile has completed (one commit per tile)
  uint32_t tmem_base;
  volatile int abort;  // a completion barrier timed out: every thread leaves at the next block-wide sync
};

static_assert(sizeof(ConvSmem) <= 227 * 1024, "ConvSmem exceeds the 227 KB a CTA may opt in to");

// ------------------------------------------------------------------------------------------------ bf16 split
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// two fp32 -> packed bf16x2 hi word and lo word (element 0 in the low half = lower address)
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 32 channels of one pixel -> ACT (8 x uint4), registers only
__device__ __forceinline__ void store_act(uint4* dst, const float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 h, l;
    split_pack2(v[8 * c + 0], v[8 * c + 1], h.x, l.x);
    split_pack2(v[8 * c + 2], v[8 * c + 3], h.y, l.y);
    split_pack2(v[8 * c + 4], v[8 * c + 5], h.z, l.z);
    split_pack2(v[8 * c + 6], v[8 * c + 7], h.w, l.w);
    dst[c] = h;
    dst[4 + c] = l;
  }
}
__device__ __forceinline__ void unpack2(uint32_t h, uint32_t l, float& a, float& b) {
  a = __uint_as_float(h << 16) + __uint_as_float(l << 16);
  b = __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u);
}
__device__ __forceinline__ void load_act(const uint4* src, float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 h = src[c], l = src[4 + c];
    unpack2(h.x, l.x, v[8 * c + 0], v[8 * c + 1]);
    unpack2(h.y, l.y, v[8 * c + 2], v[8 * c + 3]);
    unpack2(h.z, l.z, v[8 * c + 4], v[8 * c + 5]);
    unpack2(h.w, l.w, v[8 * c + 6], v[8 * c + 7]);
  }
}

// ---------------------------------------------------------------------------------------- parameter preparation
// Conv2d [co][ci][7][7] + BatchNorm2d (eval) -> folded  w' = w * s[co],  b' = (b - mean) * s + beta,  s = gamma /
// sqrt(var + eps)  (BasicBlock.main_branch, cnns.py:37-43), written as
//   w_img : [dx][hi|lo][6 - dy] 32x32 bf16 tiles in the UMMA K-major no-swizzle B layout (n = co, k = ci):
//           byte offset of (n,k) = (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2.  Tiles of one tap COLUMN are contiguous
//           in DESCENDING dy, so that [W(dy), W(dy-1), W(dy-2)] is one N = 96 B operand (see dec_conv7_tc_kernel)
//   w_f32 : [dy][dx][ci][co] fp32 (CUDA-core reference kernel)
//   bias  : [co]
__global__ void dec_fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                                     float eps, unsigned char* __restrict__ w_img, float* __restrict__ w_f32, float* __restrict__ bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kC) {
    const float s = gamma[i] / sqrtf(var[i] + eps);
    bias[i] = (b[i] - mean[i]) * s + beta[i];
  }
  if (i >= kTaps * kC * kC) return;
  const int co = i / (kC * kTaps), ci = (i / kTaps) % kC, tap = i % kTaps;  // i = flat index of w[co][ci][dy][dx]
  const float s = gamma[co] / sqrtf(var[co] + eps);
  const float v = w[i] * s;
  w_f32[(tap * kC + ci) * kC + co] = v;
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  const int off = (co >> 3) * 512 + (ci >> 3) * 128 + (co & 7) * 16 + (ci & 7) * 2;
  const int dy = tap / kK7, dx = tap % kK7;
  unsigned char* tile = w_img + (size_t)dx * kWRowBytes + (size_t)(kK7 - 1 - dy) * kWTileBytes;
  *reinterpret_cast<__nv_bfloat16*>(tile + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(tile + kK7 * kWTileBytes + off) = lo;
}

// ------------------------------------------------------------------------------------- 1x1 input conv + ReLU
// rgb_decoder.0/.1: features fp32 [P, in_dim] -> ACT [P]; w [32][in_dim], b [32] (Conv2d 1x1 layout)
__global__ void dec_input_kernel(const float* __restrict__ x, int64_t n_pix, int in_dim, const float* __restrict__ w,
                                 const float* __restrict__ b, uint4* __restrict__ out) {
  extern __shared__ float sw[];  // [in_dim][32] transposed + bias
  for (int i = threadIdx.x; i < in_dim * kC; i += blockDim.x) sw[(i % in_dim) * kC + i / in_dim] = w[i];
  for (int i = threadIdx.x; i < kC; i += blockDim.x) sw[in_dim * kC + i] = b[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  float acc[kC];
#pragma unroll
  for (int k = 0; k < kC; ++k) acc[k] = sw[in_dim * kC + k];
  for (int c = 0; c < in_dim; ++c) {
    const float v = x[p * in_dim + c];
#pragma unroll
    for (int k = 0; k < kC; ++k) acc[k] = fmaf(v, sw[c * kC + k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < kC; ++k) acc[k] = fmaxf(acc[k], 0.f);
  store_act(out + p * 8, acc);
}

// --------------------------------------------------------------------------- ConvTranspose2d, kernel = stride = 3
// rgb_decoder.4: out[3y+i][3x+j][co] = b[co] + sum_ci in[y][x][ci] * w[ci][co][i][j].  One thread owns TWO input pixels
// and produces their 2 x 9 output pixels; the weights [ij][ci][co] sit in shared memory and every read is a warp-wide
// broadcast feeding 8 FMAs (2 pixels x 4 channels).
constexpr int kUpThreads = 128;
__global__ void __launch_bounds__(kUpThreads) dec_upsample_kernel(const uint4* __restrict__ in, int batch, int H, int W,
                                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                                   uint4* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];  // [i*3+j][ci][co] + bias
  for (int i = threadIdx.x; i < kC * kC * kUp * kUp; i += blockDim.x) {
    const int ci = i / (kC * 9), co = (i / 9) % kC, ij = i % 9;
    sw[(ij * kC + ci) * kC + co] = w[i];
  }
  for (int i = threadIdx.x; i < kC; i += blockDim.x) sw[9 * kC * kC + i] = b[i];
  __syncthreads();
  const int64_t n_in = (int64_t)batch * H * W;
  const int64_t p0 = ((int64_t)blockIdx.x * kUpThreads + threadIdx.x) * 2;
  if (p0 >= n_in) return;
  const bool two = p0 + 1 < n_in;
  float v0[kC], v1[kC];
  load_act(in + p0 * 8, v0);
  load_act(in + (two ? p0 + 1 : p0) * 8, v1);
  const int64_t WO = (int64_t)W * kUp;
  int64_t base[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int64_t p = p0 + q, img = p / ((int64_t)H * W), y = (p / W) % H, x = p % W;
    base[q] = (img * H * kUp + y * kUp) * WO + x * kUp;
  }
#pragma unroll 1
  for (int ij = 0; ij < 9; ++ij) {
    float a0[kC], a1[kC];
#pragma unroll
    for (int k = 0; k < kC; ++k) a0[k] = a1[k] = sw[9 * kC * kC + k];
    const float4* wt = reinterpret_cast<const float4*>(sw + ij * kC * kC);
#pragma unroll
    for (int c = 0; c < kC; ++c) {
#pragma unroll
      for (int k4 = 0; k4 < kC / 4; ++k4) {
        const float4 ww = wt[c * (kC / 4) + k4];
        a0[4 * k4 + 0] = fmaf(v0[c], ww.x, a0[4 * k4 + 0]); a1[4 * k4 + 0] = fmaf(v1[c], ww.x, a1[4 * k4 + 0]);
        a0[4 * k4 + 1] = fmaf(v0[c], ww.y, a0[4 * k4 + 1]); a1[4 * k4 + 1] = fmaf(v1[c], ww.y, a1[4 * k4 + 1]);
        a0[4 * k4 + 2] = fmaf(v0[c], ww.z, a0[4 * k4 + 2]); a1[4 * k4 + 2] = fmaf(v1[c], ww.z, a1[4 * k4 + 2]);
        a0[4 * k4 + 3] = fmaf(v0[c], ww.w, a0[4 * k4 + 3]); a1[4 * k4 + 3] = fmaf(v1[c], ww.w, a1[4 * k4 + 3]);
      }
    }
    const int64_t off = (int64_t)(ij / 3) * WO + ij % 3;
    store_act(out + (base[0] + off) * 8, a0);
    if (two) store_act(out + (base[1] + off) * 8, a1);
  }
}

// ------------------------------------------------------------------------------------------------ epilogues
enum { EPI_RELU = 0, EPI_RES_RELU = 1, EPI_RES_RELU_RGB = 2 };
// acc = conv + folded bias for one pixel.  EPI_RELU: first conv of a BasicBlock.  EPI_RES_RELU: second conv:
// relu(x + main_branch(x)) (cnns.py:31-32).  EPI_RES_RELU_RGB: the last block, followed by rgb_decoder.7/.8
// (Conv2d 32->3 1x1 + Sigmoid) while the pixel is still in registers -> fp32 rgb.
template <int EPI>
__device__ __forceinline__ void conv_epilogue(float* acc, const uint4* __restrict__ residual, int64_t pix, uint4* __restrict__ out_act,
                                              float* __restrict__ out_rgb, const float* out_w, const float* out_b) {
  if (EPI != EPI_RELU) {
    float x[kC];
    load_act(residual + pix * 8, x);
#pragma unroll
    for (int k = 0; k < kC; ++k) acc[k] += x[k];
  }
#pragma unroll
  for (int k = 0; k < kC; ++k) acc[k] = fmaxf(acc[k], 0.f);
  if (EPI == EPI_RES_RELU_RGB) {
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      float s = out_b[o];
#pragma unroll
      for (int k = 0; k < kC; ++k) s = fmaf(acc[k], out_w[o * kC + k], s);
      out_rgb[pix * 3 + o] = 1.0f / (1.0f + expf(-s));
    }
  } else {
    store_act(out_act + pix * 8, acc);
  }
}

// ------------------------------------------------------------------------------- 7x7 conv on the tensor cores
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48), SWIZZLE_NONE
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor for kind::f16: D = F32 (1<<4), A = B = BF16 (1<<7, 1<<10), both K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Bounded completion wait (a wrong descriptor must surface as a failed status, not as a wedged GPU).
__device__ __forceinline__ bool bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  for (int it = 0; it < (1 << 17); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ uint32_t elect_one() {  // one lane of the converged warp (the same one every time)
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 writes zeros (the conv's padding) without
// reading.  Many of these are in flight per thread, which is what hides the L2 latency of the window / weight loads.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// this thread's TMEM lane, 32 accumulator columns <- the folded bias
__device__ __forceinline__ void arm_accumulator(uint32_t taddr, const float* bias) {
#pragma unroll
  for (int c = 0; c < kC; c += 8) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(bias[c + k]);
    tc::tmem_st8(taddr + (uint32_t)c, v);
  }
}

struct ConvArgs {
  const uint4* in;        // ACT [B][H][W]
  const uint4* residual;  // ACT (EPI_RES_*) or nullptr
  uint4* out_act;         // ACT or nullptr
  float* out_rgb;         // [B][H][W][3] (EPI_RES_RELU_RGB)
  const unsigned char* w_img;  // [7 dx][kWRowBytes]
  const float* w_f32;     // [49][ci][co]
  const float* bias;      // [32] folded
  const float* out_w;     // [3][32] (EPI_RES_RELU_RGB)
  const float* out_b;     // [3]
  int batch, H, W;
  int* status;
};

template <int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) dec_conv7_tc_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ConvSmem& S = *reinterpret_cast<ConvSmem*>(smem_raw);
  const int tid = threadIdx.x, ln = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  if (tid < kC) S.bias[tid] = a.bias[tid];
  if (EPI == EPI_RES_RELU_RGB) {
    if (tid < 3 * kC) S.out_w[tid] = a.out_w[tid];
    if (tid < 3) S.out_b[tid] = a.out_b[tid];
  }
  if (warp == 0) tc::tmem_alloc(&S.tmem_base, kTmemCols);
  if (tid == 0) {
    tc::mbar_init(&S.bar[0], 1);
    tc::mbar_init(&S.bar[1], 1);
    S.abort = 0;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);
  // low 32 bits of the shared-memory descriptors: start address >> 4 in [0,14), LBO >> 4 in [16,30); the high words
  // (SBO >> 4, version 1, no swizzle) are constants.  Advancing an operand = adding 16-byte units to the low word.
  const uint32_t a_lo32 = ((tc::smem_u32(S.act) & 0x3ffffu) >> 4) | ((uint32_t)(kPlaneBytes >> 4) << 16);
  const uint32_t b_lo32[2] = {((tc::smem_u32(S.w[0]) & 0x3ffffu) >> 4) | ((128u >> 4) << 16),
                              ((tc::smem_u32(S.w[1]) & 0x3ffffu) >> 4) | ((128u >> 4) << 16)};
  const uint32_t act_u32 = tc::smem_u32(S.act);
  const uint32_t w_u32[2] = {tc::smem_u32(S.w[0]), tc::smem_u32(S.w[1])};
  constexpr uint32_t kDescHiA = (128u >> 4) | (1u << 14), kDescHiB = (512u >> 4) | (1u << 14);  // bits [32,64)
  uint32_t parity[2] = {0u, 0u};

  const int tiles_x = (a.W + kStrip - 1) / kStrip, tiles_y = (a.H + kTH - 1) / kTH;
  const int64_t n_tiles = (int64_t)a.batch * tiles_y * tiles_x;
  // asynchronous fill of the input window (8 chunk planes; pixels outside the image are the conv's zero padding) and of
  // tap row 0 of the weights for one tile
  auto issue_tile_loads = [&](int64_t tile) {
    const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
    const int64_t img = tile / ((int64_t)tiles_x * tiles_y);
    const int x0 = tx * kStrip, y0 = ty * kTH;
    const uint4* in_img = a.in + img * (int64_t)a.H * a.W * 8;
    for (int i = tid; i < kIR * kPW * 8; i += kConvThreads) {
      const int c = i & 7, ip = (i >> 3) % kPW, ir = (i >> 3) / kPW;
      const int y = y0 - kPad + ir, x = x0 - kPad + ip;
      const bool inside = y >= 0 && y < a.H && x >= 0 && x < a.W;
      const uint4* src = inside ? in_img + ((int64_t)y * a.W + x) * 8 + c : in_img;
      cp_async16(act_u32 + (uint32_t)(c * kPlaneBytes + (ir * kPW + ip) * 16), src, inside ? 16u : 0u);
    }
    const uint4* src = reinterpret_cast<const uint4*>(a.w_img);
    for (int i = tid; i < kWRowBytes / 16; i += kConvThreads) cp_async16(w_u32[0] + (uint32_t)(i * 16), src + i, 16u);
  };
  if ((int64_t)blockIdx.x < n_tiles) issue_tile_loads(blockIdx.x);
  for (int r = warp >> 2; r < kTH; r += 2) arm_accumulator(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(r * kC), S.bias);
  tc::wait_st();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
    const int64_t img = tile / ((int64_t)tiles_x * tiles_y);
    const int x0 = tx * kStrip, y0 = ty * kTH;
    cp_async_wait_all();
    tc::fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
    tc::fence_before_sync();
    __syncthreads();
    for (int dx = 0; dx < kK7; ++dx) {
      const int buf = dx & 1, nb = buf ^ 1;
      if (warp == 0) {
        // The whole (converged) warp walks the loop so that every descriptor is a warp-uniform value the compiler keeps
        // in uniform registers -- tcgen05.mma takes its operands from there; built inside a single-thread branch each
        // MMA pays a register->uniform broadcast loop -- and one elected lane issues.
        tc::fence_after_sync();
        const uint32_t leader = elect_one();
#pragma unroll
        for (int i = 0; i < kIR; ++i) {  // input row i feeds output rows r_min..r_max through taps dy = i - r
          constexpr int kLast = kK7 - 1;
          const int r_min = i > kLast ? i - kLast : 0, r_max = i < kTH - 1 ? i : kTH - 1, nr = r_max - r_min + 1;
          const uint32_t d = tmem + (uint32_t)(r_min * kC);
          const uint32_t idesc = idesc_bf16(kStrip, kC * nr);
          const uint32_t slot = (uint32_t)(kLast - (i - r_min));  // first (largest-dy) tile of the N-concatenated B
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {  // 16 input channels (two 8-channel chunks) per MMA
            const uint32_t ah = a_lo32 + (uint32_t)(i * kPW + 2 * ks * (kPlaneBytes / 16)) + (uint32_t)dx;
            const uint32_t al = ah + (uint32_t)(4 * (kPlaneBytes / 16));
            const uint32_t bh = b_lo32[buf] + (slot * kWTileBytes + (uint32_t)(ks * 256)) / 16;
            const uint32_t bl = bh + (uint32_t)(kK7 * kWTileBytes / 16);
            if (leader) {
              mma_bf16_ss(d, make_desc(ah, kDescHiA), make_desc(bh, kDescHiB), idesc, 1);
              mma_bf16_ss(d, make_desc(al, kDescHiA), make_desc(bh, kDescHiB), idesc, 1);
              mma_bf16_ss(d, make_desc(ah, kDescHiA), make_desc(bl, kDescHiB), idesc, 1);
            }
          }
        }
        if (leader) tc::mma_commit(&S.bar[buf]);
        __syncwarp();
        if (dx >= 1 && dx + 1 < kK7) parity[nb] ^= 1u;  // keep the phase bookkeeping of the loading warps
      } else if (dx + 1 < kK7) {
        // warps 1..7 stream the next tap column while warp 0 is busy issuing
        if (dx >= 1) {  // the MMAs of column dx-1 read w[nb]: wait for them before overwriting it
          if (!bar_wait(&S.bar[nb], parity[nb])) S.abort = 1;
          parity[nb] ^= 1u;
        }
        const uint4* src = reinterpret_cast<const uint4*>(a.w_img + (size_t)(dx + 1) * kWRowBytes);
        for (int i = tid - 32; i < kWRowBytes / 16; i += kConvThreads - 32) cp_async16(w_u32[nb] + (uint32_t)(i * 16), src + i, 16u);
        cp_async_wait_all();
        tc::fence_async_smem();
      }
      if (dx + 1 < kK7) {
        tc::fence_before_sync();
        __syncthreads();
        if (S.abort) break;
      }
    }
    if (S.abort) break;
    // tap columns 5 (bar[1]) and 6 (bar[0]) are still outstanding; the commit of column 6 covers every earlier MMA
    if (!bar_wait(&S.bar[1], parity[1])) S.abort = 1;
    parity[1] ^= 1u;
    if (!bar_wait(&S.bar[0], parity[0])) S.abort = 1;
    parity[0] ^= 1u;
    tc::fence_after_sync();
    // every MMA of this tile has completed: the window and both weight buffers are free, so the next tile's loads fly
    // while this tile's accumulators are drained
    if (tile + gridDim.x < n_tiles) issue_tile_loads(tile + gridDim.x);
    // ---- epilogue: warps 0-3 take output rows 0 and 2, warps 4-7 rows 1 and 3; a thread owns one pixel (= TMEM lane)
    const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int m = 32 * (warp & 3) + ln;
    for (int r = warp >> 2; r < kTH; r += 2) {
      uint32_t dreg[kC];
      tc::tmem_ld16(lane_base + (uint32_t)(r * kC), dreg);
      tc::tmem_ld16(lane_base + (uint32_t)(r * kC + 16), dreg + 16);
      tc::wait_ld();
      arm_accumulator(lane_base + (uint32_t)(r * kC), S.bias);  // folded bias: the next tile's MMAs all accumulate
      const int y = y0 + r, x = x0 + m;
      if (y < a.H && x < a.W) {
        float acc[kC];
#pragma unroll
        for (int k = 0; k < kC; ++k) acc[k] = __uint_as_float(dreg[k]);
        const int64_t pix = (img * a.H + y) * a.W + x;
        conv_epilogue<EPI>(acc, a.residual, pix, a.out_act, a.out_rgb, S.out_w, S.out_b);
      }
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();  // accumulators and the input window are free again
    tc::fence_after_sync();
    if (S.abort) break;
  }
  cp_async_wait_all();
  if (S.abort && tid == 0 && a.status) atomicExch(a.status, 2);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------- 7x7 conv on the tensor cores, TMA operand loads
// Same tile loop and MMA schedule as dec_conv7_tc_kernel, but the operands are moved by the TMA engine instead of
// 16-byte LDGSTS issued by every thread (9 648 per tile: the LSU issue alone cost ~11 k cycles per tile):
//   * the input window is 8 tensor copies (one per chunk plane) from a 5-D tensor map over the ACT buffer
//     [image][y][x][chunk][8 bf16] with box (8, 1, 134, 9, 1); coordinates may be negative / beyond the image, the TMA
//     zero-fills, which IS the convolution's padding;
//   * a weight column is one 28 KB bulk copy.
// One elected lane of warp 1 is the producer, warp 0 issues the MMAs; they hand buffers to each other through mbarriers
// (transaction-count barriers for the loads, tcgen05.commit barriers for the MMAs), so the tap-column loop has no
// block-wide barrier at all.  All 8 warps run the epilogue.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(tc::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(tc::smem_u32(bar))
               : "memory");
}

struct ConvArgsTma {
  alignas(64) CUtensorMap in_map;  // ACT input as [B][H][W][8][8 x bf16], box (8,1,kPW,kIR,1)
  ConvArgs a;
};

template <int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) dec_conv7_tma_kernel(const __grid_constant__ ConvArgsTma P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ConvSmem& S = *reinterpret_cast<ConvSmem*>(smem_raw);
  const ConvArgs& a = P.a;
  const int tid = threadIdx.x, ln = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  if (tid < kC) S.bias[tid] = a.bias[tid];
  if (EPI == EPI_RES_RELU_RGB) {
    if (tid < 3 * kC) S.out_w[tid] = a.out_w[tid];
    if (tid < 3) S.out_b[tid] = a.out_b[tid];
  }
  if (warp == 0) tc::tmem_alloc(&S.tmem_base, kTmemCols);
  if (tid == 0) {
    tc::mbar_init(&S.bar[0], 1);
    tc::mbar_init(&S.bar[1], 1);
    tc::mbar_init(&S.bar_w[0], 1);
    tc::mbar_init(&S.bar_w[1], 1);
    tc::mbar_init(&S.bar_win, 1);
    tc::mbar_init(&S.bar_done, 1);
    S.abort = 0;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);
  const uint32_t a_lo32 = ((tc::smem_u32(S.act) & 0x3ffffu) >> 4) | ((uint32_t)(kPlaneBytes >> 4) << 16);
  const uint32_t b_lo32[2] = {((tc::smem_u32(S.w[0]) & 0x3ffffu) >> 4) | ((128u >> 4) << 16),
                              ((tc::smem_u32(S.w[1]) & 0x3ffffu) >> 4) | ((128u >> 4) << 16)};
  const uint32_t act_u32 = tc::smem_u32(S.act);
  const uint32_t w_u32[2] = {tc::smem_u32(S.w[0]), tc::smem_u32(S.w[1])};
  constexpr uint32_t kDescHiA = (128u >> 4) | (1u << 14), kDescHiB = (512u >> 4) | (1u << 14);  // bits [32,64)
  constexpr uint32_t kWinBytes = 8u * kIR * kPW * 16u;

  const int tiles_x = (a.W + kStrip - 1) / kStrip, tiles_y = (a.H + kTH - 1) / kTH;
  const int64_t n_tiles = (int64_t)a.batch * tiles_y * tiles_x;
  // per-barrier completion counters (identical in every thread): phase parity of completion #k is k & 1
  uint32_t n_mma[2] = {0u, 0u}, n_w[2] = {0u, 0u}, n_win = 0u;
  const bool producer_warp = warp == 1;

  // producer: window of `tile` -> act planes, tap column 0 -> w[0]
  auto issue_tile_loads = [&](int64_t tile) {
    const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
    const int img = (int)(tile / ((int64_t)tiles_x * tiles_y));
    mbar_expect_tx(&S.bar_win, kWinBytes);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      tma_load_5d(act_u32 + (uint32_t)(c * kPlaneBytes), &P.in_map, 0, c, tx * kStrip - kPad, ty * kTH - kPad, img, &S.bar_win);
    mbar_expect_tx(&S.bar_w[0], kWRowBytes);
    bulk_load(w_u32[0], a.w_img, kWRowBytes, &S.bar_w[0]);
  };
  if (producer_warp && (int64_t)blockIdx.x < n_tiles) {
    if (elect_one()) issue_tile_loads(blockIdx.x);
    __syncwarp();
  }
  for (int r = warp >> 2; r < kTH; r += 2) arm_accumulator(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(r * kC), S.bias);
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
    const int64_t img = tile / ((int64_t)tiles_x * tiles_y);
    const int x0 = tx * kStrip, y0 = ty * kTH;
    if (warp == 0) {
      // ---- MMA warp: window landed, then per tap column: column landed -> 54 MMAs -> commit
      if (!bar_wait(&S.bar_win, n_win & 1u)) S.abort = 1;
      tc::fence_after_sync();
      const uint32_t leader = elect_one();
#pragma unroll 1
      for (int dx = 0; dx < kK7; ++dx) {
        const int buf = dx & 1;
        if (!bar_wait(&S.bar_w[buf], (n_w[buf] + (uint32_t)(dx >> 1)) & 1u)) S.abort = 1;
        tc::fence_after_sync();
#pragma unroll
        for (int i = 0; i < kIR; ++i) {  // input row i feeds output rows r_min..r_max through taps dy = i - r
          constexpr int kLast = kK7 - 1;
          const int r_min = i > kLast ? i - kLast : 0, r_max = i < kTH - 1 ? i : kTH - 1, nr = r_max - r_min + 1;
          const uint32_t d = tmem + (uint32_t)(r_min * kC);
          const uint32_t idesc = idesc_bf16(kStrip, kC * nr);
          const uint32_t slot = (uint32_t)(kLast - (i - r_min));
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ah = a_lo32 + (uint32_t)(i * kPW + 2 * ks * (kPlaneBytes / 16)) + (uint32_t)dx;
            const uint32_t al = ah + (uint32_t)(4 * (kPlaneBytes / 16));
            const uint32_t bh = b_lo32[buf] + (slot * kWTileBytes + (uint32_t)(ks * 256)) / 16;
            const uint32_t bl = bh + (uint32_t)(kK7 * kWTileBytes / 16);
            if (leader) {
              mma_bf16_ss(d, make_desc(ah, kDescHiA), make_desc(bh, kDescHiB), idesc, 1);
              mma_bf16_ss(d, make_desc(al, kDescHiA), make_desc(bh, kDescHiB), idesc, 1);
              mma_bf16_ss(d, make_desc(ah, kDescHiA), make_desc(bl, kDescHiB), idesc, 1);
            }
          }
        }
        if (leader) tc::mma_commit(&S.bar[buf]);
        __syncwarp();
      }
      if (leader) tc::mma_commit(&S.bar_done);
      __syncwarp();
    } else if (producer_warp) {
      // ---- producer: column dx+1 into the buffer the MMAs of column dx-1 have finished reading
      const uint32_t leader = elect_one();
#pragma unroll 1
      for (int dx = 0; dx + 1 < kK7; ++dx) {
        const int nb = (dx & 1) ^ 1;
        if (dx >= 1 && !bar_wait(&S.bar[nb], (n_mma[nb] + (uint32_t)((dx - 1) >> 1)) & 1u)) S.abort = 1;
        if (leader) {
          mbar_expect_tx(&S.bar_w[nb], kWRowBytes);
          bulk_load(w_u32[nb], a.w_img + (size_t)(dx + 1) * kWRowBytes, kWRowBytes, &S.bar_w[nb]);
        }
        __syncwarp();
      }
    }
    // Every warp waits for the tile's ONE bar_done completion.  (A parity wait is only sound for a waiter that is at
    // most one completion behind or ahead of the barrier, so the epilogue warps, which skip the per-column
    // completions, get their own once-per-tile barrier; the producer and the MMA warp follow theirs one by one.)
    if (!bar_wait(&S.bar_done, n_win & 1u)) S.abort = 1;
    // completions of this tile: MMA barriers 4 (columns 0,2,4,6) + 3 (1,3,5); weight barriers 4 + 3; window / done 1
    n_mma[0] += 4u; n_mma[1] += 3u; n_w[0] += 4u; n_w[1] += 3u; n_win += 1u;
    tc::fence_after_sync();
    // the window and both weight buffers are free: the next tile's loads fly while the accumulators are drained
    if (producer_warp && tile + gridDim.x < n_tiles) {
      if (elect_one()) issue_tile_loads(tile + gridDim.x);
      __syncwarp();
    }
    // ---- epilogue: warps 0-3 take output rows 0 and 2, warps 4-7 rows 1 and 3; a thread owns one pixel (= TMEM lane)
    const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int m = 32 * (warp & 3) + ln;
    for (int r = warp >> 2; r < kTH; r += 2) {
      uint32_t dreg[kC];
      tc::tmem_ld16(lane_base + (uint32_t)(r * kC), dreg);
      tc::tmem_ld16(lane_base + (uint32_t)(r * kC + 16), dreg + 16);
      tc::wait_ld();
      arm_accumulator(lane_base + (uint32_t)(r * kC), S.bias);  // folded bias: the next tile's MMAs all accumulate
      const int y = y0 + r, x = x0 + m;
      if (y < a.H && x < a.W) {
        float acc[kC];
#pragma unroll
        for (int k = 0; k < kC; ++k) acc[k] = __uint_as_float(dreg[k]);
        const int64_t pix = (img * a.H + y) * a.W + x;
        conv_epilogue<EPI>(acc, a.residual, pix, a.out_act, a.out_rgb, S.out_w, S.out_b);
      }
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();  // accumulators re-armed and drained by everyone before the next tile's MMAs
    tc::fence_after_sync();
    if (S.abort) break;
  }
  if (S.abort && tid == 0 && a.status) atomicExch(a.status, 2);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------ 7x7 conv on the CUDA cores (fp32 reference)
// Same inputs, outputs and epilogues as dec_conv7_tc_kernel; one thread per output pixel, the weights of one tap row
// (7 x 32 x 32 fp32 = 28 KB) staged in shared memory.  ~20x slower; kept as the in-library cross-check of the
// tensor-core path (b200nerf_rgb_decode_fwd impl = 1) and for the numerics tests.
template <int EPI>
__global__ void __launch_bounds__(128) dec_conv7_ref_kernel(const ConvArgs a) {
  __shared__ float sw[kK7 * kC * kC];
  __shared__ float s_out[3 * kC + 4];
  const int tid = threadIdx.x;
  if (EPI == EPI_RES_RELU_RGB) {
    if (tid < 3 * kC) s_out[tid] = a.out_w[tid];
    if (tid < 3) s_out[3 * kC + tid] = a.out_b[tid];
  }
  const int tiles_x = (a.W + 127) / 128;
  const int64_t row_id = blockIdx.x / tiles_x;  // (img, y)
  const int x = (blockIdx.x % tiles_x) * 128 + tid;
  const int64_t img = row_id / a.H;
  const int y = (int)(row_id % a.H);
  float acc[kC];
#pragma unroll
  for (int k = 0; k < kC; ++k) acc[k] = a.bias[k];
  for (int dy = 0; dy < kK7; ++dy) {
    __syncthreads();
    for (int i = tid; i < kK7 * kC * kC; i += 128) sw[i] = a.w_f32[(size_t)dy * kK7 * kC * kC + i];
    __syncthreads();
    const int yy = y + dy - kPad;
    if (yy < 0 || yy >= a.H || x >= a.W) continue;
    for (int dx = 0; dx < kK7; ++dx) {
      const int xx = x + dx - kPad;
      if (xx < 0 || xx >= a.W) continue;
      float v[kC];
      load_act(a.in + ((img * a.H + yy) * (int64_t)a.W + xx) * 8, v);
      const float* wt = sw + dx * kC * kC;
#pragma unroll 4
      for (int c = 0; c < kC; ++c) {
#pragma unroll
        for (int k = 0; k < kC; ++k) acc[k] = fmaf(v[c], wt[c * kC + k], acc[k]);
      }
    }
  }
  if (x < a.W) conv_epilogue<EPI>(acc, a.residual, (img * a.H + y) * (int64_t)a.W + x, a.out_act, a.out_rgb, s_out, s_out + 3 * kC);
}

}  // namespace dec
