// tc_mlp.cuh -- tiny-MLP layers on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// A tile is 128 rows (= 4 warps x 32 lanes; in the render kernel 4 rays x 32 samples).  Each row lives in one TMEM
// lane:  activations A (fp32, read by the MMA as TF32) in columns [A_hi | A_lo], the fp32 accumulator D in another
// column range.  Weights are the B operand, staged once per CTA in shared memory in the canonical K-major
// no-swizzle ("interleave") UMMA layout.  fp32 accuracy is recovered with the 3xTF32 split
//     x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo        (x_hi = x with the 13 low mantissa bits cleared)
// i.e. three tcgen05.mma per 8-wide k-step accumulating into the same TMEM columns; the dropped x_lo*w_lo term is
// ~2^-22 relative.  One elected thread issues the MMAs; completion is signalled with tcgen05.commit on an mbarrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef NFF_MBAR_HINT
#define NFF_MBAR_HINT 0  // ns; 0 = plain try_wait
#endif

namespace tc {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: returns false instead of hanging the GPU if the MMA never signals (a wrong descriptor must show up
// as a failed test, not as a wedged box).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int it = 0; it < (1 << 20); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#if NFF_MBAR_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"  // suspend-time hint: sleep in hardware, not in the loop
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
#if NFF_MBAR_HINT
        : "r"(addr), "r"(parity), "r"((uint32_t)NFF_MBAR_HINT)
#else
        : "r"(addr), "r"(parity)
#endif
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, TF32 inputs, fp32 accumulate, M = 128
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ------------------------------------------------------------------------------------------- operand layouts
// B operand (weights, nn.Linear [N_real, K_real] row major in global memory) -> shared memory, canonical K-major
// no-swizzle layout (cute: ((8,n),2):((1,SBO),LBO) in 16-byte units): 8-row x 16-byte core matrices, core (nb, kc) at
// nb*SBO + kc*LBO with LBO = 128 B, SBO = (K_pad/4)*128 B; element (n,k) at  +(n%8)*16 + (k%4)*4.
__host__ __device__ constexpr uint32_t b_tile_floats(int n_pad, int k_pad) { return (uint32_t)(n_pad * k_pad); }
__device__ __forceinline__ uint32_t b_elem_offset(int n, int k, int k_pad) {
  return (uint32_t)((n >> 3) * (k_pad >> 2) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3));  // in floats
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// cooperative: all `nthreads` threads of the CTA
__device__ __forceinline__ void stage_b_tile(float* hi, float* lo, const float* __restrict__ w, int n_real, int k_real, int n_pad,
                                             int k_pad, int tid, int nthreads) {
  for (int i = tid; i < n_pad * k_pad; i += nthreads) {
    int n = i / k_pad, k = i % k_pad;
    float v = (n < n_real && k < k_real) ? w[n * k_real + k] : 0.0f;
    float h = tf32_hi(v);
    uint32_t off = b_elem_offset(n, k, k_pad);
    hi[off] = h;
    lo[off] = v - h;  // exact in fp32; the tensor core truncates it to TF32 (error ~2^-22 |v|)
  }
}
// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=SWIZZLE_NONE(0) [61,64)
__device__ __forceinline__ uint64_t b_desc(const float* tile, int k_pad) {
  const uint64_t addr = (uint64_t)((smem_u32(tile) & 0x3ffffu) >> 4);
  const uint64_t lbo = 128u >> 4;
  const uint64_t sbo = (uint64_t)((k_pad >> 2) * 128) >> 4;
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both
// K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ tile ops
// TMEM column map of one 128-row tile (kATotal = max K over the layers)
template <int K_MAX, int N_MAX>
struct TileCols {
  static constexpr int a_hi = 0, a_lo = K_MAX, d = 2 * K_MAX, total = 2 * K_MAX + N_MAX;
};

// Every thread writes `n` (multiple of 8) activations of ITS row into A columns [k0, k0+n).
// lane_base = tmem base | (32*(warp%4)) << 16.
template <int K_MAX>
__device__ __forceinline__ void store_a(uint32_t lane_base, int k0, const float* x, int n) {
#pragma unroll
  for (int c = 0; c < K_MAX; c += 8) {
    if (c < n) {
      uint32_t h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = x[c + i];
        float vh = tf32_hi(v);
        h[i] = __float_as_uint(vh);
        l[i] = __float_as_uint(v - vh);
      }
      tmem_st8(lane_base + (uint32_t)(k0 + c), h);
      tmem_st8(lane_base + (uint32_t)(K_MAX + k0 + c), l);
    }
  }
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
// Issue one layer: D[128 x N] = A[128 x K] * W^T with the 3xTF32 split.  Called by ONE CONVERGED WARP with
// warp-uniform arguments: the descriptors are then warp-uniform values the compiler keeps in uniform registers (where
// tcgen05.mma takes its operands from) and one elected lane issues.  Built inside a single-thread branch instead, every
// MMA pays a ~10-instruction register->uniform-register broadcast loop.
template <int K_MAX>
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, int d_col, const float* b_hi, const float* b_lo, int k_pad, int n_pad,
                                            uint64_t* bar) {
  const uint32_t leader = elect_one();
  const uint32_t idesc = idesc_tf32(128, n_pad);
  // low words: start >> 4 | LBO (128 B) >> 4 << 16; high words: SBO >> 4 | version 1
  const uint32_t h32 = ((smem_u32(b_hi) & 0x3ffffu) >> 4) | ((128u >> 4) << 16);
  const uint32_t l32 = ((smem_u32(b_lo) & 0x3ffffu) >> 4) | ((128u >> 4) << 16);
  const uint32_t hi32 = (uint32_t)(((k_pad >> 2) * 128) >> 4) | (1u << 14);
  const uint32_t d = tmem_base + (uint32_t)d_col;
  for (int ks = 0; ks < k_pad / 8; ++ks) {
    const uint32_t adv = (uint32_t)(ks * 2 * 128) >> 4;  // two 16-byte K-chunks (= 2 core matrices) per k-step
    const uint32_t a_hi = tmem_base + (uint32_t)(ks * 8), a_lo = tmem_base + (uint32_t)(K_MAX + ks * 8);
    const uint64_t dh = ((uint64_t)hi32 << 32) | (h32 + adv), dl = ((uint64_t)hi32 << 32) | (l32 + adv);
    if (leader) {
      mma_tf32_ts(d, a_hi, dh, idesc, ks != 0);
      mma_tf32_ts(d, a_lo, dh, idesc, 1);
      mma_tf32_ts(d, a_hi, dl, idesc, 1);
    }
  }
  if (leader) mma_commit(bar);
  __syncwarp();
}

}  // namespace tc
