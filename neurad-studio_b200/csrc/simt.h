// simt.h -- thin SIMT portability layer.
//
// Under nvcc this maps 1:1 onto the CUDA warp intrinsics (zero overhead).  Under a plain host C++ compiler
// (tests/host_emul only) a "warp" is 32 std::threads that run the SAME device function and meet at every warp
// collective, so the warp-level logic of the render kernel (scans, ballots, searches) can be exercised against
// the oracle on a machine without a GPU.  The host branch is test scaffolding: it is never compiled into
// libb200nerf.so and the product has no CPU path.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------- device
#define NFF_HD __host__ __device__ __forceinline__
#define NFF_D __device__ __forceinline__
#define NFF_RESTRICT __restrict__

namespace simt {
NFF_D int lane() { return threadIdx.x & 31; }
NFF_D float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
NFF_D int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
NFF_D float shfl_up(float v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
NFF_D float shfl_down(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
NFF_D float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
NFF_D unsigned vote_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
// true iff p holds on every lane of the (converged) warp
NFF_D bool vote_all_converged(bool p) { return __all_sync(0xffffffffu, p); }
NFF_D void syncwarp() { __syncwarp(); }
NFF_D void syncblock() { __syncthreads(); }
NFF_D int popc(unsigned x) { return __popc(x); }
// IEEE round-to-nearest single ops that the compiler may NOT contract into FMAs: the reference evaluates every
// elementwise op as a separate torch kernel, so keeping the same roundings keeps grid cells / bins identical.
NFF_D float fmul(float a, float b) { return __fmul_rn(a, b); }
NFF_D float fadd(float a, float b) { return __fadd_rn(a, b); }
NFF_D float fsub(float a, float b) { return __fsub_rn(a, b); }
NFF_D float fdiv(float a, float b) { return __fdiv_rn(a, b); }
NFF_D float frcp(float a) { return __frcp_rn(a); }
NFF_D float fsqrt(float a) { return __fsqrt_rn(a); }
template <typename T>
NFF_D T ldg(const T* p) { return __ldg(p); }
// scatter-accumulate of the backward operators (hash-grid gradients): RED.ADD.F32 to global memory
NFF_D void atomic_add(float* p, float v) { atomicAdd(p, v); }
// one 16-byte vector reduction (REDG.E.ADD.F32x4, sm_90+) for a whole F = 4 hash-table row; p must be 16-byte aligned
NFF_D void atomic_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 8-byte vector reduction (REDG.E.ADD.F32x2) for two ADJACENT floats; p must be 8-byte aligned
NFF_D void atomic_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// hint: bring the line at p into L1 (no register, no scoreboard entry) -- the next sample's inputs of a sequential walk
NFF_D void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
}  // namespace simt

#else
// --------------------------------------------------------------------------------------- host emulation
#include <barrier>
#include <thread>
#include <vector>

#define NFF_HD inline
#define NFF_D inline
#define NFF_RESTRICT __restrict__

struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace simt {
struct EmuWarp {
  std::barrier<> bar{32};
  uint32_t xchg[32];
};
inline thread_local int t_lane = 0;
inline thread_local EmuWarp* t_warp = nullptr;

inline int lane() { return t_lane; }
inline uint32_t exchange_(uint32_t mine, int src) {
  t_warp->xchg[t_lane] = mine;
  t_warp->bar.arrive_and_wait();
  uint32_t r = t_warp->xchg[src & 31];
  t_warp->bar.arrive_and_wait();
  return r;
}
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float shfl(float v, int src) { return u2f(exchange_(f2u(v), src)); }
inline int shfl(int v, int src) { return (int)exchange_((uint32_t)v, src); }
inline float shfl_up(float v, int d) { float r = u2f(exchange_(f2u(v), t_lane - d < 0 ? t_lane : t_lane - d)); return t_lane - d < 0 ? v : r; }
inline float shfl_down(float v, int d) { float r = u2f(exchange_(f2u(v), t_lane + d > 31 ? t_lane : t_lane + d)); return t_lane + d > 31 ? v : r; }
inline float shfl_xor(float v, int m) { return u2f(exchange_(f2u(v), t_lane ^ m)); }
inline unsigned vote_ballot(bool p) {
  t_warp->xchg[t_lane] = p ? 1u : 0u;
  t_warp->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (t_warp->xchg[i] & 1u) << i;
  t_warp->bar.arrive_and_wait();
  return r;
}
// the host emulation of the ray-per-lane kernel runs lanes one after the other: a per-lane decision is exact there
inline bool vote_all_converged(bool p) { return p; }
inline void syncwarp() { t_warp->bar.arrive_and_wait(); }
inline void syncblock() { t_warp->bar.arrive_and_wait(); }  // host emulation runs one warp per block
inline int popc(unsigned x) { return __builtin_popcount(x); }
// compiled with -ffp-contract=off: plain operators are single IEEE roundings
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float fdiv(float a, float b) { return a / b; }
inline float frcp(float a) { return 1.0f / a; }
inline float fsqrt(float a) { return std::sqrt(a); }
template <typename T>
inline T ldg(const T* p) { return *p; }
inline void atomic_add(float* p, float v) { *p += v; }  // the backward emulation runs its "threads" one after the other
inline void atomic_add4(float* p, float a, float b, float c, float d) { p[0] += a; p[1] += b; p[2] += c; p[3] += d; }
inline void atomic_add2(float* p, float a, float b) { p[0] += a; p[1] += b; }
inline void prefetch_l1(const void*) {}
}  // namespace simt

inline float fminf_(float a, float b) { return std::fmin(a, b); }
#endif
