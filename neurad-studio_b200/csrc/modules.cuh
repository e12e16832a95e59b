// modules.cuh -- kernels of the module-level operators (SURVEY.md 8b seams "Field", "Sampler", "Encoding op"): the
// reference's nn.Modules exchange [N,S,...] tensors, so these kernels read and write them (the fused renderer does not).
// Device logic lives in nff_modules.h (shared with the host emulation used by the CPU tests).
#pragma once

#include "nff_modules.h"

namespace nff {

constexpr int kModWarps = 8;

// Frustums.get_fast_isotropic_gaussian(num_multisamples=1) (cameras/rays.py:109-124): one thread per sample.
__global__ void isotropic_gaussian_kernel(const float* __restrict__ origins, const float* __restrict__ dirs,
                                          const float* __restrict__ area, const float* __restrict__ bins_e,
                                          int64_t n_rays, int S, float* __restrict__ mean, float* __restrict__ std) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * S) return;
  int64_t ray = i / S;
  int s = (int)(i % S);
  float o[3] = {origins[3 * ray], origins[3 * ray + 1], origins[3 * ray + 2]};
  float d[3] = {dirs[3 * ray], dirs[3 * ray + 1], dirs[3 * ray + 2]};
  Gauss g = sample_gaussian(o, d, area[ray], bins_e[ray * (S + 1) + s], bins_e[ray * (S + 1) + s + 1]);
  mean[3 * i] = g.x;
  mean[3 * i + 1] = g.y;
  mean[3 * i + 2] = g.z;
  std[i] = g.std;
}

struct EncodingArgs {
  const float* mean;    // [N,S,3]
  const float* std;     // [N,S]
  const float* times;   // [N] (the reference reads times[:, 0], neurad_encoding.py:194)
  const float* dirs;    // [N,3] (dirs_per_ray) or [N,S,3] or NULL
  float* features;      // [N*S, D] or NULL
  float* density;       // [N,S] or NULL: trunc_exp(Linear(D,1,bias=False)(features)) (neurad_field.py:208-213)
  float* dirs_out;      // [N,S,3] or NULL
  int32_t* actor_id;    // [N,S] or NULL
  int64_t n_rays;
  int32_t S, dirs_per_ray;
};

// NeuRADHashEncoding.forward (neurad_encoding.py:150-187), optionally followed by the proposal field's density head.
// One warp per ray: the lanes first build the ray's actor frames (lane = actor) in shared memory, then stride over
// the ray's samples.
__global__ void __launch_bounds__(kModWarps * 32) neurad_encoding_fwd_kernel(const FieldGrids fg, const Actors A,
                                                                              const EncodingArgs a) {
  __shared__ ActorFrame frames[kModWarps][kModMaxActors];
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kModWarps + warp;
  if (ray >= a.n_rays) return;
  if (A.n_actors > 0) {
    int left, right;
    float frac;
    keyframe_bracket(A, a.times[ray], left, right, frac);
    for (int k = ln; k < A.n_actors; k += 32) actor_frame(A, k, left, right, frac, frames[warp][k]);
  }
  __syncwarp();
  const int D = fg.stat.L * fg.stat.F;
  for (int s = ln; s < a.S; s += 32) {
    const int64_t i = ray * a.S + s;
    Gauss g = {a.mean[3 * i], a.mean[3 * i + 1], a.mean[3 * i + 2], a.std[i]};
    float dir[3] = {0.f, 0.f, 0.f};
    if (a.dirs) {
      const float* dp = a.dirs + 3 * (a.dirs_per_ray ? ray : i);
      dir[0] = dp[0]; dir[1] = dp[1]; dir[2] = dp[2];
    }
    float feat[kModMaxDim];
    const int aid = neurad_encode_point(fg, frames[warp], A.n_actors, g, feat, a.dirs ? dir : nullptr);
    if (a.features)
      for (int k = 0; k < D; ++k) a.features[i * D + k] = feat[k];
    if (a.density) {
      float acc = 0.f;
      for (int k = 0; k < D; ++k) acc = fmaf(feat[k], __ldg(fg.decoder + k), acc);
      a.density[i] = expf(acc);
    }
    if (a.dirs_out) {
      a.dirs_out[3 * i] = dir[0];
      a.dirs_out[3 * i + 1] = dir[1];
      a.dirs_out[3 * i + 2] = dir[2];
    }
    if (a.actor_id) a.actor_id[i] = aid;
  }
}

// NeuRADField.forward between its two MLPs (fields/neurad_field.py:139-141): geo_out [P, G+1] (sdf | geo_embedding)
// and directions [P,3] -> the feature MLP's input [P, G+16] = [geo_embedding | SH4((d + 1) / 2)]
// (get_normalized_directions base_field.py:136-142, SHEncoding encodings.py:797-805).
__global__ void field_mid_kernel(const float* __restrict__ geo_out, const float* __restrict__ dirs, int64_t n, int G,
                                 float* __restrict__ x2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int W = G + kSh;
  for (int k = 0; k < G; ++k) x2[i * W + k] = geo_out[i * (G + 1) + 1 + k];
  float c[16];
  sh4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], c);
#pragma unroll
  for (int k = 0; k < 16; ++k) x2[i * W + G + k] = c[k];
}

// ... and after them (neurad_field.py:141-149): feature = geo_embedding + mlp_feature(...); sdf = geo_out[0];
// alpha = SigmoidDensity(sdf) = sigmoid(-sdf * (|beta| + 1e-4)) (model_components/utils.py:29-41; `beta` here is the
// already offset value the context holds).
__global__ void field_tail_kernel(const float* __restrict__ geo_out, const float* __restrict__ mlp_out, int64_t n,
                                  int G, float beta, float* __restrict__ feature, float* __restrict__ sdf,
                                  float* __restrict__ alpha) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < G; ++k) feature[i * G + k] = geo_out[i * (G + 1) + 1 + k] + mlp_out[i * G + k];
  const float sd = geo_out[i * (G + 1)];
  if (sdf) sdf[i] = sd;
  if (alpha) alpha[i] = frcp(fadd(1.0f, expf(fmul(sd, beta))));
}

}  // namespace nff
