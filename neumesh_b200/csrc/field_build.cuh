// First-layer input construction shared by both MLP engines: neighbour gather + inverse-distance blend
// (neumesh.py:11-13 `interpolation`) and the positional encodings (models/base.py:52-70), written through a
// store functor so each engine can use its own shared-memory layout.
#pragma once
#include "field.cuh"

namespace nmb {

struct FieldTables {
  const float* fg;  // [V,32] sorted
  const float* fc;  // [V,32] sorted
};

// blend 8 consecutive features [q*8, q*8+8) of the 8 neighbours of point p:  sum_k table[slot_k] * w_k,
// products and sums individually rounded, k ascending (what torch's (features[idx] * w[...,None]).sum(-2) does)
__device__ __forceinline__ void blend8(const float* __restrict__ table, const FieldIn& in, int64_t p, int q,
                                       float (&acc)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    const int32_t s = in.slot[k * in.stride + p];
    const float w = in.w[k * in.stride + p];
    const float4 a = __ldg(reinterpret_cast<const float4*>(table + (int64_t)s * FEAT + q * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(table + (int64_t)s * FEAT + q * 8 + 4));
    acc[0] = __fadd_rn(acc[0], __fmul_rn(a.x, w));
    acc[1] = __fadd_rn(acc[1], __fmul_rn(a.y, w));
    acc[2] = __fadd_rn(acc[2], __fmul_rn(a.z, w));
    acc[3] = __fadd_rn(acc[3], __fmul_rn(a.w, w));
    acc[4] = __fadd_rn(acc[4], __fmul_rn(b.x, w));
    acc[5] = __fadd_rn(acc[5], __fmul_rn(b.y, w));
    acc[6] = __fadd_rn(acc[6], __fmul_rn(b.z, w));
    acc[7] = __fadd_rn(acc[7], __fmul_rn(b.w, w));
  }
}

// PE of 8 blended features: block b of width FEAT at column off + b*FEAT: [x, sin x, cos x, sin 2x, cos 2x, ...]
template <class St>
__device__ __forceinline__ void store_feat_pe(const float (&x)[8], int q, int off, int L, St&& st) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int f = q * 8 + j;
    st(off + f, x[j]);
    float fr = 1.f;
    for (int b = 0; b < L; ++b) {
      float s, c;
      sincosf(x[j] * fr, &s, &c);
      st(off + (1 + 2 * b) * FEAT + f, s);
      st(off + (2 + 2 * b) * FEAT + f, c);
      fr *= 2.f;
    }
  }
}

// PE of a scalar at columns [off, off + 1 + 2L)
template <class St>
__device__ __forceinline__ void store_scalar_pe(float x, int off, int L, St&& st) {
  st(off, x);
  float fr = 1.f;
  for (int b = 0; b < L; ++b) {
    float s, c;
    sincosf(x * fr, &s, &c);
    st(off + 1 + 2 * b, s);
    st(off + 2 + 2 * b, c);
    fr *= 2.f;
  }
}

// d/dx of store_scalar_pe (tangent seed of the forward-mode nabla)
template <class St>
__device__ __forceinline__ void store_scalar_pe_tangent(float x, int off, int L, St&& st) {
  st(off, 1.f);
  float fr = 1.f;
  for (int b = 0; b < L; ++b) {
    float s, c;
    sincosf(x * fr, &s, &c);
    st(off + 1 + 2 * b, fr * c);
    st(off + 2 + 2 * b, -fr * s);
    fr *= 2.f;
  }
}

// PE of a 3-vector: [v(3), sin v (3), cos v (3), sin 2v (3), ...]
template <class St>
__device__ __forceinline__ void store_vec3_pe(float x, float y, float z, int off, int L, St&& st) {
  st(off + 0, x);
  st(off + 1, y);
  st(off + 2, z);
  float fr = 1.f;
  for (int b = 0; b < L; ++b) {
    float s, c;
    sincosf(x * fr, &s, &c);
    st(off + 3 + 6 * b + 0, s);
    st(off + 3 + 6 * b + 3, c);
    sincosf(y * fr, &s, &c);
    st(off + 3 + 6 * b + 1, s);
    st(off + 3 + 6 * b + 4, c);
    sincosf(z * fr, &s, &c);
    st(off + 3 + 6 * b + 2, s);
    st(off + 3 + 6 * b + 5, c);
    fr *= 2.f;
  }
}

__device__ __forceinline__ void load_dir(const FieldIn& in, int64_t p, float& x, float& y, float& z) {
  const float* d = in.dirs ? (in.dirs + p * 3) : (in.rays_d + (p % in.R) * 3);
  x = d[0];
  y = d[1];
  z = d[2];
}

// nn.Softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100.
// log(1 + e) with fast intrinsics: absolute error <= ~1e-7 before the /100, i.e. <= 1e-9 on the result.
__device__ __forceinline__ float softplus100(float z) {
  const float a = 100.f * z;
  const float y = __logf(1.0f + __expf(a)) * 0.01f;
  return a > 20.f ? z : y;
}
// its derivative as autograd evaluates it: 1 above the threshold, e/(e+1) below
__device__ __forceinline__ float softplus100_grad(float z) {
  const float a = 100.f * z;
  const float e = __expf(a);
  return a > 20.f ? 1.f : __fdividef(e, e + 1.f);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace nmb
