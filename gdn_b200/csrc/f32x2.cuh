// f32x2.cuh -- packed fp32 pair arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue two IEEE fp32 operations per
// instruction) and array-wise helpers for the per-channel chains of the D-wide passes: a lane owns DPL
// consecutive channels, so every elementwise step of the chain pairs up naturally.
#pragma once
namespace gdn {

__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// out = a * b + c, elementwise over N floats (pairs go through fma.rn.f32x2; an odd tail stays scalar)
template <int N>
__device__ __forceinline__ void vfma(float (&out)[N], const float (&a)[N], const float (&b)[N], const float (&c)[N]) {
#pragma unroll
    for (int q = 0; q + 1 < N; q += 2) {
        unsigned long long r;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a[q], a[q + 1])), "l"(pk2(b[q], b[q + 1])), "l"(pk2(c[q], c[q + 1])));
        upk2(r, out[q], out[q + 1]);
    }
    if (N & 1) out[N - 1] = fmaf(a[N - 1], b[N - 1], c[N - 1]);
}
template <int N>
__device__ __forceinline__ void vmul(float (&out)[N], const float (&a)[N], const float (&b)[N]) {
#pragma unroll
    for (int q = 0; q + 1 < N; q += 2) {
        unsigned long long r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a[q], a[q + 1])), "l"(pk2(b[q], b[q + 1])));
        upk2(r, out[q], out[q + 1]);
    }
    if (N & 1) out[N - 1] = a[N - 1] * b[N - 1];
}
template <int N>
__device__ __forceinline__ void vadd(float (&out)[N], const float (&a)[N], const float (&b)[N]) {
#pragma unroll
    for (int q = 0; q + 1 < N; q += 2) {
        unsigned long long r;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a[q], a[q + 1])), "l"(pk2(b[q], b[q + 1])));
        upk2(r, out[q], out[q + 1]);
    }
    if (N & 1) out[N - 1] = a[N - 1] + b[N - 1];
}
// out = a * s (scalar broadcast)
template <int N>
__device__ __forceinline__ void vscale(float (&out)[N], const float (&a)[N], float s) {
    const unsigned long long ss = pk2(s, s);
#pragma unroll
    for (int q = 0; q + 1 < N; q += 2) {
        unsigned long long r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a[q], a[q + 1])), "l"(ss));
        upk2(r, out[q], out[q + 1]);
    }
    if (N & 1) out[N - 1] = a[N - 1] * s;
}
// out = relu(a)
template <int N>
__device__ __forceinline__ void vrelu(float (&out)[N], const float (&a)[N]) {
#pragma unroll
    for (int q = 0; q < N; ++q) out[q] = fmaxf(a[q], 0.f);
}
// out = y > 0 ? v : 0
template <int N>
__device__ __forceinline__ void vgate(float (&out)[N], const float (&y)[N], const float (&v)[N]) {
#pragma unroll
    for (int q = 0; q < N; ++q) out[q] = y[q] > 0.f ? v[q] : 0.f;
}

}  // namespace gdn
