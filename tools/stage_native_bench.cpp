// Host staging microbenchmark: float64 (pageable) -> float32 (destination), T threads, AVX-512 / scalar, NT stores or not.
// g++ -O2 -pthread tools/stage_native_bench.cpp -o /tmp/stage_native && /tmp/stage_native
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

__attribute__((target("avx512f,avx512dq"))) static void conv512(const double* s, float* d, size_t n, bool nt) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(d + i) & 63)) { d[i] = (float)s[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        __m256 a = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i));
        __m256 b = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i + 8));
        __m512 v = _mm512_insertf32x8(_mm512_castps256_ps512(a), b, 1);
        if (nt) _mm512_stream_ps(d + i, v); else _mm512_storeu_ps(d + i, v);
    }
    for (; i < n; ++i) d[i] = (float)s[i];
    if (nt) _mm_sfence();
}
static void conv_scalar(const double* s, float* d, size_t n) { for (size_t i = 0; i < n; ++i) d[i] = (float)s[i]; }

int main() {
    const size_t n = 64ull * 16384 * 16 + 64ull * 16384;
    double* src = (double*)aligned_alloc(64, n * 8);
    float* dst = (float*)aligned_alloc(64, n * 4);
    for (size_t i = 0; i < n; ++i) src[i] = (double)(i % 1000) * 1e-3;
    memset(dst, 0, n * 4);
    const bool has512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
    printf("avx512f+dq=%d hw threads=%u\n", (int)has512, std::thread::hardware_concurrency());
    for (int mode = 0; mode < 3; ++mode) {
        if (mode > 0 && !has512) break;
        for (int T : {1, 2, 4, 6, 8, 12, 16}) {
            double best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t)
                    th.emplace_back([&, t] {
                        size_t a = n * t / T, b = n * (t + 1) / T;
                        if (mode == 0) conv_scalar(src + a, dst + a, b - a); else conv512(src + a, dst + a, b - a, mode == 2);
                    });
                for (auto& x : th) x.join();
                double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                if (ms < best) best = ms;
            }
            printf("%-14s T=%2d  %.2f ms  (%.1f GB/s read+write)\n", mode == 0 ? "scalar(-O2)" : mode == 1 ? "avx512" : "avx512+nt", T, best, n * 12.0 / best / 1e6);
        }
    }
    return 0;
}
