// host_stage.cpp -- host side of the feed: the reference's loader yields pageable
// float64 batches (datasets/TimeDataset.py:64-73; train.py:63-66 casts them with `.float()` on one thread and copies
// them with a blocking `.to(device)`).  Here the cast IS the staging copy into pinned memory: one pass over the data,
// split over a persistent pool of plain threads (independent of OMP_NUM_THREADS, which torch.distributed.run pins to 1
// per rank, and of the Python interpreter lock), AVX-512 convert + non-temporal stores where the CPU has them (the
// pinned destination is read next by the copy engine, not by a core: no reason to pull it through the caches).
// Host code only: compiled with g++, linked into libgdn_b200.so.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {

__attribute__((target("avx512f,avx512dq"))) void convert_avx512(const double* s, float* d, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(d + i) & 63)) { d[i] = (float)s[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        const __m256 a = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i));
        const __m256 b = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i + 8));
        _mm512_stream_ps(d + i, _mm512_insertf32x8(_mm512_castps256_ps512(a), b, 1));
    }
    for (; i < n; ++i) d[i] = (float)s[i];
    _mm_sfence();
}

void convert_plain(const double* s, float* d, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
}

void convert(const double* s, float* d, size_t n) {
    static const bool wide = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
    if (wide) convert_avx512(s, d, n);
    else convert_plain(s, d, n);
}

// caller + (parts - 1) pool threads each convert one contiguous slice (boundaries on 16-element multiples)
struct Pool {
    std::mutex call;                    // one staging call at a time
    std::mutex m;
    std::condition_variable start, done;
    std::vector<std::thread> workers;
    unsigned long long gen = 0;
    int pending = 0, parts = 1;
    const double* src = nullptr;
    float* dst = nullptr;
    size_t n = 0;
    pid_t pid = 0;

    static void slice(size_t n, int parts, int k, size_t* a, size_t* b) {
        const size_t blocks = (n + 15) / 16;
        *a = (blocks * (size_t)k / (size_t)parts) * 16;
        *b = (blocks * (size_t)(k + 1) / (size_t)parts) * 16;
        if (*a > n) *a = n;
        if (*b > n) *b = n;
    }

    void work(int id) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> l(m);
            start.wait(l, [&] { return gen != seen; });
            seen = gen;
            const int p = parts;
            const double* s = src;
            float* d = dst;
            const size_t total = n;
            l.unlock();
            if (id + 1 < p) {
                size_t a, b;
                slice(total, p, id + 1, &a, &b);
                convert(s + a, d + a, b - a);
            }
            l.lock();
            if (--pending == 0) done.notify_one();
        }
    }

    void grow(int want) {               // under `call`; a host that refuses more threads just gets fewer slices
        try {
            while ((int)workers.size() < want) {
                const int id = (int)workers.size();
                workers.emplace_back([this, id] { work(id); });
                workers.back().detach();
            }
        } catch (...) {
        }
    }

    void run(const double* s, float* d, size_t total, int threads) {
        std::lock_guard<std::mutex> g(call);
        grow(threads - 1);
        if (threads > (int)workers.size() + 1) threads = (int)workers.size() + 1;
        {
            std::lock_guard<std::mutex> l(m);
            src = s; dst = d; n = total; parts = threads;
            pending = (int)workers.size();
            ++gen;
        }
        start.notify_all();
        size_t a, b;
        slice(total, threads, 0, &a, &b);
        convert(s + a, d + a, b - a);
        std::unique_lock<std::mutex> l(m);
        done.wait(l, [&] { return pending == 0; });
    }
};

Pool* pool() {
    static Pool* p = nullptr;
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    if (p == nullptr || p->pid != getpid()) {          // a forked child starts with no threads: fresh pool (the old one leaks)
        p = new Pool();
        p->pid = getpid();
    }
    return p;
}

}  // namespace

extern "C" int gdn_stage_f64_to_f32(const double* src, float* dst, size_t n, int threads) {
    if ((src == nullptr || dst == nullptr) && n > 0) return -1;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n < ((size_t)1 << 16) || threads == 1) {       // small: not worth a hand-off
        convert(src, dst, n);
        return 0;
    }
    pool()->run(src, dst, n, threads);
    return 0;
}
