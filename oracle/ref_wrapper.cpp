// ref_wrapper.cpp -- C ABI around the reference's OWN golden functions.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ntt_oracle.c header).
//
// This file is never compiled on its own.  oracle/Makefile streams
//   <cstdint>/<vector> includes  +  the golden functions modPow / make_roots /
//   ntt taken at build time from where they lie in
//   /root/reference/src/test.cpp (everything between `int32_t modPow(` and
//   `int main(`, i.e. src/test.cpp:15-60)  +  this wrapper
// into one g++ translation unit on stdin and writes only
// oracle/_ref/libntt_ref.so.  No reference source is copied into the repo.
//
// The wrappers mirror how src/test.cpp:main drives the golden:
//   root[0] = 1; make_roots(N, root, p, g);          (src/test.cpp:137-139)
//   ntt(a_ref, N, root, p, test_stage);              (src/test.cpp:203-207)
#include <thread>

extern "C" {

__attribute__((visibility("default"))) int32_t ref_modpow(int32_t x, int32_t n, int32_t mod) {
    return modPow(x, n, mod);
}

// verbatim make_roots; only valid for p <= 65536 (uint32 product) and
// g^2 < 2^31 etc. (int32 modPow) -- the caller checks the domain.
__attribute__((visibility("default"))) void ref_make_roots(int32_t n, int32_t *roots_out,
                                                           int32_t p, int32_t g) {
    std::vector<int32_t> root(n);
    root[0] = 1;
    make_roots(n, root, p, g);
    for (int i = 0; i < n; i++) roots_out[i] = root[i];
}

// verbatim ntt(), in place on a caller buffer
__attribute__((visibility("default"))) void ref_ntt(int32_t *a, int32_t n, const int32_t *roots,
                                                    int32_t p, int32_t stage) {
    std::vector<int32_t> va(a, a + n);
    std::vector<int32_t> vr(roots, roots + n);
    ntt(va, n, vr, p, stage);
    for (int i = 0; i < n; i++) a[i] = va[i];
}

// Batched verbatim ntt() over host threads: the CPU baseline ("kind":
// "reference") that bench.py times on the GPU box's own cores.  The vectors
// are built outside the per-polynomial loop so the timed work is the golden
// butterfly network itself.
__attribute__((visibility("default"))) void ref_ntt_batch(int32_t *a, int32_t n, int64_t batch,
                                                          const int32_t *roots, int32_t p,
                                                          int32_t stage, int32_t nthreads) {
    if (nthreads < 1) nthreads = 1;
    auto work = [=](int64_t begin, int64_t end) {
        std::vector<int32_t> vr(roots, roots + n);
        std::vector<int32_t> va(n);
        for (int64_t b = begin; b < end; b++) {
            int32_t *poly = a + b * (int64_t) n;
            va.assign(poly, poly + n);
            ntt(va, n, vr, p, stage);
            for (int i = 0; i < n; i++) poly[i] = va[i];
        }
    };
    if (nthreads == 1) {
        work(0, batch);
        return;
    }
    std::vector<std::thread> pool;
    for (int k = 0; k < nthreads; k++) {
        pool.emplace_back(work, batch * k / nthreads, batch * (k + 1) / nthreads);
    }
    for (auto &th : pool) th.join();
}

}  // extern "C"
