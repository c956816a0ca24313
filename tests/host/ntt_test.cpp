// ntt_test.cpp -- C++ host harness over the C ABI: the successor of the reference's
// host program (reference src/test.cpp:62-248) with XRT replaced by libnttb200.so.
//
// Same flow as the reference main():
//   fill a[i] = i and root = make_roots(N, p, g)          (src/test.cpp:137-144)
//   10 timed runs, microseconds printed per run            (src/test.cpp:157-175)
//   one verification run                                   (src/test.cpp:181-190)
//   CPU golden ntt(a_ref, N, root, p, test_stage)          (src/test.cpp:203-207)
//   optional ans_order block permutation of the golden     (src/test.cpp:212-219)
//   exact compare, mismatch count, PASS!/FAIL., exit 0/1   (src/test.cpp:221-247)
//
// This is TEST code: it links the oracle (oracle/ntt_oracle.c) as the CPU golden.
// The product library is reached only through include/nttb200.h.
//
//   ntt_test [--logn 11] [--p 3329] [--g 3] [--batch 1] [--stage -1] [--aie-order]
//            [--iters 10] [--random]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nttb200.h"

extern "C" {
void oracle_make_roots(int32_t n, int32_t *roots, int32_t p, int32_t g);
void oracle_ntt_gs(int32_t *a, int32_t n, const int32_t *roots_rev, int32_t p, int32_t stage);
void oracle_ans_order_permute(const int32_t *golden, int32_t *answers, int32_t n);
}

int main(int argc, const char *argv[]) {
    // ============================ Test Parameters (src/test.cpp:66-80 defaults)
    int logn = 11, stage = -1, iters = 10;
    int32_t p = 3329, g = 3;
    size_t batch = 1;
    bool aie_order = false, random_input = false;
    for (int i = 1; i < argc; i++) {
        auto next = [&](const char *name) -> long long {
            if (i + 1 >= argc) {
                fprintf(stderr, "missing value for %s\n", name);
                exit(2);
            }
            return atoll(argv[++i]);
        };
        if (!strcmp(argv[i], "--logn")) logn = (int) next("--logn");
        else if (!strcmp(argv[i], "--p")) p = (int32_t) next("--p");
        else if (!strcmp(argv[i], "--g")) g = (int32_t) next("--g");
        else if (!strcmp(argv[i], "--batch")) batch = (size_t) next("--batch");
        else if (!strcmp(argv[i], "--stage")) stage = (int) next("--stage");
        else if (!strcmp(argv[i], "--iters")) iters = (int) next("--iters");
        else if (!strcmp(argv[i], "--aie-order")) aie_order = true;
        else if (!strcmp(argv[i], "--random")) random_input = true;
        else {
            fprintf(stderr, "unknown option %s\n", argv[i]);
            return 2;
        }
    }
    const int32_t n = 1 << logn;
    const int test_stage = stage < 0 ? logn - 1 : stage;  // src/test.cpp:67

    // ============================ Buffers (successors of bo_inA / bo_root / bo_outC)
    std::vector<int32_t> root(n), in(batch * n), out(batch * n, 0);
    if (nttb200_make_roots(n, root.data(), p, g) != NTTB200_OK) {
        printf("make_roots failed\n");
        return 1;
    }
    uint64_t x = 0x5EED0001ull;
    for (size_t i = 0; i < in.size(); i++) {
        if (random_input) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            in[i] = (int32_t) ((x >> 33) % (uint64_t) p);
        } else {
            in[i] = (int32_t) (i % n);  // a[i] = i (src/test.cpp:141), NOT reduced: the library, like
                                        // the golden's `%`, reduces on first touch
        }
    }

    nttb200_plan *plan = nullptr;
    int rc = nttb200_plan_create(&plan, 0, (uint32_t) logn, (uint32_t) p, root.data(),
                                 aie_order ? NTTB200_ORDER_AIE_DEVICE : NTTB200_ORDER_GOLDEN);
    if (rc != NTTB200_OK) {
        printf("plan_create failed: %s [%s]\n", nttb200_strerror(rc), nttb200_last_error());
        return 1;
    }

    // ============================ Execute the kernel `iters` times (src/test.cpp:157-175)
    printf("Running Kernel.\n");
    for (int i = 0; i < iters; i++) {
        auto start = std::chrono::high_resolution_clock::now();
        rc = nttb200_gs_host(plan, in.data(), out.data(), batch, test_stage);
        auto stop = std::chrono::high_resolution_clock::now();
        if (rc != NTTB200_OK) {
            printf("kernel did not complete. returned status: %s [%s]\n", nttb200_strerror(rc),
                   nttb200_last_error());
            return 1;
        }
        printf("%.1f\n", std::chrono::duration<double, std::micro>(stop - start).count());
    }

    // ============================ Execute the kernel for test (src/test.cpp:181-190)
    std::fill(out.begin(), out.end(), 0);
    rc = nttb200_gs_host(plan, in.data(), out.data(), batch, test_stage);
    if (rc != NTTB200_OK) {
        printf("kernel did not complete. returned status: %s\n", nttb200_strerror(rc));
        return 1;
    }
    printf("=================================\n");
    printf("kernel path: %s, launches so far: %llu\n", nttb200_plan_last_path(plan),
           (unsigned long long) nttb200_kernel_launches());

    // ============================ CPU Reference (src/test.cpp:203-207)
    std::vector<int32_t> root_ref(n);
    root_ref[0] = 1;
    oracle_make_roots(n, root_ref.data(), p, g);
    std::vector<int32_t> a_ref(in), answers(batch * n);
    for (size_t b = 0; b < batch; b++) {
        oracle_ntt_gs(a_ref.data() + b * n, n, root_ref.data(), p, test_stage);
        if (aie_order && test_stage >= logn - 1) {
            oracle_ans_order_permute(a_ref.data() + b * n, answers.data() + b * n, n);
        } else {
            memcpy(answers.data() + b * n, a_ref.data() + b * n, sizeof(int32_t) * n);
        }
    }

    // ============================ Verify Results (src/test.cpp:221-247)
    size_t errors = 0;
    printf("Verifying results\n");
    for (size_t i = 0; i < answers.size(); i++) {
        if (out[i] != answers[i]) errors++;
    }
    for (int32_t i = 0; i < n; i++) {
        if (root[i] != root_ref[i]) errors++;
    }
    printf("  logN: %d\n", logn);
    printf("  p: %d\n", p);
    nttb200_plan_destroy(plan);
    if (!errors) {
        printf("  PASS!\n");
        return 0;
    }
    printf("  mismatches: %zu\n", errors);
    printf("  FAIL.\n\n");
    return 1;
}
