/*
 * ntt_oracle.c -- CPU restatement of the ntt-aie golden model.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker or the reported CPU
 * baseline.  The product (libnttb200.so) never links or calls it.
 *
 * Parity status: PINNED.  oracle_ntt_gs / oracle_make_roots are checked
 *   (a) against the reference's own golden code compiled from where it lies
 *       (/root/reference/src/test.cpp:15-60 -> oracle/_ref/libntt_ref.so, see
 *       oracle/Makefile) on the reference's one test configuration
 *       (N=2048, p=3329, g=3, a[i]=i, src/test.cpp:66,76-77,141) and on
 *       seeded random inputs / 29-30 bit primes, and
 *   (b) against the committed fixtures under tests/golden/ that were generated
 *       from that library (tests/golden/make_golden.py).
 * The operators the reference does not have (CT network, pointwise product,
 * negacyclic product) are "parity unpinned upstream"; they are restated from
 * Longa-Naehrig (2016) Alg. 1 and validated by GS_ref(CT(x)) == n*x and by the
 * O(N^2) schoolbook product in tests/.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* modPow -- src/test.cpp:15-25.  Same recursion (square first, then   */
/* recurse on n/2), restated with 64-bit intermediates: the reference   */
/* multiplies in int32 (x * x at :20,:22) and is therefore only valid   */
/* for mod < 46341; on that domain the two agree (asserted in tests).   */
/* ------------------------------------------------------------------ */
ORACLE_API int32_t oracle_modpow(int32_t x, int32_t n, int32_t mod) {
    if (n == 0) {
        return 1;
    }
    int64_t sq = ((int64_t) x * x) % mod;
    if (n % 2 == 1) {
        return (int32_t) (((int64_t) x * oracle_modpow((int32_t) sq, n / 2, mod)) % mod);
    }
    return oracle_modpow((int32_t) sq, n / 2, mod);
}

/* ------------------------------------------------------------------ */
/* make_roots -- src/test.cpp:27-32.  w = g^((p-1)/n) with INTEGER     */
/* division, roots[i] = roots[i-1]*w mod p for i=1..n-1.  roots[0] is   */
/* the caller's (the reference sets it to 1 at :138).  Product widened  */
/* to 64 bit (the reference's uint32 product at :30 overflows for       */
/* p > 65536).                                                          */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_make_roots(int32_t n, int32_t *roots, int32_t p, int32_t g) {
    int32_t w = oracle_modpow(g, (p - 1) / n, p);
    for (int i = 1; i < n; i++) {
        roots[i] = (int32_t) (((uint64_t) (uint32_t) roots[i - 1] * (uint32_t) w) % (uint32_t) p);
    }
}

/* ------------------------------------------------------------------ */
/* ntt -- src/test.cpp:34-60.  The golden Gentleman-Sande network:      */
/* stride t grows 1 -> n/2, block count h = m/2 shrinks n/2 -> 1,       */
/* twiddle roots_rev[h + i] for block i, per butterfly                  */
/*   a[j]   = (v0 + v1) % p                                (:48)        */
/*   a[j+t] = ((v0 + p - v1) % p) * root % p  (u64 product) (:49-50)    */
/* and early return once stage index idx == stage (:55-58).  All        */
/* additions are int32 exactly as in the reference, so the valid domain */
/* is the reference's: 2p-1 <= INT32_MAX, i.e. p <= 2^30.               */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_ntt_gs(int32_t *a, int32_t n, const int32_t *roots_rev, int32_t p,
                              int32_t stage) {
    int32_t t = 1;
    int idx = 0;
    for (int m = n; m > 1; m >>= 1) {
        int32_t j1 = 0;
        int32_t h = m / 2;
        for (int i = 0; i < h; i++) {
            int32_t j2 = j1 + t - 1;
            int32_t root = roots_rev[h + i];
            for (int j = j1; j <= j2; j++) {
                int32_t v0 = a[j];
                int32_t v1 = a[j + t];
                a[j] = (v0 + v1) % p;
                a[j + t] = (int32_t) (((uint64_t) ((v0 + p - v1) % p) * (uint64_t) root) % (uint64_t) p);
            }
            j1 += 2 * t;
        }
        t <<= 1;
        if (idx == stage) {
            return;
        }
        idx += 1;
    }
}

/* ------------------------------------------------------------------ */
/* CT network -- NOT in the reference (SURVEY 8a "NEW operators").      */
/* Longa-Naehrig Alg. 1: stride t shrinks n/2 -> 1, block count m grows */
/* 1 -> n/2, twiddle table[m + i] (the same "h+i" index rule as the     */
/* golden at src/test.cpp:45), butterfly V = a[j+t]*S; a[j] = U+V;      */
/* a[j+t] = U-V.  It is the exact inverse partner of oracle_ntt_gs up   */
/* to the factor n when the two tables hold inverse entries.  `stage`   */
/* has the golden's early-exit meaning (src/test.cpp:55-58).            */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_ntt_ct(int32_t *a, int32_t n, const int32_t *table, int32_t p,
                              int32_t stage) {
    int32_t t = n;
    int idx = 0;
    for (int m = 1; m < n; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int32_t j1 = 2 * i * t;
            uint64_t s = (uint64_t) (uint32_t) table[m + i];
            for (int j = j1; j < j1 + t; j++) {
                int32_t u = a[j];
                int32_t v = (int32_t) (((uint64_t) (uint32_t) a[j + t] * s) % (uint64_t) p);
                a[j] = (u + v) % p;
                a[j + t] = (u + p - v) % p;
            }
        }
        if (idx == stage) {
            return;
        }
        idx += 1;
    }
}

/* ------------------------------------------------------------------ */
/* WIDE moduli, 2^30 < p < 2^31 -- OUTSIDE the reference's domain       */
/* (its int32 sums v0 + v1 and v0 + p - v1 at src/test.cpp:48-49        */
/* overflow there), so parity for this range is UNPINNED upstream.      */
/* Same loops as oracle_ntt_gs / oracle_ntt_ct with every sum widened   */
/* to 64 bit; the tests assert that the two forms agree wherever the    */
/* int32 form is defined (p <= 2^30) and check the wide range against   */
/* Python big-integer arithmetic and the schoolbook product.            */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_ntt_gs_wide(int32_t *a, int32_t n, const int32_t *roots_rev, int32_t p,
                                   int32_t stage) {
    const uint64_t q = (uint64_t) (uint32_t) p;
    int32_t t = 1;
    int idx = 0;
    for (int m = n; m > 1; m >>= 1) {
        int32_t j1 = 0;
        int32_t h = m / 2;
        for (int i = 0; i < h; i++) {
            int32_t j2 = j1 + t - 1;
            uint64_t root = (uint64_t) (uint32_t) roots_rev[h + i];
            for (int j = j1; j <= j2; j++) {
                uint64_t v0 = (uint64_t) (uint32_t) a[j];
                uint64_t v1 = (uint64_t) (uint32_t) a[j + t];
                a[j] = (int32_t) ((v0 + v1) % q);
                a[j + t] = (int32_t) ((((v0 + q - v1) % q) * root) % q);
            }
            j1 += 2 * t;
        }
        t <<= 1;
        if (idx == stage) {
            return;
        }
        idx += 1;
    }
}

ORACLE_API void oracle_ntt_ct_wide(int32_t *a, int32_t n, const int32_t *table, int32_t p,
                                   int32_t stage) {
    const uint64_t q = (uint64_t) (uint32_t) p;
    int32_t t = n;
    int idx = 0;
    for (int m = 1; m < n; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int32_t j1 = 2 * i * t;
            uint64_t s = (uint64_t) (uint32_t) table[m + i];
            for (int j = j1; j < j1 + t; j++) {
                uint64_t u = (uint64_t) (uint32_t) a[j];
                uint64_t v = ((uint64_t) (uint32_t) a[j + t] * s) % q;
                a[j] = (int32_t) ((u + v) % q);
                a[j + t] = (int32_t) ((u + q - v) % q);
            }
        }
        if (idx == stage) {
            return;
        }
        idx += 1;
    }
}

/* c[i] = a[i]*b[i] mod p ; c[i] = a[i]*s mod p  (new operators, see above) */
ORACLE_API void oracle_pointwise(const int32_t *a, const int32_t *b, int32_t *c, int64_t count,
                                 int32_t p) {
    for (int64_t i = 0; i < count; i++) {
        c[i] = (int32_t) (((uint64_t) (uint32_t) a[i] * (uint32_t) b[i]) % (uint32_t) p);
    }
}

ORACLE_API void oracle_scale(const int32_t *a, int32_t *c, int64_t count, int32_t s, int32_t p) {
    for (int64_t i = 0; i < count; i++) {
        c[i] = (int32_t) (((uint64_t) (uint32_t) a[i] * (uint32_t) s) % (uint32_t) p);
    }
}

/* O(n^2) negacyclic product c = a*b mod (x^n + 1, p): independent check */
ORACLE_API void oracle_negacyclic_schoolbook(const int32_t *a, const int32_t *b, int32_t *c,
                                             int32_t n, int32_t p) {
    for (int k = 0; k < n; k++) {
        uint64_t acc = 0;
        for (int i = 0; i < n; i++) {
            int j = k - i;
            uint64_t prod;
            if (j >= 0) {
                prod = ((uint64_t) (uint32_t) a[i] * (uint32_t) b[j]) % (uint32_t) p;
            } else {
                prod = ((uint64_t) (uint32_t) a[i] * (uint32_t) b[j + n]) % (uint32_t) p;
                prod = (prod == 0) ? 0 : (uint64_t) p - prod;
            }
            acc = (acc + prod) % (uint32_t) p;
        }
        c[k] = (int32_t) acc;
    }
}

/* ------------------------------------------------------------------ */
/* ans_order -- src/test.cpp:69-71 and :212-219: the AIE device leaves  */
/* its output permuted in 16 blocks of n/16;                            */
/*   answers[ans_order[i]*B + j] = golden[i*B + j].                     */
/* ------------------------------------------------------------------ */
static const int oracle_ans_order[16] = {0, 2, 1, 3, 8, 10, 9, 11, 4, 6, 5, 7, 12, 14, 13, 15};

ORACLE_API void oracle_ans_order_permute(const int32_t *golden, int32_t *answers, int32_t n) {
    int block_size = n / 16;
    for (int i = 0; i < 16; i++) {
        int base_i = oracle_ans_order[i] * block_size;
        for (int j = 0; j < block_size; j++) {
            answers[base_i + j] = golden[i * block_size + j];
        }
    }
}

/* ------------------------------------------------------------------ */
/* Scalar modular helpers of the device kernel: src/aie_core.cc:11-39.  */
/* barrett_2k(a,b,q,w,u): w = ceil(log2 q), u = floor(2^(2w)/q)         */
/* (src/aie2.py:18-19).  Used by tests to document that the device      */
/* arithmetic and the golden `%` agree on the reference's domain.       */
/* ------------------------------------------------------------------ */
ORACLE_API int32_t oracle_modadd(int32_t a, int32_t b, int32_t q) {
    int ret = a + b;
    return ret >= q ? ret - q : ret;
}

ORACLE_API int32_t oracle_modsub(int32_t a, int32_t b, int32_t q) {
    int ret = a + q - b;
    return ret >= q ? ret - q : ret;
}

ORACLE_API int32_t oracle_barrett_2k(int32_t a, int32_t b, int32_t q, int32_t w, int32_t u) {
    int64_t t = (int64_t) a * (int64_t) b;
    int64_t x_1 = t >> (w - 2);
    int64_t x_2 = (int64_t) u * x_1;
    int64_t s = x_2 >> (w + 2);
    int64_t r = s * q;
    int64_t c = t - r;
    return (int32_t) (c >= q ? c - q : c);
}

/* ------------------------------------------------------------------ */
/* Tables for the polynomial multiplier (new operators).                */
/* bit-reversed psi powers in the golden's table[h+i] index rule:       */
/*   table[k] = base^bitrev_logn(k), k = 1..n-1, table[0] = 1 (unused). */
/* With base = psi (a primitive 2n-th root) the CT network is the       */
/* forward negacyclic NTT; with base = psi^-1 the golden GS network     */
/* (src/test.cpp:34-60) is the un-scaled inverse.                       */
/* ------------------------------------------------------------------ */
static uint32_t bitrev_u32(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | ((x >> i) & 1u);
    }
    return r;
}

static uint64_t powmod_u64(uint64_t b, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    b %= m;
    while (e) {
        if (e & 1) {
            r = (unsigned __int128) r * b % m;
        }
        b = (unsigned __int128) b * b % m;
        e >>= 1;
    }
    return r;
}

ORACLE_API int32_t oracle_powmod(int32_t b, int64_t e, int32_t m) {
    return (int32_t) powmod_u64((uint64_t) b, (uint64_t) e, (uint64_t) m);
}

ORACLE_API void oracle_make_bitrev_table(int32_t n, int32_t *table, int32_t p, int32_t base) {
    int logn = 0;
    while ((1 << logn) < n) {
        logn++;
    }
    /* natural powers first, then scatter */
    int32_t *pw = (int32_t *) malloc(sizeof(int32_t) * (size_t) n);
    pw[0] = 1;
    for (int i = 1; i < n; i++) {
        pw[i] = (int32_t) (((uint64_t) (uint32_t) pw[i - 1] * (uint32_t) base) % (uint32_t) p);
    }
    for (int k = 0; k < n; k++) {
        table[k] = pw[bitrev_u32((uint32_t) k, logn)];
    }
    free(pw);
}

/* ------------------------------------------------------------------ */
/* Batched golden over host threads: the CPU baseline bench.py reports  */
/* when the verbatim reference library is unavailable ("kind": "port"). */
/* Independent polynomials a[b*n .. b*n+n) split contiguously over      */
/* nthreads pthreads, each running oracle_ntt_gs (src/test.cpp:34-60).  */
/* ------------------------------------------------------------------ */
typedef struct {
    int32_t *a;
    int32_t n;
    const int32_t *roots;
    int32_t p;
    int32_t stage;
    int64_t begin, end;
} oracle_job;

static void *oracle_worker(void *arg) {
    oracle_job *job = (oracle_job *) arg;
    for (int64_t b = job->begin; b < job->end; b++) {
        oracle_ntt_gs(job->a + b * job->n, job->n, job->roots, job->p, job->stage);
    }
    return NULL;
}

ORACLE_API void oracle_ntt_gs_batch(int32_t *a, int32_t n, int64_t batch, const int32_t *roots,
                                    int32_t p, int32_t stage, int32_t nthreads) {
    if (nthreads <= 1) {
        oracle_job job = {a, n, roots, p, stage, 0, batch};
        oracle_worker(&job);
        return;
    }
    pthread_t *tid = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) nthreads);
    oracle_job *jobs = (oracle_job *) malloc(sizeof(oracle_job) * (size_t) nthreads);
    for (int k = 0; k < nthreads; k++) {
        jobs[k].a = a;
        jobs[k].n = n;
        jobs[k].roots = roots;
        jobs[k].p = p;
        jobs[k].stage = stage;
        jobs[k].begin = batch * k / nthreads;
        jobs[k].end = batch * (k + 1) / nthreads;
        pthread_create(&tid[k], NULL, oracle_worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) {
        pthread_join(tid[k], NULL);
    }
    free(jobs);
    free(tid);
}
