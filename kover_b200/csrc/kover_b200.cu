// libkoverb200 -- Kover's rule-scoring hot path as hand-written CUDA for B200 (sm_100a).
//
// What the kernels compute and which reference lines they restate (paths relative to
// /root/reference/core/kover):
//   k_scan_tma     rules.py:201-267 (sum_rows for two row masks in ONE pass over the packed matrix:
//                  popcount(word & mask), popcount.pyx:76-95) fused with the SCM utility and the
//                  per-utility-block maximum of scm.py:262-270; TMA bulk copies + mbarrier ring
//   k_scm_select   scm.py:270-286 (sequential merge of the block maxima), on the device
//   k_scm_ties     scm.py:273 (isclose tie set against the block maximum)
//   k_count_tma    rules.py:201-267 for up to 16 masks per pass (CV folds / CART levels / batched sum_rows)
//   k_gini         cart.py:85-161 (_gini_impurity / _gini_rule_score), min + exact ties cart.py:238
//   k_gather_cols  rules.py:164 (dataset[:, cols])
//   k_popc_inplace popcount.pyx:31-50, 76-95
// Layout in HBM: uint64 [W, ld] row-major, ld = n_cols rounded up to 1024 words so that every
// tile row (512 or 1024 columns) is a 4 / 8 KB aligned segment; padding columns are zero and masked out of every
// reduction.  The path is a bitwise reduction bound by HBM bandwidth: no tensor cores.
//
// No CPU fallback lives here: every entry point either runs on the GPU or returns an error.
#include "kover_b200.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>   // only to order LARGE tie lists (not on the scan path)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                     __FILE__, __LINE__);                                                          \
            return fail(e__ == cudaErrorMemoryAllocation ? KV_E_NOMEM : KV_E_CUDA, b__);           \
        }                                                                                          \
    } while (0)

// Wait for a stream by polling: the per-step calls end in a microseconds-long tail, where the wake-up
// latency of a blocking cudaStreamSynchronize would be a visible fraction of the step.
static inline cudaError_t stream_sync_spin(cudaStream_t st) {
    cudaError_t e;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    return e;
}

#define KV_TRY(call)                                                                               \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != KV_OK) return r__;                                                              \
    } while (0)

extern "C" const char *kv_last_error(void) { return g_err.c_str(); }
extern "C" int kv_abi_version(void) { return 1; }

// Page-locked host memory for the CALLER's result arrays: a D2H copy into pageable memory runs at 4-6 GB/s (driver
// staging + first-touch page faults), into pinned memory at the PCIe rate.  The Python layer keeps a pool of these
// and wraps them as numpy arrays (sum_rows returns 2K counts: 40 MB per call at config 2).
extern "C" int kv_host_alloc(int64_t bytes, void **out) {
    if (!out || bytes <= 0) return fail(KV_E_INVALID, "kv_host_alloc: bad arguments");
    *out = nullptr;
    if (cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return fail(KV_E_NOMEM, "kv_host_alloc: cannot pin host memory");
    }
    return KV_OK;
}
extern "C" int kv_host_free(void *p) {
    if (p) cudaFreeHost(p);
    return KV_OK;
}

extern "C" int kv_device_count(int *n) {
    if (!n) return fail(KV_E_INVALID, "kv_device_count: null pointer");
    CU(cudaGetDeviceCount(n));
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------
static constexpr int TILE_COLS = 512;    // columns per tile of the direct-load kernels: 256 threads x one 128-bit load
static constexpr int LD_ALIGN = 1024;    // row pitch granularity: a multiple of every tile width (512 / 1024 columns)
static constexpr int MAX_CLASSES = 64;

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
// Streaming 128-bit load: read-only path, no L1 allocation (every matrix word is used exactly once
// per scan).
__device__ __forceinline__ ulonglong2 ldg_stream(const uint64_t *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}

// Order-preserving map double -> uint64 so that atomicMax / atomicMin on integers order doubles.
// The smallest image of any double (that of -inf) is 0x000F..F > 0, so a zeroed slot decodes as
// "nothing seen" = -inf (max) and an all-ones slot as +inf (min).
__device__ __host__ __forceinline__ unsigned long long enc_f64(double d) {
    unsigned long long b;
#ifdef __CUDA_ARCH__
    b = (unsigned long long)__double_as_longlong(d);
#else
    memcpy(&b, &d, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
static inline double dec_f64_max(unsigned long long v) {
    if (v == 0ull) return -std::numeric_limits<double>::infinity();
    unsigned long long b = (v & 0x8000000000000000ull) ? (v & 0x7FFFFFFFFFFFFFFFull) : ~v;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline double dec_f64_min(unsigned long long v) {
    if (v == ~0ull) return std::numeric_limits<double>::infinity();
    return dec_f64_max(v);
}

// scm.py:263-264 -- float64(neg_cover) - float64(p) * float64(pos_err): one rounded multiply, one
// rounded subtract.  A fused multiply-add changes results, hence the _rn
// intrinsics, which the compiler never contracts.
__device__ __forceinline__ double scm_utility(uint32_t neg_cover, uint32_t pos_err, double p) {
    return __dsub_rn((double)neg_cover, __dmul_rn(p, (double)pos_err));
}

// floor(r / d) for 0 <= r < 2^52 using a double reciprocal and one correction step.
__device__ __forceinline__ long long fast_div(long long r, long long d, double inv) {
    long long q = (long long)((double)r * inv);
    long long rem = r - q * d;
    if (rem < 0) --q;
    else if (rem >= d) ++q;
    return q;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// k_scan_tma: the hot kernel
// ------------------------------------------------------------------------------------------------
// One pass over the shard for a pair of row masks.  Warp-specialised, persistent, one CTA per SM:
//   * producer warp: one elected lane walks the CTA's tiles x active packed rows and issues one TMA
//     bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier) per (row, tile) segment
//     of CW*512 contiguous bytes into a ring of S stages x R rows;
//   * CW consumer warps: wait for a stage, read their two 64-bit columns of each row from shared
//     memory with one 128-bit load, AND with the row's mask words (staged in shared memory once per
//     CTA) and popcount; release the stage; after the last row of a tile run the epilogue.
// The ring keeps ~128 KB of reads in flight per SM without holding registers, which the sweep in
// tools/scan_sweep.cu measured at 96-98 % of the copy-bandwidth peak against 84-95 % for direct
// loads (profiles/README.md).
// up to this many active packed rows (4 096 genomes) ride in the kernel parameter block: no H2D copy per scan
static constexpr int SCAN_INLINE_ROWS = 64;

// Fused tail of an SCM scan (single-shard step in ONE launch): after its last tile every CTA arrives at a grid-wide
// barrier, replays the block merge on the complete block maxima, collects the ties of ITS OWN tiles straight into
// mapped page-locked host memory, and the last CTA to finish publishes {best utility, tie count, epoch flag} there.
struct TieRec {
    long long idx;
    uint32_t pe, nc;
};
struct FusedTail {
    int mode;                                   // 0 = off, 1 = exact ties
    unsigned long long *barrier, *done;         // monotonic device counters (never reset)
    unsigned long long barrier_target, done_target;
    unsigned long long *zero_next;              // the block-maximum accumulators of the NEXT fused scan (other parity)
    long long n_blocks;
    double *bm, *tol;                           // block tables in d_misc: what k_scm_select leaves (overflow path)
    uint8_t *sel;
    unsigned long long *tie_count;              // device tie counter
    TieRec *out;                                // tie records (mapped host memory)
    long long cap;
    volatile unsigned long long *host;          // mapped host header: [0] best utility bits, [1] tie count, [2] epoch,
    unsigned long long epoch;                   //                     [3] status (0 ok, 1 peer timeout, 2 overflow)
    // mode 2 (one process per GPU): candidates -> packet, the last CTA publishes it into every peer's mailbox over
    // NVLink, waits for theirs and merges all of them (block maxima, block merge, exact filter) on the device
    const struct P2PDev *p2p;
    unsigned long long p2p_seq, timeout_ns;
    char *packet;                               // local packet: block maxima f64 | PacketHeader | TieRec[packet_cap]
    long long packet_cap;
    double *host_bm;                            // mapped host: merged block maxima / selected flags (fallback inputs)
    uint8_t *host_sel;
};
static constexpr int P2P_MAX_WORLD = 64;
struct P2PDev {                                 // device-resident, written once by kv_p2p_connect
    char *peer_box[P2P_MAX_WORLD];              // mailbox base of every rank, mapped into this process
    unsigned long long *peer_flags[P2P_MAX_WORLD];
    int rank, world;
    size_t slot_bytes;
};

struct Scan2Params {
    const uint64_t *mat;
    long long ld, n_cols, col0, k_total;
    // active packed rows (mask word != 0, rules.py:244): rows with only mask 0, then only mask 1, then both
    const uint32_t *rec_row;
    const uint64_t *rec_m0, *rec_m1;
    int n_act;
    int inline_rec;                   // 1: the records travel in the kernel parameters (n_act <= SCAN_INLINE_ROWS)
    int l2_hints;                     // bit 0: evict-first matrix loads, bit 1: evict-last count stores
    int csa;                          // carry-save popcounts over groups of four rows (KOVER_B200_SCAN_CSA, default on)
    uint64_t inl_m0[SCAN_INLINE_ROWS], inl_m1[SCAN_INLINE_ROWS];
    uint32_t inl_row[SCAN_INLINE_ROWS];
    uint32_t *cnt0, *cnt1;            // [ld] each; cnt1 may be null (single-mask scan)
    // SCM epilogue
    double p;
    uint32_t n_pos, n_neg;
    long long ubs;
    const uint32_t *black_pres, *black_abs;   // bitmaps over local columns or null
    unsigned long long *blockmax;             // enc_f64 per utility block
    unsigned long long *segmax;               // enc_f64 per 64-column segment: [2][ld / 64] (presence, absence)
    // Gini epilogue (two classes = the two masks, in the caller's dict order): cart.py:71-77, 99-110
    double g_prior[2], g_ntot[2], g_nnode[2];
    unsigned long long *gmin;                 // enc_f64 of the minimum split score; segmax[0..ld/64) holds segment minima
    // EPI_SCM_STREAM: candidate records per CTA, cand[blockIdx.x * cand_cap ...], and how many each CTA produced
    TieRec *cand;
    unsigned *cand_n;
    int cand_cap;
    FusedTail tail;
};

// EPI_SCM_STREAM = EPI_SCM without a single store to HBM inside the tile loop: no count arrays, no segment maxima.
// A rule can only survive the merge if u >= best - 5 T(best) (the rank-local superset criterion of DESIGN section 7,
// x - 5 T(x) increasing), and every running maximum is a lower bound of `best`, so each warp compares its utilities with
// the threshold of the CTA's running maximum and appends the few that pass {rule, pos_err, neg_cover} to the CTA's
// candidate list; the LAST CTA of the grid filters all the lists exactly.  At W = 32 the 83 MB of count / segment stores
// cost 7 % of the scan (read / write turnarounds: profiles/r02_stream_vs_counts.log).
enum { EPI_COUNTS = 0, EPI_SCM = 1, EPI_GINI2 = 2, EPI_SCM_STREAM = 3 };

// _gini_rule_score for two classes (cart.py:112-161) in numpy's operation order; l0/l1 = examples of
// class 0/1 WITH the k-mer.  Identical arithmetic to gini_score<2> below (k_gini), kept separate so the
// scan kernel does not carry the 64-class parameter block.
__device__ __forceinline__ double gini2_child(double n0, double n1, const Scan2Params &P) {
    double q0 = __ddiv_rn(__dmul_rn(P.g_prior[0], n0), P.g_ntot[0]);
    double q1 = __ddiv_rn(__dmul_rn(P.g_prior[1], n1), P.g_ntot[1]);
    const double pt = __dadd_rn(q0, q1);
    q0 = __ddiv_rn(q0, pt);
    q1 = __ddiv_rn(q1, pt);
    const double div = __dadd_rn(__dmul_rn(q0, q1), __dmul_rn(q1, q0));
    return __dmul_rn(div, pt);
}
__device__ __forceinline__ double gini2_score(uint32_t c0, uint32_t c1, bool blacklisted, const Scan2Params &P) {
    const double l0 = (double)c0, l1 = (double)c1;
    const double r0 = P.g_nnode[0] - l0, r1 = P.g_nnode[1] - l1;
    if (blacklisted || l0 + l1 == 0.0 || r0 + r1 == 0.0) return INFINITY;       // cart.py:231, 158-159
    return __dadd_rn(gini2_child(l0, l1, P), gini2_child(r0, r1, P));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
    }
}
// TMA bulk copy global -> shared (1-D), completion signalled as transaction bytes on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Same with an L2 eviction policy: the matrix is streamed once per scan (evict-first), which leaves the
// L2 to the count / segment arrays the next kernels re-read.
__device__ __forceinline__ void tma_bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_u32x2_hint(uint32_t *ptr, uint32_t a, uint32_t b, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1, %2}, %3;" ::"l"(ptr), "r"(a), "r"(b), "l"(policy) : "memory");
}
__device__ __forceinline__ void consumer_bar(int n_threads) {
    asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// max over the warp of an enc_f64 image: two redux.sync on the 32-bit halves
__device__ __forceinline__ unsigned long long warp_max_enc(unsigned long long e) {
    const uint32_t hi = (uint32_t)(e >> 32), lo = (uint32_t)e;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

__device__ __forceinline__ unsigned long long warp_min_enc(unsigned long long e) {
    const uint32_t hi = (uint32_t)(e >> 32), lo = (uint32_t)e;
    const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}

// Running maximum of one half (presence / absence) of the rule space, kept per thread while the
// CTA's tiles stay inside one utility block; reduced and pushed with ONE atomic when they leave it.
struct RunMax {
    long long lo, hi, bid;
    double mx;
};

template <int CW>
__device__ __forceinline__ void runmax_flush(RunMax &r, unsigned long long *blockmax, unsigned long long *s_red) {
    // called by all consumer threads with CTA-uniform arguments
    if (r.bid >= 0) {
        const unsigned long long v = warp_max_enc(enc_f64(r.mx));
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        consumer_bar(CW * 32);
        if (lane == 0) s_red[wid] = v;
        consumer_bar(CW * 32);
        if (threadIdx.x == 0) {
            unsigned long long m = s_red[0];
            for (int w = 1; w < CW; ++w) m = s_red[w] > m ? s_red[w] : m;
            atomicMax(blockmax + r.bid, m);
        }
    }
}

// 3:2 carry-save compressor on 32 bit lanes: a + b + d = s + 2 c, two LOP3
__device__ __forceinline__ void csa32(uint32_t &s, uint32_t &c, uint32_t a, uint32_t b, uint32_t d) {
    s = a ^ b ^ d;
    c = (a & b) | (d & (a ^ b));
}
// popcount(a) + popcount(b) + popcount(c) + popcount(d) of four (already masked) 64-bit words with 4 POPC instead
// of 8: the eight 32-bit halves go through four carry-save compressors (8 LOP3 on the ALU pipe, which idles
// while the 16-lane POPC pipe is the bottleneck), leaving two words of weight 1, one of weight 2, one of weight 4.
__device__ __forceinline__ uint32_t popc4x64_csa(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint32_t s1, c1, s2, c2, s3, c3, s4, c4;
    csa32(s1, c1, (uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b);
    csa32(s2, c2, (uint32_t)(b >> 32), (uint32_t)c, (uint32_t)(c >> 32));
    csa32(s3, c3, s1, s2, (uint32_t)d);
    csa32(s4, c4, c1, c2, c3);
    return (__popc(s3) + __popc((uint32_t)(d >> 32))) + 2 * (__popc(s4) + 2 * __popc(c4));
}

// Four rows of a ring stage at once.  Whenever a mask is non-zero on all four rows (warp-uniform: the mask words are per
// row) its four masked words per column go through the carry-save compressors: with interleaved labels both masks touch
// every packed row, 4 POPC per streamed word, and the 16-lane POPC pipe -- not HBM -- bounded the scan (72 % busy at
// 6.0 TB/s); the compressors move half of that work to the idle ALU pipe.
__device__ __forceinline__ void accum_rows4(const unsigned char *p, int row_bytes, const uint64_t *m0, const uint64_t *m1,
                                            uint32_t &c0x, uint32_t &c0y, uint32_t &c1x, uint32_t &c1y) {
    const ulonglong2 v0 = *reinterpret_cast<const ulonglong2 *>(p);
    const ulonglong2 v1 = *reinterpret_cast<const ulonglong2 *>(p + row_bytes);
    const ulonglong2 v2 = *reinterpret_cast<const ulonglong2 *>(p + 2 * row_bytes);
    const ulonglong2 v3 = *reinterpret_cast<const ulonglong2 *>(p + 3 * row_bytes);
    {
        const uint64_t a0 = m0[0], a1 = m0[1], a2 = m0[2], a3 = m0[3];
        if (a0 && a1 && a2 && a3) {
            c0x += popc4x64_csa(v0.x & a0, v1.x & a1, v2.x & a2, v3.x & a3);
            c0y += popc4x64_csa(v0.y & a0, v1.y & a1, v2.y & a2, v3.y & a3);
        } else {
            if (a0) { c0x += __popcll(v0.x & a0); c0y += __popcll(v0.y & a0); }
            if (a1) { c0x += __popcll(v1.x & a1); c0y += __popcll(v1.y & a1); }
            if (a2) { c0x += __popcll(v2.x & a2); c0y += __popcll(v2.y & a2); }
            if (a3) { c0x += __popcll(v3.x & a3); c0y += __popcll(v3.y & a3); }
        }
    }
    {
        const uint64_t b0 = m1[0], b1 = m1[1], b2 = m1[2], b3 = m1[3];
        if (b0 && b1 && b2 && b3) {
            c1x += popc4x64_csa(v0.x & b0, v1.x & b1, v2.x & b2, v3.x & b3);
            c1y += popc4x64_csa(v0.y & b0, v1.y & b1, v2.y & b2, v3.y & b3);
        } else {
            if (b0) { c1x += __popcll(v0.x & b0); c1y += __popcll(v0.y & b0); }
            if (b1) { c1x += __popcll(v1.x & b1); c1y += __popcll(v1.y & b1); }
            if (b2) { c1x += __popcll(v2.x & b2); c1y += __popcll(v2.y & b2); }
            if (b3) { c1x += __popcll(v3.x & b3); c1y += __popcll(v3.y & b3); }
        }
    }
}

__device__ __forceinline__ void accum_word(const unsigned char *p, uint64_t a, uint64_t m, uint32_t &c0x, uint32_t &c0y,
                                           uint32_t &c1x, uint32_t &c1y) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    if (a) {                                   // warp-uniform: the mask words are per row
        c0x += __popcll(v.x & a);
        c0y += __popcll(v.y & a);
    }
    if (m) {
        c1x += __popcll(v.x & m);
        c1y += __popcll(v.y & m);
    }
}

template <int CW>
__device__ void scm_fused_tail(const Scan2Params &P, long long t0, long long t1, unsigned char *ring);
template <int CW>
__device__ void scm_stream_tail(const Scan2Params &P, unsigned char *ring, unsigned n_cand);
__device__ __forceinline__ double dec_f64_max_dev(unsigned long long v);
__device__ __forceinline__ double np_tol_dev(double b);

// GREC = true: the row records stay in global memory (read through L1) instead of being staged in shared memory --
// the fallback for masks with more active packed rows than fit beside the ring (~5 000 rows = 320 000 genomes); the
// reference has no such limit (rules.py:243-262 walks any number of rows).
template <int EPI, int CW, int R, int S, int MINB, bool GREC = false>
__global__ void __launch_bounds__((CW + 1) * 32, MINB) k_scan_tma(const Scan2Params P) {
    extern __shared__ __align__(128) unsigned char smem_tma[];
    constexpr int ROW_BYTES = CW * 32 * 16;
    constexpr int TC = CW * 64;                          // columns per tile
    __shared__ unsigned long long s_red[CW];
    __shared__ unsigned long long s_cmax;                // EPI_SCM_STREAM: enc_f64 of the CTA's running maximum
    __shared__ unsigned s_ncand;                         //                 candidates appended by this CTA
    const int n_act = P.n_act;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_tma);
    uint64_t *empty = full + S;
    uint64_t *sw_m0 = empty + S;
    uint64_t *sw_m1 = sw_m0 + n_act;
    uint32_t *sw_row = reinterpret_cast<uint32_t *>(sw_m1 + n_act);
    const uint64_t *s_m0 = GREC ? P.rec_m0 : sw_m0;
    const uint64_t *s_m1 = GREC ? P.rec_m1 : sw_m1;
    const uint32_t *s_row = GREC ? P.rec_row : sw_row;
    unsigned char *ring = smem_tma + ((size_t)(2 * S) * 8 + (GREC ? 0 : (size_t)n_act * 20) + 127) / 128 * 128;
    if (GREC) {
        // nothing to stage
    } else if (P.inline_rec) {
        for (int i = threadIdx.x; i < n_act; i += blockDim.x) {
            sw_m0[i] = P.inl_m0[i];
            sw_m1[i] = P.inl_m1[i];
            sw_row[i] = P.inl_row[i];
        }
    } else {
        for (int i = threadIdx.x; i < n_act; i += blockDim.x) {
            sw_m0[i] = P.rec_m0[i];
            sw_m1[i] = P.rec_m1[i];
            sw_row[i] = P.rec_row[i];
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cmax = 0ull;
        s_ncand = 0u;
    }
    __syncthreads();

    // contiguous tile range per CTA: a CTA's rules stay inside one or two utility blocks
    const long long n_tiles = P.ld / TC;
    const long long t0 = (n_tiles * (long long)blockIdx.x) / gridDim.x;
    const long long t1 = (n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
    const int nb = (n_act + R - 1) / R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == CW) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const bool hint = P.l2_hints & 1;
            const uint64_t pol = l2_policy_evict_first();
            for (long long tile = t0; tile < t1; ++tile) {
                const uint64_t *src = P.mat + tile * TC;
                for (int b = 0; b < nb; ++b) {
                    const int rows_here = min(R, n_act - b * R);
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, (uint32_t)rows_here * ROW_BYTES);
                    unsigned char *dst = ring + (size_t)s * R * ROW_BYTES;
                    if (hint) {
                        for (int r = 0; r < rows_here; ++r)
                            tma_bulk_g2s_hint(dst + r * ROW_BYTES, src + (long long)s_row[b * R + r] * P.ld, ROW_BYTES,
                                              full + s, pol);
                    } else if (rows_here == R) {
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            tma_bulk_g2s(dst + r * ROW_BYTES, src + (long long)s_row[b * R + r] * P.ld, ROW_BYTES, full + s);
                    } else {
                        for (int r = 0; r < rows_here; ++r)
                            tma_bulk_g2s(dst + r * ROW_BYTES, src + (long long)s_row[b * R + r] * P.ld, ROW_BYTES, full + s);
                    }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    RunMax run[2];
    run[0].lo = run[1].lo = 1; run[0].hi = run[1].hi = 0; run[0].bid = run[1].bid = -1;
    run[0].mx = run[1].mx = -INFINITY;
    const double neg_inf = -INFINITY;
    double gini_min = INFINITY;
    double thr = -INFINITY;                              // EPI_SCM_STREAM: candidate threshold (warp-uniform)
    TieRec *const my_cand = EPI == EPI_SCM_STREAM ? P.cand + (size_t)blockIdx.x * P.cand_cap : nullptr;
    const uint64_t pol_last = l2_policy_evict_last();
    int s = 0;
    uint32_t ph = 0;
    for (long long tile = t0; tile < t1; ++tile) {
        uint32_t c0x = 0, c0y = 0, c1x = 0, c1y = 0;
        for (int b = 0; b < nb; ++b) {
            const int rows_here = min(R, n_act - b * R);
            mbar_wait(full + s, ph);
            const unsigned char *st = ring + (size_t)s * R * ROW_BYTES + threadIdx.x * 16;
            if (R % 4 == 0 && rows_here == R && P.csa) {      // (row groups of four must not straddle ring stages)
#pragma unroll
                for (int r = 0; r + 3 < R; r += 4)
                    accum_rows4(st + r * ROW_BYTES, ROW_BYTES, s_m0 + b * R + r, s_m1 + b * R + r, c0x, c0y, c1x, c1y);
            } else if (rows_here == R) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    accum_word(st + r * ROW_BYTES, s_m0[b * R + r], s_m1[b * R + r], c0x, c0y, c1x, c1y);
            } else {
                for (int r = 0; r < rows_here; ++r)
                    accum_word(st + r * ROW_BYTES, s_m0[b * R + r], s_m1[b * R + r], c0x, c0y, c1x, c1y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if (++s == S) { s = 0; ph ^= 1; }
        }
        const long long col = tile * TC + 2 * (long long)threadIdx.x;
        if (EPI == EPI_SCM_STREAM) {
            // nothing is stored
        } else if (P.l2_hints & 2) {
            st_u32x2_hint(P.cnt0 + col, c0x, c0y, pol_last);
            if (P.cnt1) st_u32x2_hint(P.cnt1 + col, c1x, c1y, pol_last);
        } else {
            *reinterpret_cast<uint2 *>(P.cnt0 + col) = make_uint2(c0x, c0y);
            if (P.cnt1) *reinterpret_cast<uint2 *>(P.cnt1 + col) = make_uint2(c1x, c1y);
        }

        if (EPI == EPI_GINI2) {
            const bool vx = col < P.n_cols, vy = col + 1 < P.n_cols;
            uint32_t bp = 0;
            if (P.black_pres) bp = P.black_pres[col >> 5] >> (col & 31);
            const double sx = vx ? gini2_score(c0x, c1x, bp & 1u, P) : INFINITY;
            const double sy = vy ? gini2_score(c0y, c1y, bp & 2u, P) : INFINITY;
            const double sm = fmin(sx, sy);
            const unsigned long long seg = warp_min_enc(enc_f64(sm));
            if (lane == 0) P.segmax[col >> 6] = seg;
            gini_min = fmin(gini_min, sm);
        }
        if (EPI == EPI_SCM || EPI == EPI_SCM_STREAM) {
            const bool vx = col < P.n_cols, vy = col + 1 < P.n_cols;
            uint32_t bp = 0, ba = 0;
            if (P.black_pres) {
                bp = P.black_pres[col >> 5] >> (col & 31);
                ba = P.black_abs[col >> 5] >> (col & 31);
            }
            // presence rule k: covers the negatives WITHOUT the k-mer, errs on the positives
            // without it; absence rule K+k: the complements (rules.py:265, scm.py:245-251).
            // Padding columns and blacklisted rules (scm.py:267) score -inf.
            double ux[2], uy[2];
            ux[0] = (!vx || (bp & 1u)) ? neg_inf : scm_utility(P.n_neg - c1x, P.n_pos - c0x, P.p);
            uy[0] = (!vy || (bp & 2u)) ? neg_inf : scm_utility(P.n_neg - c1y, P.n_pos - c0y, P.p);
            ux[1] = (!vx || (ba & 1u)) ? neg_inf : scm_utility(c1x, c0x, P.p);
            uy[1] = (!vy || (ba & 2u)) ? neg_inf : scm_utility(c1y, c0y, P.p);
            const long long tile_c0 = tile * TC;
            const long long tile_n = min((long long)TC, P.n_cols - tile_c0);     // <= 0 for pure padding tiles
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double um = fmax(ux[h], uy[h]);
                if (EPI == EPI_SCM) {
                    // per-64-column segment maximum: lets k_scm_ties skip segments that cannot hold a tie
                    const unsigned long long seg = warp_max_enc(enc_f64(um));
                    if (lane == 0) P.segmax[(size_t)h * (P.ld >> 6) + (col >> 6)] = seg;
                }
                if (tile_n <= 0) continue;
                RunMax &r = run[h];
                const long long r0 = (long long)h * P.k_total + P.col0 + tile_c0;
                const long long r1 = r0 + tile_n - 1;
                bool fast = (r0 >= r.lo) && (r1 < r.hi);
                if (!fast) {                                   // CTA-uniform: depends on the tile only
                    runmax_flush<CW>(r, P.blockmax, s_red);
                    const long long b0 = r0 / P.ubs, b1 = r1 / P.ubs;
                    if (b0 == b1) {
                        r.bid = b0; r.lo = b0 * P.ubs; r.hi = r.lo + P.ubs; r.mx = neg_inf;
                        fast = true;
                    } else {
                        r.bid = -1; r.lo = 1; r.hi = 0; r.mx = neg_inf;
                    }
                }
                if (fast) {
                    r.mx = fmax(r.mx, um);
                } else {                                       // tile straddles a block boundary
                    const long long rx = r0 + 2 * (long long)threadIdx.x;
                    if (vx) atomicMax(P.blockmax + rx / P.ubs, enc_f64(ux[h]));
                    if (vy) atomicMax(P.blockmax + (rx + 1) / P.ubs, enc_f64(uy[h]));
                }
            }
            if (EPI == EPI_SCM_STREAM) {
                const double um = fmax(fmax(ux[0], uy[0]), fmax(ux[1], uy[1]));
                if (__any_sync(0xffffffffu, um >= thr)) {            // rare once the running maximum has settled
                    // raise the CTA's running maximum, pick up what the other warps found, re-derive the threshold
                    unsigned long long e = warp_max_enc(enc_f64(um));
                    if (lane == 0) atomicMax(&s_cmax, e);
                    __syncwarp();
                    const unsigned long long ec = *reinterpret_cast<volatile unsigned long long *>(&s_cmax);
                    const double cur = dec_f64_max_dev(ec > e ? ec : e);
                    if (cur > -INFINITY) thr = __dsub_rn(cur, __dmul_rn(5.0, np_tol_dev(cur)));
                    const long long rule = P.col0 + col;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int h = q >> 1, y = q & 1;
                        const double u = y ? uy[h] : ux[h];
                        if (u >= thr && u > -INFINITY) {
                            const uint32_t c0 = y ? c0y : c0x, c1 = y ? c1y : c1x;
                            const unsigned slot = atomicAdd(&s_ncand, 1u);
                            if (slot < (unsigned)P.cand_cap) {
                                TieRec r;
                                r.idx = (long long)h * P.k_total + rule + y;
                                r.pe = h ? c0 : P.n_pos - c0;
                                r.nc = h ? c1 : P.n_neg - c1;
                                my_cand[slot] = r;
                            }
                        }
                    }
                }
            }
        }
    }
    if (EPI == EPI_SCM || EPI == EPI_SCM_STREAM) {
        runmax_flush<CW>(run[0], P.blockmax, s_red);
        runmax_flush<CW>(run[1], P.blockmax, s_red);
        if (EPI == EPI_SCM_STREAM) {
            consumer_bar(CW * 32);
            scm_stream_tail<CW>(P, ring, s_ncand);
        } else if (P.tail.mode) {
            scm_fused_tail<CW>(P, t0, t1, ring);      // (the ring is idle: every stage was consumed)
        }
    }
    if (EPI == EPI_GINI2) {
        const unsigned long long v = warp_min_enc(enc_f64(gini_min));
        consumer_bar(CW * 32);
        if (lane == 0) s_red[warp] = v;
        consumer_bar(CW * 32);
        if (threadIdx.x == 0) {
            unsigned long long m = s_red[0];
            for (int w = 1; w < CW; ++w) m = s_red[w] < m ? s_red[w] : m;
            atomicMin(P.gmin, m);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_scm_select: device replay of the sequential utility-block merge (scm.py:270-286), one thread.
// Same arithmetic as kv_scm_select_blocks / numpy.isclose: |a-b| <= atol + rtol*|b| and isfinite(b),
// or a == b, every operation rounded separately.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dec_f64_max_dev(unsigned long long v) {
    if (v == 0ull) return -INFINITY;
    const unsigned long long b = (v & 0x8000000000000000ull) ? (v & 0x7FFFFFFFFFFFFFFFull) : ~v;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double np_tol_dev(double b) { return __dadd_rn(1e-8, __dmul_rn(1e-5, fabs(b))); }
__device__ __forceinline__ bool np_isclose_dev(double a, double b) {
    if (a == b) return true;
    if (!isfinite(b)) return false;
    return fabs(__dsub_rn(a, b)) <= np_tol_dev(b);
}

struct TieHeader {               // first 32 bytes of d_ties
    double best_utility;
    unsigned long long count;
    unsigned long long pad[2];
};
// Both select kernels run as ONE warp: the lanes fetch + decode the block maxima in parallel (and zero
// the accumulators for the next scan), lane 0 then walks them in shared memory.
static constexpr int SELECT_SMEM_BLOCKS = 2048;

__device__ __forceinline__ const double *select_prologue(unsigned long long *blockmax, long long n_blocks, double *bm,
                                                         double *s_bm) {
    const bool in_smem = n_blocks <= SELECT_SMEM_BLOCKS;
    for (long long b = threadIdx.x; b < n_blocks; b += 32) {
        const double m = dec_f64_max_dev(blockmax[b]);
        blockmax[b] = 0ull;
        bm[b] = m;
        if (in_smem) s_bm[b] = m;
    }
    __syncwarp();
    return in_smem ? s_bm : bm;
}

__device__ __forceinline__ void scm_select_body(unsigned long long *blockmax, long long n_blocks, double *bm, double *tol,
                                                uint8_t *sel, TieHeader *hdr, double *s_bm) {
    const double *v = select_prologue(blockmax, n_blocks, bm, s_bm);
    if (threadIdx.x != 0) return;
    double best = -INFINITY;
    long long first = 0;                       // blocks before `first` are deselected
    for (long long b = 0; b < n_blocks; ++b) {
        const double m = v[b];
        tol[b] = np_tol_dev(m);
        uint8_t keep = 0;
        if (m > best || np_isclose_dev(best, m)) {
            if (np_isclose_dev(m, best)) {
                keep = 1;
            } else {
                best = m;
                for (long long q = first; q < b; ++q) sel[q] = 0;
                first = b;
                keep = 1;
            }
        }
        sel[b] = keep;
    }
    hdr->best_utility = best;
    hdr->count = 0;
}

__global__ void __launch_bounds__(32) k_scm_select(unsigned long long *blockmax, long long n_blocks, double *bm,
                                                   double *tol, uint8_t *sel, TieHeader *hdr) {
    __shared__ double s_bm[SELECT_SMEM_BLOCKS];
    scm_select_body(blockmax, n_blocks, bm, tol, sel, hdr, s_bm);
}

// Multi-rank variant: this rank only knows ITS block maxima.  Every rule that can survive the global
// merge has u >= Gmax - ~3*T(Gmax) with T(x) = 1e-8 + 1e-5*|x| and Gmax >= Lmax (the local maximum),
// and x - c*T(x) is increasing, so u >= Lmax - 5*T(Lmax) is a superset criterion that needs no
// exchange.  The exact filter runs on the host after the gather.
struct PacketHeader {            // follows the n_blocks doubles in the packet
    unsigned long long count;
    double threshold;
};
__global__ void __launch_bounds__(32) k_scm_select_local(unsigned long long *blockmax, long long n_blocks,
                                                         double *packet_bm, PacketHeader *hdr) {
    __shared__ double s_bm[SELECT_SMEM_BLOCKS];
    const double *v = select_prologue(blockmax, n_blocks, packet_bm, s_bm);
    if (threadIdx.x != 0) return;
    double lmax = -INFINITY;
    for (long long b = 0; b < n_blocks; ++b) lmax = fmax(lmax, v[b]);
    hdr->count = 0;
    hdr->threshold = isfinite(lmax) ? lmax - 5.0 * np_tol_dev(lmax) : -INFINITY;
}

// k_scm_select_local for up to 12 jobs (QUEUE_BATCH): one warp (block) per job
struct SelectLocalMultiParams {
    unsigned long long *blockmax[12];
    double *packet_bm[12];
    PacketHeader *hdr[12];
    long long n_blocks;
};
__global__ void __launch_bounds__(32) k_scm_select_local_multi(const __grid_constant__ SelectLocalMultiParams P) {
    __shared__ double s_bm[SELECT_SMEM_BLOCKS];
    const int j = blockIdx.x;
    const double *v = select_prologue(P.blockmax[j], P.n_blocks, P.packet_bm[j], s_bm);
    if (threadIdx.x != 0) return;
    double lmax = -INFINITY;
    for (long long b = 0; b < P.n_blocks; ++b) lmax = fmax(lmax, v[b]);
    P.hdr[j]->count = 0;
    P.hdr[j]->threshold = isfinite(lmax) ? lmax - 5.0 * np_tol_dev(lmax) : -INFINITY;
}

// ------------------------------------------------------------------------------------------------
// k_scm_ties: rules of the selected utility blocks whose utility is isclose to the block maximum
// ------------------------------------------------------------------------------------------------
struct TieParams {
    const uint32_t *cnt0, *cnt1;
    const unsigned long long *segmax;   // [2][ld / 64]
    long long n_cols, ld, col0, k_total, ubs;
    double inv_ubs, p;
    uint32_t n_pos, n_neg;
    const uint32_t *black_pres, *black_abs;
    const double *bm, *tol;        // per utility block: maximum and atol + rtol * |maximum|
    const uint8_t *sel;            // 1 = collect, 0 = skip
    unsigned long long *count;     // tie counter (TieHeader::count or PacketHeader::count)
    const double *threshold;       // candidate mode (multi-rank): emit every rule with u >= *threshold
    TieRec *out;
    long long cap;
};
static constexpr long long TIE_FLAG_NEG_INF = 1ll << 62;   // candidate record of a blacklisted rule (u = -inf)

__device__ __forceinline__ bool tie_close(double u, double m, double tol) {
    // numpy isclose(u, m): |u - m| <= atol + rtol*|m| and isfinite(m), or u == m
    return (u == m) || (isfinite(m) && fabs(__dsub_rn(u, m)) <= tol);
}

// can the 64-column segment hold a tie?  `smax` is the segment's maximum or an UPPER bound of it (k_scm_rescore_multi
// stores the maximum of the surrounding 256 columns, which may reach into the next utility block).  If all the
// segment's rules sit in one block: only if the block is selected and the bound reaches the block's tie window --
// every u of the segment is <= smax, so when smax < m - tol no u can be close to m.  Straddling segments are scanned.
__device__ __forceinline__ bool seg_candidate(const TieParams &P, long long r0, long long r1, double smax) {
    if (P.threshold) return smax >= *P.threshold;
    const long long b0 = fast_div(r0, P.ubs, P.inv_ubs), b1 = fast_div(r1, P.ubs, P.inv_ubs);
    if (b0 != b1) return true;
    return P.sel[b0] && (smax >= P.bm[b0] || tie_close(smax, P.bm[b0], P.tol[b0]));
}

__device__ __forceinline__ void tie_emit(const TieParams &P, long long rule, double u, uint32_t pe, uint32_t nc) {
    if (P.threshold) {
        if (!(u >= *P.threshold)) return;
        if (u == -INFINITY) rule |= TIE_FLAG_NEG_INF;
    } else {
        const long long b = fast_div(rule, P.ubs, P.inv_ubs);
        if (!P.sel[b] || !tie_close(u, P.bm[b], P.tol[b])) return;
    }
    // one atomic per group of converged lanes that emit together
    const unsigned peers = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(P.count, (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const unsigned long long slot = base + __popc(peers & ((1u << lane) - 1u));
    if ((long long)slot < P.cap) {
        TieRec r;
        r.idx = rule; r.pe = pe; r.nc = nc;
        P.out[slot] = r;
    }
}

// segments [seg0, n_seg) screened 32 at a time by warp `warp0` of `n_warps`
__device__ __forceinline__ void scm_ties_range(const TieParams &P, long long seg0, long long n_seg, long long warp0,
                                               long long n_warps) {
    const int lane = threadIdx.x & 31;
    for (long long sb = seg0 + warp0 * 32; sb < n_seg; sb += n_warps * 32) {
        // each lane screens one segment; the warp then scans the candidates together
        const long long g = sb + lane;
        unsigned cand = 0;
        if (g < n_seg) {
            const long long c0 = g << 6, c1 = min(c0 + 63, P.n_cols - 1);
            const double sp = dec_f64_max_dev(P.segmax[g]);
            const double sa = dec_f64_max_dev(P.segmax[(P.ld >> 6) + g]);
            if (seg_candidate(P, P.col0 + c0, P.col0 + c1, sp)) cand |= 1u;
            if (seg_candidate(P, P.k_total + P.col0 + c0, P.k_total + P.col0 + c1, sa)) cand |= 2u;
        }
        unsigned todo = __ballot_sync(0xffffffffu, cand != 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned which = __shfl_sync(0xffffffffu, cand, src);
            const long long cbase = (sb + src) << 6;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const long long col = cbase + lane + 32 * j;
                if (col >= P.n_cols) continue;
                const uint32_t c0 = P.cnt0[col], c1 = P.cnt1[col];
                bool bp = false, ba = false;
                if (P.black_pres) {
                    bp = (P.black_pres[col >> 5] >> (col & 31)) & 1u;
                    ba = (P.black_abs[col >> 5] >> (col & 31)) & 1u;
                }
                const uint32_t pe_p = P.n_pos - c0, nc_p = P.n_neg - c1;
                if (which & 1u) tie_emit(P, P.col0 + col, bp ? -INFINITY : scm_utility(nc_p, pe_p, P.p), pe_p, nc_p);
                if (which & 2u) tie_emit(P, P.k_total + P.col0 + col, ba ? -INFINITY : scm_utility(c1, c0, P.p), c0, c1);
            }
        }
    }
}

__device__ __forceinline__ void scm_ties_body(const TieParams &P) {
    scm_ties_range(P, 0, (P.n_cols + 63) >> 6, ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                   ((long long)gridDim.x * blockDim.x) >> 5);
}

__global__ void __launch_bounds__(256) k_scm_ties(const TieParams P) { scm_ties_body(P); }

// ---- the fused tail of k_scan_tma<EPI_SCM> (see FusedTail) ----------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

static constexpr long long FUSED_MAX_BLOCKS = 4096;      // block tables of the tail live in the idle TMA ring

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The block merge of scm.py:270-286 by ONE WARP (all 32 lanes call it with identical arguments).  The reference walks
// the blocks in order; a block only matters when its maximum beats or is close to the running best -- a handful of
// blocks out of hundreds on 8 GPUs -- so the lanes test 32 blocks at a time against the current best (ballot) and only
// the hits are processed in order: the same decisions in the same order, without a 100-cycle dependent chain per block.
// tol[b] = atol + rtol * |bm[b]| must be filled, sel[] is written for every block.  Returns best.
__device__ __forceinline__ double block_merge_warp(const double *bm, const double *tol, uint8_t *sel, long long n_blocks) {
    const int lane = threadIdx.x & 31;
    for (long long b = lane; b < n_blocks; b += 32) sel[b] = 0;
    __syncwarp();
    double best = -INFINITY, best_tol = 0.0;        // best_tol: tolerance of isclose(x, best) = atol + rtol * |best|
    long long first = 0, pos = 0;
    while (pos < n_blocks) {
        long long found = -1;
        for (long long base = pos & ~31ll; base < n_blocks; base += 32) {
            const long long b = base + lane;
            bool hit = false;
            if (b >= pos && b < n_blocks) {
                const double m = bm[b];
                // m > best or isclose(best, m): |best - m| <= tol(m) and isfinite(m), or best == m   (scm.py:271)
                hit = (m > best) || (best == m) || (isfinite(m) && fabs(__dsub_rn(best, m)) <= tol[b]);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
                found = base + (__ffs(mask) - 1);
                break;
            }
        }
        if (found < 0) break;
        const double m = bm[found];
        // isclose(m, best): |m - best| <= tol(best) and isfinite(best), or m == best   (scm.py:276)
        const bool close = (m == best) || (isfinite(best) && fabs(__dsub_rn(m, best)) <= best_tol);
        if (!close) {                                // a new best: everything selected so far is dropped (scm.py:282-286)
            for (long long q = first + lane; q < found; q += 32) sel[q] = 0;
            best = m;
            best_tol = tol[found];
            first = found;
        }
        if (lane == 0) sel[found] = 1;
        __syncwarp();
        pos = found + 1;
    }
    __syncwarp();
    return best;
}

// the same, one thread (kept for k_scm_select's single-lane walk semantics in the unfused kernels)
__device__ __forceinline__ double block_merge_seq(const double *bm, double *tol, uint8_t *sel, long long n_blocks) {
    double best = -INFINITY;
    long long first = 0;
    for (long long b = 0; b < n_blocks; ++b) {
        const double m = bm[b];
        tol[b] = np_tol_dev(m);
        uint8_t keep = 0;
        if (m > best || np_isclose_dev(best, m)) {
            if (np_isclose_dev(m, best)) {
                keep = 1;
            } else {
                best = m;
                for (long long q = first; q < b; ++q) sel[q] = 0;
                first = b;
                keep = 1;
            }
        }
        sel[b] = keep;
    }
    return best;
}

// Mode 2, last CTA of the rank: publish the local packet into every peer's mailbox (remote stores over NVLink),
// exchange sequence flags, then merge the packets of all ranks exactly as kv_scm_merge_packets does on the host --
// element-wise max of the block maxima, the sequential block merge, and the isclose filter of the candidates with
// u recomputed from the integer counts by the same two rounded operations (scm.py:263-264).
template <int NT>
__device__ void p2p_exchange_merge(const Scan2Params &P, double *s_bm, double *s_tol, uint8_t *s_sel,
                                   unsigned long long t_tail) {
    const FusedTail &F = P.tail;
    const P2PDev &X = *F.p2p;
    const int tid = threadIdx.x;
    const int world = X.world, parity = (int)(F.p2p_seq & 1ull);
    __shared__ unsigned long long s_out;
    __shared__ int s_status;
    __shared__ double s_best2;
    const size_t nb_bytes = (size_t)F.n_blocks * 8;
    const PacketHeader *hdr = reinterpret_cast<const PacketHeader *>(F.packet + nb_bytes);
    const long long n_rec = min((long long)__ldcg(&hdr->count), F.packet_cap);
    const size_t n16 = (nb_bytes + sizeof(PacketHeader) + (size_t)n_rec * sizeof(TieRec) + 15) / 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(F.packet);
    for (int r = 0; r < world; ++r) {
        uint4 *dst = reinterpret_cast<uint4 *>(X.peer_box[r] + ((size_t)parity * world + X.rank) * X.slot_bytes);
        for (size_t i = tid; i < n16; i += NT) dst[i] = __ldcg(src + i);
    }
    __shared__ unsigned long long s_stamp[2];    // (kept on chip: a store to host memory here would sit in front of
    if (tid == 0) {                              //  the system-scope fence below and cost a PCIe round trip)
        s_out = 0ull;
        s_status = 0;
        s_stamp[0] = globaltimer_ns();           // packet published (stores issued), about to exchange flags
    }
    // (no separate system fence here: the packet stores of all threads happen-before the barrier, and the flag store
    //  below is a release at system scope, which is cumulative over them)
    consumer_bar(NT);
    if (tid < world) {
        unsigned long long *f = X.peer_flags[tid] + X.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(F.p2p_seq) : "memory");
        const unsigned long long *mine = X.peer_flags[X.rank] + tid;
        const unsigned long long t0 = globaltimer_ns();
        unsigned long long v = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= F.p2p_seq) break;
            if (globaltimer_ns() - t0 > F.timeout_ns) {
                s_status = 1;
                break;
            }
        }
    }
    consumer_bar(NT);                            // (the acquire loads above order the packet reads below, through the barrier)
    if (tid == 0) s_stamp[1] = globaltimer_ns();    // every peer's packet has arrived
    const char *box = X.peer_box[X.rank] + (size_t)parity * world * X.slot_bytes;
    for (long long b = tid; b < F.n_blocks; b += NT) {
        double m = -INFINITY;
#pragma unroll 4
        for (int r = 0; r < world; ++r) {
            const double v = __ldcv(reinterpret_cast<const double *>(box + (size_t)r * X.slot_bytes) + b);
            m = v > m ? v : m;
        }
        s_bm[b] = m;
        s_tol[b] = np_tol_dev(m);
    }
    consumer_bar(NT);
    __shared__ long long s_cnt[P2P_MAX_WORLD];
    if (tid < 32) {
        const double best = block_merge_warp(s_bm, s_tol, s_sel, F.n_blocks);
        if (tid == 0) s_best2 = best;
    } else if (tid - 32 < world) {               // every rank's candidate count, fetched in parallel
        const int r = tid - 32;
        const PacketHeader *h = reinterpret_cast<const PacketHeader *>(box + (size_t)r * X.slot_bytes + nb_bytes);
        const long long c = (long long)__ldcv(&h->count);
        s_cnt[r] = c;
        if (c > F.packet_cap) atomicCAS(&s_status, 0, 2);
    }
    consumer_bar(NT);
    if (s_status == 0) {
        const double inv_ubs = 1.0 / (double)P.ubs;
        // one warp per rank's packet (round robin), one lane per candidate: the packets are filtered side by side
        for (int r = tid >> 5; r < world; r += NT / 32) {
            const PacketHeader *h = reinterpret_cast<const PacketHeader *>(box + (size_t)r * X.slot_bytes + nb_bytes);
            const long long cnt = s_cnt[r];
            const TieRec *recs = reinterpret_cast<const TieRec *>(h + 1);
            for (long long i = tid & 31; i < cnt; i += 32) {
                // (records are 8-byte aligned only: the packet head is n_blocks doubles + a 16-byte header)
                const unsigned long long *rw = reinterpret_cast<const unsigned long long *>(recs + i);
                long long idx = (long long)__ldcv(rw);
                const unsigned long long cnts = __ldcv(rw + 1);
                const uint32_t pe = (uint32_t)cnts, nc = (uint32_t)(cnts >> 32);
                const bool ninf = (idx & TIE_FLAG_NEG_INF) != 0;
                idx &= ~TIE_FLAG_NEG_INF;
                const long long b = fast_div(idx, P.ubs, inv_ubs);
                if (b < 0 || b >= F.n_blocks || !s_sel[b]) continue;
                const double u = ninf ? -INFINITY : __dsub_rn((double)nc, __dmul_rn(P.p, (double)pe));
                if (!np_isclose_dev(u, s_bm[b])) continue;
                const unsigned long long slot = atomicAdd(&s_out, 1ull);
                if ((long long)slot < F.cap) {
                    TieRec o;
                    o.idx = idx; o.pe = pe; o.nc = nc;
                    F.out[slot] = o;
                }
            }
        }
    }
    for (long long b = tid; b < F.n_blocks; b += NT) {
        F.host_bm[b] = s_bm[b];
        F.host_sel[b] = s_sel[b];
    }
    consumer_bar(NT);
    if (tid == 0) {
        if (s_status == 0 && (long long)s_out > F.cap) s_status = 2;
        F.host[0] = (unsigned long long)__double_as_longlong(s_best2);
        F.host[1] = s_out;
        F.host[3] = (unsigned long long)s_status;
        F.host[4] = t_tail;
        F.host[5] = s_stamp[0];
        F.host[6] = s_stamp[1];
        F.host[7] = globaltimer_ns();
        __threadfence_system();
        F.host[2] = F.epoch;
    }
}

template <int CW>
__device__ void scm_fused_tail(const Scan2Params &P, long long t0, long long t1, unsigned char *ring) {
    const FusedTail &F = P.tail;
    const int tid = threadIdx.x;
    constexpr int NT = CW * 32;
    __shared__ double s_best;               // mode 1: best utility; mode 2: the rank's candidate threshold
    __shared__ int s_last;
    if (blockIdx.x == 0) {
        for (long long b = tid; b < F.n_blocks; b += NT) F.zero_next[b] = 0ull;
        if (tid == 0) *F.tie_count = 0ull;
    }
    consumer_bar(NT);                            // this CTA's block maxima (atomicMax) and segment maxima are issued
    if (tid == 0) {
        __threadfence();
        atomicAdd(F.barrier, 1ull);
        while (ld_acquire_gpu_u64(F.barrier) < F.barrier_target) {
        }
    }
    consumer_bar(NT);
    // every CTA works on the complete block maxima of this shard: ~20 blocks at config 2
    double *s_bm = reinterpret_cast<double *>(ring);
    double *s_tol = s_bm + F.n_blocks;
    uint8_t *s_sel = reinterpret_cast<uint8_t *>(s_tol + F.n_blocks);
    for (long long b = tid; b < F.n_blocks; b += NT) {
        const double m = dec_f64_max_dev(__ldcg(P.blockmax + b));
        s_bm[b] = m;
        s_tol[b] = np_tol_dev(m);
    }
    consumer_bar(NT);
    if (tid < 32) {
        if (F.mode == 1) {
            const double best = block_merge_warp(s_bm, s_tol, s_sel, F.n_blocks);       // scm.py:270-286
            if (tid == 0) s_best = best;
        } else {
            // this rank only knows ITS maxima: superset threshold (see k_scm_select_local)
            double lmax = -INFINITY;
            for (long long b = tid; b < F.n_blocks; b += 32) lmax = fmax(lmax, s_bm[b]);
            lmax = warp_max(lmax);
            if (tid == 0) s_best = isfinite(lmax) ? lmax - 5.0 * np_tol_dev(lmax) : -INFINITY;
        }
    }
    consumer_bar(NT);
    if (blockIdx.x == 0) {
        if (F.mode == 1) {                       // the tables later phases expect in d_misc (tie-buffer overflow path)
            for (long long b = tid; b < F.n_blocks; b += NT) {
                F.bm[b] = s_bm[b];
                F.tol[b] = s_tol[b];
                F.sel[b] = s_sel[b];
            }
        } else {                                 // the head of the packet: block maxima + threshold
            double *pk = reinterpret_cast<double *>(F.packet);
            for (long long b = tid; b < F.n_blocks; b += NT) pk[b] = s_bm[b];
            if (tid == 0) reinterpret_cast<PacketHeader *>(F.packet + (size_t)F.n_blocks * 8)->threshold = s_best;
        }
    }
    TieParams T;
    T.cnt0 = P.cnt0;
    T.cnt1 = P.cnt1;
    T.segmax = P.segmax;
    T.n_cols = P.n_cols;
    T.ld = P.ld;
    T.col0 = P.col0;
    T.k_total = P.k_total;
    T.ubs = P.ubs;
    T.inv_ubs = 1.0 / (double)P.ubs;
    T.p = P.p;
    T.n_pos = P.n_pos;
    T.n_neg = P.n_neg;
    T.black_pres = P.black_pres;
    T.black_abs = P.black_abs;
    T.bm = s_bm;
    T.tol = s_tol;
    T.sel = s_sel;
    T.count = F.tie_count;
    T.threshold = F.mode == 1 ? nullptr : &s_best;
    T.out = F.mode == 1 ? F.out : reinterpret_cast<TieRec *>(F.packet + (size_t)F.n_blocks * 8 + sizeof(PacketHeader));
    T.cap = F.mode == 1 ? F.cap : F.packet_cap;
    const long long n_seg = (P.n_cols + 63) >> 6;
    const long long seg0 = t0 * (long long)CW, seg1 = min(t1 * (long long)CW, n_seg);
    scm_ties_range(T, seg0, seg1, tid >> 5, CW);
    consumer_bar(NT);
    if (tid == 0) {
        __threadfence_system();                  // this CTA's tie records before its ticket
        const unsigned long long ticket = atomicAdd(F.done, 1ull);
        s_last = ticket + 1 == F.done_target;
        __threadfence();
    }
    consumer_bar(NT);
    if (!s_last) return;
    __shared__ unsigned long long s_t_last;
    if (tid == 0) s_t_last = globaltimer_ns();      // the scan and this rank's tie / candidate collection are complete
    if (F.mode == 1) {
        if (tid == 0) {                          // last CTA: publish the result
            const unsigned long long found = atomicAdd(F.tie_count, 0ull);
            F.host[0] = (unsigned long long)__double_as_longlong(s_best);
            F.host[1] = found;
            F.host[3] = 0ull;
            F.host[4] = s_t_last;
            F.host[5] = F.host[6] = F.host[7] = globaltimer_ns();
            __threadfence_system();
            F.host[2] = F.epoch;
        }
        return;
    }
    p2p_exchange_merge<NT>(P, s_bm, s_tol, s_sel, s_t_last);
}

// The tail of k_scan_tma<EPI_SCM_STREAM>: ONE synchronisation point.  Every CTA records how many candidates it appended
// and takes a ticket; the last CTA of the grid owns the rest of the step: block merge (scm.py:270-286) on the complete
// maxima, exact filter of ALL the candidate lists (a few thousand records in L2), result into mapped host memory
// (mode 1) or packet -> exchange -> merge of the ranks' packets (mode 2).  No grid barrier, so no co-residency
// requirement and nothing spins while the slowest CTA finishes.  A candidate list that overflowed, or a best utility of
// -inf (every rule blacklisted: the tie set is the whole block), is reported as "more ties than the buffer holds" and
// the caller redoes the step with the count arrays.
template <int CW>
__device__ void scm_stream_tail(const Scan2Params &P, unsigned char *ring, unsigned n_cand) {
    const FusedTail &F = P.tail;
    const int tid = threadIdx.x;
    constexpr int NT = CW * 32;
    __shared__ double s_best;               // mode 1: best utility; mode 2: the rank's candidate threshold
    __shared__ int s_last;
    __shared__ unsigned s_over;
    __shared__ unsigned long long s_emit, s_t_last;
    if (tid == 0) {
        P.cand_n[blockIdx.x] = n_cand;
        s_over = 0u;
        s_emit = 0ull;
        __threadfence();                         // this CTA's candidates and block maxima before its ticket
        const unsigned long long ticket = atomicAdd(F.done, 1ull);
        s_last = ticket + 1 == F.done_target;
        __threadfence();
    }
    consumer_bar(NT);
    if (!s_last) return;
    if (tid == 0) s_t_last = globaltimer_ns();
    const int G = (int)gridDim.x;
    double *s_bm = reinterpret_cast<double *>(ring);
    double *s_tol = s_bm + F.n_blocks;
    uint8_t *s_sel = reinterpret_cast<uint8_t *>(s_tol + F.n_blocks);
    unsigned *s_pre = reinterpret_cast<unsigned *>(ring + (((size_t)F.n_blocks * 17 + 63) / 16) * 16);   // [G + 1]
    for (long long b = tid; b < F.n_blocks; b += NT) {
        const double m = dec_f64_max_dev(__ldcg(P.blockmax + b));
        s_bm[b] = m;
        s_tol[b] = np_tol_dev(m);
        F.zero_next[b] = 0ull;
    }
    for (int c = tid; c < G; c += NT) {
        unsigned n = __ldcg(P.cand_n + c);
        if (n > (unsigned)P.cand_cap) {
            n = (unsigned)P.cand_cap;
            s_over = 1u;
        }
        s_pre[c + 1] = n;
    }
    if (tid == 0) s_pre[0] = 0u;
    consumer_bar(NT);
    if (tid < 32) {
        if (F.mode == 1) {
            const double best = block_merge_warp(s_bm, s_tol, s_sel, F.n_blocks);
            if (tid == 0) s_best = best;
        } else {
            double lmax = -INFINITY;
            for (long long b = tid; b < F.n_blocks; b += 32) lmax = fmax(lmax, s_bm[b]);
            lmax = warp_max(lmax);
            if (tid == 0) s_best = isfinite(lmax) ? lmax - 5.0 * np_tol_dev(lmax) : -INFINITY;
        }
    } else if (tid < 64) {                       // inclusive scan of the list lengths, 32 at a time
        const int lane = tid & 31;
        unsigned carry = 0u;
        for (int base = 0; base < G; base += 32) {
            const int c = base + lane;
            unsigned v = c < G ? s_pre[c + 1] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += o;
            }
            if (c < G) s_pre[c + 1] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    consumer_bar(NT);
    const bool exact = F.mode == 1;
    TieRec *out = exact ? F.out : reinterpret_cast<TieRec *>(F.packet + (size_t)F.n_blocks * 8 + sizeof(PacketHeader));
    const long long cap = exact ? F.cap : F.packet_cap;
    const unsigned total = s_pre[G];
    const double inv_ubs = 1.0 / (double)P.ubs, thr = s_best;
    for (unsigned i = tid; i < total; i += NT) {
        int lo = 0, hi = G;                      // s_pre[lo] <= i < s_pre[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_pre[mid] <= i) lo = mid; else hi = mid;
        }
        const unsigned long long *rw =
            reinterpret_cast<const unsigned long long *>(P.cand + (size_t)lo * P.cand_cap + (i - s_pre[lo]));
        const long long idx = (long long)__ldcg(rw);
        const unsigned long long cnts = __ldcg(rw + 1);
        const uint32_t pe = (uint32_t)cnts, nc = (uint32_t)(cnts >> 32);
        const double u = scm_utility(nc, pe, P.p);
        bool keep;
        if (exact) {
            const long long b = fast_div(idx, P.ubs, inv_ubs);
            keep = s_sel[b] && tie_close(u, s_bm[b], s_tol[b]);
        } else {
            keep = u >= thr;
        }
        if (keep) {
            const unsigned long long slot = atomicAdd(&s_emit, 1ull);
            if ((long long)slot < cap) {
                TieRec o;
                o.idx = idx; o.pe = pe; o.nc = nc;
                out[slot] = o;
            }
        }
    }
    if (!exact) {                                // the head of the packet: block maxima + threshold
        double *pk = reinterpret_cast<double *>(F.packet);
        for (long long b = tid; b < F.n_blocks; b += NT) pk[b] = s_bm[b];
    }
    consumer_bar(NT);
    const bool redo = s_over != 0u || !(s_best > -INFINITY);
    if (exact) {
        if (tid == 0) {
            F.host[0] = (unsigned long long)__double_as_longlong(s_best);
            F.host[1] = redo ? (unsigned long long)(F.cap + 1) : s_emit;
            F.host[3] = redo ? 2ull : 0ull;
            F.host[4] = s_t_last;
            F.host[5] = F.host[6] = F.host[7] = globaltimer_ns();
            __threadfence_system();
            F.host[2] = F.epoch;
        }
        return;
    }
    if (tid == 0) {
        PacketHeader *h = reinterpret_cast<PacketHeader *>(F.packet + (size_t)F.n_blocks * 8);
        h->threshold = s_best;
        h->count = redo ? (unsigned long long)(F.packet_cap + 1) : s_emit;
    }
    consumer_bar(NT);
    p2p_exchange_merge<NT>(P, s_bm, s_tol, s_sel, s_t_last);
}

// the tie collection of up to QUEUE_BATCH jobs in one launch: blockIdx.y selects the job
static constexpr int QUEUE_BATCH = 12;
struct TieParamsMulti {
    TieParams j[QUEUE_BATCH];
};
__global__ void __launch_bounds__(256) k_scm_ties_multi(const __grid_constant__ TieParamsMulti M) {
    scm_ties_body(M.j[blockIdx.y]);
}

// ------------------------------------------------------------------------------------------------
// k_scm_rescore: phase 1 from RESIDENT counts -- another p, or the two masks' roles swapped (the
// disjunction swaps the example labels, scm.py:69-73), without touching the matrix: 8 bytes per column
// instead of 8*W.  Produces exactly what k_scan_tma<EPI_SCM>'s epilogue produces (block maxima, segment
// maxima) so phase 2 is unchanged.  One CTA per 2 048 columns, one warp per 64-column segment at a time.
// ------------------------------------------------------------------------------------------------
struct RescoreParams {
    const uint32_t *cnt_pos, *cnt_neg;     // presence counts under the (possibly swapped) positive / negative mask
    long long n_cols, ld, col0, k_total, ubs;
    double p;
    uint32_t n_pos, n_neg;
    const uint32_t *black_pres, *black_abs;
    unsigned long long *blockmax, *segmax;
};

// max over a 16-lane half warp of an enc_f64 image (the two halves reduce independently)
__device__ __forceinline__ unsigned long long half_warp_max_enc(unsigned long long e, unsigned half_mask) {
    const uint32_t hi = (uint32_t)(e >> 32), lo = (uint32_t)e;
    const uint32_t mh = __reduce_max_sync(half_mask, hi);
    const uint32_t ml = __reduce_max_sync(half_mask, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

__global__ void __launch_bounds__(256) k_scm_rescore(const RescoreParams P) {
    // 2 048 columns per CTA; a warp takes 128 columns per step: each 16-lane half owns one 64-column
    // segment, each lane four adjacent columns (one 128-bit load per count array).
    __shared__ unsigned long long s_red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, sub = lane & 15;
    const unsigned half_mask = half ? 0xFFFF0000u : 0x0000FFFFu;
    const long long chunk_c0 = (long long)blockIdx.x * 2048;
    const long long chunk_n = min(2048ll, P.n_cols - chunk_c0);
    unsigned long long run[2] = {0ull, 0ull};          // enc_f64 images; 0 = nothing seen
    bool uniform[2];
    long long bid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long r0 = (long long)h * P.k_total + P.col0 + chunk_c0;
        bid[h] = r0 / P.ubs;
        uniform[h] = chunk_n > 0 && (r0 + chunk_n - 1) / P.ubs == bid[h];
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const long long seg_c0 = chunk_c0 + (long long)(it * 16 + warp * 2 + half) * 64;
        if (seg_c0 >= P.ld) continue;                   // half-warp uniform
        const long long col = seg_c0 + sub * 4;         // ld is a multiple of 1024: the vector loads stay in bounds
        const uint4 cp = *reinterpret_cast<const uint4 *>(P.cnt_pos + col);
        const uint4 cn = *reinterpret_cast<const uint4 *>(P.cnt_neg + col);
        const uint32_t cpv[4] = {cp.x, cp.y, cp.z, cp.w}, cnv[4] = {cn.x, cn.y, cn.z, cn.w};
        uint32_t bp = 0, ba = 0;
        if (P.black_pres) {
            bp = P.black_pres[col >> 5] >> (col & 31);
            ba = P.black_abs[col >> 5] >> (col & 31);
        }
        double u[2][4];
        double um[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool valid = col + j < P.n_cols;
            u[0][j] = (!valid || ((bp >> j) & 1u)) ? -INFINITY : scm_utility(P.n_neg - cnv[j], P.n_pos - cpv[j], P.p);
            u[1][j] = (!valid || ((ba >> j) & 1u)) ? -INFINITY : scm_utility(cnv[j], cpv[j], P.p);
            um[0] = fmax(um[0], u[0][j]);
            um[1] = fmax(um[1], u[1][j]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned long long seg = half_warp_max_enc(enc_f64(um[h]), half_mask);
            if (sub == 0) P.segmax[(size_t)h * (P.ld >> 6) + (seg_c0 >> 6)] = seg;
            if (uniform[h]) {
                run[h] = seg > run[h] ? seg : run[h];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < P.n_cols) {
                        const long long r = (long long)h * P.k_total + P.col0 + col + j;
                        atomicMax(P.blockmax + r / P.ubs, enc_f64(u[h][j]));
                    }
            }
        }
    }
    // lanes 0 and 16 hold their half's running maxima
    const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, run[0], 16), o1 = __shfl_xor_sync(0xffffffffu, run[1], 16);
    if (lane == 0) {
        s_red[0][warp] = run[0] > o0 ? run[0] : o0;
        s_red[1][warp] = run[1] > o1 ? run[1] : o1;
    }
    __syncthreads();
    if (threadIdx.x < 2 && uniform[threadIdx.x]) {
        unsigned long long m = 0ull;
        for (int w = 0; w < 8; ++w) m = s_red[threadIdx.x][w] > m ? s_red[threadIdx.x][w] : m;
        if (m) atomicMax(P.blockmax + bid[threadIdx.x], m);
    }
}

// k_scm_rescore for up to QUEUE_BATCH jobs that share a mask pair (the p values of a grid, both model types): the
// counts are read ONCE, every job gets its own segment maxima and block maxima.  Same arithmetic per job.
struct RescoreMultiParams {
    const uint32_t *cnt_a, *cnt_b;         // presence counts under the pair's first / second mask
    long long n_cols, ld, col0, k_total, ubs;
    const uint32_t *black_pres, *black_abs;
    int n_jobs, n_plain;                   // jobs [0, n_plain): positives = the pair's FIRST mask; [n_plain, n_jobs): swapped
    double p[QUEUE_BATCH];
    double n_pos[QUEUE_BATCH], n_neg[QUEUE_BATCH];
    unsigned long long *blockmax[QUEUE_BATCH], *segmax[QUEUE_BATCH];
    // every job of the launch counts its examples over the same mask pair (|first mask| = dn_a, |second| = dn_b; the
    // swapped jobs with exchanged roles): the p-independent halves of the utilities are then formed once per column
    int hoist;
    double dn_a, dn_b;
};

// 2 048 columns per CTA: every warp owns 256 consecutive columns, every lane eight of them (two 128-bit loads per
// count array).  The kernel is issue-bound, not memory-bound (the first versions spent ~150, then ~100 instructions
// per warp, job and 64 columns -- profiles/README.md), so the per-job overhead (parameter loads, order-preserving
// image, the two redux.sync of a 64-bit maximum, the store) is paid once per 256 columns: the maximum of the 256
// columns is written to all four of their 64-column segment slots, an UPPER bound of each, which is all the tie
// kernel's screening needs (it only skips segments whose bound cannot reach the block maximum).  The counts are
// converted to double ONCE (int -> double conversions run on the 16-lane XU pipe and would otherwise dominate);
// n_neg - cnt is then formed in double, which is exact for integers and therefore the same value the reference
// converts (scm.py:245-264).  Jobs are ordered plain-then-swapped by the host so that the roles of the two count
// arrays are fixed per loop.
// exact uint32 -> double without the XU pipe's I2F: 2^52 + x has x in its low mantissa bits; subtracting 2^52 is exact
__device__ __forceinline__ double u32_to_f64(uint32_t x) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)x), 4503599627370496.0);
}

struct RescoreLane {
    double cp[8], cn[8];            // presence counts of the lane's columns under the positive / negative mask
    uint32_t dead_p, dead_a;        // bit q: rule of column q must score -inf (padding column / blacklisted)
    bool uniform[2];
    long long col[2], rule0[2];     // first column of the lane's two quads; first rule of the warp's span per half
    size_t seg_idx, seg_half;
    int lane, warp, n_seg;
};

__device__ __noinline__ void rescore_straddle(unsigned long long *blockmax, long long rule0, long long ubs, int n_valid,
                                              double u0, double u1, double u2, double u3) {
    const double u[4] = {u0, u1, u2, u3};
    for (int q = 0; q < n_valid; ++q) atomicMax(blockmax + (rule0 + q) / ubs, enc_f64(u[q]));
}

template <int H>
__device__ __forceinline__ void rescore_half(const RescoreMultiParams &P, int j, const RescoreLane &L, const double (&u)[8],
                                             unsigned long long (*s_red)[2][8]) {
    // (utilities are never NaN: a plain compare-and-select maximum, without fmax's NaN handling)
    const double m01 = u[0] > u[1] ? u[0] : u[1], m23 = u[2] > u[3] ? u[2] : u[3];
    const double m45 = u[4] > u[5] ? u[4] : u[5], m67 = u[6] > u[7] ? u[6] : u[7];
    const double m03 = m01 > m23 ? m01 : m23, m47 = m45 > m67 ? m45 : m67;
    const double m = m03 > m47 ? m03 : m47;
    const unsigned long long seg = warp_max_enc(enc_f64(m));
    if (L.lane < L.n_seg) P.segmax[j][H * L.seg_half + L.seg_idx + L.lane] = seg;
    if (L.lane == 0) s_red[j][H][L.warp] = seg;
    if (!L.uniform[H]) {                                       // CTA-uniform, rare: the CTA's span straddles two blocks
        const long long base = (long long)H * P.k_total + P.col0;
        rescore_straddle(P.blockmax[j], base + L.col[0], P.ubs, (int)max(0ll, min(4ll, P.n_cols - L.col[0])), u[0], u[1],
                         u[2], u[3]);
        rescore_straddle(P.blockmax[j], base + L.col[1], P.ubs, (int)max(0ll, min(4ll, P.n_cols - L.col[1])), u[4], u[5],
                         u[6], u[7]);
    }
}

__device__ __forceinline__ void rescore_job(const RescoreMultiParams &P, int j, const RescoreLane &L,
                                            unsigned long long (*s_red)[2][8]) {
    const double p = P.p[j], dn_pos = P.n_pos[j], dn_neg = P.n_neg[j];
    // presence rule: covers the negatives WITHOUT the k-mer, errs on the positives without it; absence: complements
    double up[8], ua[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        up[q] = __dsub_rn(__dsub_rn(dn_neg, L.cn[q]), __dmul_rn(p, __dsub_rn(dn_pos, L.cp[q])));
        ua[q] = __dsub_rn(L.cn[q], __dmul_rn(p, L.cp[q]));
    }
    if (L.dead_p | L.dead_a) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if ((L.dead_p >> q) & 1u) up[q] = -INFINITY;
            if ((L.dead_a >> q) & 1u) ua[q] = -INFINITY;
        }
    }
    rescore_half<0>(P, j, L, up, s_red);
    rescore_half<1>(P, j, L, ua, s_red);
}

// The same with the p-independent terms hoisted out of the job loop: A = |negatives| - cn (negatives the presence rule
// covers), B = |positives| - cp (positives it errs on); plain jobs: presence A - p B, absence cn - p cp; swapped jobs
// (roles of the two masks exchanged): presence B - p A, absence cp - p cn.  Bit-identical to rescore_job: the same
// rounded operations on the same operands, only formed once.  DEAD (warp-uniform): some lane holds a padding column or
// a blacklisted rule.
template <bool DEAD>
__device__ __forceinline__ void rescore_job_hoisted(const RescoreMultiParams &P, int j, const RescoreLane &L,
                                                    const double (&a)[8], const double (&b)[8], const double (&cn)[8],
                                                    const double (&cp)[8], unsigned long long (*s_red)[2][8]) {
    const double p = P.p[j];
    double up[8], ua[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        up[q] = __dsub_rn(a[q], __dmul_rn(p, b[q]));
        ua[q] = __dsub_rn(cn[q], __dmul_rn(p, cp[q]));
    }
    if (DEAD) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if ((L.dead_p >> q) & 1u) up[q] = -INFINITY;
            if ((L.dead_a >> q) & 1u) ua[q] = -INFINITY;
        }
    }
    rescore_half<0>(P, j, L, up, s_red);
    rescore_half<1>(P, j, L, ua, s_red);
}

template <bool DEAD>
__device__ __forceinline__ void rescore_jobs_hoisted(const RescoreMultiParams &P, const RescoreLane &L,
                                                     unsigned long long (*s_red)[2][8]) {
    double a[8], b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        a[q] = __dsub_rn(P.dn_b, L.cn[q]);
        b[q] = __dsub_rn(P.dn_a, L.cp[q]);
    }
    for (int j = 0; j < P.n_plain; ++j) rescore_job_hoisted<DEAD>(P, j, L, a, b, L.cn, L.cp, s_red);
    for (int j = P.n_plain; j < P.n_jobs; ++j) rescore_job_hoisted<DEAD>(P, j, L, b, a, L.cp, L.cn, s_red);
}

template <bool HOIST>
__global__ void __launch_bounds__(256, 2) k_scm_rescore_multi(const __grid_constant__ RescoreMultiParams P) {
    __shared__ unsigned long long s_red[QUEUE_BATCH][2][8];
    RescoreLane L;
    L.lane = threadIdx.x & 31;
    L.warp = threadIdx.x >> 5;
    const long long chunk_c0 = (long long)blockIdx.x * 2048;
    const long long chunk_n = min(2048ll, P.n_cols - chunk_c0);
    long long bid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long r0 = (long long)h * P.k_total + P.col0 + chunk_c0;
        bid[h] = r0 / P.ubs;
        L.uniform[h] = chunk_n > 0 && (r0 + chunk_n - 1) / P.ubs == bid[h];
    }
    const long long span_c0 = chunk_c0 + (long long)L.warp * 256;       // the warp's 256 columns (ld is a multiple of 1024)
    const bool in_ld = span_c0 < P.ld;
    L.seg_idx = (size_t)(span_c0 >> 6);
    L.seg_half = (size_t)(P.ld >> 6);
    L.n_seg = in_ld ? 4 : 0;
    L.dead_p = L.dead_a = 0;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        L.col[g] = span_c0 + g * 128 + L.lane * 4;
        uint4 va = make_uint4(0, 0, 0, 0), vb = va;
        uint32_t bp = 0, ba = 0;
        if (in_ld) {
            va = *reinterpret_cast<const uint4 *>(P.cnt_a + L.col[g]);
            vb = *reinterpret_cast<const uint4 *>(P.cnt_b + L.col[g]);
            if (P.black_pres) {
                bp = (P.black_pres[L.col[g] >> 5] >> (L.col[g] & 31)) & 15u;
                ba = (P.black_abs[L.col[g] >> 5] >> (L.col[g] & 31)) & 15u;
            }
        }
        L.cp[g * 4 + 0] = u32_to_f64(va.x); L.cp[g * 4 + 1] = u32_to_f64(va.y); L.cp[g * 4 + 2] = u32_to_f64(va.z); L.cp[g * 4 + 3] = u32_to_f64(va.w);
        L.cn[g * 4 + 0] = u32_to_f64(vb.x); L.cn[g * 4 + 1] = u32_to_f64(vb.y); L.cn[g * 4 + 2] = u32_to_f64(vb.z); L.cn[g * 4 + 3] = u32_to_f64(vb.w);
        // rules this lane may NOT score: padding columns and blacklisted rules (scm.py:267) -> -inf
        const int n_valid = (int)max(0ll, min(4ll, P.n_cols - L.col[g]));
        const uint32_t pad = 15u & ~((1u << n_valid) - 1u);
        L.dead_p |= (bp | pad) << (4 * g);
        L.dead_a |= (ba | pad) << (4 * g);
    }
    if (HOIST) {
        if (__any_sync(0xffffffffu, (L.dead_p | L.dead_a) != 0u)) rescore_jobs_hoisted<true>(P, L, s_red);
        else rescore_jobs_hoisted<false>(P, L, s_red);
    } else {
        for (int j = 0; j < P.n_plain; ++j) rescore_job(P, j, L, s_red);
#pragma unroll
        for (int q = 0; q < 8; ++q) {                           // swapped jobs: the two count arrays exchange roles
            const double t = L.cp[q];
            L.cp[q] = L.cn[q];
            L.cn[q] = t;
        }
        for (int j = P.n_plain; j < P.n_jobs; ++j) rescore_job(P, j, L, s_red);
    }
    __syncthreads();
    if (threadIdx.x < 2 * P.n_jobs) {
        const int j = threadIdx.x >> 1, h = threadIdx.x & 1;
        if (L.uniform[h]) {
            unsigned long long m = 0ull;
            for (int w = 0; w < 8; ++w) m = s_red[j][h][w] > m ? s_red[j][h][w] : m;
            if (m) atomicMax(P.blockmax[j] + bid[h], m);
        }
    }
}

// k_scm_select for up to QUEUE_BATCH jobs: one warp (block) per job
struct SelectMultiParams {
    unsigned long long *blockmax[QUEUE_BATCH];
    double *bm[QUEUE_BATCH], *tol[QUEUE_BATCH];
    uint8_t *sel[QUEUE_BATCH];
    TieHeader *hdr[QUEUE_BATCH];
    long long n_blocks;
};
__global__ void __launch_bounds__(32) k_scm_select_multi(const __grid_constant__ SelectMultiParams P) {
    __shared__ double s_bm[SELECT_SMEM_BLOCKS];
    const int j = blockIdx.x;
    scm_select_body(P.blockmax[j], P.n_blocks, P.bm[j], P.tol[j], P.sel[j], P.hdr[j], s_bm);
}

// ------------------------------------------------------------------------------------------------
// k_count_tma: presence counts under NM row masks in ONE pass over the shard (rules.py:201-267 for many
// row sets at once): the mask pairs of different cross-validation folds / greedy states of a
// hyper-parameter grid, the class masks of all the nodes of a CART level, batched sum_rows.
// Same producer / ring / consumer structure as k_scan_tma; the masks of a packed row sit in shared memory
// as [row][NM] and a zero mask word is skipped (warp-uniform), like the reference skips packed rows whose
// mask is empty (rules.py:244).  With many masks the pass is bound by the POPC pipe (16 lanes/clk/SM), not
// by HBM: DESIGN.md section 4.5.
// ------------------------------------------------------------------------------------------------
static constexpr int MAX_MULTI = 16;     // masks per pass of k_count_tma

struct CountParams {
    const uint64_t *mat;
    long long ld;
    const uint32_t *rec_row;      // active rows (any mask non-zero)
    const uint64_t *rec_m;        // [n_act][NM]
    int n_act;
    uint32_t *cnt[MAX_MULTI];     // [ld] each; null = discard (padding mask)
};

// Mask staging: [group of 4 consecutive active rows][NM][4] so that a (group, mask) needs two 128-bit broadcast
// reads; s_nz[group] has bit m set when mask m is non-zero on any of the group's rows.
template <int NM, int CW, int R, int S, bool CSA>
__global__ void __launch_bounds__((CW + 1) * 32, 1) k_count_tma(const CountParams P) {
    static_assert(R % 4 == 0, "row groups must not straddle ring stages");
    extern __shared__ __align__(128) unsigned char smem_tma[];
    constexpr int ROW_BYTES = CW * 32 * 16;
    constexpr int TC = CW * 64;
    const int n_act = P.n_act;
    const int n_grp = (n_act + 3) >> 2;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_tma);
    uint64_t *empty = full + S;
    uint64_t *s_m = empty + S;                               // 16-byte aligned: 2*S*8 is a multiple of 16
    uint32_t *s_row = reinterpret_cast<uint32_t *>(s_m + (size_t)n_grp * NM * 4);
    uint32_t *s_nz = s_row + n_grp * 4;
    unsigned char *ring = smem_tma + ((size_t)(2 * S) * 8 + (size_t)n_grp * (NM * 32 + 20) + 127) / 128 * 128;
    for (int i = threadIdx.x; i < n_grp * NM * 4; i += blockDim.x) s_m[i] = P.rec_m[i];
    for (int i = threadIdx.x; i < n_grp * 4; i += blockDim.x) s_row[i] = P.rec_row[i];
    __syncthreads();
    for (int g = threadIdx.x; g < n_grp; g += blockDim.x) {
        uint32_t nz = 0;
        for (int m = 0; m < NM; ++m) {
            const uint64_t *q = s_m + ((size_t)g * NM + m) * 4;
            if (q[0] | q[1] | q[2] | q[3]) nz |= 1u << m;
        }
        s_nz[g] = nz;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = P.ld / TC;
    const long long t0 = (n_tiles * (long long)blockIdx.x) / gridDim.x;
    const long long t1 = (n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
    const int nb = (n_act + R - 1) / R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == CW) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = t0; tile < t1; ++tile) {
                const uint64_t *src = P.mat + tile * TC;
                for (int b = 0; b < nb; ++b) {
                    const int rows_here = min(R, n_act - b * R);
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, (uint32_t)rows_here * ROW_BYTES);
                    unsigned char *dst = ring + (size_t)s * R * ROW_BYTES;
                    for (int r = 0; r < rows_here; ++r)
                        tma_bulk_g2s(dst + r * ROW_BYTES, src + (long long)s_row[b * R + r] * P.ld, ROW_BYTES, full + s);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    int s = 0;
    uint32_t ph = 0;
    for (long long tile = t0; tile < t1; ++tile) {
        uint32_t cx[NM], cy[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) cx[m] = cy[m] = 0;
        for (int b = 0; b < nb; ++b) {
            const int rows_here = min(R, n_act - b * R);
            mbar_wait(full + s, ph);
            const unsigned char *st = ring + (size_t)s * R * ROW_BYTES + threadIdx.x * 16;
#pragma unroll 1
            for (int g = 0; g * 4 < rows_here; ++g) {
                const int grp = ((b * R) >> 2) + g;
                const uint32_t nz = s_nz[grp];
                const ulonglong2 *mk = reinterpret_cast<const ulonglong2 *>(s_m + (size_t)grp * NM * 4);
                const unsigned char *sg = st + g * 4 * ROW_BYTES;
                if (CSA && g * 4 + 4 <= rows_here) {
                    const ulonglong2 v0 = *reinterpret_cast<const ulonglong2 *>(sg);
                    const ulonglong2 v1 = *reinterpret_cast<const ulonglong2 *>(sg + ROW_BYTES);
                    const ulonglong2 v2 = *reinterpret_cast<const ulonglong2 *>(sg + 2 * ROW_BYTES);
                    const ulonglong2 v3 = *reinterpret_cast<const ulonglong2 *>(sg + 3 * ROW_BYTES);
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        if (nz & (1u << m)) {                      // warp-uniform
                            const ulonglong2 ma = mk[2 * m], mb = mk[2 * m + 1];
                            cx[m] += popc4x64_csa(v0.x & ma.x, v1.x & ma.y, v2.x & mb.x, v3.x & mb.y);
                            cy[m] += popc4x64_csa(v0.y & ma.x, v1.y & ma.y, v2.y & mb.x, v3.y & mb.y);
                        }
                    }
                } else {
                    const int rows_g = min(4, rows_here - g * 4);
                    const uint64_t *mk1 = reinterpret_cast<const uint64_t *>(mk);
                    for (int r = 0; r < rows_g; ++r) {
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sg + r * ROW_BYTES);
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            const uint64_t mm = mk1[m * 4 + r];
                            if (mm) {                              // warp-uniform
                                cx[m] += __popcll(v.x & mm);
                                cy[m] += __popcll(v.y & mm);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if (++s == S) { s = 0; ph ^= 1; }
        }
        const long long col = tile * TC + 2 * (long long)threadIdx.x;
#pragma unroll
        for (int m = 0; m < NM; ++m)
            if (P.cnt[m]) *reinterpret_cast<uint2 *>(P.cnt[m] + col) = make_uint2(cx[m], cy[m]);
    }
}

// ------------------------------------------------------------------------------------------------
// k_gini: split score of every presence rule from the per-class counts (cart.py:85-161)
// ------------------------------------------------------------------------------------------------
struct GiniParams {
    const uint32_t *cnt[MAX_CLASSES];
    double prior[MAX_CLASSES], n_total[MAX_CLASSES], n_node[MAX_CLASSES];
    int n_classes;
    long long n_cols, col0;
    const uint32_t *black_pres;
    unsigned long long *min_enc;    // MODE_MIN
    double min_score;               // MODE_TIES
    const double *min_ptr;          // MODE_TIES: read the minimum from device memory instead (fused call)
    const unsigned long long *segmin;   // MODE_TIES: enc_f64 minima of the 64-column segments, or null
    long long *out_idx;
    unsigned long long *counter;
    long long cap;
    double *scores;                 // MODE_SCORES
    // two classes, small node: score of every (count 0, count 1) pair, [n_node0 + 1][table_stride], or null
    const double *table;
    int table_stride;
    int table_global;               // the table is too large for shared memory: looked up where it is (L2-resident)
    // occurrence tiebreak on the device (experiment_cart.py:82-94): k-mer occurrence counts of the tree's training
    // set, and the maximum occurrence among the node's ties
    const uint32_t *occ;
    unsigned int *occ_max;
};
enum { GINI_MIN = 0, GINI_TIES = 1, GINI_SCORES = 2 };

// _gini_impurity(n, multiply_by_node_proba=True), cart.py:99-110, in numpy's operation order:
//   p_j_t = ((1.0 * prior) * n) / N ; p_t = 0 + p_0 + p_1 ... ; q_j = p_j_t / p_t ;
//   div = 0 + sum_{i != j} q_i * q_j (i outer, j inner) ; result = div * p_t
template <int NC>
__device__ __forceinline__ double gini_child(const double *n, const GiniParams &P, int nc_rt) {
    const int nc = NC > 0 ? NC : nc_rt;
    double q[NC > 0 ? NC : MAX_CLASSES];
    double pt = 0.0;
#pragma unroll
    for (int c = 0; c < nc; ++c) {
        q[c] = __ddiv_rn(__dmul_rn(P.prior[c], n[c]), P.n_total[c]);
        pt = (c == 0) ? q[c] : __dadd_rn(pt, q[c]);
    }
#pragma unroll
    for (int c = 0; c < nc; ++c) q[c] = __ddiv_rn(q[c], pt);
    double div = 0.0;
    bool first = true;
#pragma unroll
    for (int i = 0; i < nc; ++i) {
#pragma unroll
        for (int j = 0; j < nc; ++j) {
            if (i == j) continue;
            const double t = __dmul_rn(q[i], q[j]);
            div = first ? t : __dadd_rn(div, t);
            first = false;
        }
    }
    return __dmul_rn(div, pt);
}

template <int NC>
__device__ __forceinline__ double gini_score(long long col, const GiniParams &P) {
    const int nc = NC > 0 ? NC : P.n_classes;
    double l[NC > 0 ? NC : MAX_CLASSES], r[NC > 0 ? NC : MAX_CLASSES];
    double suml = 0.0, sumr = 0.0;
#pragma unroll
    for (int c = 0; c < nc; ++c) {
        l[c] = (double)P.cnt[c][col];                  // examples WITH the k-mer (cart.py:129-131)
        r[c] = P.n_node[c] - l[c];                     // exact (integers)
        suml += l[c];
        sumr += r[c];
    }
    if (P.black_pres && ((P.black_pres[col >> 5] >> (col & 31)) & 1u)) return INFINITY;   // cart.py:231
    if (suml == 0.0 || sumr == 0.0) return INFINITY;                                      // cart.py:158-159
    return __dadd_rn(gini_child<NC>(l, P, nc), gini_child<NC>(r, P, nc));
}

template <int NC, int MODE>
__global__ void __launch_bounds__(256) k_gini(const GiniParams P) {
    __shared__ double s_red[8];
    const long long stride = (long long)gridDim.x * blockDim.x;
    double mn = INFINITY;
    double target = 0.0;
    if (MODE == GINI_TIES) {
        target = P.min_score;
        if (P.min_ptr) {                                   // enc_f64 image left by the scan kernel
            const unsigned long long v = *reinterpret_cast<const unsigned long long *>(P.min_ptr);
            target = (v == ~0ull) ? INFINITY : dec_f64_max_dev(v);
        }
        if (target == INFINITY) return;                    // no admissible split: nothing ties (cart.py:234-235)
    }
    for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < P.n_cols; col += stride) {
        if (MODE == GINI_TIES && P.segmin) {
            // the segment's minimum is exact, so a segment whose minimum differs holds no tie (cart.py:238)
            const unsigned long long v = P.segmin[col >> 6];
            if (v == ~0ull || dec_f64_max_dev(v) != target) continue;
        }
        const double s = gini_score<NC>(col, P);
        if (MODE == GINI_MIN) {
            mn = fmin(mn, s);      // NaN-ignoring; the reference's scores hold no NaN for valid priors
        } else if (MODE == GINI_TIES) {
            if (s == target) {
                const unsigned long long slot = atomicAdd(P.counter, 1ull);
                if ((long long)slot < P.cap) P.out_idx[slot] = P.col0 + col;
            }
        } else {
            P.scores[col] = s;
        }
    }
    if (MODE == GINI_MIN) {
        mn = warp_min(mn);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mn;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) mn = fmin(mn, s_red[w]);
            atomicMin(P.min_enc, enc_f64(mn));
        }
    }
}

// A node with few examples has few distinct (count 0, count 1) pairs: the split score -- 8 IEEE double divisions
// per column -- is then tabulated once per node (k_gini2_table) and looked up.  Same values bit for bit: the table
// entries are computed by the same gini_score arithmetic.
static constexpr int GINI2_TABLE_MAX = 4096;
// Larger nodes: the table stays in device memory (a few MB, L2-resident) as long as tabulating is cheaper than scoring
// every column -- at most GINI2_GTABLE_MAX pairs and at most half as many pairs as columns.
static constexpr long long GINI2_GTABLE_MAX = 4ll << 20;

__global__ void __launch_bounds__(256) k_gini2_table(const GiniParams P, double *table) {
    const long long n0 = (long long)P.n_node[0] + 1, stride = P.table_stride;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n0 * stride; e += (long long)gridDim.x * blockDim.x) {
        const double l[2] = {(double)(e / stride), (double)(e % stride)};
        const double r[2] = {P.n_node[0] - l[0], P.n_node[1] - l[1]};
        double v = INFINITY;                                           // cart.py:158-159
        if (l[0] + l[1] != 0.0 && r[0] + r[1] != 0.0) v = __dadd_rn(gini_child<2>(l, P, 2), gini_child<2>(r, P, 2));
        table[e] = v;
    }
}

__device__ __forceinline__ void gini2_table_load(const GiniParams &P, double *s_tab) {
    if (P.table && !P.table_global) {
        const int n = ((int)P.n_node[0] + 1) * P.table_stride;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s_tab[e] = P.table[e];
        __syncthreads();
    }
}

__device__ __forceinline__ double gini2_score_any(long long col, const GiniParams &P, const double *s_tab) {
    if (!P.table) return gini_score<2>(col, P);
    if (P.black_pres && ((P.black_pres[col >> 5] >> (col & 31)) & 1u)) return INFINITY;   // cart.py:231
    const uint32_t e = P.cnt[0][col] * (uint32_t)P.table_stride + P.cnt[1][col];
    return P.table_global ? __ldg(P.table + e) : s_tab[e];
}

// the same score from counts already in registers (lets the callers issue the loads of several columns up front)
__device__ __forceinline__ double gini2_score_cnt(uint32_t c0, uint32_t c1, bool blacklisted, const GiniParams &P,
                                                  const double *s_tab) {
    if (blacklisted) return INFINITY;                                                      // cart.py:231
    if (P.table) {
        const uint32_t e = c0 * (uint32_t)P.table_stride + c1;
        return P.table_global ? __ldg(P.table + e) : s_tab[e];
    }
    const double l[2] = {(double)c0, (double)c1};
    const double r[2] = {P.n_node[0] - l[0], P.n_node[1] - l[1]};
    if (l[0] + l[1] == 0.0 || r[0] + r[1] == 0.0) return INFINITY;                         // cart.py:158-159
    return __dadd_rn(gini_child<2>(l, P, 2), gini_child<2>(r, P, 2));
}

// Tie collection when the scan left exact per-segment minima: one lane screens one 64-column segment
// (156 k segments for 10 M k-mers), the warp then rescans only the segments whose minimum IS the target.
// OCC = 0: every tie.  OCC = 1 / 2, the reference's CART tiebreaker on the device: among the ties keep the k-mers whose
// occurrence count in the tree's training set is isclose to the largest one (experiment_cart.py:82-94,
// numpy.isclose(o, o.max()): |o - max| <= 1e-8 + 1e-5 * |max|) -- pass 1 finds the maximum, pass 2 emits the survivors.
// A node with few examples ties on millions of k-mers, of which the tiebreaker keeps a handful: filtering here
// removes their sort and their trip to the host.
template <int OCC, int NSEG>
__device__ __forceinline__ void gini2_ties_walk(const GiniParams &P, const double *s_tab) {
    double target = P.min_score;
    if (P.min_ptr) {
        const unsigned long long v = *reinterpret_cast<const unsigned long long *>(P.min_ptr);
        target = (v == ~0ull) ? INFINITY : dec_f64_max_dev(v);
    }
    if (target == INFINITY) return;
    double occ_best = 0.0, occ_tol = 0.0;
    if (OCC == 2) {
        occ_best = (double)*P.occ_max;
        occ_tol = __dadd_rn(1e-8, __dmul_rn(1e-5, fabs(occ_best)));
    }
    unsigned int occ_run = 0;
    const int lane = threadIdx.x & 31;
    const long long n_seg = (P.n_cols + 63) >> 6;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long sb = warp0 * 32; sb < n_seg; sb += n_warps * 32) {
        const long long g = sb + lane;
        bool cand = false;
        if (g < n_seg) {
            const unsigned long long v = P.segmin[g];
            cand = v != ~0ull && dec_f64_max_dev(v) == target;
        }
        unsigned todo = __ballot_sync(0xffffffffu, cand);
        // NSEG candidate segments (2 NSEG columns per lane) at a time, every load issued before the first use: a node with
        // few examples ties on most of the k-mers and this loop then walks nearly every segment
        while (todo) {
            constexpr int NE = 2 * NSEG;
            long long col[NE];
            uint32_t c0[NE], c1[NE], oc[NE], bw[NE];
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const int src = todo ? __ffs(todo) - 1 : -1;
                todo &= todo - 1;                       // (0 & anything stays 0)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const long long c = ((sb + src) << 6) + lane + 32 * j;
                    col[q * 2 + j] = (src >= 0 && c < P.n_cols) ? c : -1;
                }
            }
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const bool v = col[e] >= 0;
                c0[e] = v ? P.cnt[0][col[e]] : 0u;
                c1[e] = v ? P.cnt[1][col[e]] : 0u;
                bw[e] = (v && P.black_pres) ? P.black_pres[col[e] >> 5] : 0u;
                oc[e] = (OCC != 0 && v) ? P.occ[col[e]] : 0u;
            }
            unsigned mine = 0;                          // bit e: column e of this lane is a tie (that survives)
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (col[e] < 0) continue;
                bool hit = gini2_score_cnt(c0[e], c1[e], (bw[e] >> (col[e] & 31)) & 1u, P, s_tab) == target;
                if (OCC == 1) {
                    if (hit) occ_run = max(occ_run, oc[e]);
                    continue;
                }
                if (OCC == 2 && hit) hit = fabs(__dsub_rn((double)oc[e], occ_best)) <= occ_tol;
                if (hit) mine |= 1u << e;
            }
            if (OCC == 1) continue;
            // one atomic per warp and group
            const unsigned n_mine = __popc(mine);
            unsigned incl = n_mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.counter, (unsigned long long)total);
                base = __shfl_sync(0xffffffffu, base, 0) + (incl - n_mine);
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if ((mine >> e) & 1u) {
                        if ((long long)base < P.cap) P.out_idx[base] = P.col0 + col[e];
                        ++base;
                    }
            }
        }
    }
    if (OCC == 1) {
        occ_run = __reduce_max_sync(0xffffffffu, occ_run);
        if (lane == 0 && occ_run) atomicMax(P.occ_max, occ_run);
    }
}

// small nodes (score table; they are the ones that tie on millions of k-mers): four segments per step; large nodes score
// with eight IEEE divisions per column and tie on few: one segment per step keeps the code small
template <int OCC>
__device__ __forceinline__ void gini2_ties_body(const GiniParams &P, double *s_tab) {
    gini2_table_load(P, s_tab);
    if (P.table) gini2_ties_walk<OCC, 4>(P, s_tab);
    else gini2_ties_walk<OCC, 1>(P, s_tab);
}

__global__ void __launch_bounds__(256) k_gini2_ties_seg(const GiniParams P) {
    __shared__ double s_tab[GINI2_TABLE_MAX];
    gini2_ties_body<0>(P, s_tab);
}
__global__ void __launch_bounds__(256) k_gini2_ties_occmax(const GiniParams P) {
    __shared__ double s_tab[GINI2_TABLE_MAX];
    gini2_ties_body<1>(P, s_tab);
}
__global__ void __launch_bounds__(256) k_gini2_ties_occ(const GiniParams P) {
    __shared__ double s_tab[GINI2_TABLE_MAX];
    gini2_ties_body<2>(P, s_tab);
}

// Two-class Gini scores from RESIDENT counts (the nodes of a CART level are counted several per matrix pass by
// k_count_tma): the minimum split score of the node and the exact minima of the 64-column segments, i.e. what
// k_scan_tma<EPI_GINI2> leaves behind, at 8 bytes per column.  One warp per segment: lane l scores columns l, l+32.
__global__ void __launch_bounds__(256) k_gini2_counts(const GiniParams P, unsigned long long *segmin_out) {
    __shared__ unsigned long long s_red[8];
    __shared__ double s_tab[GINI2_TABLE_MAX];
    gini2_table_load(P, s_tab);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_seg = (P.n_cols + 63) >> 6;
    const long long n_span = (P.n_cols + 255) >> 8;                    // 256 columns = 4 segments per warp and iteration
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    double mn = INFINITY;
    // lane l: columns 4l .. 4l+3 of the span's first and second half (two 128-bit loads per count array, all four issued
    // before the first score: the kernel used to be bound by the latency of one dependent load chain per segment); the
    // count arrays are [ld] long, ld a multiple of 1024, so the vector loads stay inside them
    for (long long sp = warp0; sp < n_span; sp += n_warps) {
        const long long cbase[2] = {(sp << 8) + 4 * lane, (sp << 8) + 128 + 4 * lane};
        uint4 a[2], b[2];
        uint32_t bl[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            a[h] = *reinterpret_cast<const uint4 *>(P.cnt[0] + cbase[h]);
            b[h] = *reinterpret_cast<const uint4 *>(P.cnt[1] + cbase[h]);
            bl[h] = P.black_pres ? (P.black_pres[cbase[h] >> 5] >> (cbase[h] & 31)) & 15u : 0u;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t c0[4] = {a[h].x, a[h].y, a[h].z, a[h].w}, c1[4] = {b[h].x, b[h].y, b[h].z, b[h].w};
            double sm = INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (cbase[h] + q < P.n_cols) sm = fmin(sm, gini2_score_cnt(c0[q], c1[q], (bl[h] >> q) & 1u, P, s_tab));
            mn = fmin(mn, sm);
            unsigned long long seg = enc_f64(sm);                          // minimum of the 16 lanes that share a segment
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, seg, d);
                seg = o < seg ? o : seg;
            }
            const long long g = (sp << 2) + 2 * h + (lane >> 4);
            if ((lane & 15) == 0 && g < n_seg) segmin_out[g] = seg;
        }
    }
    const unsigned long long v = warp_min_enc(enc_f64(mn));
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = s_red[0];
        for (int w = 1; w < 8; ++w) m = s_red[w] < m ? s_red[w] : m;
        atomicMin(P.min_enc, m);
    }
}

// class counts of a tree node = its parent's - its sibling's (the children partition the parent's examples)
__global__ void __launch_bounds__(256) k_sub_counts(const uint4 *__restrict__ par0, const uint4 *__restrict__ par1,
                                                    const uint4 *__restrict__ sib0, const uint4 *__restrict__ sib1,
                                                    uint4 *__restrict__ out0, uint4 *__restrict__ out1, long long n4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uint4 a = par0[i], b = sib0[i], c = par1[i], d = sib1[i];
        out0[i] = make_uint4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        out1[i] = make_uint4(c.x - d.x, c.y - d.y, c.z - d.z, c.w - d.w);
    }
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
struct InlineCols {
    long long c[8];
};
__global__ void k_gather_cols(const uint64_t *mat, long long ld, long long n_words, const long long *cols,
                              const InlineCols inl, long long n, uint64_t *out) {
    const long long j = blockIdx.x;
    const long long col = cols ? cols[j] : inl.c[j];          // up to 8 columns ride in the parameters
    for (long long w = threadIdx.x; w < n_words; w += blockDim.x) out[j * n_words + w] = mat[w * ld + col];
}

// 32-bit packed rows -> halves of the 64-bit image (rules.py:123-125): row 2w is the high half
__global__ void k_merge32(uint64_t *mat, long long ld, long long lcol0, long long row0_32, long long n_rows,
                          long long n_cols, const uint32_t *src, long long src_ld) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    for (long long r = 0; r < n_rows; ++r) {
        const long long r32 = row0_32 + r;
        const uint64_t v = (uint64_t)src[r * src_ld + c] << ((r32 & 1) ? 0 : 32);
        uint64_t *dst = mat + (r32 >> 1) * ld + lcol0 + c;
        const uint64_t keep = (r32 & 1) ? 0xFFFFFFFF00000000ull : 0x00000000FFFFFFFFull;
        *dst = (*dst & keep) | v;
    }
}

template <typename T>
__global__ void k_popc_inplace(T *arr, long long m, long long n, const T *mask) {
    const long long total = m * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const T v = arr[i];
        if (v != 0) {                                            // popcount.pyx:47-48 / 93-94
            const T x = v & mask[i / n];
            arr[i] = sizeof(T) == 8 ? (T)__popcll((unsigned long long)x) : (T)__popc((unsigned int)x);
        }
    }
}

// presence counts -> the reference's result dtype, plus the absence half len(rows) - presence computed in
// that dtype (rules.py:265: unsigned wrap-around included)
template <typename T>
__global__ void k_narrow2(const uint32_t *src, T *pres, T *absn, long long n, unsigned long long n_rows) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T v = (T)src[i];
        pres[i] = v;
        absn[i] = (T)((T)n_rows - v);
    }
}

// dataset/split.py:178-188: a k-mer's individual risk depends only on its integer error count
// e = (n_pos - pos_present) + neg_present, so the device produces e per k-mer plus the set of values that
// occur; the host evaluates the reference's float formula on those <= n+1 values and sends back a table.
__global__ void k_risk_errors(uint32_t *cnt_pos_inout, const uint32_t *cnt_neg, long long n, uint32_t n_pos,
                              uint8_t *present) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t e = (n_pos - cnt_pos_inout[i]) + cnt_neg[i];
        cnt_pos_inout[i] = e;
        present[e] = 1;                  // benign race: every writer stores 1
    }
}
template <typename T>
__global__ void k_risk_map(const uint32_t *e, const uint32_t *lut_kmer, const uint32_t *lut_anti, T *out_kmer,
                           T *out_anti, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t v = e[i];
        out_kmer[i] = (T)lut_kmer[v];
        out_anti[i] = (T)lut_anti[v];
    }
}

template <typename T>
__global__ void k_narrow(const uint32_t *src, T *dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (T)src[i];
}

// Synthetic matrix: see kover_b200/synthetic.py (numpy twin of exactly this function).
__device__ __host__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t gcol, uint64_t w) {
    const unsigned level = (unsigned)(splitmix64(seed ^ (gcol * 0xD6E8FEB86659FD93ull)) & 7ull);
    const uint64_t base = seed + ((gcol << 13) + w) * 4ull;
    const uint64_t r1 = splitmix64(base), r2 = splitmix64(base + 1), r3 = splitmix64(base + 2);
    switch (level) {
        case 0: return r1 & r2 & r3;                       // 1/8
        case 1: return r1 & r2;                            // 1/4
        case 2: return r1 & (r2 | r3);                     // 3/8
        case 3: return r1;                                 // 1/2
        case 4: return r1 | (r2 & r3);                     // 5/8
        case 5: return r1 | r2;                            // 3/4
        case 6: return r1 | r2 | r3;                       // 7/8
        default: return r1 & r2 & r3 & splitmix64(base + 3);   // 1/16
    }
}
__global__ void k_generate(uint64_t *mat, long long ld, long long n_words, long long n_cols, long long col0,
                           uint64_t seed, uint64_t last_word_mask) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = blockIdx.y;
    if (c >= n_cols) return;
    uint64_t v = synth_word(seed, (uint64_t)(col0 + c), (uint64_t)w);
    if (w == n_words - 1) v &= last_word_mask;
    mat[w * ld + c] = v;
}

// ------------------------------------------------------------------------------------------------
// k_mask_and_column: the greedy step's example bookkeeping on the device (scm.py:133-139).  After a rule is picked
// the reference fetches its column and keeps the examples the rule accepts; here the shard keeps the remaining
// positive / negative example masks resident and ANDs them with the chosen column (presence rule: mask &= column,
// absence rule: mask &= ~column), counts what is left, and rebuilds the compacted active-row records the scan
// kernel consumes (rows with only mask 0, only mask 1, both -- ascending row order inside each group).  One CTA:
// W is a few hundred to a few thousand words.  The counts (and the column words, for the other shards / ranks)
// go to mapped host memory; the host polls the epoch word.
// ------------------------------------------------------------------------------------------------
struct MaskColParams {
    const uint64_t *mat;             // column source: mat[w * ld + col] when col >= 0,
    long long ld, col;
    const uint64_t *words;           // else these W words (device), or null = all ones (state initialisation)
    int invert;
    long long W;
    uint64_t *pos, *neg;             // resident masks [W], updated in place
    uint64_t *rec_m0, *rec_m1;       // records out [W]
    uint32_t *rec_row;
    volatile unsigned long long *host;   // mapped: [0] n_pos [1] n_neg [2] n_act [3] n_both [4] epoch, [8 ..] column words
    unsigned long long epoch;
};

__global__ void __launch_bounds__(1024) k_mask_and_column(const MaskColParams P) {
    __shared__ unsigned int s_cnt[5][32];          // per warp: pos bits, neg bits, rows A, rows B, rows C
    __shared__ unsigned int s_tot[5];
    __shared__ unsigned int s_run[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int c_pos = 0, c_neg = 0, nA = 0, nB = 0, nC = 0;
    for (long long w = tid; w < P.W; w += 1024) {
        uint64_t c = ~0ull;
        if (P.col >= 0) c = P.mat[w * P.ld + P.col];
        else if (P.words) c = P.words[w];
        if (P.col >= 0 || P.words) P.host[8 + w] = c;                  // the raw column, for the other shards
        if (P.invert) c = ~c;
        const uint64_t a = P.pos[w] & c, b = P.neg[w] & c;
        P.pos[w] = a;
        P.neg[w] = b;
        c_pos += __popcll(a);
        c_neg += __popcll(b);
        nA += (a != 0 && b == 0);
        nB += (a == 0 && b != 0);
        nC += (a != 0 && b != 0);
    }
    unsigned int v[5] = {c_pos, c_neg, nA, nB, nC};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        v[q] = __reduce_add_sync(0xffffffffu, v[q]);
        if (lane == 0) s_cnt[q][warp] = v[q];
    }
    __syncthreads();
    if (tid < 5) {
        unsigned int t = 0;
        for (int w = 0; w < 32; ++w) t += s_cnt[tid][w];
        s_tot[tid] = t;
    }
    if (tid < 3) s_run[tid] = 0;
    __syncthreads();
    const unsigned int base[3] = {0u, s_tot[2], s_tot[2] + s_tot[3]};
    // second pass: positions inside the three groups, 1024 rows at a time, ascending rows
    for (long long w0 = 0; w0 < P.W; w0 += 1024) {
        const long long w = w0 + tid;
        uint64_t a = 0, b = 0;
        if (w < P.W) {
            a = P.pos[w];
            b = P.neg[w];
        }
        const int cat = (a != 0 && b != 0) ? 2 : (a != 0 ? 0 : (b != 0 ? 1 : -1));
        unsigned int before[3], wtot[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const unsigned int m = __ballot_sync(0xffffffffu, cat == q);
            before[q] = __popc(m & ((1u << lane) - 1u));
            wtot[q] = __popc(m);
            if (lane == 0) s_cnt[q][warp] = wtot[q];
        }
        __syncthreads();
        if (cat >= 0) {
            unsigned int off = s_run[cat];
            for (int ww = 0; ww < warp; ++ww) off += s_cnt[cat][ww];
            const unsigned int i = base[cat] + off + before[cat];
            P.rec_m0[i] = a;
            P.rec_m1[i] = b;
            P.rec_row[i] = (uint32_t)w;
        }
        __syncthreads();
        if (tid < 3) {
            unsigned int t = 0;
            for (int ww = 0; ww < 32; ++ww) t += s_cnt[tid][ww];
            s_run[tid] += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        P.host[0] = s_tot[0];
        P.host[1] = s_tot[1];
        P.host[2] = s_tot[2] + s_tot[3] + s_tot[4];
        P.host[3] = s_tot[4];
        __threadfence_system();
        P.host[4] = P.epoch;
    }
}

// ------------------------------------------------------------------------------------------------
// shard
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

// Page-locked staging buffers outlive the shard that first needed them: pinning host memory costs ~1 ms per MB
// (70 ms for the four 16 MB slots of the deflate upload, measured), which a process that opens one dataset after the
// other (or one shard per device) would pay every time.  Buffers are handed back at kv_destroy and kept for the next
// shard, up to kPinnedCacheBytes in total; portable, so any device's streams may use them.
static constexpr size_t kPinnedCacheBytes = (size_t)512 << 20;
struct PinnedCache {
    struct Ent {
        void *p;
        size_t bytes;
        bool busy;
    };
    std::mutex mu;
    std::vector<Ent> ents;
};
static PinnedCache g_pinned;

static cudaError_t pinned_borrow(void **out, size_t bytes) {
    {
        std::lock_guard<std::mutex> lock(g_pinned.mu);
        PinnedCache::Ent *best = nullptr;
        for (auto &e : g_pinned.ents)
            if (!e.busy && e.bytes >= bytes && (!best || e.bytes < best->bytes)) best = &e;
        if (best) {
            best->busy = true;
            *out = best->p;
            return cudaSuccess;
        }
    }
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_pinned.mu);
    g_pinned.ents.push_back({p, bytes, true});
    *out = p;
    return cudaSuccess;
}

static void pinned_return(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pinned.mu);
    size_t idle = 0;
    for (auto &e : g_pinned.ents)
        if (!e.busy) idle += e.bytes;
    for (size_t i = 0; i < g_pinned.ents.size(); ++i) {
        if (g_pinned.ents[i].p != p) continue;
        if (idle + g_pinned.ents[i].bytes > kPinnedCacheBytes) {
            cudaFreeHost(p);
            g_pinned.ents.erase(g_pinned.ents.begin() + (long)i);
        } else {
            g_pinned.ents[i].busy = false;
        }
        return;
    }
    cudaFreeHost(p);                             // (not one of ours)
}

// The same for the device scratch of the deflate upload (compressed slabs, job lists): cudaMalloc of a few hundred MB
// costs 2 ms on some boxes and 20 ms on others (measured: `setup` of kv_upload_deflate), as much as inflating a sparse
// matrix takes.  Per device; at most kDevCacheBytes idle in total and kDevCacheEntry per buffer -- larger ones are freed.
static constexpr size_t kDevCacheBytes = (size_t)3 << 30, kDevCacheEntry = (size_t)1 << 30;
struct DevCache {
    struct Ent {
        void *p;
        size_t bytes;
        int device;
        bool busy;
    };
    std::mutex mu;
    std::vector<Ent> ents;
};
static DevCache g_devcache;

static cudaError_t dev_borrow(void **out, size_t bytes, int device) {
    {
        std::lock_guard<std::mutex> lock(g_devcache.mu);
        DevCache::Ent *best = nullptr;
        for (auto &e : g_devcache.ents)
            if (!e.busy && e.device == device && e.bytes >= bytes && e.bytes <= 2 * bytes + (1 << 20) &&
                (!best || e.bytes < best->bytes))
                best = &e;
        if (best) {
            best->busy = true;
            *out = best->p;
            return cudaSuccess;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {                      // memory pressure: give the idle cache back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lock(g_devcache.mu);
            for (size_t i = g_devcache.ents.size(); i-- > 0;)
                if (!g_devcache.ents[i].busy && g_devcache.ents[i].device == device) {
                    cudaFree(g_devcache.ents[i].p);
                    g_devcache.ents.erase(g_devcache.ents.begin() + (long)i);
                }
        }
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lock(g_devcache.mu);
    g_devcache.ents.push_back({p, bytes, device, true});
    *out = p;
    return cudaSuccess;
}

static void dev_return(void *p) {                // (the caller has synchronised the streams that used the buffer)
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_devcache.mu);
    size_t idle = 0;
    for (auto &e : g_devcache.ents)
        if (!e.busy) idle += e.bytes;
    for (size_t i = 0; i < g_devcache.ents.size(); ++i) {
        if (g_devcache.ents[i].p != p) continue;
        if (g_devcache.ents[i].bytes > kDevCacheEntry || idle + g_devcache.ents[i].bytes > kDevCacheBytes) {
            cudaFree(p);
            g_devcache.ents.erase(g_devcache.ents.begin() + (long)i);
        } else {
            g_devcache.ents[i].busy = false;
        }
        return;
    }
    cudaFree(p);
}

struct kv_p2p;
static void p2p_destroy(kv_p2p *m);
struct kv_shard;
static int finish_scan_timing(kv_shard *s);
struct Scan2Params;
static void fill_scan2_state(kv_shard *s, Scan2Params &P);

struct kv_shard {
    kv_p2p *p2p = nullptr;                    // peer mailbox (multi-rank), optional
    int device = 0, sm_count = 0;
    long long n_examples = 0, W = 0, k_total = 0, col0 = 0, n_cols = 0, ld = 0, n_tiles = 0;
    uint64_t *d_mat = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    int scan_ctas_per_sm = 4;
    int scan_csa = 1;                         // KOVER_B200_SCAN_CSA=0: plain popcounts in k_scan_tma
    int l2_hints = 0;                         // KOVER_B200_L2HINT (bit 0 loads evict-first, bit 1 count stores evict-last)
    int scan_variant = -1;                    // -1: choose per call; >= 0: forced (KOVER_B200_SCAN_VARIANT)

    std::vector<uint32_t *> d_cnt;            // count arrays, [ld] each
    DevBuf d_rec, d_misc, d_ties, d_stage, d_sort_tmp, d_sort_io, d_queue;
    void *h_pin = nullptr;                    // pinned staging (records, block tables, small results)
    size_t h_pin_bytes = 0;
    void *h_up[2] = {nullptr, nullptr};       // pinned upload ring
    cudaEvent_t ev_up[2] = {nullptr, nullptr};
    size_t h_up_bytes = 0;
    int up_slot = 0;

    uint32_t *d_black_pres = nullptr, *d_black_abs = nullptr;
    unsigned long long *d_segmax = nullptr;   // [2][ld / 64] segment maxima left by the last SCM scan
    bool have_black = false;

    // state left by the last scan for phase 2
    bool scm_valid = false, scm_swapped = false, reuse_armed = false, reuse_swap = false;
    const uint32_t *scm_cnt_a = nullptr, *scm_cnt_b = nullptr;   // presence counts under the scanned pair's first / second mask
    long long scm_n_pos = 0, scm_n_neg = 0, scm_ubs = 0;
    double scm_p = 0.0;
    bool gini_valid = false, risk_valid = false;
    long long risk_n = 0;
    GiniParams gini;

    DevBuf d_qmisc, d_qseg;                   // block tables / segment maxima of the jobs of one queue batch
    long long misc_n_blocks = -1, qmisc_n_blocks = -1;   // block count d_misc / d_qmisc are currently laid out for

    // fused single-launch SCM step (FusedTail)
    int fused = 1;                            // KOVER_B200_FUSED=0 switches back to scan + select + ties launches
    int cooperative = 1;                      // KOVER_B200_COOP=0: plain launch (the grid never exceeds the resident CTAs)
    void *h_fused = nullptr;                  // mapped page-locked: 64-byte header + kFusedCap tie records
    unsigned long long *d_fused_ctr = nullptr;   // [0] grid barrier, [16] done tickets, [32] tie counter (128 B apart)
    unsigned long long *d_fused_bm = nullptr;    // block-maximum accumulators, two parities of fused_n_blocks
    long long fused_n_blocks = -1;
    unsigned long long fused_epoch = 0, fused_barrier_total = 0, fused_done_total = 0;
    std::vector<uint64_t> h_masks;            // masks built from index arrays (kv_scm_best_rules_idx)
    // candidate-streaming step (EPI_SCM_STREAM): no count arrays are written, so what later phases read from them
    // (kv_scm_reuse_counts, kv_scm_ties) is produced on demand by one more scan of the same masks (scm_materialize)
    int stream_cand = 1;                      // KOVER_B200_STREAM=0: always write the count arrays
    int stream_backoff = 0, stream_backoff_len = 8;   // steps left on the counting path after a candidate overflow
    TieRec *d_cand = nullptr;                 // [2 * sm_count][kCandCap]
    unsigned *d_cand_n = nullptr;             // [2 * sm_count]
    bool scm_lazy = false, lazy_from_state = false;
    std::vector<uint64_t> lazy_masks;         // pos [W] | neg [W] of the last streamed step
    long long stream_steps = 0, stream_redos = 0;
    // device-resident greedy state (kv_scm_state_set / kv_mask_and_column)
    uint64_t *d_state = nullptr;              // pos [W] | neg [W] | column words [W] | rec m0 [W] | rec m1 [W] | rec row u32 [W]
    void *h_state = nullptr;                  // mapped page-locked: 64-byte header + column words [W]
    unsigned long long state_epoch = 0;
    bool state_valid = false;
    long long state_n_pos = 0, state_n_neg = 0;
    int state_n_act = 0, state_n_both = 0;
    bool profiling = false;                   // CUDA events around the scan kernels (kv_set_profiling / KOVER_B200_PROFILE)
    // deflate upload: page-locked staging slots + copy stream (kv_upload_deflate)
    cudaStream_t inf_copy_stream = nullptr;
    void *inf_pin[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t inf_pin_free[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t inf_slot_bytes = 0;
    double up_setup_ms = 0.0, up_gather_ms = 0.0, up_total_ms = 0.0;   // phases of the last kv_upload_deflate (host clock)
    DevBuf d_gtables;                         // kv_gini_level: score tables of nodes too large for shared memory
    DevBuf d_level;                           // kv_gini_level: per-node minima, counters, tie slots, segment minima
    std::vector<uint32_t *> occ_tables;       // kv_occ_table_set: k-mer occurrence counts of a training set, [ld] each
    // kv_gini_level_family: the class counts of the previous tree level's nodes stay resident (slots of one arena) so
    // that of two sibling nodes only the smaller is counted and the other one is parent - sibling
    uint32_t *d_fam_arena = nullptr;          // [fam_slots][2][ld]
    int fam_slots = 0;
    bool fam_tried = false;
    std::vector<int> fam_free;
    struct FamNode { int slot; long long n[2]; };
    std::unordered_map<long long, FamNode> fam_prev;
    std::vector<long long> fam_key, fam_parent;   // armed for the next kv_gini_level
    long long level_nodes_counted = 0, level_nodes_derived = 0;
    int64_t *h_level = nullptr;               // pinned arena: the tie lists of the last kv_gini_level, node after node
    size_t h_level_cap = 0, h_level_used = 0; // in entries
    std::vector<std::pair<size_t, size_t>> level_span;   // (offset, count) of every node in the arena

    double exch_publish_us = 0.0, exch_wait_us = 0.0, exch_merge_us = 0.0;   // fused multi-rank tail, summed over steps
    long long exch_steps = 0;
    double last_kernel_ms = 0.0, scan_ms_total = 0.0;
    long long last_launches = 0, total_launches = 0, scan_calls = 0, rescore_calls = 0;
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int ensure(DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes) return KV_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CU(cudaMalloc(&b.p, want));
    CU(cudaMemset(b.p, 0, want));          // the block-maxima accumulators rely on starting zeroed
    CU(cudaDeviceSynchronize());           // (re)allocation is rare; the memset must not race the shard's stream
    b.bytes = want;
    return KV_OK;
}

static int ensure_pinned(kv_shard *s, size_t bytes) {
    if (s->h_pin_bytes >= bytes) return KV_OK;
    if (s->h_pin) cudaFreeHost(s->h_pin);
    s->h_pin = nullptr;
    s->h_pin_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 16);
    CU(cudaMallocHost(&s->h_pin, want));
    s->h_pin_bytes = want;
    return KV_OK;
}

static int ensure_counts(kv_shard *s, int n) {
    while ((int)s->d_cnt.size() < n) {
        uint32_t *p = nullptr;
        CU(cudaMalloc(&p, (size_t)s->ld * sizeof(uint32_t)));
        s->d_cnt.push_back(p);
    }
    return KV_OK;
}

static inline void count_launch(kv_shard *s, int n = 1) {
    s->last_launches += n;
    s->total_launches += n;
}

extern "C" int kv_create(int device, int64_t n_examples, int64_t n_kmers_total, int64_t col0, int64_t n_cols,
                         kv_shard **out) {
    if (!out) return fail(KV_E_INVALID, "kv_create: null out pointer");
    *out = nullptr;
    if (n_examples <= 0 || n_kmers_total <= 0 || n_cols <= 0 || col0 < 0 || col0 + n_cols > n_kmers_total)
        return fail(KV_E_INVALID, "kv_create: bad shape (n_examples, n_kmers, col0, n_cols)");
    if (2 * n_kmers_total >= (1ll << 52)) return fail(KV_E_INVALID, "kv_create: too many k-mers");
    if (n_examples > 0xFFFFFFFFll) return fail(KV_E_INVALID, "kv_create: too many examples");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(KV_E_CUDA, std::string("kv_create: no usable CUDA device (no CPU fallback exists): ") +
                                   cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(KV_E_INVALID, "kv_create: device ordinal out of range");
    DeviceGuard g(device);
    if (!g.ok) return fail(KV_E_CUDA, "kv_create: cudaSetDevice failed");
    kv_shard *s = new kv_shard();
    s->device = device;
    s->n_examples = n_examples;
    s->W = (n_examples + 63) / 64;
    s->k_total = n_kmers_total;
    s->col0 = col0;
    s->n_cols = n_cols;
    s->ld = (n_cols + LD_ALIGN - 1) / LD_ALIGN * LD_ALIGN;
    s->n_tiles = s->ld / TILE_COLS;
    if (const char *v = getenv("KOVER_B200_SCAN_VARIANT")) s->scan_variant = atoi(v);
    if (const char *v = getenv("KOVER_B200_L2HINT")) s->l2_hints = atoi(v);
    if (const char *v = getenv("KOVER_B200_SCAN_CSA")) s->scan_csa = atoi(v);
    if (const char *v = getenv("KOVER_B200_FUSED")) s->fused = atoi(v);
    if (const char *v = getenv("KOVER_B200_COOP")) s->cooperative = atoi(v);
    if (const char *v = getenv("KOVER_B200_STREAM")) s->stream_cand = atoi(v);
    if (const char *v = getenv("KOVER_B200_PROFILE")) s->profiling = atoi(v) != 0;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    const size_t bytes = (size_t)s->W * (size_t)s->ld * 8;
    if (cudaMalloc(&s->d_mat, bytes) != cudaSuccess) {
        delete s;
        char b[256];
        snprintf(b, sizeof b, "kv_create: cannot allocate %.2f GB of HBM for the packed matrix", bytes / 1e9);
        return fail(KV_E_NOMEM, b);
    }
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    CU(cudaEventCreate(&s->ev_t0));
    CU(cudaEventCreate(&s->ev_t1));
    CU(cudaMemsetAsync(s->d_mat, 0, bytes, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    *out = s;
    return KV_OK;
}

extern "C" int kv_destroy(kv_shard *s) {
    if (!s) return KV_OK;
    DeviceGuard g(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    p2p_destroy(s->p2p);
    cudaFree(s->d_mat);
    for (auto p : s->d_cnt) cudaFree(p);
    for (auto p : s->occ_tables)
        if (p) cudaFree(p);
    if (s->d_fam_arena) cudaFree(s->d_fam_arena);
    for (DevBuf *b : {&s->d_rec, &s->d_misc, &s->d_ties, &s->d_stage, &s->d_sort_tmp, &s->d_sort_io, &s->d_queue, &s->d_level, &s->d_gtables, &s->d_qmisc, &s->d_qseg})
        if (b->p) cudaFree(b->p);
    if (s->d_black_pres) cudaFree(s->d_black_pres);
    if (s->d_black_abs) cudaFree(s->d_black_abs);
    if (s->d_segmax) cudaFree(s->d_segmax);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    if (s->h_level) cudaFreeHost(s->h_level);
    if (s->h_fused) cudaFreeHost(s->h_fused);
    if (s->h_state) cudaFreeHost(s->h_state);
    for (int i = 0; i < 4; ++i) {
        pinned_return(s->inf_pin[i]);
        if (s->inf_pin_free[i]) cudaEventDestroy(s->inf_pin_free[i]);
    }
    if (s->inf_copy_stream) cudaStreamDestroy(s->inf_copy_stream);
    if (s->d_state) cudaFree(s->d_state);
    if (s->d_fused_ctr) cudaFree(s->d_fused_ctr);
    if (s->d_cand) cudaFree(s->d_cand);
    if (s->d_cand_n) cudaFree(s->d_cand_n);
    if (s->d_fused_bm) cudaFree(s->d_fused_bm);
    for (int i = 0; i < 2; ++i) {
        pinned_return(s->h_up[i]);
        if (s->ev_up[i]) cudaEventDestroy(s->ev_up[i]);
    }
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_t0) cudaEventDestroy(s->ev_t0);
    if (s->ev_t1) cudaEventDestroy(s->ev_t1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return KV_OK;
}

extern "C" int kv_get_info(kv_shard *s, kv_shard_info *o) {
    if (!s || !o) return fail(KV_E_INVALID, "kv_get_info: null pointer");
    o->device = s->device;
    o->sm_count = s->sm_count;
    o->n_examples = s->n_examples;
    o->n_words = s->W;
    o->n_kmers_total = s->k_total;
    o->col0 = s->col0;
    o->n_cols = s->n_cols;
    o->ld = s->ld;
    o->matrix_bytes = (int64_t)((size_t)s->W * (size_t)s->ld * 8);
    o->last_kernel_ms = s->last_kernel_ms;
    o->last_launches = s->last_launches;
    o->total_launches = s->total_launches;
    o->scan_ms_total = s->scan_ms_total;
    o->scan_calls = s->scan_calls;
    o->rescore_calls = s->rescore_calls;
    o->exch_steps = s->exch_steps;
    o->exch_publish_us = s->exch_publish_us;
    o->exch_wait_us = s->exch_wait_us;
    o->exch_merge_us = s->exch_merge_us;
    o->stream_steps = s->stream_steps;
    o->stream_redos = s->stream_redos;
    o->level_nodes_counted = s->level_nodes_counted;
    o->level_nodes_derived = s->level_nodes_derived;
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
static int ensure_upload_ring(kv_shard *s) {
    if (s->h_up[0]) return KV_OK;
    s->h_up_bytes = (size_t)32 << 20;
    if (const char *v = getenv("KOVER_B200_UPLOAD_SLOT_MB")) s->h_up_bytes = (size_t)std::max(1, atoi(v)) << 20;
    for (int i = 0; i < 2; ++i) {
        CU(pinned_borrow(&s->h_up[i], s->h_up_bytes));
        CU(cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming));
    }
    return KV_OK;
}

// Pageable host rows -> the pinned slot.  One memcpy stream tops out near 5 GB/s, a fraction of what the
// PCIe Gen5 link takes, so large slots are split over a few threads (byte ranges of the flattened slot).
static void stage_rows(char *pin, const char *src, long long n_rows, size_t row_bytes, size_t src_pitch) {
    const size_t total = (size_t)n_rows * row_bytes;
    auto copy_range = [&](size_t b0, size_t b1) {
        while (b0 < b1) {
            const size_t row = b0 / row_bytes, off = b0 % row_bytes;
            const size_t n = std::min(row_bytes - off, b1 - b0);
            memcpy(pin + b0, src + row * src_pitch + off, n);
            b0 += n;
        }
    };
    const unsigned hw = std::thread::hardware_concurrency();
    static const unsigned max_threads = getenv("KOVER_B200_UPLOAD_THREADS") ? (unsigned)atoi(getenv("KOVER_B200_UPLOAD_THREADS")) : 6u;
    const int n_threads = total < ((size_t)4 << 20) ? 1 : (int)std::max(1u, std::min(hw ? hw : 1u, std::max(1u, max_threads)));
    if (n_threads == 1) {
        copy_range(0, total);
        return;
    }
    std::vector<std::thread> pool;
    const size_t chunk = ((total + n_threads - 1) / n_threads + 63) / 64 * 64;
    for (int t = 0; t < n_threads; ++t) {
        const size_t b0 = std::min(total, (size_t)t * chunk), b1 = std::min(total, b0 + chunk);
        if (b0 < b1) pool.emplace_back(copy_range, b0, b1);
    }
    for (auto &th : pool) th.join();
}

extern "C" int kv_upload_block(kv_shard *s, int pack_bits, int64_t row0, int64_t n_rows, int64_t gcol0,
                               int64_t n_cols, const void *host, int64_t ld_elems) {
    if (!s || !host) return fail(KV_E_INVALID, "kv_upload_block: null pointer");
    if (pack_bits != 64 && pack_bits != 32)
        return fail(KV_E_INVALID, "Unsupported data type for packed attribute classifications array. The supported "
                                  "data types are np.uint32 and np.uint64.");
    const long long max_rows = pack_bits == 64 ? s->W : (s->n_examples + 31) / 32;
    if (row0 < 0 || n_rows < 0 || row0 + n_rows > max_rows || gcol0 < 0 || n_cols < 0 ||
        gcol0 + n_cols > s->k_total || ld_elems < n_cols)
        return fail(KV_E_INVALID, "kv_upload_block: block outside the packed matrix");
    // intersect with the shard's columns
    const long long a = std::max<long long>(gcol0, s->col0), b = std::min<long long>(gcol0 + n_cols, s->col0 + s->n_cols);
    if (a >= b || n_rows == 0) return KV_OK;
    const long long nc = b - a, src_off = a - gcol0, lcol0 = a - s->col0;
    DeviceGuard g(s->device);
    KV_TRY(ensure_upload_ring(s));
    s->last_launches = 0;
    const size_t esz = pack_bits / 8;
    // stream the block through the pinned ring in row groups
    const long long rows_per_slot = std::max<long long>(1, (long long)(s->h_up_bytes / ((size_t)nc * esz)));
    if ((size_t)nc * esz > s->h_up_bytes) {
        // a single row wider than a slot: split by columns recursively
        const long long half = nc / 2;
        KV_TRY(kv_upload_block(s, pack_bits, row0, n_rows, a, half, (const char *)host + src_off * esz, ld_elems));
        return kv_upload_block(s, pack_bits, row0, n_rows, a + half, nc - half,
                               (const char *)host + (src_off + half) * esz, ld_elems);
    }
    for (long long r = 0; r < n_rows; r += rows_per_slot) {
        const long long nr = std::min(rows_per_slot, n_rows - r);
        const int slot = s->up_slot;
        s->up_slot ^= 1;
        CU(cudaEventSynchronize(s->ev_up[slot]));
        char *pin = (char *)s->h_up[slot];
        stage_rows(pin, (const char *)host + ((size_t)r * ld_elems + src_off) * esz, nr, (size_t)nc * esz,
                   (size_t)ld_elems * esz);
        if (pack_bits == 64) {
            CU(cudaMemcpy2DAsync(s->d_mat + (row0 + r) * s->ld + lcol0, (size_t)s->ld * 8, pin, (size_t)nc * 8,
                                 (size_t)nc * 8, (size_t)nr, cudaMemcpyHostToDevice, s->stream));
        } else {
            KV_TRY(ensure(s->d_stage, s->h_up_bytes));
            CU(cudaMemcpyAsync(s->d_stage.p, pin, (size_t)nr * nc * 4, cudaMemcpyHostToDevice, s->stream));
            k_merge32<<<(unsigned)((nc + 255) / 256), 256, 0, s->stream>>>(s->d_mat, s->ld, lcol0, row0 + r, nr, nc,
                                                                            (const uint32_t *)s->d_stage.p, nc);
            count_launch(s);
        }
        CU(cudaEventRecord(s->ev_up[slot], s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    s->scm_valid = s->gini_valid = false;
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// upload of gzip-filtered HDF5 chunks, inflated on the device (kv_inflate.cuh)
// ------------------------------------------------------------------------------------------------
#include "kv_inflate.cuh"

namespace {
// Worker threads that copy byte ranges (page cache -> pinned memory), created once per upload call: spawning
// threads per 16 MB slot cost more than the copies themselves (2 240 thread starts for a 2.2 GB file).
class GatherPool {
public:
    explicit GatherPool(int n_threads) : n_(std::max(1, n_threads)) {
        for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~GatherPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto &th : workers_) th.join();
    }
    // copy ranges i = 0 .. n-1: src[i] (bytes[i] bytes) -> dst + off[i]; returns when all are done
    void run(char *dst, const void *const *src, const int64_t *bytes, const size_t *off, int64_t n) {
        if (n_ == 1 || n < 2) {
            for (int64_t i = 0; i < n; ++i) memcpy(dst + off[i], src[i], (size_t)bytes[i]);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = dst; src_ = src; bytes_ = bytes; off_ = off; count_ = n;
            next_.store(0);
            pending_ = n_ - 1;
            ++generation_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void work() {
        for (;;) {
            const int64_t i = next_.fetch_add(1);
            if (i >= count_) return;
            memcpy(dst_ + off_[i], src_[i], (size_t)bytes_[i]);
        }
    }
    void loop(int) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    unsigned long long generation_ = 0;
    bool stop_ = false;
    int pending_ = 0;
    std::atomic<int64_t> next_{0};
    char *dst_ = nullptr;
    const void *const *src_ = nullptr;
    const int64_t *bytes_ = nullptr;
    const size_t *off_ = nullptr;
    int64_t count_ = 0;
};

static int gather_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned want = getenv("KOVER_B200_UPLOAD_THREADS") ? (unsigned)atoi(getenv("KOVER_B200_UPLOAD_THREADS")) : 16u;
    return (int)std::max(1u, std::min(hw ? hw : 1u, std::max(1u, want)));
}

struct DevFree {                      // frees device buffers / events of one kv_upload_deflate call on every exit path
    std::vector<void *> bufs;
    std::vector<cudaEvent_t> events;
    bool idle = false;                // every stream that touched the buffers has been synchronised
    ~DevFree() {
        if (!idle) cudaDeviceSynchronize();          // (error exit: kernels of the call may still be running)
        for (void *p : bufs) dev_return(p);          // back to the process-wide cache (dev_borrow)
        for (cudaEvent_t e : events)
            if (e) cudaEventDestroy(e);
    }
};
}   // namespace

static constexpr int kInflateSlots = 4;

// page-locked staging slots + the copy stream of the deflate upload: allocated once per shard (pinning memory is slow)
static int ensure_inflate_staging(kv_shard *s, size_t slot_bytes) {
    if (!s->inf_copy_stream) CU(cudaStreamCreateWithFlags(&s->inf_copy_stream, cudaStreamNonBlocking));
    if (s->inf_slot_bytes >= slot_bytes) return KV_OK;
    for (int i = 0; i < kInflateSlots; ++i) {
        pinned_return(s->inf_pin[i]);
        s->inf_pin[i] = nullptr;
        CU(pinned_borrow(&s->inf_pin[i], slot_bytes));
        if (!s->inf_pin_free[i]) CU(cudaEventCreateWithFlags(&s->inf_pin_free[i], cudaEventDisableTiming));
    }
    s->inf_slot_bytes = slot_bytes;
    return KV_OK;
}

extern "C" int kv_upload_deflate(kv_shard *s, int pack_bits, int64_t n_chunks, const void *const *src,
                                 const int64_t *src_bytes, const int64_t *row0, const int64_t *gcol0,
                                 int64_t chunk_rows, int64_t chunk_cols, int32_t *status) {
    if (!s || (n_chunks > 0 && (!src || !src_bytes || !row0 || !gcol0 || !status)))
        return fail(KV_E_INVALID, "kv_upload_deflate: null pointer");
    if (pack_bits != 64 && pack_bits != 32)
        return fail(KV_E_INVALID, "Unsupported data type for packed attribute classifications array. The supported "
                                  "data types are np.uint32 and np.uint64.");
    if (n_chunks < 0 || chunk_rows <= 0 || chunk_cols <= 0 || chunk_rows > 0x7FFFFFFF)
        return fail(KV_E_INVALID, "kv_upload_deflate: bad chunk shape / count");
    if (n_chunks == 0) return KV_OK;
    const size_t esz = (size_t)pack_bits / 8;
    const size_t chunk_bytes = (size_t)chunk_rows * (size_t)chunk_cols * esz;
    if (chunk_bytes >= ((size_t)1 << 31)) return fail(KV_E_INVALID, "kv_upload_deflate: chunk larger than 2 GB");
    const long long rows_total = pack_bits == 64 ? s->W : (s->n_examples + 31) / 32;
    for (int64_t i = 0; i < n_chunks; ++i) {
        if (!src[i] || src_bytes[i] < 6 || src_bytes[i] >= ((int64_t)1 << 31) || row0[i] < 0 || gcol0[i] < 0)
            return fail(KV_E_INVALID, "kv_upload_deflate: bad chunk descriptor");
    }
    DeviceGuard g(s->device);
    s->last_launches = 0;
    s->scm_valid = s->gini_valid = false;
    static const bool stats = getenv("KOVER_B200_INFLATE_STATS") != nullptr;      // phase times on stderr
    static const bool trace = getenv("KOVER_B200_UPLOAD_TRACE") != nullptr;       // when each slab reached the device
    const auto now = [] { return std::chrono::steady_clock::now(); };
    const auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    const auto t_begin = now();
    double t_gather = 0.0;
    float t_kernels = 0.f;
    DevFree F;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    if (stats) {
        cudaEventCreate(&ev_k0);
        cudaEventCreate(&ev_k1);
        F.events.push_back(ev_k0);
        F.events.push_back(ev_k1);
    }
    // A chunk that lies wholly inside this shard's part of the dataset (one 64-bit packed row, 16-byte aligned
    // destination) inflates straight into its row of the resident image; edge chunks, chunks cut by a shard boundary,
    // multi-row chunks and 32-bit packing inflate into a scratch slot and are placed by k_scatter_chunks.
    auto direct = [&](int64_t i) {
        return pack_bits == 64 && chunk_rows == 1 && row0[i] < rows_total && gcol0[i] >= s->col0 &&
               gcol0[i] + chunk_cols <= s->col0 + s->n_cols && ((gcol0[i] - s->col0) & 1) == 0;
    };
    const size_t chunk_stride = (chunk_bytes + 15) / 16 * 16;
    auto comp_bytes = [&](int64_t i) { return ((size_t)src_bytes[i] + 4 + 15) / 16 * 16; };
    // decode slabs: enough chunks to occupy every SM several times over, bounded by a quarter of the free HBM
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    // One warp needs ~0.1 s per 800 KB chunk whatever the load (a single lane walks the Huffman stream), so the call
    // lasts (time to stage the last slab) + (one slab's kernel): small uploads are cut into up to four slabs whose
    // kernels run concurrently on their own streams and start as soon as their streams have arrived.
    int64_t slab = std::min<int64_t>(n_chunks, 8192);
    if (n_chunks <= 8192 && !getenv("KOVER_B200_UPLOAD_ONE_SLAB")) slab = std::max<int64_t>(256, (n_chunks + 3) / 4);
    while (slab > 64 && (size_t)slab * chunk_stride * 4 > free_b / 4) slab /= 2;
    size_t max_comp = 0, max_scratch = 0, biggest = 0;
    for (int64_t a = 0; a < n_chunks; a += slab) {
        size_t c = 0, q = 0;
        for (int64_t i = a; i < std::min(n_chunks, a + slab); ++i) {
            c += comp_bytes(i);
            q += direct(i) ? 0 : 1;
            biggest = std::max(biggest, comp_bytes(i));
        }
        max_comp = std::max(max_comp, c);
        max_scratch = std::max(max_scratch, q);
    }
    KV_TRY(ensure_inflate_staging(s, std::max((size_t)16 << 20, biggest)));
    constexpr int kMaxPar = 4;
    const size_t n_slabs = (size_t)((n_chunks + slab - 1) / slab);
    const int n_par = (int)std::min<size_t>(kMaxPar, n_slabs);
    unsigned char *d_comp[kMaxPar] = {nullptr}, *d_scratch[kMaxPar] = {nullptr};
    InflateJob *d_jobs[kMaxPar] = {nullptr};
    ScatterJob *d_sjobs[kMaxPar] = {nullptr};
    cudaEvent_t copied[kMaxPar] = {nullptr}, done[kMaxPar] = {nullptr};
    cudaStream_t k_stream[kMaxPar] = {s->stream, nullptr, nullptr, nullptr};
    struct StreamFree {
        cudaStream_t *st;
        ~StreamFree() {
            for (int i = 1; i < kMaxPar; ++i)
                if (st[i]) cudaStreamDestroy(st[i]);
        }
    } stream_free{k_stream};
    for (int i = 1; i < n_par; ++i) CU(cudaStreamCreateWithFlags(&k_stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < n_par; ++i) {
        CU(dev_borrow((void **)&d_comp[i], max_comp + 64, s->device));
        F.bufs.push_back(d_comp[i]);
        if (max_scratch) {
            CU(dev_borrow((void **)&d_scratch[i], max_scratch * chunk_stride, s->device));
            F.bufs.push_back(d_scratch[i]);
            CU(dev_borrow((void **)&d_sjobs[i], max_scratch * sizeof(ScatterJob), s->device));
            F.bufs.push_back(d_sjobs[i]);
        }
        CU(dev_borrow((void **)&d_jobs[i], (size_t)slab * sizeof(InflateJob), s->device));
        F.bufs.push_back(d_jobs[i]);
        CU(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        F.events.push_back(copied[i]);
        CU(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        F.events.push_back(done[i]);
    }
    int *d_status = nullptr;
    unsigned int *d_next = nullptr;
    CU(dev_borrow((void **)&d_status, (size_t)n_chunks * sizeof(int), s->device));
    F.bufs.push_back(d_status);
    CU(dev_borrow((void **)&d_next, n_slabs * sizeof(unsigned int), s->device));
    F.bufs.push_back(d_next);
    CU(cudaMemsetAsync(d_status, 0xFF, (size_t)n_chunks * sizeof(int), s->stream));
    CU(cudaMemsetAsync(d_next, 0, n_slabs * sizeof(unsigned int), s->stream));
    CU(cudaStreamSynchronize(s->stream));
    GatherPool pool(gather_threads());
    const double t_setup = ms_since(t_begin);
    std::vector<InflateJob> jobs((size_t)slab);
    std::vector<ScatterJob> sjobs(max_scratch);
    std::vector<size_t> off((size_t)slab);
    int slot = 0;
    int64_t slab_no = 0;
    for (int64_t a = 0; a < n_chunks; a += slab, ++slab_no) {
        const int64_t b = std::min(n_chunks, a + slab);
        const int par = (int)(slab_no % n_par);
        cudaStream_t ks = k_stream[par];
        if (slab_no >= n_par) CU(cudaStreamWaitEvent(s->inf_copy_stream, done[par], 0));   // this parity's buffers are free again
        // layout of the slab's streams in the device buffer; destination of every chunk
        size_t c = 0, n_scr = 0;
        for (int64_t i = a; i < b; ++i) {
            off[(size_t)(i - a)] = c;
            InflateJob &J = jobs[(size_t)(i - a)];
            J.src_off = c;
            J.src_bytes = (unsigned int)src_bytes[i];
            J.dst_bytes = (unsigned int)chunk_bytes;
            if (direct(i)) {
                J.dst = (unsigned char *)(s->d_mat + row0[i] * s->ld + (gcol0[i] - s->col0));
            } else {
                J.dst = d_scratch[par] + n_scr * chunk_stride;
                ScatterJob &Q = sjobs[n_scr++];
                Q.src = J.dst;
                Q.row0 = row0[i];
                Q.gcol0 = gcol0[i];
            }
            c += comp_bytes(i);
        }
        // stream the compressed bytes through the pinned slots: whole chunks per slot
        int64_t i = a;
        while (i < b) {
            const size_t base = off[(size_t)(i - a)];
            int64_t j = i;
            while (j < b && off[(size_t)(j - a)] + comp_bytes(j) - base <= s->inf_slot_bytes) ++j;
            CU(cudaEventSynchronize(s->inf_pin_free[slot]));
            std::vector<size_t> rel((size_t)(j - i));
            for (int64_t q = i; q < j; ++q) rel[(size_t)(q - i)] = off[(size_t)(q - a)] - base;
            const auto t_g = now();
            pool.run((char *)s->inf_pin[slot], src + i, src_bytes + i, rel.data(), j - i);
            t_gather += ms_since(t_g);
            const size_t span = (j < b ? off[(size_t)(j - a)] : c) - base;
            CU(cudaMemcpyAsync(d_comp[par] + base, s->inf_pin[slot], span, cudaMemcpyHostToDevice, s->inf_copy_stream));
            CU(cudaEventRecord(s->inf_pin_free[slot], s->inf_copy_stream));
            slot = (slot + 1) % kInflateSlots;
            i = j;
        }
        const int n_here = (int)(b - a);
        CU(cudaMemcpyAsync(d_jobs[par], jobs.data(), (size_t)n_here * sizeof(InflateJob), cudaMemcpyHostToDevice, s->inf_copy_stream));
        if (n_scr)
            CU(cudaMemcpyAsync(d_sjobs[par], sjobs.data(), n_scr * sizeof(ScatterJob), cudaMemcpyHostToDevice, s->inf_copy_stream));
        CU(cudaEventRecord(copied[par], s->inf_copy_stream));
        CU(cudaStreamSynchronize(s->inf_copy_stream));           // `jobs` / `sjobs` are reused by the next slab
        if (trace) fprintf(stderr, "  slab %lld staged at %.1f ms (gather so far %.1f ms)\n", (long long)slab_no, ms_since(t_begin), t_gather);
        CU(cudaStreamWaitEvent(ks, copied[par], 0));
        const int grid = (int)std::min<long long>((n_here + INF_WARPS - 1) / INF_WARPS, (long long)s->sm_count * 6);
        if (stats) cudaEventRecord(ev_k0, ks);
        // (KOVER_B200_INFLATE_V=1: the first decoder, one lane walking the whole stream -- kept for comparison)
        // default 4: every symbol of a window decoded speculatively (best on compressible = realistic k-mer matrices);
        // =3: literal runs two codes per dependent load (best on literal-only streams); =2 / =1: the earlier decoders
        const int inflate_v = getenv("KOVER_B200_INFLATE_V") ? atoi(getenv("KOVER_B200_INFLATE_V")) : 4;
        if (inflate_v == 1) {
            k_inflate<<<grid, INF_WARPS * 32, 0, ks>>>(d_comp[par], d_jobs[par], n_here, d_status + a, d_next + slab_no);
        } else if (inflate_v == 2) {
            k_inflate2<<<grid, INF_WARPS * 32, 0, ks>>>(d_comp[par], d_jobs[par], n_here, d_status + a, d_next + slab_no);
        } else if (inflate_v != 3) {
            constexpr size_t smem4 = sizeof(InflateWarp4) * INF_WARPS;
            static std::atomic<bool> armed4[64];
            if (s->device < 64 && !armed4[s->device].exchange(true))
                CU(cudaFuncSetAttribute(k_inflate4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            const int grid4 = std::min(grid, s->sm_count * 4);
            k_inflate4<<<grid4, INF_WARPS * 32, smem4, ks>>>(d_comp[par], d_jobs[par], n_here, d_status + a, d_next + slab_no);
        } else {
            constexpr size_t smem3 = sizeof(InflateWarp3) * INF_WARPS;
            static std::atomic<bool> armed[64];
            if (s->device < 64 && !armed[s->device].exchange(true))
                CU(cudaFuncSetAttribute(k_inflate3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            const int grid3 = std::min(grid, s->sm_count * 4);
            k_inflate3<<<grid3, INF_WARPS * 32, smem3, ks>>>(d_comp[par], d_jobs[par], n_here, d_status + a, d_next + slab_no);
        }
        if (stats) {
            cudaEventRecord(ev_k1, ks);
            cudaEventSynchronize(ev_k1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev_k0, ev_k1);
            t_kernels += ms;
        }
        count_launch(s);
        if (n_scr) {
            // (a chunk the kernel rejected is scattered too: the caller re-does it on the host, overwriting its rows)
            const long long per_chunk = (long long)chunk_rows * chunk_cols;
            const dim3 sgrid((unsigned)std::min<long long>((per_chunk + 255) / 256, 64), (unsigned)n_scr);
            if (pack_bits == 64)
                k_scatter_chunks<uint64_t><<<sgrid, 256, 0, ks>>>(d_sjobs[par], s->d_mat, s->ld, s->col0, s->n_cols,
                                                                  rows_total, s->k_total, (int)chunk_rows, chunk_cols);
            else
                k_scatter_chunks<uint32_t><<<sgrid, 256, 0, ks>>>(d_sjobs[par], s->d_mat, s->ld, s->col0, s->n_cols,
                                                                  rows_total, s->k_total, (int)chunk_rows, chunk_cols);
            count_launch(s);
        }
        CU(cudaEventRecord(done[par], ks));
    }
    for (int i = 1; i < n_par; ++i) CU(cudaStreamSynchronize(k_stream[i]));
    static_assert(sizeof(int) == sizeof(int32_t), "status array");
    CU(cudaMemcpyAsync(status, d_status, (size_t)n_chunks * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    F.idle = true;                                  // (every kernel stream and the copy stream were synchronised above)
    s->up_setup_ms = t_setup;
    s->up_gather_ms = t_gather;
    s->up_total_ms = ms_since(t_begin);
    if (stats) {
        size_t comp = 0;
        for (int64_t q = 0; q < n_chunks; ++q) comp += (size_t)src_bytes[q];
        fprintf(stderr, "kv_upload_deflate: %lld chunks (%.1f MB deflated -> %.1f MB), slabs of %lld, %zu via scratch: setup "
                        "%.1f ms, gather %.1f ms, inflate kernels %.1f ms (serialised for this report), total %.1f ms\n",
                (long long)n_chunks, comp / 1e6, (double)n_chunks * chunk_bytes / 1e6, (long long)slab, max_scratch, t_setup,
                t_gather, t_kernels, ms_since(t_begin));
    }
    return KV_OK;
}

extern "C" int kv_upload_deflate_stats(kv_shard *s, double *setup_ms, double *gather_ms, double *total_ms) {
    if (!s || !setup_ms || !gather_ms || !total_ms) return fail(KV_E_INVALID, "kv_upload_deflate_stats: null pointer");
    *setup_ms = s->up_setup_ms;
    *gather_ms = s->up_gather_ms;
    *total_ms = s->up_total_ms;
    return KV_OK;
}

extern "C" int kv_generate_synthetic(kv_shard *s, uint64_t seed) {
    if (!s) return fail(KV_E_INVALID, "kv_generate_synthetic: null shard");
    if (s->W > 8192) return fail(KV_E_INVALID, "kv_generate_synthetic: more than 8192 packed rows");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    const int rem = (int)(s->n_examples % 64);
    const uint64_t last_mask = rem ? (~0ull << (64 - rem)) : ~0ull;
    dim3 grid((unsigned)((s->n_cols + 255) / 256), (unsigned)s->W);
    k_generate<<<grid, 256, 0, s->stream>>>(s->d_mat, s->ld, s->W, s->n_cols, s->col0, seed, last_mask);
    count_launch(s);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    s->scm_valid = s->gini_valid = false;
    return KV_OK;
}

extern "C" int kv_read_block(kv_shard *s, int64_t row0, int64_t n_rows, int64_t lcol0, int64_t n_cols,
                             uint64_t *host, int64_t ld_elems) {
    if (!s || !host) return fail(KV_E_INVALID, "kv_read_block: null pointer");
    if (row0 < 0 || n_rows < 0 || row0 + n_rows > s->W || lcol0 < 0 || n_cols < 0 || lcol0 + n_cols > s->n_cols ||
        ld_elems < n_cols)
        return fail(KV_E_INVALID, "kv_read_block: block outside the shard");
    if (n_rows == 0 || n_cols == 0) return KV_OK;
    DeviceGuard g(s->device);
    CU(cudaMemcpy2DAsync(host, (size_t)ld_elems * 8, s->d_mat + row0 * s->ld + lcol0, (size_t)s->ld * 8,
                         (size_t)n_cols * 8, (size_t)n_rows, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return KV_OK;
}

extern "C" int kv_timer_start(kv_shard *s) {
    if (!s) return fail(KV_E_INVALID, "kv_timer_start: null shard");
    DeviceGuard g(s->device);
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaEventRecord(s->ev_t0, s->stream));
    return KV_OK;
}
extern "C" int kv_timer_stop(kv_shard *s, double *elapsed_ms) {
    if (!s || !elapsed_ms) return fail(KV_E_INVALID, "kv_timer_stop: null pointer");
    DeviceGuard g(s->device);
    CU(cudaEventRecord(s->ev_t1, s->stream));
    CU(cudaEventSynchronize(s->ev_t1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1));
    *elapsed_ms = ms;
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// a1: in-place popcount of a host block
// ------------------------------------------------------------------------------------------------
template <typename T>
static int popc_inplace_host(int device, T *arr, int64_t m, int64_t n, const T *mask) {
    if (m < 0 || n < 0) return fail(KV_E_INVALID, "inplace_popcount: negative shape");
    if (m == 0 || n == 0) return KV_OK;
    if (!arr || !mask) return fail(KV_E_INVALID, "inplace_popcount: null pointer");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || device < 0 || device >= n_dev)
        return fail(KV_E_CUDA, "inplace_popcount: no usable CUDA device (no CPU fallback exists)");
    DeviceGuard g(device);
    T *d_arr = nullptr, *d_mask = nullptr;
    const size_t bytes = (size_t)m * n * sizeof(T);
    CU(cudaMalloc(&d_arr, bytes));
    if (cudaMalloc(&d_mask, (size_t)m * sizeof(T)) != cudaSuccess) {
        cudaFree(d_arr);
        return fail(KV_E_NOMEM, "inplace_popcount: cudaMalloc failed");
    }
    int rc = KV_OK;
    do {
        if (cudaMemcpy(d_arr, arr, bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d_mask, mask, (size_t)m * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
            rc = fail(KV_E_CUDA, "inplace_popcount: host to device copy failed");
            break;
        }
        const long long total = (long long)m * n;
        const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
        k_popc_inplace<T><<<grid, 256>>>(d_arr, m, n, d_mask);
        if (cudaGetLastError() != cudaSuccess || cudaMemcpy(arr, d_arr, bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(KV_E_CUDA, "inplace_popcount: kernel or device to host copy failed");
    } while (0);
    cudaFree(d_arr);
    cudaFree(d_mask);
    return rc;
}
extern "C" int kv_inplace_popcount_64(int device, uint64_t *arr, int64_t m, int64_t n, const uint64_t *row_mask) {
    return popc_inplace_host<uint64_t>(device, arr, m, n, row_mask);
}
extern "C" int kv_inplace_popcount_32(int device, uint32_t *arr, int64_t m, int64_t n, const uint32_t *row_mask) {
    return popc_inplace_host<uint32_t>(device, arr, m, n, row_mask);
}

// ------------------------------------------------------------------------------------------------
// row records
// ------------------------------------------------------------------------------------------------
// Build the active-row records for a mask pair in pinned memory and push them.  Rows are grouped
// (only mask 0, only mask 1, both) so that consecutive rows of a stage take the same branches.
// Layout in d_rec: m0[n_act] | m1[n_act] | row[n_act].
static int push_records2(kv_shard *s, const uint64_t *m0, const uint64_t *m1, int &n_act, int &n_both) {
    const long long W = s->W;
    KV_TRY(ensure_pinned(s, (size_t)W * 20 + 64));
    int nA = 0, nB = 0, nC = 0;
    for (long long w = 0; w < W; ++w) {
        const bool a = m0[w] != 0, b = m1 && m1[w] != 0;
        if (a && b) ++nC;
        else if (a) ++nA;
        else if (b) ++nB;
    }
    n_act = nA + nB + nC;
    n_both = nC;
    uint64_t *pm0 = (uint64_t *)s->h_pin, *pm1 = pm0 + n_act;
    uint32_t *prow = (uint32_t *)(pm1 + n_act);
    int iA = 0, iB = nA, iC = nA + nB;
    for (long long w = 0; w < W; ++w) {
        const bool a = m0[w] != 0, b = m1 && m1[w] != 0;
        int i;
        if (a && b) i = iC++;
        else if (a) i = iA++;
        else if (b) i = iB++;
        else continue;
        pm0[i] = m0[w];
        pm1[i] = m1 ? m1[w] : 0;
        prow[i] = (uint32_t)w;
    }
    if (n_act > SCAN_INLINE_ROWS) {          // otherwise fill_scan2 copies them into the kernel parameters
        KV_TRY(ensure(s->d_rec, (size_t)n_act * 20 + 64));
        CU(cudaMemcpyAsync(s->d_rec.p, s->h_pin, (size_t)n_act * 20, cudaMemcpyHostToDevice, s->stream));
    }
    return KV_OK;
}

static void fill_scan2(kv_shard *s, Scan2Params &P, int n_act) {
    memset(&P, 0, sizeof P);
    P.mat = s->d_mat;
    P.ld = s->ld;
    P.n_cols = s->n_cols;
    P.col0 = s->col0;
    P.k_total = s->k_total;
    P.n_act = n_act;
    P.l2_hints = s->l2_hints;
    P.csa = s->scan_csa;
    if (n_act <= SCAN_INLINE_ROWS) {
        const uint64_t *pm0 = (const uint64_t *)s->h_pin, *pm1 = pm0 + n_act;
        const uint32_t *prow = (const uint32_t *)(pm1 + n_act);
        P.inline_rec = 1;
        memcpy(P.inl_m0, pm0, (size_t)n_act * 8);
        memcpy(P.inl_m1, pm1, (size_t)n_act * 8);
        memcpy(P.inl_row, prow, (size_t)n_act * 4);
    } else {
        P.rec_m0 = (const uint64_t *)s->d_rec.p;
        P.rec_m1 = P.rec_m0 + n_act;
        P.rec_row = (const uint32_t *)(P.rec_m1 + n_act);
    }
    P.ubs = 1;
}

// Scan-kernel variants <consumer warps, rows per stage, stages, CTAs per SM>, measured on B200 by
// tools/scan_sweep.cu (profiles/README.md): variant 0 is best when each packed row carries one
// non-zero mask word (label-sorted datasets: 2 popc per streamed word), variant 1 when most rows
// carry both (4 popc per word: more consumer warps to issue them).
// X(id, consumer warps, rows per stage, stages, CTAs per SM)
#define KV_SCAN_VARIANTS(X) \
    X(0, 8, 4, 4, 2)        \
    X(1, 16, 8, 3, 1)       \
    X(2, 16, 4, 3, 2)       \
    X(3, 16, 2, 4, 2)       \
    X(4, 16, 4, 6, 1)       \
    X(5, 16, 4, 4, 1)
static constexpr int kNumScanVariants = 6;

static constexpr size_t kSmemLimit = 227 * 1024 - 2048;   // opt-in limit per CTA, minus static + driver-reserved shared memory

template <int CW, int R, int S>
static constexpr size_t scan_smem_bytes(size_t n_act_staged) {
    return ((size_t)(2 * S) * 8 + n_act_staged * 20 + 127) / 128 * 128 + (size_t)S * R * CW * 512;
}

template <int EPI, int CW, int R, int S, int MINB, bool GREC = false>
static int launch_scan_variant(kv_shard *s, const Scan2Params &P) {
    const size_t smem = scan_smem_bytes<CW, R, S>(GREC ? 0 : (size_t)P.n_act);
    if (smem > kSmemLimit)
        return fail(KV_E_INVALID, "too many active packed rows for this scan variant (shared-memory staging limit)");
    if (GREC && P.inline_rec) return fail(KV_E_INVALID, "global-record scan needs the records in device memory");
    static thread_local int attr_device = -1;
    static thread_local size_t attr_bytes = 0;
    if (attr_device != s->device || attr_bytes < smem) {   // per (instantiation, device)
        CU(cudaFuncSetAttribute(k_scan_tma<EPI, CW, R, S, MINB, GREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_device = s->device;
        attr_bytes = smem;
    }
    const long long n_tiles = s->ld / (CW * 64);
    const int grid = (int)std::max(1ll, std::min<long long>((long long)s->sm_count * MINB, n_tiles));
    if ((EPI == EPI_SCM || EPI == EPI_SCM_STREAM) && P.tail.mode) {
        // the tail's grid barrier needs every CTA resident: at most MINB CTAs per SM are launched (persistent grid)
        // and a cooperative launch makes the driver verify it.  (The streaming tail has no barrier: plain launch.)
        const size_t tables = ((size_t)P.tail.n_blocks * 17 + 63) / 16 * 16 + (EPI == EPI_SCM_STREAM ? (size_t)(grid + 1) * 4 : 0);
        if (tables > (size_t)S * R * CW * 512)
            return fail(KV_E_INVALID, "fused scan: block tables do not fit the ring");
        Scan2Params Q = P;
        Q.tail.barrier_target = s->fused_barrier_total + (unsigned long long)grid;
        Q.tail.done_target = s->fused_done_total + (unsigned long long)grid;
        if (EPI == EPI_SCM && s->cooperative) {
            void *args[] = {(void *)&Q};
            CU(cudaLaunchCooperativeKernel((const void *)k_scan_tma<EPI, CW, R, S, MINB, GREC>, dim3(grid),
                                           dim3((CW + 1) * 32), args, smem, s->stream));
        } else {
            k_scan_tma<EPI, CW, R, S, MINB, GREC><<<grid, (CW + 1) * 32, smem, s->stream>>>(Q);
        }
        if (EPI == EPI_SCM) s->fused_barrier_total += (unsigned long long)grid;
        s->fused_done_total += (unsigned long long)grid;
        return KV_OK;
    }
    k_scan_tma<EPI, CW, R, S, MINB, GREC><<<grid, (CW + 1) * 32, smem, s->stream>>>(P);
    return KV_OK;
}

template <int EPI>
static int launch_scan(kv_shard *s, const Scan2Params &P, int n_both) {
    // Default: variant 1 (16 consumer warps, 8 rows x 8 KB per stage, 3 stages = 192 KB in flight per SM) --
    // within 1 % of the best variant for label-sorted masks and the best when both masks hit every
    // row (profiles/README.md).  With more than ~1500 active packed rows the mask staging no longer fits
    // beside that ring: variant 5 (128 KB ring).
    // Beyond what variant 5 can stage (~4 900 active rows) the records stay in global memory: variant 6 = variant 1's
    // ring with GREC (no row limit; KOVER_B200_SCAN_VARIANT=6 forces it for tests).
    (void)n_both;
    int v = s->scan_variant;
    if (v == 6 && P.inline_rec) v = -1;               // forced GREC needs device-resident records (n_act > 64)
    if (v < 0 || v > kNumScanVariants) v = ((size_t)P.n_act * 20 > 30 * 1024) ? 5 : 1;
    if (v == 5 && scan_smem_bytes<16, 4, 4>((size_t)P.n_act) > kSmemLimit) v = 6;
    if (s->profiling) CU(cudaEventRecord(s->ev0, s->stream));
    int rc = KV_OK;
    switch (v) {
#define KV_CASE(ID, CW_, R_, S_, MINB_) \
    case ID: rc = launch_scan_variant<EPI, CW_, R_, S_, MINB_>(s, P); break;
        KV_SCAN_VARIANTS(KV_CASE)
#undef KV_CASE
        case 6: rc = launch_scan_variant<EPI, 16, 8, 3, 1, true>(s, P); break;
        default: return fail(KV_E_INVALID, "unknown scan variant");
    }
    KV_TRY(rc);
    count_launch(s);
    if (s->profiling) CU(cudaEventRecord(s->ev1, s->stream));
    return KV_OK;
}

// The candidate-streaming step only exists for the default ring shapes (variants 1 / 5 / 6 of launch_scan, same
// choice); a forced KOVER_B200_SCAN_VARIANT other than those keeps the counting kernel (-> false).
static constexpr int kCandCap = 2048;                 // candidate records per CTA
static bool stream_variant_ok(const kv_shard *s) {
    const int v = s->scan_variant;
    return v < 0 || v > kNumScanVariants || v == 1 || v == 4 || v == 5 || v == 6;
}
static int launch_scan_stream(kv_shard *s, const Scan2Params &P) {
    int v = s->scan_variant;
    if (v == 6 && P.inline_rec) v = -1;
    if (v < 0 || v > kNumScanVariants) v = ((size_t)P.n_act * 20 > 30 * 1024) ? 5 : 1;
    if (v == 5 && scan_smem_bytes<16, 4, 4>((size_t)P.n_act) > kSmemLimit) v = 6;
    if (s->profiling) CU(cudaEventRecord(s->ev0, s->stream));
    int rc;
    switch (v) {
        case 1: rc = launch_scan_variant<EPI_SCM_STREAM, 16, 8, 3, 1>(s, P); break;
        case 4: rc = launch_scan_variant<EPI_SCM_STREAM, 16, 4, 6, 1>(s, P); break;
        case 5: rc = launch_scan_variant<EPI_SCM_STREAM, 16, 4, 4, 1>(s, P); break;
        case 6: rc = launch_scan_variant<EPI_SCM_STREAM, 16, 8, 3, 1, true>(s, P); break;
        default: return fail(KV_E_INVALID, "no streaming kernel for this scan variant");
    }
    KV_TRY(rc);
    count_launch(s);
    if (s->profiling) CU(cudaEventRecord(s->ev1, s->stream));
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// a3: sum_rows
// ------------------------------------------------------------------------------------------------
extern "C" int kv_sum_rows(kv_shard *s, const uint64_t *mask, void *out, int out_itemsize) {
    if (!s || !mask || !out) return fail(KV_E_INVALID, "kv_sum_rows: null pointer");
    if (out_itemsize != 1 && out_itemsize != 2 && out_itemsize != 4 && out_itemsize != 8)
        return fail(KV_E_INVALID, "kv_sum_rows: out_itemsize must be 1, 2, 4 or 8");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(ensure_counts(s, 1));
    int n_act, n_both;
    KV_TRY(push_records2(s, mask, nullptr, n_act, n_both));
    Scan2Params P;
    fill_scan2(s, P, n_act);
    P.cnt0 = s->d_cnt[0];
    P.cnt1 = nullptr;
    KV_TRY(launch_scan<EPI_COUNTS>(s, P, n_both));
    s->scm_valid = s->gini_valid = false;
    if (out_itemsize == 4) {
        CU(cudaMemcpyAsync(out, s->d_cnt[0], (size_t)s->n_cols * 4, cudaMemcpyDeviceToHost, s->stream));
    } else {
        KV_TRY(ensure(s->d_stage, (size_t)s->n_cols * out_itemsize));
        const unsigned grid = (unsigned)std::min<long long>((s->n_cols + 255) / 256, (long long)s->sm_count * 8);
        if (out_itemsize == 1) k_narrow<uint8_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint8_t *)s->d_stage.p, s->n_cols);
        else if (out_itemsize == 2) k_narrow<uint16_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint16_t *)s->d_stage.p, s->n_cols);
        else k_narrow<uint64_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint64_t *)s->d_stage.p, s->n_cols);
        count_launch(s);
        CU(cudaMemcpyAsync(out, s->d_stage.p, (size_t)s->n_cols * out_itemsize, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (s->profiling) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->last_kernel_ms = ms;
    }
    return KV_OK;
}

extern "C" int kv_sum_rows_full(kv_shard *s, const uint64_t *mask, int64_t n_rows, void *out_presence,
                                void *out_absence, int out_itemsize) {
    if (!s || !mask || !out_presence || !out_absence) return fail(KV_E_INVALID, "kv_sum_rows_full: null pointer");
    if (out_itemsize != 1 && out_itemsize != 2 && out_itemsize != 4 && out_itemsize != 8)
        return fail(KV_E_INVALID, "kv_sum_rows_full: out_itemsize must be 1, 2, 4 or 8");
    if (n_rows < 0) return fail(KV_E_INVALID, "kv_sum_rows_full: negative n_rows");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(ensure_counts(s, 1));
    int n_act, n_both;
    KV_TRY(push_records2(s, mask, nullptr, n_act, n_both));
    Scan2Params P;
    fill_scan2(s, P, n_act);
    P.cnt0 = s->d_cnt[0];
    P.cnt1 = nullptr;
    KV_TRY(launch_scan<EPI_COUNTS>(s, P, n_both));
    s->scm_valid = s->gini_valid = false;
    const size_t half = (size_t)s->n_cols * out_itemsize;
    KV_TRY(ensure(s->d_stage, 2 * half + 64));
    char *d_p = (char *)s->d_stage.p, *d_a = d_p + (half + 15) / 16 * 16;
    const unsigned grid = (unsigned)std::min<long long>((s->n_cols + 255) / 256, (long long)s->sm_count * 8);
    const unsigned long long nr = (unsigned long long)n_rows;
    if (out_itemsize == 1) k_narrow2<uint8_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint8_t *)d_p, (uint8_t *)d_a, s->n_cols, nr);
    else if (out_itemsize == 2) k_narrow2<uint16_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint16_t *)d_p, (uint16_t *)d_a, s->n_cols, nr);
    else if (out_itemsize == 4) k_narrow2<uint32_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint32_t *)d_p, (uint32_t *)d_a, s->n_cols, nr);
    else k_narrow2<uint64_t><<<grid, 256, 0, s->stream>>>(s->d_cnt[0], (uint64_t *)d_p, (uint64_t *)d_a, s->n_cols, nr);
    count_launch(s);
    CU(cudaMemcpyAsync(out_presence, d_p, half, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(out_absence, d_a, half, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return finish_scan_timing(s);
}

// ------------------------------------------------------------------------------------------------
// Large tie lists (tens of thousands of equivalent rules are normal on real k-mer data): the tie
// kernels append in arbitrary order, the reference's order is ascending rule index.  Small lists are
// sorted on the host; large ones by a device radix sort so that the host only copies.
// ------------------------------------------------------------------------------------------------
__global__ void k_split_tierecs(const TieRec *recs, long long n, long long *keys, unsigned long long *vals) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const TieRec r = recs[i];
        keys[i] = r.idx;
        vals[i] = ((unsigned long long)r.pe << 32) | r.nc;
    }
}

// sorts keys (and vals if given) ascending; results land in keys_out / vals_out (device)
// keys are rule indices < 2 * k_total: the radix sort only visits the bits they can use
static int device_sort(kv_shard *s, const long long *keys_in, long long *keys_out, const unsigned long long *vals_in,
                       unsigned long long *vals_out, long long n) {
    int end_bit = 1;
    while (end_bit < 63 && (1ll << end_bit) <= 2 * s->k_total) ++end_bit;
    size_t tmp = 0;
    if (vals_in) CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, s->stream));
    else CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys_in, keys_out, (int)n, 0, end_bit, s->stream));
    KV_TRY(ensure(s->d_sort_tmp, tmp + 64));
    if (vals_in)
        CU(cub::DeviceRadixSort::SortPairs(s->d_sort_tmp.p, tmp, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, s->stream));
    else CU(cub::DeviceRadixSort::SortKeys(s->d_sort_tmp.p, tmp, keys_in, keys_out, (int)n, 0, end_bit, s->stream));
    count_launch(s, 4);
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// dataset split: per-k-mer individual risks (dataset/split.py:171-188, 213-228)
// ------------------------------------------------------------------------------------------------
extern "C" int kv_risk_errors(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                              int64_t n_total, uint8_t *present) {
    if (!s || !pos_mask || !neg_mask || !present) return fail(KV_E_INVALID, "kv_risk_errors: null pointer");
    if (n_pos < 0 || n_total < n_pos || n_total > s->n_examples) return fail(KV_E_INVALID, "kv_risk_errors: bad counts");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    s->scm_valid = s->gini_valid = s->risk_valid = false;
    KV_TRY(ensure_counts(s, 2));
    int n_act, n_both;
    KV_TRY(push_records2(s, pos_mask, neg_mask, n_act, n_both));
    Scan2Params P;
    fill_scan2(s, P, n_act);
    P.cnt0 = s->d_cnt[0];
    P.cnt1 = s->d_cnt[1];
    KV_TRY(launch_scan<EPI_COUNTS>(s, P, n_both));
    const size_t nb = (size_t)n_total + 1;
    KV_TRY(ensure(s->d_stage, nb + 64));
    CU(cudaMemsetAsync(s->d_stage.p, 0, nb, s->stream));
    const unsigned grid = (unsigned)std::min<long long>((s->n_cols + 255) / 256, (long long)s->sm_count * 8);
    k_risk_errors<<<grid, 256, 0, s->stream>>>(s->d_cnt[0], s->d_cnt[1], s->n_cols, (uint32_t)n_pos, (uint8_t *)s->d_stage.p);
    count_launch(s);
    KV_TRY(ensure_pinned(s, nb));
    CU(cudaMemcpyAsync(s->h_pin, s->d_stage.p, nb, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    const uint8_t *h = (const uint8_t *)s->h_pin;
    for (size_t i = 0; i < nb; ++i) present[i] |= h[i];
    s->risk_valid = true;
    s->risk_n = n_total;
    return finish_scan_timing(s);
}

extern "C" int kv_risk_map(kv_shard *s, const uint32_t *lut_kmer, const uint32_t *lut_anti, int64_t n_lut,
                           int out_itemsize, void *out_kmer, void *out_anti) {
    if (!s || !lut_kmer || !lut_anti || !out_kmer || !out_anti) return fail(KV_E_INVALID, "kv_risk_map: null pointer");
    if (!s->risk_valid) return fail(KV_E_INVALID, "kv_risk_map: no kv_risk_errors result on the device");
    if (n_lut != s->risk_n + 1) return fail(KV_E_INVALID, "kv_risk_map: table must have n_total + 1 entries");
    if (out_itemsize != 1 && out_itemsize != 2 && out_itemsize != 4 && out_itemsize != 8)
        return fail(KV_E_INVALID, "kv_risk_map: out_itemsize must be 1, 2, 4 or 8");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    const size_t lut_bytes = (size_t)n_lut * 4, half = ((size_t)s->n_cols * out_itemsize + 15) / 16 * 16;
    const size_t lut_pad = (lut_bytes + 15) / 16 * 16;
    KV_TRY(ensure(s->d_stage, 2 * lut_pad + 2 * half + 64));
    char *d_lk = (char *)s->d_stage.p, *d_la = d_lk + lut_pad, *d_ok = d_la + lut_pad, *d_oa = d_ok + half;
    CU(cudaMemcpyAsync(d_lk, lut_kmer, lut_bytes, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(d_la, lut_anti, lut_bytes, cudaMemcpyHostToDevice, s->stream));
    const unsigned grid = (unsigned)std::min<long long>((s->n_cols + 255) / 256, (long long)s->sm_count * 8);
    const uint32_t *e = s->d_cnt[0], *lk = (const uint32_t *)d_lk, *la = (const uint32_t *)d_la;
    if (out_itemsize == 1) k_risk_map<uint8_t><<<grid, 256, 0, s->stream>>>(e, lk, la, (uint8_t *)d_ok, (uint8_t *)d_oa, s->n_cols);
    else if (out_itemsize == 2) k_risk_map<uint16_t><<<grid, 256, 0, s->stream>>>(e, lk, la, (uint16_t *)d_ok, (uint16_t *)d_oa, s->n_cols);
    else if (out_itemsize == 4) k_risk_map<uint32_t><<<grid, 256, 0, s->stream>>>(e, lk, la, (uint32_t *)d_ok, (uint32_t *)d_oa, s->n_cols);
    else k_risk_map<uint64_t><<<grid, 256, 0, s->stream>>>(e, lk, la, (uint64_t *)d_ok, (uint64_t *)d_oa, s->n_cols);
    count_launch(s);
    CU(cudaMemcpyAsync(out_kmer, d_ok, (size_t)s->n_cols * out_itemsize, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(out_anti, d_oa, (size_t)s->n_cols * out_itemsize, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return KV_OK;
}

// ---- k_count_tma launch ------------------------------------------------------------------------
template <int NM, int R, int S, bool CSA>
static int launch_count_variant2(kv_shard *s, const CountParams &P, size_t smem) {
    constexpr int CW = 16;
    static thread_local int attr_device = -1;
    static thread_local size_t attr_bytes = 0;
    if (attr_device != s->device || attr_bytes < smem) {
        CU(cudaFuncSetAttribute(k_count_tma<NM, CW, R, S, CSA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_device = s->device;
        attr_bytes = smem;
    }
    const long long n_tiles = s->ld / (CW * 64);
    const int grid = (int)std::max(1ll, std::min<long long>((long long)s->sm_count, n_tiles));
    k_count_tma<NM, CW, R, S, CSA><<<grid, (CW + 1) * 32, smem, s->stream>>>(P);
    return KV_OK;
}
template <int NM, int R, int S>
static int launch_count_variant(kv_shard *s, const CountParams &P, size_t smem) {
    static const bool plain = getenv("KOVER_B200_COUNT_PLAIN") != nullptr;    // measurement knob (profiles/README.md)
    return plain ? launch_count_variant2<NM, R, S, false>(s, P, smem) : launch_count_variant2<NM, R, S, true>(s, P, smem);
}

static size_t count_smem_bytes(int n_act, int nm, int R, int S) {
    const size_t n_grp = ((size_t)n_act + 3) / 4;
    return ((size_t)(2 * S) * 8 + n_grp * ((size_t)nm * 32 + 20) + 127) / 128 * 128 + (size_t)S * R * 16 * 512;
}

template <int NM>
static int launch_count_nm(kv_shard *s, const CountParams &P) {
    // ring depth by what the mask staging leaves: 192 KB, 128 KB or 64 KB of reads in flight per SM
    const size_t limit = 227 * 1024 - 2048;
    size_t smem = count_smem_bytes(P.n_act, NM, 8, 3);
    if (smem <= limit) return launch_count_variant<NM, 8, 3>(s, P, smem);
    smem = count_smem_bytes(P.n_act, NM, 4, 4);
    if (smem <= limit) return launch_count_variant<NM, 4, 4>(s, P, smem);
    smem = count_smem_bytes(P.n_act, NM, 4, 2);
    if (smem <= limit) return launch_count_variant<NM, 4, 2>(s, P, smem);
    return fail(KV_E_INVALID, "too many active packed rows x masks for one pass (shared-memory staging limit)");
}

static int padded_nm(int nm) {
    static const int sizes[] = {2, 4, 6, 8, 12, 16};
    for (int v : sizes)
        if (nm <= v) return v;
    return -1;
}

// masks per pass so that the [row][mask] staging of W packed rows fits beside the smallest ring
static int max_masks_per_pass(long long W) {
    const long long budget = 227 * 1024 - 2048 - 64 * 1024 - 256;
    const long long n_grp = (std::max<long long>(W, 1) + 3) / 4;
    long long nm = (budget / n_grp - 20) / 32;
    nm = std::min<long long>(nm, MAX_MULTI);
    static const int sizes[] = {16, 12, 8, 6, 4, 2};
    for (int v : sizes)
        if (nm >= v) return v;
    return 0;
}

// staging bytes of one pass (row records [n_act][NM] u64 + rows u32), rounded so that consecutive passes stay aligned
static size_t count_batch_stride(long long W) {
    return (((size_t)W + 3) / 4 * (32 * (size_t)MAX_MULTI + 16) + 127) / 128 * 128;
}
static size_t count_pin_bytes(long long W, int n_masks) {
    const int per = std::max(2, max_masks_per_pass(W));
    const int n_batches = (n_masks + per - 1) / per;
    return (size_t)n_batches * count_batch_stride(W);
}

// Presence counts of n_masks row masks (host pointers, W words each) into the device arrays dst[m]
// ([ld] uint32 each), ceil(n_masks / masks-per-pass) passes of k_count_tma enqueued on the shard's stream.
// Nothing is synchronised: `pin` (pinned, count_pin_bytes(W, n_masks) bytes) stages the row records and
// must stay untouched until the stream has drained.  ev0 / ev1 bracket the passes.
static int count_masks_async(kv_shard *s, int n_masks, const uint64_t *const *masks, uint32_t *const *dst, char *pin) {
    const long long W = s->W;
    const int per = max_masks_per_pass(W);
    if (per < 2) {
        // More packed rows than the [row][mask] staging of k_count_tma can hold (W > ~7 800 = 500 000 genomes): two
        // masks per pass through the two-mask scan, whose records can stay in global memory (no row limit).  Rare
        // and huge, so every pass is synchronised (the pinned record buffer is reused).
        (void)pin;
        for (int m0 = 0; m0 < n_masks; m0 += 2) {
            CU(cudaStreamSynchronize(s->stream));
            int n_act, n_both;
            const bool two = m0 + 1 < n_masks;
            KV_TRY(push_records2(s, masks[m0], two ? masks[m0 + 1] : nullptr, n_act, n_both));
            Scan2Params P;
            fill_scan2(s, P, n_act);
            P.cnt0 = dst[m0];
            P.cnt1 = two ? dst[m0 + 1] : nullptr;
            KV_TRY(launch_scan<EPI_COUNTS>(s, P, n_both));
        }
        CU(cudaStreamSynchronize(s->stream));
        return KV_OK;
    }
    KV_TRY(ensure(s->d_rec, count_pin_bytes(W, n_masks)));
    if (s->profiling) CU(cudaEventRecord(s->ev0, s->stream));
    size_t off = 0;
    for (int m0 = 0; m0 < n_masks; m0 += per) {
        const int nm = std::min(per, n_masks - m0);
        const int nmp = padded_nm(nm);
        // active rows (any mask non-zero), ordered by which masks touch them so that the groups of four rows
        // the kernel skips masks on are as homogeneous as possible (stable: ties keep the matrix order)
        std::vector<std::pair<uint32_t, uint32_t>> act;      // (pattern, row)
        for (long long w = 0; w < W; ++w) {
            uint32_t pat = 0;
            for (int m = 0; m < nm; ++m) pat |= (uint32_t)(masks[m0 + m][w] != 0) << m;
            if (pat) act.emplace_back(pat, (uint32_t)w);
        }
        std::stable_sort(act.begin(), act.end(),
                         [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
        const int n_act = (int)act.size();
        const int n_grp = (n_act + 3) / 4;
        uint64_t *pm = (uint64_t *)(pin + off);
        uint32_t *prow = (uint32_t *)(pm + (size_t)n_grp * nmp * 4);
        memset(pm, 0, (size_t)n_grp * nmp * 32 + (size_t)n_grp * 16);
        for (int i = 0; i < n_act; ++i) {
            const uint32_t w = act[i].second;
            for (int m = 0; m < nm; ++m) pm[((size_t)(i >> 2) * nmp + m) * 4 + (i & 3)] = masks[m0 + m][w];
            prow[i] = w;
        }
        const size_t rec_bytes = (size_t)n_grp * ((size_t)nmp * 32 + 16);
        char *d_rec = (char *)s->d_rec.p + off;
        if (n_act) CU(cudaMemcpyAsync(d_rec, pin + off, rec_bytes, cudaMemcpyHostToDevice, s->stream));
        CountParams P;
        memset(&P, 0, sizeof P);
        P.mat = s->d_mat;
        P.ld = s->ld;
        P.rec_m = (const uint64_t *)d_rec;
        P.rec_row = (const uint32_t *)(P.rec_m + (size_t)n_grp * nmp * 4);
        P.n_act = n_act;
        for (int m = 0; m < nm; ++m) P.cnt[m] = dst[m0 + m];
        int rc = KV_OK;
        switch (nmp) {
            case 2: rc = launch_count_nm<2>(s, P); break;
            case 4: rc = launch_count_nm<4>(s, P); break;
            case 6: rc = launch_count_nm<6>(s, P); break;
            case 8: rc = launch_count_nm<8>(s, P); break;
            case 12: rc = launch_count_nm<12>(s, P); break;
            case 16: rc = launch_count_nm<16>(s, P); break;
            default: return fail(KV_E_INVALID, "count_masks: bad mask count");
        }
        KV_TRY(rc);
        count_launch(s);
        off += count_batch_stride(W);
    }
    if (s->profiling) CU(cudaEventRecord(s->ev1, s->stream));
    return KV_OK;
}

// counts of n_masks masks ([n_masks][W] contiguous) into d_cnt[first .. first+n_masks), synchronised
static int count_masks(kv_shard *s, int n_masks, const uint64_t *masks, int first) {
    KV_TRY(ensure_counts(s, first + n_masks));
    KV_TRY(ensure_pinned(s, count_pin_bytes(s->W, n_masks)));
    CU(cudaStreamSynchronize(s->stream));          // an in-flight copy may still read h_pin
    std::vector<const uint64_t *> ptrs(n_masks);
    for (int m = 0; m < n_masks; ++m) ptrs[m] = masks + (size_t)m * s->W;
    KV_TRY(count_masks_async(s, n_masks, ptrs.data(), s->d_cnt.data() + first, (char *)s->h_pin));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (s->profiling) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->last_kernel_ms = ms;
    }
    return KV_OK;
}

extern "C" int kv_sum_rows_multi(kv_shard *s, int n_masks, const uint64_t *masks, uint32_t *out) {
    if (!s || !masks || !out) return fail(KV_E_INVALID, "kv_sum_rows_multi: null pointer");
    if (n_masks <= 0 || n_masks > 4096) return fail(KV_E_INVALID, "kv_sum_rows_multi: n_masks out of range");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    s->scm_valid = s->gini_valid = false;
    KV_TRY(count_masks(s, n_masks, masks, 0));
    for (int m = 0; m < n_masks; ++m)
        CU(cudaMemcpyAsync(out + (size_t)m * s->n_cols, s->d_cnt[m], (size_t)s->n_cols * 4, cudaMemcpyDeviceToHost,
                           s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// a4: get_columns
// ------------------------------------------------------------------------------------------------
extern "C" int kv_get_columns(kv_shard *s, const int64_t *cols, int64_t n, uint64_t *out) {
    if (!s || (n > 0 && (!cols || !out))) return fail(KV_E_INVALID, "kv_get_columns: null pointer");
    if (n < 0) return fail(KV_E_INVALID, "kv_get_columns: negative count");
    if (n == 0) return KV_OK;
    for (int64_t i = 0; i < n; ++i)
        if (cols[i] < 0 || cols[i] >= s->n_cols) return fail(KV_E_INVALID, "kv_get_columns: column outside the shard");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    const size_t out_bytes = (size_t)n * s->W * 8, idx_bytes = (size_t)n * 8;
    KV_TRY(ensure(s->d_stage, out_bytes + idx_bytes));
    long long *d_cols = nullptr;
    InlineCols inl;
    memset(&inl, 0, sizeof inl);
    if (n <= 8) {
        for (int64_t i = 0; i < n; ++i) inl.c[i] = cols[i];
    } else {
        d_cols = (long long *)((char *)s->d_stage.p + out_bytes);
        CU(cudaMemcpyAsync(d_cols, cols, idx_bytes, cudaMemcpyHostToDevice, s->stream));
    }
    const int threads = (int)std::min<long long>(256, std::max<long long>(32, s->W));
    k_gather_cols<<<(unsigned)n, threads, 0, s->stream>>>(s->d_mat, s->ld, s->W, d_cols, inl, n, (uint64_t *)s->d_stage.p);
    count_launch(s);
    if (out_bytes <= ((size_t)1 << 16)) {          // small result: through pinned memory, polled
        KV_TRY(ensure_pinned(s, out_bytes));
        CU(cudaMemcpyAsync(s->h_pin, s->d_stage.p, out_bytes, cudaMemcpyDeviceToHost, s->stream));
        CU(stream_sync_spin(s->stream));
        memcpy(out, s->h_pin, out_bytes);
    } else {
        CU(cudaMemcpyAsync(out, s->d_stage.p, out_bytes, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    CU(cudaGetLastError());
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// a6: SCM scan
// ------------------------------------------------------------------------------------------------
extern "C" int kv_set_blacklist(kv_shard *s, const int64_t *rule_idx, int64_t n) {
    if (!s || (n > 0 && !rule_idx)) return fail(KV_E_INVALID, "kv_set_blacklist: null pointer");
    DeviceGuard g(s->device);
    s->scm_valid = s->gini_valid = false;
    if (n <= 0) {
        s->have_black = false;
        return KV_OK;
    }
    const size_t words = (size_t)s->ld / 32;
    std::vector<uint32_t> bp(words, 0), ba(words, 0);
    for (int64_t i = 0; i < n; ++i) {
        int64_t r = rule_idx[i];
        if (r < 0) r += 2 * s->k_total;                       // numpy negative indexing (scm.py:240)
        if (r < 0 || r >= 2 * s->k_total) return fail(KV_E_INVALID, "kv_set_blacklist: rule index out of range");
        const bool absence = r >= s->k_total;
        const int64_t c = (absence ? r - s->k_total : r) - s->col0;
        if (c < 0 || c >= s->n_cols) continue;
        (absence ? ba : bp)[c >> 5] |= 1u << (c & 31);
    }
    if (!s->d_black_pres) {
        CU(cudaMalloc(&s->d_black_pres, words * 4));
        CU(cudaMalloc(&s->d_black_abs, words * 4));
    }
    CU(cudaMemcpyAsync(s->d_black_pres, bp.data(), words * 4, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_black_abs, ba.data(), words * 4, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->have_black = true;
    return KV_OK;
}

extern "C" int64_t kv_scm_n_blocks(int64_t n_kmers_total, int64_t util_block_size) {
    if (n_kmers_total <= 0 || util_block_size <= 0) return 0;
    return (2 * n_kmers_total + util_block_size - 1) / util_block_size;
}

// d_misc layout: blockmax enc [n_blocks] | bm f64 [n_blocks] | tol f64 [n_blocks] | sel u8 [n_blocks]
struct MiscLayout {
    unsigned long long *blockmax;
    double *bm, *tol;
    uint8_t *sel;
    size_t bytes;
};
static MiscLayout misc_layout(void *base, long long n_blocks) {
    MiscLayout L;
    char *p = (char *)base;
    L.blockmax = (unsigned long long *)p; p += (size_t)n_blocks * 8;
    L.bm = (double *)p; p += (size_t)n_blocks * 8;
    L.tol = (double *)p; p += (size_t)n_blocks * 8;
    L.sel = (uint8_t *)p; p += ((size_t)n_blocks + 15) / 16 * 16;
    L.bytes = (size_t)(p - (char *)base);
    return L;
}

static constexpr long long kTieFirstBatch = 256;   // tie records fetched together with the header

static int ensure_ties(kv_shard *s, long long cap) {
    return ensure(s->d_ties, sizeof(TieHeader) + (size_t)std::max<long long>(cap, kTieFirstBatch) * sizeof(TieRec));
}

static int scm_materialize(kv_shard *s);

static int scm_scan_async(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos, int64_t n_neg,
                          double p, int64_t ubs, int64_t n_blocks) {
    if (s->reuse_armed && s->scm_valid && s->scm_lazy) KV_TRY(scm_materialize(s));   // (before d_misc is laid out below)
    if (n_pos < 0 || n_neg < 0 || n_pos > s->n_examples || n_neg > s->n_examples) {
        s->reuse_armed = false;
        return fail(KV_E_INVALID, "kv_scm_scan: n_pos / n_neg out of range");
    }
    if (ubs <= 0) return fail(KV_E_INVALID, "kv_scm_scan: util_block_size must be positive");
    if (n_blocks != kv_scm_n_blocks(s->k_total, ubs)) return fail(KV_E_INVALID, "kv_scm_scan: wrong n_blocks");
    KV_TRY(ensure_counts(s, 2));
    if (!s->d_segmax) CU(cudaMalloc(&s->d_segmax, (size_t)(s->ld / 64) * 2 * 8));
    MiscLayout L0 = misc_layout(nullptr, n_blocks);
    KV_TRY(ensure(s->d_misc, L0.bytes));
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    // L.blockmax is all-zero here: zeroed at allocation and re-zeroed by whoever consumed it last -- as long as the
    // layout stays the same.  Another block count (another util_block_size) moves the accumulators over what used to
    // be block maxima / tolerances: start from a clean buffer.
    if (s->misc_n_blocks != n_blocks) {
        CU(cudaMemsetAsync(s->d_misc.p, 0, s->d_misc.bytes, s->stream));
        s->misc_n_blocks = n_blocks;
    }
    if (s->reuse_armed) {
        // kv_scm_reuse_counts: same masks as the previous scan (possibly with swapped roles) -> rescore only
        s->reuse_armed = false;
        if (!s->scm_valid) return fail(KV_E_INVALID, "kv_scm_reuse_counts: no SCM counts on the device");
        const bool swapped = s->reuse_swap;      // relative to the masks of the scan that produced the counts
        RescoreParams R;
        memset(&R, 0, sizeof R);
        R.cnt_pos = swapped ? s->scm_cnt_b : s->scm_cnt_a;
        R.cnt_neg = swapped ? s->scm_cnt_a : s->scm_cnt_b;
        R.n_cols = s->n_cols;
        R.ld = s->ld;
        R.col0 = s->col0;
        R.k_total = s->k_total;
        R.ubs = ubs;
        R.p = p;
        R.n_pos = (uint32_t)n_pos;
        R.n_neg = (uint32_t)n_neg;
        R.black_pres = s->have_black ? s->d_black_pres : nullptr;
        R.black_abs = s->have_black ? s->d_black_abs : nullptr;
        R.blockmax = L.blockmax;
        R.segmax = s->d_segmax;
        if (s->profiling) CU(cudaEventRecord(s->ev0, s->stream));
        k_scm_rescore<<<(unsigned)((s->ld + 2047) / 2048), 256, 0, s->stream>>>(R);
        count_launch(s);
        if (s->profiling) CU(cudaEventRecord(s->ev1, s->stream));
        s->scm_swapped = swapped;
        s->rescore_calls += 1;
    } else {
        if ((pos_mask == nullptr) != (neg_mask == nullptr)) return fail(KV_E_INVALID, "kv_scm_scan: null mask");
        int n_act, n_both;
        Scan2Params P;
        if (!pos_mask) {                                         // the device-resident greedy state
            if (!s->state_valid) return fail(KV_E_INVALID, "kv_scm_scan: no greedy state on the device (kv_scm_state_set)");
            n_act = s->state_n_act;
            n_both = s->state_n_both;
            n_pos = s->state_n_pos;
            n_neg = s->state_n_neg;
            fill_scan2_state(s, P);
        } else {
            KV_TRY(push_records2(s, pos_mask, neg_mask, n_act, n_both));
            fill_scan2(s, P, n_act);
        }
        P.cnt0 = s->d_cnt[0];
        P.cnt1 = s->d_cnt[1];
        P.p = p;
        P.n_pos = (uint32_t)n_pos;
        P.n_neg = (uint32_t)n_neg;
        P.ubs = ubs;
        P.black_pres = s->have_black ? s->d_black_pres : nullptr;
        P.black_abs = s->have_black ? s->d_black_abs : nullptr;
        P.blockmax = L.blockmax;
        P.segmax = s->d_segmax;
        KV_TRY(launch_scan<EPI_SCM>(s, P, n_both));
        s->scm_cnt_a = s->d_cnt[0];
        s->scm_cnt_b = s->d_cnt[1];
        s->scm_swapped = false;
        s->scan_calls += 1;
    }
    s->scm_valid = true;
    s->gini_valid = false;
    s->scm_n_pos = n_pos;
    s->scm_n_neg = n_neg;
    s->scm_ubs = ubs;
    s->scm_p = p;
    return KV_OK;
}

// The last SCM step streamed its candidates and wrote no count arrays (EPI_SCM_STREAM): scan the same masks once more
// with the counting kernel, so that the count arrays, segment maxima and scm_* bookkeeping are what a counting step
// would have left.  Only the rare consumers of resident counts pay for it (kv_scm_reuse_counts, kv_scm_ties after a
// packet overflow).
static int scm_materialize(kv_shard *s) {
    if (!s->scm_lazy) return KV_OK;
    s->scm_lazy = false;
    const bool armed = s->reuse_armed;
    s->reuse_armed = false;
    const uint64_t *pm = s->lazy_from_state ? nullptr : s->lazy_masks.data();
    const uint64_t *nm = pm ? pm + s->W : nullptr;
    const int64_t nb = kv_scm_n_blocks(s->k_total, s->scm_ubs);
    KV_TRY(scm_scan_async(s, pm, nm, s->scm_n_pos, s->scm_n_neg, s->scm_p, s->scm_ubs, nb));
    MiscLayout L = misc_layout(s->d_misc.p, nb);
    CU(cudaMemsetAsync(L.blockmax, 0, (size_t)nb * 8, s->stream));    // nobody reads these maxima: leave the accumulators clean
    s->reuse_armed = armed;
    return KV_OK;
}

extern "C" int kv_scm_reuse_counts(kv_shard *s, int swap_roles) {
    if (!s) return fail(KV_E_INVALID, "kv_scm_reuse_counts: null shard");
    if (!s->scm_valid) return fail(KV_E_INVALID, "kv_scm_reuse_counts: no SCM counts on the device");
    s->reuse_armed = true;
    s->reuse_swap = swap_roles != 0;
    return KV_OK;
}

extern "C" int kv_scm_scan(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                           int64_t n_neg, double p, int64_t util_block_size, double *block_max, int64_t n_blocks) {
    if (!s || !block_max) return fail(KV_E_INVALID, "kv_scm_scan: null pointer");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(scm_scan_async(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks));
    KV_TRY(ensure_pinned(s, (size_t)n_blocks * 8));
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    CU(cudaMemcpyAsync(s->h_pin, L.blockmax, (size_t)n_blocks * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemsetAsync(L.blockmax, 0, (size_t)n_blocks * 8, s->stream));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    KV_TRY(finish_scan_timing(s));
    const unsigned long long *enc = (const unsigned long long *)s->h_pin;
    for (int64_t b = 0; b < n_blocks; ++b) block_max[b] = dec_f64_max(enc[b]);
    return KV_OK;
}

// numpy.isclose(a, b) for scalars, rtol=1e-5, atol=1e-8 (numpy 2.3: |a-b| <= atol + rtol*|b| and
// isfinite(b), or a == b).  Compiled with -ffp-contract=off: one rounded multiply, one rounded add.
static inline double np_tol(double b) {
    volatile double t = 1e-5 * std::fabs(b);
    volatile double tol = 1e-8 + t;
    return tol;
}
static inline bool np_isclose(double a, double b) {
    if (a == b) return true;
    if (!std::isfinite(b)) return false;
    return std::fabs(a - b) <= np_tol(b);
}

extern "C" int kv_scm_select_blocks(const double *block_max, int64_t n_blocks, double *best_utility,
                                    uint8_t *selected) {
    if (!block_max || !best_utility || !selected || n_blocks <= 0)
        return fail(KV_E_INVALID, "kv_scm_select_blocks: bad arguments");
    double best = -std::numeric_limits<double>::infinity();
    memset(selected, 0, (size_t)n_blocks);
    for (int64_t b = 0; b < n_blocks; ++b) {
        const double m = block_max[b];
        if (m > best || np_isclose(best, m)) {          // scm.py:271
            if (np_isclose(m, best)) {                  // scm.py:276: append, best_utility unchanged
                selected[b] = 1;
            } else {                                    // scm.py:282-286: replace
                best = m;
                memset(selected, 0, (size_t)n_blocks);
                selected[b] = 1;
            }
        }
    }
    *best_utility = best;
    return KV_OK;
}

// Launch k_scm_ties on the block tables already in d_misc, then bring header + records back with as
// few round trips as possible and sort by rule index (the kernel appends in arbitrary order; the
// reference's order is ascending rule index, scm.py:273-280).
static void launch_exact_ties(kv_shard *s, int64_t n_blocks, TieHeader *d_hdr, TieRec *d_rec, long long cap);
static int scm_read_ties(kv_shard *s, TieHeader *d_hdr, TieRec *d_rec, int64_t capacity, double *best_utility,
                         int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover, int64_t *n_found);

static int scm_collect_ties(kv_shard *s, int64_t n_blocks, int64_t capacity, double *best_utility, int64_t *rule_idx,
                            int64_t *pos_err, int64_t *neg_cover, int64_t *n_found) {
    TieHeader *d_hdr = (TieHeader *)s->d_ties.p;
    TieRec *d_rec = (TieRec *)(d_hdr + 1);
    launch_exact_ties(s, n_blocks, d_hdr, d_rec, std::max<long long>(capacity, kTieFirstBatch));
    return scm_read_ties(s, d_hdr, d_rec, capacity, best_utility, rule_idx, pos_err, neg_cover, n_found);
}

// k_scm_ties in exact mode on the block tables in d_misc, appending into (d_hdr, d_rec[cap])
static void launch_exact_ties(kv_shard *s, int64_t n_blocks, TieHeader *d_hdr, TieRec *d_rec, long long cap) {
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    TieParams T;
    memset(&T, 0, sizeof T);
    T.cnt0 = s->scm_swapped ? s->scm_cnt_b : s->scm_cnt_a;
    T.cnt1 = s->scm_swapped ? s->scm_cnt_a : s->scm_cnt_b;
    T.segmax = s->d_segmax;
    T.n_cols = s->n_cols;
    T.ld = s->ld;
    T.col0 = s->col0;
    T.k_total = s->k_total;
    T.ubs = s->scm_ubs;
    T.inv_ubs = 1.0 / (double)s->scm_ubs;
    T.p = s->scm_p;
    T.n_pos = (uint32_t)s->scm_n_pos;
    T.n_neg = (uint32_t)s->scm_n_neg;
    T.black_pres = s->have_black ? s->d_black_pres : nullptr;
    T.black_abs = s->have_black ? s->d_black_abs : nullptr;
    T.bm = L.bm;
    T.tol = L.tol;
    T.sel = L.sel;
    T.count = &d_hdr->count;
    T.threshold = nullptr;
    T.out = d_rec;
    T.cap = cap;
    const long long n_seg = (s->n_cols + 63) / 64;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n_seg + 255) / 256, (long long)s->sm_count * 4));
    k_scm_ties<<<grid, 256, 0, s->stream>>>(T);
    count_launch(s);
}

// header + records back to the host with as few round trips as possible, ascending rule index
static int scm_read_ties(kv_shard *s, TieHeader *d_hdr, TieRec *d_rec, int64_t capacity, double *best_utility,
                         int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover, int64_t *n_found) {
    const size_t first_bytes = sizeof(TieHeader) + (size_t)kTieFirstBatch * sizeof(TieRec);
    KV_TRY(ensure_pinned(s, first_bytes));
    CU(cudaMemcpyAsync(s->h_pin, d_hdr, first_bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    const TieHeader hdr = *(const TieHeader *)s->h_pin;
    const long long found = (long long)hdr.count;
    *n_found = found;
    if (best_utility) *best_utility = hdr.best_utility;
    if (found > capacity) return fail(KV_E_OVERFLOW, "tie buffer too small");
    if (found == 0) return KV_OK;
    if (found > kTieFirstBatch) {
        // split {rule, (pos_err, neg_cover)} -> radix sort by rule -> copy out in order
        KV_TRY(ensure(s->d_sort_io, (size_t)found * 32));
        long long *k_in = (long long *)s->d_sort_io.p, *k_out = k_in + found;
        unsigned long long *v_in = (unsigned long long *)(k_out + found), *v_out = v_in + found;
        const unsigned g2 = (unsigned)std::min<long long>((found + 255) / 256, (long long)s->sm_count * 8);
        k_split_tierecs<<<g2, 256, 0, s->stream>>>(d_rec, found, k_in, v_in);
        count_launch(s);
        KV_TRY(device_sort(s, k_in, k_out, v_in, v_out, found));
        KV_TRY(ensure_pinned(s, (size_t)found * 8));
        CU(cudaMemcpyAsync(rule_idx, k_out, (size_t)found * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(s->h_pin, v_out, (size_t)found * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        const unsigned long long *v = (const unsigned long long *)s->h_pin;
        for (long long i = 0; i < found; ++i) {
            pos_err[i] = (int64_t)(v[i] >> 32);
            neg_cover[i] = (int64_t)(v[i] & 0xFFFFFFFFull);
        }
        return KV_OK;
    }
    TieRec *recs = (TieRec *)((TieHeader *)s->h_pin + 1);
    std::sort(recs, recs + found, [](const TieRec &a, const TieRec &b) { return a.idx < b.idx; });
    for (long long i = 0; i < found; ++i) {
        rule_idx[i] = recs[i].idx;
        pos_err[i] = recs[i].pe;
        neg_cover[i] = recs[i].nc;
    }
    return KV_OK;
}

extern "C" int kv_scm_ties(kv_shard *s, const double *block_max, const uint8_t *selected, int64_t n_blocks,
                           int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                           int64_t *n_found) {
    if (!s || !block_max || !selected || !n_found || (capacity > 0 && (!rule_idx || !pos_err || !neg_cover)))
        return fail(KV_E_INVALID, "kv_scm_ties: null pointer");
    if (!s->scm_valid) return fail(KV_E_INVALID, "kv_scm_ties: no SCM scan result on the device");
    if (n_blocks != kv_scm_n_blocks(s->k_total, s->scm_ubs)) return fail(KV_E_INVALID, "kv_scm_ties: wrong n_blocks");
    if (capacity < 0) return fail(KV_E_INVALID, "kv_scm_ties: negative capacity");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(scm_materialize(s));
    // merged block table (host) -> pinned -> device: bm | tol | sel are contiguous in d_misc
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    const size_t tab_bytes = (size_t)n_blocks * 16 + ((size_t)n_blocks + 15) / 16 * 16;
    KV_TRY(ensure_pinned(s, tab_bytes + sizeof(TieHeader)));
    double *bm = (double *)s->h_pin, *tol = bm + n_blocks;
    uint8_t *sel = (uint8_t *)(tol + n_blocks);
    for (int64_t b = 0; b < n_blocks; ++b) {
        bm[b] = block_max[b];
        tol[b] = np_tol(block_max[b]);
        sel[b] = selected[b];
    }
    TieHeader *h_hdr = (TieHeader *)((char *)s->h_pin + tab_bytes);
    memset(h_hdr, 0, sizeof(TieHeader));
    KV_TRY(ensure_ties(s, capacity));
    CU(cudaMemcpyAsync(L.bm, s->h_pin, tab_bytes, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_ties.p, h_hdr, sizeof(TieHeader), cudaMemcpyHostToDevice, s->stream));
    // the pinned buffer is reused for the results below
    CU(cudaStreamSynchronize(s->stream));
    return scm_collect_ties(s, n_blocks, capacity, nullptr, rule_idx, pos_err, neg_cover, n_found);
}

// ---- device-resident greedy state --------------------------------------------------------------------------------
static int ensure_state(kv_shard *s) {
    if (s->d_state) return KV_OK;
    const size_t W = (size_t)s->W;
    CU(cudaMalloc(&s->d_state, W * 8 * 5 + W * 4 + 64));
    CU(cudaMemset(s->d_state, 0, W * 8 * 5 + W * 4 + 64));
    CU(cudaHostAlloc(&s->h_state, 64 + W * 8, cudaHostAllocMapped));
    memset(s->h_state, 0, 64 + W * 8);
    CU(cudaDeviceSynchronize());
    return KV_OK;
}

// launch k_mask_and_column and wait for its epoch word; fills the shard's state counters
static int run_mask_kernel(kv_shard *s, long long col, const uint64_t *d_words, int invert) {
    const long long W = s->W;
    if (s->scm_lazy && s->lazy_from_state) {     // the masks of the last streamed step are about to change: its counts
        s->scm_lazy = false;                     // can no longer be produced on demand
        s->scm_valid = false;
    }
    MaskColParams M;
    memset(&M, 0, sizeof M);
    M.mat = s->d_mat;
    M.ld = s->ld;
    M.col = col;
    M.words = d_words;
    M.invert = invert;
    M.W = W;
    M.pos = s->d_state;
    M.neg = s->d_state + W;
    M.rec_m0 = s->d_state + 3 * W;
    M.rec_m1 = s->d_state + 4 * W;
    M.rec_row = (uint32_t *)(s->d_state + 5 * W);
    M.host = (volatile unsigned long long *)s->h_state;
    M.epoch = ++s->state_epoch;
    k_mask_and_column<<<1, 1024, 0, s->stream>>>(M);
    count_launch(s);
    volatile unsigned long long *h = (volatile unsigned long long *)s->h_state;
    unsigned spins = 0;
    while (h[4] != M.epoch) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0x7FF) == 0) {
            const cudaError_t e = cudaStreamQuery(s->stream);
            if (e == cudaErrorNotReady) continue;
            if (h[4] == M.epoch) break;
            if (e == cudaSuccess) return fail(KV_E_CUDA, "k_mask_and_column retired without publishing its result");
            CU(e);
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    s->state_n_pos = (long long)h[0];
    s->state_n_neg = (long long)h[1];
    s->state_n_act = (int)h[2];
    s->state_n_both = (int)h[3];
    s->state_valid = true;
    return KV_OK;
}

extern "C" int kv_scm_state_set(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t *n_pos,
                                int64_t *n_neg) {
    if (!s || !pos_mask || !neg_mask) return fail(KV_E_INVALID, "kv_scm_state_set: null pointer");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(ensure_state(s));
    const size_t W = (size_t)s->W;
    KV_TRY(ensure_pinned(s, 2 * W * 8));
    CU(cudaStreamSynchronize(s->stream));                       // an in-flight copy may still read h_pin
    memcpy(s->h_pin, pos_mask, W * 8);
    memcpy((char *)s->h_pin + W * 8, neg_mask, W * 8);
    CU(cudaMemcpyAsync(s->d_state, s->h_pin, 2 * W * 8, cudaMemcpyHostToDevice, s->stream));
    KV_TRY(run_mask_kernel(s, -1, nullptr, 0));
    if (n_pos) *n_pos = s->state_n_pos;
    if (n_neg) *n_neg = s->state_n_neg;
    return KV_OK;
}

extern "C" int kv_mask_and_column(kv_shard *s, int64_t local_col, const uint64_t *col_words_in, int invert,
                                  uint64_t *col_words_out, int64_t *n_pos, int64_t *n_neg) {
    if (!s) return fail(KV_E_INVALID, "kv_mask_and_column: null shard");
    if (!s->state_valid) return fail(KV_E_INVALID, "kv_mask_and_column: no greedy state on the device (kv_scm_state_set)");
    if (local_col >= s->n_cols) return fail(KV_E_INVALID, "kv_mask_and_column: column outside the shard");
    if (local_col < 0 && !col_words_in) return fail(KV_E_INVALID, "kv_mask_and_column: neither a column nor its words");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    const size_t W = (size_t)s->W;
    uint64_t *d_words = nullptr;
    if (local_col < 0) {
        KV_TRY(ensure_pinned(s, W * 8));
        CU(cudaStreamSynchronize(s->stream));
        memcpy(s->h_pin, col_words_in, W * 8);
        d_words = s->d_state + 2 * W;
        CU(cudaMemcpyAsync(d_words, s->h_pin, W * 8, cudaMemcpyHostToDevice, s->stream));
    }
    KV_TRY(run_mask_kernel(s, local_col < 0 ? -1 : local_col, d_words, invert != 0));
    if (col_words_out) memcpy(col_words_out, (const char *)s->h_state + 64, W * 8);
    if (n_pos) *n_pos = s->state_n_pos;
    if (n_neg) *n_neg = s->state_n_neg;
    return KV_OK;
}

// scan parameters over the resident greedy state's records (device memory)
static void fill_scan2_state(kv_shard *s, Scan2Params &P) {
    const long long W = s->W;
    memset(&P, 0, sizeof P);
    P.mat = s->d_mat;
    P.ld = s->ld;
    P.n_cols = s->n_cols;
    P.col0 = s->col0;
    P.k_total = s->k_total;
    P.n_act = s->state_n_act;
    P.l2_hints = s->l2_hints;
    P.csa = s->scan_csa;
    P.inline_rec = 0;
    P.rec_m0 = s->d_state + 3 * W;
    P.rec_m1 = s->d_state + 4 * W;
    P.rec_row = (const uint32_t *)(s->d_state + 5 * W);
    P.ubs = 1;
}

// ---- the single-shard step in ONE launch (FusedTail) -------------------------------------------------------------
static constexpr long long kFusedCap = 4096;                 // tie records the mapped result buffer holds
static constexpr size_t kFusedHdrBytes = 64;

static int ensure_fused(kv_shard *s, int64_t n_blocks) {
    if (!s->h_fused) {
        // header | tie records | merged block maxima | selected flags (the last two: multi-rank mode)
        CU(cudaHostAlloc(&s->h_fused, kFusedHdrBytes + (size_t)kFusedCap * sizeof(TieRec) + (size_t)FUSED_MAX_BLOCKS * 9,
                         cudaHostAllocMapped));
        memset(s->h_fused, 0, kFusedHdrBytes);
        CU(cudaMalloc(&s->d_fused_ctr, 3 * 128));
        CU(cudaMemset(s->d_fused_ctr, 0, 3 * 128));
        CU(cudaDeviceSynchronize());
    }
    if (s->fused_n_blocks != n_blocks) {                     // both parities zeroed: the kernels keep them that way
        CU(cudaStreamSynchronize(s->stream));
        if (s->d_fused_bm) cudaFree(s->d_fused_bm);
        s->d_fused_bm = nullptr;
        CU(cudaMalloc(&s->d_fused_bm, (size_t)n_blocks * 2 * 8));
        CU(cudaMemset(s->d_fused_bm, 0, (size_t)n_blocks * 2 * 8));
        CU(cudaDeviceSynchronize());
        s->fused_n_blocks = n_blocks;
    }
    return KV_OK;
}

// scan + block merge + tie collection in one kernel, results through mapped host memory: one launch, no copy, the
// host polls the epoch flag the last CTA writes.  Falls back to the tie kernel when more than kFusedCap rules tie.
struct P2PTail {                 // what the multi-rank variant of the fused step adds
    const P2PDev *dev;
    unsigned long long seq, timeout_ns;
    char *packet;
    long long packet_cap;
    int status = 0;              // out: 0 ok, 1 peer timeout, 2 packet / result overflow
};

static int scm_best_rules_fused(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                                int64_t n_neg, double p, int64_t ubs, int64_t n_blocks, double *best_utility,
                                int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                                int64_t *n_found, P2PTail *p2p_tail = nullptr, bool allow_stream = true) {
    if (n_pos < 0 || n_neg < 0 || n_pos > s->n_examples || n_neg > s->n_examples)
        return fail(KV_E_INVALID, "kv_scm_scan: n_pos / n_neg out of range");
    KV_TRY(ensure_fused(s, n_blocks));
    // candidate streaming (EPI_SCM_STREAM) unless switched off, backing off after an overflow, or p is not finite
    bool stream = allow_stream && s->stream_cand && std::isfinite(p) && stream_variant_ok(s);
    if (stream && s->stream_backoff > 0) {
        --s->stream_backoff;
        stream = false;
    }
    if (stream && !s->d_cand) {
        CU(cudaMalloc(&s->d_cand, (size_t)s->sm_count * 2 * kCandCap * sizeof(TieRec)));
        CU(cudaMalloc(&s->d_cand_n, (size_t)s->sm_count * 2 * sizeof(unsigned)));
    }
    const uint64_t *const in_pos = pos_mask, *const in_neg = neg_mask;
    const int64_t in_n_pos = n_pos, in_n_neg = n_neg;
    KV_TRY(ensure_counts(s, 2));
    if (!s->d_segmax) CU(cudaMalloc(&s->d_segmax, (size_t)(s->ld / 64) * 2 * 8));
    KV_TRY(ensure(s->d_misc, misc_layout(nullptr, n_blocks).bytes));
    if (s->misc_n_blocks != n_blocks) {
        CU(cudaMemsetAsync(s->d_misc.p, 0, s->d_misc.bytes, s->stream));
        s->misc_n_blocks = n_blocks;
    }
    const MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    int n_act, n_both;
    Scan2Params P;
    if (!pos_mask) {                                             // the device-resident greedy state
        if (!s->state_valid) return fail(KV_E_INVALID, "kv_scm_best_rules: no greedy state on the device (kv_scm_state_set)");
        n_act = s->state_n_act;
        n_both = s->state_n_both;
        n_pos = s->state_n_pos;
        n_neg = s->state_n_neg;
        fill_scan2_state(s, P);
    } else {
        KV_TRY(push_records2(s, pos_mask, neg_mask, n_act, n_both));
        fill_scan2(s, P, n_act);
    }
    P.cnt0 = s->d_cnt[0];
    P.cnt1 = s->d_cnt[1];
    P.p = p;
    P.n_pos = (uint32_t)n_pos;
    P.n_neg = (uint32_t)n_neg;
    P.ubs = ubs;
    P.black_pres = s->have_black ? s->d_black_pres : nullptr;
    P.black_abs = s->have_black ? s->d_black_abs : nullptr;
    P.segmax = s->d_segmax;
    const unsigned long long epoch = ++s->fused_epoch;
    const int parity = (int)(epoch & 1ull);
    P.blockmax = s->d_fused_bm + (size_t)parity * n_blocks;
    FusedTail &F = P.tail;
    F.mode = 1;
    F.barrier = s->d_fused_ctr;
    F.done = s->d_fused_ctr + 16;
    F.tie_count = s->d_fused_ctr + 32;
    F.zero_next = s->d_fused_bm + (size_t)(parity ^ 1) * n_blocks;
    F.n_blocks = n_blocks;
    F.bm = L.bm;
    F.tol = L.tol;
    F.sel = L.sel;
    F.out = (TieRec *)((char *)s->h_fused + kFusedHdrBytes);      // (mapped with UVA: the host pointer is the device pointer)
    F.cap = kFusedCap;
    F.host = (volatile unsigned long long *)s->h_fused;
    F.epoch = epoch;
    if (p2p_tail) {                                               // one process per GPU: exchange + merge on the device
        F.mode = 2;
        F.p2p = p2p_tail->dev;
        F.p2p_seq = p2p_tail->seq;
        F.timeout_ns = p2p_tail->timeout_ns;
        F.packet = p2p_tail->packet;
        F.packet_cap = p2p_tail->packet_cap;
        F.tie_count = &((PacketHeader *)(p2p_tail->packet + (size_t)n_blocks * 8))->count;
        F.host_bm = (double *)((char *)s->h_fused + kFusedHdrBytes + (size_t)kFusedCap * sizeof(TieRec));
        F.host_sel = (uint8_t *)(F.host_bm + FUSED_MAX_BLOCKS);
    }
    if (stream) {
        P.cand = s->d_cand;
        P.cand_n = s->d_cand_n;
        P.cand_cap = kCandCap;
        KV_TRY(launch_scan_stream(s, P));
        s->lazy_from_state = in_pos == nullptr;
        if (in_pos) {
            s->lazy_masks.resize((size_t)s->W * 2);
            memcpy(s->lazy_masks.data(), in_pos, (size_t)s->W * 8);
            memcpy(s->lazy_masks.data() + s->W, in_neg, (size_t)s->W * 8);
        }
    } else {
        KV_TRY(launch_scan<EPI_SCM>(s, P, n_both));
    }
    s->scm_lazy = stream;
    s->scm_cnt_a = s->d_cnt[0];
    s->scm_cnt_b = s->d_cnt[1];
    s->scm_swapped = false;
    s->scan_calls += 1;
    s->scm_valid = true;
    s->gini_valid = false;
    s->scm_n_pos = n_pos;
    s->scm_n_neg = n_neg;
    s->scm_ubs = ubs;
    s->scm_p = p;
    // wait for the epoch flag (the kernel is then a few hundred nanoseconds from retiring)
    volatile unsigned long long *h = (volatile unsigned long long *)s->h_fused;
    unsigned spins = 0;
    while (h[2] != epoch) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0x7FF) == 0) {
            const cudaError_t e = cudaStreamQuery(s->stream);
            if (e == cudaErrorNotReady) continue;
            if (h[2] == epoch) break;
            if (e == cudaSuccess) return fail(KV_E_CUDA, "fused scan retired without publishing its result");
            CU(e);
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    const unsigned long long best_bits = h[0];
    const long long found = (long long)h[1];
    double best;
    memcpy(&best, &best_bits, 8);
    *best_utility = best;
    *n_found = found;
    if (s->profiling) {
        CU(cudaStreamSynchronize(s->stream));
        KV_TRY(finish_scan_timing(s));
    }
    if (stream) {
        // a candidate list overflowed / every rule blacklisted / more ties than the result buffer holds: this step is
        // redone with the count arrays (single shard: here; ranks: by the two-phase fallback through scm_materialize)
        const bool redo = p2p_tail ? h[3] == 2ull : found > kFusedCap;
        s->stream_steps += 1;
        if (redo) {
            s->stream_redos += 1;
            s->stream_backoff = s->stream_backoff_len;
            s->stream_backoff_len = std::min(s->stream_backoff_len * 2, 1024);
            if (!p2p_tail) {
                if (s->profiling) {
                    CU(cudaStreamSynchronize(s->stream));
                    KV_TRY(finish_scan_timing(s));
                }
                return scm_best_rules_fused(s, in_pos, in_neg, in_n_pos, in_n_neg, p, ubs, n_blocks, best_utility, capacity,
                                            rule_idx, pos_err, neg_cover, n_found, nullptr, false);
            }
        } else {
            s->stream_backoff_len = 8;
        }
    }
    if (p2p_tail) {
        s->exch_publish_us += (double)(h[5] - h[4]) * 1e-3;      // device globaltimer stamps of the last CTA
        s->exch_wait_us += (double)(h[6] - h[5]) * 1e-3;
        s->exch_merge_us += (double)(h[7] - h[6]) * 1e-3;
        s->exch_steps += 1;
        p2p_tail->status = (int)h[3];
        if (p2p_tail->status != 0) return KV_OK;                  // the caller reports the timeout / redoes in two phases
    }
    if (found > kFusedCap) {
        // a huge equivalent-rule set: the counts, segment maxima and block tables are resident -- tie kernel + device sort
        if (found > capacity) return fail(KV_E_OVERFLOW, "tie buffer too small");
        KV_TRY(ensure_ties(s, found));
        TieHeader *d_hdr = (TieHeader *)s->d_ties.p;
        CU(cudaMemsetAsync(d_hdr, 0, sizeof(TieHeader), s->stream));
        launch_exact_ties(s, n_blocks, d_hdr, (TieRec *)(d_hdr + 1), std::max<long long>(found, kTieFirstBatch));
        return scm_read_ties(s, d_hdr, (TieRec *)(d_hdr + 1), capacity, nullptr, rule_idx, pos_err, neg_cover, n_found);
    }
    if (found > capacity) return fail(KV_E_OVERFLOW, "tie buffer too small");
    TieRec *recs = (TieRec *)((char *)s->h_fused + kFusedHdrBytes);
    std::sort(recs, recs + found, [](const TieRec &a, const TieRec &b) { return a.idx < b.idx; });
    for (long long i = 0; i < found; ++i) {
        rule_idx[i] = recs[i].idx;
        pos_err[i] = recs[i].pe;
        neg_cover[i] = recs[i].nc;
    }
    return KV_OK;
}

extern "C" int kv_set_profiling(kv_shard *s, int on) {
    if (!s) return fail(KV_E_INVALID, "kv_set_profiling: null shard");
    s->profiling = on != 0;
    return KV_OK;
}

extern "C" int kv_scm_best_rules(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                                 int64_t n_neg, double p, int64_t util_block_size, double *best_utility,
                                 int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                                 int64_t *n_found) {
    if (!s || !best_utility || !n_found || (capacity > 0 && (!rule_idx || !pos_err || !neg_cover)))
        return fail(KV_E_INVALID, "kv_scm_best_rules: null pointer");
    if (s->col0 != 0 || s->n_cols != s->k_total)
        return fail(KV_E_INVALID, "kv_scm_best_rules: single-shard call on a partial shard; use scan/select/ties");
    if (capacity < 0) return fail(KV_E_INVALID, "kv_scm_best_rules: negative capacity");
    const int64_t n_blocks = kv_scm_n_blocks(s->k_total, util_block_size);
    if (n_blocks <= 0) return fail(KV_E_INVALID, "kv_scm_best_rules: util_block_size must be positive");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    if (s->fused && !s->reuse_armed && n_blocks <= FUSED_MAX_BLOCKS) {
        if ((pos_mask == nullptr) != (neg_mask == nullptr)) return fail(KV_E_INVALID, "kv_scm_scan: null mask");
        return scm_best_rules_fused(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks, best_utility,
                                    capacity, rule_idx, pos_err, neg_cover, n_found);
    }
    // scan -> device-side block merge -> tie collection, one host synchronisation for the whole call
    KV_TRY(ensure_ties(s, capacity));
    KV_TRY(scm_scan_async(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks));
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    k_scm_select<<<1, 32, 0, s->stream>>>(L.blockmax, n_blocks, L.bm, L.tol, L.sel, (TieHeader *)s->d_ties.p);
    count_launch(s);
    const int rc = scm_collect_ties(s, n_blocks, capacity, best_utility, rule_idx, pos_err, neg_cover, n_found);
    if (rc == KV_OK || rc == KV_E_OVERFLOW) finish_scan_timing(s);
    return rc;
}

// Row mask over the packed rows from example indices (rules.py:210-222): example i = bit 63-(i%64) of word i/64;
// duplicates count once.  -> KV_E_INDEX when an index is outside [0, n_examples).
static int fill_row_mask(kv_shard *s, const int64_t *idx, int64_t n, uint64_t *mask) {
    memset(mask, 0, (size_t)s->W * 8);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t e = idx[i];
        if (e < 0 || e >= s->n_examples) return fail(KV_E_INDEX, "index out of bounds");
        mask[e >> 6] |= 1ull << (63 - (e & 63));
    }
    return KV_OK;
}

// kv_scm_best_rules straight from the learner's example index arrays (scm.py:238-253 hands sum_rows the index
// lists): the two row masks are built here instead of in two numpy round trips.
extern "C" int kv_scm_best_rules_idx(kv_shard *s, const int64_t *pos_idx, int64_t n_pos, const int64_t *neg_idx,
                                     int64_t n_neg, double p, int64_t util_block_size, double *best_utility,
                                     int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                                     int64_t *n_found) {
    if (!s || (n_pos > 0 && !pos_idx) || (n_neg > 0 && !neg_idx))
        return fail(KV_E_INVALID, "kv_scm_best_rules_idx: null pointer");
    if (n_pos < 0 || n_neg < 0) return fail(KV_E_INVALID, "kv_scm_best_rules_idx: negative count");
    s->h_masks.resize((size_t)s->W * 2);
    KV_TRY(fill_row_mask(s, pos_idx, n_pos, s->h_masks.data()));
    KV_TRY(fill_row_mask(s, neg_idx, n_neg, s->h_masks.data() + s->W));
    return kv_scm_best_rules(s, s->h_masks.data(), s->h_masks.data() + s->W, n_pos, n_neg, p, util_block_size,
                             best_utility, capacity, rule_idx, pos_err, neg_cover, n_found);
}

// ------------------------------------------------------------------------------------------------
// Job queue: several single-shard SCM scans enqueued back to back on the shard's stream and collected
// with ONE synchronisation -- the lock-step fits of a hyper-parameter grid / cross-validation.  Each
// job's result (header + up to kQueueSlotRecs ties) lands in its own device slot.
// ------------------------------------------------------------------------------------------------
static constexpr long long kQueueSlotRecs = 256;
static constexpr size_t kQueueSlotBytes = sizeof(TieHeader) + kQueueSlotRecs * sizeof(TieRec);

extern "C" int kv_scm_queue_push(kv_shard *s, int64_t slot, int64_t n_slots, const uint64_t *pos_mask,
                                 const uint64_t *neg_mask, int64_t n_pos, int64_t n_neg, double p,
                                 int64_t util_block_size, int reuse) {
    if (!s) return fail(KV_E_INVALID, "kv_scm_queue_push: null shard");
    if (s->col0 != 0 || s->n_cols != s->k_total) return fail(KV_E_INVALID, "kv_scm_queue_push: single-shard only");
    if (slot < 0 || n_slots <= 0 || slot >= n_slots) return fail(KV_E_INVALID, "kv_scm_queue_push: bad slot");
    if (reuse < 0 || reuse > 2) return fail(KV_E_INVALID, "kv_scm_queue_push: reuse must be 0, 1 or 2");
    const int64_t n_blocks = kv_scm_n_blocks(s->k_total, util_block_size);
    if (n_blocks <= 0) return fail(KV_E_INVALID, "kv_scm_queue_push: util_block_size must be positive");
    DeviceGuard g(s->device);
    if (slot == 0) s->last_launches = 0;
    KV_TRY(ensure(s->d_queue, (size_t)n_slots * kQueueSlotBytes));
    if (reuse) {
        KV_TRY(kv_scm_reuse_counts(s, reuse == 2));
    } else if (s->W > SCAN_INLINE_ROWS) {
        // the row records of a large scan are staged through one pinned buffer: let the previous job's copy finish
        CU(cudaStreamSynchronize(s->stream));
    }
    KV_TRY(scm_scan_async(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks));
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    TieHeader *d_hdr = (TieHeader *)((char *)s->d_queue.p + (size_t)slot * kQueueSlotBytes);
    k_scm_select<<<1, 32, 0, s->stream>>>(L.blockmax, n_blocks, L.bm, L.tol, L.sel, d_hdr);
    count_launch(s);
    launch_exact_ties(s, n_blocks, d_hdr, (TieRec *)(d_hdr + 1), kQueueSlotRecs);
    CU(cudaGetLastError());
    return KV_OK;
}

extern "C" int kv_scm_queue_collect(kv_shard *s, int64_t n_slots, double *best_utility, int64_t *n_found,
                                    int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover) {
    if (!s || !best_utility || !n_found || !rule_idx || !pos_err || !neg_cover)
        return fail(KV_E_INVALID, "kv_scm_queue_collect: null pointer");
    if (n_slots <= 0 || (size_t)n_slots * kQueueSlotBytes > s->d_queue.bytes)
        return fail(KV_E_INVALID, "kv_scm_queue_collect: more slots than were pushed");
    DeviceGuard g(s->device);
    const size_t bytes = (size_t)n_slots * kQueueSlotBytes;
    KV_TRY(ensure_pinned(s, bytes));
    CU(cudaMemcpyAsync(s->h_pin, s->d_queue.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    for (int64_t j = 0; j < n_slots; ++j) {
        const TieHeader *h = (const TieHeader *)((const char *)s->h_pin + (size_t)j * kQueueSlotBytes);
        TieRec *recs = (TieRec *)(h + 1);
        best_utility[j] = h->best_utility;
        n_found[j] = (int64_t)h->count;            // > kQueueSlotRecs: truncated, the caller redoes the job alone
        const long long m = std::min<long long>((long long)h->count, kQueueSlotRecs);
        std::sort(recs, recs + m, [](const TieRec &a, const TieRec &b) { return a.idx < b.idx; });
        for (long long i = 0; i < m; ++i) {
            rule_idx[j * kQueueSlotRecs + i] = recs[i].idx;
            pos_err[j * kQueueSlotRecs + i] = recs[i].pe;
            neg_cover[j * kQueueSlotRecs + i] = recs[i].nc;
        }
    }
    return KV_OK;
}

// Score jobs job_ids[0..n) -- all over the SAME mask pair, whose presence counts sit in cnt_a / cnt_b -- from the
// resident counts: per batch of QUEUE_BATCH jobs one k_scm_rescore_multi (the counts are read once for the whole
// batch), one k_scm_select_multi, one k_scm_ties_multi, each job's result landing in its queue slot.
static int queue_score_group(kv_shard *s, const uint32_t *cnt_a, const uint32_t *cnt_b, const int64_t *job_ids, int n,
                             const int64_t *n_pos, const int64_t *n_neg, const double *p, const int32_t *reuse,
                             int64_t ubs, int64_t n_blocks, char *d_packets = nullptr, int64_t packet_capacity = 0) {
    // d_packets != null (partial shards): instead of the exact block merge + ties into the job's queue slot, every job
    // gets a PACKET {block maxima, count, threshold, local superset of the ties} at d_packets + job * packet_bytes
    // (kv_scm_scan_packet's format), to be merged with the other shards' packets on the host.
    static_assert(QUEUE_BATCH == 12, "SelectLocalMultiParams is sized for 12 jobs");
    const int64_t packet_bytes = kv_scm_packet_bytes(n_blocks, packet_capacity);
    const MiscLayout L0 = misc_layout(nullptr, n_blocks);
    const size_t misc_stride = (L0.bytes + 255) / 256 * 256, seg_stride = (size_t)(s->ld / 64) * 2 * 8;
    KV_TRY(ensure(s->d_qmisc, misc_stride * QUEUE_BATCH));
    KV_TRY(ensure(s->d_qseg, seg_stride * QUEUE_BATCH));
    if (s->qmisc_n_blocks != n_blocks) {              // (see scm_scan_async: the accumulators must start zeroed)
        CU(cudaMemsetAsync(s->d_qmisc.p, 0, s->d_qmisc.bytes, s->stream));
        s->qmisc_n_blocks = n_blocks;
    }
    const long long n_seg = (s->n_cols + 63) / 64;
    const unsigned grid_ties = (unsigned)std::max<long long>(1, std::min<long long>((n_seg + 255) / 256, (long long)s->sm_count * 4));
    for (int b0 = 0; b0 < n; b0 += QUEUE_BATCH) {
        const int nb = std::min(QUEUE_BATCH, n - b0);
        RescoreMultiParams R;
        SelectMultiParams S;
        SelectLocalMultiParams SL;
        TieParamsMulti T;
        memset(&R, 0, sizeof R);
        memset(&S, 0, sizeof S);
        memset(&SL, 0, sizeof SL);
        memset(&T, 0, sizeof T);
        SL.n_blocks = n_blocks;
        R.cnt_a = cnt_a;
        R.cnt_b = cnt_b;
        R.n_cols = s->n_cols;
        R.ld = s->ld;
        R.col0 = s->col0;
        R.k_total = s->k_total;
        R.ubs = ubs;
        R.black_pres = s->have_black ? s->d_black_pres : nullptr;
        R.black_abs = s->have_black ? s->d_black_abs : nullptr;
        R.n_jobs = nb;
        S.n_blocks = n_blocks;
        // plain jobs first, then the jobs whose positives are the pair's second mask (fixed roles per kernel loop)
        std::vector<int64_t> order;
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; i < nb; ++i)
                if ((reuse[job_ids[b0 + i]] == 2) == (pass == 1)) order.push_back(job_ids[b0 + i]);
        for (int i = 0; i < nb; ++i) R.n_plain += reuse[order[i]] != 2;
        for (int i = 0; i < nb; ++i) {
            const int64_t j = order[i];
            if (n_pos[j] < 0 || n_neg[j] < 0 || n_pos[j] > s->n_examples || n_neg[j] > s->n_examples)
                return fail(KV_E_INVALID, "kv_scm_queue_run: n_pos / n_neg out of range");
            const bool swap = reuse[j] == 2;
            const MiscLayout L = misc_layout((char *)s->d_qmisc.p + (size_t)i * misc_stride, n_blocks);
            unsigned long long *seg = (unsigned long long *)((char *)s->d_qseg.p + (size_t)i * seg_stride);
            TieHeader *d_hdr = (TieHeader *)((char *)s->d_queue.p + (size_t)j * kQueueSlotBytes);
            R.p[i] = p[j];
            R.n_pos[i] = (double)n_pos[j];
            R.n_neg[i] = (double)n_neg[j];
            R.blockmax[i] = L.blockmax;
            R.segmax[i] = seg;
            S.blockmax[i] = L.blockmax;
            S.bm[i] = L.bm;
            S.tol[i] = L.tol;
            S.sel[i] = L.sel;
            S.hdr[i] = d_hdr;
            TieParams &t = T.j[i];
            t.cnt0 = swap ? cnt_b : cnt_a;
            t.cnt1 = swap ? cnt_a : cnt_b;
            t.segmax = seg;
            t.n_cols = s->n_cols;
            t.ld = s->ld;
            t.col0 = s->col0;
            t.k_total = s->k_total;
            t.ubs = ubs;
            t.inv_ubs = 1.0 / (double)ubs;
            t.p = p[j];
            t.n_pos = (uint32_t)n_pos[j];
            t.n_neg = (uint32_t)n_neg[j];
            t.black_pres = R.black_pres;
            t.black_abs = R.black_abs;
            t.bm = L.bm;
            t.tol = L.tol;
            t.sel = L.sel;
            t.count = &d_hdr->count;
            t.threshold = nullptr;
            t.out = (TieRec *)(d_hdr + 1);
            t.cap = kQueueSlotRecs;
            if (d_packets) {
                char *pk = d_packets + (size_t)j * packet_bytes;
                PacketHeader *ph = (PacketHeader *)(pk + (size_t)n_blocks * 8);
                SL.blockmax[i] = L.blockmax;
                SL.packet_bm[i] = (double *)pk;
                SL.hdr[i] = ph;
                t.bm = t.tol = nullptr;
                t.sel = nullptr;
                t.count = &ph->count;
                t.threshold = &ph->threshold;
                t.out = (TieRec *)(ph + 1);
                t.cap = packet_capacity;
            }
        }
        {   // one mask pair behind all the jobs?  (always, for the queue's own grouping; checked because the sizes are the caller's)
            const bool first_plain = R.n_plain > 0;
            R.dn_a = first_plain ? R.n_pos[0] : R.n_neg[0];
            R.dn_b = first_plain ? R.n_neg[0] : R.n_pos[0];
            R.hoist = 1;
            for (int i = 0; i < nb; ++i) {
                const bool plain = i < R.n_plain;
                if ((plain ? R.n_pos[i] : R.n_neg[i]) != R.dn_a || (plain ? R.n_neg[i] : R.n_pos[i]) != R.dn_b) R.hoist = 0;
            }
            if (const char *v = getenv("KOVER_B200_RESCORE_HOIST")) R.hoist = R.hoist && atoi(v) != 0;
        }
        const unsigned grid_rescore = (unsigned)((s->ld + 2047) / 2048);
        if (R.hoist) k_scm_rescore_multi<true><<<grid_rescore, 256, 0, s->stream>>>(R);
        else k_scm_rescore_multi<false><<<grid_rescore, 256, 0, s->stream>>>(R);
        if (d_packets) k_scm_select_local_multi<<<nb, 32, 0, s->stream>>>(SL);
        else k_scm_select_multi<<<nb, 32, 0, s->stream>>>(S);
        k_scm_ties_multi<<<dim3(grid_ties, nb), 256, 0, s->stream>>>(T);
        count_launch(s, 3);
        s->rescore_calls += nb;
    }
    return KV_OK;
}

// Jobs whose reuse flag is 0 open a group (a new mask pair); the jobs that follow with reuse 1 / 2 re-score that
// group's counts.  One group: the leader is the fused scan (k_scan_tma<EPI_SCM>), its followers are scored from
// the counts it leaves.  Two or more groups: the pairs are counted SEVERAL PER MATRIX PASS by k_count_tma (up to
// MAX_MULTI / 2 pairs: the folds of a cross-validation, the diverged greedy states of a grid) and every job -- the
// group leaders included -- is scored from the resident counts.
static int scm_queue_run_grouped(kv_shard *s, int64_t n_jobs, const uint64_t *masks, const int64_t *n_pos,
                                 const int64_t *n_neg, const double *p, const int32_t *reuse, int64_t util_block_size,
                                 int pairs_per_pass, bool batch_pairs, char *d_packets = nullptr,
                                 int64_t packet_capacity = 0) {
    const int64_t n_blocks = kv_scm_n_blocks(s->k_total, util_block_size);
    if (n_blocks <= 0) return fail(KV_E_INVALID, "kv_scm_queue_run: util_block_size must be positive");
    std::vector<int64_t> leaders;
    for (int64_t j = 0; j < n_jobs; ++j) {
        if (reuse[j] < 0 || reuse[j] > 2) return fail(KV_E_INVALID, "kv_scm_queue_run: reuse must be 0, 1 or 2");
        if (reuse[j] == 0) leaders.push_back(j);
    }
    s->last_launches = 0;
    KV_TRY(ensure(s->d_queue, (size_t)n_jobs * kQueueSlotBytes));
    if (!s->d_segmax) CU(cudaMalloc(&s->d_segmax, (size_t)(s->ld / 64) * 2 * 8));
    std::vector<int64_t> ids;
    if (!batch_pairs) {
        // one matrix pass per group, fused with the leader's utility epilogue
        KV_TRY(ensure_pinned(s, (size_t)n_jobs * kQueueSlotBytes));
        for (size_t l = 0; l < leaders.size(); ++l) {
            const int64_t j0 = leaders[l], j1 = l + 1 < leaders.size() ? leaders[l + 1] : n_jobs;
            if (l > 0 && s->W > SCAN_INLINE_ROWS) CU(cudaStreamSynchronize(s->stream));   // pinned row records are reused
            const uint64_t *pm = masks + (size_t)j0 * 2 * s->W;
            KV_TRY(scm_scan_async(s, pm, pm + s->W, n_pos[j0], n_neg[j0], p[j0], util_block_size, n_blocks));
            MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
            TieHeader *d_hdr = (TieHeader *)((char *)s->d_queue.p + (size_t)j0 * kQueueSlotBytes);
            k_scm_select<<<1, 32, 0, s->stream>>>(L.blockmax, n_blocks, L.bm, L.tol, L.sel, d_hdr);
            count_launch(s);
            launch_exact_ties(s, n_blocks, d_hdr, (TieRec *)(d_hdr + 1), kQueueSlotRecs);
            ids.clear();
            for (int64_t j = j0 + 1; j < j1; ++j) ids.push_back(j);
            if (!ids.empty())
                KV_TRY(queue_score_group(s, s->d_cnt[0], s->d_cnt[1], ids.data(), (int)ids.size(), n_pos, n_neg, p, reuse,
                                         util_block_size, n_blocks));
        }
        CU(cudaGetLastError());
        return KV_OK;
    }
    const size_t n_passes = (leaders.size() + pairs_per_pass - 1) / pairs_per_pass;
    KV_TRY(ensure_counts(s, 2 + 2 * pairs_per_pass));
    KV_TRY(ensure_pinned(s, std::max(n_passes * count_batch_stride(s->W), (size_t)n_jobs * kQueueSlotBytes)));
    KV_TRY(ensure(s->d_rec, count_batch_stride(s->W)));
    CU(cudaStreamSynchronize(s->stream));                 // an earlier call's copies may still read the pinned buffer
    for (size_t pass = 0; pass < n_passes; ++pass) {
        const size_t l0 = pass * pairs_per_pass, l1 = std::min(leaders.size(), l0 + pairs_per_pass);
        std::vector<const uint64_t *> ptrs;
        std::vector<uint32_t *> dst;
        for (size_t l = l0; l < l1; ++l) {
            const uint64_t *pm = masks + (size_t)leaders[l] * 2 * s->W;
            ptrs.push_back(pm);
            ptrs.push_back(pm + s->W);
            dst.push_back(s->d_cnt[2 + 2 * (l - l0)]);
            dst.push_back(s->d_cnt[3 + 2 * (l - l0)]);
        }
        KV_TRY(count_masks_async(s, (int)ptrs.size(), ptrs.data(), dst.data(),
                                 (char *)s->h_pin + pass * count_batch_stride(s->W)));
        s->scan_calls += 1;
        for (size_t l = l0; l < l1; ++l) {
            const int64_t j1 = l + 1 < leaders.size() ? leaders[l + 1] : n_jobs;
            ids.clear();
            for (int64_t j = leaders[l]; j < j1; ++j) ids.push_back(j);
            KV_TRY(queue_score_group(s, dst[2 * (l - l0)], dst[2 * (l - l0) + 1], ids.data(), (int)ids.size(), n_pos, n_neg, p,
                                     reuse, util_block_size, n_blocks, d_packets, packet_capacity));
        }
    }
    s->scm_valid = false;
    CU(cudaGetLastError());
    return KV_OK;
}

extern "C" int kv_scm_queue_run(kv_shard *s, int64_t n_jobs, const uint64_t *masks, const int64_t *n_pos,
                                const int64_t *n_neg, const double *p, const int32_t *reuse, int64_t util_block_size,
                                double *best_utility, int64_t *n_found, int64_t *rule_idx, int64_t *pos_err,
                                int64_t *neg_cover) {
    if (!s || !masks || !n_pos || !n_neg || !p || !reuse) return fail(KV_E_INVALID, "kv_scm_queue_run: null pointer");
    if (n_jobs <= 0) return fail(KV_E_INVALID, "kv_scm_queue_run: no jobs");
    if (s->col0 != 0 || s->n_cols != s->k_total) return fail(KV_E_INVALID, "kv_scm_queue_run: single-shard only");
    if (reuse[0] == 0) {
        int64_t n_groups = 0;
        for (int64_t j = 0; j < n_jobs; ++j) n_groups += reuse[j] == 0;
        const int per = max_masks_per_pass(s->W);
        static const bool no_multi = getenv("KOVER_B200_NO_MULTI") != nullptr;    // measurement knob
        DeviceGuard g(s->device);
        KV_TRY(scm_queue_run_grouped(s, n_jobs, masks, n_pos, n_neg, p, reuse, util_block_size, std::max(1, per / 2),
                                     n_groups >= 2 && per >= 4 && !no_multi));
        return kv_scm_queue_collect(s, n_jobs, best_utility, n_found, rule_idx, pos_err, neg_cover);
    }
    // the first job re-scores counts an earlier call left on the device: one job after the other
    for (int64_t j = 0; j < n_jobs; ++j) {
        const uint64_t *pm = masks + (size_t)j * 2 * s->W, *nm = pm + s->W;
        KV_TRY(kv_scm_queue_push(s, j, n_jobs, pm, nm, n_pos[j], n_neg[j], p[j], util_block_size, reuse[j]));
    }
    return kv_scm_queue_collect(s, n_jobs, best_utility, n_found, rule_idx, pos_err, neg_cover);
}

// The lock-step grid on a PARTIAL shard (several shards per process, or one process per GPU): same grouping as
// kv_scm_queue_run -- distinct mask pairs counted several per matrix pass, every job scored from the counts --
// but each job ends as a packet (kv_scm_scan_packet's format) at packets + job * kv_scm_packet_bytes(...), in host
// memory or (packets_on_device) on the shard's GPU.  The caller gathers the packets of all shards / ranks with ONE
// exchange for the whole grid and merges them job by job with kv_scm_merge_packets.
extern "C" int kv_scm_queue_packets(kv_shard *s, int64_t n_jobs, const uint64_t *masks, const int64_t *n_pos,
                                    const int64_t *n_neg, const double *p, const int32_t *reuse,
                                    int64_t util_block_size, int64_t capacity, void *packets, int packets_on_device) {
    if (!s || !masks || !n_pos || !n_neg || !p || !reuse || !packets)
        return fail(KV_E_INVALID, "kv_scm_queue_packets: null pointer");
    if (n_jobs <= 0 || capacity < 0) return fail(KV_E_INVALID, "kv_scm_queue_packets: bad job count / capacity");
    if (reuse[0] != 0) return fail(KV_E_INVALID, "kv_scm_queue_packets: the first job must open a group (reuse = 0)");
    const int per = std::max(2, max_masks_per_pass(s->W));      // (beyond the staging limit: two-mask passes)
    const int64_t n_blocks = kv_scm_n_blocks(s->k_total, util_block_size);
    if (n_blocks <= 0) return fail(KV_E_INVALID, "kv_scm_queue_packets: util_block_size must be positive");
    const size_t bytes = (size_t)n_jobs * (size_t)kv_scm_packet_bytes(n_blocks, capacity);
    DeviceGuard g(s->device);
    char *d_packets = (char *)packets;
    if (!packets_on_device) {
        KV_TRY(ensure(s->d_stage, bytes));
        d_packets = (char *)s->d_stage.p;
    }
    KV_TRY(scm_queue_run_grouped(s, n_jobs, masks, n_pos, n_neg, p, reuse, util_block_size, per / 2, true, d_packets,
                                 capacity));
    if (!packets_on_device) CU(cudaMemcpyAsync(packets, d_packets, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    float ms = 0;                                   // ev0 / ev1 bracket the (last) multi-mask pass
    if (s->profiling && cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->last_kernel_ms = ms;
    return KV_OK;
}

extern "C" int64_t kv_scm_packet_bytes(int64_t n_blocks, int64_t capacity) {
    if (n_blocks <= 0 || capacity < 0) return 0;
    return n_blocks * 8 + (int64_t)sizeof(PacketHeader) + capacity * (int64_t)sizeof(TieRec);
}

// scan + local select + candidate collection into a packet in DEVICE memory; nothing is synchronised
static int scm_packet_async(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                            int64_t n_neg, double p, int64_t util_block_size, int64_t n_blocks, int64_t capacity,
                            char *d_packet) {
    KV_TRY(scm_scan_async(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks));
    MiscLayout L = misc_layout(s->d_misc.p, n_blocks);
    PacketHeader *d_hdr = (PacketHeader *)(d_packet + (size_t)n_blocks * 8);
    k_scm_select_local<<<1, 32, 0, s->stream>>>(L.blockmax, n_blocks, (double *)d_packet, d_hdr);
    count_launch(s);
    TieParams T;
    memset(&T, 0, sizeof T);
    T.cnt0 = s->scm_swapped ? s->scm_cnt_b : s->scm_cnt_a;
    T.cnt1 = s->scm_swapped ? s->scm_cnt_a : s->scm_cnt_b;
    T.segmax = s->d_segmax;
    T.n_cols = s->n_cols;
    T.ld = s->ld;
    T.col0 = s->col0;
    T.k_total = s->k_total;
    T.ubs = s->scm_ubs;
    T.inv_ubs = 1.0 / (double)s->scm_ubs;
    T.p = s->scm_p;
    T.n_pos = (uint32_t)s->scm_n_pos;
    T.n_neg = (uint32_t)s->scm_n_neg;
    T.black_pres = s->have_black ? s->d_black_pres : nullptr;
    T.black_abs = s->have_black ? s->d_black_abs : nullptr;
    T.count = &d_hdr->count;
    T.threshold = &d_hdr->threshold;
    T.out = (TieRec *)(d_hdr + 1);
    T.cap = capacity;
    const long long n_seg = (s->n_cols + 63) / 64;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n_seg + 255) / 256, (long long)s->sm_count * 4));
    k_scm_ties<<<grid, 256, 0, s->stream>>>(T);
    count_launch(s);
    return KV_OK;
}

// CUDA-event time of the last scan kernel (ev0 / ev1), when profiling is on; the stream must have drained
static int finish_scan_timing(kv_shard *s) {
    if (!s->profiling) return KV_OK;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last_kernel_ms = ms;
    s->scan_ms_total += ms;
    return KV_OK;
}

extern "C" int kv_scm_scan_packet(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                                  int64_t n_neg, double p, int64_t util_block_size, int64_t n_blocks, int64_t capacity,
                                  void *packet, int packet_on_device) {
    if (!s || !packet) return fail(KV_E_INVALID, "kv_scm_scan_packet: null pointer");
    if (capacity < 0) return fail(KV_E_INVALID, "kv_scm_scan_packet: negative capacity");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    const size_t bytes = (size_t)kv_scm_packet_bytes(n_blocks, capacity);
    char *d_packet = (char *)packet;
    if (!packet_on_device) {
        KV_TRY(ensure(s->d_ties, bytes));
        d_packet = (char *)s->d_ties.p;
    }
    KV_TRY(scm_packet_async(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks, capacity, d_packet));
    if (!packet_on_device) CU(cudaMemcpyAsync(packet, d_packet, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    return finish_scan_timing(s);
}

// Host-only: merge of the gathered packets (max of the block maxima, scm.py:270-286 replay, exact
// isclose filter of the candidates with u recomputed as scm.py:263-264 does).
static int merge_packets_strided(const void *rows, size_t stride, int64_t n_packets, int64_t n_blocks,
                                 int64_t packet_capacity, double p, int64_t util_block_size, double *best_utility,
                                 int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                                 int64_t *n_found, double *merged_block_max, uint8_t *selected, int *overflowed);

extern "C" int kv_scm_merge_packets(const void *rows, int64_t n_packets, int64_t n_blocks, int64_t packet_capacity,
                                    double p, int64_t util_block_size, double *best_utility, int64_t capacity,
                                    int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover, int64_t *n_found,
                                    double *merged_block_max, uint8_t *selected, int *overflowed) {
    return merge_packets_strided(rows, (size_t)kv_scm_packet_bytes(n_blocks, packet_capacity), n_packets, n_blocks,
                                 packet_capacity, p, util_block_size, best_utility, capacity, rule_idx, pos_err,
                                 neg_cover, n_found, merged_block_max, selected, overflowed);
}

static int merge_packets_strided(const void *rows, size_t stride, int64_t n_packets, int64_t n_blocks,
                                 int64_t packet_capacity, double p, int64_t util_block_size, double *best_utility,
                                 int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                                 int64_t *n_found, double *merged_block_max, uint8_t *selected, int *overflowed) {
    if (!rows || !best_utility || !n_found || !merged_block_max || !selected || !overflowed ||
        (capacity > 0 && (!rule_idx || !pos_err || !neg_cover)))
        return fail(KV_E_INVALID, "kv_scm_merge_packets: null pointer");
    if (n_packets <= 0 || n_blocks <= 0 || packet_capacity < 0 || util_block_size <= 0 || capacity < 0)
        return fail(KV_E_INVALID, "kv_scm_merge_packets: bad sizes");
    if (stride < (size_t)kv_scm_packet_bytes(n_blocks, packet_capacity))
        return fail(KV_E_INVALID, "kv_scm_merge_packets: stride smaller than a packet");
    const char *base = (const char *)rows;
    const double ninf = -std::numeric_limits<double>::infinity();
    for (int64_t b = 0; b < n_blocks; ++b) merged_block_max[b] = ninf;
    *overflowed = 0;
    for (int64_t r = 0; r < n_packets; ++r) {
        const double *bm = (const double *)(base + r * stride);
        for (int64_t b = 0; b < n_blocks; ++b) merged_block_max[b] = std::max(merged_block_max[b], bm[b]);
        const PacketHeader *h = (const PacketHeader *)(bm + n_blocks);
        if ((int64_t)h->count > packet_capacity) *overflowed = 1;
    }
    KV_TRY(kv_scm_select_blocks(merged_block_max, n_blocks, best_utility, selected));
    *n_found = 0;
    if (*overflowed) return KV_OK;
    std::vector<TieRec> keep;
    for (int64_t r = 0; r < n_packets; ++r) {
        const double *bm = (const double *)(base + r * stride);
        const PacketHeader *h = (const PacketHeader *)(bm + n_blocks);
        const TieRec *recs = (const TieRec *)(h + 1);
        for (unsigned long long i = 0; i < h->count; ++i) {
            TieRec t = recs[i];
            const bool neg_inf = (t.idx & TIE_FLAG_NEG_INF) != 0;
            t.idx &= ~TIE_FLAG_NEG_INF;
            const int64_t b = t.idx / util_block_size;
            if (b < 0 || b >= n_blocks) return fail(KV_E_INVALID, "kv_scm_merge_packets: corrupt packet");
            if (!selected[b]) continue;
            volatile double prod = p * (double)t.pe;              // one rounded multiply ...
            volatile double u = (double)t.nc - prod;              // ... one rounded subtract (scm.py:263-264)
            if (np_isclose(neg_inf ? ninf : (double)u, merged_block_max[b])) keep.push_back(t);
        }
    }
    std::sort(keep.begin(), keep.end(), [](const TieRec &a, const TieRec &b) { return a.idx < b.idx; });
    *n_found = (int64_t)keep.size();
    if ((int64_t)keep.size() > capacity) return fail(KV_E_OVERFLOW, "kv_scm_merge_packets: tie buffer too small");
    for (size_t i = 0; i < keep.size(); ++i) {
        rule_idx[i] = keep[i].idx;
        pos_err[i] = keep[i].pe;
        neg_cover[i] = keep[i].nc;
    }
    return KV_OK;
}

// ------------------------------------------------------------------------------------------------
// Peer mailbox: the all-gather of the packets over NVLink peer memory, fused behind the tie kernel.
// Every rank owns a mailbox [2 parities][world slots][slot_bytes] + flags[world]; after its candidates
// are complete a rank STORES its packet into slot[parity][rank] of every peer's mailbox (remote stores
// over NVLink/NVSwitch, fire and forget), fences, and writes the step's sequence number into
// flags[rank] of every peer; then it waits until its own flags show every peer's sequence number.  One
// kernel does both, so the host synchronises once per step and no NCCL call sits on the step's path.
// (Ranks run on different GPUs; never use this with several ranks time-slicing one GPU.)
// ------------------------------------------------------------------------------------------------
struct P2PParams {
    char *peer_box[P2P_MAX_WORLD];        // mailbox base of every rank, mapped into this process
    unsigned long long *peer_flags[P2P_MAX_WORLD];
    int rank, world;
    unsigned long long seq;
    size_t slot_bytes, n_blocks_bytes;
    const char *packet;                   // local packet (device)
    long long packet_capacity;
    int *error;                           // device int: 1 = timed out waiting for a peer
    unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(256) k_p2p_publish_wait(const P2PParams P) {
    // bytes worth sending: block maxima + header + the records actually present
    const PacketHeader *hdr = (const PacketHeader *)(P.packet + P.n_blocks_bytes);
    const long long n_rec = min((long long)hdr->count, P.packet_capacity);
    const size_t bytes = P.n_blocks_bytes + sizeof(PacketHeader) + (size_t)n_rec * sizeof(TieRec);
    const size_t n16 = (bytes + 15) / 16;
    const int parity = (int)(P.seq & 1ull);
    const uint4 *src = (const uint4 *)P.packet;
    for (int r = 0; r < P.world; ++r) {
        uint4 *dst = (uint4 *)(P.peer_box[r] + ((size_t)parity * P.world + P.rank) * P.slot_bytes);
        for (size_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P.world) {
        // release: the packet stores above are visible before the flag
        unsigned long long *f = P.peer_flags[threadIdx.x] + P.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(P.seq) : "memory");
        // wait for peer threadIdx.x's flag in MY mailbox
        const unsigned long long *mine = P.peer_flags[P.rank] + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        unsigned long long v = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= P.seq) break;
            if (globaltimer_ns() - t0 > P.timeout_ns) {
                *P.error = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

struct kv_p2p {
    int rank = 0, world = 0;
    size_t slot_bytes = 0;
    char *box = nullptr;                   // own mailbox: [2][world][slot_bytes] then flags[world] (256-aligned)
    size_t flags_off = 0, total_bytes = 0;
    char *peer_box[P2P_MAX_WORLD] = {nullptr};
    bool opened[P2P_MAX_WORLD] = {false};
    char *d_packet = nullptr;              // local packet staging
    int *d_error = nullptr;
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;   // waiting for a peer's packet (KOVER_B200_P2P_TIMEOUT_MS)
    void *h_rows = nullptr;                // pinned: gathered rows of one parity
    P2PDev *d_dev = nullptr;               // the peer table for the fused step (device copy)
};

extern "C" int kv_p2p_create(kv_shard *s, int rank, int world, int64_t slot_bytes, void *ipc_handle_out) {
    if (!s || !ipc_handle_out) return fail(KV_E_INVALID, "kv_p2p_create: null pointer");
    if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || slot_bytes <= 0)
        return fail(KV_E_INVALID, "kv_p2p_create: bad rank / world / slot size");
    DeviceGuard g(s->device);
    if (s->p2p) return fail(KV_E_INVALID, "kv_p2p_create: mailbox already exists");
    kv_p2p *m = new kv_p2p();
    if (const char *v = getenv("KOVER_B200_P2P_TIMEOUT_MS"))
        if (atoll(v) > 0) m->timeout_ns = (unsigned long long)atoll(v) * 1000ull * 1000ull;
    m->rank = rank;
    m->world = world;
    m->slot_bytes = ((size_t)slot_bytes + 255) / 256 * 256;
    m->flags_off = 2 * (size_t)world * m->slot_bytes;
    m->total_bytes = m->flags_off + 256 + (size_t)world * 8;
    if (cudaMalloc(&m->box, m->total_bytes) != cudaSuccess || cudaMalloc(&m->d_packet, m->slot_bytes) != cudaSuccess ||
        cudaMalloc(&m->d_error, 4) != cudaSuccess ||
        cudaMallocHost(&m->h_rows, (size_t)world * m->slot_bytes) != cudaSuccess) {
        delete m;
        return fail(KV_E_NOMEM, "kv_p2p_create: allocation failed");
    }
    CU(cudaMemset(m->box, 0, m->total_bytes));
    CU(cudaMemset(m->d_error, 0, 4));
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, m->box));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &h, 64);
    m->peer_box[rank] = m->box;
    s->p2p = m;
    return KV_OK;
}

extern "C" int kv_p2p_connect(kv_shard *s, const void *all_handles) {
    if (!s || !all_handles || !s->p2p) return fail(KV_E_INVALID, "kv_p2p_connect: no mailbox");
    DeviceGuard g(s->device);
    kv_p2p *m = s->p2p;
    for (int r = 0; r < m->world; ++r) {
        if (r == m->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * 64, 64);
        void *ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        m->peer_box[r] = (char *)ptr;
        m->opened[r] = true;
    }
    P2PDev dev;
    memset(&dev, 0, sizeof dev);
    for (int r = 0; r < m->world; ++r) {
        dev.peer_box[r] = m->peer_box[r];
        dev.peer_flags[r] = (unsigned long long *)(m->peer_box[r] + m->flags_off + 256);
    }
    dev.rank = m->rank;
    dev.world = m->world;
    dev.slot_bytes = m->slot_bytes;
    if (!m->d_dev) CU(cudaMalloc(&m->d_dev, sizeof(P2PDev)));
    CU(cudaMemcpy(m->d_dev, &dev, sizeof dev, cudaMemcpyHostToDevice));
    return KV_OK;
}

static void p2p_destroy(kv_p2p *m) {
    if (!m) return;
    for (int r = 0; r < m->world; ++r)
        if (m->opened[r]) cudaIpcCloseMemHandle(m->peer_box[r]);
    cudaFree(m->box);
    cudaFree(m->d_packet);
    cudaFree(m->d_error);
    if (m->d_dev) cudaFree(m->d_dev);
    if (m->h_rows) cudaFreeHost(m->h_rows);
    delete m;
}

// One multi-rank SCM step: scan, local candidates, publish to every peer + wait for theirs (one kernel),
// read back the gathered packets, merge on the host.  All ranks must call it in the same order.
extern "C" int kv_scm_best_rules_p2p(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                                     int64_t n_neg, double p, int64_t util_block_size, int64_t packet_capacity,
                                     double *best_utility, int64_t capacity, int64_t *rule_idx, int64_t *pos_err,
                                     int64_t *neg_cover, int64_t *n_found, double *merged_block_max,
                                     uint8_t *selected, int *overflowed) {
    if (!s || !s->p2p) return fail(KV_E_INVALID, "kv_scm_best_rules_p2p: no mailbox / null pointer");
    kv_p2p *m = s->p2p;
    const int64_t n_blocks = kv_scm_n_blocks(s->k_total, util_block_size);
    if (n_blocks <= 0) return fail(KV_E_INVALID, "kv_scm_best_rules_p2p: util_block_size must be positive");
    const size_t bytes = (size_t)kv_scm_packet_bytes(n_blocks, packet_capacity);
    if (bytes > m->slot_bytes) return fail(KV_E_INVALID, "kv_scm_best_rules_p2p: packet larger than the mailbox slot");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    if (s->fused && !s->reuse_armed && n_blocks <= FUSED_MAX_BLOCKS && m->d_dev &&
        (int64_t)m->world * packet_capacity <= kFusedCap) {
        // ONE kernel per step and rank: scan, candidates, publish into every peer's mailbox, wait, merge on the device
        if ((pos_mask == nullptr) != (neg_mask == nullptr) || !best_utility || !n_found || !merged_block_max || !selected ||
            !overflowed)
            return fail(KV_E_INVALID, "kv_scm_best_rules_p2p: null pointer");
        P2PTail t;
        t.dev = m->d_dev;
        t.seq = ++m->seq;
        t.timeout_ns = m->timeout_ns;
        t.packet = m->d_packet;
        t.packet_cap = packet_capacity;
        *overflowed = 0;
        const int rc = scm_best_rules_fused(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks, best_utility,
                                            capacity, rule_idx, pos_err, neg_cover, n_found, &t);
        if (rc != KV_OK && rc != KV_E_OVERFLOW) return rc;
        if (t.status == 1) return fail(KV_E_CUDA, "kv_scm_best_rules_p2p: timed out waiting for a peer rank's packet");
        const double *hb = (const double *)((const char *)s->h_fused + kFusedHdrBytes + (size_t)kFusedCap * sizeof(TieRec));
        const uint8_t *hs = (const uint8_t *)(hb + FUSED_MAX_BLOCKS);
        memcpy(merged_block_max, hb, (size_t)n_blocks * 8);
        memcpy(selected, hs, (size_t)n_blocks);
        if (t.status == 2) {
            *overflowed = 1;
            *n_found = 0;
            return KV_OK;
        }
        return rc;
    }
    KV_TRY(scm_packet_async(s, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, n_blocks, packet_capacity,
                            m->d_packet));
    P2PParams P;
    memset(&P, 0, sizeof P);
    for (int r = 0; r < m->world; ++r) {
        P.peer_box[r] = m->peer_box[r];
        P.peer_flags[r] = (unsigned long long *)(m->peer_box[r] + m->flags_off + 256);
    }
    P.rank = m->rank;
    P.world = m->world;
    P.seq = ++m->seq;
    P.slot_bytes = m->slot_bytes;
    P.n_blocks_bytes = (size_t)n_blocks * 8;
    P.packet = m->d_packet;
    P.packet_capacity = packet_capacity;
    P.error = m->d_error;
    P.timeout_ns = m->timeout_ns;
    k_p2p_publish_wait<<<1, 256, 0, s->stream>>>(P);
    count_launch(s);
    const int parity = (int)(P.seq & 1ull);
    // gathered slots of this parity in ONE linear copy (slot stride) + the error word
    CU(cudaMemcpyAsync(m->h_rows, m->box + (size_t)parity * m->world * m->slot_bytes, (size_t)m->world * m->slot_bytes,
                       cudaMemcpyDeviceToHost, s->stream));
    KV_TRY(ensure_pinned(s, 64));
    CU(cudaMemcpyAsync(s->h_pin, m->d_error, 4, cudaMemcpyDeviceToHost, s->stream));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    KV_TRY(finish_scan_timing(s));
    if (*(int *)s->h_pin != 0) return fail(KV_E_CUDA, "kv_scm_best_rules_p2p: timed out waiting for a peer rank's packet");
    return merge_packets_strided(m->h_rows, m->slot_bytes, m->world, n_blocks, packet_capacity, p, util_block_size,
                                 best_utility, capacity, rule_idx, pos_err, neg_cover, n_found, merged_block_max,
                                 selected, overflowed);
}

// kv_scm_best_rules_p2p from the learner's example index arrays (see kv_scm_best_rules_idx)
extern "C" int kv_scm_best_rules_p2p_idx(kv_shard *s, const int64_t *pos_idx, int64_t n_pos, const int64_t *neg_idx,
                                         int64_t n_neg, double p, int64_t util_block_size, int64_t packet_capacity,
                                         double *best_utility, int64_t capacity, int64_t *rule_idx, int64_t *pos_err,
                                         int64_t *neg_cover, int64_t *n_found, double *merged_block_max,
                                         uint8_t *selected, int *overflowed) {
    if (!s || (n_pos > 0 && !pos_idx) || (n_neg > 0 && !neg_idx))
        return fail(KV_E_INVALID, "kv_scm_best_rules_p2p_idx: null pointer");
    if (n_pos < 0 || n_neg < 0) return fail(KV_E_INVALID, "kv_scm_best_rules_p2p_idx: negative count");
    s->h_masks.resize((size_t)s->W * 2);
    KV_TRY(fill_row_mask(s, pos_idx, n_pos, s->h_masks.data()));
    KV_TRY(fill_row_mask(s, neg_idx, n_neg, s->h_masks.data() + s->W));
    return kv_scm_best_rules_p2p(s, s->h_masks.data(), s->h_masks.data() + s->W, n_pos, n_neg, p, util_block_size,
                                 packet_capacity, best_utility, capacity, rule_idx, pos_err, neg_cover, n_found,
                                 merged_block_max, selected, overflowed);
}

// ------------------------------------------------------------------------------------------------
// a9 / a11: Gini split scan
// ------------------------------------------------------------------------------------------------
template <int MODE>
static void launch_gini(kv_shard *s, const GiniParams &G) {
    const unsigned grid = (unsigned)std::min<long long>((s->n_cols + 255) / 256, (long long)s->sm_count * 8);
    switch (G.n_classes) {
        case 2: k_gini<2, MODE><<<grid, 256, 0, s->stream>>>(G); break;
        case 3: k_gini<3, MODE><<<grid, 256, 0, s->stream>>>(G); break;
        case 4: k_gini<4, MODE><<<grid, 256, 0, s->stream>>>(G); break;
        default: k_gini<0, MODE><<<grid, 256, 0, s->stream>>>(G); break;
    }
    count_launch(s);
}

// Launch the counting (+ for two classes the fused Gini minimum) of one node; nothing is synchronised.
// d_misc: [0] minimum (enc_f64), [1] tie counter.
static int gini_scan_async(kv_shard *s, int n_classes, const uint64_t *class_masks, const double *altered_prior,
                           const double *n_total, const int64_t *n_node) {
    if (n_classes < 1 || n_classes > MAX_CLASSES) return fail(KV_E_INVALID, "kv_gini_scan: 1..64 classes supported");
    s->scm_valid = false;
    s->gini_valid = false;
    KV_TRY(ensure(s->d_misc, 64));
    unsigned long long *d_min = (unsigned long long *)s->d_misc.p;
    // (d_misc doubles as the SCM block-maxima accumulator, which must be left zeroed: see kv_gini tail)
    CU(cudaMemsetAsync(d_min, 0xFF, 8, s->stream));
    CU(cudaMemsetAsync(d_min + 1, 0, 8, s->stream));
    GiniParams &G = s->gini;
    memset(&G, 0, sizeof G);
    G.n_classes = n_classes;
    for (int c = 0; c < n_classes; ++c) {
        G.prior[c] = altered_prior[c];
        G.n_total[c] = n_total[c];
        G.n_node[c] = (double)n_node[c];
    }
    G.n_cols = s->n_cols;
    G.col0 = s->col0;
    G.black_pres = s->have_black ? s->d_black_pres : nullptr;
    G.min_enc = d_min;
    G.counter = d_min + 1;
    if (n_classes == 2) {
        // two class masks: one pass of the two-mask scan with the Gini score, its minimum and the
        // per-segment minima fused into the epilogue
        KV_TRY(ensure_counts(s, 2));
        if (!s->d_segmax) CU(cudaMalloc(&s->d_segmax, (size_t)(s->ld / 64) * 2 * 8));
        int n_act, n_both;
        KV_TRY(push_records2(s, class_masks, class_masks + s->W, n_act, n_both));
        Scan2Params P;
        fill_scan2(s, P, n_act);
        P.cnt0 = s->d_cnt[0];
        P.cnt1 = s->d_cnt[1];
        P.black_pres = G.black_pres;
        P.segmax = s->d_segmax;
        P.gmin = d_min;
        for (int c = 0; c < 2; ++c) {
            P.g_prior[c] = G.prior[c];
            P.g_ntot[c] = G.n_total[c];
            P.g_nnode[c] = G.n_node[c];
        }
        KV_TRY(launch_scan<EPI_GINI2>(s, P, n_both));
        s->scan_calls += 1;
        G.segmin = s->d_segmax;
    } else {
        KV_TRY(count_masks(s, n_classes, class_masks, 0));
    }
    for (int c = 0; c < n_classes; ++c) G.cnt[c] = s->d_cnt[c];
    if (n_classes != 2) launch_gini<GINI_MIN>(s, G);
    s->gini_valid = true;
    return KV_OK;
}

// d_misc[0..1] were used as min / counter: restore the all-zero state the SCM accumulators expect
static int gini_restore_misc(kv_shard *s) {
    CU(cudaMemsetAsync(s->d_misc.p, 0, 16, s->stream));
    return KV_OK;
}

extern "C" int kv_gini_scan(kv_shard *s, int n_classes, const uint64_t *class_masks, const double *altered_prior,
                            const double *n_total, const int64_t *n_node, double *min_score) {
    if (!s || !class_masks || !altered_prior || !n_total || !n_node || !min_score)
        return fail(KV_E_INVALID, "kv_gini_scan: null pointer");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(gini_scan_async(s, n_classes, class_masks, altered_prior, n_total, n_node));
    KV_TRY(ensure_pinned(s, 64));
    CU(cudaMemcpyAsync(s->h_pin, s->gini.min_enc, 8, cudaMemcpyDeviceToHost, s->stream));
    KV_TRY(gini_restore_misc(s));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    if (n_classes == 2) KV_TRY(finish_scan_timing(s));
    *min_score = dec_f64_min(*(unsigned long long *)s->h_pin);
    return KV_OK;
}

static constexpr long long kGiniFirstBatch = 512;

// ties of the last scan: target from `min_score`, or from the device-resident minimum when from_device
static int gini_ties_impl(kv_shard *s, bool from_device, double min_score, int64_t capacity, int64_t *rule_idx,
                          int64_t *n_found, double *min_out) {
    const long long cap = std::max<long long>(capacity, kGiniFirstBatch);
    // d_ties: [0] minimum copy, [1] counter, then the indices
    KV_TRY(ensure(s->d_ties, 16 + (size_t)cap * 8));
    unsigned long long *d_hdr = (unsigned long long *)s->d_ties.p;
    GiniParams G = s->gini;
    G.min_score = min_score;
    G.min_ptr = from_device ? (const double *)d_hdr : nullptr;
    G.counter = d_hdr + 1;
    G.out_idx = (long long *)(d_hdr + 2);
    G.cap = cap;
    if (from_device) CU(cudaMemcpyAsync(d_hdr, s->gini.min_enc, 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemsetAsync(d_hdr + 1, 0, 8, s->stream));
    if (G.segmin && G.n_classes == 2) {
        const long long n_seg = (s->n_cols + 63) / 64;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n_seg + 255) / 256, (long long)s->sm_count * 4));
        k_gini2_ties_seg<<<grid, 256, 0, s->stream>>>(G);
        count_launch(s);
    } else {
        launch_gini<GINI_TIES>(s, G);
    }
    const size_t first = 16 + (size_t)kGiniFirstBatch * 8;
    KV_TRY(ensure_pinned(s, first));
    CU(cudaMemcpyAsync(s->h_pin, d_hdr, first, cudaMemcpyDeviceToHost, s->stream));
    if (from_device) KV_TRY(gini_restore_misc(s));
    CU(stream_sync_spin(s->stream));
    CU(cudaGetLastError());
    const unsigned long long *h = (const unsigned long long *)s->h_pin;
    if (min_out) *min_out = dec_f64_min(h[0]);
    const long long found = (long long)h[1];
    *n_found = found;
    if (found > capacity) return fail(KV_E_OVERFLOW, "kv_gini_ties: tie buffer too small");
    if (found == 0) return KV_OK;
    if (found > kGiniFirstBatch) {
        KV_TRY(ensure(s->d_sort_io, (size_t)found * 8));
        KV_TRY(device_sort(s, G.out_idx, (long long *)s->d_sort_io.p, nullptr, nullptr, found));
        CU(cudaMemcpyAsync(rule_idx, s->d_sort_io.p, (size_t)found * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        return KV_OK;
    }
    memcpy(rule_idx, h + 2, (size_t)found * 8);
    std::sort(rule_idx, rule_idx + found);
    return KV_OK;
}

extern "C" int kv_gini_ties(kv_shard *s, double min_score, int64_t capacity, int64_t *rule_idx, int64_t *n_found) {
    if (!s || !n_found || (capacity > 0 && !rule_idx)) return fail(KV_E_INVALID, "kv_gini_ties: null pointer");
    if (!s->gini_valid) return fail(KV_E_INVALID, "kv_gini_ties: no Gini scan result on the device");
    if (capacity < 0) return fail(KV_E_INVALID, "kv_gini_ties: negative capacity");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    return gini_ties_impl(s, false, min_score, capacity, rule_idx, n_found, nullptr);
}

extern "C" int kv_gini_best_split(kv_shard *s, int n_classes, const uint64_t *class_masks, const double *altered_prior,
                                  const double *n_total, const int64_t *n_node, double *min_score, int64_t capacity,
                                  int64_t *rule_idx, int64_t *n_found) {
    if (!s || !class_masks || !altered_prior || !n_total || !n_node || !min_score || !n_found ||
        (capacity > 0 && !rule_idx))
        return fail(KV_E_INVALID, "kv_gini_best_split: null pointer");
    if (capacity < 0) return fail(KV_E_INVALID, "kv_gini_best_split: negative capacity");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(gini_scan_async(s, n_classes, class_masks, altered_prior, n_total, n_node));
    const int rc = gini_ties_impl(s, true, 0.0, capacity, rule_idx, n_found, min_score);
    if (n_classes == 2 && (rc == KV_OK || rc == KV_E_OVERFLOW)) finish_scan_timing(s);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// CART level: the split search of SEVERAL nodes (two classes) with their class masks counted 8 nodes per
// matrix pass.  Nodes of a tree level hold disjoint example sets, so most (packed row, node) mask words are
// empty deep in the tree and are skipped; a level costs about one pass per 8 nodes instead of one per node.
// Per node: k_gini2_counts (minimum + segment minima from the resident counts) and k_gini2_ties_seg.
// ------------------------------------------------------------------------------------------------
static constexpr int kLevelNodesPerPass = MAX_MULTI / 2;
static constexpr long long kLevelSlotIdx = 1024;            // tie indices per node kept in the first-read slot

// room for `need` entries in the pinned result arena, keeping what it holds; the shard's stream must be idle
static int level_arena_reserve(kv_shard *s, size_t need) {
    if (need <= s->h_level_cap) return KV_OK;
    const size_t want = std::max({need, s->h_level_cap * 2, (size_t)1 << 20});
    int64_t *p = nullptr;
    CU(cudaMallocHost(&p, want * 8));
    if (s->h_level_used) memcpy(p, s->h_level, s->h_level_used * 8);
    if (s->h_level) cudaFreeHost(s->h_level);
    s->h_level = p;
    s->h_level_cap = want;
    return KV_OK;
}

static constexpr int kMaxOccTables = 64;

// Device-resident k-mer occurrence table `table` (0..63) = presence counts of the shard's k-mers under `mask` (the
// training set of a tree): the input of the reference's CART tiebreaker (experiment_cart.py:82-94, 264), kept on
// the device so that kv_gini_level can apply the tiebreaker there.  mask = NULL frees the table.
extern "C" int kv_occ_table_set(kv_shard *s, int table, const uint64_t *mask) {
    if (!s) return fail(KV_E_INVALID, "kv_occ_table_set: null shard");
    if (table < 0 || table >= kMaxOccTables) return fail(KV_E_INVALID, "kv_occ_table_set: table must be 0..63");
    DeviceGuard g(s->device);
    if ((int)s->occ_tables.size() <= table) s->occ_tables.resize((size_t)table + 1, nullptr);
    if (!mask) {
        if (s->occ_tables[table]) {
            CU(cudaStreamSynchronize(s->stream));
            cudaFree(s->occ_tables[table]);
            s->occ_tables[table] = nullptr;
        }
        return KV_OK;
    }
    if (!s->occ_tables[table]) CU(cudaMalloc(&s->occ_tables[table], (size_t)s->ld * sizeof(uint32_t)));
    s->last_launches = 0;
    int n_act, n_both;
    KV_TRY(push_records2(s, mask, nullptr, n_act, n_both));
    Scan2Params P;
    fill_scan2(s, P, n_act);
    P.cnt0 = s->occ_tables[table];
    P.cnt1 = nullptr;
    KV_TRY(launch_scan<EPI_COUNTS>(s, P, n_both));
    s->scan_calls += 1;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return KV_OK;
}

// occurrence counts of the given GLOBAL rule indices (presence rules of this shard) from table `table`
extern "C" int kv_occ_table_gather(kv_shard *s, int table, const int64_t *rule_idx, int64_t n, uint32_t *out) {
    if (!s || (n > 0 && (!rule_idx || !out))) return fail(KV_E_INVALID, "kv_occ_table_gather: null pointer");
    if (table < 0 || table >= (int)s->occ_tables.size() || !s->occ_tables[table])
        return fail(KV_E_INVALID, "kv_occ_table_gather: no such table");
    DeviceGuard g(s->device);
    CU(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < n; ++i) {
        const int64_t c = rule_idx[i] - s->col0;
        if (c < 0 || c >= s->n_cols) return fail(KV_E_INVALID, "kv_occ_table_gather: rule outside the shard");
        CU(cudaMemcpy(out + i, s->occ_tables[table] + c, 4, cudaMemcpyDeviceToHost));
    }
    return KV_OK;
}

// ---- sibling subtraction (kv_gini_level_family) --------------------------------------------------------------------
static void fam_release_prev(kv_shard *s) {
    for (auto &kv : s->fam_prev) s->fam_free.push_back(kv.second.slot);
    s->fam_prev.clear();
}

// one arena for the node count slots, sized once from the memory that is free at the first armed level search
static void fam_arena_init(kv_shard *s) {
    if (s->fam_tried) return;
    s->fam_tried = true;
    if (const char *v = getenv("KOVER_B200_LEVEL_SUB"))
        if (atoi(v) == 0) return;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    double frac = 0.25;
    long long cap = 320;
    if (const char *v = getenv("KOVER_B200_LEVEL_POOL_FRAC")) frac = std::min(0.8, std::max(0.0, atof(v)));
    if (const char *v = getenv("KOVER_B200_LEVEL_POOL_SLOTS")) cap = std::max(0, atoi(v));
    const size_t slot_bytes = (size_t)2 * s->ld * sizeof(uint32_t);
    const long long n = std::min<long long>(cap, (long long)(frac * (double)free_b / (double)slot_bytes));
    if (n < 3) return;
    if (cudaMalloc(&s->d_fam_arena, (size_t)n * slot_bytes) != cudaSuccess) {
        cudaGetLastError();
        s->d_fam_arena = nullptr;
        return;
    }
    s->fam_slots = (int)n;
    for (int i = (int)n - 1; i >= 0; --i) s->fam_free.push_back(i);
}

// Arms the NEXT kv_gini_level call of the shard: node_key[i] >= 0 names node i (unique over the life of the shard's
// current trees), parent_key[i] the node it was split from (-1: none).  Two nodes of the call with the same parent whose
// counts are still resident (it was searched by the previous armed call) are siblings: only the smaller one is counted
// in the matrix pass, the other one's class counts are parent - sibling.  n_nodes = 0 forgets the resident level,
// n_nodes < 0 also returns the arena's memory.
extern "C" int kv_gini_level_family(kv_shard *s, int64_t n_nodes, const int64_t *node_key, const int64_t *parent_key) {
    if (!s) return fail(KV_E_INVALID, "kv_gini_level_family: null shard");
    s->fam_key.clear();
    s->fam_parent.clear();
    if (n_nodes <= 0 || !node_key) {
        fam_release_prev(s);
        if (n_nodes < 0 && s->d_fam_arena) {
            DeviceGuard g(s->device);
            CU(cudaStreamSynchronize(s->stream));
            cudaFree(s->d_fam_arena);
            s->d_fam_arena = nullptr;
            s->fam_slots = 0;
            s->fam_free.clear();
            s->fam_tried = false;
        }
        return KV_OK;
    }
    s->fam_key.assign(node_key, node_key + n_nodes);
    if (parent_key) s->fam_parent.assign(parent_key, parent_key + n_nodes);
    else s->fam_parent.assign((size_t)n_nodes, -1);
    return KV_OK;
}

extern "C" int kv_gini_level(kv_shard *s, int64_t n_nodes, const uint64_t *class_masks, const double *altered_prior,
                             const double *n_total, const int64_t *n_node, const int32_t *occ_table, double *min_score,
                             int64_t *n_found, uint32_t *occ_max) {
    if (!s || !class_masks || !altered_prior || !n_total || !n_node || !min_score || !n_found)
        return fail(KV_E_INVALID, "kv_gini_level: null pointer");
    if (n_nodes <= 0) return fail(KV_E_INVALID, "kv_gini_level: no nodes");
    if (occ_table)
        for (int64_t i = 0; i < n_nodes; ++i)
            if (occ_table[i] >= 0 && (occ_table[i] >= (int)s->occ_tables.size() || !s->occ_tables[occ_table[i]]))
                return fail(KV_E_INVALID, "kv_gini_level: unknown occurrence table");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    s->scm_valid = s->gini_valid = false;
    s->level_span.assign((size_t)n_nodes, std::pair<size_t, size_t>(0, 0));
    s->h_level_used = 0;
    const int per = std::max(1, std::min(kLevelNodesPerPass, max_masks_per_pass(s->W) / 2));
    const int per2 = 2 * per;                 // nodes per batch: up to `per` counted + their derived siblings
    const long long W = s->W, n_seg_ld = s->ld / 64;

    // ---- plan: which nodes are counted, which are parent - sibling, where the counts of each node go ----
    std::vector<long long> key, parent;
    const bool armed = (int64_t)s->fam_key.size() == n_nodes;
    if (armed) {
        key.swap(s->fam_key);
        parent.swap(s->fam_parent);
        fam_arena_init(s);
    } else {
        fam_release_prev(s);                  // an unarmed search ends the family: nothing refers to the old level any more
    }
    s->fam_key.clear();
    s->fam_parent.clear();
    std::vector<int> slot((size_t)n_nodes, -1), derived_sib((size_t)n_nodes, -1), parent_slot((size_t)n_nodes, -1);
    std::vector<char> is_derived((size_t)n_nodes, 0);
    if (armed && s->fam_slots) {
        std::unordered_map<long long, std::vector<int>> kids;
        for (int64_t i = 0; i < n_nodes; ++i)
            if (key[i] >= 0 && parent[i] >= 0 && s->fam_prev.count(parent[i])) kids[parent[i]].push_back((int)i);
        for (int64_t i = 0; i < n_nodes; ++i) {                    // (node order: which pairs get the last slots is deterministic)
            if (parent[i] < 0) continue;
            auto it = kids.find(parent[i]);
            if (it == kids.end() || it->second.size() != 2 || it->second[0] != (int)i) continue;
            const int a = it->second[0], b = it->second[1];
            const kv_shard::FamNode &pn = s->fam_prev[parent[i]];
            if (pn.n[0] != n_node[a * 2] + n_node[b * 2] || pn.n[1] != n_node[a * 2 + 1] + n_node[b * 2 + 1]) continue;
            if (s->fam_free.empty()) break;
            const long long na = n_node[a * 2] + n_node[a * 2 + 1], nb_ = n_node[b * 2] + n_node[b * 2 + 1];
            const int cnt = na <= nb_ ? a : b, der = na <= nb_ ? b : a;
            slot[der] = s->fam_free.back();
            s->fam_free.pop_back();
            is_derived[der] = 1;
            derived_sib[cnt] = der;
            parent_slot[der] = pn.slot;
        }
        for (int64_t i = 0; i < n_nodes && !s->fam_free.empty(); ++i)
            if (!is_derived[i] && key[i] >= 0) {
                slot[i] = s->fam_free.back();
                s->fam_free.pop_back();
            }
    }
    struct SlotGuard {                        // an error return gives the slots taken above back
        kv_shard *s;
        std::vector<int> *slot;
        bool keep = false;
        ~SlotGuard() {
            if (!keep)
                for (int sl : *slot)
                    if (sl >= 0) s->fam_free.push_back(sl);
        }
    } slot_guard{s, &slot};
    std::vector<int> counted;
    for (int64_t i = 0; i < n_nodes; ++i)
        if (!is_derived[i]) counted.push_back((int)i);
    auto slot_ptr = [&](int sl, int c) { return s->d_fam_arena + ((size_t)sl * 2 + c) * (size_t)s->ld; };

    KV_TRY(ensure_counts(s, 2 + 2 * per));
    // d_level: min enc [per2] | counter [per2] | ties [per2][kLevelSlotIdx] | segmin [per2][ld/64] | score tables
    //          [per2][4096] | occurrence maxima u32 [2 * per2]
    const size_t level_bytes = (size_t)per2 * 16 + (size_t)per2 * kLevelSlotIdx * 8 + (size_t)per2 * n_seg_ld * 8 +
                               (size_t)per2 * GINI2_TABLE_MAX * 8 + (size_t)per2 * 8;
    KV_TRY(ensure(s->d_level, level_bytes));
    unsigned long long *d_min = (unsigned long long *)s->d_level.p, *d_cntr = d_min + per2;
    long long *d_ties = (long long *)(d_cntr + per2);
    unsigned long long *d_segmin = (unsigned long long *)(d_ties + (size_t)per2 * kLevelSlotIdx);
    double *d_tables = (double *)(d_segmin + (size_t)per2 * n_seg_ld);
    unsigned int *d_occmax = (unsigned int *)(d_tables + (size_t)per2 * GINI2_TABLE_MAX);
    const size_t hdr_bytes = (size_t)per2 * 16 + (size_t)per2 * kLevelSlotIdx * 8;
    KV_TRY(ensure_pinned(s, count_batch_stride(W) + hdr_bytes + (size_t)per2 * 8));
    CU(cudaStreamSynchronize(s->stream));
    char *pin_rec = (char *)s->h_pin;
    const unsigned long long *h_hdr = (const unsigned long long *)((char *)s->h_pin + count_batch_stride(W));
    const long long n_seg = (s->n_cols + 63) / 64;
    const unsigned grid_score = (unsigned)std::max<long long>(1, std::min<long long>((n_seg + 7) / 8, (long long)s->sm_count * 8));
    const unsigned grid_ties = (unsigned)std::max<long long>(1, std::min<long long>((n_seg + 255) / 256, (long long)s->sm_count * 4));
    long long gtable_max = std::min<long long>(GINI2_GTABLE_MAX, s->n_cols / 2);
    if (const char *v = getenv("KOVER_B200_GINI_GTABLE")) gtable_max = std::min<long long>(gtable_max, atoll(v));
    const unsigned grid_sub = (unsigned)std::max<long long>(1, std::min<long long>((s->ld / 4 + 255) / 256, (long long)s->sm_count * 8));
    float scan_ms = 0;
    for (size_t b0 = 0; b0 < counted.size(); b0 += (size_t)per) {
        const int nc = (int)std::min<size_t>((size_t)per, counted.size() - b0);
        // the batch's nodes: the counted ones, then the siblings derived from them
        std::vector<int> node;
        std::vector<int> from;                                       // derived: batch index of the counted sibling; else -1
        std::vector<const uint64_t *> ptrs;
        std::vector<uint32_t *> dst;
        for (int i = 0; i < nc; ++i) {
            const int nd = counted[b0 + i];
            node.push_back(nd);
            from.push_back(-1);
            const uint64_t *m = class_masks + (size_t)nd * 2 * W;
            ptrs.push_back(m);
            ptrs.push_back(m + W);
            dst.push_back(slot[nd] >= 0 ? slot_ptr(slot[nd], 0) : s->d_cnt[2 + 2 * i]);
            dst.push_back(slot[nd] >= 0 ? slot_ptr(slot[nd], 1) : s->d_cnt[3 + 2 * i]);
        }
        for (int i = 0; i < nc; ++i)
            if (derived_sib[node[i]] >= 0) {
                const int nd = derived_sib[node[i]];
                node.push_back(nd);
                from.push_back(i);
                dst.push_back(slot_ptr(slot[nd], 0));
                dst.push_back(slot_ptr(slot[nd], 1));
            }
        const int nb = (int)node.size();
        KV_TRY(count_masks_async(s, 2 * nc, ptrs.data(), dst.data(), pin_rec));
        s->scan_calls += 1;
        s->level_nodes_counted += nc;
        s->level_nodes_derived += nb - nc;
        CU(cudaMemsetAsync(d_min, 0xFF, (size_t)per2 * 8, s->stream));
        CU(cudaMemsetAsync(d_cntr, 0, (size_t)per2 * 8, s->stream));
        CU(cudaMemsetAsync(d_occmax, 0, (size_t)per2 * 8, s->stream));
        std::vector<GiniParams> params((size_t)nb);
        for (int i = 0; i < nb; ++i) {
            const int nd = node[i];
            if (from[i] >= 0) {
                k_sub_counts<<<grid_sub, 256, 0, s->stream>>>(
                    (const uint4 *)slot_ptr(parent_slot[nd], 0), (const uint4 *)slot_ptr(parent_slot[nd], 1),
                    (const uint4 *)dst[2 * from[i]], (const uint4 *)dst[2 * from[i] + 1], (uint4 *)dst[2 * i],
                    (uint4 *)dst[2 * i + 1], s->ld / 4);
                count_launch(s);
            }
            GiniParams &G = params[i];
            memset(&G, 0, sizeof G);
            G.n_classes = 2;
            for (int c = 0; c < 2; ++c) {
                G.cnt[c] = dst[2 * i + c];
                G.prior[c] = altered_prior[nd * 2 + c];
                G.n_total[c] = n_total[nd * 2 + c];
                G.n_node[c] = (double)n_node[nd * 2 + c];
            }
            G.n_cols = s->n_cols;
            G.col0 = s->col0;
            G.black_pres = s->have_black ? s->d_black_pres : nullptr;
            G.min_enc = d_min + i;
            G.min_ptr = (const double *)(d_min + i);
            G.counter = d_cntr + i;
            G.segmin = d_segmin + (size_t)i * n_seg_ld;
            G.out_idx = d_ties + (size_t)i * kLevelSlotIdx;
            G.cap = kLevelSlotIdx;
            const long long pairs = (n_node[nd * 2] + 1) * (n_node[nd * 2 + 1] + 1);
            if (pairs <= GINI2_TABLE_MAX) {
                G.table_stride = (int)n_node[nd * 2 + 1] + 1;
                k_gini2_table<<<(unsigned)((pairs + 255) / 256), 256, 0, s->stream>>>(G, d_tables + (size_t)i * GINI2_TABLE_MAX);
                count_launch(s);
                G.table = d_tables + (size_t)i * GINI2_TABLE_MAX;
            } else if (pairs <= gtable_max) {
                if (!s->d_gtables.p || s->d_gtables.bytes < (size_t)per2 * gtable_max * 8)
                    KV_TRY(ensure(s->d_gtables, (size_t)per2 * gtable_max * 8));
                double *gt = (double *)s->d_gtables.p + (size_t)i * gtable_max;
                G.table_stride = (int)n_node[nd * 2 + 1] + 1;
                k_gini2_table<<<(unsigned)std::min<long long>((pairs + 255) / 256, (long long)s->sm_count * 16), 256, 0, s->stream>>>(G, gt);
                count_launch(s);
                G.table = gt;
                G.table_global = 1;
            }
            k_gini2_counts<<<grid_score, 256, 0, s->stream>>>(G, d_segmin + (size_t)i * n_seg_ld);
            if (occ_table && occ_table[nd] >= 0) {
                // the reference's tiebreaker on the device: largest occurrence among the ties, then the survivors
                G.occ = s->occ_tables[occ_table[nd]];
                G.occ_max = d_occmax + i;
                k_gini2_ties_occmax<<<grid_ties, 256, 0, s->stream>>>(G);
                k_gini2_ties_occ<<<grid_ties, 256, 0, s->stream>>>(G);
                count_launch(s, 3);
            } else {
                k_gini2_ties_seg<<<grid_ties, 256, 0, s->stream>>>(G);
                count_launch(s, 2);
            }
        }
        CU(cudaMemcpyAsync((void *)h_hdr, d_min, hdr_bytes, cudaMemcpyDeviceToHost, s->stream));
        unsigned int *h_occmax = (unsigned int *)((char *)s->h_pin + count_batch_stride(W) + hdr_bytes);
        CU(cudaMemcpyAsync(h_occmax, d_occmax, (size_t)per2 * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(stream_sync_spin(s->stream));
        CU(cudaGetLastError());
        float ms = 0;
        if (s->profiling && cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) scan_ms += ms;
        const long long *h_ties = (const long long *)(h_hdr + 2 * per2);
        size_t batch_total = 0, biggest = 0;
        for (int i = 0; i < nb; ++i) {
            batch_total += (size_t)h_hdr[per2 + i];
            if ((long long)h_hdr[per2 + i] > kLevelSlotIdx) biggest = std::max(biggest, (size_t)h_hdr[per2 + i]);
        }
        KV_TRY(level_arena_reserve(s, s->h_level_used + batch_total));
        if (biggest * 16 > s->d_sort_io.bytes) KV_TRY(ensure(s->d_sort_io, std::max(biggest * 16, s->d_sort_io.bytes * 2)));
        bool pending = false;
        for (int i = 0; i < nb; ++i) {
            const int nd = node[i];
            const long long found = (long long)h_hdr[per2 + i];
            min_score[nd] = dec_f64_min(h_hdr[i]);
            n_found[nd] = found;
            if (occ_max) occ_max[nd] = h_occmax[i];
            int64_t *out = s->h_level + s->h_level_used;
            s->level_span[(size_t)nd] = std::pair<size_t, size_t>(s->h_level_used, (size_t)found);
            s->h_level_used += (size_t)found;
            if (found == 0) continue;
            if (found <= kLevelSlotIdx) {
                memcpy(out, h_ties + (size_t)i * kLevelSlotIdx, (size_t)found * 8);
                std::sort(out, out + found);
                continue;
            }
            // a node with few examples ties on a large share of the k-mers: its counts are still resident, collect
            // again into a buffer of the now known size, sort on the device, copy into the pinned arena
            GiniParams G = params[i];
            long long *raw = (long long *)s->d_sort_io.p, *sorted = raw + found;
            G.out_idx = raw;
            G.cap = found;
            CU(cudaMemsetAsync(d_cntr + i, 0, 8, s->stream));
            if (G.occ) k_gini2_ties_occ<<<grid_ties, 256, 0, s->stream>>>(G);      // (its maximum is still in d_occmax)
            else k_gini2_ties_seg<<<grid_ties, 256, 0, s->stream>>>(G);
            count_launch(s);
            KV_TRY(device_sort(s, raw, sorted, nullptr, nullptr, found));
            CU(cudaMemcpyAsync(out, sorted, (size_t)found * 8, cudaMemcpyDeviceToHost, s->stream));
            pending = true;
        }
        if (pending) {
            CU(stream_sync_spin(s->stream));
            CU(cudaGetLastError());
        }
    }
    // the previous level's counts are not needed any more; this level's become the resident generation
    fam_release_prev(s);
    slot_guard.keep = true;
    if (armed)
        for (int64_t i = 0; i < n_nodes; ++i)
            if (slot[i] >= 0) {
                auto it = s->fam_prev.find(key[i]);
                if (it != s->fam_prev.end()) {                      // a repeated key: the later node wins
                    s->fam_free.push_back(it->second.slot);
                    s->fam_prev.erase(it);
                }
                s->fam_prev[key[i]] = kv_shard::FamNode{slot[i], {n_node[i * 2], n_node[i * 2 + 1]}};
            }
    s->last_kernel_ms = scan_ms;
    s->scan_ms_total += scan_ms;
    return KV_OK;
}

extern "C" int kv_gini_level_result(kv_shard *s, int64_t node, const int64_t **rule_idx, int64_t *n_found) {
    if (!s || !rule_idx || !n_found) return fail(KV_E_INVALID, "kv_gini_level_result: null pointer");
    if (node < 0 || (size_t)node >= s->level_span.size()) return fail(KV_E_INVALID, "kv_gini_level_result: no such node");
    *rule_idx = s->h_level + s->level_span[(size_t)node].first;
    *n_found = (int64_t)s->level_span[(size_t)node].second;
    return KV_OK;
}

extern "C" int kv_gini_level_fetch(kv_shard *s, int64_t node, int64_t *rule_idx, int64_t capacity) {
    if (!s || (capacity > 0 && !rule_idx)) return fail(KV_E_INVALID, "kv_gini_level_fetch: null pointer");
    if (node < 0 || (size_t)node >= s->level_span.size()) return fail(KV_E_INVALID, "kv_gini_level_fetch: no such node");
    const std::pair<size_t, size_t> &sp = s->level_span[(size_t)node];
    if ((int64_t)sp.second > capacity) return fail(KV_E_OVERFLOW, "kv_gini_level_fetch: buffer too small");
    if (sp.second) memcpy(rule_idx, s->h_level + sp.first, sp.second * 8);
    return KV_OK;
}

extern "C" int kv_gini_scores(kv_shard *s, double *out) {
    if (!s || !out) return fail(KV_E_INVALID, "kv_gini_scores: null pointer");
    if (!s->gini_valid) return fail(KV_E_INVALID, "kv_gini_scores: no Gini scan result on the device");
    DeviceGuard g(s->device);
    s->last_launches = 0;
    KV_TRY(ensure(s->d_stage, (size_t)s->n_cols * 8));
    GiniParams G = s->gini;
    G.scores = (double *)s->d_stage.p;
    launch_gini<GINI_SCORES>(s, G);
    CU(cudaMemcpyAsync(out, G.scores, (size_t)s->n_cols * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return KV_OK;
}
