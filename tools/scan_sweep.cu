// scan_sweep -- design-space sweep for the masked-popcount column scan on B200 (sm_100a).
//
// Standalone (no library dependency): builds a [W, K] packed matrix on the device and times kernel
// variants that all compute, per column, popcount(word & mask0) and popcount(word & mask1) summed
// over the W rows.  Every variant is checked against variant 0 (checksum of all counts).  Output:
// one line per variant with GB/s of algorithmic bytes (8*W*K) and the fraction of --peak.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/scan_sweep.cu -o build/scan_sweep
//   build/scan_sweep [W=32] [K=10000000] [general=0|1] [peak_gbs=6538.3]
//
// general=0: label-sorted masks (each row has exactly one non-zero mask word); general=1: both
// masks non-zero in every row (the 4-popc-per-word case).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);        \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void k_fill(uint64_t *m, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) m[i] = mix((uint64_t)i);
}

__device__ __forceinline__ ulonglong2 ldg128(const uint64_t *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
struct u64x4 {
    uint64_t a, b, c, d;
};
__device__ __forceinline__ u64x4 ldg256(const uint64_t *p) {
    u64x4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d)
                 : "l"(p));
    return v;
}

struct P {
    const uint64_t *mat;
    long long ld, n_tiles;
    const uint64_t *m0, *m1;   // [W]
    int W;
    uint32_t *c0, *c1;
};

// ------------------------------------------------------------------------------------------------
// Family A: direct LDG.128, UN loads in flight per thread, THREADS threads, 2 columns per thread
// ------------------------------------------------------------------------------------------------
template <int THREADS, int UN, int MINB, bool INTERLEAVE>
__global__ void __launch_bounds__(THREADS, MINB) k_ldg(const P p) {
    extern __shared__ uint64_t sm[];
    uint64_t *s0 = sm, *s1 = sm + p.W;
    for (int i = threadIdx.x; i < p.W; i += THREADS) {
        s0[i] = p.m0[i];
        s1[i] = p.m1[i];
    }
    __syncthreads();
    constexpr int TC = THREADS * 2;
    long long t0, t1, ts;
    if (INTERLEAVE) {
        t0 = blockIdx.x; t1 = p.n_tiles; ts = gridDim.x;
    } else {
        t0 = (p.n_tiles * (long long)blockIdx.x) / gridDim.x;
        t1 = (p.n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
        ts = 1;
    }
    for (long long tile = t0; tile < t1; tile += ts) {
        const long long col = tile * TC + 2 * threadIdx.x;
        const uint64_t *base = p.mat + col;
        uint32_t ax = 0, ay = 0, bx = 0, by = 0;
        for (int i = 0; i < p.W; i += UN) {
            ulonglong2 v[UN];
#pragma unroll
            for (int j = 0; j < UN; ++j) v[j] = ldg128(base + (long long)(i + j) * p.ld);
#pragma unroll
            for (int j = 0; j < UN; ++j) {
                const uint64_t a = s0[i + j], b = s1[i + j];
                if (a) { ax += __popcll(v[j].x & a); ay += __popcll(v[j].y & a); }
                if (b) { bx += __popcll(v[j].x & b); by += __popcll(v[j].y & b); }
            }
        }
        *reinterpret_cast<uint2 *>(p.c0 + col) = make_uint2(ax, ay);
        *reinterpret_cast<uint2 *>(p.c1 + col) = make_uint2(bx, by);
    }
}

// Family A2: 256-bit loads, 4 columns per thread
template <int THREADS, int UN, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_ldg256(const P p) {
    extern __shared__ uint64_t sm[];
    uint64_t *s0 = sm, *s1 = sm + p.W;
    for (int i = threadIdx.x; i < p.W; i += THREADS) {
        s0[i] = p.m0[i];
        s1[i] = p.m1[i];
    }
    __syncthreads();
    constexpr int TC = THREADS * 4;
    const long long nt = (p.n_tiles * 512) / TC;
    const long long t0 = (nt * (long long)blockIdx.x) / gridDim.x, t1 = (nt * (long long)(blockIdx.x + 1)) / gridDim.x;
    for (long long tile = t0; tile < t1; ++tile) {
        const long long col = tile * TC + 4 * threadIdx.x;
        const uint64_t *base = p.mat + col;
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        for (int i = 0; i < p.W; i += UN) {
            u64x4 v[UN];
#pragma unroll
            for (int j = 0; j < UN; ++j) v[j] = ldg256(base + (long long)(i + j) * p.ld);
#pragma unroll
            for (int j = 0; j < UN; ++j) {
                const uint64_t a = s0[i + j], b = s1[i + j];
                if (a) { a0 += __popcll(v[j].a & a); a1 += __popcll(v[j].b & a); a2 += __popcll(v[j].c & a); a3 += __popcll(v[j].d & a); }
                if (b) { b0 += __popcll(v[j].a & b); b1 += __popcll(v[j].b & b); b2 += __popcll(v[j].c & b); b3 += __popcll(v[j].d & b); }
            }
        }
        *reinterpret_cast<uint4 *>(p.c0 + col) = make_uint4(a0, a1, a2, a3);
        *reinterpret_cast<uint4 *>(p.c1 + col) = make_uint4(b0, b1, b2, b3);
    }
}

// Family A3: register double buffer -- the next UN loads are issued before the current UN are consumed
template <int THREADS, int UN, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_ldg_pipe(const P p) {
    extern __shared__ uint64_t sm[];
    uint64_t *s0 = sm, *s1 = sm + p.W;
    for (int i = threadIdx.x; i < p.W; i += THREADS) {
        s0[i] = p.m0[i];
        s1[i] = p.m1[i];
    }
    __syncthreads();
    constexpr int TC = THREADS * 2;
    const long long t0 = (p.n_tiles * (long long)blockIdx.x) / gridDim.x;
    const long long t1 = (p.n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
    if (t0 >= t1) return;
    const int nb = p.W / UN;                       // W multiple of UN in this tool
    const long long total = (t1 - t0) * nb;        // batches of UN rows, across tiles
    ulonglong2 cur[UN], nxt[UN];
    const uint64_t *base = p.mat + t0 * TC + 2 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < UN; ++j) cur[j] = ldg128(base + (long long)j * p.ld);
    uint32_t ax = 0, ay = 0, bx = 0, by = 0;
    long long tile = t0;
    int b = 0;
    for (long long it = 0; it < total; ++it) {
        // prefetch batch it+1
        int nb_b = b + 1;
        long long nb_tile = tile;
        if (nb_b == nb) { nb_b = 0; ++nb_tile; }
        if (it + 1 < total) {
            const uint64_t *nbase = p.mat + nb_tile * TC + 2 * threadIdx.x + (long long)nb_b * UN * p.ld;
#pragma unroll
            for (int j = 0; j < UN; ++j) nxt[j] = ldg128(nbase + (long long)j * p.ld);
        }
#pragma unroll
        for (int j = 0; j < UN; ++j) {
            const uint64_t a = s0[b * UN + j], m = s1[b * UN + j];
            if (a) { ax += __popcll(cur[j].x & a); ay += __popcll(cur[j].y & a); }
            if (m) { bx += __popcll(cur[j].x & m); by += __popcll(cur[j].y & m); }
        }
        if (b == nb - 1) {
            const long long col = tile * TC + 2 * threadIdx.x;
            *reinterpret_cast<uint2 *>(p.c0 + col) = make_uint2(ax, ay);
            *reinterpret_cast<uint2 *>(p.c1 + col) = make_uint2(bx, by);
            ax = ay = bx = by = 0;
        }
#pragma unroll
        for (int j = 0; j < UN; ++j) cur[j] = nxt[j];
        b = nb_b;
        tile = nb_tile;
    }
}

// ------------------------------------------------------------------------------------------------
// Family B: TMA bulk copies (cp.async.bulk, 1-D) into a shared-memory ring, mbarrier pipeline.
// One producer warp (one elected lane issues), CW consumer warps; stage = R rows x (CW*32*16) bytes.
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int CW, int R, int S>
__global__ void __launch_bounds__((CW + 1) * 32, 1) k_tma(const P p) {
    extern __shared__ __align__(128) unsigned char raw[];
    constexpr int ROW_BYTES = CW * 32 * 16;
    constexpr int TC = CW * 64;
    uint64_t *full = reinterpret_cast<uint64_t *>(raw);
    uint64_t *empty = full + S;
    uint64_t *s0 = empty + S;
    uint64_t *s1 = s0 + p.W;
    unsigned char *ring = raw + ((2 * S + 2 * p.W) * 8 + 127) / 128 * 128;
    for (int i = threadIdx.x; i < p.W; i += blockDim.x) {
        s0[i] = p.m0[i];
        s1[i] = p.m1[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long nt = (p.n_tiles * 512) / TC;
    const long long t0 = (nt * (long long)blockIdx.x) / gridDim.x, t1 = (nt * (long long)(blockIdx.x + 1)) / gridDim.x;
    const int nb = p.W / R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == CW) {                        // producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = t0; tile < t1; ++tile) {
                for (int b = 0; b < nb; ++b) {
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, R * ROW_BYTES);
                    const uint64_t *src = p.mat + tile * TC + (long long)b * R * p.ld;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        bulk_g2s(ring + ((size_t)s * R + r) * ROW_BYTES, src + (long long)r * p.ld, ROW_BYTES, full + s);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {                                 // consumers
        int s = 0;
        uint32_t ph = 0;
        for (long long tile = t0; tile < t1; ++tile) {
            uint32_t ax = 0, ay = 0, bx = 0, by = 0;
            for (int b = 0; b < nb; ++b) {
                mbar_wait(full + s, ph);
                const unsigned char *st = ring + (size_t)s * R * ROW_BYTES + threadIdx.x * 16;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(st + r * ROW_BYTES);
                    const uint64_t a = s0[b * R + r], m = s1[b * R + r];
                    if (a) { ax += __popcll(v.x & a); ay += __popcll(v.y & a); }
                    if (m) { bx += __popcll(v.x & m); by += __popcll(v.y & m); }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
                if (++s == S) { s = 0; ph ^= 1; }
            }
            const long long col = tile * TC + 2 * threadIdx.x;
            *reinterpret_cast<uint2 *>(p.c0 + col) = make_uint2(ax, ay);
            *reinterpret_cast<uint2 *>(p.c1 + col) = make_uint2(bx, by);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// popc / lop3 issue-rate probes
// ------------------------------------------------------------------------------------------------
__global__ void k_popc_rate(uint32_t *out, int iters) {
    uint32_t x0 = threadIdx.x * 2654435761u + 1, x1 = x0 ^ 0x9e3779b9u, x2 = x0 + 77, x3 = x1 + 13, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc += __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3);
            x0 += acc; x1 ^= x0; x2 += x1; x3 ^= x2;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------
static uint64_t checksum(const uint32_t *d, long long n) {
    std::vector<uint32_t> h(n);
    CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost));
    uint64_t s = 0;
    for (long long i = 0; i < n; ++i) s = s * 1000003ull + h[i];
    return s;
}

int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 32;
    const long long K = argc > 2 ? atoll(argv[2]) : 10000000;
    const int general = argc > 3 ? atoi(argv[3]) : 0;
    const double peak = argc > 4 ? atof(argv[4]) : 6538.3;
    const long long n_tiles = (K + 511) / 512, ld = ((n_tiles * 512 + 2047) / 2048) * 2048;   // multiple of every tile width
    const long long nt512 = ld / 512;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs; W=%d K=%lld ld=%lld general=%d; %.3f GB per pass; peak %.1f GB/s\n", prop.name, sms, W, K,
           ld, general, 8.0 * W * ld / 1e9, peak);
    uint64_t *mat;
    CK(cudaMalloc(&mat, (size_t)W * ld * 8));
    k_fill<<<sms * 8, 256>>>(mat, (long long)W * ld);
    std::vector<uint64_t> h0(W), h1(W);
    for (int w = 0; w < W; ++w) {
        const uint64_t r = 0x9E3779B97F4A7C15ull * (w + 1);
        if (general) { h0[w] = r; h1[w] = ~r; }
        else if (w < W / 2) { h0[w] = r | 1; h1[w] = 0; }
        else { h0[w] = 0; h1[w] = r | 1; }
    }
    uint64_t *m0, *m1;
    uint32_t *c0, *c1;
    CK(cudaMalloc(&m0, W * 8)); CK(cudaMalloc(&m1, W * 8));
    CK(cudaMalloc(&c0, ld * 4)); CK(cudaMalloc(&c1, ld * 4));
    CK(cudaMemcpy(m0, h0.data(), W * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m1, h1.data(), W * 8, cudaMemcpyHostToDevice));
    P p{mat, ld, nt512, m0, m1, W, c0, c1};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint64_t ref0 = 0, ref1 = 0;
    bool have_ref = false;
    const double gbytes = 8.0 * W * ld / 1e9;

    auto run = [&](const char *name, auto launch) {
        CK(cudaMemset(c0, 0xAB, ld * 4));
        CK(cudaMemset(c1, 0xAB, ld * 4));
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("%-44s LAUNCH ERROR %s\n", name, cudaGetErrorString(e)); return; }
        float best = 1e30f, sum = 0;
        const int reps = 10;
        for (int i = 0; i < reps; ++i) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
            sum += ms;
        }
        const uint64_t s0 = checksum(c0, ld), s1 = checksum(c1, ld);
        if (!have_ref) { ref0 = s0; ref1 = s1; have_ref = true; }
        const bool ok = s0 == ref0 && s1 == ref1;
        printf("%-44s best %.4f ms  avg %.4f ms  %7.1f GB/s (avg)  %.3f of peak  %s\n", name, best, sum / reps,
               gbytes / (sum / reps) * 1e3, gbytes / (sum / reps) * 1e3 / peak, ok ? "ok" : "MISMATCH");
    };

#define LDG(T, U, B, IL)                                                                          \
    run("ldg128 thr=" #T " un=" #U " ctas/sm=" #B " interleave=" #IL, [&] {                      \
        k_ldg<T, U, B, IL><<<(unsigned)std::min<long long>((long long)sms * B, nt512 * 512 / (T * 2)), T, W * 16>>>( \
            P{mat, ld, nt512 * 512 / (T * 2), m0, m1, W, c0, c1});                                \
    })
    LDG(256, 8, 4, false);
    LDG(256, 8, 4, true);
    LDG(256, 4, 4, false);
    LDG(256, 4, 8, false);
    LDG(256, 8, 3, false);
    LDG(256, 16, 2, false);
    LDG(256, 16, 3, false);
    LDG(128, 8, 8, false);
    LDG(128, 16, 6, false);
    LDG(512, 8, 2, false);
    LDG(512, 4, 4, false);
    LDG(1024, 4, 2, false);
    LDG(1024, 8, 1, false);
#define LDG256(T, U, B)                                                                           \
    run("ldg256 thr=" #T " un=" #U " ctas/sm=" #B, [&] {                                         \
        k_ldg256<T, U, B><<<(unsigned)std::min<long long>((long long)sms * B, nt512 * 512 / (T * 4)), T, W * 16>>>(p); \
    })
    LDG256(256, 4, 4);
    LDG256(256, 8, 2);
    LDG256(128, 4, 8);
    LDG256(128, 8, 4);
    LDG256(512, 4, 2);
#define PIPE(T, U, B)                                                                             \
    run("ldg128-pipelined thr=" #T " un=" #U " ctas/sm=" #B, [&] {                               \
        k_ldg_pipe<T, U, B><<<(unsigned)std::min<long long>((long long)sms * B, nt512 * 512 / (T * 2)), T, W * 16>>>( \
            P{mat, ld, nt512 * 512 / (T * 2), m0, m1, W, c0, c1});                                \
    })
    PIPE(256, 4, 4);
    PIPE(256, 8, 3);
    PIPE(256, 8, 2);
    PIPE(512, 4, 2);
    PIPE(128, 8, 6);
#define TMA(CW, R, S)                                                                             \
    run("tma-bulk consumer_warps=" #CW " rows/stage=" #R " stages=" #S, [&] {                    \
        const size_t smem = ((2 * S + 2 * W) * 8 + 127) / 128 * 128 + (size_t)S * R * CW * 512;   \
        CK(cudaFuncSetAttribute(k_tma<CW, R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_tma<CW, R, S><<<sms, (CW + 1) * 32, smem>>>(p);                                        \
    })
    TMA(8, 8, 4);
    TMA(8, 8, 6);
    TMA(8, 4, 8);
    TMA(8, 4, 12);
    TMA(8, 16, 3);
    TMA(16, 4, 6);
    TMA(16, 8, 3);
    TMA(4, 8, 12);
    TMA(4, 16, 6);
    TMA(8, 2, 24);

    // popc issue rate
    {
        uint32_t *o;
        CK(cudaMalloc(&o, (size_t)sms * 8 * 256 * 4));
        const int iters = 4096;
        k_popc_rate<<<sms * 8, 256>>>(o, 16);
        CK(cudaEventRecord(e0));
        k_popc_rate<<<sms * 8, 256>>>(o, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double popc = (double)sms * 8 * 256 * iters * 32.0;
        printf("popc32 issue rate: %.2f T popc/s  (= %.1f lanes/clk/SM at 1.965 GHz); needs %.2f T/s for 2 popcll per "
               "streamed word at %.0f GB/s\n", popc / ms / 1e9, popc / ms / 1e9 * 1e12 / sms / 1.965e9 / 1e0 * 1e-0 / 1e0,
               peak / 8 * 4 / 1e3, peak);
    }
    return 0;
}
