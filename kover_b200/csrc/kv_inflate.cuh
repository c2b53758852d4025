// kv_inflate.cuh -- zlib / DEFLATE (RFC 1950 / 1951) decoding on the device, for the HDF5 -> HBM upload.
//
// A Kover dataset stores `kmer_matrix` as independent gzip-filtered chunks of (1, <=100 000) packed words
// (reference dataset/create.py:222-230, tools/kmer_tools/src/kl_64.cpp:229-263); the reference inflates them on
// the host inside every `dataset[rows, c0:c1]` read (learning/common/rules.py:253, h5py -> libhdf5 -> zlib).  Here
// the RAW streams go to the GPU and ONE WARP decodes ONE chunk:
//   * lane 0 walks the bit stream (64-bit bit buffer refilled with aligned 32-bit words) and decodes Huffman
//     symbols through per-warp tables in shared memory -- 10-bit primary table for literal/length codes, 8-bit for
//     distance codes, longer (rare) codes through the canonical bit-by-bit walk;
//   * literals are staged (up to 128) and written by the whole warp; an LZ77 match ends the batch and is copied by
//     the whole warp from the already written output (the 32 KB window is simply the output itself);
//   * stored, fixed-Huffman and dynamic-Huffman blocks, zlib header and the Adler-32 trailer are all checked;
//     any violation marks the chunk (status != 0) and the caller inflates that chunk on the host instead.
// Output goes to a scratch slab (one chunk after the other); k_scatter_chunks then places the rows of every chunk
// into the [W, ld] image (edge chunks, shard boundaries inside a chunk and 32-bit packing are handled there).
#pragma once

enum {
    INF_OK = 0,
    INF_BAD_HEADER = 1,
    INF_BAD_BLOCK = 2,
    INF_BAD_LENGTHS = 3,
    INF_OUT_OVERFLOW = 4,
    INF_IN_OVERRUN = 5,
    INF_BAD_DISTANCE = 6,
    INF_BAD_CHECKSUM = 7,
    INF_OUT_SHORT = 8,
    INF_BAD_CODE = 9,
};

struct InflateJob {
    unsigned long long src_off;     // byte offset of the zlib stream in the compressed slab (16-byte aligned)
    unsigned int src_bytes;
    unsigned int dst_bytes;         // expected size after inflation
    unsigned char *dst;             // where the chunk inflates to (16-byte aligned): its row of the resident image when
};                                  // the chunk lies wholly inside the shard, else a slot of the scratch slab

static constexpr int INF_WARPS = 8;            // chunks decoded concurrently per CTA
static constexpr int INF_LIT_BITS = 10, INF_DIST_BITS = 8;
static constexpr int INF_STAGE = 128;          // literals staged per batch

struct InflateWarp {                            // per-warp shared memory (~3.8 KB)
    unsigned short lit_tab[1 << INF_LIT_BITS];  // (symbol << 4) | code length, bit 15 = literal; 0 = longer code / unused
    unsigned short dist_tab[1 << INF_DIST_BITS];
    unsigned short lit_sym[288], dist_sym[32];  // symbols in canonical order
    unsigned short lit_cnt[16], dist_cnt[16];   // codes per length
    unsigned char lens[320];                    // code lengths of the block being set up
    unsigned char stage[INF_STAGE];
    unsigned int run[256];                      // k_inflate2: per candidate bit position (next position << 8) | literal
};

__constant__ unsigned short c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                              35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ unsigned char c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ unsigned short c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                               1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ unsigned char c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ unsigned char c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitReader {                  // lane 0 only; lives in registers (never pass it to a non-inlined function)
    const unsigned int *words;      // stream start, 4-byte aligned
    unsigned int n_words;           // words that may be read (stream bytes rounded up)
    unsigned int next;              // index of the word AFTER `ahead`
    unsigned int ahead;             // the next word, loaded one refill early (hides the load latency)
    unsigned long long buf;
    int cnt;                        // valid bits in buf
};

__device__ __forceinline__ unsigned int br_load(const BitReader &b, unsigned int i) {
    return i < b.n_words ? __ldg(b.words + i) : 0u;
}
__device__ __forceinline__ void br_start(BitReader &b, unsigned int word) {      // position the reader at `word`
    b.buf = 0;
    b.cnt = 0;
    b.ahead = br_load(b, word);
    b.next = word + 1;
}
__device__ __forceinline__ void br_refill(BitReader &b) {                       // afterwards cnt >= 33
    if (b.cnt <= 32) {
        b.buf |= (unsigned long long)b.ahead << b.cnt;
        b.cnt += 32;
        b.ahead = br_load(b, b.next);
        ++b.next;
    }
}
// more words consumed than the stream holds (two words of slack: `ahead` and the final refill may run past the end)
__device__ __forceinline__ bool br_overrun(const BitReader &b) { return b.next > b.n_words + 2; }
__device__ __forceinline__ unsigned int br_bits(BitReader &b, int n) {      // n <= 32, after a refill
    const unsigned int v = (unsigned int)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n;
    b.cnt -= n;
    return v;
}
// stream position (bytes) of the next unread byte, when the reader is byte aligned
__device__ __forceinline__ unsigned int br_byte_pos(const BitReader &b) { return (b.next - 1) * 4 - (unsigned int)(b.cnt >> 3); }

// Canonical bit-by-bit decode of the code that starts at bit 0 of `bits` (codes longer than the primary table, or
// tables with unused entries).  Pure function, deliberately NOT inlined: it is rare, and inlining it (unrolled) into
// the symbol loop bloats the loop beyond the instruction cache.  -> symbol | length << 16, or -1.
__device__ __noinline__ int slow_decode(unsigned int bits, const unsigned short *cnt, const unsigned short *sym) {
    int code = 0, first = 0, index = 0;
#pragma unroll 1
    for (int len = 1; len <= 15; ++len) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = cnt[len];
        if (code - c < first) return (int)sym[index + (code - first)] | (len << 16);
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}
__device__ __forceinline__ int br_slow_symbol(BitReader &b, const unsigned short *cnt, const unsigned short *sym) {
    const int r = slow_decode((unsigned int)b.buf, cnt, sym);
    if (r < 0) return -1;
    b.buf >>= (r >> 16);
    b.cnt -= (r >> 16);
    return r & 0xFFFF;
}

// Canonical Huffman set-up from code lengths (whole warp).  lens[n] -> cnt[16], sym[] (canonical order), primary table
// tab[1 << bits].  Returns false for an over-subscribed set of lengths.
__device__ __forceinline__ bool build_tables(const unsigned char *lens, int n, unsigned short *cnt, unsigned short *sym,
                             unsigned short *tab, int bits, int lane, bool flag_literals) {
    __shared__ unsigned short s_first_code[INF_WARPS][16], s_first_idx[INF_WARPS][16];
    const int w = threadIdx.x >> 5;
    for (int i = lane; i < (1 << bits); i += 32) tab[i] = 0;
    bool ok = true;
    if (lane == 0) {
#pragma unroll 1
        for (int l = 0; l < 16; ++l) cnt[l] = 0;
#pragma unroll 1
        for (int s = 0; s < n; ++s) cnt[lens[s]]++;
        cnt[0] = 0;
        int left = 1;
#pragma unroll 1
        for (int l = 1; l <= 15; ++l) {
            left <<= 1;
            left -= cnt[l];
            if (left < 0) ok = false;
        }
        unsigned short offs[16];
        int code = 0, idx = 0;
#pragma unroll 1
        for (int l = 1; l <= 15; ++l) {
            code = (code + cnt[l - 1]) << 1;
            s_first_code[w][l] = (unsigned short)code;
            s_first_idx[w][l] = (unsigned short)idx;
            offs[l] = (unsigned short)idx;
            idx += cnt[l];
        }
#pragma unroll 1
        for (int s = 0; s < n; ++s)
            if (lens[s]) sym[offs[lens[s]]++] = (unsigned short)s;
    }
    ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
    __syncwarp();
    if (!ok) return false;
    int total = 0;
#pragma unroll 1
    for (int l = 1; l <= 15; ++l) total += cnt[l];
    for (int i = lane; i < total; i += 32) {
        const int s = sym[i];
        const int l = lens[s];
        if (l > bits) continue;
        const unsigned int code = (unsigned int)s_first_code[w][l] + (unsigned int)(i - s_first_idx[w][l]);
        const unsigned int rev = __brev(code) >> (32 - l);
        // bit 15 marks a literal (symbol < 256) in the literal/length table: the symbol loop's fast path
        const unsigned short e = (unsigned short)((s << 4) | l | ((flag_literals && s < 256) ? 0x8000 : 0));
        for (unsigned int j = rev; j < (1u << bits); j += (1u << l)) tab[j] = e;
    }
    __syncwarp();
    return true;
}

// dynamic block header (RFC 1951 3.2.7), lane 0: fills lens[0 .. hlit + hdist)
__device__ __forceinline__ int read_dynamic_lengths(BitReader &b, unsigned char *lens, int &hlit, int &hdist) {
    br_refill(b);
    hlit = (int)br_bits(b, 5) + 257;
    hdist = (int)br_bits(b, 5) + 1;
    const int hclen = (int)br_bits(b, 4) + 4;
    if (hlit > 286 || hdist > 30) return INF_BAD_LENGTHS;
    unsigned char cl[19];
#pragma unroll 1
    for (int i = 0; i < 19; ++i) cl[i] = 0;
#pragma unroll 1
    for (int i = 0; i < hclen; ++i) {
        br_refill(b);
        cl[c_clen_order[i]] = (unsigned char)br_bits(b, 3);
    }
    // the 19-symbol code is decoded canonically, bit by bit (at most 7 bits, a few hundred symbols per block)
    unsigned short ccnt[16], csym[19];
#pragma unroll 1
    for (int l = 0; l < 16; ++l) ccnt[l] = 0;
#pragma unroll 1
    for (int s = 0; s < 19; ++s) ccnt[cl[s]]++;
    ccnt[0] = 0;
    {
        int left = 1;
#pragma unroll 1
        for (int l = 1; l <= 7; ++l) {
            left <<= 1;
            left -= ccnt[l];
            if (left < 0) return INF_BAD_LENGTHS;
        }
        unsigned short offs[16];
        int idx = 0;
#pragma unroll 1
        for (int l = 1; l <= 15; ++l) {
            offs[l] = (unsigned short)idx;
            idx += ccnt[l];
        }
#pragma unroll 1
        for (int s = 0; s < 19; ++s)
            if (cl[s]) csym[offs[cl[s]]++] = (unsigned short)s;
    }
    int n = 0;
    const int total = hlit + hdist;
    while (n < total) {
        br_refill(b);
        const int s = br_slow_symbol(b, ccnt, csym);
        if (s < 0) return INF_BAD_CODE;
        if (s < 16) {
            lens[n++] = (unsigned char)s;
            continue;
        }
        int rep, val = 0;
        if (s == 16) {
            if (n == 0) return INF_BAD_LENGTHS;
            val = lens[n - 1];
            rep = 3 + (int)br_bits(b, 2);
        } else if (s == 17) {
            rep = 3 + (int)br_bits(b, 3);
        } else {
            rep = 11 + (int)br_bits(b, 7);
        }
        if (n + rep > total) return INF_BAD_LENGTHS;
        for (int i = 0; i < rep; ++i) lens[n++] = (unsigned char)val;
    }
    if (lens[256] == 0) return INF_BAD_LENGTHS;          // no end-of-block code
    return br_overrun(b) ? INF_IN_OVERRUN : INF_OK;
}

// Adler-32 of dst[0 .. n) by the whole warp (RFC 1950): s1 = 1 + sum b_i, s2 = n + sum (n - i) b_i, mod 65521
__device__ __forceinline__ unsigned int warp_adler32(const unsigned char *dst, unsigned int n, int lane) {
    unsigned long long s1 = 0, s2 = 0;
    const unsigned int n16 = n / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(dst);
#pragma unroll 4
    for (unsigned int i = lane; i < n16; i += 32) {        // (unrolled: four independent 16-byte loads in flight)
        const uint4 q = __ldcg(v + i);
        const unsigned int w[4] = {q.x, q.y, q.z, q.w};
        unsigned int a = 0, wsum = 0;                      // a = sum b_j, wsum = sum j * b_j over the 16 bytes
#pragma unroll
        for (int k = 0; k < 4; ++k) {                      // byte-wise dot products: weights 1,1,1,1 and 4k .. 4k+3
            a = __dp4a(w[k], 0x01010101u, a);
            wsum = __dp4a(w[k], 0x03020100u + 0x04040404u * (unsigned int)k, wsum);
        }
        const unsigned long long base = (unsigned long long)n - (unsigned long long)i * 16ull;   // weight of byte 0
        s1 += a;
        s2 += base * a - wsum;
    }
    for (unsigned int i = n16 * 16 + lane; i < n; i += 32) {
        const unsigned int byte = dst[i];
        s1 += byte;
        s2 += (unsigned long long)(n - i) * byte;
    }
    s1 %= 65521ull;
    s2 %= 65521ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 = (s1 + 1ull) % 65521ull;
    s2 = (s2 + (unsigned long long)n) % 65521ull;
    return (unsigned int)((s2 << 16) | s1);
}

__global__ void __launch_bounds__(INF_WARPS * 32) k_inflate(const unsigned char *src_base, const InflateJob *jobs,
                                                             int n_jobs, int *status, unsigned int *next_job) {
    __shared__ InflateWarp s_w[INF_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    InflateWarp &S = s_w[wid];
    for (;;) {
        // dynamic work distribution: chunks differ in cost (sparse rows inflate faster)
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(next_job, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= n_jobs) return;
        const InflateJob J = jobs[job];
        unsigned char *dst = J.dst;
        BitReader b;
        b.words = reinterpret_cast<const unsigned int *>(src_base + J.src_off);
        b.n_words = (J.src_bytes + 3) / 4;
        br_start(b, 0);
        int st = INF_OK;
        unsigned int out = 0;
        if (lane == 0) {
            br_refill(b);
            const unsigned int cmf = br_bits(b, 8), flg = br_bits(b, 8);
            if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) st = INF_BAD_HEADER;
        }
        st = __shfl_sync(0xffffffffu, st, 0);
        bool last = false;
        while (st == INF_OK && !last) {
            int btype = 0, hlit = 0, hdist = 0;
            if (lane == 0) {
                br_refill(b);
                last = br_bits(b, 1) != 0;
                btype = (int)br_bits(b, 2);
            }
            last = __shfl_sync(0xffffffffu, (int)last, 0) != 0;
            btype = __shfl_sync(0xffffffffu, btype, 0);
            if (btype == 0) {
                // stored block: skip to the byte boundary, LEN / NLEN, raw bytes
                unsigned int len = 0, byte_pos = 0;
                if (lane == 0) {
                    br_bits(b, b.cnt & 7);
                    br_refill(b);
                    len = br_bits(b, 16);
                    br_refill(b);
                    const unsigned int nlen = br_bits(b, 16);
                    if ((len ^ nlen) != 0xFFFFu) st = INF_BAD_BLOCK;
                    byte_pos = br_byte_pos(b);
                    if (byte_pos + len > J.src_bytes) st = INF_IN_OVERRUN;
                    if (out + len > J.dst_bytes) st = INF_OUT_OVERFLOW;
                }
                st = __shfl_sync(0xffffffffu, st, 0);
                len = __shfl_sync(0xffffffffu, len, 0);
                byte_pos = __shfl_sync(0xffffffffu, byte_pos, 0);
                if (st != INF_OK) break;
                const unsigned char *sb = src_base + J.src_off + byte_pos;
                for (unsigned int i = lane; i < len; i += 32) dst[out + i] = sb[i];
                out += len;
                if (lane == 0) {                                              // restart the bit reader after the raw bytes
                    const unsigned int p = byte_pos + len;
                    br_start(b, p / 4);
                    br_refill(b);
                    br_bits(b, (int)(p & 3u) * 8);
                }
                __syncwarp();
                continue;
            }
            if (btype == 3) {
                st = INF_BAD_BLOCK;
                break;
            }
            if (btype == 1) {
                for (int i = lane; i < 288; i += 32) S.lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                for (int i = lane; i < 32; i += 32) S.lens[288 + i] = i < 30 ? 5 : 0;
                hlit = 288;
                hdist = 30;
            } else {
                if (lane == 0) st = read_dynamic_lengths(b, S.lens, hlit, hdist);
                st = __shfl_sync(0xffffffffu, st, 0);
                hlit = __shfl_sync(0xffffffffu, hlit, 0);
                hdist = __shfl_sync(0xffffffffu, hdist, 0);
                if (st != INF_OK) break;
            }
            __syncwarp();
            if (!build_tables(S.lens, hlit, S.lit_cnt, S.lit_sym, S.lit_tab, INF_LIT_BITS, lane, true) ||
                !build_tables(S.lens + hlit, hdist, S.dist_cnt, S.dist_sym, S.dist_tab, INF_DIST_BITS, lane, false)) {
                st = INF_BAD_LENGTHS;
                break;
            }
            // ---- symbols of the block, in batches: up to INF_STAGE literals, then at most one match
            bool end_of_block = false;
            while (!end_of_block && st == INF_OK) {
                int nlit = 0, mlen = 0, mdist = 0;
                if (lane == 0) {
#pragma unroll 1
                    for (;;) {
                        // tight loop over short literal codes (entry bit 15): refill, look up, consume, stage
                        unsigned int e;
#pragma unroll 1
                        for (;;) {
                            br_refill(b);
                            e = S.lit_tab[(unsigned int)b.buf & ((1u << INF_LIT_BITS) - 1u)];
                            if (!(e & 0x8000u)) break;
                            b.buf >>= (e & 15u);
                            b.cnt -= (int)(e & 15u);
                            S.stage[nlit] = (unsigned char)(e >> 4);
                            if (++nlit == INF_STAGE) break;
                        }
                        if (nlit == INF_STAGE) break;
                        int sym;
                        if (e & 15u) {
                            b.buf >>= (e & 15u);
                            b.cnt -= (int)(e & 15u);
                            sym = (int)(e >> 4);
                        } else {
                            sym = br_slow_symbol(b, S.lit_cnt, S.lit_sym);
                            if (sym < 0) {
                                st = INF_BAD_CODE;
                                break;
                            }
                        }
                        if (sym < 256) {                          // a literal with a code longer than the table
                            S.stage[nlit] = (unsigned char)sym;
                            if (++nlit == INF_STAGE) break;
                            continue;
                        }
                        if (sym == 256) {
                            end_of_block = true;
                            break;
                        }
                        sym -= 257;
                        if (sym >= 29) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        mlen = c_len_base[sym] + (int)br_bits(b, c_len_extra[sym]);
                        br_refill(b);
                        e = S.dist_tab[(unsigned int)b.buf & ((1u << INF_DIST_BITS) - 1u)];
                        int ds;
                        if (e & 15u) {
                            b.buf >>= (e & 15u);
                            b.cnt -= (int)(e & 15u);
                            ds = (int)(e >> 4);
                        } else {
                            ds = br_slow_symbol(b, S.dist_cnt, S.dist_sym);
                        }
                        if (ds < 0 || ds >= 30) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        mdist = c_dist_base[ds] + (int)br_bits(b, c_dist_extra[ds]);
                        break;
                    }
                    if (br_overrun(b) && st == INF_OK) st = INF_IN_OVERRUN;
                    if (st == INF_OK && (unsigned long long)out + nlit + mlen > J.dst_bytes) st = INF_OUT_OVERFLOW;
                    if (st == INF_OK && mlen && (unsigned int)mdist > out + nlit) st = INF_BAD_DISTANCE;
                }
                st = __shfl_sync(0xffffffffu, st, 0);
                if (st != INF_OK) break;
                nlit = __shfl_sync(0xffffffffu, nlit, 0);
                mlen = __shfl_sync(0xffffffffu, mlen, 0);
                mdist = __shfl_sync(0xffffffffu, mdist, 0);
                end_of_block = __shfl_sync(0xffffffffu, (int)end_of_block, 0) != 0;
                __syncwarp();                                   // lane 0's staged literals are visible
                for (int i = lane; i < nlit; i += 32) dst[out + i] = S.stage[i];
                out += (unsigned int)nlit;
                if (mlen) {
                    __syncwarp();                               // the literals just written may be the match source
                    const unsigned char *from = dst + out - mdist;
                    for (int i = lane; i < mlen; i += 32) dst[out + i] = from[i % mdist];
                    out += (unsigned int)mlen;
                }
                __syncwarp();
            }
        }
        // ---- trailer: Adler-32 of the output, stored big-endian after the last block (byte aligned)
        unsigned int want = 0;
        if (st == INF_OK) {
            if (lane == 0) {
                br_bits(b, b.cnt & 7);
                br_refill(b);
                const unsigned int a0 = br_bits(b, 16);
                br_refill(b);
                const unsigned int a1 = br_bits(b, 16);
                const unsigned int le = a0 | (a1 << 16);
                want = __byte_perm(le, 0, 0x0123);
                if (br_overrun(b)) st = INF_IN_OVERRUN;
                if (st == INF_OK && out != J.dst_bytes) st = INF_OUT_SHORT;
            }
            st = __shfl_sync(0xffffffffu, st, 0);
            want = __shfl_sync(0xffffffffu, want, 0);
        }
        if (st == INF_OK) {
            __syncwarp();
            __threadfence_block();
            if (warp_adler32(dst, J.dst_bytes, lane) != want) st = INF_BAD_CHECKSUM;
        }
        if (lane == 0) status[job] = st;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_inflate2: the same decoder with the LITERAL RUNS decoded by the whole warp.
//
// In k_inflate one lane walks the Huffman stream: ~200 cycles per symbol, whatever else the GPU does (a dependent
// chain of a refill, a shared-memory lookup and half a dozen branches at ~16 cycles per dependent instruction).  Here
// the decoder state is just the bit position p, identical in all lanes (no lane-0 section, no broadcasts), and a
// literal run is decoded speculatively: the lanes look up the codes that WOULD start at each of the next 256 bit
// positions (eight per lane) and write {next position, literal} per position to shared memory; the codes that really
// are on the chain p -> p + len -> ... are then found by walking that table, ONE dependent shared-memory load per
// symbol, and up to 32 literals are written by the warp at once.  Everything that is not a table-resident literal
// (length / distance pairs, end of block, codes longer than the table) is decoded one symbol at a time, uniformly by
// all lanes, from a 64-bit peek at p.  Block headers are parsed the same way.
// ------------------------------------------------------------------------------------------------
struct BitSrc {
    const unsigned int *w;          // stream start, 4-byte aligned
    unsigned int n_words;
};
__device__ __forceinline__ unsigned int src_word(const BitSrc &s, unsigned int i) { return i < s.n_words ? __ldg(s.w + i) : 0u; }
__device__ __forceinline__ unsigned int peek32(const BitSrc &s, unsigned int pos) {          // 32 bits from bit `pos`
    const unsigned int i = pos >> 5;
    return __funnelshift_r(src_word(s, i), src_word(s, i + 1), pos & 31u);
}
__device__ __forceinline__ unsigned long long peek64(const BitSrc &s, unsigned int pos) {    // 64 bits from bit `pos`
    const unsigned int i = pos >> 5, sh = pos & 31u;
    const unsigned int w0 = src_word(s, i), w1 = src_word(s, i + 1), w2 = src_word(s, i + 2);
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}

// dynamic block header at bit p (RFC 1951 3.2.7), executed identically by every lane; lane 0 stores the lengths
__device__ __forceinline__ int read_dynamic_lengths2(const BitSrc &src, unsigned int &p, unsigned char *lens, int &hlit,
                                                     int &hdist, int lane) {
    unsigned int w = peek32(src, p);
    hlit = (int)(w & 31u) + 257;
    hdist = (int)((w >> 5) & 31u) + 1;
    const int hclen = (int)((w >> 10) & 15u) + 4;
    p += 14;
    if (hlit > 286 || hdist > 30) return INF_BAD_LENGTHS;
    unsigned char cl[19];
#pragma unroll 1
    for (int i = 0; i < 19; ++i) cl[i] = 0;
    const unsigned long long q = peek64(src, p);                  // hclen * 3 <= 57 bits
#pragma unroll 1
    for (int i = 0; i < hclen; ++i) cl[c_clen_order[i]] = (unsigned char)((q >> (3 * i)) & 7ull);
    p += 3u * (unsigned int)hclen;
    unsigned short ccnt[16], csym[19];
#pragma unroll 1
    for (int l = 0; l < 16; ++l) ccnt[l] = 0;
#pragma unroll 1
    for (int s = 0; s < 19; ++s) ccnt[cl[s]]++;
    ccnt[0] = 0;
    {
        int left = 1;
#pragma unroll 1
        for (int l = 1; l <= 7; ++l) {
            left <<= 1;
            left -= ccnt[l];
            if (left < 0) return INF_BAD_LENGTHS;
        }
        unsigned short offs[16];
        int idx = 0;
#pragma unroll 1
        for (int l = 1; l <= 15; ++l) {
            offs[l] = (unsigned short)idx;
            idx += ccnt[l];
        }
#pragma unroll 1
        for (int s = 0; s < 19; ++s)
            if (cl[s]) csym[offs[cl[s]]++] = (unsigned short)s;
    }
    int n = 0, prev = 0;
    const int total = hlit + hdist;
#pragma unroll 1
    while (n < total) {
        w = peek32(src, p);
        const int r = slow_decode(w, ccnt, csym);
        if (r < 0) return INF_BAD_CODE;
        const int s = r & 0xFFFF, used = r >> 16;
        w >>= used;
        p += (unsigned int)used;
        if (s < 16) {
            if (lane == 0) lens[n] = (unsigned char)s;
            ++n;
            prev = s;
            continue;
        }
        int rep, val = 0;
        if (s == 16) {
            if (n == 0) return INF_BAD_LENGTHS;
            val = prev;
            rep = 3 + (int)(w & 3u);
            p += 2;
        } else if (s == 17) {
            rep = 3 + (int)(w & 7u);
            p += 3;
        } else {
            rep = 11 + (int)(w & 127u);
            p += 7;
        }
        if (n + rep > total) return INF_BAD_LENGTHS;
        if (lane == 0)
            for (int i = 0; i < rep; ++i) lens[n + i] = (unsigned char)val;
        n += rep;
        prev = val;
    }
    __syncwarp();
    if (lens[256] == 0) return INF_BAD_LENGTHS;          // no end-of-block code
    return INF_OK;
}

__global__ void __launch_bounds__(INF_WARPS * 32) k_inflate2(const unsigned char *src_base, const InflateJob *jobs,
                                                              int n_jobs, int *status, unsigned int *next_job) {
    __shared__ InflateWarp s_w[INF_WARPS];
    // length / distance code tables next to the decoder (constant-memory lookups with a data-dependent index cost a
    // constant-cache round trip each, four per match)
    __shared__ unsigned short s_len_base[29], s_dist_base[30];
    __shared__ unsigned char s_len_extra[29], s_dist_extra[30];
    if (threadIdx.x < 29) {
        s_len_base[threadIdx.x] = c_len_base[threadIdx.x];
        s_len_extra[threadIdx.x] = c_len_extra[threadIdx.x];
    }
    if (threadIdx.x < 30) {
        s_dist_base[threadIdx.x] = c_dist_base[threadIdx.x];
        s_dist_extra[threadIdx.x] = c_dist_extra[threadIdx.x];
    }
    __syncthreads();
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    InflateWarp &S = s_w[wid];
    for (;;) {
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(next_job, 1u);
        job = __shfl_sync(full, job, 0);
        if (job >= n_jobs) return;
        const InflateJob J = jobs[job];
        unsigned char *dst = J.dst;
        BitSrc src;
        src.w = reinterpret_cast<const unsigned int *>(src_base + J.src_off);
        src.n_words = (J.src_bytes + 3) / 4;
        const unsigned int nbits = J.src_bytes * 8u;
        int st = INF_OK;
        unsigned int out = 0, p = 16;
        {
            const unsigned int w = peek32(src, 0);
            const unsigned int cmf = w & 255u, flg = (w >> 8) & 255u;
            if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) st = INF_BAD_HEADER;
            if (J.src_bytes >= (1u << 28)) st = INF_IN_OVERRUN;            // bit positions are 32-bit
        }
        bool last = false;
        while (st == INF_OK && !last) {
            const unsigned int hw = peek32(src, p);
            last = (hw & 1u) != 0;
            const int btype = (int)((hw >> 1) & 3u);
            p += 3;
            if (btype == 0) {
                p = (p + 7u) & ~7u;
                const unsigned int lw = peek32(src, p);
                const unsigned int len = lw & 0xFFFFu, nlen = lw >> 16;
                p += 32;
                const unsigned int byte_pos = p >> 3;
                if ((len ^ nlen) != 0xFFFFu) st = INF_BAD_BLOCK;
                else if (byte_pos + len > J.src_bytes) st = INF_IN_OVERRUN;
                else if (out + len > J.dst_bytes) st = INF_OUT_OVERFLOW;
                if (st != INF_OK) break;
                const unsigned char *sb = src_base + J.src_off + byte_pos;
                for (unsigned int i = lane; i < len; i += 32) dst[out + i] = sb[i];
                out += len;
                p += len * 8u;
                __syncwarp();
                continue;
            }
            if (btype == 3) {
                st = INF_BAD_BLOCK;
                break;
            }
            int hlit = 288, hdist = 30;
            if (btype == 1) {
                for (int i = lane; i < 288; i += 32) S.lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                for (int i = lane; i < 32; i += 32) S.lens[288 + i] = i < 30 ? 5 : 0;
            } else {
                st = read_dynamic_lengths2(src, p, S.lens, hlit, hdist, lane);
                if (st != INF_OK) break;
            }
            __syncwarp();
            if (!build_tables(S.lens, hlit, S.lit_cnt, S.lit_sym, S.lit_tab, INF_LIT_BITS, lane, true) ||
                !build_tables(S.lens + hlit, hdist, S.dist_cnt, S.dist_sym, S.dist_tab, INF_DIST_BITS, lane, false)) {
                st = INF_BAD_LENGTHS;
                break;
            }
            // ---- symbols of the block
            bool wide = true;
            for (;;) {
                // (1) the literal run that starts at p.  Every lane decodes the codes that WOULD start at 8 of the next
                // 256 bit positions (one 32-bit peek, eight table lookups) and leaves {next position, literal} per
                // position in shared memory; the chain p -> p + len -> ... is then walked through that table -- one
                // dependent shared-memory load per symbol instead of a refill + lookup + branch chain -- by all
                // lanes alike; lane i keeps the i-th literal of the run and the warp writes up to 32 bytes at once.
                // (after a short run -- match-heavy data -- only 64 positions, two per lane, are looked up)
                const unsigned int window = wide ? 256u : 64u;
                if (wide) {
                    const unsigned int w = peek32(src, p + 8u * (unsigned int)lane);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const unsigned int e = S.lit_tab[(w >> j) & ((1u << INF_LIT_BITS) - 1u)];
                        const unsigned int c = 8u * (unsigned int)lane + (unsigned int)j;
                        S.run[c] = (e & 0x8000u) ? (((c + (e & 15u)) << 8) | ((e >> 4) & 0xFFu)) : 0xFFFFFFFFu;
                    }
                } else {
                    const unsigned int w = peek32(src, p + 2u * (unsigned int)lane);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const unsigned int e = S.lit_tab[(w >> j) & ((1u << INF_LIT_BITS) - 1u)];
                        const unsigned int c = 2u * (unsigned int)lane + (unsigned int)j;
                        S.run[c] = (e & 0x8000u) ? (((c + (e & 15u)) << 8) | ((e >> 4) & 0xFFu)) : 0xFFFFFFFFu;
                    }
                }
                __syncwarp();
                unsigned int pos = 0, n = 0, mine = 0;
                bool stop = false;
#pragma unroll 1
                for (;;) {
                    const unsigned int e = S.run[pos];
                    if (e == 0xFFFFFFFFu) {                                // not a table-resident literal
                        stop = true;
                        break;
                    }
                    if (n == (unsigned int)lane) mine = e & 0xFFu;
                    ++n;
                    pos = e >> 8;
                    if (pos >= window || n == 32u) break;
                }
                wide = n >= 8u;
                __syncwarp();                                              // (S.run is rewritten by the next iteration)
                if (out + n > J.dst_bytes) {
                    st = INF_OUT_OVERFLOW;
                    break;
                }
                if ((unsigned int)lane < n) dst[out + lane] = (unsigned char)mine;
                out += n;
                p += pos;
                if (p > nbits) {
                    st = INF_IN_OVERRUN;
                    break;
                }
                if (!stop) continue;                                       // the run goes on: next window
                // (2) the symbol at p is not a table-resident literal: one symbol, decoded identically by every lane
                unsigned long long q = peek64(src, p);
                unsigned int used;
                int sym;
                {
                    const unsigned int e1 = S.lit_tab[(unsigned int)q & ((1u << INF_LIT_BITS) - 1u)];
                    if (e1 & 15u) {
                        used = e1 & 15u;
                        sym = (int)((e1 >> 4) & 0x1FFu);
                    } else {
                        const int r = slow_decode((unsigned int)q, S.lit_cnt, S.lit_sym);
                        if (r < 0) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        used = (unsigned int)(r >> 16);
                        sym = r & 0xFFFF;
                    }
                }
                q >>= used;
                if (sym < 256) {                                           // a literal with a code longer than the table
                    if (out + 1 > J.dst_bytes) {
                        st = INF_OUT_OVERFLOW;
                        break;
                    }
                    if (lane == 0) dst[out] = (unsigned char)sym;
                    out += 1;
                    p += used;
                    continue;
                }
                if (sym == 256) {
                    p += used;
                    break;                                                 // end of block
                }
                sym -= 257;
                if (sym >= 29) {
                    st = INF_BAD_CODE;
                    break;
                }
                unsigned int eb = s_len_extra[sym];
                const unsigned int mlen = s_len_base[sym] + ((unsigned int)q & ((1u << eb) - 1u));
                q >>= eb;
                used += eb;
                int ds;
                {
                    const unsigned int e2 = S.dist_tab[(unsigned int)q & ((1u << INF_DIST_BITS) - 1u)];
                    unsigned int dl;
                    if (e2 & 15u) {
                        dl = e2 & 15u;
                        ds = (int)(e2 >> 4);
                    } else {
                        const int r = slow_decode((unsigned int)q, S.dist_cnt, S.dist_sym);
                        if (r < 0) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        dl = (unsigned int)(r >> 16);
                        ds = r & 0xFFFF;
                    }
                    q >>= dl;
                    used += dl;
                }
                if (ds >= 30) {
                    st = INF_BAD_CODE;
                    break;
                }
                eb = s_dist_extra[ds];
                const unsigned int mdist = s_dist_base[ds] + ((unsigned int)q & ((1u << eb) - 1u));
                used += eb;
                p += used;
                if (p > nbits) {
                    st = INF_IN_OVERRUN;
                    break;
                }
                if (mdist > out) {
                    st = INF_BAD_DISTANCE;
                    break;
                }
                if (out + mlen > J.dst_bytes) {
                    st = INF_OUT_OVERFLOW;
                    break;
                }
                __syncwarp();                                              // the match source may have just been written
                const unsigned char *from = dst + out - mdist;
                for (unsigned int i = lane; i < mlen; i += 32) dst[out + i] = from[i % mdist];
                out += mlen;
                __syncwarp();
            }
            if (p > nbits && st == INF_OK) st = INF_IN_OVERRUN;
        }
        // ---- trailer: Adler-32 of the output, big-endian, byte aligned
        if (st == INF_OK) {
            p = (p + 7u) & ~7u;
            const unsigned int want = __byte_perm(peek32(src, p), 0, 0x0123);
            p += 32;
            if (p > nbits) st = INF_IN_OVERRUN;
            else if (out != J.dst_bytes) st = INF_OUT_SHORT;
            if (st == INF_OK) {
                __syncwarp();
                __threadfence_block();
                if (warp_adler32(dst, J.dst_bytes, lane) != want) st = INF_BAD_CHECKSUM;
            }
        }
        if (lane == 0) status[job] = st;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_inflate3: k_inflate2 with a shorter critical path per literal run (a chunk's latency, not the GPU's throughput, bounds
// uploads of a few GB: DESIGN.md section 4.7).
//   * the next 128-byte line of the compressed stream is prefetched into L1 when the decoder enters a line (a register
//     window read with shuffles was tried: as fast alone, slower with 18 chunks per SM -- shuffles share the LSU pipe);
//   * the chain p -> p + len -> ... is walked TWO codes per dependent shared-memory load: after the per-position
//     entries E1[c] = {where the code starting at c leads, its literal} every lane composes E2[c] = {E1[c], E1[next(c)]}
//     for its positions (one more parallel round), and a chain step is then LDS.64 -> LOP -> LDS.64: the low 18 bits
//     of the loaded word ARE the shared-memory address of the next load.  Entries that are not table-resident literals
//     (STOP) or lie beyond the window (END) point at themselves, so the walk needs no test inside a group of four
//     steps.  Short runs (match-heavy data) keep the one-code walk over a 64-position window.
// Same LSU instruction count per run as k_inflate2 (the decoder is also bound by shared-memory instructions once
// every SM holds dozens of chunks), half the dependent loads.
// ------------------------------------------------------------------------------------------------
static constexpr unsigned int INF_E_STOP = 0x80000000u, INF_E_END = 0x40000000u;

struct InflateWarp3 {                           // per-warp shared memory (6 848 B), dynamic
    unsigned short lit_tab[1 << INF_LIT_BITS];
    unsigned short dist_tab[1 << INF_DIST_BITS];
    unsigned short lit_sym[288], dist_sym[32];
    unsigned short lit_cnt[16], dist_cnt[16];
    unsigned char lens[320];
    unsigned int e1[256 + 16];                  // address of E2[next] | literal << 18 | flags
    uint2 e2[256 + 16];                         // {E1[c], E1[next(c)]}
};

__global__ void __launch_bounds__(INF_WARPS * 32) k_inflate3(const unsigned char *src_base, const InflateJob *jobs,
                                                              int n_jobs, int *status, unsigned int *next_job) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    InflateWarp3 *s_w = reinterpret_cast<InflateWarp3 *>(s_dyn);
    // length / distance code tables next to the decoder (constant-memory lookups with a data-dependent index cost a
    // constant-cache round trip each, four per match)
    __shared__ unsigned short s_len_base[29], s_dist_base[30];
    __shared__ unsigned char s_len_extra[29], s_dist_extra[30];
    if (threadIdx.x < 29) {
        s_len_base[threadIdx.x] = c_len_base[threadIdx.x];
        s_len_extra[threadIdx.x] = c_len_extra[threadIdx.x];
    }
    if (threadIdx.x < 30) {
        s_dist_base[threadIdx.x] = c_dist_base[threadIdx.x];
        s_dist_extra[threadIdx.x] = c_dist_extra[threadIdx.x];
    }
    __syncthreads();
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    InflateWarp3 &S = s_w[wid];
    const unsigned int e1_base = (unsigned int)__cvta_generic_to_shared(S.e1);
    const unsigned int e2_base = (unsigned int)__cvta_generic_to_shared(S.e2);
    for (;;) {
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(next_job, 1u);
        job = __shfl_sync(full, job, 0);
        if (job >= n_jobs) return;
        const InflateJob J = jobs[job];
        unsigned char *dst = J.dst;
        BitSrc src;
        src.w = reinterpret_cast<const unsigned int *>(src_base + J.src_off);
        src.n_words = (J.src_bytes + 3) / 4;
        const unsigned int nbits = J.src_bytes * 8u;
        int st = INF_OK;
        unsigned int out = 0, p = 16;
        unsigned int pf_line = 0xFFFFFFFFu;                          // last 128-byte line of the stream that was prefetched
        {
            const unsigned int w = peek32(src, 0);
            const unsigned int cmf = w & 255u, flg = (w >> 8) & 255u;
            if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) st = INF_BAD_HEADER;
            if (J.src_bytes >= (1u << 28)) st = INF_IN_OVERRUN;            // bit positions are 32-bit
        }
        bool last = false;
        while (st == INF_OK && !last) {
            const unsigned int hw = peek32(src, p);
            last = (hw & 1u) != 0;
            const int btype = (int)((hw >> 1) & 3u);
            p += 3;
            if (btype == 0) {
                p = (p + 7u) & ~7u;
                const unsigned int lw = peek32(src, p);
                const unsigned int len = lw & 0xFFFFu, nlen = lw >> 16;
                p += 32;
                const unsigned int byte_pos = p >> 3;
                if ((len ^ nlen) != 0xFFFFu) st = INF_BAD_BLOCK;
                else if (byte_pos + len > J.src_bytes) st = INF_IN_OVERRUN;
                else if (out + len > J.dst_bytes) st = INF_OUT_OVERFLOW;
                if (st != INF_OK) break;
                const unsigned char *sb = src_base + J.src_off + byte_pos;
                for (unsigned int i = lane; i < len; i += 32) dst[out + i] = sb[i];
                out += len;
                p += len * 8u;
                __syncwarp();
                continue;
            }
            if (btype == 3) {
                st = INF_BAD_BLOCK;
                break;
            }
            int hlit = 288, hdist = 30;
            if (btype == 1) {
                for (int i = lane; i < 288; i += 32) S.lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                for (int i = lane; i < 32; i += 32) S.lens[288 + i] = i < 30 ? 5 : 0;
            } else {
                st = read_dynamic_lengths2(src, p, S.lens, hlit, hdist, lane);
                if (st != INF_OK) break;
            }
            __syncwarp();
            if (!build_tables(S.lens, hlit, S.lit_cnt, S.lit_sym, S.lit_tab, INF_LIT_BITS, lane, true) ||
                !build_tables(S.lens + hlit, hdist, S.dist_cnt, S.dist_sym, S.dist_tab, INF_DIST_BITS, lane, false)) {
                st = INF_BAD_LENGTHS;
                break;
            }
            // ---- symbols of the block
            int mode = 2;
            for (;;) {
                // (1) the literal run that starts at p.  Every lane decodes the codes that WOULD start at 8 of the next
                // 256 bit positions (one 32-bit peek, eight table lookups) and leaves {next position, literal} per
                // position in shared memory; the chain p -> p + len -> ... is then walked through that table -- one
                // dependent shared-memory load per symbol instead of a refill + lookup + branch chain -- by all
                // lanes alike; lane i keeps the i-th literal of the run and the warp writes up to 32 bytes at once.
                // (after a short run -- match-heavy data -- only 64 positions, two per lane, are looked up)
                // mode 0: 64 positions, one code per step (after a short run: match-heavy data); mode 1: 256 positions,
                // one code per step; mode 2 (after a run of >= 24 literals): 256 positions, two codes per step
                const unsigned int window = mode ? 256u : 64u;
                if ((p >> 10) != pf_line) {                            // entering a new 128-byte line: fetch the next one into L1
                    pf_line = p >> 10;
                    const unsigned int nw = (pf_line + 1u) << 5;
                    if (nw < src.n_words) asm volatile("prefetch.global.L1 [%0];" ::"l"(src.w + nw));
                }
                // E1[c], c < window + 16: the literal of the code that WOULD start at bit p + c and the shared-memory
                // address of the entry it leads to (of E2 in mode 2, of E1 otherwise); flagged STOP when it is not a
                // table-resident literal, END beyond the window -- flagged entries point at themselves
                const unsigned int nbase = mode == 2 ? e2_base : e1_base, nshift = mode == 2 ? 3u : 2u;
                if (mode) {
                    const unsigned int w = peek32(src, p + 8u * (unsigned int)lane);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const unsigned int e = S.lit_tab[(w >> j) & ((1u << INF_LIT_BITS) - 1u)];
                        const unsigned int c = 8u * (unsigned int)lane + (unsigned int)j;
                        S.e1[c] = (e & 0x8000u) ? ((nbase + ((c + (e & 15u)) << nshift)) | (((e >> 4) & 0xFFu) << 18))
                                                : ((nbase + (c << nshift)) | INF_E_STOP);
                    }
                } else {
                    const unsigned int w = peek32(src, p + 2u * (unsigned int)lane);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const unsigned int e = S.lit_tab[(w >> j) & ((1u << INF_LIT_BITS) - 1u)];
                        const unsigned int c = 2u * (unsigned int)lane + (unsigned int)j;
                        S.e1[c] = (e & 0x8000u) ? ((nbase + ((c + (e & 15u)) << nshift)) | (((e >> 4) & 0xFFu) << 18))
                                                : ((nbase + (c << nshift)) | INF_E_STOP);
                    }
                }
                if (lane < 16) S.e1[window + (unsigned int)lane] = (nbase + ((window + (unsigned int)lane) << nshift)) | INF_E_END;
                __syncwarp();
                unsigned int pos = 0, n = 0, mine = 0;
                if (mode == 2) {
                    // E2[c] = {E1[c], E1[next(c)]}: the chain is then walked TWO codes per dependent shared-memory load,
                    // and the load address is the low 18 bits of the loaded word itself (one LOP between two loads)
                    for (unsigned int c = (unsigned int)lane; c < 256u + 16u; c += 32u) {
                        const unsigned int a1 = S.e1[c];
                        const unsigned int a2 = S.e1[((a1 & 0x3FFFFu) - e2_base) >> 3];
                        S.e2[c] = make_uint2(a1, a2);
                    }
                    __syncwarp();
                    unsigned int addr = e2_base, mlo = INF_E_STOP, mhi = INF_E_STOP;
#pragma unroll
                    for (int s2 = 0; s2 < 16; ++s2) {
                        unsigned int lo, hi;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
                        if ((lane >> 1) == s2) {
                            mlo = lo;
                            mhi = hi;
                        }
                        addr = hi & 0x3FFFFu;
                        if ((s2 & 3) == 3 && (hi >> 30)) break;            // (uniform) the run has ended
                    }
                    const unsigned int e = (lane & 1) ? mhi : mlo;
                    n = (unsigned int)__popc(__ballot_sync(full, (e >> 30) == 0u));
                    mine = (e >> 18) & 0xFFu;
                    pos = (addr - e2_base) >> 3;
                } else {
                    unsigned int addr = e1_base;
#pragma unroll 1
                    for (;;) {
                        unsigned int e;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(addr) : "memory");
                        if (e >> 30) break;
                        if (n == (unsigned int)lane) mine = (e >> 18) & 0xFFu;
                        addr = e & 0x3FFFFu;
                        if (++n == 32u) break;
                    }
                    pos = (addr - e1_base) >> 2;
                }
                const bool stop = (S.e1[pos] & INF_E_STOP) != 0u;
                mode = n < 8u ? 0 : (n < 24u ? 1 : 2);
                __syncwarp();                                              // (E1 / E2 are rewritten by the next iteration)
                if (out + n > J.dst_bytes) {
                    st = INF_OUT_OVERFLOW;
                    break;
                }
                if ((unsigned int)lane < n) dst[out + lane] = (unsigned char)mine;
                out += n;
                p += pos;
                if (p > nbits) {
                    st = INF_IN_OVERRUN;
                    break;
                }
                if (!stop) continue;                                       // the run goes on: next window
                // (2) the symbol at p is not a table-resident literal: one symbol, decoded identically by every lane
                unsigned long long q = peek64(src, p);
                unsigned int used;
                int sym;
                {
                    const unsigned int e1 = S.lit_tab[(unsigned int)q & ((1u << INF_LIT_BITS) - 1u)];
                    if (e1 & 15u) {
                        used = e1 & 15u;
                        sym = (int)((e1 >> 4) & 0x1FFu);
                    } else {
                        const int r = slow_decode((unsigned int)q, S.lit_cnt, S.lit_sym);
                        if (r < 0) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        used = (unsigned int)(r >> 16);
                        sym = r & 0xFFFF;
                    }
                }
                q >>= used;
                if (sym < 256) {                                           // a literal with a code longer than the table
                    if (out + 1 > J.dst_bytes) {
                        st = INF_OUT_OVERFLOW;
                        break;
                    }
                    if (lane == 0) dst[out] = (unsigned char)sym;
                    out += 1;
                    p += used;
                    continue;
                }
                if (sym == 256) {
                    p += used;
                    break;                                                 // end of block
                }
                sym -= 257;
                if (sym >= 29) {
                    st = INF_BAD_CODE;
                    break;
                }
                unsigned int eb = s_len_extra[sym];
                const unsigned int mlen = s_len_base[sym] + ((unsigned int)q & ((1u << eb) - 1u));
                q >>= eb;
                used += eb;
                int ds;
                {
                    const unsigned int e2 = S.dist_tab[(unsigned int)q & ((1u << INF_DIST_BITS) - 1u)];
                    unsigned int dl;
                    if (e2 & 15u) {
                        dl = e2 & 15u;
                        ds = (int)(e2 >> 4);
                    } else {
                        const int r = slow_decode((unsigned int)q, S.dist_cnt, S.dist_sym);
                        if (r < 0) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        dl = (unsigned int)(r >> 16);
                        ds = r & 0xFFFF;
                    }
                    q >>= dl;
                    used += dl;
                }
                if (ds >= 30) {
                    st = INF_BAD_CODE;
                    break;
                }
                eb = s_dist_extra[ds];
                const unsigned int mdist = s_dist_base[ds] + ((unsigned int)q & ((1u << eb) - 1u));
                used += eb;
                p += used;
                if (p > nbits) {
                    st = INF_IN_OVERRUN;
                    break;
                }
                if (mdist > out) {
                    st = INF_BAD_DISTANCE;
                    break;
                }
                if (out + mlen > J.dst_bytes) {
                    st = INF_OUT_OVERFLOW;
                    break;
                }
                __syncwarp();                                              // the match source may have just been written
                const unsigned char *from = dst + out - mdist;
                unsigned char *to = dst + out;
                if (mdist == 1u) {
                    // a run of one byte value (the zero runs of a sparse k-mer matrix): aligned 32-bit stores
                    const unsigned int b = from[0], b4 = b * 0x01010101u;
                    const unsigned int head = min(mlen, (4u - ((unsigned int)(size_t)to & 3u)) & 3u);
                    const unsigned int nw = (mlen - head) >> 2, tail0 = head + (nw << 2);
                    if ((unsigned int)lane < head) to[lane] = (unsigned char)b;
                    unsigned int *tw = reinterpret_cast<unsigned int *>(to + head);
                    for (unsigned int i = lane; i < nw; i += 32) tw[i] = b4;
                    if (tail0 + (unsigned int)lane < mlen) to[tail0 + lane] = (unsigned char)b;
                } else if (mdist >= mlen) {
                    for (unsigned int i = lane; i < mlen; i += 32) to[i] = from[i];
                } else {
                    // overlapping copy = the last mdist bytes repeated; i % mdist kept incrementally
                    unsigned int j = (unsigned int)lane % mdist;
                    const unsigned int step = 32u % mdist;
                    for (unsigned int i = lane; i < mlen; i += 32) {
                        to[i] = from[j];
                        j += step;
                        if (j >= mdist) j -= mdist;
                    }
                }
                out += mlen;
                __syncwarp();
            }
            if (p > nbits && st == INF_OK) st = INF_IN_OVERRUN;
        }
        // ---- trailer: Adler-32 of the output, big-endian, byte aligned
        if (st == INF_OK) {
            p = (p + 7u) & ~7u;
            const unsigned int want = __byte_perm(peek32(src, p), 0, 0x0123);
            p += 32;
            if (p > nbits) st = INF_IN_OVERRUN;
            else if (out != J.dst_bytes) st = INF_OUT_SHORT;
            if (st == INF_OK) {
                __syncwarp();
                __threadfence_block();
                if (warp_adler32(dst, J.dst_bytes, lane) != want) st = INF_BAD_CHECKSUM;
            }
        }
        if (lane == 0) status[job] = st;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_inflate4: EVERY symbol of a 256-bit window is decoded speculatively, length / distance pairs included.
//
// Source-level stall samples of k_inflate3 on the dense bench matrix (short matches every few literals) are flat: ~460
// cycles per symbol spread over ~100 dependent instructions -- a serial program, one warp, 4 - 6 cycles per dependent
// issue.  The cure is fewer SERIAL instructions per symbol.  Here the lanes decode, for each of the next 256 bit
// positions, the complete symbol that would start there (literal, or length code + extra bits + distance code + extra
// bits, at most 48 bits, from one 64-bit peek); the real symbols are then found by the chain walk (one dependent 64-bit
// shared-memory load each), their output offsets by a warp prefix sum, literals are written at once, short matches whose
// source precedes the batch are copied by their own lanes, and only the remaining matches (long, or reading what the
// batch itself produces: the distance-1 zero runs of a sparse matrix) are copied one after the other by the whole warp.
// Positions that do not hold a valid table-resident symbol stop the chain and take k_inflate2's one-symbol path.
// ------------------------------------------------------------------------------------------------
static constexpr unsigned int INF4_STOP = 0x80000000u, INF4_END = 0x40000000u, INF4_MATCH = 0x20000000u;

struct InflateWarp4 {                           // per-warp shared memory (6 016 B), dynamic
    unsigned short lit_tab[1 << INF_LIT_BITS];
    unsigned short dist_tab[1 << INF_DIST_BITS];
    unsigned short lit_sym[288], dist_sym[32];
    unsigned short lit_cnt[16], dist_cnt[16];
    unsigned char lens[320];
    uint2 sym[256 + 48];                        // .x = address of the next entry | literal << 18 | flags, .y = length | distance << 16
};

__global__ void __launch_bounds__(INF_WARPS * 32) k_inflate4(const unsigned char *src_base, const InflateJob *jobs,
                                                              int n_jobs, int *status, unsigned int *next_job) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    InflateWarp4 *s_w = reinterpret_cast<InflateWarp4 *>(s_dyn);
    // length / distance code tables next to the decoder (constant-memory lookups with a data-dependent index cost a
    // constant-cache round trip each, four per match)
    __shared__ unsigned short s_len_base[29], s_dist_base[30];
    __shared__ unsigned char s_len_extra[29], s_dist_extra[30];
    __shared__ unsigned int s_lenx[32], s_distx[32];              // base | extra bits << 16, one lookup instead of two
    if (threadIdx.x < 32) {
        s_lenx[threadIdx.x] = threadIdx.x < 29 ? (unsigned int)c_len_base[threadIdx.x] | ((unsigned int)c_len_extra[threadIdx.x] << 16) : 0u;
        s_distx[threadIdx.x] = threadIdx.x < 30 ? (unsigned int)c_dist_base[threadIdx.x] | ((unsigned int)c_dist_extra[threadIdx.x] << 16) : 0u;
    }
    if (threadIdx.x < 29) {
        s_len_base[threadIdx.x] = c_len_base[threadIdx.x];
        s_len_extra[threadIdx.x] = c_len_extra[threadIdx.x];
    }
    if (threadIdx.x < 30) {
        s_dist_base[threadIdx.x] = c_dist_base[threadIdx.x];
        s_dist_extra[threadIdx.x] = c_dist_extra[threadIdx.x];
    }
    __syncthreads();
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    InflateWarp4 &S = s_w[wid];
    const unsigned int sym_base = (unsigned int)__cvta_generic_to_shared(S.sym);
    for (;;) {
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(next_job, 1u);
        job = __shfl_sync(full, job, 0);
        if (job >= n_jobs) return;
        const InflateJob J = jobs[job];
        unsigned char *dst = J.dst;
        BitSrc src;
        src.w = reinterpret_cast<const unsigned int *>(src_base + J.src_off);
        src.n_words = (J.src_bytes + 3) / 4;
        const unsigned int nbits = J.src_bytes * 8u;
        int st = INF_OK;
        unsigned int out = 0, p = 16;
        unsigned int pf_line = 0xFFFFFFFFu;                          // last 128-byte line of the stream that was prefetched
        {
            const unsigned int w = peek32(src, 0);
            const unsigned int cmf = w & 255u, flg = (w >> 8) & 255u;
            if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) st = INF_BAD_HEADER;
            if (J.src_bytes >= (1u << 28)) st = INF_IN_OVERRUN;            // bit positions are 32-bit
        }
        bool last = false;
        while (st == INF_OK && !last) {
            const unsigned int hw = peek32(src, p);
            last = (hw & 1u) != 0;
            const int btype = (int)((hw >> 1) & 3u);
            p += 3;
            if (btype == 0) {
                p = (p + 7u) & ~7u;
                const unsigned int lw = peek32(src, p);
                const unsigned int len = lw & 0xFFFFu, nlen = lw >> 16;
                p += 32;
                const unsigned int byte_pos = p >> 3;
                if ((len ^ nlen) != 0xFFFFu) st = INF_BAD_BLOCK;
                else if (byte_pos + len > J.src_bytes) st = INF_IN_OVERRUN;
                else if (out + len > J.dst_bytes) st = INF_OUT_OVERFLOW;
                if (st != INF_OK) break;
                const unsigned char *sb = src_base + J.src_off + byte_pos;
                for (unsigned int i = lane; i < len; i += 32) dst[out + i] = sb[i];
                out += len;
                p += len * 8u;
                __syncwarp();
                continue;
            }
            if (btype == 3) {
                st = INF_BAD_BLOCK;
                break;
            }
            int hlit = 288, hdist = 30;
            if (btype == 1) {
                for (int i = lane; i < 288; i += 32) S.lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                for (int i = lane; i < 32; i += 32) S.lens[288 + i] = i < 30 ? 5 : 0;
            } else {
                st = read_dynamic_lengths2(src, p, S.lens, hlit, hdist, lane);
                if (st != INF_OK) break;
            }
            __syncwarp();
            if (!build_tables(S.lens, hlit, S.lit_cnt, S.lit_sym, S.lit_tab, INF_LIT_BITS, lane, true) ||
                !build_tables(S.lens + hlit, hdist, S.dist_cnt, S.dist_sym, S.dist_tab, INF_DIST_BITS, lane, false)) {
                st = INF_BAD_LENGTHS;
                break;
            }
            // ---- symbols of the block
            for (;;) {
                // (1) every symbol that starts inside the next 256 bits -- literals AND length / distance pairs -- is decoded
                // speculatively: lane l decodes the symbols that WOULD start at bit p + 8 l + j (j < 8) from one 64-bit
                // peek (a symbol is at most 15 + 5 + 15 + 13 = 48 bits) and leaves {where it leads, literal | length,
                // distance} per position; positions that do not hold a table-resident, valid symbol (end of block, codes
                // longer than the tables, bad codes) are STOP entries and go through the one-symbol path (2) below.
                if ((p >> 10) != pf_line) {                            // entering a new 128-byte line: fetch the next one into L1
                    pf_line = p >> 10;
                    const unsigned int nw = (pf_line + 1u) << 5;
                    if (nw < src.n_words) asm volatile("prefetch.global.L1 [%0];" ::"l"(src.w + nw));
                }
                {
                    const unsigned long long w = peek64(src, p + 8u * (unsigned int)lane);
                    const unsigned int wlo = (unsigned int)w, whi = (unsigned int)(w >> 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // 32-bit views only (64-bit variable shifts cost two or three instructions each): a = bits
                        // [j, j + 32) of the peek, ah = bits [j + 32, 64); after the length code and its extra bits
                        // (<= 20 bits) t = the next 32 bits, enough for the distance code and its extra bits (<= 28)
                        const unsigned int a = __funnelshift_r(wlo, whi, j), ah = whi >> j;
                        const unsigned int c = 8u * (unsigned int)lane + (unsigned int)j;
                        const unsigned int e = S.lit_tab[a & ((1u << INF_LIT_BITS) - 1u)];
                        unsigned int lo = (sym_base + (c << 3)) | INF4_STOP, hi = 0u;
                        unsigned int used = e & 15u;
                        if (e & 0x8000u) {                                             // literal
                            lo = (sym_base + ((c + used) << 3)) | (((e >> 4) & 0xFFu) << 18);
                        } else if (used != 0u) {
                            const unsigned int ls = ((e >> 4) & 0x1FFu) - 257u;        // (256 = end of block wraps to a large value)
                            if (ls < 29u) {
                                const unsigned int lx = s_lenx[ls];
                                unsigned int eb = lx >> 16;
                                const unsigned int mlen = (lx & 0xFFFFu) + ((a >> used) & ((1u << eb) - 1u));
                                used += eb;
                                const unsigned int t = __funnelshift_r(a, ah, used);
                                const unsigned int e2 = S.dist_tab[t & ((1u << INF_DIST_BITS) - 1u)];
                                const unsigned int dl = e2 & 15u, ds = e2 >> 4;
                                if (dl != 0u && ds < 30u) {
                                    const unsigned int dx = s_distx[ds];
                                    eb = dx >> 16;
                                    const unsigned int mdist = (dx & 0xFFFFu) + ((t >> dl) & ((1u << eb) - 1u));
                                    used += dl + eb;
                                    lo = (sym_base + ((c + used) << 3)) | INF4_MATCH;
                                    hi = mlen | (mdist << 16);
                                }
                            }
                        }
                        S.sym[c] = make_uint2(lo, hi);
                    }
                }
                for (unsigned int c = 256u + (unsigned int)lane; c < 256u + 48u; c += 32u)
                    S.sym[c] = make_uint2((sym_base + (c << 3)) | INF4_END, 0u);
                __syncwarp();
                // the chain p -> next -> ...: one dependent 64-bit shared-memory load per symbol; lane k keeps symbol k
                unsigned int n = 0, mlo = INF4_STOP, mhi = 0u, addr = sym_base;
#pragma unroll 1
                for (;;) {
                    unsigned int lo, hi;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
                    if (lo & (INF4_STOP | INF4_END)) break;
                    if (n == (unsigned int)lane) {
                        mlo = lo;
                        mhi = hi;
                    }
                    addr = lo & 0x3FFFFu;
                    if (++n == 32u) break;
                }
                const unsigned int pos = (addr - sym_base) >> 3;
                const bool stop = (S.sym[pos].x & INF4_STOP) != 0u;
                __syncwarp();                                              // (S.sym is rewritten by the next iteration)
                // where every symbol's output goes: exclusive prefix sum of the output lengths
                const bool have = (unsigned int)lane < n, is_match = have && (mlo & INF4_MATCH) != 0u;
                const unsigned int s_len = mhi & 0xFFFFu, s_dist = mhi >> 16;
                const unsigned int mylen = have ? (is_match ? s_len : 1u) : 0u;
                unsigned int incl = mylen;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned int t = __shfl_up_sync(full, incl, d);
                    if (lane >= d) incl += t;
                }
                const unsigned int total = __shfl_sync(full, incl, 31), off = incl - mylen;
                if (out + total > J.dst_bytes) {
                    st = INF_OUT_OVERFLOW;
                    break;
                }
                if (__any_sync(full, is_match && s_dist > out + off)) {
                    st = INF_BAD_DISTANCE;
                    break;
                }
                if (have && !is_match) dst[out + off] = (unsigned char)((mlo >> 18) & 0xFFu);
                // matches whose source lies wholly before this batch's output and that are short: every lane copies its own
                const bool indep = is_match && s_dist >= off + s_len && s_len <= 16u;
                if (indep) {
                    const unsigned char *from = dst + out + off - s_dist;
                    unsigned char *to = dst + out + off;
                    for (unsigned int i = 0; i < s_len; ++i) to[i] = from[i];
                }
                // the others in symbol order, copied by the whole warp (they may read what earlier symbols of the batch wrote)
                unsigned int todo = __ballot_sync(full, is_match && !indep);
                while (todo) {
                    const int k = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const unsigned int koff = __shfl_sync(full, off, k), khi = __shfl_sync(full, mhi, k);
                    const unsigned int klen = khi & 0xFFFFu, kdist = khi >> 16;
                    __syncwarp();                                          // the match source may have just been written
                    const unsigned char *from = dst + out + koff - kdist;
                    unsigned char *to = dst + out + koff;
                    if (kdist == 1u) {
                        // a run of one byte value (the zero runs of a sparse k-mer matrix): aligned 32-bit stores
                        const unsigned int b = from[0], b4 = b * 0x01010101u;
                        const unsigned int head = min(klen, (4u - ((unsigned int)(size_t)to & 3u)) & 3u);
                        const unsigned int nw = (klen - head) >> 2, tail0 = head + (nw << 2);
                        if ((unsigned int)lane < head) to[lane] = (unsigned char)b;
                        unsigned int *tw = reinterpret_cast<unsigned int *>(to + head);
                        for (unsigned int i = lane; i < nw; i += 32) tw[i] = b4;
                        if (tail0 + (unsigned int)lane < klen) to[tail0 + lane] = (unsigned char)b;
                    } else if (kdist >= klen) {
                        for (unsigned int i = lane; i < klen; i += 32) to[i] = from[i];
                    } else {
                        unsigned int jj = (unsigned int)lane % kdist;
                        const unsigned int step = 32u % kdist;
                        for (unsigned int i = lane; i < klen; i += 32) {
                            to[i] = from[jj];
                            jj += step;
                            if (jj >= kdist) jj -= kdist;
                        }
                    }
                }
                __syncwarp();
                out += total;
                p += pos;
                if (p > nbits) {
                    st = INF_IN_OVERRUN;
                    break;
                }
                if (!stop) continue;                                       // the run goes on: next window
                // (2) the symbol at p is not a table-resident literal: one symbol, decoded identically by every lane
                unsigned long long q = peek64(src, p);
                unsigned int used;
                int sym;
                {
                    const unsigned int e1 = S.lit_tab[(unsigned int)q & ((1u << INF_LIT_BITS) - 1u)];
                    if (e1 & 15u) {
                        used = e1 & 15u;
                        sym = (int)((e1 >> 4) & 0x1FFu);
                    } else {
                        const int r = slow_decode((unsigned int)q, S.lit_cnt, S.lit_sym);
                        if (r < 0) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        used = (unsigned int)(r >> 16);
                        sym = r & 0xFFFF;
                    }
                }
                q >>= used;
                if (sym < 256) {                                           // a literal with a code longer than the table
                    if (out + 1 > J.dst_bytes) {
                        st = INF_OUT_OVERFLOW;
                        break;
                    }
                    if (lane == 0) dst[out] = (unsigned char)sym;
                    out += 1;
                    p += used;
                    continue;
                }
                if (sym == 256) {
                    p += used;
                    break;                                                 // end of block
                }
                sym -= 257;
                if (sym >= 29) {
                    st = INF_BAD_CODE;
                    break;
                }
                unsigned int eb = s_len_extra[sym];
                const unsigned int mlen = s_len_base[sym] + ((unsigned int)q & ((1u << eb) - 1u));
                q >>= eb;
                used += eb;
                int ds;
                {
                    const unsigned int e2 = S.dist_tab[(unsigned int)q & ((1u << INF_DIST_BITS) - 1u)];
                    unsigned int dl;
                    if (e2 & 15u) {
                        dl = e2 & 15u;
                        ds = (int)(e2 >> 4);
                    } else {
                        const int r = slow_decode((unsigned int)q, S.dist_cnt, S.dist_sym);
                        if (r < 0) {
                            st = INF_BAD_CODE;
                            break;
                        }
                        dl = (unsigned int)(r >> 16);
                        ds = r & 0xFFFF;
                    }
                    q >>= dl;
                    used += dl;
                }
                if (ds >= 30) {
                    st = INF_BAD_CODE;
                    break;
                }
                eb = s_dist_extra[ds];
                const unsigned int mdist = s_dist_base[ds] + ((unsigned int)q & ((1u << eb) - 1u));
                used += eb;
                p += used;
                if (p > nbits) {
                    st = INF_IN_OVERRUN;
                    break;
                }
                if (mdist > out) {
                    st = INF_BAD_DISTANCE;
                    break;
                }
                if (out + mlen > J.dst_bytes) {
                    st = INF_OUT_OVERFLOW;
                    break;
                }
                __syncwarp();                                              // the match source may have just been written
                const unsigned char *from = dst + out - mdist;
                unsigned char *to = dst + out;
                if (mdist == 1u) {
                    // a run of one byte value (the zero runs of a sparse k-mer matrix): aligned 32-bit stores
                    const unsigned int b = from[0], b4 = b * 0x01010101u;
                    const unsigned int head = min(mlen, (4u - ((unsigned int)(size_t)to & 3u)) & 3u);
                    const unsigned int nw = (mlen - head) >> 2, tail0 = head + (nw << 2);
                    if ((unsigned int)lane < head) to[lane] = (unsigned char)b;
                    unsigned int *tw = reinterpret_cast<unsigned int *>(to + head);
                    for (unsigned int i = lane; i < nw; i += 32) tw[i] = b4;
                    if (tail0 + (unsigned int)lane < mlen) to[tail0 + lane] = (unsigned char)b;
                } else if (mdist >= mlen) {
                    for (unsigned int i = lane; i < mlen; i += 32) to[i] = from[i];
                } else {
                    // overlapping copy = the last mdist bytes repeated; i % mdist kept incrementally
                    unsigned int j = (unsigned int)lane % mdist;
                    const unsigned int step = 32u % mdist;
                    for (unsigned int i = lane; i < mlen; i += 32) {
                        to[i] = from[j];
                        j += step;
                        if (j >= mdist) j -= mdist;
                    }
                }
                out += mlen;
                __syncwarp();
            }
            if (p > nbits && st == INF_OK) st = INF_IN_OVERRUN;
        }
        // ---- trailer: Adler-32 of the output, big-endian, byte aligned
        if (st == INF_OK) {
            p = (p + 7u) & ~7u;
            const unsigned int want = __byte_perm(peek32(src, p), 0, 0x0123);
            p += 32;
            if (p > nbits) st = INF_IN_OVERRUN;
            else if (out != J.dst_bytes) st = INF_OUT_SHORT;
            if (st == INF_OK) {
                __syncwarp();
                __threadfence_block();
                if (warp_adler32(dst, J.dst_bytes, lane) != want) st = INF_BAD_CHECKSUM;
            }
        }
        if (lane == 0) status[job] = st;
        __syncwarp();
    }
}

// Places the inflated chunks into the [W, ld] image.  A chunk holds chunk_rows x chunk_cols elements of the dataset
// starting at (row0, gcol0); only the part inside the dataset and inside this shard's columns is kept (HDF5 stores
// edge chunks at full size; shards are cut on 512-column tiles, so a chunk can belong to two shards).
struct ScatterJob {
    const unsigned char *src;       // the inflated chunk in the scratch slab
    long long row0, gcol0;          // dataset coordinates of the chunk's first element (row in pack_bits units)
};

template <typename T>
__global__ void __launch_bounds__(256) k_scatter_chunks(const ScatterJob *jobs, uint64_t *mat,
                                                         long long ld, long long shard_col0, long long shard_cols,
                                                         long long n_rows_total, long long n_cols_total,
                                                         int chunk_rows, long long chunk_cols) {
    const ScatterJob J = jobs[blockIdx.y];
    const T *src = reinterpret_cast<const T *>(J.src);
    const long long per_chunk = (long long)chunk_rows * chunk_cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_chunk; e += (long long)gridDim.x * blockDim.x) {
        const long long r = J.row0 + e / chunk_cols, c = J.gcol0 + e % chunk_cols;
        if (r >= n_rows_total || c >= n_cols_total) continue;
        const long long lc = c - shard_col0;
        if (lc < 0 || lc >= shard_cols) continue;
        if (sizeof(T) == 8) {
            mat[r * ld + lc] = (uint64_t)src[e];
        } else {
            // 32-bit packing (rules.py:123-128): packed row 2w is the HIGH half of word w -- written as a 32-bit store
            uint32_t *m32 = reinterpret_cast<uint32_t *>(mat + (r >> 1) * ld + lc);
            m32[(r & 1) ? 0 : 1] = (uint32_t)src[e];
        }
    }
}
