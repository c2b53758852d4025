/* TEST INFRASTRUCTURE ONLY -- CPU restatement of Kover's native popcount kernel.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library; the product path (kover_b200/) never does.
 *
 * Follows (reference file:line, relative to /root/reference/core/kover):
 *   learning/common/popcount.pyx:76-95  _inplace_popcount_64_2d
 *   learning/common/popcount.pyx:31-50  _inplace_popcount_32_2d
 *   learning/common/rules.py:247-262    the (row block x column block) accumulate loop
 *
 * Parity pinning: checked bit-for-bit against the reference's own compiled popcount.pyx
 * (oracle/_ref/popcount*.so) in tests/test_oracle_vs_reference.py and against the golden
 * vectors in tests/golden/ minted from the real reference.
 */
#include <stdint.h>
#include <stddef.h>

/* popcount.pyx:76-95 -- every non-zero word of an [m, n] C-contiguous block is replaced,
 * in place, by popcount(word & row_mask[i]); zero words are left untouched. */
void oracle_inplace_popcount_64_2d(uint64_t *arr, long m, long n, const uint64_t *row_mask)
{
    for (long i = 0; i < m; ++i) {
        uint64_t *row = arr + (size_t)i * (size_t)n;
        const uint64_t mk = row_mask[i];
        for (long j = 0; j < n; ++j) {
            if (row[j] != 0)
                row[j] = (uint64_t)__builtin_popcountl(row[j] & mk);
        }
    }
}

/* popcount.pyx:31-50 -- 32-bit twin. */
void oracle_inplace_popcount_32_2d(uint32_t *arr, long m, long n, const uint32_t *row_mask)
{
    for (long i = 0; i < m; ++i) {
        uint32_t *row = arr + (size_t)i * (size_t)n;
        const uint32_t mk = row_mask[i];
        for (long j = 0; j < n; ++j) {
            if (row[j] != 0)
                row[j] = (uint32_t)__builtin_popcount(row[j] & mk);
        }
    }
}

/* rules.py:247-262 collapsed: presence count of every column over the packed rows whose mask
 * word is non-zero (rules.py:244).  out[j] += sum_i popcount(mat[i, j] & mask[i]).
 * Used for oracle runs at sizes where the numpy block loop would take minutes; the block
 * loop itself (the thing the CPU baseline times) lives in kover_oracle.py. */
void oracle_masked_column_counts_64(const uint64_t *mat, long n_words, long n_cols, long ld,
                                    const uint64_t *mask, uint64_t *out)
{
    for (long i = 0; i < n_words; ++i) {
        const uint64_t mk = mask[i];
        if (mk == 0) continue;
        const uint64_t *row = mat + (size_t)i * (size_t)ld;
        for (long j = 0; j < n_cols; ++j)
            out[j] += (uint64_t)__builtin_popcountl(row[j] & mk);
    }
}
