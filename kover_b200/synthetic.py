"""Deterministic synthetic k-mer matrices (bench and full-size parity tests).

``synthetic_block`` is the numpy twin of the device generator (k_generate / synth_word in
csrc/kover_b200.cu): any column slice of the matrix that kv_generate_synthetic wrote into HBM can
be regenerated bit-for-bit on the host, so matrices far larger than host RAM (SURVEY section 8d:
configs C4/C5) can still be spot-checked against the CPU checker.

word(seed, k, w):  level = splitmix64(seed ^ (k * 0xD6E8FEB86659FD93)) & 7 picks the density of
column k; r1..r4 = splitmix64(seed + ((k << 13) + w) * 4 + {0,1,2,3}); the word is a fixed AND/OR
combination of r1..r4 giving bit densities 1/8, 1/4, 3/8, 1/2, 5/8, 3/4, 7/8, 1/16.  Padding bits
of the last packed row are cleared.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synthetic_block(seed, n_examples, col0, n_cols):
    """uint64 [W, n_cols]: columns col0 .. col0+n_cols of the synthetic matrix for ``seed``."""
    n_words = (n_examples + 63) // 64
    if n_words > 8192:
        raise ValueError("more than 8192 packed rows")
    seed = np.uint64(seed)
    with np.errstate(over="ignore"):
        k = np.arange(col0, col0 + n_cols, dtype=np.uint64)
        level = (splitmix64(seed ^ (k * np.uint64(0xD6E8FEB86659FD93))) & np.uint64(7)).astype(np.int64)
        w = np.arange(n_words, dtype=np.uint64).reshape(-1, 1)
        base = seed + ((k.reshape(1, -1) << np.uint64(13)) + w) * np.uint64(4)
        r1, r2, r3 = splitmix64(base), splitmix64(base + np.uint64(1)), splitmix64(base + np.uint64(2))
        r4 = splitmix64(base + np.uint64(3))
    variants = [r1 & r2 & r3, r1 & r2, r1 & (r2 | r3), r1, r1 | (r2 & r3), r1 | r2, r1 | r2 | r3, r1 & r2 & r3 & r4]
    out = np.zeros((n_words, n_cols), dtype=np.uint64)
    for lv, v in enumerate(variants):
        sel = level == lv
        out[:, sel] = v[:, sel]
    rem = n_examples % 64
    if rem:
        out[-1] &= np.uint64(((1 << rem) - 1) << (64 - rem))
    return out


def synthetic_labels(seed, n_examples, sorted_by_label=True):
    """Binary phenotype, ~50/50.  The reference's dataset builder sorts genomes by label
    (create.py:190-194), which is the default here; ``sorted_by_label=False`` interleaves them."""
    if sorted_by_label:
        y = np.zeros(n_examples, dtype=np.uint8)
        y[n_examples // 2:] = 1
        return y
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) ^ (np.arange(n_examples, dtype=np.uint64) * np.uint64(0xA24BAED4963EE407)))
    return (h >> np.uint64(63)).astype(np.uint8)


def train_split(seed, n_examples, train_fraction=2.0 / 3.0):
    """Sorted training example indices: a seeded random subset like `kover dataset split`
    (split.py:99-109 draws a permutation with numpy's RandomState)."""
    rs = np.random.RandomState(seed % (2 ** 32))
    perm = rs.permutation(n_examples)
    return np.sort(perm[:int(round(n_examples * train_fraction))])
