"""Seeded inputs shared by the golden-vector generator (make_golden.py, which runs the REAL
reference in the build container) and by the parity tests (which rebuild the same inputs and
compare the oracle / the CUDA path against the committed outputs).  numpy only."""
import numpy as np


def pack64(X):
    """Example i -> word i//64, bit 63-(i%64)  (reference utils.py:133-156).  Vectorised; checked
    against the reference's own packer in the 'pack' golden case."""
    n, k = X.shape
    W = (n + 63) // 64
    pad = np.zeros((W * 64, k), dtype=np.uint8)
    pad[:n] = X
    bits = pad.reshape(W, 64, k).astype(np.uint64)
    shifts = (np.uint64(63) - np.arange(64, dtype=np.uint64)).reshape(1, 64, 1)
    return np.bitwise_or.reduce(bits << shifts, axis=1)


def pack32_from64(P):
    W, k = P.shape
    out = np.zeros((2 * W, k), dtype=np.uint32)
    out[0::2] = (P >> np.uint64(32)).astype(np.uint32)
    out[1::2] = (P & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return out


def dense_bits(seed, n, k, density):
    rs = np.random.RandomState(seed)
    return (rs.rand(n, k) < density).astype(np.uint8)


def random_words(seed, n, k):
    """Packed matrix drawn as raw 64-bit words (for wide matrices), padding bits zeroed."""
    rs = np.random.RandomState(seed)
    W = (n + 63) // 64
    P = rs.randint(0, 2 ** 63, size=(W, k), dtype=np.int64).astype(np.uint64)
    P ^= rs.randint(0, 2, size=(W, k), dtype=np.int64).astype(np.uint64) << np.uint64(63)
    rem = n % 64
    if rem:
        P[-1] &= np.uint64(((1 << rem) - 1) << (64 - rem))
    return P


# ----------------------------------------------------------------------------- matrices
def matrix(name):
    """-> (X uint8 [n,k] or None, packed uint64 [W,k], n_genomes, chunks)"""
    if name == "M1":      # SURVEY appendix A5 SCM matrix
        X = dense_bits(0, 200, 5000, 0.3)
        return X, pack64(X), 200, (1, 1000)
    if name == "M2":      # SURVEY appendix A5 CART matrix; >255 genomes => uint16 counts
        X = dense_bits(1, 300, 4000, 0.4)
        return X, pack64(X), 300, (1, 1000)
    if name == "M3":      # exact word boundary, odd column count, per-column densities incl. all-0 / all-1
        rs = np.random.RandomState(3)
        dens = rs.rand(777)
        dens[:5] = 0.0
        dens[5:10] = 1.0
        X = (rs.rand(192, 777) < dens[None, :]).astype(np.uint8)
        return X, pack64(X), 192, (1, 100)
    if name == "M4":      # ragged last word (n % 64 = 1), no chunks attribute (block = whole row)
        X = dense_bits(4, 65, 333, 0.5)
        return X, pack64(X), 65, None
    if name == "MD":      # duplicate-heavy: 300 base columns each repeated ~10x, plus complements =>
        rs = np.random.RandomState(6)   # large equivalent-rule sets straddling utility blocks and the pres/abs halves
        dens = rs.choice([0.05, 0.2, 0.5, 0.8, 0.95], size=300)
        base = (rs.rand(120, 300) < dens[None, :]).astype(np.uint8)
        src = rs.randint(0, 300, size=3000)
        X = base[:, src]
        comp = rs.rand(3000) < 0.2
        X[:, comp] = 1 - X[:, comp]
        return X, pack64(X), 120, (1, 512)
    if name == "MS":      # few genomes => utilities take few distinct values => big tie sets
        X = dense_bits(7, 20, 2500, 0.5)
        return X, pack64(X), 20, (1, 1000)
    if name == "MW":      # wide: 2K = 1.2M rules => two 1M-rule utility blocks, absence half unaligned
        P = random_words(5, 100, 600000)
        return None, P, 100, (1, 100000)
    raise KeyError(name)


def row_sets(n, seed=11):
    rs = np.random.RandomState(seed)
    half = np.sort(rs.choice(n, n // 2, replace=False))
    shuffled = rs.permutation(half)
    sets = {
        "empty": np.array([], dtype=np.int64),
        "first": np.array([0]),
        "last": np.array([n - 1]),
        "all": np.arange(n),
        "half": half,
        "half_unsorted": shuffled,
        "first64": np.arange(min(64, n)),
        "straddle": np.array([i for i in (62, 63, 64, 65) if i < n]),
    }
    return sets


def column_sets(k):
    return {
        "int_presence": 7,
        "int_absence": k + 11,
        "list": [3, 9, 20],
        # NB: a multi-element ndarray is NOT accepted by the reference: ndarray has __index__, so
        # rules.py:141-142 calls it and numpy raises TypeError.  Lists are the supported container.
        "mixed": [5, k + 5, k - 1, 2 * k - 1],
        "np_int": np.int64(k + 3),
        "dups": [4, 4, k + 4],
    }


# ----------------------------------------------------------------------------- labels
def labels(name):
    if name == "M1_planted":          # appendix A5: X7 & ~X11
        X = matrix("M1")[0]
        return (X[:, 7] & (1 - X[:, 11])).astype(np.uint8)
    if name == "M1_noisy":
        y = labels("M1_planted").copy()
        rs = np.random.RandomState(21)
        flip = rs.rand(y.shape[0]) < 0.15
        y[flip] = 1 - y[flip]
        return y
    if name == "M2_planted":          # appendix A5: (X3 & X9) | ~X20
        X = matrix("M2")[0]
        return ((X[:, 3] & X[:, 9]) | (1 - X[:, 20])).astype(np.uint8)
    if name == "M1_3class":
        X = matrix("M1")[0]
        y = (X[:, 17] + 2 * (X[:, 17] == 0) * X[:, 42]).astype(np.uint8)   # 0,1,2
        rs = np.random.RandomState(22)
        flip = rs.rand(y.shape[0]) < 0.1
        y[flip] = rs.randint(0, 3, size=int(flip.sum()))
        return y
    if name == "MD_planted":
        X = matrix("MD")[0]
        y = (X[:, 10] & (1 - X[:, 20]) & X[:, 30]).astype(np.uint8)
        rs = np.random.RandomState(25)
        flip = rs.rand(y.shape[0]) < 0.08
        y[flip] = 1 - y[flip]
        return y
    if name == "MS_random":
        return (np.random.RandomState(26).rand(20) < 0.5).astype(np.uint8)
    if name == "MW_random":
        return (np.random.RandomState(23).rand(100) < 0.5).astype(np.uint8)
    raise KeyError(name)


def rule_risks(n_rules, seed=31, n_levels=40):
    """Stand-in for hstack(unique_risk_by_kmer, unique_risk_by_anti_kmer) (experiment_scm.py:135-137):
    small unsigned indices into a sorted unique-risk table."""
    return np.random.RandomState(seed).randint(0, n_levels, size=n_rules).astype(np.uint8)


# ----------------------------------------------------------------------------- case tables
UTILITY_CASES = [
    # name, matrix, labels, p, blacklist, util_block_size
    ("u_M1_p1", "M1", "M1_planted", 1.0, [], 1000000),
    ("u_M1_p01", "M1", "M1_noisy", 0.1, [], 1000000),
    ("u_M1_pbig", "M1", "M1_noisy", 999999.0, [], 1000000),
    ("u_M1_black", "M1", "M1_planted", 1.0, [7, 5007, 11], 1000000),
    ("u_M1_blk1000_p1", "M1", "M1_noisy", 1.0, [], 1000),
    ("u_M1_blk1000_p0316", "M1", "M1_noisy", 0.316, [], 1000),
    ("u_M1_blk1000_pbig", "M1", "M1_noisy", 999999.0, [], 1000),
    ("u_M1_blk1000_firstblack", "M1", "M1_noisy", 1.0, list(range(0, 1000)) + [4321, 9999], 1000),
    ("u_M1_blk777_p1778", "M1", "M1_noisy", 1.778, [], 777),
    ("u_M3_p1", "M3", None, 1.0, [], 100),
    ("u_MD_p1", "MD", "MD_planted", 1.0, [], 1000000),
    ("u_MD_blk1000_p1", "MD", "MD_planted", 1.0, [], 1000),
    ("u_MD_blk1000_p0562", "MD", "MD_planted", 0.562, [], 1000),
    ("u_MD_blk777_pbig", "MD", "MD_planted", 999999.0, [], 777),
    ("u_MD_blk512_p10_black", "MD", "MD_planted", 10.0, list(range(512, 1024)) + [10, 3010, 5999], 512),
    ("u_MS_p1", "MS", "MS_random", 1.0, [], 1000000),
    ("u_MS_blk1000_p3162", "MS", "MS_random", 3.162, [], 1000),
    ("u_MS_blk100_pbig", "MS", "MS_random", 999999.0, [], 100),
    ("u_MS_blk100_allblack_head", "MS", "MS_random", 1.0, list(range(0, 250)), 100),
    ("u_MW_p1", "MW", "MW_random", 1.0, [], 1000000),
    ("u_MW_p05", "MW", "MW_random", 0.5, [3, 600003], 1000000),
    ("u_MW_pbig", "MW", "MW_random", 999999.0, [], 1000000),
]

SCM_CASES = [
    # name, matrix, labels, model_type, p, max_rules, tiebreak ('identity'|'risk'), blacklist, util_block_size
    ("scm_appendix", "M1", "M1_planted", "conjunction", 1.0, 5, "identity", [], 1000000),
    ("scm_noisy_conj_p1", "M1", "M1_noisy", "conjunction", 1.0, 4, "risk", [], 1000000),
    ("scm_noisy_conj_p01", "M1", "M1_noisy", "conjunction", 0.1, 4, "risk", [], 1000),
    ("scm_noisy_disj_p1", "M1", "M1_noisy", "disjunction", 1.0, 4, "risk", [], 1000000),
    ("scm_noisy_disj_p10", "M1", "M1_noisy", "disjunction", 10.0, 3, "risk", [7, 5007], 1000),
    ("scm_MD_conj_p1", "MD", "MD_planted", "conjunction", 1.0, 5, "risk", [], 1000),
    ("scm_MD_disj_p05", "MD", "MD_planted", "disjunction", 0.5, 5, "risk", [], 777),
    ("scm_MD_conj_pbig", "MD", "MD_planted", "conjunction", 999999.0, 4, "risk", [], 1000000),
    ("scm_MD_conj_identity", "MD", "MD_planted", "conjunction", 2.0, 10, "identity", [10, 3010], 1000),
    ("scm_MS_conj_p1", "MS", "MS_random", "conjunction", 1.0, 6, "risk", [], 100),
    ("scm_MS_disj_p1778", "MS", "MS_random", "disjunction", 1.778, 6, "identity", [], 1000000),
    ("scm_noisy_conj_pbig", "M1", "M1_noisy", "conjunction", 999999.0, 3, "risk", [], 1000000),
]

CART_CASES = [
    # name, matrix, labels, max_depth, min_samples_split, class_importance, tiebreak ('none'|'occurrence'), train subset seed
    ("cart_appendix", "M2", "M2_planted", 10, 2, {0: 1.0, 1: 1.0}, "none", None),
    ("cart_noisy_occ", "M1", "M1_noisy", 4, 2, {0: 1.0, 1: 1.0}, "occurrence", 41),
    ("cart_3class", "M1", "M1_3class", 4, 5, {0: 1.0, 1: 2.0, 2: 0.5}, "occurrence", 42),
    ("cart_MD_occ", "MD", "MD_planted", 5, 2, {0: 1.0, 1: 1.0}, "occurrence", 43),
    ("cart_MD_none", "MD", "MD_planted", 3, 4, {0: 1.0, 1: 2.5}, "none", None),
    ("cart_MS", "MS", "MS_random", 6, 2, {0: 1.0, 1: 1.0}, "occurrence", None),
    ("cart_weighted", "M2", "M2_planted", 3, 10, {0: 3.0, 1: 1.0}, "occurrence", None),
]


CART_CE_CASES = [
    # the cross-entropy criterion (cart.py:167-207); same fields as CART_CASES
    ("cartce_appendix", "M2", "M2_planted", 4, 2, {0: 1.0, 1: 1.0}, "none", None),
    ("cartce_noisy_occ", "M1", "M1_noisy", 3, 2, {0: 1.0, 1: 2.0}, "occurrence", 41),
    ("cartce_3class", "M1", "M1_3class", 3, 5, {0: 1.0, 1: 2.0, 2: 0.5}, "occurrence", 42),
    ("cartce_MD", "MD", "MD_planted", 4, 4, {0: 1.0, 1: 1.0}, "occurrence", 43),
]


EXPERIMENT_CASES = [
    # name, matrix, labels, train subset seed, p values, max_rules, max_equiv_rules, n_folds
    ("exp_M1", "M1", "M1_noisy", 41, [0.1, 1.0, 10.0], 4, 3, 3),
    ("exp_MD", "MD", "MD_planted", 43, [0.5, 2.0], 3, 10000, 4),
]


def experiment_folds(train, n_folds, seed=51):
    """(fold_train_idx, fold_test_idx) like dataset/split.py:196-207."""
    rs = np.random.RandomState(seed)
    fold_of = np.arange(len(train)) % n_folds
    rs.shuffle(fold_of)
    return [(train[fold_of != f], train[fold_of == f]) for f in range(n_folds)]


CART_EXPERIMENT_CASES = [
    # name, matrix, labels, train subset seed, max_depth, min_samples_split, class_importance
    ("cartexp_M2", "M2", "M2_planted", 41, 10, 2, {0: 1.0, 1: 1.0}),
    ("cartexp_M1_noisy", "M1", "M1_noisy", 42, 6, 2, {0: 1.0, 1: 2.0}),
    ("cartexp_MD", "MD", "MD_planted", 43, 5, 4, {0: 1.0, 1: 1.0}),
]


CART_CV_CASES = [
    # name, matrix, labels, train subset seed, n folds, max_depth, min_samples_split, class_importance
    ("cartcv_M2", "M2", "M2_planted", 51, 3, 10, 2, {0: 1.0, 1: 1.0}),
    ("cartcv_M1_noisy", "M1", "M1_noisy", 52, 4, 6, 2, {0: 1.0, 1: 2.0}),
    ("cartcv_MD", "MD", "MD_planted", 53, 3, 5, 4, {0: 1.0, 1: 1.0}),
]


def cv_folds(train, n_folds, seed):
    """(train idx, test idx) per fold, drawn like `kover dataset split` (split.py:198-199)."""
    fold_of = np.random.RandomState(seed).randint(0, n_folds, size=len(train))
    return [(train[fold_of != f], train[fold_of == f]) for f in range(n_folds)]


# ----------------------------------------------------------------------------- dataset-file drivers
DRIVER_CASES = [
    # name, matrix, labels, seed, n folds, learner, parameter_selection, max_equiv_rules
    ("drv_scm_bound_M1", "M1", "M1_planted", 61, 3, "scm", "bound", 10000),
    ("drv_scm_cv_M1", "M1", "M1_noisy", 62, 4, "scm", "cv", 2),
    ("drv_scm_none_MD", "MD", "MD_planted", 63, 3, "scm", "none", 10000),
    ("drv_cart_bound_M2", "M2", "M2_planted", 64, 3, "cart", "bound", None),
    ("drv_cart_cv_M1", "M1", "M1_noisy", 65, 3, "cart", "cv", None),
]
DRIVER_SCM_ARGS = dict(model_type=["conjunction", "disjunction"], p=[0.5, 1.0, 2.0], max_rules=4, random_seed=42,
                       bound_delta=0.05, bound_max_genome_size=1000)
# (one class-importance dictionary: the reference's np.unique over dictionaries, experiment_cart.py:545, only runs on
# Python 2, where dictionaries are ordered)
DRIVER_CART_ARGS = dict(criterion=["gini"], max_depth=[6, 2, 4], min_samples_split=[2, 6],
                        class_importance=[{0: 1.0, 1: 2.0}], bound_delta=0.05, bound_max_genome_size=1000)


def driver_dataset(case):
    """Everything a Kover dataset file of a DRIVER case holds (numpy, no HDF5): the golden generator serves it to the
    reference through a stand-in KoverDataset, the tests write it to an HDF5 file."""
    name, mname, lname, seed, n_folds = case[:5]
    X, P, n, chunks = matrix(mname)
    y = labels(lname)
    k = P.shape[1]
    train = train_subset(n, seed)
    test = np.setdiff1d(np.arange(n), train)
    folds = cv_folds(train, n_folds, seed)
    seqs = ["".join("ACGT"[(i >> (2 * j)) & 3] for j in range(12)) for i in range(k)]     # sequence of matrix column i
    perm = np.random.RandomState(seed).permutation(k)      # kmer_by_matrix_column: column -> index in kmer_sequences
    ordered = [None] * k
    for col, seq_idx in enumerate(perm):
        ordered[seq_idx] = seqs[col]
    ids = ["genome_%d" % i for i in range(n)]
    return dict(P=P, n=n, k=k, chunks=chunks, y=y, train=train, test=test, folds=folds, kmer_sequences=ordered,
                kmer_by_matrix_column=perm, genome_identifiers=ids, column_sequences=seqs)


RISK_CASES = [
    # name, matrix, labels, train subset seed
    ("risk_M1", "M1", "M1_noisy", 41),
    ("risk_M2_all", "M2", "M2_planted", None),
    ("risk_MD", "MD", "MD_planted", 43),
    ("risk_MS", "MS", "MS_random", None),
]


def utility_inputs(case):
    name, mname, lname, p, blacklist, ubs = case
    X, P, n, chunks = matrix(mname)
    if lname is None:
        y = (np.random.RandomState(24).rand(n) < 0.4).astype(np.uint8)
    else:
        y = labels(lname)
    return P, n, chunks, np.where(y == 1)[0], np.where(y == 0)[0], p, blacklist, ubs


def train_subset(n, seed):
    if seed is None:
        return np.arange(n)
    return np.sort(np.random.RandomState(seed).choice(n, (2 * n) // 3, replace=False))
