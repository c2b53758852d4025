"""GPU: the CUDA path against the CPU checker at the packed-row counts of BASELINE configs 3-5 and beyond.

W = ceil(n_genomes / 64) = 79 (config 3), 313 (config 4), 782 (config 5) and 1 719: above 64 active packed rows the
row records leave the kernel parameter block for device memory, above ~1 500 a smaller TMA ring is chosen
(launch_scan), above ~4 900 they are no longer staged in shared memory at all (the global-record scan variant), and
above ~7 800 the multi-mask pass falls back to two-mask passes.  The reference walks all of these shapes through the
same loop (rules.py:243-262, scm.py:262-286, cart.py:112-161); the checker finishes them in seconds on 20-40 k columns.
Every comparison is against ``oracle.OracleKmerRuleClassifications`` -- never one CUDA kernel against another -- and
bit-exact (array_equal on counts and index sets, == on float64)."""
import numpy as np
import pytest

from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.synthetic import synthetic_block, synthetic_labels, train_split
from oracle import kover_oracle as ko

pytestmark = pytest.mark.gpu

SHAPES = [(5000, 40000), (20000, 30000), (50000, 24000), (110000, 20000)]
IDS = ["W79", "W313", "W782", "W1719"]


class World(object):
    """One synthetic matrix, its checker and its example sets (label-sorted and interleaved)."""

    def __init__(self, n, k, seed):
        self.n, self.k, self.seed = n, k, seed
        self.P = synthetic_block(seed, n, 0, k)
        self.orc = ko.OracleKmerRuleClassifications(self.P, n, fast=True)
        self.train = train_split(seed, n)
        self.sets = {}
        for name, srt in (("sorted", True), ("interleaved", False)):
            y = synthetic_labels(seed, n, sorted_by_label=srt)
            self.sets[name] = (self.train[y[self.train] == 1], self.train[y[self.train] == 0])
        self._rcs = {}

    def rc(self, n_shards=1):
        if n_shards not in self._rcs:
            self._rcs[n_shards] = KmerRuleClassifications(self.P, self.n, devices=[0] * n_shards)
        return self._rcs[n_shards]

    def close(self):
        for rc in self._rcs.values():
            rc.close()
        self._rcs = {}

    def want_utility(self, p, pos, neg, blacklist, ubs):
        bl = np.zeros(2 * self.k, dtype=bool)
        bl[np.asarray(blacklist, dtype=np.int64)] = True
        return ko.merge_utility_blocks(len(neg) - self.orc.sum_rows(neg), len(pos) - self.orc.sum_rows(pos), p, bl,
                                       block_size=ubs)


@pytest.fixture(scope="module", params=SHAPES, ids=IDS)
def world(request):
    n, k = request.param
    w = World(n, k, seed=1000 + n % 977)
    yield w
    w.close()


def assert_utility(got, want):
    bu, bi, bp, bn = got
    wu, wi, wp, wn = want
    assert bu == wu
    assert np.array_equal(bi, np.asarray(wi, dtype=np.int64))
    assert np.array_equal(bp, wp) and np.array_equal(bn, wn)


@pytest.mark.parametrize("n_shards", [1, 3])
def test_sum_rows_and_columns(world, n_shards):
    rc = world.rc(n_shards)
    pos, neg = world.sets["interleaved"]
    for rows in (pos, neg, world.train, np.arange(world.n), pos[:1], np.zeros(0, dtype=np.int64)):
        got, want = rc.sum_rows(rows), world.orc.sum_rows(rows)
        assert got.dtype == want.dtype and np.array_equal(got, want)
    cols = [0, 777, world.k - 1, world.k + 5, 2 * world.k - 1]
    assert np.array_equal(rc.get_columns(cols), world.orc.get_columns(cols))
    assert np.array_equal(rc.get_columns(world.k + 123), world.orc.get_columns(world.k + 123))


@pytest.mark.parametrize("labels", ["sorted", "interleaved"])
@pytest.mark.parametrize("n_shards", [1, 3])
def test_best_utility_rules(world, labels, n_shards):
    rc = world.rc(n_shards)
    pos, neg = world.sets[labels]
    rs = np.random.RandomState(world.n)
    blacklist = np.unique(rs.randint(0, 2 * world.k, size=300))
    for p, ubs, bl in ((1.0, 1000000, []), (0.316, 7777, blacklist), (999999.0, 10007, blacklist), (10.0, 512, [])):
        got = rc.best_utility_rules(p, pos, neg, rule_blacklist=bl, util_block_size=ubs)
        assert_utility(got, world.want_utility(p, pos, neg, bl, ubs))
    # a later greedy iteration: few examples left, most packed rows inactive for one of the masks
    got = rc.best_utility_rules(2.0, pos[::37], neg[5::11], util_block_size=9999)
    assert_utility(got, world.want_utility(2.0, pos[::37], neg[5::11], [], 9999))


@pytest.mark.parametrize("n_shards", [1, 3])
def test_grid_six_mask_pairs(world, n_shards):
    """The folds of a cross-validation: six distinct mask pairs counted per multi-mask pass, every job (two p, both
    model types) scored from the resident counts -- each against its own checker scan."""
    rc = world.rc(n_shards)
    pos, neg = world.sets["interleaved"]
    rs = np.random.RandomState(7)
    fp, fn = rs.randint(0, 5, size=len(pos)), rs.randint(0, 5, size=len(neg))
    folds = [(pos, neg)] + [(pos[fp != f], neg[fn != f]) for f in range(5)]
    jobs = [(p, a, b) for a0, b0 in folds for p in (0.5, 999999.0) for a, b in ((a0, b0), (b0, a0))]
    got = rc.best_utility_rules_grid(jobs, util_block_size=7777)
    for (p, a, b), g in zip(jobs, got):
        assert_utility(g, world.want_utility(p, a, b, [], 7777))


@pytest.mark.parametrize("labels", ["sorted", "interleaved"])
@pytest.mark.parametrize("n_shards", [1, 3])
def test_gini_best_split(world, labels, n_shards):
    rc = world.rc(n_shards)
    pos, neg = world.sets[labels]
    ex = {0: neg, 1: pos}
    n_total, altered = ko.cart_altered_priors(ex, {0: 1.0, 1: 2.5})
    blacklist = np.unique(np.random.RandomState(3).randint(0, world.k, size=100))
    score = ko.gini_rule_score(world.orc, ex, altered, n_total)
    score[blacklist] = np.inf                                       # cart.py:231
    m, cand = rc.gini_best_split(ex, altered, n_total, rule_blacklist=blacklist)
    assert m == score.min() and np.array_equal(cand, np.where(score == m)[0])
    if n_shards == 1:
        assert np.array_equal(rc.gini_scores(), score)


@pytest.mark.parametrize("n_shards", [1, 3])
def test_gini_level_eleven_nodes(world, n_shards):
    """A tree level: eleven nodes with disjoint example sets (two multi-mask passes of up to 8 nodes), with and
    without the occurrence tiebreaker (experiment_cart.py:82-94), each node against the checker's score vector."""
    rc = world.rc(n_shards)
    pos, neg = world.sets["interleaved"]
    rs = np.random.RandomState(11)
    gp, gn = rs.randint(0, 11, size=len(pos)), rs.randint(0, 11, size=len(neg))
    nodes = [{0: neg[gn == i], 1: pos[gp == i]} for i in range(11)]
    nodes[4] = {0: neg[gn == 4][:3], 1: pos[gp == 4][:2]}           # a nearly pure, tiny node: millions of ties
    n_total, altered = ko.cart_altered_priors({0: neg, 1: pos}, {0: 1.0, 1: 1.0})
    scores = [ko.gini_rule_score(world.orc, ex, altered, n_total) for ex in nodes]
    for (m, cand), score in zip(rc.gini_best_splits(nodes, altered, n_total), scores):
        assert m == score.min() and np.array_equal(cand, np.where(score == m)[0])
    occ = world.orc.sum_rows(world.train)[:world.k]
    table = rc.occurrence_table(world.train)
    try:
        for (m, cand), score in zip(rc.gini_best_splits(nodes, altered, n_total, tie_tables=[table] * 11), scores):
            ties = np.where(score == score.min())[0]
            assert m == score.min() and np.array_equal(cand, ko.cart_tiebreaker(ties, occ))
    finally:
        table.release()


def test_kmer_risk_tables(world):
    pos, neg = world.sets["interleaved"]
    want = ko.kmer_risk_tables(world.orc, pos, neg)
    for n_shards in (1, 3):
        for w, g in zip(want, world.rc(n_shards).kmer_risk_tables(pos, neg)):
            assert w.dtype == g.dtype and np.array_equal(w, g)


def test_whole_tree_with_sibling_subtraction(world):
    """A whole Gini tree (depth 6, occurrence tiebreaker on the device, blacklist) grown level by level with the counts
    of every larger sibling derived as parent - smaller sibling (kv_gini_level_family) == the checker's tree, one node
    at a time with the host tiebreaker; 1 and 3 shards."""
    if world.n > 20000:
        pytest.skip("the checker's tree takes minutes at this size; W = 79 and 313 cover the path")
    from kover_b200.learning.common.rules import LazyKmerRuleList
    from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
    from kover_b200.learning.learners.cart import DecisionTreeClassifier
    pos, neg = world.sets["interleaved"]
    ex = {0: neg, 1: pos}
    train = np.sort(np.concatenate([pos, neg]))
    rs = np.random.RandomState(world.seed)
    blacklist = np.unique(rs.randint(0, world.k, size=500))
    occ = world.orc.sum_rows(train)
    imp = {0: 1.0, 1: 1.5}
    want = ko.cart_fit(world.orc, ex, "gini", 6, 2, imp, rule_blacklist=list(blacklist),
                       tiebreaker=lambda idx: ko.cart_tiebreaker(idx, occ))
    rules = LazyKmerRuleList(np.arange(world.k), np.arange(world.k))

    def same(a, b):
        assert (a.rule is None) == b.is_leaf
        assert [len(v) for v in a.class_examples_idx.values()] == [len(v) for v in b.class_examples_idx.values()]
        if a.rule is not None:
            assert int(a.rule_idx) == int(b.rule_idx) and float(a.criterion_value) == float(b.criterion_value)
            assert np.array_equal(np.asarray(a.equivalent_rules_idx), np.asarray(b.equivalent_rules_idx))
            same(a.left_child, b.left_child)
            same(a.right_child, b.right_child)

    for n_shards in (1, 3):
        rc = world.rc(n_shards)
        derived0 = [s_.info().level_nodes_derived for s_ in rc._group.shards]
        tb = OccurrenceTiebreaker(rc, train)
        try:
            t = DecisionTreeClassifier("gini", 6, 2, imp)
            t.fit(rules, rc, ex, rule_blacklist=list(blacklist), tiebreaker=tb)
        finally:
            tb.release()
        same(t.decision_tree, want)
        assert all(s_.info().level_nodes_derived - d0 >= 5 for s_, d0 in zip(rc._group.shards, derived0))


# ------------------------------------------------------------------------------------------------ beyond the staging limits
def _check_all_paths(n, k, seed, rc, orc):
    y = synthetic_labels(seed, n, sorted_by_label=False)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    got, want = rc.sum_rows(train), orc.sum_rows(train)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    for p, ubs in ((1.0, 1000000), (999999.0, 3001)):
        want = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p,
                                       np.zeros(2 * k, dtype=bool), block_size=ubs)
        assert_utility(rc.best_utility_rules(p, pos, neg, util_block_size=ubs), want)
    ex = {0: neg, 1: pos}
    n_total, altered = ko.cart_altered_priors(ex, {0: 1.0, 1: 1.0})
    score = ko.gini_rule_score(orc, ex, altered, n_total)
    m, cand = rc.gini_best_split(ex, altered, n_total)
    assert m == score.min() and np.array_equal(cand, np.where(score == m)[0])
    # lock-step grid over two mask pairs and a three-node level: the multi-mask pass (or its two-mask fallback)
    jobs = [(p, a, b) for a, b in ((pos, neg), (pos[::2], neg[1::2])) for p in (0.5, 2.0)]
    for (p, a, b), g in zip(jobs, rc.best_utility_rules_grid(jobs, util_block_size=3001)):
        want = ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), p,
                                       np.zeros(2 * k, dtype=bool), block_size=3001)
        assert_utility(g, want)
    nodes = [{0: neg[i::3], 1: pos[i::3]} for i in range(3)]
    for node, (m, cand) in zip(nodes, rc.gini_best_splits(nodes, altered, n_total)):
        score = ko.gini_rule_score(orc, node, altered, n_total)
        assert m == score.min() and np.array_equal(cand, np.where(score == m)[0])


def test_forced_global_record_variant(monkeypatch):
    """The global-record scan variant (row records read from device memory instead of shared memory), forced at
    config 3's row count so that it is exercised on a matrix the other tests also cover."""
    monkeypatch.setenv("KOVER_B200_SCAN_VARIANT", "6")
    n, k, seed = 5000, 30000, 31
    P = synthetic_block(seed, n, 0, k)
    rc = KmerRuleClassifications(P, n, devices=[0])
    try:
        _check_all_paths(n, k, seed, rc, ko.OracleKmerRuleClassifications(P, n, fast=True))
    finally:
        rc.close()


def test_more_rows_than_shared_memory_can_stage():
    """330 000 genomes = 5 157 packed rows: 20 B of row record per active row no longer fit beside the TMA ring
    (the limit ADVICE / VERDICT flagged as KV_E_INVALID in round 1); the scan reads the records from global memory.
    The reference has no limit (rules.py:243-262)."""
    n, k, seed = 330000, 8192, 32
    P = synthetic_block(seed, n, 0, k)
    rc = KmerRuleClassifications(P, n, devices=[0])
    try:
        _check_all_paths(n, k, seed, rc, ko.OracleKmerRuleClassifications(P, n, fast=True))
    finally:
        rc.close()


def test_more_rows_than_the_multi_mask_pass_can_stage():
    """520 000 genomes = 8 125 packed rows: the [row][mask] staging of k_count_tma does not fit either; grids and
    tree levels fall back to two masks per pass through the two-mask scan."""
    n, k, seed = 520000, 4096, 33
    P = synthetic_block(seed, n, 0, k)
    rc = KmerRuleClassifications(P, n, devices=[0])
    try:
        _check_all_paths(n, k, seed, rc, ko.OracleKmerRuleClassifications(P, n, fast=True))
    finally:
        rc.close()


def test_forced_variant_row_limit_is_an_error(monkeypatch):
    """A scan variant forced through KOVER_B200_SCAN_VARIANT keeps its shared-memory row limit and says so
    (KV_E_INVALID -> ValueError) instead of launching with an oversized ring."""
    monkeypatch.setenv("KOVER_B200_SCAN_VARIANT", "1")               # 192 KB ring: room for ~1 690 staged rows
    n, k = 110000, 2048
    rc = KmerRuleClassifications(synthetic_block(5, n, 0, k), n, devices=[0])
    try:
        with pytest.raises(ValueError, match="too many active packed rows"):
            rc.sum_rows(np.arange(n))
        assert rc.sum_rows(np.arange(64 * 100)).shape == (2 * k,)   # 100 active rows still fit
    finally:
        rc.close()
