"""GPU: the CUDA path (through the C ABI, via the reference-shaped Python classes) against
  (1) the golden vectors minted from the REAL reference (tests/golden/*.npz),
  (2) the CPU checker (oracle/) on seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's config-2 size (2k genomes x 10M k-mers).
Everything is bit-exact: integer counts and index sets with array_equal, float64 utilities / Gini
scores with ==."""
import json
import os

import numpy as np
import pytest

import cases
from kover_b200 import _native
from kover_b200.learning.common import popcount as kpop
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.learners import cart as kcart
from kover_b200.learning.learners import scm as kscm
from kover_b200.synthetic import synthetic_block, synthetic_labels, train_split
from oracle import kover_oracle as ko

pytestmark = pytest.mark.gpu


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


class Chunked(np.ndarray):
    chunks = None


def as_dataset(P, chunks):
    d = P.view(Chunked)
    d.chunks = chunks
    return d


def rule_list(k):
    return LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))


# ---------------------------------------------------------------------------------- a1
def test_popcount_kernel_golden(golden_dir):
    g = load(golden_dir, "popcount")
    a = g["a64"].copy()
    kpop.inplace_popcount_64(a, g["m64"])
    assert np.array_equal(a, g["o64"])
    b = g["a32"].copy()
    kpop.inplace_popcount_32(b, g["m32"])
    assert np.array_equal(b, g["o32"])
    with pytest.raises(ValueError):
        kpop.inplace_popcount_64(g["a32"].copy(), g["m32"])


# ---------------------------------------------------------------------------------- a3 / a4 / a5
@pytest.mark.parametrize("mname", ["M1", "M2", "M3", "M4"])
@pytest.mark.parametrize("devices", [[0], [0, 0, 0]], ids=["1shard", "3shards"])
def test_sum_rows_get_columns_golden(golden_dir, mname, devices):
    g = load(golden_dir, "rules_" + mname)
    X, P, n, chunks = cases.matrix(mname)
    k = P.shape[1]
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
    assert list(rc.shape) == g["shape"].tolist()
    for rname, rows in cases.row_sets(n).items():
        got = rc.sum_rows(rows)
        want = g["sum_" + rname]
        assert got.dtype == want.dtype, rname
        assert np.array_equal(got, want), rname
    for cname, cols in cases.column_sets(k).items():
        got = rc.get_columns(cols)
        want = g["col_" + cname]
        assert got.shape == want.shape and got.dtype == want.dtype and np.array_equal(got, want), cname
    if mname == "M3":
        rc32 = KmerRuleClassifications(as_dataset(cases.pack32_from64(P), chunks), n, devices=devices)
        for rname, rows in cases.row_sets(n).items():
            assert np.array_equal(rc32.sum_rows(rows), g["sum32_" + rname]), rname
        assert np.array_equal(rc32.get_columns(cases.column_sets(k)["mixed"]), g["col32_mixed"])
    if mname == "M4":
        rc2 = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
        rc2.remove_rows([3, 10])
        rc2.remove_rows([0])
        assert list(rc2.shape) == g["rm_shape"].tolist()
        assert np.array_equal(rc2.sum_rows(np.array([0, 1, 2, 9, 40, 61])), g["rm_sum"])
        assert np.array_equal(rc2.get_columns([1, k + 2]), g["rm_col"])
        with pytest.raises(TypeError):
            rc2.get_columns(np.array([1, 2]))


def test_sum_rows_multi_matches_single():
    X, P, n, chunks = cases.matrix("M2")
    rc = KmerRuleClassifications(P, n, devices=[0, 0])
    sets = list(cases.row_sets(n).values()) + [np.arange(0, n, 3), np.arange(1, n, 7), np.arange(n - 70, n)] * 3
    multi = rc.sum_rows_multi(sets)        # 17 masks: two passes of k_count_multi
    for i, rows in enumerate(sets):
        assert np.array_equal(multi[i], rc.sum_rows(rows)[:P.shape[1]].astype(np.uint32)), i


@pytest.mark.parametrize("n,n_masks", [(260, 5), (16000, 16), (51200, 16), (51200, 7)],
                         ids=["W5", "W250_ring128K", "W800_ring64K", "W800_7masks"])
def test_multi_mask_pass_shapes(n, n_masks):
    """k_count_tma over the shapes that select its three ring depths (the mask staging grows with packed rows x
    masks), with empty masks, masks that touch a single packed row, row groups that are not a multiple of four and
    masks that leave most rows untouched -- against one single-mask pass per mask."""
    seed, k = 31, 2100
    P = cases.random_words(seed, n, k)
    rs = np.random.RandomState(seed)
    sets = []
    for m in range(n_masks):
        kind = m % 5
        if kind == 0:
            sets.append(np.sort(rs.choice(n, size=n // 3, replace=False)))
        elif kind == 1:
            sets.append(np.arange(64 * (m % ((n + 63) // 64)), min(n, 64 * (m % ((n + 63) // 64)) + 5)))    # one packed row
        elif kind == 2:
            sets.append(np.zeros(0, dtype=np.int64))                                                    # empty
        elif kind == 3:
            lo = (n // 7) * (m % 7)
            sets.append(np.arange(lo, min(n, lo + 64 * 3 + 1)))                                           # 3-4 rows
        else:
            sets.append(np.arange(n))                                                                     # everything
    rc = KmerRuleClassifications(P, n, devices=[0])
    got = rc.sum_rows_multi(sets)
    for i, rows in enumerate(sets):
        want = rc.sum_rows(rows)[:k].astype(np.uint32) if len(rows) else np.zeros(k, dtype=np.uint32)
        assert np.array_equal(got[i], want), i
    rc.close()


# ---------------------------------------------------------------------------------- a6
@pytest.mark.parametrize("case", cases.UTILITY_CASES, ids=[c[0] for c in cases.UTILITY_CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_best_utility_rules_golden(golden_dir, case, devices):
    g = load(golden_dir, case[0])
    P, n, chunks, pos, neg, p, blacklist, ubs = cases.utility_inputs(case)
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
    bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, rule_blacklist=blacklist, util_block_size=ubs)
    assert bu == float(g["best_utility"])
    assert np.array_equal(bi, g["best_idx"])
    assert np.array_equal(bp, g["best_pos_err"])
    assert np.array_equal(bn, g["best_neg_cover"])


@pytest.mark.parametrize("case", cases.UTILITY_CASES, ids=[c[0] for c in cases.UTILITY_CASES])
def test_one_exchange_packets_golden(golden_dir, case):
    """The multi-rank protocol (kv_scm_scan_packet + sharding.merge_packets) with three shards on one GPU
    standing in for three ranks: identical to the reference's result whenever no packet overflows."""
    from kover_b200.sharding import merge_packets
    g = load(golden_dir, case[0])
    P, n, chunks, pos, neg, p, blacklist, ubs = cases.utility_inputs(case)
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=[0, 0, 0])
    rc.set_rule_blacklist(blacklist)
    nb = _native.n_utility_blocks(P.shape[1], ubs)
    for cap in (4096, 3):
        rows = np.zeros((3, _native.packet_bytes(nb, cap)), dtype=np.uint8)
        for i, shard in enumerate(rc._group.shards):
            shard.scm_scan_packet(rc.row_mask(pos), rc.row_mask(neg), len(pos), len(neg), p, ubs, cap, rows[i])
        out, bm, best, sel = merge_packets(rows, nb, cap, p, ubs)
        assert best == float(g["best_utility"])
        if out is None:                      # overflow: the caller falls back to kv_scm_ties on the merged tables
            assert cap == 3
            parts = [shard.scm_ties(bm, sel) for shard in rc._group.shards]
            idx = np.sort(np.concatenate([q[0] for q in parts]))
            assert np.array_equal(idx, g["best_idx"])
        else:
            assert np.array_equal(out[1], g["best_idx"]) and np.array_equal(out[2], g["best_pos_err"])
            assert np.array_equal(out[3], g["best_neg_cover"])


def test_best_utility_no_positives_left():
    # scm.py:250-253: with no positive example left the error counts are all zero
    X, P, n, chunks = cases.matrix("M1")
    rc = KmerRuleClassifications(P, n, devices=[0])
    neg = np.arange(0, n, 2)
    bu, bi, bp, bn = rc.best_utility_rules(1.0, np.array([], dtype=np.int64), neg, util_block_size=1000)
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), np.zeros(2 * P.shape[1], dtype=np.uint8),
                                             1.0, np.zeros(2 * P.shape[1], dtype=bool), block_size=1000)
    assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64)) and not bp.any()
    assert np.array_equal(bn, wn)


def test_tie_buffer_overflow_regrows():
    # every rule of an all-zero matrix ties: 2K ties must come back complete and ascending
    n, k = 70, 3000
    P = np.zeros((2, k), dtype=np.uint64)
    rc = KmerRuleClassifications(P, n, devices=[0])
    pos, neg = np.arange(0, 30), np.arange(30, 70)
    shard = rc._group.shards[0]
    bu, bi, bp, bn = shard.scm_best_rules(rc.row_mask(pos), rc.row_mask(neg), 30, 40, 1.0, 1000, capacity=16)
    assert bu == 10.0 and np.array_equal(bi, np.arange(k))      # presence rules: 40 covered - 1.0 * 30 errors
    assert bp.tolist() == [30] * k and bn.tolist() == [40] * k


@pytest.mark.parametrize("seed,n,k,p,ubs", [(1, 1000, 70000, 1.0, 1000000), (2, 2500, 30011, 0.316, 9973),
                                            (3, 129, 100001, 999999.0, 50000), (4, 640, 51200, 2.0, 512)])
@pytest.mark.parametrize("sorted_labels", [True, False])
def test_best_utility_rules_vs_checker(seed, n, k, p, ubs, sorted_labels):
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=sorted_labels)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    blacklist = np.unique(np.random.RandomState(seed).randint(0, 2 * k, size=50))
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    bl = np.zeros(2 * k, dtype=bool)
    bl[blacklist] = True
    wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p, bl,
                                             block_size=ubs)
    for devices in ([0], [0, 0, 0]):
        rc = KmerRuleClassifications(P, n, devices=devices)
        bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, rule_blacklist=blacklist, util_block_size=ubs)
        assert bu == wu
        assert np.array_equal(bi, np.asarray(wi, dtype=np.int64))
        assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
        rc.close()


@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_util_block_size_changes_between_calls(devices):
    """Another util_block_size re-lays the block-maximum accumulators over what used to be block maxima / tolerances
    of the previous call (round-2 regression: a stale raw double decoded as a small negative "maximum" and beat the
    real, hugely negative maxima of a p = 999999 scan).  Grow, shrink and grow the block count on one shard."""
    n, k, seed = 700, 20000, 9
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    nc, pe = len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos)
    rc = KmerRuleClassifications(P, n, devices=devices)
    for p, ubs in ((1.0, 1000000), (999999.0, 3001), (0.5, 40000), (999999.0, 777), (999999.0, 1000000), (2.0, 100)):
        wu, wi, wp, wn = ko.merge_utility_blocks(nc, pe, p, np.zeros(2 * k, dtype=bool), block_size=ubs)
        bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, util_block_size=ubs)
        assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64)), (p, ubs)
        assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
        jobs = [(q, a, b) for a, b in ((pos, neg), (pos[::2], neg[1::2])) for q in (p, 0.25)]
        for (q, a, b), got in zip(jobs, rc.best_utility_rules_grid(jobs, util_block_size=ubs)):
            wu, wi, wp, wn = ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), q,
                                                     np.zeros(2 * k, dtype=bool), block_size=ubs)
            assert got[0] == wu and np.array_equal(got[1], np.asarray(wi, dtype=np.int64)), (q, ubs)
    rc.close()


@pytest.mark.parametrize("fused,coop,stream", [("1", "1", "1"), ("1", "1", "0"), ("1", "0", "0"), ("0", "1", "0")],
                         ids=["fused-stream", "fused-coop", "fused-plain", "unfused"])
def test_single_launch_step_modes(monkeypatch, fused, coop, stream):
    """The single-shard step as ONE kernel (scan -> grid barrier -> block merge -> ties into mapped host memory,
    cooperative or plain launch) and as the scan + select + ties sequence: both against the checker, over repeated
    calls (the monotonic barrier / epoch counters and the parity-alternating accumulators), several block counts,
    few and thousands of ties, blacklists, through the mask and the index-array entry points."""
    monkeypatch.setenv("KOVER_B200_FUSED", fused)
    monkeypatch.setenv("KOVER_B200_COOP", coop)
    monkeypatch.setenv("KOVER_B200_STREAM", stream)
    n, k, seed = 900, 200000, 12
    P = synthetic_block(seed, n, 0, k)
    P[:, 150000:150050] = P[:, 7:8]                    # 50 copies of one k-mer: equivalent rules
    y = synthetic_labels(seed, n, sorted_by_label=False)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    rc = KmerRuleClassifications(P, n, devices=[0])
    shard = rc._group.shards[0]
    rs = np.random.RandomState(4)
    for it in range(40):
        a = pos[rs.rand(len(pos)) < rs.choice([1.0, 0.5, 0.05])]
        b = neg[rs.rand(len(neg)) < rs.choice([1.0, 0.5, 0.05])]
        p = float(rs.choice([0.1, 1.0, 2.0, 999999.0]))
        ubs = int(rs.choice([1000000, 33333, 4099, 150]))
        bl = np.unique(rs.randint(0, 2 * k, size=int(rs.choice([0, 60]))))
        blm = np.zeros(2 * k, dtype=bool)
        blm[bl] = True
        want = ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), p, blm, block_size=ubs)
        got = rc.best_utility_rules(p, a, b, rule_blacklist=bl, util_block_size=ubs)
        assert got[0] == want[0] and np.array_equal(got[1], np.asarray(want[1], dtype=np.int64)), (it, p, ubs)
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
        if it % 5 == 0:                                # the mask entry point of the C ABI
            got = shard.scm_best_rules(rc.row_mask(a), rc.row_mask(b), len(a), len(b), p, ubs)
            assert got[0] == want[0] and np.array_equal(got[1], np.asarray(want[1], dtype=np.int64))
    with pytest.raises(IndexError):
        rc.best_utility_rules(1.0, np.array([0, n]), neg)
    with pytest.raises(IndexError):
        rc.best_utility_rules(1.0, pos, np.array([-1]))
    info = shard.info()
    if stream == "1":
        assert info.stream_steps > 0 and info.stream_redos < info.stream_steps
    else:
        assert info.stream_steps == 0
    rc.close()


def test_candidate_streaming_step():
    """The step that writes no count arrays (k_scan_tma<EPI_SCM_STREAM>: candidates streamed against the running
    maximum, filtered by the grid's last CTA) against the checker where it has to give up and redo the step with the
    counting kernel -- every rule blacklisted (best utility -inf), one CTA's candidate list overflowing (tens of
    thousands of identical columns), more ties than the result buffer -- and where later calls want the counts it
    did not write: kv_scm_reuse_counts (another p, exchanged roles) and kv_scm_ties produce them on demand."""
    n, k, seed = 700, 120000, 5
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    rc = KmerRuleClassifications(P, n, devices=[0])
    shard = rc._group.shards[0]
    none = np.zeros(2 * k, dtype=bool)

    def want(p, a, b, blm=none, ubs=25000):
        return ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), p, blm, block_size=ubs)

    def same(got, w):
        assert got[0] == w[0] and np.array_equal(got[1], np.asarray(w[1], dtype=np.int64))
        assert np.array_equal(got[2], w[2]) and np.array_equal(got[3], w[3])

    # plain streamed steps
    for p in (0.1, 1.0, 999999.0):
        same(rc.best_utility_rules(p, pos, neg, util_block_size=25000), want(p, pos, neg))
    i0 = shard.info()
    assert i0.stream_steps == 3 and i0.stream_redos == 0
    # the counts of a streamed step on demand: another p from the "resident" counts, then with exchanged roles
    mp, mn = rc.row_mask(pos), rc.row_mask(neg)
    shard.scm_best_rules(mp, mn, len(pos), len(neg), 1.0, 25000)
    shard.scm_reuse_counts(False)
    same(shard.scm_best_rules(mp, mn, len(pos), len(neg), 2.5, 25000), want(2.5, pos, neg))
    shard.scm_reuse_counts(True)
    same(shard.scm_best_rules(mn, mp, len(neg), len(pos), 0.5, 9999), want(0.5, neg, pos, ubs=9999))
    assert shard.info().rescore_calls == i0.rescore_calls + 2
    # ... and the two-phase tie collection on the block table of the host
    w = want(1.0, pos, neg)
    shard.scm_best_rules(mp, mn, len(pos), len(neg), 1.0, 25000)
    cnt_neg, cnt_pos = orc.sum_rows(neg), orc.sum_rows(pos)
    u = np.concatenate([(len(neg) - cnt_neg) - 1.0 * (len(pos) - cnt_pos), cnt_neg - 1.0 * cnt_pos])
    nb = (2 * k + 24999) // 25000
    bm = np.array([u[b * 25000:(b + 1) * 25000].max() for b in range(nb)])
    best, sel = _native.select_blocks(bm)
    idx, pe, nc = shard.scm_ties(bm, sel)
    assert best == w[0] and np.array_equal(idx, np.asarray(w[1], dtype=np.int64)) and np.array_equal(pe, w[2])
    # every rule blacklisted: best utility -inf, every rule of the selected blocks ties
    everything = np.arange(2 * k)
    blm = np.ones(2 * k, dtype=bool)
    r0 = shard.info().stream_redos
    got = rc.best_utility_rules(1.0, pos, neg, rule_blacklist=everything, util_block_size=25000)
    w = want(1.0, pos, neg, blm)
    assert got[0] == w[0] == -np.inf and np.array_equal(got[1], np.asarray(w[1], dtype=np.int64))
    assert shard.info().stream_redos == r0 + 1
    same(rc.best_utility_rules(1.0, pos, neg, util_block_size=25000), want(1.0, pos, neg))   # (blacklist cleared)
    rc.close()

    # one k-mer copied 30 000 times next to each other: the candidate lists of the CTAs that hold the copies overflow
    # (4 tiles = 8 192 rules per CTA against 2 048 candidate slots)
    n, k = 300, 600000
    P2 = synthetic_block(seed, n, 0, k)
    P2[:, 200000:230000] = P2[:, 11:12]
    y = synthetic_labels(seed, n, sorted_by_label=False)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    orc = ko.OracleKmerRuleClassifications(P2, n, fast=True)
    rc = KmerRuleClassifications(P2, n, devices=[0])
    shard = rc._group.shards[0]
    none = np.zeros(2 * k, dtype=bool)
    col = orc.get_columns(11).astype(bool)
    a, b = pos[~col[pos]][:40], neg[col[neg]]          # the copied k-mer's absence rule is far ahead: 30 001 ties
    w = want(1.0, a, b, none, 1000000)
    assert len(w[1]) >= 30001
    for it in range(12):                               # first call: redo; then the back-off window; then a new attempt
        same(rc.best_utility_rules(1.0, a, b), w)
    info = shard.info()
    assert info.stream_redos == 2 and info.stream_steps == 2, (info.stream_steps, info.stream_redos)
    same(rc.best_utility_rules(1.0, pos, neg), want(1.0, pos, neg, none, 1000000))
    rc.close()


@pytest.mark.parametrize("n,k", [(700, 30000), (5000, 12000), (20000, 6000)], ids=["W11", "W79", "W313"])
def test_device_resident_greedy_state(n, k):
    """g1 / north-star (3): the remaining-example masks live on the device; kv_mask_and_column ANDs them with the picked
    rule's column (presence: keep the genomes WITH the k-mer; absence: without) and rebuilds the scan's row records.
    Against the reference's host bookkeeping (scm.py:133-139: get_columns + index filtering) and the checker's scan
    after every pick, on a random walk of presence and absence rules."""
    seed = n % 97
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    rc = KmerRuleClassifications(P, n, devices=[0])
    state = rc.scm_state(pos, neg)
    assert state is not None and (state.n_pos, state.n_neg) == (len(pos), len(neg))
    rs = np.random.RandomState(seed)
    for it in range(8):
        p = float(rs.choice([0.1, 1.0, 999999.0]))
        want = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p,
                                       np.zeros(2 * k, dtype=bool), block_size=7777)
        got = state.scan(p, [], 7777)
        assert got[0] == want[0] and np.array_equal(got[1], np.asarray(want[1], dtype=np.int64)), it
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
        # pick a rule of middling density so that examples remain: presence and absence alternate
        rule = int(rs.randint(0, k)) + (k if it % 2 else 0)
        accepted = orc.get_columns(rule)
        pos, neg = pos[accepted[pos] != 0], neg[accepted[neg] != 0]
        assert state.apply(rule) == (len(pos), len(neg)), (it, rule)
        if len(pos) == 0 or len(neg) == 0:
            break
    lp, ln = state.index_lists()
    assert np.array_equal(lp, pos) and np.array_equal(ln, neg)
    rc.close()


def test_large_equivalent_rule_sets():
    """Tens of thousands of equivalent rules (duplicated k-mers, as on real genomes): the device-sorted path of
    both tie collectors, single- and multi-shard, against the checker."""
    rs = np.random.RandomState(77)
    n, k = 150, 90000
    base = (rs.rand(n, 12) < 0.5).astype(np.uint8)
    X = base[:, rs.randint(0, 12, size=k)]
    P = cases.pack64(X)
    y = (base[:, 3] & (1 - base[:, 7])).astype(np.uint8)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), 1.0,
                                             np.zeros(2 * k, dtype=bool), block_size=25000)
    ex = {0: neg, 1: pos}
    n_total, altered = ko.cart_altered_priors(ex, {0: 1.0, 1: 1.0})
    score = ko.gini_rule_score(orc, ex, altered, n_total)
    assert len(wi) > 5000 and (score == score.min()).sum() > 5000
    for devices in ([0], [0, 0, 0]):
        rc = KmerRuleClassifications(P, n, devices=devices)
        bu, bi, bp, bn = rc.best_utility_rules(1.0, pos, neg, util_block_size=25000)
        assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64))
        assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
        m, cand = rc.gini_best_split(ex, altered, n_total)
        assert m == score.min() and np.array_equal(cand, np.where(score == m)[0])
        rc.close()


@pytest.mark.parametrize("seed", range(48))
def test_randomized_small_shapes(seed):
    """Ragged / tiny / boundary shapes: 1..260 genomes, 1..5000 k-mers, random utility-block sizes, random
    blacklists (sometimes whole blocks), random p, 1-3 shards -- SCM scan, sum_rows, get_columns and the Gini
    scan against the CPU checker, bit for bit."""
    rs = np.random.RandomState(1000 + seed)
    n = int(rs.choice([1, 2, 63, 64, 65, 127, 128, 129, 200, 260]))
    k = int(rs.choice([1, 2, 31, 63, 64, 65, 511, 512, 513, 1023, 1024, 1025, 2047, 3000, 5000]))
    dens = rs.choice([0.0, 0.02, 0.3, 0.5, 0.9, 1.0], size=k)
    X = (rs.rand(n, k) < dens[None, :]).astype(np.uint8)
    P = cases.pack64(X)
    y = (rs.rand(n) < rs.choice([0.2, 0.5, 0.8])).astype(np.uint8)
    if rs.rand() < 0.5:
        y = np.sort(y)
    keep = rs.rand(n) < 0.8
    pos, neg = np.where((y == 1) & keep)[0], np.where((y == 0) & keep)[0]
    ubs = int(rs.choice([1, 7, 64, 100, 1000, 1000000]))
    p = float(rs.choice([0.0, 0.1, 0.316, 1.0, 1.778, 10.0, 999999.0]))
    blacklist = np.unique(rs.randint(0, 2 * k, size=rs.randint(0, max(1, k // 3))))
    if rs.rand() < 0.3 and 2 * k > ubs:
        blacklist = np.unique(np.concatenate([blacklist, np.arange(0, min(ubs, 2 * k - 1))]))   # a fully blacklisted head block
    if len(blacklist) == 2 * k:
        blacklist = blacklist[:-1]
    devices = [0] * int(rs.choice([1, 2, 3]))
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    rc = KmerRuleClassifications(P, n, devices=devices)
    assert np.array_equal(rc.sum_rows(pos), orc.sum_rows(pos))
    bl = np.zeros(2 * k, dtype=bool)
    bl[blacklist] = True
    pe = len(pos) - orc.sum_rows(pos) if len(pos) else np.zeros(2 * k, dtype=np.uint8)
    wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), pe, p, bl, block_size=ubs)
    bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, rule_blacklist=blacklist, util_block_size=ubs)
    assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64))
    assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
    cols = sorted(set(int(c) for c in rs.randint(0, 2 * k, size=3)))
    assert np.array_equal(rc.get_columns(cols), orc.get_columns(cols))
    if len(pos) and len(neg):
        ex = {0: neg, 1: pos}
        n_total, altered = ko.cart_altered_priors({0: np.where(y == 0)[0], 1: np.where(y == 1)[0]}, {0: 1.0, 1: 1.7})
        want = ko.gini_rule_score(orc, ex, altered, n_total)
        kb = blacklist[blacklist < k]
        want[kb] = np.inf
        m, cand = rc.gini_best_split(ex, altered, n_total, rule_blacklist=kb)
        assert m == want.min()
        if m != np.inf:
            assert np.array_equal(cand, np.where(want == m)[0])
        else:
            assert len(cand) == 0
    rc.close()


# ---------------------------------------------------------------------------------- a7 / a8
@pytest.mark.parametrize("case", cases.SCM_CASES, ids=[c[0] for c in cases.SCM_CASES])
def test_scm_fit_golden(golden_dir, case, monkeypatch):
    name, mname, lname, model_type, p, max_rules, tb, blacklist, ubs = case
    g = load(golden_dir, name)
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    k = P.shape[1]
    monkeypatch.setattr(kscm, "UTIL_BLOCK_SIZE", ubs)
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=[0, 0])
    risks = cases.rule_risks(2 * k)

    def tiebreaker(idx):
        if tb == "identity":
            return idx
        return ko.scm_tiebreaker(idx, risks, model_type)

    infos = []
    m = kscm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules)
    m.fit(rule_list(k), rc, np.where(y == 1)[0], np.where(y == 0)[0], rule_blacklist=blacklist,
          tiebreaker=tiebreaker, iteration_callback=lambda i: infos.append(dict(i)), iteration_rule_importances=True)
    assert len(infos) == int(g["n_iterations"])
    assert str(m.model) == str(g["model_str"])
    for t, info in enumerate(infos):
        assert info["utility_max"] == float(g["it%d_utility_max" % t])
        assert np.array_equal(np.asarray(info["utility_argmax"], dtype=np.int64), g["it%d_utility_argmax" % t])
        assert np.array_equal(np.asarray(info["equivalent_rules_idx"], dtype=np.int64), g["it%d_equivalent" % t])
        assert str(info["selected_rule"]) == str(g["it%d_selected" % t])
        assert np.array_equal(info["rule_importances"], g["it%d_importances" % t])
    assert np.array_equal(np.asarray(m.rule_importances, dtype=np.float64), g["rule_importances"])
    # predictions of the learnt model on the unpacked matrix (models.py:155-182)
    if X is not None:
        pred = m.predict(X)
        cls = np.stack([r.classify(X) for r in m.model.rules], axis=1)
        want = cls.all(axis=1) if model_type == "conjunction" else cls.any(axis=1)
        assert np.array_equal(pred, want.astype(np.uint8))


def check_scm_against_golden(g, m, infos):
    assert len(infos) == int(g["n_iterations"])
    assert str(m.model) == str(g["model_str"])
    for t, info in enumerate(infos):
        assert info["utility_max"] == float(g["it%d_utility_max" % t])
        assert np.array_equal(np.asarray(info["utility_argmax"], dtype=np.int64), g["it%d_utility_argmax" % t])
        assert np.array_equal(np.asarray(info["equivalent_rules_idx"], dtype=np.int64), g["it%d_equivalent" % t])
        assert str(info["selected_rule"]) == str(g["it%d_selected" % t])
        assert np.array_equal(info["rule_importances"], g["it%d_importances" % t])
    assert np.array_equal(np.asarray(m.rule_importances, dtype=np.float64), g["rule_importances"])


@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_fit_many_lockstep_golden(golden_dir, devices, monkeypatch):
    """north-star item 4: all fits over one matrix in lock-step, scans over identical example sets sharing a
    matrix pass (kv_scm_reuse_counts).  Every learner must end exactly where its own fit() ends."""
    groups = {}
    for case in cases.SCM_CASES:
        groups.setdefault((case[1], case[8], tuple(case[7])), []).append(case)
    for (mname, ubs, blacklist), members in groups.items():
        X, P, n, chunks = cases.matrix(mname)
        k = P.shape[1]
        monkeypatch.setattr(kscm, "UTIL_BLOCK_SIZE", ubs)
        rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
        risks = cases.rule_risks(2 * k)
        learners, args, infos = [], [], []
        for name, _, lname, model_type, p, max_rules, tb, _, _ in members:
            y = cases.labels(lname)
            tiebreaker = (lambda idx: idx) if tb == "identity" else \
                (lambda idx, mt=model_type: ko.scm_tiebreaker(idx, risks, mt))
            log = []
            infos.append(log)
            learners.append(kscm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules))
            args.append(dict(rules=rule_list(k), positive_example_idx=np.where(y == 1)[0],
                             negative_example_idx=np.where(y == 0)[0], rule_blacklist=list(blacklist),
                             tiebreaker=tiebreaker, iteration_callback=lambda i, log=log: log.append(dict(i)),
                             iteration_rule_importances=True))
        kscm.fit_many(learners, args, rc)
        for case, m, log in zip(members, learners, infos):
            check_scm_against_golden(load(golden_dir, case[0]), m, log)
        rc.close()


def test_grid_shares_matrix_passes():
    """12 jobs = 6 p values x {conjunction, disjunction (labels swapped)} over the same examples: ONE matrix
    scan, 11 re-scores of the resident counts, results identical to 12 independent scans."""
    seed, n, k = 9, 1500, 120000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    rc = KmerRuleClassifications(P, n, devices=[0])
    shard = rc._group.shards[0]
    jobs = [(p, a, b) for p in (0.1, 0.5, 1.0, 2.0, 5.0, 10.0) for a, b in ((pos, neg), (neg, pos))]
    want = [rc.best_utility_rules(p, a, b, util_block_size=50000) for p, a, b in jobs]
    before = shard.info()
    got = rc.best_utility_rules_grid(jobs, util_block_size=50000)
    after = shard.info()
    assert after.scan_calls - before.scan_calls == 1 and after.rescore_calls - before.rescore_calls == 11
    for w, g_ in zip(want, got):
        assert w[0] == g_[0] and all(np.array_equal(a, b) for a, b in zip(w[1:], g_[1:]))
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    for (p, a, b), g_ in zip(jobs, got):
        wu, wi, wp, wn = ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), p,
                                                 np.zeros(2 * k, dtype=bool), block_size=50000)
        assert g_[0] == wu and np.array_equal(g_[1], np.asarray(wi, dtype=np.int64))
        assert np.array_equal(g_[2], wp) and np.array_equal(g_[3], wn)


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]], ids=["1shard", "3shards"])
@pytest.mark.parametrize("sorted_labels", [True, False], ids=["sorted", "interleaved"])
@pytest.mark.parametrize("n_folds", [2, 5, 11])
def test_grid_batches_distinct_mask_pairs(sorted_labels, n_folds, devices):
    """The fits of a cross-validation in lock-step: every fold is a DIFFERENT mask pair.  They are counted several
    pairs per matrix pass (k_count_tma) and scored from the resident counts; results identical to independent
    scans, ceil(pairs / 8) matrix passes instead of one per pair."""
    seed, n, k = 11, 1300, 70000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=sorted_labels)
    train = train_split(seed, n)
    rs = np.random.RandomState(seed)
    fold_of = rs.randint(0, n_folds, size=len(train))
    rc = KmerRuleClassifications(P, n, devices=devices)
    jobs = []
    for f in range(n_folds + 1):
        tr = train if f == n_folds else train[fold_of != f]
        if f == 1:
            tr = tr[tr >= 64 * 3]          # a fold whose masks leave whole packed rows empty
        pos, neg = tr[y[tr] == 1], tr[y[tr] == 0]
        for p_ in (0.5, 2.0):
            jobs.append((p_, pos, neg))
            jobs.append((p_, neg, pos))
    bl = rs.choice(2 * k, size=500, replace=False)
    want = [rc.best_utility_rules(p_, a, b, rule_blacklist=bl, util_block_size=30000) for p_, a, b in jobs]
    before = [s_.info() for s_ in rc._group.shards]
    got = rc.best_utility_rules_grid(jobs, rule_blacklist=bl, util_block_size=30000)
    after = [s_.info() for s_ in rc._group.shards]
    n_pairs = n_folds + 1
    for b_, a_ in zip(before, after):          # every shard: ceil(pairs / 8) passes, every job scored from the counts
        assert a_.scan_calls - b_.scan_calls == -(-n_pairs // 8)
        assert a_.rescore_calls - b_.rescore_calls == len(jobs)
    if len(devices) > 1:                       # several shards: ONE packet gather for the whole grid
        assert rc._group.exchange_stats.get("grid_exchanges", 0) == 1
    for w, g_ in zip(want, got):
        assert w[0] == g_[0] and all(np.array_equal(a, b) for a, b in zip(w[1:], g_[1:]))
    rc.close()


def test_grid_ties_next_to_a_larger_but_close_block():
    """scm.py:271-286 appends a block whose maximum is larger than, but isclose to, the running best without raising
    it: selected blocks can then differ by more than the tie tolerance.  The scoring kernel of the grid stores, per
    64-column segment, the maximum of the surrounding 256 columns -- which may come from the NEXT block -- so the
    tie kernel must treat it as an upper bound.  Block maxima -999899 / -999908 / -999890 (tolerance 10): the tie of
    block 1 sits 110 columns before block 2's larger maximum."""
    n_pos, n_neg, k, ubs, p_ = 5, 120, 3000, 1000, 999999.0
    n = n_pos + n_neg
    X = np.zeros((n, k), dtype=np.uint8)
    X[:2, :] = 1                                   # ordinary k-mers: in 2 of the 5 positives, in every negative
    X[n_pos:, :] = 1
    for col, nc in ((10, 100), (1900, 91), (2010, 109)):
        X[:n_pos, col] = [1, 1, 1, 1, 0]           # presence rule: one positive error ...
        X[n_pos:n_pos + nc, col] = 0               # ... and nc negatives covered
    from kover_b200.utils import _pack_binary_bytes_to_ints
    P = _pack_binary_bytes_to_ints(X, 64)
    pos, neg = np.arange(n_pos), np.arange(n_pos, n)
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    want = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p_,
                                   np.zeros(2 * k, dtype=bool), block_size=ubs)
    assert list(want[1]) == [10, 1900, 2010] and want[0] == 100 - 999999.0      # all three blocks keep their tie
    rc = KmerRuleClassifications(P, n, devices=[0])
    one = rc.best_utility_rules(p_, pos, neg, util_block_size=ubs)
    grid = rc.best_utility_rules_grid([(p_, pos, neg), (p_, pos, neg), (1.0, pos[:4], neg)], util_block_size=ubs)
    for got in (one, grid[0], grid[1]):
        assert got[0] == want[0] and np.array_equal(got[1], np.asarray(want[1], dtype=np.int64))
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    rc.close()


# ---------------------------------------------------------------------------------- a9 / a11
def tree_to_dict(node):
    d = {"depth": node.depth, "criterion_value": float(node.criterion_value),
         "n_by_class": {int(c): int(len(v)) for c, v in node.class_examples_idx.items()}}
    if node.rule is not None:
        d["rule_idx"] = int(node.rule.kmer_index)
        d["equivalent_rules_idx"] = [int(x) for x in node.equivalent_rules_idx]
        d["importance"] = float(node.rule.importance)
        d["left"] = tree_to_dict(node.left_child)
        d["right"] = tree_to_dict(node.right_child)
    return d


def tree_equal(got, want, path="root", rtol=0.0):
    """rtol = 0: bit-exact criterion values (Gini).  rtol > 0 only for the cross-entropy criterion (np.log)."""
    assert got["depth"] == want["depth"], path
    assert got["n_by_class"] == {int(c): v for c, v in want["n_by_class"].items()}, path
    if rtol:
        assert abs(got["criterion_value"] - want["criterion_value"]) <= rtol * abs(want["criterion_value"]), path
    else:
        assert got["criterion_value"] == want["criterion_value"], path
    assert ("rule_idx" in got) == ("rule_idx" in want), path
    if "rule_idx" in want:
        assert got["rule_idx"] == want["rule_idx"], path
        assert got["equivalent_rules_idx"] == want["equivalent_rules_idx"], path
        assert got["importance"] == want["importance"], path
        tree_equal(got["left"], want["left"], path + ".L", rtol)
        tree_equal(got["right"], want["right"], path + ".R", rtol)


@pytest.mark.parametrize("case", cases.CART_CASES, ids=[c[0] for c in cases.CART_CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_cart_fit_golden(golden_dir, case, devices):
    name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
    g = load(golden_dir, name)
    want = json.loads(str(g["tree_json"]))
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    k = P.shape[1]
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
    train = cases.train_subset(n, subset_seed)
    example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
    tiebreaker = None
    if tb == "occurrence":
        occ = rc.sum_rows(train)
        tiebreaker = lambda idx: ko.cart_tiebreaker(idx, occ)
    t = kcart.DecisionTreeClassifier("gini", max_depth, mss, class_importance)
    t.fit(rule_list(k), rc, example_idx, rule_blacklist=[], tiebreaker=tiebreaker)
    tree_equal(tree_to_dict(t.decision_tree), want)


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]], ids=["1shard", "3shards"])
@pytest.mark.parametrize("sorted_labels", [True, False], ids=["sorted", "interleaved"])
def test_level_search_matches_per_node_search(devices, sorted_labels):
    """The nodes of a tree level searched together (class masks of 8 nodes per matrix pass, kv_gini_level) give
    exactly what one gini_best_split per node gives: same minimum, same tie set -- including tiny nodes that tie
    on a large share of the k-mers (more ties than the per-node slot) and a node with no admissible split."""
    seed, n, k = 21, 900, 60000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=sorted_labels)
    rs = np.random.RandomState(seed)
    perm = rs.permutation(n)
    sizes = [300, 200, 150, 100, 60, 30, 20, 12, 8, 5, 4, 3, 2, 2, 2]          # 15 disjoint nodes: two passes
    nodes, at = [], 0
    for sz in sizes:
        members = np.sort(perm[at:at + sz])
        at += sz
        nodes.append({0: members[y[members] == 0], 1: members[y[members] == 1]})
    nodes = [ex for ex in nodes if len(ex[0]) and len(ex[1])]
    assert len(nodes) > 8
    rc = KmerRuleClassifications(P, n, devices=devices)
    n_total = {0: float((y == 0).sum()), 1: float((y == 1).sum())}
    priors = {0: 0.4, 1: 0.6}
    bl = rs.choice(k, size=300, replace=False)
    want = [rc.gini_best_split(ex, priors, n_total, rule_blacklist=bl) for ex in nodes]
    before = [s_.info().scan_calls for s_ in rc._group.shards]
    got = rc.gini_best_splits(nodes, priors, n_total, rule_blacklist=bl)
    after = [s_.info().scan_calls for s_ in rc._group.shards]
    assert all(a - b == -(-len(nodes) // 8) for a, b in zip(after, before))
    assert max(len(w[1]) for w in want) > 1024
    for w, g_ in zip(want, got):
        assert w[0] == g_[0] and np.array_equal(w[1], g_[1])
    rc.close()


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]], ids=["1shard", "3shards"])
@pytest.mark.parametrize("slots", [None, "3", "0"], ids=["arena", "3slots", "off"])
def test_level_sibling_subtraction(devices, slots, monkeypatch):
    """kv_gini_level_family: of two sibling nodes only the smaller is counted in the matrix pass, the other one's class
    counts are parent - sibling.  Three levels of a hand-made family (a root, its two children, grandchildren incl. a
    lone child, a pair whose sizes do not add up to the parent's, an unrelated node) searched armed == the same nodes
    searched unarmed, with blacklist and the device tiebreaker; the number of derived nodes is what the arena allows."""
    from kover_b200.learning.common.rules import NodeExamples
    from kover_b200.learning.experiments import experiment_cart as cexp
    if slots == "0":
        monkeypatch.setenv("KOVER_B200_LEVEL_SUB", "0")
    elif slots is not None:
        monkeypatch.setenv("KOVER_B200_LEVEL_POOL_SLOTS", slots)
    seed, n, k = 31, 1100, 70000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    rs = np.random.RandomState(seed)
    rc = KmerRuleClassifications(P, n, devices=devices)
    train = np.sort(rs.choice(n, size=900, replace=False))
    tb = cexp.OccurrenceTiebreaker(rc, train)
    n_total = {0: float((y[train] == 0).sum()), 1: float((y[train] == 1).sum())}
    priors = {0: 0.45, 1: 0.55}
    bl = rs.choice(k, size=200, replace=False)

    def node(members, parent=None):
        members = np.sort(members)
        return NodeExamples({0: members[y[members] == 0], 1: members[y[members] == 1]}, parent)

    def split(ex, frac):
        members = np.concatenate([ex[0], ex[1]])
        take = rs.rand(len(members)) < frac
        return node(members[take], ex), node(members[~take], ex)

    root = node(train)
    a, b = split(root, 0.3)
    a1, a2 = split(a, 0.5)
    b1, b2 = split(b, 0.9)
    stranger = node(rs.choice(train, size=50, replace=False))
    c1, c2 = split(a1, 0.4)
    d1, d2 = split(b1, 0.02)                            # a tiny node next to a large one
    wrong1, wrong2 = split(b2, 0.5)
    wrong2 = node(np.concatenate([wrong2[0], wrong2[1]])[:-3], b2)       # sizes no longer add up: both are counted
    levels = [[root], [a, b], [a1, a2, b1, stranger, b2], [c1, c2, d2, d1, wrong1, wrong2, node(np.concatenate([a2[0], a2[1]])[::2], a2)]]
    tables = tb.device_occurrences
    derived0 = [s_.info().level_nodes_derived for s_ in rc._group.shards]
    for use_table in (False, True):
        for lv in levels:
            assert all(len(ex[0]) and len(ex[1]) for ex in lv)
            kw = dict(tie_tables=[tables] * len(lv)) if use_table else {}
            got = [(m, np.array(c)) for m, c in rc.gini_best_splits(lv, priors, n_total, rule_blacklist=bl, **kw)]
            plain = [dict(ex) for ex in lv]             # the same nodes without keys: every node counted
            want = rc.gini_best_splits(plain, priors, n_total, rule_blacklist=bl, **kw)
            if len(lv) == 1 and not use_table:
                want = [rc.gini_best_split(plain[0], priors, n_total, rule_blacklist=bl)]
            for (gm, gc), (wm, wc) in zip(got, want):
                assert gm == wm and np.array_equal(gc, np.asarray(wc))
            # (the unarmed search ended the family: arm this level again so that the next one finds its parents)
            rc.gini_best_splits(lv, priors, n_total, rule_blacklist=bl, **kw)
    derived = [s_.info().level_nodes_derived - d0 for s_, d0 in zip(rc._group.shards, derived0)]
    # per sweep: level 1 -> (a, b); level 2 -> (a1, a2), b1 + b2 straddle the stranger; level 3 -> (c1, c2), (d1, d2)
    # (only the first, compared search of a level finds its parents resident: the unarmed one in between ends the family)
    full = 2 * (1 + 2 + 2)
    if slots is None:
        assert derived == [full] * len(devices), derived
    elif slots == "0":
        assert derived == [0] * len(devices)
    else:
        assert all(0 < d < full for d in derived), derived
    tb.release()
    rc.close()


@pytest.mark.parametrize("seed", range(12))
def test_randomized_trees_vs_checker(seed):
    """Whole trees on random small / ragged shapes: level-batched split search + occurrence tiebreaker on the device,
    alone and grown in lock-step with a second tree on other examples (fit_many), 1-3 shards, random blacklists,
    depths and min_samples_split -- against the CPU checker's tree (one node at a time, host tiebreaker)."""
    from kover_b200.learning.experiments import experiment_cart as cexp
    rs = np.random.RandomState(2000 + seed)
    n = int(rs.choice([30, 64, 65, 127, 200, 300]))
    k = int(rs.choice([40, 511, 513, 1500, 4000]))
    dens = rs.choice([0.05, 0.3, 0.5, 0.7, 0.95], size=k)
    X = (rs.rand(n, k) < dens[None, :]).astype(np.uint8)
    P = cases.pack64(X)
    y = ((X[:, rs.randint(k)] ^ (rs.rand(n) < 0.15)) & 1).astype(np.uint8)
    if rs.rand() < 0.5:
        order = np.argsort(y, kind="stable")
        X, y = X[order], y[order]
        P = cases.pack64(X)
    depth, mss = int(rs.choice([2, 4, 7])), int(rs.choice([2, 2, 5]))
    imp = {0: 1.0, 1: float(rs.choice([1.0, 2.5]))}
    blacklist = np.unique(rs.randint(0, k, size=rs.randint(0, max(1, k // 10))))
    devices = [0] * int(rs.choice([1, 2, 3]))
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    rc = KmerRuleClassifications(P, n, devices=devices)
    rules = rule_list(k)
    subsets = [np.sort(rs.choice(n, size=max(4, int(n * f)), replace=False)) for f in (0.8, 0.6)]
    examples = [{0: s_[y[s_] == 0], 1: s_[y[s_] == 1]} for s_ in subsets]
    if any(len(e[0]) == 0 or len(e[1]) == 0 for e in examples):
        pytest.skip("degenerate labels")
    want = []
    for s_, ex in zip(subsets, examples):
        occ = orc.sum_rows(s_)
        want.append(ko.cart_fit(orc, ex, "gini", depth, mss, imp, rule_blacklist=list(blacklist),
                                tiebreaker=lambda idx, occ=occ: ko.cart_tiebreaker(idx, occ)))

    def same(a, b):
        assert (a.rule is None) == b.is_leaf
        assert {int(c): len(v) for c, v in a.class_examples_idx.items()} == {int(c): len(v) for c, v in b.class_examples_idx.items()}
        if a.rule is not None:
            assert int(a.rule_idx) == int(b.rule_idx) and float(a.criterion_value) == float(b.criterion_value)
            assert np.array_equal(np.asarray(a.equivalent_rules_idx), np.asarray(b.equivalent_rules_idx))
            same(a.left_child, b.left_child)
            same(a.right_child, b.right_child)

    # alone
    tb = cexp.OccurrenceTiebreaker(rc, subsets[0])
    t = kcart.DecisionTreeClassifier("gini", depth, mss, imp)
    t.fit(rules, rc, examples[0], rule_blacklist=list(blacklist), tiebreaker=tb)
    tb.release()
    same(t.decision_tree, want[0])
    # two trees in lock-step, each with its own occurrence table
    tbs = [cexp.OccurrenceTiebreaker(rc, s_) for s_ in subsets]
    learners = [kcart.DecisionTreeClassifier("gini", depth, mss, imp) for _ in subsets]
    kcart.fit_many(learners, [dict(rules=rules, example_idx=ex, rule_blacklist=list(blacklist), tiebreaker=b_)
                              for ex, b_ in zip(examples, tbs)], rc)
    for b_ in tbs:
        b_.release()
    for learner, w in zip(learners, want):
        same(learner.decision_tree, w)
    rc.close()


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]], ids=["1shard", "3shards"])
def test_device_occurrence_tiebreaker(devices):
    """The reference's CART tiebreaker (experiment_cart.py:82-94: among the best splits keep the k-mers that occur in
    the most training genomes, numpy.isclose to the maximum) applied by the level search on the device == the host
    tiebreaker applied to the plain tie lists; per-node tables (two trees in one search), tiny nodes with huge tie
    sets, and the single-node call."""
    from kover_b200.learning.experiments import experiment_cart as cexp
    seed, n, k = 23, 700, 50000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    rs = np.random.RandomState(seed)
    perm = rs.permutation(n)
    train_a, train_b = np.sort(perm[:450]), np.sort(perm[200:650])
    nodes, tables_of = [], []
    rc = KmerRuleClassifications(P, n, devices=devices)
    tb_a, tb_b = cexp.OccurrenceTiebreaker(rc, train_a), cexp.OccurrenceTiebreaker(rc, train_b)
    assert tb_a.device_occurrences is not None and tb_a.device_occurrences.table != tb_b.device_occurrences.table
    at = 0
    for sz, tb in ((200, tb_a), (120, tb_b), (60, tb_a), (9, tb_b), (4, tb_a), (3, tb_b), (2, tb_a), (40, None), (2, tb_b), (30, tb_a)):
        pool = train_a if tb is tb_a or tb is None else train_b
        members = np.sort(pool[at % 300:at % 300 + sz])
        at += sz
        ex = {0: members[y[members] == 0], 1: members[y[members] == 1]}
        if len(ex[0]) and len(ex[1]):
            nodes.append(ex)
            tables_of.append(tb)
    n_total = {0: float((y == 0).sum()), 1: float((y == 1).sum())}
    priors = {0: 0.5, 1: 0.5}
    plain = rc.gini_best_splits(nodes, priors, n_total)
    plain = [(s_, np.array(c)) for s_, c in plain]
    got = rc.gini_best_splits(nodes, priors, n_total,
                              tie_tables=[None if t is None else t.device_occurrences for t in tables_of])
    assert max(len(c) for _, c in plain) > 1024
    for (ws, wc), (gs, gc), tb in zip(plain, got, tables_of):
        want = wc if tb is None else tb(wc)                      # the host tiebreaker on the full tie list
        assert gs == ws and np.array_equal(np.sort(want), np.asarray(gc)), (len(wc), len(want), len(gc))
    # occurrences on the device == sum_rows on the host
    assert np.array_equal(tb_a.device_occurrences.host()[:k], rc.sum_rows(train_a)[:k])
    # a single node goes through the same path
    one = rc.gini_best_splits(nodes[:1], priors, n_total, tie_tables=[tables_of[0].device_occurrences])
    assert one[0][0] == plain[0][0] and np.array_equal(np.asarray(one[0][1]), np.sort(tables_of[0](plain[0][1])))
    tb_a.release()
    tb_b.release()
    assert len(rc._free_tables) == 64
    rc.close()


@pytest.mark.parametrize("n_classes", [2, 3, 5])
def test_gini_scores_vs_checker(n_classes):
    seed, n, k = 11, 700, 40000
    P = synthetic_block(seed, n, 0, k)
    rs = np.random.RandomState(seed)
    y = rs.randint(0, n_classes, size=n)
    node = np.sort(rs.choice(n, 400, replace=False))          # a node deeper in the tree: a subset of examples
    ex_all = {c: np.where(y == c)[0] for c in range(n_classes)}
    ex_node = {c: node[y[node] == c] for c in range(n_classes)}
    importance = {c: 1.0 + 0.37 * c for c in range(n_classes)}
    n_total, altered = ko.cart_altered_priors(ex_all, importance)
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    want = ko.gini_rule_score(orc, ex_node, altered, n_total)
    blacklist = [5, 17, 39999]
    want[blacklist] = np.inf
    rc = KmerRuleClassifications(P, n, devices=[0, 0])
    m, cand = rc.gini_best_split(ex_node, altered, n_total, rule_blacklist=blacklist)
    got = rc.gini_scores()
    assert np.array_equal(got, want)                           # bit-exact float64, inf included
    assert m == want.min() and np.array_equal(cand, np.where(want == want.min())[0])


@pytest.mark.parametrize("case", cases.CART_CE_CASES, ids=[c[0] for c in cases.CART_CE_CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_cross_entropy_fit_golden(golden_dir, case, devices):
    """a10 (cart.py:167-207): device counts (one multi-mask pass per node) + the reference's numpy formula on the
    host.  Whole trees against the REAL reference's: selected rules, tie sets (equivalent rules), importances
    identical; criterion values to 4 ulp (np.log's last bit depends on the host's libm / SIMD dispatch)."""
    name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
    g = load(golden_dir, name)
    want = json.loads(str(g["tree_json"]))
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    k = P.shape[1]
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
    train = cases.train_subset(n, subset_seed)
    example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
    tiebreaker = None
    if tb == "occurrence":
        occ = rc.sum_rows(train)
        tiebreaker = lambda idx: ko.cart_tiebreaker(idx, occ)
    t = kcart.DecisionTreeClassifier("cross-entropy", max_depth, mss, class_importance)
    t.fit(rule_list(k), rc, example_idx, rule_blacklist=[], tiebreaker=tiebreaker)
    tree_equal(tree_to_dict(t.decision_tree), want, rtol=1e-15)
    # the root's split search against the reference's raw score vector: same minimum, same tie set
    from kover_b200.learning.learners.cart import _NodeImpurity
    best, ties = t._cross_entropy_split(rc, _NodeImpurity(example_idx, class_importance), example_idx, [])
    ref_score = g["node_scores"][0]
    assert abs(best - ref_score.min()) <= 1e-15 * abs(ref_score.min())
    assert np.array_equal(ties, np.where(ref_score == ref_score.min())[0])
    rc.close()


def test_cross_entropy_vs_checker_randomized():
    """Cross-entropy split search on random nodes (2 and 3 classes, class importances, blacklist) against the
    checker's score vector: same host, same np.log -> bit-exact minimum and tie set."""
    from kover_b200.learning.learners.cart import _NodeImpurity
    for seed in range(6):
        rs = np.random.RandomState(500 + seed)
        n, k, nc = int(rs.randint(100, 900)), int(rs.randint(2000, 9000)), 2 + seed % 2
        P = synthetic_block(seed, n, 0, k)
        y = rs.randint(0, nc, size=n)
        imp = {c: float(rs.choice([0.5, 1.0, 2.0])) for c in range(nc)}
        ex_all = {c: np.where(y == c)[0] for c in range(nc)}
        node = {c: v[rs.rand(len(v)) < 0.4] for c, v in ex_all.items()}
        if seed == 3:
            node[0] = node[0][:0]                     # a class without examples in the node (cart.py:197-198)
        blacklist = np.unique(rs.randint(0, k, size=40))
        orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
        n_total, altered = ko.cart_altered_priors(ex_all, imp)
        score = ko.cross_entropy_rule_score(orc, node, altered, n_total)
        score[blacklist] = np.inf
        rc = KmerRuleClassifications(P, n, devices=[0] if seed % 2 else [0, 0, 0])
        t = kcart.DecisionTreeClassifier("cross-entropy", 3, 2, imp)
        best, ties = t._cross_entropy_split(rc, _NodeImpurity(ex_all, imp), node, blacklist)
        assert best == score.min() and np.array_equal(ties, np.where(score == score.min())[0])
        rc.close()


# ---------------------------------------------------------------------------------- bound / CV model selection
@pytest.mark.parametrize("case", cases.EXPERIMENT_CASES, ids=[c[0] for c in cases.EXPERIMENT_CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_bound_and_cv_selection_golden(golden_dir, case, devices):
    """Bound values, predictions, CV risks and the selected hyper-parameters over the device matrix, all fits
    in lock-step (single shard: through the job queue), against the reference's own functions."""
    from experiment_checks import check_experiment_case
    check_experiment_case(golden_dir, case, lambda P, n, chunks: KmerRuleClassifications(as_dataset(P, chunks), n,
                                                                                        devices=devices))


@pytest.mark.parametrize("case", cases.CART_EXPERIMENT_CASES, ids=[c[0] for c in cases.CART_EXPERIMENT_CASES])
def test_pruned_tree_bound_selection_golden(golden_dir, case):
    """Master tree grown on the device, cost-complexity pruning sequence, bound of every pruned tree, selected
    tree and its predictions, against the reference's own _prune_tree / _bound / _predictions."""
    from experiment_checks import check_cart_experiment_case
    check_cart_experiment_case(golden_dir, case, lambda P, n, chunks: KmerRuleClassifications(as_dataset(P, chunks), n,
                                                                                             devices=[0, 0]))


@pytest.mark.parametrize("case", cases.CART_CV_CASES, ids=[c[0] for c in cases.CART_CV_CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_pruned_tree_cv_selection_golden(golden_dir, case, devices):
    """Fold trees and master tree grown in LOCK-STEP on the device (the nodes of one depth of all the trees share
    matrix passes), CV choice of the pruning parameter, against the reference's _learn_pruned_tree_cv."""
    from experiment_checks import check_cart_cv_case
    check_cart_cv_case(golden_dir, case, lambda P, n, chunks: KmerRuleClassifications(as_dataset(P, chunks), n,
                                                                                     devices=devices))


# ---------------------------------------------------------------------------------- dataset split risk tables
@pytest.mark.parametrize("case", cases.RISK_CASES, ids=[c[0] for c in cases.RISK_CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0]], ids=["1shard", "2shards"])
def test_kmer_risk_tables_golden(golden_dir, case, devices):
    name, mname, lname, seed = case
    g = load(golden_dir, name)
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    train = cases.train_subset(n, seed)
    rc = KmerRuleClassifications(as_dataset(P, chunks), n, devices=devices)
    ur, bk, ba = rc.kmer_risk_tables(train[y[train] == 1], train[y[train] == 0])
    assert np.array_equal(ur, g["unique_risks"])
    assert bk.dtype == g["by_kmer"].dtype and np.array_equal(bk, g["by_kmer"])
    assert ba.dtype == g["by_anti_kmer"].dtype and np.array_equal(ba, g["by_anti_kmer"])


def test_kmer_risk_tables_vs_checker_large():
    seed, n, k = 21, 3000, 150000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    want = ko.kmer_risk_tables(ko.OracleKmerRuleClassifications(P, n, fast=True), pos, neg)
    got = KmerRuleClassifications(P, n, devices=[0]).kmer_risk_tables(pos, neg)
    for w, g_ in zip(want, got):
        assert w.dtype == g_.dtype and np.array_equal(w, g_)


# ---------------------------------------------------------------------------------- full-size properties (config 2)
@pytest.fixture(scope="module")
def config2():
    n, k, seed = 2000, 10_000_000, 72
    shard = _native.Shard(0, n, k)
    shard.generate_synthetic(seed)
    from kover_b200.sharding import ShardGroup

    class Meta(object):
        shape = (32, k)
        dtype = np.dtype(np.uint64)
        chunks = (1, 100000)

    rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([shard], n, k))
    yield rc, n, k, seed
    rc.close()


def test_config2_slices_match_checker(config2):
    rc, n, k, seed = config2
    y = synthetic_labels(seed, n)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    sp, sn = rc.sum_rows(pos), rc.sum_rows(neg)
    assert sp.dtype == np.uint16 and sp.shape == (2 * k,)
    for c0 in (0, 4_999_680, k - 100_000):
        P = synthetic_block(seed, n, c0, 100_000)
        orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
        assert np.array_equal(sp[c0:c0 + 100_000], orc.sum_rows(pos)[:100_000])
        assert np.array_equal(sn[k + c0:k + c0 + 100_000], orc.sum_rows(neg)[100_000:])
        cols = [c0, c0 + 77_777, k + c0 + 99_999]
        local = [0, 77_777, 100_000 + 99_999]
        assert np.array_equal(rc.get_columns(cols), orc.get_columns(local))


def test_config2_scan_properties(config2):
    rc, n, k, seed = config2
    y = synthetic_labels(seed, n)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    sp, sn = rc.sum_rows(pos), rc.sum_rows(neg)
    # presence + absence = |rows| for every k-mer; linearity over a disjoint union
    assert np.array_equal(sp[:k].astype(np.int64) + sp[k:], np.full(k, len(pos)))
    both = rc.sum_rows(train)
    assert np.array_equal(both[:k].astype(np.int64), sp[:k].astype(np.int64) + sn[:k])
    for p in (1.0, 0.1, 999999.0):
        neg_cover = (len(neg) - sn).astype(np.float64)
        pos_err = (len(pos) - sp).astype(np.float64)
        bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg)
        wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - sn, len(pos) - sp, p, np.zeros(2 * k, dtype=bool))
        assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64))
        assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
        u = neg_cover - p * pos_err
        assert bu <= u.max() and np.all(np.isclose(u[bi], u[bi].max(), rtol=2e-5))


def test_config2_gini_properties(config2):
    rc, n, k, seed = config2
    y = synthetic_labels(seed, n)
    train = train_split(seed, n)
    ex = {0: train[y[train] == 0], 1: train[y[train] == 1]}
    n_total, altered = ko.cart_altered_priors(ex, {0: 1.0, 1: 1.0})
    m, cand = rc.gini_best_split(ex, altered, n_total)
    scores = rc.gini_scores()
    assert m == scores.min() and np.array_equal(cand, np.where(scores == m)[0])
    c0 = 1_234_944
    P = synthetic_block(seed, n, c0, 50_000)
    want = ko.gini_rule_score(ko.OracleKmerRuleClassifications(P, n, fast=True), ex, altered, n_total)
    assert np.array_equal(scores[c0:c0 + 50_000], want)
