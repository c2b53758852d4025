"""Worker of tests/test_multigpu_nccl.py: one rank per GPU (NCCL).  Each rank holds a column shard on its own
GPU; the sharded learners must reproduce the CPU checker bit for bit through BOTH exchange paths (peer
mailbox over NVLink, NCCL all-gather)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList  # noqa: E402
from kover_b200.learning.learners import cart as kcart, scm as kscm  # noqa: E402
from kover_b200.sharding import Comm  # noqa: E402
from kover_b200.synthetic import synthetic_block, synthetic_labels, train_split  # noqa: E402
from oracle import kover_oracle as ko  # noqa: E402


def many_equivalent_rules(lr, comm, p2p, fused):
    """Thousands of equivalent rules: every rank's candidate packet overflows, the step falls back to the exact
    two-phase collection; and a shape with a few hundred ties spread over the ranks (merged on the device)."""
    import cases
    rs = np.random.RandomState(77)
    n, k = 150, 60000
    base = (rs.rand(n, 12) < 0.5).astype(np.uint8)
    y = (base[:, 3] & (1 - base[:, 7])).astype(np.uint8)
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    for n_copies in (k, 300):
        X = (rs.rand(n, k) < 0.5).astype(np.uint8)
        cols = rs.choice(k, size=n_copies, replace=False)
        X[:, cols] = base[:, rs.randint(0, 12, size=n_copies)]
        P = cases.pack64(X)
        orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
        rc = KmerRuleClassifications(P, n, devices=[lr], comm=comm)
        for p in (1.0, 999999.0):
            wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p,
                                                     np.zeros(2 * k, dtype=bool), block_size=25000)
            bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, util_block_size=25000)
            assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64)), (p2p, fused, n_copies, p)
            assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
        if n_copies == k and p2p == "1":
            assert rc._group.exchange_stats["two_phase_fallback"] >= 1
        rc.close()


def main():
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = Comm()
    seed, n, k = 5, 900, 300000
    P = synthetic_block(seed, n, 0, k)
    y = synthetic_labels(seed, n, sorted_by_label=False)
    train = train_split(seed, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    orc = ko.OracleKmerRuleClassifications(P, n, fast=True)
    blacklist = [3, k + 7, 2 * k - 1]
    bl = np.zeros(2 * k, dtype=bool)
    bl[blacklist] = True
    used = {}
    # the multi-rank step as ONE kernel per rank (scan, candidates, mailbox exchange and merge on the device), as
    # scan + select + ties + exchange kernels with the merge on the host, and through NCCL
    for p2p, fused in (("1", "1"), ("1", "0"), ("0", "1")):
        os.environ["KOVER_B200_P2P"] = p2p
        os.environ["KOVER_B200_FUSED"] = fused
        rc = KmerRuleClassifications(P, n, devices=[lr], comm=comm)
        used[p2p] = rc._group.p2p_slot_bytes > 0
        many_equivalent_rules(lr, comm, p2p, fused)
        got = rc.sum_rows(pos)
        assert np.array_equal(got, orc.sum_rows(pos))
        for p, ubs in ((1.0, 1000000), (0.316, 99991), (999999.0, 1000000), (2.0, 4096)):
            wu, wi, wp, wn = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p, bl,
                                                     block_size=ubs)
            for rep in range(3):            # consecutive steps alternate the mailbox parity
                bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, rule_blacklist=blacklist, util_block_size=ubs)
                assert bu == wu and np.array_equal(bi, np.asarray(wi, dtype=np.int64)), (p2p, p, ubs)
                assert np.array_equal(bp, wp) and np.array_equal(bn, wn)
        # a hyper-parameter grid over the same examples: one matrix pass per rank, the rest re-scored
        jobs = [(pp, a, b) for pp in (0.5, 1.0, 10.0) for a, b in ((pos, neg), (neg, pos))]
        shard = rc._group.shards[0]
        before = shard.info()
        got = rc.best_utility_rules_grid(jobs, rule_blacklist=blacklist, util_block_size=99991)
        after = shard.info()
        # (one packet gather for the whole grid; every job, the group leader included, is scored from the counts)
        assert after.scan_calls - before.scan_calls == 1 and after.rescore_calls - before.rescore_calls == 6
        for (pp, a, b), g_ in zip(jobs, got):
            wu, wi, wp, wn = ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), pp, bl,
                                                     block_size=99991)
            assert g_[0] == wu and np.array_equal(g_[1], np.asarray(wi, dtype=np.int64)), (p2p, pp)
            assert np.array_equal(g_[2], wp) and np.array_equal(g_[3], wn)
        # the fits of a cross-validation in lock-step: 4 DIFFERENT mask pairs x 2 p x both roles, one pass per rank
        fold_of = np.random.RandomState(seed).randint(0, 3, size=len(train))
        cv_jobs = []
        for f in range(4):
            tr = train if f == 3 else train[fold_of != f]
            fp, fn = tr[y[tr] == 1], tr[y[tr] == 0]
            cv_jobs += [(pp, a, b) for pp in (0.5, 2.0) for a, b in ((fp, fn), (fn, fp))]
        before, ex_before = shard.info(), rc._group.exchange_stats.get("grid_exchanges", 0)
        got = rc.best_utility_rules_grid(cv_jobs, rule_blacklist=blacklist, util_block_size=99991)
        after = shard.info()
        assert after.scan_calls - before.scan_calls == 1 and after.rescore_calls - before.rescore_calls == len(cv_jobs)
        assert rc._group.exchange_stats.get("grid_exchanges", 0) - ex_before == 1
        for (pp, a, b), g_ in zip(cv_jobs, got):
            wu, wi, wp, wn = ko.merge_utility_blocks(len(b) - orc.sum_rows(b), len(a) - orc.sum_rows(a), pp, bl,
                                                     block_size=99991)
            assert g_[0] == wu and np.array_equal(g_[1], np.asarray(wi, dtype=np.int64)), (p2p, pp)
            assert np.array_equal(g_[2], wp) and np.array_equal(g_[3], wn)
        # a full fit and a tree
        rules = LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
        m = kscm.SetCoveringMachine(model_type="conjunction", p=1.0, max_rules=4)
        m.fit(rules, rc, pos, neg, tiebreaker=lambda idx: idx)
        want_model, _, want_imp = ko.scm_fit(orc, "conjunction", 1.0, 4, pos, neg, tiebreaker=lambda idx: idx)
        assert [int(r.kmer_index) + (k if r.type == "absence" else 0) for r in m.model.rules] == want_model
        assert np.array_equal(np.asarray(m.rule_importances), np.asarray(want_imp))
        ex = {0: neg, 1: pos}
        t = kcart.DecisionTreeClassifier("gini", 3, 2, {0: 1.0, 1: 1.0})
        t.fit(rules, rc, ex)
        root = ko.cart_fit(orc, ex, "gini", 3, 2, {0: 1.0, 1: 1.0})

        def walk(a, b):
            assert (a.rule is None) == b.is_leaf
            if a.rule is not None:
                assert int(a.rule_idx) == int(b.rule_idx) and float(a.criterion_value) == float(b.criterion_value)
                walk(a.left_child, b.left_child)
                walk(a.right_child, b.right_child)
        walk(t.decision_tree, root)
        # the experiments' occurrence tiebreaker applied on every rank's device (one exchange per level)
        from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
        occ = orc.sum_rows(train)
        tb = OccurrenceTiebreaker(rc, train)
        assert tb.device_occurrences is not None
        t = kcart.DecisionTreeClassifier("gini", 5, 2, {0: 1.0, 1: 1.0})
        t.fit(rules, rc, ex, tiebreaker=tb)
        tb.release()
        root = ko.cart_fit(orc, ex, "gini", 5, 2, {0: 1.0, 1: 1.0}, tiebreaker=lambda idx: ko.cart_tiebreaker(idx, occ))

        def walk_eq(a, b):
            assert (a.rule is None) == b.is_leaf
            if a.rule is not None:
                assert int(a.rule_idx) == int(b.rule_idx)
                assert np.array_equal(np.asarray(a.equivalent_rules_idx), np.asarray(b.equivalent_rules_idx))
                walk_eq(a.left_child, b.left_child)
                walk_eq(a.right_child, b.right_child)
        walk_eq(t.decision_tree, root)
        # the same with the experiments' occurrence tiebreaker applied on every rank's device (two small reductions + one
        # exchange per level) and the larger sibling of every split derived from its parent's resident counts
        from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
        occ = orc.sum_rows(train)
        tb = OccurrenceTiebreaker(rc, train)
        on_device = tb.device_occurrences is not None
        try:
            t2 = kcart.DecisionTreeClassifier("gini", 5, 2, {0: 1.0, 1: 2.0})
            t2.fit(rules, rc, {0: neg, 1: pos}, rule_blacklist=blacklist[:1], tiebreaker=tb)
        finally:
            tb.release()
        root2 = ko.cart_fit(orc, {0: neg, 1: pos}, "gini", 5, 2, {0: 1.0, 1: 2.0}, rule_blacklist=blacklist[:1],
                            tiebreaker=lambda idx: ko.cart_tiebreaker(idx, occ))
        walk_eq(t2.decision_tree, root2)
        assert on_device and shard.info().level_nodes_derived > 0
        stats = dict(rc._group.exchange_stats)
        rc.close()
    # a peer that never publishes its packet: the waiting rank gives up after the timeout and reports it (both the
    # one-kernel exchange and k_p2p_publish_wait), instead of hanging the GPU
    import time
    os.environ["KOVER_B200_P2P"] = "1"
    os.environ["KOVER_B200_P2P_TIMEOUT_MS"] = "400"
    for fused in ("1", "0"):
        os.environ["KOVER_B200_FUSED"] = fused
        rc = KmerRuleClassifications(P, n, devices=[lr], comm=comm)
        assert rc._group.p2p_slot_bytes > 0
        if comm.rank == 0:
            t0 = time.perf_counter()
            try:
                rc.best_utility_rules(1.0, pos, neg, util_block_size=1000000)
                message = "no error"
            except RuntimeError as e:
                message = str(e)
            waited = time.perf_counter() - t0
            assert "timed out waiting for a peer" in message, message
            assert 0.3 < waited < 10.0, waited
        comm.barrier()
        rc.close()
    del os.environ["KOVER_B200_P2P_TIMEOUT_MS"]
    comm.barrier()
    if comm.rank == 0:
        print("NCCL_WORKER_OK p2p_used=%s nccl_used=%s stats=%s" % (used["1"], not used["0"], stats))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
