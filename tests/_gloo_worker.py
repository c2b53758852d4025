"""Worker of tests/test_multirank_gloo.py: one of WORLD_SIZE ranks (gloo, CPU).  Each rank owns a column
slice behind a checker-backed shard and runs the product's host logic (ShardGroup + Comm + the learners)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList  # noqa: E402
from kover_b200.learning.learners import cart as kcart, scm as kscm  # noqa: E402
from kover_b200.sharding import Comm, ShardGroup, partition_columns  # noqa: E402
from oracle import kover_oracle as ko  # noqa: E402
from oracle_shard import OracleShard  # noqa: E402


STATS = {"one_exchange": 0, "two_phase_fallback": 0}
GROUPS = []


def make_rc(P, n, comm):
    k = P.shape[1]
    slices = partition_columns(k, comm.world_size)
    c0, nc = slices[comm.rank]
    group = ShardGroup([OracleShard(P, n, c0, nc)], n, k, comm=comm, rank_slices=slices)
    GROUPS.append(group)
    return KmerRuleClassifications(P, n, shard_group=group)


def main():
    dist.init_process_group("gloo")
    comm = Comm()
    golden = os.path.join(ROOT, "tests", "golden")
    out = {"rank": comm.rank, "checks": 0}

    # sum_rows / get_columns across ranks
    X, P, n, chunks = cases.matrix("M3")
    g = np.load(os.path.join(golden, "rules_M3.npz"))
    rc = make_rc(P, n, comm)
    for rname, rows in cases.row_sets(n).items():
        got = rc.sum_rows(rows)
        assert got.dtype == g["sum_" + rname].dtype and np.array_equal(got, g["sum_" + rname]), rname
        out["checks"] += 1
    for cname, cols in cases.column_sets(P.shape[1]).items():
        assert np.array_equal(rc.get_columns(cols), g["col_" + cname]), cname
        out["checks"] += 1

    # utility scan: block maxima merged over ranks, ties concatenated
    for case, cap in [(c, cap) for cap in (256, 2) for c in cases.UTILITY_CASES]:
        if case[1] == "MW":
            continue
        gg = np.load(os.path.join(golden, case[0] + ".npz"))
        P, n, chunks, pos, neg, p, blacklist, ubs = cases.utility_inputs(case)
        rc = make_rc(P, n, comm)
        rc._group.PACKET_CAPACITY = cap       # cap = 2 forces the two-phase fallback on every rank
        bu, bi, bp, bn = rc.best_utility_rules(p, pos, neg, rule_blacklist=blacklist, util_block_size=ubs)
        assert bu == float(gg["best_utility"]), case[0]
        assert np.array_equal(bi, gg["best_idx"]) and np.array_equal(bp, gg["best_pos_err"]), case[0]
        assert np.array_equal(bn, gg["best_neg_cover"]), case[0]
        out["checks"] += 1

    # full SCM fits
    for case in cases.SCM_CASES:
        name, mname, lname, model_type, p, max_rules, tb, blacklist, ubs = case
        gg = np.load(os.path.join(golden, name + ".npz"))
        X, P, n, chunks = cases.matrix(mname)
        y = cases.labels(lname)
        k = P.shape[1]
        kscm.UTIL_BLOCK_SIZE = ubs
        rc = make_rc(P, n, comm)
        risks = cases.rule_risks(2 * k)
        tiebreaker = (lambda idx: idx) if tb == "identity" else (lambda idx: ko.scm_tiebreaker(idx, risks, model_type))
        infos = []
        m = kscm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules)
        m.fit(LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k)), rc, np.where(y == 1)[0],
              np.where(y == 0)[0], rule_blacklist=blacklist, tiebreaker=tiebreaker,
              iteration_callback=lambda i: infos.append(dict(i)), iteration_rule_importances=True)
        assert str(m.model) == str(gg["model_str"]), name
        assert len(infos) == int(gg["n_iterations"]), name
        for t, info in enumerate(infos):
            assert info["utility_max"] == float(gg["it%d_utility_max" % t]), name
            assert np.array_equal(np.asarray(info["utility_argmax"], dtype=np.int64), gg["it%d_utility_argmax" % t])
            assert np.array_equal(np.asarray(info["equivalent_rules_idx"], dtype=np.int64), gg["it%d_equivalent" % t])
        assert np.array_equal(np.asarray(m.rule_importances), gg["rule_importances"]), name
        out["checks"] += 1
    kscm.UTIL_BLOCK_SIZE = 1000000

    # the same fits in lock-step (fit_many): scans over identical example sets are shared
    groups = {}
    for case in cases.SCM_CASES:
        groups.setdefault((case[1], case[8], tuple(case[7])), []).append(case)
    for (mname, ubs, blacklist), members in groups.items():
        X, P, n, chunks = cases.matrix(mname)
        k = P.shape[1]
        kscm.UTIL_BLOCK_SIZE = ubs
        rc = make_rc(P, n, comm)
        risks = cases.rule_risks(2 * k)
        learners, args = [], []
        for name, _, lname, model_type, p, max_rules, tb, _, _ in members:
            y = cases.labels(lname)
            tiebreaker = (lambda idx: idx) if tb == "identity" else \
                (lambda idx, mt=model_type: ko.scm_tiebreaker(idx, risks, mt))
            learners.append(kscm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules))
            args.append(dict(rules=LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k)),
                             positive_example_idx=np.where(y == 1)[0], negative_example_idx=np.where(y == 0)[0],
                             rule_blacklist=list(blacklist), tiebreaker=tiebreaker))
        kscm.fit_many(learners, args, rc)
        for case, m in zip(members, learners):
            gg = np.load(os.path.join(golden, case[0] + ".npz"))
            assert str(m.model) == str(gg["model_str"]), case[0]
            assert np.array_equal(np.asarray(m.rule_importances), gg["rule_importances"]), case[0]
            out["checks"] += 1
    kscm.UTIL_BLOCK_SIZE = 1000000

    # CART fits
    for case in cases.CART_CASES:
        name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
        want = json.loads(str(np.load(os.path.join(golden, name + ".npz"))["tree_json"]))
        X, P, n, chunks = cases.matrix(mname)
        y = cases.labels(lname)
        k = P.shape[1]
        rc = make_rc(P, n, comm)
        train = cases.train_subset(n, subset_seed)
        example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
        tiebreaker = None
        if tb == "occurrence":
            occ = rc.sum_rows(train)
            tiebreaker = lambda idx: ko.cart_tiebreaker(idx, occ)
        t = kcart.DecisionTreeClassifier("gini", max_depth, mss, class_importance)
        t.fit(LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k)), rc, example_idx, tiebreaker=tiebreaker)

        def walk(node, w):
            assert node.depth == w["depth"] and float(node.criterion_value) == w["criterion_value"], name
            assert (node.rule is not None) == ("rule_idx" in w), name
            if node.rule is not None:
                assert int(node.rule.kmer_index) == w["rule_idx"], name
                assert [int(x) for x in node.equivalent_rules_idx] == w["equivalent_rules_idx"], name
                walk(node.left_child, w["left"])
                walk(node.right_child, w["right"])
        walk(t.decision_tree, want)
        out["checks"] += 1

    for g_ in GROUPS:
        for key in STATS:
            STATS[key] += g_.exchange_stats[key]
    out.update(STATS)
    assert STATS["one_exchange"] > 0 and STATS["two_phase_fallback"] > 0, STATS   # both protocols exercised
    comm.barrier()
    with open(os.path.join(os.environ["KV_TEST_OUT"], "rank%d.json" % comm.rank), "w") as f:
        json.dump(out, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
