"""CPU: the model-selection host logic (bound values, predictions, CV risks, best-HP rules, lock-step fit_many)
of kover_b200.learning.experiments over checker-backed shards, against golden vectors replayed from the real
reference's _bound / _predictions / metrics / fit."""
import numpy as np
import pytest

import cases
from experiment_checks import check_cart_cv_case, check_cart_experiment_case, check_driver_case, check_experiment_case
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.sharding import ShardGroup, partition_columns
from oracle_shard import OracleShard


def make_rc(P, n, chunks):
    k = P.shape[1]
    shards = [OracleShard(P, n, c0, nc) for c0, nc in partition_columns(k, 2) if nc > 0]
    return KmerRuleClassifications(P, n, shard_group=ShardGroup(shards, n, k))


@pytest.mark.parametrize("case", cases.EXPERIMENT_CASES, ids=[c[0] for c in cases.EXPERIMENT_CASES])
def test_bound_and_cv_selection(golden_dir, case):
    check_experiment_case(golden_dir, case, make_rc)


@pytest.mark.parametrize("case", cases.CART_EXPERIMENT_CASES, ids=[c[0] for c in cases.CART_EXPERIMENT_CASES])
def test_pruned_tree_bound_selection(golden_dir, case):
    check_cart_experiment_case(golden_dir, case, make_rc)


@pytest.mark.parametrize("case", cases.CART_CV_CASES, ids=[c[0] for c in cases.CART_CV_CASES])
def test_pruned_tree_cv_selection(golden_dir, case):
    check_cart_cv_case(golden_dir, case, make_rc)


@pytest.mark.parametrize("case", cases.DRIVER_CASES, ids=[c[0] for c in cases.DRIVER_CASES])
def test_drivers_against_the_reference(golden_dir, case, tmp_path):
    """learn_SCM / learn_CART reading a Kover-layout HDF5 file (hdf5_lite + the dataset adapter), with the matrix
    served by checker-backed shards: the host side of the drivers against the real reference's drivers."""
    check_driver_case(golden_dir, case, str(tmp_path), make_rc)
