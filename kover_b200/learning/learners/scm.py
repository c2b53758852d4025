"""Set Covering Machine learner driven by the fused device scan.

Mirrors reference kover/learning/learners/scm.py:29-288: ``UTIL_BLOCK_SIZE``,
``_compute_rule_importances``, ``BaseSetCoveringMachine`` and ``SetCoveringMachine`` with the same
constructor / ``fit`` signature, iteration_info keys, stopping rules and exceptions.  The greedy
control flow stays on the host exactly as in the reference (scm.py:88-150); the two full-matrix
``sum_rows`` passes plus the ~5 numpy passes over 2K utilities of ``_get_best_utility_rules``
(scm.py:238-288) are replaced by one call to ``rule_classifications.best_utility_rules``, i.e. one
pass of k_scan2 + k_scm_ties over the HBM-resident matrix.
"""
import logging

import numpy as np

from ..common.models import ConjunctionModel, DisjunctionModel, conjunction, disjunction

UTIL_BLOCK_SIZE = 1000000


def _compute_rule_importances(rule_classifications, model_rules_idx, training_example_idx):
    """scm.py:32-36"""
    model_rule_classifications = rule_classifications.get_columns(model_rules_idx)[training_example_idx]
    model_neg_prediction_idx = np.where(np.prod(model_rule_classifications, axis=1) == 0)[0]
    return (float(len(model_neg_prediction_idx)) -
            model_rule_classifications[model_neg_prediction_idx].sum(axis=0)) / len(model_neg_prediction_idx)


class BaseSetCoveringMachine(object):
    def __init__(self, model_type, max_rules):
        if model_type == conjunction:
            self._add_rule_to_model = self._append_conjunction_model
            self.model_type = conjunction
        elif model_type == disjunction:
            self._add_rule_to_model = self._append_disjunction_model
            self.model_type = disjunction
        else:
            raise ValueError("Unsupported model type.")
        self.max_rules = max_rules
        self._flags = {}
        super(BaseSetCoveringMachine, self).__init__()

    def fit(self, rules, rule_classifications, positive_example_idx, negative_example_idx, rule_blacklist=[],
            tiebreaker=None, iteration_callback=None, iteration_rule_importances=False, **kwargs):
        """scm.py:54-159.  The greedy loop is written once, as a generator that asks for one utility scan per
        iteration (``_fit_steps``); ``fit`` answers each request with its own scan, ``fit_many`` answers the
        requests of many learners together so that scans over the same examples share a matrix pass."""
        steps = self._fit_steps(rules, rule_classifications, positive_example_idx, negative_example_idx,
                                rule_blacklist, tiebreaker, iteration_callback, iteration_rule_importances, **kwargs)
        try:
            request = next(steps)
            while True:
                pos_idx, neg_idx, blacklist, extra = request
                request = steps.send(self._get_best_utility_rules(rule_classifications=rule_classifications,
                                                                  positive_example_idx=pos_idx,
                                                                  negative_example_idx=neg_idx,
                                                                  rule_blacklist=blacklist, **extra))
        except StopIteration:
            pass

    def _fit_steps(self, rules, rule_classifications, positive_example_idx, negative_example_idx, rule_blacklist=[],
                   tiebreaker=None, iteration_callback=None, iteration_rule_importances=False, **kwargs):
        utility_function_additional_args = {}
        if kwargs is not None:
            for key, value in kwargs.items():
                if key[:9] == "utility__":
                    utility_function_additional_args[key[9:]] = value

        # Validate that there are some positive and negative examples
        if len(positive_example_idx) == 0 or len(negative_example_idx) == 0:
            raise ValueError("There must be positive and negative examples to train the SCM.")

        positive_example_idx = np.asarray(positive_example_idx)
        negative_example_idx = np.asarray(negative_example_idx)
        if self.model_type == disjunction:
            # Switch the example labels (scm.py:69-73)
            positive_example_idx, negative_example_idx = negative_example_idx, positive_example_idx

        logging.debug("Got " + str(len(rules)) + " binary rules.")
        if rule_classifications.shape[1] != len(rules):
            raise ValueError("The number of rules must match between rule_classifications and rules.")

        # Validate the rule blacklist
        if len(rule_blacklist) > 0:
            rule_blacklist = np.unique(rule_blacklist)
            if len(rule_blacklist) == rule_classifications.shape[1]:
                raise ValueError("The blacklist cannot include all the rules.")
            logging.debug("Blacklisting: {0:d} kmers ({1:d} rules)".format(len(rule_blacklist) // 2,
                                                                           len(rule_blacklist)))

        training_example_idx = np.hstack((positive_example_idx, negative_example_idx))  # needed for rule importances
        model_rules_idx = []
        model_rule_importances = []
        while len(negative_example_idx) > 0 and len(self.model) < self.max_rules:
            iteration_info = {"iteration_number": len(self.model) + 1}

            best_utility, best_utility_idx, best_utility_pos_error_counts, best_utility_neg_cover_counts = \
                yield (positive_example_idx, negative_example_idx, rule_blacklist, utility_function_additional_args)

            iteration_info["utility_max"] = best_utility
            iteration_info["utility_argmax"] = best_utility_idx
            logging.debug("Greatest utility is %.5f" % iteration_info["utility_max"])
            logging.debug("There are %d rules with the same utility." % len(iteration_info["utility_argmax"]))

            # Do not select rules that cover no negative examples and make errors on no positive examples
            best_utility_idx = iteration_info["utility_argmax"][
                np.logical_or(best_utility_neg_cover_counts != 0, best_utility_pos_error_counts != 0)]
            if len(best_utility_idx) == 0:
                logging.debug("The rule of maximal utility does not cover negative examples or make errors" +
                              " on positive examples. It will not be added to the model. Stopping here.")
                break

            # Apply a user-specified tiebreaker if necessary
            if len(best_utility_idx) == 1:
                best_rule_idx = best_utility_idx[0]
                iteration_info["equivalent_rules_idx"] = np.array([best_rule_idx])
            elif len(best_utility_idx) > 1:
                best_rule_idx = tiebreaker(best_utility_idx)
                logging.debug("The tiebreaker returned %d equivalent rules." % len(best_rule_idx))
                iteration_info["equivalent_rules_idx"] = best_rule_idx
                best_rule_idx = best_rule_idx[0]  # If many are equivalent, just take the first one.

            # Add the best rule to the model
            iteration_info["selected_rule"] = self._add_rule_to_model(rules[best_rule_idx])
            model_rules_idx.append(best_rule_idx)

            # Get the best rule's classification for each example
            best_rule_classifications = rule_classifications.get_columns(best_rule_idx)

            # Discard examples predicted as negative
            negative_example_idx = negative_example_idx[best_rule_classifications[negative_example_idx] != 0]
            positive_example_idx = positive_example_idx[best_rule_classifications[positive_example_idx] != 0]
            logging.debug("Remaining negative examples:" + str(len(negative_example_idx)))
            logging.debug("Remaining positive examples:" + str(len(positive_example_idx)))

            # If required, compute the current model's rule importances
            if iteration_rule_importances:
                model_rule_importances = _compute_rule_importances(rule_classifications, model_rules_idx,
                                                                   training_example_idx)
                iteration_info["rule_importances"] = model_rule_importances

            if iteration_callback is not None:
                iteration_callback(iteration_info)

        # Get the complete model's rule importances
        if len(model_rules_idx) > 0:
            if iteration_rule_importances:
                self.rule_importances = model_rule_importances
            else:
                self.rule_importances = _compute_rule_importances(rule_classifications, model_rules_idx,
                                                                  training_example_idx)
        else:
            self.rule_importances = []

    def predict(self, X):
        return self._predict(X)

    def _append_conjunction_model(self, new_rule):
        self.model.add(new_rule)
        logging.debug("Rule added to the model: " + str(new_rule))
        return new_rule

    def _append_disjunction_model(self, new_rule):
        new_rule = new_rule.inverse()
        self.model.add(new_rule)
        logging.debug("Rule added to the model: " + str(new_rule))
        return new_rule

    def _is_fitted(self):
        return len(self.model) > 0

    def _predict(self, X):
        if not self._is_fitted():
            raise RuntimeError("A model must be fitted prior to calling predict.")
        return self.model.predict(X)

    def _predict_proba(self, X):
        if not self._is_fitted():
            raise RuntimeError("A model must be fitted prior to calling predict.")
        if not self._flags.get("PROBABILISTIC_PREDICTIONS", False):
            raise RuntimeError("The predictor does not support probabilistic predictions.")
        return self.model.predict_proba(X)


def fit_many(learners, fit_args, rule_classifications):
    """Fit several SetCoveringMachine learners over the SAME rule_classifications in lock-step (the reference
    runs them one after the other, or in forked workers: experiment_scm.py:196-248, 568-629).

    learners: list of SetCoveringMachine; fit_args: one dict of ``fit`` keyword arguments per learner
    (rules, positive_example_idx, negative_example_idx, rule_blacklist, tiebreaker, iteration_callback,
    iteration_rule_importances).  At every greedy iteration the pending utility scans of all learners are
    answered together by ``rule_classifications.best_utility_rules_grid``: learners whose remaining example
    sets coincide -- every p of the grid and both model types at iteration 1, and whenever they picked the
    same rules -- share one pass over the matrix.  Each learner ends in exactly the state its own ``fit``
    would have produced (same scans, same order of operations per learner)."""
    gens, pending = [], {}
    for i, (learner, kw) in enumerate(zip(learners, fit_args)):
        kw = dict(kw)
        kw["rule_classifications"] = rule_classifications
        g = learner._fit_steps(**kw)
        gens.append(g)
        try:
            pending[i] = next(g)
        except StopIteration:
            pass
    while pending:
        # one grid call per distinct blacklist (the experiments use a single blacklist for all fits)
        by_blacklist = {}
        for i, (pos_idx, neg_idx, blacklist, extra) in pending.items():
            if extra:
                raise ValueError("fit_many does not support utility__ keyword arguments")
            by_blacklist.setdefault(np.asarray(blacklist, dtype=np.int64).tobytes(), []).append(i)
        answers = {}
        for members in by_blacklist.values():
            jobs = [(learners[i].p, pending[i][0], pending[i][1]) for i in members]
            res = rule_classifications.best_utility_rules_grid(jobs, rule_blacklist=pending[members[0]][2],
                                                               util_block_size=UTIL_BLOCK_SIZE)
            answers.update(zip(members, res))
        nxt = {}
        for i, ans in answers.items():
            try:
                nxt[i] = gens[i].send(ans)
            except StopIteration:
                pass
        pending = nxt


class SetCoveringMachine(BaseSetCoveringMachine):
    """The Set Covering Machine (Marchand & Shawe-Taylor, JMLR 2003); parameters as scm.py:210-236."""

    def __init__(self, model_type=conjunction, p=1.0, max_rules=10):
        super(SetCoveringMachine, self).__init__(model_type=model_type, max_rules=max_rules)
        if model_type == conjunction:
            self.model = ConjunctionModel()
        elif model_type == disjunction:
            self.model = DisjunctionModel()
        else:
            raise ValueError("Unsupported model type.")
        self.p = p

    def _get_best_utility_rules(self, rule_classifications, positive_example_idx, negative_example_idx,
                                rule_blacklist=[]):
        """scm.py:238-288 -> (best_utility, best_utility_idx, pos_error_counts, neg_cover_counts)."""
        assert isinstance(rule_blacklist, list) or isinstance(rule_blacklist, np.ndarray)
        if not hasattr(rule_classifications, "best_utility_rules"):
            raise TypeError("kover_b200's SetCoveringMachine scans a device-resident matrix: pass a "
                            "kover_b200.learning.common.rules.KmerRuleClassifications (there is no CPU path).")
        return rule_classifications.best_utility_rules(self.p, positive_example_idx, negative_example_idx,
                                                       rule_blacklist=rule_blacklist,
                                                       util_block_size=UTIL_BLOCK_SIZE)
