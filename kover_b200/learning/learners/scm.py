"""Set Covering Machine learner driven by the fused device scan.

Interface parity with reference kover/learning/learners/scm.py:29-288: ``UTIL_BLOCK_SIZE``,
``SetCoveringMachine(model_type, p, max_rules)``, ``fit(rules, rule_classifications, positive_example_idx,
negative_example_idx, rule_blacklist, tiebreaker, iteration_callback, iteration_rule_importances)``, the keys of
the ``iteration_info`` dictionaries handed to the callback, ``rule_importances``, the stopping rules and the
exceptions.  The greedy set-cover control flow stays on the host (it is a dozen iterations); what the reference
spends its time on -- two full-matrix ``sum_rows`` passes and several numpy passes over all 2K utilities per
iteration (scm.py:238-288) -- is ONE call to ``rule_classifications.best_utility_rules``, i.e. one pass of
k_scan_tma + k_scm_select + k_scm_ties over the HBM-resident matrix.

The loop is written once, as a generator that asks for one scan per iteration; ``fit`` answers the requests one
by one, ``fit_many`` answers the requests of many learners together (lock-step hyper-parameter grids).
"""
import logging

import numpy as np

from ..common.models import ConjunctionModel, DisjunctionModel, conjunction, disjunction

UTIL_BLOCK_SIZE = 1000000          # rules per utility block of the reference's merge (scm.py:29)

_MODEL_CLASSES = {conjunction: ConjunctionModel, disjunction: DisjunctionModel}


def _compute_rule_importances(rule_classifications, model_rules_idx, training_example_idx):
    """Importance of each rule of the model = share of the examples the model rejects that THIS rule rejects
    (scm.py:32-36)."""
    outputs = rule_classifications.get_columns(model_rules_idx)[training_example_idx]
    rejected = np.flatnonzero(outputs.prod(axis=1) == 0)
    return (float(rejected.shape[0]) - outputs[rejected].sum(axis=0)) / rejected.shape[0]


class BaseSetCoveringMachine(object):
    def __init__(self, model_type, max_rules):
        if model_type not in _MODEL_CLASSES:
            raise ValueError("Unsupported model type.")
        self.model_type = model_type
        self.max_rules = max_rules
        self._flags = {}
        self.rule_importances = []

    # ------------------------------------------------------------------ model bookkeeping
    def _add_rule_to_model(self, rule):
        """A disjunction is learnt as a conjunction on swapped labels, so the rule that enters the model is the
        inverse of the selected one (scm.py:180-184).  Returns the rule as stored."""
        stored = rule.inverse() if self.model_type == disjunction else rule
        self.model.add(stored)
        logging.debug("Rule added to the model: %s", stored)
        return stored

    def _is_fitted(self):
        return len(self.model) > 0

    # ------------------------------------------------------------------ fitting
    def fit(self, rules, rule_classifications, positive_example_idx, negative_example_idx, rule_blacklist=[],
            tiebreaker=None, iteration_callback=None, iteration_rule_importances=False, **kwargs):
        steps = self._fit_steps(rules, rule_classifications, positive_example_idx, negative_example_idx,
                                rule_blacklist, tiebreaker, iteration_callback, iteration_rule_importances,
                                device_state=True, **kwargs)
        answer = None
        while True:
            try:
                request = steps.send(answer)
            except StopIteration:
                return
            if request[0] == "state":                    # the device-resident greedy state
                _, state, blacklist = request
                answer = state.scan(self.p, blacklist, UTIL_BLOCK_SIZE)
            else:
                _, pos_idx, neg_idx, blacklist, extra = request
                answer = self._get_best_utility_rules(rule_classifications=rule_classifications,
                                                      positive_example_idx=pos_idx, negative_example_idx=neg_idx,
                                                      rule_blacklist=blacklist, **extra)

    def _fit_steps(self, rules, rule_classifications, positive_example_idx, negative_example_idx, rule_blacklist=[],
                   tiebreaker=None, iteration_callback=None, iteration_rule_importances=False, device_state=False,
                   host_masks=False, **kwargs):
        """Generator of the greedy loop's requests; the caller answers each through ``send``:
          ("idx", positive_idx, negative_idx, blacklist, utility kwargs) -> the 4-tuple of _get_best_utility_rules;
          ("state", ScmGreedyState, blacklist)   -> the same 4-tuple   (device_state: the remaining examples live as
                                                    masks on the device, rule_classifications.scm_state);
          ("masks", pos_mask, neg_mask, n_pos, n_neg, blacklist) -> the same 4-tuple   (host_masks, used by fit_many:
                                                    the remaining examples are two packed row masks of W words);
          ("column", rule_idx)                   -> the packed matrix column of that rule's k-mer, uint64 [W]
                                                    (host_masks only: fit_many fetches the columns of all learners
                                                    of an iteration with one gather).
        In the two mask modes the example index lists are never filtered on the host (scm.py:133-139)."""
        scan_kwargs = dict((name[len("utility__"):], value) for name, value in (kwargs or {}).items()
                           if name.startswith("utility__"))
        if len(positive_example_idx) == 0 or len(negative_example_idx) == 0:
            raise ValueError("There must be positive and negative examples to train the SCM.")
        to_keep = np.asarray(positive_example_idx)       # examples the model must keep accepting
        to_cover = np.asarray(negative_example_idx)      # examples the model must come to reject
        if self.model_type == disjunction:               # learn the complement concept (scm.py:69-73)
            to_keep, to_cover = to_cover, to_keep
        n_rules = rule_classifications.shape[1]
        if n_rules != len(rules):
            raise ValueError("The number of rules must match between rule_classifications and rules.")
        if len(rule_blacklist) > 0:
            rule_blacklist = np.unique(rule_blacklist)
            if len(rule_blacklist) == n_rules:
                raise ValueError("The blacklist cannot include all the rules.")
            logging.debug("Blacklisting %d rules.", len(rule_blacklist))

        all_training_idx = np.hstack((to_keep, to_cover))
        chosen, importances = [], []
        state, keep_mask, cover_mask = None, None, None
        if device_state and not scan_kwargs and hasattr(rule_classifications, "scm_state"):
            state = rule_classifications.scm_state(to_keep, to_cover)
        elif host_masks and not scan_kwargs:
            keep_mask = rule_classifications.row_mask(to_keep)
            cover_mask = rule_classifications.row_mask(to_cover)
        n_to_keep, n_to_cover = len(to_keep), len(to_cover)
        n_kmers = n_rules // 2
        while n_to_cover > 0 and len(self.model) < self.max_rules:
            info = {"iteration_number": len(self.model) + 1}
            if state is not None:
                request = ("state", state, rule_blacklist)
            elif keep_mask is not None:
                request = ("masks", keep_mask, cover_mask, n_to_keep, n_to_cover, rule_blacklist)
            else:
                request = ("idx", to_keep, to_cover, rule_blacklist, scan_kwargs)
            best_utility, tied, tied_pos_errors, tied_neg_cover = yield request
            info["utility_max"], info["utility_argmax"] = best_utility, tied
            logging.debug("Greatest utility is %.5f (%d rules)", best_utility, len(tied))

            # a rule that rejects nothing at all cannot make progress (scm.py:109)
            useful = tied[(tied_neg_cover != 0) | (tied_pos_errors != 0)]
            if len(useful) == 0:
                logging.debug("The best rule covers no example: stopping.")
                break
            if len(useful) == 1:
                equivalent = np.array([useful[0]])
            else:
                equivalent = tiebreaker(useful)
            info["equivalent_rules_idx"] = equivalent
            selected = equivalent[0]                     # equivalent rules: the first one is kept (scm.py:124)

            info["selected_rule"] = self._add_rule_to_model(rules[selected])
            chosen.append(selected)

            # examples the new rule rejects leave both sets (scm.py:133-139); the column spans ALL genomes
            if state is not None:
                n_to_keep, n_to_cover = state.apply(selected)          # masks &= column, on the device
            elif keep_mask is not None:
                column = yield ("column", int(selected))               # uint64 [W]: the k-mer's packed presence column
                accepted = ~column if selected >= n_kmers else column  # (padding bits are zero in the masks)
                keep_mask = keep_mask & accepted
                cover_mask = cover_mask & accepted
                n_to_keep = int(np.bitwise_count(keep_mask).sum())
                n_to_cover = int(np.bitwise_count(cover_mask).sum())
            else:
                accepted = rule_classifications.get_columns(selected)
                to_cover = to_cover[accepted[to_cover] != 0]
                to_keep = to_keep[accepted[to_keep] != 0]
                n_to_keep, n_to_cover = len(to_keep), len(to_cover)
            logging.debug("Remaining examples to cover: %d, to keep: %d", n_to_cover, n_to_keep)

            if iteration_rule_importances:
                importances = _compute_rule_importances(rule_classifications, chosen, all_training_idx)
                info["rule_importances"] = importances
            if iteration_callback is not None:
                iteration_callback(info)

        if len(chosen) == 0:
            self.rule_importances = []
        elif iteration_rule_importances:
            self.rule_importances = importances
        else:
            self.rule_importances = _compute_rule_importances(rule_classifications, chosen, all_training_idx)

    # ------------------------------------------------------------------ predictions
    def predict(self, X):
        return self._predict(X)

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


class SetCoveringMachine(BaseSetCoveringMachine):
    """The Set Covering Machine (Marchand & Shawe-Taylor, JMLR 2003).

    model_type: "conjunction" or "disjunction"; p: trade-off between covering negatives and erring on positives in
    the utility ``N_neg_covered - p * N_pos_errors``; max_rules: early stopping (scm.py:210-236)."""

    def __init__(self, model_type=conjunction, p=1.0, max_rules=10):
        super(SetCoveringMachine, self).__init__(model_type=model_type, max_rules=max_rules)
        self.model = _MODEL_CLASSES[model_type]()
        self.p = p

    def _get_best_utility_rules(self, rule_classifications, positive_example_idx, negative_example_idx,
                                rule_blacklist=[]):
        """-> (best_utility, best_utility_idx, pos_error_counts, neg_cover_counts), scm.py:238-288."""
        assert isinstance(rule_blacklist, (list, np.ndarray))
        scan = getattr(rule_classifications, "best_utility_rules", None)
        if scan is None:
            raise TypeError("kover_b200's SetCoveringMachine scans a device-resident matrix: pass a "
                            "kover_b200.learning.common.rules.KmerRuleClassifications (there is no CPU path).")
        return scan(self.p, positive_example_idx, negative_example_idx, rule_blacklist=rule_blacklist,
                    util_block_size=UTIL_BLOCK_SIZE)


def fit_many(learners, fit_args, rule_classifications):
    """Fit several SetCoveringMachine learners over the SAME rule_classifications in lock-step (the reference
    runs them one after the other, or in forked workers: experiment_scm.py:196-248, 568-629).

    learners: list of SetCoveringMachine; fit_args: one dict of ``fit`` keyword arguments per learner (rules,
    positive_example_idx, negative_example_idx, rule_blacklist, tiebreaker, iteration_callback,
    iteration_rule_importances).  At every greedy iteration the pending utility scans of all learners are
    answered together by ``rule_classifications.best_utility_rules_grid``: learners whose remaining example sets
    coincide -- every p of the grid and both model types at iteration 1, and whenever they picked the same
    rules -- share one pass over the matrix.  Each learner ends in exactly the state its own ``fit`` would have
    produced (same scans, same order of operations per learner)."""
    # remaining examples as packed row masks (W words per set) when the matrix can score masks directly: no index
    # filtering per learner, and ONE column gather per iteration for all the learners' picks
    masks_mode = hasattr(rule_classifications, "best_utility_rules_grid_masks") and \
        not getattr(rule_classifications, "dataset_removed_rows", None)
    running = {}
    for i, (learner, kw) in enumerate(zip(learners, fit_args)):
        gen = learner._fit_steps(rule_classifications=rule_classifications, host_masks=masks_mode, **kw)
        try:
            running[i] = (gen, next(gen))
        except StopIteration:
            pass
    while running:
        # (1) learners that just picked a rule wait for its column: one gather for all of them
        waiting = [i for i, (_, req) in running.items() if req[0] == "column"]
        if waiting:
            wanted = [running[i][1][1] for i in waiting]
            columns = rule_classifications.packed_rule_columns(wanted)
            for i, column in zip(waiting, columns):
                gen = running[i][0]
                try:
                    running[i] = (gen, gen.send(column))
                except StopIteration:
                    del running[i]
            continue
        # (2) everybody is at a scan: one grid call per distinct blacklist (the experiments use a single blacklist)
        by_blacklist = {}
        for i, (_, req) in running.items():
            if req[0] == "idx" and req[4]:
                raise ValueError("fit_many does not support utility__ keyword arguments")
            blacklist = req[5] if req[0] == "masks" else req[3]
            by_blacklist.setdefault(np.asarray(blacklist, dtype=np.int64).tobytes(), []).append(i)
        answers = {}
        for members in by_blacklist.values():
            first = running[members[0]][1]
            if first[0] == "masks":
                jobs = [(learners[i].p,) + tuple(running[i][1][1:5]) for i in members]
                results = rule_classifications.best_utility_rules_grid_masks(jobs, rule_blacklist=first[5],
                                                                             util_block_size=UTIL_BLOCK_SIZE)
            else:
                jobs = [(learners[i].p, running[i][1][1], running[i][1][2]) for i in members]
                results = rule_classifications.best_utility_rules_grid(jobs, rule_blacklist=first[3],
                                                                       util_block_size=UTIL_BLOCK_SIZE)
            answers.update(zip(members, results))
        still_running = {}
        for i, answer in answers.items():
            gen = running[i][0]
            try:
                still_running[i] = (gen, gen.send(answer))
            except StopIteration:
                pass
        running = still_running
