#!/usr/bin/env python
"""bench.py -- k-mer rules scored per second per SCM iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One STEP = one SCM iteration scan (`_get_best_utility_rules`, reference scm.py:238-288) over the
whole HBM-resident packed matrix: 2*K rules scored (presence + absence), the p value and the
(fold, model-type) example masks cycling through config 2's hyper-parameter grid.

N = 1 workload: BASELINE.json configs[1] -- 2 000 genomes x 10 000 000 k-mers (W = 32 packed rows,
2.56 GB), synthetic (hash-generated on the device, kover_b200/synthetic.py), genomes sorted by label
as the reference's dataset builder does (create.py:190-194).  N > 1: weak scaling, every rank holds
a 2 000 x 10 000 000 column shard of a 2 000 x (N * 10 M) matrix; the only exchange per step is the
per-utility-block maxima + tie lists (sharding.ShardGroup over NCCL).

value     rules/s, K steps timed with CUDA events on the shard's stream (inputs resident in HBM)
roofline  algorithmic bytes of the scan kernel (8 * W_active * K_local per launch) / its own
          CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
e2e       the same steps through the reference-shaped Python API with HOST index arrays in and
          host result arrays out (wall clock, synchronised both sides)
cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/), driven with
          the reference's own compiled popcount kernel when oracle/_ref holds it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GENOMES = 2000
N_KMERS_PER_GPU = 10_000_000
SEED = 72
P_GRID = [0.1, 0.5, 1.0, 2.0, 5.0, 10.0]
N_FOLDS = 5
UTIL_BLOCK_SIZE = 1000000
CPU_SAMPLE_COLS = 1_000_000


# ------------------------------------------------------------------------------------------ workload
def step_inputs(n_genomes, seed, sorted_by_label=True):
    """The (p, pos_idx, neg_idx) of every step in the cycle: p grid x {conjunction, disjunction} x
    {full training set, 5 CV folds} -- the fits config 2 runs (experiment_scm.py:196-248)."""
    from kover_b200.synthetic import synthetic_labels, train_split
    y = synthetic_labels(seed, n_genomes, sorted_by_label=sorted_by_label)
    train = train_split(seed, n_genomes)
    rs = np.random.RandomState(seed)
    fold_of = rs.randint(0, N_FOLDS, size=train.shape[0])
    subsets = [train] + [train[fold_of != f] for f in range(N_FOLDS)]
    steps = []
    for sub in subsets:
        pos, neg = sub[y[sub] == 1], sub[y[sub] == 0]
        for model_type in ("conjunction", "disjunction"):
            for p in P_GRID:
                steps.append((p, pos, neg) if model_type == "conjunction" else (p, neg, pos))
    return steps


def active_words(pos, neg, n_genomes):
    w = np.zeros((n_genomes + 63) // 64, dtype=bool)
    w[np.asarray(pos) // 64] = True
    w[np.asarray(neg) // 64] = True
    return int(w.sum())


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clocks and throttle reasons DURING the timed region.  NVML through pynvml (a few microseconds per
    query); spawning nvidia-smi from inside the region measurably slows the steps it is supposed to watch
    (8 % at N=4), so it is only the fallback.  Rank 0 samples every GPU of the job."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, devices, enabled=True, period=0.05):
        self.devices, self.samples, self._stop, self._t = list(devices), [], threading.Event(), None
        self.enabled, self.period, self.source = enabled, period, None
        self._nvml = None
        if enabled:
            try:
                import pynvml
                pynvml.nvmlInit()
                self._handles = [pynvml.nvmlDeviceGetHandleByIndex(d) for d in self.devices]
                self._nvml, self.source = pynvml, "nvml"
            except Exception:  # noqa: BLE001
                self.source, self.period = "nvidia-smi", max(period, 0.25)

    def _poll_nvml(self):
        nv = self._nvml
        for h in self._handles:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            flag = lambda name: "Active" if (r & getattr(nv, name, 0)) else "Not Active"
            self.samples.append([str(sm), str(mx), flag("nvmlClocksThrottleReasonHwSlowdown"),
                                 flag("nvmlClocksThrottleReasonHwThermalSlowdown"),
                                 flag("nvmlClocksThrottleReasonSwThermalSlowdown"),
                                 flag("nvmlClocksThrottleReasonSwPowerCap")])

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", ",".join(str(d) for d in self.devices), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        for row in out.strip().splitlines():
            f = [x.strip() for x in row.split(",")]
            if len(f) >= 6:
                self.samples.append(f)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._poll_nvml() if self._nvml is not None else self._poll_smi()
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.enabled:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


# ------------------------------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_init(n_genomes, seed, c0, n_cols, use_ref_kernel):
    from kover_b200.synthetic import synthetic_block
    from oracle import kover_oracle as ko
    P = synthetic_block(seed, n_genomes, c0, n_cols)

    class Chunked(np.ndarray):
        chunks = None
    d = P.view(Chunked)
    d.chunks = (1, 100000)                       # the HDF5 chunk shape the reference iterates (create.py:230)
    rc = ko.OracleKmerRuleClassifications(d, n_genomes, fast=False)
    kind = "port"
    if use_ref_kernel:
        try:
            from oracle import ref_loader
            pc = ref_loader.load_reference_popcount()
            rc.inplace_popcount = pc.inplace_popcount_64      # the reference's own compiled popcount.pyx
            kind = "reference-kernel"
        except Exception:  # noqa: BLE001
            pass
    _CPU["rc"], _CPU["ko"], _CPU["kind"] = rc, ko, kind


def _cpu_step(args):
    p, pos, neg = args
    rc, ko = _CPU["rc"], _CPU["ko"]
    t0 = time.perf_counter()
    out = ko.best_utility_rules(rc, p, pos, neg)
    return time.perf_counter() - t0, float(out[0]), int(len(out[1]))


def ref_kernel_available():
    import sysconfig
    so = os.path.join(ROOT, "oracle", "_ref", "popcount" + sysconfig.get_config_var("EXT_SUFFIX"))
    return os.path.isfile(so) or os.path.isdir("/root/reference")


def cpu_baseline(n_genomes, seed, steps, n_steps=4):
    """1 core (the reference hot path is single-threaded), bounded sample: CPU_SAMPLE_COLS columns."""
    _cpu_init(n_genomes, seed, 0, CPU_SAMPLE_COLS, ref_kernel_available())
    times = [_cpu_step(steps[i % len(steps)])[0] for i in range(n_steps)]
    best = min(times)
    return {"value": 2.0 * CPU_SAMPLE_COLS / best, "unit": "rules/s", "cores": 1,
            "kind": "port", "native_kernel": _CPU["kind"],
            "sample": "%d of %d k-mer columns x all %d genomes (in RAM, no HDF5/gzip), best of %d scans of "
                      "oracle.best_utility_rules (2 sum_rows block loops + utility merge); scales linearly in columns"
                      % (CPU_SAMPLE_COLS, N_KMERS_PER_GPU, n_genomes, n_steps)}


def run_reference_arm(args):
    """The reference's CPU implementation of the path on all host cores: the reference parallelises over
    hyper-parameter combinations with a fork pool (experiment_scm.py:217), so every worker runs an
    independent scan (its own p / masks) over the same bounded sample of the matrix."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    steps = step_inputs(N_GENOMES, SEED)
    use_ref = ref_kernel_available()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(N_GENOMES, SEED, 0, CPU_SAMPLE_COLS, use_ref)) as pool:
        batch = lambda i: [steps[(i * cores + j) % len(steps)] for j in range(cores)]
        for w in range(args.warmup):
            pool.map(_cpu_step, batch(w))
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_step, batch(args.warmup + s))
        dt = time.perf_counter() - t0
    _cpu_init(N_GENOMES, SEED, 0, 1024, use_ref)
    rules = 2.0 * CPU_SAMPLE_COLS * cores * args.steps
    value = rules / dt
    sample = ("each step = %d concurrent scans (one per host core, the reference's Pool-over-HP parallelism) of "
              "%d of %d k-mer columns x %d genomes held in RAM (no HDF5/gzip)" %
              (cores, CPU_SAMPLE_COLS, N_KMERS_PER_GPU * args.gpus, N_GENOMES))
    line = {
        "impl": "reference", "metric": "kmer_rules_scored_per_sec_per_scm_iteration", "value": value,
        "unit": "rules/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 popcount + f64 utility", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "rules/s", "cores": cores, "kind": "port",
                         "native_kernel": _CPU["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "rules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def extra_measurements(rc, shard, steps, k_local, reps=12):
    """Outside the contract's timed regions (N=1 only): the other consumers of the resident matrix."""
    from kover_b200.synthetic import synthetic_labels, train_split
    from kover_b200.utils import build_row_mask
    out = {}
    shard.set_profiling(True)          # the extras report scan-kernel times (CUDA events around every scan)

    def timed(fn, n=reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3

    # CART root-node split scan (cart.py:219-250): K presence rules scored per node
    y = synthetic_labels(SEED, N_GENOMES)
    train = train_split(SEED, N_GENOMES)
    ex = {0: train[y[train] == 0], 1: train[y[train] == 1]}
    n_total = {c: float(len(v)) for c, v in ex.items()}
    priors = {c: n_total[c] / sum(n_total.values()) for c in ex}
    a = shard.info().scan_ms_total
    ms = timed(lambda: rc.gini_best_split(ex, priors, n_total))
    kms = (shard.info().scan_ms_total - a) / (reps + 1)
    out["cart_gini_node_scan"] = {"ms_per_node": ms, "splits_per_s": k_local / (ms * 1e-3), "scan_kernel_ms": kms,
                                  "scan_kernel_GBps": 8.0 * 32 * k_local / (kms * 1e-3) / 1e9}
    # the general two-mask regime: labels NOT sorted, so both masks hit every packed row (4 popc per word)
    p, pos, neg = step_inputs(N_GENOMES, SEED, sorted_by_label=False)[0]
    a = shard.info().scan_ms_total
    ms = timed(lambda: rc.best_utility_rules(p, pos, neg, util_block_size=UTIL_BLOCK_SIZE))
    kms = (shard.info().scan_ms_total - a) / (reps + 1)
    out["scm_scan_unsorted_labels"] = {"ms_per_step": ms, "rules_per_s": 2.0 * k_local / (ms * 1e-3), "scan_kernel_ms": kms,
                                       "scan_kernel_GBps": 8.0 * 32 * k_local / (kms * 1e-3) / 1e9}
    # north-star item 4: one fold's whole hyper-parameter grid at iteration 1 = 6 p values x {conjunction,
    # disjunction}: 12 jobs over the same examples -> one matrix scan + 11 re-scores of the resident counts
    _, pos, neg = steps[0]
    jobs = [(pp, a, b) for pp in P_GRID for a, b in ((pos, neg), (neg, pos))]
    ia = shard.info()
    ms = timed(lambda: rc.best_utility_rules_grid(jobs, util_block_size=UTIL_BLOCK_SIZE), 6)
    ib = shard.info()
    out["hp_grid_iteration_12_jobs"] = {"ms_per_grid": ms, "job_rules_per_s": len(jobs) * 2.0 * k_local / (ms * 1e-3),
                                        "matrix_scans_per_grid": (ib.scan_calls - ia.scan_calls) / 7.0,
                                        "rescores_per_grid": (ib.rescore_calls - ia.rescore_calls) / 7.0}
    # configs[1] in full at iteration 1: {train, 5 CV folds} x 6 p x {conjunction, disjunction} = 72 jobs over SIX
    # distinct mask pairs: k_count_tma counts the 12 masks in ONE matrix pass, every job is scored from the counts
    rs = np.random.RandomState(SEED)
    train_all = np.sort(np.concatenate((pos, neg)))
    fold_of = rs.randint(0, 5, size=len(train_all))
    is_pos = np.isin(train_all, pos)
    cv_jobs = []
    for f in range(6):
        keep = np.ones(len(train_all), dtype=bool) if f == 5 else fold_of != f
        fp, fn = train_all[keep & is_pos], train_all[keep & ~is_pos]
        cv_jobs += [(pp, a, b) for pp in P_GRID for a, b in ((fp, fn), (fn, fp))]
    ia = shard.info()
    ms = timed(lambda: rc.best_utility_rules_grid(cv_jobs, util_block_size=UTIL_BLOCK_SIZE), 6)
    ib = shard.info()
    out["cv_grid_iteration_72_jobs"] = {"ms_per_grid": ms, "job_rules_per_s": len(cv_jobs) * 2.0 * k_local / (ms * 1e-3),
                                        "distinct_mask_pairs": 6,
                                        "matrix_passes_per_grid": (ib.scan_calls - ia.scan_calls) / 7.0,
                                        "rescores_per_grid": (ib.rescore_calls - ia.rescore_calls) / 7.0}
    # whole greedy fits through the reference-shaped learners (host control flow + get_columns + bookkeeping)
    from kover_b200.learning.common.rules import LazyKmerRuleList
    from kover_b200.learning.learners.cart import DecisionTreeClassifier
    from kover_b200.learning.learners.scm import SetCoveringMachine
    rules = LazyKmerRuleList(np.arange(k_local), np.arange(k_local))
    # (one untimed fit first: the device-resident greedy state allocates its buffers on first use)
    SetCoveringMachine(model_type="conjunction", p=1.0, max_rules=2).fit(rules, rc, pos, neg, tiebreaker=lambda idx: idx)
    its = []
    m = SetCoveringMachine(model_type="conjunction", p=1.0, max_rules=10)
    la = shard.info().total_launches
    t0 = time.perf_counter()
    m.fit(rules, rc, pos, neg, tiebreaker=lambda idx: idx, iteration_callback=lambda i: its.append(1))
    dt = time.perf_counter() - t0
    out["scm_fit_max_rules_10"] = {"seconds": dt, "iterations": len(its), "ms_per_iteration": 1e3 * dt / max(1, len(its)),
                                   "kernel_launches": int(shard.info().total_launches - la),
                                   "note": "remaining-example masks resident on the device (kv_mask_and_column); includes "
                                           "the final rule importances (one get_columns of the model's rules)"}
    # (one untimed fit first: the level search allocates its pinned tie arena and sort buffers on first use)
    DecisionTreeClassifier("gini", 5, 2, {0: 1.0, 1: 1.0}).fit(rules, rc, ex)
    t = DecisionTreeClassifier("gini", 5, 2, {0: 1.0, 1: 1.0})
    t0 = time.perf_counter()
    t.fit(rules, rc, ex)
    dt = time.perf_counter() - t0
    n_split = len(t.decision_tree.rules)
    out["cart_fit_max_depth_5"] = {"seconds": dt, "split_nodes": n_split, "ms_per_split_node": 1e3 * dt / max(1, n_split)}
    # configs[2]'s depth with the experiments' tiebreaker (experiment_cart.py:82-94, 264): the nodes of a level are
    # searched together, class masks of 8 nodes per matrix pass, the occurrence tiebreaker applied on the device
    from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
    warm = OccurrenceTiebreaker(rc, train)             # (untimed: first-use allocation of the device occurrence table)
    DecisionTreeClassifier("gini", 3, 2, {0: 1.0, 1: 1.0}).fit(rules, rc, ex, tiebreaker=warm)
    warm.release()
    t = DecisionTreeClassifier("gini", 10, 2, {0: 1.0, 1: 1.0})
    ia = shard.info()
    t0 = time.perf_counter()
    tb = OccurrenceTiebreaker(rc, train)
    t.fit(rules, rc, ex, tiebreaker=tb)
    tb.release()
    dt = time.perf_counter() - t0
    ib = shard.info()
    n_split = len(t.decision_tree.rules)
    out["cart_fit_max_depth_10"] = {"seconds": dt, "split_nodes": n_split, "ms_per_split_node": 1e3 * dt / max(1, n_split),
                                    "matrix_passes": ib.scan_calls - ia.scan_calls}
    # the reference's own operator API: sum_rows returns a 2K-long host array (D2H of 2K x uint16 per call)
    out["sum_rows_dropin_ms"] = timed(lambda: rc.sum_rows(steps[0][1]), 4)
    # upload path: host numpy block -> pinned ring -> HBM (kv_upload_block), on 1/8 of the matrix
    cols = k_local // 8
    block = shard.read_block(0, shard.n_words, 0, cols)
    shard.upload_block(block, 0, shard.col0)           # (untimed: the pinned staging ring is allocated on first use)
    t0 = time.perf_counter()
    shard.upload_block(block, 0, shard.col0)
    dt = time.perf_counter() - t0
    out["upload_GBps"] = block.nbytes / dt / 1e9
    out["upload_full_matrix_s_extrapolated"] = 8.0 * 32 * k_local / (block.nbytes / dt)
    return out


def _peak_gbs():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def parity_check(rc, comm, rank, n_genomes, k_local, k_total, seed, pos, neg, p_values=(0.1, 10.0), slice_cols=50_000):
    """Before any timing: (1) the presence / absence counts of two column slices of THIS rank's shard against the CPU
    checker run on the CPU twin of the device generator; (2) the merged best utility / tie set of one multi-rank step
    against the checker's block merge (oracle.merge_utility_blocks, scm.py:258-288) over the gathered count vectors
    (rank 0 computes it).  The oracle is used as the checker only.  -> True on every rank, or raises."""
    from kover_b200.synthetic import synthetic_block
    from oracle import kover_oracle as ko
    sp, sn = rc.sum_rows(pos), rc.sum_rows(neg)                      # collective: the reference's full 2K vectors
    ok = True
    col0 = rank * k_local
    for c0 in (col0 + min(1_234_944, max(0, k_local - slice_cols)) // 512 * 512, col0 + max(0, k_local - slice_cols)):
        n_c = min(slice_cols, k_local)
        P = synthetic_block(seed, n_genomes, c0, n_c)
        orc = ko.OracleKmerRuleClassifications(P, n_genomes, fast=True)
        ok &= bool(np.array_equal(sp[c0:c0 + n_c], orc.sum_rows(pos)[:n_c]))
        ok &= bool(np.array_equal(sn[k_total + c0:k_total + c0 + n_c], orc.sum_rows(neg)[n_c:]))
    for p in p_values:
        got = rc.best_utility_rules(p, pos, neg, util_block_size=UTIL_BLOCK_SIZE)      # collective
        if rank == 0:
            want = ko.merge_utility_blocks(len(neg) - sn, len(pos) - sp, p, np.zeros(2 * k_total, dtype=bool),
                                           block_size=UTIL_BLOCK_SIZE)
            ok &= bool(got[0] == want[0] and np.array_equal(got[1], np.asarray(want[1], dtype=np.int64))
                       and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3]))
    if comm is not None:
        ok = bool(comm.allreduce_min(np.array([int(ok)], dtype=np.int64))[0])
    if not ok:
        raise SystemExit("bench: parity check against the CPU checker FAILED (rank %d)" % rank)
    return True


def measure_shape(name, n_genomes, k_local, world, rank, local_rank, comm, seed, kind, steps=10):
    """One BASELINE config (or its per-GPU shard) resident on this rank's GPU: parity on slices + one merged step
    against the checker, then `steps` scans timed on the device.  kind: "scm" (utility scan) or "cart" (Gini scan)."""
    import time as _t
    from kover_b200 import _native
    from kover_b200.learning.common.rules import KmerRuleClassifications
    from kover_b200.sharding import ShardGroup
    from kover_b200.synthetic import synthetic_labels, train_split
    k_total = k_local * world
    W = (n_genomes + 63) // 64
    shard = _native.Shard(local_rank, n_genomes, k_total, rank * k_local, k_local)
    t0 = _t.perf_counter()
    shard.generate_synthetic(seed)
    gen_s = _t.perf_counter() - t0
    group = ShardGroup([shard], n_genomes, k_total, comm=comm, rank_slices=[(r * k_local, k_local) for r in range(world)])

    class Meta(object):
        shape = (W, k_total)
        dtype = np.dtype(np.uint64)
        chunks = (1, 100000)
    rc = KmerRuleClassifications(Meta(), n_genomes, shard_group=group)
    y = synthetic_labels(seed, n_genomes)
    train = train_split(seed, n_genomes)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    out = {"n_genomes": n_genomes, "packed_rows": W, "n_kmers_per_gpu": k_local, "n_kmers_total": k_total, "n_gpus": world,
           "matrix_GB_per_gpu": 8.0 * W * k_local / 1e9, "generate_s": gen_s}
    out["parity_checked"] = parity_check(rc, comm, rank, n_genomes, k_local, k_total, seed, pos, neg, p_values=(1.0,),
                                         slice_cols=20_000 if W > 100 else 50_000)
    peak, _ = _peak_gbs()
    w_act = active_words(pos, neg, n_genomes)
    if kind == "scm":
        call = lambda: rc.best_utility_rules(1.0, pos, neg, util_block_size=UTIL_BLOCK_SIZE)
        units = 2.0 * k_total
    else:
        from kover_b200.learning.learners.cart import _NodeImpurity
        ex = {0: neg, 1: pos}
        impurity = _NodeImpurity(ex, {0: 1.0, 1: 1.0})                 # altered priors / class totals (cart.py:67-77)
        call = lambda: rc.gini_best_split(ex, impurity.priors, impurity.n_total)
        units = 1.0 * k_total
    for _ in range(3):
        call()
    if comm is not None:
        comm.barrier()
    shard.timer_start()
    for _ in range(steps):
        call()
    dev_ms = shard.timer_stop()
    shard.set_profiling(True)
    a = shard.info().scan_ms_total
    for _ in range(steps):
        call()
    kms = (shard.info().scan_ms_total - a) / steps
    shard.set_profiling(False)
    if comm is not None:
        dev_ms = float(comm.allreduce_max(np.array([dev_ms]))[0])
        kms = float(comm.allreduce_max(np.array([kms]))[0])
    out.update({"ms_per_step": dev_ms / steps, "scan_kernel_ms": kms,
                ("rules_per_s" if kind == "scm" else "splits_per_s"): units * steps / (dev_ms * 1e-3),
                "scan_kernel_GBps_per_gpu": 8.0 * w_act * k_local / (kms * 1e-3) / 1e9,
                "frac_of_hbm_peak": 8.0 * w_act * k_local / (kms * 1e-3) / 1e9 / peak,
                "step_GBps_per_gpu": 8.0 * w_act * k_local * steps / (dev_ms * 1e-3) / 1e9,
                "step_frac_of_hbm_peak": 8.0 * w_act * k_local * steps / (dev_ms * 1e-3) / 1e9 / peak})
    if kind == "cart" and world == 1:
        # configs[2] in full: depth-10 tree with the experiments' occurrence tiebreaker (experiment_cart.py:82-94, 264)
        from kover_b200.learning.common.rules import LazyKmerRuleList
        from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
        from kover_b200.learning.learners.cart import DecisionTreeClassifier
        rules = LazyKmerRuleList(np.arange(k_local), np.arange(k_local))
        warm = OccurrenceTiebreaker(rc, train)
        DecisionTreeClassifier("gini", 3, 2, {0: 1.0, 1: 1.0}).fit(rules, rc, ex, tiebreaker=warm)
        warm.release()
        for attempt in ("first", "steady"):      # the first deep fit grows the pinned tie arena / sort buffers once
            t = DecisionTreeClassifier("gini", 10, 2, {0: 1.0, 1: 1.0})
            ia = shard.info()
            t0 = _t.perf_counter()
            tb = OccurrenceTiebreaker(rc, train)
            t.fit(rules, rc, ex, tiebreaker=tb)
            tb.release()
            out["depth10_fit_s" if attempt == "steady" else "depth10_first_fit_s"] = _t.perf_counter() - t0
        out["depth10_split_nodes"] = len(t.decision_tree.rules)
        ib = shard.info()
        out["depth10_matrix_passes"] = ib.scan_calls - ia.scan_calls
        # sibling subtraction (kv_gini_level_family): nodes whose class masks went through a matrix pass / nodes whose
        # counts were parent - sibling
        out["depth10_nodes_counted"] = ib.level_nodes_counted - ia.level_nodes_counted
        out["depth10_nodes_derived"] = ib.level_nodes_derived - ia.level_nodes_derived
    rc.close()
    return out


def config_extras(world, rank, local_rank, comm):
    """BASELINE.json's other configs, as far as this job's GPUs hold them -- untimed by the headline."""
    out = {}
    if world == 1:
        c1 = measure_shape("C1", 500, 1_000_000, 1, 0, local_rank, None, 71, "scm", steps=200)
        c1["note"] = "launch-bound: 64 MB of matrix is a 10 us scan; the step is the fixed cost of one launch + result poll"
        out["C1_500x1M_scm"] = c1
        out["C3_5kx20M_cart_root_node"] = measure_shape("C3", 5000, 20_000_000, 1, 0, local_rank, None, 73, "cart", steps=10)
        out["C4_shard_20kx12.5M_scm_(per_GPU_share_at_N=4)"] = measure_shape("C4", 20000, 12_500_000, 1, 0, local_rank, None,
                                                                             74, "scm", steps=10)
        out["C5_shard_50kx12.5M_scm_(per_GPU_share_at_N=8)"] = measure_shape("C5", 50000, 12_500_000, 1, 0, local_rank, None,
                                                                             75, "scm", steps=10)
    elif world in (2, 4):
        out["C4_20kx50M_scm_sharded"] = measure_shape("C4", 20000, 50_000_000 // world, world, rank, local_rank, comm, 74,
                                                      "scm", steps=10)
    elif world == 8:
        out["C5_50kx100M_scm_sharded"] = measure_shape("C5", 50000, 100_000_000 // world, world, rank, local_rank, comm, 75,
                                                       "scm", steps=10)
    return out


def hdf5_to_hbm(shard, n_genomes, k_cols, keep_percent=None):
    """north-star item (1): a Kover-layout HDF5 file (kmer_matrix uint64, chunks (1, 100000), gzip 4: create.py:222-230)
    of the first k_cols columns of the resident matrix, written with the test writer, then
    KmerRuleClassifications(dataset.kmer_matrix, n) timed with the chunks inflated on the device and on the host.
    keep_percent: zero all but that share of the words first (hash of (row, column)) -- the bench matrix is random
    bits and deflates to 0.87 of its size; real k-mer matrices are sparse and redundant."""
    import tempfile
    import time as _t
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hdf5_lite_writer import Writer
    from kover_b200.dataset import hdf5_lite
    from kover_b200.learning.common.rules import KmerRuleClassifications
    W = shard.n_words
    P = shard.read_block(0, W, 0, k_cols)
    if keep_percent is not None:
        h = (np.arange(k_cols, dtype=np.uint64)[None, :] * np.uint64(2654435761)
             + np.arange(W, dtype=np.uint64)[:, None] * np.uint64(40503)) >> np.uint64(7)
        P[(h % np.uint64(100)) >= np.uint64(keep_percent)] = 0
        del h
    d_tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(d_tmp, "bench.kover")
    t0 = _t.perf_counter()
    with Writer(path) as w:
        w.create_dataset("kmer_matrix", P, chunks=(1, 100000), compression=4)
    out = {"packed_GB": P.nbytes / 1e9, "file_GB": os.path.getsize(path) / 1e9, "write_s": _t.perf_counter() - t0,
           "layout": "uint64 [%d, %d], chunks (1, 100000), gzip 4" % (W, k_cols)}
    try:
        small = os.path.join(d_tmp, "warm.kover")                        # first-use costs out of the timing
        with Writer(small) as w:
            w.create_dataset("kmer_matrix", P[:, :200_000], chunks=(1, 100000), compression=4)
        for mode in ("1", "0"):
            os.environ["KOVER_B200_DEVICE_INFLATE"] = mode
            f = hdf5_lite.File(small)
            KmerRuleClassifications(f["kmer_matrix"], n_genomes, devices=[shard.device]).close()
            f.close()
        ok = True
        # (the device path twice, best kept: its set-up allocates ~file-sized device buffers, whose cost depends on
        #  what this process freed before -- a `kover learn` run uploads first, in a fresh context)
        for mode, key in (("1", "device_inflate"), ("1", "device_inflate"), ("0", "host_inflate")):
            os.environ["KOVER_B200_DEVICE_INFLATE"] = mode
            f = hdf5_lite.File(path)
            t0 = _t.perf_counter()
            rc = KmerRuleClassifications(f["kmer_matrix"], n_genomes, devices=[shard.device])
            dt = _t.perf_counter() - t0
            sh = rc._group.shards[0]
            if mode == "1":
                out.setdefault("device_inflate_runs", []).append(dict(sh.upload_deflate_stats(), constructor_s=dt))
            for c0 in (0, k_cols // 2, k_cols - 100_000):
                ok &= bool(np.array_equal(sh.read_block(0, W, c0, 100_000), P[:, c0:c0 + 100_000]))
            rc.close()
            f.close()
            if key + "_s" not in out or dt < out[key + "_s"]:
                out[key + "_s"] = dt
                out[key + "_GBps"] = P.nbytes / dt / 1e9
        out["resident_image_identical"] = ok
        out["hdf5_to_hbm_GBps"] = out["device_inflate_GBps"]
    finally:
        os.environ.pop("KOVER_B200_DEVICE_INFLATE", None)
        for name in os.listdir(d_tmp):
            os.remove(os.path.join(d_tmp, name))
        os.rmdir(d_tmp)
    return out


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: SCM utility scan, 2000 genomes x 10M k-mers per GPU (W=32 packed rows, "
                        "2.56 GB per GPU), p grid {0.1,0.5,1,2,5,10} x {conjunction,disjunction} x {train, 5 CV folds}",
            "n_genomes": N_GENOMES, "n_kmers_per_gpu": N_KMERS_PER_GPU, "n_kmers_total": N_KMERS_PER_GPU * n_gpus,
            "rules_per_step": 2 * N_KMERS_PER_GPU * n_gpus, "util_block_size": UTIL_BLOCK_SIZE,
            "labels": "genomes sorted by label (create.py:190-194)", "sharding": "k-mer columns, %d shard(s)" % n_gpus,
            "l2": "inputs (2.56 GB/GPU) are 20x larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one process per GPU)" % args.gpus)
    import torch
    from kover_b200 import _native
    from kover_b200.learning.common.rules import KmerRuleClassifications
    from kover_b200.sharding import Comm, ShardGroup

    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = Comm()

    k_local, k_total = N_KMERS_PER_GPU, N_KMERS_PER_GPU * world
    col0 = rank * k_local
    shard = _native.Shard(local_rank, N_GENOMES, k_total, col0, k_local)
    t0 = time.perf_counter()
    shard.generate_synthetic(SEED)
    gen_s = time.perf_counter() - t0
    slices = [(r * k_local, k_local) for r in range(world)]
    group = ShardGroup([shard], N_GENOMES, k_total, comm=comm, rank_slices=slices)

    class Meta(object):
        shape = ((N_GENOMES + 63) // 64, k_total)
        dtype = np.dtype(np.uint64)
        chunks = (1, 100000)
    rc = KmerRuleClassifications(Meta(), N_GENOMES, shard_group=group)

    steps = step_inputs(N_GENOMES, SEED)
    info0 = shard.info()

    def one_step(i):
        p, pos, neg = steps[i % len(steps)]
        return rc.best_utility_rules(p, pos, neg, util_block_size=UTIL_BLOCK_SIZE)

    def sync_all():
        if comm is not None:
            torch.cuda.synchronize()
            comm.barrier()
        torch.cuda.synchronize()

    # parity before timing, on every rank count: slices of this rank's shard and the merged result of a step
    _, pos0, neg0 = steps[0]
    parity_ok = parity_check(rc, comm, rank, N_GENOMES, k_local, k_total, SEED, pos0, neg0)

    for i in range(args.warmup):
        one_step(i)

    # ---- timed region 1: device time of K steps (CUDA events on the shard's stream), inputs resident.
    # Nothing but the steps runs inside it; the per-kernel times are accumulated by the library itself
    # (CUDA events around every scan launch) and read once at the end.
    n_blocks = _native.n_utility_blocks(k_total, UTIL_BLOCK_SIZE)
    timed = [steps[(args.warmup + i) % len(steps)] for i in range(args.steps)]
    alg_bytes = [8.0 * active_words(pos, neg, N_GENOMES) * k_local for _, pos, neg in timed]
    h2d = sum(active_words(pos, neg, N_GENOMES) * 20 for _, pos, neg in timed)
    tie_counts = []
    with ClockSampler(range(world), enabled=(rank == 0)) as clocks:
        sync_all()
        info_a = shard.info()
        shard.timer_start()
        wall0 = time.perf_counter()
        for p, pos, neg in timed:
            out = rc.best_utility_rules(p, pos, neg, util_block_size=UTIL_BLOCK_SIZE)
            tie_counts.append(len(out[1]))
        dev_ms = shard.timer_stop()
        sync_all()
        wall_s = time.perf_counter() - wall0
        info_b = shard.info()
        launches = info_b.total_launches - info_a.total_launches
        # the scan kernel's own duration: the same K steps once more with CUDA events around every scan launch
        # (kv_set_profiling; the events are off in the region above -- they cost host and stream time per step)
        shard.set_profiling(True)
        sync_all()
        info_c = shard.info()
        for p, pos, neg in timed:
            rc.best_utility_rules(p, pos, neg, util_block_size=UTIL_BLOCK_SIZE)
        sync_all()
        kernel_ms_total = shard.info().scan_ms_total - info_c.scan_ms_total
        shard.set_profiling(False)
        # keep the sampler alive for at least a few samples on very short runs: the SAME number of extra
        # (untimed) steps on every rank, since a step is collective
        w = float(comm.allreduce_max(np.array([wall_s]))[0]) if comm is not None else wall_s
        for i in range(max(0, int((0.6 - w) / max(w / args.steps, 1e-5)))):
            one_step(i)
    # results: the kernel writes a 64-byte header + 16 bytes per tie into mapped page-locked host memory (multi-rank:
    # + the merged block maxima / selected flags, 9 bytes per utility block)
    d2h = sum(64 + 16 * c + (9 * n_blocks if world > 1 else 0) for c in tie_counts)
    kernel_ms = [kernel_ms_total / args.steps] * args.steps
    if comm is not None:
        dev_ms = float(comm.allreduce_max(np.array([dev_ms]))[0])
        wall_s = float(comm.allreduce_max(np.array([wall_s]))[0])
        launches = int(comm.allreduce_sum(np.array([launches], dtype=np.int64))[0])

    # ---- timed region 2: end to end through the reference-shaped API, host arrays in and out
    from kover_b200.learning.learners.scm import SetCoveringMachine
    learners = {p: SetCoveringMachine(p=p) for p in P_GRID}
    sync_all()
    e0 = time.perf_counter()
    for i in range(args.steps):
        p, pos, neg = steps[(args.warmup + i) % len(steps)]
        res = learners[p]._get_best_utility_rules(rc, pos, neg, rule_blacklist=[])
        assert isinstance(res[1], np.ndarray)
    sync_all()
    e2e_s = time.perf_counter() - e0
    if comm is not None:
        e2e_s = float(comm.allreduce_max(np.array([e2e_s]))[0])

    rules_per_step = 2.0 * k_total
    value = rules_per_step * args.steps / (dev_ms * 1e-3)
    peak, peak_src = _peak_gbs()
    achieved = float(np.sum(alg_bytes) / (np.sum(kernel_ms) * 1e-3) / 1e9)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.isfile(tpath):          # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
        t = json.load(open(tpath))
        traffic = t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]
        traffic_src = t["source"]
    line = {
        "metric": "kmer_rules_scored_per_sec_per_scm_iteration", "value": value, "unit": "rules/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 popcount + f64 utility", "data": "synthetic (hash-generated on device, seed %d)" % SEED,
        "config": workload_config(world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "k_scan_tma<EPI_SCM_STREAM> (the whole step: scan + candidate filter + result)",
                     "note": "peak = measured COPY bandwidth (half reads, half writes); this kernel only reads, so a "
                             "fraction slightly above 1 is possible",
                     "peak_source": peak_src,
                     "bytes_per_launch": float(np.mean(alg_bytes)), "kernel_ms": float(np.mean(kernel_ms)),
                     "kernel_share_of_step": float(np.sum(kernel_ms) / dev_ms)},
        "e2e": {"value": rules_per_step * args.steps / e2e_s, "unit": "rules/s",
                "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps),
                "h2d_note": "the step's inputs are two index arrays (host) -> two row masks -> 20 B per active packed row, "
                            "which ride in the kernel parameter block of the ONE launch per step",
                "d2h_note": "the scan kernel itself writes header + tie records into mapped page-locked host memory; "
                            "the host polls the epoch flag (no copy call)",
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "SetCoveringMachine._get_best_utility_rules(rule_classifications, pos_idx, neg_idx) with host "
                       "index arrays in, host tie arrays out; matrix resident (uploaded once per dataset)"},
        "gpu_launches": int(launches),
        "steps_streamed": int(info_b.stream_steps - info_a.stream_steps),           # rank 0: steps that wrote no count
        "steps_redone_with_counts": int(info_b.stream_redos - info_a.stream_redos),  # arrays / had to be redone with them
        "parity_checked": bool(parity_ok),
        "clocks": clocks.summary(),
        "wall_ms_per_step": 1e3 * wall_s / args.steps,
        "generate_s": gen_s,
    }
    line["roofline"]["step_GBps"] = float(np.sum(alg_bytes) / (dev_ms * 1e-3) / 1e9)
    line["roofline"]["step_frac"] = line["roofline"]["step_GBps"] / peak
    if world > 1:
        m = max(1, info_b.exch_steps - info_a.exch_steps)          # over the K timed steps
        line["exchange"] = {"steps": int(info_b.exch_steps - info_a.exch_steps),
                            "publish_us": (info_b.exch_publish_us - info_a.exch_publish_us) / m,
                            "wait_for_peers_us": (info_b.exch_wait_us - info_a.exch_wait_us) / m,
                            "merge_and_result_us": (info_b.exch_merge_us - info_a.exch_merge_us) / m,
                            "note": "rank 0, device clock of its last CTA, per step: packet stores into the peers' mailboxes; "
                                    "waiting for every peer's packet (NVLink latency + rank skew); merge + result to the host"}
    if rank == 0 and world == 1:
        line["extra"] = extra_measurements(rc, shard, steps, k_local)
        if not args.no_extras:
            line["extra"]["hdf5_to_hbm"] = hdf5_to_hbm(shard, N_GENOMES, 8_400_000)
            line["extra"]["hdf5_to_hbm_sparse_5pct"] = hdf5_to_hbm(shard, N_GENOMES, 8_400_000, keep_percent=5)
        line["cpu_baseline"] = cpu_baseline(N_GENOMES, SEED, steps)
    rc.close()
    if not args.no_extras:
        cfg = config_extras(world, rank, local_rank, comm)
        if rank == 0:
            line.setdefault("extra", {})["configs"] = cfg
    if rank == 0:
        print(json.dumps(line))
    if comm is not None:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=72)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the untimed extras (other configs, HDF5 upload)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
