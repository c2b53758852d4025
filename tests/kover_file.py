"""Test helper: write a Kover-layout HDF5 dataset (reference dataset/create.py:160-240, split.py:150-230) with this
package's minimal writer.  The risk tables come from the CPU checker (oracle/), so the helper needs no GPU."""
import numpy as np

from hdf5_lite_writer import Writer
from kover_b200.utils import _minimum_uint_size
from oracle import kover_oracle as ko


def write_kover_file(path, packed, n_genomes, labels, train_idx, test_idx, folds, seed=42, chunks=(1, 1000), gzip=4,
                     fixed_strings=False, names=None):
    """packed: uint64/uint32 [W, K]; folds: list of (train_idx, test_idx); names: optional (kmer_sequences in file
    order, kmer_by_matrix_column, genome_identifiers) -- generated from `seed` when omitted."""
    k = packed.shape[1]
    rc = ko.OracleKmerRuleClassifications(packed, n_genomes, fast=True)
    w = Writer(path)
    w.attrs("/", {"created": 1.0, "uuid": "00000000-test", "genome_source_type": "tsv", "genomic_data": "synthetic.tsv",
                  "phenotype_description": "synthetic phenotype", "phenotype_metadata_source": "synthetic.tsv",
                  "compression": "gzip (level %d)" % gzip, "classification_type": "binary"})
    w.create_dataset("kmer_matrix", packed, chunks=chunks, compression=gzip)
    seqs = ["".join("ACGT"[(i >> (2 * j)) & 3] for j in range(12)) for i in range(k)]
    perm = np.random.RandomState(seed).permutation(k)            # kmer_by_matrix_column: column -> index in kmer_sequences
    ordered = [None] * k
    for col, seq_idx in enumerate(perm):
        ordered[seq_idx] = seqs[col]
    ids = ["genome_%d" % i for i in range(n_genomes)]
    if names is not None:
        ordered, perm, ids = names
        perm = np.asarray(perm)
        seqs = [ordered[perm[col]] for col in range(k)]
    if fixed_strings:
        w.create_dataset("kmer_sequences", np.array(ordered, dtype="S12"), chunks=(max(1, min(k, 4096)),), compression=gzip)
        w.create_dataset("genome_identifiers", np.array(ids, dtype="S16"))
    else:
        w.create_dataset("kmer_sequences", np.array(ordered, dtype=object), chunks=(max(1, min(k, 4096)),), compression=gzip)
        w.create_dataset("genome_identifiers", np.array(ids, dtype=object))
    w.create_dataset("kmer_by_matrix_column", perm.astype(_minimum_uint_size(k)), chunks=(max(1, min(k, 4096)),),
                     compression=gzip)
    w.create_dataset("phenotype", np.asarray(labels, dtype=np.uint8), attrs={"description": "synthetic phenotype"})
    w.create_dataset("phenotype_tags", np.array(["0", "1"], dtype=object))

    def put(group, tr, te):
        idx_dtype = _minimum_uint_size(n_genomes)
        w.create_dataset(group + "/train_genome_idx", np.sort(tr).astype(idx_dtype))
        w.create_dataset(group + "/test_genome_idx", np.sort(te).astype(idx_dtype))
        ur, bk, ba = ko.kmer_risk_tables(rc, tr[labels[tr] == 1], tr[labels[tr] == 0])
        w.create_dataset(group + "/unique_risks", ur)
        w.create_dataset(group + "/unique_risk_by_kmer", bk)
        w.create_dataset(group + "/unique_risk_by_anti_kmer", ba)

    g = "splits/split1"
    w.attrs(g, {"random_seed": np.int64(seed), "n_folds": np.int64(len(folds)),
                "train_proportion": 1.0 * len(train_idx) / n_genomes, "test_proportion": 1.0 * len(test_idx) / n_genomes})
    put(g, np.asarray(train_idx), np.asarray(test_idx))
    for i, (tr, te) in enumerate(folds):
        put(g + "/folds/fold_%d" % (i + 1), np.asarray(tr), np.asarray(te))
    w.close()
    return seqs, perm
