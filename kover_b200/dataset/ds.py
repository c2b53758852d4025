"""Read access to a Kover dataset file (interface of reference dataset/ds.py:27-196).

The file is the reference's HDF5 layout, unchanged (dataset/create.py:160-240, split.py:150-230): root attributes
(uuid, compression, classification_type, ...), ``kmer_matrix`` (packed uint64 / uint32, chunked + gzip),
``kmer_sequences``, ``kmer_by_matrix_column``, ``genome_identifiers``, ``phenotype``, ``phenotype_tags`` and
``splits/<name>[/folds/<fold>]`` with the example indices and the individual k-mer risk tables.  It is opened
with h5py when h5py is importable (its chunk cache disabled like utils.py:58-73), otherwise with this package's
minimal reader (hdf5_lite).  Everything here is host-side I/O; the matrix reaches HBM through
``KmerRuleClassifications`` (learning/common/rules.py), which streams it in chunk-row slabs.
"""
from functools import partial

import numpy as np


def _open(path):
    try:
        import h5py
    except ImportError:
        from . import hdf5_lite
        return hdf5_lite.File(path, "r")
    return h5py.File(path, "r", rdcc_nbytes=0)        # no chunk cache: every chunk is read exactly once per upload


class KoverDatasetPhenotype(object):
    def __init__(self, description, tags, metadata, metadata_source):
        self.description, self.tags, self.metadata, self.metadata_source = description, tags, metadata, metadata_source


class KoverDatasetFold(object):
    def __init__(self, name, train_genome_idx, test_genome_idx, unique_risks, unique_risk_by_kmer,
                 unique_risk_by_anti_kmer):
        self.name = name
        self.train_genome_idx, self.test_genome_idx = train_genome_idx, test_genome_idx
        self.unique_risks = unique_risks
        self.unique_risk_by_kmer, self.unique_risk_by_anti_kmer = unique_risk_by_kmer, unique_risk_by_anti_kmer


class KoverDatasetSplit(KoverDatasetFold):
    def __init__(self, name, train_proportion, test_proportion, train_genome_idx, test_genome_idx, unique_risks,
                 unique_risk_by_kmer, unique_risk_by_anti_kmer, folds, random_seed):
        super(KoverDatasetSplit, self).__init__(name, train_genome_idx, test_genome_idx, unique_risks,
                                                unique_risk_by_kmer, unique_risk_by_anti_kmer)
        self.train_proportion, self.test_proportion = train_proportion, test_proportion
        self.folds, self.random_seed = folds, random_seed

    def __str__(self):
        return "%s   Train genomes: %d (%.3f)   Test genomes: %d (%.3f)   Folds: %d   Random Seed: %d" % (
            self.name, len(self.train_genome_idx), self.train_proportion, len(self.test_genome_idx),
            self.test_proportion, len(self.folds), self.random_seed)


class KoverDataset(object):
    def __init__(self, file):
        self.path = file
        self.dataset_open = partial(_open, file)
        self._handle = None

    def _file(self):
        # one handle per KoverDataset (the reference reopens the file on every property access, ds.py:33-124)
        if self._handle is None:
            self._handle = self.dataset_open()
        return self._handle

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def _attr(self, name, default=None):
        attrs = self._file().attrs
        try:
            return attrs[name]
        except KeyError:
            if default is not None:
                return default
            raise

    @property
    def classification_type(self):
        return self._attr("classification_type", "binary")      # pre-2.0.0 datasets have no such attribute

    @property
    def compression(self):
        return self._attr("compression")

    @property
    def kmer_filter(self):
        return self._attr("filter")

    @property
    def genome_count(self):
        return self._file()["genome_identifiers"].shape[0]

    @property
    def genome_identifiers(self):
        return self._file()["genome_identifiers"]

    @property
    def genome_source(self):
        return self._attr("genomic_data")

    @property
    def genome_source_type(self):
        return self._attr("genome_source_type")

    @property
    def kmer_by_matrix_column(self):
        return self._file()["kmer_by_matrix_column"]

    @property
    def kmer_count(self):
        return self._file()["kmer_sequences"].shape[0]

    @property
    def kmer_length(self):
        return len(self._file()["kmer_sequences"][0])

    @property
    def kmer_matrix(self):
        return self._file()["kmer_matrix"]

    @property
    def kmer_sequences(self):
        return self._file()["kmer_sequences"]

    @property
    def phenotype(self):
        f = self._file()
        try:
            description = f.attrs["phenotype_description"]
        except KeyError:
            description = f.attrs["phenotype_name"]              # pre-2.0.0
        tags = f["phenotype_tags"] if "phenotype_tags" in f else np.array(["0", "1"])
        return KoverDatasetPhenotype(description=description, tags=tags, metadata=f["phenotype"],
                                     metadata_source=f.attrs["phenotype_metadata_source"])

    @property
    def splits(self):
        f = self._file()
        if "splits" not in f:
            return []
        return [self.get_split(name) for name in f["splits"].keys()]

    @property
    def uuid(self):
        return self._attr("uuid")

    def get_split(self, name):
        split = self._file()["splits"][name]
        folds = []
        if "folds" in split:
            # group iteration order = name order (fold_1, fold_10, fold_2, ...): it fixes the order in which the fold
            # risks are averaged (ds.py:142-147, experiment_scm.py:188)
            for fold_name, fold in split["folds"].items():
                folds.append(KoverDatasetFold(fold_name, fold["train_genome_idx"], fold["test_genome_idx"],
                                              fold["unique_risks"], fold["unique_risk_by_kmer"],
                                              fold["unique_risk_by_anti_kmer"]))
        test_proportion = split.attrs["test_proportion"] if "test_proportion" in split.attrs \
            else 1.0 - split.attrs["train_proportion"]
        return KoverDatasetSplit(name, split.attrs["train_proportion"], test_proportion, split["train_genome_idx"],
                                 split["test_genome_idx"], split["unique_risks"], split["unique_risk_by_kmer"],
                                 split["unique_risk_by_anti_kmer"], folds, split.attrs["random_seed"])
