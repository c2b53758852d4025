"""Test helper: a dataset object in h5py's style, backed by the built-in reader."""
from kover_b200.dataset import hdf5_lite


class _H5pyLikeFile(object):
    def __init__(self, filename):
        self.filename = filename


class H5pyLikeDataset(object):
    """What KmerRuleClassifications sees when a Kover installation hands it an h5py dataset: shape / dtype / chunks /
    name / file.filename and slicing that inflates on the host -- and no way to get at the raw chunks."""

    def __init__(self, path, name, shape=None):
        self._d = hdf5_lite.File(path)[name]
        self.file, self.name = _H5pyLikeFile(path), "/" + name.lstrip("/")
        self.shape, self.dtype, self.chunks = (shape or self._d.shape), self._d.dtype, self._d.chunks
        self.reads = 0

    def __getitem__(self, key):
        self.reads += 1
        return self._d[key]
