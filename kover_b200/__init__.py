"""kover_b200 -- Kover's rule-scoring hot path (SCM utility scan, CART split scan) on B200.

The package mirrors the import paths of the reference for the pieces on that path:

    kover.learning.common.popcount  ->  kover_b200.learning.common.popcount
    kover.learning.common.rules     ->  kover_b200.learning.common.rules
    kover.learning.learners.scm     ->  kover_b200.learning.learners.scm
    kover.learning.learners.cart    ->  kover_b200.learning.learners.cart
    kover.utils (pack/unpack)       ->  kover_b200.utils

All compute goes through libkoverb200.so (hand-written sm_100a CUDA, C ABI in
include/kover_b200.h).  There is no CPU fallback: importing works anywhere, but the first call
that needs the device raises if the library or a GPU is missing.
"""
__version__ = "0.1.0"
