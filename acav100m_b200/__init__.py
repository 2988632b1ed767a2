"""acav100m_b200 -- Blackwell (sm_100a) implementation of ACAV100M's two GPU-bound curation stages.

Only the hot path lives here: the mini-batch SGD k-means operator
(``acav100m_b200.clustering``) and the exact greedy mutual-information selection
(``acav100m_b200.subset_selection``), both thin Python mirrors of the reference's operator API on
top of the C-ABI CUDA library ``libacav_b200.so`` (``include/acav_b200.h``).  There is no CPU
fallback: every operator raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
