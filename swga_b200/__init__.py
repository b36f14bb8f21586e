"""swga_b200 -- B200 (sm_100a) implementation of swga's data-parallel hot path: all-k k-mer
counting (behind swga/kmers.py), binding-site location (swga/locate.py) and batched primer-set
scoring (swga/score.py, swga/sets.py).  Host code is Python over one ctypes C-ABI shared library,
libswga_b200.so (include/swga_b200.h); there is no CPU fallback."""
__version__ = "0.1.0"
