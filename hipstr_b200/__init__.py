"""hipstr_b200: B200-native (sm_100a) implementation of HipSTR's read x haplotype HMM
alignment and genotype-posterior hot path.  The product is the C-ABI shared library
`libhipstr_b200.so` (include/hipstr_b200.h); this package is a thin ctypes driver."""
from .capi import (AlignBatch, BatchBuilder, Context, EmBatch, Genotyper, HipstrError, LeftAligned, Synth, em_train,  # noqa: F401
                   load, load_synth, make_em_batch, make_locus_reads)

__version__ = "0.1.0"
