"""clique_b200 -- B200-native drop-in for the alignment hot path of mckennalab/clique.

The product is libclq.so (hand-written sm_100a CUDA kernels behind the C ABI of include/clq.h).  This package is
the thin Python host layer over that ABI, mirroring the reference's Rust call surface (AffineScoring,
ReferenceManager, align_two_strings, align_to_reference_choices, exhaustive/quick search, AlignmentResult).
There is no CPU fallback: importing works anywhere, running an alignment needs the built library and a GPU.
"""
from ._lib import (CLQ_OK, READ_TOO_LONG, SCORING_NOT_REPRESENTABLE, TRACEBACK_DIVERGED, CIGAR_POOL_FULL, NO_CANDIDATE,
                   ClqError, Limits, load_library, library_path)
from .aligner import (AffineScoring, ConvexScoring, TwoPieceScoring, RustBioScoring, AlignmentResult, AlignmentWithRef, Aligner, BatchResult, Reference,
                      ReferenceManager, ShardedAligner, cigar_to_string, PackedReads, pack_reads_2bit)

__all__ = ["AffineScoring", "ConvexScoring", "TwoPieceScoring", "RustBioScoring", "AlignmentResult", "AlignmentWithRef", "Aligner", "BatchResult", "Reference",
           "ReferenceManager", "ShardedAligner", "cigar_to_string", "PackedReads", "pack_reads_2bit", "ClqError", "Limits", "load_library",
           "library_path", "CLQ_OK", "READ_TOO_LONG", "SCORING_NOT_REPRESENTABLE", "TRACEBACK_DIVERGED",
           "CIGAR_POOL_FULL", "NO_CANDIDATE"]
