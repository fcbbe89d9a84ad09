import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from clique_b200 import Aligner, AffineScoring, Reference, ReferenceManager
from clique_b200.aligner import pack_reads
import _oracle as O
al = Aligner(device=0, max_reads=1024, max_read_bytes=1<<22, cigar_ops_per_read=64, n_slots=1)
cases = [(b"AAAANAAAA", b"AAAAAAAA", (6.0,-6.0,5.0,-10.0,-10.0,1.0)), (b"ACGTACGTAC", b"ACGTTACGTAC", (10.0,-9.0,9.0,-20.0,-2.0,1.0))]
for ref, read, sc in cases:
    for gen in (1, 0):
        al.set_option("force_generic", gen)
        al.set_references(ReferenceManager([Reference(ref, b"r")]))
        rb, ro = pack_reads([read])
        for so in (False, True):
            br = al.align_batch(rb, ro, AffineScoring(*sc), "fixed", "maxlen", fixed_ref=[0], score_only=so)
            print("generic" if gen else "fast", "score_only" if so else "tb", int(br.score_scaled[0]), hex(int(br.score_scaled[0]) & 0xffffffff), br.cigar_string(0), int(br.status[0]))
    print("oracle", O.align_pair(ref, read, sc, "maxlen")["score"])
