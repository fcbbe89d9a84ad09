"""The N>1 path on CPU: two gloo ranks shard a batch the way bench.py / ShardedAligner do (contiguous read ranges balanced
by bases, no data-path collective), align their shard with the CPU oracle standing in for the GPU, and the gathered,
re-based results must equal the single-process run.  Exercises shard_bounds, concat_results and the rendezvous."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from clique_b200 import synth
    from clique_b200.aligner import BatchResult, concat_results, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = synth.config_c2(301)
    rb, ro = O.pack_seqs(c["refs"])
    bounds = shard_bounds(c["read_off"], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    off = c["read_off"][lo:hi + 1] - c["read_off"][lo]
    qb = c["read_bytes"][int(c["read_off"][lo]):int(c["read_off"][hi])]
    if rank == 1 and len(qb):
        # this rank ships its shard 2-bit packed (clq_pack2, host only): the packed form of a shard is self-contained (its own
        # base offsets from 0, whatever byte of the whole stream it starts at) and expands to the same bytes
        from clique_b200 import pack_reads_2bit
        qb = qb.copy()
        qb[5::97] = ord("N")                       # bytes outside the 2-bit alphabet travel in the exception list
        c["read_bytes"][int(c["read_off"][lo]):int(c["read_off"][hi])] = qb   # (the full run on rank 0 repeats this edit below)
        pk = pack_reads_2bit(qb, int(off[-1]))
        assert pk.words.nbytes <= (len(qb) + 15) // 16 * 4 + 4 and len(pk.exc_pos) > 0
        qb = pk.unpack()
    out = O.align_batch(rb, ro, qb if len(qb) else np.zeros(1, np.uint8), off, c["scoring"], search="fixed",
                        fixed_ref=c["fixed_ref"][lo:hi], band_mode="readlen", threads=2)
    mine = BatchResult(1, out["score"].astype(np.int64), out["ref_index"], out["cigar_off"].astype(np.uint32), out["cigar_len"],
                       out["status"], out["cigar_pool"])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # test plumbing only: the product has no collective
    dist.barrier()
    if rank == 0:
        allr = concat_results(gathered)
        l1, h1 = int(c["read_off"][bounds[1]]), int(c["read_off"][bounds[2]])
        c["read_bytes"][l1:h1][5::97] = ord("N")   # rank 1's shard as rank 1 aligned it
        full = O.align_batch(rb, ro, c["read_bytes"], c["read_off"], c["scoring"], search="fixed", fixed_ref=c["fixed_ref"],
                             band_mode="readlen", threads=2)
        ok = (allr.score_scaled == full["score"]).all() and (allr.cigar_len == full["cigar_len"]).all()
        for i in range(len(full["score"])):
            o, l = int(full["cigar_off"][i]), int(full["cigar_len"][i])
            ok = ok and np.array_equal(allr.cigar(i), full["cigar_pool"][o:o + l])
        # shards are balanced by bases within one read length
        sizes = [int(c["read_off"][bounds[k + 1]] - c["read_off"][bounds[k]]) for k in range(world)]
        ok = ok and bounds[0] == 0 and bounds[-1] == 301 and max(sizes) - min(sizes) <= 2 * 300
        with open(tmp, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_bounds_edges():
    sys.path.insert(0, ROOT)
    from clique_b200.aligner import shard_bounds
    off = np.array([0, 10, 20, 30, 40], dtype=np.uint64)
    assert shard_bounds(off, 1) == [0, 4]
    assert shard_bounds(off, 2) == [0, 2, 4]
    assert shard_bounds(off, 4) == [0, 1, 2, 3, 4]
    b = shard_bounds(off, 8)
    assert b[0] == 0 and b[-1] == 4 and all(b[i] <= b[i + 1] for i in range(8))
    assert shard_bounds(np.array([0], dtype=np.uint64), 3) == [0, 0, 0, 0]


def test_shard_bounds_balance_cells_not_bytes():
    """contiguous ranges balanced by sum(L1 * L2).  On length-sorted mixed input (C5: 300 bp reads then
    5 kb reads, each against its own amplicon) a byte split would give the last shard ~16x the cells per byte of the first."""
    sys.path.insert(0, ROOT)
    from clique_b200.aligner import shard_bounds
    lens = np.concatenate([np.full(1000, 300), np.full(1000, 5000)])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    fixed = np.concatenate([np.zeros(1000, np.int32), np.ones(1000, np.int32)])
    b = shard_bounds(off, 4, [300, 5000], fixed)
    cells = np.where(fixed == 0, 300, 5000) * lens
    share = [cells[b[i]:b[i + 1]].sum() / cells.sum() for i in range(4)]
    assert max(share) - min(share) < 0.01, share
    assert b[1] > 1000  # the whole block of short reads is far less than a quarter of the work
