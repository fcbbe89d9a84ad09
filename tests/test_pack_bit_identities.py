"""CPU checks of the packed (s16x2) direction-bit arithmetic of clique_b200/csrc/clq_pack.cuh::pack_row_step.

The PACK kernel derives the `ext2` bit ("the F layer of this cell extends the gap", three_way_max_and_direction's
Left-over-Up / Diag-over-Left tie order, alignment/alignment_matrix.rs:671-683 as used by update_3d_score :618-665) from
one relu-clamped DPX instruction.  The argument that this is exact -- no borrow crosses the two 16-bit halves, the clamp
yields the bit itself -- is integer arithmetic that can be replayed on the CPU with numpy, which is what this file does;
the GPU parity tests then pin the kernel as compiled.  Nothing here touches the oracle or the CUDA library."""
import numpy as np
import pytest

M32 = np.uint64(0xFFFFFFFF)


def dup16(v):
    return (np.uint64(v & 0xFFFF) * np.uint64(0x10001)) & M32


def halves(w):
    """signed 16-bit halves (lo, hi) of packed 32-bit words"""
    w = np.asarray(w, np.uint64)
    lo = (w & np.uint64(0xFFFF)).astype(np.int64)
    hi = ((w >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
    return np.where(lo >= 32768, lo - 65536, lo), np.where(hi >= 32768, hi - 65536, hi)


def pack(lo, hi):
    return ((np.asarray(hi, np.int64) & 0xFFFF).astype(np.uint64) << np.uint64(16)) | (np.asarray(lo, np.int64) & 0xFFFF).astype(np.uint64)


def wrap16(v):
    v = np.asarray(v, np.int64) & 0xFFFF
    return np.where(v >= 32768, v - 65536, v)


def viaddmax_s16x2(a, b, c):
    """__viaddmax_s16x2: per half max(a + b, c), signed"""
    al, ah = halves(a); bl, bh = halves(b); cl, ch = halves(c)
    return pack(np.maximum(wrap16(al + bl), cl), np.maximum(wrap16(ah + bh), ch))


def viaddmin_s16x2_relu(a, b, c):
    """__viaddmin_s16x2_relu: per half max(min(a + b, c), 0), signed"""
    al, ah = halves(a); bl, bh = halves(b); cl, ch = halves(c)
    return pack(np.maximum(np.minimum(wrap16(al + bl), cl), 0), np.maximum(np.minimum(wrap16(ah + bh), ch), 0))


def vminu2(a, b):
    a = np.asarray(a, np.uint64); b = np.asarray(b, np.uint64)
    lo = np.minimum(a & np.uint64(0xFFFF), b & np.uint64(0xFFFF))
    hi = np.minimum(a >> np.uint64(16), b >> np.uint64(16))
    return (hi << np.uint64(16)) | lo


def sub32(a, b):
    return (np.asarray(a, np.uint64) - np.asarray(b, np.uint64)) & M32


ONE = np.uint64(0x00010001)

# (gap_open + gap_extend, gap_extend) scaled: the CLI set, default_dna x2, merger, a unit-extend set, a large-open set
GAPS = [(-22, -2), (-21, -1), (-16, -1), (-6, -1), (-40, -3), (-2, -1)]


def _left_cells(rng, n, x1):
    """Random left-neighbour states (M, E, F) of both reads as the kernel holds them: biased values in [64, 32767], E and F
    additionally stored shifted by -x1 (so E - x1, F - x1 must also fit), M possibly the boundary sentinel 0."""
    top = 32767 + x1  # x1 < 0: E - x1 <= 32767
    base = rng.integers(64, top - 200, size=(2, n))
    spread = lambda: rng.integers(-60, 61, size=(2, n))
    M = np.clip(base + spread(), 64, top)
    E = np.clip(base + spread(), 64, top)
    F = np.clip(base + spread(), 64, top)
    M = np.where(rng.random((2, n)) < 0.05, 0, M)               # matrix boundary: M = MAX_NEG -> 0
    return M, E, F


@pytest.mark.parametrize("x1,le", GAPS)
def test_ext2_relu_clamp_is_the_bit(x1, le):
    rng = np.random.default_rng(1000 - x1 * 7 - le)
    n = 200000
    M, E, F = _left_cells(rng, n, x1)
    # force exact ties F + le == E + x1 and F + le == M + x1 on a share of the cells, independently per half
    r = rng.random((2, n))
    F = np.where(r < 0.10, E + x1 - le, F)
    F = np.where((r >= 0.10) & (r < 0.20) & (M > 0), M + x1 - le, F)
    F = np.clip(F, 64, 32767 + x1)
    want = ((F + le >= E + x1) & (F + le > M + x1)).astype(np.int64)            # update_3d_score's Left layer takes the extension
    Fh, Ehl, Ml = pack(F[0] - x1, F[1] - x1), pack(E[0] - x1, E[1] - x1), pack(M[0], M[1])
    LE, X1M1 = dup16(le), dup16(x1 - 1)
    t2 = viaddmax_s16x2(Ehl, X1M1, Ml)                                            # max(E - 1, M)
    # the session-2 form: second max, subtract, unsigned min
    u2 = viaddmax_s16x2(Fh, LE, t2)
    old = vminu2(sub32(u2, t2), ONE)
    # the shipped form: LE - t2 as ONE 32-bit subtraction, then a relu-clamped add
    t2l, t2h = halves(t2)
    d = sub32(LE, t2)
    dl, dh = halves(d)
    assert np.array_equal(dl, wrap16(le - t2l)) and np.array_equal(dh, wrap16(le - t2h)), "a borrow crossed the halves"
    new = viaddmin_s16x2_relu(Fh, d, ONE)
    ol, oh = halves(old); nl, nh = halves(new)
    assert np.array_equal(ol, want[0]) and np.array_equal(oh, want[1])
    assert np.array_equal(nl, want[0]) and np.array_equal(nh, want[1])


@pytest.mark.parametrize("x1,le", GAPS)
def test_ext2_from_left_cell_B_and_eP(x1, le):
    """DESIGN.md 'next experiment': max(E_left - 1, M_left) can be replaced under the comparison by B_left - eP_left, both of
    which the left cell already produced (B = max(M, E, F), eP = [E > max(M, F)]), because F + (le - x1) > F always."""
    rng = np.random.default_rng(77 - x1 - le)
    n = 200000
    M, E, F = _left_cells(rng, n, x1)
    r = rng.random((2, n))
    F = np.where(r < 0.10, E + x1 - le, F)
    F = np.where((r >= 0.10) & (r < 0.20) & (M > 0), M + x1 - le, F)
    F = np.clip(F, 64, 32767 + x1)
    want = ((F + le >= E + x1) & (F + le > M + x1)).astype(np.int64)
    B = np.maximum(np.maximum(M, E), F)
    eP = (E > np.maximum(M, F)).astype(np.int64)
    # as the F step holds them: Fh = F_left - x1 (shifted storage), Bl = B_left unshifted (Fhn = max(Fh + le, Bl))
    Fh, Bl, e = pack(F[0] - x1, F[1] - x1), pack(B[0], B[1]), pack(eP[0], eP[1])
    r2 = sub32(Bl, e)                                    # B - eP per half: B >= 64, no borrow
    d = sub32(dup16(le), r2)
    got = viaddmin_s16x2_relu(Fh, d, ONE)                # clamp(F - x1 + le - (B - eP), 0, 1)
    gl, gh = halves(got)
    assert np.array_equal(gl, want[0]) and np.array_equal(gh, want[1])


def byte_perm(x, y, sel):
    """__byte_perm(x, y, s): byte k of the result = byte s[k] of the 8-byte value {y, x} (x = bytes 0..3, y = bytes 4..7)"""
    src = [(int(x) >> (8 * i)) & 0xFF for i in range(4)] + [(int(y) >> (8 * i)) & 0xFF for i in range(4)]
    return sum(src[(sel >> (4 * k)) & 7] << (8 * k) for k in range(4))


def test_nibble_packing_matches_the_walkers_layout():
    """pack_row_step shifts four clamp results per cell pair into `nib`, four nibble pairs into acc0 / acc1 and unzips them with
    two PRMTs; walk_kernel reads column j of a word as (w >> (28 - 4 * (j & 7))) & 15 = [ext1 ext2 eP fM]."""
    rng = np.random.default_rng(5)
    for _ in range(200):
        flags = rng.integers(0, 2, size=(8, 4, 2))            # [column][ext1, ext2, eP, fM][read A, read B]
        acc = [0, 0]
        for jj in range(8):
            nib = 0
            for k in range(4):
                nib = (nib * 2 + (int(flags[jj, k, 0]) | (int(flags[jj, k, 1]) << 16))) & 0xFFFFFFFF
            a = acc[jj >> 2]
            a = nib if (jj & 3) == 0 else (a * 16 + nib) & 0xFFFFFFFF
            acc[jj >> 2] = a
        wA = byte_perm(acc[1], acc[0], 0x5410)                # low halves: read A
        wB = byte_perm(acc[1], acc[0], 0x7632)                # high halves: read B
        for jj in range(8):
            for h, w in ((0, wA), (1, wB)):
                nibble = (w >> (28 - 4 * jj)) & 15
                want = sum(int(flags[jj, k, h]) << (3 - k) for k in range(4))
                assert nibble == want, (jj, h)


@pytest.mark.parametrize("x1,le", GAPS)
def test_other_three_bits(x1, le):
    """ext1 / eP / fM as pack_row_step derives them: min_u16x2(x - y, 1) with x >= y per half (no borrow), against the tie order of
    three_way_max_and_direction (alignment/alignment_matrix.rs:671-683: Diag > Left > Up on ties)."""
    rng = np.random.default_rng(9 - x1)
    n = 100000
    top = 32767 + x1
    base = rng.integers(200, top - 200, size=(2, n))
    sp = lambda: rng.integers(-60, 61, size=(2, n))
    E_up, B_up = np.clip(base + sp(), 64, top), np.clip(base + sp(), 64, top)     # upper neighbour: E layer and best
    M, F = np.clip(base + sp(), 64, top), np.clip(base + sp(), 64, top)           # this cell: M and F (true values)
    r = rng.random((2, n))
    E_up = np.where(r < 0.15, B_up + x1 - le, E_up)                                # tie: extension == opening
    F = np.where((r > 0.8), M, F)                                                  # tie: F == M
    EhU, BU = pack(E_up[0] - x1, E_up[1] - x1), pack(B_up[0], B_up[1])
    LE, X1 = dup16(le), dup16(x1)
    Ehn = viaddmax_s16x2(EhU, LE, BU)                                              # E - x1 = max(E_up + le - x1, B_up)
    E = np.maximum(E_up + le, B_up + x1)
    el, eh = halves(Ehn)
    assert np.array_equal(el, E[0] - x1) and np.array_equal(eh, E[1] - x1)
    ext1 = vminu2(sub32(Ehn, BU), ONE)                                             # E extends: E_up + le > B_up + x1
    l, h = halves(ext1)
    assert np.array_equal(l, (E_up[0] + le > B_up[0] + x1).astype(np.int64)) and np.array_equal(h, (E_up[1] + le > B_up[1] + x1).astype(np.int64))
    Fhn, Mv = pack(F[0] - x1, F[1] - x1), pack(M[0], M[1])
    Pv = viaddmax_s16x2(Fhn, X1, Mv)                                               # max(F, M)
    Bn = viaddmax_s16x2(Ehn, X1, Pv)                                               # max(E, F, M)
    fM = vminu2(sub32(Pv, Mv), ONE)                                                # F > M   (tie: Diag)
    eP = vminu2(sub32(Bn, Pv), ONE)                                                # E > max(M, F)   (tie: not Up)
    l, h = halves(fM)
    assert np.array_equal(l, (F[0] > M[0]).astype(np.int64)) and np.array_equal(h, (F[1] > M[1]).astype(np.int64))
    l, h = halves(eP)
    mf = np.maximum(M, F)
    assert np.array_equal(l, (E[0] > mf[0]).astype(np.int64)) and np.array_equal(h, (E[1] > mf[1]).astype(np.int64))
    bl, bh = halves(Bn)
    assert np.array_equal(bl, np.maximum(E, mf)[0]) and np.array_equal(bh, np.maximum(E, mf)[1])
