/*
 * clq_oracle.c -- CPU restatement of clique's Gotoh alignment hot path (see clq_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 * All file:line citations are relative to /root/reference/rust_cmd/src/.
 *
 * The f64 functions (orc_*) follow the reference line by line and keep its data layout: a column-major
 * 3-layer f64 score array plus a column-major 3-layer direction array whose element is as fat as the
 * reference's `AlignmentDirection` enum (it carries the Inv(loc, loc, move) variant: 48 bytes), one
 * scratch matrix per worker thread, fill AND traceback for every candidate reference.  That is what the
 * cpu_baseline times.  The orci_* functions restate the same recurrence on scaled integers in compact
 * arrays; tests prove both give identical results.
 */
#define _GNU_SOURCE
#include "clq_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define MAX_NEG_SCORE (-100000.0) /* alignment/alignment_matrix.rs:34 */
#define FASTA_UNSET '-'
#define FASTA_N 'N'

/* ------------------------------------------------------------------------------------------------
 * matrix: Alignment<Ix3>, alignment/alignment_matrix.rs:219-233
 * ---------------------------------------------------------------------------------------------- */
enum { DIR_UP = 0, DIR_LEFT = 1, DIR_DIAG = 2 }; /* AlignmentDirection::{Up,Left,Diag}; zero() = Up(0) */

typedef struct {
    uint64_t tag;
    uint64_t size;
    uint64_t inv_payload[4]; /* room the reference's Inv(AlignmentLocation, AlignmentLocation, InvMove) takes */
} orc_dir_t;

struct orc_matrix {
    size_t a, b;      /* shape (a, b, 3), column-major: index = x + a*(y + b*z) */
    double* scores;
    orc_dir_t* tb;
};

static inline size_t IDX(const orc_matrix_t* m, size_t x, size_t y, size_t z) { return x + m->a * (y + m->b * z); }

orc_matrix_t* orc_matrix_create(size_t dim_a, size_t dim_b) {
    orc_matrix_t* m = (orc_matrix_t*)calloc(1, sizeof(*m));
    if (!m) return NULL;
    m->a = dim_a;
    m->b = dim_b;
    m->scores = (double*)calloc(dim_a * dim_b * 3, sizeof(double));
    m->tb = (orc_dir_t*)calloc(dim_a * dim_b * 3, sizeof(orc_dir_t));
    if (!m->scores || !m->tb) { orc_matrix_free(m); return NULL; }
    return m;
}

void orc_matrix_free(orc_matrix_t* m) {
    if (!m) return;
    free(m->scores);
    free(m->tb);
    free(m);
}

/* alignment/scoring_functions.rs:100-102 */
double orc_match_mismatch(const orc_affine_t* sc, uint8_t a, uint8_t b) {
    if (a == FASTA_N || b == FASTA_N || a < 58 || b < 58) return sc->special_character_score;
    if (a == b) return sc->match_score;
    return sc->mismatch_score;
}

/* alignment/alignment_matrix.rs:671-683: strict '>' => ties go Diag > Left > Up */
double orc_three_way_max(double up, double left, double diag, int* dir) {
    if (up > left) {
        if (up > diag) { *dir = DIR_UP; return up; }
        *dir = DIR_DIAG; return diag;
    } else if (left > diag) {
        *dir = DIR_LEFT; return left;
    }
    *dir = DIR_DIAG; return diag;
}

/* band of row x, alignment/alignment_matrix.rs:413-417 -- the centre is computed in f64 */
void orc_band(size_t x, size_t l1, size_t l2, size_t bandwidth, int64_t* lo, int64_t* hi) {
    int64_t y_bounds = (int64_t)(((double)x / (double)(l1 + 1)) * (double)(l2 + 1));
    int64_t bw = (int64_t)bandwidth;
    int64_t a = y_bounds - bw, b = y_bounds + bw;
    *lo = a > 1 ? a : 1;
    *hi = b < (int64_t)l2 + 1 ? b : (int64_t)l2 + 1;
}

/* update_3d_score, alignment/alignment_matrix.rs:618-665 */
static inline void update_3d_score(orc_matrix_t* m, const uint8_t* s1, size_t l1, const uint8_t* s2, size_t l2,
                                   const orc_affine_t* sc, size_t x, size_t y) {
    double gap_multiplier = (x == l1 || y == l2) ? sc->final_gap_multiplier : 1.0;
    double x1 = sc->gap_open + (sc->gap_extend * gap_multiplier);
    double local_gap_ext = sc->gap_extend * gap_multiplier;
    int d;
    {
        double ms = orc_match_mismatch(sc, s1[x - 1], s2[y - 1]);
        double v = orc_three_way_max(m->scores[IDX(m, x - 1, y - 1, 1)] + ms, m->scores[IDX(m, x - 1, y - 1, 2)] + ms,
                                     m->scores[IDX(m, x - 1, y - 1, 0)] + ms, &d);
        m->scores[IDX(m, x, y, 0)] = v;
        m->tb[IDX(m, x, y, 0)].tag = d;
        m->tb[IDX(m, x, y, 0)].size = 1;
    }
    {
        double v = orc_three_way_max(m->scores[IDX(m, x - 1, y, 1)] + local_gap_ext, m->scores[IDX(m, x - 1, y, 2)] + x1,
                                     m->scores[IDX(m, x - 1, y, 0)] + x1, &d);
        m->scores[IDX(m, x, y, 1)] = v;
        m->tb[IDX(m, x, y, 1)].tag = d;
        m->tb[IDX(m, x, y, 1)].size = 1;
    }
    {
        double v = orc_three_way_max(m->scores[IDX(m, x, y - 1, 1)] + x1, m->scores[IDX(m, x, y - 1, 2)] + local_gap_ext,
                                     m->scores[IDX(m, x, y - 1, 0)] + x1, &d);
        m->scores[IDX(m, x, y, 2)] = v;
        m->tb[IDX(m, x, y, 2)].tag = d;
        m->tb[IDX(m, x, y, 2)].size = 1;
    }
}

/* perform_affine_alignment_bandwidth, alignment/alignment_matrix.rs:376-425 */
int orc_fill(orc_matrix_t* m, const uint8_t* s1, size_t l1, const uint8_t* s2, size_t l2, const orc_affine_t* sc,
             size_t bandwidth) {
    if (!(m->a > l1) || !(m->b > l2)) return -1; /* :381-383 asserts */

    m->scores[IDX(m, 0, 0, 0)] = 0.0;
    m->scores[IDX(m, 0, 0, 1)] = MAX_NEG_SCORE;
    m->scores[IDX(m, 0, 0, 2)] = MAX_NEG_SCORE;
    /* T[0,0,*] is never written by the reference: fresh matrix => Up(0) */
    for (int z = 0; z < 3; z++) { m->tb[IDX(m, 0, 0, z)].tag = DIR_UP; m->tb[IDX(m, 0, 0, z)].size = 0; }

    for (size_t x = 1; x < l1 + 1; x++) { /* first column, :389-396 */
        double g = (sc->gap_open + ((double)x * sc->gap_extend)) * sc->final_gap_multiplier;
        m->scores[IDX(m, x, 0, 0)] = MAX_NEG_SCORE;
        m->scores[IDX(m, x, 0, 1)] = g;
        m->scores[IDX(m, x, 0, 2)] = g;
        for (int z = 0; z < 3; z++) { m->tb[IDX(m, x, 0, z)].tag = DIR_UP; m->tb[IDX(m, x, 0, z)].size = 1; }
    }
    for (size_t y = 1; y < l2 + 1; y++) { /* top row, :398-405 */
        double g = (sc->gap_open + ((double)y * sc->gap_extend)) * sc->final_gap_multiplier;
        m->scores[IDX(m, 0, y, 0)] = MAX_NEG_SCORE;
        m->scores[IDX(m, 0, y, 1)] = g;
        m->scores[IDX(m, 0, y, 2)] = g;
        for (int z = 0; z < 3; z++) { m->tb[IDX(m, 0, y, z)].tag = DIR_LEFT; m->tb[IDX(m, 0, y, z)].size = 1; }
    }

    for (size_t x = 1; x < l1 + 1; x++) { /* :413-424 */
        int64_t lo, hi;
        orc_band(x, l1, l2, bandwidth, &lo, &hi);
        /* canonicalisation: cells this row's band skips hold the fresh-matrix state (0.0, Up(0)) */
        for (int64_t y = 1; y <= (int64_t)l2; y++) {
            if (y >= lo && y < hi) { y = hi - 1; continue; }
            for (int z = 0; z < 3; z++) {
                m->scores[IDX(m, x, (size_t)y, z)] = 0.0;
                m->tb[IDX(m, x, (size_t)y, z)].tag = DIR_UP;
                m->tb[IDX(m, x, (size_t)y, z)].size = 0;
            }
        }
        for (int64_t y = lo; y < hi; y++) update_3d_score(m, s1, l1, s2, l2, sc, x, (size_t)y);
    }
    return 0;
}

/* simplify_cigar_string, alignment_manager.rs:386-423 (duplicate alignment_functions.rs:874-911) */
size_t orc_simplify_cigar(const uint32_t* in, size_t n, uint32_t* out) {
    size_t k = 0;
    int have = 0;
    uint32_t last = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t tok = in[i];
        if (!have) { last = tok; have = 1; }
        else if ((last & 0xF) == (tok & 0xF)) last = (((last >> 4) + (tok >> 4)) << 4) | (tok & 0xF);
        else { out[k++] = last; last = tok; }
    }
    if (have) out[k++] = last;
    return k;
}

/* get_reference_alignment_rate, consensus/consensus_builders.rs:288-307 */
double orc_alignment_rate(const uint8_t* ref_aligned, const uint8_t* read_aligned, size_t n, uint32_t* matches, uint32_t* mismatches) {
    uint32_t m = 0, mm = 0;
    for (size_t i = 0; i < n; i++) {
        if (ref_aligned[i] > 64 && ref_aligned[i] != FASTA_N && read_aligned[i] > 64) {
            if (ref_aligned[i] == read_aligned[i]) m++; else mm++;
        }
    }
    if (matches) *matches = m;
    if (mismatches) *mismatches = mm;
    return (double)m / (double)(m + mm);
}

/* perform_3d_global_traceback (global), alignment/alignment_matrix.rs:941-1086 */
int orc_traceback(orc_matrix_t* m, const uint8_t* s1, size_t l1, const uint8_t* s2, size_t l2, orc_result_t* res,
                  uint32_t* cigar, size_t cigar_cap, uint8_t* ref_aligned, uint8_t* read_aligned, size_t aligned_cap) {
    size_t x = l1, y = l2;
    size_t cap = l1 + l2 + 2;
    uint32_t* cig = (uint32_t*)malloc(cap * sizeof(uint32_t));
    uint8_t* a1 = (uint8_t*)malloc(cap);
    uint8_t* a2 = (uint8_t*)malloc(cap);
    size_t nc = 0, na = 0, npath = 0;
    res->status = ORC_OK;

    /* :963-972  max_by(partial_cmp) returns the LAST maximum: layer 2 > 1 > 0 on ties */
    size_t z = 0;
    double best = m->scores[IDX(m, x, y, 0)];
    for (size_t k = 1; k < 3; k++) {
        double v = m->scores[IDX(m, x, y, k)];
        if (!(v < best)) { best = v; z = k; }
    }
    double score = m->scores[IDX(m, x, y, z)];

    while (x > 0 && y > 0) { /* :977 */
        m->scores[IDX(m, x, y, 0)] = 0.0; /* :979-981 -- the reference zeroes the path as it goes */
        m->scores[IDX(m, x, y, 1)] = 0.0;
        m->scores[IDX(m, x, y, 2)] = 0.0;
        npath++;
        orc_dir_t d = m->tb[IDX(m, x, y, z)];
        size_t next_z = d.tag == DIR_DIAG ? 0 : (d.tag == DIR_UP ? 1 : 2); /* :986-989 */
        size_t size = (size_t)d.size;
        if (size == 0) { /* stale Up(0): no movement, z -> 1, forever.  Report instead of hanging. */
            res->status = ORC_TRACEBACK_DIVERGED;
            break;
        }
        switch (z) { /* :1019-1049 */
            case 0:
                cig[nc++] = (1u << 4) | ORC_OP_M;
                for (size_t i = 0; i < size; i++) { a1[na] = s1[x - 1]; a2[na] = s2[y - 1]; na++; x--; y--; }
                break;
            case 1:
                cig[nc++] = (1u << 4) | ORC_OP_D;
                for (size_t i = 0; i < size; i++) { a1[na] = s1[x - 1]; a2[na] = FASTA_UNSET; na++; x--; }
                break;
            default:
                cig[nc++] = (1u << 4) | ORC_OP_I;
                for (size_t i = 0; i < size; i++) { a1[na] = FASTA_UNSET; a2[na] = s2[y - 1]; na++; y--; }
                break;
        }
        z = next_z;
    }
    if (res->status == ORC_OK) {
        while (x > 0) { a1[na] = s1[x - 1]; a2[na] = FASTA_UNSET; na++; x--; cig[nc++] = (1u << 4) | ORC_OP_D; } /* :1054-1059 */
        while (y > 0) { a1[na] = FASTA_UNSET; a2[na] = s2[y - 1]; na++; y--; cig[nc++] = (1u << 4) | ORC_OP_I; } /* :1060-1065 */
    }
    /* reverse everything, :1066-1069 */
    for (size_t i = 0; i < nc / 2; i++) { uint32_t t = cig[i]; cig[i] = cig[nc - 1 - i]; cig[nc - 1 - i] = t; }
    uint32_t* merged = (uint32_t*)malloc((nc + 1) * sizeof(uint32_t));
    size_t nm = orc_simplify_cigar(cig, nc, merged);

    orc_alignment_rate(a1, a2, na, &res->matches, &res->mismatches); /* order-independent: the strings are still reversed */
    res->score = score;
    res->n_cigar = (uint32_t)nm;
    res->aligned_len = (uint32_t)na;
    res->path_len = (uint32_t)npath;
    int rc = 0;
    if (cigar) {
        if (nm > cigar_cap) { res->status = ORC_CIGAR_POOL_FULL; rc = -2; }
        else memcpy(cigar, merged, nm * sizeof(uint32_t));
    }
    if (ref_aligned && read_aligned && na <= aligned_cap) {
        for (size_t i = 0; i < na; i++) { ref_aligned[i] = a1[na - 1 - i]; read_aligned[i] = a2[na - 1 - i]; }
    }
    free(merged); free(cig); free(a1); free(a2);
    return rc;
}

/* align_two_strings, alignment_manager.rs:231-273: fresh matrix, perform_affine_alignment (:366-372) */
int orc_align_two_strings(const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2, const orc_affine_t* sc,
                          orc_result_t* res, uint32_t* cigar, size_t cigar_cap, uint8_t* ref_aligned,
                          uint8_t* read_aligned, size_t aligned_cap) {
    orc_matrix_t* m = orc_matrix_create(l1 + 1, l2 + 1);
    if (!m) return -1;
    size_t bw = l1 > l2 ? l1 : l2;
    int rc = orc_fill(m, ref, l1, read, l2, sc, bw);
    if (rc == 0) rc = orc_traceback(m, ref, l1, read, l2, res, cigar, cigar_cap, ref_aligned, read_aligned, aligned_cap);
    orc_matrix_free(m);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * scaled-int scoring
 * ---------------------------------------------------------------------------------------------- */
static int is_integral(double v) { return v == floor(v) && fabs(v) < 1e9; }

int orc_affine_to_int(const orc_affine_t* sc, orc_affine_int_t* out) {
    for (int scale = 1; scale <= 64; scale *= 2) {
        double s = (double)scale;
        double v[10];
        v[0] = sc->match_score * s;
        v[1] = sc->mismatch_score * s;
        v[2] = sc->special_character_score * s;
        v[3] = (sc->gap_open + sc->gap_extend) * s;
        v[4] = sc->gap_extend * s;
        v[5] = (sc->gap_open + sc->gap_extend * sc->final_gap_multiplier) * s;
        v[6] = (sc->gap_extend * sc->final_gap_multiplier) * s;
        v[7] = (sc->gap_open * sc->final_gap_multiplier) * s;
        v[8] = (sc->gap_extend * sc->final_gap_multiplier) * s;
        v[9] = MAX_NEG_SCORE * s;
        int ok = 1;
        for (int i = 0; i < 10; i++) ok = ok && is_integral(v[i]);
        if (!ok) continue;
        out->scale = scale;
        out->match = (int32_t)v[0]; out->mismatch = (int32_t)v[1]; out->special = (int32_t)v[2];
        out->oe_in = (int32_t)v[3]; out->e_in = (int32_t)v[4];
        out->oe_fin = (int32_t)v[5]; out->e_fin = (int32_t)v[6];
        out->b0 = (int32_t)v[7]; out->b1 = (int32_t)v[8];
        out->max_neg = (int32_t)v[9];
        return ORC_OK;
    }
    return ORC_SCORING_NOT_REPRESENTABLE;
}

/* ------------------------------------------------------------------------------------------------
 * compact scaled-int restatement (same recurrence, same tie rules, same band, same traceback)
 * ---------------------------------------------------------------------------------------------- */
static inline int64_t tw_i(int64_t up, int64_t left, int64_t diag, uint8_t* dir) {
    if (up > left) {
        if (up > diag) { *dir = DIR_UP; return up; }
        *dir = DIR_DIAG; return diag;
    } else if (left > diag) { *dir = DIR_LEFT; return left; }
    *dir = DIR_DIAG; return diag;
}

static inline int is_special(uint8_t c) { return c == FASTA_N || c < 58; }

int orci_align_pair(const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2, const orc_affine_int_t* sc,
                    int band_mode, size_t band_k, int64_t* score_scaled, int32_t* status, uint32_t* n_cigar,
                    uint32_t* cigar, size_t cigar_cap, int want_traceback) {
    size_t W = l2 + 1;
    size_t bw = band_mode == ORC_BAND_MAXLEN ? (l1 > l2 ? l1 : l2) : (band_mode == ORC_BAND_READLEN ? l2 : band_k);
    int32_t* S = (int32_t*)malloc((l1 + 1) * W * 3 * sizeof(int32_t)); /* [x][y][z] row-major */
    uint8_t* T = (uint8_t*)malloc((l1 + 1) * W * 3);                   /* dir | size<<2 */
    if (!S || !T) { free(S); free(T); return -1; }
#define SI(x, y, z) S[((x) * W + (y)) * 3 + (z)]
#define TI(x, y, z) T[((x) * W + (y)) * 3 + (z)]
    SI(0, 0, 0) = 0; SI(0, 0, 1) = sc->max_neg; SI(0, 0, 2) = sc->max_neg;
    TI(0, 0, 0) = TI(0, 0, 1) = TI(0, 0, 2) = DIR_UP;
    for (size_t x = 1; x <= l1; x++) {
        int32_t g = sc->b0 + (int32_t)x * sc->b1;
        SI(x, 0, 0) = sc->max_neg; SI(x, 0, 1) = g; SI(x, 0, 2) = g;
        TI(x, 0, 0) = TI(x, 0, 1) = TI(x, 0, 2) = DIR_UP | 4;
    }
    for (size_t y = 1; y <= l2; y++) {
        int32_t g = sc->b0 + (int32_t)y * sc->b1;
        SI(0, y, 0) = sc->max_neg; SI(0, y, 1) = g; SI(0, y, 2) = g;
        TI(0, y, 0) = TI(0, y, 1) = TI(0, y, 2) = DIR_LEFT | 4;
    }
    for (size_t x = 1; x <= l1; x++) {
        int64_t lo, hi;
        orc_band(x, l1, l2, bw, &lo, &hi);
        for (int64_t yy = 1; yy <= (int64_t)l2; yy++) {
            size_t y = (size_t)yy;
            if (yy < lo || yy >= hi) {
                SI(x, y, 0) = SI(x, y, 1) = SI(x, y, 2) = 0;
                TI(x, y, 0) = TI(x, y, 1) = TI(x, y, 2) = DIR_UP; /* size 0 */
                continue;
            }
            int fin = (x == l1 || y == l2);
            int64_t x1 = fin ? sc->oe_fin : sc->oe_in, le = fin ? sc->e_fin : sc->e_in;
            uint8_t a = ref[x - 1], b = read[y - 1], d;
            int64_t ms = (is_special(a) || is_special(b)) ? sc->special : (a == b ? sc->match : sc->mismatch);
            SI(x, y, 0) = (int32_t)tw_i(SI(x - 1, y - 1, 1) + ms, SI(x - 1, y - 1, 2) + ms, SI(x - 1, y - 1, 0) + ms, &d);
            TI(x, y, 0) = d | 4;
            SI(x, y, 1) = (int32_t)tw_i(SI(x - 1, y, 1) + le, SI(x - 1, y, 2) + x1, SI(x - 1, y, 0) + x1, &d);
            TI(x, y, 1) = d | 4;
            SI(x, y, 2) = (int32_t)tw_i(SI(x, y - 1, 1) + x1, SI(x, y - 1, 2) + le, SI(x, y - 1, 0) + x1, &d);
            TI(x, y, 2) = d | 4;
        }
    }
    size_t x = l1, y = l2, z = 0;
    int32_t best = SI(x, y, 0);
    for (size_t k = 1; k < 3; k++) if (SI(x, y, k) >= best) { best = SI(x, y, k); z = k; }
    *score_scaled = best;
    *status = ORC_OK;
    *n_cigar = 0;
    int rc = 0;
    if (want_traceback) {
        uint32_t* cig = (uint32_t*)malloc((l1 + l2 + 2) * sizeof(uint32_t));
        size_t nc = 0;
        while (x > 0 && y > 0) {
            uint8_t t = TI(x, y, z);
            if (!(t & 4)) { *status = ORC_TRACEBACK_DIVERGED; break; }
            size_t nz = (t & 3) == DIR_DIAG ? 0 : ((t & 3) == DIR_UP ? 1 : 2);
            if (z == 0) { cig[nc++] = (1u << 4) | ORC_OP_M; x--; y--; }
            else if (z == 1) { cig[nc++] = (1u << 4) | ORC_OP_D; x--; }
            else { cig[nc++] = (1u << 4) | ORC_OP_I; y--; }
            z = nz;
        }
        if (*status == ORC_OK) {
            while (x > 0) { cig[nc++] = (1u << 4) | ORC_OP_D; x--; }
            while (y > 0) { cig[nc++] = (1u << 4) | ORC_OP_I; y--; }
        }
        for (size_t i = 0; i < nc / 2; i++) { uint32_t t = cig[i]; cig[i] = cig[nc - 1 - i]; cig[nc - 1 - i] = t; }
        uint32_t* merged = (uint32_t*)malloc((nc + 1) * sizeof(uint32_t));
        size_t nm = orc_simplify_cigar(cig, nc, merged);
        *n_cigar = (uint32_t)nm;
        if (cigar) {
            if (nm > cigar_cap) { *status = ORC_CIGAR_POOL_FULL; rc = -2; }
            else memcpy(cigar, merged, nm * sizeof(uint32_t));
        }
        free(merged); free(cig);
    }
#undef SI
#undef TI
    free(S); free(T);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * "convex" = two-piece affine global alignment (definition owned by this repo; PARITY UNPINNED: the reference has
 * no convex DP, only the never-called ConvexScoring::gap).
 *
 * A gap of length k costs w(k) = max(o1 + k*e1, o2 + k*e2) (o_i < 0).  Five states per cell:
 *   M (diagonal), E1/E2 (gap consuming the reference = Del, piece 1/2), F1/F2 (gap consuming the read = Ins).
 *   A(cell) = argmax over (M, F1, F2, E1, E2) scanned in that order with strict '>', so earlier states win ties;
 *   B(cell) = that maximum.
 *   M [x,y] = B(x-1,y-1) + m(x,y)                                   came from A(x-1,y-1)
 *   Ei[x,y] = max(Ei[x-1,y] + ei, B(x-1,y) + oi + ei)               extends iff strictly greater, else from A(x-1,y)
 *   Fi[x,y] = max(Fi[x,y-1] + ei, B(x,y-1) + oi + ei)               extends iff strictly greater, else from A(x,y-1)
 * i.e. the affine recurrence's "any state may open any gap" connectivity (alignment/alignment_matrix.rs:618-665),
 * the same MAX_NEG sentinel and boundary shape: S[0,0] = (0, NEG, NEG, NEG, NEG); S[k,0] and S[0,k] hold NEG in M and
 * oi + k*ei in Ei and Fi.  No band, no final-gap multiplier.  Score = B(L1,L2), start state = A(L1,L2); the walk and
 * the trailing boundary run follow perform_3d_global_traceback (:977-1065).
 * ---------------------------------------------------------------------------------------------- */
enum { CV_M = 0, CV_E1 = 1, CV_E2 = 2, CV_F1 = 3, CV_F2 = 4 };

static inline int32_t cv_argmax(const int32_t* s5, uint8_t* who) {
    static const uint8_t order[5] = {CV_M, CV_F1, CV_F2, CV_E1, CV_E2};
    int32_t best = s5[CV_M];
    *who = CV_M;
    for (int i = 1; i < 5; i++) if (s5[order[i]] > best) { best = s5[order[i]]; *who = order[i]; }
    return best;
}

int orc_convex_align_pair(const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2, const orc_convex_t* sc,
                          int64_t* score, int32_t* status, uint32_t* n_cigar, uint32_t* cigar, size_t cigar_cap,
                          int want_traceback) {
    size_t W = l2 + 1;
    int32_t* S = (int32_t*)malloc((l1 + 1) * W * 5 * sizeof(int32_t));
    uint8_t* T = (uint8_t*)malloc((l1 + 1) * W * 5); /* source state per layer; 8 | state = "extended" */
    if (!S || !T) { free(S); free(T); return -1; }
#define SI(x, y, z) S[((x) * W + (y)) * 5 + (z)]
#define TI(x, y, z) T[((x) * W + (y)) * 5 + (z)]
    const int32_t o[2] = {sc->o1, sc->o2}, e[2] = {sc->e1, sc->e2};
    for (int z = 0; z < 5; z++) { SI(0, 0, z) = z == CV_M ? 0 : sc->max_neg; TI(0, 0, z) = CV_M; }
    for (size_t x = 1; x <= l1; x++) {
        SI(x, 0, CV_M) = sc->max_neg;
        SI(x, 0, CV_E1) = SI(x, 0, CV_F1) = sc->o1 + (int32_t)x * sc->e1;
        SI(x, 0, CV_E2) = SI(x, 0, CV_F2) = sc->o2 + (int32_t)x * sc->e2;
        for (int z = 0; z < 5; z++) TI(x, 0, z) = CV_M;
    }
    for (size_t y = 1; y <= l2; y++) {
        SI(0, y, CV_M) = sc->max_neg;
        SI(0, y, CV_E1) = SI(0, y, CV_F1) = sc->o1 + (int32_t)y * sc->e1;
        SI(0, y, CV_E2) = SI(0, y, CV_F2) = sc->o2 + (int32_t)y * sc->e2;
        for (int z = 0; z < 5; z++) TI(0, y, z) = CV_M;
    }
    for (size_t x = 1; x <= l1; x++) {
        for (size_t y = 1; y <= l2; y++) {
            uint8_t a = ref[x - 1], b = read[y - 1], who;
            int32_t ms = (is_special(a) || is_special(b)) ? sc->special : (a == b ? sc->match : sc->mismatch);
            int32_t bd = cv_argmax(&SI(x - 1, y - 1, 0), &who);
            SI(x, y, CV_M) = bd + ms;
            TI(x, y, CV_M) = who;
            int32_t bu = cv_argmax(&SI(x - 1, y, 0), &who);
            for (int p = 0; p < 2; p++) {
                int zz = p == 0 ? CV_E1 : CV_E2;
                int32_t ext = SI(x - 1, y, zz) + e[p], opn = bu + o[p] + e[p];
                if (ext > opn) { SI(x, y, zz) = ext; TI(x, y, zz) = 8 | (uint8_t)zz; }
                else { SI(x, y, zz) = opn; TI(x, y, zz) = who; }
            }
            int32_t bl = cv_argmax(&SI(x, y - 1, 0), &who);
            for (int p = 0; p < 2; p++) {
                int zz = p == 0 ? CV_F1 : CV_F2;
                int32_t ext = SI(x, y - 1, zz) + e[p], opn = bl + o[p] + e[p];
                if (ext > opn) { SI(x, y, zz) = ext; TI(x, y, zz) = 8 | (uint8_t)zz; }
                else { SI(x, y, zz) = opn; TI(x, y, zz) = who; }
            }
        }
    }
    size_t x = l1, y = l2;
    uint8_t zst;
    int32_t best = cv_argmax(&SI(x, y, 0), &zst);
    size_t z = zst;
    *score = best;
    *status = ORC_OK;
    *n_cigar = 0;
    int rc = 0;
    if (want_traceback) {
        uint32_t* cig = (uint32_t*)malloc((l1 + l2 + 2) * sizeof(uint32_t));
        size_t nc = 0;
        while (x > 0 && y > 0) {
            size_t nz = TI(x, y, z) & 7;
            if (z == CV_M) { cig[nc++] = (1u << 4) | ORC_OP_M; x--; y--; }
            else if (z == CV_E1 || z == CV_E2) { cig[nc++] = (1u << 4) | ORC_OP_D; x--; }
            else { cig[nc++] = (1u << 4) | ORC_OP_I; y--; }
            z = nz;
        }
        while (x > 0) { cig[nc++] = (1u << 4) | ORC_OP_D; x--; }
        while (y > 0) { cig[nc++] = (1u << 4) | ORC_OP_I; y--; }
        for (size_t i = 0; i < nc / 2; i++) { uint32_t t = cig[i]; cig[i] = cig[nc - 1 - i]; cig[nc - 1 - i] = t; }
        uint32_t* merged = (uint32_t*)malloc((nc + 1) * sizeof(uint32_t));
        size_t nm = orc_simplify_cigar(cig, nc, merged);
        *n_cigar = (uint32_t)nm;
        if (cigar) {
            if (nm > cigar_cap) { *status = ORC_CIGAR_POOL_FULL; rc = -2; }
            else memcpy(cigar, merged, nm * sizeof(uint32_t));
        }
        free(merged); free(cig);
    }
#undef SI
#undef TI
    free(S); free(T);
    return rc;
}

/* ConvexScoring::gap, alignment/scoring_functions.rs:50-52 */
double orc_convex_gap(double gap_open, size_t length) { return gap_open + log10((double)length); }

/* ------------------------------------------------------------------------------------------------
 * k-mer index: sequence_to_kmers + unique_kmers, reference/fasta_reference.rs:159-202
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* p; /* points into an upper-cased copy */
    uint32_t ref;
    uint32_t count;
} kmer_ent_t;

struct orc_kmer_index {
    uint32_t k, skip, n_refs, n;
    uint8_t* keys;  /* n * k bytes, sorted */
    uint32_t* owner; /* n */
};

static uint32_t g_cmp_k;
static int kmer_cmp(const void* a, const void* b) {
    const kmer_ent_t* x = (const kmer_ent_t*)a;
    const kmer_ent_t* y = (const kmer_ent_t*)b;
    int c = memcmp(x->p, y->p, g_cmp_k);
    if (c) return c;
    return x->ref < y->ref ? -1 : (x->ref > y->ref);
}

static inline uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* sequence_to_kmers (:159-167): uppercase, windows(k).step_by(skip), consecutive dedup_with_count.
 * Calls cb(kmer, run_count) for each run. */
typedef void (*kmer_cb)(const uint8_t* kmer, uint32_t count, void* ctx);
static void sequence_to_kmers(const uint8_t* up, size_t len, uint32_t k, uint32_t skip, kmer_cb cb, void* ctx) {
    if (len < k || k == 0 || skip == 0) return;
    const uint8_t* run = NULL;
    uint32_t cnt = 0;
    for (size_t pos = 0; pos + k <= len; pos += skip) {
        const uint8_t* w = up + pos;
        if (run && memcmp(run, w, k) == 0) cnt++;
        else {
            if (run) cb(run, cnt, ctx);
            run = w;
            cnt = 1;
        }
    }
    if (run) cb(run, cnt, ctx);
}

typedef struct { kmer_ent_t* v; size_t n, cap; uint32_t ref; } ent_vec_t;
static void push_ent(const uint8_t* kmer, uint32_t count, void* ctx) {
    ent_vec_t* ev = (ent_vec_t*)ctx;
    if (ev->n == ev->cap) { ev->cap = ev->cap ? ev->cap * 2 : 1024; ev->v = (kmer_ent_t*)realloc(ev->v, ev->cap * sizeof(kmer_ent_t)); }
    ev->v[ev->n].p = kmer; ev->v[ev->n].ref = ev->ref; ev->v[ev->n].count = count; ev->n++;
}

orc_kmer_index_t* orc_kmer_index_build(uint32_t n_refs, const uint8_t* ref_bytes, const uint64_t* ref_off, uint32_t k,
                                       uint32_t skip) {
    orc_kmer_index_t* ix = (orc_kmer_index_t*)calloc(1, sizeof(*ix));
    ix->k = k; ix->skip = skip; ix->n_refs = n_refs;
    size_t total = (size_t)ref_off[n_refs];
    uint8_t* up = (uint8_t*)malloc(total + 1);
    for (size_t i = 0; i < total; i++) up[i] = upper(ref_bytes[i]);
    ent_vec_t ev = {0};
    for (uint32_t r = 0; r < n_refs; r++) {
        ev.ref = r;
        sequence_to_kmers(up + ref_off[r], (size_t)(ref_off[r + 1] - ref_off[r]), k, skip, push_ent, &ev);
    }
    g_cmp_k = k;
    qsort(ev.v, ev.n, sizeof(kmer_ent_t), kmer_cmp);
    ix->keys = (uint8_t*)malloc((ev.n + 1) * (size_t)k);
    ix->owner = (uint32_t*)malloc((ev.n + 1) * sizeof(uint32_t));
    /* unique <=> total run count over all references == 1 (:173-178, :187) */
    for (size_t i = 0; i < ev.n;) {
        size_t j = i;
        uint64_t tot = 0;
        while (j < ev.n && memcmp(ev.v[j].p, ev.v[i].p, k) == 0) { tot += ev.v[j].count; j++; }
        if (tot == 1) {
            memcpy(ix->keys + (size_t)ix->n * k, ev.v[i].p, k);
            ix->owner[ix->n] = ev.v[i].ref;
            ix->n++;
        }
        i = j;
    }
    free(ev.v);
    free(up);
    return ix;
}

void orc_kmer_index_free(orc_kmer_index_t* ix) {
    if (!ix) return;
    free(ix->keys); free(ix->owner); free(ix);
}

uint32_t orc_kmer_index_size(const orc_kmer_index_t* ix) { return ix->n; }

typedef struct { const orc_kmer_index_t* ix; uint32_t* votes; uint32_t total; } vote_ctx_t;
static void vote_cb(const uint8_t* kmer, uint32_t count, void* ctx) {
    (void)count; /* the run count `_c` is ignored, alignment_functions.rs:705-709 */
    vote_ctx_t* vc = (vote_ctx_t*)ctx;
    const orc_kmer_index_t* ix = vc->ix;
    size_t lo = 0, hi = ix->n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        int c = memcmp(ix->keys + mid * ix->k, kmer, ix->k);
        if (c == 0) { vc->votes[ix->owner[mid]]++; vc->total++; return; }
        if (c < 0) lo = mid + 1; else hi = mid;
    }
}

uint32_t orc_kmer_votes(const orc_kmer_index_t* ix, const uint8_t* read, size_t l2, uint32_t* votes) {
    uint8_t* up = (uint8_t*)malloc(l2 + 1);
    for (size_t i = 0; i < l2; i++) up[i] = upper(read[i]);
    vote_ctx_t vc = {ix, votes, 0};
    sequence_to_kmers(up, l2, ix->k, ix->skip, vote_cb, &vc);
    free(up);
    return vc.total;
}

/* ------------------------------------------------------------------------------------------------
 * batch driver
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t* v;
    size_t n, cap;
} u32vec_t;

typedef struct {
    const orc_batch_t* in;
    const orc_affine_t* sc;
    orc_batch_out_t* out;
    const orc_kmer_index_t* kix;
    atomic_uint_fast64_t* next;
    atomic_uint_fast64_t* cells;
    size_t max_l1, max_l2;
    u32vec_t cig;        /* thread-local cigar store */
    uint64_t* tl_off;    /* shared: per-read offset into its thread's store */
    uint32_t* tl_owner;  /* shared: per-read thread id */
    uint32_t tid;
} worker_t;

static size_t band_of(const orc_batch_t* in, size_t l1, size_t l2) {
    if (in->band_mode == ORC_BAND_MAXLEN) return l1 > l2 ? l1 : l2;
    if (in->band_mode == ORC_BAND_READLEN) return l2;
    return (size_t)in->band_k;
}

/* one (read, reference) pair into the thread's scratch matrix: align_two_strings_passed_matrix,
 * alignment_functions.rs:383-449 */
static void align_pair(worker_t* w, orc_matrix_t* m, uint32_t r, const uint8_t* read, size_t l2, int traceback,
                       orc_result_t* res, uint32_t* cig, size_t cig_cap) {
    const orc_batch_t* in = w->in;
    const uint8_t* ref = in->ref_bytes + in->ref_off[r];
    size_t l1 = (size_t)(in->ref_off[r + 1] - in->ref_off[r]);
    orc_fill(m, ref, l1, read, l2, w->sc, band_of(in, l1, l2));
    atomic_fetch_add(w->cells, (uint64_t)l1 * l2);
    if (traceback) orc_traceback(m, ref, l1, read, l2, res, cig, cig_cap, NULL, NULL, 0);
    else {
        size_t z = 0;
        double best = m->scores[IDX(m, l1, l2, 0)];
        for (size_t k = 1; k < 3; k++) { double v = m->scores[IDX(m, l1, l2, k)]; if (!(v < best)) { best = v; z = k; } }
        (void)z;
        res->score = best; res->status = ORC_OK; res->n_cigar = 0;
    }
}

static void* worker_main(void* arg) {
    worker_t* w = (worker_t*)arg;
    const orc_batch_t* in = w->in;
    orc_batch_out_t* out = w->out;
    orc_matrix_t* m = orc_matrix_create(w->max_l1 + 1, w->max_l2 + 1); /* thread-local matrix, alignment_functions.rs:136-141 */
    size_t cig_cap = w->max_l1 + w->max_l2 + 2;
    uint32_t* cig_best = (uint32_t*)malloc(cig_cap * sizeof(uint32_t));
    uint32_t* cig_tmp = (uint32_t*)malloc(cig_cap * sizeof(uint32_t));
    uint32_t* votes = (uint32_t*)calloc(in->n_refs + 1, sizeof(uint32_t));
    uint8_t* cand = (uint8_t*)malloc(in->n_refs + 1);
    for (;;) {
        uint64_t i = atomic_fetch_add(w->next, 1); /* par_bridge(): one task per read */
        if (i >= in->n_reads) break;
        const uint8_t* read = in->read_bytes + in->read_off[i];
        size_t l2 = (size_t)(in->read_off[i + 1] - in->read_off[i]);
        orc_result_t best;
        memset(&best, 0, sizeof(best));
        uint32_t best_ref = 0;
        int have = 0;
        if (in->search_mode == ORC_SEARCH_FIXED) {
            if (in->fixed_ref[i] >= 0 && (uint32_t)in->fixed_ref[i] < in->n_refs) { /* otherwise: no candidate (Option::None) */
                best_ref = (uint32_t)in->fixed_ref[i];
                align_pair(w, m, best_ref, read, l2, 1, &best, cig_best, cig_cap);
                have = 1;
            }
        } else {
            int single = -1;
            memset(cand, 1, in->n_refs);
            if (in->search_mode == ORC_SEARCH_QUICK) { /* quick_alignment_search, alignment_functions.rs:693-767 */
                memset(votes, 0, in->n_refs * sizeof(uint32_t));
                uint32_t total = orc_kmer_votes(w->kix, read, l2, votes);
                if (total > 0) {
                    double count = (double)total;
                    double bestp = -1.0;
                    uint32_t bi = 0;
                    for (uint32_t r = 0; r < in->n_refs; r++) { /* ascending index, last maximum wins (total_cmp) */
                        if (!votes[r]) continue;
                        double p = (double)votes[r] / count;
                        if (p >= bestp) { bestp = p; bi = r; }
                    }
                    if (bestp > in->match_threshold) single = (int)bi;
                    else for (uint32_t r = 0; r < in->n_refs; r++) cand[r] = votes[r] > 0;
                }
            }
            if (single >= 0) {
                best_ref = (uint32_t)single;
                align_pair(w, m, best_ref, read, l2, 1, &best, cig_best, cig_cap);
                have = 1;
            } else { /* exhaustive_alignment_search, alignment_functions.rs:769-827 */
                for (uint32_t r = 0; r < in->n_refs; r++) {
                    if (!cand[r]) continue;
                    orc_result_t res;
                    memset(&res, 0, sizeof(res));
                    align_pair(w, m, r, read, l2, in->traceback_all_candidates, &res, cig_tmp, cig_cap);
                    if (!have || !(res.score < best.score)) { /* max_by(partial_cmp): last maximum */
                        best = res; best_ref = r; have = 1;
                        if (in->traceback_all_candidates) { uint32_t* t = cig_best; cig_best = cig_tmp; cig_tmp = t; }
                    }
                }
                if (have && !in->traceback_all_candidates) { /* traceback only the winner */
                    align_pair(w, m, best_ref, read, l2, 1, &best, cig_best, cig_cap);
                }
            }
        }
        if (!have) {
            out->score[i] = 0.0; out->ref_index[i] = 0xFFFFFFFFu; out->status[i] = ORC_NO_CANDIDATE; out->cigar_len[i] = 0;
            w->tl_owner[i] = w->tid; w->tl_off[i] = 0;
            continue;
        }
        out->score[i] = best.score;
        out->ref_index[i] = best_ref;
        out->status[i] = (uint32_t)best.status;
        out->cigar_len[i] = best.n_cigar;
        if (out->matches) out->matches[i] = best.matches;
        if (out->mismatches) out->mismatches[i] = best.mismatches;
        if (w->cig.n + best.n_cigar > w->cig.cap) {
            w->cig.cap = (w->cig.n + best.n_cigar) * 2 + 1024;
            w->cig.v = (uint32_t*)realloc(w->cig.v, w->cig.cap * sizeof(uint32_t));
        }
        memcpy(w->cig.v + w->cig.n, cig_best, best.n_cigar * sizeof(uint32_t));
        w->tl_owner[i] = w->tid;
        w->tl_off[i] = w->cig.n;
        w->cig.n += best.n_cigar;
    }
    free(cand); free(votes); free(cig_tmp); free(cig_best);
    orc_matrix_free(m);
    return NULL;
}

int orc_align_batch(const orc_batch_t* in, const orc_affine_t* sc, orc_batch_out_t* out) {
    int nt = in->threads > 0 ? in->threads : 1;
    size_t max_l1 = 0, max_l2 = 0;
    for (uint32_t r = 0; r < in->n_refs; r++) { size_t l = (size_t)(in->ref_off[r + 1] - in->ref_off[r]); if (l > max_l1) max_l1 = l; }
    for (uint32_t i = 0; i < in->n_reads; i++) { size_t l = (size_t)(in->read_off[i + 1] - in->read_off[i]); if (l > max_l2) max_l2 = l; }
    orc_kmer_index_t* kix = NULL;
    if (in->search_mode == ORC_SEARCH_QUICK) kix = orc_kmer_index_build(in->n_refs, in->ref_bytes, in->ref_off, in->kmer_k, in->kmer_skip);
    atomic_uint_fast64_t next = 0, cells = 0;
    uint64_t* tl_off = (uint64_t*)calloc(in->n_reads + 1, sizeof(uint64_t));
    uint32_t* tl_owner = (uint32_t*)calloc(in->n_reads + 1, sizeof(uint32_t));
    worker_t* ws = (worker_t*)calloc((size_t)nt, sizeof(worker_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    for (int t = 0; t < nt; t++) {
        ws[t].in = in; ws[t].sc = sc; ws[t].out = out; ws[t].kix = kix; ws[t].next = &next; ws[t].cells = &cells;
        ws[t].max_l1 = max_l1; ws[t].max_l2 = max_l2; ws[t].tl_off = tl_off; ws[t].tl_owner = tl_owner; ws[t].tid = (uint32_t)t;
        pthread_create(&th[t], NULL, worker_main, &ws[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    /* concatenate thread-local CIGAR stores in read order */
    uint64_t used = 0;
    int rc = 0;
    for (uint32_t i = 0; i < in->n_reads; i++) {
        uint32_t n = out->cigar_len[i];
        out->cigar_off[i] = used;
        if (used + n > out->cigar_cap) { out->status[i] = ORC_CIGAR_POOL_FULL; out->cigar_len[i] = 0; rc = -2; continue; }
        if (n) memcpy(out->cigar_pool + used, ws[tl_owner[i]].cig.v + tl_off[i], n * sizeof(uint32_t));
        used += n;
    }
    out->cigar_used = used;
    out->cells = (uint64_t)cells;
    for (int t = 0; t < nt; t++) free(ws[t].cig.v);
    free(ws); free(th); free(tl_off); free(tl_owner);
    orc_kmer_index_free(kix);
    return rc;
}


/* ---------------------------------------------------------------------------------------------------------
 * Post-alignment string functions on the path's output.
 * --------------------------------------------------------------------------------------------------------- */

/* is_valid_fasta_base, utils/base_utils.rs:17-23 */
static int orc_is_valid_fasta_base(uint8_t b) {
    if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
    switch (b) {
        case 'A': case 'C': case 'G': case 'T': case 'U': case 'R': case 'Y': case 'S': case 'W': case 'K': case 'M':
        case 'B': case 'D': case 'H': case 'V': case 'N': return 1;
        default: return 0;
    }
}

typedef struct { uint8_t* p; size_t n, cap; } orc_bytes_t;

static void orc_push(orc_bytes_t* v, uint8_t b) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->p = (uint8_t*)realloc(v->p, v->cap);
    }
    v->p[v->n++] = b;
}

/* extract_tagged_sequences, extractor.rs:271-332: the match arms in their source order */
size_t orc_extract_tagged_sequences(const uint8_t* aligned_read, const uint8_t* aligned_ref, size_t n, uint8_t* out, size_t cap) {
    orc_bytes_t vals[256];
    memset(vals, 0, sizeof(vals));
    int in_extractor = 0;
    unsigned next_read = 'a', next_ref = 'A';
    for (size_t i = 0; i < n; i++) {
        const uint8_t rb = aligned_ref[i], qb = aligned_read[i];
        const int valid = orc_is_valid_fasta_base(rb);
        const int upper = (rb >= 'A' && rb <= 'Z') || (rb == '-' && in_extractor);
        const int special = rb >= '0' && rb <= '9'; /* SPECIAL_CHARACTERS, :19-34 */
        if (upper) {                                 /* (_x, true, _z) */
            in_extractor = 1;
            orc_push(&vals[next_ref & 255], rb);
            orc_push(&vals[next_read & 255], qb);
        } else if (!valid && !in_extractor && special) { /* (false, _y, false) if special */
            orc_push(&vals[rb], qb);
        } else if (!valid && in_extractor && special) {  /* (false, _y, true) if special */
            orc_push(&vals[next_ref & 255], rb);
            orc_push(&vals[next_read & 255], qb);
            orc_push(&vals[rb], qb);
        } else {                                     /* (_x, false, _z) */
            if (in_extractor) { next_read++; next_ref++; }
            in_extractor = 0;
        }
    }
    size_t w = 0;
    int overflow = 0;
    for (int k = 0; k < 256; k++) {
        if (!vals[k].n) { free(vals[k].p); continue; }
        if (w + 5 + vals[k].n > cap) overflow = 1;
        if (!overflow) {
            out[w] = (uint8_t)k;
            const uint32_t len = (uint32_t)vals[k].n;
            memcpy(out + w + 1, &len, 4);
            memcpy(out + w + 5, vals[k].p, vals[k].n);
            w += 5 + vals[k].n;
        }
        free(vals[k].p);
    }
    return overflow ? 0 : w;
}

/* reverse_complement, utils/read_utils.rs:50-72 */
void orc_reverse_complement(const uint8_t* dna, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) {
        uint8_t b = dna[n - 1 - i];
        if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
        uint8_t c = b;
        switch (b) {
            case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'G': c = 'C'; break; case 'C': c = 'G'; break;
            case 'R': c = 'Y'; break; case 'Y': c = 'R'; break; case 'K': c = 'M'; break; case 'M': c = 'K'; break;
            case 'B': c = 'V'; break; case 'V': c = 'B'; break; case 'D': c = 'H'; break; case 'H': c = 'D'; break;
            default: break; /* S, W, N and unknown bytes map to themselves */
        }
        out[i] = c;
    }
}


/* ---------------------------------------------------------------------------------------------------------
 * rust-bio `pairwise::Aligner::global` (PARITY UNPINNED, see clq_oracle.h): a restatement of `custom` with the
 * clip penalties `global` installs.  Variable names follow the crate's (S, I, D, Sn, Lx, Ly, traceback cells).
 * --------------------------------------------------------------------------------------------------------- */
enum { RB_START = 0, RB_INS = 1, RB_DEL = 2, RB_SUBST = 3, RB_MATCH = 4, RB_XCLIP_PREFIX = 5, RB_XCLIP_SUFFIX = 6,
       RB_YCLIP_PREFIX = 7, RB_YCLIP_SUFFIX = 8 };
typedef struct { uint8_t s, i, d; } rb_cell_t;

static int32_t rb_max(int32_t a, int32_t b) { return a > b ? a : b; }

int orc_rustbio_global(const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2, int32_t match_score,
                       int32_t mismatch_score, int32_t gap_open, int32_t gap_extend, int32_t* score, uint32_t* n_cigar,
                       uint32_t* cigar, size_t cigar_cap) {
    const uint8_t* x = read; /* global(reference = forward_oriented_seq, read = ref_base.sequence): x is the READ */
    const uint8_t* y = ref;
    const size_t m = l2, n = l1;
    const int32_t MIN = ORC_RUSTBIO_MIN_SCORE;
    const int32_t xclip_prefix = MIN, xclip_suffix = MIN, yclip_prefix = MIN, yclip_suffix = MIN;
    int32_t* Sv[2]; int32_t* Iv[2]; int32_t* Dv[2];
    for (int k = 0; k < 2; k++) {
        Sv[k] = (int32_t*)malloc((m + 1) * sizeof(int32_t));
        Iv[k] = (int32_t*)malloc((m + 1) * sizeof(int32_t));
        Dv[k] = (int32_t*)malloc((m + 1) * sizeof(int32_t));
    }
    int32_t* Sn = (int32_t*)malloc((m + 1) * sizeof(int32_t));
    size_t* Lx = (size_t*)calloc(n + 1, sizeof(size_t));
    size_t* Ly = (size_t*)calloc(m + 1, sizeof(size_t));
    rb_cell_t* tb = (rb_cell_t*)calloc((m + 1) * (n + 1), sizeof(rb_cell_t));
#define TB(i_, j_) tb[(size_t)(i_) * (n + 1) + (j_)]
    /* initial conditions (both rolling columns) */
    for (int k = 0; k < 2; k++) {
        for (size_t i = 0; i <= m; i++) { Dv[k][i] = MIN; Iv[k][i] = MIN; Sv[k][i] = MIN; }
        Sv[k][0] = 0;
        if (k == 0) {
            rb_cell_t c = {RB_START, RB_START, RB_START};
            TB(0, 0) = c;
            for (size_t i = 0; i <= m; i++) Sn[i] = MIN;
            Sn[0] = yclip_suffix;
            Ly[0] = n;
        }
        for (size_t i = 1; i <= m; i++) {
            rb_cell_t c = {RB_START, RB_START, RB_START};
            if (i == 1) {
                Iv[k][i] = gap_open + gap_extend;
                c.i = RB_START;
            } else {
                const int32_t i_score = gap_open + gap_extend * (int32_t)i;
                const int32_t c_score = xclip_prefix + gap_open + gap_extend;
                if (i_score > c_score) { Iv[k][i] = i_score; c.i = RB_INS; }
                else { Iv[k][i] = c_score; c.i = RB_XCLIP_PREFIX; }
            }
            if (i == m) c.s = RB_XCLIP_SUFFIX; else Sv[k][i] = MIN;
            if (Iv[k][i] > Sv[k][i]) { Sv[k][i] = Iv[k][i]; c.s = RB_INS; }
            if (xclip_prefix > Sv[k][i]) { Sv[k][i] = xclip_prefix; c.s = RB_XCLIP_PREFIX; }
            if (i != m && Sv[k][i] + xclip_suffix > Sv[k][m]) { Sv[k][m] = Sv[k][i] + xclip_suffix; Lx[0] = m - i; }
            if (k == 0) TB(i, 0) = c;
            if (Sv[k][i] + yclip_suffix > Sn[i]) { Sn[i] = Sv[k][i] + yclip_suffix; Ly[i] = n; }
        }
    }
    for (size_t j = 1; j <= n; j++) {
        const int curr = (int)(j % 2), prev = 1 - curr;
        {   /* i = 0 */
            rb_cell_t c = {RB_START, RB_START, RB_START};
            Iv[curr][0] = MIN;
            if (j == 1) { Dv[curr][0] = gap_open + gap_extend; c.d = RB_START; }
            else {
                const int32_t d_score = gap_open + gap_extend * (int32_t)j;
                const int32_t c_score = yclip_prefix + gap_open + gap_extend;
                if (d_score > c_score) { Dv[curr][0] = d_score; c.d = RB_DEL; }
                else { Dv[curr][0] = c_score; c.d = RB_YCLIP_PREFIX; }
            }
            if (Dv[curr][0] > yclip_prefix) { Sv[curr][0] = Dv[curr][0]; c.s = RB_DEL; }
            else { Sv[curr][0] = yclip_prefix; c.s = RB_YCLIP_PREFIX; }
            if (j == n && Sn[0] > Sv[curr][0]) { Sv[curr][0] = Sn[0]; c.s = RB_YCLIP_SUFFIX; }
            else if (Sv[curr][0] + yclip_suffix > Sn[0]) { Sn[0] = Sv[curr][0] + yclip_suffix; Ly[0] = n - j; }
            TB(0, j) = c;
        }
        for (size_t i = 1; i <= m; i++) Sv[curr][i] = MIN;
        const uint8_t q = y[j - 1];
        const int32_t xclip_score = xclip_prefix + rb_max(yclip_prefix, gap_open + gap_extend * (int32_t)j);
        for (size_t i = 1; i <= m; i++) {
            const uint8_t p = x[i - 1];
            rb_cell_t c = {RB_START, RB_START, RB_START};
            /* the reference's closure (alignment_functions.rs:55): a = x byte (the read), N in the read matches anything */
            const int32_t sub = (p == q || p == 'N') ? match_score : mismatch_score;
            const int32_t m_score = Sv[prev][i - 1] + sub;
            const int32_t i_score = Iv[curr][i - 1] + gap_extend;
            int32_t s_score = Sv[curr][i - 1] + gap_open + gap_extend;
            int32_t best_i_score;
            if (i_score > s_score) { best_i_score = i_score; c.i = RB_INS; }
            else { best_i_score = s_score; c.i = TB(i - 1, j).s; }
            const int32_t d_score = Dv[prev][i] + gap_extend;
            s_score = Sv[prev][i] + gap_open + gap_extend;
            int32_t best_d_score;
            if (d_score > s_score) { best_d_score = d_score; c.d = RB_DEL; }
            else { best_d_score = s_score; c.d = TB(i, j - 1).s; }
            c.s = RB_XCLIP_SUFFIX;
            int32_t best_s_score = Sv[curr][i];
            if (m_score > best_s_score) { best_s_score = m_score; c.s = (p == q) ? RB_MATCH : RB_SUBST; }
            if (best_i_score > best_s_score) { best_s_score = best_i_score; c.s = RB_INS; }
            if (best_d_score > best_s_score) { best_s_score = best_d_score; c.s = RB_DEL; }
            if (xclip_score > best_s_score) { best_s_score = xclip_score; c.s = RB_XCLIP_PREFIX; }
            const int32_t yclip_score = yclip_prefix + gap_open + gap_extend * (int32_t)i;
            if (yclip_score > best_s_score) { best_s_score = yclip_score; c.s = RB_YCLIP_PREFIX; }
            Sv[curr][i] = best_s_score;
            Iv[curr][i] = best_i_score;
            Dv[curr][i] = best_d_score;
            if (Sv[curr][i] + xclip_suffix > Sv[curr][m]) { Sv[curr][m] = Sv[curr][i] + xclip_suffix; Lx[j] = m - i; }
            if (Sv[curr][i] + yclip_suffix > Sn[i]) { Sn[i] = Sv[curr][i] + yclip_suffix; Ly[i] = n - j; }
            TB(i, j) = c;
        }
    }
    {   /* suffix clipping in the j = n column, then the recomputation of its I values */
        const size_t j = n;
        const int curr = (int)(j % 2);
        for (size_t i = 0; i <= m; i++) {
            if (Sn[i] > Sv[curr][i]) { Sv[curr][i] = Sn[i]; TB(i, j).s = RB_YCLIP_SUFFIX; }
            if (Sv[curr][i] + xclip_suffix > Sv[curr][m]) { Sv[curr][m] = Sv[curr][i] + xclip_suffix; Lx[j] = m - i; TB(m, j).s = RB_XCLIP_SUFFIX; }
        }
        for (size_t i = 1; i <= m; i++) {
            const int32_t s_score = Sv[curr][i - 1] + gap_open + gap_extend;
            if (s_score > Iv[curr][i]) { Iv[curr][i] = s_score; TB(i, j).i = TB(i - 1, j).s; }
            if (s_score > Sv[curr][i]) {
                Sv[curr][i] = s_score;
                TB(i, j).s = RB_INS;
                if (Sv[curr][i] + xclip_suffix > Sv[curr][m]) { Sv[curr][m] = Sv[curr][i] + xclip_suffix; Lx[j] = m - i; TB(m, j).s = RB_XCLIP_SUFFIX; }
            }
        }
    }
    *score = Sv[n % 2][m];
    /* traceback: unit ops pushed back to front, then cigar_to_alignment's simplify_cigar_string (Match and Subst -> M) */
    size_t i = m, j = n, cap = m + n + 2, nu = 0;
    uint8_t* unit = (uint8_t*)malloc(cap);
    int rc = ORC_OK;
    uint8_t last_layer = TB(i, j).s;
    for (;;) {
        uint8_t next_layer;
        if (last_layer == RB_START) break;
        else if (last_layer == RB_INS) { unit[nu++] = ORC_OP_I; next_layer = TB(i, j).i; i -= 1; }
        else if (last_layer == RB_DEL) { unit[nu++] = ORC_OP_D; next_layer = TB(i, j).d; j -= 1; }
        else if (last_layer == RB_MATCH || last_layer == RB_SUBST) { unit[nu++] = ORC_OP_M; next_layer = TB(i - 1, j - 1).s; i -= 1; j -= 1; }
        else { rc = ORC_TRACEBACK_DIVERGED; break; } /* clip operations cannot win with MIN_SCORE penalties */
        last_layer = next_layer;
        if (nu >= cap) { rc = ORC_TRACEBACK_DIVERGED; break; }
    }
    uint32_t nc = 0;
    if (rc == ORC_OK) {
        for (size_t k = nu; k-- > 0;) {
            const uint32_t op = unit[k];
            if (nc && (cigar[nc - 1] & 15u) == op) cigar[nc - 1] += 16u;
            else if (nc < cigar_cap) cigar[nc++] = 16u | op;
            else { rc = ORC_CIGAR_POOL_FULL; break; }
        }
    }
    *n_cigar = nc;
#undef TB
    free(unit); free(tb); free(Ly); free(Lx); free(Sn);
    for (int k = 0; k < 2; k++) { free(Sv[k]); free(Iv[k]); free(Dv[k]); }
    return rc;
}


/* ---------------------------------------------------------------------------------------------------------
 * Paired-read merging after align_two_strings: utils/read_utils.rs:6-38, merger.rs:428-498
 * --------------------------------------------------------------------------------------------------------- */
double orc_phred_to_prob(uint8_t phred) {
    const double phred_f64 = (double)((size_t)phred - 33); /* (*phred as usize) - 33: panics below 33 in debug builds */
    return pow(10.0, (-1.0 * phred_f64) / 10.0);
}

static uint8_t orc_sat_u8(double v) { /* Rust `as u8`: saturating, NaN -> 0 */
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

uint8_t orc_prob_to_phred(double qual) { return orc_sat_u8(((-10.0) * log10(qual)) + 33.0); }

uint8_t orc_combine_phred_scores(uint8_t phred_one, uint8_t phred_two, int agree) {
    const double prob1 = orc_phred_to_prob(phred_one), prob2 = orc_phred_to_prob(phred_two);
    if (agree) return orc_prob_to_phred(prob1 * prob2);
    return orc_prob_to_phred(1.0 - ((1.0 - prob2) * (1.0 * prob1)));
}

size_t orc_alignment_rate_and_consensus(const uint8_t* a1, const uint8_t* q1, size_t nq1, const uint8_t* a2, const uint8_t* q2,
                                        size_t nq2, size_t n, uint8_t* out_bases, uint8_t* out_quals) {
    size_t p1 = 0, p2 = 0;
    for (size_t i = 0; i < n; i++) {
        const uint8_t a = a1[i], b = a2[i];
        if (a == b) { /* includes gap/gap, as the reference does */
            if (p1 >= nq1 || p2 >= nq2) return (size_t)-1;
            out_bases[i] = a;
            out_quals[i] = orc_combine_phred_scores(q1[p1], q2[p2], 1);
            p1++; p2++;
        } else if (a == '-') {
            if (p2 >= nq2) return (size_t)-1;
            out_bases[i] = b; out_quals[i] = q2[p2]; p2++;
        } else if (b == '-') {
            if (p1 >= nq1) return (size_t)-1;
            out_bases[i] = a; out_quals[i] = q1[p1]; p1++;
        } else {
            if (p1 >= nq1 || p2 >= nq2) return (size_t)-1;
            out_bases[i] = q1[p1] >= q2[p2] ? a : b; /* bases disagree: the higher quality base */
            out_quals[i] = orc_combine_phred_scores(q1[p1], q2[p2], 0);
            p1++; p2++;
        }
    }
    return n;
}


/* ---------------------------------------------------------------------------------------------------------
 * Strand orientation: linked_alignment.rs:24-32, :97-130, :341-362
 * --------------------------------------------------------------------------------------------------------- */
static int orc_acgt(uint8_t b) { /* DEGENERATEBASES value sets hold only A/C/G/T in both cases, fasta_comparisons.rs:21-67 */
    switch (b) { case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 3; case 'T': case 't': return 4; default: return 0; }
}

/* extend_hit: both lookups must succeed both ways, i.e. plain bases equal up to case */
size_t orc_extend_hit(const uint8_t* search, size_t n, size_t sloc, const uint8_t* ref, size_t m, size_t rloc) {
    size_t len = 0;
    while (len + sloc < n && len + rloc < m) {
        const int a = orc_acgt(search[sloc + len]), b = orc_acgt(ref[rloc + len]);
        if (!a || a != b) return len;
        len++;
    }
    return len;
}

typedef struct { const uint8_t* ref; size_t m; } orc_sa_ctx_t;
static orc_sa_ctx_t g_sa_ctx; /* qsort has no context argument; the oracle is single-threaded here */
static int orc_suffix_cmp(const void* pa, const void* pb) {
    const uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
    const size_t la = g_sa_ctx.m - a, lb = g_sa_ctx.m - b, l = la < lb ? la : lb;
    const int c = memcmp(g_sa_ctx.ref + a, g_sa_ctx.ref + b, l);
    if (c) return c;
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

size_t orc_find_greedy_non_overlapping_segments(const uint8_t* search, size_t n, const uint8_t* reference, size_t m,
                                                size_t seed_size, orc_segment_t* out, size_t cap, size_t* start_position) {
    size_t nh = 0, position = 0, least_ref_pos = m, greatest_ref_pos = 0;
    uint32_t* pos = (uint32_t*)malloc((m + 1) * sizeof(uint32_t));
    while ((int64_t)position <= (int64_t)n - (int64_t)seed_size) {
        /* SuffixTable::positions(seed): every occurrence, in suffix-array order */
        size_t np = 0;
        for (size_t r = 0; seed_size > 0 && r + seed_size <= m; r++)
            if (memcmp(reference + r, search + position, seed_size) == 0) pos[np++] = (uint32_t)r;
        g_sa_ctx.ref = reference; g_sa_ctx.m = m;
        qsort(pos, np, sizeof(uint32_t), orc_suffix_cmp);
        size_t longest_hit = 0;
        for (size_t k = 0; k < np; k++) {
            const size_t ref_position = pos[k];
            if (ref_position >= greatest_ref_pos) {
                const size_t ext = orc_extend_hit(search, n, position, reference, m, ref_position); /* `position` may have moved */
                if (ext > longest_hit) {
                    if (nh < cap) { out[nh].search_start = (uint32_t)position; out[nh].ref_start = (uint32_t)ref_position; out[nh].length = (uint32_t)ext; }
                    nh++;
                    position += ext;
                    if (ref_position < least_ref_pos) least_ref_pos = ref_position;
                    if (ref_position + ext > greatest_ref_pos) greatest_ref_pos = ref_position + ext;
                    longest_hit = ext;
                }
            }
        }
        position += 1;
    }
    free(pos);
    if (start_position) *start_position = least_ref_pos;
    return nh < cap ? nh : cap;
}

/* bio::alphabets::dna::revcomp: complement over "AGCTYRWSKMDVHBN" <-> "TCGARYWSMKHBDVN" in both cases, other bytes unchanged */
static uint8_t orc_bio_complement(uint8_t b) {
    static const char from[] = "AGCTYRWSKMDVHBNagctyrwskmdvhbn";
    static const char to[] = "TCGARYWSMKHBDVNtcgarywsmkhbdvn";
    for (int i = 0; from[i]; i++) if ((uint8_t)from[i] == b) return (uint8_t)to[i];
    return b;
}

int orc_orient_by_longest_segment(const uint8_t* search, size_t n, const uint8_t* reference, size_t m, size_t seed_size,
                                  size_t* fwd_score, size_t* rev_score) {
    const size_t cap = n + 1;
    orc_segment_t* seg = (orc_segment_t*)malloc(cap * sizeof(orc_segment_t));
    size_t nf = orc_find_greedy_non_overlapping_segments(search, n, reference, m, seed_size, seg, cap, NULL), f = 0, r = 0;
    for (size_t i = 0; i < nf; i++) f += seg[i].length;
    uint8_t* rc = (uint8_t*)malloc(n + 1);
    for (size_t i = 0; i < n; i++) rc[i] = orc_bio_complement(search[n - 1 - i]);
    size_t nr = orc_find_greedy_non_overlapping_segments(rc, n, reference, m, seed_size, seg, cap, NULL);
    for (size_t i = 0; i < nr; i++) r += seg[i].length;
    free(rc); free(seg);
    if (fwd_score) *fwd_score = f;
    if (rev_score) *rev_score = r;
    return f > r;
}
