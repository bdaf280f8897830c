/*
 * dbt_oracle.c -- CPU oracle for the dbtproj tuple operators.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under
 * database-technology-algorithms_b200/ links, imports or executes it.
 *
 * It restates, in plain C, the *defect-free* semantics ("CANON") of the
 * reference's four operators, each function citing the reference lines it follows.  The
 * reference as shipped ("REF") loses/duplicates a few rows; REF
 * itself is available as oracle/_ref/ref_runner (built from the untouched reference sources by
 * oracle/Makefile) and the two are pinned against each other by tests/test_oracle_golden.py (frozen
 * fixtures in tests/golden/) and tests/test_gpu_entrypoints.py (REF run live).
 *
 * Parity status: PINNED against the reference binary's behaviour (the reference has no tests or
 * golden vectors of its own).
 *
 * All functions work on in-memory images of block files: a flat array of 14016-byte blocks.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_STR 120
#define ORC_RPB 100

/* reference: dbtproj.h:20-27 */
typedef struct {
    uint32_t recid;
    uint32_t num;
    char str[ORC_STR];
    uint8_t valid;
    uint8_t pad_[3];
    uint32_t dummy1;
    uint32_t dummy2;
} orc_record;

/* reference: dbtproj.h:31-38 */
typedef struct {
    uint32_t blockid;
    uint32_t nreserved;
    orc_record entries[ORC_RPB];
    uint8_t valid;
    uint8_t misc;
    uint8_t pad_[2];
    uint32_t dummy;
} orc_block;

_Static_assert(sizeof(orc_record) == 140, "record layout");
_Static_assert(sizeof(orc_block) == 14016, "block layout");

int orc_abi_version(void) { return 1; }

/* ------------------------------------------------------------------------------------------
 * Key order.  reference: DatabaseProject.cpp:44-92 (compareNUM/STR/ID/NUMSTR) and :18-42.
 * strcmp semantics: unsigned bytes up to the first NUL; bytes after the NUL are ignored.
 * A str with no NUL in its 120 bytes is treated as 120 characters long (the reference would
 * read past the field; CANON bounds it).
 * ------------------------------------------------------------------------------------------ */
static int cmp_str(const char *a, const char *b) {
    for (int i = 0; i < ORC_STR; ++i) {
        unsigned char ca = (unsigned char)a[i], cb = (unsigned char)b[i];
        if (ca != cb) return ca < cb ? -1 : 1;
        if (ca == 0) return 0;
    }
    return 0;
}

static int cmp_key(const orc_record *a, const orc_record *b, int field) {
    switch (field) {
    case '0': return a->recid < b->recid ? -1 : (a->recid > b->recid ? 1 : 0);
    case '1': return a->num < b->num ? -1 : (a->num > b->num ? 1 : 0);
    case '2': return cmp_str(a->str, b->str);
    case '3':
        if (a->num != b->num) return a->num < b->num ? -1 : 1;
        return cmp_str(a->str, b->str);
    default: return 0;
    }
}

int orc_field_ok(int field) { return field >= '0' && field <= '3'; }

/* Row index: pointers to the live entries of a block image, in file order.
 * reference: DatabaseProject.cpp:198-205 (only entries[0..nreserved) of each block are rows). */
static int64_t collect_rows(const orc_block *blocks, int64_t nblocks, const orc_record ***out) {
    int64_t n = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        uint32_t r = blocks[b].nreserved;
        n += r > ORC_RPB ? ORC_RPB : r;
    }
    const orc_record **rows = (const orc_record **)malloc(sizeof(*rows) * (size_t)(n ? n : 1));
    int64_t k = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        uint32_t r = blocks[b].nreserved;
        if (r > ORC_RPB) r = ORC_RPB;
        for (uint32_t i = 0; i < r; ++i) rows[k++] = &blocks[b].entries[i];
    }
    *out = rows;
    return n;
}

int64_t orc_count_rows(const void *blocks, int64_t nblocks) {
    const orc_block *bl = (const orc_block *)blocks;
    int64_t n = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        uint32_t r = bl[b].nreserved;
        n += r > ORC_RPB ? ORC_RPB : r;
    }
    return n;
}

/* Stable merge sort of row pointers by (key(field), recid); remaining ties keep file order.
 * CANON tie rule: the reference's own tie order is platform luck, north_star
 * canonicalises ties by recid. */
static int g_field;
static int cmp_rows(const orc_record *a, const orc_record *b) {
    int c = cmp_key(a, b, g_field);
    if (c) return c;
    return a->recid < b->recid ? -1 : (a->recid > b->recid ? 1 : 0);
}

static void msort(const orc_record **a, const orc_record **tmp, int64_t n) {
    if (n < 2) return;
    if (n <= 16) { /* insertion sort, stable */
        for (int64_t i = 1; i < n; ++i) {
            const orc_record *x = a[i];
            int64_t j = i;
            while (j > 0 && cmp_rows(a[j - 1], x) > 0) { a[j] = a[j - 1]; --j; }
            a[j] = x;
        }
        return;
    }
    int64_t h = n / 2;
    msort(a, tmp, h);
    msort(a + h, tmp, n - h);
    if (cmp_rows(a[h - 1], a[h]) <= 0) return;
    memcpy(tmp, a, sizeof(*a) * (size_t)h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) a[k++] = (cmp_rows(a[j], tmp[i]) < 0) ? a[j++] : tmp[i++];
    while (i < h) a[k++] = tmp[i++];
}

static void sort_rows(const orc_record **rows, int64_t n, int field) {
    g_field = field;
    const orc_record **tmp = (const orc_record **)malloc(sizeof(*tmp) * (size_t)(n ? n : 1));
    msort(rows, tmp, n);
    free(tmp);
}

/* CANON output image: rows packed 100 per block; block k has blockid=k, nreserved=live rows,
 * valid=1, misc=0, dummy=nreserved; unused entries and padding are zero.  (The reference leaves
 * output headers as junk; CANON writes sane ones.)  Rows are copied
 * verbatim, all 140 bytes. */
static int64_t pack_rows(const orc_record **rows, int64_t n, orc_block *out) {
    int64_t nb = (n + ORC_RPB - 1) / ORC_RPB;
    if (nb) memset(out, 0, sizeof(orc_block) * (size_t)nb);
    for (int64_t k = 0; k < nb; ++k) {
        int64_t lo = k * ORC_RPB, hi = lo + ORC_RPB;
        if (hi > n) hi = n;
        out[k].blockid = (uint32_t)k;
        out[k].nreserved = (uint32_t)(hi - lo);
        out[k].valid = 1;
        out[k].dummy = (uint32_t)(hi - lo);
        for (int64_t i = lo; i < hi; ++i) out[k].entries[i - lo] = *rows[i];
    }
    return nb;
}

int64_t orc_out_blocks(int64_t nrows) { return (nrows + ORC_RPB - 1) / ORC_RPB; }

/* MergeSort, CANON.  reference: DatabaseProject.cpp:172-381 (run generation + k-way merge have
 * one observable result when defect-free: every input row once, ordered by key).
 * `out` must hold orc_out_blocks(orc_count_rows(in)) blocks.  Returns the row count. */
int64_t orc_sort(const void *in, int64_t nblocks, int field, void *out) {
    const orc_record **rows;
    int64_t n = collect_rows((const orc_block *)in, nblocks, &rows);
    sort_rows(rows, n, field);
    pack_rows(rows, n, (orc_block *)out);
    free(rows);
    return n;
}

/* first-of-group filter over sorted rows; reference: DatabaseProject.cpp:121-162 (keep a row iff
 * its key differs from the previous row's).  CANON keeps the minimum-recid row of each key
 * group (the sort above puts it first) and has none of D5/D6/D7. */
static int64_t unique_rows(const orc_record **rows, int64_t n, int field) {
    int64_t u = 0;
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || cmp_key(rows[i], rows[u - 1], field) != 0) rows[u++] = rows[i];
    return u;
}

/* EliminateDuplicates, CANON.  Returns nunique; `out` sized for the input row count. */
int64_t orc_dedup(const void *in, int64_t nblocks, int field, void *out) {
    const orc_record **rows;
    int64_t n = collect_rows((const orc_block *)in, nblocks, &rows);
    sort_rows(rows, n, field);
    int64_t u = unique_rows(rows, n, field);
    pack_rows(rows, u, (orc_block *)out);
    free(rows);
    return u;
}

/* MergeJoin, CANON.  reference: DatabaseProject.cpp:384-502: dedup both inputs (side files
 * "1outfile.bin"/"2outfile.bin", :385-394), two-pointer intersection (:414-482) emitting R's row
 * per common key (:454).  Also simulates the loop's block-read count for nios (:405,465,475).
 * out_uR/out_uS sized for the input row counts; out sized for min of them.
 * res[0]=nres res[1]=nunique_R res[2]=nunique_S res[3]=later block reads (both files). */
void orc_mergejoin(const void *inR, int64_t nbR, const void *inS, int64_t nbS, int field,
                   void *out_uR, void *out_uS, void *out, int64_t *res) {
    const orc_record **r, **s;
    int64_t nr = collect_rows((const orc_block *)inR, nbR, &r);
    int64_t ns = collect_rows((const orc_block *)inS, nbS, &s);
    sort_rows(r, nr, field);
    sort_rows(s, ns, field);
    int64_t ur = unique_rows(r, nr, field), us = unique_rows(s, ns, field);
    pack_rows(r, ur, (orc_block *)out_uR);
    pack_rows(s, us, (orc_block *)out_uS);
    const orc_record **m = (const orc_record **)malloc(sizeof(*m) * (size_t)((ur < us ? ur : us) + 1));
    int64_t i = 0, j = 0, k = 0, reads = 0;
    /* the walk: a block boundary of either list costs one read; the walk ends at the first
     * read that returns an empty block (R is checked before S, :416-417,462-479). */
    int64_t loadedR = ur ? 1 : 0, loadedS = us ? 1 : 0; /* blocks loaded so far (first reads are the "+2") */
    if (ur && us) {
        for (;;) {
            if (i >= loadedR * ORC_RPB || i >= ur) { /* R index ran past the current block */
                ++reads;
                if (i >= ur) break;
                ++loadedR;
                continue;
            }
            if (j >= loadedS * ORC_RPB || j >= us) {
                ++reads;
                if (j >= us) break;
                ++loadedS;
                continue;
            }
            int c = cmp_key(r[i], s[j], field);
            if (c < 0) ++i;
            else if (c > 0) ++j;
            else { m[k++] = r[i]; ++i; ++j; }
        }
    }
    pack_rows(m, k, (orc_block *)out);
    res[0] = k; res[1] = ur; res[2] = us; res[3] = reads;
    free(m); free(r); free(s);
}

/* HashJoin, CANON.  reference: DatabaseProject.cpp:504-647.  Build = the key set of R
 * (fields '0'..'2', :531-540) or the key multiset (field '3', :541-544); probe S in file order
 * and emit S's row once if its key is in the set (:584-615), or once per matching R row for
 * field '3' (:616-629).  Implemented as sort(R) + binary search (same result as hashing).
 * `out` must hold orc_out_blocks(nres) blocks: call with out=NULL first to get nres. */
int64_t orc_hashjoin(const void *inR, int64_t nbR, const void *inS, int64_t nbS, int field, void *out) {
    const orc_record **r, **s;
    int64_t nr = collect_rows((const orc_block *)inR, nbR, &r);
    int64_t ns = collect_rows((const orc_block *)inS, nbS, &s);
    sort_rows(r, nr, field);
    int64_t cap = 1024, k = 0;
    const orc_record **m = out ? (const orc_record **)malloc(sizeof(*m) * (size_t)cap) : NULL;
    for (int64_t q = 0; q < ns; ++q) {
        int64_t lo = 0, hi = nr; /* lower bound of s[q]'s key in sorted R */
        while (lo < hi) {
            int64_t mid = (lo + hi) / 2;
            if (cmp_key(r[mid], s[q], field) < 0) lo = mid + 1; else hi = mid;
        }
        int64_t cnt = 0;
        if (field == '3') {
            for (int64_t t = lo; t < nr && cmp_key(r[t], s[q], field) == 0; ++t) ++cnt;
        } else {
            cnt = (lo < nr && cmp_key(r[lo], s[q], field) == 0) ? 1 : 0;
        }
        for (int64_t t = 0; t < cnt; ++t) {
            if (m) {
                if (k == cap) { cap *= 2; m = (const orc_record **)realloc(m, sizeof(*m) * (size_t)cap); }
                m[k] = s[q];
            }
            ++k;
        }
    }
    if (out) pack_rows(m, k, (orc_block *)out);
    free(m); free(r); free(s);
    return k;
}

/* Pair-producing inner join (EXTENSION: north_star's comparison format; the reference never emits pairs,
 * SURVEY.md F9).  pairs[2k] = recid of the R row, pairs[2k+1] = recid of the S row, for every (R row, S row)
 * with equal key(field); order: S file order, then R rows by (key, recid).  Call with pairs=NULL to count.
 * Consistency with the reference: the distinct S recids of the pairs are HashJoin's output for
 * fields '0'..'2'; the number of pairs is HashJoin's nres for field '3'. */
int64_t orc_innerjoin_pairs(const void *inR, int64_t nbR, const void *inS, int64_t nbS, int field, uint32_t *pairs) {
    const orc_record **r, **s;
    int64_t nr = collect_rows((const orc_block *)inR, nbR, &r);
    int64_t ns = collect_rows((const orc_block *)inS, nbS, &s);
    sort_rows(r, nr, field);
    int64_t k = 0;
    for (int64_t q = 0; q < ns; ++q) {
        int64_t lo = 0, hi = nr;
        while (lo < hi) {
            int64_t mid = (lo + hi) / 2;
            if (cmp_key(r[mid], s[q], field) < 0) lo = mid + 1; else hi = mid;
        }
        for (int64_t t = lo; t < nr && cmp_key(r[t], s[q], field) == 0; ++t) {
            if (pairs) { pairs[2 * k] = r[t]->recid; pairs[2 * k + 1] = s[q]->recid; }
            ++k;
        }
    }
    free(r); free(s);
    return k;
}

/* ------------------------------------------------------------------------------------------
 * Counters.  B input blocks, M = nmem_blocks, F = M-1.
 * reference: DatabaseProject.cpp:191-233 (runs of M blocks, each written as exactly M blocks),
 * :245-369 (phases of F-way merges, at least one), :371-376 (outputs).
 * ------------------------------------------------------------------------------------------ */
void orc_sort_counters(int64_t B, int64_t M, int64_t *segs, int64_t *passes, int64_t *nios) {
    int64_t F = M - 1;
    int64_t R = (B + M - 1) / M; /* runs */
    int64_t f = R, total = R, phases = 0;
    do {
        f = (f + F - 1) / F;
        total += f;
        ++phases;
    } while (f > 1); /* a file with zero runs still "produces" per the loop shape; B=0 is guarded by callers */
    *segs = total;
    *passes = 1 + phases;
    *nios = R * M + phases * B;
}

/* nios of the other operators, CANON flavour (Appendix B). */
int64_t orc_dedup_nios(int64_t B, int64_t M, int64_t nunique) {
    int64_t s, p, io;
    orc_sort_counters(B, M, &s, &p, &io);
    return io + (nunique + ORC_RPB - 1) / ORC_RPB;
}

int64_t orc_hashjoin_nios(int64_t BR, int64_t BS, int64_t M, int64_t nres) {
    int64_t F = M - 1;
    /* reference: DatabaseProject.cpp:518-525,561-568: one count per bulk fread of F blocks, and
     * the loop always pays a final short/empty read */
    return (BR / F + 1) + (BS / F + 1) + (nres + ORC_RPB - 1) / ORC_RPB;
}

/* ------------------------------------------------------------------------------------------
 * Generators.
 * ------------------------------------------------------------------------------------------ */

/* G_ref: reference: main.cpp:10-18,41-77 with srand(seed) instead of time(0); glibc rand() call
 * order per row: num1, 5 x str1, num2, 5 x str2.  Rows are zero-filled first (tail_mode 0) or
 * have every byte after the NUL, plus misc/dummy1/dummy2 and padding, set to junk (tail_mode 1:
 * the reference's rows carry uninitialised stack bytes there). */
static void junk_fill(orc_record *rec, uint32_t *state) {
    size_t len = strnlen(rec->str, ORC_STR);
    for (size_t i = len + 1; i < ORC_STR; ++i) {
        *state = *state * 1664525u + 1013904223u;
        rec->str[i] = (char)(0x80 | (*state >> 24)) ; /* never NUL */
    }
    *state = *state * 1664525u + 1013904223u; rec->dummy1 = *state;
    *state = *state * 1664525u + 1013904223u; rec->dummy2 = *state;
    rec->pad_[0] = 0xA5; rec->pad_[1] = 0x5A; rec->pad_[2] = 0xC3;
}

void orc_gen_ref(unsigned seed, int64_t nblocks, int tail_mode, uint32_t num_mod, void *file1, void *file2) {
    orc_block *f1 = (orc_block *)file1, *f2 = (orc_block *)file2;
    static const char alpha[] = "abcdefghijklmnopqrstuvwxyz";
    uint32_t js = 0x12345u ^ seed;
    uint32_t mod = num_mod ? num_mod : (uint32_t)(nblocks * 30);
    srand(seed);
    uint32_t recid = 0;
    memset(f1, 0, sizeof(orc_block) * (size_t)nblocks);
    if (f2) memset(f2, 0, sizeof(orc_block) * (size_t)nblocks);
    for (int64_t b = 0; b < nblocks; ++b) {
        for (int r = 0; r < ORC_RPB; ++r) {
            orc_record r1, r2;
            memset(&r1, 0, sizeof r1); memset(&r2, 0, sizeof r2);
            r1.recid = recid;
            r1.num = (uint32_t)rand() % mod;
            for (int i = 0; i < 5; ++i) r1.str[i] = alpha[rand() % 26];
            r2.recid = recid++;
            r2.num = (uint32_t)rand() % mod;
            for (int i = 0; i < 5; ++i) r2.str[i] = alpha[rand() % 26];
            if (r == 1) { strcpy(r1.str, "Hola"); strcpy(r2.str, "Hola"); r1.str[5] = r2.str[5] = 0; }
            r1.valid = r2.valid = 1;
            if (tail_mode) { junk_fill(&r1, &js); junk_fill(&r2, &js); }
            f1[b].entries[r] = r1;
            if (f2) f2[b].entries[r] = r2;
        }
        f1[b].blockid = (uint32_t)b; f1[b].nreserved = ORC_RPB; f1[b].valid = 1; f1[b].dummy = ORC_RPB;
        if (f2) { f2[b].blockid = (uint32_t)b; f2[b].nreserved = ORC_RPB; f2[b].valid = 1; f2[b].dummy = ORC_RPB; }
    }
}

/* G_syn: counter-based generator for sizes the reference cannot run.  The same
 * arithmetic is implemented on the device in csrc/gen.cu; this CPU copy lets tests spot-check
 * any sub-range.  Row r of a file of n rows:
 *   j    = (r * A + C) mod n                      (affine bijection of row positions, gcd(A,n)=1)
 *   num  = kind 0: mix32(seed, j mod U)           (U distinct keys; j and j+U collide => n-U duplicate rows)
 *          kind 1: mix64(seed, r) mod U           (uniform over [0,U))
 *          kind 2: zipf-ish: floor(U * u^s')      (skewed, heavy head), u from mix64
 *   str  = 5 letters from mix64(seed^0x5bd1, r)   ("Hola" when r mod 100 == 1, like main.cpp:57-61)
 *   recid = recid0 + r
 */
static uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
/* invertible 32-bit mixer (bijection on u32) */
static uint32_t bij32(uint32_t x, uint32_t seed) {
    x ^= seed;
    x *= 0x9E3779B1u; x ^= x >> 15;
    x *= 0x85EBCA77u; x ^= x >> 13;
    x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}

/* kind 4: exact Zipf(s = 1.1) ranks over [1, U] by rejection-inversion (Hoermann & Derflinger, "Rejection-inversion to
 * generate variates from monotone discrete distributions", 1996): H(x) = (x^(1-s) - 1)/(1-s) is the integral of the
 * hat x^-s, a uniform u in (H(U + 1/2), H(3/2) - 1] is mapped through H^-1 and rounded to the nearest rank k, which is
 * accepted when it is within s_cut of x or when u >= H(k + 1/2) - k^-s.  SURVEY.md 8(d) cfg 4: "S.num = rho(rank),
 * rank ~ Zipf(s=1.1) over [1,D], rho a seeded bijection".  log and exp are spelled out with IEEE +, *, / in a fixed
 * order (this file is compiled with -ffp-contract=off) so that the device generator, which uses the same sequence of
 * correctly rounded operations, produces the same bits. */
static double zbits2d(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static uint64_t zd2bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
static double zlog(double x) {
    uint64_t b = zd2bits(x);
    int e = (int)((b >> 52) & 0x7FF) - 1023;
    double m = zbits2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double f = (m + -1.0) / (m + 1.0), f2 = f * f;
    double p = 1.0 / 23.0;
    for (int k = 21; k >= 1; k -= 2) p = p * f2 + 1.0 / (double)k;
    return (double)e * 0.6931471805599453 + (2.0 * f) * p;
}
static double zexp(double y) {
    static const double inv_fact[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                        1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,
                                        0.5, 1.0, 1.0};
    double kf = y * 1.4426950408889634;
    long long k = (long long)(kf < 0 ? kf + -0.5 : kf + 0.5);
    double t1 = (double)k * 0.693147180369123816490, t2 = (double)k * 1.90821492927058770002e-10;
    double r = (y + -t1) + -t2;
    double p = 1.0 / 6227020800.0;
    for (int i = 0; i < 13; ++i) p = p * r + inv_fact[i];
    return p * zbits2d((uint64_t)(k + 1023) << 52);
}
#define ORC_ZIPF_S 1.1
static double zh(double x) { return zexp(-ORC_ZIPF_S * zlog(x)); }
static double zH(double x) { return (zexp((1.0 - ORC_ZIPF_S) * zlog(x)) + -1.0) / (1.0 - ORC_ZIPF_S); }
static double zHinv(double u) {
    double t = 1.0 + (1.0 - ORC_ZIPF_S) * u;
    if (t < 1e-300) t = 1e-300;
    return zexp(zlog(t) / (1.0 - ORC_ZIPF_S));
}
uint64_t orc_zipf_rank(uint64_t seed, uint64_t U, uint64_t r) {
    double h_x1 = zH(1.5) + -1.0, h_n = zH((double)U + 0.5), s_cut = 2.0 + -zHinv(zH(2.5) + -zh(2.0));
    for (uint64_t it = 0;; ++it) {
        uint64_t hsh = mix64(mix64(seed * 0x100000001B3ull + r) + it * 0xD6E8FEB86659FD93ull);
        double u01 = (double)(hsh >> 11) * 1.1102230246251565e-16;
        double u = h_n + u01 * (h_x1 + -h_n);
        double x = zHinv(u);
        long long k = (long long)(x + 0.5);
        if (k < 1) k = 1;
        if ((uint64_t)k > U) k = (long long)U;
        if ((double)k + -x <= s_cut || u >= zH((double)k + 0.5) + -zh((double)k) || it >= 63) return (uint64_t)k;
    }
}

uint32_t orc_syn_num(uint64_t seed, uint64_t n, uint64_t U, int kind, uint64_t r) {
    const uint64_t A = 2654435761ull, C = 40503ull; /* A is prime > any n we use => gcd(A,n)=1 unless n multiple of A */
    if (kind == 4) return (uint32_t)(((orc_zipf_rank(seed, U, r) - 1) * A + C) % U);
    if (kind == 0) {
        uint64_t j = (uint64_t)(((unsigned __int128)r * A + C) % n);
        return bij32((uint32_t)(j % U), (uint32_t)seed);
    } else if (kind == 1) {
        return (uint32_t)(mix64(seed * 0x100000001B3ull + r) % U);
    } else {
        /* rank = floor(U * u^8): P(rank < x) = (x/U)^(1/8): a heavy head (power-law-like skew,
         * integer-only so CPU and GPU agree bit for bit); key = bij over the rank, mod U */
        uint64_t h = mix64(seed * 0x100000001B3ull + r);
        uint64_t u = h >> 32; /* 32-bit uniform */
        unsigned __int128 p = u;
        for (int i = 0; i < 3; ++i) p = (p * p) >> 32; /* u^8 in 0.32 fixed point */
        uint64_t rank = (uint64_t)((p * U) >> 32);
        return (uint32_t)(bij32((uint32_t)rank, (uint32_t)seed) % U);
    }
}

void orc_gen_syn(uint64_t seed, uint64_t n_total, uint64_t U, int kind, uint64_t row0, uint64_t nrows,
                 uint32_t recid0, void *blocks) {
    orc_block *bl = (orc_block *)blocks;
    uint64_t nb = (nrows + ORC_RPB - 1) / ORC_RPB;
    memset(bl, 0, sizeof(orc_block) * (size_t)nb);
    for (uint64_t k = 0; k < nrows; ++k) {
        uint64_t r = row0 + k;
        orc_record *rec = &bl[k / ORC_RPB].entries[k % ORC_RPB];
        rec->recid = recid0 + (uint32_t)r;
        /* kind 3: with probability 1/2 the row copies (num, str) of a pseudo-random row of the partner relation
         * (kind 1, seed ^ 0x5EED): about half of the composite keys of the two relations match */
        uint64_t kseed = seed, krow = r;
        int kkind = kind;
        if (kind == 3) {
            uint64_t h3 = mix64(seed * 0x100000001B3ull + r + 0x777ull);
            kkind = 1;
            if (h3 & 1) { kseed = seed ^ 0x5EEDull; krow = (h3 >> 1) % n_total; }
        }
        rec->num = orc_syn_num(kseed, n_total, U, kkind, krow);
        r = krow;
        uint64_t h = mix64((kseed ^ 0x5bd1e995ull) * 0x100000001B3ull + r);
        if (r % ORC_RPB == 1) { memcpy(rec->str, "Hola", 5); }
        else for (int i = 0; i < 5; ++i) { rec->str[i] = (char)('a' + (h % 26)); h /= 26; }
        rec->valid = 1;
    }
    for (uint64_t b = 0; b < nb; ++b) {
        uint64_t lo = b * ORC_RPB, hi = lo + ORC_RPB; if (hi > nrows) hi = nrows;
        bl[b].blockid = (uint32_t)(row0 / ORC_RPB + b);
        bl[b].nreserved = (uint32_t)(hi - lo);
        bl[b].valid = 1;
        bl[b].dummy = (uint32_t)(hi - lo);
    }
}

/* ------------------------------------------------------------------------------------------
 * Small helpers for the tests.
 * ------------------------------------------------------------------------------------------ */

/* 1 iff the rows of the image are ordered by key(field) (non-decreasing). */
int orc_is_sorted(const void *blocks, int64_t nblocks, int field) {
    const orc_record **rows;
    int64_t n = collect_rows((const orc_block *)blocks, nblocks, &rows);
    int ok = 1;
    for (int64_t i = 1; i < n && ok; ++i) if (cmp_key(rows[i - 1], rows[i], field) > 0) ok = 0;
    free(rows);
    return ok;
}

/* Stable re-sort of equal-key groups by recid, in place on a packed image (used to canonicalise
 * REF's arbitrary tie order, SURVEY.md 8c "REFc"). Rows must already be key-sorted. */
void orc_canonicalise_ties(void *blocks, int64_t nblocks, int field) {
    orc_block *bl = (orc_block *)blocks;
    const orc_record **rows;
    int64_t n = collect_rows(bl, nblocks, &rows);
    orc_record *copy = (orc_record *)malloc(sizeof(orc_record) * (size_t)(n ? n : 1));
    for (int64_t i = 0; i < n; ++i) copy[i] = *rows[i];
    const orc_record **p = (const orc_record **)malloc(sizeof(*p) * (size_t)(n ? n : 1));
    for (int64_t i = 0; i < n; ++i) p[i] = &copy[i];
    sort_rows(p, n, field);
    for (int64_t i = 0; i < n; ++i) *(orc_record *)rows[i] = *p[i];
    free(p); free(copy); free(rows);
}
