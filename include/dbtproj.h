/*
 * dbtproj.h -- drop-in contract header of the B200-native tuple operators.
 *
 * Layout- and signature-compatible with the reference's contract header
 * (reference: dbtproj.h:16-38 for the two POD types, dbtproj.h:55,68,82,96 for the four
 * operator prototypes). A caller written against the reference header (main.cpp) compiles
 * against this one unchanged and links against libdbt_b200.so instead of DatabaseProject.o.
 *
 * Facts the build relies on (verified by static_assert below):
 *   - record_t is 140 bytes, 4-byte aligned; block_t is 14016 bytes (= 876 x 16).
 *   - `field` is an ASCII character '0'..'3' (reference: DatabaseProject.cpp:23-40,207-214),
 *     not an integer 0..3.
 *   - MergeSort's `outfile` is an OUT buffer (>= 30 bytes) that receives "segment<N>.bin"
 *     (reference: DatabaseProject.cpp:375-376); for the other three operators it is the
 *     path to create.
 *   - `buffer` is ignored by every operator (reference: DatabaseProject.cpp:109,182,402,509).
 */
#ifndef _DBTPROJ_H
#define _DBTPROJ_H

#define STR_LENGTH 120
#define MAX_RECORDS_PER_BLOCK 100

/* One tuple. Key fields: recid ('0'), num ('1'), str ('2'), (num, str) ('3'). */
typedef struct {
    unsigned int recid;        /* @0   */
    unsigned int num;          /* @4   */
    char         str[STR_LENGTH]; /* @8, NUL-terminated, bytes after the NUL are don't-care */
    bool         valid;        /* @128 */
    unsigned int dummy1;       /* @132 */
    unsigned int dummy2;       /* @136 */
} record_t;

/* One I/O unit of a block file: a flat array of these, no file header. */
typedef struct {
    unsigned int  blockid;     /* @0     */
    unsigned int  nreserved;   /* @4, number of live entries (<= MAX_RECORDS_PER_BLOCK) */
    record_t      entries[MAX_RECORDS_PER_BLOCK]; /* @8 */
    bool          valid;       /* @14008 */
    unsigned char misc;        /* @14009 */
    unsigned int  dummy;       /* @14012 */
} block_t;

#ifdef __cplusplus
static_assert(sizeof(record_t) == 140, "record_t must be 140 bytes");
static_assert(sizeof(block_t) == 14016, "block_t must be 14016 bytes");
static_assert(__builtin_offsetof(record_t, str) == 8 && __builtin_offsetof(record_t, valid) == 128 &&
              __builtin_offsetof(record_t, dummy1) == 132 && __builtin_offsetof(record_t, dummy2) == 136,
              "record_t field offsets");
static_assert(__builtin_offsetof(block_t, entries) == 8 && __builtin_offsetof(block_t, valid) == 14008 &&
              __builtin_offsetof(block_t, misc) == 14009 && __builtin_offsetof(block_t, dummy) == 14012,
              "block_t field offsets");
#endif

/*
 * The four operators (C++ linkage on purpose: the reference header has no extern "C", so the
 * symbols a caller imports are the Itanium-mangled ones, e.g. _Z9MergeSortPchP7block_tjS_PjS2_S2_).
 *
 * field        '0' recid | '1' num | '2' str (strcmp order) | '3' num then str
 * buffer       ignored (may be NULL)
 * nmem_blocks  must be > 2; it no longer bounds memory, it only determines the reported
 *              nsorted_segs / npasses / nios of the external sort the reference would have run
 * outputs      always written on success
 */
void MergeSort(char *infile, unsigned char field, block_t *buffer, unsigned int nmem_blocks,
               char *outfile, unsigned int *nsorted_segs, unsigned int *npasses, unsigned int *nios);

void EliminateDuplicates(char *infile, unsigned char field, block_t *buffer, unsigned int nmem_blocks,
                         char *outfile, unsigned int *nunique, unsigned int *nios);

void MergeJoin(char *infile1, char *infile2, unsigned char field, block_t *buffer,
               unsigned int nmem_blocks, char *outfile, unsigned int *nres, unsigned int *nios);

void HashJoin(char *infile1, char *infile2, unsigned char field, block_t *buffer,
              unsigned int nmem_blocks, char *outfile, unsigned int *nres, unsigned int *nios);

#endif /* _DBTPROJ_H */
