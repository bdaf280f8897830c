// dbt_main.cpp -- command-line driver for the drop-in library, covering the workflow of the
// reference's main.cpp (generate two block files, MergeJoin, then HashJoin on MergeJoin's side
// files; reference main.cpp:36-79,109-123) with the things that driver cannot do as shipped
// fixed: a seed flag instead of time(0) (main.cpp:22), file2.bin opened for writing (main.cpp:37),
// all four operators selectable, nmem_blocks and field as flags.  It only talks to the library
// through include/dbtproj.h, exactly like the reference's main.cpp talks to DatabaseProject.o.
//
//   dbt_main [--nblocks N] [--seed S] [--field 0|1|2|3] [--nmem M] [--ops sort,dedup,mjoin,hjoin] [--keep]
#include "dbtproj.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static void random_letters(char *s, int len) { // reference: main.cpp:10-18
    for (int i = 0; i < len; ++i) s[i] = (char)('a' + rand() % 26);
    s[len] = 0;
}

static void generate(const char *name1, const char *name2, int nblocks, unsigned seed) {
    // reference: main.cpp:41-77 -- per row: num1, 5 letters, num2, 5 letters; "Hola" at row 1 of every block
    srand(seed);
    FILE *f1 = fopen(name1, "wb"), *f2 = fopen(name2, "wb");
    if (!f1 || !f2) { perror("fopen"); exit(1); }
    block_t *b1 = (block_t *)calloc(1, sizeof(block_t)), *b2 = (block_t *)calloc(1, sizeof(block_t));
    unsigned recid = 0;
    char s[16];
    for (int b = 0; b < nblocks; ++b) {
        memset(b1, 0, sizeof *b1);
        memset(b2, 0, sizeof *b2);
        for (int r = 0; r < MAX_RECORDS_PER_BLOCK; ++r) {
            record_t &x = b1->entries[r], &y = b2->entries[r];
            x.recid = recid;
            x.num = (unsigned)rand() % (unsigned)(nblocks * 30);
            random_letters(s, 5);
            strcpy(x.str, s);
            y.recid = recid++;
            y.num = (unsigned)rand() % (unsigned)(nblocks * 30);
            random_letters(s, 5);
            strcpy(y.str, s);
            if (r == 1) { strcpy(x.str, "Hola"); strcpy(y.str, "Hola"); }
            x.valid = y.valid = true;
        }
        b1->blockid = b2->blockid = (unsigned)b;
        b1->nreserved = b2->nreserved = MAX_RECORDS_PER_BLOCK;
        b1->valid = b2->valid = true;
        b1->dummy = b2->dummy = MAX_RECORDS_PER_BLOCK;
        fwrite(b1, 1, sizeof *b1, f1);
        fwrite(b2, 1, sizeof *b2, f2);
    }
    fclose(f1); fclose(f2); free(b1); free(b2);
}

template <class F> static double timed(F f) {
    auto t0 = std::chrono::steady_clock::now();
    f();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char **argv) {
    int nblocks = 6000; // reference default, main.cpp:23
    unsigned seed = 42, nmem = 100;
    unsigned char field = '1';
    std::string ops = "mjoin,hjoin"; // what the reference's main.cpp runs
    bool keep = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return (i + 1 < argc) ? argv[++i] : (char *)""; };
        if (a == "--nblocks") nblocks = atoi(next());
        else if (a == "--seed") seed = (unsigned)atoi(next());
        else if (a == "--field") field = (unsigned char)next()[0];
        else if (a == "--nmem") nmem = (unsigned)atoi(next());
        else if (a == "--ops") ops = next();
        else if (a == "--keep") keep = true;
        else if (i == 1 && a[0] != '-') nblocks = atoi(argv[1]); // `dbt <nblocks>` like the reference
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    printf("Creating input files... (%d blocks each, seed %u)\n", nblocks, seed);
    char file1[] = "file.bin", file2[] = "file2.bin";
    generate(file1, file2, nblocks, seed);
    unsigned nios = 0, a = 0, b = 0;
    const long total = (long)nblocks * MAX_RECORDS_PER_BLOCK;
    auto has = [&](const char *op) { return ops.find(op) != std::string::npos; };
    if (has("sort")) {
        char out[64];
        double s = timed([&] { MergeSort(file1, field, nullptr, nmem, out, &a, &b, &nios); });
        printf("MERGE SORT: passes %u, sorted segments %u, IOs %u, outfile %s  (%.3f s, %.1f M records/s)\n", b, a, nios, out, s,
               total / s / 1e6);
    }
    if (has("dedup")) {
        char out[] = "NOduplicates.bin";
        double s = timed([&] { EliminateDuplicates(file1, field, nullptr, nmem, out, &a, &nios); });
        printf("ELIMINATE DUPLICATES: %u unique of %ld, IOs %u  (%.3f s)\n", a, total, nios, s);
    }
    if (has("mjoin")) {
        char out[] = "outmerge.bin";
        double s = timed([&] { MergeJoin(file1, file2, field, nullptr, nmem, out, &a, &nios); });
        printf("MERGE JOIN: pairs in the output %u of %ld, IOs %u  (%.3f s)\n", a, total, nios, s);
    }
    if (has("hjoin")) {
        char out[] = "outhash.bin";
        // the reference joins MergeJoin's side files (main.cpp:121); without a preceding mjoin use the raw files
        char in1[] = "1outfile.bin", in2[] = "2outfile.bin";
        bool side = has("mjoin");
        double s = timed([&] { HashJoin(side ? in1 : file1, side ? in2 : file2, field, nullptr, nmem, out, &a, &nios); });
        printf("HASH JOIN: pairs in the output %u of %ld, IOs %u  (%.3f s)\n", a, total, nios, s);
    }
    if (!keep) {
        int rc = system("rm -f segment*.bin"); // reference: main.cpp:128-132
        (void)rc;
    }
    return 0;
}
