// ref_runner.cpp -- tiny command-line harness around the UNMODIFIED reference operators ("REF").
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/dbt_oracle.c header).
//
// Built by oracle/Makefile as
//   g++ -std=c++11 -O3 -I/root/reference oracle/ref_runner.cpp /root/reference/DatabaseProject.cpp
// i.e. the reference source is compiled from where it lies, nothing is copied into this repo,
// and the output goes to oracle/_ref/ (git-ignored).  The reference's own main.cpp cannot be the
// driver: it seeds with time(0) and opens file2.bin for reading.
//
// Usage (run in a FRESH scratch working directory -- the reference uses fixed file names in CWD
// and calls exit(0) on errors, reference: DatabaseProject.cpp:38-39,378-379,385-386,653-657):
//   ref_runner sort  <field> <nmem> <infile>
//   ref_runner dedup <field> <nmem> <infile> <outfile>
//   ref_runner mjoin <field> <nmem> <infile1> <infile2> <outfile>
//   ref_runner hjoin <field> <nmem> <infile1> <infile2> <outfile>
// Prints one JSON line with the operator's out-params and the wall time of the call alone.
#include "dbtproj.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: ref_runner sort|dedup|mjoin|hjoin <field> <nmem> <in1> [<in2>] [<out>]\n");
        return 2;
    }
    const char *op = argv[1];
    unsigned char field = (unsigned char)argv[2][0];
    unsigned nmem = (unsigned)strtoul(argv[3], nullptr, 10);
    unsigned a = 0, b = 0, c = 0;
    char outname[256];
    outname[0] = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (!strcmp(op, "sort")) {
        MergeSort(argv[4], field, nullptr, nmem, outname, &a, &b, &c);
    } else if (!strcmp(op, "dedup") && argc >= 6) {
        EliminateDuplicates(argv[4], field, nullptr, nmem, argv[5], &a, &c);
        snprintf(outname, sizeof outname, "%s", argv[5]);
    } else if (!strcmp(op, "mjoin") && argc >= 7) {
        MergeJoin(argv[4], argv[5], field, nullptr, nmem, argv[6], &a, &c);
        snprintf(outname, sizeof outname, "%s", argv[6]);
    } else if (!strcmp(op, "hjoin") && argc >= 7) {
        HashJoin(argv[4], argv[5], field, nullptr, nmem, argv[6], &a, &c);
        snprintf(outname, sizeof outname, "%s", argv[6]);
    } else {
        fprintf(stderr, "bad arguments\n");
        return 2;
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // a = nsorted_segs | nunique | nres ; b = npasses (sort only) ; c = nios
    printf("\n{\"op\": \"%s\", \"a\": %u, \"b\": %u, \"nios\": %u, \"outfile\": \"%s\", \"seconds\": %.6f}\n",
           op, a, b, c, outname, sec);
    return 0;
}
