/* A caller without Python or Java: FASTA -> .gcz/.gcx on the GPU, then one pattern counted and located.
 *
 *   gcc -std=c11 -Iinclude examples/index_and_count.c -Lgecoz_b200 -lgcz_b200 -Wl,-rpath,$PWD/gecoz_b200 -o /tmp/index_and_count
 *   /tmp/index_and_count genome.fa genome.gcz ACGTACGTACGTACGT
 *
 * This is the flow of `gecotools -i genome.fa -o genome.gcz` followed by `gecotools -i genome.gcz -s PATTERN`
 * (tools/Gecotools.java:112-183 of the reference), through include/gcz_file.h. */
#include "gcz_file.h"

#include <stdio.h>
#include <string.h>

static int die(const char* what) {
    fprintf(stderr, "%s: %s\n", what, gcz_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s in.fa out.gcz PATTERN\n", argv[0]);
        return 2;
    }
    gcz_fasta* fa = NULL;
    if (gcz_fasta_open(argv[1], &fa) != GCZ_OK) return die("reading the FASTA file");
    gcz_index_report rep;
    const int rc = gcz_index_fasta(fa, argv[2], NULL, 32, 0, NULL, NULL, &rep);      /* device 0, CUDA engine */
    gcz_fasta_close(fa);
    if (rc != GCZ_OK) return die("building the index");
    printf("%lld sequences in %lld blocks, %lld symbols, %.3f s\n", (long long)rep.sequences, (long long)rep.blocks,
           (long long)rep.symbols, rep.seconds);

    gcz_reader* rd = NULL;
    if (gcz_reader_open(argv[2], &rd) != GCZ_OK) return die("opening the index");
    char* text = NULL;
    int64_t len = 0;
    if (gcz_match(rd, 0, NULL, (const uint8_t*)argv[3], (int64_t)strlen(argv[3]), 1, NULL, &text, &len) != GCZ_OK) {
        gcz_reader_close(rd);
        return die("searching");
    }
    fwrite(text, 1, (size_t)len, stdout);
    gcz_free(text);
    gcz_reader_close(rd);
    return 0;
}
