/* Plain C client of libkmergma_cuda: findGenes on a FASTA genome with a FASTA reference family.
 * Shows the call sequence the Julia shim (julia/KmerGMACuda.jl) performs, with no Python in between.
 *   gcc -O2 -Iinclude examples/findgenes.c -o findgenes -Lkmergma.jl_b200 -lkmergma_cuda -Wl,-rpath,$PWD/kmergma.jl_b200
 *   ./findgenes genome.fasta refs.fasta [k=6] [thr=30] [buffer=50]
 * Prints one header line per hit in the reference's format (src/Alignment.jl:71-78). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kmergma.h"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s genome.fasta refs.fasta [k] [thr] [buffer]\n", argv[0]); return 2; }
    const int k = argc > 3 ? atoi(argv[3]) : 6;
    const double thr = argc > 4 ? atof(argv[4]) : 30.0;
    const long buffer = argc > 5 ? atol(argv[5]) : 50;

    kgma_ctx *ctx = NULL;
    if (kgma_create(0, &ctx) != KGMA_OK) { fprintf(stderr, "kgma_create: %s\n", kgma_last_error(NULL)); return 1; }

    /* gen_ref_ws_cons (src/ReferenceGeneration.jl:4-41) */
    kgma_refs *refs = NULL;
    if (kgma_refs_from_fasta(argv[2], &refs) != KGMA_OK) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    const size_t nb = (size_t)1 << (2 * k);
    int32_t *S = calloc(nb, sizeof *S), n_refs = 0;
    int64_t window = 0, maxlen = kgma_refs_maxlen(refs);
    char *consensus = calloc((size_t)maxlen + 2, 1);
    if (kgma_refs_profile(refs, k, S, &n_refs, &window, consensus) != KGMA_OK) { fprintf(stderr, "profile failed\n"); return 1; }

    /* FASTA.Reader + getSeq + NUCLEOTIDE_BITS (src/Consts.jl:22-39) */
    kgma_genome *g = NULL;
    if (kgma_genome_from_fasta(argv[1], &g) != KGMA_OK) { fprintf(stderr, "cannot ingest %s\n", argv[1]); return 1; }

    /* ac_gma_testing! (src/GenomeMiner.jl:4-109) */
    kgma_profile prof = { k, n_refs, window, S, consensus, (int32_t)strlen(consensus), thr };
    kgma_scan_params par = { KGMA_MODE_SINGLE, KGMA_F_ALIGN, buffer, -69, -1, 0, 1, -1, 0 };
    kgma_result *res = NULL;
    if (kgma_scan(ctx, g, &prof, 1, &par, &res) != KGMA_OK) { fprintf(stderr, "kgma_scan: %s\n", kgma_last_error(ctx)); return 1; }

    const kgma_hit *h = kgma_result_hits(res);
    for (int64_t i = 0; i < kgma_result_n_hits(res); i++)
        printf("%s | D = %lld/%lld | MatchPos = %lld:%lld | GenomePos = %lld | Len = %lld | flags = %u\n",
               kgma_genome_identifier(g, h[i].record), (long long)h[i].D, 2LL * k * n_refs * n_refs,
               (long long)h[i].first, (long long)h[i].last, (long long)h[i].genome_pos,
               (long long)(h[i].last - h[i].first + 1), h[i].flags);

    kgma_result_free(res); kgma_genome_destroy(g); kgma_refs_destroy(refs); kgma_destroy(ctx);
    free(S); free(consensus);
    return 0;
}
