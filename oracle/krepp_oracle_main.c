/* oracle/krepp_oracle_main.c -- TEST INFRASTRUCTURE ONLY: tiny CLI around the C restatement (dist TSV body).
 * usage: krepp_oracle dist INDEX_DIR READS.fq [nthreads]        (plain-text FASTA/FASTQ only) */
#include "krepp_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char** argv)
{
  if (argc < 4 || strcmp(argv[1], "dist")) { fprintf(stderr, "usage: %s dist INDEX_DIR READS [nthreads]\n", argv[0]); return 2; }
  char err[512] = {0};
  ko_index_t* ix = ko_index_load(argv[2], err, sizeof err);
  if (!ix) { fprintf(stderr, "[ERROR] %s\n", err); return 1; }
  FILE* f = fopen(argv[3], "rb");
  if (!f) { fprintf(stderr, "[ERROR] Failed to open the file at %s\n", argv[3]); return 1; }
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  char* txt = (char*)malloc((size_t)sz + 1);
  if (fread(txt, 1, (size_t)sz, f) != (size_t)sz) return 1;
  fclose(f);
  ko_seqrec_t* recs; int64_t n = ko_parse_reads(txt, (size_t)sz, &recs);
  ko_params_t p = {4, 2.706, NAN, 2, 1, 1, 0, 0};
  char* tsv = ko_dist_tsv(ix, &p, recs, n, argc > 4 ? atoi(argv[4]) : 1);
  printf("SEQ_ID\tREFERENCE_NAME\tDIST\n%s", tsv);
  free(tsv); ko_free_reads(recs, n); free(txt); ko_index_free(ix);
  return 0;
}
