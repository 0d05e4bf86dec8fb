/* Counts the nodes of the N-Queens search tree per depth (= explored-tree nodes the reference pushes,
 * nqueens_chpl.chpl:77-87: a child at depth d+1 exists for every queen that isSafe on row d), with a
 * bitboard DFS.  gcc -O3 -fopenmp make_depth_hist.c -o /tmp/dh && /tmp/dh 8 18 > nqueens_depth_hist.json
 * The per-depth sums reproduce the reference's "Size of the explored tree" (counts.json). */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
static void dfs(int N, int d, uint32_t cols, uint32_t ld, uint32_t rd, uint64_t* hist) {
  uint32_t full = (1u << N) - 1, free_ = ~(cols | ld | rd) & full;
  while (free_) {
    uint32_t b = free_ & -free_;
    free_ ^= b;
    hist[d + 1]++;
    if (d + 1 < N) dfs(N, d + 1, cols | b, ((ld | b) << 1) & full, (rd | b) >> 1, hist);
  }
}
int main(int argc, char** argv) {
  int lo = atoi(argv[1]), hi = atoi(argv[2]);
  printf("{\n");
  for (int N = lo; N <= hi; N++) {
    uint64_t hist[32] = {0};
    uint64_t part[32][32] = {{0}};
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int a = 0; a < N; a++)
      for (int b = 0; b < N; b++) {
        if (b == a || b == a + 1 || b == a - 1) continue;
        uint32_t full = (1u << N) - 1, ba = 1u << a, bb = 1u << b;
        uint64_t h[32] = {0};
        uint32_t ld = ((ba << 1) & full), rd = ba >> 1;
        h[2] = 1;
        if (N > 2) dfs(N, 2, ba | bb, ((ld | bb) << 1) & full, (rd | bb) >> 1, h);
#pragma omp critical
        for (int d = 0; d < 32; d++) hist[d] += h[d];
      }
    hist[1] = N;
    (void)part;
    uint64_t tot = 0;
    printf(" \"%d\": {", N);
    for (int d = 1; d <= N; d++) { printf("%s\"%d\": %llu", d > 1 ? ", " : "", d, (unsigned long long)hist[d]); tot += hist[d]; }
    printf("}%s\n", N < hi ? "," : "");
    fprintf(stderr, "N=%d tree=%llu sol=%llu\n", N, (unsigned long long)tot, (unsigned long long)hist[N]);
  }
  printf("}\n");
  return 0;
}
