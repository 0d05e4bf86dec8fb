/*
 * ref_batch.c — batch drivers around the REFERENCE's own CPU functions (test infrastructure).
 *
 * Compiled by oracle/Makefile together with the reference's C sources (from /root/reference,
 * never copied) into oracle/_ref/libref_{nqueens,pfsp}.so.  The loops below do what the
 * reference's sequential `decompose` does to evaluate the children of each parent
 * (baselines/nqueens/nqueens_c.c:89-111, baselines/pfsp/pfsp_c.c:106-191) — they call the
 * reference's isSafe / lb1_bound / lb1_children_bounds / lb2_bound — minus the pool pushes, and
 * store the values in the labels / bounds layout of the GPU path.  Used as the "reference" CPU
 * baseline of bench.py and to generate tests/golden/.
 */
#include <stdint.h>
#include <string.h>

#ifdef REF_BATCH_NQUEENS
#include "lib/NQueens_node.h"
uint8_t isSafe(const int G, const uint8_t* board, const uint8_t queen_num, const uint8_t row_pos);

void ref_nq_evaluate_range(const Node* parents, int begin, int end, int N, int G, uint8_t* labels) {
  for (int p = begin; p < end; p++) {
    const uint8_t depth = parents[p].depth;
    for (int j = depth; j < N; j++)
      labels[(size_t)p * N + j] = isSafe(G, parents[p].board, depth, parents[p].board[j]);
  }
}
/* `repeat` sweeps in one call, so that a timing thread pays the call overhead once */
void ref_nq_evaluate_range_rep(const Node* parents, int begin, int end, int N, int G, uint8_t* labels, int repeat) {
  for (int r = 0; r < repeat; r++) ref_nq_evaluate_range(parents, begin, end, N, G, labels);
}
#endif

#ifdef REF_BATCH_PFSP
#include "lib/PFSP_node.h"
#include "lib/c_bound_johnson.h"
#include "lib/c_bound_simple.h"

void ref_pfsp_evaluate_range(const lb1_bound_data* d1, const lb2_bound_data* d2, int lb, const Node* parents,
                             int begin, int end, int best, int* bounds) {
  const int jobs = d1->nb_jobs;
  for (int p = begin; p < end; p++) {
    const Node* parent = &parents[p];
    if (lb == 0) { /* decompose_lb1_d, pfsp_c.c:134-162 */
      int lb_begin[MAX_JOBS];
      lb1_children_bounds(d1, parent->prmu, parent->limit1, jobs, lb_begin);
      for (int i = parent->limit1 + 1; i < jobs; i++) bounds[(size_t)p * jobs + i] = lb_begin[parent->prmu[i]];
    } else { /* decompose_lb1 :106-132, decompose_lb2 :164-191 */
      for (int i = parent->limit1 + 1; i < jobs; i++) {
        int prmu[MAX_JOBS];
        memcpy(prmu, parent->prmu, jobs * sizeof(int));
        int x = prmu[parent->depth];
        prmu[parent->depth] = prmu[i];
        prmu[i] = x;
        bounds[(size_t)p * jobs + i] = lb == 1 ? lb1_bound(d1, prmu, parent->limit1 + 1, jobs)
                                               : lb2_bound(d1, d2, prmu, parent->limit1 + 1, jobs, best);
      }
    }
  }
}
void ref_pfsp_evaluate_range_rep(const lb1_bound_data* d1, const lb2_bound_data* d2, int lb, const Node* parents,
                                 int begin, int end, int best, int* bounds, int repeat) {
  for (int r = 0; r < repeat; r++) ref_pfsp_evaluate_range(d1, d2, lb, parents, begin, end, best, bounds);
}
#endif
