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

/* ---- the reference's sequential SEARCH (nqueens_search, baselines/nqueens/nqueens_c.c:112-141: popBack +
 * decompose until the pool is empty) started from given nodes instead of the root, so that several host threads
 * can each explore a share of the tree with the reference's own pool, isSafe and decompose.  bench.py's CPU legs
 * hand out the subtrees of a frontier (ref_nq_frontier) to one thread per host core. */
#include "lib/Pool.h"
void decompose(const int N, const int G, const Node parent, unsigned long long int* tree_loc,
               unsigned long long int* num_sol, SinglePool* pool);

void ref_nq_search_from(int N, int G, const Node* nodes, int n, unsigned long long* tree, unsigned long long* sol) {
  SinglePool pool;
  initSinglePool(&pool);
  for (int i = 0; i < n; i++) pushBack(&pool, nodes[i]);
  while (1) {
    int hasWork = 0;
    Node parent = popBack(&pool, &hasWork);
    if (!hasWork) break;
    decompose(N, G, parent, tree, sol, &pool);
  }
  deleteSinglePool(&pool);
}
/* all nodes of depth `depth` (breadth first from the root, with the reference's decompose); returns their number
 * (-1 if `cap` is too small); *tree / *sol count what was explored on the way */
int ref_nq_frontier(int N, int G, int depth, Node* out, int cap, unsigned long long* tree, unsigned long long* sol) {
  SinglePool pool;
  initSinglePool(&pool);
  Node root;
  initRoot(&root, N);
  pushBack(&pool, root);
  int n = 0;
  while (1) {
    int hasWork = 0;
    Node parent = popFront(&pool, &hasWork);
    if (!hasWork) break;
    if (parent.depth >= depth) {
      if (n >= cap) {
        deleteSinglePool(&pool);
        return -1;
      }
      out[n++] = parent;
    } else {
      decompose(N, G, parent, tree, sol, &pool);
    }
  }
  deleteSinglePool(&pool);
  return n;
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
