/*
  TSB200.chpl — Chapel binding of libtsb200.so (include/tsb200.h).

  NOT compile-tested: no Chapel compiler exists in the build image (SURVEY.md fact 1).  It is kept
  minimal and mechanical on purpose: extern declarations + two thin wrappers that halt on error,
  mirroring 1:1 the C++ drivers (csrc/tsb_host.cpp) that ARE tested.

  Build the patched drivers with, e.g.
    chpl --fast -M lib/commons -M lib/nqueens -M <this dir> nqueens_gpu_chpl.chpl \
         -I<repo>/include -L<repo>/gpu-accelerated-tree-search-chapel_b200 -ltsb200
  CHPL_LOCALE_MODEL=flat is sufficient (here.gpus is no longer needed); with CHPL_LOCALE_MODEL=gpu the
  library shares the primary CUDA context of the Chapel runtime (it only uses the CUDA runtime API).
*/
module TSB200 {
  use CTypes;

  require "tsb200.h", "-ltsb200";

  extern type tsb_nq;    // opaque
  extern type tsb_pfsp;  // opaque

  extern const TSB_OK: c_int;
  extern const TSB_LB1_D: c_int;  // 0   (encoding of baselines/pfsp/pfsp_c.c:86-88)
  extern const TSB_LB1: c_int;    // 1
  extern const TSB_LB2: c_int;    // 2

  extern proc tsb_strerror(code: c_int): c_ptrConst(c_char);
  extern proc tsb_last_cuda_error(): c_ptrConst(c_char);
  extern proc tsb_device_count(): c_int;
  extern proc tsb_init_devices(n: c_int): c_int;
  extern proc tsb_bind_thread_to_device(device: c_int): c_int;  // first statement of every per-GPU task

  extern proc tsb_nq_create(ref h: c_ptr(tsb_nq), device: c_int, N: c_int, g: c_int, M_max: c_int): c_int;
  extern proc tsb_nq_destroy(h: c_ptr(tsb_nq)): void;
  extern proc tsb_nq_evaluate(h: c_ptr(tsb_nq), parents: c_ptrConst(void), count: c_int,
                              labels: c_ptr(uint(8))): c_int;

  // optional, once per search: page-lock + map the driver's long-lived `parents` / `labels` (`bounds`) arrays
  // so that tsb_*_evaluate works on them in place (zero-copy); they must outlive the handle
  extern proc tsb_nq_register_host(h: c_ptr(tsb_nq), p: c_ptr(void), bytes: c_size_t): c_int;
  extern proc tsb_nq_unregister_host(h: c_ptr(tsb_nq), p: c_ptr(void)): c_int;
  extern proc tsb_pfsp_register_host(h: c_ptr(tsb_pfsp), p: c_ptr(void), bytes: c_size_t): c_int;
  extern proc tsb_pfsp_unregister_host(h: c_ptr(tsb_pfsp), p: c_ptr(void)): c_int;

  extern proc tsb_pfsp_create(ref h: c_ptr(tsb_pfsp), device: c_int, jobs: c_int, machines: c_int,
                              M_max: c_int, p_times: c_ptrConst(int(32)), min_heads: c_ptrConst(int(32)),
                              min_tails: c_ptrConst(int(32)), nb_pairs: c_int,
                              johnson: c_ptrConst(int(32)), lags: c_ptrConst(int(32)),
                              mp0: c_ptrConst(int(32)), mp1: c_ptrConst(int(32)),
                              mp_order: c_ptrConst(int(32))): c_int;
  // a build with `-sMAX_JOBS=50` (208-byte nodes, ta031..ta060) passes max_jobs = MAX_JOBS here
  extern proc tsb_pfsp_create_wide(ref h: c_ptr(tsb_pfsp), device: c_int, max_jobs: c_int, jobs: c_int, machines: c_int,
                                   M_max: c_int, p_times: c_ptrConst(int(32)), min_heads: c_ptrConst(int(32)),
                                   min_tails: c_ptrConst(int(32)), nb_pairs: c_int,
                                   johnson: c_ptrConst(int(32)), lags: c_ptrConst(int(32)),
                                   mp0: c_ptrConst(int(32)), mp1: c_ptrConst(int(32)),
                                   mp_order: c_ptrConst(int(32))): c_int;
  extern proc tsb_pfsp_destroy(h: c_ptr(tsb_pfsp)): void;
  extern proc tsb_pfsp_evaluate(h: c_ptr(tsb_pfsp), lb_kind: c_int, parents: c_ptrConst(void), count: c_int,
                                best: int(64), bounds: c_ptr(int(32))): c_int;

  // ---- beyond the three offload lines: fused evaluate + generate_children, device-resident pool
  // (INTEGRATION.md section 3b; SURVEY.md section 8f rows 1 and 3)
  extern proc tsb_nq_expand(h: c_ptr(tsb_nq), parents: c_ptrConst(void), count: c_int, children: c_ptr(void),
                            capacity_nodes: uint(64), ref n_children: uint(64), ref n_solutions: uint(64)): c_int;
  extern proc tsb_nq_pool_push(h: c_ptr(tsb_nq), nodes: c_ptrConst(void), n: int(64)): c_int;
  extern proc tsb_nq_pool_size(h: c_ptr(tsb_nq)): int(64);
  extern proc tsb_nq_pool_step(h: c_ptr(tsb_nq), m: c_int, M: c_int, ref n_parents: int(64),
                               ref n_children: uint(64), ref n_solutions: uint(64)): c_int;
  extern proc tsb_nq_pool_steal(victim: c_ptr(tsb_nq), thief: c_ptr(tsb_nq), m: c_int, ref n_stolen: int(64)): c_int;
  extern proc tsb_pfsp_pool_steal(victim: c_ptr(tsb_pfsp), thief: c_ptr(tsb_pfsp), m: c_int, ref n_stolen: int(64)): c_int;
  extern proc tsb_nq_pool_run(h: c_ptr(tsb_nq), m: c_int, M: c_int, max_rounds: int(64), ref n_rounds: uint(64),
                              ref n_parents: uint(64), ref n_children: uint(64), ref n_solutions: uint(64)): c_int;
  extern proc tsb_nq_pool_run_multi(handles: c_ptr(c_ptr(tsb_nq)), n_pools: c_int, m: c_int, M: c_int,
                                    max_rounds: int(64), outCounts: c_ptr(uint(64))): c_int;
  extern proc tsb_nq_sibling(h: c_ptr(tsb_nq), index: c_int, ref sibling: c_ptr(tsb_nq)): c_int;
  extern proc tsb_nq_pools_per_launch(h: c_ptr(tsb_nq), M: c_int): c_int;
  extern proc tsb_nq_pool_drain(h: c_ptr(tsb_nq), nodes: c_ptr(void), capacity_nodes: int(64),
                                ref n: int(64)): c_int;
  extern proc tsb_pfsp_expand(h: c_ptr(tsb_pfsp), lb_kind: c_int, parents: c_ptrConst(void), count: c_int,
                              ref best: int(64), children: c_ptr(void), capacity_nodes: uint(64),
                              ref n_children: uint(64), ref n_solutions: uint(64)): c_int;
  extern proc tsb_pfsp_pool_push(h: c_ptr(tsb_pfsp), nodes: c_ptrConst(void), n: int(64)): c_int;
  extern proc tsb_pfsp_pool_size(h: c_ptr(tsb_pfsp)): int(64);
  extern proc tsb_pfsp_pool_step(h: c_ptr(tsb_pfsp), lb_kind: c_int, m: c_int, M: c_int, ref best: int(64),
                                 ref n_parents: int(64), ref n_children: uint(64),
                                 ref n_solutions: uint(64)): c_int;
  extern proc tsb_pfsp_pool_drain(h: c_ptr(tsb_pfsp), nodes: c_ptr(void), capacity_nodes: int(64),
                                  ref n: int(64)): c_int;

  proc tsbCheck(rc: c_int, what: string) {
    if rc != TSB_OK then
      halt(what, ": ", string.createCopyingBuffer(tsb_strerror(rc)), " — ",
           string.createCopyingBuffer(tsb_last_cuda_error()));
  }

  // "lb1_d" | "lb1" | "lb2"  (pfsp_gpu_chpl.chpl:15)  ->  C encoding
  proc tsbLbKind(lb: string): c_int {
    select lb {
      when "lb1_d" do return TSB_LB1_D;
      when "lb1" do return TSB_LB1;
      otherwise do return TSB_LB2;
    }
  }
}
