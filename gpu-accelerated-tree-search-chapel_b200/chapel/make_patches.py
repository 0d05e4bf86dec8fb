#!/usr/bin/env python
"""Produce chapel/patches/*.diff: the edits that put libtsb200 behind the reference's GPU drivers, as unified diffs
against the reference tree (apply with `patch -p1 < x.diff` in its root).  The edits are mechanical: the device array
declarations become one tsb_*_create (+ tsb_*_register_host of the two long-lived chunk arrays), the three offload
statements become one tsb_*_evaluate, a tsb_*_destroy follows the loop, and in the multi-GPU drivers every task
pins itself next to its GPU.  NOT compile-tested (no Chapel compiler in the build image); the same call sequence is
what csrc/tsb_host.cpp executes under test.  Run in the build container:  python chapel/make_patches.py"""
import difflib
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "patches")


def sub(text, old, new, count=1):
    assert text.count(old) >= 1, old
    return text.replace(old, new, count)


NQ_DECL = """  on device var parents_d: [0..#M] Node;
  on device var labels_d: [0..#(M*N)] uint(8);
"""
NQ_EVAL = """      parents_d = parents; // host-to-device
      on device do evaluate_gpu(parents_d, numLabels, labels_d); // GPU kernel
      labels = labels_d; // device-to-host
"""
PF_DECL_HEAD = """  on device var parents_d: [0..#M] Node;
  on device var bounds_d: [0..#(M*jobs)] int(32);

  on device var lbound1_d = new lb1_bound_data(jobs, machines);
  lbound1_d.p_times   = lbound1.p_times;
  lbound1_d.min_heads = lbound1.min_heads;
  lbound1_d.min_tails = lbound1.min_tails;

  on device var lbound2_d = new lb2_bound_data(jobs, machines);
  lbound2_d.johnson_schedules  = lbound2.johnson_schedules;
  lbound2_d.lags               = lbound2.lags;
  lbound2_d.machine_pairs      = lbound2.machine_pairs;
  lbound2_d.machine_pair_order = lbound2.machine_pair_order;
"""


def indent(s, n):
    return "".join((" " * n + ln if ln.strip() else ln) for ln in s.splitlines(True))


def nq_create(dev, ind):
    return indent(f"""  var h: c_ptr(tsb_nq);  // libtsb200: owns what `on device var parents_d, labels_d` owned
  tsbCheck(tsb_nq_create(h, {dev}:c_int, N:c_int, g:c_int, M:c_int), "tsb_nq_create");
  // the two chunk arrays live for the whole step 2: page-lock them once (zero-copy offloads)
  tsbCheck(tsb_nq_register_host(h, c_ptrTo(parents[0]):c_ptr(void), (M * c_sizeof(Node)):c_size_t), "register parents");
  tsbCheck(tsb_nq_register_host(h, c_ptrTo(labels[0]):c_ptr(void), (M * N):c_size_t), "register labels");
""", ind)


def nq_eval(ind):
    return indent("""      tsbCheck(tsb_nq_evaluate(h, c_ptrToConst(parents[0]):c_ptrConst(void), poolSize:c_int, c_ptrTo(labels[0])),
               "tsb_nq_evaluate");  // H2D + evaluate_gpu + D2H of the live prefix
""", ind)


def pf_create(dev, ind):
    return indent(f"""  var h: c_ptr(tsb_pfsp);  // libtsb200: owns parents_d / bounds_d / lbound1_d / lbound2_d (tables are copied)
  tsbCheck(tsb_pfsp_create(h, {dev}:c_int, jobs:c_int, machines:c_int, M:c_int,
      c_ptrToConst(lbound1.p_times[0]), c_ptrToConst(lbound1.min_heads[0]), c_ptrToConst(lbound1.min_tails[0]),
      lbound2.nb_machine_pairs:c_int, c_ptrToConst(lbound2.johnson_schedules[0]), c_ptrToConst(lbound2.lags[0]),
      c_ptrToConst(lbound2.machine_pairs[0][0]), c_ptrToConst(lbound2.machine_pairs[1][0]),
      c_ptrToConst(lbound2.machine_pair_order[0])), "tsb_pfsp_create");
  tsbCheck(tsb_pfsp_register_host(h, c_ptrTo(parents[0]):c_ptr(void), (M * c_sizeof(Node)):c_size_t), "register parents");
  tsbCheck(tsb_pfsp_register_host(h, c_ptrTo(bounds[0]):c_ptr(void), (M * jobs * 4):c_size_t), "register bounds");
  const lbKind = tsbLbKind(lb);
""", ind)


def pf_eval(best, ind):
    return indent(f"""      tsbCheck(tsb_pfsp_evaluate(h, lbKind, c_ptrToConst(parents[0]):c_ptrConst(void), poolSize:c_int,
                                 {best}:int(64), c_ptrTo(bounds[0])), "tsb_pfsp_evaluate");
""", ind)


def patch(name, edit):
    src = open(os.path.join(REF, name)).read()
    dst = edit(src)
    d = difflib.unified_diff(src.splitlines(True), dst.splitlines(True), "a/" + name, "b/" + name, n=2)
    open(os.path.join(OUT, name.replace(".chpl", ".diff")), "w").write("".join(d))


def nq_gpu(s):
    s = sub(s, "use NQueens_node;\n", "use NQueens_node;\nuse TSB200, CTypes;\n")
    s = sub(s, NQ_DECL, nq_create("0", 0))
    s = sub(s, NQ_EVAL, nq_eval(0))
    return sub(s, "  timer.stop();\n  const res2 =", "  tsb_nq_destroy(h);\n\n  timer.stop();\n  const res2 =")


def nq_multi(s):
    s = sub(s, "use NQueens_node;\n", "use NQueens_node;\nuse TSB200, CTypes;\n")
    s = sub(s, "    const device = here.gpus[gpuID];\n",
            "    tsb_bind_thread_to_device(gpuID:c_int);  // this task next to its GPU (NUMA); one qthreads worker per task\n")
    s = sub(s, indent(NQ_DECL, 2), nq_create("gpuID", 2))
    s = sub(s, indent(NQ_EVAL, 2), nq_eval(2))
    return sub(s, "    const poolLocSize = pool_loc.size;\n", "    tsb_nq_destroy(h);\n\n    const poolLocSize = pool_loc.size;\n")


def pf_gpu(s):
    s = sub(s, "use Taillard;\n", "use Taillard;\nuse TSB200, CTypes;\n")
    s = sub(s, PF_DECL_HEAD, pf_create("0", 0))
    s = sub(s, """      parents_d = parents; // host-to-device
      on device do evaluate_gpu(parents_d, numBounds, best, lbound1_d, lbound2_d, bounds_d); // GPU kernel
      bounds = bounds_d; // device-to-host
""", pf_eval("best", 0))
    return sub(s, "  timer.stop();\n  const res2 =", "  tsb_pfsp_destroy(h);\n\n  timer.stop();\n  const res2 =")


def pf_multi(s):
    s = sub(s, "use Taillard;\n", "use Taillard;\nuse TSB200, CTypes;\n")
    s = sub(s, "    const device = here.gpus[gpuID];\n",
            "    tsb_bind_thread_to_device(gpuID:c_int);  // this task next to its GPU (NUMA); one qthreads worker per task\n")
    s = sub(s, indent(PF_DECL_HEAD, 2), pf_create("gpuID", 2))
    s = sub(s, """        parents_d = parents; // host-to-device
        on device do evaluate_gpu(parents_d, numBounds, best_l, lbound1_d, lbound2_d, bounds_d); // GPU kernel
        bounds = bounds_d; // device-to-host
""", pf_eval("best_l", 2))
    return sub(s, "    const poolLocSize = pool_loc.size;\n", "    tsb_pfsp_destroy(h);\n\n    const poolLocSize = pool_loc.size;\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    patch("nqueens_gpu_chpl.chpl", nq_gpu)
    patch("nqueens_multigpu_chpl.chpl", nq_multi)
    patch("pfsp_gpu_chpl.chpl", pf_gpu)
    patch("pfsp_multigpu_chpl.chpl", pf_multi)
    for f in sorted(os.listdir(OUT)):
        print(f, sum(1 for _ in open(os.path.join(OUT, f))), "lines")
