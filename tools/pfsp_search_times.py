import sys, time
sys.path[:0] = ['.', 'gpu-accelerated-tree-search-chapel_b200']
import tsb200
for inst, lb in ((20, "lb2"), (14, "lb1"), (20, "lb1_d")):
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        st = tsb200.pfsp_search_device_part(inst, lb, 1, 25, 50000, 1, 0, 0)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(inst, lb, st.explored_tree, st.offloads, ["%.1f" % t for t in ts], "steps(ms): %.1f %.1f %.1f" % (st.t_step1 * 1e3, st.t_step2 * 1e3, st.t_step3 * 1e3))
