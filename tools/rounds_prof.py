import json, sys, time
sys.path.insert(0, "gpu-accelerated-tree-search-chapel_b200")
import tsb200
tsb200.lib().tsb_init_devices(1)
N, M = int(sys.argv[1]), int(sys.argv[2])
for _ in range(2):
    t0 = time.perf_counter(); st = tsb200.nqueens_search_device(N, 1, 25, M, 1); dt = time.perf_counter() - t0
print(json.dumps({"N": N, "M": M, "seconds": dt, "us_per_round": dt / st.offloads * 1e6, "Mnodes_s": st.explored_tree / dt / 1e6}))
