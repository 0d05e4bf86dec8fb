"""times whole N-Queens searches with the pool on the device at the reference's default --M 50000 (persistent
multi-round kernel) and with large chunks; prints one JSON line per search"""
import json
import sys
import time

sys.path.insert(0, "gpu-accelerated-tree-search-chapel_b200")
import tsb200  # noqa: E402

tsb200.lib().tsb_init_devices(1)
for N, M in [(12, 50000), (14, 50000), (15, 50000), (16, 50000), (17, 50000), (17, 1 << 22)] + \
        [(int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[1:]]:
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        st = tsb200.nqueens_search_device(N, 1, 25, M, 1)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(json.dumps({"N": N, "M": M, "tree": st.explored_tree, "sol": st.explored_sol, "seconds": best,
                      "Mnodes_s": st.explored_tree / best / 1e6, "offloads": st.offloads,
                      "us_per_round": best / max(1, st.offloads) * 1e6, "launches": st.kernel_launches}), flush=True)
