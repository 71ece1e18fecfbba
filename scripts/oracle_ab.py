"""CPU arm A/B on this machine: the oracle as committed vs an older source (path given), same workload, same threads.
usage: python scripts/oracle_ab.py OLD_DIR [seconds] [threads]"""
import importlib.util, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import bench
from enhancershoebox_b200 import ljsoft, protocol as P

def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(path, "esb_oracle.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m

def rate(O, secs, threads, shared):
    O.build_native()
    w = bench.WORKLOADS["box12_p3_x512"]
    d, x, v, t = bench.load_snapshot(w)
    pt = P.pair_tables("genes_stages")
    secs_ = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
    lookup = [O.build_lookup_table(secs_[k], c) for (k, c) in pt.specs]
    plan = P.chunk_schedule("genes_stages", n_chunks=1001)[1000]
    kw = dict(shared_tables=True) if shared else {}
    os_ = []
    for k in range(threads):
        o = O.Oracle(d, pt, lookup, seed=1 + k, **kw); o.set_state(t, x, v); os_.append(o)
    dl = [0.0]
    def work(o):
        n = 0
        while True:
            o.run_chunk(plan); n += 1
            if time.perf_counter() >= dl[0]: return n
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda o: (o.apply_plan(plan), o.run(20)), os_))
        t0 = time.perf_counter(); dl[0] = t0 + secs
        ch = list(ex.map(work, os_)); dt = time.perf_counter() - t0
    return sum(ch) * 200 / dt

old_dir = sys.argv[1]; secs = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0; threads = int(sys.argv[3]) if len(sys.argv) > 3 else os.cpu_count()
new = load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"), "new_oracle")
old = load(old_dir, "old_oracle")
for rep in range(2):
    print(f"threads {threads}: old {rate(old, secs, threads, False):.0f}  new {rate(new, secs, threads, True):.0f} replica-steps/s", flush=True)
