"""Times the replay plan phase (K2) on the benchmark workload: python tests/replay_timing_tool.py [scale] [workload]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qtip_b200 import sim, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
which = sys.argv[2] if len(sys.argv) > 2 else "c3"
ng = int(sys.argv[3]) if len(sys.argv) > 3 else 1
wl = bench.workload(ng, scale, which)
fa = synth.make_fasta(wl["records"], seed=wl["genome_seed"], lead_n=wl["lead_n"])
eng = sim.Engine()
eng.load_fasta_bytes(fa)
added = wl.get("added") or (len(wl["tu"]), len(wl["tb"]), len(wl["tc"]), len(wl["td"]))
eng.set_unpaired(0, wl["tu"], added[0])
eng.set_unpaired(1, wl["tb"], added[1])
eng.set_paired(2, wl["tc"], added[2])
eng.set_paired(3, wl["td"], added[3])
kw = wl["kw"]
for i in range(4):
    p = sim.make_params(128 * 1024, kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"], 0,
                        cg=bench.cli_seed(100 + i), shard_rank=0, shard_count=ng)
    t = time.perf_counter()
    st = eng.run(p)
    dt = time.perf_counter() - t
    print(f"run {i}: wrote {list(st.wrote)} plan {st.ms_plan:.3f} ms simulate {st.ms_simulate:.3f} ms total {st.ms_total:.3f} ms "
          f"wall {dt * 1e3:.1f} ms launches {st.kernel_launches} draws {st.draws}", flush=True)
eng.close()
