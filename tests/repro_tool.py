"""Debug helper (not a test): run one case through the CUDA path and print a summary.
usage: python tests/repro_tool.py <case> <rng_mode> [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402
import util  # noqa: E402

case = cases.CASES[sys.argv[1]]()
d = "/dev/shm/repro"
os.makedirs(d, exist_ok=True)
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 12345
if int(sys.argv[2]) == 0 and len(sys.argv) > 3:
    case["kw"] = dict(case["kw"], cg=cases.cli_seed(seed))
s, st, p = util.run_cuda(case, d, sys.argv[1], rng_mode=int(sys.argv[2]), philox_seed=seed)
print("ok", sys.argv[1:], list(st.wrote), list(st.bytes))
