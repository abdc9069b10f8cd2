"""Launch-shape sweep of the simulate kernel: python tests/tools/k3_sweep.py c3|c4 scale"""
import os, subprocess, sys, json
which, scale = sys.argv[1], sys.argv[2]
combos = [(0, 0)] + [(rw, w) for rw in (8, 16, 32) for w in (1, 2, 4)]
for rw, w in combos:
    env = dict(os.environ)
    if rw: env["QTIPSIM_RW"] = str(rw); env["QTIPSIM_WARPS"] = str(w)
    r = subprocess.run([sys.executable, "bench.py", "--workload", which, "--scale", scale, "--steps", "3", "--warmup", "2", "--no-cpu-baseline",
                        "--no-secondary", "--no-c4"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{which} rw={rw or 'dflt'} warps={w or 'dflt'}: simulate {d['phase_ms_per_step_rank0']['simulate']:.3f} ms frac {d['roofline']['frac']:.3f}", flush=True)
    except Exception as e:
        print(which, rw, w, "failed", r.stderr[-300:], flush=True)
