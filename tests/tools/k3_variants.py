"""Simulate-kernel timing over library variants / tile sizes: python tests/tools/k3_variants.py"""
import os, subprocess, sys, json, glob
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
runs = [("default", {}), ("rw20", {"QTIPSIM_RW": "20", "QTIPSIM_WARPS": "1"}), ("rw24", {"QTIPSIM_RW": "24", "QTIPSIM_WARPS": "1"}),
        ("rw12", {"QTIPSIM_RW": "12", "QTIPSIM_WARPS": "1"})]
for lib in sorted(glob.glob(os.path.join(ROOT, "qtip_b200", "libqtipsim_*.so"))):
    if "hosttest" in lib or "dbg" in lib: continue
    runs.append((os.path.basename(lib), {"QTIPSIM_LIB": lib}))
for which in ("c3", "c4"):
    for tag, envx in runs:
        env = dict(os.environ); env.update(envx)
        r = subprocess.run([sys.executable, "bench.py", "--workload", which, "--scale", "0.5", "--steps", "3", "--warmup", "2", "--no-cpu-baseline",
                            "--no-secondary", "--no-c4"], env=env, capture_output=True, text=True, cwd=ROOT)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{which} {tag}: simulate {d['phase_ms_per_step_rank0']['simulate']:.3f} ms frac {d['roofline']['frac']:.3f}", flush=True)
        except Exception:
            print(which, tag, "failed", r.stderr[-200:], flush=True)
