"""Generates tests/golden/golden_full.json from the UNMODIFIED reference: the BASELINE.json configurations at
their stated sizes (tests/cases.py FULL_CASES: C2 full, C5 full, C3 full, C4 at 1/10).

    python tests/golden/make_golden_full.py [case ...]

Needs /root/reference (oracle/_ref is built from it) and a few minutes of one CPU core (C3 full: ~70 s of
reference time plus ~6 GB of FASTQ on tmpfs).  Records sha256 + byte count of each of the seven FASTQ streams, the
number of records per category, the generator state at exit and sha256 of the inputs.
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from make_golden import templates_digest  # noqa: E402

OUT = os.path.join(HERE, "golden_full.json")


def file_digest(path):
    h, n, lines = hashlib.sha256(), 0, 0
    with open(path, "rb") as fh:
        while True:
            b = fh.read(1 << 24)
            if not b:
                break
            h.update(b)
            n += len(b)
            lines += b.count(b"\n")
    return {"bytes": n, "sha256": h.hexdigest()}, lines // 4


def main():
    po.build(ref=True)
    want = sys.argv[1:] or list(cases.FULL_CASES)
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as fh:
            out = json.load(fh)
    for name in want:
        case = cases.FULL_CASES[name]()
        d = tempfile.mkdtemp(prefix="goldenfull_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            fa_path = os.path.join(d, "in.fa")
            case["fasta"].tofile(fa_path)
            fa_sha = hashlib.sha256(case["fasta"]).hexdigest()
            r = po.ref_simulate(fa_path, tu=case["tu"], tb=case["tb"], tc=case["tc"], td=case["td"], outdir=d,
                                read_back=False, **cases.sim_kwargs(case))
            assert r.rc == 0, (name, r.rc)
            streams, recs = [], []
            for nm in po.OUT_NAMES:
                dg, n = file_digest(os.path.join(d, f"ref_reads_{nm}.fastq"))
                streams.append(dg)
                recs.append(n)
            out[name] = {"fasta_sha256": fa_sha, "templates_sha256": templates_digest(case), "streams": streams,
                         "wrote": [recs[0], recs[1], recs[3], recs[5]], "cg_exit": list(r.cg),
                         "reference_seconds": round(r.seconds, 2)}
            print(name, [s["bytes"] for s in streams], out[name]["wrote"], r.cg, f"{r.seconds:.1f} s", flush=True)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        with open(OUT, "w") as fh:
            json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
