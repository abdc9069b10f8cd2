"""Generates tests/golden/golden.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference to build oracle/_ref, see oracle/Makefile):

    python tests/golden/make_golden.py

For every case in tests/cases.py it runs the reference's StreamingSimulator
(oracle/_ref/libref.so: the reference's own simplesim.cpp/fasta.cpp/rnglib.cpp/ranlib.cpp behind
oracle/ref_wrap.cpp) and records sha256 + size of each of the seven FASTQ streams, the
generator state at exit and sha256 of the inputs (to detect drift in the synthetic generators).
The smallest case's streams are also stored verbatim (tests/golden/<case>_<stream>.fastq).
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

VERBATIM = ("inversion",)


def templates_digest(case) -> str:
    h = hashlib.sha256()
    for k in ("tu", "tb", "tc", "td"):
        for t in case[k]:
            h.update(repr(t).encode())
    return h.hexdigest()


def main():
    po.build(ref=True)
    out = {}
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    for name, fn in cases.CASES.items():
        case = fn()
        fa_path = os.path.join(tmp, f"golden_{name}.fa")
        case["fasta"].tofile(fa_path)
        r = po.ref_simulate(fa_path, tu=case["tu"], tb=case["tb"], tc=case["tc"], td=case["td"],
                            **cases.sim_kwargs(case))
        assert r.rc == 0, (name, r.rc)
        out[name] = {
            "fasta_sha256": hashlib.sha256(case["fasta"].tobytes()).hexdigest(),
            "templates_sha256": templates_digest(case),
            "streams": [{"bytes": len(s), "sha256": hashlib.sha256(s).hexdigest()} for s in r.streams],
            "cg_exit": list(r.cg),
        }
        if name in VERBATIM:
            for s, nm in zip(r.streams, po.OUT_NAMES):
                with open(os.path.join(HERE, f"{name}_{nm}.fastq"), "wb") as fh:
                    fh.write(s)
        os.unlink(fa_path)
        print(name, [len(s) for s in r.streams], r.cg)
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
