"""Drop-in test: the reference's UNCHANGED qtip_parse.cpp, compiled against src/ (our header-
compatible shim) and linked to libqtipsim, must write byte-identical files to the reference's own
qtip-parse for the same command line -- the process-level boundary the Python `qtip` driver uses
(qtip:474-478; argv protocol qtip_parse.cpp:1594-1712)."""
import filecmp
import os
import subprocess

import pytest

import util
from oracle import pyoracle as po
from qtip_b200 import synth

pytestmark = pytest.mark.gpu

OURS = os.path.join(util.ROOT, "build", "qtip-parse")


def _inputs(workdir):
    fa = synth.make_fasta([("chr1 first", 260000), ("chr2", 140000)], seed=91, lead_n=1500)
    fa_path = os.path.join(workdir, "dropin.fa")
    fa.tofile(fa_path)
    tu = synth.templates_unpaired(400, 100, 92, max_x=4) + synth.templates_unpaired(100, 120, 93, indel=True)
    tb = synth.templates_badend(150, 100, 94, opp_len=60)
    tc = synth.templates_paired(300, 100, 95, frag_mean=320, frag_sd=30)
    td = synth.templates_paired(120, 100, 96, discordant=True, max_fraglen=4000)
    sam = os.path.join(workdir, "dropin.sam")
    synth.write_sam(sam, unpaired=tu, badend=tb, pairs=[(t, True) for t in tc] + [(t, False) for t in td], rname="chr1")
    return fa_path, sam


@pytest.mark.skipif(not (po.have_ref() and os.path.exists(OURS)), reason="needs oracle/_ref and build/qtip-parse")
@pytest.mark.parametrize("mode,opts", [
    ("s", {"input-model-size": 30000, "sim-function": "const", "sim-factor": 2500, "sim-unp-min": 1, "sim-conc-min": 1,
           "sim-disc-min": 1, "sim-bad-end-min": 1}),
    # reservoir smaller than the input: the sampler consumes draws before the simulator starts;
    # Python-side defaults (qtip:854-898): sqrt / factor 45 / minimums
    ("ifs", {"input-model-size": 200, "sim-function": "sqrt", "sim-factor": 45, "sim-unp-min": 3000, "sim-conc-min": 3000,
             "sim-disc-min": 1000, "sim-bad-end-min": 1000, "max-allowed-fraglen": 100000, "wiggle": 30}),
])
def test_qtip_parse_dropin_is_byte_identical(mode, opts, workdir):
    fa, sam = _inputs(workdir)
    outs = {}
    for tag, exe in (("ref", po.REF_PARSE), ("ours", OURS)):
        d = os.path.join(workdir, f"dropin_{mode}_{tag}")
        os.makedirs(d, exist_ok=True)
        paths, cp = po.ref_qtip_parse(sam, fa, d, seed=424242, opts=opts, exe=exe, mode=mode)
        assert cp.returncode == 0, cp.stderr[-2000:]
        outs[tag] = (d, paths, cp.stderr)
    for a, b in zip(outs["ref"][1], outs["ours"][1]):
        assert os.path.exists(a) and os.path.exists(b)  # all seven files always exist (qtip:488-508)
        assert filecmp.cmp(a, b, shallow=False), os.path.basename(a)
    assert sum(os.path.getsize(p) for p in outs["ref"][1]) > 100000
    # the four summary lines (simplesim.cpp:399-406) are part of the interface too
    ref_wrote = [l for l in outs["ref"][2].splitlines() if "Wrote" in l]
    our_wrote = [l for l in outs["ours"][2].splitlines() if "Wrote" in l]
    assert ref_wrote == our_wrote and len(ref_wrote) == 4
    if "f" in mode:  # feature records come from the unchanged caller; they must not be disturbed
        for suffix in ("rec_rec_u.npy", "rec_rec_b.npy", "rec_rec_c.npy", "rec_rec_d.npy", "rec_rec_u.meta"):
            assert filecmp.cmp(os.path.join(outs["ref"][0], suffix), os.path.join(outs["ours"][0], suffix), shallow=False)


@pytest.mark.skipif(not os.path.exists(OURS), reason="needs build/qtip-parse")
def test_qtip_parse_passthrough_protocol():
    """No-argument call prints the option names the Python driver forwards (qtip_parse.cpp:1552-1568)."""
    out = subprocess.run([OURS], capture_output=True, text=True).stdout.split()
    assert out == ["wiggle", "input-model-size", "fraction-even", "low-score-bias", "max-allowed-fraglen", "sim-factor",
                   "sim-function", "sim-unp-min", "sim-conc-min", "sim-disc-min", "sim-bad-end-min", "seed"]
