"""Full-size parity properties (BASELINE.json config 3: 3.1 Gbp reference, 10 M unpaired 250 bp reads).

The oracle needs minutes at this size, so the CUDA path is checked through properties that do not depend
on size: determinism, shard-count invariance (native) and shard-chain invariance (replay) of the whole
5.8 GB FASTQ stream by running CRC, the record structure of every record, and -- on a sample of records
spread over the whole stream -- the read name grammar, the reference bases behind every unedited column and
the template's quality string.  Everything goes through the C ABI, like the small bit-exact cases.
"""
import sys
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (the workload builder of the benchmark: same inputs as the measured run)

pytestmark = pytest.mark.gpu

CHUNKSZ = 128 * 1024
_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


@pytest.fixture(scope="module")
def full():
    """Engine with the full-size genome loaded and the templates set; built once for the module."""
    from qtip_b200 import sim, synth
    wl = bench.workload(1, 1.0)
    fa = synth.make_fasta(wl["records"], seed=wl["genome_seed"], lead_n=wl["lead_n"])
    eng = sim.Engine()
    eng.load_fasta_bytes(fa)
    eng.set_unpaired(0, wl["tu"], len(wl["tu"]))
    eng.set_unpaired(1, [], 0)
    eng.set_paired(2, [], 0)
    eng.set_paired(3, [], 0)
    yield eng, wl, fa
    eng.close()


def _run(eng, wl, rng_mode, seed=77, shard=(0, 1), cg=None):
    from qtip_b200 import sim
    kw = wl["kw"]
    p = sim.make_params(CHUNKSZ, kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"],
                        rng_mode, cg=cg or bench.cli_seed(seed), philox_seed=seed, shard_rank=shard[0], shard_count=shard[1])
    st = eng.run(p)
    return st, p, eng.result(0, int(st.bytes[0]))


def test_native_full_size_deterministic_and_shard_invariant(full):
    eng, wl, _ = full
    st, _, out = _run(eng, wl, 1)
    assert st.wrote[0] > 9_000_000 and int(st.bytes[0]) == out.size
    crc = zlib.crc32(out)
    del out
    st2, _, out2 = _run(eng, wl, 1)
    assert list(st2.wrote) == list(st.wrote) and zlib.crc32(out2) == crc
    del out2
    # three shards, concatenated in rank order (counter-based draws are keyed by the global read index)
    run, nreads = 0, 0
    for r in range(3):
        s, _, part = _run(eng, wl, 1, shard=(r, 3))
        run = zlib.crc32(part, run)
        nreads += s.wrote[0]
        del part
    assert nreads == st.wrote[0] and run == crc


def test_replay_full_size_shard_chain_is_invariant(full):
    eng, wl, _ = full
    st, p, out = _run(eng, wl, 0)
    crc = zlib.crc32(out)
    del out
    run, cg, draws = 0, None, 0
    for r in range(2):  # the chain is sequential: rank r starts from rank r-1's exit state
        s, pp, part = _run(eng, wl, 0, shard=(r, 2), cg=cg)
        cg = (pp.cg1, pp.cg2)
        run = zlib.crc32(part, run)
        draws += s.draws
        del part
    assert run == crc and cg == (p.cg1, p.cg2) and draws == st.draws


def test_full_size_record_structure_and_sampled_content(full):
    eng, wl, fa = full
    st, _, out = _run(eng, wl, 1, seed=4242)
    n = int(st.wrote[0])
    # --- every record: four lines, '+' line, sequence and quality of the template length
    nl = np.flatnonzero(out == 10)
    assert nl.size == 4 * n
    starts = np.empty(n, np.int64)
    starts[0] = 0
    starts[1:] = nl[3:-1:4] + 1
    assert (out[starts] == ord("@")).all()
    l1, l2, l3, l4 = nl[0::4], nl[1::4], nl[2::4], nl[3::4]
    assert ((l2 - l1 - 1) == 250).all() and ((l4 - l3 - 1) == 250).all() and ((l3 - l2) == 2).all()
    assert (out[l2 + 1] == ord("+")).all()
    # --- the genome, for the sampled checks: record name -> (first base offset in the text, line width 70)
    text = fa
    hdr = {}
    pos = 0
    for name, length in wl["records"]:
        assert text[pos] == ord(">")
        eol = pos + 1 + len(name)
        assert bytes(text[pos + 1:eol]) == name.encode() and text[eol] == 10
        hdr[name.encode()] = (eol + 1, length)
        pos = eol + 1 + length + (length + 69) // 70

    def ref_bases(name, off, ln):
        base, length = hdr[name]
        assert off + ln <= length
        first = base + off + off // 70
        raw = bytes(text[first:first + ln + ln // 70 + 2]).replace(b"\n", b"")
        return raw[:ln].upper()

    quals = {}
    for t in wl["tu"]:
        quals.setdefault((t.fw_flag, t.best_score), set()).add(t.qual if isinstance(t.qual, bytes) else t.qual.encode())
    rng = np.random.default_rng(5)
    checked = 0
    for i in rng.integers(0, n, 3000):
        a, b, c, d = l1[i], l2[i], l3[i], l4[i]
        name = bytes(out[starts[i]:a])
        seq = bytes(out[a + 1:b])
        qual = bytes(out[c + 1:d])
        f = name.split(b":")
        assert f[0] == b"@qtip!" and f[2] in (b"+", b"-") and f[-1] == b"u", name
        fw = f[2] == b"+"
        ref = ref_bases(f[1], int(f[3]), 250)
        read = seq if fw else seq.translate(_COMP)[::-1]
        mism = sum(1 for x, y in zip(read, ref) if x != y)
        assert mism <= 5, (name, mism)  # templates carry 0-5 X and nothing else
        key = ("T" if fw else "F", int(f[4]))
        want = qual if fw else qual[::-1]
        assert want in quals.get(key, ()), name  # the quality string is one of the templates' with this strand and score
        checked += 1
    assert checked == 3000
