"""Synthetic inputs of the shapes BASELINE.json names: uniform-ACGT
genomes as (multi-)FASTA text and input-model templates, all from fixed seeds.

Also writes the templates as a SAM file with ``ZT:Z`` tags and ``=/X`` CIGARs, the input
the reference's ``qtip-parse`` CLI needs.
"""
from __future__ import annotations

import random
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .templates import TemplatePaired, TemplateUnpaired

_LUT4 = None


def _lut4() -> np.ndarray:
    global _LUT4
    if _LUT4 is None:
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        idx = np.arange(256, dtype=np.uint32)
        lut = np.empty((256, 4), dtype=np.uint8)
        for k in range(4):
            lut[:, k] = acgt[(idx >> (2 * k)) & 3]
        _LUT4 = lut
    return _LUT4


def random_bases(n: int, rng: np.random.Generator) -> np.ndarray:
    """n i.i.d. uniform ACGT bytes (4 bases per random byte)."""
    nb = (n + 3) // 4
    raw = rng.integers(0, 256, size=nb, dtype=np.uint8)
    return _lut4()[raw].reshape(-1)[:n]


def fasta_record(name: str, seq: np.ndarray, width: int = 70) -> np.ndarray:
    """One FASTA record as a uint8 array: '>name\\n' + lines of `width` bases."""
    n = seq.size
    nfull, tail = divmod(n, width)
    hdr = np.frombuffer((">" + name + "\n").encode(), dtype=np.uint8)
    body = np.empty(nfull * (width + 1) + (tail + 1 if tail else 0), dtype=np.uint8)
    if nfull:
        blk = body[: nfull * (width + 1)].reshape(nfull, width + 1)
        blk[:, :width] = seq[: nfull * width].reshape(nfull, width)
        blk[:, width] = 10
    if tail:
        body[nfull * (width + 1): -1] = seq[nfull * width:]
        body[-1] = 10
    return np.concatenate([hdr, body])


def make_fasta(records: Sequence[Tuple[str, int]], seed: int, lead_n: int = 0, width: int = 70,
               lowercase_frac: float = 0.0) -> np.ndarray:
    """FASTA text (uint8 array) with the given (name, length) records.  `lead_n`: each
    record starts with that many N (10 kbp leading N run)."""
    rng = np.random.default_rng(seed)
    parts = []
    for name, length in records:
        seq = random_bases(length, rng)
        if lead_n:
            seq[: min(lead_n, length)] = ord("N")
        if lowercase_frac > 0:
            m = rng.random(length) < lowercase_frac
            seq[m] |= 0x20
        parts.append(fasta_record(name, seq, width))
    return np.concatenate(parts) if parts else np.empty(0, dtype=np.uint8)


# GRCh38 primary chromosome lengths (1..22, X, Y), for the "GRCh38-sized" config C3/C4
GRCH38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                  138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                  83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


def grch38_like_records(scale: float = 1.0) -> List[Tuple[str, int]]:
    names = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
    return [(nm, max(1000, int(l * scale))) for nm, l in zip(names, GRCH38_LENGTHS)]


def _qual(rnd: random.Random, L: int) -> bytes:
    return bytes(rnd.randint(35, 73) for _ in range(L))  # qual U[35,73]


def _xscript_subst(rnd: random.Random, L: int, max_x: int) -> bytes:
    xs = bytearray(b"=" * L)
    for _ in range(rnd.randint(0, max_x)):
        xs[rnd.randrange(L)] = ord("X")
    return bytes(xs)


def _xscript_indel(rnd: random.Random, L: int, edit_frac: float = 0.10, soft_clip: bool = True) -> bytes:
    """Read of length L with ~edit_frac edited columns: X, and I/D runs of 1-6
    (SURVEY 8d C5), optionally soft-clipped ends."""
    ops = bytearray()
    rd = 0
    lead = rnd.randint(0, 5) if soft_clip and rnd.random() < 0.3 else 0
    trail = rnd.randint(0, 5) if soft_clip and rnd.random() < 0.3 else 0
    ops += b"S" * lead
    rd += lead
    body_end = L - trail
    while rd < body_end:
        r = rnd.random()
        if r < edit_frac * 0.5:
            ops += b"X"
            rd += 1
        elif r < edit_frac * 0.75:
            run = min(rnd.randint(1, 6), body_end - rd)
            ops += b"I" * run
            rd += run
        elif r < edit_frac and ops and ops[-1:] != b"S":
            ops += b"D" * rnd.randint(1, 6)
        else:
            ops += b"="
            rd += 1
    # a transcript must not end in a deletion next to the clip; harmless either way
    ops += b"S" * trail
    return bytes(ops)


def templates_unpaired(n: int, L: int, seed: int, max_x: int = 3, indel: bool = False, mate_flag: str = "0",
                       opp_len: int = 0, len_jitter: int = 0) -> List[TemplateUnpaired]:
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        l = L - (rnd.randint(0, len_jitter) if len_jitter else 0)
        xs = _xscript_indel(rnd, l) if indel else _xscript_subst(rnd, l, max_x)
        out.append(TemplateUnpaired(best_score=-rnd.randint(0, 60), len=l, fw_flag="T" if rnd.random() < 0.5 else "F",
                                    mate_flag=mate_flag, opp_len=opp_len, qual=_qual(rnd, l), edit_xscript=xs))
    return out


def templates_badend(n: int, L: int, seed: int, opp_len: int = 150, max_x: int = 4) -> List[TemplateUnpaired]:
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        out.append(TemplateUnpaired(best_score=-rnd.randint(0, 60), len=L, fw_flag="T" if rnd.random() < 0.5 else "F",
                                    mate_flag="1" if rnd.random() < 0.5 else "2", opp_len=opp_len + rnd.randint(-3, 3),
                                    qual=_qual(rnd, L), edit_xscript=_xscript_subst(rnd, L, max_x)))
    return out


def _rflen(xs: bytes) -> int:
    return sum(xs.count(c) for c in (b"S", b"=", b"X", b"D"))


def templates_paired(n: int, L: int, seed: int, frag_mean: float = 400.0, frag_sd: float = 40.0, max_x: int = 4,
                     discordant: bool = False, max_fraglen: int = 50000, indel: bool = False) -> List[TemplatePaired]:
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        xs1 = _xscript_indel(rnd, L) if indel else _xscript_subst(rnd, L, max_x)
        xs2 = _xscript_indel(rnd, L) if indel else _xscript_subst(rnd, L, max_x)
        if discordant:
            # anything from overlapping mates to far-apart ones, capped like max-allowed-fraglen
            frag = min(max_fraglen, int(rnd.choice([rnd.uniform(L, 2 * L), rnd.uniform(1000, max_fraglen)])))
            fw1 = "T" if rnd.random() < 0.5 else "F"
            fw2 = "T" if rnd.random() < 0.5 else "F"
        else:
            frag = max(L, int(rnd.gauss(frag_mean, frag_sd)))
            fw1 = "T" if rnd.random() < 0.5 else "F"
            fw2 = "F" if fw1 == "T" else "T"
        s1, s2 = -rnd.randint(0, 60), -rnd.randint(0, 60)
        out.append(TemplatePaired(score_12=s1 + s2, score_1=s1, len_1=L, fw_flag_1=fw1, qual_1=_qual(rnd, L),
                                  edit_xscript_1=xs1, score_2=s2, len_2=L, fw_flag_2=fw2, qual_2=_qual(rnd, L),
                                  edit_xscript_2=xs2, upstream1=rnd.random() < 0.5, fraglen=frag))
    return out


# --------------------------------------------------------------------------- SAM

def _cigar(xs: bytes) -> str:
    out = []
    i = 0
    while i < len(xs):
        j = i
        while j < len(xs) and xs[j] == xs[i]:
            j += 1
        out.append(f"{j - i}{chr(xs[i])}")
        i = j
    return "".join(out)


def sam_lines(unpaired: Iterable[TemplateUnpaired] = (), badend: Iterable[TemplateUnpaired] = (),
              pairs: Iterable[Tuple[TemplatePaired, bool]] = (), rname: str = "ref") -> List[str]:
    """SAM records the reference's parser turns back into (approximately) these templates
    (qtip_parse.cpp:814-861, :932-1166).  `pairs` items are
    (template, concordant?)."""
    lines = ["@HD\tVN:1.0\tSO:unsorted"]
    k = 0
    for t in unpaired:
        flag = 0 if t.fw_flag == "T" else 16
        lines.append(f"r{k}\t{flag}\t{rname}\t1000\t42\t{_cigar(t.edit_xscript)}\t*\t0\t0\t{'A' * t.len}\t"
                     f"{t.qual.decode()}\tZT:Z:{t.best_score},1,2")
        k += 1
    for t in badend:
        al_flag = 1 | 8 | (0 if t.fw_flag == "T" else 16) | (64 if t.mate_flag == "1" else 128)
        un_flag = 1 | 4 | (128 if t.mate_flag == "1" else 64)
        al = (f"r{k}\t{al_flag}\t{rname}\t1000\t42\t{_cigar(t.edit_xscript)}\t=\t1000\t0\t{'A' * t.len}\t"
              f"{t.qual.decode()}\tZT:Z:{t.best_score},1,2")
        un = f"r{k}\t{un_flag}\t{rname}\t1000\t0\t*\t=\t1000\t0\t{'C' * t.opp_len}\t{'I' * t.opp_len}\tYT:Z:UP"
        lines += [al, un] if t.mate_flag == "1" else [un, al]
        k += 1
    for t, conc in pairs:
        # place the mates so the reference derives upstream1 and (roughly) fraglen
        up_rf = _rflen(t.edit_xscript_1 if t.upstream1 else t.edit_xscript_2)
        dn_rf = _rflen(t.edit_xscript_2 if t.upstream1 else t.edit_xscript_1)
        up_pos = 1000
        dn_pos = up_pos + max(1, t.fraglen - dn_rf)
        p1, p2 = (up_pos, dn_pos) if t.upstream1 else (dn_pos, up_pos)
        base = 1 | (2 if conc else 0)
        f1 = base | 64 | (0 if t.fw_flag_1 == "T" else 16) | (0 if t.fw_flag_2 == "T" else 32)
        f2 = base | 128 | (0 if t.fw_flag_2 == "T" else 16) | (0 if t.fw_flag_1 == "T" else 32)
        lines.append(f"r{k}\t{f1}\t{rname}\t{p1}\t42\t{_cigar(t.edit_xscript_1)}\t=\t{p2}\t0\t{'A' * t.len_1}\t"
                     f"{t.qual_1.decode()}\tZT:Z:{t.score_1},1,2")
        lines.append(f"r{k}\t{f2}\t{rname}\t{p2}\t42\t{_cigar(t.edit_xscript_2)}\t=\t{p1}\t0\t{'A' * t.len_2}\t"
                     f"{t.qual_2.decode()}\tZT:Z:{t.score_2},1,2")
        k += 1
    return lines


def write_sam(path: str, **kw) -> None:
    with open(path, "w") as fh:
        fh.write("\n".join(sam_lines(**kw)) + "\n")
