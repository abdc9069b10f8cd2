#!/usr/bin/env python
"""bench.py -- tandem-read simulator throughput on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

One "step" = one pass of the hot path over the workload: (pack the FASTA), resolve the
reference's draw chain, simulate + emit FASTQ for every read.

Primary workload, at every N: BASELINE.json configs[2] (the one quoted for 1 B200): GRCh38-sized synthetic
3.1 Gbp reference, 10 M unpaired 250 bp reads PER GPU (weak scaling: windows sharded by rank), native
(Philox) RNG.  The same JSON line carries two more first-class blocks, each with its own value,
step-level and kernel-level roofline and e2e:
  "replay" : the same workload in REPLAY mode (bit-exact with the reference), same number of steps;
  "c4"     : BASELINE.json configs[3] -- 3.1 Gbp, 12.5 M pairs per GPU (100 M on 8), 70/15/15
             concordant / discordant / bad-end 2x150 -- native mode, plus its own "replay" block.

Prints ONE JSON line on rank 0.  `value` = reads (or pairs) per second with the inputs
resident in HBM; `e2e` = the same through the C ABI with host buffers (FASTA text and
templates in host memory -> FASTQ bytes in pinned host memory).  `--impl reference` times the
UNMODIFIED reference (oracle/_ref/libref.so = the reference's own simplesim.cpp, fasta.cpp,
rnglib.cpp, ranlib.cpp) on the host CPU on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from qtip_b200 import synth  # noqa: E402

CHUNKSZ = 128 * 1024
METRIC = "tandem reads/sec"
UNIT = "reads/s"


# ------------------------------------------------------------------ workload --

def workload(n_gpus: int, scale: float, which: str = "auto", cats: str = "bcd"):
    """Returns dict(name, records, templates, targets) for the bench configuration."""
    recs = synth.grch38_like_records(scale)
    if which in ("c3", "auto"):
        # weak scaling: 10M reads per GPU over the same genome, windows sharded across the ranks
        tu = synth.templates_unpaired(30000, 250, seed=3003, max_x=5)
        per = f", x{n_gpus} GPUs (windows sharded by rank)" if n_gpus > 1 else ""
        return dict(name=f"C3: GRCh38-sized synthetic {sum(l for _, l in recs) / 1e9:.2f} Gbp ref (24 records, 10 kbp "
                         f"leading N each), {10 * scale:.3g}M unpaired 250 bp reads per GPU from 30k templates (0-5 X){per}",
                    records=recs, lead_n=10000, tu=tu, tb=[], tc=[], td=[],
                    kw=dict(fraction=float(int(10_000_000 * scale) * n_gpus), function=3, min_u=1, min_c=1, min_d=1, min_b=1),
                    unit="reads/s", genome_seed=3001)
    tc = synth.templates_paired(30000, 150, seed=4002, frag_mean=400, frag_sd=40, max_x=4)
    td = synth.templates_paired(30000, 150, seed=4003, discordant=True, max_fraglen=50000)
    tb = synth.templates_badend(30000, 150, seed=4004, opp_len=150)
    total = int(12_500_000 * n_gpus * scale)
    # 70 / 15 / 15 mix through `added` and the linear function: target = added
    added = (0, int(total * 0.15), int(total * 0.70), int(total * 0.15))
    if set(cats) != set("bcd"):  # diagnostic runs: one category at a time
        tb, tc, td = (tb if "b" in cats else []), (tc if "c" in cats else []), (td if "d" in cats else [])
        added = (0, added[1] if "b" in cats else 0, added[2] if "c" in cats else 0, added[3] if "d" in cats else 0)
    return dict(name=f"C4: GRCh38-sized synthetic {sum(l for _, l in recs) / 1e9:.2f} Gbp ref, {total / 1e6:.3g}M pairs "
                     f"70/15/15 concordant/discordant/bad-end 2x150, sharded across {n_gpus} GPU(s)",
                records=recs, lead_n=10000, tu=[], tb=tb, tc=tc, td=td, added=added,
                kw=dict(fraction=1.0, function=1, min_u=1, min_c=1, min_d=1, min_b=1), unit="pairs/s", genome_seed=3001)


def cli_seed(seed: int):
    return (seed, (seed * 77) % 2147483562 + 1)


# -------------------------------------------------------------------- clocks --

class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------- reference --

def reference_sample(wl, scale_note=""):
    """A bounded sample of the workload for the CPU arm: the last records of the same genome
    (about 3% of the bases) with the read targets scaled by the same fraction, so reads per base --
    and therefore per-read and per-base costs -- are those of the full workload."""
    recs = wl["records"]
    tot = sum(l for _, l in recs)
    pick, acc = [], 0
    for nm, l in reversed(recs):
        pick.append((nm, l))
        acc += l
        if acc >= 0.03 * tot:
            break
    pick.reverse()
    frac = acc / tot
    return pick, frac


def run_reference_steps(wl, steps, warmup, threads_note="1 (the reference is single-threaded, qtip_parse.cpp:73)"):
    from oracle import pyoracle as po
    if not po.have_ref():
        return None, "oracle/_ref not built"
    pick, frac = reference_sample(wl)
    fa = synth.make_fasta(pick, seed=wl["genome_seed"], lead_n=wl["lead_n"])
    d = tempfile.mkdtemp(prefix="qtipref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa_path = os.path.join(d, "sample.fa")
    fa.tofile(fa_path)
    kw = dict(wl["kw"])
    added = wl.get("added")
    if kw["function"] == 3:
        kw["fraction"] = float(max(1, int(kw["fraction"] * frac)))
    else:
        added = tuple(int(a * frac) for a in added)
    times, nreads, nbytes = [], 0, 0
    for i in range(warmup + steps):
        r = po.ref_simulate(fa_path, tu=wl["tu"], tb=wl["tb"], tc=wl["tc"], td=wl["td"], added=added, chunksz=CHUNKSZ,
                            cg=cli_seed(1000 + i), outdir=d, read_back=False, **kw)
        dt = r.seconds
        assert r.rc == 0
        if i >= warmup:
            times.append(dt)
            nbytes = sum(r.streams)
    # count reads from the first stream(s): 4 lines per record
    def count(path):
        n = 0
        with open(path, "rb") as fh:
            while True:
                b = fh.read(1 << 24)
                if not b:
                    break
                n += b.count(b"\n")
        return n // 4
    names = ("u", "b_1", "c_1", "d_1")
    nreads = sum(count(os.path.join(d, f"ref_reads_{n}.fastq")) for n in names)
    import shutil
    shutil.rmtree(d, ignore_errors=True)
    sec = float(np.mean(times))
    return dict(reads_per_s=nreads / sec, gb_per_s=nbytes / sec / 1e9, sec_per_step=sec, reads=nreads, bytes=nbytes,
                sample=f"{len(pick)} of {len(wl['records'])} records ({frac * 100:.1f}% of the bases, same read density), "
                       f"{nreads} reads/pairs, outputs on tmpfs, FASTA scan included",
                cores=1), None


def bind_to_gpu_numa_node(local_rank: int):
    """Pins this rank to the CPUs of its GPU's NUMA node before any pinned host memory is allocated, so that the rank's
    FASTQ result buffers (GBs per step at N = 8) live in the memory next to its GPU's PCIe root.  No-op on single-node
    hosts or where sysfs does not say."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node_path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        if not os.path.exists(node_path):
            return
        node = int(open(node_path).read().strip())
        if node < 0:
            return
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & ids
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


# ---------------------------------------------------------------------- main --

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("QTIP_BENCH_SCALE", "1.0")),
                    help="shrink genome and read counts together (1.0 = the BASELINE.json configuration)")
    ap.add_argument("--mode", default="native", choices=["replay", "native"],
                    help="RNG mode of the primary numbers; the other mode is measured too (1 step) and reported alongside")
    ap.add_argument("--workload", default="auto", choices=["auto", "c3", "c4"],
                    help="primary workload.  auto / c3: BASELINE config 3 (10M unpaired 250 bp reads per GPU, weak scaling), with "
                         "config 4 (12.5M pairs per GPU, 70/15/15 c/d/b) as the \"c4\" block; c4: config 4 as the primary (diagnostic)")
    ap.add_argument("--cats", default="bcd", help="diagnostic: pair categories kept in the C4 mix (subset of 'bcd')")
    ap.add_argument("--no-secondary", action="store_true", help="skip the replay-mode blocks")
    ap.add_argument("--no-c4", action="store_true", help="skip the config-4 block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    wl = workload(n_gpus, args.scale, args.workload, args.cats)

    if args.impl == "reference":
        if rank != 0:
            return 0
        ref_warm = max(1, min(args.warmup, 1))  # one warm-up: a CPU timing does not need more, and each costs ~20 s
        res, why = run_reference_steps(wl, args.steps, ref_warm)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": why}))
            return 0
        line = {
            "impl": "reference", "metric": METRIC, "value": res["reads_per_s"], "unit": wl["unit"], "n_gpus": n_gpus,
            "steps": args.steps, "warmup": ref_warm, "ms_per_step": res["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": res["sample"]},
            "output_gb_per_s": res["gb_per_s"],
            "cpu_baseline": {"value": res["reads_per_s"], "unit": wl["unit"], "cores": res["cores"], "kind": "reference",
                             "sample": res["sample"]},
            "e2e": {"value": res["reads_per_s"], "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from qtip_b200 import _abi, sim

    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rng_mode = _abi.RNG_REPLAY if args.mode == "replay" else _abi.RNG_PHILOX

    # ---- inputs: FASTA text in pinned host memory (rank 0 builds it) ----
    eng = sim.Engine(local_rank)
    fa_pinned = None
    t_setup = time.perf_counter()
    if rank == 0:
        fa = synth.make_fasta(wl["records"], seed=wl["genome_seed"], lead_n=wl["lead_n"])
        fa_pinned = torch.empty(fa.size, dtype=torch.uint8).pin_memory()
        fa_pinned.numpy()[:] = fa
        del fa
        eng.load_fasta_bytes(fa_pinned.numpy())
    repl_state = {}

    def replicate_genome():
        """Rank 0 has packed the genome on its GPU: NCCL broadcast of the packed arrays over NVLink, the others adopt."""
        import ctypes as C
        if "meta" not in repl_state:  # sizes and record table: the same for every step
            meta = [None]
            if rank == 0:
                ptrs = []
                for which in (0, 1, 2):
                    p, n, isd = C.c_void_p(), C.c_uint64(), C.c_int()
                    eng.check(eng.lib.qtipsim_genome_export(eng.h, which, C.byref(p), C.byref(n), C.byref(isd)))
                    ptrs.append((p.value, n.value))
                blob = (C.c_char * ptrs[2][1]).from_address(ptrs[2][0]).raw
                meta = [(ptrs[0][1], ptrs[1][1], blob)]
            dist.broadcast_object_list(meta, src=0)
            repl_state["meta"] = meta[0]
        nb_bytes, nm_bytes, blob = repl_state["meta"]

        class _Raw:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}

        if rank == 0:
            ptrs = []
            for which in (0, 1):
                p, n, isd = C.c_void_p(), C.c_uint64(), C.c_int()
                eng.check(eng.lib.qtipsim_genome_export(eng.h, which, C.byref(p), C.byref(n), C.byref(isd)))
                ptrs.append(p.value)
            tb_ = torch.as_tensor(_Raw(ptrs[0], nb_bytes), device="cuda")
            tn_ = torch.as_tensor(_Raw(ptrs[1], nm_bytes), device="cuda")
        else:
            if "bufs" not in repl_state:
                repl_state["bufs"] = (torch.empty(nb_bytes, dtype=torch.uint8, device="cuda"),
                                      torch.empty(nm_bytes, dtype=torch.uint8, device="cuda"))
            tb_, tn_ = repl_state["bufs"]
        dist.broadcast(tb_, src=0)
        dist.broadcast(tn_, src=0)
        torch.cuda.synchronize()
        if rank != 0:
            eng.check(eng.lib.qtipsim_genome_adopt(eng.h, blob, len(blob), C.c_void_p(tb_.data_ptr()), nb_bytes,
                                                   C.c_void_p(tn_.data_ptr()), nm_bytes))

    if world > 1:
        replicate_genome()
    cur = {"wl": wl}  # the workload being measured (the config-4 block swaps the templates in, same genome)

    def set_templates():
        w = cur["wl"]
        added = w.get("added") or (len(w["tu"]), len(w["tb"]), len(w["tc"]), len(w["td"]))
        eng.set_unpaired(0, w["tu"], added[0])
        eng.set_unpaired(1, w["tb"], added[1])
        eng.set_paired(2, w["tc"], added[2])
        eng.set_paired(3, w["td"], added[3])

    set_templates()
    setup_s = time.perf_counter() - t_setup

    state_buf = torch.zeros(2, dtype=torch.int64, device="cuda")
    host_out = [None] * _abi.NOUT

    trace = os.environ.get("QTIP_BENCH_TRACE") == "1"
    stream_ready = [False]  # pinned result buffers sized by a first (one-shot) end-to-end step

    def one_step(seed, rng_mode, e2e=False):
        """Runs one step on this rank's shard; returns stats."""
        kw = cur["wl"]["kw"]
        tr0 = time.perf_counter()
        if e2e and world == 1 and stream_ready[0] and os.environ.get("QTIP_BENCH_STREAM", "1") == "1":
            # the streaming path (default at N = 1; QTIP_BENCH_STREAM=0 selects the one-shot calls): chunked upload,
            # record-by-record pack + simulate, download of earlier records' reads while later text is still uploading.
            # Same bytes as the one-shot path (tests/test_gpu_parity.py::test_streaming_path_equals_one_shot_path).
            eng.stream_begin(fa_pinned.numpy())
            tr1 = time.perf_counter()
            set_templates()                                    # the host compiles the templates while the text uploads
            if trace:
                print(f"[trace] stream_begin {1e3 * (tr1 - tr0):.1f} ms, set_templates {1e3 * (time.perf_counter() - tr1):.1f} ms", file=sys.stderr)
            p = sim.make_params(CHUNKSZ, kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"],
                                rng_mode, cg=cli_seed(seed), philox_seed=seed, shard_rank=0, shard_count=1)
            st = eng.stream_run(p, [t.data_ptr() if t is not None else 0 for t in host_out],
                                [t.numel() if t is not None else 0 for t in host_out])
            if trace:
                print(f"[trace] streaming e2e step {1e3 * (time.perf_counter() - tr0):.1f} ms", file=sys.stderr)
            return st
        if e2e:
            if rank == 0:
                eng.load_fasta_bytes_async(fa_pinned.numpy())  # H2D of the FASTA text + device pack ...
            tr1 = time.perf_counter()
            set_templates()                                    # ... while the host compiles the templates
            tr1b = time.perf_counter()
            eng.wait_genome()
            if world > 1:
                replicate_genome()                             # N > 1: rank 0's packed genome to every GPU (NCCL)
            if trace:
                print(f"[trace] load_fasta {1e3 * (tr1 - tr0):.1f} ms, set_templates {1e3 * (tr1b - tr1):.1f} ms, genome resident after "
                      f"{1e3 * (time.perf_counter() - tr0):.1f} ms", file=sys.stderr)
        cg = cli_seed(seed)
        chain = world > 1 and rng_mode == _abi.RNG_REPLAY
        if chain and rank > 0:
            dist.recv(state_buf, src=rank - 1)  # the replay chain is sequential across shards
            cg = (int(state_buf[0].item()), int(state_buf[1].item()))
        p = sim.make_params(CHUNKSZ, kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"],
                            rng_mode, cg=cg, philox_seed=seed, shard_rank=rank, shard_count=world)
        tr2 = time.perf_counter()
        st = eng.run(p)
        tr3 = time.perf_counter()
        if chain and rank < world - 1:
            state_buf[0], state_buf[1] = p.cg1, p.cg2
            dist.send(state_buf, dst=rank + 1)
        if e2e:
            import ctypes as C
            for s in range(_abi.NOUT):
                n = int(st.bytes[s])
                if n:
                    if host_out[s] is None or host_out[s].numel() < n:
                        host_out[s] = torch.empty(int(n * 1.05) + (1 << 20), dtype=torch.uint8).pin_memory()
                    eng.check(eng.lib.qtipsim_result_copy(eng.h, s, C.c_void_p(host_out[s].data_ptr()), host_out[s].numel()))
            if trace:
                print(f"[trace] run {1e3 * (tr3 - tr2):.1f} ms (pack {st.ms_pack:.1f}), result_copy {1e3 * (time.perf_counter() - tr3):.1f} ms",
                      file=sys.stderr)
        return st

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(rng_mode, steps, warmup, with_clocks, do_e2e=True):
        """Device-resident arm + end-to-end arm for one RNG mode of the current workload; returns a dict of reduced numbers."""
        wl = cur["wl"]
        # clocks / throttle reasons are sampled (nvidia-smi, every 100 ms) from the warm-up through the
        # end-to-end arm: the device-timed region alone lasts only a few milliseconds per step
        sampler = ClockSampler(local_rank) if with_clocks else None
        if sampler:
            sampler.start()
        for i in range(warmup):
            one_step(100 + i, rng_mode)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        agg = dict(reads=0, bytes=0, alg=0, launches=0, ms_plan=0.0, ms_sim=0.0, ms_total=0.0)
        for i in range(steps):
            st = one_step(200 + i, rng_mode)
            agg["reads"] += sum(st.wrote)
            agg["bytes"] += sum(st.bytes)
            agg["alg"] += int(st.algorithmic_bytes)
            agg["launches"] += int(st.kernel_launches)
            agg["ms_plan"] += st.ms_plan
            agg["ms_sim"] += st.ms_simulate
            agg["ms_total"] += st.ms_total
        ev1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ev_ms = ev0.elapsed_time(ev1)
        barrier()
        # end-to-end arm (host buffers; H2D of the FASTA text and templates, D2H of the FASTQ bytes inside)
        e2e_reads, e2e_d2h, e2e_ms = 0, 0, 0.0
        e2e_steps = max(1, min(steps, 2))
        if do_e2e:
            one_step(300, rng_mode, e2e=True)  # warm (allocates the pinned result buffers)
            stream_ready[0] = True
            if world == 1 and os.environ.get("QTIP_BENCH_STREAM", "1") == "1":
                one_step(301, rng_mode, e2e=True)  # warm (the path the timed steps take)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(e2e_steps):
                st = one_step(400 + i, rng_mode, e2e=True)
                e2e_reads += sum(st.wrote)
                e2e_d2h += sum(st.bytes)
            e1.record()
            torch.cuda.synchronize()
            e2e_ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        barrier()
        t = torch.tensor([ev_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
        w = torch.tensor([agg["reads"], agg["bytes"], agg["alg"], agg["launches"], e2e_reads, e2e_d2h], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
        ev_ms, e2e_ms, wall_ms = [float(x) for x in t.tolist()]
        reads, obytes, alg, launches, e2e_reads, e2e_d2h = [float(x) for x in w.tolist()]
        sim_ms_step = agg["ms_sim"] / steps
        alg_step_rank0 = agg["alg"] / steps
        achieved = alg_step_rank0 / (sim_ms_step * 1e-3) / 1e9 if sim_ms_step > 0 else 0.0
        # FASTA text bytes: uploaded by rank 0 every e2e step (N > 1: then broadcast GPU to GPU as packed arrays)
        h2d = sum(l for _, l in wl["records"]) * 71 // 70 + 64 * len(wl["records"])
        e2e = None
        if do_e2e:
            e2e = {"value": e2e_reads / (e2e_ms * 1e-3), "unit": wl["unit"], "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(e2e_d2h / e2e_steps), "ms_per_step": e2e_ms / e2e_steps,
                   "output_gb_per_s": e2e_d2h / (e2e_ms * 1e-3) / 1e9, "steps": e2e_steps}
        return dict(value=reads / (ev_ms * 1e-3), ms_per_step=ev_ms / steps, wall_ms_per_step=wall_ms / steps,
                    output_gb_per_s=obytes / (ev_ms * 1e-3) / 1e9, reads_per_step=reads / steps,
                    phase={"plan": agg["ms_plan"] / steps, "simulate": sim_ms_step, "total": agg["ms_total"] / steps},
                    achieved=achieved, alg_step=alg_step_rank0, alg_all=alg / steps, launches=int(launches), clocks=clocks,
                    e2e=e2e, steps=steps, warmup=warmup)

    primary_mode = _abi.RNG_PHILOX if args.mode == "native" else _abi.RNG_REPLAY
    other_mode = _abi.RNG_REPLAY if args.mode == "native" else _abi.RNG_PHILOX
    other = "replay" if args.mode == "native" else "native"
    names = {"native": "philox (native counter-based RNG: the product default)",
             "replay": "replay (consumes the reference's rnglib/ranlib stream: bit-exact with the reference)"}
    sec_steps = max(3, args.steps)  # the other RNG mode is measured like the primary one, never on a single sample
    m = measure(primary_mode, args.steps, args.warmup, True)
    m2 = None if args.no_secondary else measure(other_mode, sec_steps, max(1, min(args.warmup, 3)), False)
    # ---- BASELINE configs[3] on the same genome: 12.5 M pairs per GPU, 70/15/15 concordant / discordant / bad-end ----
    m4 = m4r = wl4 = None
    if args.workload in ("auto", "c3") and not args.no_c4:
        wl4 = workload(n_gpus, args.scale, "c4", args.cats)
        cur["wl"] = wl4
        set_templates()
        host_out[:] = [None] * _abi.NOUT  # the primary workload's pinned result buffers go back to the system
        try:
            torch._C._host_emptyCache()
        except Exception:
            pass
        stream_ready[0] = False
        m4 = measure(primary_mode, args.steps, args.warmup, False)
        # the replay chain of this mix is long (~150 skipped draws per bad-end pair: 3 s per 13.5 M pairs) and sequential
        # across ranks, so N ranks would take N times as long for N times the pairs: timed at N = 1 only, on the
        # device-resident arm
        m4r = None if (args.no_secondary or world > 1) else measure(other_mode, 3, 1, False, do_e2e=False)
        cur["wl"] = wl

    # ---- who limits the end-to-end arm at N > 1?  device-to-host copies of 2 GB into pinned memory: rank 0 alone, then every
    # rank at once (the FASTQ bytes of a step go home that way: N x 5.8 GB) ----
    d2h_probe = None
    if world > 1:
        nprobe = 2 << 30
        dsrc = torch.empty(nprobe, dtype=torch.uint8, device="cuda")
        hdst = torch.empty(nprobe, dtype=torch.uint8).pin_memory()

        def d2h_once():
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hdst.copy_(dsrc, non_blocking=True)
            torch.cuda.synchronize()
            return nprobe / (time.perf_counter() - t0) / 1e9

        d2h_once()
        barrier()
        alone = d2h_once() if rank == 0 else 0.0
        barrier()
        together = d2h_once()
        barrier()
        tt = torch.tensor([alone, together, -together], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tsum = torch.tensor([together], dtype=torch.float64, device="cuda")
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        d2h_probe = {"what": "2 GiB device-to-host into pinned memory: rank 0 alone, then all ranks at once", "alone_gb_per_s": float(tt[0]),
                     "concurrent_per_rank_gb_per_s": {"max": float(tt[1]), "min": float(-tt[2])}, "concurrent_aggregate_gb_per_s": float(tsum[0])}
        del dsrc, hdst

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

        def traffic_of(fname):
            # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture (same workload only)
            if args.scale != 1.0:
                return None
            try:
                with open(os.path.join(ROOT, "profiles", fname)) as fh:
                    return json.load(fh)["traffic_bytes_per_launch"]
            except Exception:
                return None

        def block(mm, w, mode, traffic=None):
            """One measured (workload, RNG mode): throughput, where the step's time goes, the dominant kernel against the HBM
            roofline and the whole step against it, and the end-to-end arm."""
            step_gbs = mm["alg_step"] / (mm["ms_per_step"] * 1e-3) / 1e9 if mm["ms_per_step"] > 0 else 0.0
            out = {"rng_mode": names[mode], "workload": w["name"], "value": mm["value"], "unit": w["unit"], "steps": mm["steps"],
                   "warmup": mm["warmup"], "ms_per_step": mm["ms_per_step"], "reads_per_step": mm["reads_per_step"],
                   "output_gb_per_s": mm["output_gb_per_s"], "phase_ms_per_step_rank0": mm["phase"],
                   "roofline": {"bound": "hbm", "kernel": "k_simulate_tile", "achieved": mm["achieved"], "peak": peak, "unit": "GB/s",
                                "frac": mm["achieved"] / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                                "algorithmic_bytes_per_step": mm["alg_step"]},
                   "roofline_step": {"what": "algorithmic bytes of rank 0's step / rank 0's whole step (plan + offsets + simulate)",
                                     "achieved": mm["alg_step"] / (mm["phase"]["total"] * 1e-3) / 1e9 if mm["phase"]["total"] > 0 else None,
                                     "frac": (mm["alg_step"] / (mm["phase"]["total"] * 1e-3) / 1e9 / peak) if mm["phase"]["total"] > 0 and peak else None,
                                     "achieved_all_ranks_wall": step_gbs},
                   "gpu_launches": mm["launches"]}
            if mm["e2e"] is not None:
                out["e2e"] = mm["e2e"]
            return out

        tr_native = traffic_of("r02_k_simulate_tile_traffic.json")
        # config 4: the step's three launches (bad-end, concordant, discordant) together, like its `achieved`
        tr_c4 = traffic_of("r02_k_simulate_tile_pairs_traffic.json") if (args.mode == "native" and args.cats == "bcd") else None
        if args.workload in ("auto", "c3"):
            b = block(m, wl, args.mode, tr_native if args.mode == "native" else None)
        else:
            b = block(m, wl, args.mode, tr_c4)
        line = {
            "metric": METRIC, "value": m["value"], "unit": wl["unit"], "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["name"], "rng_mode": names[args.mode], "chunksz": CHUNKSZ, "scale": args.scale,
                       "l2": "per-step working set (packed genome + FASTQ output) is GBs, far larger than the 126 MB L2",
                       "parallelism": f"windows sharded across {world} rank(s); packed genome broadcast over NCCL",
                       "blocks": "replay = the same workload in REPLAY (bit-exact) mode; c4 = BASELINE configs[3], 12.5M pairs per GPU"},
            "output_gb_per_s": m["output_gb_per_s"],
            "reads_per_step": m["reads_per_step"],
            "wall_ms_per_step": m["wall_ms_per_step"],
            "phase_ms_per_step_rank0": m["phase"],
            "roofline": b["roofline"],
            "roofline_step": b["roofline_step"],
            "e2e": m["e2e"],
            "gpu_launches": m["launches"] + (m2["launches"] if m2 else 0) + (m4["launches"] if m4 else 0) + (m4r["launches"] if m4r else 0),
            "clocks": m["clocks"],
            "setup_s": setup_s,
        }
        if d2h_probe is not None and line["e2e"] is not None:
            line["e2e"]["d2h_probe"] = d2h_probe
        if m2 is not None:
            line[other] = block(m2, wl, other)
        if m4 is not None:
            line["c4"] = block(m4, wl4, args.mode, tr_c4)
            if m4r is not None:
                line["c4"][other] = block(m4r, wl4, other)
        if n_gpus == 1 and not args.no_cpu_baseline:
            res, why = run_reference_steps(wl, 1, 0)
            if res is not None:
                line["cpu_baseline"] = {"value": res["reads_per_s"], "unit": wl["unit"], "cores": res["cores"],
                                        "kind": "reference", "sample": res["sample"], "gb_per_s": res["gb_per_s"]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": wl["unit"], "cores": 0, "kind": "reference", "sample": why}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
