#!/usr/bin/env python
"""bench.py — decoded shots/s of the batched A* most-likely-error decode path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload bb144] [--shots S] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path (tsr_decode_batch_b8*, include/tesseract_b200.h) over one batch of S
synthetic shots per GPU, sampled i.i.d. from the workload's detector error model (stim is not available;
SURVEY.md F3).  Shots shard across ranks (rank r decodes its own S shots: weak scaling, no data-path
collective); per step one all-reduce sums {shots, logical errors, low confidence}.

  value   shots/s with the syndromes already resident in HBM (tsr_decode_batch_b8_device), whole job.
  e2e     the same through the host-buffer entry point (tsr_decode_batch_b8) with pinned host buffers:
          H2D of the bit-packed syndromes and D2H of observables / costs / flags inside the timed region.
  roofline  the task's HBM-class figure: algorithmic bytes (B_shot from the search's own counters
          Q,P,L,S,A,V — the bytes the REFERENCE's loops touch for the same search) over the search kernel's
          CUDA-event time, against the measured HBM copy bandwidth.  The tables are L2-resident, so this is a
          work rate, not DRAM use; `issue` beside it is what actually bounds the kernel (warp instructions
          per second against the SM issue peak, from the committed ncu capture).
  cpu_baseline  the reference's own decoder (oracle/_ref, built from the unmodified reference sources) or the
          CPU port, shot-parallel on all host cores, on a bounded prefix of the same shots; every shot it
          decodes is compared with the GPU's answer (observables, costs, flags and error lists).
  workloads  (N=1, default workload only) the other BASELINE.json configs at their literal flags, each with
          value / e2e / roofline frac / cpu / mismatches, on batches sized to a few seconds.

Any mismatch against the CPU decoder makes the run exit non-zero (after printing the line).
--impl reference times that CPU decoder alone (rank 0 only), on the same workload; it loads only oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import workloads as W  # noqa: E402

# Bench workloads: BASELINE.json configs at their literal flags.  The headline metric is quoted on
# the bivariate-bicycle [[144,12,12]] code (configs[3]); the config leaves p open — p=1e-4 is the noise level at
# which the literal flags (unbounded beam and queue) are a million-shot workload at all.
# `side`: (shots, steps, warmup, cpu_seconds) when the workload runs as an entry of the headline's `workloads` array.
WORKLOADS = {
    "bb144": dict(cfg=W.CONFIGS["C4"], shots=262144, tag="C4 bb144 p=1e-4",
                  name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=1e-4 X, --num-det-orders 24 --no-revisit-dets "
                       "(DetIndex, seed 518278944), beam/pqlimit unbounded (CLI defaults)"),
    "surface_d5": dict(cfg=W.CONFIGS["C1"], shots=1 << 20, tag="C1 surface d5", side=(262144, 3, 1, 2.0),
                       name="surface code d=5 r=5 p=0.002 X, --beam 5 --pqlimit 200000"),
    # Heavy-tailed at its literal flags: in this 2000-shot batch (seed 1002, the golden set g_c2_color_d7_literal) the
    # median search pushes 837 nodes, shot 631 pushes 35.6 M (226 s on one CPU core).  As a side workload the first 600
    # shots are timed (largest search: 305 544 pushes); `--workload color_d7` decodes all 2000.
    "color_d7": dict(cfg=W.CONFIGS["C2"], shots=2000, sample=(1002, 2000), tag="C2 color d7 literal (600 of 2000 shots)",
                     side=(600, 2, 1, 3.0),
                     name="superdense color code d=7 r=7 p=0.001 X, --num-det-orders 1, CLI defaults otherwise "
                          "(beam and pqlimit unbounded, revisits allowed)"),
    "surface_d11": dict(cfg=W.CONFIGS["C3"], shots=8192, tag="C3 surface d11 climb", side=(2048, 2, 1, 3.0),
                        name="surface code d=11 r=11 p=0.001 X, --beam 23 --beam-climbing"),
    "transcx_d5": dict(cfg=W.CONFIGS["C5"], shots=262144, tag="C5 transcx d5 penalty", side=(131072, 3, 1, 2.0),
                       name="transversal-CNOT surface code d=5 p=0.001 X, --det-penalty 1"),
    "transcx_d5_amt": dict(cfg=dict(W.CONFIGS["C5"], at_most_two_errors_per_detector=True), shots=262144,
                           tag="C5 transcx d5 penalty at-most-two", side=(131072, 3, 1, 2.0), cpu_kind="port",
                           name="transversal-CNOT surface code d=5 p=0.001 X, --det-penalty 1 --at-most-two-errors-per-detector "
                                "(no reference code for this flag: checked against the CPU port)"),
    # the reference's own published sweep point (benchmarking/sparsify_errors/aggregated_results.jsonl:195, submit.sh:223):
    # 4.105 CPU-seconds per shot on their hardware
    "bb144_p1e-3_longbeam": dict(cfg=dict(dem="bb144_r12_p0.001_X", det_beam=20, beam_climbing=True, pqlimit=1000000,
                                          no_revisit_dets=True, num_det_orders=21), shots=1024, tag="bb144 p=1e-3 long-beam",
                                 side=(256, 1, 1, 4.0), counters_from="port", cpu_first=16, counter_shots=8,
                                 name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=1e-3 X, --beam 20 --beam-climbing --num-det-orders 21 "
                                      "--pqlimit 1000000 --no-revisit-dets (the reference's long-beam sweep point)"),
    # ... and the same point with --sparsify-errors at the reference's own best setting for it (base degree 3, reactivate limit
    # 128: 0.116 CPU-seconds per shot, aggregated_results.jsonl; tesseract.cc:256-292, 673-737)
    "bb144_p1e-3_longbeam_sparsify": dict(cfg=dict(dem="bb144_r12_p0.001_X", det_beam=20, beam_climbing=True, pqlimit=1000000,
                                                   no_revisit_dets=True, num_det_orders=21, sparsify_errors=True, sparsify_base_degree=3,
                                                   sparsify_reactivate_limit=128),
                                          shots=4096, tag="bb144 p=1e-3 long-beam --sparsify-errors (base 3, limit 128)",
                                          side=(2048, 2, 1, 4.0), counters_from="port", counter_shots=64,
                                          name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=1e-3 X, long-beam flags + --sparsify-errors "
                                               "--sparsify-base-degree 3 --sparsify-reactivate-limit 128"),
    "bb144_p5e-4": dict(cfg=dict(W.CONFIGS["C4"], dem="bb144_r12_p0.0005_X"), shots=4096, tag="C4 bb144 p=5e-4",
                        name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=5e-4 X, --num-det-orders 24 --no-revisit-dets"),
}
SIDE_ORDER = ["surface_d5", "color_d7", "surface_d11", "transcx_d5", "transcx_d5_amt", "bb144_p1e-3_longbeam",
              "bb144_p1e-3_longbeam_sparsify"]
L2_FLUSH_BYTES = 256 << 20
MAX_CHUNK = 1 << 18  # tsr_config.max_batch_shots default: shots per search launch
# Issue-slot peak of one B200: 148 SMs x 4 schedulers x 1 warp instruction per cycle at the measured max SM clock.
ISSUE_PEAK_PER_CLOCK = 148 * 4


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
            return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(c, shots, D, O):
    """B = shots*(ceil(D/8)+ceil(O/8)+9) + Σ_searches[16(Q+P) + 12(Q-1) + 12L + 16S + 4A + W·V]."""
    Wv = 8 * ((D + 63) // 64)
    return (shots * ((D + 7) // 8 + (O + 7) // 8 + 9) + 16 * (c["Q"] + c["P"]) + 12 * (c["Q"] - c["searches"]) + 12 * c["L"]
            + 16 * c["S"] + 4 * c["A"] + Wv * c["V"])


def cpu_decoder(dem, kw, threads, kind=None):
    from oracle import oracle
    if kind is None:
        kind = "reference" if oracle.available("reference") else "port"
    if kw.get("at_most_two_errors_per_detector") and kind == "reference":
        kind = "port"  # the reference snapshot has no code for this flag
    kw = {k: v for k, v in kw.items() if not (k == "at_most_two_errors_per_detector" and not v)}
    return oracle.OracleDecoder(dem, kind=kind, num_threads=threads, **kw), kind


def cpu_sample(dec, dets, target_seconds, threads, n_first=None):
    """Decode a prefix of `dets` sized for about target_seconds of wall time; returns (n, wall, result)."""
    n0 = min(len(dets), n_first or max(2 * threads, 16))
    r0 = dec.decode_batch_b8(dets[:n0])
    if r0["wall_seconds"] >= target_seconds or n0 == len(dets):
        return n0, r0["wall_seconds"], r0
    rate = n0 / max(r0["wall_seconds"], 1e-6)
    n = int(min(len(dets), max(n0, rate * target_seconds)))
    r = dec.decode_batch_b8(dets[:n])
    return n, r["wall_seconds"], r


def run_reference(args, wl, rank):
    """The reference's own CPU decoder on this workload.  Loads oracle/ only: det orders and the synthetic shots come
    from the checker side (oracle.build_det_orders, oracle.sample_dem_b8), not from the product library."""
    if rank != 0:
        return
    from oracle import oracle
    kind = "reference" if oracle.available("reference") else "port"
    dem, kw = W.decoder_kwargs(wl["cfg"], lambda d, n, m, s: oracle.build_det_orders(d, n, m, s, kind=kind))
    threads = os.cpu_count() or 1
    dec, kind = cpu_decoder(dem, kw, threads, kind)
    D, O = dec.num_detectors, dec.num_observables
    shots = args.shots or wl["shots"]
    seed, n_sampled = wl.get("sample", (1000, None))
    dets, _ = oracle.sample_dem_b8(dem, seed, max(n_sampled or shots, shots), kind=kind)
    dets = dets[:shots]
    budget = args.ref_seconds / (args.steps + args.warmup)  # seconds of CPU wall per step
    n, _, _ = cpu_sample(dec, dets, budget, threads)
    for _ in range(args.warmup):
        dec.decode_batch_b8(dets[:n])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_batch_b8(dets[:n])
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"first {n} of the {shots} shots of rank 0's batch per step"
    print(json.dumps({
        "impl": "reference", "metric": "decoded shots/sec", "value": value, "unit": "shots/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "shots_per_step": n, "detectors": D, "observables": O},
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class Bench:
    """One workload on this rank's GPU: buffers, the two timed legs, counters and the CPU comparison."""

    def __init__(self, args, wl, S, rank, world, local, dist):
        import torch
        from tesseract_decoder_b200 import capi
        self.torch, self.capi, self.args, self.wl, self.S = torch, capi, args, wl, S
        self.rank, self.world, self.local, self.dist = rank, world, local, dist
        self.dem, self.kw = W.decoder_kwargs(wl["cfg"], capi.build_det_orders)
        self.dec = capi.Decoder(self.dem, device=local, warps_per_block=args.warps, force_generic_kernel=args.scan == "generic",
                                row_scan={"lanes": 1, "bits": 2}.get(args.scan, 0), **self.kw)
        self.dec.set_collect_errors(False)
        self.D, self.O = self.dec.num_detectors, self.dec.num_observables
        self.stride, self.ostride = (self.D + 7) // 8, (self.O + 7) // 8
        seed, n_sampled = wl.get("sample", (1000, None))
        self.dets, self.true_obs = capi.sample_dem_b8(self.dem, seed + rank, max(n_sampled or S, S), self.D, self.O)
        self.dets, self.true_obs = np.ascontiguousarray(self.dets[:S]), np.ascontiguousarray(self.true_obs[:S])
        # Device-resident buffers (value leg) and pinned host buffers (e2e leg).
        self.d_dets = torch.from_numpy(self.dets).cuda()
        self.d_true = torch.from_numpy(np.ascontiguousarray(self.true_obs)).cuda()
        self.d_obs = torch.zeros((S, self.ostride), dtype=torch.uint8, device="cuda")
        self.d_cost = torch.zeros(S, dtype=torch.float64, device="cuda")
        self.d_low = torch.zeros(S, dtype=torch.uint8, device="cuda")
        self.h_dets = torch.from_numpy(self.dets).pin_memory()
        self.h_obs = torch.zeros((S, self.ostride), dtype=torch.uint8).pin_memory()
        self.h_cost = torch.zeros(S, dtype=torch.float64).pin_memory()
        self.h_low = torch.zeros(S, dtype=torch.uint8).pin_memory()
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.tally = torch.zeros(3, dtype=torch.int64, device="cuda")
        self.acc = {"search_ms": 0.0, "pipeline_ms": 0.0, "launches": 0, "rerun": 0}

    def note_timing(self):
        t = self.dec.last_batch_timing()
        chunks = (self.S + MAX_CHUNK - 1) // MAX_CHUNK
        self.acc["search_ms"] += t["search_ms"]
        self.acc["pipeline_ms"] += t["pipeline_ms"]
        self.acc["launches"] += 4 * chunks + 2 * t["search_launches"]  # stage + 3 ordering kernels per chunk, search + resolve per tier
        self.acc["rerun"] += t["rerun_items"]

    def init_comm(self):
        """N > 1: the product's own NCCL communicator (tsr_comm_*, csrc/group.cu).  torch.distributed only ships the
        128-byte unique id from rank 0 (launcher plumbing) and provides the barrier / max-over-ranks of the timing."""
        if self.dist is None:
            return
        box = [self.capi.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        self.dec.comm_init_rank(box[0], self.rank, self.world)

    def reduce_counts(self, wrong, low, shots=None):
        counts = [self.S if shots is None else shots, int(wrong), int(low)]
        if self.dist is not None:
            # the path's only collective: one ncclAllReduce of {shots, logical errors, low confidence} (tsr_comm_allreduce_sum_u64)
            counts = self.dec.allreduce_sum(counts).tolist()
        return counts

    def step_device(self):
        self.flush.zero_()  # evict L2 between steps (256 MB > 126 MB)
        self.torch.cuda.synchronize()
        self.dec.decode_batch_b8_device(self.d_dets.data_ptr(), self.stride, self.S, self.d_obs.data_ptr(), self.d_cost.data_ptr(),
                                        self.d_low.data_ptr())
        self.note_timing()
        bad = ((self.d_obs != self.d_true).any(dim=1) & (self.d_low == 0)).sum()
        return self.reduce_counts(bad, self.d_low.sum())

    def step_e2e(self):
        self.flush.zero_()
        self.torch.cuda.synchronize()
        self.dec.decode_batch_b8_raw(self.h_dets.data_ptr(), self.stride, self.S, self.h_obs.data_ptr(), self.h_cost.data_ptr(),
                                     self.h_low.data_ptr())
        self.note_timing()
        low = self.h_low.numpy()
        bad = int(((self.h_obs.numpy() != self.true_obs).any(axis=1) & (low == 0)).sum())
        return self.reduce_counts(bad, int(low.sum()))

    def timed(self, step, steps, warmup):
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            step()
        for k in self.acc:
            self.acc[k] = 0
        self.dec.counters(reset=True)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        t0 = time.perf_counter()
        for _ in range(steps):
            counts = step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([dt, self.acc["pipeline_ms"], self.acc["search_ms"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        return t.tolist(), counts, clocks, dict(self.acc), self.dec.counters()

    def instrumented_counters(self):
        """Counters of one batch, S included (the instrumented kernel variant: one untimed pass over the same batch)."""
        self.dec.set_instrumented(True)
        self.dec.counters(reset=True)
        self.step_device()
        c = self.dec.counters(reset=True)
        self.dec.set_instrumented(False)
        return c

    def compare_with_cpu(self, seconds, kind=None, want_counters=False):
        """Times the CPU decoder on a prefix of this batch and compares every shot it decoded with the GPU's answers
        (observables, cost, low-confidence flag and the error lists)."""
        threads = os.cpu_count() or 1
        cpu, kind = cpu_decoder(self.dem, self.kw, threads, kind)
        n, wall, ref = cpu_sample(cpu, self.dets, seconds, threads, self.wl.get("cpu_first"))
        self.dec.set_collect_errors(True)
        out = self.dec.decode_batch_b8(self.dets[:n])
        self.dec.set_collect_errors(False)
        lists_g = W.split_lists(out["err_offsets"], out["err_idx"])
        lists_c = W.split_lists(ref["err_offsets"], ref["err_idx"])
        bad = ((ref["obs"] != out["obs"]).any(axis=1) | (ref["low_confidence"] != out["low_confidence"]) | (ref["cost"] != out["cost"])
               | np.array([a != b for a, b in zip(lists_g, lists_c)], dtype=bool))
        # the timed e2e leg wrote the same answers for these shots
        bad |= (self.h_obs.numpy()[:n] != ref["obs"]).any(axis=1) | (self.h_low.numpy()[:n] != ref["low_confidence"]) \
            | (self.h_cost.numpy()[:n] != ref["cost"])
        mism = int(bad.sum())
        rec = {"value": n / wall, "unit": "shots/s", "cores": threads, "kind": kind,
               "sample": f"first {n} shots of the same batch, {wall:.1f} s wall, {mism} mismatches vs the GPU "
                         "(obs, cost, low_conf, error lists)",
               "cpu_seconds_per_shot": ref["sum_shot_seconds"] / n, "mismatches": mism, "shots_compared": n}
        ctr = None
        if want_counters:
            from oracle import oracle
            port = oracle.OracleDecoder(self.dem, kind="port", num_threads=threads,
                                        **{k: v for k, v in self.kw.items() if not (k == "at_most_two_errors_per_detector" and not v)})
            n_ctr = min(n, self.wl.get("counter_shots", n))
            port.decode_batch_b8(self.dets[:n_ctr])
            ctr = dict(port.counters(), shots=n_ctr)
        return rec, n, ctr


def run_side_workload(args, key, rank, world, local):
    """One entry of the `workloads` array: a BASELINE config at its literal flags on a batch sized to a few seconds."""
    wl = WORKLOADS[key]
    S, steps, warmup, cpu_s = wl["side"]
    peak, _, _ = measured_peaks()
    t_start = time.perf_counter()
    b = Bench(args, wl, S, rank, world, local, None)
    (dt, _, search_ms), counts, _, a, ctr = b.timed(b.step_device, steps, warmup)
    (dt2, _, _), counts2, _, _, _ = b.timed(b.step_e2e, steps, warmup)
    cpu, n_cpu, port_ctr = b.compare_with_cpu(cpu_s, wl.get("cpu_kind"), want_counters=wl.get("counters_from") == "port")
    if port_ctr is not None:  # counters of the CPU port on its prefix, scaled to the batch (distinct trials == literal trials here)
        b_shot = algorithmic_bytes(port_ctr, port_ctr["shots"], b.D, b.O) / port_ctr["shots"]
        src = f"CPU port on {port_ctr['shots']} shots"
    else:
        c1 = b.instrumented_counters()
        assert all(ctr[k] == c1[k] * steps for k in ("Q", "P", "X", "L", "V")), (ctr, c1)
        b_shot = algorithmic_bytes(c1, S, b.D, b.O) / S
        src = "instrumented pass"
    achieved = b_shot * S * steps / (a["search_ms"] * 1e-3) / 1e9 if a["search_ms"] > 0 else 0.0
    rec = {"w": wl["tag"], "shots": S, "steps": steps, "value": round(S * steps / dt, 1), "e2e": round(S * steps / dt2, 1),
           "frac": round(achieved / peak, 4), "MB_per_shot": round(b_shot / 1e6, 3), "cpu": round(cpu["value"], 1), "cpu_kind": cpu["kind"],
           "cpu_shots": n_cpu, "mism": cpu["mismatches"], "low_conf": counts[2], "errors": counts[1],
           "path": b.dec.search_path, "rerun": a["rerun"]}
    print(f"[bench] {wl['tag']}: {json.dumps(rec)} (B_shot from the {src}; {time.perf_counter() - t_start:.1f} s)", file=sys.stderr, flush=True)
    del b
    return rec


def api_legs(args, local):
    """The entry points a reference user actually calls, on BASELINE configs[0] (surface code d=5): the pybind11
    sinter decoder's decode_shots_bit_packed with a pageable numpy array (tesseract_sinter_compat.pybind.h:45-97) and the
    latency of a single-shot decode_to_errors (tesseract.cc:295-347; the reference CLI and Python decode call it per shot)."""
    import tesseract_decoder_b200 as td
    from tesseract_decoder_b200 import capi
    cfg = W.CONFIGS["C1"]
    dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
    S = 1 << 18
    dets, _ = capi.sample_dem_b8(dem, 1000, S)

    class Dem:
        def __str__(self):
            return dem
    sinter = td.tesseract_sinter_compat.TesseractSinterDecoder(det_beam=kw["det_beam"], beam_climbing=False, no_revisit_dets=kw["no_revisit_dets"],
                                                                 pqlimit=kw["pqlimit"], num_det_orders=0)
    compiled = sinter.compile_decoder_for_dem(dem=Dem())
    compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets[:4096])
    t0 = time.perf_counter()
    for _ in range(3):
        compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    pybind_rate = 3 * S / (time.perf_counter() - t0)
    dec = capi.Decoder(dem, device=local, **kw)
    bits = np.unpackbits(dets[:256], axis=1, bitorder="little")[:, :dec.num_detectors]
    hits = [np.nonzero(r)[0] for r in bits]
    for h in hits[:16]:
        dec.decode_to_errors(h)
    t0 = time.perf_counter()
    for h in hits:
        dec.decode_to_errors(h)
    single_ms = 1e3 * (time.perf_counter() - t0) / len(hits)
    return {"workload": "surface code d=5 (configs[0])", "pybind_decode_shots_bit_packed_shots_per_s": round(pybind_rate, 1),
            "pybind_input": f"pageable numpy uint8[{S}, {dets.shape[1]}]", "single_shot_decode_to_errors_ms": round(single_ms, 4)}


def strong_leg(args, b, rank, world):
    """Strong scaling: a FIXED batch of T shots (the same on every rank: seed 2000) dealt out in interleaved chunks of
    4096 shots (rank r decodes chunks r, r+N, ...), decoded end to end from host buffers, counters summed
    by the one all-reduce.  value = T x steps / max-over-ranks time."""
    from tesseract_decoder_b200 import sharding
    torch, dist = b.torch, b.dist
    T, steps, warmup, chunk = args.strong_shots, args.strong_steps, 1, 4096
    dets, true_obs = b.capi.sample_dem_b8(b.dem, 2000, T, b.D, b.O)
    mine = np.ascontiguousarray(sharding.shard_rows(dets, rank, world, chunk))
    truth = np.ascontiguousarray(sharding.shard_rows(true_obs, rank, world, chunk))
    n = len(mine)
    h_dets = torch.from_numpy(mine).pin_memory()
    h_obs = torch.zeros((max(n, 1), b.ostride), dtype=torch.uint8).pin_memory()
    h_low = torch.zeros(max(n, 1), dtype=torch.uint8).pin_memory()

    def step():
        b.flush.zero_()
        torch.cuda.synchronize()
        if n:
            b.dec.decode_batch_b8_raw(h_dets.data_ptr(), b.stride, n, h_obs.data_ptr(), 0, h_low.data_ptr())
        low = h_low.numpy()[:n]
        bad = int(((h_obs.numpy()[:n] != truth).any(axis=1) & (low == 0)).sum())
        return b.reduce_counts(bad, int(low.sum()), shots=n)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        counts = step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {"value": T * steps / dt, "unit": "shots/s", "total_shots": T, "shots_this_rank": n, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * dt / steps, "partition": f"interleaved chunks of {chunk} shots over {world} rank(s)",
            "tally": {"shots": counts[0], "logical_errors": counts[1], "low_confidence": counts[2]},
            "limiter": "last wave of the search kernel on each rank (per-rank batch shrinks with N); the collective is 24 bytes"}


def run_b200(args, wl, rank, world, local):
    import torch

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    S = args.shots or wl["shots"]
    b = Bench(args, wl, S, rank, world, local, dist)
    b.init_comm()
    dec, D, O = b.dec, b.D, b.O

    # Algorithmic bytes of one batch, from the search's own operation counters.  Counter S
    # (row entries the reference's early-exit loop visits) needs the instrumented kernel variant, so it is
    # taken from one untimed pass over the same batch; the timed steps run the production variant.
    ctr1 = b.instrumented_counters()
    (dt, pipe_ms, search_ms), counts, clocks, a, ctr = b.timed(b.step_device, args.steps, args.warmup)
    (dt2, _, _), counts2, _, _, _ = b.timed(b.step_e2e, args.steps, args.warmup)
    total = S * world * args.steps
    value, e2e = total / dt, total / dt2

    peak, sm_max_mhz, peak_src = measured_peaks()
    alg = algorithmic_bytes(ctr1, S, D, O) * args.steps  # rank 0's launches: the same batch every step
    assert all(ctr[k] == ctr1[k] * args.steps for k in ("Q", "P", "X", "L", "V")), (ctr, ctr1)
    n_launch = args.steps * ((S + MAX_CHUNK - 1) // MAX_CHUNK)  # tier-0 search launches (one per 262144-shot chunk)
    achieved = alg / (a["search_ms"] * 1e-3) / 1e9 if a["search_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "tsr::search_fast_kernel" if dec.search_path.startswith("fast") else "tsr::search_kernel",
                "what": "reference-equivalent algorithmic bytes (SURVEY.md 8d) / search-kernel time; tables are L2-resident, see `issue`",
                "algorithmic_bytes_per_launch": alg / n_launch, "algorithmic_bytes_per_shot": alg / (S * args.steps),
                "kernel_ms_per_launch": a["search_ms"] / n_launch, "launches_timed": n_launch,
                "scanned_entries_per_shot": ctr1["S"] / S, "search_path": dec.search_path}
    issue = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            rec = json.load(open(prof)).get(args.workload, {})
            if rec.get("dram_bytes_per_shot"):
                roofline["traffic"] = rec["dram_bytes_per_shot"] * min(S, MAX_CHUNK)  # DRAM bytes per launch, from the committed ncu capture
            if rec.get("warp_instructions_per_shot"):
                # what bounds the kernel: warp instructions issued per second against 148 SMs x 4 schedulers x SM clock
                wi = rec["warp_instructions_per_shot"] * S * args.steps / (a["search_ms"] * 1e-3)
                pk = ISSUE_PEAK_PER_CLOCK * sm_max_mhz * 1e6
                issue = {"bound": "issue", "achieved": wi / 1e9, "peak": pk / 1e9, "unit": "G warp-instr/s", "frac": wi / pk,
                         "warp_instructions_per_shot": rec["warp_instructions_per_shot"], "source": rec.get("source", "profiles/")}
                for k in ("l2_hit_pct", "l2_throughput_pct", "achieved_occupancy_pct", "dram_pct_of_peak", "issue_active_pct"):
                    if k in rec:
                        issue[k] = rec[k]
        except Exception:
            pass

    line = {
        "metric": "decoded shots/sec", "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (i.i.d. DEM-sampled shots, seed 1000+rank)",
        "config": {"workload": wl["name"], "shots_per_gpu_per_step": S, "detectors": D, "errors": dec.num_errors, "observables": O,
                   "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MB write)", "sharding": f"shots x{world}, tables replicated"},
        "device_ms_per_step": pipe_ms / args.steps, "search_kernel_ms_per_step": search_ms / args.steps,
        "e2e": {"value": e2e, "unit": "shots/s", "h2d_bytes_per_step": S * b.stride, "d2h_bytes_per_step": S * (b.ostride + 9),
                "ms_per_step": 1e3 * dt2 / args.steps},
        "gpu_launches": a["launches"], "rerun_items": a["rerun"],
        "tally": {"shots": counts[0], "logical_errors": counts[1], "low_confidence": counts[2], "e2e_equal": counts == counts2,
                  "collective": f"ncclAllReduce u64[3] per step via tsr_comm (world {dec.comm_world})" if world > 1 else "none (1 GPU)"},
        "roofline": roofline, "clocks": clocks,
    }
    if issue:
        line["issue"] = issue

    if args.strong_shots and args.workload == "bb144":
        line["scaling_strong"] = strong_leg(args, b, rank, world)

    mismatches = 0
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, n_cpu, _ = b.compare_with_cpu(args.cpu_seconds)
        mismatches += cpu["mismatches"]
        line["cpu_baseline"] = cpu
        line["config"]["parity"] = f"{cpu['mismatches']} mismatches / {n_cpu} shots vs the {cpu['kind']} CPU decoder"
        line["tally"]["mismatches"] = cpu["mismatches"]
    if rank == 0 and world == 1 and not args.no_side and args.workload == "bb144" and not args.shots:
        del b
        try:
            line["api"] = api_legs(args, local)
        except Exception as e:
            line["api"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        side = []
        for key in SIDE_ORDER:
            try:
                side.append(run_side_workload(args, key, rank, world, local))
            except Exception as e:  # a workload that cannot run is a failure of the run, reported in place
                side.append({"w": WORKLOADS[key]["tag"], "error": f"{type(e).__name__}: {e}"[:200], "mism": -1})
        line["workloads"] = side
        mismatches += sum(abs(r["mism"]) for r in side)
        line["config"]["parity"] += f"; side workloads: {sum(max(r['mism'], 0) for r in side)} mismatches, " \
                                    f"{sum(1 for r in side if 'error' in r)} failed"
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if mismatches:
        print(f"[bench] FAILED: {mismatches} parity mismatches / failed workloads against the CPU decoder", file=sys.stderr)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bb144", choices=sorted(WORKLOADS))
    ap.add_argument("--shots", type=int, default=0, help="shots per GPU per step (default: per workload)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="wall-time target of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="--impl reference: wall-time budget of the whole run")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the `workloads` array (the other BASELINE configs)")
    ap.add_argument("--strong-shots", type=int, default=1 << 20,
                    help="fixed total batch of the strong-scaling leg (`scaling_strong`), split over the ranks; 0 = skip")
    ap.add_argument("--strong-steps", type=int, default=2)
    ap.add_argument("--warps", type=int, default=0, help="warps per search CTA (0 = automatic)")
    ap.add_argument("--scan", default="auto", choices=["auto", "bits", "lanes", "generic"],
                    help="search kernel variant (auto = the decoder's choice)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
    else:
        run_b200(args, wl, rank, world, local)


if __name__ == "__main__":
    main()
