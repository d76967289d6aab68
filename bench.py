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
  roofline  algorithmic bytes (B_shot from the search's own counters Q,P,L,S,A,V) over the
          search kernel's CUDA-event time, against the measured HBM copy bandwidth.
  cpu_baseline  the reference's own decoder (oracle/_ref, built from the unmodified reference sources) or the
          CPU port, shot-parallel on all host cores, on a bounded prefix of the same shots; every shot it
          decodes is also compared with the GPU's answer.

--impl reference times that CPU decoder alone (rank 0 only), on the same workload.
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

# Bench workloads: BASELINE.json configs, SURVEY.md §8d flags.  The headline metric is quoted on the
# bivariate-bicycle [[144,12,12]] code (configs[3]); the config leaves p open — p=1e-4 is the noise level at
# which the literal flags (unbounded beam and queue) are a million-shot workload at all.
WORKLOADS = {
    "bb144": dict(cfg=W.CONFIGS["C4"], shots=262144,
                  name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=1e-4 X, --num-det-orders 24 --no-revisit-dets "
                       "(DetIndex, seed 518278944), beam/pqlimit unbounded (CLI defaults)"),
    "bb144_p5e-4": dict(cfg=dict(W.CONFIGS["C4"], dem="bb144_r12_p0.0005_X"), shots=4096,
                        name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=5e-4 X, --num-det-orders 24 --no-revisit-dets"),
    # the reference's own published sweep point (benchmarking/sparsify_errors/aggregated_results.jsonl:195, submit.sh:223):
    # 4.105 CPU-seconds per shot on their hardware
    "bb144_p1e-3_longbeam": dict(cfg=dict(dem="bb144_r12_p0.001_X", det_beam=20, beam_climbing=True, pqlimit=1000000,
                                          no_revisit_dets=True, num_det_orders=21), shots=1024,
                                 name="bivariate-bicycle [[144,12,12]] r=12 si1000 p=1e-3 X, --beam 20 --beam-climbing --num-det-orders 21 "
                                      "--pqlimit 1000000 --no-revisit-dets (the reference's long-beam sweep point)"),
    "surface_d5": dict(cfg=W.CONFIGS["C1"], shots=1 << 20, name="surface code d=5 r=5 p=0.002 X, --beam 5 --pqlimit 200000"),
    "color_d7": dict(cfg=dict(W.CONFIGS["C2"], det_beam=5, pqlimit=200000, no_revisit_dets=True), shots=65536,
                     name="superdense color code d=7 r=7 p=0.001 X, --num-det-orders 1 --beam 5 --pqlimit 200000 --no-revisit-dets"),
    "surface_d11": dict(cfg=W.CONFIGS["C3"], shots=8192, name="surface code d=11 r=11 p=0.001 X, --beam 23 --beam-climbing"),
    "transcx_d5": dict(cfg=W.CONFIGS["C5"], shots=262144, name="transversal-CNOT surface code d=5 p=0.001 X, --det-penalty 1"),
}
L2_FLUSH_BYTES = 256 << 20


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def cpu_decoder(dem, kw, threads):
    from oracle import oracle
    kind = "reference" if oracle.available("reference") else "port"
    return oracle.OracleDecoder(dem, kind=kind, num_threads=threads, **kw), kind


def cpu_sample(dec, dets, target_seconds, threads):
    """Decode a prefix of `dets` sized for about target_seconds of wall time; returns (n, wall, result)."""
    n0 = min(len(dets), max(2 * threads, 16))
    r0 = dec.decode_batch_b8(dets[:n0])
    rate = n0 / max(r0["wall_seconds"], 1e-6)
    n = int(min(len(dets), max(n0, rate * target_seconds)))
    r = dec.decode_batch_b8(dets[:n])
    return n, r["wall_seconds"], r


def run_reference(args, wl, rank):
    if rank != 0:
        return
    from tesseract_decoder_b200 import capi  # host-side helpers only (DEM sampler, det orders): no GPU use
    dem, kw = W.decoder_kwargs(wl["cfg"], capi.build_det_orders)
    threads = os.cpu_count() or 1
    dec, kind = cpu_decoder(dem, kw, threads)
    D, O = dec.num_detectors, dec.num_observables
    shots = args.shots or wl["shots"]
    dets, _ = capi.sample_dem_b8(dem, 1000, shots, D, O)
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


def run_b200(args, wl, rank, world, local):
    import torch
    from tesseract_decoder_b200 import capi

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dem, kw = W.decoder_kwargs(wl["cfg"], capi.build_det_orders)
    dec = capi.Decoder(dem, device=local, warps_per_block=args.warps, force_generic_kernel=args.scan == "generic",
                       row_scan={"lanes": 1, "bits": 2}.get(args.scan, 0), **kw)
    dec.set_collect_errors(False)
    D, O = dec.num_detectors, dec.num_observables
    S = args.shots or wl["shots"]
    stride, ostride = (D + 7) // 8, (O + 7) // 8
    dets, true_obs = capi.sample_dem_b8(dem, 1000 + rank, S, D, O)

    # Device-resident buffers (value leg) and pinned host buffers (e2e leg).
    d_dets = torch.from_numpy(dets).cuda()
    d_true = torch.from_numpy(np.ascontiguousarray(true_obs)).cuda()
    d_obs = torch.zeros((S, ostride), dtype=torch.uint8, device="cuda")
    d_cost = torch.zeros(S, dtype=torch.float64, device="cuda")
    d_low = torch.zeros(S, dtype=torch.uint8, device="cuda")
    h_dets = torch.from_numpy(dets).pin_memory()
    h_obs = torch.zeros((S, ostride), dtype=torch.uint8).pin_memory()
    h_cost = torch.zeros(S, dtype=torch.float64).pin_memory()
    h_low = torch.zeros(S, dtype=torch.uint8).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    tally = torch.zeros(3, dtype=torch.int64, device="cuda")
    acc = {"search_ms": 0.0, "pipeline_ms": 0.0, "launches": 0, "rerun": 0}

    def note_timing():
        t = dec.last_batch_timing()
        chunks = (S + (1 << 18) - 1) >> 18
        acc["search_ms"] += t["search_ms"]
        acc["pipeline_ms"] += t["pipeline_ms"]
        acc["launches"] += 4 * chunks + 2 * t["search_launches"]  # stage + 3 ordering kernels per chunk, search + resolve per tier
        acc["rerun"] += t["rerun_items"]

    def reduce_counts(wrong, low):
        tally[0], tally[1], tally[2] = S, wrong, low
        if dist is not None:
            dist.all_reduce(tally)  # the path's only collective: {shots, logical errors, low confidence}
        return tally.tolist()

    def step_device():
        flush.zero_()  # evict L2 between steps (256 MB > 126 MB)
        torch.cuda.synchronize()
        dec.decode_batch_b8_device(d_dets.data_ptr(), stride, S, d_obs.data_ptr(), d_cost.data_ptr(), d_low.data_ptr())
        note_timing()
        bad = ((d_obs != d_true).any(dim=1) & (d_low == 0)).sum()
        return reduce_counts(bad, d_low.sum())

    def step_e2e():
        flush.zero_()
        torch.cuda.synchronize()
        dec.decode_batch_b8_raw(h_dets.data_ptr(), stride, S, h_obs.data_ptr(), h_cost.data_ptr(), h_low.data_ptr())
        note_timing()
        low = h_low.numpy()
        bad = int(((h_obs.numpy() != true_obs).any(axis=1) & (low == 0)).sum())
        return reduce_counts(bad, int(low.sum()))

    # Algorithmic bytes of one batch, from the search's own operation counters.  Counter S
    # (row entries the reference's early-exit loop visits) needs the instrumented kernel variant, so it is
    # taken from one untimed pass over the same batch; the timed steps run the production variant.
    dec.set_instrumented(True)
    dec.counters(reset=True)
    step_device()
    ctr1 = dec.counters(reset=True)
    dec.set_instrumented(False)

    def timed(step):
        for _ in range(args.warmup):
            step()
        for k in acc:
            acc[k] = 0
        dec.counters(reset=True)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        t0 = time.perf_counter()
        for _ in range(args.steps):
            counts = step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([dt, acc["pipeline_ms"], acc["search_ms"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        return t.tolist(), counts, clocks, dict(acc), dec.counters()

    (dt, pipe_ms, search_ms), counts, clocks, a, ctr = timed(step_device)
    (dt2, _, _), counts2, _, _, _ = timed(step_e2e)
    total = S * world * args.steps
    value, e2e = total / dt, total / dt2

    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes(ctr1, S, D, O) * args.steps  # rank 0's launches: the same batch every step
    assert all(ctr[k] == ctr1[k] * args.steps for k in ("Q", "P", "X", "L", "V")), (ctr, ctr1)
    n_launch = args.steps * ((S + (1 << 18) - 1) >> 18)  # tier-0 search launches (one per 262144-shot chunk)
    achieved = alg / (a["search_ms"] * 1e-3) / 1e9 if a["search_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "tsr::search_fast_kernel" if dec.search_path.startswith("fast") else "tsr::search_kernel", "algorithmic_bytes_per_launch": alg / n_launch,
                "algorithmic_bytes_per_shot": alg / (S * args.steps), "kernel_ms_per_launch": a["search_ms"] / n_launch,
                "launches_timed": n_launch, "scanned_entries_per_shot": ctr1["S"] / S, "search_path": dec.search_path,
                "counters_per_batch": ctr1}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            per_shot = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_shot")
            if per_shot:
                roofline["traffic"] = per_shot * min(S, 1 << 18)  # DRAM bytes per launch, from the committed ncu capture
        except Exception:
            pass

    line = {
        "metric": "decoded shots/sec", "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (i.i.d. DEM-sampled shots, seed 1000+rank)",
        "config": {"workload": wl["name"], "shots_per_gpu_per_step": S, "detectors": D, "errors": dec.num_errors, "observables": O,
                   "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MB write)", "sharding": f"shots x{world}, tables replicated"},
        "device_ms_per_step": pipe_ms / args.steps, "search_kernel_ms_per_step": search_ms / args.steps,
        "e2e": {"value": e2e, "unit": "shots/s", "h2d_bytes_per_step": S * stride, "d2h_bytes_per_step": S * (ostride + 9),
                "ms_per_step": 1e3 * dt2 / args.steps},
        "gpu_launches": a["launches"], "rerun_items": a["rerun"],
        "tally": {"shots": counts[0], "logical_errors": counts[1], "low_confidence": counts[2], "e2e_equal": counts == counts2},
        "roofline": roofline, "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        cpu, kind = cpu_decoder(dem, kw, threads)
        n, wall, ref = cpu_sample(cpu, dets, args.cpu_seconds, threads)
        mism = int((ref["obs"] != h_obs.numpy()[:n]).any(axis=1).sum() + (ref["low_confidence"] != h_low.numpy()[:n]).sum()
                   + (ref["cost"] != h_cost.numpy()[:n]).sum())
        line["cpu_baseline"] = {"value": n / wall, "unit": "shots/s", "cores": threads, "kind": kind,
                                "sample": f"first {n} shots of the same batch, {wall:.1f} s wall",
                                "cpu_seconds_per_shot": ref["sum_shot_seconds"] / n,
                                "parity": {"shots": n, "mismatches_obs_cost_lowconf": mism}}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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
