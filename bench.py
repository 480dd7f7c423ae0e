#!/usr/bin/env python
"""bench.py -- distinct states/sec of the explicit-state BFS hot path (BASELINE.json metric).

A "step" is one complete breadth-first model-checking job of the workload model (all reachable
states, invariants checked on every state) on N GPUs.  Workload: the committed compiled form of
BASELINE config #3 scaled to a single-GPU-sized state space (examples/Paxos, 3 acceptors / 2 values,
ballots 0..4, invariants Inv1-Inv4; tests/golden/MCPaxos3_b4.tlagz: 352,133,865 distinct / 3,462,635,854
generated states, depth 41 -- the >= 10^8-state configuration the metric is quoted on) --
configs #4/#5 (raft, SSI) are not lowered to the device yet.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (dominant kernel of the
step, k_wave), k1_roofline (the fingerprint/probe kernel alone on SURVEY §8d's synthetic batch),
cpu_baseline, e2e, clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    def __init__(self, dev=0):
        super().__init__(daemon=True)
        self.dev = dev
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


CPU_SAMPLE_STATES = 30_000_000   # bounded sample for the CPU arm: BFS stops after the level that reaches this many states


def cpu_reference(cm, init, info, threads, expect_distinct=0):
    """The reference arm: the path's CPU implementation (ORACLE O2, oracle/tlag_cpu.c -- TLC itself needs a JVM,
    which neither this image nor the reference provides) on all host cores, same model."""
    from oracle import cpu_engine
    t0 = time.time()
    stop = CPU_SAMPLE_STATES if expect_distinct > 2 * CPU_SAMPLE_STATES else 0
    want = min(expect_distinct, 3 * CPU_SAMPLE_STATES) if stop else expect_distinct
    cap = max(1 << 16, int(want * 1.25) + 4096) if want else 1 << 26   # right-sized store/table
    r = cpu_engine.run(cm, init, n_threads=threads, deadlock=info["deadlock"], max_states=cap, stop_after=stop)
    r["sample"] = (f"BFS prefix: the first {r['distinct']} distinct states (levels 1..{len(r['levels']) - 1}) of the workload"
                   if stop else "the whole workload model once") + " (oracle/tlag_cpu.c, all cores)"
    dt = r["seconds"]
    return r, dt, time.time() - t0


def k1_microbench(dev, peak):
    """n = 2^27 candidates x W = 20 words, 50 % duplicates, table 2^28 slots.
    Algorithmic bytes per candidate = S + 8 + p*8 + 1 = 93 (S = 80, p = 0.5)."""
    import torch
    from tla_rust_b200.engine import Engine, ProbeOnlyModel
    W, n = 20, 1 << 27
    g = torch.Generator(device=dev).manual_seed(0x5EED)
    half = torch.randint(-2**31, 2**31 - 1, (n // 2, W), dtype=torch.int32, device=dev, generator=g)
    perm = torch.randint(0, n // 2, (n // 2,), device=dev, generator=g)
    states = torch.cat([half, half[perm]])
    del half, perm
    flags = torch.zeros(n, dtype=torch.uint8, device=dev)
    e = Engine(ProbeOnlyModel(W), table_log2=28, device=torch.device(dev).index or 0, native=False)
    times = []
    for it in range(2 + 5):
        e.reset_table()
        ms = e.probe_batch_device(states.data_ptr(), n, flags.data_ptr())
        if it >= 2:
            times.append(ms)
    n_new = int(flags.sum().item())
    med = float(np.median(times))
    bytes_per = 80 + 8 + 0.5 * 8 + 1
    ach = bytes_per * n / (med * 1e-3) / 1e9
    launches = e.launches()
    e.close()
    del states, flags
    torch.cuda.empty_cache()
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["k_probe_staged"]
        if tj["n"] == n and tj["W"] == W:
            traffic = tj["bytes_per_launch"]
    except Exception:
        pass
    return {"kernel": "k_probe_staged", "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
            "frac": round(ach / peak, 4), "traffic": traffic, "n": n, "W": W, "p_new": n_new / n,
            "ms_per_launch": round(med, 4), "candidates_per_s": round(n / (med * 1e-3), 1),
            "bytes_per_candidate": bytes_per}, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="MCPaxos3_b4")
    ap.add_argument("--no-k1", action="store_true")
    args = ap.parse_args()
    if os.environ.get("TLAG_BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["TLAG_BENCH_WATCHDOG"]), exit=True)

    from tla_rust_b200.compiled import load_compiled
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cm, init, exp, info = load_compiled(os.path.join(ROOT, "tests", "golden", args.workload + ".tlagz"))
    cfg = {"workload": f"{args.workload}: {info.get('source', '')} (compiled fixture), W={cm.W} words/state, "
                       f"{exp['o2']['distinct']} distinct / {exp['o2']['generated']} generated states, "
                       f"{len(cm.invariants)} invariants", "l2": "state store + seen-set rebuilt every step (restart), "
                                                                 "working set streamed; see DESIGN.md",
           "parallelism": f"fp-hash-range x{args.gpus}",
           # TLAG_NATIVE=1 runs the headline on the model-specialised native build
           "engine_build": "native" if os.environ.get("TLAG_NATIVE", "0") == "1" else "interpreter"}
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for _ in range(max(1, min(args.steps, 2))):
            r, dt, wall = cpu_reference(cm, init, info, threads, exp["o2"]["distinct"])
            vals.append(r["distinct"] / dt)
        v = float(np.median(vals))
        line = {"impl": "reference", "metric": "distinct states/sec", "value": round(v, 1), "unit": "states/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 * r["distinct"] / v, 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": round(v, 1), "unit": "states/s", "cores": threads, "kind": "port",
                                 "sample": r["sample"] + "; TLC itself needs a JVM: absent"},
                "e2e": {"value": round(v, 1), "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from tla_rust_b200.engine import Engine
    dev = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    peak, peak_src = load_peaks()
    multi = world > 1 or bool(os.environ.get("TLAG_FORCE_ROUTE"))   # knob: exercise the routed path on one GPU
    if multi:
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(dev))
        else:
            dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    W = cm.W
    S = 4 * W
    h2d = int(cm.code.nbytes + cm.cpool.nbytes + cm.layout.nbytes + init.nbytes)
    stats = {}

    if not multi:
        e = Engine(cm, deadlock=info["deadlock"], device=local_rank)
        e.seed(init)
        for _ in range(max(args.warmup, 1)):
            e.restart()
            r = e.run()
        assert (r["verdict"], r["generated"], r["distinct"], r["depth"]) == (
            exp["o2"]["verdict"], exp["o2"]["generated"], exp["o2"]["distinct"], exp["o2"]["depth"]), r
        if sampler:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        kern_s = 0.0
        l0 = e.launches()
        for _ in range(args.steps):
            e.restart()
            r = e.run()
            kern_s += r["device_seconds"]
        barrier()
        dt = time.perf_counter() - t0
        launches = e.launches() - l0
        distinct, generated = r["distinct"], r["generated"]
        # end to end through the C ABI with host buffers: create + seed(H2D) + run + result(D2H) + destroy
        e.close()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            e2 = Engine(cm, deadlock=info["deadlock"], device=local_rank)
            e2.seed(init)
            r2 = e2.run()
            e2.close()
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t1
        assert r2["distinct"] == distinct
        stats = dict(kern_s=kern_s)
    else:
        from tla_rust_b200.dist import DistributedBFS
        fps = None   # ownership is computed from the packed words (tla_rust_b200.fingerprint.owner_of_words)

        e = Engine(cm, deadlock=info["deadlock"], device=local_rank)
        d = DistributedBFS(e, cm, rank, world, dev, cap_records=1 << 26, chunk_states=1 << 22)
        d.seed(init, fps)
        first = [True]

        def one():
            if not first[0]:
                e.restart()          # keeps the grown store / table / exchange buffers, re-seeds this rank's initial states
            first[0] = False
            l0 = e.launches()
            c0 = d.comm_ms
            out = d.run()
            return out, out["local"]["device_seconds"], e.launches() - l0, d.comm_ms - c0
        for _ in range(max(args.warmup, 1)):
            out, _, _, _ = one()
        assert (out["generated"], out["distinct"]) == (exp["o2"]["generated"], exp["o2"]["distinct"]), out
        if sampler:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        kern_s, launches, comm_ms = 0.0, 0, 0.0
        for _ in range(args.steps):
            out, ks, ln, cms = one()
            kern_s += ks
            launches += ln
            comm_ms += cms
        barrier()
        dt = time.perf_counter() - t0
        tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax.item())
        distinct, generated = out["distinct"], out["generated"]
        # end to end: fresh engine + exchange buffers + host-side seed every step (host buffers in, result out)
        e.close()
        del d
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            e2 = Engine(cm, deadlock=info["deadlock"], device=local_rank)
            d2 = DistributedBFS(e2, cm, rank, world, dev, cap_records=1 << 26, chunk_states=1 << 22)
            d2.seed(init, fps)
            out2 = d2.run()
            e2.close()
            del d2
        barrier()
        dt_e2e = time.perf_counter() - t1
        tmax = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt_e2e = float(tmax.item())
        assert out2["distinct"] == distinct
        stats = dict(kern_s=kern_s, comm_ms=comm_ms)

    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    if rank != 0:
        if multi:
            dist.destroy_process_group()
        return

    value = distinct * args.steps / dt
    # roofline of the dominant kernel (k_wave, fused expand + fingerprint + probe + compaction):
    # algorithmic bytes per step = S*expanded (frontier read) + 8*generated (slot probe) +
    #                              discovered*(8 slot write + S state write + 8 parent/meta)   [SURVEY §8d, fused form]
    bytes_step = S * distinct + 8 * generated + (distinct - len(np.unique(init, axis=0))) * (8 + S + 8)
    ach = bytes_step * args.steps / max(stats["kern_s"], 1e-9) / 1e9
    roof = {"kernel": "k_wave (fused expand+fingerprint+probe+compact)", "bound": "hbm", "achieved": round(ach, 3),
            "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 6), "traffic": None, "peak_source": peak_src,
            "bytes_per_step": int(bytes_step), "kernel_s_per_step": round(stats["kern_s"] / args.steps, 6),
            "note": "instruction-issue bound (ncu: issue-active 84%): the wave kernel interprets the Next/invariant "
                    "bytecode per state; per-launch DRAM traffic is in profiles/r1_traffic.json (frame spills dominate)"}
    k1 = None
    k1_launches = 0
    if not args.no_k1 and not multi:
        try:
            k1, k1_launches = k1_microbench(dev, peak)
            k1["peak_source"] = peak_src
        except Exception as ex:  # noqa: BLE001
            k1 = {"error": str(ex)}
    cpu_r, cpu_dt, _ = cpu_reference(cm, init, info, threads, exp["o2"]["distinct"])
    other = None
    if not multi and not args.no_k1:
        # second workload, reported next to the headline (not part of `value`): BASELINE config #4 (raft, 3 servers)
        # at the committed fixture's bounds -- container-typed state (W = 44 words), counts checked against the oracle.
        # Runs tools/fixture_bench.py in a child process with a time limit, so that it can never cost the headline.
        try:
            fx = os.path.join(ROOT, "tests", "golden", "MCraft_s3_l.tlagz")
            if os.path.exists(fx):
                p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fixture_bench.py"), "MCraft_s3_l", "--reps", "2"],
                                   capture_output=True, text=True, timeout=240,
                                   env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank))))
                r3 = json.loads(p.stdout.strip().splitlines()[-1])
                other = {"workload": "MCraft_s3_l: examples/raft.tla via models/MCraft.tla, 3 servers, MaxTerm 3, MaxLogLen 2, "
                                     "MaxMessages 4 (compiled fixture)", "W": r3["W"], "distinct": r3["distinct"],
                         "generated": r3["generated"], "depth": r3["depth"], "kernel_s": r3["device_s"],
                         "distinct_per_s": r3["distinct_per_s"], "counts_match_oracle": r3["counts_match_oracle"]}
        except Exception as ex:  # noqa: BLE001
            other = {"error": str(ex)[:300]}
    native = None
    if not multi and not args.no_k1 and args.workload == "MCPaxos3_b4":
        # Third, clearly labelled leg (not part of `value`): the same workload on the model-specialised native build of
        # the engine (compile/native.py: the program as straight-line CUDA instead of the bytecode interpreter), if
        # __graft_entry__.build() left its library in csrc/native/.  Child process with a time limit; counts and the
        # fingerprint digest are checked against the oracle record, and a failure only shows up here.
        try:
            from tla_rust_b200.engine import native_library_path
            if os.path.exists(native_library_path(cm)):
                p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fixture_bench.py"), "MCPaxos3_b4", "--native",
                                    "--reps", "2"], capture_output=True, text=True, timeout=300,
                                   env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank))))
                rn = json.loads(p.stdout.strip().splitlines()[-1])
                native = {"workload": "MCPaxos3_b4 on the model-specialised native build (experimental leg)",
                          "kernel_s": rn.get("device_s"), "distinct_per_s": rn.get("distinct_per_s"),
                          "counts_match_oracle": rn.get("counts_match_oracle"),
                          "digest_matches_oracle": rn.get("digest_matches_oracle"), "error": rn.get("error")}
        except Exception as ex:  # noqa: BLE001
            native = {"workload": "MCPaxos3_b4 on the model-specialised native build (experimental leg)",
                      "error": str(ex)[:300]}
    line = {"metric": "distinct states/sec", "value": round(value, 1), "unit": "states/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": cfg, "generated_per_s": round(generated * args.steps / dt, 1),
            "roofline": roof, "k1_roofline": k1,
            "cpu_baseline": {"value": round(cpu_r["distinct"] / cpu_dt, 1), "unit": "states/s", "cores": threads,
                             "kind": "port", "sample": cpu_r["sample"]},
            "e2e": {"value": round(distinct * args.steps / dt_e2e, 1), "unit": "states/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 96},
            "gpu_launches": int(launches), "clocks": sampler.summary() if sampler else None,
            "other_workloads": ([other] if other else []) + ([native] if native else [])}
    if multi:
        line["comm_ms_per_step"] = round(stats["comm_ms"] / args.steps, 3)
        dist.destroy_process_group()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
