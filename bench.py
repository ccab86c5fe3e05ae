#!/usr/bin/env python
"""bench.py -- headline benchmark of the bn-pp factor-algebra hot path on B200.

Metric (BASELINE.json): factor entries/sec of fused product+sum-out, measured on
config 4: exact PR by variable elimination (min-fill) on the synthetic "wide" Bayesian
network whose largest elimination step is a 2^28-entry fp64 union table.

One step = one full PR query: every bucket of the elimination order is one fused
product+sum-out launch, intermediates stay in HBM.  "entries" = sum over the
elimination steps of the union-table size.

  value : device-timed, CPTs already resident, plan prebuilt.
  e2e   : BN.partition(evidence, "mf") from HOST buffers each step: pinned CPTs -> H2D,
          min-fill ordering on the host, planning, the launches, scalar -> D2H.
  roofline : the widest fused launch, algorithmic bytes 8*(sum #operands + #out) over its
          CUDA-event duration inside the timed region, against MEASURED_PEAKS.json.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref) on a bounded sample of the same
          generator (narrower network), one core -- it is single-threaded.

N > 1 (torchrun): the network is too wide for one GPU by log2(N) variables; each rank
eliminates the slab of the wide factors selected by fixing those variables to its rank's
bits, and the cross-shard sum-out of the shard variables is one NCCL all-reduce of the
partition.  Per-GPU work stays fixed as N grows => "scaling": "weak".
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

METRIC = "factor entries/sec (fused product+sum-out), VE PR on 2^28-entry tables"
UNIT = "entries/s"

# config 4 per GPU count: (N, W, K, seed) of bnpp_b200.synth.random_bn_uai and the min-fill width
# (chosen with the host orderer so that fixing log2(N) shard variables leaves every rank a width-27 problem
# of ~2.2e9 union entries: 1 GPU 2.217e9, 2 GPUs 2.190e9, 4 GPUs 2.153e9, 8 GPUs 2.031e9 per rank)
WIDE = {1: (64, 40, 4, 5), 2: (68, 40, 4, 27), 4: (72, 40, 4, 23), 8: (76, 44, 4, 3)}
# bounded CPU sample: same generator, narrower (the reference needs ~1 us per entry)
CPU_SAMPLE = (48, 26, 4, 3)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons, sampled back to back from just before the timed
    regions until after them (B200_PROFILING.md recipe; one-shot queries: `-lms` output is
    block-buffered into a pipe and lost on terminate)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def _run_nvml(self):
        """in-process NVML: a sample every millisecond, so even a 15 ms timed region is covered"""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for nm, b in bits.items():
                if r & b:
                    self.reasons.add(nm)
            time.sleep(0.001)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(self.NAMES, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_leg(steps=1, procs=None):
    """times the unmodified reference (oracle/_ref/ref_harness) on the bounded sample.  The reference is
    single-threaded per query, so all host cores are used the only way it can use them: `procs`
    independent queries side by side (one process each); value = all their entries / wall time."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import concurrent.futures
    import oracle as orc
    from bnpp_b200 import model, synth
    N, W, K, seed = CPU_SAMPLE
    text = synth.random_bn_uai(N, W, K, seed)
    scopes, _ = synth.random_bn_scopes(N, W, K, seed)
    order, width = model.elim_order([2] * N, scopes, list(range(N)), "mf")
    entries = union_entries(scopes, order)
    path = "/tmp/bnpp_cpu_sample_%d.uai" % os.getpid()
    with open(path, "w") as f:
        f.write(text)
    if procs is None:
        procs = max(1, min(os.cpu_count() or 1, 32))
    if orc.have_ref():
        kind = "reference"

        def one(_):
            rows = orc.RefHarness().run(["model " + path, "opt mf", "pr"], timeout=3000)
            pr = [r for r in rows if r[0] == "PR"][0]
            return float(pr[1]), float(pr[2])
    else:
        kind = "port"
        procs = 1
        m = orc.read_uai(path)

        def one(_):
            t0 = time.perf_counter()
            z = orc.partition(m, {}, order)
            return z, (time.perf_counter() - t0) * 1e3
    walls, z, per_query_ms = [], None, []
    for _ in range(steps):
        t0 = time.perf_counter()
        with concurrent.futures.ThreadPoolExecutor(procs) as ex:     # each task is its own OS process (the harness)
            res = list(ex.map(one, range(procs)))
        walls.append(time.perf_counter() - t0)
        z = res[0][0]
        per_query_ms += [r[1] for r in res]
    os.unlink(path)
    wall = sum(walls) / len(walls)
    t_ms = wall * 1e3
    return {"value": entries * procs / wall, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": "same generator N=%d W=%d K=%d seed=%d (min-fill width %d, %d union entries per query); %d concurrent "
                      "single-threaded queries per step (the reference has no threading), %d step(s), %.1f s wall each, "
                      "%.1f s per query inside the reference; Z=%.12g"
                      % (N, W, K, seed, width, entries, procs, steps, wall, sum(per_query_ms) / len(per_query_ms) / 1e3, z)}, t_ms


def union_entries(scopes, order):
    """sum over elimination steps of the union-table size (binary variables), bucket elimination"""
    rank = {v: i for i, v in enumerate(order)}
    buckets = {v: [] for v in order}
    for sc in scopes:
        live = [v for v in sc if v in rank]
        if live:
            buckets[min(live, key=rank.get)].append(set(live))
    total = 0
    for v in order:
        if not buckets[v]:
            continue
        u = set().union(*buckets[v])
        total += 2 ** len(u)
        u.discard(v)
        if u:
            buckets[min(u, key=rank.get)].append(u)
    return total


def config5_cpu_leg(text, evs, procs=None, per_proc=2):
    """the reference's BN::partition (-mf) on the first evidence sets of config 5: `procs` harness processes side by
    side (the reference is single-threaded), `per_proc` queries each -> (queries/s, cores, kind, Z of set 0)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import concurrent.futures
    import oracle as orc
    if not orc.have_ref():
        return None
    if procs is None:
        procs = max(1, min(os.cpu_count() or 1, 32))
    path = "/tmp/bnpp_cpu_config5_%d.uai" % os.getpid()
    with open(path, "w") as f:
        f.write(text)

    def one(i):
        script = ["model " + path]
        for ev in evs[i * per_proc:(i + 1) * per_proc]:
            script += ["evidset %d %s" % (len(ev), " ".join("%d %d" % kv for kv in sorted(ev.items()))), "opt mf", "pr"]
        rows = orc.RefHarness().run(script, timeout=600)
        return [float(r[1]) for r in rows if r[0] == "PR"]
    t0 = time.perf_counter()
    with concurrent.futures.ThreadPoolExecutor(procs) as ex:
        res = list(ex.map(one, range(procs)))
    wall = time.perf_counter() - t0
    os.unlink(path)
    n = sum(len(r) for r in res)
    return {"value": n / wall, "unit": "queries/s", "cores": procs, "kind": "reference",
            "sample": "%d evidence sets of the same batch, %d concurrent single-threaded reference processes x %d queries "
                      "(model load included), %.2f s wall" % (n, procs, per_proc, wall)}, res[0][0]


def config5_leg(ctx, steps, warmup, cpu_arm=True):
    """The other half of BASELINE.json's metric -- VE PR queries/sec -- on config 5: 65 536 evidence sets on
    the 500-variable BN, PR per set, one GPU (tools/batch_bench.py is the multi-GPU version).  Device-timed
    with the evidence resident, and end to end from pinned host evidence (ordering, planning, H2D, the
    launch, D2H of every Z) with nothing cached."""
    import torch
    from bnpp_b200 import model, synth
    N, W, K, seed, nobs, nsets = 500, 6, 3, 11, 20, 65536
    text = synth.random_bn_uai(N, W, K, seed)
    _, bn = model.from_uai_text(ctx, text)
    evs = synth.evidence_batch(N, nobs, nsets, seed=5, fixed_ids=True)
    observed = sorted(evs[0])
    host = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8).pin_memory()
    dev = host.cuda()
    s = ctx.torch_stream
    for _ in range(max(3, warmup)):
        z = bn.partition_batch(observed, dev, "mf")
    ctx.sync()
    torch.cuda.synchronize()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(steps):
        z = bn.partition_batch(observed, dev, "mf")
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (ctx.launches - l0) // steps
    plan = list(bn._plans.values())[0]
    lanes, arena, n_steps = plan.fused_info(nsets)
    union_entries = plan.union_entries
    per_iter = []
    for _ in range(steps):
        t0 = time.perf_counter()
        bn.drop_plans()
        z = bn.partition_batch(observed, None, "mf", host_values=host)
        with torch.cuda.stream(s):
            zh = z.to("cpu", non_blocking=False)
        per_iter.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = sum(per_iter) / len(per_iter)
    bn.close()
    cpu = None
    try:
        got = config5_cpu_leg(text, evs) if cpu_arm else None
        if got:
            cpu, z_ref = got
            cpu["Z_set0_reference"], cpu["Z_set0_gpu"] = z_ref, float(zh[0])
    except Exception as e:
        cpu = {"error": "%s: %s" % (type(e).__name__, e)}
    return {"metric": "VE PR queries/sec", "value": nsets / ms * 1e3, "unit": "queries/s", "ms_per_batch": ms,
            "config": {"workload": "config 5: %d evidence sets, synthetic BN N=%d W=%d K=%d seed=%d, %d observed ids fixed, "
                                   "PR per set via VE (min-fill)" % (nsets, N, W, K, seed, nobs),
                       "l2_flush": "none: the launch reads 1.3 MB of evidence, 64 KB of CPTs and its 150 KB program and writes "
                                   "0.5 MB of results; the intermediates stay in shared memory, so this leg is instruction-bound, "
                                   "not HBM- or L2-bound (profiles/r1_fused_ncu.md)"},
            "e2e": {"value": nsets / e2e_ms * 1e3, "unit": "queries/s", "ms_per_batch": e2e_ms,
                    "h2d_bytes_per_step": host.numel(), "d2h_bytes_per_step": 8 * nsets},
            "gpu_launches_per_batch": launches, "union_entries_per_s": union_entries * nsets / ms * 1e3,
            "kernel": ("ve_fused: one launch, %d lanes per evidence set, %d doubles of shared memory per set, %d steps"
                       % (lanes, arena, n_steps)) if lanes else "contract_batched: one launch per bucket",
            "sample_Z": zh[:2].tolist(), "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the PR-queries/sec leg (config 5)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, t_ms = cpu_reference_leg(max(1, min(args.steps, 3)))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config 4 generator, bounded CPU sample: " + cb["sample"]},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from bnpp_b200 import capi, model, sharding, synth

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = capi.Context(local)
    stream = ctx.torch_stream
    shard_comm = None
    if world > 1:
        from bnpp_b200.nccl import ShardComm
        shard_comm = ShardComm(ctx, rank, world)     # the product's own NCCL communicator for the one collective of the path
        with torch.cuda.stream(stream):
            warm = torch.zeros(1, dtype=torch.float64, device="cuda")
        for _ in range(3):                           # NCCL builds its channels on the first collective: not part of a step
            shard_comm.allreduce_sum(warm.data_ptr(), 1)
        ctx.sync()

    n_gpus = world
    if n_gpus not in WIDE:
        raise SystemExit("bench.py supports 1, 2, 4 or 8 GPUs")
    N, W, K, seed = WIDE[n_gpus]
    text = synth.random_bn_uai(N, W, K, seed)
    _, bn = model.from_uai_text(ctx, text)
    full_order, full_width = bn.order(list(range(N)), {}, "mf")

    # wide-factor sharding: the log2(N) variables of the widest clique that are eliminated last
    shard_vars, evidence = [], {}
    if n_gpus > 1:
        g = n_gpus.bit_length() - 1
        shard_vars = sharding.pick_shard_vars(bn.scopes, full_order, g)
        evidence = sharding.shard_evidence(shard_vars, rank)
    variables = [v for v in range(N) if v not in evidence]

    def one_query_device(plan, obs_val, res):
        plan.run(bn.table_ptrs, obs_val, res.data_ptr(), res.data_ptr() + 8)

    # ---- device-timed value: plan prebuilt, tables resident --------------------------------
    order, width = bn.order(variables, evidence, "mf")
    plan = bn.plan(sorted(evidence), order)
    obs_val = [evidence[v] for v in sorted(evidence)]
    with torch.cuda.stream(stream):
        res = torch.zeros(2, dtype=torch.float64, device="cuda")
    for _ in range(args.warmup):
        one_query_device(plan, obs_val, res)
    ctx.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_query_device(plan, obs_val, res)     # the plan replays as one CUDA graph; no host sync inside the region
    if world > 1:
        with torch.cuda.stream(stream):
            zall = res[1:].clone()
        # cross-shard sum-out of the shard variables: one double over NVLink, through bnpp_shard_allreduce_sum
        shard_comm.allreduce_sum(zall.data_ptr(), 1)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = e0.elapsed_time(e1)
    gpu_launches = ctx.launches - launches0
    # second timed pass, same K steps, with a CUDA event pair around EVERY launch on the launching stream
    # (graph replay off): this is where the roofline's per-launch durations come from
    plan.set_profiling(True)
    per_launch = None
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        one_query_device(plan, obs_val, res)
        st = plan.step_stats()
        per_launch = st if per_launch is None else [dict(a, ms=a["ms"] + b["ms"]) for a, b in zip(per_launch, st)]
    p1.record(stream)
    torch.cuda.synchronize()
    profiled_ms = p0.elapsed_time(p1) / args.steps
    per_launch = [dict(a, ms=a["ms"] / args.steps) for a in per_launch]
    plan.set_profiling(False)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    z_local = float(res[1].item())
    z_total = float(zall.item()) if world > 1 else z_local
    entries_rank = plan.union_entries
    ent = torch.tensor([float(entries_rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ent)
    entries_all = float(ent.item())
    value = entries_all * args.steps / (dev_ms / 1e3)

    # ---- e2e: through the public API from host buffers -----------------------------------------
    for _ in range(2):
        bn.drop_plans()
        bn.reupload()
        bn.partition(evidence, "mf")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_parts = {}
    for _ in range(args.steps):
        bn.drop_plans()               # nothing cached: ordering + planning are paid every step
        bn.reupload()                   # pinned host CPTs -> HBM
        z_e2e, _ = bn.partition(evidence, "mf")     # ... launches ... scalar -> host
        for kk, vv in bn.last_timing.items():
            e2e_parts[kk] = e2e_parts.get(kk, 0.0) + vv / args.steps
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = entries_all * args.steps / e2e_s
    clocks = sampler.summary() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: the widest fused launch ---------------------------------
    peak, peak_kind = peaks()
    widest = max(range(len(per_launch)), key=lambda i: per_launch[i]["bytes"])
    wl = per_launch[widest]
    w_ms = wl["ms"]
    achieved = wl["bytes"] / w_ms / 1e6
    big = [s for s in per_launch if s["entries"] >= (1 << 24)]
    big_bytes = sum(s["bytes"] for s in big)
    big_ms = sum(s["ms"] for s in big)
    all_ms = sum(s["ms"] for s in per_launch)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if n_gpus == 1 and os.path.exists(tp):
        tj = json.load(open(tp))      # dram__bytes_read.sum + dram__bytes_write.sum of this launch, one `ncu --set full` capture
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes": wl["bytes"], "peak_kind": peak_kind,
                "measured_in": "second timed pass of the same K steps with a CUDA event pair around every launch "
                               "(%.3f ms per step with events vs %.3f ms in the graph-replayed pass `value` is taken from)"
                               % (profiled_ms, dev_ms / args.steps),
                "kernel": "contract_fast (fused product+sum-out), widest launch: k=%d operands, %d union entries, "
                          "%.3f GB algorithmic, %.3f ms" % (wl["k"], wl["entries"], wl["bytes"] / 1e9, w_ms),
                "launches_ge_2p24_entries": {"n": len(big), "GBs": big_bytes / big_ms / 1e6 if big_ms else None,
                                             "frac": big_bytes / big_ms / 1e6 / peak if big_ms else None,
                                             "share_of_step_ms": big_ms / all_ms if all_ms else None}}

    config5 = None
    if n_gpus == 1 and not args.no_config5:
        try:
            config5 = config5_leg(ctx, args.steps, args.warmup, cpu_arm=not args.no_cpu_baseline)
        except Exception as e:      # the headline line must not depend on the auxiliary leg
            config5 = {"error": "%s: %s" % (type(e).__name__, e)}

    cb = None
    if not args.no_cpu_baseline:
        cb, _ = cpu_reference_leg(1)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 4: synthetic random BN N=%d W=%d K=%d seed=%d, exact PR via VE (min-fill width %d%s); "
                                   "%d fused launches per query, %d union entries per rank, largest union table 2^%d entries"
                                   % (N, W, K, seed, full_width,
                                      (", sharded on variables %s -> width %d per rank" % (shard_vars, width)) if shard_vars else "",
                                      plan.n_launches, entries_rank, plan.max_step_entries.bit_length() - 1),
                       "l2_flush": "inputs larger than L2: every step streams %.1f GB of tables through a 126 MB L2"
                                   % (plan.bytes / 1e9),
                       "partition": z_total, "partition_e2e": z_e2e, "peak_intermediate_GB": plan.peak_bytes / 1e9},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bn.h2d_bytes, "d2h_bytes_per_step": 16,
                    "ms_per_step": e2e_s / args.steps * 1e3, "host_breakdown_ms": e2e_parts},
            "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cb,
            "pr_queries": config5}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
