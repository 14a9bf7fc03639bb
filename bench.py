#!/usr/bin/env python
"""bench.py -- (T) wall-time and FP64 GFLOP/s of the CCSD(T) triples hot path on N B200s.

Default workload (config.workload) = BASELINE.json's target shape, (H2O)10/aug-cc-pVTZ (configs[4]; 40 occupied / 870
virtual alpha orbitals, tilesize 40, 7 590 tile tuples, 4.5e17 FLOP): the block stores are generated ON THE DEVICE
(keyed hash, no store ever on the host), V2 in the reference's spin-free `2eorb` form (105 GB; the spin-orbital form
would be 527 GB), resident whole on one GPU and sharded over the N GPUs otherwise, remote blocks pulled over NVLink.
A step = one pass of the hot path over the SAME fixed sample of the heaviest-first task list at every N (--tasks M:
tasks floor(i*7590/M), i < M -- diagonal, off-diagonal, 40- and 39-wide tiles in the list's proportions; the whole list
does not fit the per-N time limit: 4.5e17 FLOP ~ 3.7 h on one GPU), partitioned over the ranks in equal-cost contiguous
blocks (strong scaling), one ncclAllReduce of the two energies per step (ga_dgop, ccsd_t.F:297).

  value : whole-job GFLOP/s (algorithmic FLOPs of SURVEY 8d / device time, CUDA events on the library's stream, max
          over ranks), stores resident in HBM (Tier 2 / native API)
  e2e   : the same metric through the reference-facing call surface (Tier 1: host block stores, host TCE_SORT_4,
          sd_t_*_cuda_ calls with HOST operands, H2D inside the timed region, D2H of the energies), in the reference's
          own contract (pageable operands reused right after each call); e2e.optin = with the pinned/async opt-in
  energy_check : the N-rank NCCL-reduced energies against the N=1 energies of the same prefix (tests/golden/
          bench_energies.json, written by a 1-GPU run), |dE| <= 1e-9 Eh
  parity (N=1): GPU vs the CPU oracle on a p4 slab of task 0, inside the cpu_baseline leg

Other workloads: --workload microbench_t40 (BASELINE configs[1], round 1's bench), uracil_augccpvdz (configs[2], whole
list), benzene_dimer_augccpvtz (configs[3]).
--impl reference: the reference's CPU implementation of the path (the oracle port of ccsd_t_kernels_omp.F; the Fortran
cannot be compiled in this image) on all host cores, each step a bounded sample of the same workload.
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
import numpy as np  # noqa: E402

SEED = 20240229
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_energies.json")
# per workload: storage, default task prefix (0 = whole list), generator scales (t1, t2, v2) chosen so |E| = O(1) Eh
WORKLOADS = {
    "h2o10_augccpvtz": dict(intorb=True, tasks=8, scale=(1e-3, 5e-5, 5e-3), device_gen=True,
                            what="(H2O)10/aug-cc-pVTZ-shaped: alpha occ/virt 40/870, tilesize 40 (2+44 tiles, 7590 tuples)"),
    "benzene_dimer_augccpvtz": dict(intorb=True, tasks=12, scale=(1e-3, 5e-5, 5e-3), device_gen=True,
                                    what="benzene-dimer/aug-cc-pVTZ-shaped: alpha occ/virt 30/786, tilesize 40 (2+40 tiles, 5740 tuples)"),
    "uracil_augccpvdz": dict(intorb=True, tasks=0, scale=(1e-3, 1e-4, 1e-2), device_gen=True,
                             what="uracil/aug-cc-pVDZ-shaped: alpha occ/virt 21/191, tilesize 40 (2+10 tiles, 110 tuples)"),
    "microbench_t40": dict(intorb=False, tasks=0, scale=(0.05, 0.02, 0.1), device_gen=False,
                           what="microbench: o=v=40 alpha orbitals, tilesize 40, random T1/T2/V2 tiles (2 tuples)"),
}


class ClockSampler:
    def __init__(self, dev):
        self.dev, self.rows, self.p = dev, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------------------------
# CPU arm (--impl reference): kernels of ccsd_t_kernels_omp.F (oracle port) on a bounded sample of the workload
# --------------------------------------------------------------------------------------------------------------------
def task0_ranges(workload):
    """(h3d,h2d,h1d,p6d,p5d,p4d) of the heaviest task and the contracted ranges (h7, p7) it meets."""
    from nwchem_b200 import synth
    t = synth.shape_tiling(workload)
    occ = max(int(r) for r in t.range[:t.noab]); virt = max(int(r) for r in t.range[t.noab:])
    return (occ, occ, occ, virt, virt, virt), occ, virt


def cpu_kernel_sample(workload, steps, warmup, target_s=3.0):
    """Each step = all 27 kernels (9 sd_t_d2 + 9 sd_t_d1 + 9 sd_t_s1, one contracted tile each) on a p4 slab of the
    heaviest task's t3 tile, slab width calibrated to ~target_s seconds per step.  All host threads."""
    from oracle import oracle as ora
    ora.lib()
    ora.set_num_threads(host_cores())
    (h3d, h2d, h1d, p6d, p5d, p4d), kh, kp = task0_ranges(workload)
    rng = np.random.default_rng(SEED)
    # operands in task-tuple names; for permuted kernels the ranges are equal-sized particle (hole) tiles, so one set
    # of operand arrays of the largest shape serves all nine k of a family
    t2_d2 = rng.uniform(-1, 1, kp * p4d * h1d * h2d); v2_d2 = rng.uniform(-1, 1, kp * h3d * p6d * p5d)
    t2_d1 = rng.uniform(-1, 1, kh * p4d * p5d * h1d); v2_d1 = rng.uniform(-1, 1, h3d * h2d * p6d * kh)
    t1 = rng.uniform(-1, 1, p4d * h1d); v2_s = rng.uniform(-1, 1, h3d * h2d * p6d * p5d)
    dims = (h3d, h2d, h1d, p6d, p5d, p4d)

    def one_step(w):
        n = h3d * h2d * h1d * p6d * p5d * w
        t3d = np.zeros(n); t3s = np.zeros(n)
        fl = 0.0
        t0 = time.perf_counter()
        for k in range(1, 10):
            ora.kernel_slab(2, k, dims, kp, 0, w, t3d, t2_d2, v2_d2); fl += 2.0 * n * kp
            ora.kernel_slab(1, k, dims, kh, 0, w, t3d, t2_d1, v2_d1); fl += 2.0 * n * kh
            ora.kernel_slab(0, k, dims, 1, 0, w, t3s, t1, v2_s); fl += 2.0 * n
        return time.perf_counter() - t0, fl

    dt1, _ = one_step(1)
    w = int(max(1, min(p4d, round(target_s / max(dt1, 1e-3)))))
    for _ in range(warmup):
        one_step(w)
    tot = 0.0; fl = 0.0
    for _ in range(steps):
        dt, f = one_step(w)
        tot += dt; fl += f
    return dict(seconds=tot, flops=fl, gflops=fl / tot * 1e-9, cores=ora.num_threads(), steps=steps,
                sample=f"per step: the 27 CPU kernels (sd_t_d2_1..9 with K={kp}, sd_t_d1_1..9 with K={kh}, sd_t_s1_1..9) on a "
                       f"p4 slab of {w}/{p4d} of the heaviest {workload} tuple ({fl / steps:.3e} FLOP, {tot / steps:.2f} s), "
                       f"OpenMP on {ora.num_threads()} host threads; kernels only (no fetch/sort/energy)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="h2o10_augccpvtz", choices=sorted(WORKLOADS))
    ap.add_argument("--tasks", type=int, default=-1, help="size M of the strided sample of the heaviest-first task list run per step (0 = whole list)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--replicated", action="store_true", help="N > 1: keep the whole V2 store on every GPU instead of sharding it")
    ap.add_argument("--weak", action="store_true", help="N > 1: every rank runs its own copy of the prefix (round 1's replica mode)")
    ap.add_argument("--write-golden", action="store_true", help="N = 1: record the energies of this prefix in tests/golden/bench_energies.json")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = WORKLOADS[a.workload]
    ntasks_req = W["tasks"] if a.tasks < 0 else a.tasks

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_kernel_sample(a.workload, a.steps, a.warmup)
        cfg = {"workload": f"{a.workload}: {W['what']}", "tilesize": 40}
        line = {"impl": "reference", "metric": "(T) FP64 GFLOP/s", "value": r["gflops"], "unit": "GFLOP/s", "n_gpus": a.gpus,
                "steps": r["steps"], "warmup": a.warmup, "ms_per_step": r["seconds"] / r["steps"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": r["gflops"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU arm runs on rank 0 only with all host threads; it does not scale with --gpus"}
        print(json.dumps(line))
        return

    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("NWC_BENCH_WATCHDOG_S", "1500")), exit=True)   # never hang a GPU box
    import torch
    import torch.distributed as dist
    from nwchem_b200 import capi, synth, tiling as tl
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncores = host_cores()
    capi.set_host_threads(max(1, ncores // max(1, world)))   # torchrun exports OMP_NUM_THREADS=1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(x):
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    def allmax(x):
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    # ---- roofline denominator, measured here and now ----
    peak = capi.fp64_peak_probe(local)
    peak_how = "DMMA.8x8x4 register loop measured in this run (nwc_fp64_peak_probe)"

    # ---- state: tables on the host, stores generated on the device ----
    t = synth.shape_tiling(a.workload)
    sharded = world > 1 and not a.replicated and not a.weak
    tr = capi.Triples(local)
    t_setup = time.perf_counter()
    if W["device_gen"]:
        st = synth.empty_stores(t, intorb=W["intorb"])
        if W["intorb"]:
            tr.set_state_2eorb(st, rank if sharded else 0, world if sharded else 1)
        elif sharded:
            tr.set_state_sharded(st, rank, world)
        else:
            tr.set_state(st)
        tr.synth_fill(SEED, W["scale"])
    else:
        st = synth.random_blocks(t, seed=SEED)
        if sharded:
            tr.set_state_sharded(synth.shard_v2(st, rank, world), rank, world)
        else:
            tr.set_state(st)
    if sharded:
        mine = torch.tensor(list(tr.v2_ipc_handle()), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        tr.v2_open_peers(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    if world > 1:  # the library's own communicator: rank 0's id is broadcast with the torch plumbing
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.Triples.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        tr.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    tr.set_batch_bytes(6 << 30)
    ntot = tr.num_tasks
    ntasks = ntot if ntasks_req <= 0 else min(ntasks_req, ntot)
    ids = np.array([(i * ntot) // ntasks for i in range(ntasks)], np.int64)   # strided sample, the same at every N
    tasks = tr.task_list()[ids]
    setup_s = time.perf_counter() - t_setup
    resident_gb = tr.stats()["resident_bytes"] * 1e-9

    def step():
        if a.weak or world == 1:
            e1, e2 = tr.run_partition_list(0, 1, ids)
        else:
            e1, e2 = tr.run_partition_list(rank, world, ids)
        if world > 1 and not a.weak:
            e1, e2 = tr.allreduce(e1, e2)   # replaces ga_dgop (ccsd_t.F:297)
        return e1, e2

    e_first = None
    for _ in range(a.warmup):
        e_first = step()
    tr.set_timing(True)
    tr.stats(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    tr.timer_start()
    for _ in range(a.steps):
        e = step()
    ms = tr.timer_stop_ms()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tr.set_timing(False)
    mem_free = {"after_timed_loop_GB": torch.cuda.mem_get_info()[0] * 1e-9}
    s = tr.stats()
    ms_max, flops_all = allmax(ms), allsum(s["flops"])
    # per-rank device time of the fused kernel and of the whole timed region (load balance of the static partition)
    per_rank = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
    per_rank[2 * rank] = s["fused_ms"] / a.steps
    per_rank[2 * rank + 1] = ms / a.steps
    if world > 1:
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    per_rank = [round(float(x), 2) for x in per_rank.tolist()]
    value = flops_all / (ms_max * 1e-3) * 1e-9
    launches = int(s["fused_launches"] + s["repack_launches"] + s["reduce_launches"] + s["pull_launches"] + s["antisym_launches"])
    fused_avg_ms = s["fused_ms"] / max(1, s["fused_launches"])
    achieved = (s["flops"] / max(1, s["fused_launches"])) / (fused_avg_ms * 1e-3) * 1e-12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_fused_r02.json"))).get(a.workload, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "fused_kernel (DMMA.8x8x4 + UBLKCP)", "peak_source": peak_how,
                "flops_per_launch": s["flops"] / max(1, s["fused_launches"]), "avg_launch_ms": fused_avg_ms,
                "fused_launches": int(s["fused_launches"]), "fused_share_of_step": s["fused_ms"] / ms,
                "repack_share_of_step": s["repack_ms"] / ms, "pull_share_of_step": s["pull_ms"] / ms}
    nvlink = None
    if sharded:
        pb, pm = allsum(s["peer_bytes"]), allmax(s["pull_ms"])
        nvlink = {"pulled_GB_per_step_all_ranks": pb * 1e-9 / a.steps,
                  "rank0_GBps_in_pull_kernel": (s["peer_bytes"] / (s["pull_ms"] * 1e-3) * 1e-9) if s["pull_ms"] > 0 else None,
                  "pull_ms_per_step_max_rank": pm / a.steps}

    # ---- energies: reproducible step to step, and equal to the N=1 energies of the same prefix ----
    key = f"{a.workload}:sample={ntasks}of{ntot}:seed={SEED}"
    golden = {}
    try:
        golden = json.load(open(GOLDEN))
    except Exception:
        pass
    energy_check = {"steps_bitwise_equal": bool(e_first is None or tuple(e_first) == tuple(e)), "key": key}
    if a.weak:
        energy_check["note"] = "weak mode: ranks run replicas, nothing to reduce"
    elif key in golden:
        ref = golden[key]["energy"]
        dE = [e[0] - ref[0], e[1] - ref[1]]
        energy_check.update({"n1_energy": ref, "dE": dE, "ok": bool(abs(dE[0]) <= 1e-9 and abs(dE[1]) <= 1e-9),
                             "tol_Eh": 1e-9, "source": "N=1 run recorded in tests/golden/bench_energies.json"})
        if world > 1 and not energy_check["ok"] and rank == 0:
            print(f"ENERGY MISMATCH vs N=1: {dE}", file=sys.stderr)
    else:
        energy_check["ok"] = None
        energy_check["note"] = "no N=1 record for this prefix (run with --gpus 1 --write-golden)"
    if a.write_golden and world == 1 and rank == 0:
        golden[key] = {"energy": [e[0], e[1]], "flops_per_step": flops_all / a.steps, "when": time.strftime("%Y-%m-%d")}
        os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
        json.dump(golden, open(GOLDEN, "w"), indent=1, sort_keys=True)
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump(golden, open(os.path.join(ROOT, "gpurun_out", "bench_energies.json"), "w"), indent=1, sort_keys=True)
        except Exception:
            pass

    # ---- host copy of exactly the blocks this rank's Tier-1 tasks read (exported from the device) ----
    need_host = (not a.no_e2e) or (rank == 0 and world == 1 and not a.no_cpu_baseline)
    my_tasks = tasks[rank::world] if world > 1 else tasks
    tr.trim()   # Tier 1 has its own engine: give Tier 2's batch arenas back first (130 GB of stores are resident)
    if need_host and W["device_gen"]:   # bound the host copy by the memory the box has (~8 GB per (H2O)10 tuple)
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 64 << 30
        fit = max(1, int(0.5 * avail / max(1, world) / 9e9))
        if fit < len(my_tasks):
            my_tasks = my_tasks[:fit]
    host = None
    if need_host and len(my_tasks) and W["device_gen"]:
        host = sparse_host_store(tr, t, st, my_tasks, capi, synth, tl)
    elif need_host and len(my_tasks):
        host = st

    # ---- e2e through the reference-facing Tier-1 surface (host buffers) ----
    e2e = None
    if not a.no_e2e:
        import ctypes
        capi.lib().nwc_triples_set_local_rank(ctypes.c_long(local))
        nrep = 1 if len(my_tasks) >= 6 else max(1, min(a.steps, 2))

        def tier1(reference_contract):
            capi.set_reference_contract(reference_contract)
            pinned = []
            if not reference_contract and host is not None:
                for arr in (host.t2, host.v2):
                    try:
                        arr.setflags(write=True); capi.host_register(arr); pinned.append(arr)
                    except Exception:
                        pass
            if host is not None:
                capi.ccsd_t_gpu_tasks(host, my_tasks[:1], icuda=max(world, local + 1))   # warm the pools
            capi.compat_stats(reset=True)
            barrier()
            t0 = time.perf_counter()
            c = (0.0, 0.0)
            for _ in range(nrep):
                if host is not None:
                    c = capi.ccsd_t_gpu_tasks(host, my_tasks, icuda=max(world, local + 1))[:2]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            cs = capi.compat_stats()
            for arr in pinned:
                capi.host_unregister(arr)
            capi.set_reference_contract(False)
            dtm, fl = allmax(dt), allsum(cs["flops"])
            return {"value": fl / dtm * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": allsum(cs["h2d_bytes"]) / nrep,
                    "d2h_bytes_per_step": allsum(cs["d2h_bytes"]) / nrep, "ms_per_step": dtm / nrep * 1e3, "reps": nrep,
                    "energy": [allsum(c[0]), allsum(c[1])]}
        d = tier1(True)
        o = tier1(False)
        e2e = dict(d)
        e2e["api"] = ("nwc_ccsd_t_gpu_tasks -> sd_t_*_cuda_/compute_en_ (Tier 1), host block stores, host TCE_SORT_4 included; "
                      "reference contract: pageable operands, refilled right after each call (ccsd_t_doubles_gpu.F:282-327,723-726); "
                      "tasks dealt whole to ranks (the reference's granularity)")
        n_e2e = int(allsum(len(my_tasks)))
        e2e["tasks"] = n_e2e
        e2e["energy_matches_native"] = bool(abs(d["energy"][0] - e[0]) <= 1e-9 * max(1.0, abs(e[0])) and
                                            abs(d["energy"][1] - e[1]) <= 1e-9 * max(1.0, abs(e[1]))) if (not a.weak and n_e2e == ntasks) else None
        e2e["optin"] = {k: o[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")}
        e2e["optin"]["contract"] = "nwc_compat_set_async_uploads(1) + pinned host stores (operands untouched until compute_en_)"

    if not a.no_e2e:
        mem_free["after_e2e_GB"] = torch.cuda.mem_get_info()[0] * 1e-9
        capi.compat_trim()
    # ---- CPU baseline + GPU-vs-oracle parity on a p4 slab of task 0 (N = 1 only) ----
    cpu = None; parity = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and host is not None:
        from oracle import oracle as ora
        ora.lib()
        ora.set_num_threads(ncores)
        tup = [int(x) for x in tasks[0][:6]]
        width = 4
        t0 = time.perf_counter()
        o1, o2 = ora.tuple_slab(host, tup, 0, width)
        dt = time.perf_counter() - t0
        items = tr.tuple_items(tup)
        nb4 = (t.r(tup[0]) + 3) // 4
        tr.stats(reset=True)
        g1, g2 = tr.run_items(tup, 0, items // nb4)
        fl = tr.stats()["flops"]
        cpu = {"value": fl / dt * 1e-9, "unit": "GFLOP/s", "cores": ora.num_threads(), "kind": "port", "seconds": dt,
               "sample": f"task 0 of the list {tup}, p4 slab {width}/{t.r(tup[0])} of its t3 tile, whole path: GET_HASH_BLOCK + "
                         f"TCE_SORT_4 + every fired sd_t_s1/d1/d2 kernel + ccsd_t_dot ({fl:.3e} FLOP), OpenMP on all host threads"}
        parity = {"slab": f"task 0, p4 in [0,{width})", "oracle": [o1, o2], "gpu": [g1, g2], "dE1": g1 - o1, "dE2": g2 - o2,
                  "rel1": abs(g1 - o1) / max(abs(o1), 1e-300), "rel2": abs(g2 - o2) / max(abs(o2), 1e-300)}

    if rank == 0:
        cfg = {"workload": f"{a.workload}: {W['what']}; " +
                           (f"strided sample of {ntasks} of the {ntot} tuples of the heaviest-first list per step (tasks floor(i*{ntot}/{ntasks}))" if ntasks < ntot else f"all {ntot} tuples per step"),
               "tilesize": 40, "tasks_per_step": int(ntasks),
               "v2": ("2eorb (spin-free orbital form, antisymmetrised on the device per batch)" if W["intorb"] else "spin-orbital blocks") +
                     (f", sharded over {world} GPUs, remote blocks pulled over NVLink" if sharded else ", whole store resident on every GPU"),
               "stores": "generated on the device (keyed hash)" if W["device_gen"] else "host random blocks uploaded once",
               "resident_GB_per_gpu": resident_gb, "setup_s": setup_s, "panel_index_order": tr.order,
               "partition": "weak: replicas" if a.weak else "strong: equal-cost contiguous blocks of the task prefix, boundary tuples split at sub-tile granularity",
               "l2": "operand panels per tuple (>= 1.4 GB) exceed the 126 MB L2; no explicit flush"}
        line = {"metric": "(T) FP64 GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
                "scaling": "weak" if a.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "wall_s_per_step": ms_max / a.steps * 1e-3, "flops_per_step": flops_all / a.steps,
                "frac_of_fp64_peak": value * 1e-3 / (peak * world), "energy": list(e), "energy_check": energy_check,
                "full_list_extrapolated_s": (4.5208e17 / (value * 1e9)) if a.workload == "h2o10_augccpvtz" else None,
                "rank_fused_ms_per_step": per_rank[0::2], "rank_ms_per_step": per_rank[1::2],
                "roofline": roofline, "nvlink": nvlink, "cpu_baseline": cpu, "parity": parity, "e2e": e2e,
                "gpu_launches": launches, "gpu_mem_free": mem_free, "clocks": clocks}
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


def sparse_host_store(tr, t, st, my_tasks, capi, synth, tl):
    """Host block stores holding exactly the T1/T2/V2 blocks `my_tasks` read, copied out of the device stores (V2
    blocks in the spin-orbital form the Fortran passes: antisymmetrised on the device from the 2eorb store)."""
    full_v2h, _ = None, None
    base = synth.BlockStores(t, st.t1_hash, None, st.t2_hash, None, np.zeros(1, np.int64), None)

    def table(keys, sizes):
        n = len(keys)
        h = np.zeros(2 * n + 1, np.int64)
        h[0] = n; h[1:n + 1] = keys
        off = np.zeros(n, np.int64)
        if n:
            off[1:] = np.cumsum(np.array(sizes, np.int64))[:-1]
        h[n + 1:] = off
        return h, off, int(sum(sizes))

    def full_offsets(h):
        n = int(h[0])
        return {int(h[1 + i]): int(h[1 + n + i]) for i in range(n)}

    # T1 whole (tiny), T2 and V2 sparse
    n1 = tl.t1_offset(t)[1]
    t1 = tr.debug_read(1, 0, n1)
    k2 = capi.host_collect_blocks(base, my_tasks, 2)
    sz2 = [int(np.prod([t.r(b) for b in tl.decode_t2_key(t, int(k))])) for k in k2]
    t2h, off2, tot2 = table(k2, sz2)
    f2 = full_offsets(st.t2_hash)
    t2 = np.empty(tot2)
    for k, o, n in zip(k2, off2, sz2):
        t2[o:o + n] = tr.debug_read(2, f2[int(k)], n)
    k3 = capi.host_collect_blocks(base, my_tasks, 3)
    sz3 = [int(np.prod([t.r(b) for b in tl.decode_v2_key(t, int(k))])) for k in k3]
    v2h, off3, tot3 = table(k3, sz3)
    v2 = np.empty(tot3)
    for k, o, n in zip(k3, off3, sz3):
        v2[o:o + n] = tr.export_v2_block(*tl.decode_v2_key(t, int(k)))
    return synth.BlockStores(t, st.t1_hash, t1, t2h, t2, v2h, v2)


if __name__ == "__main__":
    main()
