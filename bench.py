#!/usr/bin/env python
"""bench.py -- (T) wall-time and FP64 GFLOP/s of the CCSD(T) triples hot path on N B200s.

Workload (config.workload): BASELINE.json configs[1], the synthetic (T) kernel microbench: o = v = 40
spatial orbitals, tilesize 40, random T1/T2/V2 tiles, RHF-restricted -> 2 tile tuples, 34 contraction
groups, 1.1256e13 algorithmic FP64 FLOPs per step.  A step = one pass of the hot path over that task list.

  value : whole-job GFLOP/s with T1/T2/V2 resident in HBM (Tier 2 / native API), device-event timed
  e2e   : the same metric through the reference-facing call surface (Tier 1: host block stores, host
          TCE_SORT, per-call H2D of the sorted operands, D2H of the energies) -- the headline vs the CPU arm
  N > 1 : weak scaling -- every rank runs its own copy of the task list (independent tuples, inputs
          replicated per GPU), one ncclAllReduce of the two energies per step replaces ga_dgop.

--impl reference: the reference's CPU implementation of the path (the oracle port of ccsd_t_kernels_omp.F +
ccsd_t_dot.F; the Fortran cannot be compiled in this image) on the host cores, on a bounded sample.
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

WORKLOAD = "microbench_t40"
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")


def fp64_peak_tflops():
    """Roofline denominator: MEASURED_PEAKS.json carries no FP64 figure, so the measured DMMA.8x8x4 rate of
    tools/fp64_peak.cu on this pool's B200 (profiles/fp64_peak_r01.json) is used; cuBLAS DGEMM is beside it."""
    try:
        d = json.load(open(FP64_PEAK_FILE))
        return max(v for k, v in d.items() if k.startswith("dmma_")), "measured DMMA.8x8x4 loop (profiles/fp64_peak_r01.json)"
    except Exception:
        return 37.1, "fallback 37.1 (measured DMMA on this pool, file missing)"


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


def _cpu_slab(P4, steps):
    from oracle import oracle as ora
    rng = np.random.default_rng(20240229)
    T = 40
    dims = (T, T, T, T, T, P4)  # h3d,h2d,h1d,p6d,p5d,p4d
    n = T ** 5 * P4
    t3d = np.zeros(n); t3s = np.zeros(n)
    t2a = rng.uniform(-1, 1, T * P4 * T * T); v2a = rng.uniform(-1, 1, T ** 4)
    t2b = rng.uniform(-1, 1, T * P4 * T * T); v2b = rng.uniform(-1, 1, T ** 4)
    t1 = rng.uniform(-1, 1, P4 * T); v2s = rng.uniform(-1, 1, T ** 4)
    flops = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        for k in range(1, 10):
            # kernels 4-9 permute the particle ranges; with p4d != p5d = p6d only k<=3 keep the slab shape,
            # so the sample cycles the three hole permutations over the same slab (same FLOPs per call)
            kk = (k - 1) % 3 + 1
            ora.kernel(2, kk, dims, T, t3d, t2a, v2a); flops += 2.0 * n * T
            ora.kernel(1, kk, dims, T, t3d, t2b, v2b); flops += 2.0 * n * T
            ora.kernel(0, kk, dims, 1, t3s, t1, v2s); flops += 2.0 * n
    return time.perf_counter() - t0, flops


def cpu_sample(steps=1, target_s=15.0):
    """Bounded CPU sample of the same workload: tuple 1 of the microbench restricted to a p4 slab
    (all 9 sd_t_d2 + 9 sd_t_d1 + 9 sd_t_s1 kernels of ccsd_t_kernels_omp.F restated), all host threads.
    The slab width is calibrated so that one step is about `target_s` seconds of CPU work."""
    from oracle import oracle as ora
    ora.lib()
    dt, fl = _cpu_slab(1, 1)                      # calibration: slab of 1
    P4 = int(max(1, min(12, round(target_s / max(dt, 1e-3)))))
    dt, fl = _cpu_slab(P4, steps)
    return dict(seconds=dt, flops=fl, gflops=fl / dt * 1e-9, cores=ora.num_threads(),
                sample=f"tuple 1 of {WORKLOAD} restricted to a p4 slab of {P4}/40: 9 sd_t_d2 + 9 sd_t_d1 + 9 sd_t_s1 "
                       f"kernel calls at tilesize 40 ({fl / steps:.3e} FLOP per step, {dt / steps:.1f} s), OpenMP on all host threads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--strong", action="store_true", help="one task list dealt over the ranks (fixed total work) instead of one copy per rank")
    ap.add_argument("--sharded", action="store_true", help="shard the V2 store over the ranks; remote blocks are read over NVLink (CUDA IPC)")
    ap.add_argument("--intorb", action="store_true", help="V2 in the reference's spin-free `2eorb` form, antisymmetrised on the device")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    peak, peak_how = fp64_peak_tflops()
    if a.workload == WORKLOAD:
        cfg = {"workload": f"{a.workload}: o=v=40 alpha orbitals, tilesize 40, random T1/T2/V2 tiles, RHF-restricted, 2 tile tuples",
               "tilesize": 40, "tuples_per_step_per_gpu": 2,
               "l2": "operand panels touched per step (1.4 GB) exceed the 126 MB L2; no explicit flush"}
    else:
        from nwchem_b200 import synth as _s
        sh = _s.SHAPES[a.workload]
        cfg = {"workload": f"{a.workload}: alpha occ/virt {sh['occ']}/{sh['virt']}, tilesize {sh['tilesize']}, random tiles, RHF-restricted",
               "tilesize": sh["tilesize"], "v2": "sharded over ranks, NVLink peer reads" if a.sharded else "replicated",
               "partition": "strong: heaviest-first task list dealt round-robin" if a.strong else "weak: one task list per rank",
               "l2": "operand panels per tuple exceed the 126 MB L2; no explicit flush"}
    if a.intorb:
        cfg["v2"] = "2eorb: spin-free orbital-form store resident, blocks antisymmetrised on the device per batch"

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_sample(max(1, min(a.steps, 3)), target_s=12.0)
        line = {"impl": "reference", "metric": "(T) FP64 GFLOP/s", "value": r["gflops"], "unit": "GFLOP/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["seconds"] / max(1, min(a.steps, 3)) * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": r["gflops"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("NWC_BENCH_WATCHDOG_S", "900")), exit=True)   # never hang a GPU box
    import torch
    import torch.distributed as dist
    from nwchem_b200 import capi, synth
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t = synth.shape_tiling(a.workload)
    st = synth.random_blocks(t, seed=20240229 + (0 if (a.strong or a.sharded) else rank))
    tr = capi.Triples(local)
    if a.sharded and world > 1:
        tr.set_state_sharded(synth.shard_v2(st, rank, world), rank, world)
        mine = torch.tensor(list(tr.v2_ipc_handle()), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        tr.v2_open_peers(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    elif a.intorb:
        st.orb = synth.random_orbital(t, seed=20240229 + (0 if a.strong else rank))
        tr.set_state_2eorb(st)
    else:
        tr.set_state(st)
    if world > 1:  # the library's own communicator: rank 0's id is broadcast with the torch plumbing
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.Triples.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        tr.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def step():
        e1, e2 = tr.run(first=rank, stride=world) if a.strong else tr.run()
        if world > 1:
            e1, e2 = tr.allreduce(e1, e2)   # replaces ga_dgop (ccsd_t.F:297)
        return e1, e2

    for _ in range(a.warmup):
        step()
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
    s = tr.stats()
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    fl = torch.tensor([s["flops"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    ms_max, flops_all = float(tms.item()), float(fl.item())
    value = flops_all / (ms_max * 1e-3) * 1e-9
    launches = int(s["fused_launches"] + s["repack_launches"] + s["reduce_launches"])
    fused_avg_ms = s["fused_ms"] / max(1, s["fused_launches"])
    achieved = (s["flops"] / max(1, s["fused_launches"])) / (fused_avg_ms * 1e-3) * 1e-12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_fused_r01.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "fused_kernel (DMMA.8x8x4 + UBLKCP)", "peak_source": peak_how,
                "flops_per_launch": s["flops"] / max(1, s["fused_launches"]), "avg_launch_ms": fused_avg_ms}

    # ---- e2e through the reference-facing Tier-1 surface (host buffers), rank-local ----
    e2e = None
    if not a.no_e2e and a.workload == WORKLOAD and not a.strong:
        for arr in (st.t1, st.t2, st.v2):
            arr.setflags(write=True)
        pinned = []
        for arr in (st.t2, st.v2):
            try:
                capi.host_register(arr); pinned.append(arr)
            except Exception:
                pass
        import ctypes
        capi.lib().nwc_triples_set_local_rank(ctypes.c_long(local))
        for _ in range(min(a.warmup, 1)):
            capi.ccsd_t_gpu(st, icuda=max(world, local + 1))
        capi.compat_stats(reset=True)
        barrier()
        t0 = time.perf_counter()
        nrep = max(1, min(a.steps, 3))
        for _ in range(nrep):
            c1, c2, _ = capi.ccsd_t_gpu(st, icuda=max(world, local + 1))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        cs = capi.compat_stats()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        ff = torch.tensor([cs["flops"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(ff, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ff.item()) / float(tt.item()) * 1e-9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": cs["h2d_bytes"] / nrep, "d2h_bytes_per_step": cs["d2h_bytes"] / nrep,
               "ms_per_step": dt / nrep * 1e3, "api": "nwc_ccsd_t_gpu -> sd_t_*_cuda_/compute_en_ (Tier 1), host TCE_SORT included",
               "energy_matches_native": bool(abs(c1 - e[0]) <= 1e-9 * max(1.0, abs(e[0]))) if world == 1 else None}
        for arr in pinned:
            capi.host_unregister(arr)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_sample(1)
        cpu = {"value": r["gflops"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "seconds": r["seconds"]}
    if rank == 0:
        line = {"metric": "(T) FP64 GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "strong" if a.strong else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "wall_s_per_step": ms_max / a.steps * 1e-3, "flops_per_step": flops_all / a.steps,
                "frac_of_fp64_peak": value * 1e-3 / (peak * world), "energy": list(e),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
