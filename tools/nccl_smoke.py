"""2+ rank smoke test of the library's own NCCL path (unique id via torch.distributed, allreduce of energies)."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from nwchem_b200 import capi, synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
def log(*a): print(f"[r{rank}]", *a, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg up")
st = synth.physical(synth.shape_tiling("h2o_ccpvdz_c2v"))
tr = capi.Triples(local)
sharded = "--sharded" in sys.argv
if sharded:   # V2 sharded over the ranks, remote blocks read over NVLink through CUDA-IPC mappings
    tr.set_state_sharded(synth.shard_v2(st, rank, world), rank, world)
    mine = torch.tensor(list(tr.v2_ipc_handle()), dtype=torch.uint8, device="cuda")
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    tr.v2_open_peers(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    log("sharded V2: peers mapped, resident bytes", tr.stats()["resident_bytes"])
else:
    tr.set_state(st)
log("state set")
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    raw = capi.Triples.nccl_unique_id()
    uid = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
torch.cuda.synchronize()
log("uid broadcast", int(uid.sum()))
tr.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
log("nccl init done")
e1, e2 = tr.run(first=rank, stride=world)
log("partial", e1, e2)
t1, t2 = tr.allreduce(e1, e2)
log("allreduced", t1, t2)
# restartable (T): every rank gets the rank-summed table (one allreduce per outer virtual tile, ccsd_t_restart.F:255)
rb, rtab, _, rte = tr.run_restart(first=rank, stride=world)
log("restart table sum", rte)
assert abs(rte - t2) < 1e-12 and rb == st.t.nvab + 1, (rte, t2, rb)
if rank == 0:
    from oracle import oracle as ora
    ref = ora.ccsd_t(st)
    assert abs(ref["e1"] - t1) < 1e-12 and abs(ref["e2"] - t2) < 1e-12, (ref["e1"], t1)
    f1, f2 = tr.run()
    assert abs(f1 - t1) < 1e-12 and abs(f2 - t2) < 1e-12, (f1, t1)
    log("OK matches single-rank total")
dist.barrier(); tr.close(); dist.destroy_process_group()
