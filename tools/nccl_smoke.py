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
tr = capi.Triples(local); tr.set_state(st)
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
if rank == 0:
    f1, f2 = tr.run()
    assert abs(f1 - t1) < 1e-12 and abs(f2 - t2) < 1e-12, (f1, t1)
    log("OK matches single-rank total")
dist.barrier(); tr.close(); dist.destroy_process_group()
