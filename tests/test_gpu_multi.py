"""Multi-GPU tests (skipped on boxes with fewer than two GPUs): the CUDA-IPC / NVLink flavour of the sharded V2 store
(spin-orbital and `2eorb`), the static block partition across ranks and the library's own NCCL reduction (ga_dgop,
ccsd_t.F:297), one process per GPU, against the single-GPU result on the same device-generated stores."""
import os
import socket
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, intorb, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from nwchem_b200 import capi, synth
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    t = synth.shape_tiling("h2o_ccpvdz_c2v")
    tr = capi.Triples(rank)
    st = synth.empty_stores(t, intorb=intorb)
    if intorb:
        tr.set_state_2eorb(st, rank, world)
    else:
        tr.set_state_sharded(st, rank, world)
    tr.synth_fill(123)
    mine = torch.tensor(list(tr.v2_ipc_handle()), dtype=torch.uint8, device="cuda")
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    tr.v2_open_peers(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(capi.Triples.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    tr.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    e1, e2, pt = tr.run_partition(rank, world, per_task=True)
    r1, r2 = tr.allreduce(e1, e2)                  # the library's ncclAllReduce of the two scalars
    ptsum = tr.allreduce_sum(pt.ravel()).reshape(pt.shape)
    peer = tr.stats()["peer_bytes"]
    tr.close()
    if rank == 0:
        one = capi.Triples(0)                      # the same stores, whole, on one GPU
        s1 = synth.empty_stores(t, intorb=intorb)
        if intorb:
            one.set_state_2eorb(s1)
        else:
            one.set_state(s1)
        one.synth_fill(123)
        f1, f2, fpt = one.run(per_task=True)
        one.close()
        out.put((r1, r2, ptsum, f1, f2, fpt, peer))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("intorb", [False, True])
def test_two_gpus_ipc_sharded_partition_nccl(intorb):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, intorb, q)) for r in range(world)]
    for p in procs:
        p.start()
    r1, r2, ptsum, f1, f2, fpt, peer = q.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert peer > 0                                  # blocks really came from the other GPU's shard
    assert abs(r1 - f1) <= 1e-12 * max(1.0, abs(f1)) and abs(r2 - f2) <= 1e-12 * max(1.0, abs(f2))
    assert np.max(np.abs(ptsum - fpt)) <= 1e-14 + 1e-12 * np.max(np.abs(fpt))
