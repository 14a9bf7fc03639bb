"""ctypes binding of nwchem_b200/lib/libnwc_triples.so (include/nwc_triples.h).

The library is the product; this module only marshals numpy arrays into its C ABI.  There is no CPU
fallback: if the shared library is missing the import of this module raises.
"""
from __future__ import annotations
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libnwc_triples.so")
L = C.c_long
PL = C.POINTER(C.c_long)
PD = C.POINTER(C.c_double)


class TceState(C.Structure):  # nwc_tce_state
    _fields_ = [("noab", L), ("nvab", L), ("restricted", L), ("irrep_t", L), ("irrep_v", L),
                ("spin", PL), ("sym", PL), ("range", PL), ("offset", PL), ("alpha", PL), ("evl_sorted", PD),
                ("t1_hash", PL), ("t1", PD), ("t2_hash", PL), ("t2", PD), ("v2_hash", PL), ("v2", PD)]


class OrbState(C.Structure):  # nwc_tce_orb_state
    _fields_ = [("noa", L), ("nva", L), ("b2am", PL), ("spin_alpha", PL), ("sym_alpha", PL), ("range_alpha", PL),
                ("v2orb_hash", PL), ("v2orb", PD)]


class TraceRec(C.Structure):  # nwc_trace_rec
    _fields_ = [("kind", L), ("k0", L), ("side", L), ("K", L), ("neg", L), ("a", C.c_void_p), ("b", C.c_void_p),
                ("sa", C.c_longlong * 6), ("sb", C.c_longlong * 6), ("ka", C.c_longlong), ("kb", C.c_longlong),
                ("scale", C.c_double)]


class Stats(C.Structure):  # nwc_triples_stats
    _fields_ = [("fused_ms", C.c_double), ("repack_ms", C.c_double), ("fused_launches", C.c_longlong),
                ("repack_launches", C.c_longlong), ("reduce_launches", C.c_longlong), ("work_items", C.c_longlong),
                ("descs", C.c_longlong), ("tuples", C.c_longlong), ("flops", C.c_double), ("h2d_bytes", C.c_double),
                ("d2h_bytes", C.c_double), ("resident_bytes", C.c_double), ("pull_ms", C.c_double),
                ("peer_bytes", C.c_double), ("pull_launches", C.c_longlong), ("antisym_launches", C.c_longlong)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C nwchem_b200/csrc` "
                               "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        _lib = C.CDLL(os.environ.get("NWC_TRIPLES_LIB", LIB_PATH))   # override: experiments with alternative builds
        _lib.nwc_triples_last_error.restype = C.c_char_p
        _lib.nwc_triples_num_tasks.restype = L
        _lib.nwc_triples_num_tasks.argtypes = [C.c_void_p]
        _lib.nwc_triples_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        _lib.nwc_triples_destroy.argtypes = [C.c_void_p]
        _lib.nwc_triples_set_state.argtypes = [C.c_void_p, C.POINTER(TceState)]
        _lib.nwc_triples_task_list.argtypes = [C.c_void_p, PL]
        _lib.nwc_triples_set_state_sharded.argtypes = [C.c_void_p, C.POINTER(TceState), C.c_int, C.c_int]
        _lib.nwc_triples_v2_ipc_handle.argtypes = [C.c_void_p, C.c_char_p]
        _lib.nwc_triples_v2_open_peers.argtypes = [C.c_void_p, C.c_char_p]
        _lib.nwc_triples_run.argtypes = [C.c_void_p, L, L, L, PD, PD]
        _lib.nwc_triples_run_tuple.argtypes = [C.c_void_p, PL, PD, PD, PD]
        _lib.nwc_triples_set_timing.argtypes = [C.c_void_p, C.c_int]
        _lib.nwc_triples_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats), C.c_int]
        _lib.nwc_triples_set_batch_bytes.argtypes = [C.c_void_p, C.c_size_t]
        _lib.nwc_triples_timer_start.argtypes = [C.c_void_p]
        _lib.nwc_triples_timer_stop_ms.argtypes = [C.c_void_p, PD]
        _lib.nwc_host_register.argtypes = [C.c_void_p, C.c_size_t]
        _lib.nwc_host_unregister.argtypes = [C.c_void_p]
        _lib.nwc_triples_nccl_unique_id.argtypes = [C.c_char_p]
        _lib.nwc_triples_nccl_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        _lib.nwc_triples_allreduce_energy.argtypes = [C.c_void_p, PD]
        _lib.nwc_ccsd_t_gpu.argtypes = [C.POINTER(TceState), L, L, L, PD, PD]
        _lib.nwc_ccsd_t_gpu_tasks.argtypes = [C.POINTER(TceState), L, PL, L, PD, PD]
        _lib.nwc_triples_set_state_2eorb.argtypes = [C.c_void_p, C.POINTER(TceState), C.POINTER(OrbState)]
        _lib.nwc_triples_set_state_2eorb_sharded.argtypes = [C.c_void_p, C.POINTER(TceState), C.POINTER(OrbState), C.c_int, C.c_int]
        _lib.nwc_triples_run_partition.argtypes = [C.c_void_p, L, L, L, L, PD, PD]
        _lib.nwc_triples_run_items.argtypes = [C.c_void_p, PL, C.c_longlong, C.c_longlong, PD]
        _lib.nwc_triples_tuple_items.argtypes = [C.c_void_p, PL]
        _lib.nwc_triples_tuple_items.restype = C.c_longlong
        _lib.nwc_triples_synth_fill.argtypes = [C.c_void_p, C.c_ulonglong, C.c_double, C.c_double, C.c_double]
        _lib.nwc_triples_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, PD]
        _lib.nwc_triples_export_v2_block.argtypes = [C.c_void_p, PL, PD]
        _lib.nwc_triples_allreduce_sum.argtypes = [C.c_void_p, PD, C.c_size_t]
        _lib.nwc_triples_set_arena_cap.argtypes = [C.c_void_p, C.c_size_t]
        _lib.nwc_triples_set_host_threads.argtypes = [C.c_int]
        _lib.nwc_driver_set_reference_contract.argtypes = [C.c_int]
        _lib.nwc_ccsd_t_gpu_tuple.argtypes = [C.POINTER(TceState), PL, PD, PD, PD]
    return _lib


def _pl(a):
    return a.ctypes.data_as(PL)


def _pd(a):
    return a.ctypes.data_as(PD)


def make_state(st):
    """BlockStores -> (nwc_tce_state, keepalive dict of the contiguous arrays it points into).
    A data array that is None becomes a NULL pointer: the library allocates that store without uploading."""
    t = st.t
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
    i64 = lambda a: None if a is None else np.ascontiguousarray(a, np.int64)
    k = dict(spin=i64(t.spin), sym=i64(t.sym), range=i64(t.range), offset=i64(t.offset), alpha=i64(t.alpha),
             evl=f64(t.evl_sorted), t1h=i64(st.t1_hash), t1=f64(st.t1), t2h=i64(st.t2_hash), t2=f64(st.t2),
             v2h=i64(st.v2_hash), v2=f64(st.v2))
    pl = lambda a: None if a is None else _pl(a)
    pd = lambda a: None if a is None else _pd(a)
    s = TceState(t.noab, t.nvab, int(t.restricted), 0, 0, _pl(k["spin"]), _pl(k["sym"]), _pl(k["range"]),
                 _pl(k["offset"]), _pl(k["alpha"]), _pd(k["evl"]), pl(k["t1h"]), pd(k["t1"]), pl(k["t2h"]),
                 pd(k["t2"]), pl(k["v2h"]), pd(k["v2"]))
    return s, k


def make_orb_state(orb):
    a = orb.a
    k = dict(b2am=np.ascontiguousarray(a.b2am, np.int64), spa=np.ascontiguousarray(a.spin_alpha, np.int64),
             sya=np.ascontiguousarray(a.sym_alpha, np.int64), rga=np.ascontiguousarray(a.range_alpha, np.int64),
             voh=np.ascontiguousarray(orb.v2orb_hash, np.int64),
             vo=None if orb.v2orb is None else np.ascontiguousarray(orb.v2orb, np.float64))
    o = OrbState(a.noa, a.nva, _pl(k["b2am"]), _pl(k["spa"]), _pl(k["sya"]), _pl(k["rga"]), _pl(k["voh"]),
                 None if k["vo"] is None else _pd(k["vo"]))
    return o, k


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().nwc_triples_last_error().decode()}")


# ------------------------------------------------------------------------------------------------
# Tier 1 through the host driver (mirrors ccsd_t_gpu.F)
# ------------------------------------------------------------------------------------------------
def ccsd_t_gpu(st, icuda=1, my_rank=0, nranks=1, ntasks=None):
    """Whole (T) through the reference call surface (host block stores, per-call H2D copies).
    Returns (E[T], E(T), per_task[ntasks,2])."""
    s, keep = make_state(st)
    e = np.zeros(2)
    pt = np.zeros((max(ntasks or 0, 1), 2)) if ntasks else None
    rc = lib().nwc_ccsd_t_gpu(C.byref(s), icuda, my_rank, nranks, _pd(e), _pd(pt) if pt is not None else None)
    _check(rc, "nwc_ccsd_t_gpu")
    return float(e[0]), float(e[1]), pt


def ccsd_t_gpu_tasks(st, tasks, icuda=1):
    """A given list of tasks (rows of >= 6 tile ids) through Tier 1.  Returns (E[T], E(T), per_task[n,2])."""
    s, keep = make_state(st)
    tt = np.ascontiguousarray(np.asarray(tasks, np.int64)[:, :6])
    e = np.zeros(2)
    pt = np.zeros((max(len(tt), 1), 2))
    _check(lib().nwc_ccsd_t_gpu_tasks(C.byref(s), icuda, _pl(tt), len(tt), _pd(e), _pd(pt)), "nwc_ccsd_t_gpu_tasks")
    return float(e[0]), float(e[1]), pt[:len(tt)]


def bind_backend(so_path):
    """Route the host driver's Tier-1 calls into another library exporting the reference's symbols (None: back)."""
    l = lib()
    l.nwc_driver_bind_backend.argtypes = [C.c_char_p]
    rc = l.nwc_driver_bind_backend(so_path.encode() if so_path else None)
    if rc != 0:
        raise RuntimeError(f"nwc_driver_bind_backend({so_path}) failed")


def set_reference_contract(on=True):
    lib().nwc_driver_set_reference_contract(int(on))


def set_host_threads(n):
    lib().nwc_triples_set_host_threads(int(n))


def ccsd_t_gpu_tuple(st, tup, dump=False):
    """One tuple (p4b,p5b,p6b,h1b,h2b,h3b) through Tier 1; with dump=True also returns the singles and
    doubles t3 tiles as arrays indexed [p4,p5,p6,h1,h2,h3]."""
    s, keep = make_state(st)
    tt = np.array(tup, np.int64)
    e = np.zeros(2)
    if dump:
        dims = [st.t.r(int(b)) for b in tup]
        d = np.zeros(int(np.prod(dims))); sg = np.zeros_like(d)
        _check(lib().nwc_ccsd_t_gpu_tuple(C.byref(s), _pl(tt), _pd(e), _pd(d), _pd(sg)), "nwc_ccsd_t_gpu_tuple")
        return float(e[0]), float(e[1]), sg.reshape(dims), d.reshape(dims)
    _check(lib().nwc_ccsd_t_gpu_tuple(C.byref(s), _pl(tt), _pd(e), None, None), "nwc_ccsd_t_gpu_tuple")
    return float(e[0]), float(e[1])


# ------------------------------------------------------------------------------------------------
# Tier 2
# ------------------------------------------------------------------------------------------------
class Triples:
    """Native tier: block stores resident in HBM; static task partition; optional NCCL reduction."""

    def __init__(self, device: int = 0, trace: bool = False):
        """trace=True: host-only trace context (nwc_triples_create_trace): records the driver's operand descriptors,
        executes nothing; the arrays handed to set_state / set_lambda / set_cr are kept by reference."""
        self._h = C.c_void_p()
        self._keep = []
        self.is_trace = bool(trace)
        if trace:
            lib().nwc_triples_create_trace.argtypes = [C.POINTER(C.c_void_p)]
            _check(lib().nwc_triples_create_trace(C.byref(self._h)), "nwc_triples_create_trace")
        else:
            _check(lib().nwc_triples_create(C.byref(self._h), device), "nwc_triples_create")
        self.t = None

    def trace_tuple(self, tup, method: int):
        """Records of one tuple (method 0 (T), 1 Lambda-CCSD(T), 2 / 3 CR-CCSD(T) numerator / denominator pass) as a
        list of TraceRec; see include/nwc_triples.h nwc_trace_rec."""
        l = lib()
        l.nwc_triples_trace_tuple.argtypes = [C.c_void_p, PL, C.c_int]
        l.nwc_triples_trace_take.argtypes = [C.c_void_p, C.POINTER(TraceRec), C.c_size_t, C.POINTER(C.c_size_t)]
        tt = np.array([int(x) for x in tup[:6]], np.int64)
        _check(l.nwc_triples_trace_tuple(self._h, _pl(tt), int(method)), "nwc_triples_trace_tuple")
        cap = 4096
        buf = (TraceRec * cap)()
        n = C.c_size_t(0)
        _check(l.nwc_triples_trace_take(self._h, buf, cap, C.byref(n)), "nwc_triples_trace_take")
        if n.value > cap:
            raise RuntimeError("trace buffer too small")
        return [buf[i] for i in range(n.value)], buf

    def close(self):
        if self._h:
            lib().nwc_triples_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, st):
        s, keep = make_state(st)
        _check(lib().nwc_triples_set_state(self._h, C.byref(s)), "nwc_triples_set_state")
        self.t = st.t
        if self.is_trace:
            self._keep.append(keep)

    def set_state_2eorb(self, st, rank: int = 0, world: int = 1):
        """`2eorb` storage: V2 is read from st.orb (synth.OrbitalV2), never from st.v2.  world > 1: the orbital
        blocks are sharded over the ranks (block i of the needed ones -> rank i % world); st.orb.v2orb is the full
        store (or None: allocate only, fill with synth_fill)."""
        s, keep = make_state(st)
        o, keep2 = make_orb_state(st.orb)
        s.v2_hash = None; s.v2 = None
        if world > 1:
            _check(lib().nwc_triples_set_state_2eorb_sharded(self._h, C.byref(s), C.byref(o), rank, world),
                   "nwc_triples_set_state_2eorb_sharded")
        else:
            _check(lib().nwc_triples_set_state_2eorb(self._h, C.byref(s), C.byref(o)), "nwc_triples_set_state_2eorb")
        self.t = st.t

    def synth_fill(self, seed: int, scale=(0.05, 0.02, 0.1)):
        """Fill the resident stores on the device with the keyed generator (synth.keyed_* restates it in numpy)."""
        _check(lib().nwc_triples_synth_fill(self._h, seed, float(scale[0]), float(scale[1]), float(scale[2])),
               "nwc_triples_synth_fill")

    def debug_read(self, which: int, offset: int, n: int) -> np.ndarray:
        out = np.zeros(n)
        _check(lib().nwc_triples_debug_read(self._h, which, offset, n, _pd(out)), "nwc_triples_debug_read")
        return out

    def export_v2_block(self, g3b, g4b, g1b, g2b) -> np.ndarray:
        g = np.array([g3b, g4b, g1b, g2b], np.int64)
        out = np.zeros(int(np.prod([self.t.r(int(b)) for b in g])))
        _check(lib().nwc_triples_export_v2_block(self._h, _pl(g), _pd(out)), "nwc_triples_export_v2_block")
        return out

    def run_partition(self, rank: int, world: int, first_task: int = 0, ntasks: int = 0, per_task=False):
        """Static equal-cost block partition of tasks [first_task, first_task+ntasks) over `world` ranks, tuples on a
        boundary shared at sub-tile granularity.  Returns this rank's (E[T], E(T)[, per_task partials])."""
        e = np.zeros(2)
        n = self.num_tasks - first_task if ntasks <= 0 else min(ntasks, self.num_tasks - first_task)
        pt = np.zeros((max(n, 1), 2)) if per_task else None
        _check(lib().nwc_triples_run_partition(self._h, rank, world, first_task, ntasks, _pd(e), _pd(pt) if per_task else None),
               "nwc_triples_run_partition")
        return (float(e[0]), float(e[1]), pt[:n]) if per_task else (float(e[0]), float(e[1]))

    def run_partition_list(self, rank: int, world: int, task_ids, per_task=False):
        """Like run_partition for an explicit list of task indices (partitioned in the order given)."""
        ids = np.ascontiguousarray(task_ids, np.int64)
        e = np.zeros(2)
        pt = np.zeros((max(len(ids), 1), 2)) if per_task else None
        l = lib()
        l.nwc_triples_run_partition_list.argtypes = [C.c_void_p, L, L, PL, L, PD, PD]
        _check(l.nwc_triples_run_partition_list(self._h, rank, world, _pl(ids), len(ids), _pd(e), _pd(pt) if per_task else None),
               "nwc_triples_run_partition_list")
        return (float(e[0]), float(e[1]), pt[:len(ids)]) if per_task else (float(e[0]), float(e[1]))

    def set_lambda(self, lam):
        """lam: synth.LambdaStores (lambda_1, lambda_2, Fock (h,p) blocks + offset tables)."""
        k = [np.ascontiguousarray(a, np.int64 if i % 2 == 0 else np.float64)
             for i, a in enumerate((lam.y1_hash, lam.y1, lam.y2_hash, lam.y2, lam.f1_hash, lam.f1))]
        l = lib()
        l.nwc_triples_set_lambda.argtypes = [C.c_void_p, PL, PD, PL, PD, PL, PD]
        _check(l.nwc_triples_set_lambda(self._h, _pl(k[0]), _pd(k[1]), _pl(k[2]), _pd(k[3]), _pl(k[4]), _pd(k[5])),
               "nwc_triples_set_lambda")
        if self.is_trace:
            self._keep.append(k)

    def run_lambda(self, first=0, stride=1, max_tasks=0, per_task=False):
        """Lambda-CCSD[T] / Lambda-CCSD(T) correction energies (lambda_ccsd_t.F), tasks first, first+stride, ..."""
        e = np.zeros(2)
        cnt = len(range(first, self.num_tasks, stride))
        if max_tasks and max_tasks > 0:
            cnt = min(cnt, max_tasks)
        pt = np.zeros((max(cnt, 1), 2)) if per_task else None
        l = lib()
        l.nwc_triples_run_lambda.argtypes = [C.c_void_p, L, L, L, PD, PD]
        _check(l.nwc_triples_run_lambda(self._h, first, stride, max_tasks, _pd(e), _pd(pt) if per_task else None),
               "nwc_triples_run_lambda")
        return (float(e[0]), float(e[1]), pt[:cnt]) if per_task else (float(e[0]), float(e[1]))

    def run_lambda_partition(self, rank: int, world: int, first_task: int = 0, ntasks: int = 0, per_task=False):
        e = np.zeros(2)
        n = self.num_tasks - first_task if ntasks <= 0 else min(ntasks, self.num_tasks - first_task)
        pt = np.zeros((max(n, 1), 2)) if per_task else None
        l = lib()
        l.nwc_triples_run_lambda_partition.argtypes = [C.c_void_p, L, L, L, L, PD, PD]
        _check(l.nwc_triples_run_lambda_partition(self._h, rank, world, first_task, ntasks, _pd(e), _pd(pt) if per_task else None),
               "nwc_triples_run_lambda_partition")
        return (float(e[0]), float(e[1]), pt[:n]) if per_task else (float(e[0]), float(e[1]))

    def set_cr(self, cr):
        """cr: the three CR-CCSD(T) intermediates with their offset tables (attributes n1_hash, n1, n2_hash, n2, e2_hash,
        e2 -- tiling.cr_n1_offset / cr_n2_offset / cr_e2_offset layouts; cr_ccsd_t_N.F / cr_ccsd_t_E.F toggle 1)."""
        k = [np.ascontiguousarray(a, np.int64 if i % 2 == 0 else np.float64)
             for i, a in enumerate((cr.n1_hash, cr.n1, cr.n2_hash, cr.n2, cr.e2_hash, cr.e2))]
        l = lib()
        l.nwc_triples_set_cr.argtypes = [C.c_void_p, PL, PD, PL, PD, PL, PD]
        _check(l.nwc_triples_set_cr(self._h, _pl(k[0]), _pd(k[1]), _pl(k[2]), _pd(k[3]), _pl(k[4]), _pd(k[5])),
               "nwc_triples_set_cr")
        if self.is_trace:
            self._keep.append(k)

    def set_cr_sharded(self, cr, rank: int, world: int):
        """As set_cr, but cr.n2 holds only this rank's blocks of the pphp intermediate (synth.shard_store); afterwards
        exchange cr_shard_ptr / cr_set_peer_ptr (one process) or cr_ipc_handle / cr_open_peers (one process per GPU)."""
        k = [np.ascontiguousarray(a, np.int64 if i % 2 == 0 else np.float64)
             for i, a in enumerate((cr.n1_hash, cr.n1, cr.n2_hash, cr.n2, cr.e2_hash, cr.e2))]
        l = lib()
        l.nwc_triples_set_cr_sharded.argtypes = [C.c_void_p, PL, PD, PL, PD, PL, PD, C.c_int, C.c_int]
        _check(l.nwc_triples_set_cr_sharded(self._h, _pl(k[0]), _pd(k[1]), _pl(k[2]), _pd(k[3]), _pl(k[4]), _pd(k[5]), rank, world),
               "nwc_triples_set_cr_sharded")

    def cr_shard_ptr(self) -> int:
        lib().nwc_triples_cr_shard_ptr.restype = C.c_void_p
        lib().nwc_triples_cr_shard_ptr.argtypes = [C.c_void_p]
        return int(lib().nwc_triples_cr_shard_ptr(self._h) or 0)

    def cr_set_peer_ptr(self, rank: int, ptr: int):
        lib().nwc_triples_cr_set_peer_ptr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _check(lib().nwc_triples_cr_set_peer_ptr(self._h, rank, C.c_void_p(ptr)), "nwc_triples_cr_set_peer_ptr")

    def cr_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        lib().nwc_triples_cr_ipc_handle.argtypes = [C.c_void_p, C.c_char_p]
        _check(lib().nwc_triples_cr_ipc_handle(self._h, buf), "nwc_triples_cr_ipc_handle")
        return buf.raw

    def cr_open_peers(self, handles: bytes):
        lib().nwc_triples_cr_open_peers.argtypes = [C.c_void_p, C.c_char_p]
        _check(lib().nwc_triples_cr_open_peers(self._h, handles), "nwc_triples_cr_open_peers")

    def run_cr(self, first=0, stride=1, max_tasks=0, per_task=False):
        """CR-CCSD(T) tuple loop (cr_ccsd_t.F:93-233): sums = (num1, num2, den1, den2) without den0 [, per_task[n,4]]."""
        s = np.zeros(4)
        cnt = len(range(first, self.num_tasks, stride))
        if max_tasks and max_tasks > 0:
            cnt = min(cnt, max_tasks)
        pt = np.zeros((max(cnt, 1), 4)) if per_task else None
        l = lib()
        l.nwc_triples_run_cr.argtypes = [C.c_void_p, L, L, L, PD, PD]
        _check(l.nwc_triples_run_cr(self._h, first, stride, max_tasks, _pd(s), _pd(pt) if per_task else None),
               "nwc_triples_run_cr")
        return (s, pt[:cnt]) if per_task else s

    def run_cr_partition(self, rank: int, world: int, first_task: int = 0, ntasks: int = 0, per_task=False):
        s = np.zeros(4)
        n = self.num_tasks - first_task if ntasks <= 0 else min(ntasks, self.num_tasks - first_task)
        pt = np.zeros((max(n, 1), 4)) if per_task else None
        l = lib()
        l.nwc_triples_run_cr_partition.argtypes = [C.c_void_p, L, L, L, L, PD, PD]
        _check(l.nwc_triples_run_cr_partition(self._h, rank, world, first_task, ntasks, _pd(s), _pd(pt) if per_task else None),
               "nwc_triples_run_cr_partition")
        return (s, pt[:n]) if per_task else s

    @staticmethod
    def cr_energies(sums, den0):
        """cr_ccsd_t.F:260-263: (CR-CCSD[T], CR-CCSD(T)) corrections from the four sums and the scalar of cr_ccsd_t_D."""
        return float(sums[0] / (1.0 + sums[2] + den0)), float(sums[1] / (1.0 + sums[3] + den0))

    def set_creom(self, q):
        """q: the CR-EOMCCSD(T) inputs (attributes x1_hash, x1, x2_hash, x2, m1..m4 (+_hash) = d_i2_1..4, q2 (+_hash) = d_i3_1,
        r0, excit); call set_cr first when r0 != 0."""
        names = ("x1", "x2", "m1", "m2", "m3", "m4", "q2")
        k = []
        for n in names:
            k += [np.ascontiguousarray(getattr(q, n + "_hash"), np.int64), np.ascontiguousarray(getattr(q, n), np.float64)]
        l = lib()
        l.nwc_triples_set_creom.argtypes = [C.c_void_p] + [PL, PD] * 7 + [C.c_double, C.c_double]
        args = []
        for i in range(7):
            args += [_pl(k[2 * i]), _pd(k[2 * i + 1])]
        _check(l.nwc_triples_set_creom(self._h, *args, float(q.r0), float(q.excit)), "nwc_triples_set_creom")
        if self.is_trace:
            self._keep.append(k)

    def run_creom(self, first=0, stride=1, max_tasks=0, per_task=False):
        """CR-EOMCCSD(T) tuple loop (cr_eomccsd_t.F:325-493): sums = (sum f R R/denex, sum f L R, sum f L R/denex, sum f L L)."""
        s = np.zeros(4)
        cnt = len(range(first, self.num_tasks, stride))
        if max_tasks and max_tasks > 0:
            cnt = min(cnt, max_tasks)
        pt = np.zeros((max(cnt, 1), 4)) if per_task else None
        l = lib()
        l.nwc_triples_run_creom.argtypes = [C.c_void_p, L, L, L, PD, PD]
        _check(l.nwc_triples_run_creom(self._h, first, stride, max_tasks, _pd(s), _pd(pt) if per_task else None),
               "nwc_triples_run_creom")
        return (s, pt[:cnt]) if per_task else s

    def run_creom_partition(self, rank: int, world: int, first_task: int = 0, ntasks: int = 0, per_task=False):
        s = np.zeros(4)
        n = self.num_tasks - first_task if ntasks <= 0 else min(ntasks, self.num_tasks - first_task)
        pt = np.zeros((max(n, 1), 4)) if per_task else None
        l = lib()
        l.nwc_triples_run_creom_partition.argtypes = [C.c_void_p, L, L, L, L, PD, PD]
        _check(l.nwc_triples_run_creom_partition(self._h, rank, world, first_task, ntasks, _pd(s), _pd(pt) if per_task else None),
               "nwc_triples_run_creom_partition")
        return (s, pt[:n]) if per_task else s

    def tuple_items(self, tup) -> int:
        tt = np.array(tup, np.int64)
        return int(lib().nwc_triples_tuple_items(self._h, _pl(tt)))

    def run_items(self, tup, item_lo: int, item_hi: int):
        """One tuple restricted to sub-tiles [item_lo, item_hi) (linear 4-wide-block order, p4 block slowest)."""
        tt = np.array(tup, np.int64)
        e = np.zeros(2)
        _check(lib().nwc_triples_run_items(self._h, _pl(tt), item_lo, item_hi, _pd(e)), "nwc_triples_run_items")
        return float(e[0]), float(e[1])

    @property
    def order(self) -> int:
        lib().nwc_triples_get_order.argtypes = [C.c_void_p]
        return int(lib().nwc_triples_get_order(self._h))

    def trim(self):
        """Release the batch arenas (the resident stores stay)."""
        lib().nwc_triples_trim.argtypes = [C.c_void_p]
        _check(lib().nwc_triples_trim(self._h), "nwc_triples_trim")

    def set_arena_cap(self, n):
        lib().nwc_triples_set_arena_cap(self._h, int(n))

    def allreduce_sum(self, arr: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(arr, np.float64).copy()
        _check(lib().nwc_triples_allreduce_sum(self._h, _pd(a), a.size), "nwc_triples_allreduce_sum")
        return a

    def set_state_sharded(self, st_shard, rank: int, world: int):
        """st_shard.v2 holds only this rank's V2 blocks (see synth.shard_v2); tables are the full ones."""
        s, keep = make_state(st_shard)
        _check(lib().nwc_triples_set_state_sharded(self._h, C.byref(s), rank, world), "nwc_triples_set_state_sharded")
        self.t = st_shard.t

    def v2_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().nwc_triples_v2_ipc_handle(self._h, buf), "nwc_triples_v2_ipc_handle")
        return buf.raw

    def v2_shard_ptr(self) -> int:
        lib().nwc_triples_v2_shard_ptr.restype = C.c_void_p
        lib().nwc_triples_v2_shard_ptr.argtypes = [C.c_void_p]
        return int(lib().nwc_triples_v2_shard_ptr(self._h) or 0)

    def v2_set_peer_ptr(self, rank: int, ptr: int):
        lib().nwc_triples_v2_set_peer_ptr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _check(lib().nwc_triples_v2_set_peer_ptr(self._h, rank, C.c_void_p(ptr)), "nwc_triples_v2_set_peer_ptr")

    def v2_open_peers(self, handles: bytes):
        _check(lib().nwc_triples_v2_open_peers(self._h, handles), "nwc_triples_v2_open_peers")

    @property
    def num_tasks(self) -> int:
        return int(lib().nwc_triples_num_tasks(self._h))

    def task_list(self):
        n = self.num_tasks
        kl = np.zeros((max(n, 1), 7), np.int64)
        lib().nwc_triples_task_list(self._h, _pl(kl))
        return kl[:n]

    def run(self, first=0, stride=1, max_tasks=0, per_task=False):
        e = np.zeros(2)
        n = self.num_tasks
        cnt = len(range(first, n, stride))
        if max_tasks and max_tasks > 0:
            cnt = min(cnt, max_tasks)
        pt = np.zeros((max(cnt, 1), 2)) if per_task else None
        _check(lib().nwc_triples_run(self._h, first, stride, max_tasks, _pd(e), _pd(pt) if per_task else None),
               "nwc_triples_run")
        return (float(e[0]), float(e[1]), pt[:cnt]) if per_task else (float(e[0]), float(e[1]))

    def run_restart(self, begin=1, table=None, max_outer=0, first=0, stride=1):
        """Restartable (T) (ccsd_t_restart.F): returns (new begin, table[nvab] of CCSD(T) partials per outer virtual
        tile, table of CCSD[T] partials, t_energy).  Pass the returned begin/table back in to resume."""
        nv = self.t.nvab
        tab = np.zeros(nv) if table is None else np.ascontiguousarray(table, np.float64).copy()
        tab1 = np.zeros(nv)
        b = L(begin)
        te = C.c_double(0.0)
        l = lib()
        l.nwc_triples_run_restart.argtypes = [C.c_void_p, L, L, C.POINTER(L), C.POINTER(C.c_double),
                                              C.POINTER(C.c_double), L, C.POINTER(C.c_double)]
        _check(l.nwc_triples_run_restart(self._h, first, stride, C.byref(b), _pd(tab), _pd(tab1), max_outer, C.byref(te)),
               "nwc_triples_run_restart")
        return int(b.value), tab, tab1, float(te.value)

    def run_tuple(self, tup, dump=False):
        tt = np.array(tup, np.int64)
        e = np.zeros(2)
        if dump:
            dims = [self.t.r(int(b)) for b in tup]
            d = np.zeros(int(np.prod(dims))); sg = np.zeros_like(d)
            _check(lib().nwc_triples_run_tuple(self._h, _pl(tt), _pd(e), _pd(d), _pd(sg)), "nwc_triples_run_tuple")
            return float(e[0]), float(e[1]), sg.reshape(dims), d.reshape(dims)
        _check(lib().nwc_triples_run_tuple(self._h, _pl(tt), _pd(e), None, None), "nwc_triples_run_tuple")
        return float(e[0]), float(e[1])

    def set_timing(self, on=True):
        lib().nwc_triples_set_timing(self._h, int(on))

    def set_batch_bytes(self, n):
        lib().nwc_triples_set_batch_bytes(self._h, int(n))

    def timer_start(self):
        lib().nwc_triples_timer_start(self._h)

    def timer_stop_ms(self) -> float:
        ms = C.c_double(0.0)
        lib().nwc_triples_timer_stop_ms(self._h, C.byref(ms))
        return ms.value

    def stats(self, reset=False) -> dict:
        s = Stats()
        lib().nwc_triples_get_stats(self._h, C.byref(s), int(reset))
        return s.asdict()

    # multi-GPU (one process per GPU)
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().nwc_triples_nccl_unique_id(buf), "nwc_triples_nccl_unique_id")
        return buf.raw

    def nccl_init(self, uid: bytes, rank: int, nranks: int):
        _check(lib().nwc_triples_nccl_init(self._h, uid, rank, nranks), "nwc_triples_nccl_init")

    def allreduce(self, e1, e2):
        e = np.array([e1, e2], np.float64)
        _check(lib().nwc_triples_allreduce_energy(self._h, _pd(e)), "nwc_triples_allreduce_energy")
        return float(e[0]), float(e[1])


# raw Tier-1 entry points for kernel-level tests (by-reference Fortran convention)
def tier1_single_call(family, k, dims_task, dims_perm, kd, tsub, v2sub, eps, factor):
    """Open a tuple with the TASK ranges dims_task=(h1d,h2d,h3d,p4d,p5d,p6d), issue ONE sd_t_{s1,d1,d2}_k_cuda_
    call with the PERMUTED ranges dims_perm (same order), then nwc_compute_en_dump_.
    Returns (e1, e2, singles_tile, doubles_tile) with tiles indexed [p4,p5,p6,h1,h2,h3]."""
    l = lib()
    ref = lambda v: C.byref(C.c_long(int(v)))
    T = [C.c_long(int(x)) for x in dims_task]
    P = [C.c_long(int(x)) for x in dims_perm]
    K = C.c_long(int(kd))
    l.initmemmodule_()
    l.dev_mem_s_(*[C.byref(x) for x in T])
    l.dev_mem_d_(*[C.byref(x) for x in T])
    tsub = np.ascontiguousarray(tsub, np.float64); v2sub = np.ascontiguousarray(v2sub, np.float64)
    name = {0: "s1", 1: "d1", 2: "d2"}[family]
    fn = getattr(l, f"sd_t_{name}_{k}_cuda_")
    h1, h2, h3, p4, p5, p6 = [C.byref(x) for x in P]
    if family == 0:
        fn(h1, h2, h3, p4, p5, p6, None, _pd(tsub), _pd(v2sub))
    elif family == 1:
        fn(h1, h2, h3, C.byref(K), p4, p5, p6, None, _pd(tsub), _pd(v2sub))
    else:
        fn(h1, h2, h3, p4, p5, p6, C.byref(K), None, _pd(tsub), _pd(v2sub))
    e = np.zeros(2)
    h1d, h2d, h3d, p4d, p5d, p6d = [int(x) for x in dims_task]
    n = h1d * h2d * h3d * p4d * p5d * p6d
    d = np.zeros(n); s = np.zeros(n)
    ev = [np.ascontiguousarray(x, np.float64) for x in eps]  # h1,h2,h3,p4,p5,p6
    f = C.c_double(factor)
    l.nwc_compute_en_dump_(C.byref(f), _pd(e), *[_pd(x) for x in ev], *[C.byref(x) for x in T], _pd(d), _pd(s))
    l.dev_release_()
    l.finalizememmodule_()
    shp = (p4d, p5d, p6d, h1d, h2d, h3d)
    return float(e[0]), float(e[1]), s.reshape(shp), d.reshape(shp)


# ------------------------------------------------------------------------------------------------
# host-only helpers (no device needed)
# ------------------------------------------------------------------------------------------------
def host_2eorb_plan(st, g3b, g4b, g1b, g2b):
    """Product host logic of the `2eorb` path (no device): (off_direct, off_exchange, strides[2][4]) into st.orb.v2orb."""
    s, keep = make_state(st)
    o, keep2 = make_orb_state(st.orb)
    g = np.array([g3b, g4b, g1b, g2b], np.int64)
    off = np.zeros(2, np.int64); strides = np.zeros(8, np.int64)
    l = lib()
    l.nwc_host_2eorb_plan.argtypes = [C.POINTER(TceState), C.POINTER(OrbState), PL, PL, PL]
    rc = l.nwc_host_2eorb_plan(C.byref(s), C.byref(o), _pl(g), _pl(off), _pl(strides))
    if rc != 0:
        raise RuntimeError("nwc_host_2eorb_plan: offset table does not match the tiling")
    return int(off[0]), int(off[1]), strides.reshape(2, 4)


def host_task_list(st):
    s, keep = make_state(st)
    l = lib()
    l.nwc_host_task_list.restype = L
    n = int(l.nwc_host_task_list(C.byref(s), None, L(0)))
    kl = np.zeros((max(n, 1), 7), np.int64)
    l.nwc_host_task_list(C.byref(s), _pl(kl), L(n))
    return kl[:n]


def host_collect_blocks(st, tasks, which):
    """Sorted unique block keys of store `which` (1 T1, 2 T2, 3 spin-orbital V2) the given tasks read (host only)."""
    s, keep = make_state(st)
    tt = np.ascontiguousarray(np.asarray(tasks, np.int64)[:, :6])
    l = lib()
    l.nwc_host_collect_blocks.restype = L
    l.nwc_host_collect_blocks.argtypes = [C.POINTER(TceState), PL, L, C.c_int, PL, L]
    n = int(l.nwc_host_collect_blocks(C.byref(s), _pl(tt), len(tt), which, None, 0))
    if n < 0:
        raise RuntimeError("nwc_host_collect_blocks failed")
    out = np.zeros(max(n, 1), np.int64)
    l.nwc_host_collect_blocks(C.byref(s), _pl(tt), len(tt), which, _pl(out), n)
    return out[:n]


def fp64_peak_probe(device=0) -> float:
    v = C.c_double(0.0)
    l = lib()
    l.nwc_fp64_peak_probe.argtypes = [C.c_int, PD]
    if l.nwc_fp64_peak_probe(device, C.byref(v)) != 0:
        raise RuntimeError("nwc_fp64_peak_probe failed")
    return float(v.value)


def host_count_tuple(st, tup, state=None):
    s, keep = state if state is not None else make_state(st)
    calls = np.zeros(3, np.int64); flops = np.zeros(3)
    tt = np.array(tup, np.int64)
    lib().nwc_host_count_tuple(C.byref(s), _pl(tt), _pl(calls), _pd(flops))
    return calls, flops


def compat_stats(reset=False) -> dict:
    s = Stats()
    lib().nwc_compat_get_stats(C.byref(s), int(reset))
    return s.asdict()


def compat_trim():
    lib().nwc_compat_trim()


def compat_set_timing(on=True):
    lib().nwc_compat_set_timing(int(on))


def compat_timer_start():
    lib().nwc_compat_timer_start()


def compat_timer_stop_ms() -> float:
    ms = C.c_double(0.0)
    lib().nwc_compat_timer_stop_ms(C.byref(ms))
    return ms.value


def host_register(arr: np.ndarray):
    _check(lib().nwc_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes), "nwc_host_register")


def host_unregister(arr: np.ndarray):
    lib().nwc_host_unregister(C.c_void_p(arr.ctypes.data))
