// Tier 2: native API.  T1/T2/V2 block stores live in HBM (replicated per GPU) in the reference's own TCE
// block layout; per tuple the host walks the driver logic (host_driver.h) and emits device-side repack jobs
// (the reference's GET_HASH_BLOCK + TCE_SORT_4, done once on the device, sign folded in) and contraction
// descriptors; whole batches of tuples run in one fused launch.  Replaces get_block.F:79-81 (GA gets),
// util_gnxtval.c:31 (task counter -> static deal) and ccsd_t.F:297 (ga_dgop -> one ncclAllReduce).
#include "engine.h"
#include "host_driver.h"
#include <cstring>
#include <algorithm>
#include <unordered_map>
#include <string>
#include <dlfcn.h>

using namespace nwc;
namespace nwc { Engine& compat_engine(); }  // compat_abi.cu

static thread_local std::string g_err;
#define NWC_TRY(x)                                                                           \
  do {                                                                                       \
    cudaError_t _e = (x);                                                                    \
    if (_e != cudaSuccess) {                                                                 \
      g_err = std::string(#x) + ": " + cudaGetErrorString(_e);                               \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

// ---- NCCL, bound lazily so the library loads on a box without NCCL/GPU ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
namespace {
struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { g_err = "cannot dlopen libnccl.so.2"; return false; }
    GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce) { g_err = "libnccl lacks required symbols"; return false; }
    return true;
  }
} g_nccl;
const int NCCL_DOUBLE = 8, NCCL_SUM = 0;  // ncclFloat64, ncclSum (nccl.h)
}  // namespace

struct nwc_triples_ctx {
  Engine* eng = nullptr;
  HostState S;
  double *d_t1 = nullptr, *d_t2 = nullptr, *d_v2 = nullptr, *d_evl = nullptr, *d_red = nullptr;
  size_t n_t1 = 0, n_t2 = 0, n_v2 = 0;
  std::vector<Integer> klist;
  size_t batch_bytes = (size_t)8 << 30;
  ncclComm_t comm = nullptr;
  int nranks = 1;
  // sharded V2: block i of the V2 offset table lives on rank i % v2_nshards, compacted in table order;
  // remote shards are mapped with CUDA IPC and read over NVLink by the repack kernel / singles staging.
  int v2_nshards = 1, v2_rank = 0;
  std::vector<Integer> v2_shard_off;        // per table index: offset inside its owner's shard
  std::vector<double*> v2_peer;             // per rank: base of that rank's shard as seen from this process
  std::vector<char> v2_peer_opened;
  // `2eorb` storage: orbital-form integrals resident, spin-orbital blocks built per batch in the arena
  double* d_v2orb = nullptr;
  size_t n_v2orb = 0;
  std::unordered_map<Integer, const double*> v2_built;   // spin-orbital key -> block built since the last arena reset
};

namespace {

// position of `key` in a TCE offset table (1-based), -1 if absent
Integer hash_index(const Integer* hash, Integer key) {
  Integer n = hash[0], lo = 1, hi = n;
  while (lo <= hi) {
    Integer mid = (lo + hi) >> 1;
    if (hash[mid] == key) return mid;
    if (hash[mid] < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

const double* v2_block(const nwc_triples_ctx* c, Integer key, const char* what) {
  const HostState& S = c->S;
  if (c->v2_nshards <= 1) return c->d_v2 + hash_lookup_or_die(S.v2_hash, key, what);
  const Integer idx = hash_index(S.v2_hash.data(), key);
  if (idx < 0) { printf("nwc_triples: %s: block key %ld not found\n", what, key); fflush(stdout); exit(1); }
  const int owner = (int)((idx - 1) % c->v2_nshards);
  const double* base = c->v2_peer[owner];
  if (!base) { printf("nwc_triples: V2 shard of rank %d is not mapped (nwc_triples_v2_open_peers)\n", owner); fflush(stdout); exit(1); }
  return base + c->v2_shard_off[idx - 1];
}

// `2eorb`: <g3 g4||g1 g2> as one antisym job built from HostState::block_plan (get_block_ind.F:818-1538)
const double* orb_block_ptr(const nwc_triples_ctx* c, Integer key) {
  auto it = c->S.orb_off.find(key);
  if (it == c->S.orb_off.end()) { printf("nwc_triples: orbital V2 block key %ld is not resident\n", key); fflush(stdout); exit(1); }
  return c->d_v2orb + it->second;
}

const double* v2_block_2eorb(nwc_triples_ctx* c, Integer g3b, Integer g4b, Integer g1b, Integer g2b) {
  const HostState& S = c->S;
  const Integer skey = v2_key(S, g3b, g4b, g1b, g2b);
  auto hit = c->v2_built.find(skey);
  if (hit != c->v2_built.end()) return hit->second;
  AntisymJob j{};
  j.n[0] = (int)S.rg(g3b); j.n[1] = (int)S.rg(g4b); j.n[2] = (int)S.rg(g1b); j.n[3] = (int)S.rg(g2b);
  const size_t bytes = sizeof(double) * (size_t)j.n[0] * j.n[1] * j.n[2] * j.n[3];
  j.dst = (double*)c->eng->arena().alloc(bytes);
  const HostState::OrbPlan p = S.block_plan(g3b, g4b, g1b, g2b);
  if (p.key_a >= 0) { j.a = orb_block_ptr(c, p.key_a); j.ca = 1.0; for (int q = 0; q < 4; q++) j.sa[q] = p.sa[q]; }
  if (p.key_b >= 0) { j.b = orb_block_ptr(c, p.key_b); j.cb = -1.0; for (int q = 0; q < 4; q++) j.sb[q] = p.sb[q]; }
  c->eng->add_antisym(j);
  c->v2_built[skey] = j.dst;
  return j.dst;
}

// V2 block <g3 g4||g1 g2> (tile ids after tce_restricted_4), from whichever storage the context holds
const double* v2_operand(nwc_triples_ctx* c, Integer g3b, Integer g4b, Integer g1b, Integer g2b, const char* what) {
  if (c->S.intorb) return v2_block_2eorb(c, g3b, g4b, g1b, g2b);
  return v2_block(c, v2_key(c->S, g3b, g4b, g1b, g2b), what);
}

// the arena is about to be rewound: blocks built in it are gone
void reset_arena(nwc_triples_ctx* c) {
  c->eng->arena().reset();
  c->v2_built.clear();
}

struct NativeSink {
  nwc_triples_ctx* c;
  Engine& e;
  const HostState& S;

  void singles(const Row& r, Integer p4b_1, Integer h1b_1, Integer p5b_2, Integer p6b_2, Integer h2b_2, Integer h3b_2,
               const bool fire[9]) {
    OperandView t, v;
    // T1 block stored (p4,h1) with h1 fastest; the reference's TCE_SORT_2(2,1) becomes a stride swap
    t.base = c->d_t1 + hash_lookup_or_die(S.t1_hash, t1_key(S, p4b_1, h1b_1), "t1");
    t.stride[N_H1] = 1; t.stride[N_P4] = S.rg(r.h1b);
    // V2 block <p5 p6||h2 h3> stored (p5,p6,h2,h3), h3 fastest == v2sub(h3,h2,p6,p5)
    v.base = v2_operand(c, p5b_2, p6b_2, h2b_2, h3b_2, "v2(pphh)");
    v.stride[N_H3] = 1; v.stride[N_H2] = S.rg(r.h3b); v.stride[N_P6] = S.rg(r.h3b) * S.rg(r.h2b);
    v.stride[N_P5] = S.rg(r.h3b) * S.rg(r.h2b) * S.rg(r.p6b);
    for (int k = 0; k < 9; k++)
      if (fire[k]) e.add_singles(k, t, v);
  }

  void d1_pair(const Row& r, Integer h7b, const Integer am[4], const Integer bm[4], const bool fire[9]) {
    const Integer rp5 = S.rg(r.p5b), rh1 = S.rg(r.h1b), rh7 = S.rg(h7b);
    OperandView t, v;
    double sign;
    if (h7b < r.h1b) {  // block <p4 p5||h7 h1>, h1 fastest; TCE_SORT_4(4,2,1,3), factor -1 (tce_hashnsort.F:47-53)
      t.base = c->d_t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[3], am[2]), "t2");
      t.stride[N_H1] = 1; t.kstride = rh1; t.stride[N_P5] = rh7 * rh1; t.stride[N_P4] = rp5 * rh7 * rh1;
      sign = -1.0;
    } else {            // block <p4 p5||h1 h7>, h7 fastest; TCE_SORT_4(3,2,1,4), factor +1 (:55-62)
      t.base = c->d_t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[2], am[3]), "t2");
      t.kstride = 1; t.stride[N_H1] = rh7; t.stride[N_P5] = rh1 * rh7; t.stride[N_P4] = rp5 * rh1 * rh7;
      sign = 1.0;
    }
    // block <h7 p6||h2 h3> stored (h7,p6,h2,h3), h3 fastest == v2sub(h3,h2,p6,h7)  (:67-80)
    v.base = v2_operand(c, bm[1], bm[0], bm[2], bm[3], "v2(hphh)");
    v.stride[N_H3] = 1; v.stride[N_H2] = S.rg(r.h3b); v.stride[N_P6] = S.rg(r.h3b) * S.rg(r.h2b);
    v.kstride = S.rg(r.h3b) * S.rg(r.h2b) * S.rg(r.p6b);
    std::vector<PanelSlot> tc, vc;
    for (int k = 0; k < 9; k++)
      if (fire[k]) e.add_contraction(1, k, (int)rh7, t, v, sign, &tc, &vc);
  }

  void d2_pair(const Row& r, Integer p7b, const Integer am[4], const Integer bm[4], const bool fire[9]) {
    const Integer rp4 = S.rg(r.p4b), rp7 = S.rg(p7b), rh1 = S.rg(r.h1b), rh2 = S.rg(r.h2b);
    OperandView t, v;
    double sign;
    if (p7b < r.p4b) {  // block <p7 p4||h1 h2>; TCE_SORT_4(4,3,2,1), factor -1 (tce_hashnsort.F:129-135)
      t.base = c->d_t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[1], am[0], am[2], am[3]), "t2");
      t.stride[N_H2] = 1; t.stride[N_H1] = rh2; t.stride[N_P4] = rh1 * rh2; t.kstride = rp4 * rh1 * rh2;
      sign = -1.0;
    } else {            // block <p4 p7||h1 h2>; TCE_SORT_4(4,3,1,2), factor +1 (:137-144)
      t.base = c->d_t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[2], am[3]), "t2");
      t.stride[N_H2] = 1; t.stride[N_H1] = rh2; t.kstride = rh1 * rh2; t.stride[N_P4] = rp7 * rh1 * rh2;
      sign = 1.0;
    }
    // block <p5 p6||h3 p7> stored (p5,p6,h3,p7), p7 fastest == v2sub(p7,h3,p6,p5)  (:149-161)
    v.base = v2_operand(c, bm[0], bm[1], bm[2], bm[3], "v2(pphp)");
    v.kstride = 1; v.stride[N_H3] = rp7; v.stride[N_P6] = rp7 * S.rg(r.h3b);
    v.stride[N_P5] = rp7 * S.rg(r.h3b) * S.rg(r.p6b);
    std::vector<PanelSlot> tc, vc;
    for (int k = 0; k < 9; k++)
      if (fire[k]) e.add_contraction(2, k, (int)rp7, t, v, sign, &tc, &vc);
  }
};

void emit_tuple(nwc_triples_ctx* c, const Integer t[6]) {
  const HostState& S = c->S;
  int R[6];
  R[POS_P4] = (int)S.rg(t[0]); R[POS_P5] = (int)S.rg(t[1]); R[POS_P6] = (int)S.rg(t[2]);
  R[POS_H1] = (int)S.rg(t[3]); R[POS_H2] = (int)S.rg(t[4]); R[POS_H3] = (int)S.rg(t[5]);
  c->eng->begin_tuple(R);
  NativeSink sink{c, *c->eng, S};
  walk_singles(S, t, sink);
  walk_doubles(S, t, sink);
  const double* eps[6] = {c->d_evl + S.offset[t[3] - 1], c->d_evl + S.offset[t[4] - 1], c->d_evl + S.offset[t[5] - 1],
                          c->d_evl + S.offset[t[0] - 1], c->d_evl + S.offset[t[1] - 1], c->d_evl + S.offset[t[2] - 1]};
  c->eng->end_tuple(eps, tuple_factor(S, t));
}

int upload(double** dst, size_t* n_out, const double* src, size_t n, Engine* e) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  NWC_TRY(cudaMalloc((void**)dst, (n ? n : 1) * sizeof(double)));
  if (n) NWC_TRY(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
  *n_out = n;
  e->stats.h2d_bytes += n * sizeof(double);
  return 0;
}

size_t store_size(const Integer* hash, const HostState& S, int which) {
  // size = offset of last block + its size; recompute from keys
  const Integer n = hash[0];
  if (n == 0) return 0;
  Integer key = hash[n], off = hash[2 * n];
  Integer sz = 0;
  if (which == 1) { Integer h = key % S.noab + 1, p = key / S.noab + S.noab + 1; sz = S.rg(h) * S.rg(p); }
  else if (which == 2) {
    Integer h4 = key % S.noab + 1; key /= S.noab; Integer h3 = key % S.noab + 1; key /= S.noab;
    Integer p2 = key % S.nvab + S.noab + 1; key /= S.nvab; Integer p1 = key + S.noab + 1;
    sz = S.rg(p1) * S.rg(p2) * S.rg(h3) * S.rg(h4);
  } else {
    const Integer N = S.N();
    Integer g2 = key % N + 1; key /= N; Integer g1 = key % N + 1; key /= N; Integer g4 = key % N + 1; key /= N;
    Integer g3 = key + 1;
    sz = S.rg(g3) * S.rg(g4) * S.rg(g1) * S.rg(g2);
  }
  return (size_t)(off + sz);
}

}  // namespace

extern "C" {

const char* nwc_triples_last_error(void) { return g_err.c_str(); }

int nwc_triples_create(nwc_triples_ctx** out, int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { g_err = "no CUDA device (there is no CPU fallback)"; return 1; }
  if (device < 0 || device >= count) { g_err = "device index out of range"; return 1; }
  nwc_triples_ctx* c = new nwc_triples_ctx();
  c->eng = new Engine(device);
  *out = c;
  return 0;
}

int nwc_triples_destroy(nwc_triples_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->eng->device());
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (size_t r = 0; r < c->v2_peer.size(); r++)
    if (c->v2_peer_opened[r] && c->v2_peer[r]) cudaIpcCloseMemHandle(c->v2_peer[r]);
  cudaFree(c->d_t1); cudaFree(c->d_t2); cudaFree(c->d_v2); cudaFree(c->d_v2orb); cudaFree(c->d_evl); cudaFree(c->d_red);
  delete c->eng;
  delete c;
  return 0;
}

int nwc_triples_set_state(nwc_triples_ctx* c, const nwc_tce_state* st) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  c->S.load_tables(st);
  const HostState& S = c->S;
  if (upload(&c->d_t1, &c->n_t1, st->t1, store_size(st->t1_hash, S, 1), c->eng)) return 1;
  if (upload(&c->d_t2, &c->n_t2, st->t2, store_size(st->t2_hash, S, 2), c->eng)) return 1;
  if (upload(&c->d_v2, &c->n_v2, st->v2, store_size(st->v2_hash, S, 3), c->eng)) return 1;
  size_t ne;
  if (upload(&c->d_evl, &ne, S.evl.data(), S.evl.size(), c->eng)) return 1;
  c->v2_nshards = 1; c->v2_rank = 0;
  build_task_list(S, c->klist);
  return 0;
}

int nwc_triples_set_state_2eorb(nwc_triples_ctx* c, const nwc_tce_state* st, const nwc_tce_orb_state* orb) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  nwc_tce_state s2 = *st;
  s2.v2_hash = nullptr;   // not read in this mode
  c->S.load_tables(&s2);
  HostState& S = c->S;
  const std::string err = S.load_orbital(orb->noa, orb->nva, orb->b2am, orb->spin_alpha, orb->sym_alpha,
                                         orb->range_alpha, orb->v2orb_hash);
  if (!err.empty()) { g_err = err; return 1; }
  if (upload(&c->d_t1, &c->n_t1, st->t1, store_size(st->t1_hash, S, 1), c->eng)) return 1;
  if (upload(&c->d_t2, &c->n_t2, st->t2, store_size(st->t2_hash, S, 2), c->eng)) return 1;
  // only the blocks (T) can touch become resident, compacted run by run (the rest of d_v2orb stays on the host)
  if (c->d_v2orb) { cudaFree(c->d_v2orb); c->d_v2orb = nullptr; }
  NWC_TRY(cudaMalloc((void**)&c->d_v2orb, (size_t)(S.orb_size ? S.orb_size : 1) * sizeof(double)));
  for (const HostState::OrbRun& r : S.orb_runs)
    NWC_TRY(cudaMemcpy(c->d_v2orb + r.dst, orb->v2orb + r.src, (size_t)r.n * sizeof(double), cudaMemcpyHostToDevice));
  c->n_v2orb = (size_t)S.orb_size;
  c->eng->stats.h2d_bytes += (size_t)S.orb_size * sizeof(double);
  if (c->d_v2) { cudaFree(c->d_v2); c->d_v2 = nullptr; }
  c->n_v2 = 0;
  size_t ne;
  if (upload(&c->d_evl, &ne, S.evl.data(), S.evl.size(), c->eng)) return 1;
  c->v2_nshards = 1; c->v2_rank = 0;
  c->v2_built.clear();
  build_task_list(S, c->klist);
  return 0;
}

// Sharded variant: st->v2 points at THIS rank's shard only (its blocks, table order, compacted).
int nwc_triples_set_state_sharded(nwc_triples_ctx* c, const nwc_tce_state* st, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
  NWC_TRY(cudaSetDevice(c->eng->device()));
  c->S.load_tables(st);
  const HostState& S = c->S;
  if (upload(&c->d_t1, &c->n_t1, st->t1, store_size(st->t1_hash, S, 1), c->eng)) return 1;
  if (upload(&c->d_t2, &c->n_t2, st->t2, store_size(st->t2_hash, S, 2), c->eng)) return 1;
  // shard offsets of every block (all ranks compute the same table)
  const Integer n = S.v2_hash[0];
  const Integer total = (Integer)store_size(st->v2_hash, S, 3);
  c->v2_shard_off.assign((size_t)n, 0);
  std::vector<Integer> fill((size_t)nranks, 0);
  for (Integer i = 0; i < n; i++) {
    const Integer off = S.v2_hash[n + 1 + i], next = (i + 1 < n) ? S.v2_hash[n + 2 + i] : total;
    const int owner = (int)(i % nranks);
    c->v2_shard_off[(size_t)i] = fill[owner];
    fill[owner] += next - off;
  }
  if (upload(&c->d_v2, &c->n_v2, st->v2, (size_t)fill[rank], c->eng)) return 1;
  size_t ne;
  if (upload(&c->d_evl, &ne, S.evl.data(), S.evl.size(), c->eng)) return 1;
  c->v2_nshards = nranks; c->v2_rank = rank;
  c->v2_peer.assign((size_t)nranks, nullptr);
  c->v2_peer_opened.assign((size_t)nranks, 0);
  c->v2_peer[rank] = c->d_v2;
  build_task_list(S, c->klist);
  return 0;
}

// CUDA IPC handle (64 bytes) of this rank's V2 shard; the host all-gathers them (MPI/GA in NWChem)
int nwc_triples_v2_ipc_handle(nwc_triples_ctx* c, char handle64[64]) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  cudaIpcMemHandle_t h;
  NWC_TRY(cudaIpcGetMemHandle(&h, c->d_v2));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}

// map every peer's shard (handles = nranks x 64 bytes, rank order); peer reads then go over NVLink
int nwc_triples_v2_open_peers(nwc_triples_ctx* c, const char* handles) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  for (int r = 0; r < c->v2_nshards; r++) {
    if (r == c->v2_rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    NWC_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->v2_peer[(size_t)r] = (double*)p;
    c->v2_peer_opened[(size_t)r] = 1;
  }
  return 0;
}

// same-process alternative to the IPC exchange (several contexts in one process, tests): device pointer of this
// context's shard, and direct registration of a peer's shard pointer
void* nwc_triples_v2_shard_ptr(nwc_triples_ctx* c) { return c->d_v2; }
int nwc_triples_v2_set_peer_ptr(nwc_triples_ctx* c, int rank, void* dev_ptr) {
  if (rank < 0 || rank >= c->v2_nshards) { g_err = "bad peer rank"; return 1; }
  c->v2_peer[(size_t)rank] = (double*)dev_ptr;
  return 0;
}

Integer nwc_triples_num_tasks(nwc_triples_ctx* c) { return (Integer)(c->klist.size() / 7); }

int nwc_triples_task_list(nwc_triples_ctx* c, Integer* klist7) {
  memcpy(klist7, c->klist.data(), c->klist.size() * sizeof(Integer));
  return 0;
}

int nwc_triples_run(nwc_triples_ctx* c, Integer first, Integer stride, Integer max_tasks, double energy[2],
                    double* per_task) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  if (stride <= 0) stride = 1;
  const Integer nt = (Integer)(c->klist.size() / 7);
  Engine& e = *c->eng;
  energy[0] = energy[1] = 0.0;
  std::vector<double> eb;
  Integer done = 0, out_pos = 0;
  auto flush = [&]() {
    const int n = e.pending_tuples();
    if (n == 0) return;
    eb.assign(2 * (size_t)n, 0.0);
    e.run(eb.data());
    for (int i = 0; i < n; i++) {
      energy[0] += eb[2 * i];
      energy[1] += eb[2 * i + 1];
      if (per_task) { per_task[2 * (out_pos + i)] = eb[2 * i]; per_task[2 * (out_pos + i) + 1] = eb[2 * i + 1]; }
    }
    out_pos += n;
    reset_arena(c);
  };
  for (Integer k = first; k < nt && (max_tasks <= 0 || done < max_tasks); k += stride, done++) {
    emit_tuple(c, &c->klist[7 * k]);
    if (e.arena().used() >= c->batch_bytes || e.pending_items() > (size_t)32000000 || e.pending_tuples() >= 4096) flush();
  }
  flush();
  return 0;
}

int nwc_triples_run_restart(nwc_triples_ctx* c, Integer first, Integer stride, Integer* restart_begin, double* table,
                            double* table_bracket, Integer max_outer, double* t_energy) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  if (stride <= 0) stride = 1;
  if (*restart_begin < 1) *restart_begin = 1;
  const HostState& S = c->S;
  Engine& e = *c->eng;
  const Integer n0 = S.noab, n1 = S.noab + S.nvab;
  std::vector<double> eb;
  Integer done = 0;
  for (Integer p4 = n0 + *restart_begin; p4 <= n1; p4++) {
    if (max_outer > 0 && done >= max_outer) break;
    double en[2] = {0.0, 0.0};
    auto flush = [&]() {
      const int n = e.pending_tuples();
      if (n == 0) return;
      eb.assign(2 * (size_t)n, 0.0);
      e.run(eb.data());
      for (int i = 0; i < n; i++) { en[0] += eb[2 * i]; en[1] += eb[2 * i + 1]; }
      reset_arena(c);
    };
    Integer count = 0;   // position in this outer tile's loop order (ccsd_t_restart.F:120-150)
    for (Integer p5 = p4; p5 <= n1; p5++)
      for (Integer p6 = p5; p6 <= n1; p6++)
        for (Integer h1 = 1; h1 <= n0; h1++)
          for (Integer h2 = h1; h2 <= n0; h2++)
            for (Integer h3 = h2; h3 <= n0; h3++) {
              const Integer ps = S.sp(p4) + S.sp(p5) + S.sp(p6), hs = S.sp(h1) + S.sp(h2) + S.sp(h3);
              if (ps != hs) continue;
              if (S.restricted && ps + hs > 8) continue;
              if ((S.sy(p4) ^ S.sy(p5) ^ S.sy(p6) ^ S.sy(h1) ^ S.sy(h2) ^ S.sy(h3)) != 0) continue;
              const Integer k = count++;
              if (k < first || (k - first) % stride != 0) continue;
              const Integer t[6] = {p4, p5, p6, h1, h2, h3};
              emit_tuple(c, t);
              if (e.arena().used() >= c->batch_bytes || e.pending_items() > (size_t)32000000 || e.pending_tuples() >= 4096)
                flush();
            }
    flush();
    if (c->comm) {
      if (nwc_triples_allreduce_energy(c, en) != 0) return 1;
    }
    const Integer outer = p4 - n0;
    table[outer - 1] = en[1];
    if (table_bracket) table_bracket[outer - 1] = en[0];
    *restart_begin = outer + 1;
    done++;
  }
  *t_energy = 0.0;
  for (Integer i = 0; i < S.nvab; i++) *t_energy += table[i];
  return 0;
}

int nwc_triples_run_tuple(nwc_triples_ctx* c, const Integer t[6], double energy[2], double* host_doubles,
                          double* host_singles) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  Engine& e = *c->eng;
  emit_tuple(c, t);
  double *dd = nullptr, *ds = nullptr;
  size_t sz = 1;
  if (host_doubles) {
    for (int q = 0; q < 6; q++) sz *= (size_t)c->S.rg(t[q]);
    dd = (double*)e.arena().alloc(sz * sizeof(double));
    ds = (double*)e.arena().alloc(sz * sizeof(double));
    NWC_TRY(cudaMemsetAsync(dd, 0, sz * sizeof(double), e.stream()));
    NWC_TRY(cudaMemsetAsync(ds, 0, sz * sizeof(double), e.stream()));
  }
  double out[2] = {0, 0};
  e.run(out, dd, ds);
  if (host_doubles) {
    NWC_TRY(cudaMemcpy(host_doubles, dd, sz * sizeof(double), cudaMemcpyDeviceToHost));
    NWC_TRY(cudaMemcpy(host_singles, ds, sz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  reset_arena(c);
  energy[0] = out[0];
  energy[1] = out[1];
  return 0;
}

int nwc_triples_set_timing(nwc_triples_ctx* c, int on) { c->eng->timing = on != 0; return 0; }

int nwc_triples_get_stats(nwc_triples_ctx* c, nwc_triples_stats* o, int reset) {
  const EngineStats& s = c->eng->stats;
  o->fused_ms = s.fused_ms; o->repack_ms = s.repack_ms;
  o->fused_launches = s.fused_launches; o->repack_launches = s.repack_launches; o->reduce_launches = s.reduce_launches;
  o->work_items = s.work_items; o->descs = s.descs; o->tuples = s.tuples; o->flops = s.flops;
  o->h2d_bytes = (double)s.h2d_bytes; o->d2h_bytes = (double)s.d2h_bytes;
  o->resident_bytes = 8.0 * (double)(c->n_t1 + c->n_t2 + c->n_v2 + c->n_v2orb);
  if (reset) c->eng->stats = EngineStats();
  return 0;
}

static void fill_stats(const EngineStats& s, nwc_triples_stats* o, double resident) {
  o->fused_ms = s.fused_ms; o->repack_ms = s.repack_ms;
  o->fused_launches = s.fused_launches; o->repack_launches = s.repack_launches; o->reduce_launches = s.reduce_launches;
  o->work_items = s.work_items; o->descs = s.descs; o->tuples = s.tuples; o->flops = s.flops;
  o->h2d_bytes = (double)s.h2d_bytes; o->d2h_bytes = (double)s.d2h_bytes;
  o->resident_bytes = resident;
}
int nwc_triples_timer_start(nwc_triples_ctx* c) { cudaSetDevice(c->eng->device()); c->eng->timer_start(); return 0; }
int nwc_triples_timer_stop_ms(nwc_triples_ctx* c, double* ms) { cudaSetDevice(c->eng->device()); *ms = c->eng->timer_stop_ms(); return 0; }
int nwc_host_register(void* ptr, size_t bytes) { NWC_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault)); return 0; }
int nwc_host_unregister(void* ptr) { NWC_TRY(cudaHostUnregister(ptr)); return 0; }
int nwc_compat_get_stats(nwc_triples_stats* o, int reset) {
  Engine& e = nwc::compat_engine();
  fill_stats(e.stats, o, 0.0);
  if (reset) e.stats = EngineStats();
  return 0;
}
int nwc_compat_set_timing(int on) { nwc::compat_engine().timing = on != 0; return 0; }
int nwc_compat_timer_start(void) { nwc::compat_engine().timer_start(); return 0; }
int nwc_compat_timer_stop_ms(double* ms) { *ms = nwc::compat_engine().timer_stop_ms(); return 0; }

// debugging aid (not part of the public header): per-CTA phase clocks of the next launches; cap_items = 0 turns it off
int nwc_debug_phase_timing(unsigned long long* host_out, unsigned int cap_items, int fetch) {
  static unsigned long long* d_buf = nullptr;
  static unsigned int cap = 0;
  if (fetch) {
    if (!d_buf) return 1;
    NWC_TRY(cudaDeviceSynchronize());
    NWC_TRY(cudaMemcpy(host_out, d_buf, (size_t)cap * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
  }
  if (d_buf) { cudaFree(d_buf); d_buf = nullptr; }
  cap = cap_items;
  if (cap_items) {
    NWC_TRY(cudaMalloc((void**)&d_buf, (size_t)cap_items * 8 * sizeof(unsigned long long)));
    NWC_TRY(cudaMemset(d_buf, 0, (size_t)cap_items * 8 * sizeof(unsigned long long)));
  }
  nwc::set_phase_timing(d_buf, cap_items);
  return 0;
}

int nwc_triples_set_batch_bytes(nwc_triples_ctx* c, size_t bytes) { c->batch_bytes = bytes; return 0; }

int nwc_triples_nccl_unique_id(char id128[128]) {
  if (!g_nccl.load()) return 1;
  ncclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) { g_err = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return 1; }
  memcpy(id128, id.internal, 128);
  return 0;
}

int nwc_triples_nccl_init(nwc_triples_ctx* c, const char id128[128], int rank, int nranks) {
  if (!g_nccl.load()) return 1;
  NWC_TRY(cudaSetDevice(c->eng->device()));
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) { g_err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return 1; }
  c->nranks = nranks;
  if (!c->d_red) NWC_TRY(cudaMalloc((void**)&c->d_red, 2 * sizeof(double)));
  return 0;
}

int nwc_triples_allreduce_energy(nwc_triples_ctx* c, double energy[2]) {
  if (!c->comm) { if (c->nranks == 1) return 0; g_err = "nccl not initialised"; return 1; }
  NWC_TRY(cudaSetDevice(c->eng->device()));
  cudaStream_t s = c->eng->stream();
  NWC_TRY(cudaMemcpyAsync(c->d_red, energy, 2 * sizeof(double), cudaMemcpyHostToDevice, s));
  int r = g_nccl.AllReduce(c->d_red, c->d_red, 2, NCCL_DOUBLE, NCCL_SUM, c->comm, s);  // replaces ga_dgop (ccsd_t.F:297)
  if (r != 0) { g_err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return 1; }
  NWC_TRY(cudaMemcpyAsync(energy, c->d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  NWC_TRY(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"
