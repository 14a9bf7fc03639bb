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

// A block store dealt block-wise over the ranks: block i of its offset table lives on rank i % nshards, compacted in
// table order; remote blocks are pulled whole into the batch arena (pull_kernel), cached per batch slot.  The V2 store
// has its own, older copy of this bookkeeping in the context (v2_*); this one serves the CR-CCSD(T) pphp intermediate.
struct PeerStore {
  int nshards = 1, rank = 0;
  std::vector<Integer> shard_off, block_n;
  std::vector<double*> peer;
  std::vector<char> opened;
  std::unordered_map<Integer, const double*> pulled[2];
  void reset() {
    for (size_t r = 0; r < peer.size(); r++)
      if (opened[r] && peer[r]) cudaIpcCloseMemHandle(peer[r]);
    nshards = 1; rank = 0;
    shard_off.clear(); block_n.clear(); peer.clear(); opened.clear();
    pulled[0].clear(); pulled[1].clear();
  }
};

struct nwc_triples_ctx {
  Engine* eng = nullptr;
  HostState S;
  double *d_t1 = nullptr, *d_t2 = nullptr, *d_v2 = nullptr, *d_evl = nullptr, *d_red = nullptr;
  size_t n_t1 = 0, n_t2 = 0, n_v2 = 0, n_red = 0;
  std::vector<Integer> klist;
  size_t batch_bytes = (size_t)8 << 30;
  ncclComm_t comm = nullptr;
  int nranks = 1;
  // Sharded V2: block i of the store -- entry i of the spin-orbital offset table, or the i-th orbital block
  // (T) can touch in `2eorb` form -- lives on rank i % v2_nshards, compacted in that order.  Remote shards are mapped
  // with CUDA IPC; the blocks a batch needs are pulled whole over NVLink into its arena (pull_kernel), which replaces
  // the ga_get per tile of get_block.F:79-81.
  int v2_nshards = 1, v2_rank = 0;
  std::vector<Integer> v2_shard_off;        // per block index: offset inside its owner's shard
  std::vector<Integer> v2_block_n;          // per block index: doubles
  std::vector<double*> v2_peer;             // per rank: base of that rank's shard as seen from this process
  std::vector<char> v2_peer_opened;
  bool peer_direct = false;                 // NWC_PEER_DIRECT=1: read remote blocks in place (strided 8-byte loads), for A/B runs
  // `2eorb` storage: orbital-form integrals resident, spin-orbital blocks built per batch in the arena
  double* d_v2orb = nullptr;
  size_t n_v2orb = 0;
  // Lambda-CCSD(T) inputs (nwc_triples_set_lambda): lambda_1 (h,p), lambda_2 (hh,pp), Fock (h,p) blocks, replicated
  double *d_y1 = nullptr, *d_y2 = nullptr, *d_f1 = nullptr;
  size_t n_y1 = 0, n_y2 = 0, n_f1 = 0;
  std::vector<Integer> y1_hash, y2_hash, f1_hash;
  // CR-CCSD(T) inputs (nwc_triples_set_cr): the three intermediates of cr_ccsd_t.F's tuple loop, replicated
  double *d_crn1 = nullptr, *d_crn2 = nullptr, *d_cre2 = nullptr;
  size_t n_crn1 = 0, n_crn2 = 0, n_cre2 = 0;
  std::vector<Integer> crn1_hash, crn2_hash, cre2_hash;
  std::vector<double> cre2_scaled;          // trace contexts keep host data by reference: the 2/3-scaled copy lives here
  PeerStore crn2s;                          // nwc_triples_set_cr_sharded: the pphp intermediate (the size of V2's <pp||hp> class) dealt over the ranks
  // CR-EOMCCSD(T) inputs (nwc_triples_set_creom): x amplitudes, the four moment intermediates, and two combined stores
  // for the left-hand outer products (eomy1 = r0*t1 + x1, eomz = r0*(2/3)*i1_tt + 2*i1_xt); eps vectors for the shifted
  // and the unit denominators
  double *d_x2 = nullptr, *d_m1 = nullptr, *d_m2 = nullptr, *d_m3 = nullptr, *d_m4 = nullptr, *d_eomy1 = nullptr, *d_eomz = nullptr;
  double *d_evl_shift = nullptr, *d_unit = nullptr, *d_zero = nullptr;
  size_t n_x2 = 0, n_m1 = 0, n_m2 = 0, n_m3 = 0, n_m4 = 0, n_eomy1 = 0, n_eomz = 0, n_evl_shift = 0, n_unit = 0, n_zero = 0;
  std::vector<Integer> x2_hash, m1_hash, m2_hash, m3_hash, m4_hash, eomz_hash;
  std::vector<double> eom_host[5];          // trace contexts: combined stores and eps vectors live here
  double eom_r0 = 0.0, eom_excit = 0.0;
  bool eom_lr0 = false, eom_set = false;
  // per batch slot (engine.h): blocks that live in that slot's arena
  std::unordered_map<Integer, const double*> v2_built[2];   // spin-orbital key -> antisymmetrised block
  std::unordered_map<Integer, const double*> pulled[2];     // block index -> local copy of a remote block
};

namespace {

void free_stores(nwc_triples_ctx* c) {   // one reset routine for every set_state* variant
  if (c->eng->trace_only()) {   // host-only trace context: every store pointer is borrowed from the caller
    c->eng->abort();
    c->d_t1 = c->d_t2 = c->d_v2 = c->d_v2orb = c->d_evl = c->d_y1 = c->d_y2 = c->d_f1 = nullptr;
    c->d_crn1 = c->d_crn2 = c->d_cre2 = nullptr;
    c->d_x2 = c->d_m1 = c->d_m2 = c->d_m3 = c->d_m4 = c->d_eomy1 = c->d_eomz = c->d_evl_shift = c->d_unit = c->d_zero = nullptr;
    c->eom_set = false;
    c->y1_hash.clear(); c->y2_hash.clear(); c->f1_hash.clear();
    c->crn1_hash.clear(); c->crn2_hash.clear(); c->cre2_hash.clear();
    c->klist.clear();
    return;
  }
  cudaSetDevice(c->eng->device());
  c->eng->abort();
  for (size_t r = 0; r < c->v2_peer.size(); r++)
    if (c->v2_peer_opened[r] && c->v2_peer[r]) cudaIpcCloseMemHandle(c->v2_peer[r]);
  c->v2_peer.clear(); c->v2_peer_opened.clear(); c->v2_shard_off.clear(); c->v2_block_n.clear();
  c->v2_nshards = 1; c->v2_rank = 0;
  cudaFree(c->d_t1); cudaFree(c->d_t2); cudaFree(c->d_v2); cudaFree(c->d_v2orb); cudaFree(c->d_evl);
  cudaFree(c->d_y1); cudaFree(c->d_y2); cudaFree(c->d_f1);
  c->crn2s.reset();
  cudaFree(c->d_crn1); cudaFree(c->d_crn2); cudaFree(c->d_cre2);
  cudaFree(c->d_x2); cudaFree(c->d_m1); cudaFree(c->d_m2); cudaFree(c->d_m3); cudaFree(c->d_m4); cudaFree(c->d_eomy1);
  cudaFree(c->d_eomz); cudaFree(c->d_evl_shift); cudaFree(c->d_unit); cudaFree(c->d_zero);
  c->d_x2 = c->d_m1 = c->d_m2 = c->d_m3 = c->d_m4 = c->d_eomy1 = c->d_eomz = c->d_evl_shift = c->d_unit = c->d_zero = nullptr;
  c->n_x2 = c->n_m1 = c->n_m2 = c->n_m3 = c->n_m4 = c->n_eomy1 = c->n_eomz = c->n_evl_shift = c->n_unit = c->n_zero = 0;
  c->x2_hash.clear(); c->m1_hash.clear(); c->m2_hash.clear(); c->m3_hash.clear(); c->m4_hash.clear();
  c->eom_set = false;
  c->d_t1 = c->d_t2 = c->d_v2 = c->d_v2orb = c->d_evl = c->d_y1 = c->d_y2 = c->d_f1 = nullptr;
  c->d_crn1 = c->d_crn2 = c->d_cre2 = nullptr;
  c->n_t1 = c->n_t2 = c->n_v2 = c->n_v2orb = c->n_y1 = c->n_y2 = c->n_f1 = c->n_crn1 = c->n_crn2 = c->n_cre2 = 0;
  c->y1_hash.clear(); c->y2_hash.clear(); c->f1_hash.clear();
  c->crn1_hash.clear(); c->crn2_hash.clear(); c->cre2_hash.clear();
  for (int s = 0; s < 2; s++) { c->v2_built[s].clear(); c->pulled[s].clear(); }
  c->S.intorb = false; c->S.orb_off.clear(); c->S.host_off.clear(); c->S.orb_runs.clear(); c->S.orb_blocks.clear();
  c->S.orb_index.clear(); c->S.orb_size = c->S.orb_host_size = 0;
  c->klist.clear();
  const char* pd = getenv("NWC_PEER_DIRECT");
  c->peer_direct = pd && *pd == '1';
}

void recover(nwc_triples_ctx* c) {
  if (!c || !c->eng) return;
  c->eng->abort();
  for (int s = 0; s < 2; s++) { c->v2_built[s].clear(); c->pulled[s].clear(); c->crn2s.pulled[s].clear(); }
}

template <class F>
int guarded(nwc_triples_ctx* c, F&& f) {
  try {
    return f();
  } catch (const std::exception& ex) {
    g_err = ex.what();
    recover(c);
    return 1;
  }
}

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

// block `idx` of the sharded store: local pointer if this rank owns it, else its pulled copy in the current batch arena
const double* sharded_block(nwc_triples_ctx* c, Integer idx, double* local_base) {
  const int owner = (int)(idx % c->v2_nshards);
  const Integer off = c->v2_shard_off[(size_t)idx];
  if (owner == c->v2_rank) return local_base + off;
  const double* base = c->v2_peer[(size_t)owner];
  if (!base) throw Error("nwc_triples: V2 shard of rank " + std::to_string(owner) + " is not mapped (nwc_triples_v2_open_peers)");
  if (c->peer_direct) return base + off;
  auto& cache = c->pulled[c->eng->current_slot()];
  auto hit = cache.find(idx);
  if (hit != cache.end()) return hit->second;
  const Integer n = c->v2_block_n[(size_t)idx];
  double* dst = (double*)c->eng->arena().alloc((size_t)n * sizeof(double));
  c->eng->add_copy(CopyJob{base + off, dst, (long long)n});
  cache[idx] = dst;
  return dst;
}

const double* v2_block(nwc_triples_ctx* c, Integer key, const char* what) {
  const HostState& S = c->S;
  if (c->v2_nshards <= 1) return c->d_v2 + hash_lookup_or_die(S.v2_hash, key, what);
  const Integer idx = hash_index(S.v2_hash.data(), key);
  if (idx < 0) throw Error(std::string("nwc_triples: ") + what + ": block key " + std::to_string(key) + " not found");
  return sharded_block(c, idx - 1, c->d_v2);
}

// `2eorb`: <g3 g4||g1 g2> as one antisym job built from HostState::block_plan (get_block_ind.F:818-1538)
const double* orb_block_ptr(nwc_triples_ctx* c, Integer key) {
  if (c->v2_nshards <= 1) {
    auto it = c->S.orb_off.find(key);
    if (it == c->S.orb_off.end()) throw Error("nwc_triples: orbital V2 block key " + std::to_string(key) + " is not resident");
    return c->d_v2orb + it->second;
  }
  auto it = c->S.orb_index.find(key);
  if (it == c->S.orb_index.end()) throw Error("nwc_triples: orbital V2 block key " + std::to_string(key) + " is not resident");
  return sharded_block(c, it->second, c->d_v2orb);
}

const double* v2_block_2eorb(nwc_triples_ctx* c, Integer g3b, Integer g4b, Integer g1b, Integer g2b) {
  const HostState& S = c->S;
  const Integer skey = v2_key(S, g3b, g4b, g1b, g2b);
  auto& built = c->v2_built[c->eng->current_slot()];
  auto hit = built.find(skey);
  if (hit != built.end()) return hit->second;
  AntisymJob j{};
  j.n[0] = (int)S.rg(g3b); j.n[1] = (int)S.rg(g4b); j.n[2] = (int)S.rg(g1b); j.n[3] = (int)S.rg(g2b);
  const size_t bytes = sizeof(double) * (size_t)j.n[0] * j.n[1] * j.n[2] * j.n[3];
  const HostState::OrbPlan p = S.block_plan(g3b, g4b, g1b, g2b);
  if (p.key_a >= 0) { j.a = orb_block_ptr(c, p.key_a); j.ca = 1.0; for (int q = 0; q < 4; q++) j.sa[q] = p.sa[q]; }
  if (p.key_b >= 0) { j.b = orb_block_ptr(c, p.key_b); j.cb = -1.0; for (int q = 0; q < 4; q++) j.sb[q] = p.sb[q]; }
  j.dst = (double*)c->eng->arena().alloc(bytes);
  c->eng->add_antisym(j);
  built[skey] = j.dst;
  return j.dst;
}

// V2 block <g3 g4||g1 g2> (tile ids after tce_restricted_4), from whichever storage the context holds
const double* v2_operand(nwc_triples_ctx* c, Integer g3b, Integer g4b, Integer g1b, Integer g2b, const char* what) {
  if (c->S.intorb) return v2_block_2eorb(c, g3b, g4b, g1b, g2b);
  return v2_block(c, v2_key(c->S, g3b, g4b, g1b, g2b), what);
}

// block `key` of the CR-CCSD(T) pphp intermediate i1(p4 p5 h1 p12): replicated, or this rank's shard / a pulled copy
const double* crn2_block(nwc_triples_ctx* c, Integer key) {
  PeerStore& P = c->crn2s;
  if (P.nshards <= 1) return c->d_crn2 + hash_lookup_or_die(c->crn2_hash, key, "cr n2(pphp)");
  const Integer pos = hash_index(c->crn2_hash.data(), key);
  if (pos < 0) throw Error("nwc_triples: cr n2(pphp): block key " + std::to_string(key) + " not found");
  const Integer idx = pos - 1;
  const int owner = (int)(idx % P.nshards);
  const Integer off = P.shard_off[(size_t)idx];
  if (owner == P.rank) return c->d_crn2 + off;
  const double* base = P.peer[(size_t)owner];
  if (!base) throw Error("nwc_triples: CR shard of rank " + std::to_string(owner) + " is not mapped (nwc_triples_cr_open_peers)");
  auto& cache = P.pulled[c->eng->current_slot()];
  auto hit = cache.find(idx);
  if (hit != cache.end()) return hit->second;
  const Integer n = P.block_n[(size_t)idx];
  double* dst = (double*)c->eng->arena().alloc((size_t)n * sizeof(double));
  c->eng->add_copy(CopyJob{base + off, dst, (long long)n});
  cache[idx] = dst;
  return dst;
}

// a batch slot has been collected: blocks built / pulled in its arena are gone
void slot_done(nwc_triples_ctx* c, int slot) {
  if (slot < 0) return;
  c->v2_built[slot].clear();
  c->pulled[slot].clear();
  c->crn2s.pulled[slot].clear();
}

// lambda_2 block key (h4b<=h5b, p1b<=p2b): lambda_ccsd_t_left.F:378-380
inline Integer y2_key(const HostState& S, Integer h4, Integer h5, Integer p1, Integer p2) {
  return p2 - S.noab - 1 + S.nvab * (p1 - S.noab - 1 + S.nvab * (h5 - 1 + S.noab * (h4 - 1)));
}

// The walkers of host_driver.h report the operand pairs of the (T) right-hand side.  With `lam` set the same walk
// produces the Lambda-CCSD(T) LEFT-hand tiles: term by term, lambda_ccsd_t_left_1/_3/_4 are ccsd_t_singles /
// ccsd_t_doubles with t1(p,h) -> lambda_1(h,p), t2(pp,hh) -> lambda_2(hh,pp) and every V2 block replaced by its
// bra<->ket transpose, which for real orbitals is the same block the (T) path already holds (same index permutations,
// same signs: compare lambda_ccsd_t_left.F:6-9 with ccsd_t_doubles.F / ccsd_t_singles.F).  So only the amplitude views
// change: where they come from (Y stores), their block keys and their element strides (transposed layout).
struct NativeSink {
  nwc_triples_ctx* c;
  Engine& e;
  const HostState& S;
  bool lam = false;          // amplitudes from the lambda stores (left-hand side)
  int side = 0;              // 0: the tuple's doubles tile; 1: the left-hand doubles tile of a two-sided (Lambda) tuple
  bool want_singles = true;  // (T) right-hand side of Lambda-CCSD(T) uses the doubles only (lambda_ccsd_t.F:109-111)
  bool want_doubles = true;
  // CR-CCSD(T) (cr_ccsd_t.F:139-152).  CR_MOMENT: the walk of the (T) doubles produces the moment tile -- cr_ccsd_t_N_1 /
  // _N_2 are ccsd_t_doubles with the <hp||hh> / <pp||hp> blocks replaced by the dressed intermediates (same rows, filters,
  // T2 fetches, dispatch tests and kernel layouts/signs; only the store, its key and, for N_1, the element order differ).
  // CR_DENOM: the walk of the (T) singles produces cr_ccsd_t_E_2 (t1 x i1_tt, -2/3) and walk_cr_e1 produces cr_ccsd_t_E_1,
  // both as outer products bound to the side-0 tile.
  // CR_EOM_RIGHT / CR_EOM_LEFT: the tiles of CR-EOMCCSD(T) (cr_eomccsd_t.F:377-419).  Its per-tuple routines are the CR-CCSD(T)
  // ones with other operands and constant factors (creomccsd_t_n2_mem.F:674,:5665,:9657,:12905; q3rexpt2.F:80,:414), so one
  // walk of the (T) doubles emits, per contracted tile, the segments  r0*t2*i1 (if r0 != 0), f*t2*i2_{1|2}, x2*i2_{3|4}
  // (f = 1 for the Sum(h) family, -2 for the Sum(p) family) -- concatenated along K in one panel pair -- and the left tile
  // is t2 x (r0*t1 + x1) through walk_cr_e1 plus t1 x (r0*2/3*i1_tt + 2*i1_xt) through walk_singles.
  enum { CR_OFF = 0, CR_MOMENT = 1, CR_DENOM = 2, CR_EOM_RIGHT = 3, CR_EOM_LEFT = 4 };
  int cr = CR_OFF;
  int op_target = Engine::OP_SIDE0;   // where CR_DENOM / CR_EOM_LEFT outer products go
  int op_mask = 3;                    // CR_EOM_LEFT: bit 0 = the walk_cr_e1 family, bit 1 = the walk_singles family
  // contracted tiles of the current row, concatenated along K when the row ends (engine.h Segment)
  std::vector<Segment> segs;
  bool row_fire[9] = {false, false, false, false, false, false, false, false, false};
  void push(const OperandView& t, const OperandView& v, double sign, Integer K, const bool fire[9]) {
    Segment sg;
    sg.K = (int)K; sg.t = t; sg.v = v; sg.tscale = sign;
    segs.push_back(sg);
    for (int k = 0; k < 9; k++) row_fire[k] = fire[k];
  }
  void row_end(int family) {
    if (!segs.empty()) {
      std::vector<GroupPanel> tc, vc;   // kernels fired by the same operand list (diagonal tuples) share panels
      for (int k = 0; k < 9; k++)
        if (row_fire[k]) e.add_contraction_group(family, k, segs.data(), (int)segs.size(), &tc, &vc, side);
    }
    segs.clear();
  }
  // two-particle amplitude block over the tiles (pA<=pB | hC<=hD) (ids after tce_restricted_4; ranges rA..rD):
  // base pointer and the element strides of pA, pB, hC, hD
  const double* amp2(Integer pA, Integer pB, Integer hC, Integer hD, Integer rA, Integer rB, Integer rC, Integer rD,
                     long long st[4]) const {
    if (!lam) {   // T2 block [pA][pB][hC][hD], hD fastest (tce_t2_offset_new.F)
      st[3] = 1; st[2] = rD; st[1] = rC * rD; st[0] = rB * rC * rD;
      return c->d_t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, pA, pB, hC, hD), "t2");
    }
    // lambda_2 block [hC][hD][pA][pB], pB fastest
    st[1] = 1; st[0] = rB; st[3] = rA * rB; st[2] = rD * rA * rB;
    return c->d_y2 + hash_lookup_or_die(c->y2_hash, y2_key(S, hC, hD, pA, pB), "lambda2");
  }

  void singles(const Row& r, Integer p4b_1, Integer h1b_1, Integer p5b_2, Integer p6b_2, Integer h2b_2, Integer h3b_2,
               const bool fire[9]) {
    if (!want_singles) return;
    OperandView t, v;
    if (!lam) {
      // T1 block stored (p4,h1) with h1 fastest; the reference's TCE_SORT_2(2,1) becomes a stride swap
      t.base = c->d_t1 + hash_lookup_or_die(S.t1_hash, t1_key(S, p4b_1, h1b_1), "t1");
      t.stride[N_H1] = 1; t.stride[N_P4] = S.rg(r.h1b);
    } else {
      // lambda_1 block stored (h4,p1), p1 fastest; key p1b-noab-1 + nvab*(h4b-1) (lambda_ccsd_t_left.F:154-155)
      t.base = c->d_y1 + hash_lookup_or_die(c->y1_hash, p4b_1 - S.noab - 1 + S.nvab * (h1b_1 - 1), "lambda1");
      t.stride[N_P4] = 1; t.stride[N_H1] = S.rg(r.p4b);
    }
    // V2 block <p5 p6||h2 h3> stored (p5,p6,h2,h3), h3 fastest == v2sub(h3,h2,p6,p5)
    v.stride[N_H3] = 1; v.stride[N_H2] = S.rg(r.h3b); v.stride[N_P6] = S.rg(r.h3b) * S.rg(r.h2b);
    v.stride[N_P5] = S.rg(r.h3b) * S.rg(r.h2b) * S.rg(r.p6b);
    if (cr == CR_DENOM || cr == CR_EOM_LEFT) {
      // cr_ccsd_t_E_2: i1(p5 p6 h2 h3)_tt block, same layout, key as a T2 block (cr_ccsd_t_E.F:605-608); sd_E2_K adds
      // twot * t1sub * v2sub with twot = -2/3 * (sign of sd_t_s1_K) (:629-:721).  The store is resident pre-scaled by
      // 2/3 (nwc_triples_set_cr), so only the sign is left.  CR-EOMCCSD(T): the combined store of nwc_triples_set_creom.
      if (cr == CR_EOM_LEFT && !(op_mask & 2)) return;
      const double* zstore = cr == CR_DENOM ? c->d_cre2 : c->d_eomz;
      v.base = zstore + hash_lookup_or_die(cr == CR_DENOM ? c->cre2_hash : c->eomz_hash, t2_key(S, p5b_2, p6b_2, h2b_2, h3b_2), "cr e2(pphh)");
      for (int k = 0; k < 9; k++)
        if (fire[k]) {
          int sa[6], sb[6];
          for (int q = 0; q < 6; q++) { sa[q] = (int)t.stride[DECL[0][k][q]]; sb[q] = (int)v.stride[DECL[0][k][q]]; }
          e.add_outer_product(t.base, sa, v.base, sb, SIGN[0][k] > 0, op_target);
        }
      return;
    }
    v.base = v2_operand(c, p5b_2, p6b_2, h2b_2, h3b_2, "v2(pphh)");
    for (int k = 0; k < 9; k++)
      if (fire[k]) e.add_singles(k, t, v);
  }

  // cr_ccsd_t_E_1 (cr_ccsd_t_E.F:261-277, kernels sd_E_K): t2sub(p4,p5,h1,h2) = the stored T2 block <p4 p5||h1 h2>
  // (h2 fastest; the reference's TCE_SORT_4(4,3,2,1) only reverses the index order), t1sub(p6,h3) = the stored T1 block
  void cr_e1(const Row& r, const Integer am[4], const Integer bm[2], const bool fire[9]) {
    if (cr == CR_EOM_LEFT && !(op_mask & 1)) return;
    OperandView a, b;
    // CR-EOMCCSD(T): r0 * cr_ccsd_t_E_1 (t2 x t1) + q3rexpt2_1 (t2 x x1) = t2 x (r0*t1 + x1), the combined T1-like store
    a.base = (cr == CR_EOM_LEFT ? c->d_eomy1 : c->d_t1) + hash_lookup_or_die(S.t1_hash, t1_key(S, bm[0], bm[1]), "t1");
    a.stride[N_H3] = 1; a.stride[N_P6] = S.rg(r.h3b);
    b.base = c->d_t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[2], am[3]), "t2");
    b.stride[N_H2] = 1; b.stride[N_H1] = S.rg(r.h2b); b.stride[N_P5] = S.rg(r.h1b) * S.rg(r.h2b);
    b.stride[N_P4] = S.rg(r.p5b) * S.rg(r.h1b) * S.rg(r.h2b);
    for (int k = 0; k < 9; k++)
      if (fire[k]) {
        int sa[6], sb[6];
        for (int q = 0; q < 6; q++) { sa[q] = (int)a.stride[DECL_E1[k][q]]; sb[q] = (int)b.stride[DECL_E1[k][q]]; }
        e.add_outer_product(a.base, sa, b.base, sb, SIGN_E1[k] < 0, op_target);
      }
  }

  void d1_pair(const Row& r, Integer h7b, const Integer am[4], const Integer bm[4], const bool fire[9]) {
    if (!want_doubles) return;
    const Integer rp4 = S.rg(r.p4b), rp5 = S.rg(r.p5b), rh1 = S.rg(r.h1b), rh7 = S.rg(h7b);
    OperandView t, v;
    double sign;
    long long st[4];
    if (h7b < r.h1b) {  // block <p4 p5||h7 h1>; TCE_SORT_4(4,2,1,3), factor -1 (tce_hashnsort.F:47-53; left_3 :614-621)
      t.base = amp2(am[0], am[1], am[3], am[2], rp4, rp5, rh7, rh1, st);
      t.stride[N_P4] = st[0]; t.stride[N_P5] = st[1]; t.kstride = st[2]; t.stride[N_H1] = st[3];
      sign = -1.0;
    } else {            // block <p4 p5||h1 h7>; TCE_SORT_4(3,2,1,4), factor +1 (:55-62; left_3 :622-629)
      t.base = amp2(am[0], am[1], am[2], am[3], rp4, rp5, rh1, rh7, st);
      t.stride[N_P4] = st[0]; t.stride[N_P5] = st[1]; t.stride[N_H1] = st[2]; t.kstride = st[3];
      sign = 1.0;
    }
    if (cr == CR_MOMENT || cr == CR_EOM_RIGHT) {
      // i1(h7 p6 h2 h3) of cr_ccsd_t_N_1, stored (p6,h7,h2,h3), h3 fastest == v2sub(h3,h2,h7,p6) of sd_t_cr1_K; key
      // h3-1 + noab*(h2-1 + noab*(h7-1 + noab*(p6-noab-1))) (cr_ccsd_t_N.F:509-512)
      const Integer key = bm[3] - 1 + S.noab * (bm[2] - 1 + S.noab * (bm[1] - 1 + S.noab * (bm[0] - S.noab - 1)));
      v.stride[N_H3] = 1; v.stride[N_H2] = S.rg(r.h3b); v.kstride = S.rg(r.h3b) * S.rg(r.h2b);
      v.stride[N_P6] = S.rg(r.h3b) * S.rg(r.h2b) * rh7;
      if (cr == CR_EOM_RIGHT) {
        // creomsd_t_n2_mem_1 (t2 x i2_1), _3 (x2 x i2_3), and r0 * cr_ccsd_t_N_1: same T2-type fetch (x2 has the T2 block
        // structure), same intermediate layout and key
        OperandView tx = t;
        tx.base = c->d_x2 + (t.base - c->d_t2);   // same block, same offset: the x2 offset table is checked equal to T2's
        if (c->eom_lr0) { v.base = c->d_crn1 + hash_lookup_or_die(c->crn1_hash, key, "cr n1(phhh)"); push(t, v, sign * c->eom_r0, rh7, fire); }
        v.base = c->d_m1 + hash_lookup_or_die(c->m1_hash, key, "creom i2_1(phhh)"); push(t, v, sign, rh7, fire);
        v.base = c->d_m3 + hash_lookup_or_die(c->m3_hash, key, "creom i2_3(phhh)"); push(tx, v, sign, rh7, fire);
        return;
      }
      v.base = c->d_crn1 + hash_lookup_or_die(c->crn1_hash, key, "cr n1(phhh)");
    } else {
      // block <h7 p6||h2 h3> stored (h7,p6,h2,h3), h3 fastest == v2sub(h3,h2,p6,h7)  (:67-80)
      v.base = v2_operand(c, bm[1], bm[0], bm[2], bm[3], "v2(hphh)");
      v.stride[N_H3] = 1; v.stride[N_H2] = S.rg(r.h3b); v.stride[N_P6] = S.rg(r.h3b) * S.rg(r.h2b);
      v.kstride = S.rg(r.h3b) * S.rg(r.h2b) * S.rg(r.p6b);
    }
    push(t, v, sign, rh7, fire);
  }

  void d2_pair(const Row& r, Integer p7b, const Integer am[4], const Integer bm[4], const bool fire[9]) {
    if (!want_doubles) return;
    const Integer rp4 = S.rg(r.p4b), rp7 = S.rg(p7b), rh1 = S.rg(r.h1b), rh2 = S.rg(r.h2b);
    OperandView t, v;
    double sign;
    long long st[4];
    if (p7b < r.p4b) {  // block <p7 p4||h1 h2>; TCE_SORT_4(4,3,2,1), factor -1 (tce_hashnsort.F:129-135; left_4 :856-863)
      t.base = amp2(am[1], am[0], am[2], am[3], rp7, rp4, rh1, rh2, st);
      t.kstride = st[0]; t.stride[N_P4] = st[1]; t.stride[N_H1] = st[2]; t.stride[N_H2] = st[3];
      sign = -1.0;
    } else {            // block <p4 p7||h1 h2>; TCE_SORT_4(4,3,1,2), factor +1 (:137-144; left_4 :864-871)
      t.base = amp2(am[0], am[1], am[2], am[3], rp4, rp7, rh1, rh2, st);
      t.stride[N_P4] = st[0]; t.kstride = st[1]; t.stride[N_H1] = st[2]; t.stride[N_H2] = st[3];
      sign = 1.0;
    }
    // block <p5 p6||h3 p7> stored (p5,p6,h3,p7), p7 fastest == v2sub(p7,h3,p6,p5)  (:149-161); CR-CCSD(T): i1(p5 p6 h3 p7)
    // of cr_ccsd_t_N_2, same layout, key p7-noab-1 + nvab*(h3-1 + noab*(p6-noab-1 + nvab*(p5-noab-1))) (cr_ccsd_t_N.F:3753-3756)
    const Integer ckey = bm[3] - S.noab - 1 + S.nvab * (bm[2] - 1 + S.noab * (bm[1] - S.noab - 1 + S.nvab * (bm[0] - S.noab - 1)));
    if (cr == CR_EOM_RIGHT) {
      // creomsd_t_n2_mem_2 (t2 x i2_2, cre_t_K factor = -2 x the sd_t_d2cp_K sign), _4 (x2 x i2_4, factor = the sign), and
      // r0 * cr_ccsd_t_N_2
      v.kstride = 1; v.stride[N_H3] = rp7; v.stride[N_P6] = rp7 * S.rg(r.h3b);
      v.stride[N_P5] = rp7 * S.rg(r.h3b) * S.rg(r.p6b);
      OperandView tx = t;
      tx.base = c->d_x2 + (t.base - c->d_t2);
      if (c->eom_lr0) { v.base = crn2_block(c, ckey); push(t, v, sign * c->eom_r0, rp7, fire); }
      v.base = c->d_m2 + hash_lookup_or_die(c->m2_hash, ckey, "creom i2_2(pphp)"); push(t, v, -2.0 * sign, rp7, fire);
      v.base = c->d_m4 + hash_lookup_or_die(c->m4_hash, ckey, "creom i2_4(pphp)"); push(tx, v, sign, rp7, fire);
      return;
    }
    if (cr == CR_MOMENT)
      v.base = crn2_block(c, ckey);
    else
      v.base = v2_operand(c, bm[0], bm[1], bm[2], bm[3], "v2(pphp)");
    v.kstride = 1; v.stride[N_H3] = rp7; v.stride[N_P6] = rp7 * S.rg(r.h3b);
    v.stride[N_P5] = rp7 * S.rg(r.h3b) * S.rg(r.p6b);
    push(t, v, sign, rp7, fire);
  }
};

void tuple_ranges(const HostState& S, const Integer t[6], int R[6]) {
  R[POS_P4] = (int)S.rg(t[0]); R[POS_P5] = (int)S.rg(t[1]); R[POS_P6] = (int)S.rg(t[2]);
  R[POS_H1] = (int)S.rg(t[3]); R[POS_H2] = (int)S.rg(t[4]); R[POS_H3] = (int)S.rg(t[5]);
}

// [item_lo, item_hi): sub-range of the tuple's sub-tiles this launch evaluates (item_hi < 0: all)
void emit_tuple(nwc_triples_ctx* c, const Integer t[6], long long item_lo = 0, long long item_hi = -1) {
  const HostState& S = c->S;
  int R[6];
  tuple_ranges(S, t, R);
  c->eng->begin_tuple(R);
  NativeSink sink{c, *c->eng, S};
  walk_singles(S, t, sink);
  walk_doubles(S, t, sink);
  const double* eps[6] = {c->d_evl + S.offset[t[3] - 1], c->d_evl + S.offset[t[4] - 1], c->d_evl + S.offset[t[5] - 1],
                          c->d_evl + S.offset[t[0] - 1], c->d_evl + S.offset[t[1] - 1], c->d_evl + S.offset[t[2] - 1]};
  c->eng->end_tuple(eps, tuple_factor(S, t), item_lo, item_hi);
}

// Lambda-CCSD(T): a two-sided tuple.  Side 0 = Td = ccsd_t_doubles(T2,V2) (lambda_ccsd_t.F:109-111); side 1 = the
// left-hand doubles Yd = y2*f (left_2, doubles-bound outer products) - sum_h7 y2*v (left_3) - sum_p7 y2*v (left_4);
// singles tile = Ys = y1*v (left_1).  The kernel accumulates Td, parks it in a second canonical tile, accumulates Yd,
// and the energy pass forms sum f Td Yd/Delta and sum f Td (Ys+Yd)/Delta -- neither tile ever exists in HBM.
void emit_tuple_lambda(nwc_triples_ctx* c, const Integer t[6], long long item_lo = 0, long long item_hi = -1) {
  const HostState& S = c->S;
  int R[6];
  tuple_ranges(S, t, R);
  c->eng->begin_tuple(R);
  {   // right-hand doubles
    NativeSink rhs{c, *c->eng, S};
    rhs.want_singles = false;
    walk_doubles(S, t, rhs);
  }
  c->eng->set_two_sided();
  {   // left-hand contractions (side 1) and left-hand singles
    NativeSink lhs{c, *c->eng, S};
    lhs.lam = true;
    lhs.side = 1;
    walk_singles(S, t, lhs);
    walk_doubles(S, t, lhs);
  }
  // lambda_ccsd_t_left_2: i0(h4 h5 h6 p1 p2 p3) += P(9) y(h4 h5 p1 p2) f(h6 p3).  Written from the algebra: one particle
  // P_a and one hole H_b of the tuple go to f, the other two of each kind (ascending, hence canonical blocks) to y2;
  // P(h4 h5 / h6) = 1 - (h6<->h4) - (h6<->h5) gives the sign -1 exactly when the MIDDLE hole (particle) is the special
  // one (cf. the nine TCE_SORTACC_6 factors at lambda_ccsd_t_left.F:416-472).  Equal tiles need no special casing: the
  // nine terms are distinct index assignments of the same t3 element.
  const int ppos[3] = {POS_P4, POS_P5, POS_P6}, hpos[3] = {POS_H1, POS_H2, POS_H3};
  const Integer N = S.N();
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2, x = b == 0 ? 1 : 0, y = b == 2 ? 1 : 2;
      const Integer Pa = t[a], Pu = t[u], Pv = t[v], Hb = t[3 + b], Hx = t[3 + x], Hy = t[3 + y];
      if (S.sp(Hb) != S.sp(Pa) || (S.sy(Hb) ^ S.sy(Pa)) != 0) continue;                                     // :366-367
      if (S.sp(Hx) + S.sp(Hy) != S.sp(Pu) + S.sp(Pv) || (S.sy(Hx) ^ S.sy(Hy) ^ S.sy(Pu) ^ S.sy(Pv)) != 0) continue;
      const Integer four[4] = {Hx, Hy, Pu, Pv}, two[2] = {Hb, Pa};
      Integer m4[4], m2[2];
      restricted_map(S, 4, four, m4);                                                                       // :368
      restricted_map(S, 2, two, m2);                                                                        // :369
      const double* fblk = c->d_f1 + hash_lookup_or_die(c->f1_hash, m2[1] - 1 + N * (m2[0] - 1), "f1(hp)");  // :390-391
      const double* yblk = c->d_y2 + hash_lookup_or_die(c->y2_hash, y2_key(S, m4[0], m4[1], m4[2], m4[3]), "lambda2");
      int sa[6] = {0, 0, 0, 0, 0, 0}, sb[6] = {0, 0, 0, 0, 0, 0};
      sa[ppos[a]] = 1; sa[hpos[b]] = (int)S.rg(Pa);                       // f block (h6,p3), p3 fastest
      sb[ppos[v]] = 1; sb[ppos[u]] = (int)S.rg(Pv);                       // y2 block (h4,h5,p1,p2), p2 fastest
      sb[hpos[y]] = (int)(S.rg(Pu) * S.rg(Pv)); sb[hpos[x]] = (int)(S.rg(Hy) * S.rg(Pu) * S.rg(Pv));
      const bool neg = (a == 1) != (b == 1);
      c->eng->add_outer_product(fblk, sa, yblk, sb, neg, Engine::OP_SIDE1);
    }
  const double* eps[6] = {c->d_evl + S.offset[t[3] - 1], c->d_evl + S.offset[t[4] - 1], c->d_evl + S.offset[t[5] - 1],
                          c->d_evl + S.offset[t[0] - 1], c->d_evl + S.offset[t[1] - 1], c->d_evl + S.offset[t[2] - 1]};
  c->eng->end_tuple(eps, tuple_factor(S, t), item_lo, item_hi);
}

// CR-CCSD(T) (cr_ccsd_t.F:125-207): per tuple four t3-sized tiles -- S, D of (T), the moment M and the denominator
// tile E -- and four sums  num1 = <M,D>, num2 = <M,S+D>, den1 = <E,D>, den2 = <E,S+D>,  <A,B> = sum f A B / Delta.
// Two two-sided tuples through the LAMBDA instantiation of the fused kernel (energy = (<T0,T1>, <T0,T1+Ts>)):
//   pass 0 (numerators):   side 0 = M (contractions with the dressed intermediates), side 1 = D, singles tile = S
//   pass 1 (denominators): side 0 = E (18 outer products, no contraction),           side 1 = D, singles tile = S
// so D is contracted twice (1.5x the minimal FLOPs of the method).  pass 2 = both at once, the default: a DUAL tuple
// (engine.h set_dual) with side 0 = M, side 1 = D, singles = S and E as a fourth tile that the kernel forms in
// registers after M has been consumed (kernels.cu, "Dual tuples") -- M and D are contracted once each, the minimal FLOP
// count.  None of the four tiles ever exists in HBM.
void emit_tuple_cr(nwc_triples_ctx* c, const Integer t[6], int pass, long long item_lo = 0, long long item_hi = -1) {
  const HostState& S = c->S;
  int R[6];
  tuple_ranges(S, t, R);
  c->eng->begin_tuple(R);
  if (pass == 0 || pass == 2) {   // cr_ccsd_t_N toggle 2: _N_1 (Sum h11) and _N_2 (Sum p12)
    NativeSink m{c, *c->eng, S};
    m.cr = NativeSink::CR_MOMENT;
    m.want_singles = false;
    walk_doubles(S, t, m);
  }
  if (pass == 1 || pass == 2) {   // cr_ccsd_t_E toggle 2: _E_1, _E_2
    NativeSink d{c, *c->eng, S};
    d.cr = NativeSink::CR_DENOM;
    walk_cr_e1(S, t, d);
    walk_singles(S, t, d, S.irrep_t ^ S.irrep_t ^ S.irrep_t);
  }
  if (pass == 2) c->eng->set_dual();
  c->eng->set_two_sided();
  {   // ccsd_t_singles_l / ccsd_t_doubles_l (cr_ccsd_t.F:139-144)
    NativeSink rhs{c, *c->eng, S};
    rhs.side = 1;
    walk_singles(S, t, rhs);
    walk_doubles(S, t, rhs);
  }
  const double* eps[6] = {c->d_evl + S.offset[t[3] - 1], c->d_evl + S.offset[t[4] - 1], c->d_evl + S.offset[t[5] - 1],
                          c->d_evl + S.offset[t[0] - 1], c->d_evl + S.offset[t[1] - 1], c->d_evl + S.offset[t[2] - 1]};
  c->eng->end_tuple(eps, tuple_factor(S, t), item_lo, item_hi);   // cr_ccsd_t.F:153-167 == ccsd_t_dot.F:52-66
}

// CR-EOMCCSD(T) (cr_eomccsd_t.F:325-493): per tuple a right tile R (contractions only) and a left tile L (outer products
// only) and four sums  A = sum f R R/denex,  B = sum f L R,  C = sum f L R/denex,  D = sum f L L  (denex = Delta + omega);
// the file adds A + B into num1 and C + D into den1 (:455-464).  Composed from tuple forms the kernels already have:
//   which 0 "X":  PLAIN (T)-type tuple, doubles = R, singles = L, orbital energies of the h1 slot shifted by omega
//                 -> (A, A + C): the (T) formulas E[T] = <D,D>, E(T) = <D,D+S> with D = R, S = L
//   which 1 "Y1": two-sided tuple with UNIT denominators (eps = 1 for the h1 slot, 0 elsewhere): side 0 = L (outer
//                 products bound to it), side 1 = R, singles = the t2 x (r0 t1 + x1) half of L -> (B, B + <L,La>)
//   which 2 "Y2": two-sided, unit denominators, no contraction: side 0 = L, singles = the other half of L -> (0, <L,Lb>)
// so D = <L,La> + <L,Lb>.  R is contracted twice (once at the plain kernel's 3 CTAs/SM): NWC_CREOM_COMPOSED=1.
//   which 3, the default: ONE dual-energy tuple in its CR-EOMCCSD(T) form (engine.h set_dual_eom): side 1 = R, singles = L,
//                 shifted orbital energies; the kernel returns (A, A + C) and, from an undenominated second energy pass
//                 over the same two tiles, (B, B + D).  R is contracted once.
void emit_tuple_creom(nwc_triples_ctx* c, const Integer t[6], int which, long long item_lo = 0, long long item_hi = -1) {
  const HostState& S = c->S;
  int R[6];
  tuple_ranges(S, t, R);
  c->eng->begin_tuple(R);
  if (which == 0 || which == 1 || which == 3) {   // R: r0 * cr_ccsd_t_N + creomsd_t_n2_mem_1..4 (cr_eomccsd_t.F:377-395)
    NativeSink r{c, *c->eng, S};
    r.cr = NativeSink::CR_EOM_RIGHT;
    r.want_singles = false;
    r.side = which == 0 ? 0 : 1;
    walk_doubles(S, t, r);
  }
  {   // L: r0 * cr_ccsd_t_E + q3rexpt2_1, _2 (:400-419)
    NativeSink l{c, *c->eng, S};
    l.cr = NativeSink::CR_EOM_LEFT;
    l.op_target = (which == 0 || which == 3) ? Engine::OP_SINGLES : Engine::OP_SIDE0;
    walk_cr_e1(S, t, l);
    walk_singles(S, t, l, S.irrep_t ^ S.irrep_t ^ S.irrep_t);
    if (which == 3) c->eng->set_dual_eom();
    if (which == 1 || which == 2) {
      c->eng->set_two_sided();
      NativeSink h{c, *c->eng, S};   // the half of L that plays the singles tile
      h.cr = NativeSink::CR_EOM_LEFT;
      h.op_target = Engine::OP_SINGLES;
      h.op_mask = which == 1 ? 1 : 2;
      walk_cr_e1(S, t, h);
      walk_singles(S, t, h, S.irrep_t ^ S.irrep_t ^ S.irrep_t);
    }
  }
  const double* eps[6];
  if (which == 0 || which == 3) {
    const double* e[6] = {c->d_evl_shift + S.offset[t[3] - 1], c->d_evl + S.offset[t[4] - 1], c->d_evl + S.offset[t[5] - 1],
                          c->d_evl + S.offset[t[0] - 1], c->d_evl + S.offset[t[1] - 1], c->d_evl + S.offset[t[2] - 1]};
    for (int q = 0; q < 6; q++) eps[q] = e[q];
  } else {
    eps[0] = c->d_unit;
    for (int q = 1; q < 6; q++) eps[q] = c->d_zero;
  }
  c->eng->end_tuple(eps, tuple_factor(S, t), item_lo, item_hi);   // cr_eomccsd_t.F:421-435 == ccsd_t_dot.F:52-66
}

// Double-buffered batch loop: while the GPU runs batch k the host walks the driver logic of batch k+1 into the other
// slot.  Energies are accumulated in task order (deterministic).  slot_of[i] (optional) = where result i goes in per_task.
struct Pipeline {
  nwc_triples_ctx* c;
  Engine& e;
  double* energy;
  double* per_task;
  std::vector<Integer> cur_pos, prev_pos;   // per_task row of each tuple of the batch being built / in flight
  int prev = -1;
  bool dual = false;   // dual-energy batches (engine.h set_dual): per_task rows hold four doubles, pair 0 then pair 1
  std::vector<double> eb;
  Pipeline(nwc_triples_ctx* c_, double* en, double* pt) : c(c_), e(*c_->eng), energy(en), per_task(pt) {}
  void emitted(Integer row) {
    cur_pos.push_back(row);
    if (e.arena().used() >= c->batch_bytes || e.pending_items() > (size_t)32000000 || e.pending_tuples() >= 4096) flush();
  }
  void wait_prev() {
    if (prev < 0) return;
    const size_t n = prev_pos.size();
    eb.assign((dual ? 4 : 2) * n + 2, 0.0);
    e.collect(prev, eb.data(), /*compact=*/true);
    for (size_t i = 0; i < n; i++) {
      energy[0] += eb[2 * i];
      energy[1] += eb[2 * i + 1];
      if (!per_task || prev_pos[i] < 0) continue;
      if (!dual) { per_task[2 * prev_pos[i]] += eb[2 * i]; per_task[2 * prev_pos[i] + 1] += eb[2 * i + 1]; }
      else
        for (int q = 0; q < 2; q++) {
          per_task[4 * prev_pos[i] + q] += eb[2 * i + q];             // pair 0 of tuple i
          per_task[4 * prev_pos[i] + 2 + q] += eb[2 * (n + i) + q];   // pair 1: the shadow tuples follow the real ones
        }
    }
    slot_done(c, prev);
    prev = -1;
  }
  void flush() {
    const int s = e.submit();
    if (s < 0) return;
    wait_prev();          // frees the slot the engine has just switched to; the GPU already has batch s queued
    prev = s;
    prev_pos.swap(cur_pos);
    cur_pos.clear();
  }
  void finish() { flush(); wait_prev(); }
};

int upload(double** dst, size_t* n_out, const double* src, size_t n, Engine* e) {
  if (e->trace_only()) {   // nothing is copied anywhere: the records will point into the caller's arrays
    *dst = const_cast<double*>(src);
    *n_out = n;
    return 0;
  }
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  NWC_TRY(cudaMalloc((void**)dst, (n ? n : 1) * sizeof(double)));
  if (n && src) {
    NWC_TRY(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    e->stats.h2d_bytes += n * sizeof(double);
  } else if (n) {   // no host data: the store is allocated only (filled on the device, nwc_triples_synth_fill)
    NWC_TRY(cudaMemset(*dst, 0, n * sizeof(double)));
  }
  *n_out = n;
  return 0;
}

size_t store_size(const Integer* hash, const HostState& S, int which) {
  // size = offset of last block + its size; recompute from keys
  const Integer n = hash[0];
  if (n == 0) return 0;
  Integer key = hash[n], off = hash[2 * n];
  Integer sz = 0;
  const Integer top = which == 1 ? S.noab * S.nvab : which == 2 ? S.noab * S.noab * S.nvab * S.nvab : S.N() * S.N() * S.N() * S.N();
  if (key < 0 || key >= top || off < 0) throw Error("nwc_triples: offset table does not belong to this tiling (last key out of range)");
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
  return guarded(nullptr, [&]() {
    nwc_triples_ctx* c = new nwc_triples_ctx();
    c->eng = new Engine(device);
    *out = c;
    return 0;
  });
}

int nwc_triples_create_trace(nwc_triples_ctx** out) {
  return guarded(nullptr, [&]() {
    nwc_triples_ctx* c = new nwc_triples_ctx();
    c->eng = new Engine(-1);
    *out = c;
    return 0;
  });
}

int nwc_triples_destroy(nwc_triples_ctx* c) {
  if (!c) return 0;
  if (c->eng->trace_only()) {
    delete c->eng;
    delete c;
    return 0;
  }
  cudaSetDevice(c->eng->device());
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  free_stores(c);
  cudaFree(c->d_red);
  delete c->eng;
  delete c;
  return 0;
}

// Panel index order for this tiling (engine.h set_order): the padding model evaluated on a tuple made of the average
// hole tile and the average particle tile decides whether holes or particles go first inside the 64-row blocks.
static int choose_order(const HostState& S) {
  const char* e = getenv("NWC_ORDER");
  if (e && (*e == '0' || *e == '1')) return *e - '0';
  double cost[2] = {0, 0};
  long long n = 0;
  for (Integer h = 1; h <= S.noab; h++)
    for (Integer p = S.noab + 1; p <= S.noab + S.nvab; p++) {
      const int R[6] = {(int)S.rg(h), (int)S.rg(h), (int)S.rg(h), (int)S.rg(p), (int)S.rg(p), (int)S.rg(p)};
      const double w = (double)S.rg(h) * S.rg(p);
      cost[0] += w * Engine::padding_cost(R, 0);
      cost[1] += w * Engine::padding_cost(R, 1);
      n++;
    }
  return cost[1] < cost[0] * 0.995 ? 1 : 0;
}

static int finish_state(nwc_triples_ctx* c) {
  size_t ne;
  if (upload(&c->d_evl, &ne, c->S.evl.data(), c->S.evl.size(), c->eng)) return 1;
  build_task_list(c->S, c->klist);
  c->eng->set_order(choose_order(c->S));
  return 0;
}

}  // extern "C"
// shard layout of `n` blocks of sizes size(i): block i -> rank i % nranks, compacted in index order
template <class SizeOf>
static void build_shards(nwc_triples_ctx* c, Integer n, int rank, int nranks, SizeOf size_of, Integer* my_doubles) {
  c->v2_shard_off.assign((size_t)n, 0);
  c->v2_block_n.assign((size_t)n, 0);
  std::vector<Integer> fill((size_t)nranks, 0);
  for (Integer i = 0; i < n; i++) {
    const int owner = (int)(i % nranks);
    const Integer sz = size_of(i);
    c->v2_shard_off[(size_t)i] = fill[owner];
    c->v2_block_n[(size_t)i] = sz;
    fill[owner] += sz;
  }
  *my_doubles = fill[rank];
  c->v2_nshards = nranks; c->v2_rank = rank;
  c->v2_peer.assign((size_t)nranks, nullptr);
  c->v2_peer_opened.assign((size_t)nranks, 0);
}

extern "C" {
int nwc_triples_set_state(nwc_triples_ctx* c, const nwc_tce_state* st) {
  return guarded(c, [&]() {
    if (!c->eng->trace_only()) NWC_TRY(cudaSetDevice(c->eng->device()));
    free_stores(c);
    c->S.load_tables(st);
    const HostState& S = c->S;
    if (upload(&c->d_t1, &c->n_t1, st->t1, store_size(st->t1_hash, S, 1), c->eng)) return 1;
    if (upload(&c->d_t2, &c->n_t2, st->t2, store_size(st->t2_hash, S, 2), c->eng)) return 1;
    if (upload(&c->d_v2, &c->n_v2, st->v2, store_size(st->v2_hash, S, 3), c->eng)) return 1;
    return finish_state(c);
  });
}

static int set_state_orbital(nwc_triples_ctx* c, const nwc_tce_state* st, const nwc_tce_orb_state* orb, int rank, int nranks) {
  if (c->eng->trace_only()) { g_err = "a trace context takes replicated spin-orbital stores only (nwc_triples_set_state)"; return 1; }
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
  NWC_TRY(cudaSetDevice(c->eng->device()));
  free_stores(c);
  nwc_tce_state s2 = *st;
  s2.v2_hash = nullptr;   // not read in this mode
  c->S.load_tables(&s2);
  HostState& S = c->S;
  const std::string err = S.load_orbital(orb->noa, orb->nva, orb->b2am, orb->spin_alpha, orb->sym_alpha,
                                         orb->range_alpha, orb->v2orb_hash);
  if (!err.empty()) { g_err = err; return 1; }
  if (upload(&c->d_t1, &c->n_t1, st->t1, store_size(st->t1_hash, S, 1), c->eng)) return 1;
  if (upload(&c->d_t2, &c->n_t2, st->t2, store_size(st->t2_hash, S, 2), c->eng)) return 1;
  // only the blocks (T) can touch become resident (the rest of d_v2orb stays on the host)
  if (nranks == 1) {
    NWC_TRY(cudaMalloc((void**)&c->d_v2orb, (size_t)(S.orb_size ? S.orb_size : 1) * sizeof(double)));
    if (orb->v2orb) {
      for (const HostState::OrbRun& r : S.orb_runs)   // compacted run by run
        NWC_TRY(cudaMemcpy(c->d_v2orb + r.dst, orb->v2orb + r.src, (size_t)r.n * sizeof(double), cudaMemcpyHostToDevice));
      c->eng->stats.h2d_bytes += (size_t)S.orb_size * sizeof(double);
    }
    c->n_v2orb = (size_t)S.orb_size;
  } else {
    // sharded: needed block i -> rank i % nranks; orb->v2orb (if given) is the caller's FULL d_v2orb file, of which
    // only this rank's blocks are read
    Integer mine = 0;
    build_shards(c, (Integer)S.orb_blocks.size(), rank, nranks, [&](Integer i) { return S.orb_blocks[(size_t)i].size; }, &mine);
    NWC_TRY(cudaMalloc((void**)&c->d_v2orb, (size_t)(mine ? mine : 1) * sizeof(double)));
    if (orb->v2orb) {
      for (size_t i = (size_t)rank; i < S.orb_blocks.size(); i += (size_t)nranks)
        NWC_TRY(cudaMemcpy(c->d_v2orb + c->v2_shard_off[i], orb->v2orb + S.orb_blocks[i].host_off,
                           (size_t)S.orb_blocks[i].size * sizeof(double), cudaMemcpyHostToDevice));
      c->eng->stats.h2d_bytes += (size_t)mine * sizeof(double);
    }
    c->n_v2orb = (size_t)mine;
    c->v2_peer[(size_t)rank] = c->d_v2orb;
  }
  return finish_state(c);
}

int nwc_triples_set_state_2eorb(nwc_triples_ctx* c, const nwc_tce_state* st, const nwc_tce_orb_state* orb) {
  return guarded(c, [&]() { return set_state_orbital(c, st, orb, 0, 1); });
}

int nwc_triples_set_state_2eorb_sharded(nwc_triples_ctx* c, const nwc_tce_state* st, const nwc_tce_orb_state* orb, int rank,
                                        int nranks) {
  return guarded(c, [&]() { return set_state_orbital(c, st, orb, rank, nranks); });
}

// Sharded variant: st->v2 points at THIS rank's shard only (its blocks, table order, compacted), or is NULL.
int nwc_triples_set_state_sharded(nwc_triples_ctx* c, const nwc_tce_state* st, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    free_stores(c);
    c->S.load_tables(st);
    const HostState& S = c->S;
    if (upload(&c->d_t1, &c->n_t1, st->t1, store_size(st->t1_hash, S, 1), c->eng)) return 1;
    if (upload(&c->d_t2, &c->n_t2, st->t2, store_size(st->t2_hash, S, 2), c->eng)) return 1;
    // shard offsets of every block (all ranks compute the same table)
    const Integer n = S.v2_hash[0];
    const Integer total = (Integer)store_size(st->v2_hash, S, 3);
    Integer mine = 0;
    build_shards(c, n, rank, nranks, [&](Integer i) {
      const Integer off = S.v2_hash[(size_t)(n + 1 + i)], next = (i + 1 < n) ? S.v2_hash[(size_t)(n + 2 + i)] : total;
      return next - off;
    }, &mine);
    if (upload(&c->d_v2, &c->n_v2, st->v2, (size_t)mine, c->eng)) return 1;
    c->v2_peer[(size_t)rank] = c->d_v2;
    return finish_state(c);
  });
}

static double* shard_base(nwc_triples_ctx* c) { return c->S.intorb ? c->d_v2orb : c->d_v2; }

// CUDA IPC handle (64 bytes) of this rank's V2 shard; the host all-gathers them (MPI/GA in NWChem)
int nwc_triples_v2_ipc_handle(nwc_triples_ctx* c, char handle64[64]) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  cudaIpcMemHandle_t h;
  NWC_TRY(cudaIpcGetMemHandle(&h, shard_base(c)));
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
void* nwc_triples_v2_shard_ptr(nwc_triples_ctx* c) { return shard_base(c); }
int nwc_triples_v2_set_peer_ptr(nwc_triples_ctx* c, int rank, void* dev_ptr) {
  if (rank < 0 || rank >= c->v2_nshards) { g_err = "bad peer rank"; return 1; }
  c->v2_peer[(size_t)rank] = (double*)dev_ptr;
  return 0;
}

// ---- synthetic stores generated on the device (bench / tests) ----
// Fills T1, T2 and the V2 store this context holds (whole, or this rank's shard) with scale * U(-1,1) values that are a
// pure function of (seed, store, block key, element): no rank ever needs the store on the host, and every rank count sees
// the same tensors.  store ids: 1 = T1, 2 = T2, 3 = spin-orbital V2, 4 = orbital-form V2.
int nwc_triples_synth_fill(nwc_triples_ctx* c, unsigned long long seed, double scale_t1, double scale_t2, double scale_v2) {
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    const HostState& S = c->S;
    cudaStream_t st = c->eng->stream();
    auto run = [&](std::vector<FillJob>& jobs, unsigned long long store, double scale) -> int {
      if (jobs.empty()) return 0;
      long long mx = 0;
      for (const FillJob& j : jobs) mx = std::max(mx, j.n);
      FillJob* d = nullptr;
      NWC_TRY(cudaMalloc((void**)&d, jobs.size() * sizeof(FillJob)));
      NWC_TRY(cudaMemcpyAsync(d, jobs.data(), jobs.size() * sizeof(FillJob), cudaMemcpyHostToDevice, st));
      launch_synth_fill(d, (int)jobs.size(), mx, seed, store, scale, st);
      NWC_TRY(cudaGetLastError());
      NWC_TRY(cudaStreamSynchronize(st));
      NWC_TRY(cudaFree(d));
      return 0;
    };
    auto table_jobs = [&](const std::vector<Integer>& hash, double* base, size_t total, std::vector<FillJob>& jobs) {
      const Integer n = hash[0];
      for (Integer i = 0; i < n; i++) {
        const Integer off = hash[(size_t)(n + 1 + i)], next = (i + 1 < n) ? hash[(size_t)(n + 2 + i)] : (Integer)total;
        jobs.push_back(FillJob{base + off, (long long)hash[(size_t)(1 + i)], (long long)(next - off)});
      }
    };
    std::vector<FillJob> jobs;
    table_jobs(S.t1_hash, c->d_t1, c->n_t1, jobs);
    if (run(jobs, 1, scale_t1)) return 1;
    jobs.clear();
    table_jobs(S.t2_hash, c->d_t2, c->n_t2, jobs);
    if (run(jobs, 2, scale_t2)) return 1;
    jobs.clear();
    if (S.intorb) {
      for (size_t i = 0; i < S.orb_blocks.size(); i++) {
        const HostState::OrbBlock& b = S.orb_blocks[i];
        if (c->v2_nshards <= 1) jobs.push_back(FillJob{c->d_v2orb + S.orb_off.at(b.key), (long long)b.key, (long long)b.size});
        else if ((int)(i % (size_t)c->v2_nshards) == c->v2_rank)
          jobs.push_back(FillJob{c->d_v2orb + c->v2_shard_off[i], (long long)b.key, (long long)b.size});
      }
      if (run(jobs, 4, scale_v2)) return 1;
    } else if (c->v2_nshards <= 1) {
      table_jobs(S.v2_hash, c->d_v2, c->n_v2, jobs);
      if (run(jobs, 3, scale_v2)) return 1;
    } else {
      const Integer n = S.v2_hash[0];
      for (Integer i = c->v2_rank; i < n; i += c->v2_nshards)
        jobs.push_back(FillJob{c->d_v2 + c->v2_shard_off[(size_t)i], (long long)S.v2_hash[(size_t)(1 + i)], (long long)c->v2_block_n[(size_t)i]});
      if (run(jobs, 3, scale_v2)) return 1;
    }
    return 0;
  });
}

// validation aid: read back part of a resident store.  which: 1 = T1, 2 = T2, 3 = V2 (this rank's shard), 4 = orbital V2
int nwc_triples_debug_read(nwc_triples_ctx* c, int which, size_t offset, size_t n, double* host_out) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  const double* base = which == 1 ? c->d_t1 : which == 2 ? c->d_t2 : which == 3 ? c->d_v2 : c->d_v2orb;
  const size_t cap = which == 1 ? c->n_t1 : which == 2 ? c->n_t2 : which == 3 ? c->n_v2 : c->n_v2orb;
  if (!base || offset + n > cap) { g_err = "nwc_triples_debug_read: out of range"; return 1; }
  NWC_TRY(cudaMemcpy(host_out, base + offset, n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// The spin-orbital block <g3 g4||g1 g2> (tile ids as stored, i.e. after tce_restricted_4) as the (T) path sees it,
// copied to the host: the device form of get_hash_block / get_hash_block_i for one block.
int nwc_triples_export_v2_block(nwc_triples_ctx* c, const Integer g3g4g1g2[4], double* host_out) {
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    const HostState& S = c->S;
    const Integer g3 = g3g4g1g2[0], g4 = g3g4g1g2[1], g1 = g3g4g1g2[2], g2 = g3g4g1g2[3];
    const double* p = v2_operand(c, g3, g4, g1, g2, "v2(export)");
    c->eng->flush_prep();
    const size_t n = (size_t)(S.rg(g3) * S.rg(g4) * S.rg(g1) * S.rg(g2));
    NWC_TRY(cudaMemcpyAsync(host_out, p, n * sizeof(double), cudaMemcpyDeviceToHost, c->eng->stream()));
    NWC_TRY(cudaStreamSynchronize(c->eng->stream()));
    c->eng->arena().reset();
    slot_done(c, c->eng->current_slot());
    return 0;
  });
}

Integer nwc_triples_num_tasks(nwc_triples_ctx* c) { return (Integer)(c->klist.size() / 7); }

int nwc_triples_task_list(nwc_triples_ctx* c, Integer* klist7) {
  memcpy(klist7, c->klist.data(), c->klist.size() * sizeof(Integer));
  return 0;
}

int nwc_triples_run(nwc_triples_ctx* c, Integer first, Integer stride, Integer max_tasks, double energy[2],
                    double* per_task) {
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    if (stride <= 0) stride = 1;
    const Integer nt = (Integer)(c->klist.size() / 7);
    energy[0] = energy[1] = 0.0;
    Integer done = 0;
    for (Integer k = first; per_task && k < nt && (max_tasks <= 0 || done < max_tasks); k += stride, done++)
      per_task[2 * done] = per_task[2 * done + 1] = 0.0;
    Pipeline pipe(c, energy, per_task);
    done = 0;
    for (Integer k = first; k < nt && (max_tasks <= 0 || done < max_tasks); k += stride, done++) {
      emit_tuple(c, &c->klist[7 * k]);
      pipe.emitted(done);
    }
    pipe.finish();
    return 0;
  });
}

// Static block partition of the task space (north star: "the tile-tuple task space is block-partitioned across the
// GPUs"), the stand-in for the nxtask counter (ccsd_t.F:174-255).  Tasks [first_task, first_task + ntasks) of the
// heaviest-first list are laid end to end, every 4^6 sub-tile weighted by the k4 planes its tuple contracts (+ a constant
// for the per-sub-tile epilogue); rank r takes the r-th of nranks equal-cost contiguous pieces.  A tuple that straddles a
// boundary is shared between two ranks at sub-tile granularity (energies are additive over sub-tiles), so the balance
// does not depend on how many tuples there are.  per_task (optional, 2*ntasks, indexed by task - first_task) receives
// this rank's (partial) energies; summed over ranks it holds the per-task energies.
static int run_partition_ids(nwc_triples_ctx* c, Integer rank, Integer nranks, const std::vector<Integer>& ids,
                             double energy[2], double* per_task) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  const HostState& S = c->S;
  const Integer nt = (Integer)(c->klist.size() / 7);
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
  for (Integer id : ids)
    if (id < 0 || id >= nt) { g_err = "task index out of range"; return 1; }
  energy[0] = energy[1] = 0.0;
  if (per_task) for (size_t i = 0; i < 2 * ids.size(); i++) per_task[i] = 0.0;
  if (ids.empty()) return 0;
  std::vector<long long> ranges;
  block_partition(S, c->klist, rank, nranks, ids, ranges);
  Pipeline pipe(c, energy, per_task);
  for (size_t i = 0; i < ids.size(); i++) {
    const long long a = ranges[2 * i], b = ranges[2 * i + 1];
    if (b <= a) continue;
    emit_tuple(c, &c->klist[7 * (size_t)ids[i]], a, b);
    pipe.emitted((Integer)i);
  }
  pipe.finish();
  return 0;
}

int nwc_triples_run_partition(nwc_triples_ctx* c, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                              double energy[2], double* per_task) {
  return guarded(c, [&]() {
    const Integer nt = (Integer)(c->klist.size() / 7);
    if (first_task < 0) first_task = 0;
    if (ntasks <= 0 || first_task + ntasks > nt) ntasks = nt - first_task;
    std::vector<Integer> ids;
    for (Integer i = 0; i < ntasks; i++) ids.push_back(first_task + i);
    return run_partition_ids(c, rank, nranks, ids, energy, per_task);
  });
}

// the same for an explicit list of task indices (e.g. a strided sample of the list), partitioned in the order given
int nwc_triples_run_partition_list(nwc_triples_ctx* c, Integer rank, Integer nranks, const Integer* task_ids, Integer n,
                                   double energy[2], double* per_task) {
  return guarded(c, [&]() {
    std::vector<Integer> ids(task_ids, task_ids + (n > 0 ? n : 0));
    return run_partition_ids(c, rank, nranks, ids, energy, per_task);
  });
}

// one tuple restricted to the sub-tiles [item_lo, item_hi) of its linear sub-tile order (h3 block fastest, p4 block
// slowest): e.g. a p4 slab [4a, 4b) of the t3 tile is the range [a*m, b*m), m = sub-tiles per p4 block
int nwc_triples_run_items(nwc_triples_ctx* c, const Integer t[6], long long item_lo, long long item_hi, double energy[2]) {
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    emit_tuple(c, t, item_lo, item_hi);
    double out[2] = {0, 0};
    c->eng->run(out);
    slot_done(c, c->eng->current_slot());
    energy[0] = out[0];
    energy[1] = out[1];
    return 0;
  });
}

long long nwc_triples_tuple_items(nwc_triples_ctx* c, const Integer t[6]) {
  int R[6];
  tuple_ranges(c->S, t, R);
  return Engine::tuple_items(R);
}

int nwc_triples_run_restart(nwc_triples_ctx* c, Integer first, Integer stride, Integer* restart_begin, double* table,
                            double* table_bracket, Integer max_outer, double* t_energy) {
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    if (stride <= 0) stride = 1;
    if (*restart_begin < 1) *restart_begin = 1;
    const HostState& S = c->S;
    const Integer n0 = S.noab, n1 = S.noab + S.nvab;
    Integer done = 0;
    for (Integer p4 = n0 + *restart_begin; p4 <= n1; p4++) {
      if (max_outer > 0 && done >= max_outer) break;
      double en[2] = {0.0, 0.0};
      Pipeline pipe(c, en, nullptr);
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
                pipe.emitted(-1);
              }
      pipe.finish();
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
  });
}

// ---- Lambda-CCSD(T) (src/tce/ccsd_t/lambda_ccsd_t.F) ----
// lambda_1 / lambda_2 / Fock(h,p) block stores with their offset tables ([n, keys.., offsets..]; keys as in
// lambda_ccsd_t_left.F:154-155, :378-380, :390-391), replicated in HBM.  Call after a set_state* variant.
int nwc_triples_set_lambda(nwc_triples_ctx* c, const Integer* y1_hash, const double* y1, const Integer* y2_hash,
                           const double* y2, const Integer* f1_hash, const double* f1) {
  return guarded(c, [&]() {
    if (!c->eng->trace_only()) NWC_TRY(cudaSetDevice(c->eng->device()));
    const HostState& S = c->S;
    auto load = [&](const Integer* h, const double* data, std::vector<Integer>& tab, double** d, size_t* n, int kind) -> int {
      const Integer nb = h[0];
      tab.assign(h, h + 2 * nb + 1);
      size_t total = 0;
      if (nb > 0) {   // size of the last block from its key
        Integer key = h[nb], sz = 0;
        if (kind == 2) {
          const Integer p2 = key % S.nvab + S.noab + 1; key /= S.nvab;
          const Integer p1 = key % S.nvab + S.noab + 1; key /= S.nvab;
          const Integer h5 = key % S.noab + 1; key /= S.noab;
          if (key < 0 || key >= S.noab) throw Error("nwc_triples: lambda_2 offset table does not belong to this tiling");
          sz = S.rg(key + 1) * S.rg(h5) * S.rg(p1) * S.rg(p2);
        } else if (kind == 1) {
          const Integer p1 = key % S.nvab + S.noab + 1, h4 = key / S.nvab + 1;
          if (h4 < 1 || h4 > S.noab) throw Error("nwc_triples: lambda_1 offset table does not belong to this tiling");
          sz = S.rg(h4) * S.rg(p1);
        } else {
          const Integer g2 = key % S.N() + 1, g1 = key / S.N() + 1;
          if (g1 < 1 || g1 > S.N()) throw Error("nwc_triples: f1 offset table does not belong to this tiling");
          sz = S.rg(g1) * S.rg(g2);
        }
        total = (size_t)(h[2 * nb] + sz);
      }
      return upload(d, n, data, total, c->eng);
    };
    if (load(y1_hash, y1, c->y1_hash, &c->d_y1, &c->n_y1, 1)) return 1;
    if (load(y2_hash, y2, c->y2_hash, &c->d_y2, &c->n_y2, 2)) return 1;
    if (load(f1_hash, f1, c->f1_hash, &c->d_f1, &c->n_f1, 3)) return 1;
    return 0;
  });
}

// Lambda-CCSD[T] / Lambda-CCSD(T) correction energies of tasks first, first+stride, ... (lambda_ccsd_t.F:59-190):
//   energy[0] = sum f Td Yd / Delta ,  energy[1] = sum f Td (Ys + Yd) / Delta ,
// the left-hand tiles taken in T3 order (the sort lambda_ccsd_t.F:35-36 announces; see oracle/triples_oracle.c for the
// literal reading of the file).  The t3-sized tiles never exist in HBM here either: every tuple is a two-sided tuple
// (emit_tuple_lambda) of the LAMBDA instantiation of the fused kernel, at the minimal FLOP count of the method.
static int run_lambda_ids(nwc_triples_ctx* c, const std::vector<Integer>& ids, const std::vector<long long>* ranges,
                          double energy[2], double* per_task) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  if (!c->d_y2 || !c->d_y1 || !c->d_f1) { g_err = "nwc_triples_run_lambda: call nwc_triples_set_lambda first"; return 1; }
  energy[0] = energy[1] = 0.0;
  if (per_task) for (size_t i = 0; i < 2 * ids.size(); i++) per_task[i] = 0.0;
  Pipeline pipe(c, energy, per_task);
  for (size_t i = 0; i < ids.size(); i++) {
    const long long a = ranges ? (*ranges)[2 * i] : 0, b = ranges ? (*ranges)[2 * i + 1] : -1;
    if (ranges && b <= a) continue;
    emit_tuple_lambda(c, &c->klist[7 * (size_t)ids[i]], a, b);
    pipe.emitted((Integer)i);
  }
  pipe.finish();
  return 0;
}

int nwc_triples_run_lambda(nwc_triples_ctx* c, Integer first, Integer stride, Integer max_tasks, double energy[2],
                           double* per_task) {
  return guarded(c, [&]() {
    if (stride <= 0) stride = 1;
    const Integer nt = (Integer)(c->klist.size() / 7);
    std::vector<Integer> ids;
    for (Integer k = first; k < nt && (max_tasks <= 0 || (Integer)ids.size() < max_tasks); k += stride) ids.push_back(k);
    return run_lambda_ids(c, ids, nullptr, energy, per_task);
  });
}

// Lambda-CCSD(T) over the static block partition of nwc_triples_run_partition (the sums are additive over sub-tiles,
// so the rank sums add up exactly as for (T)).
int nwc_triples_run_lambda_partition(nwc_triples_ctx* c, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                                     double energy[2], double* per_task) {
  return guarded(c, [&]() {
    const Integer nt = (Integer)(c->klist.size() / 7);
    if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
    if (first_task < 0) first_task = 0;
    if (ntasks <= 0 || first_task + ntasks > nt) ntasks = nt - first_task;
    std::vector<Integer> ids;
    for (Integer i = 0; i < ntasks; i++) ids.push_back(first_task + i);
    std::vector<long long> ranges;
    if (!ids.empty()) block_partition(c->S, c->klist, rank, nranks, ids, ranges);
    return run_lambda_ids(c, ids, &ranges, energy, per_task);
  });
}

// ---- CR-CCSD(T) (src/tce/ccsd_t/cr_ccsd_t.F) ----
// The three intermediates the reference's tuple loop reads, with their offset tables ([n, keys.., offsets..]):
//   n1 = d_i1_1 of cr_ccsd_t_N: i1(h11 p4 h1 h2), blocks (p4b,h11b,h1b<=h2b), OFFSET_cr_ccsd_t_N_1_1 (cr_ccsd_t_N.F:773)
//   n2 = d_i1_2 of cr_ccsd_t_N: i1(p4 p5 h1 p12), blocks (p4b<=p5b,h1b,p12b), OFFSET_cr_ccsd_t_N_2_1 (:4011)
//   e2 = d_i1_2 of cr_ccsd_t_E: i1(p4 p5 h1 h2)_tt, the T2 block structure,   OFFSET_cr_ccsd_t_E_2_1 (cr_ccsd_t_E.F:907)
// i.e. what cr_ccsd_t_N(...,1) / cr_ccsd_t_E(...,1) leave in GA, or the files gr1_1 / gr1_2 / ei1_2 of read_in3
// (cr_ccsd_t_N.F:98-104).  Replicated in HBM.  Call after a set_state* variant.
static int set_cr_impl(nwc_triples_ctx* c, const Integer* n1_hash, const double* n1, const Integer* n2_hash, const double* n2,
                       const Integer* e2_hash, const double* e2, int rank, int nranks);
int nwc_triples_set_cr(nwc_triples_ctx* c, const Integer* n1_hash, const double* n1, const Integer* n2_hash, const double* n2,
                       const Integer* e2_hash, const double* e2) {
  return set_cr_impl(c, n1_hash, n1, n2_hash, n2, e2_hash, e2, 0, 1);
}
// The pphp intermediate is as large as V2's <pp||hp> class (2.5*o*v^3 doubles: 527 GB for (H2O)10), so like V2 it can be
// dealt over the GPUs of the node: block i of its offset table lives on rank i % nranks and `n2` holds THIS rank's blocks
// only (table order, compacted); the small hphh and pphh intermediates stay replicated.  Afterwards the ranks exchange
// nwc_triples_cr_ipc_handle / nwc_triples_cr_open_peers (or, inside one process, cr_shard_ptr / cr_set_peer_ptr) exactly
// as for a sharded V2; remote blocks are pulled over NVLink per batch.
int nwc_triples_set_cr_sharded(nwc_triples_ctx* c, const Integer* n1_hash, const double* n1, const Integer* n2_hash,
                               const double* n2_shard, const Integer* e2_hash, const double* e2, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
  if (c->eng->trace_only() && nranks > 1) { g_err = "a trace context takes replicated stores only"; return 1; }
  return set_cr_impl(c, n1_hash, n1, n2_hash, n2_shard, e2_hash, e2, rank, nranks);
}
int nwc_triples_cr_ipc_handle(nwc_triples_ctx* c, char handle64[64]) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  if (!c->d_crn2) { g_err = "nwc_triples_cr_ipc_handle: call nwc_triples_set_cr_sharded first"; return 1; }
  cudaIpcMemHandle_t h;
  NWC_TRY(cudaIpcGetMemHandle(&h, c->d_crn2));
  memcpy(handle64, &h, 64);
  return 0;
}
int nwc_triples_cr_open_peers(nwc_triples_ctx* c, const char* handles) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  PeerStore& P = c->crn2s;
  for (int r = 0; r < P.nshards; r++) {
    if (r == P.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    NWC_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    P.peer[(size_t)r] = (double*)p;
    P.opened[(size_t)r] = 1;
  }
  return 0;
}
void* nwc_triples_cr_shard_ptr(nwc_triples_ctx* c) { return c->d_crn2; }
int nwc_triples_cr_set_peer_ptr(nwc_triples_ctx* c, int rank, void* dev_ptr) {
  if (rank < 0 || rank >= c->crn2s.nshards) { g_err = "bad peer rank"; return 1; }
  c->crn2s.peer[(size_t)rank] = (double*)dev_ptr;
  return 0;
}
static int set_cr_impl(nwc_triples_ctx* c, const Integer* n1_hash, const double* n1, const Integer* n2_hash, const double* n2,
                       const Integer* e2_hash, const double* e2, int rank, int nranks) {
  return guarded(c, [&]() {
    if (!c->eng->trace_only()) NWC_TRY(cudaSetDevice(c->eng->device()));
    const HostState& S = c->S;
    auto total_of = [&](const Integer* h, int kind) -> size_t {
      const Integer nb = h[0];
      if (nb <= 0) return 0;
      Integer key = h[nb], sz = 0;
      if (key < 0) throw Error("nwc_triples: CR offset table does not belong to this tiling");
      if (kind == 1) {          // h2 + noab*(h1 + noab*(h11 + noab*(p4-noab)))
        const Integer h2 = key % S.noab + 1; key /= S.noab; const Integer h1 = key % S.noab + 1; key /= S.noab;
        const Integer h11 = key % S.noab + 1; key /= S.noab;
        if (key >= S.nvab) throw Error("nwc_triples: CR n1 offset table does not belong to this tiling");
        sz = S.rg(key + S.noab + 1) * S.rg(h11) * S.rg(h1) * S.rg(h2);
      } else if (kind == 2) {   // p12-noab + nvab*(h1 + noab*(p5-noab + nvab*(p4-noab)))
        const Integer p12 = key % S.nvab + S.noab + 1; key /= S.nvab; const Integer h1 = key % S.noab + 1; key /= S.noab;
        const Integer p5 = key % S.nvab + S.noab + 1; key /= S.nvab;
        if (key >= S.nvab) throw Error("nwc_triples: CR n2 offset table does not belong to this tiling");
        sz = S.rg(key + S.noab + 1) * S.rg(p5) * S.rg(h1) * S.rg(p12);
      } else {                  // a T2 key
        const Integer h2 = key % S.noab + 1; key /= S.noab; const Integer h1 = key % S.noab + 1; key /= S.noab;
        const Integer p5 = key % S.nvab + S.noab + 1; key /= S.nvab;
        if (key >= S.nvab) throw Error("nwc_triples: CR e2 offset table does not belong to this tiling");
        sz = S.rg(key + S.noab + 1) * S.rg(p5) * S.rg(h1) * S.rg(h2);
      }
      return (size_t)(h[2 * nb] + sz);
    };
    c->crn1_hash.assign(n1_hash, n1_hash + 2 * n1_hash[0] + 1);
    c->crn2_hash.assign(n2_hash, n2_hash + 2 * n2_hash[0] + 1);
    c->cre2_hash.assign(e2_hash, e2_hash + 2 * e2_hash[0] + 1);
    if (upload(&c->d_crn1, &c->n_crn1, n1, total_of(n1_hash, 1), c->eng)) return 1;
    c->crn2s.reset();
    if (nranks <= 1) {
      if (upload(&c->d_crn2, &c->n_crn2, n2, total_of(n2_hash, 2), c->eng)) return 1;
    } else {   // shard offsets of every block (all ranks compute the same table)
      PeerStore& P = c->crn2s;
      const Integer nb = n2_hash[0];
      const Integer total = (Integer)total_of(n2_hash, 2);
      P.nshards = nranks; P.rank = rank;
      P.shard_off.assign((size_t)nb, 0); P.block_n.assign((size_t)nb, 0);
      P.peer.assign((size_t)nranks, nullptr); P.opened.assign((size_t)nranks, 0);
      std::vector<Integer> fill((size_t)nranks, 0);
      for (Integer i = 0; i < nb; i++) {
        const Integer off = n2_hash[nb + 1 + i], next = (i + 1 < nb) ? n2_hash[nb + 2 + i] : total;
        const int owner = (int)(i % nranks);
        P.shard_off[(size_t)i] = fill[(size_t)owner];
        P.block_n[(size_t)i] = next - off;
        fill[(size_t)owner] += next - off;
      }
      if (upload(&c->d_crn2, &c->n_crn2, n2, (size_t)fill[(size_t)rank], c->eng)) return 1;
      P.peer[(size_t)rank] = c->d_crn2;
    }
    // sd_E2_K multiplies by +-2/3 (cr_ccsd_t_E.F:629-721): the factor is folded into the resident copy once
    const size_t ne = total_of(e2_hash, 3);
    std::vector<double>& scaled = c->cre2_scaled;
    scaled.assign(e2 ? ne : 0, 0.0);
    for (size_t i = 0; i < scaled.size(); i++) scaled[i] = (2.0 / 3.0) * e2[i];
    if (upload(&c->d_cre2, &c->n_cre2, e2 ? scaled.data() : nullptr, ne, c->eng)) return 1;
    if (!c->eng->trace_only()) { scaled.clear(); scaled.shrink_to_fit(); }
    return 0;
  });
}

// CR-CCSD(T) sums of tasks `ids` (cr_ccsd_t.F:176-207): sums[4] = (num1, num2, den1, den2) WITHOUT den0; the caller adds
// the scalar of cr_ccsd_t_D and forms  E[T] = num1/(1+den1+den0),  E(T) = num2/(1+den2+den0)  (:260-263) after the sum over
// ranks (nwc_triples_allreduce_sum with n = 4).  per_task (optional): 4 doubles per task.
static int run_cr_ids(nwc_triples_ctx* c, const std::vector<Integer>& ids, const std::vector<long long>* ranges,
                      double sums[4], double* per_task) {
  NWC_TRY(cudaSetDevice(c->eng->device()));
  if (!c->d_crn1 || !c->d_crn2 || !c->d_cre2) { g_err = "nwc_triples_run_cr: call nwc_triples_set_cr first"; return 1; }
  std::vector<double> rows(4 * ids.size() + 4, 0.0);   // four doubles per task
  double dummy[2] = {0.0, 0.0};
  Pipeline pipe(c, dummy, rows.data());
  const char* tp = getenv("NWC_CR_TWO_PASS");   // A/B: the two-pass form (numerators, then denominators; D contracted twice)
  const bool two_pass = tp && *tp == '1';
  pipe.dual = !two_pass;
  for (size_t i = 0; i < ids.size(); i++) {
    const long long a = ranges ? (*ranges)[2 * i] : 0, b = ranges ? (*ranges)[2 * i + 1] : -1;
    if (ranges && b <= a) continue;
    if (!two_pass) {
      emit_tuple_cr(c, &c->klist[7 * (size_t)ids[i]], 2, a, b);
      pipe.emitted((Integer)i);
    } else {
      for (int pass = 0; pass < 2; pass++) {   // row 2i = pass 0 of task i, row 2i+1 = pass 1, two doubles each
        emit_tuple_cr(c, &c->klist[7 * (size_t)ids[i]], pass, a, b);
        pipe.emitted((Integer)(2 * i + pass));
      }
    }
  }
  pipe.finish();
  sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
  for (size_t i = 0; i < ids.size(); i++)
    for (int q = 0; q < 4; q++) {
      sums[q] += rows[4 * i + q];
      if (per_task) per_task[4 * i + q] = rows[4 * i + q];
    }
  return 0;
}

int nwc_triples_run_cr(nwc_triples_ctx* c, Integer first, Integer stride, Integer max_tasks, double sums[4], double* per_task) {
  return guarded(c, [&]() {
    if (stride <= 0) stride = 1;
    const Integer nt = (Integer)(c->klist.size() / 7);
    std::vector<Integer> ids;
    for (Integer k = first < 0 ? 0 : first; k < nt && (max_tasks <= 0 || (Integer)ids.size() < max_tasks); k += stride) ids.push_back(k);
    return run_cr_ids(c, ids, nullptr, sums, per_task);
  });
}

// the same over the static block partition of nwc_triples_run_partition (all four sums are additive over sub-tiles)
int nwc_triples_run_cr_partition(nwc_triples_ctx* c, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                                 double sums[4], double* per_task) {
  return guarded(c, [&]() {
    const Integer nt = (Integer)(c->klist.size() / 7);
    if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
    if (first_task < 0) first_task = 0;
    if (ntasks <= 0 || first_task + ntasks > nt) ntasks = nt - first_task;
    std::vector<Integer> ids;
    for (Integer i = 0; i < ntasks; i++) ids.push_back(first_task + i);
    std::vector<long long> ranges;
    if (!ids.empty()) block_partition(c->S, c->klist, rank, nranks, ids, ranges);
    return run_cr_ids(c, ids, &ranges, sums, per_task);
  });
}

// ---- CR-EOMCCSD(T) (src/tce/cr-eomccsd_t/cr_eomccsd_t.F) ----
// Inputs of the tuple loop :325-493 beyond T1/T2 and (when r0 != 0) the CR-CCSD(T) intermediates of nwc_triples_set_cr:
// the right-hand amplitudes x1 / x2 (the T1 / T2 block structure), the four intermediates of creomsd_t_n2_mem (toggle 1:
// d_i2_1..4; layouts of the CR ones: OFFSET_creomsd_t_n2_mem_{1,2,3,4}_1), the one of q3rexpt2 (d_i3_1, the T2 block
// structure), r0 (r0xx, :134-141) and the excitation energy.  Call after set_state (and set_cr when |r0| >= 1e-7).
static int fetch_host(nwc_triples_ctx* c, const double* dptr, size_t n, std::vector<double>& out) {
  out.assign(n, 0.0);
  if (!n) return 0;
  if (c->eng->trace_only()) { memcpy(out.data(), dptr, n * sizeof(double)); return 0; }
  NWC_TRY(cudaMemcpy(out.data(), dptr, n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int nwc_triples_set_creom(nwc_triples_ctx* c, const Integer* x1_hash, const double* x1, const Integer* x2_hash, const double* x2,
                          const Integer* m1_hash, const double* m1, const Integer* m2_hash, const double* m2,
                          const Integer* m3_hash, const double* m3, const Integer* m4_hash, const double* m4,
                          const Integer* q2_hash, const double* q2, double r0, double excit) {
  return guarded(c, [&]() {
    if (!c->eng->trace_only()) NWC_TRY(cudaSetDevice(c->eng->device()));
    const HostState& S = c->S;
    if (!c->d_t1 || !c->d_t2) { g_err = "nwc_triples_set_creom: call nwc_triples_set_state first"; return 1; }
    const bool lr0 = !(fabs(r0) < 1.0e-7);   // cr_eomccsd_t.F:146-147
    if (lr0 && (!c->d_crn1 || !c->d_crn2 || !c->d_cre2)) { g_err = "nwc_triples_set_creom: r0 != 0 needs the CR-CCSD(T) intermediates (nwc_triples_set_cr)"; return 1; }
    auto same = [](const Integer* h, const std::vector<Integer>& ref) {
      if ((size_t)(2 * h[0] + 1) != ref.size()) return false;
      for (size_t i = 0; i < ref.size(); i++) if (h[i] != ref[i]) return false;
      return true;
    };
    if (!same(x1_hash, S.t1_hash) || !same(x2_hash, S.t2_hash) || !same(q2_hash, S.t2_hash))
      throw Error("nwc_triples_set_creom: x1 / x2 / the q3rexpt2 intermediate must have the T1 / T2 / T2 block structure (irrep_x = 0)");
    if (lr0 && !same(q2_hash, c->cre2_hash)) throw Error("nwc_triples_set_creom: the i1_tt and i1_xt offset tables differ");
    c->eom_r0 = r0; c->eom_excit = excit; c->eom_lr0 = lr0;
    c->x2_hash.assign(x2_hash, x2_hash + 2 * x2_hash[0] + 1);
    c->m1_hash.assign(m1_hash, m1_hash + 2 * m1_hash[0] + 1); c->m2_hash.assign(m2_hash, m2_hash + 2 * m2_hash[0] + 1);
    c->m3_hash.assign(m3_hash, m3_hash + 2 * m3_hash[0] + 1); c->m4_hash.assign(m4_hash, m4_hash + 2 * m4_hash[0] + 1);
    c->eomz_hash.assign(q2_hash, q2_hash + 2 * q2_hash[0] + 1);
    const size_t n1 = store_size(S.t1_hash.data(), S, 1), n2 = store_size(S.t2_hash.data(), S, 2);
    // store sizes of the intermediates: same key decoding as nwc_triples_set_cr
    auto cr_total = [&](const Integer* h, int kind) -> size_t {
      const Integer nb = h[0];
      if (nb <= 0) return 0;
      Integer key = h[nb], sz = 0;
      if (kind == 1) {
        const Integer h2 = key % S.noab + 1; key /= S.noab; const Integer h1 = key % S.noab + 1; key /= S.noab;
        const Integer h11 = key % S.noab + 1; key /= S.noab;
        if (key < 0 || key >= S.nvab) throw Error("nwc_triples_set_creom: hphh offset table does not belong to this tiling");
        sz = S.rg(key + S.noab + 1) * S.rg(h11) * S.rg(h1) * S.rg(h2);
      } else {
        const Integer p12 = key % S.nvab + S.noab + 1; key /= S.nvab; const Integer h1 = key % S.noab + 1; key /= S.noab;
        const Integer p5 = key % S.nvab + S.noab + 1; key /= S.nvab;
        if (key < 0 || key >= S.nvab) throw Error("nwc_triples_set_creom: pphp offset table does not belong to this tiling");
        sz = S.rg(key + S.noab + 1) * S.rg(p5) * S.rg(h1) * S.rg(p12);
      }
      return (size_t)(h[2 * nb] + sz);
    };
    if (upload(&c->d_x2, &c->n_x2, x2, n2, c->eng)) return 1;
    if (upload(&c->d_m1, &c->n_m1, m1, cr_total(m1_hash, 1), c->eng)) return 1;
    if (upload(&c->d_m2, &c->n_m2, m2, cr_total(m2_hash, 2), c->eng)) return 1;
    if (upload(&c->d_m3, &c->n_m3, m3, cr_total(m3_hash, 1), c->eng)) return 1;
    if (upload(&c->d_m4, &c->n_m4, m4, cr_total(m4_hash, 2), c->eng)) return 1;
    // left-hand outer products, combined once: r0*E_1(t2,t1) + q3rexpt2_1(t2,x1) = t2 x (r0*t1 + x1);
    // r0*E_2 (twot = -+2/3, t1 x i1_tt) + q3rexpt2_2 (twot = -+2, t1 x i1_xt) = -+ t1 x (r0*(2/3)*i1_tt + 2*i1_xt)
    std::vector<double>& y1 = c->eom_host[0];
    std::vector<double>& z = c->eom_host[1];
    std::vector<double> tmp;
    y1.assign(x1, x1 + n1);
    if (lr0) { if (fetch_host(c, c->d_t1, n1, tmp)) return 1; for (size_t i = 0; i < n1; i++) y1[i] += r0 * tmp[i]; }
    z.assign(n2, 0.0);
    for (size_t i = 0; i < n2; i++) z[i] = 2.0 * q2[i];
    if (lr0) { if (fetch_host(c, c->d_cre2, n2, tmp)) return 1; for (size_t i = 0; i < n2; i++) z[i] += r0 * tmp[i]; }   // d_cre2 holds (2/3)*i1_tt
    if (upload(&c->d_eomy1, &c->n_eomy1, y1.data(), n1, c->eng)) return 1;
    if (upload(&c->d_eomz, &c->n_eomz, z.data(), n2, c->eng)) return 1;
    // denominators: denex = Delta + omega (the h1 slot reads eps + omega), and unit denominators (1 for the h1 slot, 0 elsewhere)
    std::vector<double>& es = c->eom_host[2];
    std::vector<double>& un = c->eom_host[3];
    std::vector<double>& ze = c->eom_host[4];
    es = S.evl;
    for (double& v : es) v += excit;
    Integer maxr = 1;
    for (Integer b = 1; b <= S.N(); b++) maxr = S.rg(b) > maxr ? S.rg(b) : maxr;
    un.assign((size_t)maxr, 1.0);
    ze.assign((size_t)maxr, 0.0);
    if (upload(&c->d_evl_shift, &c->n_evl_shift, es.data(), es.size(), c->eng)) return 1;
    if (upload(&c->d_unit, &c->n_unit, un.data(), un.size(), c->eng)) return 1;
    if (upload(&c->d_zero, &c->n_zero, ze.data(), ze.size(), c->eng)) return 1;
    if (!c->eng->trace_only()) for (auto& v : c->eom_host) { v.clear(); v.shrink_to_fit(); }
    c->eom_set = true;
    return 0;
  });
}

// sums[4] = (A, B, C, D) = (sum f R R/denex, sum f L R, sum f L R/denex, sum f L L) over the tasks; the caller forms
// num1 = A + B, den1 = C + D and energy1 = num1/(r0^2 + d12 + den1) (cr_eomccsd_t.F:455-464, :564) after the sum over ranks.
static int run_creom_ids(nwc_triples_ctx* c, const std::vector<Integer>& ids, const std::vector<long long>* ranges,
                         double sums[4], double* per_task) {
  if (!c->eng->trace_only()) NWC_TRY(cudaSetDevice(c->eng->device()));
  if (!c->eom_set) { g_err = "nwc_triples_run_creom: call nwc_triples_set_creom first"; return 1; }
  std::vector<double> ex(2 * ids.size() + 2, 0.0), ey(4 * ids.size() + 4, 0.0);
  double dummy[2] = {0.0, 0.0};
  const char* cm = getenv("NWC_CREOM_COMPOSED");
  if (!(cm && *cm == '1')) {   // one dual-energy tuple per task: rows of four = (A, A + C, B, B + D)
    Pipeline pipe(c, dummy, ey.data());
    pipe.dual = true;
    for (size_t i = 0; i < ids.size(); i++) {
      const long long a = ranges ? (*ranges)[2 * i] : 0, b = ranges ? (*ranges)[2 * i + 1] : -1;
      if (ranges && b <= a) continue;
      emit_tuple_creom(c, &c->klist[7 * (size_t)ids[i]], 3, a, b);
      pipe.emitted((Integer)i);
    }
    pipe.finish();
    sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
    for (size_t i = 0; i < ids.size(); i++) {
      const double row[4] = {ey[4 * i], ey[4 * i + 2], ey[4 * i + 1] - ey[4 * i], ey[4 * i + 3] - ey[4 * i + 2]};
      for (int q = 0; q < 4; q++) {
        sums[q] += row[q];
        if (per_task) per_task[4 * i + q] = row[q];
      }
    }
    return 0;
  }
  {   // plain tuples: (A, A + C)
    Pipeline pipe(c, dummy, ex.data());
    for (size_t i = 0; i < ids.size(); i++) {
      const long long a = ranges ? (*ranges)[2 * i] : 0, b = ranges ? (*ranges)[2 * i + 1] : -1;
      if (ranges && b <= a) continue;
      emit_tuple_creom(c, &c->klist[7 * (size_t)ids[i]], 0, a, b);
      pipe.emitted((Integer)i);
    }
    pipe.finish();
  }
  {   // two-sided tuples with unit denominators: row 2i = Y1 -> (B, B + <L,La>), row 2i+1 = Y2 -> (0, <L,Lb>)
    Pipeline pipe(c, dummy, ey.data());
    for (size_t i = 0; i < ids.size(); i++) {
      const long long a = ranges ? (*ranges)[2 * i] : 0, b = ranges ? (*ranges)[2 * i + 1] : -1;
      if (ranges && b <= a) continue;
      for (int w = 1; w <= 2; w++) {
        emit_tuple_creom(c, &c->klist[7 * (size_t)ids[i]], w, a, b);
        pipe.emitted((Integer)(2 * i + (w - 1)));
      }
    }
    pipe.finish();
  }
  sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
  for (size_t i = 0; i < ids.size(); i++) {
    const double A = ex[2 * i], C = ex[2 * i + 1] - ex[2 * i];
    const double B = ey[4 * i], D = (ey[4 * i + 1] - ey[4 * i]) + ey[4 * i + 3];
    const double row[4] = {A, B, C, D};
    for (int q = 0; q < 4; q++) {
      sums[q] += row[q];
      if (per_task) per_task[4 * i + q] = row[q];
    }
  }
  return 0;
}

int nwc_triples_run_creom(nwc_triples_ctx* c, Integer first, Integer stride, Integer max_tasks, double sums[4], double* per_task) {
  return guarded(c, [&]() {
    if (stride <= 0) stride = 1;
    const Integer nt = (Integer)(c->klist.size() / 7);
    std::vector<Integer> ids;
    for (Integer k = first < 0 ? 0 : first; k < nt && (max_tasks <= 0 || (Integer)ids.size() < max_tasks); k += stride) ids.push_back(k);
    return run_creom_ids(c, ids, nullptr, sums, per_task);
  });
}

int nwc_triples_run_creom_partition(nwc_triples_ctx* c, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                                    double sums[4], double* per_task) {
  return guarded(c, [&]() {
    const Integer nt = (Integer)(c->klist.size() / 7);
    if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "bad rank/nranks"; return 1; }
    if (first_task < 0) first_task = 0;
    if (ntasks <= 0 || first_task + ntasks > nt) ntasks = nt - first_task;
    std::vector<Integer> ids;
    for (Integer i = 0; i < ntasks; i++) ids.push_back(first_task + i);
    std::vector<long long> ranges;
    if (!ids.empty()) block_partition(c->S, c->klist, rank, nranks, ids, ranges);
    return run_creom_ids(c, ids, &ranges, sums, per_task);
  });
}

// ---- host-only trace (include/nwc_triples.h): the driver logic above, recorded instead of executed ----
int nwc_triples_trace_tuple(nwc_triples_ctx* c, const Integer t[6], int method) {
  return guarded(c, [&]() {
    if (!c->eng->trace_only()) { g_err = "nwc_triples_trace_tuple needs a context from nwc_triples_create_trace"; return 1; }
    if (!c->d_t1 || !c->d_t2 || !c->d_v2) { g_err = "nwc_triples_trace_tuple: call nwc_triples_set_state first"; return 1; }
    if (method == 0) emit_tuple(c, t);
    else if (method == 1) {
      if (!c->d_y2 || !c->d_y1 || !c->d_f1) { g_err = "nwc_triples_trace_tuple: call nwc_triples_set_lambda first"; return 1; }
      emit_tuple_lambda(c, t);
    } else if (method >= 2 && method <= 4) {
      if (!c->d_crn1 || !c->d_crn2 || !c->d_cre2) { g_err = "nwc_triples_trace_tuple: call nwc_triples_set_cr first"; return 1; }
      emit_tuple_cr(c, t, method - 2);
    } else if (method >= 5 && method <= 8) {
      if (!c->eom_set) { g_err = "nwc_triples_trace_tuple: call nwc_triples_set_creom first"; return 1; }
      emit_tuple_creom(c, t, method - 5);
    } else { g_err = "nwc_triples_trace_tuple: method must be 0..8"; return 1; }
    return 0;
  });
}

int nwc_triples_trace_take(nwc_triples_ctx* c, nwc_trace_rec* out, size_t cap, size_t* n) {
  std::vector<nwc_trace_rec>& tr = c->eng->trace;
  *n = tr.size();
  const size_t m = tr.size() < cap ? tr.size() : cap;
  if (out && m) memcpy(out, tr.data(), m * sizeof(nwc_trace_rec));
  tr.clear();
  return 0;
}

int nwc_triples_run_tuple(nwc_triples_ctx* c, const Integer t[6], double energy[2], double* host_doubles,
                          double* host_singles) {
  return guarded(c, [&]() {
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
    slot_done(c, e.current_slot());
    if (host_doubles) {   // the arena has been rewound but not released: the tiles are still there
      NWC_TRY(cudaMemcpy(host_doubles, dd, sz * sizeof(double), cudaMemcpyDeviceToHost));
      NWC_TRY(cudaMemcpy(host_singles, ds, sz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    energy[0] = out[0];
    energy[1] = out[1];
    return 0;
  });
}

int nwc_triples_set_timing(nwc_triples_ctx* c, int on) { c->eng->timing = on != 0; return 0; }

static void fill_stats(const EngineStats& s, nwc_triples_stats* o, double resident) {
  o->fused_ms = s.fused_ms; o->repack_ms = s.repack_ms;
  o->fused_launches = s.fused_launches; o->repack_launches = s.repack_launches; o->reduce_launches = s.reduce_launches;
  o->work_items = s.work_items; o->descs = s.descs; o->tuples = s.tuples; o->flops = s.flops;
  o->h2d_bytes = (double)s.h2d_bytes; o->d2h_bytes = (double)s.d2h_bytes;
  o->resident_bytes = resident;
  o->pull_ms = s.pull_ms; o->peer_bytes = (double)s.peer_bytes;
  o->pull_launches = s.pull_launches; o->antisym_launches = s.antisym_launches;
}
int nwc_triples_get_stats(nwc_triples_ctx* c, nwc_triples_stats* o, int reset) {
  fill_stats(c->eng->stats, o, 8.0 * (double)(c->n_t1 + c->n_t2 + c->n_v2 + c->n_v2orb));
  if (reset) c->eng->stats = EngineStats();
  return 0;
}
int nwc_triples_timer_start(nwc_triples_ctx* c) {
  return guarded(c, [&]() { cudaSetDevice(c->eng->device()); c->eng->timer_start(); return 0; });
}
int nwc_triples_timer_stop_ms(nwc_triples_ctx* c, double* ms) {
  return guarded(c, [&]() { cudaSetDevice(c->eng->device()); *ms = c->eng->timer_stop_ms(); return 0; });
}
int nwc_host_register(void* ptr, size_t bytes) { NWC_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault)); return 0; }
int nwc_host_unregister(void* ptr) { NWC_TRY(cudaHostUnregister(ptr)); return 0; }
int nwc_compat_get_stats(nwc_triples_stats* o, int reset) {
  return guarded(nullptr, [&]() {
    Engine& e = nwc::compat_engine();
    fill_stats(e.stats, o, 0.0);
    if (reset) e.stats = EngineStats();
    return 0;
  });
}
int nwc_compat_set_timing(int on) { return guarded(nullptr, [&]() { nwc::compat_engine().timing = on != 0; return 0; }); }
int nwc_compat_timer_start(void) { return guarded(nullptr, [&]() { nwc::compat_engine().timer_start(); return 0; }); }
int nwc_compat_timer_stop_ms(double* ms) { return guarded(nullptr, [&]() { *ms = nwc::compat_engine().timer_stop_ms(); return 0; }); }

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
// release the batch arenas (tens of GB after large tuples); the resident stores stay, the next run re-allocates
int nwc_triples_trim(nwc_triples_ctx* c) {
  return guarded(c, [&]() {
    NWC_TRY(cudaSetDevice(c->eng->device()));
    c->eng->trim();
    for (int s = 0; s < 2; s++) { c->v2_built[s].clear(); c->pulled[s].clear(); }
    return 0;
  });
}
int nwc_compat_trim(void) { return guarded(nullptr, [&]() { nwc::compat_engine().trim(); return 0; }); }
int nwc_triples_get_order(nwc_triples_ctx* c) { return c->eng->order(); }
int nwc_triples_set_arena_cap(nwc_triples_ctx* c, size_t bytes) { c->eng->set_arena_cap(bytes); return 0; }

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
  return 0;
}

// sum over ranks of n doubles, in place (host buffer): one ncclAllReduce on the library's stream
int nwc_triples_allreduce_sum(nwc_triples_ctx* c, double* buf, size_t n) {
  if (!c->comm) { if (c->nranks == 1) return 0; g_err = "nccl not initialised"; return 1; }
  NWC_TRY(cudaSetDevice(c->eng->device()));
  if (n > c->n_red) {
    if (c->d_red) cudaFree(c->d_red);
    c->d_red = nullptr; c->n_red = 0;
    NWC_TRY(cudaMalloc((void**)&c->d_red, n * sizeof(double)));
    c->n_red = n;
  }
  cudaStream_t s = c->eng->stream();
  NWC_TRY(cudaMemcpyAsync(c->d_red, buf, n * sizeof(double), cudaMemcpyHostToDevice, s));
  int r = g_nccl.AllReduce(c->d_red, c->d_red, n, NCCL_DOUBLE, NCCL_SUM, c->comm, s);  // replaces ga_dgop (ccsd_t.F:297)
  if (r != 0) { g_err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return 1; }
  NWC_TRY(cudaMemcpyAsync(buf, c->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  NWC_TRY(cudaStreamSynchronize(s));
  return 0;
}

int nwc_triples_allreduce_energy(nwc_triples_ctx* c, double energy[2]) { return nwc_triples_allreduce_sum(c, energy, 2); }

}  // extern "C"
