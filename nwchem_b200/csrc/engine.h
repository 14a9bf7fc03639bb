// Host-side engine shared by the two ABI tiers: device arenas, per-tuple descriptor building, batch launch.
//
// Two batch slots (arena + metadata buffer + pinned result buffer + events each): submit() queues everything a batch
// needs on the stream and returns at once, so the host walks the driver logic of the NEXT batch (host_driver.h) while the
// GPU runs the current one; collect() waits for a slot and hands out its per-tuple energies.  run() = submit + collect.
//
// Errors are C++ exceptions (nwc::Error).  The Tier-2 entry points turn them into a status + nwc_triples_last_error();
// the Tier-1 compat layer prints and exits like the reference does (src/tce/ccsd_t/header.h:27-37).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "tables.h"
#include "../../include/nwc_triples.h"

namespace nwc {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define NWC_CUDA(x)                                                                                        \
  do {                                                                                                     \
    cudaError_t _e = (x);                                                                                  \
    if (_e != cudaSuccess)                                                                                 \
      throw ::nwc::Error(std::string("CUDA CALL FAILED AT LINE ") + std::to_string(__LINE__) + " OF FILE " + \
                         __FILE__ + " error " + cudaGetErrorString(_e) + " (" #x ")");                      \
  } while (0)

// bump allocator over a few large cudaMalloc chunks (replaces the size-keyed free lists of memory.cu:74-163)
class Arena {
 public:
  void* alloc(size_t bytes);
  void reset(bool compact = false);   // rewind; compact: merge several chunks into one (frees memory: nothing may still point into it)
  void release();          // cudaFree everything
  size_t used() const { return used_; }
  size_t capacity() const;
  size_t min_chunk = (size_t)1 << 30;   // few, large chunks: cudaMalloc synchronises the device, so growth must stop early
  size_t max_bytes = (size_t)150 << 30;   // growth cap: beyond it alloc() throws instead of driving the GPU out of memory
 private:
  struct Chunk { char* base; size_t size; size_t off; };
  std::vector<Chunk> chunks_;
  size_t cur_ = 0, used_ = 0;
};

// a strided view of a contraction operand: element strides per permuted index name (tables.h N_*) and for k
struct OperandView {
  const double* base = nullptr;   // DEVICE pointer
  long long stride[6] = {0, 0, 0, 0, 0, 0};
  long long kstride = 0;
};

// One contracted tile of a contraction group: sd_t_d1_K / sd_t_d2_K called for one h7b / p7b tile.  All calls of one
// kernel K within a tuple hit the same split with the same external ranges, so their K ranges are concatenated into
// ONE panel pair and one descriptor: K is padded to a multiple of 8 once per group instead of once per tile
// (uracil: 5 x 38/39 -> 192 instead of 200 k values).
struct Segment {
  int K = 0;
  OperandView t, v;
  double tscale = 1.0;
};
// a built group panel: permuted-name order of its (x1,x2,x3) and the source block of every segment; lets the kernels
// one operand list fires on diagonal tuples share a panel
struct GroupPanel {
  const double* p = nullptr;
  int names[3] = {-1, -1, -1};
  std::vector<const double*> bases;
  std::vector<double> scales;
};

struct EngineStats {
  double fused_ms = 0, repack_ms = 0, pull_ms = 0;   // CUDA-event time of the kernels (when timing enabled)
  long long fused_launches = 0, repack_launches = 0, reduce_launches = 0, antisym_launches = 0, pull_launches = 0;
  long long work_items = 0, descs = 0, tuples = 0;
  double flops = 0;                     // algorithmic FLOPs: 2*prod(R)*K per fired contraction, 2*prod(R) per singles
  size_t h2d_bytes = 0, d2h_bytes = 0;
  size_t peer_bytes = 0;                // bytes pulled from other GPUs' shards over NVLink
};

class Engine {
 public:
  explicit Engine(int device);   // device < 0: host-only trace engine (include/nwc_triples.h nwc_triples_create_trace)
  ~Engine();
  int device() const { return device_; }
  // trace engine: the add_* calls record their operand descriptors here instead of building panels; nothing launches
  bool trace_only() const { return device_ < 0; }
  std::vector<nwc_trace_rec> trace;
  cudaStream_t stream() const { return stream_; }
  Arena& arena() { return slots_[cur_].arena; }   // arena of the batch being built
  int current_slot() const { return cur_; }
  void set_arena_cap(size_t bytes) { for (auto& s : slots_) s.arena.max_bytes = bytes; }
  // Index order inside the operand panels' 64-row blocks (tables.h make_split): 0 = holes first, 1 = particles first.
  // Padding rows of ragged tiles are skipped at a granularity of 4 / 2 / 1 values for the first / second / third
  // index of a group, so the less ragged index type goes first.  Fixed for all tuples of a batch.
  void set_order(int order);
  int order() const { return order_; }
  // executed / useful 8x8 DMMA blocks for a tuple of these ranges (physical order h3,h2,h1,p6,p5,p4) under `order`
  static double padding_cost(const int R_phys[6], int order);

  // ---- tuple building (all pointers are device pointers) ----
  void begin_tuple(const int R_phys[6]);
  bool tuple_open() const { return open_; }
  // family 1 (sd_t_d1_K) or 2 (sd_t_d2_K); k0 = K-1; segs = the contracted tiles (h7b / p7b) of this kernel in this
  // tuple, concatenated along K.  t_cache / v_cache (optional): group panels already built; reused when the index
  // order and the segments' source blocks match.
  // side 0: the tuple's doubles tile; side 1 (Lambda-CCSD(T)): the LEFT-hand doubles tile, accumulated separately and
  // paired with the side-0 tile in the energy pass (a tuple with side-1 contractions is "two-sided")
  void add_contraction_group(int family, int k0, const Segment* segs, int nseg, std::vector<GroupPanel>* t_cache = nullptr,
                             std::vector<GroupPanel>* v_cache = nullptr, int side = 0);
  void set_two_sided() { two_sided_ = true; }   // even if no left-hand contraction fires (its tile is then the outer products)
  // Dual-energy two-sided tuple (CR-CCSD(T) in one pass): the OP_SIDE0 outer products form a FOURTH tile E instead of
  // being added to the side-0 tile M, and the kernel returns two energy pairs, (<M,D>, <M,D+S>) and (<E,D>, <E,D+S>).
  // A batch holds only dual tuples or none; its results come out as [pair 0 of every tuple | pair 1 of every tuple].
  void set_dual() { two_sided_ = true; dual_ = true; }
  // CR-EOMCCSD(T) form of a dual tuple: ONE contraction tile R (side 1) and the outer-product tile L (singles); pair 0 =
  // (<R,R>, <R,R+L>) with the tuple's denominators, pair 1 = (sum f L R, sum f L (R+L)) without denominators
  void set_dual_eom() { two_sided_ = true; dual_ = true; eom_ = true; }
  // one contracted tile on its own (K7 = its range)
  void add_contraction(int family, int k0, int K7, const OperandView& tsub, const OperandView& v2sub, double tscale = 1.0) {
    Segment sg;
    sg.K = K7; sg.t = tsub; sg.v = v2sub; sg.tscale = tscale;
    add_contraction_group(family, k0, &sg, 1);
  }
  void add_singles(int k0, const OperandView& t1sub, const OperandView& v2sub);
  // generic outer-product term  +-a[sum g*sa] * b[sum g*sb]  (element strides per PHYSICAL position h3,h2,h1,p6,p5,p4;
  // `a` carries one hole and one particle index, `b` the other four).  target: OP_SINGLES = the singles tile;
  // OP_SIDE1 = the side-1 doubles tile of a two-sided tuple (Lambda-CCSD(T): y2*f); OP_SIDE0 = its side-0 tile
  // (CR-CCSD(T): the denominator tile E).  Doubles-bound terms are added before the energy pass.
  enum { OP_SINGLES = 0, OP_SIDE1 = 1, OP_SIDE0 = 2 };
  void add_outer_product(const double* a, const int sa[6], const double* b, const int sb[6], bool negative, int target);
  // eps: six DEVICE vectors in reference argument order (h1,h2,h3,p4,p5,p6).  [item_lo, item_hi) restricts the launch
  // to a sub-range of the tuple's 4^6 sub-tiles (linear index, h3 block fastest, p4 block slowest; item_hi < 0 = all):
  // energies are additive over sub-tiles, so a tuple can be shared between GPUs or evaluated slab by slab.
  void end_tuple(const double* const d_eps_h1h2h3p4p5p6[6], double factor, long long item_lo = 0, long long item_hi = -1);
  static long long tuple_items(const int R_phys[6]);
  int pending_tuples() const { return (int)tuples_.size(); }
  size_t pending_items() const { return (size_t)items_; }

  // ---- execution ----
  // queue every pending tuple of the current slot (pull -> antisym -> repack -> fused -> reduce -> D2H) and switch to
  // the other slot, which must have been collected.  Returns the slot to collect, or -1 if nothing was pending.
  int submit(double* dump_doubles = nullptr, double* dump_singles = nullptr);
  // wait for a submitted slot; energies_out[2*i..] = (E1,E2) of its tuple i.  Rewinds the slot's arena.
  // compact: also merge a fragmented arena into one chunk (only when nothing allocated from it is read afterwards)
  void collect(int slot, double* energies_out, bool compact = false);
  int slot_tuples(int slot) const { return slots_[slot].ntuples; }   // result pairs of the slot (2 per tuple for a dual batch)
  bool slot_busy(int slot) const { return slots_[slot].busy; }
  // synchronous convenience: submit + collect
  void run(double* energies_out, double* dump_doubles = nullptr, double* dump_singles = nullptr);
  // error recovery: drop everything pending / in flight and rewind both arenas
  void abort();
  // give the batch arenas and metadata buffers back to the device (they are re-allocated on demand)
  void trim();
  void flush_prep();      // launch pending pull + antisym + repack jobs now (asynchronous)
  // `2eorb`: queue the construction of one dense spin-orbital V2 block (job.dst must come from arena())
  void add_antisym(const AntisymJob& job);
  // sharded stores: queue the pull of one remote block into the arena
  void add_copy(const CopyJob& job);

  EngineStats stats;
  bool timing = false;
  void timer_start();
  double timer_stop_ms();

 private:
  struct Slot {
    Arena arena;
    void* d_meta = nullptr; size_t d_meta_cap = 0;
    double2* h_out = nullptr; size_t h_out_cap = 0;   // pinned
    char* h_stage = nullptr; size_t h_stage_cap = 0, h_stage_off = 0;   // pinned staging of job lists + metadata
    cudaEvent_t done = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // pull, repack, fused: begin/end
    bool timed[3] = {false, false, false};
    int ntuples = 0;
    bool busy = false;
  };
  int device_;
  cudaStream_t stream_;
  Slot slots_[2];
  int cur_ = 0;
  bool open_ = false;
  TupleHdr cur_hdr_{};
  std::vector<ContrDesc> cur_descs_[2][9];
  bool two_sided_ = false, dual_ = false, eom_ = false;
  std::vector<TupleHdr> tuples_;
  std::vector<ContrDesc> descs_;
  std::vector<SinglesDesc> sdescs_;
  std::vector<SinglesDesc> cur_sd_singles_, cur_sd_doubles_, cur_sd_side0_;   // of the tuple being built
  std::vector<RepackJob> jobs_;
  std::vector<AntisymJob> ajobs_;
  std::vector<CopyJob> cjobs_;
  long long max_ablock_ = 0, max_panel_ = 0, max_copy_ = 0;
  long long items_ = 0;
  int max_chunks_ = 1;
  int order_ = 0;
  cudaEvent_t evt0_, evt1_;
  // host bytes -> pinned staging of the current slot -> `dst` (device), truly asynchronous
  void upload(void* dst, const void* host, size_t bytes);
  void* upload_jobs(const void* host, size_t bytes);
};

}  // namespace nwc
