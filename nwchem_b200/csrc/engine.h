// Host-side engine shared by the two ABI tiers: device arena, per-tuple descriptor building, batch launch.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "tables.h"

namespace nwc {

// reference error behaviour: print and exit(1) (src/tce/ccsd_t/header.h:27-37)
#define NWC_CUDA(x)                                                                                       \
  do {                                                                                                    \
    cudaError_t _e = (x);                                                                                 \
    if (_e != cudaSuccess) {                                                                              \
      printf("CUDA CALL FAILED AT LINE %d OF FILE %s error %s\n", __LINE__, __FILE__, cudaGetErrorString(_e)); \
      fflush(stdout);                                                                                     \
      exit(1);                                                                                            \
    }                                                                                                     \
  } while (0)

// bump allocator over a few large cudaMalloc chunks (replaces the size-keyed free lists of memory.cu:74-163)
class Arena {
 public:
  void* alloc(size_t bytes);
  void reset();            // keep chunks, rewind
  void release();          // cudaFree everything
  size_t used() const { return used_; }
  size_t capacity() const;
  size_t min_chunk = (size_t)256 << 20;
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

// a built panel and the permuted-name order of its (x1,x2,x3); lets the native tier share one panel
// between the several kernels one operand pair fires on diagonal tuples
struct PanelSlot {
  const double* p = nullptr;
  int names[3] = {-1, -1, -1};
};

struct EngineStats {
  double fused_ms = 0, repack_ms = 0;   // CUDA-event time of the kernels (when timing enabled)
  long long fused_launches = 0, repack_launches = 0, reduce_launches = 0, antisym_launches = 0;
  long long work_items = 0, descs = 0, tuples = 0;
  double flops = 0;                     // algorithmic FLOPs: 2*prod(R)*K per fired contraction, 2*prod(R) per singles
  size_t h2d_bytes = 0, d2h_bytes = 0;
};

class Engine {
 public:
  explicit Engine(int device);
  ~Engine();
  int device() const { return device_; }
  cudaStream_t stream() const { return stream_; }
  Arena& arena() { return arena_; }

  // ---- tuple building (all pointers are device pointers) ----
  void begin_tuple(const int R_phys[6]);
  bool tuple_open() const { return open_; }
  // family 1 (sd_t_d1_K) or 2 (sd_t_d2_K); k0 = K-1; K7 = range of the contracted tile.
  // t_cache / v_cache (optional): panels already built from the same operand; reused when the index order matches.
  void add_contraction(int family, int k0, int K7, const OperandView& tsub, const OperandView& v2sub,
                       double tscale = 1.0, std::vector<PanelSlot>* t_cache = nullptr,
                       std::vector<PanelSlot>* v_cache = nullptr);
  void add_singles(int k0, const OperandView& t1sub, const OperandView& v2sub);
  // eps: six DEVICE vectors in reference argument order (h1,h2,h3,p4,p5,p6)
  void end_tuple(const double* const d_eps_h1h2h3p4p5p6[6], double factor);
  int pending_tuples() const { return (int)tuples_.size(); }
  size_t pending_items() const { return (size_t)items_; }

  // ---- execution: runs every pending tuple in one batch; energies[2*i..] = (E1,E2) of tuple i ----
  void run(double* energies_out, double* dump_doubles = nullptr, double* dump_singles = nullptr);
  void flush_repack();      // launch pending antisym + repack jobs now (asynchronous)
  // `2eorb`: queue the construction of one dense spin-orbital V2 block (job.dst must come from arena())
  void add_antisym(const AntisymJob& job);

  EngineStats stats;
  bool timing = false;
  void timer_start();
  double timer_stop_ms();

 private:
  int device_;
  cudaStream_t stream_;
  Arena arena_;
  bool open_ = false;
  TupleHdr cur_{};
  std::vector<ContrDesc> cur_descs_[9];
  std::vector<TupleHdr> tuples_;
  std::vector<ContrDesc> descs_;
  std::vector<SinglesDesc> sdescs_;
  std::vector<RepackJob> jobs_;
  std::vector<AntisymJob> ajobs_;
  long long max_ablock_ = 0;
  void* d_ajobs_ = nullptr; size_t d_ajobs_cap_ = 0;
  long long max_panel_ = 0;
  long long items_ = 0;
  cudaEvent_t ev0_, ev1_, evt0_, evt1_;
  // small reusable device buffers for descriptor uploads
  void* d_meta_ = nullptr; size_t d_meta_cap_ = 0;
  void* d_jobs_ = nullptr; size_t d_jobs_cap_ = 0;
  void* h_pin_ = nullptr; size_t h_pin_cap_ = 0;
};

}  // namespace nwc
