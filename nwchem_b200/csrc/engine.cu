// Host-side engine: arena, descriptor building for one or many tile tuples, batch launch.
#include "engine.h"
#include <cstring>
#include <algorithm>

namespace nwc {

// ------------------------------------------------------------------------------------------------
void* Arena::alloc(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  while (cur_ < chunks_.size()) {
    Chunk& c = chunks_[cur_];
    if (c.off + bytes <= c.size) {
      void* p = c.base + c.off;
      c.off += bytes;
      used_ += bytes;
      return p;
    }
    cur_++;
  }
  Chunk c;
  c.size = std::max(bytes, min_chunk);
  c.off = 0;
  NWC_CUDA(cudaMalloc((void**)&c.base, c.size));
  chunks_.push_back(c);
  cur_ = chunks_.size() - 1;
  chunks_[cur_].off = bytes;
  used_ += bytes;
  return chunks_[cur_].base;
}
void Arena::reset() {
  for (auto& c : chunks_) c.off = 0;
  cur_ = 0;
  used_ = 0;
}
void Arena::release() {
  for (auto& c : chunks_) cudaFree(c.base);
  chunks_.clear();
  cur_ = 0;
  used_ = 0;
}
size_t Arena::capacity() const {
  size_t s = 0;
  for (auto& c : chunks_) s += c.size;
  return s;
}

// ------------------------------------------------------------------------------------------------
Engine::Engine(int device) : device_(device) {
  NWC_CUDA(cudaSetDevice(device_));
  NWC_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  NWC_CUDA(cudaEventCreate(&ev0_));
  NWC_CUDA(cudaEventCreate(&ev1_));
  NWC_CUDA(cudaEventCreate(&evt0_));
  NWC_CUDA(cudaEventCreate(&evt1_));
}
void Engine::timer_start() { NWC_CUDA(cudaEventRecord(evt0_, stream_)); }
double Engine::timer_stop_ms() {
  NWC_CUDA(cudaEventRecord(evt1_, stream_));
  NWC_CUDA(cudaEventSynchronize(evt1_));
  float ms = 0;
  NWC_CUDA(cudaEventElapsedTime(&ms, evt0_, evt1_));
  return ms;
}
Engine::~Engine() {
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  arena_.release();
  if (d_meta_) cudaFree(d_meta_);
  if (d_jobs_) cudaFree(d_jobs_);
  if (d_ajobs_) cudaFree(d_ajobs_);
  if (h_pin_) cudaFreeHost(h_pin_);
  cudaEventDestroy(ev0_);
  cudaEventDestroy(ev1_);
  cudaEventDestroy(evt0_);
  cudaEventDestroy(evt1_);
  cudaStreamDestroy(stream_);
}

void Engine::begin_tuple(const int R_phys[6]) {
  if (open_) { printf("nwc_triples: begin_tuple while a tuple is open\n"); exit(1); }
  memset(&cur_, 0, sizeof(cur_));
  for (int q = 0; q < 6; q++) {
    cur_.R[q] = R_phys[q];
    cur_.nb[q] = (R_phys[q] + SB - 1) / SB;
  }
  for (int s = 0; s < 9; s++) cur_descs_[s].clear();
  cur_.sdesc_begin = (int)sdescs_.size();
  open_ = true;
}

static inline double prodR(const int R[6]) {
  double p = 1.0;
  for (int q = 0; q < 6; q++) p *= R[q];
  return p;
}

void Engine::add_contraction(int family, int k0, int K7, const OperandView& tsub, const OperandView& v2sub, double tscale,
                             std::vector<PanelSlot>* t_cache, std::vector<PanelSlot>* v_cache) {
  if (!open_ || (family != 1 && family != 2) || k0 < 0 || k0 > 8) { printf("nwc_triples: bad add_contraction\n"); exit(1); }
  // which operand is the G1 (one particle + two holes) one, and the singleton names
  const bool t_is_g1 = (family == 2);
  const int pa = pos_of(family, k0, family == 2 ? N_P4 : N_P6);
  const int hb = pos_of(family, k0, family == 2 ? N_H3 : N_H1);
  const int s = split_id(pa, hb);
  const Split sp = make_split(s);
  const OperandView& g1 = t_is_g1 ? tsub : v2sub;
  const OperandView& g2 = t_is_g1 ? v2sub : tsub;
  const int n1[3] = {DECL[family][k0][sp.g1[0]], DECL[family][k0][sp.g1[1]], DECL[family][k0][sp.g1[2]]};
  const int n2[3] = {DECL[family][k0][sp.g2[0]], DECL[family][k0][sp.g2[1]], DECL[family][k0][sp.g2[2]]};
  const int p1[3] = {sp.g1[0], sp.g1[1], sp.g1[2]}, p2[3] = {sp.g2[0], sp.g2[1], sp.g2[2]};
  std::vector<PanelSlot>* c1 = t_is_g1 ? t_cache : v_cache;
  std::vector<PanelSlot>* c2 = t_is_g1 ? v_cache : t_cache;
  if (K7 <= 0) return;
  auto make_job = [&](const OperandView& op, const int nm[3], const int ps[3], double scale) -> const double* {
    RepackJob j;
    j.src = op.base;
    j.s1 = op.stride[nm[0]]; j.s2 = op.stride[nm[1]]; j.s3 = op.stride[nm[2]]; j.sk = op.kstride;
    j.X1 = cur_.R[ps[0]]; j.X2 = cur_.R[ps[1]]; j.X3 = cur_.R[ps[2]]; j.K = K7;
    j.scale = scale;
    const long long n = panel_doubles(j.X1, j.X2, j.X3, j.K);
    j.dst = (double*)arena_.alloc((size_t)n * sizeof(double));
    max_panel_ = std::max(max_panel_, n);
    jobs_.push_back(j);
    return j.dst;
  };
  auto get_panel = [&](std::vector<PanelSlot>* cache, const OperandView& op, const int nm[3], const int ps[3],
                       double scale) -> const double* {
    if (cache)
      for (const PanelSlot& sl : *cache)
        if (sl.names[0] == nm[0] && sl.names[1] == nm[1] && sl.names[2] == nm[2]) return sl.p;
    PanelSlot sl;
    sl.p = make_job(op, nm, ps, scale);
    sl.names[0] = nm[0]; sl.names[1] = nm[1]; sl.names[2] = nm[2];
    if (cache) cache->push_back(sl);
    return sl.p;
  };
  const double* g1p = get_panel(c1, g1, n1, p1, t_is_g1 ? tscale : 1.0);
  const double* g2p = get_panel(c2, g2, n2, p2, t_is_g1 ? 1.0 : tscale);
  ContrDesc d;
  d.g1 = g1p; d.g2 = g2p;
  d.nk4 = (K7 + 3) / 4;
  d.neg = SIGN[family][k0] < 0 ? 1 : 0;
  cur_descs_[s].push_back(d);
  stats.flops += 2.0 * prodR(cur_.R) * K7;
}

void Engine::add_singles(int k0, const OperandView& t1sub, const OperandView& v2sub) {
  if (!open_ || k0 < 0 || k0 > 8) { printf("nwc_triples: bad add_singles\n"); exit(1); }
  SinglesDesc d;
  memset(&d, 0, sizeof(d));
  d.t1 = t1sub.base;
  d.v2 = v2sub.base;
  for (int q = 0; q < 6; q++) {
    const int name = DECL[0][k0][q];
    d.st1[q] = (int)t1sub.stride[name];
    d.sv2[q] = (int)v2sub.stride[name];
  }
  d.neg = SIGN[0][k0] < 0 ? 1 : 0;
  sdescs_.push_back(d);
  stats.flops += 2.0 * prodR(cur_.R);
}

void Engine::end_tuple(const double* const eps[6], double factor) {
  if (!open_) { printf("nwc_triples: end_tuple without begin_tuple\n"); exit(1); }
  // reference argument order (h1,h2,h3,p4,p5,p6) -> physical positions
  cur_.eps[POS_H1] = eps[0]; cur_.eps[POS_H2] = eps[1]; cur_.eps[POS_H3] = eps[2];
  cur_.eps[POS_P4] = eps[3]; cur_.eps[POS_P5] = eps[4]; cur_.eps[POS_P6] = eps[5];
  cur_.factor = factor;
  int n = (int)descs_.size();
  for (int s = 0; s < 9; s++) {
    cur_.desc_begin[s] = n;
    descs_.insert(descs_.end(), cur_descs_[s].begin(), cur_descs_[s].end());
    n += (int)cur_descs_[s].size();
  }
  cur_.desc_begin[9] = n;
  cur_.sdesc_end = (int)sdescs_.size();
  if (cur_.sdesc_end - cur_.sdesc_begin > 12) { printf("nwc_triples: more than 12 singles terms in one tuple\n"); exit(1); }
  long long items = 1;
  for (int q = 0; q < 6; q++) items *= cur_.nb[q];
  cur_.item_begin = items_;
  cur_.nitems = (int)items;
  items_ += items;
  tuples_.push_back(cur_);
  open_ = false;
}

void Engine::add_antisym(const AntisymJob& job) {
  ajobs_.push_back(job);
  const long long n = (long long)job.n[0] * job.n[1] * job.n[2] * job.n[3];
  if (n > max_ablock_) max_ablock_ = n;
}

void Engine::flush_repack() {
  if (!ajobs_.empty()) {   // the spin-orbital blocks first: the repack jobs below read them (same stream)
    const size_t bytes = ajobs_.size() * sizeof(AntisymJob);
    if (bytes > d_ajobs_cap_) {
      if (d_ajobs_) { NWC_CUDA(cudaStreamSynchronize(stream_)); NWC_CUDA(cudaFree(d_ajobs_)); }
      d_ajobs_cap_ = std::max(bytes * 2, (size_t)1 << 16);
      NWC_CUDA(cudaMalloc(&d_ajobs_, d_ajobs_cap_));
    } else {
      NWC_CUDA(cudaStreamSynchronize(stream_));
    }
    NWC_CUDA(cudaMemcpyAsync(d_ajobs_, ajobs_.data(), bytes, cudaMemcpyHostToDevice, stream_));
    launch_antisym((const AntisymJob*)d_ajobs_, (int)ajobs_.size(), max_ablock_, stream_);
    NWC_CUDA(cudaGetLastError());
    stats.antisym_launches += (long long)((ajobs_.size() + 32767) / 32768);
    stats.h2d_bytes += bytes;
    ajobs_.clear();
    max_ablock_ = 0;
  }
  if (jobs_.empty()) return;
  const size_t bytes = jobs_.size() * sizeof(RepackJob);
  if (bytes > d_jobs_cap_) {
    if (d_jobs_) { NWC_CUDA(cudaStreamSynchronize(stream_)); NWC_CUDA(cudaFree(d_jobs_)); }
    d_jobs_cap_ = std::max(bytes * 2, (size_t)1 << 16);
    NWC_CUDA(cudaMalloc(&d_jobs_, d_jobs_cap_));
  } else {
    // the previous job list may still be in use by an in-flight repack launch
    NWC_CUDA(cudaStreamSynchronize(stream_));
  }
  NWC_CUDA(cudaMemcpyAsync(d_jobs_, jobs_.data(), bytes, cudaMemcpyHostToDevice, stream_));
  if (timing) NWC_CUDA(cudaEventRecord(ev0_, stream_));
  launch_repack((const RepackJob*)d_jobs_, (int)jobs_.size(), max_panel_, stream_);
  NWC_CUDA(cudaGetLastError());
  if (timing) {
    NWC_CUDA(cudaEventRecord(ev1_, stream_));
    NWC_CUDA(cudaEventSynchronize(ev1_));
    float ms = 0;
    NWC_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    stats.repack_ms += ms;
  }
  stats.repack_launches += (long long)((jobs_.size() + 32767) / 32768);
  stats.h2d_bytes += bytes;
  jobs_.clear();
  max_panel_ = 0;
}

void Engine::run(double* energies_out, double* dump_doubles, double* dump_singles) {
  if (open_) { printf("nwc_triples: run with an open tuple\n"); exit(1); }
  const int nt = (int)tuples_.size();
  if (nt == 0) return;
  flush_repack();
  // one metadata buffer: tuples | descs | sdescs | energies | partials
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_t = 0, o_d = al(o_t + nt * sizeof(TupleHdr)), o_s = al(o_d + descs_.size() * sizeof(ContrDesc)),
               o_e = al(o_s + sdescs_.size() * sizeof(SinglesDesc)), o_p = al(o_e + nt * sizeof(double2)),
               total = al(o_p + (size_t)items_ * partials_per_item() * sizeof(double2));
  if (total > d_meta_cap_) {
    if (d_meta_) NWC_CUDA(cudaFree(d_meta_));
    d_meta_cap_ = total + total / 4;
    NWC_CUDA(cudaMalloc(&d_meta_, d_meta_cap_));
  }
  char* dm = (char*)d_meta_;
  NWC_CUDA(cudaMemcpyAsync(dm + o_t, tuples_.data(), nt * sizeof(TupleHdr), cudaMemcpyHostToDevice, stream_));
  if (!descs_.empty())
    NWC_CUDA(cudaMemcpyAsync(dm + o_d, descs_.data(), descs_.size() * sizeof(ContrDesc), cudaMemcpyHostToDevice, stream_));
  if (!sdescs_.empty())
    NWC_CUDA(cudaMemcpyAsync(dm + o_s, sdescs_.data(), sdescs_.size() * sizeof(SinglesDesc), cudaMemcpyHostToDevice, stream_));
  stats.h2d_bytes += nt * sizeof(TupleHdr) + descs_.size() * sizeof(ContrDesc) + sdescs_.size() * sizeof(SinglesDesc);
  if (timing) NWC_CUDA(cudaEventRecord(ev0_, stream_));
  if (dump_doubles)
    launch_fused_dump((const TupleHdr*)(dm + o_t), nt, (const ContrDesc*)(dm + o_d), (const SinglesDesc*)(dm + o_s),
                      (double2*)(dm + o_p), items_, dump_doubles, dump_singles, stream_);
  else {
    bool ragged = false;
    for (const TupleHdr& t : tuples_)
      for (int q = 0; q < 6; q++) ragged = ragged || (t.R[q] % SB != 0);
    launch_fused((const TupleHdr*)(dm + o_t), nt, (const ContrDesc*)(dm + o_d), (const SinglesDesc*)(dm + o_s),
                 (double2*)(dm + o_p), items_, ragged, stream_);
  }
  NWC_CUDA(cudaGetLastError());
  if (timing) NWC_CUDA(cudaEventRecord(ev1_, stream_));
  launch_reduce((const TupleHdr*)(dm + o_t), nt, (const double2*)(dm + o_p), (double2*)(dm + o_e), stream_);
  NWC_CUDA(cudaGetLastError());
  NWC_CUDA(cudaMemcpyAsync(energies_out, dm + o_e, nt * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
  NWC_CUDA(cudaStreamSynchronize(stream_));
  if (timing) {
    float ms = 0;
    NWC_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    stats.fused_ms += ms;
  }
  stats.d2h_bytes += nt * sizeof(double2);
  stats.fused_launches += 1;
  stats.reduce_launches += 1;
  stats.work_items += items_;
  stats.descs += (long long)descs_.size();
  stats.tuples += nt;
  tuples_.clear();
  descs_.clear();
  sdescs_.clear();
  items_ = 0;
}

}  // namespace nwc
