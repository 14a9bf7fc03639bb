// Host-side engine: arenas, descriptor building for one or many tile tuples, asynchronous batch launch.
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
  if (capacity() + c.size > max_bytes)
    throw Error("nwc_triples: batch arena would exceed its cap (" + std::to_string(max_bytes >> 20) +
                " MiB); lower the batch budget (nwc_triples_set_batch_bytes)");
  NWC_CUDA(cudaMalloc((void**)&c.base, c.size));
  chunks_.push_back(c);
  cur_ = chunks_.size() - 1;
  chunks_[cur_].off = bytes;
  used_ += bytes;
  return chunks_[cur_].base;
}
void Arena::reset(bool compact) {
  // (compact) A batch that needed several chunks leaves a fragmented arena, and the next batch (other sizes, other order) may
  // not fit it the same way: it would grow again, and every cudaMalloc synchronises the device under the running
  // batch.  Coalesce once into one chunk with head-room; later batches of similar size then never allocate.
  if (compact && chunks_.size() > 1) {
    size_t total = capacity();
    total += total / 16;   // little head-room: the resident stores may leave few GB (130 GB of 180 for (H2O)10 on one GPU)
    if (total > max_bytes) total = max_bytes;
    for (auto& c : chunks_) cudaFree(c.base);
    chunks_.clear();
    Chunk c;
    c.size = total; c.off = 0; c.base = nullptr;
    if (cudaMalloc((void**)&c.base, c.size) == cudaSuccess) chunks_.push_back(c);
    else cudaGetLastError();   // fall back to growing on demand
  }
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
  if (trace_only()) { stream_ = nullptr; evt0_ = evt1_ = nullptr; return; }
  NWC_CUDA(cudaSetDevice(device_));
  NWC_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  NWC_CUDA(cudaEventCreate(&evt0_));
  NWC_CUDA(cudaEventCreate(&evt1_));
  for (Slot& s : slots_) {
    NWC_CUDA(cudaEventCreate(&s.done));
    for (cudaEvent_t& e : s.ev) NWC_CUDA(cudaEventCreate(&e));
  }
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
  if (trace_only()) return;
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  for (Slot& s : slots_) {
    s.arena.release();
    if (s.d_meta) cudaFree(s.d_meta);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.h_stage) cudaFreeHost(s.h_stage);
    cudaEventDestroy(s.done);
    for (cudaEvent_t e : s.ev) cudaEventDestroy(e);
  }
  cudaEventDestroy(evt0_);
  cudaEventDestroy(evt1_);
  cudaStreamDestroy(stream_);
}

void Engine::trim() {
  abort();
  if (trace_only()) return;
  for (Slot& s : slots_) {
    s.arena.release();
    if (s.d_meta) { cudaFree(s.d_meta); s.d_meta = nullptr; s.d_meta_cap = 0; }
  }
}

void Engine::abort() {
  if (!trace_only()) {
    cudaStreamSynchronize(stream_);
    cudaGetLastError();
  }
  trace.clear();
  open_ = false;
  tuples_.clear(); descs_.clear(); sdescs_.clear(); jobs_.clear(); ajobs_.clear(); cjobs_.clear();
  items_ = 0; max_chunks_ = 1; max_ablock_ = max_panel_ = max_copy_ = 0;
  for (Slot& s : slots_) { s.arena.reset(); s.busy = false; s.ntuples = 0; s.h_stage_off = 0; s.timed[0] = s.timed[1] = s.timed[2] = false; }
}

void Engine::set_order(int order) {
  order = order ? 1 : 0;
  if (order == order_) return;
  if (open_ || !tuples_.empty() || !jobs_.empty()) throw Error("nwc_triples: the panel index order cannot change inside a batch");
  order_ = order;
}

// Fraction of the padded 8x8 blocks the K loops execute: an index with n in-range values in a 4-wide block keeps
// 1 (first in its group), ceil(n/2)/2 (second) or n/4 (third) of the rows; summed over the blocks of each range and
// multiplied over the six indices of each split.  Divided by the same with every n/4 it is the executed / useful ratio.
double Engine::padding_cost(const int R[6], int order) {
  auto S = [](int range, int role) {
    const int nb = (range + 3) / 4;
    double tot = 0;
    for (int b = 0; b < nb; b++) {
      const int n = range - 4 * b < 4 ? range - 4 * b : 4;
      tot += role == 0 ? 1.0 : role == 1 ? ((n + 1) / 2) / 2.0 : n / 4.0;
    }
    return tot / nb;
  };
  double ex = 0, ideal = 1;
  for (int q = 0; q < 6; q++) ideal *= S(R[q], 2);
  for (int s = 0; s < 9; s++) {
    const Split sp = make_split(s, order);
    double e = 1;
    for (int r = 0; r < 3; r++) e *= S(R[sp.g1[r]], r) * S(R[sp.g2[r]], r);
    ex += e;
  }
  return ex / (9.0 * ideal);
}

void Engine::begin_tuple(const int R_phys[6]) {
  if (open_) throw Error("nwc_triples: begin_tuple while a tuple is open");
  if (slots_[cur_].busy) throw Error("nwc_triples: the batch slot being built is still in flight (collect it first)");
  memset(&cur_hdr_, 0, sizeof(cur_hdr_));
  for (int q = 0; q < 6; q++) {
    if (R_phys[q] <= 0) throw Error("nwc_triples: tuple with an empty tile range");
    cur_hdr_.R[q] = R_phys[q];
    cur_hdr_.nb[q] = (R_phys[q] + SB - 1) / SB;
  }
  for (int side = 0; side < 2; side++)
    for (int s = 0; s < 9; s++) cur_descs_[side][s].clear();
  two_sided_ = false;
  dual_ = false;
  eom_ = false;
  cur_sd_singles_.clear();
  cur_sd_doubles_.clear();
  cur_sd_side0_.clear();
  open_ = true;
}

static inline double prodR(const int R[6]) {
  double p = 1.0;
  for (int q = 0; q < 6; q++) p *= R[q];
  return p;
}

long long Engine::tuple_items(const int R_phys[6]) {
  long long items = 1;
  for (int q = 0; q < 6; q++) items *= (R_phys[q] + SB - 1) / SB;
  return items;
}

void Engine::add_contraction_group(int family, int k0, const Segment* segs, int nseg, std::vector<GroupPanel>* t_cache,
                                   std::vector<GroupPanel>* v_cache, int side) {
  if (!open_ || (family != 1 && family != 2) || k0 < 0 || k0 > 8 || side < 0 || side > 1) throw Error("nwc_triples: bad add_contraction");
  if (side == 1) two_sided_ = true;
  if (trace_only()) {
    for (int i = 0; i < nseg; i++) {
      nwc_trace_rec r{};
      r.kind = family; r.k0 = k0; r.side = side; r.K = segs[i].K; r.a = segs[i].t.base; r.b = segs[i].v.base;
      for (int q = 0; q < 6; q++) { r.sa[q] = segs[i].t.stride[q]; r.sb[q] = segs[i].v.stride[q]; }
      r.ka = segs[i].t.kstride; r.kb = segs[i].v.kstride; r.scale = segs[i].tscale;
      trace.push_back(r);
    }
    return;
  }
  // which operand is the G1 (one particle + two holes) one, and the singleton names
  const bool t_is_g1 = (family == 2);
  const int pa = pos_of(family, k0, family == 2 ? N_P4 : N_P6);
  const int hb = pos_of(family, k0, family == 2 ? N_H3 : N_H1);
  const int s = split_id(pa, hb);
  const Split sp = make_split(s, order_);
  const int n1[3] = {DECL[family][k0][sp.g1[0]], DECL[family][k0][sp.g1[1]], DECL[family][k0][sp.g1[2]]};
  const int n2[3] = {DECL[family][k0][sp.g2[0]], DECL[family][k0][sp.g2[1]], DECL[family][k0][sp.g2[2]]};
  const int p1[3] = {sp.g1[0], sp.g1[1], sp.g1[2]}, p2[3] = {sp.g2[0], sp.g2[1], sp.g2[2]};
  std::vector<GroupPanel>* c1 = t_is_g1 ? t_cache : v_cache;
  std::vector<GroupPanel>* c2 = t_is_g1 ? v_cache : t_cache;
  long long Ktot = 0;
  for (int i = 0; i < nseg; i++) Ktot += segs[i].K > 0 ? segs[i].K : 0;
  if (Ktot <= 0) return;
  if (Ktot > 2000000) throw Error("nwc_triples: contracted range too long");
  // one panel per operand for the whole group, one repack job per segment
  auto build = [&](bool is_g1, const int nm[3], const int ps[3], std::vector<GroupPanel>* cache) -> const double* {
    const bool is_t = (is_g1 == t_is_g1);
    if (cache)
      for (const GroupPanel& gp : *cache) {
        if (gp.names[0] != nm[0] || gp.names[1] != nm[1] || gp.names[2] != nm[2] || (int)gp.bases.size() != nseg) continue;
        bool same = true;
        for (int i = 0; i < nseg && same; i++)
          same = gp.bases[(size_t)i] == (is_t ? segs[i].t.base : segs[i].v.base) && gp.scales[(size_t)i] == (is_t ? segs[i].tscale : 1.0);
        if (same) return gp.p;
      }
    const int X1 = cur_hdr_.R[ps[0]], X2 = cur_hdr_.R[ps[1]], X3 = cur_hdr_.R[ps[2]];
    const long long n = panel_doubles(X1, X2, X3, (int)Ktot);
    double* dst = (double*)arena().alloc((size_t)n * sizeof(double));
    GroupPanel gp;
    gp.p = dst; gp.names[0] = nm[0]; gp.names[1] = nm[1]; gp.names[2] = nm[2];
    int k_off = 0, last = -1;
    for (int i = 0; i < nseg; i++) if (segs[i].K > 0) last = i;
    for (int i = 0; i < nseg; i++) {
      const OperandView& op = is_t ? segs[i].t : segs[i].v;
      gp.bases.push_back(op.base);
      gp.scales.push_back(is_t ? segs[i].tscale : 1.0);
      if (segs[i].K <= 0) continue;
      RepackJob j;
      j.src = op.base; j.dst = dst;
      j.s1 = op.stride[nm[0]]; j.s2 = op.stride[nm[1]]; j.s3 = op.stride[nm[2]]; j.sk = op.kstride;
      j.X1 = X1; j.X2 = X2; j.X3 = X3; j.K = segs[i].K;
      j.k_off = k_off;
      j.k_end = (i == last) ? (int)((Ktot + 4 * KPL - 1) / (4 * KPL)) * 4 * KPL : k_off + segs[i].K;
      j.scale = is_t ? segs[i].tscale : 1.0;
      max_panel_ = std::max(max_panel_, panel_doubles(X1, X2, X3, j.k_end - (j.k_off / (4 * KPL)) * 4 * KPL));
      jobs_.push_back(j);
      k_off += segs[i].K;
    }
    if (cache) cache->push_back(gp);
    return dst;
  };
  ContrDesc d;
  d.g1 = build(true, n1, p1, c1);
  d.g2 = build(false, n2, p2, c2);
  d.nk4 = (int)((Ktot + 3) / 4);
  d.neg = SIGN[family][k0] < 0 ? 1 : 0;
  cur_descs_[side][s].push_back(d);
  cur_hdr_.factor += 2.0 * prodR(cur_hdr_.R) * (double)Ktot;   // FLOPs of the whole tuple, parked here until end_tuple
}

void Engine::add_singles(int k0, const OperandView& t1sub, const OperandView& v2sub) {
  if (!open_ || k0 < 0 || k0 > 8) throw Error("nwc_triples: bad add_singles");
  if (trace_only()) {
    nwc_trace_rec r{};
    r.kind = 0; r.k0 = k0; r.a = t1sub.base; r.b = v2sub.base; r.scale = 1.0;
    for (int q = 0; q < 6; q++) { r.sa[q] = t1sub.stride[q]; r.sb[q] = v2sub.stride[q]; }
    trace.push_back(r);
    return;
  }
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
  cur_sd_singles_.push_back(d);
  cur_hdr_.factor += 2.0 * prodR(cur_hdr_.R);
}

void Engine::add_outer_product(const double* a, const int sa[6], const double* b, const int sb[6], bool negative, int target) {
  if (!open_) throw Error("nwc_triples: add_outer_product outside a tuple");
  if (target != OP_SINGLES && target != OP_SIDE1 && target != OP_SIDE0) throw Error("nwc_triples: bad outer-product target");
  SinglesDesc d;
  memset(&d, 0, sizeof(d));
  d.t1 = a;
  d.v2 = b;
  int na = 0, nb = 0;
  for (int q = 0; q < 6; q++) {
    d.st1[q] = sa[q];
    d.sv2[q] = sb[q];
    na += sa[q] != 0; nb += sb[q] != 0;
    if (sa[q] != 0 && sb[q] != 0) throw Error("nwc_triples: outer product operands share an index");
  }
  // (a stride may legitimately be 0 only for an index the operand lacks; ranges of 1 still carry stride >= 1)
  if (na != 2 || nb != 4) throw Error("nwc_triples: outer product needs a 2-index and a 4-index operand");
  d.neg = negative ? 1 : 0;
  if (trace_only()) {
    nwc_trace_rec r{};
    r.kind = 3; r.side = target; r.neg = d.neg; r.a = a; r.b = b; r.scale = 1.0;
    for (int q = 0; q < 6; q++) { r.sa[q] = sa[q]; r.sb[q] = sb[q]; }
    trace.push_back(r);
    return;
  }
  (target == OP_SIDE0 ? cur_sd_side0_ : target == OP_SIDE1 ? cur_sd_doubles_ : cur_sd_singles_).push_back(d);
  cur_hdr_.factor += 2.0 * prodR(cur_hdr_.R);
}

void Engine::end_tuple(const double* const eps[6], double factor, long long item_lo, long long item_hi) {
  if (!open_) throw Error("nwc_triples: end_tuple without begin_tuple");
  if (trace_only()) {
    nwc_trace_rec r{};
    r.kind = 9; r.K = two_sided_ ? (dual_ ? (eom_ ? 3 : 2) : 1) : 0; r.scale = factor;
    for (int q = 0; q < 6; q++) r.sa[q] = cur_hdr_.R[q];
    r.sb[0] = item_lo; r.sb[1] = item_hi;
    r.a = eps[0]; r.b = eps[1];                                   // orbital-energy vectors: h1, h2, ...
    r.sb[2] = (long long)(uintptr_t)eps[2]; r.sb[3] = (long long)(uintptr_t)eps[3];   // ... h3, p4,
    r.sb[4] = (long long)(uintptr_t)eps[4]; r.sb[5] = (long long)(uintptr_t)eps[5];   // p5, p6
    trace.push_back(r);
    open_ = false;
    return;
  }
  // reference argument order (h1,h2,h3,p4,p5,p6) -> physical positions
  cur_hdr_.eps[POS_H1] = eps[0]; cur_hdr_.eps[POS_H2] = eps[1]; cur_hdr_.eps[POS_H3] = eps[2];
  cur_hdr_.eps[POS_P4] = eps[3]; cur_hdr_.eps[POS_P5] = eps[4]; cur_hdr_.eps[POS_P6] = eps[5];
  const double tuple_flops = cur_hdr_.factor;
  cur_hdr_.factor = factor;
  int n = (int)descs_.size();
  for (int s = 0; s < 9; s++) {
    cur_hdr_.desc_begin[s] = n;
    descs_.insert(descs_.end(), cur_descs_[0][s].begin(), cur_descs_[0][s].end());
    n += (int)cur_descs_[0][s].size();
  }
  cur_hdr_.desc_begin[9] = n;
  for (int s = 0; s < 9; s++) {
    cur_hdr_.desc2_begin[s] = n;
    descs_.insert(descs_.end(), cur_descs_[1][s].begin(), cur_descs_[1][s].end());
    n += (int)cur_descs_[1][s].size();
  }
  cur_hdr_.desc2_begin[9] = n;
  if (!two_sided_ && !cur_sd_side0_.empty()) { open_ = false; throw Error("nwc_triples: side-0 outer products need a two-sided tuple"); }
  cur_hdr_.two_sided = two_sided_ ? 1 + (int)cur_sd_side0_.size() : 0;   // kernels.cuh TupleHdr
  if (eom_) {
    if (!cur_sd_side0_.empty() || !cur_sd_doubles_.empty()) { open_ = false; throw Error("nwc_triples: a CR-EOMCCSD(T) tuple takes its outer products in the singles tile"); }
    cur_hdr_.two_sided += 64;
  }
  if (dual_) cur_hdr_.two_sided = -cur_hdr_.two_sided;
  cur_hdr_.sdesc_begin = (int)sdescs_.size();
  sdescs_.insert(sdescs_.end(), cur_sd_side0_.begin(), cur_sd_side0_.end());
  sdescs_.insert(sdescs_.end(), cur_sd_doubles_.begin(), cur_sd_doubles_.end());
  cur_hdr_.sdesc_mid = (int)sdescs_.size();
  sdescs_.insert(sdescs_.end(), cur_sd_singles_.begin(), cur_sd_singles_.end());
  cur_hdr_.sdesc_end = (int)sdescs_.size();
  open_ = false;
  const int max_terms = two_sided_ ? MAX_SINGLES_TERMS_2S : MAX_SINGLES_TERMS;
  if (cur_hdr_.sdesc_end - cur_hdr_.sdesc_begin > max_terms)
    throw Error("nwc_triples: more than " + std::to_string(max_terms) + " outer-product terms in one tuple");
  const long long all = tuple_items(cur_hdr_.R);
  if (item_hi < 0 || item_hi > all) item_hi = all;
  if (item_lo < 0) item_lo = 0;
  if (item_lo > item_hi) item_lo = item_hi;
  cur_hdr_.item_begin = items_;
  cur_hdr_.nitems = (int)(item_hi - item_lo);
  cur_hdr_.item_first = (int)item_lo;
  items_ += cur_hdr_.nitems;
  max_chunks_ = std::max(max_chunks_, reduce_chunks(cur_hdr_.nitems));
  stats.flops += tuple_flops * ((double)cur_hdr_.nitems / (double)all);
  tuples_.push_back(cur_hdr_);
}

void Engine::add_antisym(const AntisymJob& job) {
  ajobs_.push_back(job);
  const long long n = (long long)job.n[0] * job.n[1] * job.n[2] * job.n[3];
  if (n > max_ablock_) max_ablock_ = n;
}

void Engine::add_copy(const CopyJob& job) {
  cjobs_.push_back(job);
  if (job.n > max_copy_) max_copy_ = job.n;
  stats.peer_bytes += (size_t)job.n * sizeof(double);
}

// Host bytes -> device through the slot's pinned staging buffer: a copy from pageable memory may make the driver wait
// for the stream (CUDA API synchronisation rules), which would serialise the host walk of the next batch behind the
// running one.  The staging buffer is reused only after the slot has been collected.
void Engine::upload(void* dst, const void* host, size_t bytes) {
  if (bytes == 0) return;
  Slot& S = slots_[cur_];
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (S.h_stage_off + need > S.h_stage_cap) {
    // grow: copies queued from the old buffer must have been read first
    NWC_CUDA(cudaStreamSynchronize(stream_));
    if (S.h_stage) NWC_CUDA(cudaFreeHost(S.h_stage));
    S.h_stage = nullptr;
    const size_t cap = std::max((S.h_stage_off + need) * 2, (size_t)4 << 20);
    NWC_CUDA(cudaMallocHost((void**)&S.h_stage, cap));
    S.h_stage_cap = cap;
    S.h_stage_off = 0;
  }
  char* p = S.h_stage + S.h_stage_off;
  S.h_stage_off += need;
  memcpy(p, host, bytes);
  NWC_CUDA(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, stream_));
  stats.h2d_bytes += bytes;
}

// job lists live in the slot's arena: it is private to the batch, so nothing in flight can still read it
void* Engine::upload_jobs(const void* host, size_t bytes) {
  void* d = arena().alloc(bytes);
  upload(d, host, bytes);
  return d;
}

void Engine::flush_prep() {
  Slot& S = slots_[cur_];
  if (!cjobs_.empty()) {   // remote blocks first: antisym / repack read the local copies (same stream)
    const CopyJob* d = (const CopyJob*)upload_jobs(cjobs_.data(), cjobs_.size() * sizeof(CopyJob));
    if (timing) NWC_CUDA(cudaEventRecord(S.ev[0], stream_));
    launch_pull(d, (int)cjobs_.size(), max_copy_, stream_);
    NWC_CUDA(cudaGetLastError());
    if (timing) { NWC_CUDA(cudaEventRecord(S.ev[1], stream_)); S.timed[0] = true; }
    stats.pull_launches += (long long)((cjobs_.size() + 32767) / 32768);
    cjobs_.clear();
    max_copy_ = 0;
  }
  if (!ajobs_.empty()) {   // then the spin-orbital blocks: the repack jobs below read them
    const AntisymJob* d = (const AntisymJob*)upload_jobs(ajobs_.data(), ajobs_.size() * sizeof(AntisymJob));
    launch_antisym(d, (int)ajobs_.size(), max_ablock_, stream_);
    NWC_CUDA(cudaGetLastError());
    stats.antisym_launches += (long long)((ajobs_.size() + 32767) / 32768);
    ajobs_.clear();
    max_ablock_ = 0;
  }
  if (jobs_.empty()) return;
  const RepackJob* d = (const RepackJob*)upload_jobs(jobs_.data(), jobs_.size() * sizeof(RepackJob));
  if (timing) NWC_CUDA(cudaEventRecord(S.ev[2], stream_));
  launch_repack(d, (int)jobs_.size(), max_panel_, stream_);
  NWC_CUDA(cudaGetLastError());
  if (timing) { NWC_CUDA(cudaEventRecord(S.ev[3], stream_)); S.timed[1] = true; }
  stats.repack_launches += (long long)((jobs_.size() + 32767) / 32768);
  jobs_.clear();
  max_panel_ = 0;
}

int Engine::submit(double* dump_doubles, double* dump_singles) {
  if (open_) throw Error("nwc_triples: submit with an open tuple");
  if (trace_only()) throw Error("nwc_triples: a trace context cannot execute anything (there is no CPU fallback)");
  const int nt = (int)tuples_.size();
  if (nt == 0) return -1;
  Slot& S = slots_[cur_];
  if (S.busy) throw Error("nwc_triples: batch slot still in flight");
  flush_prep();
  // Dual-energy batch: every tuple gets a shadow header `items_` work items further on.  The fused kernel is launched over
  // the real items only (it never sees the shadows) and writes each sub-tile's second energy pair into the shadow's
  // partial slot; the reduction then treats 2*nt tuples alike.
  bool dual = false, single = false;
  for (const TupleHdr& t : tuples_) { dual = dual || t.two_sided < 0; single = single || t.two_sided >= 0; }
  if (dual && single) throw Error("nwc_triples: a batch cannot mix dual-energy tuples with others");
  if (dual && (dump_doubles || dump_singles)) throw Error("nwc_triples: the validation dump does not support dual-energy tuples");
  const int ntot = dual ? 2 * nt : nt;
  const long long slots = dual ? 2 * items_ : items_;
  if (dual) {
    tuples_.reserve((size_t)ntot);
    for (int i = 0; i < nt; i++) {
      TupleHdr sh = tuples_[(size_t)i];
      sh.item_begin += items_;
      tuples_.push_back(sh);
    }
  }
  // one metadata buffer per slot: tuples | descs | sdescs | energies | chunk sums | partials
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_t = 0, o_d = al(o_t + ntot * sizeof(TupleHdr)), o_s = al(o_d + descs_.size() * sizeof(ContrDesc)),
               o_e = al(o_s + sdescs_.size() * sizeof(SinglesDesc)), o_c = al(o_e + ntot * sizeof(double2)),
               o_p = al(o_c + (size_t)ntot * max_chunks_ * sizeof(double2)),
               total = al(o_p + (size_t)slots * partials_per_item() * sizeof(double2));
  if (total > S.d_meta_cap) {
    if (S.d_meta) NWC_CUDA(cudaFree(S.d_meta));
    S.d_meta = nullptr; S.d_meta_cap = 0;
    NWC_CUDA(cudaMalloc(&S.d_meta, total + total / 4));
    S.d_meta_cap = total + total / 4;
  }
  if ((size_t)ntot > S.h_out_cap) {
    if (S.h_out) NWC_CUDA(cudaFreeHost(S.h_out));
    S.h_out = nullptr; S.h_out_cap = 0;
    const size_t cap = std::max((size_t)ntot * 2, (size_t)256);
    NWC_CUDA(cudaMallocHost((void**)&S.h_out, cap * sizeof(double2)));
    S.h_out_cap = cap;
  }
  char* dm = (char*)S.d_meta;
  upload(dm + o_t, tuples_.data(), ntot * sizeof(TupleHdr));
  upload(dm + o_d, descs_.data(), descs_.size() * sizeof(ContrDesc));
  upload(dm + o_s, sdescs_.data(), sdescs_.size() * sizeof(SinglesDesc));
  S.timed[2] = timing;
  if (timing) NWC_CUDA(cudaEventRecord(S.ev[4], stream_));
  if (dump_doubles) {
    for (const TupleHdr& t : tuples_)
      if (t.two_sided) throw Error("nwc_triples: the validation dump does not support two-sided (Lambda) tuples");
    launch_fused_dump((const TupleHdr*)(dm + o_t), nt, (const ContrDesc*)(dm + o_d), (const SinglesDesc*)(dm + o_s),
                      (double2*)(dm + o_p), items_, dump_doubles, dump_singles, order_, stream_);
  } else {
    bool ragged = false, lambda = false, plain = false, crx = false;
    for (const TupleHdr& t : tuples_) {
      for (int q = 0; q < 6; q++) ragged = ragged || (t.R[q] % SB != 0);
      const bool l = t.two_sided != 0;
      if (!l && t.sdesc_mid > t.sdesc_begin) throw Error("nwc_triples: doubles-bound outer products need a two-sided tuple");
      lambda = lambda || l;
      plain = plain || !l;
      // side-0-bound terms, more terms than FusedSmem holds, or a dual-energy tuple: the CR-CCSD(T) instantiation
      crx = crx || (t.two_sided != 0 && t.two_sided != 1) || (t.sdesc_end - t.sdesc_begin > MAX_SINGLES_TERMS);
    }
    if (lambda && plain) throw Error("nwc_triples: a batch cannot mix (T) tuples and two-sided (Lambda) tuples");
    launch_fused((const TupleHdr*)(dm + o_t), nt, (const ContrDesc*)(dm + o_d), (const SinglesDesc*)(dm + o_s),
                 (double2*)(dm + o_p), items_, ragged, order_, lambda ? (crx ? 2 : 1) : 0, stream_);
  }
  NWC_CUDA(cudaGetLastError());
  if (timing) NWC_CUDA(cudaEventRecord(S.ev[5], stream_));
  launch_reduce((const TupleHdr*)(dm + o_t), ntot, (const double2*)(dm + o_p), (double2*)(dm + o_c), max_chunks_,
                (double2*)(dm + o_e), stream_);
  NWC_CUDA(cudaGetLastError());
  NWC_CUDA(cudaMemcpyAsync(S.h_out, dm + o_e, ntot * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
  NWC_CUDA(cudaEventRecord(S.done, stream_));
  S.ntuples = ntot;
  S.busy = true;
  stats.d2h_bytes += ntot * sizeof(double2);
  stats.fused_launches += 1;
  stats.reduce_launches += 2;
  stats.work_items += items_;
  stats.descs += (long long)descs_.size();
  stats.tuples += nt;
  tuples_.clear();
  descs_.clear();
  sdescs_.clear();
  items_ = 0;
  max_chunks_ = 1;
  const int submitted = cur_;
  cur_ ^= 1;
  return submitted;
}

void Engine::collect(int slot, double* energies_out, bool compact) {
  if (slot < 0) return;
  Slot& S = slots_[slot];
  if (!S.busy) return;
  NWC_CUDA(cudaEventSynchronize(S.done));
  double* acc[3] = {&stats.pull_ms, &stats.repack_ms, &stats.fused_ms};
  for (int k = 0; k < 3; k++)
    if (S.timed[k]) {
      float ms = 0;
      NWC_CUDA(cudaEventElapsedTime(&ms, S.ev[2 * k], S.ev[2 * k + 1]));
      *acc[k] += ms;
      S.timed[k] = false;
    }
  S.h_stage_off = 0;
  if (energies_out) memcpy(energies_out, S.h_out, (size_t)S.ntuples * sizeof(double2));
  S.busy = false;
  S.arena.reset(compact);
}

void Engine::run(double* energies_out, double* dump_doubles, double* dump_singles) {
  const int s = submit(dump_doubles, dump_singles);
  collect(s, energies_out);
  if (s >= 0 && !slots_[cur_].busy) cur_ = s;   // synchronous callers keep building in the same slot (one warm arena)
}

}  // namespace nwc
