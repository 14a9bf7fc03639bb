// Host driver above the kernel boundary: a C++ restatement of ccsd_t_gpu.F (task loop, :87-241),
// ccsd_t_singles_gpu.F and ccsd_t_doubles_gpu.F that calls the Tier-1 symbols exactly as the Fortran does:
// blocks are fetched from the (host) block stores, sorted on the host with TCE_SORT_2/4 semantics and passed
// as host arrays.  It stands in for the Fortran in this image (no Fortran compiler); INTEGRATION.md shows
// the ISO_C_BINDING module the real driver uses instead.
#include "host_driver.h"
#include "engine.h"
#include <cstring>
#include <dlfcn.h>
#include <algorithm>

using namespace nwc;
namespace nwc { Engine& compat_engine(); void compat_set_async_uploads(bool on); void compat_forget_uploads(); }

namespace {

// Pinned bump arena for the sorted T1/T2 operands of one tuple.  Every sorted block gets fresh pinned memory, so
// (i) its upload is asynchronous DMA overlapping the host sort of the next block and (ii) the promise behind
// nwc_compat_set_async_uploads(1) holds: an operand is never modified between the call that passes it and
// compute_en_, which lets the library recognise the same block passed again for each fired kernel.  When the arena
// wraps, in-flight copies are drained and the library is told to forget the host pointers it has seen.
struct PinnedScratch {
  double* base = nullptr;
  size_t cap = 0, off = 0;
  double* acquire(size_t n) {
    n = (n + 31) & ~(size_t)31;
    if (cap == 0) {
      const char* e = getenv("NWC_PINNED_MB");
      cap = (size_t)(e && *e ? atol(e) : 1024) * (1u << 20) / sizeof(double);
    }
    if (n > cap) { if (base) { NWC_CUDA(cudaStreamSynchronize(compat_engine().stream())); cudaFreeHost(base); base = nullptr; } cap = n; }
    if (!base) { NWC_CUDA(cudaMallocHost((void**)&base, cap * sizeof(double))); off = 0; }
    if (off + n > cap) {   // wrap: nothing may still be reading, and cached identities are void
      NWC_CUDA(cudaStreamSynchronize(compat_engine().stream()));
      compat_forget_uploads();
      off = 0;
    }
    double* p = base + off;
    off += n;
    return p;
  }
  void tuple_done() { off = 0; }   // compute_en_ has synchronised the stream
};
PinnedScratch g_scratch;

// Reference contract (nwc_driver_set_reference_contract(1)): behave exactly like the unmodified Fortran call sites --
// the sorted operand lives in ONE pageable scratch buffer that is refilled for every (row, h7b/p7b) pair (the MA
// push/pop of k_a_sort, ccsd_t_doubles_gpu.F:262-304,723-726), nothing is pinned, and the library is given no promise.
bool g_reference_contract = false;
struct PageableScratch {
  double* base = nullptr;
  size_t cap = 0;
  double* acquire(size_t n) {
    if (n > cap) { free(base); base = (double*)malloc(n * sizeof(double)); cap = n; }
    return base;
  }
};
PageableScratch g_ma_scratch;
double* scratch(size_t n) { return g_reference_contract ? g_ma_scratch.acquire(n) : g_scratch.acquire(n); }

// sorted(i,j,k,l order, l fastest) = factor * unsorted(a,b,c,d order, d fastest): tce_sort_4 semantics
// (src/tce/sort/new_sort4.F).  Cache-blocked: the input's fastest index (d) and the output's fastest index are moved in
// 16 x 16 tiles, so both the reads and the writes touch whole cache lines; the naive loop writes one double per cache
// line (a 40^4 block: 2.56e6 strided stores) and was the largest host cost of the Tier-1 path.
void sort4(const double* in, double* out, Integer a, Integer b, Integer c, Integer d, int i, int j, int k, int l,
           double factor) {
  const Integer jd[4] = {a, b, c, d};
  const int perm[4] = {i - 1, j - 1, k - 1, l - 1};
  Integer ostride[4], istride[4];  // strides in `out` / `in` of input index q
  Integer s = 1;
  for (int q = 3; q >= 0; q--) { ostride[perm[q]] = s; s *= jd[perm[q]]; }
  istride[3] = 1; istride[2] = d; istride[1] = c * d; istride[0] = b * c * d;
  const int f = perm[3];            // input index that is fastest in the output
  if (f == 3) {                     // same fastest index: contiguous runs of length d
#pragma omp parallel for collapse(3) schedule(static)
    for (Integer i0 = 0; i0 < a; i0++)
      for (Integer i1 = 0; i1 < b; i1++)
        for (Integer i2 = 0; i2 < c; i2++) {
          const double* src = in + i0 * istride[0] + i1 * istride[1] + i2 * istride[2];
          double* dst = out + i0 * ostride[0] + i1 * ostride[1] + i2 * ostride[2];
          for (Integer x = 0; x < d; x++) dst[x] = factor * src[x];
        }
    return;
  }
  int r[2], nr = 0;                 // the two indices that are fastest in neither array
  for (int q = 0; q < 3; q++) if (q != f) r[nr++] = q;
  const Integer T = 16, nx = (d + T - 1) / T, ny = (jd[f] + T - 1) / T;
  const Integer os3 = ostride[3], isf = istride[f];
#pragma omp parallel for collapse(3) schedule(static)
  for (Integer u0 = 0; u0 < jd[r[0]]; u0++)
    for (Integer u1 = 0; u1 < jd[r[1]]; u1++)
      for (Integer ty = 0; ty < ny; ty++) {
        const double* src0 = in + u0 * istride[r[0]] + u1 * istride[r[1]];
        double* dst0 = out + u0 * ostride[r[0]] + u1 * ostride[r[1]];
        const Integer y0 = ty * T, y1 = y0 + T < jd[f] ? y0 + T : jd[f];
        for (Integer tx = 0; tx < nx; tx++) {
          const Integer x0 = tx * T, x1 = x0 + T < d ? x0 + T : d;
          for (Integer x = x0; x < x1; x++) {
            const double* src = src0 + x;
            double* dst = dst0 + x * os3;
            for (Integer y = y0; y < y1; y++) dst[y] = factor * src[y * isf];
          }
        }
      }
}

// The Tier-1 call surface as a table.  By default it is this library's own symbols; nwc_driver_bind_backend() points it
// at any other shared library that exports the reference's names -- e.g. the reference's own sd_t_total.cu + memory.cu
// compiled unmodified (oracle/_ref) -- so the same host driver can be run against the reference kernels (a test of
// the driver's calling sequence against reference code).
typedef void (*s1_fn)(Integer*, Integer*, Integer*, Integer*, Integer*, Integer*, double*, double*, double*);
typedef void (*dx_fn)(Integer*, Integer*, Integer*, Integer*, Integer*, Integer*, Integer*, double*, double*, double*);
typedef void (*mem_fn)(Integer*, Integer*, Integer*, Integer*, Integer*, Integer*);
typedef void (*en_fn)(double*, double*, double*, double*, double*, double*, double*, double*, Integer*, Integer*, Integer*,
                      Integer*, Integer*, Integer*, double*, double*);
struct Tier1Api {
  s1_fn s1[9] = {sd_t_s1_1_cuda_, sd_t_s1_2_cuda_, sd_t_s1_3_cuda_, sd_t_s1_4_cuda_, sd_t_s1_5_cuda_,
                 sd_t_s1_6_cuda_, sd_t_s1_7_cuda_, sd_t_s1_8_cuda_, sd_t_s1_9_cuda_};
  dx_fn d1[9] = {sd_t_d1_1_cuda_, sd_t_d1_2_cuda_, sd_t_d1_3_cuda_, sd_t_d1_4_cuda_, sd_t_d1_5_cuda_,
                 sd_t_d1_6_cuda_, sd_t_d1_7_cuda_, sd_t_d1_8_cuda_, sd_t_d1_9_cuda_};
  dx_fn d2[9] = {sd_t_d2_1_cuda_, sd_t_d2_2_cuda_, sd_t_d2_3_cuda_, sd_t_d2_4_cuda_, sd_t_d2_5_cuda_,
                 sd_t_d2_6_cuda_, sd_t_d2_7_cuda_, sd_t_d2_8_cuda_, sd_t_d2_9_cuda_};
  mem_fn mem_s = dev_mem_s_, mem_d = dev_mem_d_;
  en_fn compute_en = compute_en_;
  void (*init)(void) = initmemmodule_;
  void (*fini)(void) = finalizememmodule_;
  void (*release)(void) = dev_release_;
  bool foreign = false;
  void* handle = nullptr;
};
Tier1Api g_api;

struct CompatSink {
  const HostState& S;
  const nwc_tce_state* st;
  double* a_sort = nullptr;

  void singles(const Row& r, Integer p4b_1, Integer h1b_1, Integer p5b_2, Integer p6b_2, Integer h2b_2, Integer h3b_2,
               const bool fire[9]) {
    const Integer rp4 = S.rg(r.p4b), rh1 = S.rg(r.h1b);
    const double* blk = st->t1 + hash_lookup_or_die(S.t1_hash, t1_key(S, p4b_1, h1b_1), "t1");
    a_sort = scratch((size_t)(rp4 * rh1));
    for (Integer p = 0; p < rp4; p++)      // TCE_SORT_2(...,2,1): stored (p4,h1) h1 fastest -> t1sub(p4,h1)
      for (Integer h = 0; h < rh1; h++) a_sort[p + rp4 * h] = blk[h + rh1 * p];
    const double* v = st->v2 + hash_lookup_or_die(S.v2_hash, v2_key(S, p5b_2, p6b_2, h2b_2, h3b_2), "v2(pphh)");
    Integer h1d = rh1, h2d = S.rg(r.h2b), h3d = S.rg(r.h3b), p4d = rp4, p5d = S.rg(r.p5b), p6d = S.rg(r.p6b);
    for (int k = 0; k < 9; k++)
      if (fire[k]) g_api.s1[k](&h1d, &h2d, &h3d, &p4d, &p5d, &p6d, nullptr, a_sort, const_cast<double*>(v));
  }

  void d1_pair(const Row& r, Integer h7b, const Integer am[4], const Integer bm[4], const bool fire[9]) {
    const Integer rp4 = S.rg(r.p4b), rp5 = S.rg(r.p5b), rh1 = S.rg(r.h1b), rh7 = S.rg(h7b);
    a_sort = scratch((size_t)(rp4 * rp5 * rh1 * rh7));
    if (h7b < r.h1b) {  // ccsd_t_doubles_gpu.F:282-289
      const double* blk = st->t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[3], am[2]), "t2");
      sort4(blk, a_sort, rp4, rp5, rh7, rh1, 4, 2, 1, 3, -1.0);
    } else {            // :291-298
      const double* blk = st->t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[2], am[3]), "t2");
      sort4(blk, a_sort, rp4, rp5, rh1, rh7, 3, 2, 1, 4, 1.0);
    }
    const double* v = st->v2 + hash_lookup_or_die(S.v2_hash, v2_key(S, bm[1], bm[0], bm[2], bm[3]), "v2(hphh)");  // :315-327
    Integer h1d = rh1, h2d = S.rg(r.h2b), h3d = S.rg(r.h3b), h7d = rh7, p4d = rp4, p5d = rp5, p6d = S.rg(r.p6b);
    for (int k = 0; k < 9; k++)
      if (fire[k]) g_api.d1[k](&h1d, &h2d, &h3d, &h7d, &p4d, &p5d, &p6d, nullptr, a_sort, const_cast<double*>(v));
  }

  void d2_pair(const Row& r, Integer p7b, const Integer am[4], const Integer bm[4], const bool fire[9]) {
    const Integer rp4 = S.rg(r.p4b), rp7 = S.rg(p7b), rh1 = S.rg(r.h1b), rh2 = S.rg(r.h2b);
    a_sort = scratch((size_t)(rp4 * rp7 * rh1 * rh2));
    if (p7b < r.p4b) {  // ccsd_t_doubles_gpu.F:942-949
      const double* blk = st->t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[1], am[0], am[2], am[3]), "t2");
      sort4(blk, a_sort, rp7, rp4, rh1, rh2, 4, 3, 2, 1, -1.0);
    } else {            // :950-957
      const double* blk = st->t2 + hash_lookup_or_die(S.t2_hash, t2_key(S, am[0], am[1], am[2], am[3]), "t2");
      sort4(blk, a_sort, rp4, rp7, rh1, rh2, 4, 3, 1, 2, 1.0);
    }
    const double* v = st->v2 + hash_lookup_or_die(S.v2_hash, v2_key(S, bm[0], bm[1], bm[2], bm[3]), "v2(pphp)");  // :964-976
    Integer h1d = rh1, h2d = rh2, h3d = S.rg(r.h3b), p4d = rp4, p5d = S.rg(r.p5b), p6d = S.rg(r.p6b), p7d = rp7;
    for (int k = 0; k < 9; k++)
      if (fire[k]) g_api.d2[k](&h1d, &h2d, &h3d, &p4d, &p5d, &p6d, &p7d, nullptr, a_sort, const_cast<double*>(v));
  }
  void row_end(int) {}   // the Fortran calls the kernels tile by tile; the library concatenates them itself
};

// one task: ccsd_t_gpu.F:114-230
void one_tuple(const HostState& S, const nwc_tce_state* st, const Integer t[6], double e[2], double* dump_d,
               double* dump_s) {
  Integer rp4 = S.rg(t[0]), rp5 = S.rg(t[1]), rp6 = S.rg(t[2]), rh1 = S.rg(t[3]), rh2 = S.rg(t[4]), rh3 = S.rg(t[5]);
  g_api.init();                                       // :135
  CompatSink sink{S, st};
  // opt-in fast path: this driver owns every pinned operand it passes and never touches one in flight
  compat_set_async_uploads(!g_reference_contract && !g_api.foreign);
  g_api.mem_s(&rh1, &rh2, &rh3, &rp4, &rp5, &rp6);    // ccsd_t_singles_gpu.F:192-197
  walk_singles(S, t, sink);                           // :139
  g_api.mem_d(&rh1, &rh2, &rh3, &rp4, &rp5, &rp6);    // ccsd_t_doubles_gpu.F:227-232
  walk_doubles(S, t, sink);                           // :144
  double factor = tuple_factor(S, t);                 // :153-167
  double* ev = const_cast<double*>(st->evl_sorted);
  if (dump_d && !g_api.foreign)
    nwc_compute_en_dump_(&factor, e, ev + S.offset[t[3] - 1], ev + S.offset[t[4] - 1], ev + S.offset[t[5] - 1],
                         ev + S.offset[t[0] - 1], ev + S.offset[t[1] - 1], ev + S.offset[t[2] - 1], &rh1, &rh2, &rh3,
                         &rp4, &rp5, &rp6, dump_d, dump_s);
  else
    g_api.compute_en(&factor, e, ev + S.offset[t[3] - 1], ev + S.offset[t[4] - 1], ev + S.offset[t[5] - 1],
                ev + S.offset[t[0] - 1], ev + S.offset[t[1] - 1], ev + S.offset[t[2] - 1], &rh1, &rh2, &rh3, &rp4,
                &rp5, &rp6, nullptr, nullptr);            // :205-215
  compat_set_async_uploads(false);
  g_scratch.tuple_done();
  g_api.release();                                    // :219
  g_api.fini();                                       // :220
}

}  // namespace

extern "C" {

void nwc_driver_set_reference_contract(int on) { g_reference_contract = on != 0; }

// Route the host driver's Tier-1 calls into another shared library exporting the reference's symbols (NULL or "" =
// back to this library).  Returns 0, or 1 if the library or one of its 33 symbols is missing.
int nwc_driver_bind_backend(const char* so_path) {
  if (g_api.handle) { dlclose(g_api.handle); }
  g_api = Tier1Api();
  if (!so_path || !*so_path) return 0;
  void* h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { printf("nwc_driver_bind_backend: %s\n", dlerror()); return 1; }
  Tier1Api a;
  bool ok = true;
  auto sym = [&](const std::string& n) { void* p = dlsym(h, n.c_str()); ok = ok && p != nullptr; return p; };
  for (int k = 0; k < 9; k++) {
    a.s1[k] = (s1_fn)sym("sd_t_s1_" + std::to_string(k + 1) + "_cuda_");
    a.d1[k] = (dx_fn)sym("sd_t_d1_" + std::to_string(k + 1) + "_cuda_");
    a.d2[k] = (dx_fn)sym("sd_t_d2_" + std::to_string(k + 1) + "_cuda_");
  }
  a.mem_s = (mem_fn)sym("dev_mem_s_"); a.mem_d = (mem_fn)sym("dev_mem_d_");
  a.compute_en = (en_fn)sym("compute_en_");
  a.init = (void (*)(void))sym("initmemmodule_"); a.fini = (void (*)(void))sym("finalizememmodule_");
  a.release = (void (*)(void))sym("dev_release_");
  if (!ok) { dlclose(h); printf("nwc_driver_bind_backend: %s lacks part of the call surface\n", so_path); return 1; }
  a.foreign = true; a.handle = h;
  g_api = a;
  return 0;
}

static int ccsd_t_gpu_impl(const nwc_tce_state* st, Integer icuda, Integer my_rank, Integer nranks, double energy[2],
                           double* per_task);
int nwc_ccsd_t_gpu(const nwc_tce_state* st, Integer icuda, Integer my_rank, Integer nranks, double energy[2],
                   double* per_task) {
  try {
    return ccsd_t_gpu_impl(st, icuda, my_rank, nranks, energy, per_task);
  } catch (const std::exception& ex) {   // reference behaviour above the kernel boundary: errquit
    printf("%s\n", ex.what());
    fflush(stdout);
    exit(1);
  }
}

// a given list of tasks (ntasks x 6 tile ids: p4b,p5b,p6b,h1b,h2b,h3b) through the Tier-1 surface, e.g. a prefix of
// the heaviest-first list; energy[2] = their sums
int nwc_ccsd_t_gpu_tasks(const nwc_tce_state* st, Integer icuda, const Integer* tasks6, Integer ntasks, double energy[2],
                         double* per_task) {
  try {
    HostState S;
    S.load_tables(st);
    Integer ic = icuda, devno = 0;
    if (check_device_(&ic) != 1) { printf("nwc_ccsd_t_gpu: this rank owns no GPU and there is no CPU path\n"); return 2; }
    device_init_(&ic, &devno);
    if (devno == 30) return 30;
    energy[0] = energy[1] = 0.0;
    for (Integer k = 0; k < ntasks; k++) {
      double e[2] = {0, 0};
      one_tuple(S, st, tasks6 + 6 * k, e, nullptr, nullptr);
      energy[0] += e[0];
      energy[1] += e[1];
      if (per_task) { per_task[2 * k] = e[0]; per_task[2 * k + 1] = e[1]; }
    }
    return 0;
  } catch (const std::exception& ex) {
    printf("%s\n", ex.what());
    fflush(stdout);
    exit(1);
  }
}

static int ccsd_t_gpu_impl(const nwc_tce_state* st, Integer icuda, Integer my_rank, Integer nranks, double energy[2],
                           double* per_task) {
  HostState S;
  S.load_tables(st);
  Integer ic = icuda, devno = 0;
  if (check_device_(&ic) != 1) { printf("nwc_ccsd_t_gpu: this rank owns no GPU and there is no CPU path\n"); return 2; }
  device_init_(&ic, &devno);                          // ccsd_t_gpu.F:55-57
  if (devno == 30) return 30;                         // :58-60
  // same loop nest and filters as ccsd_t_gpu.F:87-112; the nxtask counter is replaced by a static deal
  energy[0] = energy[1] = 0.0;
  Integer count = 0;
  const Integer n0 = S.noab, n1 = S.noab + S.nvab;
  for (Integer p4 = n0 + 1; p4 <= n1; p4++)
    for (Integer p5 = p4; p5 <= n1; p5++)
      for (Integer p6 = p5; p6 <= n1; p6++)
        for (Integer h1 = 1; h1 <= n0; h1++)
          for (Integer h2 = h1; h2 <= n0; h2++)
            for (Integer h3 = h2; h3 <= n0; h3++) {
              const Integer ps = S.sp(p4) + S.sp(p5) + S.sp(p6), hs = S.sp(h1) + S.sp(h2) + S.sp(h3);
              if (ps != hs) continue;
              if (S.restricted && ps + hs > 8) continue;
              if ((S.sy(p4) ^ S.sy(p5) ^ S.sy(p6) ^ S.sy(h1) ^ S.sy(h2) ^ S.sy(h3)) != 0) continue;
              if (count % nranks == my_rank) {
                const Integer t[6] = {p4, p5, p6, h1, h2, h3};
                double e[2] = {0, 0};
                one_tuple(S, st, t, e, nullptr, nullptr);
                energy[0] += e[0];                    // :216-217
                energy[1] += e[1];
                if (per_task) { per_task[2 * count] = e[0]; per_task[2 * count + 1] = e[1]; }
              }
              count++;
            }
  return 0;
}

// ---- host-only helpers (no device needed): the task list and a dry run of the dispatch logic ----
Integer nwc_host_task_list(const nwc_tce_state* st, Integer* klist7, Integer capacity_tasks) {
  HostState S;
  S.load_tables(st);
  std::vector<Integer> kl;
  build_task_list(S, kl);
  const Integer n = (Integer)(kl.size() / 7);
  if (klist7 && n <= capacity_tasks) memcpy(klist7, kl.data(), kl.size() * sizeof(Integer));
  return n;
}

namespace {
struct CountSink {
  const HostState& S;
  Integer calls[3] = {0, 0, 0};
  double flops[3] = {0, 0, 0};
  double prod(const Row& r) const {
    return (double)S.rg(r.p4b) * S.rg(r.p5b) * S.rg(r.p6b) * S.rg(r.h1b) * S.rg(r.h2b) * S.rg(r.h3b);
  }
  void singles(const Row& r, Integer, Integer, Integer, Integer, Integer, Integer, const bool fire[9]) {
    for (int k = 0; k < 9; k++) if (fire[k]) { calls[0]++; flops[0] += 2.0 * prod(r); }
  }
  void d1_pair(const Row& r, Integer h7b, const Integer*, const Integer*, const bool fire[9]) {
    for (int k = 0; k < 9; k++) if (fire[k]) { calls[1]++; flops[1] += 2.0 * prod(r) * S.rg(h7b); }
  }
  void d2_pair(const Row& r, Integer p7b, const Integer*, const Integer*, const bool fire[9]) {
    for (int k = 0; k < 9; k++) if (fire[k]) { calls[2]++; flops[2] += 2.0 * prod(r) * S.rg(p7b); }
  }
  void row_end(int) {}
};
}  // namespace

namespace {
struct CollectSink {   // block keys one tuple touches, per store
  const HostState& S;
  std::vector<Integer>* keys[3];   // T1, T2, V2
  void singles(const Row&, Integer p4b_1, Integer h1b_1, Integer p5b_2, Integer p6b_2, Integer h2b_2, Integer h3b_2,
               const bool[9]) {
    keys[0]->push_back(t1_key(S, p4b_1, h1b_1));
    keys[2]->push_back(v2_key(S, p5b_2, p6b_2, h2b_2, h3b_2));
  }
  void d1_pair(const Row& r, Integer h7b, const Integer am[4], const Integer bm[4], const bool[9]) {
    keys[1]->push_back(h7b < r.h1b ? t2_key(S, am[0], am[1], am[3], am[2]) : t2_key(S, am[0], am[1], am[2], am[3]));
    keys[2]->push_back(v2_key(S, bm[1], bm[0], bm[2], bm[3]));
  }
  void d2_pair(const Row& r, Integer p7b, const Integer am[4], const Integer bm[4], const bool[9]) {
    keys[1]->push_back(p7b < r.p4b ? t2_key(S, am[1], am[0], am[2], am[3]) : t2_key(S, am[0], am[1], am[2], am[3]));
    keys[2]->push_back(v2_key(S, bm[0], bm[1], bm[2], bm[3]));
  }
  void row_end(int) {}
};
}  // namespace

// Sorted unique block keys of store `which` (1 T1, 2 T2, 3 spin-orbital V2) that the given tasks read: lets a caller
// stage only those blocks on the host (st needs the tiling tables and the T1/T2 offset tables; no data, no device).
// Returns the number of keys; keys_out (capacity `cap`) is filled when it is large enough.
Integer nwc_host_collect_blocks(const nwc_tce_state* st, const Integer* tasks6, Integer ntasks, int which,
                                Integer* keys_out, Integer cap) {
  try {
    HostState S;
    S.load_tables(st);
    std::vector<Integer> k[3];
    CollectSink sink{S, {&k[0], &k[1], &k[2]}};
    for (Integer i = 0; i < ntasks; i++) {
      walk_singles(S, tasks6 + 6 * i, sink);
      walk_doubles(S, tasks6 + 6 * i, sink);
    }
    if (which < 1 || which > 3) return -1;
    std::vector<Integer>& v = k[which - 1];
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    if (keys_out && (Integer)v.size() <= cap) memcpy(keys_out, v.data(), v.size() * sizeof(Integer));
    return (Integer)v.size();
  } catch (const std::exception& ex) {
    printf("%s\n", ex.what());
    return -1;
  }
}

// host-only view of nwc_triples_run_partition: ranges[2*i..] = sub-tile range of task first_task+i that `rank` runs
int nwc_host_block_partition(const nwc_tce_state* st, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                             long long* ranges) {
  try {
    HostState S;
    S.load_tables(st);
    std::vector<Integer> kl;
    build_task_list(S, kl);
    const Integer nt = (Integer)(kl.size() / 7);
    if (nranks < 1 || rank < 0 || rank >= nranks || first_task < 0 || first_task > nt) return 1;
    if (ntasks <= 0 || first_task + ntasks > nt) ntasks = nt - first_task;
    std::vector<long long> r;
    std::vector<Integer> ids;
    for (Integer i = 0; i < ntasks; i++) ids.push_back(first_task + i);
    block_partition(S, kl, rank, nranks, ids, r);
    memcpy(ranges, r.data(), r.size() * sizeof(long long));
    return 0;
  } catch (const std::exception& ex) {
    printf("%s\n", ex.what());
    return 1;
  }
}

// host-only: the driver's TCE_SORT_4 (tests compare it with the oracle's restatement for all 24 permutations)
void nwc_host_sort4(const double* unsorted, double* sorted, Integer a, Integer b, Integer c, Integer d, int i, int j, int k,
                    int l, double factor) {
  sort4(unsorted, sorted, a, b, c, d, i, j, k, l, factor);
}

// calls[3], flops[3] = fired sd_t_s1 / d1 / d2 kernels of one tuple and their algorithmic FLOPs
int nwc_host_count_tuple(const nwc_tce_state* st, const Integer tuple[6], Integer calls[3], double flops[3]) {
  HostState S;
  S.load_tables(st);
  CountSink sink{S};
  walk_singles(S, tuple, sink);
  walk_doubles(S, tuple, sink);
  for (int i = 0; i < 3; i++) { calls[i] = sink.calls[i]; flops[i] = sink.flops[i]; }
  return 0;
}

// `2eorb`: where the two Mulliken halves of <g3 g4||g1 g2> sit in the caller's d_v2orb (host-only; no device needed).
// off_host[2] = element offsets of the direct / exchange source blocks (-1: that half does not fire), strides[8] =
// element strides of (g3,g4,g1,g2) in the two sources.  Returns non-zero if the table does not match the tiling.
int nwc_host_2eorb_plan(const nwc_tce_state* st, const nwc_tce_orb_state* orb, const Integer g3g4g1g2[4],
                        Integer off_host[2], Integer strides[8]) {
  HostState S;
  nwc_tce_state s2 = *st;
  s2.v2_hash = nullptr;
  S.load_tables(&s2);
  if (!S.load_orbital(orb->noa, orb->nva, orb->b2am, orb->spin_alpha, orb->sym_alpha, orb->range_alpha, orb->v2orb_hash).empty())
    return 1;
  const HostState::OrbPlan p = S.block_plan(g3g4g1g2[0], g3g4g1g2[1], g3g4g1g2[2], g3g4g1g2[3]);
  off_host[0] = p.key_a >= 0 ? S.host_off.at(p.key_a) : -1;
  off_host[1] = p.key_b >= 0 ? S.host_off.at(p.key_b) : -1;
  for (int q = 0; q < 4; q++) { strides[q] = p.sa[q]; strides[4 + q] = p.sb[q]; }
  return 0;
}

int nwc_ccsd_t_gpu_tuple(const nwc_tce_state* st, const Integer tuple[6], double energy[2], double* host_doubles,
                         double* host_singles) {
  try {
    HostState S;
    S.load_tables(st);
    one_tuple(S, st, tuple, energy, host_doubles, host_singles);
    return 0;
  } catch (const std::exception& ex) {
    printf("%s\n", ex.what());
    fflush(stdout);
    exit(1);
  }
}

}  // extern "C"
