// Host-side restatement of the per-tuple driver logic above the kernel boundary:
//   ccsd_t_singles_gpu.F:36-574, ccsd_t_doubles_gpu.F:48-742 (_1, Sum h7) and :743-1345 (_2, Sum p7),
//   tce_hashnsort.F, tce_restricted.F, tce_hash.F:271-322, ccsd_t_neword.F.
// It walks the permutation table of one task tuple, applies the reference's filters and dispatch tests and
// reports every (operand pair, fired kernels) event to a Sink.  Two sinks exist:
//   * driver_replica.cu : fetch + TCE_SORT on the host, then call the Tier-1 symbols (as the Fortran does)
//   * native_abi.cu     : emit device-side repack jobs / descriptors against the HBM-resident block stores
#pragma once
#include <vector>
#include <unordered_map>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include "../../include/nwc_triples.h"

namespace nwc {

struct HostState {
  Integer noab = 0, nvab = 0, restricted = 1, irrep_t = 0, irrep_v = 0;
  std::vector<Integer> spin, sym, range, offset, alpha;
  std::vector<double> evl;
  std::vector<Integer> t1_hash, t2_hash, v2_hash;
  void load_tables(const nwc_tce_state* s) {
    noab = s->noab; nvab = s->nvab; restricted = s->restricted; irrep_t = s->irrep_t; irrep_v = s->irrep_v;
    const Integer n = noab + nvab;
    spin.assign(s->spin, s->spin + n); sym.assign(s->sym, s->sym + n); range.assign(s->range, s->range + n);
    offset.assign(s->offset, s->offset + n); alpha.assign(s->alpha, s->alpha + n);
    Integer ne = 0;
    for (Integer i = 0; i < n; i++) ne = offset[i] + range[i] > ne ? offset[i] + range[i] : ne;
    evl.assign(s->evl_sorted, s->evl_sorted + ne);
    t1_hash.assign(s->t1_hash, s->t1_hash + 2 * s->t1_hash[0] + 1);
    t2_hash.assign(s->t2_hash, s->t2_hash + 2 * s->t2_hash[0] + 1);
    if (s->v2_hash) v2_hash.assign(s->v2_hash, s->v2_hash + 2 * s->v2_hash[0] + 1); else v2_hash.assign(1, 0);
    intorb = false;
  }
  // ---- `2eorb` storage (tce.fh intorb): V2 spin-free over the alpha tiles -------------------------------------
  bool intorb = false;
  Integer noa = 0, nva = 0;
  std::vector<Integer> b2am, spin_alpha, sym_alpha, range_alpha;
  std::unordered_map<Integer, Integer> orb_off;   // orbital block key -> offset in the RESIDENT (compacted) store
  std::unordered_map<Integer, Integer> host_off;  // scratch: key -> offset in the caller's full store
  struct OrbRun { Integer src, dst, n; };         // contiguous run of needed blocks: host offset -> resident offset
  std::vector<OrbRun> orb_runs;
  struct OrbBlock { Integer key, host_off, size; };   // the needed blocks in storage order (the unit of sharding)
  std::vector<OrbBlock> orb_blocks;
  std::unordered_map<Integer, Integer> orb_index; // orbital block key -> index into orb_blocks
  Integer orb_size = 0;                           // doubles resident on the device
  Integer orb_host_size = 0;                      // doubles in the caller's d_v2orb
  static Integer index_pair(Integer i, Integer j) { return (i * (i - 1)) / 2 + j; }   // tce_mo2e_offset_intorb.F:615
  // Expands the checkpointed table k_v2_alpha_offset into a full key -> offset map by running the block loops of
  // tce_mo2e_offset_intorb.F:32-50 once (tce_hash_v2 re-walks them from a checkpoint at every lookup) and checks
  // every checkpoint of the caller's table against it.  Returns "" or an error text.
  std::string load_orbital(Integer noa_, Integer nva_, const Integer* b2am_, const Integer* spin_a, const Integer* sym_a,
                           const Integer* range_a, const Integer* table) {
    noa = noa_; nva = nva_;
    const Integer n = noa + nva;
    b2am.assign(b2am_, b2am_ + noab + nvab);
    spin_alpha.assign(spin_a, spin_a + n); sym_alpha.assign(sym_a, sym_a + n); range_alpha.assign(range_a, range_a + n);
    orb_off.clear();
    host_off.clear();
    orb_runs.clear();
    orb_blocks.clear();
    orb_index.clear();
    // (T) reads <pp||hh>, <hp||hh> and <pp||hp> only, i.e. the Mulliken blocks (vo|vo), (oo|vo) and (vo|vv): a stored
    // block can be touched iff at least one of its two tile pairs is mixed (one hole tile, one particle tile).
    // Everything else -- (oo|oo), (oo|vv), (vv|vv), the bulk of the store -- stays on the host.
    auto is_p = [&](Integer a) { return a > noa; };
    Integer size = 0, resident = 0;
    for (Integer g3b = 1; g3b <= n; g3b++)
      for (Integer g4b = g3b; g4b <= n; g4b++)
        for (Integer g1b = 1; g1b <= n; g1b++)
          for (Integer g2b = g1b; g2b <= n; g2b++) {
            if (spin_alpha[g3b - 1] + spin_alpha[g4b - 1] != spin_alpha[g1b - 1] + spin_alpha[g2b - 1]) continue;
            if ((sym_alpha[g3b - 1] ^ sym_alpha[g4b - 1] ^ sym_alpha[g1b - 1] ^ sym_alpha[g2b - 1]) != irrep_v) continue;
            if (index_pair(g4b, g3b) < index_pair(g2b, g1b)) continue;
            const Integer key = g2b - 1 + n * (g1b - 1 + n * (g4b - 1 + n * (g3b - 1)));
            const Integer bs = range_alpha[g3b - 1] * range_alpha[g4b - 1] * range_alpha[g1b - 1] * range_alpha[g2b - 1];
            const int pr = (int)is_p(g3b) + (int)is_p(g4b), pc = (int)is_p(g1b) + (int)is_p(g2b);
            const bool needed = pr == 1 || pc == 1;
            host_off[key] = size;
            if (needed) {
              orb_off[key] = resident;
              orb_index[key] = (Integer)orb_blocks.size();
              orb_blocks.push_back({key, size, bs});
              if (!orb_runs.empty() && orb_runs.back().src + orb_runs.back().n == size) orb_runs.back().n += bs;
              else orb_runs.push_back({size, resident, bs});
              resident += bs;
            }
            size += bs;
          }
    orb_size = resident;
    orb_host_size = size;
    const Integer length1 = table[0];
    for (Integer pos = 1; pos <= length1 + 1; pos++) {
      const Integer key = table[pos], off = table[(length1 + 1) + pos];
      auto it = host_off.find(key);
      if (it == host_off.end() || it->second != off) return "k_v2_alpha_offset does not match the alpha tiling (checkpoint " + std::to_string(pos) + ")";
    }
    intorb = true;
    return "";
  }
  // One Mulliken integral class (a b|c d), a in alpha tile tile[0], ...: which stored block holds it and with which
  // element strides (get_block_ind.F:884-1016 pair ordering; tce_mo2e_trans.F:707-723 element order (k l|i j), k fastest).
  struct OrbSource { Integer key; long long stride[4]; };
  OrbSource mulliken_source(const Integer tile[4]) const {
    const Integer n = noa + nva;
    int ia = 0, ib = 1, ic = 2, id = 3;                       // argument slots: (a b | c d)
    if (tile[ia] < tile[ib]) { int t = ia; ia = ib; ib = t; } // larger tile first inside each pair
    if (tile[ic] < tile[id]) { int t = ic; ic = id; id = t; }
    if (index_pair(tile[ia], tile[ib]) < index_pair(tile[ic], tile[id])) { int t = ia; ia = ic; ic = t; t = ib; ib = id; id = t; }
    // now row pair = (ia, ib) -> (k, l), column pair = (ic, id) -> (i, j)
    OrbSource m;
    m.key = tile[ic] - 1 + n * (tile[id] - 1 + n * (tile[ia] - 1 + n * (tile[ib] - 1)));
    const long long rk = range_alpha[tile[ia] - 1], rl = range_alpha[tile[ib] - 1], ri = range_alpha[tile[ic] - 1];
    m.stride[ia] = 1; m.stride[ib] = rk; m.stride[ic] = rk * rl; m.stride[id] = rk * rl * ri;
    return m;
  }
  // <g3 g4||g1 g2> = (g3 g1|g4 g2) - (g3 g2|g4 g1) (get_block_ind.F:818-1538): the (up to) two sources with the
  // strides of the target indices x0=g3, x1=g4, x2=g1, x3=g2; key < 0 = that half does not fire for these spins
  struct OrbPlan { Integer key_a, key_b; long long sa[4], sb[4]; };
  OrbPlan block_plan(Integer g3b, Integer g4b, Integer g1b, Integer g2b) const {
    OrbPlan p{};
    p.key_a = p.key_b = -1;
    const Integer s3 = sp(g3b), s4 = sp(g4b), s1 = sp(g1b), s2 = sp(g2b);
    const Integer a3 = b2am[g3b - 1], a4 = b2am[g4b - 1], a1 = b2am[g1b - 1], a2 = b2am[g2b - 1];
    if (s3 == s1 && s4 == s2) {   // direct: uaadaa, ubbdbb, uabdab, ubadba (get_block_ind.F:983)
      const Integer tile[4] = {a3, a1, a4, a2};
      const OrbSource m = mulliken_source(tile);
      p.key_a = m.key;
      p.sa[0] = m.stride[0]; p.sa[2] = m.stride[1]; p.sa[1] = m.stride[2]; p.sa[3] = m.stride[3];
    }
    if (s3 == s2 && s4 == s1) {   // exchange: uaadaa, ubbdbb, uabdba, ubadab (:1264)
      const Integer tile[4] = {a3, a2, a4, a1};
      const OrbSource m = mulliken_source(tile);
      p.key_b = m.key;
      p.sb[0] = m.stride[0]; p.sb[3] = m.stride[1]; p.sb[1] = m.stride[2]; p.sb[2] = m.stride[3];
    }
    return p;
  }
  Integer sp(Integer b) const { return spin[b - 1]; }
  Integer sy(Integer b) const { return sym[b - 1]; }
  Integer rg(Integer b) const { return range[b - 1]; }
  Integer N() const { return noab + nvab; }
};

// offset of block `key` in a TCE offset table (sorted keys): tce_hash.F:271-322.  -1 if absent.
inline Integer hash_lookup(const Integer* hash, Integer key) {
  Integer n = hash[0], lo = 1, hi = n;
  while (lo <= hi) {
    Integer mid = (lo + hi) >> 1;
    if (hash[mid] == key) return hash[n + mid];
    if (hash[mid] < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
inline Integer hash_lookup_or_die(const std::vector<Integer>& hash, Integer key, const char* what) {
  Integer off = hash_lookup(hash.data(), key);
  if (off < 0) throw std::runtime_error(std::string("nwc_triples: ") + what + ": block key " + std::to_string(key) + " not found");
  return off;
}

// block keys
inline Integer t1_key(const HostState& S, Integer p, Integer h) { return h - 1 + S.noab * (p - S.noab - 1); }
inline Integer t2_key(const HostState& S, Integer p1, Integer p2, Integer h3, Integer h4) {
  return h4 - 1 + S.noab * (h3 - 1 + S.noab * (p2 - S.noab - 1 + S.nvab * (p1 - S.noab - 1)));
}
inline Integer v2_key(const HostState& S, Integer g3, Integer g4, Integer g1, Integer g2) {
  const Integer N = S.N();
  return g2 - 1 + N * (g1 - 1 + N * (g4 - 1 + N * (g3 - 1)));
}

// tce_restricted_2 / _4 (tce_restricted.F:1-71)
inline void restricted_map(const HostState& S, int n, const Integer* in, Integer* out) {
  Integer ssum = 0;
  for (int i = 0; i < n; i++) ssum += S.sp(in[i]);
  const bool map = S.restricted && ssum == 2 * n;
  for (int i = 0; i < n; i++) out[i] = map ? S.alpha[in[i] - 1] : in[i];
}

struct Row { Integer p4b, p5b, p6b, h1b, h2b, h3b; };

// the nine-row table: P/H choose which task tile plays which permuted role; duplicates are dropped
inline int build_rows(const Integer t[6], const int P[3][3], const int H[3][3], Row rows[9]) {
  int n = 0;
  for (int ip = 0; ip < 3; ip++)
    for (int ih = 0; ih < 3; ih++) {
      Row r{t[P[ip][0]], t[P[ip][1]], t[P[ip][2]], t[3 + H[ih][0]], t[3 + H[ih][1]], t[3 + H[ih][2]]};
      bool dup = false;
      for (int j = 0; j < n; j++)
        dup = dup || (rows[j].p4b == r.p4b && rows[j].p5b == r.p5b && rows[j].p6b == r.p6b && rows[j].h1b == r.h1b &&
                      rows[j].h2b == r.h2b && rows[j].h3b == r.h3b);
      if (!dup) rows[n++] = r;
    }
  return n;
}

inline bool row_ok_target(const HostState& S, const Row& r, Integer target) {
  const Integer ps = S.sp(r.p4b) + S.sp(r.p5b) + S.sp(r.p6b), hs = S.sp(r.h1b) + S.sp(r.h2b) + S.sp(r.h3b);
  if (S.restricted && ps + hs == 12) return false;
  if (ps != hs) return false;
  return (S.sy(r.p4b) ^ S.sy(r.p5b) ^ S.sy(r.p6b) ^ S.sy(r.h1b) ^ S.sy(r.h2b) ^ S.sy(r.h3b)) == target;
}
inline bool row_ok(const HostState& S, const Row& r) { return row_ok_target(S, r, S.irrep_v ^ S.irrep_t); }

// which of the nine kernels fire for this row: kernel K=3*kp+kh fires iff the task tuple equals the row
// permuted by TP[kp] (particles) and TH[kh] (holes)
inline int fired(const Integer t[6], const Row& r, const int TP[3][3], const int TH[3][3], bool fire[9]) {
  const Integer rp[3] = {r.p4b, r.p5b, r.p6b}, rh[3] = {r.h1b, r.h2b, r.h3b};
  int n = 0;
  for (int kp = 0; kp < 3; kp++)
    for (int kh = 0; kh < 3; kh++) {
      const bool f = t[0] == rp[TP[kp][0]] && t[1] == rp[TP[kp][1]] && t[2] == rp[TP[kp][2]] &&
                     t[3] == rh[TH[kh][0]] && t[4] == rh[TH[kh][1]] && t[5] == rh[TH[kh][2]];
      fire[3 * kp + kh] = f;
      n += f;
    }
  return n;
}

// t = (t_p4b,t_p5b,t_p6b,t_h1b,t_h2b,t_h3b)
// row_target < 0: the (T) filter (irrep_v xor irrep_t).  cr_ccsd_t_E_2 (cr_ccsd_t_E.F:408-742) is this same walk -- same
// rows :472-533, row filter :560, t1 filter :579-580, restricted maps :582-583, dispatch tests :625-721 -- with the row
// target irrep_t^irrep_t^irrep_t (:571) and the <pp||hh> block replaced by the i1(pphh)_tt intermediate.
template <class Sink>
void walk_singles(const HostState& S, const Integer t[6], Sink& sink, Integer row_target = -1) {
  static const int P[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}};   // ccsd_t_singles_gpu.F:101-162
  static const int H[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}};
  static const int TP[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}};  // tests :281,:370,:462
  static const int TH[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}};  // tests :281,:310,:340
  Row rows[9];
  const int n = build_rows(t, P, H, rows);
  for (int i = 0; i < n; i++) {
    const Row& r = rows[i];
    if (!(r.p5b <= r.p6b && r.h2b <= r.h3b)) continue;                       // :200
    if (!(row_target < 0 ? row_ok(S, r) : row_ok_target(S, r, row_target))) continue;   // :203-211
    if (S.sp(r.p4b) != S.sp(r.h1b)) continue;                                // :218
    if ((S.sy(r.p4b) ^ S.sy(r.h1b)) != S.irrep_t) continue;                  // :219
    bool fire[9];
    if (!fired(t, r, TP, TH, fire)) continue;
    const Integer a[2] = {r.p4b, r.h1b}, b[4] = {r.p5b, r.p6b, r.h2b, r.h3b};
    Integer am[2], bm[4];
    restricted_map(S, 2, a, am);                                             // :221
    restricted_map(S, 4, b, bm);                                             // :222
    sink.singles(r, am[0], am[1], bm[0], bm[1], bm[2], bm[3], fire);
  }
}

// cr_ccsd_t_E_1 (cr_ccsd_t_E.F:74-407): E += P(9) t(p4 p5 h1 h2) t(p6 h3) -- nine outer products of a T2 block with a T1
// block.  Reports (row, T2 block ids after tce_restricted_4, T1 block ids after tce_restricted_2, fired sd_E_K).
template <class Sink>
void walk_cr_e1(const HostState& S, const Integer t[6], Sink& sink) {
  static const int P[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}};   // rows :138-199: (p4,p5,p6),(p5,p6,p4),(p4,p6,p5)
  static const int H[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}};   //                (h1,h2,h3),(h2,h3,h1),(h1,h3,h2)
  static const int TP[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}};  // tests :290,:323,:356
  static const int TH[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}};  // tests :290,:301,:312
  Row rows[9];
  const int n = build_rows(t, P, H, rows);
  for (int i = 0; i < n; i++) {
    const Row& r = rows[i];
    if (!(r.p4b <= r.p5b && r.h1b <= r.h2b)) continue;                                  // :226
    if (!row_ok_target(S, r, S.irrep_t ^ S.irrep_t)) continue;                          // :229-237
    if (S.sp(r.p4b) + S.sp(r.p5b) != S.sp(r.h1b) + S.sp(r.h2b)) continue;               // :244
    if ((S.sy(r.p4b) ^ S.sy(r.p5b) ^ S.sy(r.h1b) ^ S.sy(r.h2b)) != S.irrep_t) continue; // :246
    bool fire[9];
    if (!fired(t, r, TP, TH, fire)) continue;
    const Integer a[4] = {r.p4b, r.p5b, r.h1b, r.h2b}, b[2] = {r.p6b, r.h3b};
    Integer am[4], bm[2];
    restricted_map(S, 4, a, am);                                                        // :248
    restricted_map(S, 2, b, bm);                                                        // :249
    sink.cr_e1(r, am, bm, fire);
  }
}

template <class Sink>
void walk_doubles(const HostState& S, const Integer t[6], Sink& sink) {
  {  // ---- Sum(h7): ccsd_t_doubles_gpu.F:120-742 ----
    static const int P[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}};
    static const int H[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}};
    static const int TP[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}};  // tests :357,:474,:597
    static const int TH[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}};  // tests :357,:394,:433
    Row rows[9];
    const int n = build_rows(t, P, H, rows);
    for (int i = 0; i < n; i++) {
      const Row& r = rows[i];
      if (!(r.p4b <= r.p5b && r.h2b <= r.h3b)) continue;            // :236
      if (!row_ok(S, r)) continue;                                  // :239-247
      bool fire[9];
      if (!fired(t, r, TP, TH, fire)) continue;
      for (Integer h7b = 1; h7b <= S.noab; h7b++) {                 // :255
        if (S.sp(r.p4b) + S.sp(r.p5b) != S.sp(r.h1b) + S.sp(h7b)) continue;
        if ((S.sy(r.p4b) ^ S.sy(r.p5b) ^ S.sy(r.h1b) ^ S.sy(h7b)) != S.irrep_t) continue;
        const Integer a[4] = {r.p4b, r.p5b, r.h1b, h7b}, b[4] = {r.p6b, h7b, r.h2b, r.h3b};
        Integer am[4], bm[4];
        restricted_map(S, 4, a, am);                                // :260
        restricted_map(S, 4, b, bm);                                // :261
        sink.d1_pair(r, h7b, am, bm, fire);
      }
      sink.row_end(1);   // all h7 tiles of this row have been reported
    }
  }
  {  // ---- Sum(p7): ccsd_t_doubles_gpu.F:810-1345 ----
    static const int P[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}};
    static const int H[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}};
    static const int TP[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}};  // tests :998,:1106,:1214
    static const int TH[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}};  // tests :998,:1034,:1070
    Row rows[9];
    const int n = build_rows(t, P, H, rows);
    for (int i = 0; i < n; i++) {
      const Row& r = rows[i];
      if (!(r.p5b <= r.p6b && r.h1b <= r.h2b)) continue;            // :909
      if (!row_ok(S, r)) continue;
      bool fire[9];
      if (!fired(t, r, TP, TH, fire)) continue;
      for (Integer p7b = S.noab + 1; p7b <= S.noab + S.nvab; p7b++) {  // :923
        if (S.sp(r.p4b) + S.sp(p7b) != S.sp(r.h1b) + S.sp(r.h2b)) continue;
        if ((S.sy(r.p4b) ^ S.sy(p7b) ^ S.sy(r.h1b) ^ S.sy(r.h2b)) != S.irrep_t) continue;
        const Integer a[4] = {r.p4b, p7b, r.h1b, r.h2b}, b[4] = {r.p5b, r.p6b, r.h3b, p7b};
        Integer am[4], bm[4];
        restricted_map(S, 4, a, am);                                // :928
        restricted_map(S, 4, b, bm);                                // :929
        sink.d2_pair(r, p7b, am, bm, fire);
      }
      sink.row_end(2);   // all p7 tiles of this row have been reported
    }
  }
}

// ccsd_t_dot.F:52-66
inline double tuple_factor(const HostState& S, const Integer t[6]) {
  double f = S.restricted ? 2.0 : 1.0;
  if (t[0] == t[1] && t[1] == t[2]) f /= 6.0; else if (t[0] == t[1] || t[1] == t[2]) f /= 2.0;
  if (t[3] == t[4] && t[4] == t[5]) f /= 6.0; else if (t[3] == t[4] || t[4] == t[5]) f /= 2.0;
  return f;
}

// dry walk: k4 planes of all fired contractions of one tuple -- the cost model of the static block partition
struct CostSink {
  const HostState& S;
  long long planes = 0;
  long long rowK = 0; int rowfired = 0;   // the contracted tiles of one row are concatenated along K (engine.h Segment)
  void singles(const Row&, Integer, Integer, Integer, Integer, Integer, Integer, const bool fire[9]) {
    for (int k = 0; k < 9; k++) if (fire[k]) planes += 1;
  }
  void pair(Integer K, const bool fire[9]) {
    rowK += K;
    rowfired = 0;
    for (int k = 0; k < 9; k++) rowfired += fire[k] ? 1 : 0;
  }
  void d1_pair(const Row&, Integer h7b, const Integer*, const Integer*, const bool fire[9]) { pair(S.rg(h7b), fire); }
  void d2_pair(const Row&, Integer p7b, const Integer*, const Integer*, const bool fire[9]) { pair(S.rg(p7b), fire); }
  void row_end(int) { planes += rowfired * ((rowK + 3) / 4); rowK = 0; rowfired = 0; }
  void cr_e1(const Row&, const Integer*, const Integer*, const bool fire[9]) {
    for (int k = 0; k < 9; k++) if (fire[k]) planes += 1;
  }
};

inline long long tuple_sub_tiles(const HostState& S, const Integer t[6]) {
  long long n = 1;
  for (int q = 0; q < 6; q++) n *= (S.rg(t[q]) + 3) / 4;
  return n;
}

// Static block partition of the tasks `ids` (indices into `klist`, rows of 7) over nranks: the tasks are laid end to
// end in the order given, every 4^6 sub-tile weighted by the k4 planes its tuple contracts plus a constant for the per-sub-tile epilogue,
// and rank r takes the r-th equal-cost contiguous piece.  ranges[2*i], ranges[2*i+1] = the sub-tile range
// [item_lo, item_hi) of task ids[i] that `rank` runs (empty when lo == hi).  Pure integer arithmetic: every rank
// derives the same cuts.
inline void block_partition(const HostState& S, const std::vector<Integer>& klist, Integer rank, Integer nranks,
                            const std::vector<Integer>& ids, std::vector<long long>& ranges) {
  const Integer ntasks = (Integer)ids.size();
  const long long EPILOGUE_PLANES = 24;   // transfers, singles, energy of one sub-tile in units of one k4 plane
  std::vector<long long> items((size_t)ntasks), w((size_t)ntasks);
  std::vector<__int128> cum((size_t)ntasks + 1, 0);
  for (Integer i = 0; i < ntasks; i++) {
    const Integer* t = &klist[7 * (size_t)ids[(size_t)i]];
    CostSink cs{S};
    walk_singles(S, t, cs);
    walk_doubles(S, t, cs);
    items[(size_t)i] = tuple_sub_tiles(S, t);
    w[(size_t)i] = cs.planes + EPILOGUE_PLANES;
    cum[(size_t)i + 1] = cum[(size_t)i] + (__int128)items[(size_t)i] * w[(size_t)i];
  }
  const __int128 total = cum[(size_t)ntasks];
  const __int128 lo = total * rank / nranks, hi = total * (rank + 1) / nranks;
  auto cut = [&](__int128 bound, Integer i) -> long long {   // first sub-tile of task i at or beyond `bound`
    const __int128 rel = bound - cum[(size_t)i];
    if (rel <= 0) return 0;
    const __int128 q = (rel + w[(size_t)i] - 1) / w[(size_t)i];
    return q > items[(size_t)i] ? items[(size_t)i] : (long long)q;
  };
  ranges.assign(2 * (size_t)ntasks, 0);
  for (Integer i = 0; i < ntasks; i++) {
    ranges[2 * (size_t)i] = cut(lo, i);
    ranges[2 * (size_t)i + 1] = cut(hi, i);
  }
}

// task enumeration + heaviest-first banding: ccsd_t_neword.F:42-217.  Rows of 7: 6 tile ids + weight.
inline void build_task_list(const HostState& S, std::vector<Integer>& klist) {
  std::vector<Integer> aux;
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
              const Integer w = S.rg(p4) * S.rg(p5) * S.rg(p6) * S.rg(h1) * S.rg(h2) * S.rg(h3);
              const Integer row[7] = {p4, p5, p6, h1, h2, h3, w};
              aux.insert(aux.end(), row, row + 7);
            }
  const size_t nt = aux.size() / 7;
  klist.clear();
  if (nt == 0) return;
  Integer wmax = 0, wmin;
  for (size_t i = 0; i < nt; i++) wmax = aux[7 * i + 6] > wmax ? aux[7 * i + 6] : wmax;
  wmin = wmax;
  for (size_t i = 0; i < nt; i++) wmin = aux[7 * i + 6] < wmin ? aux[7 * i + 6] : wmin;
  if (((wmax - wmin) * 100.0) / wmax < 1.0) { klist = aux; return; }   // :136-143
  klist.reserve(aux.size());
  auto sweep = [&](Integer value) {
    for (size_t i = 0; i < nt; i++)
      if (aux[7 * i + 6] > value) {
        klist.insert(klist.end(), aux.begin() + 7 * i, aux.begin() + 7 * i + 7);
        aux[7 * i + 6] = -99;
      }
  };
  for (Integer ii = 16; ii >= 1; ii--) sweep(wmin + ((wmax - wmin) * (ii - 1)) / 16);  // :168-173
  sweep(0);                                                                              // :174
}

}  // namespace nwc
