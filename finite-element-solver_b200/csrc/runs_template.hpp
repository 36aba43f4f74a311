// Stencil runs: host-side detection and template programs for k_assemble_runs (assemble_runs.cuh).
// Pure C++ (no CUDA headers) so the same code is compiled into libfes_b200.so and into the CPU
// emulator test (tests/host/runs_emulator.cpp), which replays a template program on the host and
// compares it with a plain per-pair assembly of the same mesh.
//
// A RUN is a maximal range of consecutive nodes whose RELATIVE stencil repeats with period P (1 or 2):
// node v0 + P t + phi (lane-step t, phase phi) has the incident elements E0 + t * delta + de_r with the
// same local indices and the same corner nodes relative to v0 + P t for every t.  Structured meshes are
// made of runs (a mesh row of the reference's create_rect_mesh is one run of period 2 -- the diagonals
// alternate --, a z-line of create_box_mesh one run of period 1; ref fem/mesh/structured.py:17-96); the
// detection reads the connectivity only, so a structured mesh with arbitrary coordinates qualifies.
//
// What the consecutive runs leave over is searched for STRIDED runs v0 + S t (segment_strided: the end nodes of the
// mesh rows / z-lines), and a last handful of nodes become single-node tiles, so a structured mesh is covered
// completely and one kernel launch assembles it.
//
// A run is cut into TILES of <= 32 lane-steps.  The kernel evaluates the geometry of every element LINE
// {E0 + d0 + tau * delta} touching the tile once per tau into shared-memory planes indexed by tau, then
// lane t of a warp sums the pairs of its P nodes from addresses (code + t): unit stride across the warp,
// no bank conflicts, and nothing but a 32-byte tile header, the node coordinates and the per-element
// coefficients is read from HBM.  The warps of a CTA (4 or 6) split the PAIRS of the period between them.
//
// Template blob (int32 words; offsets relative to the template's first word):
//   [0] P  [1] n_lines  [2] TS (plane stride, entries)  [3] image padding: 0 = none, else 1 + log2(rows per pad unit)
//   [4] LU (image row, units)  [5] n_spans  [6] off_lines  [7] off_spans  [8] n_warps (CTA = 32 n_warps threads)
//   [9] step (node stride per lane-step: P for consecutive nodes, S for a strided run)
//   [10] threads / LU  [11] threads mod LU (the store loop's row / unit increments per step of the CTA)
//   [12] n_words  [13] coordinate slots staged per tile  [14] delta  [15] Tmax (lane-steps per tile, 32 - the longest line overhang)
//   [16] doubles per image row (c^2 * pairs per period)  [17] record entries (32 zero entries + n_lines * N * TS)
//   [20..27] off_prog[warp]
//   lines : n_lines x {eoff, ext, cidx[N]}   element of entry tau' = E0 + eoff + tau' * delta, entries
//           tau' < T + ext; corner c sits in coordinate slot cidx[c] + tau'
//   spans : n_spans x {start, n_ext, base, 0}  nodes v0 + start + step * i, i <= (T - 1) + n_ext, staged in slot base + i
//   prog  : per warp {n_pairs, then per pair: head, codes[2^lg * L]}
//           head = img_b (15) | row_b (12) << 15 | lg (2) << 27 | diag << 29: byte offset 8 (c^2 prefix_phi + c k) of the
//                  block in the lane's image row, bytes 8 c deg between the block's rows
//           code = 16 entry_a | 16 entry_b << 16 (byte offsets in the double2 planes; the z plane sits at half),
//                  entry = 32 + (line * N + local) * TS + (shift - taumin_line); lane t adds t;
//           partial sum j of a pair = its L codes [j L, j L + L), short ones padded with code 0 (the zero entries)
// An image UNIT is 16 bytes when c is even, 8 bytes otherwise; row r of the image starts at unit r LU + r / g (g rows
// per pad unit), which keeps the lanes' stores on distinct banks.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

namespace fes_runs_host {

constexpr int kLaneSteps = 32;  // lane-steps per tile: one warp wide
constexpr int kMaxWarps = 8;    // warps per CTA = pair groups of a template: 4, 6 or 8 (the plan's choice)
constexpr int kZeroEntries = 32;  // record entries [0, 32) hold zeros: the padding of short partial sums (lane t reads entry t)
constexpr int kMinPeriods = 8;  // shorter repeats stay with the tiled kernel
constexpr int kHdrWords = 28;
constexpr int kTileInts = 8;    // {v0, T, template word offset, E0, first pair, raw word offset, pair stride per lane-step, 0}

// 64-bit mix shared by the device signature kernel and the host emulator
#if defined(__CUDACC__)
#define FES_RUNS_HD __host__ __device__
#else
#define FES_RUNS_HD
#endif
FES_RUNS_HD inline uint64_t sig_mix(uint64_t h, uint64_t x) {
  h = (h ^ x) * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}
// signature of a node's relative stencil: incident elements (ascending id) as (e - e_first, local index, corner - v)
FES_RUNS_HD inline uint64_t node_signature(int32_t v, int32_t ninc, const uint32_t* ne_code, const int32_t* en, int N) {
  uint64_t h = sig_mix(0x243F6A8885A308D3ull, (uint64_t)ninc);
  if (ninc == 0) return h;
  const int64_t e0 = (int64_t)(ne_code[0] >> 3);
  for (int32_t r = 0; r < ninc; ++r) {
    const int64_t e = (int64_t)(ne_code[r] >> 3);
    h = sig_mix(h, (uint64_t)(e - e0) * 8u + (ne_code[r] & 7u));
    for (int c = 0; c < N; ++c) h = sig_mix(h, (uint64_t)((int64_t)en[e * N + c] - v));
  }
  return h;
}

struct Run {
  int32_t v0;
  int32_t P;        // nodes per lane-step (consecutive): 1 or 2
  int32_t periods;  // lane-steps
  int64_t delta;    // element stride per lane-step
  int32_t step;     // node stride per lane-step: P for a run of consecutive nodes, S > 1 for a STRIDED run (P = 1)
  int32_t pstride;  // stored pairs between consecutive lane-steps: node_ptr[v + step] - node_ptr[v]
};

// Greedy segmentation of the node range into runs; everything else is left to the tiled kernel.
inline std::vector<Run> segment_runs(const uint64_t* sig, const int32_t* eref, const int32_t* ninc, int64_t n,
                                     int min_periods = kMinPeriods) {
  std::vector<Run> runs;
  int64_t v = 0;
  while (v < n) {
    bool found = false;
    for (int P = 1; P <= 2 && !found; ++P) {
      if (v + 2 * (int64_t)P > n) break;
      bool ok = true;
      for (int phi = 0; phi < P; ++phi) ok = ok && ninc[v + phi] > 0;
      if (!ok) continue;
      const int64_t delta = (int64_t)eref[v + P] - eref[v];
      if (delta <= 0) continue;
      int64_t w = v + P;
      while (w < n && sig[w] == sig[w - P] && (int64_t)eref[w] - eref[w - P] == delta) ++w;
      const int64_t periods = (w - v) / P;
      if (periods >= min_periods) {
        runs.push_back(Run{(int32_t)v, P, (int32_t)periods, delta, P, 0});  // pstride: filled in by the caller from node_ptr
        v += periods * P;
        found = true;
      }
    }
    if (!found) ++v;
  }
  return runs;
}

// STRIDED runs among the nodes the consecutive runs left over (`free_node[v]` != 0): v0 + S t with the same relative
// stencil, a constant element stride and a constant pair stride.  The two end nodes of every row of a structured
// mesh are such runs (S = the row pitch, or twice that when the diagonals alternate), as are the first / last nodes
// of the z-lines of a box mesh: with them in the run kernel a structured mesh needs no remainder tiles at all but
// for a handful of corner nodes.  Marks the nodes it takes in free_node (set to 0).
inline std::vector<Run> segment_strided(const uint64_t* sig, const int32_t* eref, const int32_t* ninc, const int32_t* node_ptr,
                                        int64_t n, std::vector<char>& free_node, int min_periods = kMinPeriods) {
  std::vector<Run> runs;
  std::vector<int64_t> rest;
  for (int64_t v = 0; v < n; ++v)
    if (free_node[v] && ninc[v] > 0) rest.push_back(v);
  for (size_t a = 0; a < rest.size(); ++a) {
    const int64_t v = rest[a];
    if (!free_node[v]) continue;
    // candidate strides: the distances to the next few left-over nodes with the same signature
    int tried = 0;
    for (size_t b = a + 1; b < rest.size() && b < a + 64 && tried < 3; ++b) {
      const int64_t w1 = rest[b];
      if (!free_node[w1] || sig[w1] != sig[v]) continue;
      ++tried;
      const int64_t S = w1 - v, delta = (int64_t)eref[w1] - eref[v], ps = (int64_t)node_ptr[w1] - node_ptr[v];
      if (S < 2 || S > 0x3fffffff || delta <= 0 || ps <= 0) continue;
      int64_t w = w1, cnt = 1;
      while (w < n && free_node[w] && sig[w] == sig[v] && (int64_t)eref[w] - eref[w - S] == delta &&
             (int64_t)node_ptr[w] - node_ptr[w - S] == ps) {
        ++cnt;
        w += S;
      }
      if (cnt >= min_periods) {
        runs.push_back(Run{(int32_t)v, 1, (int32_t)cnt, delta, (int32_t)S, (int32_t)ps});
        for (int64_t k = 0; k < cnt; ++k) free_node[v + k * S] = 0;
        break;
      }
    }
  }
  return runs;
}

// incident elements of node v0 + phi of a period, ascending element id
struct PhaseStencil {
  std::vector<int32_t> de;      // element id - E0 (E0 = first incident element of node v0)
  std::vector<int32_t> a;       // local index of the node in the element
  std::vector<int32_t> corner;  // N per element: corner node - v0
};

struct Template {
  int P = 0, N = 0, c = 0;
  int64_t delta = 0;
  int n_lines = 0, TS = 0, LU = 0, pad_rows = 0, img_units = 0, coord_nodes = 0, ppp = 0, n_spans = 0, Tmax = 0, rec_entries = 0;
  std::vector<int32_t> words;  // the kernel's program (layout above)
  std::vector<int32_t> raw;    // the stencil itself, for the verification kernel:
                               //   {P, delta, ppp, 0, then per phase: ninc, deg, prefix, ninc x (de, a, corner[N])}
};

inline int64_t floor_div(int64_t a, int64_t b) {  // b > 0
  int64_t q = a / b;
  if ((a % b) < 0) --q;
  return q;
}

inline void item_split_host(uint32_t cnt, uint32_t L, uint32_t& lg, uint32_t& m) {  // = item_split of plan.cu
  const uint32_t need = (cnt + L - 1) / L;
  lg = 0;
  while ((1u << lg) < need && lg < 5) ++lg;
  m = (cnt + (1u << lg) - 1) >> lg;
}

// Build the program of one template.  Returns false when the stencil cannot be encoded (the nodes then stay
// with the tiled kernel).  L = contributions per partial sum (tile_item_len(N)): the summation tree of a pair
// is the tiled kernel's, so both kernels produce the same bits for the same pair.
inline bool build_template(int N, int c, int P, int64_t delta, const std::vector<PhaseStencil>& ph, int L, int n_warps,
                           Template& out, int step = 0 /* node stride per lane-step; 0 = P (consecutive nodes) */) {
  if (step <= 0) step = P;
  if (step != P && P != 1) return false;
  out = Template();
  out.P = P; out.N = N; out.c = c; out.delta = delta;
  if (delta <= 0 || delta > 0x7fffffff || (int)ph.size() != P || n_warps < 1 || n_warps > kMaxWarps) return false;
  const int kThreads = 32 * n_warps;
  struct Line {
    int64_t d0;
    int64_t tmin, tmax;
    std::vector<int64_t> ocan;  // corner offsets relative to v0 + P tau
  };
  std::vector<Line> lines;
  // a line = one residue of the element id modulo delta AND one set of canonical corner offsets: references with
  // equal residues but other corners (tets: the same Kuhn tet of a neighbouring cell column, (nz - 1) steps along
  // the same arithmetic progression) are separate lines.  Merging references is only a deduplication: each
  // reference's corners are verified per lane, so every line is right for the lanes that read it.
  std::map<std::pair<int64_t, std::vector<int64_t>>, int> line_of;
  // per reference (phase, r): line and shift
  std::vector<std::vector<int>> ref_line(P);
  std::vector<std::vector<int64_t>> ref_shift(P);
  for (int phi = 0; phi < P; ++phi) {
    const PhaseStencil& s = ph[phi];
    const int ninc = (int)s.de.size();
    if (ninc == 0 || (int)s.a.size() != ninc || (int)s.corner.size() != ninc * N) return false;
    ref_line[phi].resize(ninc);
    ref_shift[phi].resize(ninc);
    for (int r = 0; r < ninc; ++r) {
      const int64_t q = floor_div(s.de[r], delta), d0 = s.de[r] - q * delta;
      std::vector<int64_t> oc(N);
      for (int k = 0; k < N; ++k) oc[k] = (int64_t)s.corner[r * N + k] - (int64_t)step * q;
      if (s.a[r] < 0 || s.a[r] >= N || s.corner[r * N + s.a[r]] != phi) return false;  // the node itself
      const auto lkey = std::make_pair(d0, oc);
      auto it = line_of.find(lkey);
      int li;
      if (it == line_of.end()) {
        li = (int)lines.size();
        line_of[lkey] = li;
        lines.push_back(Line{d0, q, q, oc});
      } else {
        li = it->second;
        lines[li].tmin = std::min(lines[li].tmin, q);
        lines[li].tmax = std::max(lines[li].tmax, q);
      }
      ref_line[phi][r] = li;
      ref_shift[phi][r] = q;
    }
  }
  // lines in ascending d0 (deterministic layout)
  {
    std::vector<int> order(lines.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
      const int64_t ex = lines[x].d0 + lines[x].tmin * delta, ey = lines[y].d0 + lines[y].tmin * delta;
      return ex != ey ? ex < ey : lines[x].ocan < lines[y].ocan;
    });
    std::vector<int> rank(lines.size());
    std::vector<Line> sorted;
    for (size_t i = 0; i < order.size(); ++i) { rank[order[i]] = (int)i; sorted.push_back(lines[order[i]]); }
    lines.swap(sorted);
    for (int phi = 0; phi < P; ++phi)
      for (auto& li : ref_line[phi]) li = rank[li];
  }
  const int n_lines = (int)lines.size();
  int64_t max_ext = 0;
  for (auto& ln : lines) max_ext = std::max(max_ext, ln.tmax - ln.tmin);
  if (max_ext > 8) return false;
  // lane-steps per tile: 32 - max_ext, so every element line has at most 32 entries and phase 1 of the kernel is
  // one warp-pass per line (lane = entry) with warp-uniform line descriptors
  const int Tmax = kLaneSteps - (int)max_ext;
  const int TS = kLaneSteps;
  if (Tmax < 2 * kMinPeriods) return false;
  const int64_t rec_entries = kZeroEntries + (int64_t)n_lines * N * TS;
  if (rec_entries > 4095) return false;  // 16-bit BYTE offsets of the double2 planes in the codes

  // Coordinate spans.  Corner c of line l sits at node v0 + o + step * tau for the entries tau in [tmin, T - 1 + tmax]:
  // a STRIDED node sequence.  Sequences of one residue class (mod step) that overlap are merged into one span
  // {start, n_ext}: nodes v0 + start + step * i, i <= (T - 1) + n_ext, staged in slots base + i -- so the corner
  // of entry tau' is slot(c) + tau', unit stride across the threads of phase 1 whatever the period or the stride.
  struct Iv { int64_t lo, hi; };
  std::vector<Iv> ivs;
  for (auto& ln : lines)
    for (int k = 0; k < N; ++k) ivs.push_back(Iv{ln.ocan[k] + step * ln.tmin, ln.ocan[k] + step * ln.tmax});
  auto residue = [&](int64_t x) { return ((x % step) + step) % step; };
  std::sort(ivs.begin(), ivs.end(), [&](const Iv& x, const Iv& y) {
    return residue(x.lo) != residue(y.lo) ? residue(x.lo) < residue(y.lo) : x.lo < y.lo;
  });
  struct Span { int64_t start, n_ext, base; };
  std::vector<Span> spans;
  for (auto& iv : ivs) {
    if (!spans.empty() && residue(iv.lo) == residue(spans.back().start) &&
        iv.lo <= spans.back().start + (int64_t)step * (spans.back().n_ext + kMinPeriods)) {
      spans.back().n_ext = std::max(spans.back().n_ext, (iv.hi - spans.back().start) / step);
    } else {
      spans.push_back(Span{iv.lo, (iv.hi - iv.lo) / step, 0});
    }
  }
  int64_t coord_slots = 0;
  for (auto& sp : spans) {
    if (sp.start < -0x3fffffff || sp.start > 0x3fffffff || sp.n_ext > 64) return false;
    sp.base = coord_slots;
    coord_slots += (int64_t)(Tmax - 1) + sp.n_ext + 1;
    coord_slots = (coord_slots + 1) & ~(int64_t)1;
  }
  if (coord_slots > 8192 || spans.size() > 64) return false;
  auto cidx_of = [&](int64_t node_rel) -> int64_t {
    for (auto& sp : spans)
      if (residue(node_rel) == residue(sp.start) && node_rel >= sp.start && node_rel <= sp.start + step * sp.n_ext)
        return sp.base + (node_rel - sp.start) / step;
    return -1;
  };

  // pairs of each phase: distinct corner offsets in ascending order; contributions in ascending element id
  struct Pair {
    int phi, k, deg, prefix;
    bool diag;
    std::vector<uint32_t> codes;
  };
  std::vector<Pair> pairs;
  std::vector<int> deg_phi(P), prefix_phi(P);
  int ppp = 0;
  const uint32_t lg_cap = (N == 3) ? 3u : 2u;  // partial sums per pair the kernel's summation trees are compiled for: 2^lg_cap
  for (int phi = 0; phi < P; ++phi) {
    const PhaseStencil& s = ph[phi];
    std::vector<int32_t> nb(s.corner);
    std::sort(nb.begin(), nb.end());
    nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
    deg_phi[phi] = (int)nb.size();
    prefix_phi[phi] = ppp;
    ppp += deg_phi[phi];
    for (int k = 0; k < (int)nb.size(); ++k) {
      Pair p{phi, k, deg_phi[phi], prefix_phi[phi], true, {}};
      for (int r = 0; r < (int)s.de.size(); ++r)
        for (int b = 0; b < N; ++b)
          if (s.corner[r * N + b] == nb[k]) {
            const Line& ln = lines[ref_line[phi][r]];
            const int64_t sh = ref_shift[phi][r] - ln.tmin;
            const uint32_t ea = (uint32_t)(kZeroEntries + (ref_line[phi][r] * N + s.a[r]) * TS + sh);
            const uint32_t eb = (uint32_t)(kZeroEntries + (ref_line[phi][r] * N + b) * TS + sh);
            p.diag = p.diag && ea == eb;
            p.codes.push_back((16u * ea) | ((16u * eb) << 16));
          }
      uint32_t lg, m;
      item_split_host((uint32_t)p.codes.size(), (uint32_t)L, lg, m);
      if (p.codes.empty() || m > (uint32_t)L || lg > lg_cap) return false;
      pairs.push_back(p);
    }
    if (deg_phi[phi] > 127) return false;
  }
  const int row_doubles = c * c * ppp;
  if (row_doubles > 4095 || 8 * c * 127 > 4095) return false;
  const int unit_doubles = (c % 2 == 0) ? 2 : 1;
  const int LU = row_doubles / unit_doubles;
  // Image rows are LU units long; one pad unit is inserted every g rows (row r starts at unit r LU + r / g) so that
  // the lanes' stores of one instruction fall on distinct banks: 16-byte units are served a quarter-warp at a time
  // over 8 bank groups, 8-byte units a half-warp over 16.  The largest g that keeps every such lane group
  // conflict-free is taken (config 2: LU = 28, g = 2 -- two rows leave in one bulk store); an odd row of 8-byte
  // units needs no padding at all (g = 0: the image is the tile's contiguous slice of the values).
  int pad_rows = 1;
  {
    const int lanes = (c % 2 == 0) ? 8 : 16;
    auto conflict_free = [&](int g) {
      for (int t0 = 0; t0 < kLaneSteps; t0 += lanes) {
        unsigned seen = 0;
        for (int t = t0; t < t0 + lanes; ++t) {
          const unsigned bit = 1u << (((int64_t)t * LU + (g ? t / g : 0)) % lanes);
          if (seen & bit) return false;
          seen |= bit;
        }
      }
      return true;
    };
    if (conflict_free(0)) pad_rows = 0;
    else if (step == P)  // rows of a strided run are not adjacent in the values: they leave one by one
      for (int g = 8; g >= 2; g >>= 1)
        if (conflict_free(g)) { pad_rows = g; break; }
    if (pad_rows == 1 && !conflict_free(1)) return false;
  }
  const int img_units = Tmax * LU + (pad_rows ? (Tmax + pad_rows - 1) / pad_rows : 0);

  // longest-processing-time split of the pairs over the warps (cost: gradient loads + a constant per pair)
  auto pair_cost = [&](const Pair& p) { return (double)p.codes.size() * (p.diag ? 0.6 : 1.0) + 1.5; };
  std::vector<int> order(pairs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pair_cost(pairs[x]) > pair_cost(pairs[y]); });
  std::vector<std::vector<int>> group(n_warps);
  std::vector<double> load(n_warps, 0.0);
  for (int i : order) {
    int w = 0;
    for (int x = 1; x < n_warps; ++x)
      if (load[x] < load[w]) w = x;
    group[w].push_back(i);
    load[w] += pair_cost(pairs[i]);
  }

  // emit
  std::vector<int32_t>& W = out.words;
  W.assign(kHdrWords, 0);
  int pad_shift1 = 0;  // 0: no padding; else 1 + log2(rows per pad unit)
  for (int g = pad_rows; g >= 1; g >>= 1) ++pad_shift1;
  W[0] = P; W[1] = n_lines; W[2] = TS; W[3] = pad_shift1; W[4] = LU; W[5] = (int32_t)spans.size();
  W[13] = (int32_t)coord_slots; W[14] = (int32_t)delta; W[15] = Tmax; W[16] = row_doubles; W[17] = (int32_t)rec_entries;
  W[8] = n_warps; W[9] = step; W[10] = kThreads / LU; W[11] = kThreads % LU;
  W[6] = (int32_t)W.size();
  for (auto& ln : lines) {
    const int64_t eoff = ln.d0 + ln.tmin * delta;
    if (eoff < -0x7fffffff || eoff > 0x7fffffff) return false;
    W.push_back((int32_t)eoff);
    W.push_back((int32_t)(ln.tmax - ln.tmin));
    for (int k = 0; k < N; ++k) {
      const int64_t ci = cidx_of(ln.ocan[k] + step * ln.tmin);
      if (ci < 0) return false;
      W.push_back((int32_t)ci);
    }
  }
  W[7] = (int32_t)W.size();
  for (auto& sp : spans) {
    W.push_back((int32_t)sp.start); W.push_back((int32_t)sp.n_ext); W.push_back((int32_t)sp.base); W.push_back(0);
  }
  for (int w = 0; w < n_warps; ++w) {
    W[20 + w] = (int32_t)W.size();
    W.push_back((int32_t)group[w].size());
    for (int i : group[w]) {
      const Pair& p = pairs[i];
      uint32_t lg, m;
      item_split_host((uint32_t)p.codes.size(), (uint32_t)L, lg, m);
      const uint32_t img_b = 8u * (uint32_t)(c * c * p.prefix + c * p.k), row_b = 8u * (uint32_t)(c * p.deg);
      W.push_back((int32_t)(img_b | (row_b << 15) | (lg << 27) | ((p.diag ? 1u : 0u) << 29)));
      // 2^lg partial sums of exactly L codes: partial j takes contributions [j m, j m + m), padded with the zero record
      for (uint32_t j = 0; j < (1u << lg); ++j)
        for (uint32_t x = 0; x < (uint32_t)L; ++x) {
          const uint32_t kk = j * m + x;
          W.push_back((x < m && kk < p.codes.size()) ? (int32_t)p.codes[kk] : 0);
        }
    }
  }
  while (W.size() % 4) W.push_back(0);  // 16-byte granularity
  W[12] = (int32_t)W.size();
  out.n_lines = n_lines; out.TS = TS; out.LU = LU; out.pad_rows = pad_rows; out.img_units = img_units; out.coord_nodes = (int)coord_slots; out.ppp = ppp;
  out.n_spans = (int)spans.size(); out.Tmax = Tmax; out.rec_entries = (int)rec_entries;

  std::vector<int32_t>& R = out.raw;
  R = {P, (int32_t)delta, ppp, step};
  for (int phi = 0; phi < P; ++phi) {
    const PhaseStencil& s = ph[phi];
    R.push_back((int32_t)s.de.size());
    R.push_back(deg_phi[phi]);
    R.push_back(prefix_phi[phi]);
    for (int r = 0; r < (int)s.de.size(); ++r) {
      R.push_back(s.de[r]);
      R.push_back(s.a[r]);
      for (int k = 0; k < N; ++k) R.push_back(s.corner[r * N + k]);
    }
  }
  return true;
}

// tiles of one run: balanced chunks of <= Tmax lane-steps (the template's choice)
inline void split_run(const Run& r, int Tmax, std::vector<std::pair<int32_t, int32_t>>& out /* (first lane-step, T) */) {
  const int n = (r.periods + Tmax - 1) / Tmax;
  const int base = r.periods / n, extra = r.periods % n;
  int32_t t0 = 0;
  for (int i = 0; i < n; ++i) {
    const int32_t T = base + (i < extra ? 1 : 0);
    out.emplace_back(t0, T);
    t0 += T;
  }
}

}  // namespace fes_runs_host
