// CPU emulator of k_assemble_runs (finite-element-solver_b200/csrc/assemble_runs.cuh): replays the template
// programs built by runs_template.hpp on the host, with the kernel's own index arithmetic, and compares the
// assembled values of every run-covered node with a plain per-pair assembly of the same mesh.  Test
// infrastructure only (tests/test_runs_host.py compiles and runs it with g++); it checks the run detection,
// the template construction and the addressing scheme without a GPU.
//
// Meshes follow the reference's structured generators (ref fem/mesh/structured.py:17-96): create_rect_mesh
// (nodes row-major in y, elements column-major, alternating diagonals) and create_box_mesh (Kuhn tets).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../finite-element-solver_b200/csrc/runs_template.hpp"

using namespace fes_runs_host;

struct Mesh {
  int N, sd;
  std::vector<int32_t> en;
  std::vector<double> xyz;
  int64_t n_nodes() const { return (int64_t)xyz.size() / sd; }
  int64_t n_el() const { return (int64_t)en.size() / N; }
};

static Mesh rect_mesh(int nx, int ny, double jitter, unsigned seed) {
  Mesh m{3, 2, {}, {}};
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(-jitter, jitter);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) { m.xyz.push_back(i + U(rng)); m.xyz.push_back(0.7 * j + U(rng)); }
  auto idx = [&](int i, int j) { return j * nx + i; };
  for (int i = 0; i < nx - 1; ++i)
    for (int j = 0; j < ny - 1; ++j) {
      if ((i + j) % 2 == 0) {
        const int32_t a[3] = {idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)}, b[3] = {idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)};
        m.en.insert(m.en.end(), a, a + 3); m.en.insert(m.en.end(), b, b + 3);
      } else {
        const int32_t a[3] = {idx(i, j), idx(i + 1, j), idx(i, j + 1)}, b[3] = {idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)};
        m.en.insert(m.en.end(), a, a + 3); m.en.insert(m.en.end(), b, b + 3);
      }
    }
  return m;
}

static Mesh box_mesh(int nx, int ny, int nz, double jitter, unsigned seed) {
  Mesh m{4, 3, {}, {}};
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(-jitter, jitter);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (int k = 0; k < nz; ++k) { m.xyz.push_back(i + U(rng)); m.xyz.push_back(1.1 * j + U(rng)); m.xyz.push_back(0.9 * k + U(rng)); }
  auto idx = [&](int i, int j, int k) { return (i * ny + j) * nz + k; };
  static const int KUHN[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};
  for (int i = 0; i < nx - 1; ++i)
    for (int j = 0; j < ny - 1; ++j)
      for (int k = 0; k < nz - 1; ++k) {
        int32_t corner[8];
        for (int c = 0; c < 8; ++c) corner[c] = idx(i + ((c >> 2) & 1), j + ((c >> 1) & 1), k + (c & 1));
        for (auto& tet : KUHN)
          for (int x = 0; x < 4; ++x) m.en.push_back(corner[tet[x]]);
      }
  return m;
}

// scaled gradients of one element (sqrt(vol) grad phi_n), plain arithmetic
static void element_grads(const Mesh& m, int64_t e, double* g /* N x sd */) {
  const int N = m.N, sd = m.sd;
  double X[4][3];
  for (int n = 0; n < N; ++n)
    for (int s = 0; s < sd; ++s) X[n][s] = m.xyz[(int64_t)m.en[e * N + n] * sd + s];
  double cof[3][3], det;
  if (sd == 2) {
    const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0], c = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
    det = a * d - b * c;
    cof[0][0] = d; cof[0][1] = -b; cof[1][0] = -c; cof[1][1] = a;
  } else {
    double J[3][3];
    for (int s = 0; s < 3; ++s)
      for (int r = 0; r < 3; ++r) J[s][r] = X[r + 1][s] - X[0][s];
    cof[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    cof[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    cof[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * cof[0][0] + J[0][1] * cof[1][0] + J[0][2] * cof[2][0];
    cof[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    cof[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    cof[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    cof[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    cof[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    cof[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  }
  const double ref = sd == 2 ? 0.5 : 1.0 / 6.0;
  const double f = copysign(sqrt(ref) / sqrt(fabs(det)), det);
  for (int s = 0; s < sd; ++s) {
    double first = 0.0;
    for (int r = 0; r < sd; ++r) { g[(r + 1) * sd + s] = cof[r][s] * f; first -= cof[r][s] * f; }
    g[s] = first;
  }
}

struct Emu {
  const Mesh& m;
  int c;  // 1 = Laplacian, sd = elasticity
  std::vector<int32_t> ne_ptr, node_ptr, node_adj;
  std::vector<uint32_t> ne_code;
  explicit Emu(const Mesh& mesh, int comps) : m(mesh), c(comps) {
    const int64_t nn = m.n_nodes(), ne = m.n_el();
    const int N = m.N;
    std::vector<std::vector<uint32_t>> inc(nn);
    for (int64_t e = 0; e < ne; ++e)
      for (int a = 0; a < N; ++a) inc[m.en[e * N + a]].push_back((uint32_t)e * 8u + a);
    ne_ptr.assign(nn + 1, 0);
    node_ptr.assign(nn + 1, 0);
    for (int64_t v = 0; v < nn; ++v) {
      ne_ptr[v + 1] = ne_ptr[v] + (int32_t)inc[v].size();
      ne_code.insert(ne_code.end(), inc[v].begin(), inc[v].end());
      std::vector<int32_t> nb;
      for (uint32_t code : inc[v])
        for (int b = 0; b < N; ++b) nb.push_back(m.en[(int64_t)(code >> 3) * N + b]);
      std::sort(nb.begin(), nb.end());
      nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
      node_ptr[v + 1] = node_ptr[v] + (int32_t)nb.size();
      node_adj.insert(node_adj.end(), nb.begin(), nb.end());
    }
  }

  // plain assembly of node v's rows: block (v, w_k) = sum over shared elements, closed form applied per pair
  void reference_rows(int64_t v, double lam_over_mu, std::vector<double>& out) const {
    const int N = m.N, sd = m.sd;
    const int deg = node_ptr[v + 1] - node_ptr[v];
    out.assign((size_t)c * c * deg, 0.0);
    for (int k = 0; k < deg; ++k) {
      const int32_t wn = node_adj[node_ptr[v] + k];
      double S[3][3] = {{0}};
      for (int32_t g = ne_ptr[v]; g < ne_ptr[v + 1]; ++g) {
        const int64_t e = ne_code[g] >> 3;
        const int a = ne_code[g] & 7;
        double gr[12];
        element_grads(m, e, gr);
        for (int b = 0; b < N; ++b)
          if (m.en[e * N + b] == wn)
            for (int i = 0; i < sd; ++i)
              for (int j = 0; j < sd; ++j) S[i][j] += gr[a * sd + i] * gr[b * sd + j];
      }
      if (c == 1) {
        double t = 0;
        for (int s = 0; s < sd; ++s) t += S[s][s];
        out[k] = t;
      } else {
        double tr = 0;
        for (int i = 0; i < sd; ++i) tr += S[i][i];
        for (int i = 0; i < c; ++i)
          for (int j = 0; j < c; ++j) out[(size_t)i * c * deg + c * k + j] = lam_over_mu * S[i][j] + S[j][i] + (i == j ? tr : 0.0);
      }
    }
  }
};

// the kernel, on the host: one tile
static bool emulate_tile(const Emu& E, const int32_t* hdr, const std::vector<int32_t>& blob, double lam_over_mu,
                         std::vector<double>& values) {
  const Mesh& m = E.m;
  const int N = m.N, SD = m.sd, C = E.c;
  const int32_t v0 = hdr[0], T = hdr[1], toff = hdr[2], pair0 = hdr[4], pstride = hdr[6];
  const int32_t* s_t = blob.data() + toff;
  const int step = s_t[9];
  const int n_lines = s_t[1], TS = s_t[2], psh = s_t[3], LU = s_t[4], n_spans = s_t[5];
  auto row_off = [&](int r) { return (size_t)r * LU + (psh ? (r >> (psh - 1)) : 0); };  // units
  const int off_lines = s_t[6], off_spans = s_t[7];
  const int UD = (C % 2 == 0) ? 2 : 1;  // doubles per unit
  const int64_t nn = m.n_nodes();
  if (T > s_t[15] || T < 1) { printf("tile longer than Tmax\n"); return false; }
  std::vector<double> s_xyz((size_t)s_t[13] * SD, NAN);
  std::vector<double> rec((size_t)s_t[17] * SD, NAN);  // [entry][SD]
  for (int i = 0; i < kZeroEntries * SD; ++i) rec[i] = 0.0;
  std::vector<double> img((row_off(kLaneSteps) + 1) * UD, NAN);
  // staging: strided spans
  for (int s = 0; s < n_spans; ++s) {
    const int32_t* sp = s_t + off_spans + 4 * s;
    const int cnt = T + sp[1];
    for (int i = 0; i < cnt; ++i) {
      const int64_t node = (int64_t)v0 + sp[0] + (int64_t)step * i;
      const int d = sp[2] + i;
      if (d >= s_t[13]) { printf("span overflow\n"); return false; }
      if (node >= 0 && node < nn)
        for (int x = 0; x < SD; ++x) s_xyz[(size_t)d * SD + x] = m.xyz[node * SD + x];
    }
  }
  // phase 1: warp per line, lane = entry
  if (TS != kLaneSteps) { printf("plane stride is not 32\n"); return false; }
  for (int l = 0; l < n_lines; ++l)
    for (int lane = 0; lane < 32; ++lane) {
      const int32_t* ld = s_t + off_lines + l * (2 + N);
      if (T + ld[1] > 32) { printf("a line has more than 32 entries\n"); return false; }
      if (lane >= T + ld[1]) continue;
      // geometry from the TEMPLATE's corner table (not from the connectivity), like the kernel
      Mesh tmp{N, SD, {}, {}};
      for (int n = 0; n < N; ++n) {
        const int ci = ld[2 + n] + lane;
        if (ci < 0 || ci >= s_t[13]) { printf("coordinate index out of the table\n"); return false; }
        tmp.en.push_back(n);
        for (int x = 0; x < SD; ++x) tmp.xyz.push_back(s_xyz[(size_t)ci * SD + x]);
      }
      double g[12];
      element_grads(tmp, 0, g);
      for (int n = 0; n < N; ++n) {
        const size_t ent = (size_t)kZeroEntries + (size_t)(l * N + n) * 32 + lane;
        if ((int)ent >= s_t[17]) { printf("record entry out of range\n"); return false; }
        for (int x = 0; x < SD; ++x) rec[ent * SD + x] = g[n * SD + x];
      }
    }
  // phase 2
  const int L = (N == 3) ? 2 : 6;
  for (int warp = 0; warp < s_t[8]; ++warp) {
    const int32_t* prog = s_t + s_t[20 + warp];
    const int np = prog[0];
    for (int lane = 0; lane < T; ++lane) {
      int w = 1;
      for (int pi = 0; pi < np; ++pi) {
        const uint32_t head = (uint32_t)prog[w];
        const uint32_t im = (head & 0x7fffu) / 8u, row = ((head >> 15) & 0xfffu) / 8u, lg = (head >> 27) & 3u, diag = (head >> 29) & 1u;
        double S[3][3] = {{0}};
        for (uint32_t kk = 0; kk < ((uint32_t)L << lg); ++kk) {
          const uint32_t code = (uint32_t)prog[w + 1 + (int)kk], ea = (code & 0xffffu) / 16u, eb = (code >> 16) / 16u;
          if ((code & 0xf000fu) != 0) { printf("code offsets not multiples of 16\n"); return false; }
          if (diag && ea != eb) { printf("diag pair with a != b\n"); return false; }
          if ((int)(ea + lane) >= s_t[17] || (int)(eb + lane) >= s_t[17]) { printf("entry out of range\n"); return false; }
          const double* A = &rec[(size_t)(ea + lane) * SD];
          const double* B = &rec[(size_t)(eb + lane) * SD];
          for (int i = 0; i < SD; ++i)
            for (int j = 0; j < SD; ++j) S[i][j] += A[i] * B[j];   // a NaN here = an entry nobody computed
        }
        w += 1 + (L << lg);
        double* rowp = &img[row_off(lane) * UD];
        if (C == 1) {
          double t = 0;
          for (int s = 0; s < SD; ++s) t += S[s][s];
          rowp[im] = t;
        } else {
          double tr = 0;
          for (int i = 0; i < SD; ++i) tr += S[i][i];
          for (int i = 0; i < C; ++i)
            for (int j = 0; j < C; ++j) rowp[im + (uint32_t)i * row + j] = lam_over_mu * S[i][j] + S[j][i] + (i == j ? tr : 0.0);
        }
      }
    }
  }
  {  // the stores of one instruction: distinct banks within a quarter-warp (16-byte units) / half-warp (8-byte)
    const int lanes = (UD == 2) ? 8 : 16;
    for (int t0 = 0; t0 < kLaneSteps; t0 += lanes) {
      unsigned seen = 0;
      for (int t = t0; t < t0 + lanes; ++t) {
        const unsigned bit = 1u << (row_off(t) % lanes);
        if (seen & bit) { printf("image rows collide on a bank\n"); return false; }
        seen |= bit;
      }
    }
  }
  // store: row r of the image -> values[c^2 pair0 + r * row_doubles ...]
  const int kThreads = 32 * s_t[8];
  const int GR = C * C * pstride / UD;  // units between consecutive image rows in `values`
  if (GR < LU) { printf("rows overlap in the output\n"); return false; }
  for (int tid = 0; tid < kThreads; ++tid) {  // the kernel's flat copy: thread tid takes units tid, tid + threads, ...
    int r = tid / LU, u = tid % LU;
    for (int x = tid; x < T * LU; x += kThreads) {
      if (r != x / LU || u != x % LU) { printf("store increments wrong\n"); return false; }
      for (int d = 0; d < UD; ++d) {
        double& out = values[(size_t)C * C * pair0 + ((size_t)r * GR + u) * UD + d];
        if (!std::isnan(out)) { printf("a value is written twice\n"); return false; }
        out = img[(row_off(r) + u) * UD + d];
      }
      r += s_t[10]; u += s_t[11];
      if (u >= LU) { u -= LU; ++r; }
    }
  }
  return true;
}

static bool host_verify(const Emu& E, const int32_t* h, const std::vector<int32_t>& rawblob) {
  const Mesh& m = E.m;
  const int N = m.N;
  const int32_t v0 = h[0], T = h[1], E0 = h[3], pair0 = h[4], pstride = h[6];
  const int32_t* R = rawblob.data() + h[5];
  const int P = R[0];
  const int64_t delta = R[1];
  const int step = R[3];
  for (int t = 0; t < T; ++t) {
    const int32_t* p = R + 4;
    for (int phi = 0; phi < P; ++phi) {
      const int ninc = p[0], deg = p[1], prefix = p[2];
      p += 3;
      const int64_t base = (int64_t)v0 + (int64_t)step * t, v = base + phi;
      if (v >= m.n_nodes()) return false;
      const int32_t g0 = E.ne_ptr[v];
      if (E.ne_ptr[v + 1] - g0 != ninc || E.node_ptr[v] != pair0 + t * pstride + prefix || E.node_ptr[v + 1] - E.node_ptr[v] != deg) return false;
      for (int r = 0; r < ninc; ++r) {
        const int64_t e = (int64_t)E0 + (int64_t)t * delta + p[0];
        if (e < 0 || e >= m.n_el()) return false;
        if (E.ne_code[g0 + r] != (uint32_t)e * 8u + (uint32_t)p[1]) return false;
        for (int k = 0; k < N; ++k)
          if ((int64_t)m.en[e * N + k] != base + p[2 + k]) return false;
        p += 2 + N;
      }
    }
  }
  return true;
}

// returns the covered fraction (< 0 on failure)
static double run_case(const char* name, const Mesh& m, int c, double min_cover, int n_warps = 4) {
  Emu E(m, c);
  const int64_t nn = m.n_nodes();
  const int N = m.N;
  std::vector<uint64_t> sig(nn);
  std::vector<int32_t> eref(nn), ninc(nn);
  for (int64_t v = 0; v < nn; ++v) {
    ninc[v] = E.ne_ptr[v + 1] - E.ne_ptr[v];
    sig[v] = node_signature((int32_t)v, ninc[v], E.ne_code.data() + E.ne_ptr[v], m.en.data(), N);
    eref[v] = ninc[v] ? (int32_t)(E.ne_code[E.ne_ptr[v]] >> 3) : -1;
  }
  std::vector<Run> runs = segment_runs(sig.data(), eref.data(), ninc.data(), nn);
  size_t n_strided = 0;
  {
    std::vector<char> free_node(nn, 1);
    for (Run& r : runs) {
      r.pstride = E.node_ptr[r.v0 + r.P] - E.node_ptr[r.v0];
      for (int64_t v = r.v0; v < (int64_t)r.v0 + (int64_t)r.P * r.periods; ++v) free_node[v] = 0;
    }
    for (int pass = 0; pass < 4; ++pass) {   // as plan_runs.cu: strided passes, then single-node tiles for a handful of corners
      const std::vector<Run> strided = segment_strided(sig.data(), eref.data(), ninc.data(), E.node_ptr.data(), nn, free_node);
      if (strided.empty()) break;
      n_strided += strided.size();
      runs.insert(runs.end(), strided.begin(), strided.end());
    }
    std::vector<int64_t> left;
    for (int64_t v = 0; v < nn && left.size() <= 64; ++v)
      if (free_node[v]) left.push_back(v);
    if (left.size() <= 64)
      for (int64_t v : left)
        if (ninc[v] > 0) runs.push_back(Run{(int32_t)v, 1, 1, 1, 1, E.node_ptr[v + 1] - E.node_ptr[v]});
  }
  std::vector<int32_t> blob, rawblob, tiles;
  std::vector<char> covered(nn, 0);
  int n_tmpl = 0, n_fail = 0;
  const int L = (N == 3) ? 2 : 6;
  for (const Run& r : runs) {
    std::vector<PhaseStencil> ph(r.P);
    for (int phi = 0; phi < r.P; ++phi) {
      const int64_t v = (int64_t)r.v0 + phi;
      for (int32_t g = E.ne_ptr[v]; g < E.ne_ptr[v + 1]; ++g) {
        const int64_t e = E.ne_code[g] >> 3;
        ph[phi].de.push_back((int32_t)(e - eref[r.v0]));
        ph[phi].a.push_back((int32_t)(E.ne_code[g] & 7u));
        for (int x = 0; x < N; ++x) ph[phi].corner.push_back(m.en[e * N + x] - r.v0);
      }
    }
    Template t;  // the emulator builds one template per run (the library shares them by key)
    if (!build_template(N, c, r.P, r.delta, ph, L, n_warps, t, r.step)) { ++n_fail; continue; }
    ++n_tmpl;
    if (getenv("RUNS_EMU_VERBOSE"))
      printf("  run v0=%d P=%d periods=%d delta=%lld: lines=%d Tmax=%d TS=%d LU=%d pad_rows=%d coord nodes=%d spans=%d words=%zu\n", r.v0, r.P,
             r.periods, (long long)r.delta, t.n_lines, t.Tmax, t.TS, t.LU, t.pad_rows, t.coord_nodes, t.n_spans, t.words.size());
    const int32_t woff = (int32_t)blob.size(), roff = (int32_t)rawblob.size();
    blob.insert(blob.end(), t.words.begin(), t.words.end());
    rawblob.insert(rawblob.end(), t.raw.begin(), t.raw.end());
    std::vector<std::pair<int32_t, int32_t>> parts;
    split_run(r, t.Tmax, parts);
    for (auto& pt : parts) {
      const int32_t v0 = r.v0 + r.step * pt.first;
      const int32_t hdr[kTileInts] = {v0, pt.second, woff, eref[v0], E.node_ptr[v0], roff, r.pstride, 0};
      tiles.insert(tiles.end(), hdr, hdr + kTileInts);
      for (int32_t k = 0; k < pt.second; ++k)
        for (int phi = 0; phi < r.P; ++phi) {
          if (covered[v0 + r.step * k + phi]) { printf("%s: node covered twice\n", name); return -1; }
          covered[v0 + r.step * k + phi] = 1;
        }
    }
  }
  const double lam_over_mu = 1.5;
  std::vector<double> values((size_t)c * c * E.node_ptr[nn], NAN);
  const int64_t nt = (int64_t)tiles.size() / kTileInts;
  for (int64_t t = 0; t < nt; ++t) {
    if (!host_verify(E, tiles.data() + t * kTileInts, rawblob)) { printf("%s: verification failed on tile %lld\n", name, (long long)t); return -1; }
    if (!emulate_tile(E, tiles.data() + t * kTileInts, blob, lam_over_mu, values)) { printf("%s: emulation failed\n", name); return -1; }
  }
  int64_t n_cov = 0;
  double worst = 0.0;
  std::vector<double> ref;
  for (int64_t v = 0; v < nn; ++v) {
    const int deg = E.node_ptr[v + 1] - E.node_ptr[v];
    const double* got = &values[(size_t)c * c * E.node_ptr[v]];
    if (!covered[v]) {
      for (int i = 0; i < c * c * deg; ++i)
        if (!std::isnan(got[i])) { printf("%s: a value of uncovered node %lld was written\n", name, (long long)v); return -1; }
      continue;
    }
    ++n_cov;
    E.reference_rows(v, lam_over_mu, ref);
    double scale = 0;
    for (double x : ref) scale = std::max(scale, fabs(x));
    for (int i = 0; i < c * c * deg; ++i) {
      if (std::isnan(got[i])) { printf("%s: value %d of node %lld never written\n", name, i, (long long)v); return -1; }
      worst = std::max(worst, fabs(got[i] - ref[i]) / scale);
    }
  }
  const double cover = (double)n_cov / (double)nn;
  printf("%s: nodes=%lld runs=%zu (strided %zu) templates=%d unencodable=%d tiles=%lld covered=%.3f max rel err=%.2e\n", name, (long long)nn,
         runs.size(), n_strided, n_tmpl, n_fail, (long long)nt, cover, worst);
  if (worst > 1e-13) { printf("%s: values differ\n", name); return -1; }
  if (cover < min_cover) { printf("%s: coverage below %.2f\n", name, min_cover); return -1; }
  return cover;
}

int main() {
  bool ok = true;
  ok &= run_case("rect 41x23 laplace", rect_mesh(41, 23, 0.2, 1), 1, 0.8) >= 0;
  ok &= run_case("rect 41x23 elastic", rect_mesh(41, 23, 0.2, 2), 2, 0.8) >= 0;
  ok &= run_case("rect 140x12 elastic", rect_mesh(140, 12, 0.15, 3), 2, 0.9) >= 0;
  ok &= run_case("rect 12x40 elastic (short rows)", rect_mesh(12, 40, 0.15, 4), 2, 0.0) >= 0;
  ok &= run_case("box 7x6x45 elastic", box_mesh(7, 6, 45, 0.15, 5), 3, 0.8) >= 0;
  ok &= run_case("box 5x5x21 laplace", box_mesh(5, 5, 21, 0.15, 6), 1, 0.8) >= 0;
  ok &= run_case("box 6x5x80 elastic", box_mesh(6, 5, 80, 0.1, 7), 3, 0.9) >= 0;
  ok &= run_case("box 6x5x80 elastic, 8 warps", box_mesh(6, 5, 80, 0.1, 7), 3, 0.9, 8) >= 0;
  ok &= run_case("rect 140x12 elastic, 8 warps", rect_mesh(140, 12, 0.15, 3), 2, 0.9, 8) >= 0;
  ok &= run_case("box 6x5x80 elastic, 6 warps", box_mesh(6, 5, 80, 0.1, 7), 3, 0.9, 6) >= 0;
  // strided runs: the end nodes of the mesh rows / z-lines (node stride = row pitch, twice that with alternating diagonals)
  ok &= run_case("rect 44x40 elastic (strided row ends)", rect_mesh(44, 40, 0.15, 8), 2, 0.97) >= 0;
  ok &= run_case("box 4x14x24 elastic (strided line ends)", box_mesh(4, 14, 24, 0.1, 9), 3, 0.97, 6) >= 0;
  ok &= run_case("box 4x14x24 laplace (strided line ends)", box_mesh(4, 14, 24, 0.1, 9), 1, 0.97) >= 0;
  {  // an irregular patch inside a structured mesh: the other diagonal in a block of cells.  The nodes around it fall
     // out of the runs (their rows are cut into shorter runs or left to the tiled kernel); everything still covered
     // by a tile must be exact.
    Mesh m = rect_mesh(60, 34, 0.1, 11);
    const int nx = 60, ny = 34;
    for (int i = 20; i < 29; ++i)
      for (int j = 10; j < 19; ++j) {
        const int32_t a = j * nx + i, b = a + 1, c = (j + 1) * nx + i + 1, d = (j + 1) * nx + i;
        const int64_t k = 2 * ((int64_t)i * (ny - 1) + j);
        const int32_t t0[3] = {a, b, d}, t1[3] = {b, c, d}, u0[3] = {a, b, c}, u1[3] = {a, c, d};
        const bool even = (i + j) % 2 == 0;
        for (int x = 0; x < 3; ++x) { m.en[3 * k + x] = even ? t0[x] : u0[x]; m.en[3 * (k + 1) + x] = even ? t1[x] : u1[x]; }
      }
    const double cover = run_case("rect 60x34 elastic with an irregular patch", m, 2, 0.5);
    ok &= cover >= 0 && cover < 0.99;
  }
  {  // a random node numbering: no run at all, nothing written
    Mesh m = rect_mesh(30, 30, 0.1, 12);
    std::vector<int32_t> perm(m.n_nodes());
    for (size_t i = 0; i < perm.size(); ++i) perm[i] = (int32_t)i;
    std::mt19937 rng(5);
    std::shuffle(perm.begin(), perm.end(), rng);
    std::vector<double> xyz(m.xyz.size());
    for (size_t v = 0; v < perm.size(); ++v) { xyz[2 * perm[v]] = m.xyz[2 * v]; xyz[2 * perm[v] + 1] = m.xyz[2 * v + 1]; }
    m.xyz = xyz;
    for (auto& n : m.en) n = perm[n];
    ok &= run_case("rect 30x30 elastic, random node numbering", m, 2, 0.0) == 0.0;
  }
  printf(ok ? "RUNS_EMULATOR PASS\n" : "RUNS_EMULATOR FAIL\n");
  return ok ? 0 : 1;
}
