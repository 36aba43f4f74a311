// Stencil-run detection for the fused assembly (see runs_template.hpp for the idea and the template layout).
// Device: one signature per node over its RELATIVE stencil, and an exact verification of every covered node
// against its template (a hash match alone is never trusted).  Host: O(n_nodes) segmentation into runs, one
// template program per distinct (period, signatures, element stride).
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "plan.cuh"
#include "runs_template.hpp"

namespace fes {
namespace {
using namespace fes_runs_host;

constexpr int kBlock = 256;
constexpr int kMaxTemplates = 256;

__global__ void k_node_sig(const int32_t* __restrict__ ne_ptr, const uint32_t* __restrict__ ne_code,
                           const int32_t* __restrict__ en, int N, int64_t n_nodes, uint64_t* __restrict__ sig,
                           int32_t* __restrict__ eref) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const int32_t g0 = ne_ptr[v], ninc = ne_ptr[v + 1] - g0;
  sig[v] = node_signature((int32_t)v, ninc, ne_code + g0, en, N);
  eref[v] = ninc ? (int32_t)(ne_code[g0] >> 3) : -1;
}

// one thread per (tile, lane-step): the node's incident list, element connectivity, degree and value offset
// are exactly what the template predicts
__global__ void k_verify_runs(const int32_t* __restrict__ tiles, int64_t n_tiles, const int32_t* __restrict__ raw,
                              const int32_t* __restrict__ ne_ptr, const uint32_t* __restrict__ ne_code,
                              const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ en, int N,
                              int64_t n_nodes, int64_t n_el, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tile = i / kLaneSteps;
  const int t = (int)(i - tile * kLaneSteps);
  if (tile >= n_tiles) return;
  const int32_t* h = tiles + tile * kTileInts;
  const int32_t v0 = h[0], T = h[1], E0 = h[3], pair0 = h[4], pstride = h[6];
  if (t >= T) return;
  const int32_t* R = raw + h[5];
  const int P = R[0];
  const int64_t delta = R[1];
  const int step = R[3];
  const int32_t* p = R + 4;
  bool ok = true;
  for (int phi = 0; phi < P; ++phi) {
    const int ninc = p[0], deg = p[1], prefix = p[2];
    p += 3;
    const int64_t base = (int64_t)v0 + (int64_t)step * t, v = base + phi;
    if (v >= n_nodes) { ok = false; break; }
    const int32_t g0 = ne_ptr[v];
    ok = ok && (ne_ptr[v + 1] - g0 == ninc) && (node_ptr[v] == pair0 + t * pstride + prefix) && (node_ptr[v + 1] - node_ptr[v] == deg);
    if (!ok) break;
    for (int r = 0; r < ninc; ++r) {
      const int64_t e = (int64_t)E0 + (int64_t)t * delta + p[0];
      if (e < 0 || e >= n_el) { ok = false; break; }
      ok = ok && ne_code[g0 + r] == (uint32_t)e * 8u + (uint32_t)p[1];
      for (int k = 0; k < N; ++k) ok = ok && (int64_t)en[e * N + k] == base + p[2 + k];
      p += 2 + N;
    }
    if (!ok) break;
  }
  if (!ok) atomicExch(bad, 1);
}

}  // namespace

// Detect the runs of plan P, build their templates and tiles, verify them on the device.  Leaves
// P->runs.active = false when nothing (or too little) qualifies; otherwise `rest_ranges` receives the node
// ranges no run covers, as {begin, end} pairs for build_tiles.
int build_runs(fes_plan* P, const int32_t* d_elem_nodes, const std::vector<int32_t>& h_ne, const std::vector<int32_t>& h_np,
               cudaStream_t stream, std::vector<int32_t>& rest_ranges) {
  fes_runs& Rn = P->runs;
  Rn.active = false;
  rest_ranges.clear();
  const int64_t nn = P->n_nodes;
  const int N = P->N, c = P->c;
  if (P->n_geo == 0 || (N != 3 && N != 4) || nn < 2 * kMinPeriods) return FES_OK;  // the run kernel covers P1 triangles and tets
  DevBuf<uint64_t> d_sig;
  DevBuf<int32_t> d_eref;
  FES_CUDA(d_sig.alloc(nn));
  FES_CUDA(d_eref.alloc(nn));
  k_node_sig<<<grid_for(nn, kBlock), kBlock, 0, stream>>>(P->ne_ptr.p, P->ne_code.p, d_elem_nodes, N, nn, d_sig.p, d_eref.p);
  FES_LAUNCH_CHECK();
  std::vector<uint64_t> sig(nn);
  std::vector<int32_t> eref(nn), ninc(nn);
  FES_CUDA(cudaMemcpyAsync(sig.data(), d_sig.p, nn * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaMemcpyAsync(eref.data(), d_eref.p, nn * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  d_sig.release();
  d_eref.release();
  for (int64_t v = 0; v < nn; ++v) ninc[v] = h_ne[v + 1] - h_ne[v];
  std::vector<Run> runs = segment_runs(sig.data(), eref.data(), ninc.data(), nn);
  if (runs.empty()) return FES_OK;
  // strided runs among the nodes left over (row / z-line end nodes of structured meshes); ascending v0 overall
  {
    std::vector<char> free_node(nn, 1);
    for (Run& r : runs) {
      r.pstride = h_np[r.v0 + r.P] - h_np[r.v0];
      for (int64_t v = r.v0; v < (int64_t)r.v0 + (int64_t)r.P * r.periods; ++v) free_node[v] = 0;
    }
    if (!getenv("FES_NO_STRIDED_RUNS")) {
      // repeated: what one pass leaves (the four corner nodes of every i-plane of a box: the ends of the strided runs
      // along j) is itself strided (along i)
      for (int pass = 0; pass < 4; ++pass) {
        const std::vector<Run> strided = segment_strided(sig.data(), eref.data(), ninc.data(), h_np.data(), nn, free_node);
        if (strided.empty()) break;
        runs.insert(runs.end(), strided.begin(), strided.end());
      }
      // A handful of nodes left (the corners of a structured mesh): one single-node tile each, so the step needs no
      // second kernel at all (a remainder kernel on a side stream costs ~8 us of fork / join per assembly however
      // small it is).  More than that is an irregular region: it keeps the tiled kernel.
      std::vector<int64_t> left;
      for (int64_t v = 0; v < nn && left.size() <= 64; ++v)
        if (free_node[v]) left.push_back(v);
      if (left.size() <= 64)
        for (int64_t v : left)
          if (ninc[v] > 0) runs.push_back(Run{(int32_t)v, 1, 1, 1, 1, h_np[v + 1] - h_np[v]});
    }
  }

  // one template per distinct (P, signatures, element stride, phase offset of the element base)
  using Key = std::tuple<int, uint64_t, uint64_t, int64_t, int64_t, int>;
  std::map<Key, int> tmpl_of;  // -> index into tmpls, or -1 (not encodable)
  std::vector<Template> tmpls;
  std::vector<int32_t> word_off, raw_off, blob, raw;
  std::vector<int32_t> tiles;
  std::vector<int32_t> covered;  // {begin, end} node ranges, ascending
  std::vector<std::pair<int32_t, int32_t>> covered_ranges;
  const int L = tile_item_len(N);
  // warps per CTA (= pair groups of a template), measured at 151^3 / config 2: tet elasticity (15 pairs, 96
  // contributions of 9 FMAs per node, 18 element lines = three per warp in phase 1) 1.396 ms with 6, 1.578 with 4,
  // 1.549 with 8; the scalar tet forms 1.158 ms with 4, 1.270 with 6; triangles (14 pairs per period, 8 lines)
  // 0.1345 ms with 4, 0.155 with 8
  int n_warps = (N == 4 && c > 1) ? 6 : 4;
  if (const char* env = getenv("FES_RUNS_WARPS")) n_warps = (atoi(env) >= 4 && atoi(env) <= 8 && atoi(env) != 5) ? atoi(env) : 4;  // tuning knob: 4, 6, 7, 8
  int64_t n_cov = 0;
  for (const Run& r : runs) {
    const Key key{r.P, sig[r.v0], r.P > 1 ? sig[r.v0 + 1] : 0ull, r.delta,
                  r.P > 1 ? (int64_t)eref[r.v0 + 1] - eref[r.v0] : 0, r.step};
    auto it = tmpl_of.find(key);
    int ti;
    if (it != tmpl_of.end()) {
      ti = it->second;
    } else if ((int)tmpls.size() >= kMaxTemplates) {
      ti = -1;
    } else {
      // the representative period's stencil: incident lists + connectivity rows (small copies, once per template)
      std::vector<PhaseStencil> ph(r.P);
      for (int phi = 0; phi < r.P; ++phi) {
        const int64_t v = (int64_t)r.v0 + phi;
        const int32_t g0 = h_ne[v], k = h_ne[v + 1] - g0;
        std::vector<uint32_t> codes(k);
        FES_CUDA(cudaMemcpyAsync(codes.data(), P->ne_code.p + g0, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        FES_CUDA(cudaStreamSynchronize(stream));
        std::vector<int32_t> rows((size_t)k * N);
        for (int i = 0; i < k; ++i)
          FES_CUDA(cudaMemcpyAsync(rows.data() + (size_t)i * N, d_elem_nodes + (int64_t)(codes[i] >> 3) * N, N * sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, stream));
        FES_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < k; ++i) {
          ph[phi].de.push_back((int32_t)((int64_t)(codes[i] >> 3) - eref[r.v0]));
          ph[phi].a.push_back((int32_t)(codes[i] & 7u));
          for (int x = 0; x < N; ++x) ph[phi].corner.push_back(rows[(size_t)i * N + x] - r.v0);
        }
      }
      Template t;
      if (build_template(N, c, r.P, r.delta, ph, L, n_warps, t, r.step) && blob.size() + t.words.size() <= (size_t)fes_runs::kMaxTmplWords) {
        ti = (int)tmpls.size();
        word_off.push_back((int32_t)blob.size());
        raw_off.push_back((int32_t)raw.size());
        blob.insert(blob.end(), t.words.begin(), t.words.end());
        raw.insert(raw.end(), t.raw.begin(), t.raw.end());
        tmpls.push_back(std::move(t));
      } else {
        ti = -1;
      }
      tmpl_of[key] = ti;
    }
    if (ti < 0) continue;
    std::vector<std::pair<int32_t, int32_t>> parts;
    split_run(r, tmpls[ti].Tmax, parts);
    for (auto& pt : parts) {
      const int32_t v0 = r.v0 + r.step * pt.first;
      const int32_t hdr[kTileInts] = {v0, pt.second, word_off[ti], eref[v0], h_np[v0], raw_off[ti], r.pstride, 0};
      tiles.insert(tiles.end(), hdr, hdr + kTileInts);
    }
    if (r.step == r.P) {
      covered_ranges.emplace_back(r.v0, r.v0 + r.P * r.periods);
    } else {
      for (int32_t k = 0; k < r.periods; ++k) covered_ranges.emplace_back(r.v0 + r.step * k, r.v0 + r.step * k + 1);
    }
    n_cov += (int64_t)r.P * r.periods;
  }
  // the covered nodes as ascending, merged {begin, end} ranges
  std::sort(covered_ranges.begin(), covered_ranges.end());
  for (auto& cr : covered_ranges) {
    if (!covered.empty() && covered.back() == cr.first) covered.back() = cr.second;
    else { covered.push_back(cr.first); covered.push_back(cr.second); }
  }
  const int64_t nt = (int64_t)tiles.size() / kTileInts;
  // Worth two kernels (run tiles + remainder tiles on a side stream) only when the runs cover a good part of a mesh
  // that is large enough: below ~10^5 nodes an assembly is launch-latency bound and one kernel wins (config 1:
  // 15 us tiled, 22 us split).  Both thresholds are knobs, read per plan (the tests force small meshes through).
  const double min_cover = getenv("FES_RUNS_MIN_COVER") ? atof(getenv("FES_RUNS_MIN_COVER")) : 0.25;
  const long long min_nodes = getenv("FES_RUNS_MIN_NODES") ? atoll(getenv("FES_RUNS_MIN_NODES")) : 100000;
  if (nt == 0 || (double)n_cov < min_cover * (double)nn || n_cov < min_nodes) return FES_OK;

  FES_CUDA(Rn.tiles.alloc(tiles.size()));
  DevBuf<int32_t> d_raw;
  DevBuf<int> bad;
  FES_CUDA(d_raw.alloc(raw.size()));
  FES_CUDA(bad.alloc(1));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  FES_CUDA(cudaMemcpyAsync(Rn.tiles.p, tiles.data(), tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  FES_CUDA(cudaMemcpyAsync(d_raw.p, raw.data(), raw.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  k_verify_runs<<<grid_for(nt * kLaneSteps, kBlock), kBlock, 0, stream>>>(Rn.tiles.p, nt, d_raw.p, P->ne_ptr.p, P->ne_code.p,
                                                                         P->node_ptr.p, d_elem_nodes, N, nn, P->n_el, bad.p);
  FES_LAUNCH_CHECK();
  int h_bad = 0;
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) {  // a signature collision or an inconsistent template: the tiled kernel keeps every node
    if (getenv("FES_PLAN_DEBUG")) fprintf(stderr, "[fes runs] verification failed: runs disabled\n");
    Rn.tiles.release();
    return FES_OK;
  }
  Rn.n_tiles = nt;
  Rn.n_nodes_covered = n_cov;
  Rn.n_templates = (int64_t)tmpls.size();
  Rn.max_rec_entries = Rn.max_coord_nodes = Rn.max_img_units = Rn.max_lines = 0;
  for (const Template& t : tmpls) {
    Rn.max_rec_entries = std::max(Rn.max_rec_entries, t.rec_entries);
    Rn.max_coord_nodes = std::max(Rn.max_coord_nodes, t.coord_nodes);
    Rn.max_img_units = std::max(Rn.max_img_units, t.img_units);
    Rn.max_lines = std::max(Rn.max_lines, t.n_lines);
  }
  if (Rn.slot >= 0) { runs_slot_release(Rn.slot); Rn.slot = -1; }
  FES_TRY(runs_slot_acquire(blob.data(), blob.size(), stream, &Rn.slot));
  if (Rn.slot < 0) {  // every template slot is held by a live plan: this one keeps the tiled kernel
    Rn.tiles.release();
    return FES_OK;
  }
  Rn.n_warps = n_warps;
  if (!Rn.side) {
    FES_CUDA(cudaStreamCreateWithFlags(&Rn.side, cudaStreamNonBlocking));
    FES_CUDA(cudaEventCreateWithFlags(&Rn.ev_fork, cudaEventDisableTiming));
    FES_CUDA(cudaEventCreateWithFlags(&Rn.ev_join, cudaEventDisableTiming));
  }
  int64_t prev = 0;
  for (size_t i = 0; i + 1 < covered.size(); i += 2) {
    if (covered[i] > prev) { rest_ranges.push_back((int32_t)prev); rest_ranges.push_back(covered[i]); }
    prev = covered[i + 1];
  }
  if (prev < nn) { rest_ranges.push_back((int32_t)prev); rest_ranges.push_back((int32_t)nn); }
  Rn.active = true;
  if (getenv("FES_PLAN_DEBUG"))
    fprintf(stderr, "[fes runs] N=%d c=%d runs=%zu templates=%zu (%zu of %d words) tiles=%lld nodes covered=%lld of %lld, rest ranges=%zu\n",
            N, c, runs.size(), tmpls.size(), blob.size(), fes_runs::kMaxTmplWords, (long long)nt, (long long)n_cov, (long long)nn,
            rest_ranges.size() / 2);
  return FES_OK;
}

}  // namespace fes
