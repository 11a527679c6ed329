// engine.cu -- engine object and the extern "C" boundary declared in include/lmc_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../../include/lmc_b200.h"
#include "engine.h"
#include "kernels.cuh"
#include "kmc_kernels.cuh"
#include "kmc_team_kernels.cuh"
#include "cmc_kernels.cuh"
#include "cmc_grid_kernels.cuh"
#include "cmc_domain.h"
#include "grouping.h"
#include "tables.h"

namespace lmc {

thread_local std::string g_last_error;

struct StatusError : std::runtime_error {
  int code;
  StatusError(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

#define LMC_CUDA(call)                                                                                     \
  do {                                                                                                     \
    cudaError_t err__ = (call);                                                                            \
    if (err__ != cudaSuccess)                                                                              \
      throw StatusError(LMC_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(err__));       \
  } while (0)

int guard(const std::function<void()> &fn) {
  try {
    fn();
    return LMC_OK;
  } catch (const StatusError &e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::invalid_argument &e) {
    g_last_error = e.what();
    return LMC_ERR_INVALID_ARGUMENT;
  } catch (const std::out_of_range &e) {
    g_last_error = e.what();
    return LMC_ERR_OUT_OF_RANGE;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return LMC_ERR_RUNTIME;
  }
}

// ------------------------------------------------------------------------------------------------ Engine
Engine::Engine(const int32_t factors[3], int32_t id_order, const int32_t *element_set, int32_t n_elements, int32_t solvent,
               int32_t n_walkers_, int32_t device_)
    : n_walkers(n_walkers_), device(device_) {
  for (int d = 0; d < 3; ++d)
    if (factors[d] < 4) throw std::invalid_argument("supercell factors must be >= 4 (the reference's cell list needs 3 cells of 5.3 A per axis)");
  if (id_order != LMC_ID_ORDER_GENERATE && id_order != LMC_ID_ORDER_REASSIGNED) throw std::invalid_argument("unknown id order");
  if (n_walkers < 1) throw std::invalid_argument("n_walkers must be >= 1");
  lat = make_lattice(factors[0], factors[1], factors[2], id_order);
  species = make_species(element_set, n_elements, solvent);
  const Geometry &g = geometry();
  if (g.pair_first_pos != kFirstPos || g.pair_second_pos != kSecondPos || g.site_centre_pos != kCentrePos)
    throw std::logic_error("ordered-neighbourhood structural constants changed");
  for (int r = 0; r < kSiteRows; ++r)          // the compile-time row table of gather_site_rows against the derived site list
    for (int k = 0; k < site_row_count(r); ++k) {
      const int dz = site_row_kind(r) == 0 ? 2 * (k - 1) : (site_row_kind(r) == 1 ? 0 : 2 * k - 1);
      const Int3 o = g.site_offsets[site_row_pos(r) + k];
      if (o.x != site_row_dx(r) || o.y != site_row_dy(r) || o.z != dz) throw std::logic_error("site row table does not match the 43-site list");
    }
  if (device >= 0) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count)
      throw StatusError(LMC_ERR_NO_DEVICE, "CUDA device " + std::to_string(device) + " not available");
    LMC_CUDA(cudaSetDevice(device));
    LMC_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    LMC_CUDA(cudaMalloc(&d_occ, static_cast<size_t>(n_walkers) * lat.padded_size + 16));   // +16: the KMC box scan reads aligned words
    LMC_CUDA(cudaMalloc(&d_error, sizeof(int)));
    LMC_CUDA(cudaEventCreate(&ev_begin));
    LMC_CUDA(cudaEventCreate(&ev_end));
    LMC_CUDA(cudaMemsetAsync(d_error, 0, sizeof(int), stream));
    upload_geometry_tables();
  }
}

Engine::~Engine() {
  if (device >= 0) {
    cudaSetDevice(device);
    for (int r = 0; r < 8; ++r)
      if (cmc_peer_xchg[r] && cmc_peer_xchg[r] != d_cmc_xchg) cudaIpcCloseMemHandle(cmc_peer_xchg[r]);
    for (int r = 0; r < 8; ++r) {
      if (r == dom_rank) continue;
      for (int b = 0; b < 2; ++b)
        if (dom_peer_occ[b][r]) cudaIpcCloseMemHandle(dom_peer_occ[b][r]);
      if (dom_peer_lines[r]) cudaIpcCloseMemHandle(dom_peer_lines[r]);
    }
    for (void *p : device_allocs) cudaFree(p);
    if (d_occ_buf[1]) { cudaFree(d_occ_buf[0]); cudaFree(d_occ_buf[1]); }
    else cudaFree(d_occ);
    cudaFree(d_error);
    cudaFree(d_scratch);
    cudaFree(d_group);
    if (h_group_local) cudaFreeHost(h_group_local);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (stream) cudaStreamDestroy(stream);
  }
}

// device attributes are queried once (cudaDevAttrClockRate alone costs about a millisecond per call)
int Engine::device_attr(int attr) {
  auto it = attr_cache.find(attr);
  if (it != attr_cache.end()) return it->second;
  int v = 0;
  LMC_CUDA(cudaDeviceGetAttribute(&v, static_cast<cudaDeviceAttr>(attr), device));
  attr_cache[attr] = v;
  return v;
}

void Engine::time_begin() { cudaEventRecord(ev_begin, stream); }
void Engine::time_end() {
  cudaEventRecord(ev_end, stream);
  ++launch_count;
  timing_pending = true;
}
double Engine::last_kernel_ms() {
  if (!timing_pending) return last_ms;
  require_device();
  LMC_CUDA(cudaEventSynchronize(ev_end));
  float ms = 0.f;
  LMC_CUDA(cudaEventElapsedTime(&ms, ev_begin, ev_end));
  last_ms = ms;
  timing_pending = false;
  return last_ms;
}

void Engine::require_device() const {
  if (device < 0) throw StatusError(LMC_ERR_NO_DEVICE, "compute call on a host-only engine: there is no CPU fallback");
  cudaSetDevice(device);
}
void Engine::require_coefficients() const {
  if (!has_coefficients) throw std::invalid_argument("coefficients not loaded (lmc_engine_load_coefficients)");
}

template <class T>
const T *Engine::to_device(const std::vector<T> &v) {
  void *p = nullptr;
  LMC_CUDA(cudaMalloc(&p, std::max<size_t>(v.size() * sizeof(T), 16)));
  LMC_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));   // v may be a temporary
  device_allocs.push_back(p);
  return static_cast<const T *>(p);
}

void *Engine::scratch(size_t bytes) {
  if (bytes > scratch_bytes) {
    LMC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_scratch);
    if (h_pinned) cudaFreeHost(h_pinned);
    scratch_bytes = std::max(bytes, scratch_bytes * 2);
    LMC_CUDA(cudaMalloc(&d_scratch, scratch_bytes));
    LMC_CUDA(cudaMallocHost(&h_pinned, scratch_bytes));
  }
  return d_scratch;
}

// Event order of the 12 jumps of a vacancy (KineticMcFirstOmp.cpp:52-68: ascending lattice id of the neighbour).  It
// depends on the vacancy site only through, per axis, whether a neighbour wraps around the period (coordinate 0 or
// period - 1) and the coordinate's parity: 4 classes per axis.  slot[class][k] = rank of jump k's neighbour id, ranked
// once on a representative site of each class (periods are even and >= 8: factors >= 4).  Host only.
std::vector<uint8_t> Engine::kmc_event_order_table() const {
  const Geometry &g = geometry();
  std::vector<uint8_t> slot(64 * 12, 0);
  const int period[3] = {2 * lat.fx, 2 * lat.fy, 2 * lat.fz};
  auto representative = [&](int c, int axis) { return c == 0 ? 0 : (c == 3 ? period[axis] - 1 : 1 + c); };
  for (int cls = 0; cls < 64; ++cls) {
    const int R[3] = {representative(cls >> 4, 0), representative((cls >> 2) & 3, 1), representative(cls & 3, 2)};
    if ((R[0] + R[1] + R[2]) & 1) continue;                  // classes of the other parity hold no site
    int64_t id[12];
    for (int k = 0; k < 12; ++k)
      id[k] = lat.id_of_coords(wrap_coord(R[0] + g.nn1[k].x, period[0]), wrap_coord(R[1] + g.nn1[k].y, period[1]), wrap_coord(R[2] + g.nn1[k].z, period[2]));
    for (int k = 0; k < 12; ++k) {
      int rank = 0;
      for (int o2 = 0; o2 < 12; ++o2) rank += id[o2] < id[k];
      slot[cls * 12 + k] = static_cast<uint8_t>(rank);
    }
  }
  return slot;
}

// The 12 first neighbours of `site` in the order the KMC kernels file their events: neighbour of direction k goes to slot
// table[class(site)][k].  Must equal the ascending-id list of Config::GetFirstNeighborsAdjacencyList for every site.
void Engine::kmc_event_order(int64_t site, int64_t *neighbours_in_slot_order) const {
  if (site < 0 || site >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  const Geometry &g = geometry();
  const std::vector<uint8_t> table = kmc_event_order_table();
  const int period[3] = {2 * lat.fx, 2 * lat.fy, 2 * lat.fz};
  int X, Y, Z;
  lat.coords_of_id(site, X, Y, Z);
  auto cls1 = [](int v, int p) { return v == 0 ? 0 : (v == p - 1 ? 3 : 1 + (v & 1)); };      // coord_class of the kernels
  const int cls = (cls1(X, period[0]) * 4 + cls1(Y, period[1])) * 4 + cls1(Z, period[2]);
  for (int k = 0; k < 12; ++k)
    neighbours_in_slot_order[table[cls * 12 + k]] =
        lat.id_of_coords(wrap_coord(X + g.nn1[k].x, period[0]), wrap_coord(Y + g.nn1[k].y, period[1]), wrap_coord(Z + g.nn1[k].z, period[2]));
}

void Engine::upload_geometry_tables() {
  const Geometry &g = geometry();
  std::vector<int32_t> pair_delta(24 * kPairDeltaStride, 0), site_delta(2 * 43);
  for (int k = 0; k < 12; ++k)
    for (int zp = 0; zp < 2; ++zp)
      for (int t = 0; t < 60; ++t) {
        const Int3 o = g.pair_offsets[k][0][t];   // values are frame independent: the fast kernels use flag 0 throughout
        pair_delta[(k * 2 + zp) * kPairDeltaStride + t] = lat.padded_delta(o.x, o.y, o.z, zp);
      }
  for (int zp = 0; zp < 2; ++zp)
    for (int t = 0; t < 43; ++t) site_delta[zp * 43 + t] = lat.padded_delta(g.site_offsets[t].x, g.site_offsets[t].y, g.site_offsets[t].z, zp);
  std::vector<int8_t> dir_lut(27, -1), nn1(48, 0), frame_p(48, 0), pair_off(12 * 2 * 60 * 4, 0), site_off(43 * 4, 0);
  for (int k = 0; k < 12; ++k) {
    const Int3 d = g.nn1[k], p = g.frame_p[k];
    dir_lut[(d.x + 1) * 9 + (d.y + 1) * 3 + (d.z + 1)] = static_cast<int8_t>(k);
    nn1[4 * k] = d.x; nn1[4 * k + 1] = d.y; nn1[4 * k + 2] = d.z;
    frame_p[4 * k] = p.x; frame_p[4 * k + 1] = p.y; frame_p[4 * k + 2] = p.z;
    for (int s = 0; s < 2; ++s)
      for (int t = 0; t < 60; ++t) {
        const Int3 o = g.pair_offsets[k][s][t];
        int8_t *dst = &pair_off[((k * 2 + s) * 60 + t) * 4];
        dst[0] = o.x; dst[1] = o.y; dst[2] = o.z;
      }
  }
  for (int t = 0; t < 43; ++t) {
    site_off[4 * t] = g.site_offsets[t].x; site_off[4 * t + 1] = g.site_offsets[t].y; site_off[4 * t + 2] = g.site_offsets[t].z;
  }
  // KMC box tables: the (dx, dy) rows of the 7 x 7 box that hold at least one neighbourhood site of some jump
  auto cell_of = [&](int zp, int dx, int dy, int slot, int &dz) {
    const int dzi = slot + (zp == 0 ? -2 : -1);
    dz = 2 * dzi - zp + ((zp + dx + dy) & 1);
    return (dx * lat.ny + dy) * lat.nz + dzi;
  };
  auto envpos_of = [&](int k, int dx, int dy, int dz) -> int {
    for (int t = 0; t < 60; ++t)
      if (g.pair_offsets[k][0][t] == Int3{dx, dy, dz}) return g.env_of_state[t] >= 0 ? g.env_of_state[t] : (g.env_of_state[t] == -1 ? 58 : 59);
    return -1;
  };
  std::vector<std::pair<int, int>> rows;
  for (int dx = -3; dx <= 3; ++dx)
    for (int dy = -3; dy <= 3; ++dy) {
      bool used = false;
      for (int zp = 0; zp < 2; ++zp)
        for (int slot = 0; slot < 4; ++slot) {
          int dz;
          cell_of(zp, dx, dy, slot, dz);
          for (int k = 0; k < 12; ++k) used = used || envpos_of(k, dx, dy, dz) >= 0;
        }
      if (used) rows.emplace_back(dx, dy);
    }
  if (static_cast<int>(rows.size()) > kBoxRows) throw std::logic_error("KMC box has more rows than the scan covers");
  if (rows.at(kBoxCentreRow) != std::make_pair(0, 0)) throw std::logic_error("KMC box centre row is not where the kernels expect it");
  while (static_cast<int>(rows.size()) < kBoxRows) rows.emplace_back(3, 3);      // padding rows: scanned, never mapped (no site of any jump)
  std::vector<int32_t> box_delta(2 * kBoxCells, 0);
  std::vector<int8_t> box_envpos(2 * 12 * kBoxCells, -1);
  for (int zp = 0; zp < 2; ++zp) {
    for (int c = 0; c < kBoxCells; ++c) {
      const int slot = c % 4, dx = rows[static_cast<size_t>(c / 4)].first, dy = rows[static_cast<size_t>(c / 4)].second;
      int dz;
      box_delta[zp * kBoxCells + c] = cell_of(zp, dx, dy, slot, dz);
      if (lat.padded_delta(dx, dy, dz, zp) != box_delta[zp * kBoxCells + c]) throw std::logic_error("box cell geometry is inconsistent");
      for (int k = 0; k < 12; ++k) box_envpos[(zp * 12 + k) * kBoxCells + c] = static_cast<int8_t>(envpos_of(k, dx, dy, dz));
    }
    for (int k = 0; k < 12; ++k) {   // every neighbourhood site must be inside the box, exactly once
      int found = 0;
      for (int c = 0; c < kBoxCells; ++c) found += box_envpos[(zp * 12 + k) * kBoxCells + c] >= 0;
      if (found != 60) throw std::logic_error("KMC box does not cover a jump neighbourhood");
    }
  }
  tab.box_delta = to_device(box_delta);
  tab.box_envpos = to_device(box_envpos);
  tab.pair_delta = to_device(pair_delta);
  tab.site_delta = to_device(site_delta);
  tab.dir_lut = to_device(dir_lut);
  tab.nn1 = to_device(nn1);
  tab.kmc_slot = to_device(kmc_event_order_table());
  tab.frame_p = to_device(frame_p);
  tab.pair_off = to_device(pair_off);
  tab.site_off = to_device(site_off);
  tab.pair_first_pos = g.pair_first_pos;
  tab.pair_second_pos = g.pair_second_pos;
  tab.site_centre_pos = g.site_centre_pos;
  tab.n_species = species.n;
  tab.solvent = species.solvent;
  std::vector<uint64_t> pmask(g.env_pair_mask_hi.begin(), g.env_pair_mask_hi.end()), smask(g.site_pair_mask_hi.begin(), g.site_pair_mask_hi.end());
  std::vector<uint16_t> pbase(g.env_pair_base.begin(), g.env_pair_base.end()), sbase(g.site_pair_base.begin(), g.site_pair_base.end());
  tab.pair_mask_hi = to_device(pmask);
  tab.pair_base = to_device(pbase);
  tab.site_mask_hi = to_device(smask);
  tab.site_base = to_device(sbase);
  tab.n_pair_pairs = static_cast<int32_t>(g.env_pairs.size());
  tab.n_site_pairs = static_cast<int32_t>(g.site_env_pairs.size());
  // total-energy walk in 43-list positions
  auto site_pos = [&](Int3 o) {
    for (int t = 0; t < 43; ++t)
      if (g.site_offsets[t] == o) return t;
    throw std::logic_error("offset outside the site neighbourhood");
  };
  std::vector<uint8_t> e_walk, e_shell;
  for (const auto &w : g.energy_triplets) {
    e_walk.push_back(static_cast<uint8_t>(site_pos(w.o2)));
    e_walk.push_back(static_cast<uint8_t>(site_pos(w.o3)));
    e_walk.push_back(static_cast<uint8_t>(w.label));
    e_walk.push_back(0);
  }
  for (int t = 0; t < 43; ++t) {
    if (t == g.site_centre_pos) continue;
    e_shell.push_back(static_cast<uint8_t>(t));
    e_shell.push_back(static_cast<uint8_t>(g.site_env_shell[t - (t > g.site_centre_pos)]));
  }
  tab.e_walk = to_device(e_walk);
  tab.e_shell_pos = to_device(e_shell);
  tab.n_e_walk = static_cast<int32_t>(g.energy_triplets.size());
  // debug-tap mappings
  const auto types = cluster_types(species);
  const TypeLut lut = make_type_lut(species, types);
  tab.type_lut = to_device(lut.lut);
  tab.n_types = static_cast<int32_t>(types.size());
  auto flat_state = [&](const std::vector<Cluster> &cl) {
    std::vector<int8_t> out;
    for (const auto &c : cl) {
      out.push_back(c.label);
      for (int q = 0; q < 3; ++q) out.push_back(q < c.arity ? static_cast<int8_t>(c.pos[q]) : -1);
    }
    return out;
  };
  tab.map_state_pair = to_device(flat_state(g.state_pair));
  tab.n_state_pair = static_cast<int32_t>(g.state_pair.size());
  tab.map_state_site = to_device(flat_state(g.state_site));
  tab.n_state_site = static_cast<int32_t>(g.state_site.size());
  int len_mmm = 0, len_mm2 = 0;
  const auto gm = group_layout(g, false, species.n, &len_mmm);
  const auto g2 = group_layout(g, true, species.n, &len_mm2);
  auto flat_avg = [&](const std::vector<Cluster> &cl, const std::vector<GroupInfo> &groups) {
    std::vector<int16_t> out;
    for (const auto &c : cl) {
      out.push_back(static_cast<int16_t>(groups[c.group].offset));
      out.push_back(c.pos[0]);
      out.push_back(c.arity == 2 ? c.pos[1] : static_cast<int16_t>(-1));
      out.push_back(c.symmetric ? 1 : 0);
    }
    return out;
  };
  tab.map_mmm = to_device(flat_avg(g.mmm, gm));
  tab.map_mm2 = to_device(flat_avg(g.mm2, g2));
  tab.n_avg_clusters = static_cast<int32_t>(g.mmm.size());
  tab.len_mmm = len_mmm;
  tab.len_mm2 = len_mm2;
  std::vector<int8_t> env_of_list(4 * 58), spe(58);
  for (int a = 0; a < 58; ++a) {
    env_of_list[a] = static_cast<int8_t>(g.env_of_mmm[a]);
    env_of_list[58 + a] = static_cast<int8_t>(g.env_of_mm2[a]);
    env_of_list[2 * 58 + a] = static_cast<int8_t>(g.env_of_mm2_backward[0][a]);
    env_of_list[3 * 58 + a] = static_cast<int8_t>(g.env_of_mm2_backward[1][a]);
  }
  for (int t = 0; t < 60; ++t)
    if (g.env_of_state[t] >= 0) spe[g.env_of_state[t]] = static_cast<int8_t>(t);
  tab.env_of_list = to_device(env_of_list);
  tab.state_pos_of_env = to_device(spe);
  std::vector<int8_t> coe(species.code_of_enum.begin(), species.code_of_enum.end());
  std::vector<uint8_t> eoc(species.enum_of_code.begin(), species.enum_of_code.end());
  d_code_of_enum = to_device(coe);
  d_enum_of_code = to_device(eoc);
}

void Engine::load_coefficients(const std::string &json_path, int model) {
  load_coefficients(parse_coefficients_json(json_path), model);
  coefficients_path = json_path;
}

void Engine::load_coefficients(const Coefficients &co, int model) {
  coefficients = co;
  pair_tables = build_pair_tables(species, coefficients, model);
  tab.barrier_model = model;
  site_tables = build_site_tables(species, coefficients);
  energy_tables = build_energy_tables(species, coefficients);
  has_coefficients = true;
  // [..][3] = (dE, logD, logKs)  ->  [..][2] = (dE, logKs + 2 logD) for the KMC kernels.  The folded tables are put on a
  // common binary grid: every entry is rounded to a multiple of 2^-q, q per component chosen so that the largest
  // possible sum |C| + sum_t max|A_t| + sum_pairs max|B| stays below 2^(51-q).  Then every partial sum any kernel can
  // form is a multiple of 2^-q below 2^52 of them, i.e. EXACT in double: the contracted sums do not depend on the
  // order of the additions, and the launch shapes of the KMC driver (half-warp per walker: sequential; block per walker:
  // tree over lanes) deliver bit-identical (dE, log E0).  Cost: <= 2^-(q+1) per entry (q = 44..48 for eV-sized
  // coefficients: ~1e-14 eV), against a parity tolerance of 1e-9 eV.
  auto fold = [](const std::vector<double> &v) {
    std::vector<double> out(v.size() / 3 * 2);
    for (size_t i = 0; i < v.size() / 3; ++i) {
      out[2 * i] = v[3 * i];
      out[2 * i + 1] = v[3 * i + 2] + 2.0 * v[3 * i + 1];
    }
    return out;
  };
  std::vector<double> &C2 = folded_C, &A2 = folded_A, &B2 = folded_B;
  C2 = fold(pair_tables.C); A2 = fold(pair_tables.A); B2 = fold(pair_tables.B);
  {
    const size_t n = static_cast<size_t>(species.n), n_pairs = geometry().env_pairs.size();
    for (int c = 0; c < 2; ++c) {
      double bound = 0.0;
      for (size_t m = 0; m < n; ++m) {
        double b = std::fabs(C2[m * 2 + c]);
        for (size_t t = 0; t < static_cast<size_t>(kEnvN); ++t) {
          double mx = 0.0;
          for (size_t e = 0; e < n; ++e) mx = std::max(mx, std::fabs(A2[((m * kEnvN + t) * n + e) * 2 + c]));
          b += mx;
        }
        for (size_t pq = 0; pq < n_pairs; ++pq) {
          double mx = 0.0;
          for (size_t e = 0; e < n * n; ++e) mx = std::max(mx, std::fabs(B2[((m * n_pairs + pq) * n * n + e) * 2 + c]));
          b += mx;
        }
        bound = std::max(bound, b);
      }
      int exp2 = 0;
      std::frexp(std::max(bound, 1e-300), &exp2);            // bound < 2^exp2
      const int q = std::min(60, 51 - exp2);
      pair_grid_bits[c] = q;
      const double up = std::ldexp(1.0, q), down = std::ldexp(1.0, -q);
      for (std::vector<double> *tabv : {&C2, &A2, &B2})
        for (size_t i = static_cast<size_t>(c); i < tabv->size(); i += 2) (*tabv)[i] = std::nearbyint((*tabv)[i] * up) * down;
    }
  }
  if (device >= 0) {
    cudaSetDevice(device);
    tab.pair_C2 = to_device(folded_C);
    tab.pair_A2 = to_device(folded_A);
    tab.pair_B2 = to_device(folded_B);
    tab.pair_C = to_device(pair_tables.C);
    tab.pair_A = to_device(pair_tables.A);
    tab.pair_B = to_device(pair_tables.B);
    tab.site_C = to_device(site_tables.C);
    tab.site_A = to_device(site_tables.A);
    tab.site_B = to_device(site_tables.B);
    tab.e_single = to_device(energy_tables.single);
    tab.e_pair = to_device(energy_tables.pair);
    tab.e_triplet = to_device(energy_tables.triplet);
  }
}

void Engine::check_event_errors(const char *what) {
  int err = 0;
  LMC_CUDA(cudaMemcpyAsync(&err, d_error, sizeof(int), cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  if (!err) return;
  LMC_CUDA(cudaMemsetAsync(d_error, 0, sizeof(int), stream));
  std::string msg = std::string(what) + ":";
  if (err & kErrBadSite) msg += " lattice id or element code out of range;";
  if (err & kErrNotNeighbour) msg += " Neighbor not found for lattice pair (not first neighbours);";
  if (err & kErrNotVacancy) msg += " first site of a jump pair must hold the vacancy and the second an atom;";
  if (err & kErrExtraVacancy) msg += " Cluster not found in ClusterIndexer (two vacancies within interaction range);";
  throw StatusError((err & kErrBadSite) ? LMC_ERR_INVALID_ARGUMENT : LMC_ERR_OUT_OF_RANGE, msg);
}

void Engine::set_occupancy(int32_t walker, const uint8_t *occ, int64_t n, int32_t count) {
  require_device();
  if (walker < 0 || count < 1 || walker + count > n_walkers) throw std::invalid_argument("walker index out of range");
  if (n != lat.num_sites * count) throw std::invalid_argument("occupancy length must be num_sites per walker");
  cmc_ready = false;                       // the CMC mirror / cell arrays are rebuilt by the next lmc_cmc_reset (or run)
  kmc_ready = false;                       // vacancy position, concentrations and clocks belong to the old configuration
  uint8_t *d_in = static_cast<uint8_t *>(scratch(static_cast<size_t>(n)));
  LMC_CUDA(cudaMemcpyAsync(d_in, occ, static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((lat.padded_size + threads - 1) / threads);
  for (int w0 = 0; w0 < count; w0 += 32768) {   // gridDim.y limit is 65535
    const unsigned ny = static_cast<unsigned>(std::min(32768, count - w0));
    upload_occupancy_kernel<<<dim3(blocks, ny), threads, 0, stream>>>(lat, d_in + static_cast<int64_t>(w0) * lat.num_sites,
                                                                     d_occ + static_cast<int64_t>(walker + w0) * lat.padded_size,
                                                                     d_code_of_enum, d_error);
    ++launch_count;
  }
  LMC_CUDA(cudaGetLastError());
  check_event_errors("set_occupancy (element not in element_set)");
}

void Engine::get_occupancy(int32_t walker, uint8_t *occ, int64_t n, int32_t count) {
  require_device();
  if (walker < 0 || count < 1 || walker + count > n_walkers) throw std::invalid_argument("walker index out of range");
  if (n != lat.num_sites * count) throw std::invalid_argument("occupancy length must be num_sites per walker");
  uint8_t *d_out = static_cast<uint8_t *>(scratch(static_cast<size_t>(n)));
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((lat.num_sites + threads - 1) / threads);
  for (int w0 = 0; w0 < count; w0 += 32768) {
    const unsigned ny = static_cast<unsigned>(std::min(32768, count - w0));
    download_occupancy_kernel<<<dim3(blocks, ny), threads, 0, stream>>>(lat, d_occ + static_cast<int64_t>(walker + w0) * lat.padded_size,
                                                                       d_out + static_cast<int64_t>(w0) * lat.num_sites, d_enum_of_code);
    ++launch_count;
  }
  LMC_CUDA(cudaGetLastError());
  LMC_CUDA(cudaMemcpyAsync(occ, d_out, static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
}

void Engine::get_elements(int32_t walker, int64_t n, const int64_t *sites, uint8_t *out) {
  require_device();
  if (walker < 0 || walker >= n_walkers) throw std::invalid_argument("walker index out of range");
  if (n <= 0) return;
  if (!sites || !out) throw std::invalid_argument("null buffer");
  char *d = static_cast<char *>(scratch(static_cast<size_t>(n) * 9 + 64));
  int64_t *d_sites = reinterpret_cast<int64_t *>(d);
  uint8_t *d_out = reinterpret_cast<uint8_t *>(d_sites + n);
  LMC_CUDA(cudaMemcpyAsync(d_sites, sites, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
  gather_sites_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(lat, d_occ + static_cast<int64_t>(walker) * lat.padded_size, d_sites, n,
                                                                                 d_enum_of_code, d_out, d_error);
  ++launch_count;
  LMC_CUDA(cudaGetLastError());
  LMC_CUDA(cudaMemcpyAsync(out, d_out, static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream));
  check_event_errors("lmc_engine_get_elements");
}

int64_t Engine::find_element(int32_t walker, int32_t element, int64_t *count) {
  require_device();
  if (walker < 0 || walker >= n_walkers) throw std::invalid_argument("walker index out of range");
  const int code = element >= 0 && element < 16 ? species.code_of_enum[static_cast<size_t>(element)] : -1;
  if (code < 0) throw std::invalid_argument("element is not in the element set");
  unsigned long long *d = static_cast<unsigned long long *>(scratch(64));
  const unsigned long long init[2] = {~0ULL, 0ULL};
  LMC_CUDA(cudaMemcpyAsync(d, init, 16, cudaMemcpyHostToDevice, stream));
  find_element_kernel<<<static_cast<unsigned>((lat.num_sites + 255) / 256), 256, 0, stream>>>(lat, d_occ + static_cast<int64_t>(walker) * lat.padded_size, code, d, d + 1);
  ++launch_count;
  LMC_CUDA(cudaGetLastError());
  unsigned long long res[2];
  LMC_CUDA(cudaMemcpyAsync(res, d, 16, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  if (count) *count = static_cast<int64_t>(res[1]);
  return res[1] ? static_cast<int64_t>(res[0]) : -1;
}

void Engine::lattice_jump(int32_t walker, int64_t a, int64_t b) {
  require_device();
  if (walker < 0 || walker >= n_walkers) throw std::invalid_argument("walker index out of range");
  if (a < 0 || b < 0 || a >= lat.num_sites || b >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  cmc_ready = false;
  kmc_ready = false;                       // the vacancy may have moved: the next KMC run locates it again (clocks restart)
  lattice_jump_kernel<<<1, 32, 0, stream>>>(lat, d_occ + static_cast<int64_t>(walker) * lat.padded_size, a, b);
  LMC_CUDA(cudaGetLastError());
  LMC_CUDA(cudaStreamSynchronize(stream));
}

// Batched evaluation requests in arbitrary order make every lane of a warp gather from its own cache lines.  For large
// batches whose order is not already local (a sample of the first requests decides), the requests are grouped on the device
// by the lattice region of their first site (grouping.cu) and the kernels run over that permutation; results stay in the
// caller's order.  LMC_GROUP_REQUESTS = 0 / 1 forces the choice, LMC_GROUP_MIN sets the smallest batch considered.
const uint32_t *Engine::group_if_scattered(int64_t n, const int32_t *walker, const int64_t *site) {
  last_grouped = false;
  const char *force = std::getenv("LMC_GROUP_REQUESTS");
  const int64_t min_n = std::getenv("LMC_GROUP_MIN") ? std::atoll(std::getenv("LMC_GROUP_MIN")) : (1LL << 18);
  if (n < 2 || n >= (1LL << 31) || (force && force[0] == '0') || (!force && n < min_n)) return nullptr;
  const GroupPlan plan = group_plan(lat, n_walkers, n);
  if (plan.total_bytes > group_bytes) {
    LMC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_group);
    group_bytes = std::max(plan.total_bytes, group_bytes * 2);
    LMC_CUDA(cudaMalloc(&d_group, group_bytes));
    if (!h_group_local) LMC_CUDA(cudaMallocHost(&h_group_local, sizeof(unsigned int)));
  }
  if (!force) {
    unsigned int *d_local = reinterpret_cast<unsigned int *>(static_cast<char *>(d_group) + group_bytes - 64);
    group_sample_locality(lat, plan, n, walker, site, d_local, stream);
    LMC_CUDA(cudaMemcpyAsync(h_group_local, d_local, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    LMC_CUDA(cudaStreamSynchronize(stream));
    const int64_t sample = std::min<int64_t>(n, kGroupSample);
    if (2 * static_cast<int64_t>(*h_group_local) >= sample) return nullptr;          // the caller's order is local already
  }
  const bool dbg = std::getenv("LMC_DEBUG_TIMING") != nullptr;
  cudaEvent_t g0 = nullptr, g1 = nullptr;
  if (dbg) { cudaEventCreate(&g0); cudaEventCreate(&g1); cudaEventRecord(g0, stream); }
  const uint32_t *perm = group_requests(lat, plan, n, walker, site, d_group, stream);
  LMC_CUDA(cudaGetLastError());
  if (dbg) {
    cudaEventRecord(g1, stream); cudaEventSynchronize(g1);
    float ms = 0.f; cudaEventElapsedTime(&ms, g0, g1);
    std::fprintf(stderr, "[lmc] grouping %lld requests: %.4f ms (keys + radix sort, shift %d)\n", static_cast<long long>(n), ms, plan.shift);
    cudaEventDestroy(g0); cudaEventDestroy(g1);
  }
  launch_count += 3;       // key kernel + two radix passes
  last_grouped = true;
  return perm;
}

void Engine::eval_barriers_dev(int64_t n, const int32_t *walker, const int64_t *site_i, const int64_t *site_j, double *Ea,
                               double *dE, double *D, double *Ks) {
  require_device();
  require_coefficients();
  if (!pair_tables.has_barrier) throw std::invalid_argument("the coefficient file has no per-element quartic blocks");
  if (n <= 0) return;
  const unsigned blocks = static_cast<unsigned>((n + kBarrierThreads - 1) / kBarrierThreads);
  time_begin();
  const uint32_t *perm = group_if_scattered(n, walker, site_i);
  barrier_kernel<<<blocks, kBarrierThreads, 0, stream>>>(lat, tab, d_occ, lat.padded_size, n, walker, site_i, site_j, Ea, dE, D, Ks, d_error, perm);
  time_end();
  LMC_CUDA(cudaGetLastError());
}

void Engine::eval_barriers(int64_t n, const int32_t *walker, const int64_t *site_i, const int64_t *site_j, double *Ea,
                           double *dE, double *D, double *Ks) {
  require_device();
  if (n <= 0) return;
  // one staging buffer: [i | j | walker] in, [Ea | dE | D | Ks] out
  const size_t in_bytes = static_cast<size_t>(n) * (8 + 8 + 4), out_bytes = static_cast<size_t>(n) * 8 * 4;
  char *d = static_cast<char *>(scratch(in_bytes + out_bytes + 64));
  int64_t *d_i = reinterpret_cast<int64_t *>(d);
  int64_t *d_j = d_i + n;
  double *d_out = reinterpret_cast<double *>(d_j + n);
  int32_t *d_w = reinterpret_cast<int32_t *>(d_out + 4 * n);
  LMC_CUDA(cudaMemcpyAsync(d_i, site_i, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
  LMC_CUDA(cudaMemcpyAsync(d_j, site_j, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
  if (walker) LMC_CUDA(cudaMemcpyAsync(d_w, walker, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, stream));
  eval_barriers_dev(n, walker ? d_w : nullptr, d_i, d_j, d_out, d_out + n, D ? d_out + 2 * n : nullptr, Ks ? d_out + 3 * n : nullptr);
  LMC_CUDA(cudaMemcpyAsync(Ea, d_out, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(dE, d_out + n, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream));
  if (D) LMC_CUDA(cudaMemcpyAsync(D, d_out + 2 * n, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream));
  if (Ks) LMC_CUDA(cudaMemcpyAsync(Ks, d_out + 3 * n, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream));
  check_event_errors("lmc_eval_barriers");
}

void Engine::eval_vacancy_events_dev(int64_t n, const int32_t *walker, const int64_t *vacancy, int64_t *neighbour, double *Ea, double *dE) {
  require_device();
  require_coefficients();
  if (!pair_tables.has_barrier) throw std::invalid_argument("the coefficient file has no per-element quartic blocks");
  if (n <= 0) return;
  if (lat.num_sites >= (1LL << 31)) throw std::invalid_argument("event lists are ordered by 32-bit lattice ids (num_sites < 2^31)");
  const size_t smem = static_cast<size_t>(species.n) * kEnvN * species.n * 2 * sizeof(double);
  LMC_CUDA(cudaFuncSetAttribute(vacancy_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // enough blocks to fill the device, few enough that each stages the tables once for several items
  const int64_t want = (n + kKmcWalkersPerBlock - 1) / kKmcWalkersPerBlock;
  const unsigned blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(device_attr(cudaDevAttrMultiProcessorCount)) * 7)));
  time_begin();
  vacancy_events_kernel<<<blocks, kKmcThreads, smem, stream>>>(lat, tab, d_occ, lat.padded_size, n, walker, vacancy, neighbour, Ea, dE, d_error);
  time_end();
  LMC_CUDA(cudaGetLastError());
}

void Engine::eval_vacancy_events(int64_t n, const int32_t *walker, const int64_t *vacancy, int64_t *neighbour, double *Ea, double *dE) {
  require_device();
  if (n <= 0) return;
  const size_t N = static_cast<size_t>(n);
  char *d = static_cast<char *>(scratch(N * 8 + N * 12 * (8 + 8 + 8) + N * 4 + 64));
  int64_t *d_v = reinterpret_cast<int64_t *>(d);
  int64_t *d_nb = d_v + N;
  double *d_ea = reinterpret_cast<double *>(d_nb + 12 * N), *d_de = d_ea + 12 * N;
  int32_t *d_w = reinterpret_cast<int32_t *>(d_de + 12 * N);
  LMC_CUDA(cudaMemcpyAsync(d_v, vacancy, N * 8, cudaMemcpyHostToDevice, stream));
  if (walker) LMC_CUDA(cudaMemcpyAsync(d_w, walker, N * 4, cudaMemcpyHostToDevice, stream));
  eval_vacancy_events_dev(n, walker ? d_w : nullptr, d_v, d_nb, d_ea, d_de);
  if (neighbour) LMC_CUDA(cudaMemcpyAsync(neighbour, d_nb, N * 12 * 8, cudaMemcpyDeviceToHost, stream));
  if (Ea) LMC_CUDA(cudaMemcpyAsync(Ea, d_ea, N * 12 * 8, cudaMemcpyDeviceToHost, stream));
  if (dE) LMC_CUDA(cudaMemcpyAsync(dE, d_de, N * 12 * 8, cudaMemcpyDeviceToHost, stream));
  check_event_errors("lmc_eval_vacancy_events");
}

void Engine::eval_swap_de_dev(int64_t n, const int32_t *walker, const int64_t *a, const int64_t *b, double *dE, bool first_neighbours_only) {
  require_device();
  require_coefficients();
  if (n <= 0) return;
  const int64_t want = (n + kSwapThreads - 1) / kSwapThreads;
  time_begin();
  static const bool force_general = std::getenv("LMC_SWAP_GENERAL_KERNEL") != nullptr;   // A/B switch for tests and profiling
  // grouping by the region of site a only on request (LMC_GROUP_REQUESTS=1): measured on B200 with 4.2M random pairs it saves
  // 0.13 ms of kernel time at 40^3 (0.58 -> 0.46 ms with the a side coalesced; the b side stays scattered) and costs 0.15 ms
  const char *group_env = std::getenv("LMC_GROUP_REQUESTS");
  const uint32_t *perm = (species.n + 1 <= kSwapMaxM && !force_general && group_env && group_env[0] == '1') ? group_if_scattered(n, walker, a) : nullptr;
  if (species.n + 1 <= kSwapMaxM && !force_general) {            // persistent blocks: the walk tables are staged in shared memory once per block
    static const int occ_variant = std::getenv("LMC_SWAP_OCC") ? std::atoi(std::getenv("LMC_SWAP_OCC")) : 6;     // tuning switch (6 blocks/SM = 80 registers measured best)
    const size_t smem = swap_rows_smem_bytes(species.n + 1);
    const int per_sm = occ_variant;
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(want, static_cast<int64_t>(device_attr(cudaDevAttrMultiProcessorCount)) * per_sm * 3));
    auto launch = [&](auto kernel) {
      kernel<<<blocks, kSwapThreads, smem, stream>>>(lat, tab, d_occ, lat.padded_size, n, walker, a, b, dE, d_error, first_neighbours_only ? 1 : 0, perm);
    };
    if (occ_variant >= 8) launch(swap_de_rows_kernel<8>);
    else if (occ_variant >= 6) launch(swap_de_rows_kernel<6>);
    else if (occ_variant == 5) launch(swap_de_rows_kernel<5>);
    else launch(swap_de_rows_kernel<4>);
  } else {
    swap_de_kernel<<<static_cast<unsigned>(want), kSwapThreads, 0, stream>>>(lat, tab, d_occ, lat.padded_size, n, walker, a, b, dE, d_error,
                                                                           first_neighbours_only ? 1 : 0);
  }
  time_end();
  LMC_CUDA(cudaGetLastError());
}

void Engine::eval_swap_de(int64_t n, const int32_t *walker, const int64_t *a, const int64_t *b, double *dE, bool first_neighbours_only) {
  require_device();
  if (n <= 0) return;
  char *d = static_cast<char *>(scratch(static_cast<size_t>(n) * (8 + 8 + 8 + 4) + 64));
  int64_t *d_a = reinterpret_cast<int64_t *>(d);
  int64_t *d_b = d_a + n;
  double *d_out = reinterpret_cast<double *>(d_b + n);
  int32_t *d_w = reinterpret_cast<int32_t *>(d_out + n);
  LMC_CUDA(cudaMemcpyAsync(d_a, a, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
  LMC_CUDA(cudaMemcpyAsync(d_b, b, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
  if (walker) LMC_CUDA(cudaMemcpyAsync(d_w, walker, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, stream));
  eval_swap_de_dev(n, walker ? d_w : nullptr, d_a, d_b, d_out, first_neighbours_only);
  LMC_CUDA(cudaMemcpyAsync(dE, d_out, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream));
  check_event_errors(first_neighbours_only ? "lmc_eval_pair_de" : "lmc_eval_swap_de");
}

void Engine::eval_site_de(int64_t n, const int32_t *walker, const int64_t *site, const uint8_t *new_element, double *dE) {
  require_device();
  require_coefficients();
  if (n <= 0) return;
  std::vector<uint8_t> codes(static_cast<size_t>(n));
  for (int64_t k = 0; k < n; ++k) {
    const int c = new_element[k] < 16 ? species.code_of_enum[new_element[k]] : -1;
    if (c < 0) throw std::invalid_argument("new element is not in the element set");
    codes[static_cast<size_t>(k)] = static_cast<uint8_t>(c);
  }
  char *d = static_cast<char *>(scratch(static_cast<size_t>(n) * (8 + 8 + 4 + 1) + 64));
  int64_t *d_s = reinterpret_cast<int64_t *>(d);
  double *d_out = reinterpret_cast<double *>(d_s + n);
  int32_t *d_w = reinterpret_cast<int32_t *>(d_out + n);
  uint8_t *d_c = reinterpret_cast<uint8_t *>(d_w + n);
  LMC_CUDA(cudaMemcpyAsync(d_s, site, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
  LMC_CUDA(cudaMemcpyAsync(d_c, codes.data(), static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
  if (walker) LMC_CUDA(cudaMemcpyAsync(d_w, walker, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, stream));
  const unsigned blocks = static_cast<unsigned>((n + kSwapThreads - 1) / kSwapThreads);
  site_de_kernel<<<blocks, kSwapThreads, 0, stream>>>(lat, tab, d_occ, lat.padded_size, n, walker ? d_w : nullptr, d_s, d_c, d_out, d_error);
  LMC_CUDA(cudaGetLastError());
  LMC_CUDA(cudaMemcpyAsync(dE, d_out, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream));
  check_event_errors("lmc_eval_site_de");
}

double Engine::total_energy(int32_t walker, int64_t *counts, int32_t n_types) { return cluster_energy(walker, nullptr, -1, counts, n_types); }

// EnergyPredictor::GetEnergy (sites == nullptr) / GetEnergyOfCluster (pred/src/EnergyPredictor.cpp:97-184): n listed lattice
// sites; the site set is the listed sites plus their 1-3NN shells
double Engine::cluster_energy(int32_t walker, const int64_t *sites, int64_t n, int64_t *counts, int32_t n_types) {
  require_device();
  require_coefficients();
  if (walker < 0 || walker >= n_walkers) throw std::invalid_argument("walker index out of range");
  if (counts && n_types != tab.n_types) throw std::invalid_argument("counts buffer must hold n_types entries");
  if (sites == nullptr && n >= 0) throw std::invalid_argument("null site list");
  const unsigned blocks = static_cast<unsigned>((lat.num_sites + kEnergyThreads - 1) / kEnergyThreads);
  const int m = species.n + 1;
  const size_t member_bytes = sites ? static_cast<size_t>(lat.padded_size) + 16 : 0, list_bytes = sites ? static_cast<size_t>(std::max<int64_t>(n, 1)) * 8 : 0;
  char *d = static_cast<char *>(scratch(static_cast<size_t>(blocks) * 8 + static_cast<size_t>(tab.n_types) * 8 + 64 + list_bytes + member_bytes));
  double *d_sums = reinterpret_cast<double *>(d);
  unsigned long long *d_counts = reinterpret_cast<unsigned long long *>(d_sums + blocks);
  int64_t *d_sites = reinterpret_cast<int64_t *>(d_counts + tab.n_types + 8);
  uint8_t *d_member = sites ? reinterpret_cast<uint8_t *>(d_sites) + list_bytes : nullptr;
  LMC_CUDA(cudaMemsetAsync(d_counts, 0, static_cast<size_t>(tab.n_types) * 8, stream));
  if (sites) {
    LMC_CUDA(cudaMemsetAsync(d_member, 0, member_bytes, stream));
    if (n > 0) {
      LMC_CUDA(cudaMemcpyAsync(d_sites, sites, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream));
      mark_cluster_sites_kernel<<<static_cast<unsigned>((n * 43 + 127) / 128), 128, 0, stream>>>(lat, tab, d_sites, n, d_member, d_error);
      ++launch_count;
      LMC_CUDA(cudaGetLastError());
    }
  }
  const size_t smem = sizeof(double) * (m + 3 * m * m + 4 * m * m * m) + sizeof(int32_t) * 2 * 43 + sizeof(unsigned) * tab.n_types;
  energy_kernel<<<blocks, kEnergyThreads, smem, stream>>>(lat, tab, d_occ + static_cast<int64_t>(walker) * lat.padded_size, d_sums,
                                                         counts ? d_counts : nullptr, d_member);
  LMC_CUDA(cudaGetLastError());
  if (sites) check_event_errors("lmc_energy_of_cluster");
  std::vector<double> sums(blocks);
  LMC_CUDA(cudaMemcpyAsync(sums.data(), d_sums, static_cast<size_t>(blocks) * 8, cudaMemcpyDeviceToHost, stream));
  if (counts) LMC_CUDA(cudaMemcpyAsync(counts, d_counts, static_cast<size_t>(tab.n_types) * 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  long double total = 0;
  for (double s : sums) total += s;
  const double e = static_cast<double>(total);
  if (e != e) throw StatusError(LMC_ERR_OUT_OF_RANGE, "Cluster not found in ClusterIndexer (two vacancies within interaction range)");
  return e;
}

void Engine::debug_pair(int32_t walker, int64_t i, int64_t j, int64_t *state, int64_t *mmm, int64_t *mm2, int64_t *mm2b,
                        int32_t *sc, int32_t *ec, int32_t *enc_mmm, int32_t *enc_f, int32_t *enc_b) {
  require_device();
  if (walker < 0 || walker >= n_walkers) throw std::invalid_argument("walker index out of range");
  if (i < 0 || j < 0 || i >= lat.num_sites || j >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  const size_t n_lists = 60 + 3 * 58, n_counts = 2 * static_cast<size_t>(tab.n_types), n_enc = static_cast<size_t>(tab.len_mmm) + 2 * tab.len_mm2;
  char *d = static_cast<char *>(scratch(n_lists * 8 + (n_counts + n_enc) * 4 + 64));
  int64_t *d_lists = reinterpret_cast<int64_t *>(d);
  int32_t *d_counts = reinterpret_cast<int32_t *>(d_lists + n_lists);
  int32_t *d_enc = d_counts + n_counts;
  debug_pair_kernel<<<1, 128, 0, stream>>>(lat, tab, d_occ + static_cast<int64_t>(walker) * lat.padded_size, i, j, d_lists, d_counts, d_enc, d_error);
  LMC_CUDA(cudaGetLastError());
  std::vector<int64_t> lists(n_lists);
  std::vector<int32_t> counts(n_counts), enc(n_enc);
  LMC_CUDA(cudaMemcpyAsync(lists.data(), d_lists, n_lists * 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(counts.data(), d_counts, n_counts * 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(enc.data(), d_enc, n_enc * 4, cudaMemcpyDeviceToHost, stream));
  check_event_errors("lmc_debug_pair");
  if (state) std::copy(lists.begin(), lists.begin() + 60, state);
  if (mmm) std::copy(lists.begin() + 60, lists.begin() + 118, mmm);
  if (mm2) std::copy(lists.begin() + 118, lists.begin() + 176, mm2);
  if (mm2b) std::copy(lists.begin() + 176, lists.begin() + 234, mm2b);
  if (sc) std::copy(counts.begin(), counts.begin() + tab.n_types, sc);
  if (ec) std::copy(counts.begin() + tab.n_types, counts.end(), ec);
  if (enc_mmm) std::copy(enc.begin(), enc.begin() + tab.len_mmm, enc_mmm);
  if (enc_f) std::copy(enc.begin() + tab.len_mmm, enc.begin() + tab.len_mmm + tab.len_mm2, enc_f);
  if (enc_b) std::copy(enc.begin() + tab.len_mmm + tab.len_mm2, enc.end(), enc_b);
}

void Engine::debug_site(int32_t walker, int64_t site, int32_t new_element, int64_t *state43, int32_t *sc, int32_t *ec) {
  require_device();
  if (walker < 0 || walker >= n_walkers) throw std::invalid_argument("walker index out of range");
  if (site < 0 || site >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  const int code = (new_element >= 0 && new_element < 16) ? species.code_of_enum[new_element] : -1;
  if (code < 0) throw std::invalid_argument("new element is not in the element set");
  const size_t n_counts = 2 * static_cast<size_t>(tab.n_types);
  char *d = static_cast<char *>(scratch(43 * 8 + n_counts * 4 + 64));
  int64_t *d_list = reinterpret_cast<int64_t *>(d);
  int32_t *d_counts = reinterpret_cast<int32_t *>(d_list + 43);
  debug_site_kernel<<<1, 64, 0, stream>>>(lat, tab, d_occ + static_cast<int64_t>(walker) * lat.padded_size, site, code, d_list, d_counts, d_error);
  LMC_CUDA(cudaGetLastError());
  std::vector<int64_t> list(43);
  std::vector<int32_t> counts(n_counts);
  LMC_CUDA(cudaMemcpyAsync(list.data(), d_list, 43 * 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(counts.data(), d_counts, n_counts * 4, cudaMemcpyDeviceToHost, stream));
  check_event_errors("lmc_debug_site");
  if (state43) std::copy(list.begin(), list.end(), state43);
  if (sc) std::copy(counts.begin(), counts.begin() + tab.n_types, sc);
  if (ec) std::copy(counts.begin() + tab.n_types, counts.end(), ec);
}

// ------------------------------------------------------------------------------------------------ KMC driver
__global__ void kmc_target_kernel(const int64_t *__restrict__ steps, int64_t *__restrict__ target, int n, int64_t n_steps) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n) target[w] = steps[w] + n_steps;
}
namespace {
template <class T>
T *dev_alloc(size_t n) {
  void *p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess)
    throw StatusError(LMC_ERR_CUDA, "cudaMalloc failed");
  return static_cast<T *>(p);
}
}  // namespace

void Engine::kmc_reset() {
  require_device();
  if (!d_kmc_vacancy) {
    const size_t n = static_cast<size_t>(n_walkers);
    d_kmc_vacancy = dev_alloc<int64_t>(n); d_kmc_steps = dev_alloc<int64_t>(n);
    d_kmc_time = dev_alloc<double>(n); d_kmc_energy = dev_alloc<double>(n); d_kmc_temperature = dev_alloc<double>(n);
    d_kmc_cvac = dev_alloc<double>(n); d_kmc_csol = dev_alloc<double>(n); d_kmc_error = dev_alloc<int32_t>(n);
    d_kmc_previous = dev_alloc<int64_t>(n);
    d_kmc_target = dev_alloc<int64_t>(n); d_kmc_done = dev_alloc<int>(2); d_kmc_tail_order = dev_alloc<int32_t>(n);
    device_allocs.push_back(d_kmc_target); device_allocs.push_back(d_kmc_done); device_allocs.push_back(d_kmc_tail_order);
    for (void *p : {static_cast<void *>(d_kmc_vacancy), static_cast<void *>(d_kmc_steps), static_cast<void *>(d_kmc_time),
                    static_cast<void *>(d_kmc_energy), static_cast<void *>(d_kmc_temperature), static_cast<void *>(d_kmc_cvac),
                    static_cast<void *>(d_kmc_csol), static_cast<void *>(d_kmc_error), static_cast<void *>(d_kmc_previous)})
      device_allocs.push_back(p);
  }
  KmcState st{d_kmc_vacancy, d_kmc_time, d_kmc_energy, d_kmc_steps, d_kmc_temperature, d_kmc_cvac, d_kmc_csol, d_kmc_error, d_kmc_previous};
  const int al = species.code_of_enum[1];   // RateCorrector counts "not Al, not X" as solute (KineticMcAbstract.cpp:35)
  kmc_init_kernel<<<static_cast<unsigned>(n_walkers), 256, 0, stream>>>(lat, d_occ, lat.padded_size, st, species.n, al, 1);
  LMC_CUDA(cudaGetLastError());
  std::vector<int32_t> err(static_cast<size_t>(n_walkers));
  LMC_CUDA(cudaMemcpyAsync(err.data(), d_kmc_error, err.size() * 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  for (int w = 0; w < n_walkers; ++w)
    if (err[static_cast<size_t>(w)]) throw std::out_of_range("vacancy not found (walker " + std::to_string(w) + " must hold exactly one vacancy)");
  kmc_ready = true;
}

// The launch shape of a first-order KMC run.  A block per walker with 8 / 16 / 32 lanes per candidate jump is used when
// every walker's block is resident at once (a second wave would double the run time) -- the widest group that fits;
// thousands of walkers go to the half-warp kernel.  LMC_KMC_TEAM_LANES = 0 / 8 / 16 / 32 overrides the choice (A/B runs).
const void *Engine::kmc_team_kernel_choice(bool instrumented, size_t table_smem, int *lanes_out, size_t *smem_out, bool for_tail) {
  // for_tail: the 8-lane instantiation for the tail of a hybrid launch (any number of blocks; most leave at once)
  const int forced = for_tail ? 8 : (std::getenv("LMC_KMC_TEAM_LANES") ? std::atoi(std::getenv("LMC_KMC_TEAM_LANES")) : -1);
  const int max_per_sm = std::getenv("LMC_KMC_TEAM_WALKERS_PER_SM") ? std::atoi(std::getenv("LMC_KMC_TEAM_WALKERS_PER_SM")) : 7;
  if (forced == 0) return nullptr;
  // small cells: the walker's occupancy resident in shared memory (LMC_KMC_TEAM_SMEM=0 switches it off for A/B runs)
  const bool smem_env = !(std::getenv("LMC_KMC_TEAM_SMEM") && std::atoi(std::getenv("LMC_KMC_TEAM_SMEM")) == 0);
  const size_t occ_smem = static_cast<size_t>(lat.padded_size) + 16;
  auto kernel_for = [&](int lanes, bool resident_occ) -> const void * {
#define LMC_TEAM_CASE(G) \
    case G: return resident_occ ? (instrumented ? reinterpret_cast<const void *>(kmc_team_run_kernel<G, true, true>) : reinterpret_cast<const void *>(kmc_team_run_kernel<G, false, true>)) \
                                : (instrumented ? reinterpret_cast<const void *>(kmc_team_run_kernel<G, true, false>) : reinterpret_cast<const void *>(kmc_team_run_kernel<G, false, false>));
    switch (lanes) {
      LMC_TEAM_CASE(8)
      LMC_TEAM_CASE(16)
      LMC_TEAM_CASE(32)
    }
#undef LMC_TEAM_CASE
    return nullptr;
  };
  int sms = 0, dev = 0;
  LMC_CUDA(cudaGetDevice(&dev));
  LMC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  for (int lanes : {32, 16, 8}) {
    if (forced > 0 && lanes != forced) continue;
    // measured on B200 (tools/kmc_team_sweep.sh, profiles/r2_k_team_sweep.txt): 32 lanes per jump win by ~5 % around one
    // walker per SM, 16 lanes up to ~3 walkers per SM (and for a handful of walkers), 8 lanes up to 7
    if (forced <= 0 && lanes == 32 && n_walkers < 64) continue;
    const double per_sm_limit = lanes == 32 ? 1.0 : (lanes == 16 ? 3.0 : static_cast<double>(max_per_sm));
    for (int resident_occ = (smem_env && occ_smem <= (48u << 10)) ? 1 : 0; resident_occ >= 0; --resident_occ) {
      const void *kernel = kernel_for(lanes, resident_occ != 0);
      const size_t smem = table_smem + (resident_occ ? occ_smem : 0);
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) { cudaGetLastError(); continue; }
      int per_sm = 0;
      LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 12 * lanes + 32, smem));
      const int64_t resident = static_cast<int64_t>(per_sm) * sms;
      if (per_sm > 0 && ((forced > 0 && (n_walkers <= resident || !resident_occ || (for_tail && per_sm >= 7))) || (n_walkers <= resident && n_walkers <= per_sm_limit * sms))) {
        *lanes_out = lanes;
        *smem_out = smem;
        kmc_team_resident_occ = resident_occ != 0;
        return kernel;
      }
    }
  }
  if (forced > 0) throw std::invalid_argument("LMC_KMC_TEAM_LANES must be 0, 8, 16 or 32");
  return nullptr;
}

void Engine::kmc_run(const lmc_kmc_params &params, int64_t n_steps, const double *u1, const double *u2, const lmc_kmc_trace *trace,
                     bool second_order) {
  require_device();
  require_coefficients();
  if (!pair_tables.has_barrier) throw std::invalid_argument("the coefficient file has no per-element quartic blocks");
  if (!kmc_ready) kmc_reset();
  if (!second_order && (u1 == nullptr) != (u2 == nullptr)) throw std::invalid_argument("replay needs both u1 and u2");
  if (second_order) u1 = u2;      // one uniform per step (SelectEvent); staged once below
  if (n_steps <= 0) return;
  const size_t nw = static_cast<size_t>(n_walkers), total = nw * static_cast<size_t>(n_steps);
  // temperatures
  std::vector<double> temps(nw, params.temperature);
  if (params.temperatures) std::copy(params.temperatures, params.temperatures + nw, temps.begin());
  for (double t : temps)
    if (!(t > 0.0) && params.n_time_temperature == 0) throw std::invalid_argument("temperature must be positive");
  LMC_CUDA(cudaMemcpyAsync(d_kmc_temperature, temps.data(), nw * 8, cudaMemcpyHostToDevice, stream));
  // staging: [tt_time | tt_temp | u1 | u2 | trace...]
  const size_t n_tt = static_cast<size_t>(std::max(0, params.n_time_temperature));
  const bool tracing = trace != nullptr;
  size_t bytes = 2 * n_tt * 8 + 64;
  if (u2) bytes += 2 * total * 8;
  if (tracing) bytes += total * (8 + 8 + 4 + 5 * 8) + 64;
  char *d = static_cast<char *>(scratch(bytes));
  double *d_tt_time = reinterpret_cast<double *>(d), *d_tt_temp = d_tt_time + n_tt;
  double *cursor = d_tt_temp + n_tt;
  if (n_tt) {
    LMC_CUDA(cudaMemcpyAsync(d_tt_time, params.tt_time, n_tt * 8, cudaMemcpyHostToDevice, stream));
    LMC_CUDA(cudaMemcpyAsync(d_tt_temp, params.tt_temperature, n_tt * 8, cudaMemcpyHostToDevice, stream));
  }
  double *d_u1 = nullptr, *d_u2 = nullptr;
  if (u2) {
    d_u1 = cursor; d_u2 = d_u1 + total; cursor = d_u2 + total;
    if (!second_order) LMC_CUDA(cudaMemcpyAsync(d_u1, u1, total * 8, cudaMemcpyHostToDevice, stream));
    LMC_CUDA(cudaMemcpyAsync(d_u2, u2, total * 8, cudaMemcpyHostToDevice, stream));
  }
  KmcTraceDev tr{};
  if (tracing) {
    tr.from = trace->from ? reinterpret_cast<int64_t *>(cursor) : nullptr; cursor += total;
    tr.to = trace->to ? reinterpret_cast<int64_t *>(cursor) : nullptr; cursor += total;
    tr.dt = trace->dt ? cursor : nullptr; cursor += total;
    tr.Ea = trace->Ea ? cursor : nullptr; cursor += total;
    tr.dE = trace->dE ? cursor : nullptr; cursor += total;
    tr.total_rate = trace->total_rate ? cursor : nullptr; cursor += total;
    tr.temperature = trace->temperature ? cursor : nullptr; cursor += total;
    tr.slot = trace->slot ? reinterpret_cast<int32_t *>(cursor) : nullptr;
  }
  KmcState st{d_kmc_vacancy, d_kmc_time, d_kmc_energy, d_kmc_steps, d_kmc_temperature, d_kmc_cvac, d_kmc_csol, d_kmc_error, d_kmc_previous};
  // LMC_KMC_SELECT_MARGIN (tests; read at every launch): a value > 1 sends every first-order step through the sequential select
  const double select_margin = std::getenv("LMC_KMC_SELECT_MARGIN") ? std::atof(std::getenv("LMC_KMC_SELECT_MARGIN")) : kSelectMargin;
  // LMC_KMC_FINISH_TIMES=1 (diagnostics): when each walker's half-warp left kmc_run_kernel, as deciles of the launch
  unsigned long long *d_finish = nullptr;
  if (std::getenv("LMC_KMC_FINISH_TIMES") && !second_order) {
    LMC_CUDA(cudaMalloc(&d_finish, nw * 8));
    LMC_CUDA(cudaMemsetAsync(d_finish, 0, nw * 8, stream));
  }
  KmcParams prm{static_cast<int32_t>(n_tt), d_tt_time, d_tt_temp, params.rate_corrector, params.seed, nullptr, 0, nullptr, nullptr, nullptr, d_finish,
                std::max(select_margin, kSelectMargin)};
  const int walkers_per_block = kKmcThreads / 16;
  const unsigned blocks = static_cast<unsigned>((n_walkers + walkers_per_block - 1) / walkers_per_block);
  if (lat.num_sites >= (1LL << 31)) throw std::invalid_argument("the KMC driver orders events by 32-bit lattice ids (num_sites < 2^31)");
  const size_t kmc_smem = static_cast<size_t>(species.n) * kEnvN * species.n * 2 * sizeof(double);
  const bool instrumented = tracing || d_u1 != nullptr;      // the first-order kernel reads its tables through L1: no dynamic shared memory
  size_t team_smem = kmc_smem;
  kmc_handoff = false;
  time_begin();
  if (second_order) {
    LMC_CUDA(cudaFuncSetAttribute(kmc_chain_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kmc_smem)));
    kmc_chain_run_kernel<<<static_cast<unsigned>(n_walkers), kChainThreads, kmc_smem, stream>>>(lat, tab, d_occ, lat.padded_size, n_walkers, st,
                                                                                               prm, n_steps, d_u2, tr);
  } else if (const void *team = kmc_team_kernel_choice(instrumented, kmc_smem, &kmc_team_lanes, &team_smem)) {
    // few walkers: a thread block per walker (kmc_team_kernels.cuh) -- the step is latency-bound, not throughput-bound
    void *args[] = {&lat, &tab, &d_occ, const_cast<int64_t *>(&lat.padded_size), &n_walkers, &st, &prm, &n_steps, &d_u1, &d_u2, &tr};
    LMC_CUDA(cudaLaunchKernel(team, dim3(static_cast<unsigned>(n_walkers)), dim3(static_cast<unsigned>(12 * kmc_team_lanes + 32)), args, team_smem, stream));
  } else {
    kmc_team_lanes = 0;
    // Tail hand-off.  A launch of thousands of walkers lasts as long as its most expensive walker (the first half-warp
    // is through at ~0.4 of the kernel time, the median at 0.7: profiles/r2_m_kmc_finish_times.txt).  Once all but
    // `keep` walkers are through, the half-warps still running stop at a 16-step boundary and the latency kernel takes
    // each of those walkers from where it stopped to its target -- a block per walker runs such a step 2-3 x faster.
    // Every launch shape computes bit-identical (dE, log E0, rate) (tables on a binary grid, one rate chain), so the
    // result does not depend on when the hand-off happens.  LMC_KMC_HANDOFF=0 switches it off, a value in (0, 1) sets
    // the fraction of the walkers handed over.  Applies whenever this kernel is chosen (more walkers than the latency
    // kernel can hold at once) and the launch is long enough to have a tail (>= 128 steps).
    // How many walkers to hand over (measured on B200, ms per 2048-hop launch, profiles/r2_n_handoff_sweep*.txt and
    // r2_p_handoff_by_walkers.txt): 8192 walkers -- 35 % is the optimum (14.1; none 16.0, 50 % 14.6); 4096 -- 65 % (9.65; none
    // 11.5, 35 % 9.9); 2048 and 1536 -- 80 % (8.3; none 10.3, 35 % 8.7).  In walkers that is ~19 per SM (2.7 rounds of the
    // 7 blocks an SM holds) wherever that is less than 80 % of the launch.
    int sm_count = 0, cur_dev = 0;
    LMC_CUDA(cudaGetDevice(&cur_dev));
    LMC_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, cur_dev));
    double keep_fraction = std::min(0.8, 19.0 * sm_count / std::max(1, n_walkers));
    if (const char *v = std::getenv("LMC_KMC_HANDOFF")) keep_fraction = std::atof(v);
    static const int handoff_min_walkers = std::getenv("LMC_KMC_HANDOFF_MIN_WALKERS") ? std::atoi(std::getenv("LMC_KMC_HANDOFF_MIN_WALKERS")) : 256;
    const bool handoff = !instrumented && keep_fraction > 0.0 && keep_fraction < 1.0 && n_walkers >= handoff_min_walkers && n_steps >= 128;
    kmc_handoff = handoff;
    if (handoff) {
      kmc_target_kernel<<<static_cast<unsigned>((n_walkers + 255) / 256), 256, 0, stream>>>(d_kmc_steps, d_kmc_target, n_walkers, n_steps);
      LMC_CUDA(cudaMemsetAsync(d_kmc_done, 0, sizeof(int), stream));
      prm.handoff_done = d_kmc_done;
      prm.handoff_threshold = n_walkers - std::max(1, static_cast<int>(keep_fraction * n_walkers));
    }
    if (instrumented) kmc_run_kernel<true><<<blocks, kKmcThreads, 0, stream>>>(lat, tab, d_occ, lat.padded_size, n_walkers, st, prm, n_steps, d_u1, d_u2, tr);
    else kmc_run_kernel<false><<<blocks, kKmcThreads, 0, stream>>>(lat, tab, d_occ, lat.padded_size, n_walkers, st, prm, n_steps, d_u1, d_u2, tr);
    if (handoff) {
      LMC_CUDA(cudaGetLastError());
      int lanes = 0;
      const void *tail = kmc_team_kernel_choice(false, kmc_smem, &lanes, &team_smem, true);
      prm.handoff_done = nullptr;
      prm.steps_target = d_kmc_target;
      prm.tail_order = d_kmc_tail_order;
      prm.tail_count = d_kmc_done + 1;
      kmc_tail_order_kernel<<<1, 1024, 0, stream>>>(d_kmc_steps, d_kmc_target, d_kmc_error, n_walkers, n_steps, d_kmc_tail_order, d_kmc_done + 1);
      // (fewer resident blocks per SM for the tail -- 5, 4, 3 instead of 7 -- measured: no gain, profiles/r2_n_tail_sweep.txt)
      void *args[] = {&lat, &tab, &d_occ, const_cast<int64_t *>(&lat.padded_size), &n_walkers, &st, &prm, &n_steps, &d_u1, &d_u2, &tr};
      LMC_CUDA(cudaLaunchKernel(tail, dim3(static_cast<unsigned>(n_walkers)), dim3(static_cast<unsigned>(12 * lanes + 32)), args, team_smem, stream));
      launch_count += 3;            // kmc_target_kernel, kmc_tail_order_kernel and the tail kernel, besides kmc_run_kernel (time_end)
    }
  }
  time_end();
  LMC_CUDA(cudaGetLastError());
  if (tracing) {
    auto back = [&](void *host, const void *dev, size_t elem) {
      if (host && dev) LMC_CUDA(cudaMemcpyAsync(host, dev, total * elem, cudaMemcpyDeviceToHost, stream));
    };
    back(trace->from, tr.from, 8); back(trace->to, tr.to, 8); back(trace->slot, tr.slot, 4); back(trace->dt, tr.dt, 8);
    back(trace->Ea, tr.Ea, 8); back(trace->dE, tr.dE, 8); back(trace->total_rate, tr.total_rate, 8);
    back(trace->temperature, tr.temperature, 8);
  }
  std::vector<int32_t> err(nw);
  LMC_CUDA(cudaMemcpyAsync(err.data(), d_kmc_error, nw * 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  if (d_finish) {
    std::vector<unsigned long long> fin(nw);
    LMC_CUDA(cudaMemcpy(fin.data(), d_finish, nw * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_finish);
    std::vector<unsigned long long> t;
    for (unsigned long long v : fin) if (v) t.push_back(v);
    if (!t.empty()) {
      std::sort(t.begin(), t.end());
      const double kernel_ns = last_kernel_ms() * 1e6;      // the last walker leaves at the end of the kernel
      std::fprintf(stderr, "kmc finish times (%zu walkers, kernel %.3f ms), fraction of the kernel time at which the given fraction of walkers had finished:", t.size(), kernel_ns * 1e-6);
      for (double q : {0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.98, 0.99, 0.995, 0.999})
        std::fprintf(stderr, " q%.3g=%.3f", q, 1.0 - static_cast<double>(t.back() - t[static_cast<size_t>(q * (t.size() - 1))]) / kernel_ns);
      // persistence of a walker's cost from launch to launch (decides whether the previous launch can order the next one)
      static std::vector<unsigned long long> previous;
      if (previous.size() == fin.size()) {
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        const double n = static_cast<double>(fin.size());
        for (size_t w = 0; w < fin.size(); ++w) {
          const double x = static_cast<double>(previous[w]), y = static_cast<double>(fin[w] - t.front());
          sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
        }
        const double cov = sxy - sx * sy / n, vx = sxx - sx * sx / n, vy = syy - sy * sy / n;
        std::fprintf(stderr, " | correlation with the previous launch's finish times %.3f", cov / std::sqrt(vx * vy));
      }
      previous.resize(fin.size());
      for (size_t w = 0; w < fin.size(); ++w) previous[w] = fin[w] - t.front();
      std::fprintf(stderr, "\n");
    }
  }
  for (size_t w = 0; w < nw; ++w)
    if (err[w]) {
      kmc_ready = false;
      throw std::out_of_range("KMC walker " + std::to_string(w) + ": Cluster not found in ClusterIndexer (vacancy lost or second vacancy in range)");
    }
}

void Engine::kmc_set_state(const double *time, const double *energy, const int64_t *steps) {
  require_device();
  if (!kmc_ready) kmc_reset();
  const size_t nw = static_cast<size_t>(n_walkers);
  if (steps)
    for (size_t w = 0; w < nw; ++w)
      if (steps[w] < 0) throw std::invalid_argument("steps must be >= 0");
  if (time) LMC_CUDA(cudaMemcpyAsync(d_kmc_time, time, nw * 8, cudaMemcpyHostToDevice, stream));
  if (energy) LMC_CUDA(cudaMemcpyAsync(d_kmc_energy, energy, nw * 8, cudaMemcpyHostToDevice, stream));
  if (steps) LMC_CUDA(cudaMemcpyAsync(d_kmc_steps, steps, nw * 8, cudaMemcpyHostToDevice, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
}

void Engine::kmc_get_state(double *time, double *energy, int64_t *steps, int64_t *vacancy, double *temperature) {
  require_device();
  if (!d_kmc_vacancy) throw std::invalid_argument("lmc_kmc_reset has not been called");
  const size_t nw = static_cast<size_t>(n_walkers);
  if (time) LMC_CUDA(cudaMemcpyAsync(time, d_kmc_time, nw * 8, cudaMemcpyDeviceToHost, stream));
  if (energy) LMC_CUDA(cudaMemcpyAsync(energy, d_kmc_energy, nw * 8, cudaMemcpyDeviceToHost, stream));
  if (steps) LMC_CUDA(cudaMemcpyAsync(steps, d_kmc_steps, nw * 8, cudaMemcpyDeviceToHost, stream));
  if (vacancy) LMC_CUDA(cudaMemcpyAsync(vacancy, d_kmc_vacancy, nw * 8, cudaMemcpyDeviceToHost, stream));
  if (temperature) LMC_CUDA(cudaMemcpyAsync(temperature, d_kmc_temperature, nw * 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
}

// ------------------------------------------------------------------------------------------------ grid / multi-GPU CMC
void Engine::cmc_grid_prepare() {
  require_device();
  if (d_cmc_xchg) return;
  d_cmc_xchg = dev_alloc<CmcExchange>(1);
  d_cmc_grid_counter = dev_alloc<unsigned long long>(1);
  d_cmc_sequence = dev_alloc<unsigned long long>(1);
  d_cmc_abort = dev_alloc<int>(1);
  d_cmc_accum = dev_alloc<unsigned long long>(8);
  for (void *p : {d_cmc_xchg, static_cast<void *>(d_cmc_grid_counter), static_cast<void *>(d_cmc_sequence), static_cast<void *>(d_cmc_abort),
                  static_cast<void *>(d_cmc_accum)})
    device_allocs.push_back(p);
  LMC_CUDA(cudaMemsetAsync(d_cmc_xchg, 0, sizeof(CmcExchange), stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_sequence, 0, 8, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  cmc_peer_xchg[0] = d_cmc_xchg;
}

void Engine::cmc_exchange_handle(void *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!handle64) throw std::invalid_argument("null handle buffer");
  cmc_grid_prepare();
  cudaIpcMemHandle_t h;
  LMC_CUDA(cudaIpcGetMemHandle(&h, d_cmc_xchg));
  std::memcpy(handle64, &h, 64);
}

void Engine::cmc_attach_peers(int32_t rank, int32_t world, const void *handles, int32_t grid_ctas) {
  cmc_grid_prepare();
  if (world < 1 || world > kGridMaxWorld || rank < 0 || rank >= world) throw std::invalid_argument("rank / world out of range (world <= 8)");
  if (world > 1 && !handles) throw std::invalid_argument("null handles");
  if (n_walkers != 1) throw std::invalid_argument("the multi-GPU CMC driver runs ONE replicated lattice (n_walkers == 1)");
  for (int r = 0; r < 8; ++r) {
    if (cmc_peer_xchg[r] && cmc_peer_xchg[r] != d_cmc_xchg) cudaIpcCloseMemHandle(cmc_peer_xchg[r]);
    cmc_peer_xchg[r] = nullptr;
  }
  for (int r = 0; r < world; ++r) {
    if (r == rank) { cmc_peer_xchg[r] = d_cmc_xchg; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char *>(handles) + 64 * r, 64);
    void *mapped = nullptr;
    LMC_CUDA(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
    cmc_peer_xchg[r] = mapped;
  }
  cmc_world = world;
  cmc_rank = rank;
  cmc_grid_ctas = grid_ctas;
  // a fresh world starts a fresh flag sequence (collective: every rank attaches before any rank runs)
  LMC_CUDA(cudaMemsetAsync(d_cmc_xchg, 0, sizeof(CmcExchange), stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_sequence, 0, 8, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
}

void Engine::cmc_grid_run(const lmc_cmc_params &params, int64_t n_trials) {
  const bool dbg = std::getenv("LMC_DEBUG_TIMING") != nullptr;
  auto t_dbg = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!dbg) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[cmc_grid_run] %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_dbg).count());
    t_dbg = now;
  };
  require_device();
  require_coefficients();
  if (n_walkers != 1) throw std::invalid_argument("lmc_cmc_grid_run drives ONE lattice with the whole GPU (n_walkers == 1)");
  if (!cmc_ready) cmc_reset(0.0, 0);
  if (cmc_cells_stale) cmc_sync_cells();
  cmc_grid_prepare();
  if (n_trials <= 0) return;
  const double temp = params.temperatures ? params.temperatures[0] : params.temperature;
  LMC_CUDA(cudaMemcpyAsync(d_cmc_temperature, &temp, 8, cudaMemcpyHostToDevice, stream));
  unsigned long long steps0 = 0;
  LMC_CUDA(cudaMemcpyAsync(&steps0, d_cmc_steps, 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  const unsigned long long target = steps0 + static_cast<unsigned long long>(n_trials);
  tick("steps readback");
  int sms = 0;
  sms = device_attr(cudaDevAttrMultiProcessorCount);
  int ctas = std::min(sms, kGridMaxCtas);
  if (cmc_grid_ctas > 0) ctas = std::min(ctas, cmc_grid_ctas);
  // proposals per CTA and batch (= 2 x the trials a CTA evaluates): about N / 172 trials of a batch can be mutually
  // non-interfering.  A trial is evaluated by 2 L lanes; small lattices, whose batches leave the SMs almost empty, get
  // wide groups (L = 8 at 40^3: 32 trials x 16 lanes per CTA)
  int proposals = params.batch_size;
  if (proposals <= 0) {
    const int64_t want_pairs = std::max<int64_t>(16, lat.num_sites / 172);
    proposals = 64;
    while (proposals < kCmcMaxThreads / 2 && static_cast<int64_t>(proposals) / 2 * ctas < want_pairs) proposals *= 2;   // <= 256: at least 2 lanes per side
  }
  if (proposals < 32 || proposals > kCmcMaxThreads || (proposals & (proposals - 1))) throw std::invalid_argument("batch_size must be a power of two in 32..512");
  if (static_cast<int64_t>(ctas) * (proposals / 2) > 65535) throw std::invalid_argument("batch too large for 16-bit claim priorities");
  int lanes = 8;
  if (const char *v = std::getenv("LMC_CMC_GRID_LANES")) lanes = std::max(1, std::atoi(v));   // tuning knob: lanes per trial side
  while (lanes > 1 && proposals * lanes > kCmcMaxThreads) lanes /= 2;
  if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16) throw std::invalid_argument("lanes per trial side must be 1, 2, 4, 8 or 16");
  const int threads = proposals * lanes;            // (proposals / 2) trials x 2 sides x L lanes
  const int m = species.n + 1;
  const size_t a_len = static_cast<size_t>(m) * kSiteEnvN * m, b_len = static_cast<size_t>(m) * tab.n_site_pairs * m * m;
  const size_t fixed = (m + a_len) * 8 + kSiteEnvN * 8 + 44 * 2 + kSiteEnvN * kSiteEnvN + 4 + static_cast<size_t>(threads) * 43 + 64;
  int max_optin = 0;
  max_optin = device_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin);
  const int stage_b = (fixed + b_len * 8 + 24 * 1024 <= static_cast<size_t>(max_optin)) ? 1 : 0;
  const size_t smem = fixed + (stage_b ? b_len * 8 : 0);
  using GridKernel = void (*)(LatticeDesc, DevTables, uint8_t *, uint8_t *, unsigned int *, CmcState, const double *, uint64_t, unsigned long long, CmcGridParams);
  GridKernel kernel = nullptr;
  switch (lanes * 2 + stage_b) {
    case 2: kernel = cmc_grid_kernel<1, false>; break;
    case 3: kernel = cmc_grid_kernel<1, true>; break;
    case 4: kernel = cmc_grid_kernel<2, false>; break;
    case 5: kernel = cmc_grid_kernel<2, true>; break;
    case 8: kernel = cmc_grid_kernel<4, false>; break;
    case 9: kernel = cmc_grid_kernel<4, true>; break;
    case 16: kernel = cmc_grid_kernel<8, false>; break;
    case 17: kernel = cmc_grid_kernel<8, true>; break;
    case 32: kernel = cmc_grid_kernel<16, false>; break;
    default: kernel = cmc_grid_kernel<16, true>; break;
  }
  if (cmc_grid_checked_threads != threads || cmc_grid_checked_smem != smem || cmc_grid_checked_kernel != reinterpret_cast<const void *>(kernel)) {     // once per launch shape
    LMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) throw std::runtime_error("cmc_grid_kernel does not fit on an SM");
    cmc_grid_checked_threads = threads;
    cmc_grid_checked_smem = smem;
    cmc_grid_checked_kernel = reinterpret_cast<const void *>(kernel);
  }
  CmcGridParams gp{};
  gp.world = cmc_world;
  gp.rank = cmc_rank;
  for (int r = 0; r < kGridMaxWorld; ++r) gp.xchg[r] = static_cast<CmcExchange *>(cmc_peer_xchg[r]);
  gp.xchg[cmc_rank] = static_cast<CmcExchange *>(d_cmc_xchg);
  gp.barrier_counter = d_cmc_grid_counter;
  gp.abort_flag = d_cmc_abort;
  gp.sequence = d_cmc_sequence;
  gp.accum = d_cmc_accum;
  int clock_khz = 0;
  clock_khz = device_attr(cudaDevAttrClockRate);
  gp.spin_limit = static_cast<long long>(clock_khz) * 1000LL * 5LL;        // ~5 s: a lost peer ends the launch instead of hanging it
  LMC_CUDA(cudaMemsetAsync(d_cmc_grid_counter, 0, 8, stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_abort, 0, 4, stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_accum, 0, 64, stream));
  CmcState st{d_cmc_energy, d_cmc_steps, d_cmc_accepted, d_cmc_proposals, d_cmc_epoch, static_cast<SaSchedule *>(d_cmc_sa), d_cmc_error};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // co-residency of all CTAs: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  tick("setup");
  time_begin();
  LMC_CUDA(cudaLaunchKernelEx(&cfg, kernel, lat, tab, d_occ, d_cmc_mirror, d_cmc_marks, st, static_cast<const double *>(d_cmc_temperature),
                              static_cast<uint64_t>(params.seed), target, gp));
  time_end();
  tick("launch");
  LMC_CUDA(cudaGetLastError());
  int32_t err = 0;
  int aborted = 0;
  LMC_CUDA(cudaMemcpyAsync(&err, d_cmc_error, 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(&aborted, d_cmc_abort, 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  tick("kernel + sync");
  if (aborted) {
    cmc_ready = false;
    throw std::runtime_error("CMC grid run: barrier / peer exchange timed out (a peer rank is missing or out of step)");
  }
  if (err) {
    cmc_ready = false;
    throw std::out_of_range("CMC: Cluster not found in ClusterIndexer (two vacancies within interaction range)");
  }
}

// ------------------------------------------------------------------------------------------------ domain-decomposed CMC / SA
void Engine::cmc_domain_prepare() {
  require_device();
  if (d_occ_buf[1]) return;
  const size_t bytes = static_cast<size_t>(n_walkers) * lat.padded_size + 16;
  d_occ_buf[0] = d_occ;
  LMC_CUDA(cudaMalloc(&d_occ_buf[1], bytes));
  occ_cur = 0;
  d_dom_state = dev_alloc<DomState>(2 * static_cast<size_t>(n_walkers));
  d_dom_accum = dev_alloc<unsigned long long>(3 * static_cast<size_t>(n_walkers) * 4);
  d_dom_lines = dev_alloc<DomLine>(2 * kGridMaxWorld);
  d_dom_counters = dev_alloc<unsigned long long>(8);   // [0] grid barrier, [1] inter-GPU line sequence, [2] sweeps done, [4..5] domain queues (3 x u32)
  d_dom_abort = dev_alloc<int>(1);
  for (void *p : {d_dom_state, static_cast<void *>(d_dom_accum), d_dom_lines, static_cast<void *>(d_dom_counters), static_cast<void *>(d_dom_abort)})
    device_allocs.push_back(p);
  LMC_CUDA(cudaMemsetAsync(d_occ_buf[1], 0, bytes, stream));
  LMC_CUDA(cudaMemsetAsync(d_dom_lines, 0, sizeof(DomLine) * 2 * kGridMaxWorld, stream));
  LMC_CUDA(cudaMemsetAsync(d_dom_counters, 0, 64, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  for (int b = 0; b < 2; ++b) dom_peer_occ[b][0] = d_occ_buf[b];
  dom_peer_lines[0] = d_dom_lines;
}

void Engine::cmc_domain_handles(void *handles192) {
  if (!handles192) throw std::invalid_argument("null handle buffer");
  cmc_domain_prepare();
  cudaIpcMemHandle_t h[3];
  LMC_CUDA(cudaIpcGetMemHandle(&h[0], d_occ_buf[0]));
  LMC_CUDA(cudaIpcGetMemHandle(&h[1], d_occ_buf[1]));
  LMC_CUDA(cudaIpcGetMemHandle(&h[2], d_dom_lines));
  std::memcpy(handles192, h, 192);
}

void Engine::cmc_domain_attach_peers(int32_t rank, int32_t world, const void *handles) {
  cmc_domain_prepare();
  if (world < 1 || world > kGridMaxWorld || rank < 0 || rank >= world) throw std::invalid_argument("rank / world out of range (world <= 8)");
  if (world > 1 && !handles) throw std::invalid_argument("null handles");
  if (n_walkers != 1) throw std::invalid_argument("the multi-GPU domain driver runs ONE lattice (n_walkers == 1)");
  for (int r = 0; r < 8; ++r) {
    if (r != dom_rank) {
      for (int b = 0; b < 2; ++b)
        if (dom_peer_occ[b][r]) cudaIpcCloseMemHandle(dom_peer_occ[b][r]);
      if (dom_peer_lines[r]) cudaIpcCloseMemHandle(dom_peer_lines[r]);
    }
    dom_peer_occ[0][r] = dom_peer_occ[1][r] = nullptr;
    dom_peer_lines[r] = nullptr;
  }
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      dom_peer_occ[0][r] = d_occ_buf[0]; dom_peer_occ[1][r] = d_occ_buf[1]; dom_peer_lines[r] = d_dom_lines;
      continue;
    }
    cudaIpcMemHandle_t h[3];
    std::memcpy(h, static_cast<const char *>(handles) + 192 * r, 192);
    void *mapped[3] = {nullptr, nullptr, nullptr};
    for (int q = 0; q < 3; ++q) LMC_CUDA(cudaIpcOpenMemHandle(&mapped[q], h[q], cudaIpcMemLazyEnablePeerAccess));
    dom_peer_occ[0][r] = static_cast<uint8_t *>(mapped[0]);
    dom_peer_occ[1][r] = static_cast<uint8_t *>(mapped[1]);
    dom_peer_lines[r] = mapped[2];
  }
  dom_world = world;
  dom_rank = rank;
  // a fresh world starts a fresh line sequence (collective: every rank attaches before any rank runs)
  LMC_CUDA(cudaMemsetAsync(d_dom_lines, 0, sizeof(DomLine) * 2 * kGridMaxWorld, stream));
  LMC_CUDA(cudaMemsetAsync(d_dom_counters, 0, 16, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
}

void Engine::cmc_sync_cells() {
  const unsigned blocks = static_cast<unsigned>((lat.num_sites + 255) / 256);
  for (int w0 = 0; w0 < n_walkers; w0 += 32768) {
    const unsigned ny = static_cast<unsigned>(std::min(32768, n_walkers - w0));
    cmc_mirror_kernel<<<dim3(blocks, ny), 256, 0, stream>>>(lat, d_occ + static_cast<int64_t>(w0) * lat.padded_size,
                                                            d_cmc_mirror + static_cast<int64_t>(w0) * lat.num_sites);
    ++launch_count;
  }
  // claim marks cleared, species bytes copied: the packed cell array of the batched CMC kernels
  const int64_t n_cells = static_cast<int64_t>(n_walkers) * lat.padded_size;
  cmc_cells_init_kernel<<<static_cast<unsigned>((n_cells + 255) / 256), 256, 0, stream>>>(n_cells, d_occ, d_cmc_marks);
  ++launch_count;
  LMC_CUDA(cudaGetLastError());
  LMC_CUDA(cudaMemsetAsync(d_cmc_epoch, 0, static_cast<size_t>(n_walkers) * 8, stream));   // the marks were cleared: epochs may restart
  cmc_cells_stale = false;
}

void Engine::cmc_domain_run(const lmc_cmc_params &params, const lmc_cmc_domain_params *dom, int64_t n_trials) {
  require_device();
  require_coefficients();
  if (n_walkers > kDomMaxWalkers) throw std::invalid_argument("lmc_cmc_domain_run drives at most 2048 replicas per engine");
  if (!cmc_ready) cmc_reset(0.0, 0);
  cmc_domain_prepare();
  if (dom_world > 1 && n_walkers != 1) throw std::invalid_argument("the multi-GPU domain driver runs ONE lattice (n_walkers == 1)");
  if (n_trials <= 0) return;
  const size_t nw = static_cast<size_t>(n_walkers);
  std::vector<double> temps(nw, params.temperature);
  if (params.temperatures) std::copy(params.temperatures, params.temperatures + nw, temps.begin());
  LMC_CUDA(cudaMemcpyAsync(d_cmc_temperature, temps.data(), nw * 8, cudaMemcpyHostToDevice, stream));
  std::vector<unsigned long long> steps(nw);
  unsigned long long sweep0 = 0;
  LMC_CUDA(cudaMemcpyAsync(steps.data(), d_cmc_steps, nw * 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(&sweep0, d_dom_counters + 2, 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  const unsigned long long target = *std::min_element(steps.begin(), steps.end()) + static_cast<unsigned long long>(n_trials);
  // ---- decomposition
  // defaults measured on B200 (tools/domain_probe.py): domains of 6 half lattice constants (cores of 4 x 4 x 4 / 2 = 32 sites) and 216
  // rounds per sweep beat 8 / 216 by 18 % at 100^3, by 50 % at 40^3 and by 25 % at the 2000 domains per GPU of an 8-GPU run
  const int edge = dom && dom->domain_edge > 0 ? dom->domain_edge : 6;
  if (edge < 4 || edge > 48) throw std::invalid_argument("domain_edge must be in 4..48 half lattice constants");
  const int px = 2 * lat.fx, py = 2 * lat.fy;
  CmcDomainParams dp{};
  dp.ndx = std::max(1, px / edge); dp.ndy = std::max(1, py / edge); dp.ndz = std::max(1, (2 * lat.fz) / edge);
  auto max_size = [](int period, int nd) { return (period + nd - 1) / nd; };
  const int dx_max = max_size(px, dp.ndx), dy_max = max_size(py, dp.ndy), dz_max = 2 * max_size(lat.fz, dp.ndz);
  const int dz_min = 2 * (lat.fz / dp.ndz);
  if (px / dp.ndx < 4 || py / dp.ndy < 4 || dz_min < 4) throw std::invalid_argument("domains must span at least 4 half lattice constants per axis");
  if (px >= 32768 || py >= 32768 || lat.fz >= 32768) throw std::invalid_argument("lattice too large for the 32-bit domain bounds");
  if (static_cast<long long>(dx_max - 2) * (dy_max - 2) * (dz_max - 2) / 2 > 65535) throw std::invalid_argument("domain core too large (<= 65535 sites)");
  dp.tile_y = dy_max + 2;
  dp.tile_zh = (dz_max + 2) / 2;
  const int tile_cells = (dx_max + 2) * dp.tile_y * dp.tile_zh;
  if (2 * dp.tile_y * dp.tile_zh + 2 * dp.tile_zh + 2 > 32767) throw std::invalid_argument("domain tile too large for 16-bit offsets");
  dp.tile_cells = (tile_cells + 15) & ~15;
  dp.max_core = (dx_max - 2) * (dy_max - 2) * (dz_max - 2) / 2;
  dp.rounds = dom && dom->rounds_per_sweep > 0 ? dom->rounds_per_sweep : 216;
  const double dom_passes = dom ? dom->passes : 0.0;
  dp.n_walkers = n_walkers;
  dp.world = dom_world; dp.rank = dom_rank;
  // ---- launch shape: one lane group per domain; wide groups while the domains of the whole job leave the SMs empty
  const int sms = device_attr(cudaDevAttrMultiProcessorCount);
  const long long items_total = static_cast<long long>(n_walkers) * dp.ndx * dp.ndy * dp.ndz;
  const int slab = domain_slab_begin(dp.ndx, dom_world, dom_rank + 1) - domain_slab_begin(dp.ndx, dom_world, dom_rank);
  const long long items_rank = static_cast<long long>(n_walkers) * slab * dp.ndy * dp.ndz;
  // lanes per trial (G) and trials of a domain in flight (S).  Measured on B200 (tools/domain_probe.py): 8 lanes per trial and
  // four speculative trials per domain -- a whole warp per domain -- win at every load from 7 to 125 domains per SM (when the
  // domains outnumber the 32 warps of a block they are handed out dynamically over several passes: the domains of a sweep
  // differ in cost by 2x, and a domain that runs four trials at a time leaves less of a tail).  Launch shape only: the
  // trajectory does not depend on it.
  int lanes = dom ? dom->lanes : 0, spec = dom ? dom->speculate : 0;
  if (const char *v = std::getenv("LMC_CMC_DOMAIN_LANES")) lanes = std::atoi(v);        // tuning knobs
  if (const char *v = std::getenv("LMC_CMC_DOMAIN_SPECULATE")) spec = std::atoi(v);
  if (lanes <= 0) lanes = spec == 1 ? 8 : (spec == 2 ? 16 : 8);
  if (spec <= 0) spec = lanes == 8 ? 4 : (lanes == 16 ? 2 : 1);
  if (lanes != 8 && lanes != 16 && lanes != 32) throw std::invalid_argument("lanes per trial must be 8, 16 or 32");
  if (!((spec == 1) || (spec == 2 && (lanes == 8 || lanes == 16)) || (spec == 4 && lanes == 8)))
    throw std::invalid_argument("speculate must be 1, 2 (8 or 16 lanes) or 4 (8 lanes)");
  const int team = lanes * spec;                      // lanes per domain
  dp.tile_bytes = dp.tile_cells + spec * 96 + ((2 * dp.max_core + 15) & ~15);
  const int m = species.n + 1;
  // shared memory: fixed tables, then the coefficient tables -- difference tables per species pair (kTab 1) when they fit
  const int ns = m - 1, n_sp = m * (m - 1) / 2;
  const size_t base_bytes = kSiteEnvN * 8 + kDomPidxBytes + 2 * 44 * 2 + 64 + nw * 8;
  const size_t tab0_bytes = (static_cast<size_t>(m) + static_cast<size_t>(m) * kSiteEnvN * m) * 8;
  const size_t tab1_bytes = static_cast<size_t>(n_sp) * (1 + kSiteEnvN * ns + static_cast<size_t>(tab.n_site_pairs) * ns * ns) * 8;
  const int max_optin = device_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin) - 1024;     // static shared memory of the kernel
  int max_threads = kDomMaxThreads;
  if (const char *v = std::getenv("LMC_CMC_DOMAIN_THREADS")) max_threads = std::max(32, std::min(kDomMaxThreads, std::atoi(v) / 32 * 32));
  // lane groups per block: the domains of a sweep are handed out dynamically; `passes` domains per group on average
  double passes = dom_passes > 0.0 ? dom_passes : 1.0;
  if (const char *v = std::getenv("LMC_CMC_DOMAIN_PASSES")) passes = std::max(1.0, std::atof(v));
  if (spec > 1 && lanes == 16) max_threads = std::min(max_threads, 512);       // no 1024-thread instantiation of (16 lanes, 2 trials)
  int groups = static_cast<int>(std::min<long long>(max_threads / team, std::max<long long>(1, static_cast<long long>(std::ceil(static_cast<double>(items_rank) / (sms * passes))))));
  int k_tab = (m <= 8 && base_bytes + tab1_bytes + static_cast<size_t>(std::min(groups, 8)) * dp.tile_bytes <= static_cast<size_t>(max_optin)) ? 1 : 0;
  if (const char *v = std::getenv("LMC_CMC_DOMAIN_TABLES")) k_tab = std::atoi(v) ? k_tab : 0;      // A/B switch: 0 forces the A + B form
  const size_t fixed = base_bytes + (k_tab ? tab1_bytes : tab0_bytes);
  if (fixed + dp.tile_bytes > static_cast<size_t>(max_optin)) throw std::invalid_argument("domain tile does not fit in shared memory");
  while (groups > 1 && fixed + static_cast<size_t>(groups) * dp.tile_bytes > static_cast<size_t>(max_optin)) --groups;
  int threads = (groups * team + 31) / 32 * 32;
  groups = threads / team;
  while (fixed + static_cast<size_t>(groups) * dp.tile_bytes > static_cast<size_t>(max_optin)) { threads -= 32; groups = threads / team; }
  const size_t smem = fixed + static_cast<size_t>(groups) * dp.tile_bytes;
  const int ctas = static_cast<int>(std::max<long long>(1, std::min<long long>(sms, (items_rank + groups - 1) / groups)));
  // table-walk ownership: environment positions dealt to the L lanes of a side by longest-processing-time-first on the
  // number of pair partners above each position (the stripe t mod L leaves the busiest lane 1.2 - 1.65 x the mean)
  {
    const Geometry &g = geometry();
    const int L = lanes / 2;
    std::vector<int> order(kSiteEnvN), load(L, 0);
    for (int t = 0; t < kSiteEnvN; ++t) order[t] = t;
    auto weight = [&](int t) { return 2 + __builtin_popcountll(g.site_pair_mask_hi[t]); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight(a) > weight(b); });
    for (int q = 0; q < 16; ++q) dp.own_mask[q] = 0ULL;
    for (int t : order) {
      const int best = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
      load[best] += weight(t);
      dp.own_mask[best] |= 1ULL << t;
    }
  }
  const void *kernel = nullptr;
  // blocks of at most 512 threads get the 128-register instantiation (no spills in the round loop)
  const bool small = threads <= 512 && !std::getenv("LMC_CMC_DOMAIN_64REG");
  kernel = cmc_domain_kernel_for(lanes, k_tab, spec, small);
  if (!kernel) throw std::logic_error("no cmc_domain_kernel instantiation for this launch shape");
  LMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) throw std::runtime_error("cmc_domain_kernel does not fit on an SM");
  // ---- buffers: the sweep with index s reads occ[s & 1]; the data of sweep0 sits in d_occ_buf[occ_cur]
  const int phase = (occ_cur - static_cast<int>(sweep0 & 1ULL)) & 1;
  for (int i = 0; i < 2; ++i) {
    dp.occ[i] = d_occ_buf[(i + phase) & 1];
    for (int r = 0; r < kGridMaxWorld; ++r) dp.peer_occ[i][r] = dom_peer_occ[(i + phase) & 1][r];
  }
  for (int r = 0; r < kGridMaxWorld; ++r) dp.peer_lines[r] = static_cast<DomLine *>(dom_peer_lines[r]);
  dp.lines = static_cast<DomLine *>(d_dom_lines);
  dp.state = static_cast<DomState *>(d_dom_state);
  dp.accum = d_dom_accum;
  dp.barrier_counter = d_dom_counters;
  dp.line_seq = d_dom_counters + 1;
  dp.queue = reinterpret_cast<unsigned int *>(d_dom_counters + 4);
  dp.abort_flag = d_dom_abort;
  dp.sweep = d_dom_counters + 2;
  dp.spin_limit = static_cast<long long>(device_attr(cudaDevAttrClockRate)) * 1000LL * 5LL;   // ~5 s
  LMC_CUDA(cudaMemsetAsync(d_dom_counters, 0, 8, stream));
  LMC_CUDA(cudaMemsetAsync(d_dom_counters + 4, 0, 16, stream));
  LMC_CUDA(cudaMemsetAsync(d_dom_abort, 0, 4, stream));
  LMC_CUDA(cudaMemsetAsync(d_dom_accum, 0, 3 * nw * 4 * 8, stream));
  CmcState st{d_cmc_energy, d_cmc_steps, d_cmc_accepted, d_cmc_proposals, d_cmc_epoch, static_cast<SaSchedule *>(d_cmc_sa), d_cmc_error};
  cmc_domain_state_init(n_walkers, st, d_cmc_temperature, static_cast<DomState *>(d_dom_state) + ((sweep0 + 1) & 1ULL) * nw, stream);
  ++launch_count;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // co-residency of all CTAs: the per-sweep grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  dom_last_lanes = lanes; dom_last_spec = spec; dom_last_threads = threads; dom_last_ctas = ctas; dom_last_domains = static_cast<int>(items_total);
  dom_last_rounds = dp.rounds; dom_last_edge = edge;
  time_begin();
  uint64_t seed_arg = static_cast<uint64_t>(params.seed);
  unsigned long long target_arg = target;
  void *args[] = {&lat, &tab, &dp, &st, &seed_arg, &target_arg};
  LMC_CUDA(cudaLaunchKernelExC(&cfg, kernel, args));
  time_end();
  LMC_CUDA(cudaGetLastError());
  int32_t err = 0;
  int aborted = 0;
  unsigned long long sweep_end = 0;
  LMC_CUDA(cudaMemcpyAsync(&err, d_cmc_error, 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(&aborted, d_dom_abort, 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaMemcpyAsync(&sweep_end, d_dom_counters + 2, 8, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  if (aborted) {
    cmc_ready = false;
    throw std::runtime_error("CMC domain run: barrier / peer exchange timed out (a peer rank is missing or out of step)");
  }
  if (err) {
    cmc_ready = false;
    throw std::out_of_range("CMC: Cluster not found in ClusterIndexer (two vacancies within interaction range)");
  }
  // the occupancy now lives in the buffer the next sweep would read; its periodic halo images are refreshed for the other kernels
  occ_cur = (static_cast<int>(sweep_end & 1ULL) + phase) & 1;
  d_occ = d_occ_buf[occ_cur];
  cmc_domain_refresh_halo(lat, d_occ, n_walkers, stream);
  launch_count += (n_walkers + 32767) / 32768;
  LMC_CUDA(cudaGetLastError());
  cmc_cells_stale = true;
}

// ------------------------------------------------------------------------------------------------ CMC / SA driver
void Engine::cmc_reset(double sa_initial_temperature, uint64_t sa_maximum_steps) {
  require_device();
  const size_t nw = static_cast<size_t>(n_walkers);
  if (!d_cmc_energy) {
    d_cmc_energy = dev_alloc<double>(nw); d_cmc_temperature = dev_alloc<double>(nw);
    d_cmc_steps = dev_alloc<unsigned long long>(nw); d_cmc_accepted = dev_alloc<unsigned long long>(nw);
    d_cmc_proposals = dev_alloc<unsigned long long>(nw); d_cmc_epoch = dev_alloc<unsigned long long>(nw);
    d_cmc_sa = dev_alloc<SaSchedule>(nw); d_cmc_error = dev_alloc<int32_t>(nw);
    if (lat.num_sites >= (1LL << 31)) throw std::invalid_argument("the CMC driver addresses lattice ids with 32 bits (num_sites < 2^31)");
    d_cmc_marks = dev_alloc<unsigned int>(nw * static_cast<size_t>(lat.padded_size));
    d_cmc_mirror = dev_alloc<uint8_t>(nw * static_cast<size_t>(lat.num_sites));
    for (void *p : {static_cast<void *>(d_cmc_energy), static_cast<void *>(d_cmc_temperature), static_cast<void *>(d_cmc_steps),
                    static_cast<void *>(d_cmc_accepted), static_cast<void *>(d_cmc_proposals), static_cast<void *>(d_cmc_epoch),
                    d_cmc_sa, static_cast<void *>(d_cmc_error), static_cast<void *>(d_cmc_marks), static_cast<void *>(d_cmc_mirror)})
      device_allocs.push_back(p);
  }
  LMC_CUDA(cudaMemsetAsync(d_cmc_energy, 0, nw * 8, stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_steps, 0, nw * 8, stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_accepted, 0, nw * 8, stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_proposals, 0, nw * 8, stream));
  if (d_dom_counters) LMC_CUDA(cudaMemsetAsync(d_dom_counters + 2, 0, 8, stream));   // sweep counter of the domain driver
  LMC_CUDA(cudaMemsetAsync(d_cmc_epoch, 0, nw * 8, stream));
  LMC_CUDA(cudaMemsetAsync(d_cmc_error, 0, nw * 4, stream));
  cmc_sync_cells();
  // SimulatedAnnealing constructor (mc/src/SimulatedAnnealing.cpp:53-56; ratios mc/include/SimulatedAnnealing.h:45-57)
  SaSchedule sa{};
  sa.enabled = sa_maximum_steps > 0 ? 1 : 0;
  sa.temperature = sa_initial_temperature;
  sa.maximum_steps = sa_maximum_steps;
  sa.reheat_trigger_steps = std::max<unsigned long long>(1ULL, static_cast<unsigned long long>(static_cast<double>(sa_maximum_steps) * 0.05));
  sa.reheat_cooldown_steps = std::max<unsigned long long>(1ULL, static_cast<unsigned long long>(static_cast<double>(sa_maximum_steps) * 0.10));
  sa.window_size = std::max<unsigned long long>(1ULL, static_cast<unsigned long long>(static_cast<double>(sa_maximum_steps) * 0.001));
  std::vector<SaSchedule> all(nw, sa);
  LMC_CUDA(cudaMemcpyAsync(d_cmc_sa, all.data(), nw * sizeof(SaSchedule), cudaMemcpyHostToDevice, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  cmc_ready = true;
}

void Engine::cmc_run(const lmc_cmc_params &params, int64_t n_trials, int32_t replay_walker, int64_t n_replay, const int64_t *a,
                     const int64_t *b, const double *u, double *dE, double *energy_before, double *temperature_before,
                     uint8_t *accepted) {
  require_device();
  require_coefficients();
  // one lattice: the whole GPU (cooperative grid) beats one cluster of <= 16 SMs from ~32k sites on; a world of several
  // GPUs (lmc_cmc_attach_peers) always runs the grid kernel
  if (n_replay <= 0 && n_walkers == 1 && (cmc_world > 1 || lat.num_sites >= 32000)) {
    cmc_grid_run(params, n_trials);
    return;
  }
  if (!cmc_ready) cmc_reset(0.0, 0);
  if (cmc_cells_stale) cmc_sync_cells();
  const size_t nw = static_cast<size_t>(n_walkers);
  std::vector<double> temps(nw, params.temperature);
  if (params.temperatures) std::copy(params.temperatures, params.temperatures + nw, temps.begin());
  LMC_CUDA(cudaMemcpyAsync(d_cmc_temperature, temps.data(), nw * 8, cudaMemcpyHostToDevice, stream));
  int threads = params.batch_size;
  CmcState st{d_cmc_energy, d_cmc_steps, d_cmc_accepted, d_cmc_proposals, d_cmc_epoch, static_cast<SaSchedule *>(d_cmc_sa), d_cmc_error};
  CmcReplay rp{};
  const bool replaying = n_replay > 0;
  unsigned long long target = 0;
  int first_walker = 0, n_run = n_walkers;
  if (replaying) {
    if (replay_walker < 0 || replay_walker >= n_walkers) throw std::invalid_argument("walker index out of range");
    if (!a || !b || !u) throw std::invalid_argument("replay needs site_a, site_b and u");
    const size_t n = static_cast<size_t>(n_replay);
    char *d = static_cast<char *>(scratch(n * (8 + 8 + 8 + 8 + 8 + 8 + 1) + 64));
    int64_t *d_a = reinterpret_cast<int64_t *>(d), *d_b = d_a + n;
    double *d_u = reinterpret_cast<double *>(d_b + n), *d_de = d_u + n, *d_eb = d_de + n, *d_tb = d_eb + n;
    uint8_t *d_acc = reinterpret_cast<uint8_t *>(d_tb + n);
    LMC_CUDA(cudaMemcpyAsync(d_a, a, n * 8, cudaMemcpyHostToDevice, stream));
    LMC_CUDA(cudaMemcpyAsync(d_b, b, n * 8, cudaMemcpyHostToDevice, stream));
    LMC_CUDA(cudaMemcpyAsync(d_u, u, n * 8, cudaMemcpyHostToDevice, stream));
    rp = CmcReplay{d_a, d_b, d_u, d_de, d_eb, d_tb, d_acc};
    first_walker = replay_walker;
    n_run = 1;
  } else {
    if (n_trials <= 0) return;
    std::vector<unsigned long long> steps(nw);
    LMC_CUDA(cudaMemcpyAsync(steps.data(), d_cmc_steps, nw * 8, cudaMemcpyDeviceToHost, stream));
    LMC_CUDA(cudaStreamSynchronize(stream));
    unsigned long long lo = steps[0];
    for (auto s : steps) lo = std::min(lo, s);
    target = lo + static_cast<unsigned long long>(n_trials);
  }
  // per-replica pointers are offset on the host for the single-replica replay launch
  CmcState st_run = st;
  if (replaying) {
    st_run.energy += first_walker; st_run.steps += first_walker; st_run.accepted += first_walker; st_run.proposals += first_walker;
    st_run.epoch += first_walker; st_run.sa += first_walker; st_run.error += first_walker;
  }
  // Launch shape.  The number of mutually non-interfering trials of a batch peaks near N / (2 * 43 * 2) live trials; a
  // trial occupies one lane pair, and the scattered gathers are L1-wavefront bound per SM, so a replica is spread over as
  // many CTAs as the device allows (cluster of up to 16 on B200) with correspondingly small CTAs.
  int sms = 0;
  sms = device_attr(cudaDevAttrMultiProcessorCount);
  int cluster = 1;
  while (cluster < 16 && static_cast<int64_t>(n_run) * cluster * 2 <= sms) cluster *= 2;
  if (threads <= 0) {
    const int64_t want_pairs = std::max<int64_t>(16, lat.num_sites / 172);             // the non-interference optimum
    threads = 64;
    while (threads < kCmcMaxThreads && static_cast<int64_t>(threads) / 2 * cluster < want_pairs) threads *= 2;
    while (cluster > 1 && static_cast<int64_t>(threads) / 2 * (cluster / 2) >= want_pairs) cluster /= 2;   // small lattices: fewer CTAs
  }
  if (replaying) threads = std::min(threads, 128);
  if (threads < 32 || threads > kCmcMaxThreads || (threads & (threads - 1))) throw std::invalid_argument("batch_size must be a power of two in 32..512");
  // dynamic shared memory: replay dE window, site tables, per-thread species staging (2 x 43 bytes per thread)
  const int m = species.n + 1;
  const size_t a_len = static_cast<size_t>(m) * kSiteEnvN * m, b_len = static_cast<size_t>(m) * tab.n_site_pairs * m * m;
  int max_optin = 0;
  max_optin = device_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin);
  LMC_CUDA(cudaFuncSetAttribute(cmc_run_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  int stage_b = 0;
  for (;; cluster /= 2) {
    const size_t fixed = (static_cast<size_t>(threads) * cluster + m + a_len) * 8 + kSiteEnvN * 8 + 44 * 2 + static_cast<size_t>(threads) * 86 + 16;
    stage_b = (fixed + b_len * 8 + 16 * 1024 <= static_cast<size_t>(max_optin)) ? 1 : 0;
    const size_t smem = fixed + (stage_b ? b_len * 8 : 0);
    LMC_CUDA(cudaFuncSetAttribute(cmc_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cfg.gridDim = dim3(static_cast<unsigned>(n_run * cluster));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // a cluster that cannot be co-scheduled on this device (GPC shape, MIG slice) is halved until it can
    int active = 0;
    const cudaError_t q = cudaOccupancyMaxActiveClusters(&active, cmc_run_kernel, &cfg);
    if (q == cudaSuccess && active > 0) break;
    (void)cudaGetLastError();
    if (cluster == 1) { LMC_CUDA(q); throw std::runtime_error("cmc_run_kernel does not fit on this device"); }
  }
  time_begin();
  LMC_CUDA(cudaLaunchKernelEx(&cfg, cmc_run_kernel, lat, tab, d_occ + static_cast<int64_t>(first_walker) * lat.padded_size, lat.padded_size,
                              d_cmc_mirror + static_cast<size_t>(first_walker) * lat.num_sites,
                              d_cmc_marks + static_cast<size_t>(first_walker) * lat.padded_size, st_run,
                              static_cast<const double *>(d_cmc_temperature + first_walker), static_cast<uint64_t>(params.seed),
                              static_cast<unsigned long long>(target), rp,
                              static_cast<unsigned long long>(std::max<int64_t>(0, n_replay)), stage_b));
  time_end();
  LMC_CUDA(cudaGetLastError());
  if (replaying) {
    const size_t n = static_cast<size_t>(n_replay);
    if (dE) LMC_CUDA(cudaMemcpyAsync(dE, rp.dE, n * 8, cudaMemcpyDeviceToHost, stream));
    if (energy_before) LMC_CUDA(cudaMemcpyAsync(energy_before, rp.energy_before, n * 8, cudaMemcpyDeviceToHost, stream));
    if (temperature_before) LMC_CUDA(cudaMemcpyAsync(temperature_before, rp.temperature_before, n * 8, cudaMemcpyDeviceToHost, stream));
    if (accepted) LMC_CUDA(cudaMemcpyAsync(accepted, rp.accepted, n, cudaMemcpyDeviceToHost, stream));
  }
  std::vector<int32_t> err(nw);
  LMC_CUDA(cudaMemcpyAsync(err.data(), d_cmc_error, nw * 4, cudaMemcpyDeviceToHost, stream));
  LMC_CUDA(cudaStreamSynchronize(stream));
  for (size_t w = 0; w < nw; ++w)
    if (err[w]) {
      cmc_ready = false;
      if (err[w] & kErrBadSite) throw std::invalid_argument("CMC replica " + std::to_string(w) + ": lattice id out of range in the trial stream");
      throw std::out_of_range("CMC replica " + std::to_string(w) + ": Cluster not found in ClusterIndexer (two vacancies within interaction range)");
    }
}

void Engine::cmc_get_state(double *energy, int64_t *steps, int64_t *accepted, double *temperature) {
  require_device();
  if (!d_cmc_energy) throw std::invalid_argument("lmc_cmc_reset has not been called");
  const size_t nw = static_cast<size_t>(n_walkers);
  if (energy) LMC_CUDA(cudaMemcpyAsync(energy, d_cmc_energy, nw * 8, cudaMemcpyDeviceToHost, stream));
  if (steps) LMC_CUDA(cudaMemcpyAsync(steps, d_cmc_steps, nw * 8, cudaMemcpyDeviceToHost, stream));
  if (accepted) LMC_CUDA(cudaMemcpyAsync(accepted, d_cmc_accepted, nw * 8, cudaMemcpyDeviceToHost, stream));
  std::vector<SaSchedule> sa(nw);
  std::vector<double> temps(nw);
  if (temperature) {
    LMC_CUDA(cudaMemcpyAsync(sa.data(), d_cmc_sa, nw * sizeof(SaSchedule), cudaMemcpyDeviceToHost, stream));
    LMC_CUDA(cudaMemcpyAsync(temps.data(), d_cmc_temperature, nw * 8, cudaMemcpyDeviceToHost, stream));
  }
  LMC_CUDA(cudaStreamSynchronize(stream));
  if (temperature)
    for (size_t w = 0; w < nw; ++w) temperature[w] = sa[w].enabled ? sa[w].temperature : temps[w];
}

// ------------------------------------------------------------------------------------------------ host geometry
int64_t Engine::wrapped_id(int x, int y, int z) const {
  auto w = [](int v, int p) { v %= p; return v < 0 ? v + p : v; };
  return lat.id_of_coords(w(x, 2 * lat.fx), w(y, 2 * lat.fy), w(z, 2 * lat.fz));
}

void Engine::neighbors(int32_t shell, int64_t site, int64_t *out) const {
  if (site < 0 || site >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  const Geometry &g = geometry();
  int x, y, z;
  lat.coords_of_id(site, x, y, z);
  std::vector<int64_t> ids;
  auto push = [&](const Int3 &o) { ids.push_back(wrapped_id(x + o.x, y + o.y, z + o.z)); };
  if (shell == 1) for (const auto &o : g.nn1) push(o);
  else if (shell == 2) for (const auto &o : g.nn2) push(o);
  else if (shell == 3) for (const auto &o : g.nn3) push(o);
  else throw std::invalid_argument("shell must be 1, 2 or 3");
  std::sort(ids.begin(), ids.end());   // adjacency lists are sorted ascending (cfg/src/Config.cpp:1036-1044)
  std::copy(ids.begin(), ids.end(), out);
}

int Engine::host_direction(int64_t i, int64_t j, int *xyz_i) const {
  if (i < 0 || j < 0 || i >= lat.num_sites || j >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  const Geometry &g = geometry();
  int xi, yi, zi, xj, yj, zj;
  lat.coords_of_id(i, xi, yi, zi);
  lat.coords_of_id(j, xj, yj, zj);
  auto md = [](int d, int p) { d %= p; if (d > p / 2) d -= p; if (d < -p / 2) d += p; return d; };
  const Int3 d{md(xj - xi, 2 * lat.fx), md(yj - yi, 2 * lat.fy), md(zj - zi, 2 * lat.fz)};
  if (xyz_i) { xyz_i[0] = xi; xyz_i[1] = yi; xyz_i[2] = zi; }
  for (int k = 0; k < 12; ++k)
    if (g.nn1[k] == d) return k;
  throw std::out_of_range("Neighbor not found for lattice pair in GetPairFlatIndex");   // reference message (:281-283)
}

int Engine::host_frame_flag(const int *xyz, int k) const {
  const Int3 p = geometry().frame_p[k];
  return wrapped_id(xyz[0] + p.x, xyz[1] + p.y, xyz[2] + p.z) < wrapped_id(xyz[0] - p.x, xyz[1] - p.y, xyz[2] - p.z) ? 0 : 1;
}

void Engine::pair_lists(int64_t i, int64_t j, int64_t *state, int64_t *mmm, int64_t *mm2, int64_t *mm2b) const {
  const Geometry &g = geometry();
  int ci[3], cj[3];
  const int k = host_direction(i, j, ci);
  const int kb = host_direction(j, i, cj);
  const int sf = host_frame_flag(ci, k), sb = host_frame_flag(cj, kb);
  int64_t ids[60];
  for (int t = 0; t < 60; ++t) {
    const Int3 o = g.pair_offsets[k][sf][t];
    ids[t] = wrapped_id(ci[0] + o.x, ci[1] + o.y, ci[2] + o.z);
  }
  int state_pos_of_env[58];
  for (int t = 0; t < 60; ++t)
    if (g.env_of_state[t] >= 0) state_pos_of_env[g.env_of_state[t]] = t;
  const Int3 pf = g.frame_p[k], pb = g.frame_p[kb];
  const int fs = sf ? -1 : 1, bs = sb ? -1 : 1;
  const bool same = fs * pf.x == bs * pb.x && fs * pf.y == bs * pb.y && fs * pf.z == bs * pb.z;
  for (int t = 0; t < 60 && state; ++t) state[t] = ids[t];
  for (int a = 0; a < 58; ++a) {
    if (mmm) mmm[a] = ids[state_pos_of_env[g.env_of_mmm[a]]];
    if (mm2) mm2[a] = ids[state_pos_of_env[g.env_of_mm2[a]]];
    if (mm2b) mm2b[a] = ids[state_pos_of_env[g.env_of_mm2_backward[same ? 1 : 0][a]]];
  }
}

void Engine::site_list(int64_t site, int64_t *state43) const {
  if (site < 0 || site >= lat.num_sites) throw std::invalid_argument("lattice id out of range");
  const Geometry &g = geometry();
  int x, y, z;
  lat.coords_of_id(site, x, y, z);
  for (int t = 0; t < 43; ++t) state43[t] = wrapped_id(x + g.site_offsets[t].x, y + g.site_offsets[t].y, z + g.site_offsets[t].z);
}

}  // namespace lmc

// ================================================================================================ C ABI
using lmc::Engine;
using lmc::guard;

struct lmc_engine {
  std::unique_ptr<Engine> impl;
};

extern "C" {

const char *lmc_last_error(void) { return lmc::g_last_error.c_str(); }
int lmc_abi_version(void) { return 1; }
int lmc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int lmc_engine_create(lmc_engine **out, const int32_t factors[3], int32_t id_order, const int32_t *element_set,
                      int32_t n_elements, int32_t solvent, int32_t n_walkers, int32_t device) {
  return guard([&] {
    if (!out || !factors || !element_set) throw std::invalid_argument("null argument");
    auto e = std::make_unique<lmc_engine>();
    e->impl = std::make_unique<Engine>(factors, id_order, element_set, n_elements, solvent, n_walkers, device);
    *out = e.release();
  });
}
void lmc_engine_destroy(lmc_engine *engine) { delete engine; }
int64_t lmc_engine_num_sites(const lmc_engine *engine) { return engine ? engine->impl->lat.num_sites : 0; }
int32_t lmc_engine_num_walkers(const lmc_engine *engine) { return engine ? engine->impl->n_walkers : 0; }

int lmc_engine_load_coefficients(lmc_engine *engine, const char *json_path) {
  return guard([&] {
    if (!engine || !json_path) throw std::invalid_argument("null argument");
    engine->impl->load_coefficients(json_path);
  });
}
int lmc_engine_load_coefficients_model(lmc_engine *engine, const char *json_path, int32_t model) {
  return guard([&] {
    if (!engine || !json_path) throw std::invalid_argument("null argument");
    if (model != LMC_BARRIER_QUARTIC && model != LMC_BARRIER_E0) throw std::invalid_argument("unknown barrier model");
    engine->impl->load_coefficients(json_path, model);
  });
}
int lmc_engine_set_occupancy(lmc_engine *engine, int32_t walker, const uint8_t *occupancy, int64_t n) {
  return guard([&] { engine->impl->set_occupancy(walker, occupancy, n, 1); });
}
int lmc_engine_get_occupancy(lmc_engine *engine, int32_t walker, uint8_t *occupancy, int64_t n) {
  return guard([&] { engine->impl->get_occupancy(walker, occupancy, n, 1); });
}
int lmc_engine_set_occupancy_all(lmc_engine *engine, const uint8_t *occupancy, int64_t n_total) {
  return guard([&] { engine->impl->set_occupancy(0, occupancy, n_total, engine->impl->n_walkers); });
}
int lmc_engine_get_occupancy_all(lmc_engine *engine, uint8_t *occupancy, int64_t n_total) {
  return guard([&] { engine->impl->get_occupancy(0, occupancy, n_total, engine->impl->n_walkers); });
}
int lmc_engine_lattice_jump(lmc_engine *engine, int32_t walker, int64_t site_a, int64_t site_b) {
  return guard([&] { engine->impl->lattice_jump(walker, site_a, site_b); });
}

int lmc_eval_barriers(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_i, const int64_t *site_j,
                      double *Ea, double *dE, double *D, double *Ks) {
  return guard([&] { engine->impl->eval_barriers(n, walker, site_i, site_j, Ea, dE, D, Ks); });
}
int lmc_eval_barriers_dev(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_i,
                          const int64_t *site_j, double *Ea, double *dE, double *D, double *Ks) {
  return guard([&] { engine->impl->eval_barriers_dev(n, walker, site_i, site_j, Ea, dE, D, Ks); });
}
int lmc_eval_swap_de(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_a, const int64_t *site_b,
                     double *dE) {
  return guard([&] { engine->impl->eval_swap_de(n, walker, site_a, site_b, dE); });
}
int lmc_eval_pair_de(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_a, const int64_t *site_b,
                     double *dE) {
  return guard([&] { engine->impl->eval_swap_de(n, walker, site_a, site_b, dE, true); });
}
int lmc_engine_barrier_model(const lmc_engine *engine) {
  if (!engine || !engine->impl->has_coefficients || !engine->impl->pair_tables.has_barrier) return -1;
  return engine->impl->pair_tables.model;
}
int lmc_eval_swap_de_dev(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_a,
                         const int64_t *site_b, double *dE) {
  return guard([&] { engine->impl->eval_swap_de_dev(n, walker, site_a, site_b, dE); });
}
int lmc_eval_site_de(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site, const uint8_t *new_element,
                     double *dE) {
  return guard([&] { engine->impl->eval_site_de(n, walker, site, new_element, dE); });
}
int lmc_engine_get_elements(lmc_engine *engine, int32_t walker, int64_t n, const int64_t *lattice_ids, uint8_t *elements) {
  return guard([&] { engine->impl->get_elements(walker, n, lattice_ids, elements); });
}
int64_t lmc_engine_find_element(lmc_engine *engine, int32_t walker, int32_t element, int64_t *count) {
  int64_t first = -1;
  const int rc = guard([&] { first = engine->impl->find_element(walker, element, count); });
  return rc == LMC_OK ? first : rc - 1000;       // errors as values below -1 (-1 = no such site)
}
const char *lmc_engine_coefficients_path(const lmc_engine *engine) { return engine ? engine->impl->coefficients_path.c_str() : ""; }
int lmc_energy_of_cluster(lmc_engine *engine, int32_t walker, const int64_t *lattice_ids, int64_t n, double *energy, int64_t *counts,
                          int32_t n_types) {
  return guard([&] {
    if (n < 0 || (n > 0 && !lattice_ids)) throw std::invalid_argument("bad site list");
    static const int64_t none = 0;
    const double e = engine->impl->cluster_energy(walker, n > 0 ? lattice_ids : &none, n, counts, n_types);
    if (energy) *energy = e;
  });
}
int lmc_energy_encode(lmc_engine *engine, int32_t walker, const int64_t *lattice_ids, int64_t n, double *encode, int32_t n_types) {
  return guard([&] {
    auto &impl = *engine->impl;
    if (!encode || n_types != impl.tab.n_types) throw std::invalid_argument("encode buffer must hold n_types entries");
    std::vector<int64_t> counts(static_cast<size_t>(n_types));
    static const int64_t none = 0;
    if (lattice_ids || n == 0) impl.cluster_energy(walker, n > 0 ? lattice_ids : &none, n, counts.data(), n_types);
    else impl.cluster_energy(walker, nullptr, -1, counts.data(), n_types);
    const auto types = lmc::cluster_types(impl.species);
    for (int32_t q = 0; q < n_types; ++q) encode[q] = static_cast<double>(counts[static_cast<size_t>(q)]) / lmc::energy_cluster_counter(types[static_cast<size_t>(q)].label);
  });
}
int32_t lmc_chemical_potential(lmc_engine *engine, int32_t solvent_element, int32_t *elements, double *mu, int32_t capacity) {
  int32_t count = -1;
  const int rc = guard([&] {
    auto &impl = *engine->impl;
    impl.require_device();
    impl.require_coefficients();
    // pred/src/EnergyPredictor.cpp:196-214: 15 x 15 x 15 cells of pure solvent, atom 0 replaced by each other element (and X)
    std::vector<int32_t> set;
    for (int c = 0; c < impl.species.n; ++c) set.push_back(impl.species.enum_of_code[static_cast<size_t>(c)]);
    if (std::find(set.begin(), set.end(), solvent_element) == set.end()) throw std::invalid_argument("solvent element is not in the element set");
    const int32_t f15[3] = {15, 15, 15};
    lmc::Engine ref(f15, LMC_ID_ORDER_GENERATE, set.data(), static_cast<int32_t>(set.size()), solvent_element, 1, impl.device);
    ref.load_coefficients(impl.coefficients, impl.pair_tables.model);
    std::vector<uint8_t> occ(static_cast<size_t>(ref.lat.num_sites), static_cast<uint8_t>(solvent_element));
    ref.set_occupancy(0, occ.data(), static_cast<int64_t>(occ.size()), 1);
    const double e_solvent = ref.total_energy(0, nullptr, 0);
    std::vector<int32_t> all(set);
    all.push_back(0);                                      // ElementName::X
    std::sort(all.begin(), all.end(), [](int a, int b) { return std::string(lmc::element_name(a)) < std::string(lmc::element_name(b)); });
    count = static_cast<int32_t>(all.size());
    if (!elements || !mu || capacity < count) return;      // query of the length
    for (int32_t q = 0; q < count; ++q) {
      elements[q] = all[static_cast<size_t>(q)];
      if (all[static_cast<size_t>(q)] == solvent_element) { mu[q] = 0.0; continue; }
      occ[0] = static_cast<uint8_t>(all[static_cast<size_t>(q)]);
      ref.set_occupancy(0, occ.data(), static_cast<int64_t>(occ.size()), 1);
      mu[q] = ref.total_energy(0, nullptr, 0) - e_solvent;
      occ[0] = static_cast<uint8_t>(solvent_element);
    }
  });
  return rc == LMC_OK ? count : rc;
}
int lmc_total_energy(lmc_engine *engine, int32_t walker, double *energy, int64_t *counts, int32_t n_types) {
  return guard([&] {
    const double e = engine->impl->total_energy(walker, counts, n_types);
    if (energy) *energy = e;
  });
}
void *lmc_engine_cuda_stream(lmc_engine *engine) { return engine ? static_cast<void *>(engine->impl->stream) : nullptr; }
int lmc_engine_synchronize(lmc_engine *engine) {
  return guard([&] {
    engine->impl->require_device();
    if (cudaStreamSynchronize(engine->impl->stream) != cudaSuccess) throw std::runtime_error("cudaStreamSynchronize failed");
  });
}
double lmc_engine_last_kernel_ms(lmc_engine *engine) {
  double ms = -1.0;
  guard([&] { ms = engine->impl->last_kernel_ms(); });
  return ms;
}
int64_t lmc_engine_launch_count(const lmc_engine *engine) { return engine ? engine->impl->launch_count : 0; }
int lmc_kmc_last_launch_lanes(const lmc_engine *engine) { return engine ? engine->impl->kmc_team_lanes : 0; }
int lmc_kmc_last_launch_handoff(const lmc_engine *engine) { return engine && engine->impl->kmc_team_lanes == 0 && engine->impl->kmc_handoff ? 1 : 0; }
int lmc_kmc_last_launch_resident_occupancy(const lmc_engine *engine) {
  return engine && engine->impl->kmc_team_lanes > 0 && engine->impl->kmc_team_resident_occ ? 1 : 0;
}
int lmc_eval_vacancy_events(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *vacancy_site, int64_t *neighbour_site,
                            double *Ea, double *dE) {
  return guard([&] { engine->impl->eval_vacancy_events(n, walker, vacancy_site, neighbour_site, Ea, dE); });
}
int lmc_eval_vacancy_events_dev(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *vacancy_site, int64_t *neighbour_site,
                                double *Ea, double *dE) {
  return guard([&] {
    if (!vacancy_site || !neighbour_site || !Ea || !dE) throw std::invalid_argument("null device pointer");
    engine->impl->eval_vacancy_events_dev(n, walker, vacancy_site, neighbour_site, Ea, dE);
  });
}
int lmc_kmc_reset(lmc_engine *engine) {
  return guard([&] { engine->impl->kmc_reset(); });
}
int lmc_kmc_run(lmc_engine *engine, const lmc_kmc_params *params, int64_t n_steps, const double *replay_u1,
                const double *replay_u2, const lmc_kmc_trace *trace) {
  return guard([&] {
    if (!params) throw std::invalid_argument("null params");
    engine->impl->kmc_run(*params, n_steps, replay_u1, replay_u2, trace, false);
  });
}
int lmc_kmc_chain_run(lmc_engine *engine, const lmc_kmc_params *params, int64_t n_steps, const double *replay_u,
                      const lmc_kmc_trace *trace) {
  return guard([&] {
    if (!params) throw std::invalid_argument("null params");
    engine->impl->kmc_run(*params, n_steps, nullptr, replay_u, trace, true);
  });
}
int lmc_kmc_set_state(lmc_engine *engine, const double *time, const double *energy, const int64_t *steps) {
  return guard([&] { engine->impl->kmc_set_state(time, energy, steps); });
}
int lmc_kmc_get_state(lmc_engine *engine, double *time, double *energy, int64_t *steps, int64_t *vacancy, double *temperature) {
  return guard([&] { engine->impl->kmc_get_state(time, energy, steps, vacancy, temperature); });
}
int lmc_cmc_reset(lmc_engine *engine, double sa_initial_temperature, uint64_t sa_maximum_steps) {
  return guard([&] { engine->impl->cmc_reset(sa_initial_temperature, sa_maximum_steps); });
}
int lmc_cmc_run(lmc_engine *engine, const lmc_cmc_params *params, int64_t n_trials) {
  return guard([&] {
    if (!params) throw std::invalid_argument("null params");
    engine->impl->cmc_run(*params, n_trials, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  });
}
int lmc_cmc_grid_run(lmc_engine *engine, const lmc_cmc_params *params, int64_t n_trials) {
  return guard([&] {
    if (!params) throw std::invalid_argument("null params");
    engine->impl->cmc_grid_run(*params, n_trials);
  });
}
int lmc_cmc_domain_run(lmc_engine *engine, const lmc_cmc_params *params, const lmc_cmc_domain_params *domain, int64_t n_trials) {
  return guard([&] {
    if (!params) throw std::invalid_argument("null params");
    engine->impl->cmc_domain_run(*params, domain, n_trials);
  });
}
int lmc_cmc_domain_handles(lmc_engine *engine, void *handles192) {
  return guard([&] { engine->impl->cmc_domain_handles(handles192); });
}
int lmc_cmc_domain_attach_peers(lmc_engine *engine, int32_t rank, int32_t world, const void *handles) {
  return guard([&] { engine->impl->cmc_domain_attach_peers(rank, world, handles); });
}
int lmc_cmc_domain_last_shape(const lmc_engine *engine, int32_t *shape6) {     // seven entries
  if (!engine || !shape6) return LMC_ERR_INVALID_ARGUMENT;
  const auto &e = *engine->impl;
  shape6[0] = e.dom_last_edge; shape6[1] = e.dom_last_domains; shape6[2] = e.dom_last_lanes; shape6[3] = e.dom_last_threads;
  shape6[4] = e.dom_last_ctas; shape6[5] = e.dom_last_rounds; shape6[6] = e.dom_last_spec;
  return LMC_OK;
}
int lmc_cmc_exchange_handle(lmc_engine *engine, void *handle64) {
  return guard([&] { engine->impl->cmc_exchange_handle(handle64); });
}
int lmc_cmc_attach_peers(lmc_engine *engine, int32_t rank, int32_t world, const void *handles, int32_t grid_ctas) {
  return guard([&] { engine->impl->cmc_attach_peers(rank, world, handles, grid_ctas); });
}
int lmc_cmc_replay(lmc_engine *engine, int32_t walker, const lmc_cmc_params *params, int64_t n, const int64_t *site_a,
                   const int64_t *site_b, const double *u, double *dE, double *energy_before, double *temperature_before,
                   uint8_t *accepted) {
  return guard([&] {
    if (!params) throw std::invalid_argument("null params");
    if (n <= 0) return;
    engine->impl->cmc_run(*params, 0, walker, n, site_a, site_b, u, dE, energy_before, temperature_before, accepted);
  });
}
int lmc_cmc_get_state(lmc_engine *engine, double *energy, int64_t *steps, int64_t *accepted, double *temperature) {
  return guard([&] { engine->impl->cmc_get_state(energy, steps, accepted, temperature); });
}
int lmc_debug_pair(lmc_engine *engine, int32_t walker, int64_t site_i, int64_t site_j, int64_t *state60, int64_t *mmm58,
                   int64_t *mm2_58, int64_t *mm2_backward58, int32_t *start_counts, int32_t *end_counts, int32_t *enc_mmm,
                   int32_t *enc_mm2_forward, int32_t *enc_mm2_backward) {
  return guard([&] {
    engine->impl->debug_pair(walker, site_i, site_j, state60, mmm58, mm2_58, mm2_backward58, start_counts, end_counts, enc_mmm,
                             enc_mm2_forward, enc_mm2_backward);
  });
}
int lmc_debug_site(lmc_engine *engine, int32_t walker, int64_t site, int32_t new_element, int64_t *state43,
                   int32_t *start_counts, int32_t *end_counts) {
  return guard([&] { engine->impl->debug_site(walker, site, new_element, state43, start_counts, end_counts); });
}

int lmc_engine_neighbors(const lmc_engine *engine, int32_t shell, int64_t site, int64_t *out) {
  return guard([&] { engine->impl->neighbors(shell, site, out); });
}
int lmc_engine_kmc_table_grid_bits(const lmc_engine *engine, int32_t bits[2]) {
  return guard([&] {
    engine->impl->require_coefficients();
    bits[0] = engine->impl->pair_grid_bits[0];
    bits[1] = engine->impl->pair_grid_bits[1];
  });
}
int lmc_engine_kmc_event_order(const lmc_engine *engine, int64_t site, int64_t *out) {
  return guard([&] { engine->impl->kmc_event_order(site, out); });
}
int lmc_engine_site_coords(const lmc_engine *engine, int64_t site, int32_t xyz[3]) {
  return guard([&] {
    if (site < 0 || site >= engine->impl->lat.num_sites) throw std::invalid_argument("lattice id out of range");
    int x, y, z;
    engine->impl->lat.coords_of_id(site, x, y, z);
    xyz[0] = x; xyz[1] = y; xyz[2] = z;
  });
}
int lmc_engine_pair_lists(const lmc_engine *engine, int64_t site_i, int64_t site_j, int64_t *state60, int64_t *mmm58,
                          int64_t *mm2_58, int64_t *mm2_backward58) {
  return guard([&] { engine->impl->pair_lists(site_i, site_j, state60, mmm58, mm2_58, mm2_backward58); });
}
int lmc_engine_site_list(const lmc_engine *engine, int64_t site, int64_t *state43) {
  return guard([&] { engine->impl->site_list(site, state43); });
}

int64_t lmc_tables_mapping(int32_t which, int64_t *out, int64_t capacity) {
  int64_t needed = -1;
  const int rc = guard([&] {
    const lmc::Geometry &g = lmc::geometry();
    std::vector<int64_t> flat;
    auto emit_groups = [&](const std::vector<std::vector<std::vector<int64_t>>> &groups) {
      flat.push_back(static_cast<int64_t>(groups.size()));
      for (const auto &grp : groups) {
        flat.push_back(static_cast<int64_t>(grp.size()));
        flat.push_back(grp.empty() ? 0 : static_cast<int64_t>(grp[0].size()));
        for (const auto &c : grp) flat.insert(flat.end(), c.begin(), c.end());
      }
    };
    std::vector<std::vector<std::vector<int64_t>>> groups;
    if (which == 0 || which == 3) {
      const auto &cl = which == 0 ? g.state_pair : g.state_site;
      groups.resize(8);
      for (const auto &c : cl) {
        std::vector<int64_t> v;
        for (int q = 0; q < c.arity; ++q) v.push_back(c.pos[q]);
        groups[c.label].push_back(v);
      }
    } else if (which == 1 || which == 2) {
      const auto &cl = which == 1 ? g.mmm : g.mm2;
      groups.resize(which == 1 ? g.n_groups_mmm : g.n_groups_mm2);
      for (const auto &c : cl) {
        std::vector<int64_t> v;
        if (c.symmetric) v.push_back(-1);
        for (int q = 0; q < c.arity; ++q) v.push_back(c.pos[q]);
        groups[c.group].push_back(v);
      }
    } else {
      throw std::invalid_argument("which must be 0..3");
    }
    emit_groups(groups);
    needed = static_cast<int64_t>(flat.size());
    if (out) std::copy(flat.begin(), flat.begin() + std::min<int64_t>(needed, capacity), out);
  });
  return rc == LMC_OK ? needed : rc;
}

int32_t lmc_tables_cluster_types(const int32_t *element_set, int32_t n_elements, int32_t *rows5, int32_t capacity_rows) {
  int32_t count = -1;
  const int rc = guard([&] {
    const lmc::Species sp = lmc::make_species(element_set, n_elements, 0);
    const auto types = lmc::cluster_types(sp);
    count = static_cast<int32_t>(types.size());
    for (int32_t r = 0; rows5 && r < count && r < capacity_rows; ++r) {
      rows5[5 * r] = types[r].label;
      rows5[5 * r + 1] = types[r].arity;
      for (int q = 0; q < 3; ++q) rows5[5 * r + 2 + q] = q < types[r].arity ? sp.enum_of_code[types[r].code[q]] : -1;
    }
  });
  return rc == LMC_OK ? count : rc;
}

int32_t lmc_tables_group_sizes(int32_t which, int32_t n_elements, int32_t *sizes, int32_t capacity) {
  int32_t len = -1;
  const int rc = guard([&] {
    if (which != 1 && which != 2) throw std::invalid_argument("which must be 1 (mmm) or 2 (mm2)");
    int L = 0;
    const auto groups = lmc::group_layout(lmc::geometry(), which == 2, n_elements, &L);
    len = L;
    if (sizes)
      for (const auto &gi : groups)
        for (int q = 0; q < gi.length && gi.offset + q < capacity; ++q) sizes[gi.offset + q] = gi.size;
  });
  return rc == LMC_OK ? len : rc;
}

int64_t lmc_engine_get_tables(const lmc_engine *engine, int32_t which, double *out, int64_t capacity) {
  int64_t len = -1;
  const int rc = guard([&] {
    engine->impl->require_coefficients();
    const Engine &e = *engine->impl;
    const std::vector<double> *v = nullptr;
    switch (which) {
      case 0: v = &e.pair_tables.C; break;
      case 1: v = &e.pair_tables.A; break;
      case 2: v = &e.pair_tables.B; break;
      case 3: v = &e.site_tables.C; break;
      case 4: v = &e.site_tables.A; break;
      case 5: v = &e.site_tables.B; break;
      case 6: v = &e.coefficients.base_theta; break;
      case 7: v = &e.folded_C; break;
      case 8: v = &e.folded_A; break;
      case 9: v = &e.folded_B; break;
      default: throw std::invalid_argument("which must be 0..9");
    }
    len = static_cast<int64_t>(v->size());
    if (out) std::copy(v->begin(), v->begin() + std::min<int64_t>(len, capacity), out);
  });
  return rc == LMC_OK ? len : rc;
}

int32_t lmc_tables_env_pairs(int32_t which, int16_t *pairs, int32_t capacity_pairs) {
  int32_t count = -1;
  const int rc = guard([&] {
    if (which != 0 && which != 1) throw std::invalid_argument("which must be 0 or 1");
    const auto &v = which == 0 ? lmc::geometry().env_pairs : lmc::geometry().site_env_pairs;
    count = static_cast<int32_t>(v.size());
    for (int32_t p = 0; pairs && p < count && p < capacity_pairs; ++p) {
      pairs[2 * p] = v[p][0];
      pairs[2 * p + 1] = v[p][1];
    }
  });
  return rc == LMC_OK ? count : rc;
}

}  // extern "C"
