// sb_api.cpp -- Grid / Model host objects and the extern "C" ABI declared in include/scythe_b200.h.
//
// All state is device-resident for the whole integration (physical slots, spectral B/A, AB3
// history); the host only enqueues kernels and (multi-GPU) one NCCL all-reduce per step.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/scythe_b200.h"
#include "sb_internal.hpp"

using namespace sb;

namespace sb { int sb_rows_per_cta(int L, bool fast); }

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct DomainErr : std::runtime_error { using std::runtime_error::runtime_error; };
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };
struct NanFound : std::runtime_error { using std::runtime_error::runtime_error; };

#define CU(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return SB_OK;
  } catch (const DomainErr& e) { return fail(SB_EDOMAIN, e.what());
  } catch (const Unsupported& e) { return fail(SB_EUNSUPPORTED, e.what());
  } catch (const NanFound& e) { return fail(SB_ENAN, e.what());
  } catch (const CudaError& e) { return fail(SB_ECUDA, e.what());
  } catch (const std::invalid_argument& e) { return fail(SB_EINVAL, e.what());
  } catch (const std::exception& e) { return fail(SB_ECUDA, e.what()); }
}

template <class T>
static T* dev_upload(const std::vector<T>& h) {
  T* d = nullptr;
  CU(cudaMalloc((void**)&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CU(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static double* dev_zeros(long long n, cudaStream_t s) {
  double* d = nullptr;
  CU(cudaMalloc((void**)&d, (size_t)std::max<long long>(n, 1) * sizeof(double)));
  CU(cudaMemsetAsync(d, 0, (size_t)std::max<long long>(n, 1) * sizeof(double), s));
  return d;
}

// ====================================================================================== grid
struct sb_grid {
  sb_grid_params gp{};
  std::vector<int32_t> bcl, bcr, bcb, bct;
  int device = 0;
  cudaStream_t stream = nullptr;
  DevGrid dg{};
  std::vector<int> ring_n, ring_ri, h2r;
  std::vector<long long> hoff, woff;
  std::vector<double> rad;
  ChebTables cheb;
  std::vector<SplineFactor> factors;
  std::vector<DevSplineFactor> hfactors;
  DevSplineFactor* d_factors = nullptr;   // device copy of hfactors ([V]) for the merged spline solve
  std::vector<void*> owned;   // device allocations freed at destroy
  double* physical = nullptr;
  const double* slot0_src = nullptr;   // calcTendency's `physical .= var_np1` (src/semiimplicit.jl:731) deferred until slot 0 is read
  double* spectralB = nullptr;
  bool owns_B = true;                  // false: spectralB aliases the patch's shared B (single-tile model)
  double* spectralA = nullptr;
  double* scratch = nullptr;
  long long scratch_doubles = 0;
  int vchunk = 1;
  ZTile* d_ztiles = nullptr;
  int nztiles = 0;
  std::vector<FftClass> classes;
  std::vector<std::vector<LWork>> fwork, iwork;
  std::vector<const LWork*> d_fwork, d_iwork;
  std::vector<std::vector<LWork>> fwork2, iwork2;     // v2 persistent ring-FFT items (bigger row ranges)
  std::vector<const LWork*> d_fwork2, d_iwork2;
  double* d_fft3_scratch = nullptr;                    // parking area of the composite-length forward FFT
  const PeerScatter* scatter = nullptr;                // multi-GPU: fwd_r also stores into the plane owners' buffers
  std::vector<const double*> d_tw, d_twp;
  RingPlan* d_plans = nullptr;
  double* d_blob = nullptr;
  double* d_fwdT = nullptr;
  double* d_invM = nullptr;
  double* d_parM = nullptr;     // parity tables for the BC-free fast Chebyshev inverse (FMA form)
  double* d_parB = nullptr;     // DMMA B-fragment tables of the same matrices
  double* d_fwdB = nullptr;     // DMMA B-fragment tables of the analysis (forward) matrix
  std::vector<char> z_bcfree;   // per variable: BCB == BCT == R0
  std::vector<int> zt_first;    // [rDim+1] first z tile of every ring (tiles are ordered by ring)
  SmallRings small;             // merged launch of the generic-kernel ring classes
  std::vector<SmallCls> h_smallcls;
  // overlapped step (tile_step_overlapped): FP64-bound ring FFTs and HBM-bound kernels on two streams, by ring batches
  struct Overlap {
    bool ready = false;
    cudaStream_t s_fft = nullptr, s_mem = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_invr = nullptr, ev_fl = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> ev_il, ev_z;
    std::vector<std::pair<int, int>> batches;   // ring ranges [r_lo, r_hi), outermost first
    int* d_counters = nullptr;
    int counter_next = 0;
    double* buf = nullptr;
    long long buf_doubles = 0;
    int sm_reserve = 24;
  } ov;
  long long launches = 0;
  long long* d_nan = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int ndims = 1;

  Profiler prof;
  LaunchCtx ctx() {
    LaunchCtx c{stream, &launches, &prof};
    static const bool merge_small = !(std::getenv("SB_MERGE_SMALL") && std::atoi(std::getenv("SB_MERGE_SMALL")) == 0);
    if (merge_small && small.cls) c.small = &small;
    // SB_FFT_DYNAMIC=<chunk>: ring FFT work shares from an atomic counter (<chunk> items at a time) instead of equal contiguous
    // shares, on the ordinary one-stream step as well (A/B switch; the overlapped step always uses it)
    static const int dyn = std::getenv("SB_FFT_DYNAMIC") ? std::atoi(std::getenv("SB_FFT_DYNAMIC")) : 0;
    if (dyn > 0) {
      if (!ov.d_counters) {
        if (cudaMalloc((void**)&ov.d_counters, 512 * sizeof(int)) != cudaSuccess) ov.d_counters = nullptr;
      }
      if (ov.d_counters) { c.counters = ov.d_counters; c.counter_next = &ov.counter_next; c.ncounters = 512; c.fft_chunk = dyn; }
    }
    return c;
  }
  template <class T> T* up(const std::vector<T>& h) { T* d = dev_upload(h); owned.push_back(d); return d; }
  long long slot_stride() const { return dg.N * dg.V; }

  void ensure_physical() {
    if (!physical) physical = dev_zeros(dg.N * dg.V * dg.D, stream);
  }
  void materialize_slot0() {
    if (slot0_src && physical)
      CU(cudaMemcpyAsync(physical, slot0_src, (size_t)dg.N * dg.V * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    slot0_src = nullptr;
  }
  void release_physical() {
    if (physical) { CU(cudaStreamSynchronize(stream)); cudaFree(physical); physical = nullptr; }
  }
  // per-variable scratch need (doubles)
  long long fwd_need() const {
    long long sl = (long long)dg.bz * dg.W;
    long long sz = (dg.has_l && dg.has_z) ? (long long)dg.bz * dg.hpoints : 0;
    return (dg.has_l || dg.has_z) ? sl + sz : 0;
  }
  long long inv_need() const {
    long long sl = 3LL * dg.bz * dg.W;
    long long sz = (dg.has_l && dg.has_z) ? 5LL * dg.bz * dg.hpoints : 0;
    return (dg.has_l || dg.has_z) ? sl + sz : 0;
  }
  void ensure_scratch_doubles(long long n) {   // fused K3+K4 keeps the rows of all variables at once
    ensure_scratch();
    if (scratch_doubles >= n) return;
    CU(cudaStreamSynchronize(stream));
    cudaFree(scratch);
    scratch = nullptr;
    scratch_doubles = n;
    CU(cudaMalloc((void**)&scratch, (size_t)scratch_doubles * sizeof(double)));
  }
  void ensure_scratch() {
    if (scratch) return;
    long long per_v = std::max(fwd_need(), inv_need());
    const char* env = std::getenv("SB_VCHUNK");
    vchunk = dg.V;
    // as many variables per pass as fit a 16 GiB scratch (C4: all three -- one launch per stage instead of one per
    // variable: fwd_l 2.89 -> 2.67 ms; the TC boundary-layer set: three of six), SB_SCRATCH_GB / SB_VCHUNK override
    const long long budget = (std::getenv("SB_SCRATCH_GB") ? std::atoll(std::getenv("SB_SCRATCH_GB")) : 16LL) << 30;
    if (env && std::atoi(env) > 0) vchunk = std::min(dg.V, std::atoi(env));
    else if (per_v * dg.V * 8 > budget) vchunk = (int)std::max<long long>(1, std::min<long long>(dg.V, budget / std::max<long long>(per_v * 8, 1)));
    scratch_doubles = std::max<long long>(per_v * vchunk, 1) + 16;   // + alignment slack of the SZ region
    CU(cudaMalloc((void**)&scratch, (size_t)scratch_doubles * sizeof(double)));
  }
};

static void build_grid(sb_grid* G) {
  const sb_grid_params& gp = G->gp;
  DevGrid& d = G->dg;
  if (gp.geometry < SB_GEOM_R || gp.geometry > SB_GEOM_RLZ) throw DomainErr("Unknown geometry");
  if (gp.num_cells < 1 || !(gp.xmax > gp.xmin)) throw std::invalid_argument("num_cells >= 1 and xmax > xmin required");
  if (gp.nvars < 1) throw std::invalid_argument("nvars >= 1 required");
  d.has_l = (gp.geometry == SB_GEOM_RL || gp.geometry == SB_GEOM_RLZ);
  d.has_z = (gp.geometry == SB_GEOM_RZ || gp.geometry == SB_GEOM_RLZ);
  d.V = gp.nvars;
  d.D = 3 + (d.has_l ? 2 : 0) + (d.has_z ? 2 : 0);
  d.num_cells = (int)gp.num_cells;
  d.rDim = 3 * d.num_cells;
  d.b_rDim = d.num_cells + 3;
  if (d.has_z) {
    if (gp.zDim < 4 || !(gp.zmax > gp.zmin)) throw std::invalid_argument("zDim >= 4 and zmax > zmin required");
    d.zDim = (int)gp.zDim;
    long long def = std::min<long long>(gp.zDim, (2 * gp.zDim - 1) / 3 + 1);
    d.bz = (int)(gp.b_zDim > 0 ? gp.b_zDim : def);
    if (d.bz > d.zDim) throw std::invalid_argument("b_zDim > zDim");
  } else {
    d.zDim = 1;
    d.bz = 1;
  }
  d.bzp = (d.bz + 3) & ~3;
  const long long sIL = gp.spectralIndexL > 0 ? gp.spectralIndexL : 1;
  d.coefOffset = (int)(sIL - 1);
  d.patchOffsetL = d.coefOffset * 3;
  d.kDim = d.has_l ? d.rDim + d.patchOffsetL : 0;
  d.ncolp = 1 + 2 * d.kDim;
  const double DX = (gp.xmax - gp.xmin) / gp.num_cells;
  spline_weights(DX, d.phi, d.wq);
  spline_mish_points(gp.xmin, DX, d.num_cells, G->rad);
  G->ring_n.resize(d.rDim);
  G->ring_ri.resize(d.rDim);
  G->hoff.assign(d.rDim + 1, 0);
  G->woff.assign(d.rDim + 1, 0);
  for (int r = 0; r < d.rDim; ++r) {
    int ri = r + 1 + d.patchOffsetL;
    G->ring_ri[r] = ri;
    G->ring_n[r] = d.has_l ? 4 + 4 * ri : 1;
    G->hoff[r + 1] = G->hoff[r] + G->ring_n[r];
    G->woff[r + 1] = G->woff[r] + (d.has_l ? 1 + 2 * ri : 1);
  }
  d.hpoints = G->hoff[d.rDim];
  d.W = G->woff[d.rDim];
  d.N = d.hpoints * d.zDim;
  d.S = (long long)d.bz * d.b_rDim * d.ncolp;
  G->ndims = 1 + (d.has_l ? 1 : 0) + (d.has_z ? 1 : 0);
  G->h2r.resize((size_t)d.hpoints);
  for (int r = 0; r < d.rDim; ++r)
    for (long long h = G->hoff[r]; h < G->hoff[r + 1]; ++h) G->h2r[(size_t)h] = r;

  CU(cudaSetDevice(G->device));
  d.ring_n = G->up(G->ring_n);
  d.ring_ri = G->up(G->ring_ri);
  d.ring_hoff = G->up(G->hoff);
  // ring offsets with every ring padded to whole 16-point blocks (blocked SZ layout, sb_internal.hpp RowDst)
  std::vector<long long> hoffp(G->hoff.size(), 0);
  for (size_t r = 0; r + 1 < G->hoff.size(); ++r) hoffp[r + 1] = hoffp[r] + ((G->hoff[r + 1] - G->hoff[r] + 15) / 16) * 16;
  d.hpointsp = hoffp.back();
  d.ring_hoffp = G->up(hoffp);
  d.ring_woff = G->up(G->woff);
  d.rad = G->up(G->rad);
  d.h2r = G->up(G->h2r);

  // splines: one factor per (BCL,BCR) pair
  std::map<std::pair<int, int>, int> cache;
  G->factors.clear();
  std::vector<int> fidx(d.V);
  for (int v = 0; v < d.V; ++v) {
    auto key = std::make_pair((int)G->bcl[v], (int)G->bcr[v]);
    auto it = cache.find(key);
    if (it == cache.end()) {
      if (key.first < 0 || key.first > 7 || key.second < 0 || key.second > 7) throw std::invalid_argument("bad spline BC code");
      if ((key.first == SB_BC_PERIODIC) != (key.second == SB_BC_PERIODIC)) throw std::invalid_argument("PERIODIC must be set on both ends");
      if (key.first == SB_BC_PERIODIC && d.b_rDim > 4096) throw Unsupported("PERIODIC splines are limited to 4093 cells");
      G->factors.push_back(make_spline_factor(d.num_cells, DX, gp.l_q > 0 ? gp.l_q : 2.0, key.first, key.second));
      it = cache.emplace(key, (int)G->factors.size() - 1).first;
    }
    fidx[v] = it->second;
  }
  std::vector<DevSplineFactor> df(G->factors.size());
  for (size_t i = 0; i < G->factors.size(); ++i) {
    const SplineFactor& f = G->factors[i];
    DevSplineFactor& o = df[i];
    o.M = f.M; o.rL = f.rL; o.rR = f.rR; o.nfree = f.nfree; o.periodic = f.periodic ? 1 : 0;
    o.foldL[0] = f.foldL[0]; o.foldL[1] = f.foldL[1]; o.foldR[0] = f.foldR[0]; o.foldR[1] = f.foldR[1];
    o.chol = f.chol.empty() ? nullptr : G->up(f.chol);
    o.dense = f.dense.empty() ? nullptr : G->up(f.dense);
  }
  G->hfactors.resize(d.V);
  for (int v = 0; v < d.V; ++v) G->hfactors[v] = df[fidx[v]];
  G->d_factors = G->up(G->hfactors);

  // Chebyshev tables
  if (d.has_z) {
    G->cheb = make_cheb_tables(d.zDim, d.bz, gp.zmin, gp.zmax);
    d.zlev = G->up(G->cheb.z);
    std::vector<double> fwdT((size_t)d.zDim * d.bzp, 0.0);
    for (int zb = 0; zb < d.bz; ++zb)
      for (int z = 0; z < d.zDim; ++z) fwdT[(size_t)z * d.bzp + zb] = G->cheb.fwd[(size_t)zb * d.zDim + z];
    G->d_fwdT = G->up(fwdT);
    const int zp = (d.zDim + 3) & ~3, nz = d.zDim, bz = d.bz;
    std::vector<double> invM((size_t)d.V * 3 * bz * zp, 0.0);
    for (int v = 0; v < d.V; ++v) {
      int b = G->bcb.empty() ? 0 : G->bcb[v], t = G->bct.empty() ? 0 : G->bct[v];
      if (b < 0 || b > 3 || t < 0 || t > 3) throw std::invalid_argument("bad Chebyshev BC code");
      std::vector<double> IG = cheb_bc_matrix(G->cheb, b, t);
      const std::vector<double>* T[3] = {&G->cheb.T0, &G->cheb.T1, &G->cheb.T2};
      for (int k = 0; k < 3; ++k)
        for (int z = 0; z < nz; ++z)
          for (int zb = 0; zb < bz; ++zb) {
            long double s = 0;
            for (int q = 0; q < bz; ++q) s += (long double)(*T[k])[(size_t)z * nz + q] * IG[(size_t)q * bz + zb];
            invM[(((size_t)v * 3 + k) * bz + zb) * zp + z] = (double)s;
          }
    }
    G->d_invM = G->up(invM);
    G->z_bcfree.assign(d.V, 0);
    for (int v = 0; v < d.V; ++v) G->z_bcfree[v] = ((G->bcb.empty() || G->bcb[v] == 0) && (G->bct.empty() || G->bct[v] == 0)) ? 1 : 0;
    std::vector<double> parM;
    build_inv_z_par_tables(d.zDim, d.bz, G->cheb.T0.data(), G->cheb.T1.data(), G->cheb.T2.data(), parM);
    G->d_parM = G->up(parM);
    std::vector<double> parB;
    build_inv_z_mma_tables(d.zDim, d.bz, G->cheb.T0.data(), G->cheb.T1.data(), G->cheb.T2.data(), parB);
    G->d_parB = G->up(parB);
    std::vector<double> fwdB;
    build_fwd_z_mma_tables(d.zDim, d.bz, G->cheb.fwd.data(), fwdB);
    G->d_fwdB = G->up(fwdB);
    // z tiles
    std::vector<ZTile> zt;
    if (d.has_l) {
      G->zt_first.assign(d.rDim + 1, 0);
      for (int r = 0; r < d.rDim; ++r) {
        G->zt_first[r] = (int)zt.size();
        for (int j0 = 0; j0 < G->ring_n[r]; j0 += 32) {
          ZTile t{};
          t.hcol0 = (int)(G->hoff[r] + j0);
          t.ncols = std::min(32, G->ring_n[r] - j0);
          t.out_base = (long long)d.bz * G->hoff[r] + j0;
          t.out_stride = G->ring_n[r];
          t.ring = r;
          t.rad = G->rad[r];
          t.blk = (long long)d.bz * (hoffp[r] + j0);
          zt.push_back(t);
        }
      }
      G->zt_first[d.rDim] = (int)zt.size();
    } else {
      for (int j0 = 0; j0 < d.rDim; j0 += 32) {
        ZTile t{};
        t.hcol0 = j0;
        t.ncols = std::min(32, d.rDim - j0);
        t.out_base = j0;
        t.out_stride = d.rDim;
        zt.push_back(t);
      }
    }
    if ((long long)d.hpoints > 0x7fffffffLL) throw Unsupported("more than 2^31 horizontal points per tile");
    G->nztiles = (int)zt.size();
    G->d_ztiles = G->up(zt);
  }
  // ring FFT plans
  if (d.has_l) {
    std::vector<RingPlan> plans;
    std::vector<double> blob;
    build_ring_plans(d.has_z ? 32 : 256, G->ring_ri, G->classes, plans, blob);
    G->d_plans = G->up(plans);
    G->d_blob = G->up(blob);
    G->fwork.assign(G->classes.size(), {});
    G->iwork.assign(G->classes.size(), {});
    for (int r = d.rDim - 1; r >= 0; --r) {
      int nr = sb_rows_per_cta(plans[r].L, G->classes[plans[r].cls].fast);
      for (int row0 = 0; row0 < d.bz; row0 += nr) G->fwork[plans[r].cls].push_back(LWork{r, row0, std::min(nr, d.bz - row0), 0});
      for (int row0 = 0; row0 < 5 * d.bz; row0 += nr) G->iwork[plans[r].cls].push_back(LWork{r, row0, std::min(nr, 5 * d.bz - row0), 0});
    }
    G->fwork2.assign(G->classes.size(), {});
    G->iwork2.assign(G->classes.size(), {});
    for (int r = d.rDim - 1; r >= 0; --r) {
      const int L = plans[r].L, cls = plans[r].cls;
      if (!G->classes[cls].fast) continue;
      // items of (nearly) equal size: ceil(total / k) rows for the k items the target size asks for
      auto even = [](int total, int nr) { const int k = (total + nr - 1) / nr; return (total + k - 1) / k; };
      if (fft2_supported(L, true)) {
        const int nr = even(d.bz, fft2_rows_per_item(L, true));
        for (int row0 = 0; row0 < d.bz; row0 += nr) G->fwork2[cls].push_back(LWork{r, row0, std::min(nr, d.bz - row0), 0});
      }
      if (fft2_supported(L, false)) {
        const int nr = even(5 * d.bz, fft2_rows_per_item(L, false));
        for (int row0 = 0; row0 < 5 * d.bz; row0 += nr) G->iwork2[cls].push_back(LWork{r, row0, std::min(nr, 5 * d.bz - row0), 0});
      }
    }
    {   // merged work lists of the classes the generic kernels serve (not fast, not composite)
      std::vector<LWork> iw, fw;
      std::vector<SmallCls> sc(G->classes.size(), SmallCls{0, 0, nullptr});
      size_t smem = 0;
      for (size_t ci = G->classes.size(); ci-- > 0;) {
        const FftClass& cl = G->classes[ci];
        if (cl.fast || cl.R == 3) continue;
        iw.insert(iw.end(), G->iwork[ci].begin(), G->iwork[ci].end());
        fw.insert(fw.end(), G->fwork[ci].begin(), G->fwork[ci].end());
        smem = std::max(smem, (size_t)2 * sb_rows_per_cta(cl.L, false) * cl.L * 16);
      }
      G->small.niwork = (int)iw.size();
      G->small.nfwork = (int)fw.size();
      G->small.smem = smem;
      if (!iw.empty()) {
        G->small.iwork = G->up(iw);
        G->small.fwork = G->up(fw);
      }
      G->h_smallcls = sc;     // tw pointers are filled in below, once the class tables are on the device
    }
    size_t f3 = 0;
    for (auto& cl : G->classes)
      if (cl.R == 3) f3 = std::max(f3, fft3_scratch_doubles(cl.L));
    if (f3) { G->d_fft3_scratch = dev_zeros((long long)f3, G->stream); G->owned.push_back(G->d_fft3_scratch); }
    for (size_t c = 0; c < G->classes.size(); ++c) {
      G->d_fwork2.push_back(G->up(G->fwork2[c]));
      G->d_iwork2.push_back(G->up(G->iwork2[c]));
      G->d_fwork.push_back(G->up(G->fwork[c]));
      G->d_iwork.push_back(G->up(G->iwork[c]));
      G->d_tw.push_back(G->up(G->classes[c].tw));
      G->d_twp.push_back(G->up(G->classes[c].twp));
    }
    if (G->small.niwork > 0) {
      for (size_t c = 0; c < G->classes.size(); ++c) {
        G->h_smallcls[c].log2L = G->classes[c].log2L;
        G->h_smallcls[c].tw = reinterpret_cast<const double2*>(G->d_tw[c]);
      }
      G->small.cls = G->up(G->h_smallcls);
    }
  }
  G->spectralB = dev_zeros(d.S * d.V, G->stream);
  G->spectralA = dev_zeros(d.S * d.V, G->stream);
  CU(cudaMalloc((void**)&G->d_nan, sizeof(long long)));
  CU(cudaEventCreate(&G->ev0));
  CU(cudaEventCreate(&G->ev1));
}

static void grid_fwd_z(sb_grid* G, int nv, const double* in, double* mir, double* out, long long out_vs);

// forward transform of `in` ([V][N] var-major) into G->spectralB (K1)
static void grid_forward(sb_grid* G, const double* in, double* mirror) {
  DevGrid& d = G->dg;
  LaunchCtx c = G->ctx();
  if (!d.has_l && !d.has_z) {
    if (mirror && mirror != in) launch_copy(c, mirror, in, d.N * d.V);
    launch_fwd_r(c, d, d.V, in, d.N, G->spectralB, d.S, G->scatter, 0);
    return;
  }
  G->ensure_scratch();
  const long long slN = (long long)d.bz * d.W, szN = (long long)d.bz * d.hpoints;
  for (int v0 = 0; v0 < d.V; v0 += G->vchunk) {
    const int nv = std::min(G->vchunk, d.V - v0);
    double* SL = G->scratch;
    double* SZ = G->scratch + ((slN * G->vchunk + 15) & ~15LL);   // 128-byte aligned: rows are read as double2
    const double* inv = in + (long long)v0 * d.N;
    double* mir = mirror ? mirror + (long long)v0 * d.N : nullptr;
    if (d.has_l && d.has_z) {
      grid_fwd_z(G, nv, inv, mir, SZ, szN);
      launch_fwd_l(c, d, G->fwork, G->d_fwork.data(), G->classes, G->d_tw.data(), G->d_twp.data(), G->d_plans, G->d_blob, nv, SZ, szN,
                   1, nullptr, 0, SL, slN, &G->fwork2, G->d_fwork2.data(), G->d_fft3_scratch);
    } else if (d.has_l) {
      launch_fwd_l(c, d, G->fwork, G->d_fwork.data(), G->classes, G->d_tw.data(), G->d_twp.data(), G->d_plans, G->d_blob, nv, inv, d.N,
                   0, mir, d.N, SL, slN, &G->fwork2, G->d_fwork2.data(), G->d_fft3_scratch);
    } else {
      grid_fwd_z(G, nv, inv, mir, SL, slN);
    }
    launch_fwd_r(c, d, nv, SL, slN, G->spectralB + (long long)v0 * d.S, d.S, G->scatter, v0);
  }
}

static void grid_inv_z(sb_grid* T, const LaunchCtx& c, int nv, int v0, int nfields, const double* in, long long fs, long long vs);

// Chebyshev analysis of a variable chunk: tensor-core (DMMA) kernel when the level count allows it
static void grid_fwd_z(sb_grid* G, int nv, const double* in, double* mir, double* out, long long out_vs) {
  const DevGrid& d = G->dg;
  static const bool generic = std::getenv("SB_FWDZ_GENERIC") != nullptr;   // A/B switch
  if (!generic && fwd_z_mma_ok(d) && (uintptr_t)in % 16 == 0)
    launch_fwd_z_mma(G->ctx(), d, G->d_ztiles, G->nztiles, nv, in, d.N, mir, d.N, out, out_vs, G->d_fwdB);
  else
    launch_fwd_z(G->ctx(), d, G->d_ztiles, G->nztiles, nv, in, d.N, mir, d.N, out, out_vs, G->d_fwdT);
}

// Physical slots (bit d of the mask) -> what each K3 stage has to produce.  Slot order: R {f, r, rr}; RL {f, r, rr, l, ll};
// RZ {f, r, rr, z, zz}; RLZ {f, r, rr, l, ll, z, zz}.
static K3Need k3_need_from_slots(const DevGrid& t, unsigned slots) {
  K3Need n;
  const unsigned all = (1u << t.D) - 1u;
  slots &= all;
  if (slots == all) return n;
  if (t.has_l && t.has_z) {
    const unsigned zz = (slots >> 5) & 3u;
    n.lmask = (slots & 31u) | (zz ? 1u : 0u);
    n.zmask = n.lmask;
    n.zsel = (slots & 1u) | (zz << 1);
    n.smask = ((n.lmask & 25u) ? 1u : 0u) | (n.lmask & 6u);
  } else if (t.has_l) {
    n.lmask = slots & 31u;
    n.smask = ((n.lmask & 25u) ? 1u : 0u) | (n.lmask & 6u);
  } else if (t.has_z) {
    const unsigned zz = (slots >> 3) & 3u;
    n.smask = (slots & 7u) | (zz ? 1u : 0u);
    n.zmask = n.smask;
    n.zsel = (slots & 1u) | (zz << 1);
  } else {
    n.smask = slots & 7u;
  }
  return n;
}

// inverse transform: patch A -> tile physical (K3).  `need` (optional): per variable, the physical slots that
// will be read before the next K3 (tiles_physics passes what the equation-set kernel reads; the derivative slots
// of a tile are intermediates between K3 and K4 and are visible to nothing else -- in the reference they are
// overwritten by calcTendency's broadcast right after, src/semiimplicit.jl:731).  Slots outside the mask keep
// stale values.
static void grid_inverse(sb_grid* P, sb_grid* T, const unsigned* need = nullptr, bool poison = false) {
  DevGrid& t = T->dg;
  DevGrid& p = P->dg;
  if (t.has_l != p.has_l || t.has_z != p.has_z || t.V != p.V || t.bz != p.bz || t.zDim != p.zDim)
    throw std::invalid_argument("tile and patch are incompatible");
  if (t.coefOffset < p.coefOffset || t.coefOffset + t.b_rDim > p.coefOffset + p.b_rDim)
    throw std::invalid_argument("tile is not inside the patch");
  T->ensure_physical();
  const unsigned all = (1u << t.D) - 1u;
  if (need && poison)   // test hook: slots outside the mask become NaN
    CU(cudaMemsetAsync(T->physical, 0xFF, (size_t)t.N * t.V * t.D * sizeof(double), T->stream));
  // a variable with an empty mask is never read (diagnostic output of the equation set); any other variable's
  // slot 0 is always produced
  auto chunk_slots = [&](int v0, int nv) {
    unsigned m = need ? 0u : all;
    for (int v = v0; need && v < v0 + nv; ++v) m |= (need[v] & all) ? ((need[v] & all) | 1u) : 0u;
    return m;
  };
  T->slot0_src = nullptr;                 // slot 0 of every variable that is read is rewritten below
  LaunchCtx c = T->ctx();
  if (!t.has_l && !t.has_z) {
    c.need = k3_need_from_slots(t, chunk_slots(0, t.V));
    if (c.need.smask) launch_inv_r(c, t, p, t.V, P->spectralA, p.S, T->physical, 0, 0, 1, 0);
    return;
  }
  T->ensure_scratch();
  const long long slN = (long long)t.bz * t.W, szN = (long long)t.bz * t.hpoints;
  // a pass takes consecutive variables that need the SAME slots (up to vchunk of them): the stage masks are per launch, and
  // a union over unlike variables would transform rows nobody reads (the boundary-layer set: ub, vb need five ring rows,
  // the diagnostic wb none)
  // (launch-bound grids -- C2: 181,800 points, every kernel ~15 us -- take all variables in one pass with the union of their
  //  masks instead: fewer launches beat fewer rows there)
  const char* bm = std::getenv("SB_K3_BY_MASK");        // tests: 1 forces the per-mask passes on small grids, 0 the union
  const bool by_mask = bm ? std::atoi(bm) != 0 : t.N * t.V >= (1LL << 25);
  for (int v0 = 0, nv = 1; v0 < t.V; v0 += nv) {
    unsigned slots = chunk_slots(v0, 1);
    nv = 1;
    if (by_mask) {
      while (nv < T->vchunk && v0 + nv < t.V && chunk_slots(v0 + nv, 1) == slots) ++nv;
    } else {
      nv = std::min(T->vchunk, t.V - v0);
      slots = chunk_slots(v0, nv);
    }
    if (!slots) continue;                             // nothing of these variables is read (diagnostic outputs)
    c.need = k3_need_from_slots(t, slots);
    double* SL = T->scratch;                          // [3][vchunk][slN]
    double* SZ = T->scratch + ((3 * slN * T->vchunk + 15) & ~15LL);    // [5][vchunk][szN], 128-byte aligned
    const long long sl_fs = slN * T->vchunk, sz_fs = szN * T->vchunk;
    launch_inv_r(c, t, p, nv, P->spectralA + (long long)v0 * p.S, p.S, SL, sl_fs, slN, 0, v0);
    if (t.has_l && t.has_z) {
      launch_inv_l(c, t, T->iwork, T->d_iwork.data(), T->classes, T->d_tw.data(), T->d_twp.data(), T->d_plans, T->d_blob, nv, SL, sl_fs,
                   slN, SZ, sz_fs, szN, 0, v0, &T->iwork2, T->d_iwork2.data());
      grid_inv_z(T, c, nv, v0, 5, SZ, sz_fs, szN);
    } else if (t.has_l) {
      launch_inv_l(c, t, T->iwork, T->d_iwork.data(), T->classes, T->d_tw.data(), T->d_twp.data(), T->d_plans, T->d_blob, nv, SL, sl_fs,
                   slN, T->physical, 0, 0, 1, v0, &T->iwork2, T->d_iwork2.data());
    } else {
      grid_inv_z(T, c, nv, v0, 3, SL, sl_fs, slN);
    }
  }
}

// Chebyshev inverse of a variable chunk: parity fast path when no variable of the chunk has vertical BCs
static void grid_inv_z(sb_grid* T, const LaunchCtx& c, int nv, int v0, int nfields, const double* in, long long fs, long long vs) {
  const char* mode = std::getenv("SB_INVZ");   // A/B switch: "generic" | "fma" | (default) "mma"
  const std::string md = mode ? mode : "mma";
  bool bcfree = md != "generic";
  for (int v = v0; v < v0 + nv && bcfree; ++v) bcfree = T->z_bcfree[v] != 0;
  if (bcfree && md == "mma" && inv_z_mma_ok(T->dg, nfields))
    launch_inv_z_mma(c, T->dg, T->d_ztiles, T->nztiles, nv, v0, nfields, in, fs, vs, T->physical, T->d_parB);
  else if (bcfree && inv_z_par_ok(T->dg, nfields))
    launch_inv_z_par(c, T->dg, T->d_ztiles, T->nztiles, nv, v0, nfields, in, fs, vs, T->physical, T->d_parM);
  else
    launch_inv_z(c, T->dg, T->d_ztiles, T->nztiles, nv, v0, nfields, in, fs, vs, T->physical, T->d_invM);
}

static void grid_spline(sb_grid* P, const double* B) {
  launch_spline_solve(P->ctx(), P->dg, P->d_factors, P->hfactors, B, P->spectralA);
}

static void fill_gridpoints(sb_grid* G, double* out) {
  const DevGrid& d = G->dg;
  const long long N = d.N;
  const double twopi = 6.283185307179586476925286766559;
  for (int r = 0; r < d.rDim; ++r) {
    const int n = G->ring_n[r], ri = G->ring_ri[r];
    const double dl = twopi / n, ymin = 0.5 * dl * (ri - 1);
    for (int j = 0; j < n; ++j) {
      const long long h = G->hoff[r] + j;
      for (int z = 0; z < d.zDim; ++z) {
        const long long i = h * d.zDim + z;
        int col = 0;
        out[i + N * col++] = G->rad[r];
        if (d.has_l) out[i + N * col++] = ymin + dl * j;
        if (d.has_z) out[i + N * col++] = G->cheb.z[z];
      }
    }
  }
}

static void copy_params(sb_grid* G, const sb_grid_params* gp) {
  G->gp = *gp;
  auto cp = [&](const int32_t* src, std::vector<int32_t>& dst) {
    dst.assign(gp->nvars, 0);
    if (src) std::copy(src, src + gp->nvars, dst.begin());
  };
  cp(gp->BCL, G->bcl); cp(gp->BCR, G->bcr); cp(gp->BCB, G->bcb); cp(gp->BCT, G->bct);
  G->gp.BCL = G->bcl.data(); G->gp.BCR = G->bcr.data(); G->gp.BCB = G->bcb.data(); G->gp.BCT = G->bct.data();
}

static void require_device() {
#ifndef SB_EMU
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw CudaError("no CUDA device: libscythe_b200 has no CPU fallback (" + std::string(cudaGetErrorString(e)) + ")");
#endif
}

static void grid_free(sb_grid* G);
static sb_grid* grid_new(const sb_grid_params* gp, int device, void* stream) {
  if (!gp) throw std::invalid_argument("grid params is NULL");
  if (gp->nvars < 1) throw std::invalid_argument("nvars >= 1 required");
  require_device();
  sb_grid* G = new sb_grid();
  try {
    copy_params(G, gp);
    G->device = device;
    G->stream = (cudaStream_t)stream;
    build_grid(G);
  } catch (...) {          // e.g. out of memory half-way: release every device allocation made so far
    grid_free(G);
    throw;
  }
  return G;
}

static void grid_free(sb_grid* G) {
  if (!G) return;
  cudaSetDevice(G->device);
  cudaStreamSynchronize(G->stream);
  for (void* p : G->owned) cudaFree(p);
  if (G->ov.s_fft) { cudaStreamSynchronize(G->ov.s_fft); cudaStreamDestroy(G->ov.s_fft); }
  if (G->ov.s_mem) { cudaStreamSynchronize(G->ov.s_mem); cudaStreamDestroy(G->ov.s_mem); }
  for (cudaEvent_t e : {G->ov.ev_fork, G->ov.ev_invr, G->ov.ev_fl, G->ov.ev_join})
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : G->ov.ev_il) cudaEventDestroy(e);
  for (cudaEvent_t e : G->ov.ev_z) cudaEventDestroy(e);
  cudaFree(G->ov.d_counters);
  cudaFree(G->ov.buf);
  cudaFree(G->physical); if (G->owns_B) cudaFree(G->spectralB); cudaFree(G->spectralA); cudaFree(G->scratch); cudaFree(G->d_nan);
  if (G->ev0) cudaEventDestroy(G->ev0);
  if (G->ev1) cudaEventDestroy(G->ev1);
  delete G;
}

// calcTileSizes: equal-gridpoint cuts on cell boundaries, >= 3 cells per tile
static void calc_tile_sizes(const sb_grid_params* gp, int ntiles, double* out) {
  if (!gp || !out) throw std::invalid_argument("NULL argument");
  const long long nc = gp->num_cells;
  if (ntiles < 1 || nc < 3LL * ntiles) throw DomainErr("Too many tiles for this grid (need at least 3 cells per tile)");
  const bool has_l = (gp->geometry == SB_GEOM_RL || gp->geometry == SB_GEOM_RLZ);
  const bool has_z = (gp->geometry == SB_GEOM_RZ || gp->geometry == SB_GEOM_RLZ);
  const long long zDim = has_z ? gp->zDim : 1;
  const long long sIL = gp->spectralIndexL > 0 ? gp->spectralIndexL : 1;
  const long long off = (sIL - 1) * 3;
  std::vector<long long> cum(nc + 1, 0);
  for (long long c = 0; c < nc; ++c) {
    long long pts = 0;
    for (int mu = 0; mu < 3; ++mu) {
      long long ri = 3 * c + mu + 1 + off;
      pts += has_l ? 4 + 4 * ri : 1;
    }
    cum[c + 1] = cum[c] + pts * zDim;
  }
  const double total = (double)cum[nc];
  const double DX = (gp->xmax - gp->xmin) / nc;
  long long start = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int remaining = ntiles - t - 1;
    long long end;
    if (remaining == 0) {
      end = nc;
    } else {
      const double target = total * (t + 1) / ntiles;
      end = std::lower_bound(cum.begin(), cum.end(), target, [](long long a, double b) { return (double)a < b; }) - cum.begin();
      if (end > 0 && std::fabs((double)cum[end - 1] - target) <= std::fabs((double)cum[std::min(end, nc)] - target)) end -= 1;
      end = std::max(end, start + 3);
      end = std::min(end, nc - 3LL * remaining);
    }
    out[5 * t + 0] = gp->xmin + start * DX;
    out[5 * t + 1] = gp->xmin + end * DX;
    out[5 * t + 2] = (double)(end - start);
    out[5 * t + 3] = (double)(sIL + start);
    out[5 * t + 4] = (double)(cum[end] - cum[start]);
    start = end;
  }
}

static void check_cfl(sb_grid* G, int32_t* var, int64_t* index) {
  G->ensure_physical();
  G->materialize_slot0();
  long long init = 0x7fffffffffffffffLL;
  CU(cudaMemcpyAsync(G->d_nan, &init, sizeof(init), cudaMemcpyHostToDevice, G->stream));
  launch_nan_scan(G->ctx(), G->physical, G->dg.N, G->dg.V, G->d_nan);
  long long res = 0;
  CU(cudaMemcpyAsync(&res, G->d_nan, sizeof(res), cudaMemcpyDeviceToHost, G->stream));
  CU(cudaStreamSynchronize(G->stream));
  if (res != init) {
    int v = (int)(res / G->dg.N);
    long long i = res - (long long)v * G->dg.N;
    if (var) *var = v;
    if (index) *index = i;
    throw NanFound("NaN found in variable " + std::to_string(v) + " at index" + std::to_string(i + 1) +
                   " ! CFL condition likely violated");
  }
}

// ====================================================================================== NCCL (dlopen)
struct Uid { char internal[128]; };  // ncclUniqueId (passed by value to ncclCommInitRank)
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, void*) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static void nccl_load() {
  if (g_nccl.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) throw std::runtime_error(std::string("cannot load NCCL: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Uid, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, void*))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, void*))dlsym(g_nccl.lib, "ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, void*))dlsym(g_nccl.lib, "ncclRecv");
  g_nccl.GroupStart = (int (*)())dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) throw std::runtime_error("NCCL symbols missing");
}
struct CommError : std::runtime_error { using std::runtime_error::runtime_error; };
#define NC(call)                                                                         \
  do {                                                                                   \
    int r_ = (call);                                                                     \
    if (r_ != 0) throw CommError(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?")); \
  } while (0)

// ====================================================================================== model
// Host-driven stepping without stalls (sb_model_stage_in / sb_model_stage_out): two device staging buffers each way and
// two copy streams, so that the H2D copy of step i+1's state, the kernels of step i and the D2H copy of step i-1's
// result overlap (PCIe is full duplex).  All ordering is by events; the host never blocks.
struct HostPipe {
  cudaStream_t s_in = nullptr, s_out = nullptr;
  double* in[2] = {nullptr, nullptr};
  double* out[2] = {nullptr, nullptr};
  cudaEvent_t in_ready[2] = {nullptr, nullptr}, in_free[2] = {nullptr, nullptr};
  cudaEvent_t out_ready[2] = {nullptr, nullptr}, out_free[2] = {nullptr, nullptr};
  bool in_used[2] = {false, false}, out_used[2] = {false, false};
  int ki = 0, ko = 0;
  bool on = false;
};

struct TileState {
  sb_grid* grid = nullptr;
  double* var_np1 = nullptr;
  double* expd[3] = {nullptr, nullptr, nullptr};  // n, nm1, nm2 (rotating)
  double* impd[3] = {nullptr, nullptr, nullptr};
  bool hist_untouched = true;   // nothing but the equation-set kernels has written expd[] since it was allocated (zeros)
  HostPipe pipe;
};

struct sb_model {
  sb_grid_params gp{};
  std::vector<int32_t> bcl, bcr, bcb, bct;
  std::vector<std::string> var_names;
  std::map<std::string, double> params;
  double ts = 0, integration_time = 0, output_interval = 0;
  int eq = -1;
  int semiimplicit = 0;
  int k3_slots = 0;   // sb_model_set_k3_slots: 0 = what the equation set reads (+ K4 fused where built), 1 = all D slots,
                      // 2 = needed slots, the others poisoned, 3 = needed slots (no fusion)
  EqParams ep{};
  int ntiles = 1, tile_first = 0, tile_count = 1;
  std::vector<double> tile_params;
  int device = 0;
  cudaStream_t stream = nullptr;
  sb_grid* patch = nullptr;
  std::vector<TileState> tiles;
  std::vector<void*> owned;
  double* d_colops = nullptr;
  double* d_colfrag = nullptr;
  double* d_refstate = nullptr;
  double* d_sicols = nullptr;
  std::vector<double> ref_host;  // [4][3][zDim]: sbar, xibar, mubar, mu_lbar
  void* comm = nullptr;
  double* d_barrier = nullptr;
  int rank = 0, nranks = 1;
  long long extra_launches = 0;
  // CUDA graphs of three consecutive AB3 steps (sb_model_run on launch-bound grids): after three steps the history
  // pointer rotation is back where it started, so one graph per rotation phase replays for the rest of the run
  struct StepGraph {
#ifndef SB_EMU
    cudaGraphExec_t exec[3] = {nullptr, nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
#endif
    long long launches[3] = {0, 0, 0};
    int k3_slots[3] = {-1, -1, -1};
    bool failed = false;
    long long replays = 0;
  } graph;
  int rot_phase = 0;                        // number of history rotations so far, mod 3
  // ---- distributed spline solve by z-mode planes (sb_model_colsolve_*)
  struct ColSolve {
    bool on = false;
    int rank = 0, nranks = 1;
    std::vector<int> z0;                    // [nranks+1] plane ranges
    std::vector<DevGrid> tdg;               // [ntiles] tile descriptors (host copies, device tables of the patch not needed)
    std::vector<int> tile_owner;            // [ntiles] rank that owns tile t
    std::vector<long long> recv_off;        // [ntiles] offset of tile t's chunk [V][nz][ncolp_t][M_t] in recvB / sendA
    double* recvB = nullptr;                // chunks of every tile for my planes
    double* sendA = nullptr;                // solved coefficients, packed per tile in the same shape
    double* slabB = nullptr;                // [V][nz][ncolp_P][M_P]
    double* slabA = nullptr;
    DevGrid slab{};                         // patch descriptor restricted to my planes
    // peer-memory mode (sb_model_p2p_enable): no messages -- fwd_r / extract store into the other GPUs directly
    bool p2p = false;
    std::vector<double*> peer_recvB;        // [nranks] base of rank k's recvB (CUDA IPC mapping; own pointer for k == rank)
    std::vector<double*> peer_tileA;        // [ntiles] spectralA of tile t on the rank that owns it
    std::vector<PeerScatter> scat;          // [local tiles]
    std::vector<void*> ipc_opened;
    int nz() const { return z0[rank + 1] - z0[rank]; }
  } cs;
};

static int var_index(const sb_model* M, const char* name) {
  for (size_t i = 0; i < M->var_names.size(); ++i)
    if (M->var_names[i] == name) return (int)i;
  return -1;
}
static double param_or(const sb_model* M, const char* k, double dflt) {
  auto it = M->params.find(k);
  return it == M->params.end() ? dflt : it->second;
}
static double param_req(const sb_model* M, const char* k) {
  auto it = M->params.find(k);
  if (it == M->params.end()) throw std::invalid_argument(std::string("physical_params is missing :") + k);
  return it->second;
}

// variables whose expdot history the fused kernels may leave alone (ModelArrays::passive); SB_PASSIVE=0: A/B switch
static unsigned passive_mask(const sb_model* M) {
  static const bool off = std::getenv("SB_PASSIVE") != nullptr && std::atoi(std::getenv("SB_PASSIVE")) == 0;
  return off ? 0u : equation_set_passive(M->eq, M->gp.nvars);
}

static void model_check_equation_set(sb_model* M) {
  const int geom = M->gp.geometry, V = M->gp.nvars;
  auto need = [&](int g, int nv, std::initializer_list<const char*> ps) {
    if (geom != g) throw std::invalid_argument("equation set does not match the grid geometry");
    if (V < nv) throw std::invalid_argument("equation set needs more variables than the grid has");
    for (const char* p : ps) param_req(M, p);
  };
  switch (M->eq) {
    case EQ_LinearAdvection1D: need(SB_GEOM_R, 1, {"c_0", "K"}); break;
    case EQ_LinearAdvectionRZ: need(SB_GEOM_RZ, 4, {"K"}); break;
    case EQ_LinearAdvectionRL: need(SB_GEOM_RL, 3, {"K"}); break;
    case EQ_LinearAdvectionRLZ: need(SB_GEOM_RLZ, 3, {"K"}); break;
    case EQ_LinearShallowWater1D: need(SB_GEOM_R, 2, {"g", "K", "H"}); break;
    case EQ_LinearShallowWaterRL: need(SB_GEOM_RL, 3, {"g", "K", "H"}); break;
    case EQ_Oneway_ShallowWater_Slab: need(SB_GEOM_RL, 6, {"g", "K", "Cd", "Hfree", "Hb", "f"}); break;
    case EQ_Twoway_ShallowWater_Slab: need(SB_GEOM_RL, 6, {"g", "K", "Cd", "Hfree", "Hb", "f", "S1"}); break;
    case EQ_Oneway_ShallowWater_HeightResolvedBL: need(SB_GEOM_RLZ, 6, {"g", "Kh", "Cd", "Hfree", "f", "Um", "Vm"}); break;
    case EQ_Euler_test: need(SB_GEOM_RZ, 5, {"K"}); break;
    case EQ_BF02_test: need(SB_GEOM_RZ, 7, {"K"}); break;
    case EQ_rainfall_test: need(SB_GEOM_RZ, 8, {"K"}); break;
    default: throw Unsupported("equation set is not built as a CUDA kernel (no CPU fallback)");
  }
  EqParams& e = M->ep;
  e.ts = M->ts;
  e.c_0 = param_or(M, "c_0", 0); e.K = param_or(M, "K", 0); e.g = param_or(M, "g", 0); e.Cd = param_or(M, "Cd", 0);
  e.Hfree = param_or(M, "Hfree", 0); e.Hb = param_or(M, "Hb", 1); e.f = param_or(M, "f", 0); e.S1 = param_or(M, "S1", 0);
  e.H = param_or(M, "H", 0); e.Kh = param_or(M, "Kh", 0); e.Um = param_or(M, "Um", 0); e.Vm = param_or(M, "Vm", 0);
  e.iw = var_index(M, "w"); e.ixi = var_index(M, "xi"); e.ih = var_index(M, "h");
  if (M->eq == EQ_Oneway_ShallowWater_HeightResolvedBL && e.ih < 0) throw std::invalid_argument("variable \"h\" not found");
  if (M->eq == EQ_Euler_test) {
    if (M->ref_host.empty()) throw std::invalid_argument("Euler_test needs a reference state");
    if (M->semiimplicit && (e.iw != 4 || e.ixi != 1)) throw std::invalid_argument("Euler_test expects vars s,xi,mu,u,w = 1..5");
  }
  if (M->eq == EQ_BF02_test || M->eq == EQ_rainfall_test) {
    const char* name = M->eq == EQ_BF02_test ? "BF02_test" : "rainfall_test";
    if (M->ref_host.empty()) throw std::invalid_argument(std::string(name) + " needs a reference state");
    // condensation_adjustment (src/microphysics.jl:141-165) looks these up by NAME in grid_params.vars and throws a
    // KeyError when one is missing -- which is what the reference's BF02_test does on its first step with a variable list
    // that calls column 6 "mu_l".  Same error, raised when the model is created.
    for (const char* k : {"s", "xi", "mu", "mu_c", "mu_r", "qss"})
      if (var_index(M, k) < 0) throw std::invalid_argument(std::string("KeyError: key \"") + k + "\" not found (condensation_adjustment, src/microphysics.jl:141-165)");
    const bool rain = M->eq == EQ_rainfall_test;
    const char* order[2][8] = {{"s", "xi", "mu", "u", "w", "mu_c", "qss", "mu_r"}, {"s", "xi", "mu", "u", "w", "mu_c", "mu_r", "qss"}};
    for (int v = 0; v < 8; ++v)
      if (var_index(M, order[rain][v]) != v)
        throw std::invalid_argument(std::string(name) + (rain ? " expects vars s,xi,mu,u,w,mu_c,mu_r,qss = 1..8"
                                                              : " expects vars s,xi,mu,u,w,mu_c,qss,mu_r = 1..8 (column 6 is the equation set's mu_l)"));
  }
}

// composite column operators (all stored transposed: Mt[k][z'][z] = M[z][z'])
static void build_column_ops(sb_model* M) {
  sb_grid* P = M->patch;
  const DevGrid& d = P->dg;
  if (!d.has_z) return;
  const int nz = d.zDim, bz = d.bz;
  const ChebTables& ct = P->cheb;
  auto composite = [&](const std::vector<double>& T, const std::vector<double>& IG, std::vector<double>& out) {
    // out[z][z'] = sum_{q,zb} T[z][q] IG[q][zb] fwd[zb][z']
    std::vector<double> TI((size_t)nz * bz);
    for (int z = 0; z < nz; ++z)
      for (int zb = 0; zb < bz; ++zb) {
        long double s = 0;
        for (int q = 0; q < bz; ++q) s += (long double)T[(size_t)z * nz + q] * IG[(size_t)q * bz + zb];
        TI[(size_t)z * bz + zb] = (double)s;
      }
    out.assign((size_t)nz * nz, 0.0);
    matmul(TI.data(), ct.fwd.data(), out.data(), nz, bz, nz);
  };
  auto transpose_into = [&](const std::vector<double>& A, double* dst) {
    for (int z = 0; z < nz; ++z)
      for (int k = 0; k < nz; ++k) dst[(size_t)k * nz + z] = A[(size_t)z * nz + k];
  };
  const size_t nn = (size_t)nz * nz;
  if (M->eq == EQ_Oneway_ShallowWater_HeightResolvedBL) {
    const int ih = M->ep.ih;
    std::vector<double> IG = cheb_bc_matrix(ct, P->bcb[ih], P->bct[ih]);
    std::vector<double> ops(3 * nn), tmp;
    composite(ct.T0, IG, tmp); transpose_into(tmp, ops.data());
    composite(ct.T1, IG, tmp); transpose_into(tmp, ops.data() + nn);
    composite(ct.Tint, IG, tmp); transpose_into(tmp, ops.data() + 2 * nn);
    M->d_colops = dev_upload(ops); M->owned.push_back(M->d_colops);
    if (nz % 8 == 0 && nz <= 64) {   // tensor-core column operators (k_heightresolved_bl2)
      std::vector<double> frag, f2;
      build_colop_fragments(nz, ops.data() + 2 * nn, frag);
      build_colop_fragments(nz, ops.data() + nn, f2);
      frag.insert(frag.end(), f2.begin(), f2.end());
      M->d_colfrag = dev_upload(frag); M->owned.push_back(M->d_colfrag);
    }
  }
  if (M->eq == EQ_Euler_test || M->eq == EQ_BF02_test || M->eq == EQ_rainfall_test) {
    M->d_refstate = dev_upload(M->ref_host); M->owned.push_back(M->d_refstate);
    std::vector<double> ops(7 * nn, 0.0);
    if (M->eq == EQ_rainfall_test) {   // CB -> CA -> CIx of the "mu_r" column (src/testModels.jl:525-529)
      const int imr = var_index(M, "mu_r");
      std::vector<double> IGr = cheb_bc_matrix(ct, P->bcb[imr], P->bct[imr]), Dr;
      composite(ct.T1, IGr, Dr); transpose_into(Dr, ops.data() + 6 * nn);
    }
    if (M->semiimplicit) {
      const int ixi = M->ep.ixi;
      std::vector<double> IG = cheb_bc_matrix(ct, P->bcb[ixi], P->bct[ixi]);
      std::vector<double> F, Dz;
      composite(ct.T0, IG, F); transpose_into(F, ops.data());
      composite(ct.T1, IG, Dz); transpose_into(Dz, ops.data() + nn);
      // Helmholtz (src/semiimplicit.jl:768-781): rows 0,1 = tau^2 Pxi dct[0,:], dct[nz-1,:]; rows 2.. = (tau^2 Pxi dct2 - dct)[1..nz-2]
      for (int k = 0; k < 2; ++k) {
        const double tau = (k == 0 ? 0.5 : 1.25) * M->ts;
        const double c = tau * tau * M->ep.Pxi_bar;
        std::vector<double> H(nn);
        for (int j = 0; j < nz; ++j) {
          H[j] = c * ct.T0[j];
          H[(size_t)nz + j] = c * ct.T0[(size_t)(nz - 1) * nz + j];
        }
        for (int i = 2; i < nz; ++i)
          for (int j = 0; j < nz; ++j) H[(size_t)i * nz + j] = c * ct.T2[(size_t)(i - 1) * nz + j] - ct.T0[(size_t)(i - 1) * nz + j];
        if (!invert(H, nz)) throw std::runtime_error("Helmholtz matrix is singular");
        // HS = H^-1 Shift : column j of HS = column j+1 of H^-1 for j = 1..nz-2 (g_new[i] = g_old[i-1], i>=2), else 0
        std::vector<double> HS(nn, 0.0);
        for (int i = 0; i < nz; ++i)
          for (int j = 1; j <= nz - 2; ++j) HS[(size_t)i * nz + j] = H[(size_t)i * nz + j + 1];
        std::vector<double> W(nn), X(nn);
        matmul(ct.T0.data(), HS.data(), W.data(), nz, nz, nz);
        matmul(ct.T1.data(), HS.data(), X.data(), nz, nz, nz);
        transpose_into(W, ops.data() + (2 + 2 * k) * nn);
        transpose_into(X, ops.data() + (3 + 2 * k) * nn);
      }
    }
    if (M->semiimplicit || M->eq == EQ_rainfall_test) { M->d_sicols = dev_upload(ops); M->owned.push_back(M->d_sicols); }
  }
}

static void model_free(sb_model* M);
static sb_model* model_new(const sb_model_params* mp, int ntiles, int tile_first, int tile_count, int device, void* stream) {
  if (!mp || !mp->grid || !mp->equation_set) throw std::invalid_argument("NULL model parameter");
  if (!(mp->ts > 0)) throw std::invalid_argument("ts must be positive");
  if (ntiles < 1 || tile_first < 0 || tile_count < 1 || tile_first + tile_count > ntiles)
    throw std::invalid_argument("bad tile range");
  require_device();
  struct Guard {          // a throw below (bad parameter, out of memory) frees every device allocation made so far
    sb_model* m;
    ~Guard() { if (m) model_free(m); }
    sb_model* get() const { return m; }
    sb_model* operator->() const { return m; }
    sb_model* release() { sb_model* r = m; m = nullptr; return r; }
  } M{new sb_model()};
  M->gp = *mp->grid;
  const int V = mp->grid->nvars;
  auto cp = [&](const int32_t* src, std::vector<int32_t>& dst) { dst.assign(V, 0); if (src) std::copy(src, src + V, dst.begin()); };
  cp(mp->grid->BCL, M->bcl); cp(mp->grid->BCR, M->bcr); cp(mp->grid->BCB, M->bcb); cp(mp->grid->BCT, M->bct);
  M->gp.BCL = M->bcl.data(); M->gp.BCR = M->bcr.data(); M->gp.BCB = M->bcb.data(); M->gp.BCT = M->bct.data();
  for (int v = 0; v < V; ++v) M->var_names.push_back(mp->var_names && mp->var_names[v] ? mp->var_names[v] : ("v" + std::to_string(v)));
  for (int i = 0; i < mp->n_physical_params; ++i) M->params[mp->param_names[i]] = mp->param_values[i];
  M->ts = mp->ts; M->integration_time = mp->integration_time; M->output_interval = mp->output_interval;
  M->semiimplicit = mp->semiimplicit;
  M->eq = equation_set_from_name(mp->equation_set);
  if (const char* e = std::getenv("SB_K3_FULL")) M->k3_slots = std::atoi(e) != 0 ? 1 : 0;
  if (M->eq < 0) throw Unsupported(std::string("equation set \"") + mp->equation_set + "\" is not built as a CUDA kernel (no CPU fallback)");
  const bool has_z = (M->gp.geometry == SB_GEOM_RZ || M->gp.geometry == SB_GEOM_RLZ);
  if (mp->ref_sbar && mp->ref_xibar && mp->ref_mubar && has_z) {
    const size_t n3 = (size_t)3 * M->gp.zDim;
    M->ref_host.assign(4 * n3, 0.0);
    std::copy(mp->ref_sbar, mp->ref_sbar + n3, M->ref_host.begin());
    std::copy(mp->ref_xibar, mp->ref_xibar + n3, M->ref_host.begin() + n3);
    std::copy(mp->ref_mubar, mp->ref_mubar + n3, M->ref_host.begin() + 2 * n3);
    if (mp->ref_mu_lbar) std::copy(mp->ref_mu_lbar, mp->ref_mu_lbar + n3, M->ref_host.begin() + 3 * n3);
  }
  M->ep.Pxi_bar = mp->Pxi_bar;
  model_check_equation_set(M.get());
  M->ntiles = ntiles; M->tile_first = tile_first; M->tile_count = tile_count;
  M->device = device; M->stream = (cudaStream_t)stream;
  M->tile_params.resize((size_t)5 * ntiles);
  calc_tile_sizes(&M->gp, ntiles, M->tile_params.data());
  M->patch = grid_new(&M->gp, device, stream);
  std::vector<int32_t> r0(V, SB_BC_R0);
  for (int t = tile_first; t < tile_first + tile_count; ++t) {
    sb_grid_params tp = M->gp;
    tp.xmin = M->tile_params[5 * t + 0];
    tp.xmax = M->tile_params[5 * t + 1];
    tp.num_cells = (int64_t)M->tile_params[5 * t + 2];
    tp.spectralIndexL = (int64_t)M->tile_params[5 * t + 3];
    tp.tile_num = t + 2;
    tp.BCL = r0.data();  // tiles carry no radial BCs (src/semiimplicit.jl:163-164)
    tp.BCR = r0.data();
    M->tiles.emplace_back();
    TileState& ts = M->tiles.back();     // registered before anything is allocated: model_free sees partial tiles
    ts.grid = grid_new(&tp, device, stream);
    if (ntiles == 1) {   // one tile == the patch: its B *is* the shared B (no clear / assemble pass per step)
      CU(cudaFree(ts.grid->spectralB));
      ts.grid->spectralB = M->patch->spectralB;
      ts.grid->owns_B = false;
    }
    // the tile's physical [N,V,D] array is allocated on first use (a non-fused K3, an API read): the default fused
    // step never touches it (7.2 GB at C4)
    const long long n = ts.grid->dg.N * V;
    ts.var_np1 = dev_zeros(n, M->stream);
    for (int k = 0; k < 3; ++k) ts.expd[k] = dev_zeros(n, M->stream);
    if (M->semiimplicit)
      for (int k = 0; k < 3; ++k) ts.impd[k] = dev_zeros(n, M->stream);
  }
  build_column_ops(M.get());
  return M.release();
}

#ifndef SB_EMU
static void step_graph_free(sb_model* M);
#endif
static void model_free(sb_model* M) {
  if (!M) return;
#ifndef SB_EMU
  step_graph_free(M);
#endif
  if (M->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(M->comm);
#ifndef SB_EMU
  for (void* p : M->cs.ipc_opened) cudaIpcCloseMemHandle(p);
#endif
  for (auto& t : M->tiles) {
    if (t.pipe.on) {
      cudaStreamSynchronize(t.pipe.s_in); cudaStreamSynchronize(t.pipe.s_out);
      for (int k = 0; k < 2; ++k) {
        cudaFree(t.pipe.in[k]); cudaFree(t.pipe.out[k]);
        cudaEventDestroy(t.pipe.in_ready[k]); cudaEventDestroy(t.pipe.in_free[k]);
        cudaEventDestroy(t.pipe.out_ready[k]); cudaEventDestroy(t.pipe.out_free[k]);
      }
      cudaStreamDestroy(t.pipe.s_in); cudaStreamDestroy(t.pipe.s_out);
    }
    grid_free(t.grid);
    cudaFree(t.var_np1);
    for (int k = 0; k < 3; ++k) { cudaFree(t.expd[k]); cudaFree(t.impd[k]); }
  }
  grid_free(M->patch);
  for (void* p : M->owned) cudaFree(p);
  delete M;
}

static void model_initialize(sb_model* M, const double* ic_host) {
  sb_grid* P = M->patch;
  P->ensure_physical();
  if (ic_host) CU(cudaMemcpyAsync(P->physical, ic_host, (size_t)P->dg.N * P->dg.V * sizeof(double), cudaMemcpyHostToDevice, P->stream));
  grid_forward(P, P->physical, nullptr);      // spectralTransform!(patch)   :135
  grid_spline(P, P->spectralB);               // gridTransform!(patch)       :136  (A-solve ...
  grid_inverse(P, P);                         //                                   ... + evaluate)
  if (M->cs.on)                               // plane-distributed solve: every local tile keeps its own slice of A
    for (auto& T : M->tiles) launch_extract(P->ctx(), P->dg, T.grid->dg, P->spectralA, T.grid->spectralA);
  CU(cudaStreamSynchronize(P->stream));
  if ((size_t)P->dg.N * P->dg.V * P->dg.D * sizeof(double) > ((size_t)4 << 30)) P->release_physical();
}

// LinearAdvectionRLZ: inv_r + inv_l of the seven rows the equation reads (h: value, r, rr, l, ll; u, v: value), then the
// Chebyshev synthesis with the tendency and the AB3 step in its epilogue (k_inv_z_advection).  `physical` is not touched.
static bool fused_advection_ok(const sb_model* M, const sb_grid* G) {
  static const bool off = std::getenv("SB_FUSE_K4") != nullptr && std::atoi(std::getenv("SB_FUSE_K4")) == 0;   // A/B switch
  if (off || M->eq != EQ_LinearAdvectionRLZ || !inv_z_advection_ok(G->dg)) return false;
  for (int v = 0; v < G->dg.V; ++v)
    if (!G->z_bcfree[v]) return false;        // vertical BCs: the parity split does not hold
  return true;
}
static void grid_inverse_advection_fused(sb_grid* P, sb_grid* T, const EqParams& ep, const ModelArrays& a, int tq) {
  DevGrid& t = T->dg;
  DevGrid& p = P->dg;
  // SZ in the blocked layout (two bulk copies per field and tile in the synthesis kernel instead of one per z-mode) when
  // that kernel runs; the ring-row layout otherwise
  const bool blocked = inv_z_advection_blocked(t);
  const long long slN = (long long)t.bz * t.W, szN = (long long)t.bz * (blocked ? t.hpointsp : t.hpoints);
  const int sz_kind = blocked ? 2 : 0;
  const long long sz_off = (3 * slN + 15) & ~15LL;
  T->ensure_scratch_doubles(sz_off + 7 * szN + 16);
  double* SL = T->scratch;                  // [3][slN]   one variable at a time
  double* SZ = T->scratch + sz_off;         // [7][szN]   h: 5 rows | u | v
  T->slot0_src = nullptr;                   // the state the step starts from is synthesised in registers
  LaunchCtx c = T->ctx();
  // h: five rows (value, r, rr, l, ll) from three spectra; u and v: the value row only, both in one pass (same masks)
  c.need = k3_need_from_slots(t, 31u);
  launch_inv_r(c, t, p, 1, P->spectralA, p.S, SL, slN, slN, 0, 0);
  launch_inv_l(c, t, T->iwork, T->d_iwork.data(), T->classes, T->d_tw.data(), T->d_twp.data(), T->d_plans, T->d_blob, 1, SL, slN,
               slN, SZ, szN, szN, sz_kind, 0, &T->iwork2, T->d_iwork2.data());
  c.need = k3_need_from_slots(t, 1u);
  launch_inv_r(c, t, p, 2, P->spectralA + p.S, p.S, SL, 2 * slN, slN, 0, 1);          // SL[0] = u spectrum, SL[1] = v spectrum
  launch_inv_l(c, t, T->iwork, T->d_iwork.data(), T->classes, T->d_tw.data(), T->d_twp.data(), T->d_plans, T->d_blob, 2, SL,
               2 * slN, slN, SZ + 5 * szN, szN, szN, sz_kind, 1, &T->iwork2, T->d_iwork2.data());
  launch_inv_z_advection(c, t, T->d_ztiles, T->nztiles, SZ, szN, T->d_parB, ep, a, tq, blocked);
}

// first half of advanceTimestep: tileTransform! + equation set + explicit/semi-implicit step
static void tiles_physics(sb_model* M, int64_t t) {
  sb_grid* P = M->patch;
  // K3 produces what the equation-set kernel reads, not every slot (sb_model_set_k3_slots(m, 1): all D slots of all
  // variables, the reference's materialised dataflow -- bench.py times both)
  std::vector<unsigned> need(P->dg.V);
  equation_set_needs(M->eq, M->ep, P->dg, need.data());
  for (auto& T : M->tiles) {
    sb_grid* G = T.grid;
    ModelArrays a{};
    a.var_np1 = T.var_np1;
    a.exp_n = T.expd[0]; a.exp_nm1 = T.expd[1]; a.exp_nm2 = T.expd[2];
    a.imp_n = T.impd[0]; a.imp_nm1 = T.impd[1]; a.imp_nm2 = T.impd[2];
    a.colops = M->d_colops; a.colfrag = M->d_colfrag; a.refstate = M->d_refstate; a.sicols = M->d_sicols;
    a.passive = T.hist_untouched ? passive_mask(M) : 0u;
    const int tq = (int)std::min<int64_t>(t, 3);
    if (M->k3_slots == 0 && fused_advection_ok(M, G)) {   // K3's last stage and K4 in one kernel: no slot reaches HBM
      grid_inverse_advection_fused(M->cs.on ? G : P, G, M->ep, a, tq);
    } else {
      grid_inverse(M->cs.on ? G : P, G, M->k3_slots == 1 ? nullptr : need.data(), M->k3_slots == 2);   // tileTransform!  :305 (plane-distributed solve: A arrives tile-local)
      a.phys = G->physical;
      launch_equation_set(G->ctx(), M->eq, G->dg, M->ep, a, tq);  // :308-314
    }
    // history rotation (:685-695): nm2 <- nm1 <- n ; the old nm2 buffer becomes next step's n
    std::rotate(T.expd, T.expd + 2, T.expd + 3);
    if (M->semiimplicit) std::rotate(T.impd, T.impd + 2, T.impd + 3);
  }
  M->rot_phase = (M->rot_phase + 1) % 3;
}

// second half: calcTendency (K1) + own block / halo into the shared B buffer
static void tiles_tendency(sb_model* M) {
  sb_grid* P = M->patch;
  if (M->cs.on) {   // tile B stays tile-local; the z-mode-plane owners assemble and solve (sb_model_colsolve_solve)
    for (auto& T : M->tiles) { grid_forward(T.grid, T.var_np1, nullptr); T.grid->slot0_src = T.var_np1; }
    return;
  }
  if (M->ntiles == 1 && !M->tiles[0].grid->owns_B) {   // the tile writes the shared B directly
    TileState& T = M->tiles[0];
    grid_forward(T.grid, T.var_np1, nullptr);
    T.grid->slot0_src = T.var_np1;
    return;
  }
  CU(cudaMemsetAsync(P->spectralB, 0, (size_t)P->dg.S * P->dg.V * sizeof(double), P->stream));  // :272
  sb_grid* prev = nullptr;
  for (auto& T : M->tiles) {
    sb_grid* G = T.grid;
    grid_forward(G, T.var_np1, nullptr);                    // calcTendency    :728-735 (slot-0 copy deferred)
    G->slot0_src = T.var_np1;
    launch_assemble(G->ctx(), P->dg, G->dg, G->spectralB, prev ? &prev->dg : nullptr, prev ? prev->spectralB : nullptr, 0,
                    P->spectralB);                         // :320-329
    prev = G;
  }
}

// ====================================================================================== overlapped step
// advanceTimestep of one tile (src/semiimplicit.jl:301-332) for LinearAdvectionRLZ with the two kinds of kernels of the
// step running side by side: the ring FFTs are FP64-pipe bound and leave HBM almost idle, the Chebyshev / radial stages
// are HBM bound and leave the FP64 pipe idle.  The tile's rings are cut into batches (outermost first); the ring FFTs of
// batch b run on a high-priority stream with grids that leave `sm_reserve` SMs free, and on a second stream the Chebyshev
// synthesis + equation set + AB3 and the Chebyshev analysis of batch b-1 run on those SMs:
//   mem stream:  inv_r | wait I(b0): Z(b0) | wait I(b1): Z(b1) | ...                         | wait F(last): fwd_r
//   fft stream:        | I(b0) | I(b1) | I(b2) | ... | wait Z(b0): F(b0) | wait Z(b1): F(b1) ...
//   I = inv_l of the 7 rows, Z = inv_z + K4 then fwd_z, F = fwd_l.  Same kernels, same arithmetic, same per-point order as
//   the sequential step: the state is bit-identical (tests: test_overlapped_step_*).
static bool overlap_usable(const sb_model* M) {
  // Opt-in (SB_OVERLAP=1).  Measured on B200 at C4 (gpurun_out/r2k): 18.5 ms with no SM reserved, 27.0 ms with
  // 24, 21.2 ms with 40 -- against 18.2 ms for the one-stream step.  The HBM-bound kernels reach only ~45 GB/s per SM
  // when confined to a few SMs (a plain copy kernel: 120 GB/s per SM, profiles/microbench/overlap_probe.cu), so they need
  // more SMs than the ring FFTs can spare.  Kept, tested for bit-identity, as the frame for kernels that stream better.
  const char* e = std::getenv("SB_OVERLAP");
  if (!e || std::atoi(e) != 1) return false;
  if (M->k3_slots != 0 || M->eq != EQ_LinearAdvectionRLZ) return false;
  for (auto& T : M->tiles)
    if (!fused_advection_ok(M, T.grid) || T.grid->zt_first.empty()) return false;
  return true;
}

static void overlap_prepare(sb_grid* G) {
  auto& ov = G->ov;
  if (ov.ready) return;
  const DevGrid& t = G->dg;
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CU(cudaStreamCreateWithPriority(&ov.s_fft, cudaStreamNonBlocking, hi));
  CU(cudaStreamCreateWithPriority(&ov.s_mem, cudaStreamNonBlocking, lo));
  for (cudaEvent_t* e : {&ov.ev_fork, &ov.ev_invr, &ov.ev_fl, &ov.ev_join}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  const char* eb = std::getenv("SB_OVERLAP_BATCHES");
  int nb = eb && std::atoi(eb) > 0 ? std::atoi(eb) : 6;
  if (nb > t.rDim) nb = t.rDim;
  // equal shares of the horizontal points, outermost rings first
  int r_hi = t.rDim;
  for (int b = 0; b < nb; ++b) {
    const long long target = G->hoff[t.rDim] * (long long)(nb - 1 - b) / nb;    // points below the batch
    int r_lo = (b == nb - 1) ? 0 : r_hi - 1;
    while (r_lo > 0 && G->hoff[r_lo] > target) --r_lo;
    if (b == nb - 1) r_lo = 0;
    if (r_lo < r_hi) ov.batches.emplace_back(r_lo, r_hi);
    r_hi = r_lo;
  }
  ov.ev_il.resize(ov.batches.size());
  ov.ev_z.resize(ov.batches.size());
  for (auto& e : ov.ev_il) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : ov.ev_z) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (!ov.d_counters) CU(cudaMalloc((void**)&ov.d_counters, 512 * sizeof(int)));
  const long long slN = (long long)t.bz * t.W, szN = (long long)t.bz * t.hpoints;
  ov.buf_doubles = 8 * ((slN + 15) & ~15LL) + 10 * ((szN + 15) & ~15LL) + 64;
  CU(cudaMalloc((void**)&ov.buf, (size_t)ov.buf_doubles * sizeof(double)));
  const char* ek = std::getenv("SB_OVERLAP_K");
  ov.sm_reserve = ek ? std::atoi(ek) : 24;
  if (ov.sm_reserve < 0) ov.sm_reserve = 0;
  if (ov.sm_reserve > sb_sm_count() / 2) ov.sm_reserve = sb_sm_count() / 2;
  ov.ready = true;
}

static void tile_step_overlapped(sb_model* M, TileState& T, int tq) {
  sb_grid* G = T.grid;
  sb_grid* P = M->cs.on ? G : M->patch;
  overlap_prepare(G);
  auto& ov = G->ov;
  DevGrid& t = G->dg;
  DevGrid& p = P->dg;
  const long long slN = (long long)t.bz * t.W, szN = (long long)t.bz * t.hpoints;
  const long long slA = (slN + 15) & ~15LL, szA = (szN + 15) & ~15LL;    // 128-byte aligned regions: rows are read as double2 / bulk-copied
  double* SLi = ov.buf;                 // [5]: h value, d/dr, d2/dr2 | u | v
  double* SZi = SLi + 5 * slA;          // [7]: h value, r, rr, l, ll | u | v
  double* SLf = SZi + 7 * szA;          // [3 variables]
  double* SZf = SLf + 3 * slA;          // [3 variables]
  ModelArrays a{};
  a.var_np1 = T.var_np1;
  a.exp_n = T.expd[0]; a.exp_nm1 = T.expd[1]; a.exp_nm2 = T.expd[2];
  a.passive = T.hist_untouched ? passive_mask(M) : 0u;
  G->slot0_src = nullptr;
  LaunchCtx cm = G->ctx(), cf = G->ctx();
  cm.stream = ov.s_mem;
  cf.stream = ov.s_fft;
  cf.counters = ov.d_counters; cf.counter_next = &ov.counter_next; cf.ncounters = 512; cf.sm_reserve = ov.sm_reserve;
  CU(cudaEventRecord(ov.ev_fork, G->stream));
  CU(cudaStreamWaitEvent(ov.s_mem, ov.ev_fork, 0));
  CU(cudaStreamWaitEvent(ov.s_fft, ov.ev_fork, 0));
  // ---- radial evaluation of the three variables (whole tile, nothing to share the chip with yet)
  const unsigned slots[3] = {31u, 1u, 1u};
  for (int v = 0; v < 3; ++v) {
    cm.need = k3_need_from_slots(t, slots[v]);
    launch_inv_r(cm, t, p, 1, P->spectralA + (long long)v * p.S, p.S, SLi + (v ? (2 + v) * slA : 0), slA, slA, 0, v);
  }
  CU(cudaEventRecord(ov.ev_invr, ov.s_mem));
  CU(cudaStreamWaitEvent(ov.s_fft, ov.ev_invr, 0));
  const int nb = (int)ov.batches.size();
  for (int b = 0; b < nb; ++b) {
    const int r_lo = ov.batches[b].first, r_hi = ov.batches[b].second;
    // ---- I(b): ring synthesis of the seven rows
    cf.r_lo = r_lo; cf.r_hi = r_hi;
    for (int v = 0; v < 3; ++v) {
      cf.need = k3_need_from_slots(t, slots[v]);
      launch_inv_l(cf, t, G->iwork, G->d_iwork.data(), G->classes, G->d_tw.data(), G->d_twp.data(), G->d_plans, G->d_blob, 1,
                   SLi + (v ? (2 + v) * slA : 0), slA, slA, SZi + (v ? (4 + v) * szA : 0), szA, szA, 0, v, &G->iwork2,
                   G->d_iwork2.data());
    }
    CU(cudaEventRecord(ov.ev_il[b], ov.s_fft));
    // ---- Z(b): Chebyshev synthesis + tendency + AB3, then Chebyshev analysis of the new state, on the reserved SMs
    CU(cudaStreamWaitEvent(ov.s_mem, ov.ev_il[b], 0));
    const int z0 = G->zt_first[r_lo], z1 = G->zt_first[r_hi];
    cm.mem_grid_sms = (b + 1 < nb) ? ov.sm_reserve : 0;       // the last batch has the chip to itself until F(b0) starts
    if (cm.mem_grid_sms > 0 && cm.mem_grid_sms < 4) cm.mem_grid_sms = 4;
    launch_inv_z_advection(cm, t, G->d_ztiles + z0, z1 - z0, SZi, szA, G->d_parB, M->ep, a, tq);
    launch_fwd_z_mma(cm, t, G->d_ztiles + z0, z1 - z0, 3, T.var_np1, t.N, nullptr, t.N, SZf, szA, G->d_fwdB);
    CU(cudaEventRecord(ov.ev_z[b], ov.s_mem));
  }
  // ---- F(b): ring analysis
  for (int b = 0; b < nb; ++b) {
    cf.r_lo = ov.batches[b].first; cf.r_hi = ov.batches[b].second;
    cf.need = K3Need{};
    CU(cudaStreamWaitEvent(ov.s_fft, ov.ev_z[b], 0));
    launch_fwd_l(cf, t, G->fwork, G->d_fwork.data(), G->classes, G->d_tw.data(), G->d_twp.data(), G->d_plans, G->d_blob, 3, SZf, szA,
                 1, nullptr, 0, SLf, slA, &G->fwork2, G->d_fwork2.data(), G->d_fft3_scratch);
  }
  CU(cudaEventRecord(ov.ev_fl, ov.s_fft));
  CU(cudaStreamWaitEvent(ov.s_mem, ov.ev_fl, 0));
  cm.mem_grid_sms = 0;
  cm.need = K3Need{};
  launch_fwd_r(cm, t, 3, SLf, slA, G->spectralB, t.S, G->scatter, 0);
  CU(cudaEventRecord(ov.ev_join, ov.s_mem));
  CU(cudaStreamWaitEvent(G->stream, ov.ev_join, 0));
  G->slot0_src = T.var_np1;
}

static void tiles_step_overlapped(sb_model* M, int64_t t) {
  sb_grid* P = M->patch;
  const int tq = (int)std::min<int64_t>(t, 3);
  const bool direct = M->cs.on || (M->ntiles == 1 && !M->tiles[0].grid->owns_B);
  if (!direct) CU(cudaMemsetAsync(P->spectralB, 0, (size_t)P->dg.S * P->dg.V * sizeof(double), P->stream));  // :272
  sb_grid* prev = nullptr;
  for (auto& T : M->tiles) {
    tile_step_overlapped(M, T, tq);
    std::rotate(T.expd, T.expd + 2, T.expd + 3);
    if (!direct) {
      sb_grid* G = T.grid;
      launch_assemble(G->ctx(), P->dg, G->dg, G->spectralB, prev ? &prev->dg : nullptr, prev ? prev->spectralB : nullptr, 0,
                      P->spectralB);                         // :320-329
      prev = G;
    }
  }
  M->rot_phase = (M->rot_phase + 1) % 3;
}

static void model_advance_tiles(sb_model* M, int64_t t) {
  if (overlap_usable(M)) { tiles_step_overlapped(M, t); return; }
  tiles_physics(M, t);
  tiles_tendency(M);
}

static void model_exchange(sb_model* M);

static void model_step(sb_model* M, int64_t t) {
  model_advance_tiles(M, t);
  model_exchange(M);
  if (!M->cs.on) grid_spline(M->patch, M->patch->spectralB);  // splineTransform! :285
}


// ====================================================================================== CUDA graphs of the step
// Small grids are launch bound (C2, the reference's production configuration: 22 launches of a few microseconds per
// step): sb_model_run replays three AB3 steps as one graph launch.  Only for a single process (no communicator), with
// the profiler off, on grids below SB_GRAPH_MAX_POINTS points x variables (default 2^25; SB_GRAPH=0 switches it off).
static bool step_graph_usable(sb_model* M) {
#ifdef SB_EMU
  (void)M;
  return false;
#else
  static const bool off = std::getenv("SB_GRAPH") && std::atoi(std::getenv("SB_GRAPH")) == 0;
  static const long long maxpts = std::getenv("SB_GRAPH_MAX_POINTS") ? std::atoll(std::getenv("SB_GRAPH_MAX_POINTS")) : (1LL << 25);
  if (off || M->graph.failed || M->nranks > 1 || M->comm || (M->cs.on && M->cs.nranks > 1)) return false;
  if (M->patch->dg.N * M->patch->dg.V > maxpts) return false;
  if (M->patch->prof.on) return false;
  for (auto& T : M->tiles)
    if (T.grid->prof.on || T.pipe.on) return false;
  return true;
#endif
}

#ifndef SB_EMU
static void step_graph_free(sb_model* M) {
  auto& g = M->graph;
  for (int k = 0; k < 3; ++k)
    if (g.exec[k]) { cudaGraphExecDestroy(g.exec[k]); g.exec[k] = nullptr; }
  if (g.ev_in) cudaEventDestroy(g.ev_in);
  if (g.ev_out) cudaEventDestroy(g.ev_out);
  if (g.stream) cudaStreamDestroy(g.stream);
  g.ev_in = g.ev_out = nullptr;
  g.stream = nullptr;
}

// three steps t, t+1, t+2 (all AB3: t >= 3) through the graph of the current rotation phase; captured on first use
static void step_graph_run3(sb_model* M, int64_t t) {
  auto& g = M->graph;
  const int ph = M->rot_phase;
  if (!g.stream) {
    CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g.ev_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g.ev_out, cudaEventDisableTiming));
  }
  if (g.exec[ph] && g.k3_slots[ph] != M->k3_slots) { cudaGraphExecDestroy(g.exec[ph]); g.exec[ph] = nullptr; }
  bool first = false;
  if (!g.exec[ph]) {
    // capture: every launch of the three steps goes to the capture stream instead of the model's stream
    std::vector<sb_grid*> grids{M->patch};
    for (auto& T : M->tiles) grids.push_back(T.grid);
    std::vector<cudaStream_t> saved;
    for (sb_grid* G : grids) { saved.push_back(G->stream); G->stream = g.stream; }
    cudaStream_t saved_m = M->stream;
    M->stream = g.stream;
    std::vector<TileState> tiles_before = M->tiles;
    const int phase_before = M->rot_phase;
    long long l0 = sb_model_launch_count(M);
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    std::string why = "cudaStreamBeginCapture failed";
    if (ok) {
      try {
        for (int k = 0; k < 3; ++k) model_step(M, t + k);
      } catch (const std::exception& e) { ok = false; why = e.what(); }
      cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
      if (ce != cudaSuccess || !graph) { ok = false; if (why.empty() || why == "cudaStreamBeginCapture failed") why = cudaGetErrorString(ce); }
    }
    if (ok && cudaGraphInstantiate(&g.exec[ph], graph, 0) != cudaSuccess) { ok = false; why = "cudaGraphInstantiate failed"; g.exec[ph] = nullptr; }
    if (graph) cudaGraphDestroy(graph);
    for (size_t i = 0; i < grids.size(); ++i) grids[i]->stream = saved[i];
    M->stream = saved_m;
    if (!ok) {                       // nothing ran: put the host-side state back and step eagerly from now on
      cudaGetLastError();
      for (size_t i = 0; i < M->tiles.size(); ++i) {
        for (int k = 0; k < 3; ++k) { M->tiles[i].expd[k] = tiles_before[i].expd[k]; M->tiles[i].impd[k] = tiles_before[i].impd[k]; }
      }
      M->rot_phase = phase_before;
      g.failed = true;
      for (int k = 0; k < 3; ++k) model_step(M, t + k);
      return;
    }
    g.launches[ph] = sb_model_launch_count(M) - l0;   // counted once by the capture itself
    g.k3_slots[ph] = M->k3_slots;
    first = true;
  } else {
    // host-side state of three steps: three pointer rotations = identity; slot0_src as model_step leaves it
    for (auto& T : M->tiles) T.grid->slot0_src = T.var_np1;
    M->extra_launches += g.launches[ph];
  }
  (void)first;
  CU(cudaEventRecord(g.ev_in, M->stream));
  CU(cudaStreamWaitEvent(g.stream, g.ev_in, 0));
  CU(cudaGraphLaunch(g.exec[ph], g.stream));
  CU(cudaEventRecord(g.ev_out, g.stream));
  CU(cudaStreamWaitEvent(M->stream, g.ev_out, 0));
  ++g.replays;
}
#endif

// ====================================================================================== plane-distributed K2
// The global spline solve couples all radii but not the (z-mode, wavenumber) columns, so the columns
// are dealt out by z-mode plane: rank k owns planes [z0[k], z0[k+1]).  Every tile sends each owner the
// contiguous [planes][columns][coefficients] slab of its B (no packing: zb is the slowest index), the
// owner overlap-adds the tiles in the reference order (own block, then the lower neighbour's 3-coefficient
// halo, src/semiimplicit.jl:323-329), solves its planes once, and returns to every tile exactly the
// coefficients that tile evaluates.  Per rank this moves ~2 x (N-1)/N of ONE tile's spectrum instead of
// all-reducing the whole patch, and the solve is no longer replicated (SURVEY 8e option B).
static DevGrid plane_view(const DevGrid& g, int nz) {
  DevGrid d = g;
  d.bz = nz;
  d.S = (long long)nz * g.b_rDim * g.ncolp;
  return d;
}

static void colsolve_init(sb_model* M, int rank, int nranks) {
  auto& cs = M->cs;
  const DevGrid& p = M->patch->dg;
  if (nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("bad rank");
  if (M->ntiles % nranks) throw std::invalid_argument("tiles must divide evenly over ranks");
  cs.rank = rank; cs.nranks = nranks;
  cs.z0.resize(nranks + 1);
  for (int k = 0; k <= nranks; ++k) cs.z0[k] = (int)((long long)p.bz * k / nranks);
  const int per = M->ntiles / nranks, nz = cs.nz();
  cs.tdg.resize(M->ntiles); cs.tile_owner.resize(M->ntiles); cs.recv_off.resize(M->ntiles + 1);
  long long off = 0;
  for (int t = 0; t < M->ntiles; ++t) {
    DevGrid d = p;                        // only the scalar members are used by k_assemble / k_extract
    d.num_cells = (int)M->tile_params[5 * t + 2];
    d.rDim = 3 * d.num_cells; d.b_rDim = d.num_cells + 3;
    d.coefOffset = (int)M->tile_params[5 * t + 3] - 1;
    d.patchOffsetL = 3 * d.coefOffset;
    d.kDim = p.has_l ? d.rDim + d.patchOffsetL : 0;
    d.ncolp = 1 + 2 * d.kDim;
    d.S = (long long)p.bz * d.b_rDim * d.ncolp;
    cs.tdg[t] = d;
    cs.tile_owner[t] = t / per;
    cs.recv_off[t] = off;
    off += (long long)p.V * nz * d.ncolp * d.b_rDim;
  }
  cs.recv_off[M->ntiles] = off;
  cs.slab = plane_view(p, nz);
  cs.recvB = dev_zeros(off, M->stream); M->owned.push_back(cs.recvB);
  cs.sendA = dev_zeros(off, M->stream); M->owned.push_back(cs.sendA);
  cs.slabB = dev_zeros(cs.slab.S * p.V, M->stream); M->owned.push_back(cs.slabB);
  cs.slabA = dev_zeros(cs.slab.S * p.V, M->stream); M->owned.push_back(cs.slabA);
  cs.on = true;
}

// pointer / element count of one message buffer.  what: 0 = B chunk of local tile `tile` (global index) for
// owner `peer`; 1 = where tile `tile`'s B chunk lands on this owner; 2 = solved chunk to return to `tile`;
// 3 = where owner `peer`'s planes land in local tile `tile`'s A; 4 = my solved slab; 5 = owner `peer`'s slab
// inside the replicated patch A (output only).  All per variable v.
static void colsolve_buffer(sb_model* M, int what, int tile, int v, int peer, void** ptr, long long* count) {
  auto& cs = M->cs;
  if (!cs.on) throw std::invalid_argument("sb_model_colsolve_init has not been called");
  const DevGrid& p = M->patch->dg;
  if (v < 0 || v >= p.V) throw std::invalid_argument("bad variable");
  auto local_tile = [&](int t) -> sb_grid* {
    if (t < M->tile_first || t >= M->tile_first + M->tile_count) throw std::invalid_argument("tile is not local");
    return M->tiles[t - M->tile_first].grid;
  };
  const int nz = cs.nz();
  if (what == 0 || what == 3) {
    if (peer < 0 || peer >= cs.nranks || tile < 0 || tile >= M->ntiles) throw std::invalid_argument("bad peer/tile");
    sb_grid* G = local_tile(tile);
    const DevGrid& d = G->dg;
    const long long plane = (long long)d.ncolp * d.b_rDim;
    double* base = (what == 0 ? G->spectralB : G->spectralA) + (long long)v * d.S + cs.z0[peer] * plane;
    *ptr = base; *count = (cs.z0[peer + 1] - cs.z0[peer]) * plane;
  } else if (what == 1 || what == 2) {
    if (tile < 0 || tile >= M->ntiles) throw std::invalid_argument("bad tile");
    const DevGrid& d = cs.tdg[tile];
    const long long per_v = (long long)nz * d.ncolp * d.b_rDim;
    *ptr = (what == 1 ? cs.recvB : cs.sendA) + cs.recv_off[tile] + v * per_v; *count = per_v;
  } else if (what == 4) {
    *ptr = cs.slabA + (long long)v * cs.slab.S; *count = cs.slab.S;
  } else if (what == 5) {
    if (peer < 0 || peer >= cs.nranks) throw std::invalid_argument("bad peer");
    const long long plane = (long long)p.ncolp * p.b_rDim;
    *ptr = M->patch->spectralA + (long long)v * p.S + cs.z0[peer] * plane; *count = (cs.z0[peer + 1] - cs.z0[peer]) * plane;
  } else {
    throw std::invalid_argument("what must be 0..5");
  }
}

// messages that stay on this rank: device copies on the model stream
static void colsolve_local(sb_model* M, int direction /*0: B chunks in, 1: A chunks out*/) {
  auto& cs = M->cs;
  const DevGrid& p = M->patch->dg;
  if (cs.nz() == 0) return;
  for (int t = M->tile_first; t < M->tile_first + M->tile_count; ++t)
    for (int v = 0; v < p.V; ++v) {
      void *a, *b; long long na, nb;
      colsolve_buffer(M, direction ? 3 : 0, t, v, cs.rank, &a, &na);
      colsolve_buffer(M, direction ? 2 : 1, t, v, cs.rank, &b, &nb);
      if (na != nb) throw std::runtime_error("colsolve: chunk size mismatch");
      if (direction) CU(cudaMemcpyAsync(a, b, (size_t)na * sizeof(double), cudaMemcpyDeviceToDevice, M->stream));
      else CU(cudaMemcpyAsync(b, a, (size_t)na * sizeof(double), cudaMemcpyDeviceToDevice, M->stream));
    }
}

// owner side: overlap-add the tiles' chunks, solve my planes, pack what each tile needs
static void colsolve_solve(sb_model* M) {
  auto& cs = M->cs;
  if (!cs.on) throw std::invalid_argument("sb_model_colsolve_init has not been called");
  sb_grid* P = M->patch;
  const DevGrid& p = P->dg;
  const int nz = cs.nz();
  if (nz == 0) return;                  // fewer planes than ranks (grids without a vertical dimension): nothing to own
  if (!cs.p2p) colsolve_local(M, 0);
  CU(cudaMemsetAsync(cs.slabB, 0, (size_t)cs.slab.S * p.V * sizeof(double), M->stream));
  LaunchCtx c = P->ctx();
  for (int t = 0; t < M->ntiles; ++t) {
    DevGrid tv = plane_view(cs.tdg[t], nz);
    if (t > 0) {
      DevGrid pv = plane_view(cs.tdg[t - 1], nz);
      launch_assemble(c, cs.slab, tv, cs.recvB + cs.recv_off[t], &pv, cs.recvB + cs.recv_off[t - 1], 0, cs.slabB);
    } else {
      launch_assemble(c, cs.slab, tv, cs.recvB + cs.recv_off[t], nullptr, nullptr, 0, cs.slabB);
    }
  }
  launch_spline_solve(c, cs.slab, P->d_factors, P->hfactors, cs.slabB, cs.slabA);
  if (cs.p2p) {     // every tile's slice straight into that tile's A, wherever it lives
    // rank k starts with tile k and goes round: at any moment the ranks store into different GPUs.  In tile order every rank
    // wrote into tile 0's GPU first, then all into tile 1's ... -- eight senders on one NVLink ingress, seven receivers idle
    // (N = 8: extract 1.4-1.9 ms on ranks >= 1 against 0.6 ms of work)
    for (int i = 0; i < M->ntiles; ++i) {
      const int t = (cs.rank + i) % M->ntiles;
      const DevGrid& d = cs.tdg[t];
      launch_extract(c, cs.slab, plane_view(d, nz), cs.slabA, cs.peer_tileA[t] + (long long)cs.z0[cs.rank] * d.ncolp * d.b_rDim, d.S);
    }
    return;
  }
  for (int t = 0; t < M->ntiles; ++t) launch_extract(c, cs.slab, plane_view(cs.tdg[t], nz), cs.slabA, cs.sendA + cs.recv_off[t]);
  colsolve_local(M, 1);
}

// output only: my solved planes into the replicated patch A (the other owners' planes arrive by broadcast)
static void colsolve_publish(sb_model* M) {
  auto& cs = M->cs;
  const DevGrid& p = M->patch->dg;
  if (cs.nz() == 0) return;
  for (int v = 0; v < p.V; ++v) {
    void *a, *b; long long na, nb;
    colsolve_buffer(M, 4, 0, v, cs.rank, &a, &na);
    colsolve_buffer(M, 5, 0, v, cs.rank, &b, &nb);
    CU(cudaMemcpyAsync(b, a, (size_t)na * sizeof(double), cudaMemcpyDeviceToDevice, M->stream));
  }
}


// native (library-owned NCCL communicator) exchange of one message set: dir 0 = B chunks to the plane owners,
// dir 1 = solved chunks back to the tiles.  One ncclGroup = one fused launch over NVLink.
static void colsolve_p2p(sb_model* M, int dir) {
  auto& cs = M->cs;
  if (cs.nranks <= 1) return;
  if (!M->comm || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart) throw CommError("NCCL point-to-point is not available");
  const int V = M->patch->dg.V;
  NC(g_nccl.GroupStart());
  for (int t = 0; t < M->ntiles; ++t) {
    const bool mine = cs.tile_owner[t] == cs.rank;
    for (int v = 0; v < V; ++v) {
      void* ptr; long long n;
      if (mine) {          // I hold tile t: talk to every other plane owner
        for (int k = 0; k < cs.nranks; ++k) {
          if (k == cs.rank) continue;
          colsolve_buffer(M, dir ? 3 : 0, t, v, k, &ptr, &n);
          if (n == 0) continue;
          if (dir) NC(g_nccl.Recv(ptr, (size_t)n, /*ncclFloat64*/ 8, k, M->comm, (void*)M->stream));
          else NC(g_nccl.Send(ptr, (size_t)n, 8, k, M->comm, (void*)M->stream));
        }
      } else {             // tile t lives elsewhere: its chunk of my planes
        colsolve_buffer(M, dir ? 2 : 1, t, v, cs.rank, &ptr, &n);
        if (n == 0) continue;
        if (dir) NC(g_nccl.Send(ptr, (size_t)n, 8, cs.tile_owner[t], M->comm, (void*)M->stream));
        else NC(g_nccl.Recv(ptr, (size_t)n, 8, cs.tile_owner[t], M->comm, (void*)M->stream));
      }
    }
  }
  NC(g_nccl.GroupEnd());
  ++M->extra_launches;
}

// stream-ordered rendezvous of all ranks: every rank's kernels enqueued before it (and their stores into peer
// memory) are complete when it returns on the stream
static void comm_barrier(sb_model* M) {
  if (M->nranks <= 1) return;
  if (!M->comm) throw CommError("library communicator not initialised (sb_model_comm_init)");
  if (!M->d_barrier) throw CommError("no barrier buffer");
  NC(g_nccl.AllReduce(M->d_barrier, M->d_barrier, 1, /*ncclFloat64*/ 8, /*ncclSum*/ 0, M->comm, (void*)M->stream));
  ++M->extra_launches;
}

static void model_exchange(sb_model* M) {
  if (M->cs.on && M->cs.p2p) {
    comm_barrier(M);        // every tile's fwd_r has stored its B into the owners' buffers
    colsolve_solve(M);
    comm_barrier(M);        // every owner's extract has stored A into the tiles
    return;
  }
  if (M->cs.on) {
    colsolve_p2p(M, 0);
    colsolve_solve(M);
    colsolve_p2p(M, 1);
    return;
  }
  if (M->nranks <= 1) return;
  sb_grid* P = M->patch;
  NC(g_nccl.AllReduce(P->spectralB, P->spectralB, (size_t)P->dg.S * P->dg.V, /*ncclFloat64*/ 8, /*ncclSum*/ 0, M->comm,
                      (void*)P->stream));
  ++M->extra_launches;
}



// ---- peer-memory mode: CUDA IPC handles of the buffers other ranks write into
static void p2p_handle(sb_model* M, int what, int tile, void* out64) {
  auto& cs = M->cs;
  if (!cs.on || !out64) throw std::invalid_argument("bad argument");
#ifdef SB_EMU
  (void)what; (void)tile;
  throw Unsupported("CUDA IPC is not available in the CPU emulation build");
#else
  void* ptr = nullptr;
  if (what == 0) ptr = cs.recvB;
  else if (what == 1) {
    if (tile < M->tile_first || tile >= M->tile_first + M->tile_count) throw std::invalid_argument("tile is not local");
    ptr = M->tiles[tile - M->tile_first].grid->spectralA;
  } else throw std::invalid_argument("what must be 0 or 1");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CU(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(out64), ptr));
#endif
}

static void p2p_open(sb_model* M, int what, int index, const void* handle64) {
  auto& cs = M->cs;
  if (!cs.on || !handle64) throw std::invalid_argument("bad argument");
#ifdef SB_EMU
  (void)what; (void)index;
  throw Unsupported("CUDA IPC is not available in the CPU emulation build");
#else
  cs.peer_recvB.resize(cs.nranks, nullptr);
  cs.peer_tileA.resize(M->ntiles, nullptr);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  void* ptr = nullptr;
  CU(cudaSetDevice(M->device));
  CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  cs.ipc_opened.push_back(ptr);
  if (what == 0) {
    if (index < 0 || index >= cs.nranks) throw std::invalid_argument("bad peer");
    cs.peer_recvB[index] = (double*)ptr;
  } else if (what == 1) {
    if (index < 0 || index >= M->ntiles) throw std::invalid_argument("bad tile");
    cs.peer_tileA[index] = (double*)ptr;
  } else throw std::invalid_argument("what must be 0 or 1");
#endif
}

static void p2p_enable(sb_model* M) {
  auto& cs = M->cs;
  if (!cs.on) throw std::invalid_argument("sb_model_colsolve_init has not been called");
  if (cs.nranks > SB_MAX_PEERS) throw Unsupported("peer-memory exchange supports at most 8 ranks");
  const DevGrid& p = M->patch->dg;
  cs.peer_recvB.resize(cs.nranks, nullptr);
  cs.peer_tileA.resize(M->ntiles, nullptr);
  cs.peer_recvB[cs.rank] = cs.recvB;
  for (int t = M->tile_first; t < M->tile_first + M->tile_count; ++t) cs.peer_tileA[t] = M->tiles[t - M->tile_first].grid->spectralA;
  for (int k = 0; k < cs.nranks; ++k)
    if (!cs.peer_recvB[k]) throw std::invalid_argument("peer buffer of rank " + std::to_string(k) + " has not been opened");
  for (int t = 0; t < M->ntiles; ++t)
    if (!cs.peer_tileA[t]) throw std::invalid_argument("spectral A of tile " + std::to_string(t) + " has not been opened");
  cs.scat.assign(M->tile_count, PeerScatter{});
  for (int i = 0; i < M->tile_count; ++i) {
    const int t = M->tile_first + i;
    const DevGrid& d = cs.tdg[t];
    PeerScatter& ps = cs.scat[i];
    ps.nranks = cs.nranks;
    for (int k = 0; k <= cs.nranks; ++k) ps.z0[k] = cs.z0[k];
    for (int k = 0; k < cs.nranks; ++k) {
      const long long nzk = cs.z0[k + 1] - cs.z0[k];
      long long off = 0;     // owner k's receive slots are laid out tile by tile: [t'][V][nz_k][ncolp_t'][M_t']
      for (int tt = 0; tt < t; ++tt) off += (long long)p.V * nzk * cs.tdg[tt].ncolp * cs.tdg[tt].b_rDim;
      ps.base[k] = cs.peer_recvB[k] + off;
      ps.vstride[k] = nzk * d.ncolp * d.b_rDim;
    }
    M->tiles[i].grid->scatter = &ps;
  }
  cs.p2p = true;
}

// ====================================================================================== Chebyshev column API
static ChebTables cheb_tables_of(const sb_cheb_params* cp, int* bz_out) {
  if (!cp) throw std::invalid_argument("NULL Chebyshev parameters");
  if (cp->zDim < 4 || !(cp->zmax > cp->zmin)) throw std::invalid_argument("zDim >= 4 and zmax > zmin required");
  if (cp->BCB < 0 || cp->BCB > 3 || cp->BCT < 0 || cp->BCT > 3) throw std::invalid_argument("bad Chebyshev BC code");
  const long long def = std::min<long long>(cp->zDim, (2 * cp->zDim - 1) / 3 + 1);
  const int bz = (int)(cp->b_zDim > 0 ? cp->b_zDim : def);
  if (bz > cp->zDim) throw std::invalid_argument("b_zDim > zDim");
  *bz_out = bz;
  return make_cheb_tables((int)cp->zDim, bz, cp->zmin, cp->zmax);
}

static void cheb_columns(const sb_cheb_params* cp, int op, const double* in, double* out, long long ncols, double C0, int device) {
  if (!in || !out || ncols < 0) throw std::invalid_argument("bad argument");
  require_device();
  int bz = 0;
  const ChebTables t = cheb_tables_of(cp, &bz);
  const int nz = t.nz;
  std::vector<double> M;
  int rows = nz, cols = nz;
  switch (op) {
    case SB_CHEB_CB: rows = bz; cols = nz; M = t.fwd; break;
    case SB_CHEB_CA: {   // a = (I + Gamma) b, zero filled to zDim
      rows = nz; cols = bz;
      const std::vector<double> IG = cheb_bc_matrix(t, cp->BCB, cp->BCT);
      M.assign((size_t)nz * bz, 0.0);
      std::copy(IG.begin(), IG.end(), M.begin());
      break;
    }
    case SB_CHEB_CI: M = t.T0; break;
    case SB_CHEB_CIX: M = t.T1; break;
    case SB_CHEB_CIXX: M = t.T2; break;
    case SB_CHEB_CIINT: M = t.Tint; break;
    default: throw std::invalid_argument("unknown Chebyshev column operation");
  }
  if (ncols == 0) return;
  CU(cudaSetDevice(device));
  double *dM = dev_upload(M), *din = nullptr, *dout = nullptr;
  CU(cudaMalloc((void**)&din, (size_t)cols * ncols * sizeof(double)));
  CU(cudaMalloc((void**)&dout, (size_t)rows * ncols * sizeof(double)));
  try {
    CU(cudaMemcpy(din, in, (size_t)cols * ncols * sizeof(double), cudaMemcpyHostToDevice));
    LaunchCtx c{nullptr, nullptr, nullptr};
    launch_column_op(c, dM, rows, cols, din, dout, ncols, op == SB_CHEB_CIINT ? C0 : 0.0);
    CU(cudaMemcpy(out, dout, (size_t)rows * ncols * sizeof(double), cudaMemcpyDeviceToHost));
  } catch (...) {
    cudaFree(dM); cudaFree(din); cudaFree(dout);
    throw;
  }
  cudaFree(dM); cudaFree(din); cudaFree(dout);
}

// ====================================================================================== C ABI
extern "C" {

const char* sb_last_error(void) { return g_err.c_str(); }
const char* sb_version(void) {
#ifdef SB_EMU
  return "scythe_b200 0.1 (cpu-emulation test build)";
#else
  return "scythe_b200 0.1 (sm_100a)";
#endif
}
int sb_device_count(void) {
#ifdef SB_EMU
  return 1;
#else
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
#endif
}

int sb_grid_create(const sb_grid_params* gp, int device, void* stream, sb_grid_t* out) {
  return guarded([&] { if (!out) throw std::invalid_argument("out is NULL"); *out = grid_new(gp, device, stream); });
}
int sb_grid_destroy(sb_grid_t g) { return guarded([&] { grid_free(g); }); }
int sb_grid_get_info(sb_grid_t g, sb_grid_info* o) {
  return guarded([&] {
    if (!g || !o) throw std::invalid_argument("NULL argument");
    const DevGrid& d = g->dg;
    o->N = d.N; o->V = d.V; o->D = d.D; o->S = d.S; o->rDim = d.rDim; o->b_rDim = d.b_rDim;
    o->zDim = d.has_z ? d.zDim : 0; o->b_zDim = d.has_z ? d.bz : 0; o->kDim = d.kDim;
    o->lDim = d.has_l ? d.hpoints : 0; o->num_columns = d.has_z ? d.hpoints : 0;
    o->patchOffsetL = d.patchOffsetL; o->ndims = g->ndims;
  });
}
int sb_grid_get_gridpoints(sb_grid_t g, double* out, int64_t n) {
  return guarded([&] {
    if (!g || !out) throw std::invalid_argument("NULL argument");
    if (n < g->dg.N * g->ndims) throw std::invalid_argument("gridpoints buffer too small");
    fill_gridpoints(g, out);
  });
}
static void check_slots(sb_grid_t g, const void* host, int slot0, int nslots) {
  if (!g || !host) throw std::invalid_argument("NULL argument");
  if (slot0 < 0 || nslots < 1 || slot0 + nslots > g->dg.D) throw std::invalid_argument("slot range outside 0..D-1");
}
int sb_grid_set_physical(sb_grid_t g, const double* host, int32_t slot0, int32_t nslots) {
  return guarded([&] {
    check_slots(g, host, slot0, nslots);
    g->ensure_physical();
    if (slot0 == 0) g->slot0_src = nullptr;
    CU(cudaMemcpyAsync(g->physical + g->slot_stride() * slot0, host, (size_t)g->slot_stride() * nslots * sizeof(double),
                       cudaMemcpyHostToDevice, g->stream));
    CU(cudaStreamSynchronize(g->stream));
  });
}
int sb_grid_get_physical(sb_grid_t g, double* host, int32_t slot0, int32_t nslots) {
  return guarded([&] {
    check_slots(g, host, slot0, nslots);
    g->ensure_physical();
    if (slot0 == 0) g->materialize_slot0();
    CU(cudaMemcpyAsync(host, g->physical + g->slot_stride() * slot0, (size_t)g->slot_stride() * nslots * sizeof(double),
                       cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
  });
}
int sb_grid_set_spectral(sb_grid_t g, int32_t which, const double* host) {
  return guarded([&] {
    if (!g || !host || which < 0 || which > 1) throw std::invalid_argument("bad argument");
    CU(cudaMemcpyAsync(which ? g->spectralA : g->spectralB, host, (size_t)g->dg.S * g->dg.V * sizeof(double),
                       cudaMemcpyHostToDevice, g->stream));
    CU(cudaStreamSynchronize(g->stream));
  });
}
int sb_grid_get_spectral(sb_grid_t g, int32_t which, double* host) {
  return guarded([&] {
    if (!g || !host || which < 0 || which > 1) throw std::invalid_argument("bad argument");
    CU(cudaMemcpyAsync(host, which ? g->spectralA : g->spectralB, (size_t)g->dg.S * g->dg.V * sizeof(double),
                       cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
  });
}
int sb_spectral_transform(sb_grid_t g) {
  return guarded([&] {
    if (!g) throw std::invalid_argument("NULL grid");
    g->ensure_physical();
    g->materialize_slot0();
    grid_forward(g, g->physical, nullptr);
  });
}
int sb_grid_transform(sb_grid_t g) {
  return guarded([&] { if (!g) throw std::invalid_argument("NULL grid"); grid_spline(g, g->spectralB); grid_inverse(g, g); });
}
int sb_spline_transform(sb_grid_t patch, sb_grid_t shared) {
  return guarded([&] {
    if (!patch) throw std::invalid_argument("NULL grid");
    sb_grid* s = shared ? shared : patch;
    if (s->dg.S != patch->dg.S || s->dg.V != patch->dg.V) throw std::invalid_argument("shared spectral array has the wrong size");
    grid_spline(patch, s->spectralB);
  });
}
int sb_tile_transform(sb_grid_t patch, sb_grid_t tile) {
  return guarded([&] { if (!patch || !tile) throw std::invalid_argument("NULL grid"); grid_inverse(patch, tile); });
}
int sb_calc_tile_sizes(const sb_grid_params* patch, int32_t ntiles, double* out) {
  return guarded([&] { calc_tile_sizes(patch, ntiles, out); });
}
int sb_shared_clear(sb_grid_t patch) {
  return guarded([&] {
    if (!patch) throw std::invalid_argument("NULL grid");
    CU(cudaMemsetAsync(patch->spectralB, 0, (size_t)patch->dg.S * patch->dg.V * sizeof(double), patch->stream));
  });
}
int sb_shared_assemble(sb_grid_t patch, sb_grid_t tile, sb_grid_t prev, int32_t last) {
  return guarded([&] {
    if (!patch || !tile) throw std::invalid_argument("NULL grid");
    launch_assemble(tile->ctx(), patch->dg, tile->dg, tile->spectralB, prev ? &prev->dg : nullptr,
                    prev ? prev->spectralB : nullptr, last, patch->spectralB);
  });
}
int sb_check_cfl(sb_grid_t g, int32_t* var, int64_t* index) {
  return guarded([&] { if (!g) throw std::invalid_argument("NULL grid"); check_cfl(g, var, index); });
}
int sb_grid_sync(sb_grid_t g) {
  return guarded([&] { if (!g) throw std::invalid_argument("NULL grid"); CU(cudaStreamSynchronize(g->stream)); });
}
int sb_grid_device_ptr(sb_grid_t g, int32_t which, void** ptr, int64_t* n) {
  return guarded([&] {
    if (!g || !ptr) throw std::invalid_argument("NULL argument");
    switch (which) {
      case 0: g->ensure_physical(); g->materialize_slot0(); *ptr = g->physical; if (n) *n = g->dg.N * g->dg.V * g->dg.D; break;
      case 1: *ptr = g->spectralB; if (n) *n = g->dg.S * g->dg.V; break;
      case 2: *ptr = g->spectralA; if (n) *n = g->dg.S * g->dg.V; break;
      default: throw std::invalid_argument("which must be 0, 1 or 2");
    }
  });
}

int sb_model_create(const sb_model_params* mp, int32_t ntiles, int32_t tile_first, int32_t tile_count, int device,
                    void* stream, sb_model_t* out) {
  return guarded([&] { if (!out) throw std::invalid_argument("out is NULL"); *out = model_new(mp, ntiles, tile_first, tile_count, device, stream); });
}
int sb_model_destroy(sb_model_t m) { return guarded([&] { model_free(m); }); }
int sb_model_initialize(sb_model_t m, const double* ic) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); model_initialize(m, ic); });
}
int sb_model_patch(sb_model_t m, sb_grid_t* out) {
  return guarded([&] { if (!m || !out) throw std::invalid_argument("NULL argument"); *out = m->patch; });
}
int sb_model_tile(sb_model_t m, int32_t i, sb_grid_t* out) {
  return guarded([&] {
    if (!m || !out || i < 0 || i >= (int)m->tiles.size()) throw std::invalid_argument("bad tile index");
    *out = m->tiles[i].grid;
  });
}
int sb_model_advance_tiles(sb_model_t m, int64_t t) {
  return guarded([&] { if (!m || t < 1) throw std::invalid_argument("bad argument"); model_advance_tiles(m, t); });
}
int sb_model_exchange(sb_model_t m) {
  try {
    if (!m) return fail(SB_EINVAL, "NULL model");
    model_exchange(m);
    return SB_OK;
  } catch (const std::exception& e) { return fail(SB_ECOMM, e.what()); }
}
int sb_model_spline_transform(sb_model_t m) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); grid_spline(m->patch, m->patch->spectralB); });
}
int sb_model_step(sb_model_t m, int64_t t) {
  try {
    if (!m || t < 1) return fail(SB_EINVAL, "bad argument");
    model_step(m, t);
    return SB_OK;
  } catch (const CommError& e) { return fail(SB_ECOMM, e.what());
  } catch (const std::exception& e) { return fail(SB_ECUDA, e.what()); }
}
int sb_model_run(sb_model_t m, int64_t t0, int64_t nsteps) {
  try {
    if (!m || t0 < 1 || nsteps < 0) return fail(SB_EINVAL, "bad argument");
    int64_t t = t0;
    while (t < t0 + nsteps) {
#ifndef SB_EMU
      if (t >= 4 && t0 + nsteps - t >= 3 && step_graph_usable(m)) { step_graph_run3(m, t); t += 3; continue; }
#endif
      model_step(m, t);
      ++t;
    }
    return SB_OK;
  } catch (const CommError& e) { return fail(SB_ECOMM, e.what());
  } catch (const std::exception& e) { return fail(SB_ECUDA, e.what()); }
}
int sb_model_output(sb_model_t m, double* host) {
  return guarded([&] {
    if (!m) throw std::invalid_argument("NULL model");
    sb_grid* P = m->patch;
    if (m->cs.on) colsolve_publish(m);   // (other owners' planes: broadcast by the caller beforehand)
    grid_inverse(P, P);
    check_cfl(P, nullptr, nullptr);
    if (host) {
      CU(cudaMemcpyAsync(host, P->physical, (size_t)P->dg.N * P->dg.V * P->dg.D * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
      CU(cudaStreamSynchronize(P->stream));
    }
  });
}
int sb_model_get_state(sb_model_t m, int32_t tile, int32_t which, double* host) {
  return guarded([&] {
    if (!m || !host || tile < 0 || tile >= (int)m->tiles.size() || which < 0 || which > 6) throw std::invalid_argument("bad argument");
    TileState& T = m->tiles[tile];
    const double* src = which == 0 ? T.var_np1 : (which <= 3 ? T.expd[which - 1] : T.impd[which - 4]);
    // after a step the rotation has already happened: expd[0] is scratch for the next step, so
    // "expdot_n" of the step just taken is expd[1] (== expdot_nm1, as the reference leaves it)
    if (which == 1) src = T.expd[1];
    if (which == 4) src = T.impd[1];
    if (!src) throw std::invalid_argument("state array not allocated (semi-implicit off)");
    CU(cudaMemcpyAsync(host, src, (size_t)T.grid->dg.N * T.grid->dg.V * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
  });
}
int sb_model_set_state(sb_model_t m, int32_t tile, int32_t which, const double* host) {
  return guarded([&] {
    if (!m || !host || tile < 0 || tile >= (int)m->tiles.size() || which < 0 || which > 6 || which == 1 || which == 4)
      throw std::invalid_argument("bad argument (which: 0 var_np1, 2/3 expdot_nm1/nm2, 5/6 impdot_nm1/nm2)");
    TileState& T = m->tiles[tile];
    // history as sb_model_get_state reports it after a step: nm1 = buffer 1, nm2 = buffer 2 of the rotation
    double* dst = which == 0 ? T.var_np1 : (which <= 3 ? T.expd[which - 1] : T.impd[which - 4]);
    if (!dst) throw std::invalid_argument("state array not allocated (semi-implicit off)");
    CU(cudaMemcpyAsync(dst, host, (size_t)T.grid->dg.N * T.grid->dg.V * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    if (which == 2 || which == 3) {     // a tendency history from outside: variables without a tendency keep the all-zeros
      const unsigned pm = passive_mask(m);   // guarantee only if what was stored for them is all zeros too (a restart file)
      const long long N = T.grid->dg.N;
      for (int v = 0; v < T.grid->dg.V && T.hist_untouched; ++v)
        if ((pm >> v) & 1u)
          for (long long i = 0; i < N; ++i)
            if (host[(long long)v * N + i] != 0.0) { T.hist_untouched = false; break; }
#ifndef SB_EMU
      if (!T.hist_untouched)      // step graphs captured so far have the zero-history form of the kernels baked in
        for (int k = 0; k < 3; ++k)
          if (m->graph.exec[k]) { cudaGraphExecDestroy(m->graph.exec[k]); m->graph.exec[k] = nullptr; }
#endif
    }
  });
}
static void pipe_open(sb_model* m, TileState& T) {
  HostPipe& p = T.pipe;
  if (p.on) return;
  const size_t bytes = (size_t)T.grid->dg.N * T.grid->dg.V * sizeof(double);
  (void)m;
  try {
    CU(cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));    // non-blocking: the compute stream may be the legacy
    CU(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));   // default stream, which would serialise blocking streams
    for (int k = 0; k < 2; ++k) {
      CU(cudaMalloc((void**)&p.in[k], bytes));
      CU(cudaMalloc((void**)&p.out[k], bytes));
      CU(cudaEventCreate(&p.in_ready[k])); CU(cudaEventCreate(&p.in_free[k]));
      CU(cudaEventCreate(&p.out_ready[k])); CU(cudaEventCreate(&p.out_free[k]));
    }
  } catch (...) {   // e.g. no room for the four staging buffers: leave nothing behind, the blocking calls still work
    cudaGetLastError();
    for (int k = 0; k < 2; ++k) {
      if (p.in[k]) cudaFree(p.in[k]);
      if (p.out[k]) cudaFree(p.out[k]);
      if (p.in_ready[k]) cudaEventDestroy(p.in_ready[k]);
      if (p.in_free[k]) cudaEventDestroy(p.in_free[k]);
      if (p.out_ready[k]) cudaEventDestroy(p.out_ready[k]);
      if (p.out_free[k]) cudaEventDestroy(p.out_free[k]);
    }
    if (p.s_in) cudaStreamDestroy(p.s_in);
    if (p.s_out) cudaStreamDestroy(p.s_out);
    p = HostPipe{};
    throw;
  }
  p.on = true;
}
int sb_model_stage_in(sb_model_t m, int32_t tile, const double* host) {
  return guarded([&] {
    if (!m || !host || tile < 0 || tile >= (int)m->tiles.size()) throw std::invalid_argument("bad argument");
    TileState& T = m->tiles[tile];
    pipe_open(m, T);
    HostPipe& p = T.pipe;
    const int k = p.ki;
    const size_t bytes = (size_t)T.grid->dg.N * T.grid->dg.V * sizeof(double);
    if (p.in_used[k]) CU(cudaStreamWaitEvent(p.s_in, p.in_free[k], 0));    // the step two back has consumed this buffer
    CU(cudaMemcpyAsync(p.in[k], host, bytes, cudaMemcpyHostToDevice, p.s_in));
    CU(cudaEventRecord(p.in_ready[k], p.s_in));
    CU(cudaStreamWaitEvent(m->stream, p.in_ready[k], 0));
    CU(cudaMemcpyAsync(T.var_np1, p.in[k], bytes, cudaMemcpyDeviceToDevice, m->stream));
    CU(cudaEventRecord(p.in_free[k], m->stream));
    p.in_used[k] = true;
    p.ki ^= 1;
  });
}
int sb_model_stage_out(sb_model_t m, int32_t tile, double* host) {
  return guarded([&] {
    if (!m || !host || tile < 0 || tile >= (int)m->tiles.size()) throw std::invalid_argument("bad argument");
    TileState& T = m->tiles[tile];
    pipe_open(m, T);
    HostPipe& p = T.pipe;
    const int k = p.ko;
    const size_t bytes = (size_t)T.grid->dg.N * T.grid->dg.V * sizeof(double);
    if (p.out_used[k]) CU(cudaStreamWaitEvent(m->stream, p.out_free[k], 0));   // its previous D2H copy has left the buffer
    CU(cudaMemcpyAsync(p.out[k], T.var_np1, bytes, cudaMemcpyDeviceToDevice, m->stream));
    CU(cudaEventRecord(p.out_ready[k], m->stream));
    CU(cudaStreamWaitEvent(p.s_out, p.out_ready[k], 0));
    CU(cudaMemcpyAsync(host, p.out[k], bytes, cudaMemcpyDeviceToHost, p.s_out));
    CU(cudaEventRecord(p.out_free[k], p.s_out));
    p.out_used[k] = true;
    p.ko ^= 1;
  });
}
int sb_model_stage_drain(sb_model_t m, int32_t block) {
  return guarded([&] {
    if (!m) throw std::invalid_argument("NULL model");
    for (auto& T : m->tiles) {
      HostPipe& p = T.pipe;
      if (!p.on) continue;
      for (int k = 0; k < 2; ++k) {
        if (p.out_used[k]) CU(cudaStreamWaitEvent(m->stream, p.out_free[k], 0));
        if (p.in_used[k]) CU(cudaStreamWaitEvent(m->stream, p.in_free[k], 0));
      }
    }
    if (block) CU(cudaStreamSynchronize(m->stream));
  });
}
int sb_model_tendency(sb_model_t m) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); tiles_tendency(m); });
}
int sb_model_physics(sb_model_t m, int64_t t) {
  return guarded([&] { if (!m || t < 1) throw std::invalid_argument("bad argument"); tiles_physics(m, t); });
}
int sb_model_cycle(sb_model_t m, int64_t t) {
  try {
    if (!m || t < 1) return fail(SB_EINVAL, "bad argument");
    tiles_tendency(m);
    model_exchange(m);
    if (!m->cs.on) grid_spline(m->patch, m->patch->spectralB);
    tiles_physics(m, t);
    return SB_OK;
  } catch (const CommError& e) { return fail(SB_ECOMM, e.what());
  } catch (const std::exception& e) { return fail(SB_ECUDA, e.what()); }
}
int sb_model_set_k3_slots(sb_model_t m, int32_t mode) {
  return guarded([&] {
    if (!m) throw std::invalid_argument("NULL model");
    if (mode < 0 || mode > 3) throw std::invalid_argument("k3 slot mode must be 0 (fused/needed), 1 (all), 2 (needed, rest poisoned) or 3 (needed)");
    m->k3_slots = mode;
  });
}
int sb_model_profile(sb_model_t m, int32_t on) {
  return guarded([&] {
    if (!m) throw std::invalid_argument("NULL model");
    m->patch->prof.on = on != 0;
    for (auto& t : m->tiles) t.grid->prof.on = on != 0;
  });
}
int sb_model_profile_report(sb_model_t m, char* buf, int64_t buflen) {
  return guarded([&] {
    if (!m || !buf || buflen < 2) throw std::invalid_argument("bad argument");
    CU(cudaStreamSynchronize(m->stream));
    std::map<std::string, std::pair<long long, double>> acc;
    auto drain = [&](sb_grid* g) {
      for (auto& r : g->prof.recs) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = acc[r.name];
        e.first += 1;
        e.second += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
      }
      g->prof.recs.clear();
    };
    drain(m->patch);
    for (auto& t : m->tiles) drain(t.grid);
    std::string s;
    for (auto& kv : acc) s += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
    if ((int64_t)s.size() + 1 > buflen) s.resize((size_t)buflen - 1);
    std::memcpy(buf, s.c_str(), s.size() + 1);
  });
}
int sb_model_sync(sb_model_t m) {
  return guarded([&] {
    if (!m) throw std::invalid_argument("NULL model");
    CU(cudaStreamSynchronize(m->stream));
    for (auto& T : m->tiles)
      if (T.pipe.on) { CU(cudaStreamSynchronize(T.pipe.s_in)); CU(cudaStreamSynchronize(T.pipe.s_out)); }
  });
}
int64_t sb_model_launch_count(sb_model_t m) {
  if (!m) return 0;
  long long n = m->patch->launches + m->extra_launches;
  for (auto& t : m->tiles) n += t.grid->launches;
  return n;
}


int sb_model_colsolve_init(sb_model_t m, int32_t rank, int32_t nranks) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); colsolve_init(m, rank, nranks); });
}
int sb_model_colsolve_buffer(sb_model_t m, int32_t what, int32_t tile, int32_t v, int32_t peer, void** ptr, int64_t* count) {
  return guarded([&] {
    if (!m || !ptr || !count) throw std::invalid_argument("NULL argument");
    long long n = 0;
    colsolve_buffer(m, what, tile, v, peer, ptr, &n);
    *count = n;
  });
}
int sb_model_colsolve_planes(sb_model_t m, int32_t* z0, int32_t n) {
  return guarded([&] {
    if (!m || !z0 || !m->cs.on || n < m->cs.nranks + 1) throw std::invalid_argument("bad argument");
    for (int k = 0; k <= m->cs.nranks; ++k) z0[k] = m->cs.z0[k];
  });
}
int sb_model_colsolve_solve(sb_model_t m) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); colsolve_solve(m); });
}
int sb_model_colsolve_publish(sb_model_t m) {
  return guarded([&] { if (!m || !m->cs.on) throw std::invalid_argument("bad argument"); colsolve_publish(m); });
}

int sb_model_ipc_handle(sb_model_t m, int32_t what, int32_t tile, void* out64) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); p2p_handle(m, what, tile, out64); });
}
int sb_model_ipc_open(sb_model_t m, int32_t what, int32_t index, const void* handle64) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); p2p_open(m, what, index, handle64); });
}
int sb_model_p2p_enable(sb_model_t m) {
  return guarded([&] { if (!m) throw std::invalid_argument("NULL model"); p2p_enable(m); });
}

int sb_cheb_mish_points(const sb_cheb_params* cp, double* z) {
  return guarded([&] {
    if (!z) throw std::invalid_argument("NULL argument");
    int bz;
    const ChebTables t = cheb_tables_of(cp, &bz);
    std::copy(t.z.begin(), t.z.end(), z);
  });
}
int sb_cheb_matrices(const sb_cheb_params* cp, double* dct, double* dct1, double* dct2) {
  return guarded([&] {
    int bz;
    const ChebTables t = cheb_tables_of(cp, &bz);
    const int nz = t.nz;
    double* dst[3] = {dct, dct1, dct2};
    const std::vector<double>* src[3] = {&t.T0, &t.T1, &t.T2};
    for (int m = 0; m < 3; ++m)
      if (dst[m])
        for (int j = 0; j < nz; ++j)
          for (int k = 0; k < nz; ++k) dst[m][(size_t)k * nz + j] = (*src[m])[(size_t)j * nz + k];   // column-major
  });
}
int sb_cheb_columns(const sb_cheb_params* cp, int32_t op, const double* in, double* out, int64_t ncols, double C0, int device) {
  return guarded([&] { cheb_columns(cp, op, in, out, ncols, C0, device); });
}

int sb_comm_unique_id(void* out128) {
  try {
    if (!out128) return fail(SB_EINVAL, "NULL argument");
    nccl_load();
    NC(g_nccl.GetUniqueId(out128));
    return SB_OK;
  } catch (const std::exception& e) { return fail(SB_ECOMM, e.what()); }
}
int sb_model_comm_init(sb_model_t m, const void* id128, int32_t rank, int32_t nranks) {
  try {
    if (!m || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(SB_EINVAL, "bad argument");
    nccl_load();
    Uid id;
    std::memcpy(&id, id128, sizeof(id));
    cudaSetDevice(m->device);
    NC(g_nccl.CommInitRank(&m->comm, nranks, id, rank));
    if (!m->d_barrier) { m->d_barrier = dev_zeros(8, m->stream); m->owned.push_back(m->d_barrier); }
    m->rank = rank;
    m->nranks = nranks;
    return SB_OK;
  } catch (const std::exception& e) { return fail(SB_ECOMM, e.what()); }
}

int sb_timer_start(sb_grid_t g) {
  return guarded([&] { if (!g) throw std::invalid_argument("NULL grid"); CU(cudaEventRecord(g->ev0, g->stream)); });
}
int sb_timer_stop(sb_grid_t g, float* ms) {
  return guarded([&] {
    if (!g || !ms) throw std::invalid_argument("NULL argument");
    CU(cudaEventRecord(g->ev1, g->stream));
    CU(cudaEventSynchronize(g->ev1));
    CU(cudaEventElapsedTime(ms, g->ev0, g->ev1));
  });
}

}  // extern "C"
