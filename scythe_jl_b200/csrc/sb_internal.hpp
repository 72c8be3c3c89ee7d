// sb_internal.hpp -- structures shared between host table generation, kernels and the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cuda_emu.h"

namespace sb {

// ---------------------------------------------------------------- host-side tables
struct SplineFactor {        // Gamma (P+Q) Gamma^T = L L^T for one (BCL,BCR) pair
  int M = 0;                 // b_rDim
  int rL = 0, rR = 0;        // BC ranks folded on the left / right
  int nfree = 0;
  bool periodic = false;
  double foldL[2] = {0, 0};  // rank1: (alpha1,beta1); rank2: (alpha2,beta2)
  double foldR[2] = {0, 0};
  std::vector<double> chol;  // [nfree][4]: {1/L_ii, L_{i,i-1}, L_{i,i-2}, L_{i,i-3}}
  std::vector<double> dense; // periodic only: [M][M] row-major  T = Gamma^T G^-1 Gamma
};
SplineFactor make_spline_factor(int num_cells, double DX, double l_q, int bcl, int bcr);
// basis weights at the 3 mish points of ANY cell: phi[d][mu][j] (d-th derivative of node
// (cell-1+j) at mish point mu), wq[mu] quadrature weights (incl. DX)
void spline_weights(double DX, double phi[3][3][4], double wq[3]);
void spline_mish_points(double xmin, double DX, int num_cells, std::vector<double>& r);

struct ChebTables {
  int nz = 0, bz = 0;
  std::vector<double> z;      // [nz] mish points
  std::vector<double> fwd;    // [bz][nz]   b = fwd u            (CB)
  std::vector<double> T0, T1, T2, Tint;  // [nz][nz]: value / d/dz / d2/dz2 / integral-from-bottom of mode k at z_j
};
ChebTables make_cheb_tables(int nz, int bz, double zmin, double zmax);
// (I + Gamma) [bz][bz] for one (BCB,BCT) pair (CA)
std::vector<double> cheb_bc_matrix(const ChebTables& t, int bcb, int bct);
// dense helpers (row-major)
void matmul(const double* A, const double* B, double* C, int n, int k, int m);  // C[n][m] = A[n][k] B[k][m]
bool lu_factor(std::vector<double>& A, std::vector<int>& piv, int n);
void lu_solve(const std::vector<double>& LU, const std::vector<int>& piv, int n, double* b);
bool invert(std::vector<double>& A, int n);

struct FftClass {            // one convolution length: 2^a, or 3 * 2^a (radix-3 split over three sub-FFT teams)
  int L = 0, log2L = 0;
  int R = 1, L2 = 0;         // L = R * L2, L2 a power of two
  std::vector<double> tw;    // [L] complex: exp(-2 pi i t / L)
  bool fast = false;         // L in 256..8192: register-resident radix-16 kernel (sb_ringfft.cu)
  std::vector<double> twp;   // fast: per-pass twiddle tables
  int twoff[4] = {0, 0, 0, 0};
};
// fast-kernel configuration / host mirrors (sb_ringfft.cu)
bool fast_class_supported(int L);
void fast_class_config(int L, int* log2L, int* nfull, int* rf, int* T, int* nteams, int* iters, int* nrows);
void fast_class_twiddles(int L, std::vector<double>& tab, int off[4]);
void host_fft_dif16(double* x, int L);   // natural -> the fast kernel's digit-reversed order
struct RingPlan {            // Bluestein plan of one ring (sub-DFT length m = n/4)
  int n = 0, m = 0, L = 0, cls = 0;
  long long off = 0;         // offset (in doubles) of this ring's tables in the blob
  long long off2 = 0;        // offset of the pre-combined v2 tables (PQ[4][2m] | AF[4][2m]); 0 if absent
};
// blob layout per ring: chirp[2m] | wk[2m] | ph[2m] | FHp[2L] | P0,Q0,P1,Q1[2m each] | A0..A3[2m each]  (interleaved re,im)
//   inverse prologue:  u[k] = conj(c_k D_k) P_h[k] + (c_{m-k} D_{m-k}) Q_h[k]          (sequence half h = 0,1)
//   forward epilogue:  X[k] = b0[k] A0[k] + conj(b0[m-k]) A1[k] + b1[k] A2[k] + conj(b1[m-k]) A3[k]
// min_fast_L: shortest convolution length served by the register-resident team kernels on this grid (the rest goes to the
// generic kernel, all classes in one launch): 32 on grids with levels (hundreds of sequences per ring), 256 on RL grids
// (ten sequences per ring and variable: launch-bound, fewer launches win)
void build_ring_plans(int min_fast_L, const std::vector<int>& ring_ri, std::vector<FftClass>& classes,
                      std::vector<RingPlan>& plans, std::vector<double>& blob);
// host reference of the device FFT (used to build FHp and by self-tests)
void host_fft_dif(double* x, int L, const double* tw);   // natural -> digit-reversed
void host_fft_dit(double* x, int L, const double* tw);   // digit-reversed -> natural, unnormalised inverse

// ---------------------------------------------------------------- device descriptors
struct DevGrid {
  int has_l, has_z;
  int V, D;
  int num_cells, rDim, b_rDim, zDim, bz, bzp, kDim, ncolp;
  int patchOffsetL;          // radial mish-point offset of this tile inside the patch
  int coefOffset;            // spectralIndexL-1: coefficient offset inside the patch
  long long hpoints, N, S, W;
  const int* ring_n;           // [rDim]
  const int* ring_ri;          // [rDim]
  const long long* ring_hoff;  // [rDim+1] horizontal point prefix
  const long long* ring_hoffp; // [rDim+1] the same prefix with every ring padded to a multiple of 16 points (blocked SZ layout)
  long long hpointsp;          // ring_hoffp[rDim]
  const long long* ring_woff;  // [rDim+1] retained-coefficient prefix (1+2ri per ring, or 1)
  const double* rad;           // [rDim]
  const double* zlev;          // [zDim]
  const int* h2r;              // [hpoints] ring of each horizontal point
  double phi[3][3][4];
  double wq[3];
};

// number of SMs of the current device (148 on B200), queried once per device; grids of the persistent kernels are sized from it
int sb_sm_count();

// a <= 32-column tile of one ring for the Chebyshev kernels.  ring / rad / blk ride along so that the fused kernel needs no
// dependent lookups (h2r -> rad, h2r -> ring offsets) behind the descriptor load: blk = bz * (hoffp[ring] + j0), the tile's
// block in a field row of the blocked SZ layout
struct ZTile { int hcol0; int ncols; long long out_base; int out_stride; int ring; double rad; long long blk; };

// Where the inverse ring transform puts the rows of one (field, variable, z-mode) of a ring.  out_is_phys selects
//   1: the physical array [D][V][N] (grids without levels);
//   0: the post-L scratch SZ, one contiguous row of n points per z-mode:  bz * hoff[r] + zb * n + j;
//   2: SZ in the BLOCKED layout the fused synthesis kernel (k_inv_z_advection_bulk) fetches with two bulk copies per field
//      and tile instead of one per z-mode: the ring is cut into blocks of 16 points (rings padded to a multiple of 16),
//      a block holds all bz modes of its 16 points contiguously, even modes first (the parity split of the synthesis),
//      and inside a 16-point row the point index is XORed with 4 (k & 3), k = zb >> 1 -- the DMMA fragment loads
//      (4 modes x 8 points per warp) then hit 16 different banks without padding the rows:
//        bz * (hoffp[r] + 16 (j >> 4)) + 16 ((zb & 1) * ceil(bz / 2) + k) + ((j & 15) ^ 4 (k & 3))
struct RowDst {
  double* base;
  long long blk;
  int swz;
  __device__ __forceinline__ double* at(int j) const {      // j even: a double2 (points j, j + 1) is stored here
    return blk ? base + (long long)(j >> 4) * blk + ((j & 15) ^ swz) : base + j;
  }
  // at(j + 4 T n) == at(j) + n * step(T) for T % 4 == 0: the team kernels store their eight outputs a = n T + tl with one
  // address computation and a constant stride in either layout
  __device__ __forceinline__ long long step(int T) const { return blk ? (long long)(T >> 2) * blk : 4LL * T; }
};
// hoff: the ring's point offset -- g.ring_hoff[r], or g.ring_hoffp[r] (padded rings) for the blocked layout
__device__ __forceinline__ RowDst row_dst(const DevGrid& g, double* out, long long out_fs, long long out_vs, int out_is_phys,
                                          int f, int v, int var0, long long hoff, int n, int zb) {
  RowDst d;
  d.blk = 0; d.swz = 0;
  if (out_is_phys == 1) {
    d.base = out + ((long long)f * g.V + var0 + v) * g.N + hoff;
  } else if (out_is_phys == 0) {
    d.base = out + (long long)f * out_fs + (long long)v * out_vs + (long long)g.bz * hoff + (long long)zb * n;
  } else {
    const int k = zb >> 1;
    d.base = out + (long long)f * out_fs + (long long)v * out_vs + (long long)g.bz * hoff + 16 * ((zb & 1) * ((g.bz + 1) >> 1) + k);
    d.blk = 16LL * g.bz;
    d.swz = (k & 3) << 2;
  }
  return d;
}
struct LWork { int r; int row0; int nrows; int pad; };

// ---------------------------------------------------------------- kernel launchers (sb_transforms.cu)
// per-launch CUDA-event timing on the launching stream (bench.py's live roofline numbers)
struct Profiler {
  bool on = false;
  struct Rec { const char* name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
};
// What a K3 pass has to produce (tiles_physics: only what the equation-set kernel reads; everywhere else: all).
//   smask: radial spectra  A | dA/dr | d2A/dr2                      (inv_r outputs, bit d)
//   lmask: ring rows       value | r | rr | lambda | lambda-lambda  (inv_l outputs, bit f)
//   zmask: fields the Chebyshev synthesis reads (bit f);  zsel: of field 0, bit 0 = value, 1 = d/dz, 2 = d2/dz2
struct K3Need { unsigned smask = 7, lmask = 31, zmask = 31, zsel = 7; };
// all generic-kernel (L < 256) ring classes of a grid in ONE launch: merged work lists + per-class parameters
struct SmallCls { int log2L; int pad; const double2* tw; };
struct SmallRings {
  const LWork* iwork = nullptr; int niwork = 0;      // inverse items of every small class, largest class first
  const LWork* fwork = nullptr; int nfwork = 0;
  const SmallCls* cls = nullptr;                     // device, indexed by RingPlan::cls
  size_t smem = 0;                                   // largest per-class requirement
};
struct LaunchCtx {
  cudaStream_t stream; long long* launches; Profiler* prof; K3Need need;
  // overlapped step only: restrict ring work lists / z tiles to the rings [r_lo, r_hi) (r_hi < 0: all rings), dynamic work
  // counters for the ring FFT launches, SMs the FFT grids leave to the other stream, grid limit of the HBM-bound kernels
  int r_lo = 0, r_hi = -1;
  int* counters = nullptr; int* counter_next = nullptr; int ncounters = 0;
  int sm_reserve = 0;
  int fft_chunk = 2;               // items per dynamic work share
  int mem_grid_sms = 0;            // > 0: persistent HBM-bound kernels size their grid for this many SMs
  const SmallRings* small = nullptr;
};
struct ProfScope {
  cudaStream_t s;
  cudaEvent_t b = nullptr;
  ProfScope(const LaunchCtx& c, const char* name) : s(c.stream) {
    if (c.prof && c.prof->on) {
      cudaEvent_t a;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, s);
      c.prof->recs.push_back(Profiler::Rec{name, a, b});
    }
  }
  ~ProfScope() { if (b) cudaEventRecord(b, s); }
};

void launch_fwd_z(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars,
                  const double* in, long long in_vstride, double* mirror, long long mirror_vstride,
                  double* out, long long out_vstride, const double* fwdT);
void launch_inv_z(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                  int nfields, const double* in, long long in_fstride, long long in_vstride,
                  double* phys, const double* invM /* [V][3][bz][zDim] */);
bool inv_z_mma_ok(const DevGrid& g, int nfields);
void build_inv_z_mma_tables(int zDim, int bz, const double* T0, const double* T1, const double* T2, std::vector<double>& out);
void launch_inv_z_mma(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                      int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                      const double* parB);
bool fwd_z_mma_ok(const DevGrid& g);
void build_fwd_z_mma_tables(int zDim, int bz, const double* fwd, std::vector<double>& out);
void launch_fwd_z_mma(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, const double* in,
                      long long in_vstride, double* mirror, long long mirror_vstride, double* out, long long out_vstride,
                      const double* fwdB);
bool inv_z_par_ok(const DevGrid& g, int nfields);
void build_inv_z_par_tables(int zDim, int bz, const double* T0, const double* T1, const double* T2, std::vector<double>& out);
void launch_inv_z_par(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                      int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                      const double* parM);
void launch_fwd_l(const LaunchCtx& c, const DevGrid& g, const std::vector<std::vector<LWork>>& hostwork,
                  const LWork* const* work, const std::vector<FftClass>& classes,
                  const double* const* tw, const double* const* twp, const RingPlan* plans, const double* blob, int nvars,
                  const double* in, long long in_vstride, int in_is_z, double* mirror, long long mirror_vstride,
                  double* out, long long out_vstride,
                  const std::vector<std::vector<LWork>>* hostwork2 = nullptr, const LWork* const* work2 = nullptr,
                  double* fft3_scratch = nullptr);
void launch_inv_l(const LaunchCtx& c, const DevGrid& g, const std::vector<std::vector<LWork>>& hostwork,
                  const LWork* const* work, const std::vector<FftClass>& classes,
                  const double* const* tw, const double* const* twp, const RingPlan* plans, const double* blob, int nvars,
                  const double* in, long long in_fstride, long long in_vstride,
                  double* out, long long out_fstride, long long out_vstride, int out_is_phys, int var0,
                  const std::vector<std::vector<LWork>>* hostwork2 = nullptr, const LWork* const* work2 = nullptr);
// v2 persistent ring FFT (sb_ringfft2.cu)
bool fft3_enabled();                          // composite lengths L = 3 * 2^a available (A/B switch SB_FFT3=0)
void fft3_class_tables(int L2, std::vector<double>& tab);   // appended to the class twiddles: TWB1[T] TWB2[T] C1[16] C2[16]
size_t fft3_scratch_doubles(int L);           // forward kernel's parking area (per launch)
void launch_inv_l3(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0);
void launch_fwd_l3(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs, double* scratch);
// v4 ring FFT (sb_ringfft4.cu): tables in Tensor Memory, input rows staged by the bulk-copy engine
bool fft4_supported(int L, bool forward);
void launch_inv_l4(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0);
void launch_fwd_l4(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs);
bool fft5_supported(int L);                   // composite lengths 3072 / 6144 with the v4 data movement
void launch_inv_l5(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0);
void launch_fwd_l5(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs, double* scratch);
bool fft2_supported(int L, bool forward);
int fft2_rows_per_item(int L, bool forward);
void launch_inv_l2(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0);
void launch_fwd_l2(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs);
// peer scatter of the forward radial transform (multi-GPU, plane-distributed solve): besides the tile's own B,
// every coefficient is stored straight into the buffer of the rank that owns its z-mode plane -- a plain pointer
// into that GPU's memory, mapped with CUDA IPC, so the transfer rides NVLink from inside the kernel.
#define SB_MAX_PEERS 8
struct PeerScatter {
  int nranks;
  int z0[SB_MAX_PEERS + 1];          // plane ranges
  double* base[SB_MAX_PEERS];        // owner k's receive slot of THIS tile (variable 0, its first plane)
  long long vstride[SB_MAX_PEERS];   // per-variable stride inside owner k's slot = nz_k * ncolp_t * b_rDim_t
};
void launch_fwd_r(const LaunchCtx& c, const DevGrid& g, int nvars, const double* in, long long in_vstride,
                  double* B, long long B_vstride, const PeerScatter* scatter = nullptr, int var0 = 0);
void launch_inv_r(const LaunchCtx& c, const DevGrid& tile, const DevGrid& patch, int nvars,
                  const double* A, long long A_vstride, double* out, long long out_fstride,
                  long long out_vstride, int out_is_phys, int var0);
void launch_fwd_l_fast(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                       const int twoff[4], const RingPlan* plans, const double* blob, int nvars, const double* in,
                       long long in_vs, double* mirror, long long mirror_vs, double* out, long long out_vs);
void launch_inv_l_fast(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                       const int twoff[4], const RingPlan* plans, const double* blob, int nvars, const double* in,
                       long long in_fs, long long in_vs, double* out, long long out_fs, long long out_vs, int out_is_phys,
                       int var0);
struct DevSplineFactor {
  int M, rL, rR, nfree, periodic;
  double foldL[2], foldR[2];
  const double* chol;   // [nfree][4]
  const double* dense;  // [M][M]
};
void launch_spline_solve(const LaunchCtx& c, const DevGrid& g, const DevSplineFactor* factors /*device [V]*/,
                         const std::vector<DevSplineFactor>& hfactors, const double* B, double* A);
void launch_assemble(const LaunchCtx& c, const DevGrid& patch, const DevGrid& tile, const double* tileB,
                     const DevGrid* prev, const double* prevB, int last, double* shared);
void launch_extract(const LaunchCtx& c, const DevGrid& patch, const DevGrid& tile, const double* A, double* tileA,
                    long long dst_vstride = 0);
void launch_column_op(const LaunchCtx& c, const double* M /*[rows][cols]*/, int rows, int cols, const double* in, double* out,
                      long long ncols, double C0);
void launch_copy(const LaunchCtx& c, double* dst, const double* src, long long n);
void launch_nan_scan(const LaunchCtx& c, const double* phys, long long N, int V, long long* result);

// ---------------------------------------------------------------- equation sets (sb_model.cu)
enum EquationSet {
  EQ_LinearAdvection1D = 0, EQ_LinearAdvectionRZ, EQ_LinearAdvectionRL, EQ_LinearAdvectionRLZ,
  EQ_LinearShallowWater1D, EQ_LinearShallowWaterRL, EQ_Oneway_ShallowWater_Slab,
  EQ_Twoway_ShallowWater_Slab, EQ_Oneway_ShallowWater_HeightResolvedBL, EQ_Euler_test,
  EQ_BF02_test, EQ_rainfall_test, EQ_COUNT
};
int equation_set_from_name(const char* name);
struct EqParams {
  double ts;
  double c_0, K, g, Cd, Hfree, Hb, f, S1, H, Kh, Um, Vm;
  double Pxi_bar;
  int iw, ixi, ih;     // column indices of "w", "xi", "h" (0-based, -1 when absent)
};
struct ModelArrays {
  double* phys;        // [D][V][N]
  double* var_np1;     // [V][N]
  double* exp_n;       // [V][N]  (pointers rotate each step)
  double* exp_nm1;
  double* exp_nm2;
  double* imp_n;       // semi-implicit only
  double* imp_nm1;
  double* imp_nm2;
  unsigned passive;       // bit v: variable v has no tendency in this equation set AND its expdot history is known to be all
                          // zeros (never written since allocation): a kernel may skip reading exp_nm1/exp_nm2[v] and writing
                          // exp_n[v] (= 0); var_np1[v] comes out of the same ab_step arithmetic fed with zeros.  0 = general path
  const double* colops;   // [4][zDim][zDim]: CB->CA->{CI, CIx, CIInt} of "h", (spare)
  const double* colfrag;  // DMMA B fragments of colops 2 (CIInt) and 1 (CIx): [2][zDim/8][zDim/4][32]; null if zDim % 8
  const double* refstate; // [4 profiles][3][zDim] sbar, xibar, mubar, mu_lbar (value, dz, dzz)
  const double* helm;     // [2][zDim][zDim] inverse Helmholtz matrices (tau=0.5 ts, 1.25 ts)
  const double* sicols;   // [7][zDim][zDim] composite column operators: 0..5 semi-implicit (F, Dz of xi; W, X for tau = 0.5 ts,
                          // 1.25 ts), 6 = CB->CA->CIx of "mu_r" (rainfall_test's sedimentation flux divergence)
};
void equation_set_needs(int eq, const EqParams& p, const DevGrid& g, unsigned* need /*[V]: slots the kernel reads*/);
unsigned equation_set_passive(int eq, int V);   // bit v: the equation set never writes expdot[:, v] (src/testModels.jl:40,68,93: only h)
void build_colop_fragments(int nz, const double* Mt /*[k][z]*/, std::vector<double>& out);
// K3 last stage + K4 fused for LinearAdvectionRLZ (sb_chebmma.cu): in = [7 field rows: h value,r,rr,l,ll | u | v][bz][ring rows]
bool inv_z_advection_ok(const DevGrid& g);
bool inv_z_advection_blocked(const DevGrid& g);   // the kernel that reads the blocked SZ layout (RowDst) will run for this grid
void launch_inv_z_advection(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, const double* in,
                            long long in_fstride, const double* parB, const EqParams& p, const ModelArrays& arr, int t,
                            bool blocked = false);
void launch_equation_set(const LaunchCtx& c, int eq, const DevGrid& g, const EqParams& p,
                         const ModelArrays& a, int tstep);

}  // namespace sb
