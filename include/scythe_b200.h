/* scythe_b200.h -- C ABI of libscythe_b200.so, the B200 (sm_100a) implementation of the
 * Scythe.jl semi-spectral transform + time-step hot path.
 *
 * This is the drop-in boundary a Julia shim binds with `ccall` and the
 * Python host mirror binds with ctypes (scythe_jl_b200/_lib.py).  Plain pointers and sizes only.
 * Every entry point cites the reference interface it replaces; paths are relative to the
 * Scythe.jl tree (/root/reference).  Springsteel.jl (where the reference's transforms live) is
 * an un-vendored dependency (Project.toml:20); those citations are the Scythe call sites.
 *
 * Conventions
 *   - all arrays Float64, column-major, 0-based in C / 1-based in the Julia shim;
 *     physical[N,V,D], spectral[S,V]  (SURVEY App. A.2 C1/C2, A.3 layout);
 *   - host buffers are caller-owned and never retained after the call returns;
 *   - every call returns 0 on success, a negative SB_E* code otherwise; sb_last_error() gives
 *     the message (thread-local).  CUDA errors are sticky for the handle;
 *   - one handle = one device + one stream; calls on a handle are not re-entrant;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with SB_ECUDA.
 */
#ifndef SCYTHE_B200_H
#define SCYTHE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_EINVAL (-1)   /* bad argument            -> Julia ArgumentError              */
#define SB_EDOMAIN (-2)  /* DomainError (unknown geometry, too many tiles, ...)          */
#define SB_ECUDA (-3)    /* CUDA runtime / launch failure                                */
#define SB_ENAN (-4)     /* checkCFL: NaN found                                          */
#define SB_EUNSUPPORTED (-5) /* equation set / BC not built as a CUDA kernel (no fallback) */
#define SB_ECOMM (-6)    /* NCCL failure                                                 */

/* geometry codes: createGrid dispatch, src/spectralGrid.jl:63-94 */
enum { SB_GEOM_R = 0, SB_GEOM_RZ = 1, SB_GEOM_RL = 2, SB_GEOM_RLZ = 3 };

/* radial (cubic B-spline) boundary conditions: CubicBSpline.R0 ... PERIODIC,
 * models/cha_bell2024/Oneway_ShallowWater_Slab.jl:13-26, models/LinearAdvection1D.jl:13-16 */
enum { SB_BC_R0 = 0, SB_BC_R1T0 = 1, SB_BC_R1T1 = 2, SB_BC_R1T2 = 3, SB_BC_R2T10 = 4,
       SB_BC_R2T20 = 5, SB_BC_R3 = 6, SB_BC_PERIODIC = 7 };
/* vertical (Chebyshev) boundary conditions: Chebyshev.R0 ..., src/reference_state.jl:102-103 */
enum { SB_ZBC_R0 = 0, SB_ZBC_R1T0 = 1, SB_ZBC_R1T1 = 2, SB_ZBC_R1T2 = 3 };

/* GridParameters, src/spectralGrid.jl:20-45 (derived fields are computed by the library) */
typedef struct sb_grid_params {
  int32_t geometry;        /* SB_GEOM_*                                         */
  int32_t nvars;           /* length(vars); variable v is column v (0-based)    */
  double xmin, xmax;       /* radial extent                                     */
  int64_t num_cells;       /* rDim = 3*num_cells, b_rDim = num_cells+3          */
  double l_q;              /* spline filter cutoff in cells (default 2.0)       */
  double zmin, zmax;       /* vertical extent (RZ/RLZ)                          */
  int64_t zDim;            /* vertical levels (RZ/RLZ)                          */
  int64_t b_zDim;          /* <=0: default min(zDim, floor((2zDim-1)/3)+1)      */
  int64_t spectralIndexL;  /* 1-based first patch coefficient of this tile      */
  int64_t tile_num;
  const int32_t* BCL;      /* [nvars] SB_BC_*  left/inner radial BC per variable */
  const int32_t* BCR;      /* [nvars] right/outer                               */
  const int32_t* BCB;      /* [nvars] SB_ZBC_* bottom (may be NULL = R0)        */
  const int32_t* BCT;      /* [nvars] top     (may be NULL = R0)                */
} sb_grid_params;

/* dimensions of a created grid (what Julia reads off grid.params / size(grid.physical)) */
typedef struct sb_grid_info {
  int64_t N;          /* grid points = size(physical,1)                 */
  int64_t V;          /* variables                                     */
  int64_t D;          /* derivative slots: R 3, RL 5, RZ 5, RLZ 7      */
  int64_t S;          /* spectral rows = size(spectral,1)              */
  int64_t rDim, b_rDim, zDim, b_zDim, kDim, lDim;
  int64_t num_columns; /* num_columns(grid): 0 for R/RL               */
  int64_t patchOffsetL;
  int64_t ndims;       /* columns of getGridpoints: 1,2,2,3             */
} sb_grid_info;

typedef struct sb_grid* sb_grid_t;
typedef struct sb_model* sb_model_t;

const char* sb_last_error(void);
/* library/ABI version and the CUDA arch it was built for ("sm_100a"); never fails */
const char* sb_version(void);
/* number of CUDA devices visible (0 when none / no driver) */
int sb_device_count(void);

/* ---- grids (Springsteel surface used by Scythe) --------------------------- */
/* createGrid(gp): src/semiimplicit.jl:130,150,155.  stream: a cudaStream_t or NULL (default). */
int sb_grid_create(const sb_grid_params* gp, int device, void* stream, sb_grid_t* out);
int sb_grid_destroy(sb_grid_t g);
int sb_grid_get_info(sb_grid_t g, sb_grid_info* out);
/* getGridpoints(grid): src/semiimplicit.jl:59.  out is [N, ndims] column-major. */
int sb_grid_get_gridpoints(sb_grid_t g, double* out, int64_t n_doubles);
/* grid.physical[:, :, slot0:slot0+nslots-1] <-> host [N,V,nslots] (read_physical_grid fills slot 0:
 * src/semiimplicit.jl:134) */
int sb_grid_set_physical(sb_grid_t g, const double* host, int32_t slot0, int32_t nslots);
int sb_grid_get_physical(sb_grid_t g, double* host, int32_t slot0, int32_t nslots);
/* grid.spectral (which=0: B, the spline inner products) or the A coefficients (which=1) <-> host [S,V] */
int sb_grid_set_spectral(sb_grid_t g, int32_t which, const double* host);
int sb_grid_get_spectral(sb_grid_t g, int32_t which, double* host);
/* spectralTransform!(grid): physical[:,:,1] -> spectral (B).  src/semiimplicit.jl:135,734 */
int sb_spectral_transform(sb_grid_t g);
/* gridTransform!(grid): spectral (B) -> A -> physical[:,:,1:D].  src/semiimplicit.jl:136 */
int sb_grid_transform(sb_grid_t g);
/* splineTransform!(patchSplines, patchSpectral, gp, sharedSpectral, tile): B -> A over the whole
 * patch.  `shared` is the grid whose B buffer is the input (may be `patch` itself); the A
 * coefficients land in patch's A buffer.  src/semiimplicit.jl:237,285 */
int sb_spline_transform(sb_grid_t patch, sb_grid_t shared);
/* tileTransform!(patchSplines, patchSpectral, gp, tile, splineBuffer): patch A -> tile.physical at
 * the tile's own points (tile may be the patch itself).  src/semiimplicit.jl:241,252,290,305 */
int sb_tile_transform(sb_grid_t patch, sb_grid_t tile);
/* calcTileSizes(patch, n): 5 x ntiles column-major [xmin;xmax;num_cells;spectralIndexL;npts].
 * src/semiimplicit.jl:141-144 */
int sb_calc_tile_sizes(const sb_grid_params* patch, int32_t ntiles, double* out5xN);
/* sharedSpectral .= 0  (src/semiimplicit.jl:272) on patch's B buffer */
int sb_shared_clear(sb_grid_t patch);
/* sharedSpectral[patchIndexMap] .= tileView; sharedSpectral[haloReceiveIndexMap] .+= halo(prev)
 * (src/semiimplicit.jl:320-329; calcPatchMap/calcHaloMap :79-86).  prev may be NULL.  When
 * last!=0 the tile's own halo is also added (the master's add at :279-282). */
int sb_shared_assemble(sb_grid_t patch, sb_grid_t tile, sb_grid_t prev, int32_t last);
/* checkCFL(grid): src/semiimplicit.jl:737-751.  Returns SB_ENAN and fills var/index (0-based). */
int sb_check_cfl(sb_grid_t g, int32_t* var, int64_t* index);
int sb_grid_sync(sb_grid_t g);
/* raw device pointers for zero-copy interop (torch.distributed / CUDA.jl unsafe_wrap):
 * which = 0 physical, 1 spectral B, 2 spectral A */
int sb_grid_device_ptr(sb_grid_t g, int32_t which, void** ptr, int64_t* n_doubles);

/* ---- model (Scythe driver; src/Scythe.jl:8-62, src/semiimplicit.jl:18-355) ----------------- */
/* ModelParameters, src/Scythe.jl:8-21.  physical_params is passed as parallel name/value arrays
 * (the reference's Dict{Symbol,Float64}); options mirror options[:semiimplicit]. */
typedef struct sb_model_params {
  double ts;
  double integration_time;
  double output_interval;
  const char* equation_set;       /* name resolved to a built-in CUDA kernel; src/semiimplicit.jl:357-363 */
  const sb_grid_params* grid;     /* patch grid_params                                                   */
  const char* const* var_names;   /* [nvars]: keys of grid_params.vars ordered by column                 */
  int32_t n_physical_params;
  const char* const* param_names; /* e.g. "g","K","Cd","Hfree","Hb","f","S1","c_0","Kh","Um","Vm","H"     */
  const double* param_values;
  int32_t semiimplicit;           /* options[:semiimplicit]                                              */
  /* reference state already evaluated on the model levels (ReferenceState, src/reference_state.jl:4-10):
   * each [zDim,3] column-major (value, d/dz, d2/dz2); NULL when ref_state_file is empty */
  const double* ref_sbar;
  const double* ref_xibar;
  const double* ref_mubar;
  double Pxi_bar;
  const double* ref_mu_lbar;      /* liquid-water reference profile of BF02_test (src/testModels.jl:291-293); NULL = zeros */
} sb_model_params;

/* initialize_model(model, workerids) (src/semiimplicit.jl:126-193) for the tiles this process owns:
 * tiles tile_first .. tile_first+tile_count-1 of ntiles (one tile per reference worker).
 * ic_host: patch physical[:, :, 1] as [N_patch, V] column-major (read_physical_grid, :134). */
int sb_model_create(const sb_model_params* mp, int32_t ntiles, int32_t tile_first, int32_t tile_count,
                    int device, void* stream, sb_model_t* out);
int sb_model_destroy(sb_model_t m);
int sb_model_initialize(sb_model_t m, const double* ic_host);
/* handles of the patch grid / local tile i (borrowed; owned by the model) */
int sb_model_patch(sb_model_t m, sb_grid_t* out);
int sb_model_tile(sb_model_t m, int32_t i, sb_grid_t* out);
/* advanceTimestep(mtile, sharedSpectral, haloSend, haloReceive, t) for every local tile
 * (src/semiimplicit.jl:301-332): K3 tileTransform!, equation set + explicit_timestep
 * [+ semiimplicit_adjustment], calcTendency (K1), own block + halo into the shared B buffer.
 * t is the 1-based step index (selects Euler / AB2 / AB3, :682-696). */
int sb_model_advance_tiles(sb_model_t m, int64_t t);
/* the per-step exchange when tiles live in several processes: sum of the shared B buffer over
 * ranks (NCCL all-reduce; replaces SharedArray + RemoteChannel halos, :204-229,272-282) */
int sb_model_exchange(sb_model_t m);
/* splineTransform! on every worker (src/semiimplicit.jl:285): shared B -> patch A */
int sb_model_spline_transform(sb_model_t m);
/* one full iteration of model_loop (src/semiimplicit.jl:268-297) without output */
int sb_model_step(sb_model_t m, int64_t t);
/* `nsteps` iterations starting at step index t0 (1-based) */
int sb_model_run(sb_model_t m, int64_t t0, int64_t nsteps);
/* output path (src/semiimplicit.jl:289-291): tileTransform!(patch) + checkCFL; copies
 * patch.physical [N_patch,V,D] to host if host != NULL */
int sb_model_output(sb_model_t m, double* host_physical);
/* ModelTile state arrays of local tile i <-> host [N_tile,V]:
 * which = 0 var_np1, 1 expdot_n, 2 expdot_nm1, 3 expdot_nm2, 4 impdot_n, 5 impdot_nm1, 6 impdot_nm2 */
int sb_model_get_state(sb_model_t m, int32_t tile, int32_t which, double* host);
/* state of local tile i <- host [N_tile,V]: which = 0 var_np1, 2/3 expdot_nm1/nm2, 5/6 impdot_nm1/nm2 (the
 * numbering of sb_model_get_state; 1 and 4 are transient).  Restart / host-driven stepping: the reference
 * restarts from an output file (physical_out_*.csv as the next initial_conditions,
 * notebooks/Cha_Bell_WCD2024_initialization.ipynb:196) and re-enters the Euler/AB2 start-up because the
 * AB3 history of src/semiimplicit.jl:685-695 is not saved; setting the history too makes a restart exact. */
int sb_model_set_state(sb_model_t m, int32_t tile, int32_t which, const double* host);
/* calcTendency for every local tile (src/semiimplicit.jl:728-735: physical[:,:,1] <- var_np1,
 * spectralTransform!) followed by the own-block / halo assembly into the shared B buffer (:320-329).
 * With sb_model_set_state + sb_model_exchange + sb_model_spline_transform this is the tile-parallel
 * form of initialize_model's spectralTransform!(patch) (:135): the forward transform is additive
 * across tiles, so no rank ever needs the whole patch in physical space. */
int sb_model_tendency(sb_model_t m);
/* one full iteration of model_loop entered at calcTendency (src/semiimplicit.jl:317-329, 279-285,
 * then :305-314 of the next iteration): var_np1 -> K1 -> shared sum -> K2 -> K3 -> equation set ->
 * var_np1.  Same work as sb_model_step; used for host-buffer (end-to-end) stepping. */
int sb_model_cycle(sb_model_t m, int64_t t);
/* Host-driven stepping without stalls: asynchronous, pipelined forms of sb_model_set_state(which = 0) and
 * sb_model_get_state(which = 0) for a caller whose state lives in HOST memory between steps (the reference's workers
 * own `var_np1` as Julia host arrays, src/semiimplicit.jl:18-42).  stage_in: host [N_tile,V] -> a device staging buffer
 * on a copy stream, then into var_np1 on the compute stream; stage_out: var_np1 -> a staging buffer on the compute
 * stream, then to host [N_tile,V] on a second copy stream.  Two staging buffers each way, ordered by events only, so the
 * H2D copy of step i+1, the kernels of step i and the D2H copy of step i-1 overlap (PCIe is full duplex).  The calls
 * return at once: `host` must be page-locked for the copies to be asynchronous and must stay valid (stage_in:
 * unmodified; stage_out: unread) until sb_model_stage_drain(m, 1) or sb_model_sync(m) returns.
 * sb_model_stage_drain makes the compute stream wait for every staged copy issued so far (block = 0: stream-ordered
 * only, so that a following sb_timer_stop covers them; block = 1: also waits on the host). */
int sb_model_stage_in(sb_model_t m, int32_t tile, const double* host);
int sb_model_stage_out(sb_model_t m, int32_t tile, double* host);
int sb_model_stage_drain(sb_model_t m, int32_t block);
/* the first half of advanceTimestep alone (src/semiimplicit.jl:305-314): tileTransform! + equation set + time step.
 * sb_model_tendency + caller-driven exchange + sb_model_physics = sb_model_cycle when the exchange is the caller's. */
int sb_model_physics(sb_model_t m, int64_t t);
/* Which physical slots the tileTransform! inside a step (src/semiimplicit.jl:305) produces.  The derivative slots of a
 * tile are intermediates between tileTransform! and the equation set: nothing else reads them (calcTendency
 * overwrites them right after, src/semiimplicit.jl:731).  mode 0 (default): the (variable, slot) pairs the built-in
 * equation-set kernel reads, and where a fused kernel exists (LinearAdvectionRLZ) the equation set + time step run in
 * the epilogue of the last transform stage, so no slot is written at all; 1: all D slots of every variable, the
 * reference's dataflow (env SB_K3_FULL=1 makes it the default); 2: the needed slots with every other slot set to NaN
 * first (test hook); 3: the needed slots, no fusion.  var_np1 is bit-identical in all modes; sb_tile_transform /
 * sb_grid_transform / sb_model_output always produce every slot. */
int sb_model_set_k3_slots(sb_model_t m, int32_t mode);
/* per-kernel CUDA-event timing on the model's stream: enable/disable, then read
 * "name launches total_ms\n" lines (clears the records). */
int sb_model_profile(sb_model_t m, int32_t on);
int sb_model_profile_report(sb_model_t m, char* buf, int64_t buflen);
int sb_model_sync(sb_model_t m);
/* number of kernels this model launched since creation (bench.py's gpu_launches) */
int64_t sb_model_launch_count(sb_model_t m);

/* ---- Chebyshev column API (Springsteel's Chebyshev module as Scythe calls it) ------------------
 * Chebyshev1D(ChebyshevParameters(zmin, zmax, zDim, bDim, BCB, BCT)) with CBtransform!, CAtransform!,
 * CItransform!, CIxtransform, CIxxtransform, CIInttransform(col, C0) -- src/semiimplicit.jl:569-574,593,596,
 * src/shallowWaterModels.jl:424-429,480-482, src/reference_state.jl:97-108,141-155 -- and
 * Chebyshev.dct_matrix / dct_1st_derivative / dct_2nd_derivative (src/semiimplicit.jl:772-775).
 * The column transforms are batched on the device: `in` / `out` are host arrays [len x ncols], column-major
 * (one column of the reference's col.uMish / col.b / col.a per batch entry).  Lengths: CB zDim -> b_zDim,
 * CA b_zDim -> zDim (zero filled), CI / CIx / CIxx / CIInt zDim -> zDim. */
typedef struct sb_cheb_params {
  double zmin, zmax;
  int64_t zDim;           /* number of Gauss-Lobatto levels */
  int64_t b_zDim;         /* retained modes; <= 0 -> min(zDim, (2 zDim - 1)/3 + 1) */
  int32_t BCB, BCT;       /* SB_ZBC_* */
} sb_cheb_params;
enum { SB_CHEB_CB = 0, SB_CHEB_CA = 1, SB_CHEB_CI = 2, SB_CHEB_CIX = 3, SB_CHEB_CIXX = 4, SB_CHEB_CIINT = 5 };
int sb_cheb_mish_points(const sb_cheb_params* cp, double* z /* [zDim] */);
/* zDim x zDim matrices, column-major (Julia): values / d/dz / d2/dz2 at the mish points of coefficient k */
int sb_cheb_matrices(const sb_cheb_params* cp, double* dct, double* dct1, double* dct2);
int sb_cheb_columns(const sb_cheb_params* cp, int32_t op, const double* in, double* out, int64_t ncols, double C0, int device);

/* ---- multi-GPU: NCCL over NVLink, one process per GPU ------------------------------------- */
/* ncclGetUniqueId -> 128 bytes to ship to the other ranks (the reference ships RemoteChannels the
 * same way, src/semiimplicit.jl:205-219) */
int sb_comm_unique_id(void* out128);
/* ---- plane-distributed splineTransform! (SURVEY 8e option B) --------------------------------
 * The reference replicates the whole-patch solve on every worker (src/semiimplicit.jl:285) after
 * summing the tiles' B through one SharedArray (:323-329).  Across GPUs that is an all-reduce of the
 * whole patch; instead the (z-mode, wavenumber) columns are dealt out by z-mode plane.  After
 * sb_model_colsolve_init, sb_model_advance_tiles / sb_model_tendency leave B tile-local, the caller
 * (or sb_model_exchange with the library's own communicator) moves the message buffers named by
 * sb_model_colsolve_buffer, sb_model_colsolve_solve overlap-adds + solves + packs on the owner, and the
 * tiles evaluate their own slice of A.  what: 0 send-B (local tile -> owner peer), 1 recv-B (tile ->
 * me), 2 send-A (me -> tile), 3 recv-A (owner peer -> local tile), 4 my solved planes, 5 owner
 * peer's planes inside the replicated patch A (only needed for output).  Buffers are per variable v. */
int sb_model_colsolve_init(sb_model_t m, int32_t rank, int32_t nranks);
int sb_model_colsolve_planes(sb_model_t m, int32_t* z0 /* [nranks+1] */, int32_t n);
int sb_model_colsolve_buffer(sb_model_t m, int32_t what, int32_t tile, int32_t v, int32_t peer, void** ptr, int64_t* count);
int sb_model_colsolve_solve(sb_model_t m);
int sb_model_colsolve_publish(sb_model_t m);
/* Peer-memory form of the same exchange (one process per GPU on one NVSwitch box): instead of messages, the
 * forward radial kernel stores every B coefficient straight into the receive buffer of the rank that owns its
 * z-mode plane, and the owner's cut-out kernel stores every tile's slice of A straight into that tile's A -- plain
 * pointers into the other GPU's memory, mapped with CUDA IPC, so the transfer overlaps the kernels tile by tile over
 * NVLink.  Each rank exports handles (what = 0: its receive buffer; 1: spectral A of local tile `tile`), ships them
 * to the others (the reference ships RemoteChannels the same way, src/semiimplicit.jl:205-219), opens the others'
 * (index = peer rank for what 0, global tile index for what 1) and calls sb_model_p2p_enable.  A step then needs
 * only two stream-ordered rendezvous (sb_model_exchange with the library's communicator, or the caller's own). */
int sb_model_ipc_handle(sb_model_t m, int32_t what, int32_t tile, void* out64);
int sb_model_ipc_open(sb_model_t m, int32_t what, int32_t index, const void* handle64);
int sb_model_p2p_enable(sb_model_t m);
int sb_model_comm_init(sb_model_t m, const void* id128, int32_t rank, int32_t nranks);

/* ---- timing on the handle's stream (CUDA events) ------------------------------------------ */
int sb_timer_start(sb_grid_t g);
int sb_timer_stop(sb_grid_t g, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* SCYTHE_B200_H */
