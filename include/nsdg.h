/*
 * nsdg.h -- C ABI of libnsdg_cuda.so, the B200-native (sm_100a, FP64) implementation of
 * neXtSIM_DG's dynamics hot path (DG advection + subcycled CG momentum with mEVP / BBM).
 *
 * This is the drop-in boundary: the entry points are exactly what a CUDAMEVPDynamics /
 * CUDABBMDynamics IDynamics module needs from its kernel object.  Each one names the
 * reference interface it replaces (paths relative to the nextsimdg source tree).
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success, non-zero on
 *    error; nsdg_last_error() gives the message of the last failure on this thread
 *    (the C++ module wrapper rethrows it as std::runtime_error, like the reference's
 *    exceptions at IDynamics.hpp:116).
 *  - host arrays are the raw buffers of the model's ModelArrays: row-major N x ncomp
 *    (core/src/include/ModelArray.hpp:92), element index i + nx*j (x fastest).
 *  - a handle is not re-entrant; calls are synchronous (results are complete on return),
 *    matching the single model thread that calls IDynamics::update
 *    (core/src/PrognosticData.cpp:95).  Several handles may live in one process and be used
 *    alternately, but from one thread at a time: handles that run the generic kernel on a uniform
 *    mesh (either CG/DG build) share one per-device constant-memory operator set that each
 *    re-uploads when it is not its owner.
 *    Every entry point makes the handle's device current (cudaSetDevice) and leaves it so.
 *  - there is no CPU fallback: every call fails with an error if no CUDA device is usable.
 */
#ifndef NSDG_H
#define NSDG_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsdg_handle_s* nsdg_handle;

/* Rheology = which reference kernel is replaced. */
enum nsdg_rheology {
    NSDG_MEVP = 0, /* MEVPDynamicsKernel  (dynamics/src/include/MEVPDynamicsKernel.hpp:17-37) */
    NSDG_BBM = 1, /* BBMDynamicsKernel   (dynamics/src/include/BBMDynamicsKernel.hpp:18-44)  */
    NSDG_FREEDRIFT = 2 /* FreeDriftDynamicsKernel (dynamics/src/include/FreeDriftDynamicsKernel.hpp:19-96) */
};

/* Named fields of DynamicsKernel::setData / getDG0Data / getDGData
 * (dynamics/src/include/DynamicsKernel.hpp:92-156, core/src/include/gridNames.hpp:17-37). */
enum nsdg_field {
    NSDG_HICE = 0, /* "hice"      in/out, DG */
    NSDG_CICE = 1, /* "cice"      in/out, DG */
    NSDG_DAMAGE = 2, /* "damage"    in/out, DG (BBM only) */
    NSDG_U = 3, /* "u"         in (cell means -> CG via DG2CG), out (CG -> DG0) */
    NSDG_V = 4, /* "v" */
    NSDG_UWIND = 5, /* "uwind"     in */
    NSDG_VWIND = 6, /* "vwind"     in */
    NSDG_UOCEAN = 7, /* "uocean"    in */
    NSDG_VOCEAN = 8, /* "vocean"    in */
    NSDG_SSH = 9, /* "ssh"       in, DG0 */
    NSDG_TAUX = 10, /* "uiostress" out: ice-ocean stress, x */
    NSDG_TAUY = 11, /* "viostress" out: ice-ocean stress, y */
    NSDG_NFIELDS = 12
};

/* Neighbour sides of a partition box (partition.cdl connectivity: core/test/partition_metadata_3.cdl:28-66) */
enum nsdg_side { NSDG_BOTTOM = 0, NSDG_RIGHT = 1, NSDG_TOP = 2, NSDG_LEFT = 3 };

/*
 * Construction parameters.  Zero-initialise, then set what differs from the defaults
 * (nsdg_config_default does both).  dgadv / cgdegree are the reference's compile-time
 * DGCOMP and CGDEGREE (CMakeLists.txt:12-16,112-118); nsteps is DynamicsKernel::nSteps
 * (DynamicsKernel.hpp:187, hard-coded 100 upstream).
 */
typedef struct nsdg_config {
    int rheology; /* enum nsdg_rheology */
    int dgadv; /* DGCOMP: 6 (DG2, default), 3 (DG1) */
    int cgdegree; /* CGDEGREE: 2 (default), 1 */
    int nsteps; /* subcycles per update(); default 100 */
    int device; /* CUDA device ordinal; -1 = current device */
    int use_cuda_graph; /* 1 = capture the subcycle loop in a CUDA graph (default 1) */
    int force_general; /* 1 = never use the uniform-rectangular-mesh fast path (default 0) */
    int pin_host_buffers; /* 1 = nsdg_update page-locks the caller's arrays on first use (cudaHostRegister) so
                             the per-step copies are asynchronous DMA; the arrays must then stay allocated
                             until nsdg_destroy (true for the model's ModelArrays). default 0 */
    /* mEVP relaxation parameters (MEVPStressUpdateStep.hpp:126-127, VPCGDynamicsKernel.hpp:125-126) */
    double alpha, beta; /* default 1500, 1500 */
    /* Partition of the global grid handled by this handle (ModelMetadata.cpp:42-62, run/partition.cdl).
     * Single-domain use: leave all zero. */
    int global_nx, global_ny; /* extent of the whole domain in elements (0 = not partitioned) */
    int box_x0, box_y0; /* first owned element of this box in the global grid */
    int rank, nranks; /* rank of this box, number of boxes */
    int neighbour[4]; /* rank of the box across each side, -1 = domain edge */
    /* OPT-IN EXTENSION, not reference behaviour (default 0).  The reference hands hice / cice (/ damage) to the dynamics as
     * one-component HFields every step, so DGModelArray::ma2dg zeroes the higher DG moments the advection has just built
     * (quirk Q4: DGModelArray.hpp:20-32; the thermodynamics in between works on cell means, IceGrowth.cpp:45-46).  With
     * keep_dg_moments = 1, nsdg_update takes the caller's arrays as the new CELL MEANS only: component 0 is replaced, the
     * higher moments resident on the device are kept; hice and cice are limited again (LimitMax / LimitMin as after the
     * advection, DynamicsKernel.hpp:160-172) so that their bounds hold for the new mean.  nsdg_set_field is unaffected. */
    int keep_dg_moments;
} nsdg_config;

void nsdg_config_default(nsdg_config* cfg);

/* Replaces: construction of the kernel member in MEVPDynamics::MEVPDynamics / BBMDynamics::BBMDynamics
 * (core/src/modules/DynamicsModule/MEVPDynamics.cpp:29-35, BBMDynamics.cpp:24-30). */
int nsdg_create(const nsdg_config* cfg, nsdg_handle* out);

/* Replaces: Module::finalize<IDynamics> destroying the kernel (core/src/include/Module.hpp:171-175). */
int nsdg_destroy(nsdg_handle h);

/* Replaces: DynamicsKernel::initialise(coords, isSpherical, mask)
 * (DynamicsKernel.hpp:44-75, CGDynamicsKernel.cpp:24-51, BrittleCGDynamicsKernel.hpp:72-86).
 *   coords_xy : (nx+1)*(ny+1) interleaved (x,y) pairs, vertex index x-fastest; metres, or
 *               lon/lat in RADIANS when spherical != 0 (the module multiplies degrees by
 *               pi/180 first, MEVPDynamics.cpp:43-46)
 *   mask      : nx*ny doubles, 1.0 = ocean/ice element (ParametricMesh.cpp:221)
 * With a partition in the config, nx, ny, coords and mask describe the box INCLUDING its
 * one-element overlap ring where a neighbour exists.  Calling it again on a box whose halos are connected
 * disconnects them (the arena its neighbours mapped is reallocated): every box of the partition must then
 * repeat nsdg_halo_export / nsdg_halo_connect / nsdg_halo_ready. */
int nsdg_set_mesh(nsdg_handle h, int nx, int ny, const double* coords_xy, const double* mask, int spherical);

/* Replaces: DynamicsKernel::setData(name, ModelArray) (DynamicsKernel.hpp:92-112,
 * CGDynamicsKernel.cpp:53-90, BrittleCGDynamicsKernel.hpp:138-145) with DGModelArray::ma2dg
 * semantics (DGModelArray.hpp:20-32): ncomp == 1 sets component 0 and ZEROES the higher
 * moments; ncomp == dgadv copies all components. */
int nsdg_set_field(nsdg_handle h, int field, const double* host, int ncomp);

/* Replaces: kernel.update(tst) -- VPCGDynamicsKernel::update (VPCGDynamicsKernel.hpp:63-94) or
 * BrittleCGDynamicsKernel::update (BrittleCGDynamicsKernel.hpp:91-136): advection + limiters,
 * prepareIteration, nsteps subcycles.  dt_seconds = tst.step.seconds(). */
int nsdg_step(nsdg_handle h, double dt_seconds);

/* Replaces: getDG0Data(name) (ncomp == 1; DynamicsKernel.hpp:114-126, CGDynamicsKernel.cpp:92-118)
 * and getDGData(name) (ncomp == dgadv; DynamicsKernel.hpp:134-156). */
int nsdg_get_field(nsdg_handle h, int field, double* host, int ncomp);

/* Replaces: assigning ParametricMesh::dirichlet[4] / ParametricMesh::periodic (dynamics/src/include/ParametricMesh.hpp:76-79),
 * which the reference fills from a version-2.0 .smesh file (dynamics/src/ParametricMesh.cpp:79-178) or, in its advection
 * tests, by hand (dynamics/test/Advection_test.cpp:222-240, AdvectionPeriodicBC_test.cpp:228-249).  IDynamics itself never
 * sets periodic edges (DynamicsKernel.hpp:54).
 *   dirichlet[e], ndirichlet[e] : element list of edge e (0 bottom, 1 right, 2 top, 3 left); a NULL list keeps the one
 *                                 derived from the mask by nsdg_set_mesh; dirichlet == NULL keeps all four
 *   periodic                    : all segments back to back, 4 longs per entry {type, c1, c2, edge}: type 0 = X-edge
 *                                 (c1 below, c2 above), 1 = Y-edge (c1 left of the edge, c2 right of it), edge = index of
 *                                 the edge whose normal velocity the flux uses (DGTransport.cpp:466-481)
 *   periodic_segment_sizes      : entries per segment (VectorManipulations::CGAveragePeriodic works segment by segment)
 * Effect: the advection applies upwind fluxes across the periodic edges (DGTransport.cpp:466-481), prepareIteration
 * averages cgH, cgA across the seam (CGDynamicsKernel.cpp:264-266) and every subcycle averages the stress divergence across
 * it before the momentum update (CGDynamicsKernel.cpp:395-397).  A handle with periodic edges runs the generic subcycle
 * kernels until the next nsdg_set_mesh.  Not available on partition boxes. */
int nsdg_set_boundaries(nsdg_handle h, const long* const* dirichlet, const size_t* ndirichlet, const long* periodic,
    const size_t* periodic_segment_sizes, size_t nsegments);

/* Replaces: the time loop of a DGTransport object used on its own -- nsteps x { DGTransport::reinitnormalvelocity
 * (dynamics/src/DGTransport.cpp:158-252); DGTransport::step with settimesteppingscheme("rk1" | "rk2" | "rk3")
 * (DGTransport.cpp:514-566); LimitMax / LimitMin (dynamics/src/include/dgLimit.hpp:16-84) } -- on one advected DG field
 * of the handle (NSDG_HICE, NSDG_CICE, NSDG_DAMAGE), with the DG velocity currently in the transport object
 * (nsdg_set_internal "velx" / "vely" = DGTransport::GetVx / GetVy; or the one the last nsdg_step projected).
 *   rk_order   : 1, 2, 3
 *   limit_mode : bit 0 = LimitMax(maxv), bit 1 = LimitMin(minv) after every step; 0 = none */
int nsdg_advect_field(nsdg_handle h, int field, double dt_seconds, int rk_order, int nsteps, int limit_mode, double maxv,
    double minv);

/* Device-resident synthetic forcing: evaluates the reference's benchmark atmosphere and ocean
 * (physics/src/modules/AtmosphereBoundaryModule/BenchmarkAtmosphere.cpp:38-74: moving cyclone;
 *  physics/src/modules/OceanBoundaryModule/BenchmarkOcean.cpp:27-36: steady gyre, ssh = 0; cell-corner
 *  coordinates of physics/src/BenchmarkCoordinates.cpp:20-42) directly on the GPU and feeds them through the
 * same DG0 -> CG path as nsdg_set_field(UWIND/VWIND/UOCEAN/VOCEAN/SSH), so that no forcing crosses PCIe.
 *   elapsed_seconds : time since the first update (tst.start - t0)
 *   domain_x/y      : extent of the GLOBAL domain in metres (the reference hard-codes 512e3) */
int nsdg_set_benchmark_forcing(nsdg_handle h, double elapsed_seconds, double domain_x, double domain_y);

/* The whole of MEVPDynamics::update / BBMDynamics::update in ONE call (MEVPDynamics.cpp:59-87,
 * BBMDynamics.cpp:66-101): uploads the 7 (8) input HFields from pinned staging with async copies,
 * steps, downloads the 6 (7) outputs.  Any pointer may be NULL to skip that field.
 * in/out arrays are nx*ny doubles each. */
typedef struct nsdg_update_io {
    const double *hice_in, *cice_in, *damage_in, *uwind, *vwind, *uocean, *vocean, *ssh;
    double *hice_out, *cice_out, *damage_out, *u_out, *v_out, *taux_out, *tauy_out;
} nsdg_update_io;
int nsdg_update(nsdg_handle h, const nsdg_update_io* io, double dt_seconds);

/* ---- mesh-derived integer state (bit-exact against ParametricMesh, ParametricMesh.cpp:217-292) ---- */
int nsdg_get_landmask(nsdg_handle h, unsigned char* out /* nx*ny */);
int nsdg_get_dirichlet(nsdg_handle h, int edge /* 0 bottom,1 right,2 top,3 left */, long* out, size_t capacity,
    size_t* count);

/* ---- test / diagnostic access to internal state, converted to the reference's layouts ----
 * names: "cg_u","cg_v","cgH","cgA","uGradSSH","vGradSSH","uOcean","vOcean","uAtmos","vAtmos","avgU","avgV",
 *        "lumpedcgmass" (flat CG vectors, cgVector.hpp:21-31);
 *        "hice","cice","damage","s11","s12","s22" (row-major N x comps, dgVector.hpp:89-91);
 *        "divS1","divS2","iMgradX","iMgradY","iMJwPSI","iMJwPSI_dam","divM","iMM" (N x rows x cols,
 *        ParametricMap.hpp:72-113), "AdvX","AdvY","iMass" (ParametricMap.hpp:22-27). */
int nsdg_get_internal(nsdg_handle h, const char* name, double* host, size_t capacity, size_t* count);
int nsdg_set_internal(nsdg_handle h, const char* name, const double* host, size_t count);

/* Damage healing on the device: Nextsim::ConstantHealing::updateElement
 * (physics/src/modules/DamageHealingModule/ConstantHealing.cpp:53-72) applied to the DG0 damage of a BBM handle, so that
 * the damage can stay resident between nsdg_step calls:
 *   damage = (damage (cice - g) + g) / cice,  g = max(0, delta_cice);  damage = min(1, damage + dt / td).
 * delta_cice: host array of nx*ny lateral concentration growth from the thermodynamics, or NULL for none.
 * td_seconds: ConstantHealing.td (default 15 days) in seconds. */
int nsdg_heal_damage(nsdg_handle h, double dt_seconds, double td_seconds, const double* delta_cice);

/* ---- restart state --------------------------------------------------------------------------------
 * Everything the dynamics carries from one timestep to the next, in the reference's layouts, as one flat buffer of
 * doubles: a 10-double header {magic, version, rheology, dgadv, cgdegree, nx, ny, nfields, box_x0, box_y0}, then per
 * field its length followed by its data (nsdg_set_state validates every header entry and length: a corrupted, truncated
 * or foreign buffer -- other mesh size, other partition box -- is rejected, never read out of bounds).  Fields: hice, cice [, damage] (N x DGadv, the prognostic DG fields a restart file
 * holds, BBMDynamics.cpp:104-132), the CG velocity u, v and the DG stresses s11, s12, s22 -- which the reference does
 * NOT checkpoint (CGDynamicsKernel.cpp:57 TODO; its restarts begin from zero stress) -- and, for BBM, the running-mean
 * velocity that advects the next step (BrittleCGDynamicsKernel.hpp:89).  get -> set on a fresh handle with the same
 * mesh resumes bit-reproducibly (tests/test_gpu_parity.py::test_restart_state_resumes_bitwise). */
int nsdg_get_state(nsdg_handle h, double* host, size_t capacity, size_t* count /* doubles written or needed */);
int nsdg_set_state(nsdg_handle h, const double* host, size_t count);

/* Run n bare subcycles on the current device state (no advection / prepare); the unit the
 * throughput metric counts.  Returns device milliseconds of the loop through *ms (CUDA events). */
int nsdg_subcycles(nsdg_handle h, int n, float* ms);

/* Average device time of the two kernels of one subcycle (subcycle_strip, subcycle_lines) over n
 * subcycles, each launch bracketed by CUDA events on the handle's stream (state advances by n subcycles). */
int nsdg_time_kernels(nsdg_handle h, int n, float* strip_ms, float* lines_ms);

/* Timing of the last nsdg_step, measured with CUDA events on the handle's stream. */
typedef struct nsdg_timing {
    float advection_ms, prepare_ms, subcycle_ms, total_ms;
    long kernel_launches; /* kernels launched by this library during the last nsdg_step/nsdg_subcycles */
    int uniform_path; /* 1 if the uniform-rectangular operator path is active */
    float halo_ms; /* partitioned boxes: average device time of one u,v halo exchange (nsdg_time_kernels) */
} nsdg_timing;
int nsdg_get_timing(nsdg_handle h, nsdg_timing* t);

/* ---- multi-GPU halo plumbing (no reference counterpart: the reference dynamics has no halo
 * exchange; semantics follow the 2-D box decomposition of run/partition.cdl) ----
 * Each box exports one device buffer per side that neighbours write into over NVLink; the
 * handles are CUDA IPC handles exchanged by the launcher (any transport).                    */
#define NSDG_IPC_HANDLE_BYTES 64
/* after nsdg_set_mesh: the CUDA IPC handle of this box's halo arena */
int nsdg_halo_export(nsdg_handle h, unsigned char* ipc_handle /* NSDG_IPC_HANDLE_BYTES */);
/* map the arena of the neighbour across `side` (enum nsdg_side) */
int nsdg_halo_connect(nsdg_handle h, int side, const unsigned char* peer_ipc_handle);
/* all neighbour sides connected: from now on nsdg_step / nsdg_update / nsdg_subcycles exchange halos.  Resets this box's
 * exchange epochs: after a re-mesh EVERY box of the partition calls export / connect / ready again, and none may start
 * exchanging before all have returned from nsdg_halo_ready (a barrier in the launcher).
 * A box whose neighbour does not deliver within 30 s of wall-clock time (device %globaltimer) gives up; the entry
 * point then returns an error ("halo exchange timed out") instead of a result computed on stale ring data. */
int nsdg_halo_ready(nsdg_handle h);

const char* nsdg_last_error(void);
const char* nsdg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NSDG_H */
