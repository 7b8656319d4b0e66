/*
 * gfs_b200.h -- C-ABI of the B200-native PIC/FLIP particle<->grid transfer path.
 *
 * Drop-in boundary for rlguy/GridFluidSim3D's two accelerator classes (the only two that touch
 * OpenCL in the reference) and for the FluidSimulation stages that call them.  Citations are
 * file:line under /root/reference.
 *
 * Conventions follow the reference's own C bindings (src/c_bindings/cbindings.cpp:11-19,
 * src/c_bindings/fluidsimulation_c.cpp:15-77):
 *   - extern "C", opaque handle, plain pointers and sizes, no C++ or torch types;
 *   - every call takes a trailing `int *err`, set to GFS_SUCCESS (1) or GFS_FAIL (0);
 *   - on failure gfs_get_error_message() returns a NUL-terminated description (CUDA error string +
 *     call site) from a per-thread 4096-byte buffer;
 *   - POD structs are layout-identical to Vector3_t / MarkerParticle_t / GridIndex_t
 *     (src/c_bindings/vector3_c.h, markerparticle_c.h, gridindex_c.h).
 *
 * Data layout contracts (what crosses the boundary without conversion):
 *   - positions / velocities: packed float triples, 12 B (vmath::vec3, src/vmath.h:33-57)
 *   - particles: AoS {position, velocity}, 24 B (MarkerParticle, src/markerparticle.h:25-37)
 *   - grids: dense float, flat = i + width*(j + height*k) (Array3d, src/array3d.h:394-397);
 *     U is (I+1,J,K), V is (I,J+1,K), W is (I,J,K+1) floats, i.e. exactly
 *     MACVelocityField::getRawArrayU/V/W() (src/macvelocityfield.cpp:37-45, :87-97)
 *   - material: one byte per cell, 0 air / 1 fluid / 2 solid (src/fluidmaterialgrid.h:29-33)
 *
 * There is no CPU fallback: every entry point runs CUDA kernels on the context's device and fails
 * (err = 0) if that is impossible.
 */
#ifndef GFS_B200_H
#define GFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)      /* the library is built with -fvisibility=hidden */
#endif

#define GFS_SUCCESS 1
#define GFS_FAIL    0

#define GFS_AIR   0
#define GFS_FLUID 1
#define GFS_SOLID 2

/* interpolation of the staggered MAC field */
#define GFS_TRILINEAR 0   /* MACVelocityField::evaluateVelocityAtPositionLinear, src/macvelocityfield.cpp:561-575 */
#define GFS_TRICUBIC  1   /* MACVelocityField::evaluateVelocityAtPosition,       src/macvelocityfield.cpp:549-559 */

/* arithmetic mode */
#define GFS_FAST  0       /* fp32 tap contraction / fp32 kernel weights (index + fraction still exact) */
#define GFS_EXACT 1       /* the reference's own operation sequence in fp64, no contraction: bit-for-bit */

/* field slots of a device-resident domain */
#define GFS_FIELD_NEW    0   /* FluidSimulation::_MACVelocity at G2P time (post pressure solve) */
#define GFS_FIELD_SAVED  1   /* FluidSimulation::_savedVelocityField                           */
#define GFS_FIELD_P2G    2   /* what stage 5 (_advectVelocityField) produces                    */

typedef struct gfs_context gfs_context;

typedef struct gfs_vec3_t { float x, y, z; } gfs_vec3_t;                                   /* Vector3_t */
typedef struct gfs_marker_particle_t { gfs_vec3_t position, velocity; } gfs_marker_particle_t;  /* MarkerParticle_t */
typedef struct gfs_grid_index_t { int i, j, k; } gfs_grid_index_t;                           /* GridIndex_t */

/* An active inflow source, as FluidSimulation::_applyFluidSourceToVelocityField sees it
 * (src/fluidsimulation.cpp:2489-2524).  kind 0: sphere, centre p, radius a
 * (SphericalFluidSource::containsPoint, src/sphericalfluidsource.cpp:54-58); kind 1: cuboid, min corner
 * p, extents a,b,c (AABB::isPointInside, src/aabb.cpp:123-126). */
typedef struct gfs_source_t {
    int    kind;
    float  p[3];
    double a, b, c;
    float  velocity[3];
} gfs_source_t;

/* per-substep counters (device-resident path) */
typedef struct gfs_stats_t {
    int64_t num_particles;        /* particles currently resident                                   */
    int64_t out_of_grid;          /* particles whose cell lies outside the grid (ignored by P2G)     */
    int64_t in_solid;             /* particles whose cell is solid at classification (src/fluidsimulation.cpp:2015 asserts 0) */
    int64_t solid_hits;           /* particles whose advected position fell in a solid cell (resolved, or kept at p0 with option 3 = 0) */
    int64_t fluid_cells;          /* cells classified fluid by the last P2G                         */
    int64_t kernel_launches;      /* CUDA kernels launched by this context since creation (kernels inside replayed graphs included) */
    int64_t graph_replays;        /* substeps performed by replaying a captured CUDA graph (gfs_substep, steady state) */
    int64_t removed_particles;    /* particles removed by options 5 and 6 since creation                */
    int64_t collision_overflow;   /* colliders that did not fit the collision list since creation (they kept their old
                                     position instead of going through the reference's resolve); gfs_get_particles fails
                                     while it is non-zero -- raise option 8 */
} gfs_stats_t;

/* ---- context ------------------------------------------------------------------------------- */

/* Error text of the last failed call on this thread (CBindings_get_error_message,
 * src/c_bindings/cbindings.cpp:76-80). */
const char *gfs_get_error_message(void);

/* One context = one CUDA device + one stream.  `stream` is a cudaStream_t to enqueue on (so a caller
 * can time with its own events), or NULL to let the context create one.  Replaces the OpenCL
 * context/queue set-up of ParticleAdvector::initialize (src/particleadvector.cpp:30-58) and
 * CLScalarField::initialize (src/clscalarfield.cpp:27-58). */
gfs_context *gfs_create(int device, void *stream, int *err);
void gfs_destroy(gfs_context *ctx, int *err);
/* ParticleAdvector::getDeviceInfo (src/particleadvector.cpp:87-118) */
void gfs_device_info(gfs_context *ctx, char *buf, int buflen, int *err);
void gfs_sync(gfs_context *ctx, int *err);
void gfs_get_stats(gfs_context *ctx, gfs_stats_t *out, int *err);

/* Per-kernel timing with CUDA events on the context's stream (one event pair per launch while enabled).
 * gfs_profile_read synchronises, writes up to `cap` kernel names (64 bytes each, NUL-terminated) with their
 * summed milliseconds and launch counts, optionally resets, and returns the number of names. */
void gfs_profile_enable(gfs_context *ctx, int on, int *err);
int gfs_profile_read(gfs_context *ctx, char *names, double *total_ms, int64_t *counts, int cap, int reset, int *err);

/* ---- host-pointer, synchronous operators (mirror the accelerator classes one call each) ------- */

/* ParticleAdvector::tricubicInterpolate (src/particleadvector.cpp:401-455; NoCL body :1124-1137) and
 * its trilinear sibling.  out[n*3] = velocity at pos[n*3]; validate != 0 applies _validateOutput
 * (:1139-1149: any NaN/Inf component zeroes the vector). */
void gfs_sample(gfs_context *ctx, const float *pos, int64_t n,
                const float *u, const float *v, const float *w, int isize, int jsize, int ksize, double dx,
                int interp, int arith, int validate, float *out, int *err);

/* ParticleAdvector::advectParticlesRK1..4 (src/particleadvector.cpp:209-399; NoCL bodies :1045-1122).
 * order in 1..4. */
void gfs_advect(gfs_context *ctx, const float *pos, int64_t n,
                const float *u, const float *v, const float *w, int isize, int jsize, int ksize, double dx,
                double dt, int order, int interp, int arith, float *out, int *err);

/* MACVelocityField::extrapolateVelocityField(materialGrid, numLayers) (src/macvelocityfield.cpp:786-798) on the
 * caller's own arrays, in place: u, v, w are MACVelocityField::getRawArrayU/V/W() ((I+1)JK, I(J+1)K, IJ(K+1) floats),
 * material one byte per cell with the codes of src/fluidmaterialgrid.h:29-33.  Bit-identical to the reference. */
void gfs_extrapolate_field(gfs_context *ctx, float *u, float *v, float *w, int isize, int jsize, int ksize,
                           const uint8_t *material, int num_layers, int *err);

/* CLScalarField::addPointValues(points, values, radius, offset, dx, scalarfield, weightfield)
 * (src/clscalarfield.cpp:200-267) == ScalarField::addPointValue per point (src/scalarfield.cpp:167-201).
 * field/weight are (ni,nj,nk) float grids; weight may be NULL (the no-weight overload, :147-198).
 * accumulate != 0 adds into the caller's arrays (the OpenCL path's semantics, :1427-1455);
 * accumulate == 0 overwrites them (the NoCL path's semantics, :1521-1551). */
void gfs_add_point_values(gfs_context *ctx, const float *pos, const float *values, int64_t n,
                          double radius, const float *offset3, double dx, int ni, int nj, int nk,
                          float *field, float *weight, int accumulate, int arith, int *err);

/* CLScalarField::addPoints(points, radius, offset, dx, field) (src/clscalarfield.cpp:62-145): gfs_add_point_values with
 * value 1 for every point and no weight grid.  use_threshold != 0 mirrors setMaxScalarFieldValueThreshold(threshold)
 * (src/clscalarfield.cpp:303-310; set by IsotropicParticleMesher, src/isotropicparticlemesher.cpp:334-359): a node whose
 * value before this call already exceeds `threshold` receives nothing from it -- the per-node, order-independent form of
 * the reference's skip rules (src/scalarfield.cpp:182-184 on the CPU, whole chunks at src/clscalarfield.cpp:1002-1010 in
 * OpenCL; its own no-OpenCL body ignores the threshold, :1490-1505).  Needs accumulate != 0 to have a "before". */
void gfs_add_points(gfs_context *ctx, const float *pos, int64_t n, double radius, const float *offset3, double dx,
                    int ni, int nj, int nk, float *field, int accumulate, int use_threshold, float threshold, int arith, int *err);

/* ---- device-resident domain (particles, fields and material stay in HBM between calls) -------- */

/* Allocate grids for an isize x jsize x ksize domain of cell size dx (FluidSimulation(isize,jsize,ksize,dx),
 * src/fluidsimulation.cpp:25-32) and mark the border cells solid (_initializeSolidCells, :1191-1213). */
void gfs_domain_init(gfs_context *ctx, int isize, int jsize, int ksize, double dx, int *err);
void gfs_set_material(gfs_context *ctx, const uint8_t *material, int *err);
void gfs_get_material(gfs_context *ctx, uint8_t *material, int *err);
/* FluidSimulation::_fluidCellIndices (src/fluidsimulation.cpp:2019-2039): the fluid cells of the resident material grid in
 * the reference's k, j, i scan order, compacted on the device (flags -> exclusive scan -> triples).  At most `capacity`
 * cells are written; *count receives the number of fluid cells (capacity 0 sizes the buffer).  Single domain. */
void gfs_get_fluid_cells(gfs_context *ctx, gfs_grid_index_t *cells, int64_t capacity, int64_t *count, int *err);
/* inflow sources used by P2G (copied) */
void gfs_set_sources(gfs_context *ctx, const gfs_source_t *sources, int nsources, int *err);
/* FluidSimulation::_updateFluidSources on the resident particles (src/fluidsimulation.cpp:1771-1879).
 * gfs_emit_from_sources: every ACTIVE INFLOW source given to gfs_set_sources seeds the air cells it overlaps with 8 particles
 * (_addNewFluidCells, :1749-1759) and puts one particle into every empty half-dx sub-cell of its fluid-or-air cells
 * (_getNewFluidParticles, :1771-1821), all with the source's velocity.  jitter = 0.25 * jitter factor * dx (:1219-1221).
 * The reference draws the jitter from rand(); here it is a hash of (seed, cell, sub-cell): the set of emitting sub-cells is
 * the reference's, positions agree to within the jitter.  gfs_remove_in_sources: the particles in the FLUID cells that the
 * given outflow sources overlap are removed (:1853-1877, _removeMarkerParticlesFromCells :1717-1730).  Both read the
 * resident material grid as the previous classification left it, like the reference; single domain. */
void gfs_emit_from_sources(gfs_context *ctx, double jitter, uint64_t seed, int64_t *emitted, int *err);
void gfs_remove_in_sources(gfs_context *ctx, const gfs_source_t *outflow_sources, int nsources, int64_t *removed, int *err);
/* AoS MarkerParticle_t[n] <-> device SoA */
void gfs_set_particles(gfs_context *ctx, const gfs_marker_particle_t *particles, int64_t n, int *err);
int64_t gfs_num_particles(gfs_context *ctx, int *err);
void gfs_get_particles(gfs_context *ctx, gfs_marker_particle_t *particles, int *err);
/* order[r] = index, in the array last passed to gfs_set_particles, of the particle now stored at r */
void gfs_get_particle_order(gfs_context *ctx, int32_t *order, int *err);
/* raw u,v,w arrays in the reference's layout; slot is GFS_FIELD_* */
void gfs_set_field(gfs_context *ctx, int slot, const float *u, const float *v, const float *w, int *err);
void gfs_get_field(gfs_context *ctx, int slot, float *u, float *v, float *w, int *err);

/* the same for the cell layers [k_first, k_first + k_count) only (u, v, w, material: the caller's WHOLE arrays; the w array's
 * extra face layer k_first + k_count travels too): what a z-slab rank moves -- owned layers + halo up, owned layers down */
void gfs_set_field_layers(gfs_context *ctx, int slot, const float *u, const float *v, const float *w, int k_first, int k_count, int *err);
void gfs_get_field_layers(gfs_context *ctx, int slot, float *u, float *v, float *w, int k_first, int k_count, int *err);
void gfs_get_material_layers(gfs_context *ctx, uint8_t *material, int k_first, int k_count, int *err);

/* K0: bin particles by cell (brick-major key) and build the cell table.  gfs_sort is stable (radix sort: particles
 * of one cell keep their relative order, which the exact-arithmetic P2G relies on); gfs_sort_unstable is the
 * counting sort the fast substep uses (the order inside a cell is unspecified; nothing in fast mode depends on it). */
void gfs_sort(gfs_context *ctx, int *err);
void gfs_sort_unstable(gfs_context *ctx, int *err);

/* Grid-only neighbours of the path, so that the P2G output and the G2P input can stay on the device:
 * gfs_extrapolate = MACVelocityField::extrapolateVelocityField(materialGrid, num_layers) (src/macvelocityfield.cpp:786-798)
 * on the resident field `slot` with the resident material grid; bit-identical to the reference.  FluidSimulation calls
 * it with num_layers = ceil(CFL + 2) on the saved field after P2G and on the solved field before G2P
 * (src/fluidsimulation.cpp:3067-3070, 3306-3307, 3334).  gfs_copy_field: dst slot := src slot (":3306"). */
void gfs_extrapolate(gfs_context *ctx, int slot, int num_layers, int *err);

/* Stages 6-8 of FluidSimulation::_stepFluid on the resident grid, single domain.
 *
 * gfs_apply_body_force = FluidSimulation::_applyConstantBodyForces (src/fluidsimulation.cpp:2765-2805): adds
 *     (float)(force * dt) to every face of `slot` that borders a fluid cell; zero components are skipped.
 * gfs_pressure_solve  = PressureSolver::solve behind FluidSimulation::_updatePressureGrid (src/pressuresolver.cpp:116-505,
 *     src/fluidsimulation.cpp:2870-2889): the reference's MICCG(0) -- same right-hand side, same modified incomplete
 *     Cholesky factor, same substitutions and updates, operation for operation in double; the sequential sweeps run as
 *     tile wavefronts, which respects every data dependency and therefore gives the same bits.  Only the summation order
 *     of the dot products differs (last-place effects on alpha / beta).  The reference's parameters are density 20.0
 *     (src/fluidsimulation.h:1154), tolerance 1e-6 and 200 iterations (src/pressuresolver.h:159-160).  *iterations
 *     receives the reference's iterationNumber at return (-1: right-hand side below the tolerance, pressure 0;
 *     max_iterations: limit reached, estimate kept), *residual the last max |residual|; both may be NULL.
 * gfs_apply_pressure  = FluidSimulation::_applyPressureToVelocityField (src/fluidsimulation.cpp:2895-3061) with the
 *     float pressure grid of the last solve: dst_slot := src_slot projected (faces bordering fluid: 0 next to a solid,
 *     U - dt/(density dx) (p1 - p0) otherwise; all other faces copied).  src_slot == dst_slot is allowed.
 * gfs_get_pressure    : that float grid, isize*jsize*ksize values, i fastest (0 outside fluid cells). */
void gfs_apply_body_force(gfs_context *ctx, int slot, float fx, float fy, float fz, double dt, int *err);
void gfs_pressure_solve(gfs_context *ctx, int slot, double dt, double density, double tolerance, int max_iterations,
                        int *iterations, double *residual, int *err);
void gfs_apply_pressure(gfs_context *ctx, int src_slot, int dst_slot, double dt, double density, int *err);
void gfs_get_pressure(gfs_context *ctx, float *pressure, int *err);
/* gfs_pressure_solve_field: the same solve on caller-owned host arrays -- the one-call body of PressureSolver::solve
 * (src/pressuresolver.cpp:116-139) for a simulator that keeps its grids on the host (dropin/pressuresolver.cpp).
 * u, v, w = MACVelocityField::getRawArrayU/V/W(), material = one byte per cell; pressure receives isize*jsize*ksize
 * DOUBLES (the solver's precision; i fastest; 0 outside fluid cells). */
void gfs_pressure_solve_field(gfs_context *ctx, const float *u, const float *v, const float *w, int isize, int jsize, int ksize,
                              double dx, const uint8_t *material, double dt, double density, double tolerance,
                              int max_iterations, double *pressure, int *iterations, double *residual, int *err);
void gfs_copy_field(gfs_context *ctx, int dst_slot, int src_slot, int *err);
/* gfs_sort_index: the counting sort without moving the particles -- only the sorted index is materialised and the
 * P2G / G2P kernels fetch through it (G2P stores its results in sorted order).  What gfs_substep does internally;
 * gfs_get_particles afterwards returns the storage order, not the sorted one. */
void gfs_sort_index(gfs_context *ctx, int *err);
/* Tuning switches.  option 0: fast-P2G variant, 0 = global atomics only; brick tiles in shared memory: 1 = round-1
 * kernel, 2 = round-2 kernel, 3 = round-2 kernel with lanes transposed through shared memory (default); all produce
 * bit-identical grids.  option 1: fast-G2P variant, 0 = global loads only; TMA-staged brick tiles (dx a power of two,
 * sorted particles): 1 = round-1 kernel, 2 = round-2 trilinear kernel (default), 3 = the same on a bank-skewed tile;
 * bit-identical results.
 * option 2: 1 = gfs_substep / gfs_sort_index sort by index only once the storage is nearly sorted (default), 0 = always
 * move the particles.  option 3: 1 = particles advected into a solid cell go through the reference's collision resolve
 * (FluidSimulation::_resolveParticleSolidCellCollision, src/fluidsimulation.cpp:3145-3179; default), 0 = they keep
 * their old position (bare solid test).  gfs_stats_t.solid_hits counts them either way.  option 4: 1 = gfs_substep
 * captures its launch sequence into a CUDA graph (one per buffer parity) and replays it while the particle count, the
 * step parameters and every buffer stay the same (default), 0 = always launch kernel by kernel; identical results.
 * option 5: per-cell particle cap, FluidSimulation::_removeMarkerParticles (src/fluidsimulation.cpp:3221-3243; the
 * reference uses 100): every binning pass keeps at most `value` particles per cell and removes the rest -- which ones
 * is arbitrary, as with the reference's rand() shuffle; 0 = no cap (default).  option 6: 1 = a particle found inside a
 * solid cell by a sort is removed (FluidSimulation::_removeMarkerParticlesInSolidCells, :1933-1957), 0 = it is kept and
 * counted in gfs_stats_t.in_solid (default).  Both cost one 4-byte read-back per binning pass and apply to the
 * single-domain entry points (gfs_sort*, gfs_substep, gfs_g2p_advect; the sharded gfs_comm_* entry points refuse to run
 * with them on); gfs_stats_t.removed_particles counts.  option 7: limit, in seconds (default 4), of the device-side
 * waits of the peer exchange; a wait that exceeds it fails gfs_comm_migrate_finish / gfs_comm_substep.  option 8:
 * capacity of the collision list in particles (default 0 = n/16 + 4096); see gfs_stats_t.collision_overflow.  option 9:
 * 1 = gfs_comm_substep posts this rank's max |v| right after its G2P (default), 0 = the all-ranks maximum is exchanged
 * where it is needed, between the sort and the splat.  option 10: 1 = the device-side wait for a neighbour's layers runs
 * in a single-thread kernel of its own instead of inside the unpack kernel (needed when several slabs share one GPU, where
 * the spinning CTAs of one context would starve the other's push; gfs_mg_create sets it), 0 = fused (default).  option 11:
 * 1 = normalisation, isValueSet, inflow override and face assembly in one pass straight from the accumulators (default),
 * 0 = the two-kernel form with the node grid and its mask in HBM; identical results.  option 12: substitution sweeps of
 * gfs_pressure_solve: 0 = inputs read from global memory inside the dependent steps, tiles synchronised by completion
 * flags; 1 = each tile's inputs staged in shared memory before it waits; 2 = staged, and synchronised by the data itself
 * (cells hold a sentinel until produced; no flags, no fences), a tile waiting for its whole halo first (default); 3 = as
 * 2, but a lane waits for a halo value only at the step that needs it (measured slower: an L2 round trip lands on every
 * boundary step of the dependent chain).  Identical results.  option 13: 1 = record per-tile timestamps of the last
 * substitution sweeps (gfs_device_ptr 46; profiles/press_trace.py reads them), 0 = off (default). */
void gfs_set_option(gfs_context *ctx, int option, int value, int *err);
/* K1: stage 1 + stage 5 of _stepFluid on the resident particles: material classification
 * (src/fluidsimulation.cpp:1998-2017), u/v/w splat + normalisation + inflow override + bordering-fluid
 * assembly (:2526-2730).  Result in slot GFS_FIELD_P2G; material updated.  Requires gfs_sort. */
void gfs_p2g(gfs_context *ctx, int arith, int *err);
/* K2: stage 11 + stage 12 (src/fluidsimulation.cpp:3104-3129, :3181-3209): PIC/FLIP velocity update from slots NEW
 * and SAVED, RK advance through NEW, solid test + collision resolve (:3145-3179, option 3).  The reference's shuffle
 * is not reproduced; its per-cell cap is option 5. */
void gfs_g2p_advect(gfs_context *ctx, double dt, double ratio_picflip, int order, int interp, int arith, int *err);
/* gfs_sort + gfs_p2g + gfs_g2p_advect, stream-ordered, no host synchronisation (fast arithmetic: index-only counting sort
 * binned by the previous call's G2P epilogue; the steady-state launch sequence is replayed from a CUDA graph, option 4). */
void gfs_substep(gfs_context *ctx, double dt, double ratio_picflip, int order, int interp, int arith, int *err);

/* Advection only, device resident (ParticleAdvector::advectParticlesRK1..4, src/particleadvector.cpp:209-399, on the
 * resident cell-sorted particles): index sort + RK `rk_order` through field slot NEW with the trilinear brick kernel,
 * positions only (12 B read + 12 B written per particle; binned for the next call in the kernel's epilogue).  No solid
 * test, no velocity update: the velocity arrays are undefined afterwards (gfs_p2g / gfs_g2p_advect refuse to run until
 * the particles are uploaded again).  Power-of-two dx only.  This is BASELINE configs[4]'s advection sweep operator. */
void gfs_advect_substep(gfs_context *ctx, double dt, int rk_order, int *err);

/* ---- z-slab sharding across GPUs (one context per GPU; the exchange itself is the caller's: NCCL) ----------
 * Every rank allocates the whole grid but OWNS the cell layers [k0,k1): its particles are those whose cell lies in
 * them, and its grid kernels only touch the owned layers plus one halo layer.  A sharded P2G is
 *   gfs_sort_unstable; gfs_p2g_begin;  exchange(accumulator layers: add; material layers: copy);  gfs_p2g_end
 * Partial sums are 64-bit integers, so the merged grid is bit-identical to the single-GPU one.
 * Layer arrays (`what`): 0..2 NEW u,v,w; 3..5 SAVED u,v,w; 6..8 P2G u,v,w; 9 material; 10..12 accumulators of u,v,w. */
void gfs_set_owned_layers(gfs_context *ctx, int k0, int k1, int *err);
void gfs_p2g_begin(gfs_context *ctx, int arith, int *err);
void gfs_p2g_end(gfs_context *ctx, int *err);
int64_t gfs_layer_bytes(gfs_context *ctx, int what, int *err);
/* copy z-layers [k_first, k_first+k_count) of a resident array to / from a caller-owned DEVICE buffer; add != 0
 * (accumulators only) adds the buffer as 64-bit integers instead of overwriting */
void gfs_pack_layers(gfs_context *ctx, int what, int k_first, int k_count, void *dst_device, int *err);
void gfs_unpack_layers(gfs_context *ctx, int what, int k_first, int k_count, const void *src_device, int add, int *err);
/* the same for up to 16 layer ranges at byte offsets of ONE device buffer, in one kernel launch; direction 0 = pack,
 * 1 = unpack (add[i] != 0: 64-bit integer add, accumulators only) */
void gfs_copy_layers_batch(gfs_context *ctx, int direction, int n, const int *what, const int *k_first, const int *k_count,
                           const int64_t *offsets, const int *add, void *buffer_device, int *err);
/* particle migration: remove the particles whose cell layer is < k_lo (written as MarkerParticle_t AoS to down_device)
 * or >= k_hi (to up_device); cap = capacity of each buffer in particles.  Synchronises to return the counts. */
void gfs_extract_particles(gfs_context *ctx, int k_lo, int k_hi, void *down_device, void *up_device, int64_t cap,
                           int64_t *n_down, int64_t *n_up, int *err);
/* the same split without any host synchronisation: counts {kept, down, up, spare} (4 x uint32) stay in caller-owned
 * DEVICE memory; after reading them back the caller commits the new particle count */
void gfs_extract_particles_async(gfs_context *ctx, int k_lo, int k_hi, void *down_device, void *up_device, int64_t cap,
                                 void *counters_device, int *err);
void gfs_extract_commit(gfs_context *ctx, int64_t n_kept, int *err);
void gfs_append_particles_device(gfs_context *ctx, const void *aos_device, int64_t n, int *err);

/* ---- peer-memory exchange (CUDA IPC, one node): the neighbour GPU writes layers / particles / flags straight into
 * this context's "comm blocks" over NVLink; waits are device-side spins on those flags.  side: 0 = down, 1 = up.
 *   gfs_comm_alloc(ctx, bytes of the largest layer message, arrival capacity in particles)   once, after domain_init
 *   gfs_comm_export(ctx, side, handle[64])  -> ship the 64 bytes to the neighbour on that side (any channel)
 *   gfs_comm_connect(ctx, side, handle[64]) <- the neighbour's handle for ITS opposite side
 *   per substep:  gfs_comm_push_layers on every side, then gfs_comm_pull_layers on every side   (C1 / C2)
 *                 gfs_comm_migrate_begin, gfs_comm_migrate_finish (the one host synchronisation)   (C3)
 * Every rank must issue the same sequence of exchanges. */
void gfs_comm_alloc(gfs_context *ctx, int64_t layer_bytes, int64_t particle_cap, int *err);
void gfs_comm_export(gfs_context *ctx, int side, void *handle64, int *err);
void gfs_comm_connect(gfs_context *ctx, int side, const void *handle64, int *err);
void gfs_comm_connect_local(gfs_context *ctx, int side, gfs_context *neighbour, int *err);   /* same process (tests) */
void gfs_comm_push_layers(gfs_context *ctx, int side, int n, const int *what, const int *k_first, const int *k_count,
                          const int64_t *offsets, int *err);
void gfs_comm_pull_layers(gfs_context *ctx, int side, int n, const int *what, const int *k_first, const int *k_count,
                          const int64_t *offsets, const int *add, int *err);
void gfs_comm_migrate_begin(gfs_context *ctx, int has_down, int has_up, int *err);
/* gfs_g2p_advect with the migration fused into the G2P kernel (leavers are stored into the neighbour's arrival buffer
 * by the kernel that advects them; stayers are binned for the next gfs_sort_index).  Replaces gfs_g2p_advect +
 * gfs_comm_migrate_begin; falls back to exactly that pair where the brick kernel does not apply. */
void gfs_comm_g2p_advect(gfs_context *ctx, double dt, double picflip_ratio, int rk_order, int interp, int arith,
                         int has_down, int has_up, int *err);
void gfs_comm_migrate_finish(gfs_context *ctx, int64_t *moved2, int *err);
/* The whole sharded substep of one rank in one call (no interpreter between the launches): gfs_comm_set_plan stores the
 * merged C1 + C2 exchange of a side once (arguments as push_layers / pull_layers), gfs_comm_substep then runs
 * sort_index, allmax_scale, p2g_begin, push/pull, p2g_end, comm_g2p_advect and migrate_finish. */
void gfs_comm_set_plan(gfs_context *ctx, int side, int n_push, const int *push_what, const int *push_first, const int *push_count,
                       const int64_t *push_offsets, int n_pull, const int *pull_what, const int *pull_first, const int *pull_count,
                       const int64_t *pull_offsets, const int *pull_add, int *err);
void gfs_comm_substep(gfs_context *ctx, double dt, double picflip_ratio, int rk_order, int interp, int arith,
                      int has_down, int has_up, int64_t *moved2, int *err);

/* All-ranks maximum over peer memory (<= 16 GPUs of one node).  The fixed-point scale of the P2G accumulators derives
 * from max |v| over ALL particles of the domain, so that the integer partial sums of different GPUs are commensurable
 * and the sharded result equals the single-GPU one bit for bit: call gfs_comm_allmax_scale between the sort and
 * gfs_p2g_begin.  Setup: world_alloc, ship world_export's 64 bytes to everyone, world_connect for every other rank. */
void gfs_comm_world_alloc(gfs_context *ctx, int rank, int world, int *err);
void gfs_comm_world_export(gfs_context *ctx, void *handle64, int *err);
void gfs_comm_world_connect(gfs_context *ctx, int rank, const void *handle64, int *err);
void gfs_comm_world_connect_local(gfs_context *ctx, int rank, gfs_context *other, int *err);
void gfs_comm_allmax_scale(gfs_context *ctx, int *err);
/* The two halves apart: gfs_comm_allmax_post sends this rank's max |v| to every rank as soon as it is final (after the
 * G2P of the previous substep: its epilogue takes the maximum over every particle it advected), gfs_comm_allmax_scale
 * then only waits for the others' values -- a whole sort later, off the critical path.  gfs_comm_substep does this. */
void gfs_comm_allmax_post(gfs_context *ctx, int *err);

/* Raw device pointers of resident buffers for zero-copy interop (halo exchange by the multi-GPU driver).
 * which: 0..2 NEW u,v,w; 3..5 SAVED u,v,w; 6..8 P2G u,v,w; 9 material; 10..15 particle x,y,z,vx,vy,vz;
 * 16 the word holding max |v| of the resident particles (float bits; the P2G fixed-point scale derives from it --
 * sharded runs must replace it by the maximum over all ranks between the sort and gfs_p2g_begin); 40..45 the dense
 * double vectors of the last pressure solve: residual, auxillary, search, pressure, forward-solve temporary, MIC(0)
 * diagonal (verification hook of tests/test_gpu_pressure.py); 46 the per-tile timestamps of option 13. */
void *gfs_device_ptr(gfs_context *ctx, int which, int *err);
/* Verification hook: order- and distribution-independent 64-bit hashes of the resident state.  out5[0] = material of the
 * owned cell layers, out5[1..3] = P2G u, v, w faces of the owned layers, out5[4] = the particle set.  Each is a sum mod
 * 2^64 of per-element hashes keyed by GLOBAL element index (particles: by their six words), so the per-rank hashes of a
 * z-slab run add up to the single-GPU run's.  Synchronises; compacts dead slots like gfs_get_particles. */
void gfs_state_hash(gfs_context *ctx, uint64_t *out5, int *err);
/* Allocate now what a substep would size lazily (particle arrays for particle_capacity slots, sort / G2P / exchange
 * scratch): no cudaMalloc / cudaFree -- device-wide synchronisations -- inside the steps that follow. */
void gfs_reserve(gfs_context *ctx, int64_t particle_capacity, int *err);
/* Resize the resident particle set to n (contents of [0,min(old,n)) kept) -- used by slab migration. */
void gfs_resize_particles(gfs_context *ctx, int64_t n, int *err);

/* ---- native multi-GPU group (gfs_mg_create / scatter / substep / gather) --------------------------------
 * One process, one context and one host thread per GPU, z-slab sharding with particle-weighted cuts, neighbour exchange
 * through peer memory (cudaDeviceEnablePeerAccess; the kernels of one GPU write layers, migrating particles and flags
 * straight into its neighbour's HBM).  No NCCL, no IPC handles, no interpreter: what a C++11 host of the reference calls
 * in place of the single-GPU entry points.  The sharded result is bit-identical to the single-GPU one (integer partial
 * sums; gfs_mg_state_hash equals gfs_state_hash of the unsharded run).  `devices` may be NULL (GPUs 0..ndev-1) and may
 * name a GPU more than once (several slabs on one GPU: tests).  halo_layers: cell layers of the NEW / SAVED fields a
 * slab needs from each neighbour = gfs_slab_halo_cells(interp, max displacement per substep, dx).  Errors of these
 * entry points are read with gfs_mg_get_error_message(). */
typedef struct gfs_mg gfs_mg;
const char *gfs_mg_get_error_message(void);
gfs_mg *gfs_mg_create(int ndev, const int *devices, int isize, int jsize, int ksize, double dx, int halo_layers, int *err);
void gfs_mg_destroy(gfs_mg *mg, int *err);
int gfs_mg_num_devices(gfs_mg *mg);
gfs_context *gfs_mg_context(gfs_mg *mg, int rank);            /* the per-GPU context (stats, options, profiling) */
void gfs_mg_get_slab(gfs_mg *mg, int rank, int *k0, int *k1, int *err);
void gfs_mg_set_option(gfs_mg *mg, int option, int value, int *err);
void gfs_mg_set_material(gfs_mg *mg, const uint8_t *material, int *err);
void gfs_mg_set_sources(gfs_mg *mg, const gfs_source_t *sources, int nsources, int *err);
/* u, v, w: whole arrays in the reference's layout; every GPU takes (delivers) its own layers (+ halo on the way in) */
void gfs_mg_set_field(gfs_mg *mg, int slot, const float *u, const float *v, const float *w, int *err);
void gfs_mg_get_field(gfs_mg *mg, int slot, float *u, float *v, float *w, int *err);
void gfs_mg_get_material(gfs_mg *mg, uint8_t *material, int *err);
/* distribute the particles over the slabs (cuts balance the particle count) and build the exchange plans */
void gfs_mg_scatter_particles(gfs_mg *mg, const gfs_marker_particle_t *particles, int64_t n, int *err);
int64_t gfs_mg_num_particles(gfs_mg *mg, int *err);
void gfs_mg_gather_particles(gfs_mg *mg, gfs_marker_particle_t *particles, int *err);
/* gfs_substep on every GPU with the slab exchange (gfs_comm_substep per rank, one host thread each) */
void gfs_mg_substep(gfs_mg *mg, double dt, double ratio_picflip, int rk_order, int interp, int arith, int64_t *moved2, int *err);
void gfs_mg_state_hash(gfs_mg *mg, uint64_t *out5, int *err);
void gfs_mg_sync(gfs_mg *mg, int *err);

/* ---- z-slab helpers for the multi-GPU driver (host-only arithmetic, usable without a GPU) ----- */

/* Cell range [k0,k1) owned by `rank` of `nranks` z-slabs over ksize cells. */
void gfs_slab_range(int ksize, int nranks, int rank, int *k0, int *k1, int *err);
/* Rank owning cell layer k. */
int gfs_slab_owner(int ksize, int nranks, int k, int *err);
/* Ghost-layer width (cells) a slab needs for G2P: stencil radius of `interp` + ceil(max_displacement/dx). */
int gfs_slab_halo_cells(int interp, double max_displacement, double dx, int *err);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
