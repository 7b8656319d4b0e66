/* fluidsimulation_resident.cpp -- the particle<->grid stages AND the grid stages between them of
 * FluidSimulation::_stepFluid (/root/reference/src/fluidsimulation.cpp:3262-3390) on the device-resident path of
 * libgfs_b200:
 *
 *    stage  1  _updateFluidCells                  :1990-2040   -> gfs_set_particles (only when the host set changed),
 *                                                                 gfs_sort_index + gfs_p2g (classification + splat)
 *    stage  5  _advectVelocityField               :2597-2740   -> nothing left to do: the splat of stage 1 is the P2G slot
 *              _extrapolateFluidVelocities(saved) :3067-3070   -> gfs_copy_field(SAVED <- P2G) + gfs_extrapolate(SAVED)
 *    stage  6  _applyBodyForcesToVelocityField    :2846-2849   -> gfs_apply_body_force(P2G) (constant forces; variable
 *                                                                 force fields are host callbacks: the field makes one
 *                                                                 round trip through the reference's own loop)
 *    stage  7  _updatePressureGrid                :2870-2889   -> gfs_pressure_solve(P2G): the reference's MICCG(0)
 *    stage  8  _applyPressureToVelocityField      :3015-3061   -> gfs_apply_pressure(P2G -> NEW)
 *    stage  9  _extrapolateFluidVelocities(MAC)   :3067-3070   -> gfs_extrapolate(NEW); NEW is then copied down into
 *                                                                 _MACVelocity for the host stages that read it
 *                                                                 (diffuse particles, getVelocityField())
 *    stage 11  _updateMarkerParticleVelocities    :3104-3138   -> nothing to ship: NEW and SAVED are on the device
 *    stage 12  _advanceMarkerParticles            :3245-3256   -> gfs_g2p_advect (PIC/FLIP + RK4 + collision resolve),
 *                                                                 gfs_get_particles, then the reference's own
 *                                                                 _removeMarkerParticles
 *
 * These are DEFINITIONS OF THE REFERENCE'S OWN PRIVATE MEMBER FUNCTIONS: the unmodified fluidsimulation.cpp is compiled
 * as it lies, the definitions listed in oracle/Makefile (RESIDENT_WEAK) are made weak in the object file (objcopy
 * --weaken-symbol) and the strong ones below win at link time; every call site inside the reference (all PLT calls)
 * lands here.  Everything else of the simulator -- surface reconstruction, level set, diffuse particles, sources,
 * output -- is the reference's code, unchanged, on the host.
 *
 * Data kept in HBM across substeps: the marker particles (uploaded once; re-uploaded only when the host vector changed
 * size or a source / cell queue edited it), the material grid, the three field slots, the pressure system.  Per substep
 * over PCIe: the material grid both ways (1 B per cell), the final velocity field down, the advanced particles down (the
 * host stages between -- meshing, per-cell cap -- read the host vector).
 *
 * The Array3d<float> pressure grid the reference passes from stage 7 to stage 8 stays zero on the host: both stages are
 * here and the pressure lives on the device (gfs_get_pressure reads it).
 *
 * C++11, no CUDA types: only the C-ABI of include/gfs_b200.h.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include "fluidsimulation.h"
#include "gfs_b200.h"

namespace {

struct ResidentState {
    bool domain;                      // gfs_domain_init done
    bool deviceValid;                 // the device particle set equals the host vector as a multiset
    size_t count;                     // particles on the device
    std::vector<unsigned char> material;
    std::vector<gfs_marker_particle_t> staging;
    std::vector<gfs_grid_index_t> fluidCells;
    ResidentState() : domain(false), deviceValid(false), count(0) {}
};

void check(int err, const char *what) {
    if (err != GFS_SUCCESS) {
        std::cerr << "FluidSimulation (resident path): " << what << " failed: " << gfs_get_error_message() << std::endl;
        std::abort();                                     /* FLUIDSIM_ASSERT semantics, src/fluidsimassert.h */
    }
}

ResidentState *state(ParticleAdvector &adv) {
    if (!adv.residentState) {
        adv.residentState = std::shared_ptr<void>(new ResidentState(), [](void *p) { delete static_cast<ResidentState *>(p); });
    }
    return static_cast<ResidentState *>(adv.residentState.get());
}

}  // namespace

/* Stage 1.  The host-side edits of the particle set stay the reference's (removal in solid cells, added / removed cell
 * queues, sources: src/fluidsimulation.cpp:1991-1994); the marking loop (:1998-2017) and the fluid-cell list (:2019-2039)
 * come from the device: gfs_p2g classifies from the cell table of the sort and splats in the same pass. */
void FluidSimulation::_updateFluidCells() {
    const size_t before = _markerParticles.size();
    const bool edits = !_fluidSources.empty() || !_addedFluidCellQueue.empty() || !_removedFluidCellQueue.empty();
    _removeParticlesInSolidCells();
    _updateAddedFluidCellQueue();
    _updateRemovedFluidCellQueue();
    _updateFluidSources();

    ResidentState *rs = state(_particleAdvector);
    gfs_context *ctx = _particleAdvector.context();
    int err;
    if (!rs->domain) {
        gfs_domain_init(ctx, _isize, _jsize, _ksize, _dx, &err);
        check(err, "gfs_domain_init");
        rs->domain = true;
        rs->material.resize((size_t)_isize * _jsize * _ksize);
    }

    /* material: solids may have been edited through the public API since the last step */
    size_t c = 0;
    for (int k = 0; k < _ksize; k++)
        for (int j = 0; j < _jsize; j++)
            for (int i = 0; i < _isize; i++)
                rs->material[c++] = (unsigned char)_materialGrid(i, j, k);
    gfs_set_material(ctx, &rs->material[0], &err);
    check(err, "gfs_set_material");

    /* active inflow sources override the splatted velocity on set faces (:2588-2594), in insertion order */
    std::vector<gfs_source_t> sources;
    for (unsigned int s = 0; s < _fluidSources.size(); s++) {
        FluidSource *src = _fluidSources[s];
        if (!(src->isInflow() && src->isActive())) {
            continue;
        }
        gfs_source_t g;
        std::memset(&g, 0, sizeof(g));
        vmath::vec3 v = src->getVelocity();
        g.velocity[0] = v.x; g.velocity[1] = v.y; g.velocity[2] = v.z;
        if (SphericalFluidSource *sp = dynamic_cast<SphericalFluidSource *>(src)) {
            vmath::vec3 p = sp->getPosition();
            g.kind = 0; g.p[0] = p.x; g.p[1] = p.y; g.p[2] = p.z; g.a = sp->getRadius();
        } else {
            AABB bb = src->getAABB();
            g.kind = 1; g.p[0] = bb.position.x; g.p[1] = bb.position.y; g.p[2] = bb.position.z;
            g.a = bb.width; g.b = bb.height; g.c = bb.depth;
        }
        sources.push_back(g);
    }
    FLUIDSIM_ASSERT(sources.size() <= 8);
    gfs_set_sources(ctx, sources.empty() ? NULL : &sources[0], (int)sources.size(), &err);
    check(err, "gfs_set_sources");

    /* particles: the device set is reused unless the host vector was edited since the last download */
    const size_t n = _markerParticles.size();
    if (!rs->deviceValid || edits || n != before || n != rs->count) {
        rs->staging.resize(n);
        for (size_t p = 0; p < n; p++) {
            MarkerParticle mp = _markerParticles[p];
            gfs_marker_particle_t &d = rs->staging[p];
            d.position.x = mp.position.x; d.position.y = mp.position.y; d.position.z = mp.position.z;
            d.velocity.x = mp.velocity.x; d.velocity.y = mp.velocity.y; d.velocity.z = mp.velocity.z;
        }
        gfs_set_particles(ctx, n ? &rs->staging[0] : NULL, (int64_t)n, &err);
        check(err, "gfs_set_particles");
        rs->count = n;
        rs->deviceValid = true;
    }

    const int arith = _particleAdvector.isOpenCLEnabled() ? GFS_FAST : GFS_EXACT;
    if (arith == GFS_EXACT) {
        gfs_sort(ctx, &err);                              /* the exact gather sums in stable cell order */
    } else {
        gfs_sort_index(ctx, &err);
    }
    check(err, "gfs_sort");
    gfs_p2g(ctx, arith, &err);
    check(err, "gfs_p2g");

    gfs_stats_t st;
    gfs_get_stats(ctx, &st, &err);
    check(err, "gfs_get_stats");
    FLUIDSIM_ASSERT(st.in_solid == 0);                    /* :2015 */

    gfs_get_material(ctx, &rs->material[0], &err);
    check(err, "gfs_get_material");
    c = 0;
    for (int k = 0; k < _ksize; k++)
        for (int j = 0; j < _jsize; j++)
            for (int i = 0; i < _isize; i++) {
                const Material m = (Material)rs->material[c++];
                if (m != _materialGrid(i, j, k)) {
                    _materialGrid.set(i, j, k, m);
                }
            }
    /* the fluid-cell list in the reference's k, j, i order (:2019-2039), compacted on the device */
    int64_t nfluid = 0;
    rs->fluidCells.resize((size_t)st.fluid_cells + 1);
    gfs_get_fluid_cells(ctx, &rs->fluidCells[0], (int64_t)rs->fluidCells.size(), &nfluid, &err);
    check(err, "gfs_get_fluid_cells");
    FLUIDSIM_ASSERT(nfluid == st.fluid_cells);
    _fluidCellIndices.clear();
    _fluidCellIndices.reserve((size_t)nfluid);
    for (int64_t f = 0; f < nfluid; f++) {
        _fluidCellIndices.push_back(rs->fluidCells[(size_t)f].i, rs->fluidCells[(size_t)f].j, rs->fluidCells[(size_t)f].k);
    }
}

/* Stage 5: _advectVelocityFieldU/V/W each clear their component and refill it (:2597-2730); the three refilled arrays
 * are the P2G slot the splat of stage 1 left on the device.  The host _MACVelocity is refreshed after stage 9. */
void FluidSimulation::_advectVelocityField() {
}

/* Stages 5 (saved field) and 9 (solved field): MACVelocityField::extrapolateVelocityField(_materialGrid,
 * ceil(CFL + 2)) (:3067-3070) -- bit-identical on the device (gfs_extrapolate). */
void FluidSimulation::_extrapolateFluidVelocities(MACVelocityField &MACGrid) {
    const int numLayers = (int)ceil(_CFLConditionNumber + 2);
    gfs_context *ctx = _particleAdvector.context();
    int err;
    if (&MACGrid == &_savedVelocityField) {               /* "_savedVelocityField = _MACVelocity" (:3306) on the device */
        gfs_copy_field(ctx, GFS_FIELD_SAVED, GFS_FIELD_P2G, &err);
        check(err, "gfs_copy_field");
        gfs_extrapolate(ctx, GFS_FIELD_SAVED, numLayers, &err);
        check(err, "gfs_extrapolate(saved)");
    } else if (&MACGrid == &_MACVelocity) {
        gfs_extrapolate(ctx, GFS_FIELD_NEW, numLayers, &err);
        check(err, "gfs_extrapolate(new)");
        gfs_get_field(ctx, GFS_FIELD_NEW, _MACVelocity.getRawArrayU(), _MACVelocity.getRawArrayV(), _MACVelocity.getRawArrayW(), &err);
        check(err, "gfs_get_field");
    } else {
        MACGrid.extrapolateVelocityField(_materialGrid, numLayers);
    }
}

/* Stage 6 (:2846-2849).  Constant forces on the device; variable force fields are user callbacks on the host, so with
 * any registered the field makes one round trip through the reference's own two loops. */
void FluidSimulation::_applyBodyForcesToVelocityField(double dt) {
    gfs_context *ctx = _particleAdvector.context();
    int err;
    if (!_variableBodyForces.empty()) {
        gfs_get_field(ctx, GFS_FIELD_P2G, _MACVelocity.getRawArrayU(), _MACVelocity.getRawArrayV(), _MACVelocity.getRawArrayW(), &err);
        check(err, "gfs_get_field");
        _applyConstantBodyForces(dt);
        _applyVariableBodyForces(dt);
        gfs_set_field(ctx, GFS_FIELD_P2G, _MACVelocity.getRawArrayU(), _MACVelocity.getRawArrayV(), _MACVelocity.getRawArrayW(), &err);
        check(err, "gfs_set_field");
        return;
    }
    vmath::vec3 bf = _getConstantBodyForce();
    gfs_apply_body_force(ctx, GFS_FIELD_P2G, bf.x, bf.y, bf.z, dt, &err);
    check(err, "gfs_apply_body_force");
}

/* Stage 7 (:2870-2889): PressureSolver::solve with the reference's own parameters (PressureSolver's tolerance and
 * iteration limit are private constants, src/pressuresolver.h:159-160). */
void FluidSimulation::_updatePressureGrid(Array3d<float> &pressureGrid, double dt) {
    (void)pressureGrid;
    gfs_context *ctx = _particleAdvector.context();
    int err, iterations = 0;
    double residual = 0.0;
    gfs_pressure_solve(ctx, GFS_FIELD_P2G, dt, _density, 1e-6, 200, &iterations, &residual, &err);
    check(err, "gfs_pressure_solve");
    if (iterations >= 200) {
        _logfile.log("Iterations limit reached.\t Estimated error : ", residual, 1);      /* :502-503 */
    } else if (iterations >= 0) {
        _logfile.log("CG Iterations: ", iterations, 1);                                    /* :479 */
    }
}

/* Stage 8 (:3015-3061) */
void FluidSimulation::_applyPressureToVelocityField(Array3d<float> &pressureGrid, double dt) {
    (void)pressureGrid;
    gfs_context *ctx = _particleAdvector.context();
    int err;
    gfs_apply_pressure(ctx, GFS_FIELD_P2G, GFS_FIELD_NEW, dt, _density, &err);
    check(err, "gfs_apply_pressure");
}

/* Stage 11 needs both fields and no dt; stage 12 has the dt.  The fused G2P kernel does both stages in one pass over the
 * particles and both fields are already in their slots (the caller clears _savedVelocityField right after, :3350). */
void FluidSimulation::_updateMarkerParticleVelocities() {
}

/* Stages 11 + 12 on the device: PIC/FLIP update (:3118-3128), RK4 (src/particleadvector.cpp:1045-1054), solid test and
 * collision resolve (:3145-3209); then the host vector is refreshed and the reference's own shuffle + per-cell cap runs. */
void FluidSimulation::_advanceMarkerParticles(double dt) {
    ResidentState *rs = state(_particleAdvector);
    gfs_context *ctx = _particleAdvector.context();
    int err;
    const int arith = _particleAdvector.isOpenCLEnabled() ? GFS_FAST : GFS_EXACT;
    gfs_g2p_advect(ctx, dt, _ratioPICFLIP, 4, GFS_TRICUBIC, arith, &err);
    check(err, "gfs_g2p_advect");

    const size_t n = (size_t)gfs_num_particles(ctx, &err);
    check(err, "gfs_num_particles");
    FLUIDSIM_ASSERT(n == _markerParticles.size());
    rs->staging.resize(n);
    gfs_get_particles(ctx, n ? &rs->staging[0] : NULL, &err);
    check(err, "gfs_get_particles");
    for (size_t p = 0; p < n; p++) {
        const gfs_marker_particle_t &s = rs->staging[p];
        _markerParticles[p].position = vmath::vec3(s.position.x, s.position.y, s.position.z);
        _markerParticles[p].velocity = vmath::vec3(s.velocity.x, s.velocity.y, s.velocity.z);
    }
    rs->count = n;
    rs->deviceValid = true;

    _removeMarkerParticles();                             /* shuffle + at most 100 per cell (:3221-3243): may shrink the vector */
}
