/* fluidsimulation_resident.cpp -- the four particle<->grid stages of FluidSimulation::_stepFluid
 * (/root/reference/src/fluidsimulation.cpp:3262-3390) on the device-resident path of libgfs_b200:
 *
 *    stage  1  _updateFluidCells                  :1990-2040   -> gfs_set_particles (only when the host set changed),
 *                                                                 gfs_sort_index + gfs_p2g (classification + splat)
 *    stage  5  _advectVelocityField               :2597-2740   -> gfs_get_field(P2G) into _MACVelocity
 *    stage 11  _updateMarkerParticleVelocities    :3104-3138   -> gfs_set_field(NEW, SAVED)
 *    stage 12  _advanceMarkerParticles            :3245-3256   -> gfs_g2p_advect (PIC/FLIP + RK4 + collision resolve),
 *                                                                 gfs_get_particles, then the reference's own
 *                                                                 _removeMarkerParticles
 *
 * These are DEFINITIONS OF THE REFERENCE'S OWN PRIVATE MEMBER FUNCTIONS: the unmodified fluidsimulation.cpp is compiled
 * as it lies, its four definitions are made weak in the object file (objcopy --weaken-symbol, oracle/Makefile target
 * `resident`) and the strong ones below win at link time; every call site inside the reference (all PLT calls) lands
 * here.  Everything else of the simulator -- surface reconstruction, level set, body forces, pressure solve,
 * extrapolation, diffuse particles, output -- is the reference's code, unchanged, on the host.
 *
 * Data kept in HBM across substeps: the marker particles (uploaded once; re-uploaded only when the host vector changed
 * size or a source / cell queue edited it), the material grid, the three field slots.  Per substep over PCIe: the
 * material grid and the P2G field down, the post-pressure and the saved field up, the advanced particles down (the
 * host stages between -- meshing, per-cell cap -- read the host vector).
 *
 * C++11, no CUDA types: only the C-ABI of include/gfs_b200.h.
 */
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
    _fluidCellIndices.clear();
    _fluidCellIndices.reserve((size_t)st.fluid_cells);
    c = 0;
    for (int k = 0; k < _ksize; k++)
        for (int j = 0; j < _jsize; j++)
            for (int i = 0; i < _isize; i++) {
                const Material m = (Material)rs->material[c++];
                if (m != _materialGrid(i, j, k)) {
                    _materialGrid.set(i, j, k, m);
                }
                if (m == Material::fluid) {
                    _fluidCellIndices.push_back(i, j, k);
                }
            }
}

/* Stage 5: _advectVelocityFieldU/V/W each clear their component and refill it (:2597-2730); the three refilled arrays
 * are the P2G slot the splat of stage 1 left on the device. */
void FluidSimulation::_advectVelocityField() {
    gfs_context *ctx = _particleAdvector.context();
    int err;
    gfs_get_field(ctx, GFS_FIELD_P2G, _MACVelocity.getRawArrayU(), _MACVelocity.getRawArrayV(), _MACVelocity.getRawArrayW(), &err);
    check(err, "gfs_get_field");
}

/* Stage 11 needs both fields and no dt; stage 12 has the dt.  The fused G2P kernel does both stages in one pass over the
 * particles, so stage 11 only ships the two fields (the caller clears _savedVelocityField right after, :3350). */
void FluidSimulation::_updateMarkerParticleVelocities() {
    gfs_context *ctx = _particleAdvector.context();
    int err;
    gfs_set_field(ctx, GFS_FIELD_NEW, _MACVelocity.getRawArrayU(), _MACVelocity.getRawArrayV(), _MACVelocity.getRawArrayW(), &err);
    check(err, "gfs_set_field(new)");
    gfs_set_field(ctx, GFS_FIELD_SAVED, _savedVelocityField.getRawArrayU(), _savedVelocityField.getRawArrayV(),
                  _savedVelocityField.getRawArrayW(), &err);
    check(err, "gfs_set_field(saved)");
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
