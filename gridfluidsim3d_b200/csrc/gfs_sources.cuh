// gfs_sources.cuh -- inflow emission and outflow removal on the resident particles:
// FluidSimulation::_updateFluidSources (/root/reference/src/fluidsimulation.cpp:1838-1879), _updateInflowFluidSource
// (:1823-1836), _addNewFluidCells (:1749-1759), _getNewFluidParticles (:1771-1821), _addMarkerParticlesToCell (:1223-1247),
// _removeMarkerParticlesFromCells (:1717-1730).
//
// What the reference does for an active inflow source: (B) every AIR cell the source overlaps gets the 8 sub-cell particles
// of _addMarkerParticlesToCell and becomes fluid; (C) over the source's (grid-fitted) bounding box a half-dx occupancy grid
// is built from the particles INSIDE the box, and every empty sub-cell of a fluid-or-air source cell gets one particle at
// its centre.  Both carry the source's velocity and a jitter of +-0.25*0.1*dx per axis drawn from rand(); here the jitter
// is a counter-based hash of (seed, cell, sub-cell) -- which sub-cells emit is identical to the reference's, the positions
// agree to the jitter.  An outflow source removes the particles of the FLUID cells it overlaps.
#pragma once
#include "gfs_kernels.cuh"

namespace gfs {

struct Emitter {
    gfs_source_t src;
    int smin[3], smax[3];          // cells the shape may overlap (FluidSource::_getOverlappingCells index bounds)
    float bpos[3];                 // grid-fitted AABB of the source (Grid3d::fitAABBtoGrid): min corner ...
    double bext[3];                // ... and extents
    int bmin[3], bmax[3];          // its cell index bounds (Grid3d::getGridIndexBounds(AABB))
    float offset[3];               // GridIndexToPosition(bmin)
};

__device__ __forceinline__ bool aabb_inside(const Emitter &e, float x, float y, float z) {          // AABB::isPointInside, aabb.cpp:123-126
    return x >= e.bpos[0] && y >= e.bpos[1] && z >= e.bpos[2] && (double)x < __dadd_rn((double)e.bpos[0], e.bext[0]) &&
           (double)y < __dadd_rn((double)e.bpos[1], e.bext[1]) && (double)z < __dadd_rn((double)e.bpos[2], e.bext[2]);
}

// cell (i,j,k) is one of the source's overlapping cells: inside the shape's index bounds and, for a sphere, its centre
// inside the sphere (sphericalfluidsource.cpp:84-109; cuboid: every cell of the bounds, cuboidfluidsource.cpp:111-125)
__device__ __forceinline__ bool source_cell(const Grid &g, const Emitter &e, int i, int j, int k) {
    if (i < e.smin[0] || j < e.smin[1] || k < e.smin[2] || i > e.smax[0] || j > e.smax[1] || k > e.smax[2]) return false;
    if (e.src.kind != 0) return true;
    const double hw = 0.5 * g.dx;                              // GridIndexToCellCenter, grid3d.h:103-106
    const float cx = (float)__dadd_rn(__dmul_rn((double)(float)i, g.dx), hw), cy = (float)__dadd_rn(__dmul_rn((double)(float)j, g.dx), hw),
                cz = (float)__dadd_rn(__dmul_rn((double)(float)k, g.dx), hw);
    return source_contains(e.src, cx, cy, cz);
}

__device__ __forceinline__ float jitter(unsigned long long key, double jit) {
    const unsigned long long h = mix64(key);
    const double u = (double)(h >> 40) * (1.0 / 16777216.0);
    return (float)((2.0 * u - 1.0) * jit);
}

// occupancy of the half-dx sub-grid over the source's box: one bit per sub-cell, set by every particle inside the box
// (slots [base, base + min(n, *n_dev)): the resident particles, or the ones earlier sources of this call emitted)
__global__ void __launch_bounds__(256) k_source_mark(Grid g, Emitter e, int64_t base, int64_t n, const unsigned int *__restrict__ n_dev,
                                                     const float *__restrict__ x, const float *__restrict__ y,
                                                     const float *__restrict__ z, unsigned int *__restrict__ occ) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || (n_dev && r >= (int64_t)*n_dev)) return;
    r += base;
    const float px = x[r], py = y[r], pz = z[r];
    if (!aabb_inside(e, px, py, pz)) return;
    const double inv = 1.0 / (0.5 * g.dx);                    // positionToGridIndex(p - offset, 0.5 dx), :1813-1814
    const int si = (int)floor(__dmul_rn((double)__fsub_rn(px, e.offset[0]), inv)), sj = (int)floor(__dmul_rn((double)__fsub_rn(py, e.offset[1]), inv)),
              sk = (int)floor(__dmul_rn((double)__fsub_rn(pz, e.offset[2]), inv));
    const int w2 = 2 * (e.bmax[0] - e.bmin[0] + 1), h2 = 2 * (e.bmax[1] - e.bmin[1] + 1), d2 = 2 * (e.bmax[2] - e.bmin[2] + 1);
    if (si < 0 || sj < 0 || sk < 0 || si >= w2 || sj >= h2 || sk >= d2) return;          // Array3d::set ignores nothing: cannot happen inside the box
    const unsigned long long bit = (unsigned long long)si + (unsigned long long)w2 * ((unsigned long long)sj + (unsigned long long)h2 * (unsigned long long)sk);
    atomicOr(occ + (bit >> 5), 1u << (bit & 31));
}

struct EmitOut {
    float *x, *y, *z, *vx, *vy, *vz;
    int32_t *tag;
    unsigned int *count;          // particles appended so far (tickets)
    int64_t at, cap;              // first free slot, capacity in particles beyond `at`
};

__device__ __forceinline__ void emit_one(const Grid &g, const uint8_t *__restrict__ material, const Emitter &e, const EmitOut &o, float px, float py, float pz) {
    // _addMarkerParticle (:1253-1259): only inside the grid and outside solids
    const int i = cell_floor((double)px, g.invdx), j = cell_floor((double)py, g.invdx), k = cell_floor((double)pz, g.invdx);
    if (cell_solid_or_outside(g, material, i, j, k)) return;
    const unsigned int t = atomicAdd(o.count, 1u);
    if ((int64_t)t >= o.cap) return;
    const int64_t s = o.at + t;
    o.x[s] = px; o.y[s] = py; o.z[s] = pz;
    o.vx[s] = e.src.velocity[0]; o.vy[s] = e.src.velocity[1]; o.vz[s] = e.src.velocity[2];
    o.tag[s] = -1;
}

// one thread per cell of the source's box
__global__ void __launch_bounds__(128) k_source_emit(Grid g, Emitter e, uint8_t *__restrict__ material, unsigned int *__restrict__ occ, EmitOut o,
                                                     unsigned long long seed, double jit) {
    const int w = e.bmax[0] - e.bmin[0] + 1, h = e.bmax[1] - e.bmin[1] + 1, d = e.bmax[2] - e.bmin[2] + 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)w * h * d) return;
    const int li = (int)(t % w), lj = (int)((t / w) % h), lk = (int)(t / ((long long)w * h));
    const int i = e.bmin[0] + li, j = e.bmin[1] + lj, k = e.bmin[2] + lk;
    if (!source_cell(g, e, i, j, k)) return;
    const size_t cell = (size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)(k - g.k0));
    const uint8_t m = material[cell];
    if (m == GFS_SOLID) return;
    const unsigned long long ckey = seed ^ ((unsigned long long)cell << 8);
    const int w2 = 2 * w, h2 = 2 * h;
    const double inv = 1.0 / (0.5 * g.dx);
    if (m == GFS_AIR) {
        // (B) _addNewFluidCells -> _addMarkerParticlesToCell: 8 particles at c +- q with the source's velocity; the cell becomes fluid
        const double hw = 0.5 * g.dx, q = 0.25 * g.dx;
        const float cx = (float)__dadd_rn(__dmul_rn((double)(float)i, g.dx), hw), cy = (float)__dadd_rn(__dmul_rn((double)(float)j, g.dx), hw),
                    cz = (float)__dadd_rn(__dmul_rn((double)(float)k, g.dx), hw);
        const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, -1, -1, 1, 1, 1, 1}, sz[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
#pragma unroll
        for (int idx = 0; idx < 8; idx++) {
            const float px = __fadd_rn((float)(sx[idx] < 0 ? __dsub_rn((double)cx, q) : __dadd_rn((double)cx, q)), jitter(ckey ^ (unsigned long long)(16 + 3 * idx), jit));
            const float py = __fadd_rn((float)(sy[idx] < 0 ? __dsub_rn((double)cy, q) : __dadd_rn((double)cy, q)), jitter(ckey ^ (unsigned long long)(17 + 3 * idx), jit));
            const float pz = __fadd_rn((float)(sz[idx] < 0 ? __dsub_rn((double)cz, q) : __dadd_rn((double)cz, q)), jitter(ckey ^ (unsigned long long)(18 + 3 * idx), jit));
            emit_one(g, material, e, o, px, py, pz);          // (_addMarkerParticlesToCell pushes unconditionally; the cell is air, hence in range and not solid)
            if (aabb_inside(e, px, py, pz)) {                 // ... and counts for the occupancy of step (C)
                const int si = (int)floor(__dmul_rn((double)__fsub_rn(px, e.offset[0]), inv)), sj = (int)floor(__dmul_rn((double)__fsub_rn(py, e.offset[1]), inv)),
                          sk = (int)floor(__dmul_rn((double)__fsub_rn(pz, e.offset[2]), inv));
                if (si >= 0 && sj >= 0 && sk >= 0 && si < w2 && sj < h2 && sk < 2 * d) {
                    const unsigned long long bit = (unsigned long long)si + (unsigned long long)w2 * ((unsigned long long)sj + (unsigned long long)h2 * (unsigned long long)sk);
                    atomicOr(occ + (bit >> 5), 1u << (bit & 31));
                }
            }
        }
        material[cell] = GFS_FLUID;
    }
    // (C) _getNewFluidParticles: every empty sub-cell of this (fluid-or-air) source cell gets a particle at its centre.  The
    // sub-cells of one cell are only ever touched by this thread and by k_source_mark, which has finished.
#pragma unroll 1
    for (int idx = 0; idx < 8; idx++) {
        const int si = 2 * li + (idx & 1), sj = 2 * lj + ((idx >> 1) & 1), sk = 2 * lk + (idx >> 2);
        const unsigned long long bit = (unsigned long long)si + (unsigned long long)w2 * ((unsigned long long)sj + (unsigned long long)h2 * (unsigned long long)sk);
        if (occ[bit >> 5] & (1u << (bit & 31))) continue;
        const double hdx = 0.5 * g.dx, hw = 0.5 * hdx;         // GridIndexToCellCenter(i, j, k, 0.5 dx) + offset + jit (:1356-1357)
        const float px = __fadd_rn(__fadd_rn((float)__dadd_rn(__dmul_rn((double)(float)si, hdx), hw), e.offset[0]), jitter(ckey ^ (unsigned long long)(64 + 3 * idx), jit));
        const float py = __fadd_rn(__fadd_rn((float)__dadd_rn(__dmul_rn((double)(float)sj, hdx), hw), e.offset[1]), jitter(ckey ^ (unsigned long long)(65 + 3 * idx), jit));
        const float pz = __fadd_rn(__fadd_rn((float)__dadd_rn(__dmul_rn((double)(float)sk, hdx), hw), e.offset[2]), jitter(ckey ^ (unsigned long long)(66 + 3 * idx), jit));
        emit_one(g, material, e, o, px, py, pz);
        if (material[cell] != GFS_FLUID) material[cell] = GFS_FLUID;          // _addNewFluidParticles sets the particle's cell fluid
    }
}

// outflow: flag the FLUID cells each outflow source overlaps (FluidSource::getFluidCells), then compact the particles away
__global__ void __launch_bounds__(128) k_outflow_cells(Grid g, Emitter e, const uint8_t *__restrict__ material, uint8_t *__restrict__ removal) {
    const int w = e.smax[0] - e.smin[0] + 1, h = e.smax[1] - e.smin[1] + 1, d = e.smax[2] - e.smin[2] + 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w <= 0 || h <= 0 || d <= 0 || t >= (long long)w * h * d) return;
    const int i = e.smin[0] + (int)(t % w), j = e.smin[1] + (int)((t / w) % h), k = e.smin[2] + (int)(t / ((long long)w * h));
    if (!source_cell(g, e, i, j, k)) return;
    const size_t cell = (size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)(k - g.k0));
    if (material[cell] == GFS_FLUID) removal[cell] = 1;
}

// keep the particles whose cell is not flagged (block-aggregated tickets: one atomic per block); order of the kept ones preserved per block
__global__ void __launch_bounds__(256) k_remove_in_cells(Grid g, int64_t n, const uint8_t *__restrict__ removal,
                                                         const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                                                         const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz, const int32_t *__restrict__ tag,
                                                         float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
                                                         float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz, int32_t *__restrict__ otag,
                                                         unsigned int *__restrict__ kept) {
    __shared__ unsigned int s_warp[8], s_base;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool keep = false;
    if (r < n) {
        const int i = cell_floor((double)x[r], g.invdx), j = cell_floor((double)y[r], g.invdx), k = cell_floor((double)z[r], g.invdx);
        keep = true;
        if (i >= 0 && j >= 0 && k >= g.k0 && i < g.I && j < g.J && k < g.k1)
            keep = removal[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)(k - g.k0))] == 0;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, keep);
    const unsigned int before = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tot = 0;
        for (int w = 0; w < 8; w++) { const unsigned int c = s_warp[w]; s_warp[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(kept, tot) : 0u;
    }
    __syncthreads();
    if (!keep) return;
    const unsigned int s = s_base + s_warp[warp] + before;
    ox[s] = x[r]; oy[s] = y[r]; oz[s] = z[r]; ovx[s] = vx[r]; ovy[s] = vy[r]; ovz[s] = vz[r]; otag[s] = tag[r];
}

}  // namespace gfs
