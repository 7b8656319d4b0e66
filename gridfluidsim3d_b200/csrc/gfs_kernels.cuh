// gfs_kernels.cuh -- the sm_100a kernels of the particle<->grid transfer path (round-1 set).
//
//   K0  k_hist + exclusive scan + k_scatter_sorted (counting sort; stable variant: cub radix sort + k_reorder)
//   K1  k_classify, k_p2g_scatter (fast, order-independent fixed point) or k_p2g_gather (exact,
//       reference summation order), k_p2g_finalize, k_assemble   P2G + classification
//   K2  k_g2p_advect                                            PIC/FLIP + RK1..4 + solid test
//   plus the host-pointer operators k_sample / k_advect / k_splat_points
#pragma once
#include <cuda.h>
#include "gfs_device.cuh"
#include "../../include/gfs_b200.h"

namespace gfs {

// ------------------------------------------------------------------------------------------------
// Fixed-point accumulation (fast P2G).  Weights and weight*value products are converted to 64-bit
// fixed point and added with integer atomics: integer addition is associative, so every node sum is
// independent of particle order, brick decomposition and GPU count -- bit-reproducible without
// ordering constraints and without float atomics.
//   weight:  S_w = 2^44            (sum of weights < 2^18)
//   num:     S_n = 2^(44 - vexp)   where 2^vexp >= max |velocity component|  (|sum| < 2^(18+vexp))
// One contribution is < 2^44 in magnitude, so it splits exactly into a signed high word (>> 22) and an
// unsigned 22-bit low word; the brick-tile kernel accumulates the two words separately with native
// 32-bit shared-memory atomics (exact for up to 511 contributions per node per CTA, guaranteed by sending
// bricks that hold a cell with more than 63 particles down the 64-bit path) and the 64-bit value is
// rebuilt as (sum_hi << 22) + sum_lo.  Global accumulators are plain 64-bit integer atomics.
// Resolution 2^-44 = 5.7e-14: a node whose total weight is 1e-8 still carries 5.7e-6 relative precision
// (fp32 accumulation, which the reference uses, carries 6e-8 relative but depends on particle order).
// ------------------------------------------------------------------------------------------------
constexpr int kWeightFracBits = 44;
constexpr int kLoBits = 22;
constexpr float kWeightScaleF = 17592186044416.0f;       // 2^44
constexpr double kWeightScaleD = 17592186044416.0;

struct SplatParams {
    double radius, rsq;           // ScalarField::setPointRadius (scalarfield.cpp:40-46)
    double c1, c2, c3;            // (4/9)/r^6, (17/9)/r^4, (22/9)/r^2
    float  c1f, c2f, c3f;
    float  rsqf;                  // float r such that (d2 < rsqf) == ((double)d2 < rsq) for every float d2
    float  inv_rsqf;              // 1/R^2 (exact when R is a power of two)
    double inv_rsq;
    const unsigned int *vmax_bits;   // device word: float bits of max |velocity component| (k_keys / caller)
};

// 2^vexp >= max|v|, clamped to [2^-24, 2^40]; the numerator scale is 2^(48 - vexp)
__device__ __forceinline__ int num_exponent(const SplatParams &sp) {
    int e = (int)((__ldg(sp.vmax_bits) >> 23) & 0xffu) - 127 + 1;
    return max(-24, min(40, e));
}
__device__ __forceinline__ float num_scale_f(int vexp) { return __uint_as_float((unsigned)(127 + kWeightFracBits - vexp) << 23); }
__device__ __forceinline__ double inv_num_scale_d(int vexp) {
    return __longlong_as_double((long long)(1023 - kWeightFracBits + vexp) << 52);
}

struct Sources {
    int n;
    gfs_source_t s[8];
};

// d^2 in float exactly as the reference forms it (vmath::dot, vmath.h:71-73; no contraction)
__device__ __forceinline__ float dist2(float vx, float vy, float vz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
}

// ScalarField::_evaluateTricubicFieldFunctionForRadiusSquared (scalarfield.cpp:569-571)
__device__ __forceinline__ double kernel_weight_exact(const SplatParams &sp, double d) {
    double a = __dmul_rn(__dmul_rn(__dmul_rn(sp.c1, d), d), d);
    double b = __dmul_rn(__dmul_rn(sp.c2, d), d);
    double c = __dmul_rn(sp.c3, d);
    return __dsub_rn(__dadd_rn(__dsub_rn(1.0, a), b), c);
}
// Same polynomial with u = 1 - d^2/R^2:  1 - (4/9)s^3 + (17/9)s^2 - (22/9)s  ==  u^2 (5/9 + (4/9) u).  The
// reference's form cancels three O(1) terms down to ~0 near the rim of the support (harmless in its fp64, a
// 3e-7 ABSOLUTE error in fp32, i.e. tens of percent of a rim weight); this form has no cancellation, so the
// fp32 weight is relatively accurate everywhere.  R^2 - d^2 is formed in fp64 unless R^2 is exact in fp32.
template <bool POW2>
__device__ __forceinline__ float kernel_weight_fast(const SplatParams &sp, float d) {
    float u = POW2 ? __fmul_rn(__fsub_rn(sp.rsqf, d), sp.inv_rsqf)
                   : (float)__dmul_rn(__dsub_rn(sp.rsq, (double)d), sp.inv_rsq);
    return __fmul_rn(__fmul_rn(u, u), fmaf(u, 0.44444444444444444f, 0.55555555555555556f));
}

// node coordinate (float)((float)i*dx)  (Grid3d::GridIndexToPosition(int,int,int,dx), grid3d.h:73-75)
__device__ __forceinline__ float node_pos(int i, double dx) { return (float)__dmul_rn((double)(float)i, dx); }

// ------------------------------------------------------------------------------------------------
// K0: keys
// ------------------------------------------------------------------------------------------------
// block-wide max of a non-negative float into *vmax_bits (float bits order like unsigned ints)
__device__ __forceinline__ void block_vmax(float m, unsigned int *__restrict__ vmax_bits) {
    __shared__ float s_m[32];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x + 31) / 32 ? s_m[threadIdx.x] : 0.0f;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0 && m > __uint_as_float(*(volatile unsigned int *)vmax_bits)) atomicMax(vmax_bits, __float_as_uint(m));
    }
}

// key of a position: brick-major id of its cell, or nkeys (the overflow bin, sorts last) when the cell is
// outside the grid / outside this slab's stored layers.  fp64 index arithmetic: bit-exact.
__device__ __forceinline__ uint32_t position_key(const Grid &g, uint32_t nkeys, float px, float py, float pz) {
    int i = cell_floor((double)px, g.invdx), j = cell_floor((double)py, g.invdx), k = cell_floor((double)pz, g.invdx);
    if (i >= 0 && j >= 0 && i < g.I && j < g.J && k >= g.k0 && k < g.k1 && k >= 0 && k < g.K) return brick_key(g, i, j, k - g.k0);
    return nkeys;
}

// z-slab runs keep a cell table only for the bricks around the owned layers [key_lo, key_hi): a key outside that range
// (a particle more than one hop away from its owner: never with CFL-limited steps) goes to the out-of-grid bin, where the
// global-path CTA still advects and forwards it.
struct KeyRange { uint32_t lo, hi; };
__device__ __forceinline__ uint32_t clamp_key(uint32_t key, uint32_t nkeys, KeyRange kr) {
    return (key < nkeys && (key < kr.lo || key >= kr.hi)) ? nkeys : key;
}

// Ticket of a particle in its cell; with a per-cell cap (FluidSimulation::_removeMarkerParticles,
// fluidsimulation.cpp:3221-3243: at most _maxMarkerParticlesPerCell particles survive per cell, WHICH ones is decided
// by the reference's rand() shuffle and here by the ticket order -- equally arbitrary) a particle whose ticket is past
// the cap is re-binned into the dead bin nkeys + 1.  The cell counter is left over-incremented; k_clamp_counts trims
// it before the scan.
__device__ __forceinline__ uint32_t take_ticket(uint32_t *__restrict__ counts, uint32_t nkeys, uint32_t cap, uint32_t &key) {
    uint32_t t = atomicAdd(counts + key, 1u);
    if (cap && key < nkeys && t >= cap) { key = nkeys + 1; t = atomicAdd(counts + key, 1u); }
    return t;
}

__global__ void __launch_bounds__(256) k_clamp_counts(uint32_t nkeys, uint32_t cap, uint32_t *__restrict__ counts) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nkeys && counts[t] > cap) counts[t] = cap;
}

// K0a.  One thread per particle: key, rank inside its cell (atomic ticket on the cell counter -- the count
// is deterministic, the ticket order is not and does not need to be: the fast P2G is order-independent),
// identity permutation for the stable path, and max |velocity|.  solid != null: a particle inside a solid cell goes to
// the dead bin (FluidSimulation::_removeMarkerParticlesInSolidCells, fluidsimulation.cpp:1933-1957).
__global__ void __launch_bounds__(256) k_hist(Grid g, uint32_t nkeys, const float *__restrict__ x, const float *__restrict__ y,
                       const float *__restrict__ z, const float *__restrict__ vx, const float *__restrict__ vy,
                       const float *__restrict__ vz, int64_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ rank,
                       int32_t *__restrict__ perm, uint32_t *__restrict__ counts, unsigned int *__restrict__ vmax_bits,
                       const uint8_t *__restrict__ solid, uint32_t cap, KeyRange kr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (r < n) {
        uint32_t key = clamp_key(position_key(g, nkeys, x[r], y[r], z[r]), nkeys, kr);
        if (solid && key < nkeys) {
            const int i = cell_floor((double)x[r], g.invdx), j = cell_floor((double)y[r], g.invdx), k = cell_floor((double)z[r], g.invdx);
            if (solid[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)(k - g.k0))] == GFS_SOLID) key = nkeys + 1;
        }
        rank[r] = take_ticket(counts, nkeys, cap, key);
        keys[r] = key;
        perm[r] = (int32_t)r;
        m = fmaxf(fabsf(vx[r]), fmaxf(fabsf(vy[r]), fabsf(vz[r])));
        if (!(m < 3.0e38f)) m = 0.0f;          // NaN/Inf velocities do not steer the scale
    }
    block_vmax(m, vmax_bits);
}

// K0c (counting sort).  sorted slot = cell_start[key] + rank.
__global__ void __launch_bounds__(256) k_scatter_sorted(int64_t n, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rank,
                                 const int32_t *__restrict__ cell_start,
                                 const float *__restrict__ sx, const float *__restrict__ sy, const float *__restrict__ sz,
                                 const float *__restrict__ svx, const float *__restrict__ svy, const float *__restrict__ svz,
                                 const int32_t *__restrict__ stag,
                                 float *__restrict__ dx_, float *__restrict__ dy, float *__restrict__ dz,
                                 float *__restrict__ dvx, float *__restrict__ dvy, float *__restrict__ dvz,
                                 int32_t *__restrict__ dtag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int64_t d = (int64_t)cell_start[keys[r]] + rank[r];
    dx_[d] = sx[r]; dy[d] = sy[r]; dz[d] = sz[r];
    dvx[d] = svx[r]; dvy[d] = svy[r]; dvz[d] = svz[r];
    dtag[d] = stag[r];
}

// _fluidCellIndices (fluidsimulation.cpp:2019-2039): the fluid cells in k, j, i scan order.  flag pass for the scan, then
// the compaction writes (i, j, k) triples at the scanned offsets -- ascending linear index = the reference's order.
__global__ void __launch_bounds__(256) k_fluid_flags(const uint8_t *__restrict__ material, long long cells, uint32_t *__restrict__ flags) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells) flags[c] = material[c] == GFS_FLUID ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_fluid_cells(const uint8_t *__restrict__ material, const uint32_t *__restrict__ offset, long long cells,
                                                     int I, int J, long long capacity, int *__restrict__ ijk) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells || material[c] != GFS_FLUID) return;
    const long long o = offset[c];
    if (o >= capacity) return;
    ijk[3 * o] = (int)(c % I); ijk[3 * o + 1] = (int)((c / I) % J); ijk[3 * o + 2] = (int)(c / ((long long)I * J));
}

// K0c, first sort after an upload: the uploaded AoS records (24 contiguous bytes each) are still on the device, so the
// sorted SoA arrays are GATHERED from them through the index -- one or two 32-byte sectors read per particle, fully
// coalesced writes -- instead of scattering seven 4-byte streams to random slots (k_scatter_sorted: a 32-byte sector
// touched per 4 bytes written; 25 ms against 1 ms at 96 M shuffled particles).  tag = upload index = the source slot.
__global__ void __launch_bounds__(256) k_gather_sorted_aos(int64_t n, const int32_t *__restrict__ index, const float2 *__restrict__ aos,
                                 float *__restrict__ dx_, float *__restrict__ dy, float *__restrict__ dz,
                                 float *__restrict__ dvx, float *__restrict__ dvy, float *__restrict__ dvz, int32_t *__restrict__ dtag) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const int32_t r = index[d];
    const float2 *p = aos + 3 * (int64_t)r;
    const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    dx_[d] = a.x; dy[d] = a.y; dz[d] = b.x; dvx[d] = b.y; dvy[d] = c.x; dvz[d] = c.y;
    dtag[d] = r;
}

// K0c (lazy variant used inside the fused substep): only the permutation is materialised, index[sorted slot] = current
// slot; P2G and G2P then fetch their particles through it, and G2P -- which rewrites every particle anyway -- stores
// its results at the sorted slot.  Saves moving 56 bytes per particle once per substep.
__global__ void __launch_bounds__(256) k_build_index(int64_t n, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rank,
                                                     const int32_t *__restrict__ cell_start, int32_t *__restrict__ index) {
    // four particles per thread, loads first: the key -> cell_start -> store chain is pure latency
    const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
    uint32_t k[4], q[4];
    int32_t cs[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int64_t r = r0 + u * 256;
        if (r < n) { k[u] = keys[r]; q[u] = rank[r]; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) if (r0 + u * 256 < n) cs[u] = cell_start[k[u]];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int64_t r = r0 + u * 256;
        if (r < n) index[(int64_t)cs[u] + q[u]] = (int32_t)r;
    }
}

// K0c (stable path).  sorted SoA <- unsorted SoA through the radix-sorted permutation
__global__ void k_reorder(int64_t n, const int32_t *__restrict__ perm,
                          const float *__restrict__ sx, const float *__restrict__ sy, const float *__restrict__ sz,
                          const float *__restrict__ svx, const float *__restrict__ svy, const float *__restrict__ svz,
                          const int32_t *__restrict__ stag,
                          float *__restrict__ dx_, float *__restrict__ dy, float *__restrict__ dz,
                          float *__restrict__ dvx, float *__restrict__ dvy, float *__restrict__ dvz,
                          int32_t *__restrict__ dtag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int32_t s = perm[r];
    dx_[r] = sx[s]; dy[r] = sy[s]; dz[r] = sz[s];
    dvx[r] = svx[s]; dvy[r] = svy[s]; dvz[r] = svz[s];
    dtag[r] = stag[s];
}

// ------------------------------------------------------------------------------------------------
// K1a: classification  (FluidSimulation::_updateFluidCells marking loop, fluidsimulation.cpp:1998-2017)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_classify(Grid g, const int32_t *__restrict__ cell_start, uint8_t *__restrict__ material,
                           unsigned long long *__restrict__ counters /* [0]=in_solid particles, [1]=fluid cells */,
                           long long first_cell, long long ncells, int own_k0, int own_k1) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool counted = false;
    if (t < ncells) {
        const size_t idx = (size_t)(first_cell + t);
        const uint32_t idx32 = (uint32_t)idx;                       // cell counts fit 31 bits (checked at domain_init)
        const int i = (int)(idx32 % (uint32_t)g.I), j = (int)((idx32 / (uint32_t)g.I) % (uint32_t)g.J), kl = (int)(idx32 / ((uint32_t)g.I * (uint32_t)g.J));
        const int k = kl + g.k0;
        const uint32_t key = brick_key(g, i, j, kl);
        const int cnt = cell_start[key + 1] - cell_start[key];
        const uint8_t m0 = material[idx];
        uint8_t m = m0;
        const bool interior = i >= 1 && i < g.I - 1 && j >= 1 && j < g.J - 1 && k >= 1 && k < g.K - 1;
        if (interior && m == GFS_FLUID) m = GFS_AIR;
        if (cnt > 0) {
            if (m == GFS_SOLID) atomicAdd(&counters[0], (unsigned long long)cnt);      // rare: particles inside a solid
            else m = GFS_FLUID;
        }
        if (m != m0) material[idx] = m;
        counted = m == GFS_FLUID && k >= own_k0 && k < own_k1;
    }
    // one atomic per block (a ticket per fluid cell on one address serialises ~10^7 atomics)
    const int nfluid = __syncthreads_count(counted);
    if (threadIdx.x == 0 && nfluid > 0) atomicAdd(&counters[1], (unsigned long long)nfluid);
}

// ------------------------------------------------------------------------------------------------
// K1b (fast): particle-centric splat of u,v,w into 64-bit fixed-point accumulators.
// acc layout: for comp c, node n: acc[c][2*n] = sum w*v, acc[c][2*n+1] = sum w  (int64, global).
// Candidate nodes per axis are the contiguous run of {c-1,c,c+1} whose 1-D distance already passes
// d^2 < R^2 (a necessary condition: the float sum of non-negative squares is monotone); the full float
// distance test of scalarfield.cpp:186-189 then decides, as in the reference.  (The reference also clips
// to Grid3d::getGridIndexBounds; that only ever removes pairs at distance R +- 1 ulp whose weight is
// ~1e-14 -- the exact-mode gather below applies the clip literally.)
// ------------------------------------------------------------------------------------------------
// tile: shared-memory accumulators of one brick (nullptr = accumulate straight into global memory).
// Layout tile[word][node], word = 0 num_hi, 1 num_lo, 2 wt_hi, 3 wt_lo, node = li + 10*(lj + 10*lk) with
// l = global node index - (tile origin ti,tj,tk).
constexpr int kTileEdge = 10;
constexpr int kTileNodes = kTileEdge * kTileEdge * kTileEdge;

template <int ARITH>
__device__ __forceinline__ void splat_component(const Grid &g, const SplatParams &sp, int comp, float px, float py, float pz,
                                                float value, int ni, int nj, int nkl, int koff,
                                                float off, float num_scale, unsigned long long *__restrict__ acc,
                                                uint32_t *tile = nullptr, int ti = 0, int tj = 0, int tk = 0) {
    // p -= offset (scalarfield.cpp:168); offset is (0,.5,.5)dx / (.5,0,.5)dx / (.5,.5,0)dx narrowed to float
    float q[3] = {comp == 0 ? px : __fsub_rn(px, off), comp == 1 ? py : __fsub_rn(py, off), comp == 2 ? pz : __fsub_rn(pz, off)};
    int   lo[3], hi[3], base[3];
    float d1[3][3];
    const int nmax[3] = {ni - 1, nj - 1, nkl - 1 + koff};
    const int nmin[3] = {0, 0, koff};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int c = cell_floor((double)q[a], g.invdx);
        int l = 2, h = -1;
#pragma unroll
        for (int s = 0; s < 3; s++) {
            float d = __fsub_rn(node_pos(c - 1 + s, g.dx), q[a]);
            d1[a][s] = d;
            if ((double)__fmul_rn(d, d) < sp.rsq) { l = min(l, s); h = max(h, s); }
        }
        base[a] = c - 1;
        lo[a] = max(c - 1 + l, nmin[a]) - base[a];      // slot range in {0,1,2}, clamped to the node grid
        hi[a] = min(c - 1 + h, nmax[a]) - base[a];
    }
    const int bi = base[0], bj = base[1], bk = base[2];
    for (int sk = lo[2]; sk <= hi[2]; sk++)
        for (int sj = lo[1]; sj <= hi[1]; sj++)
            for (int si = lo[0]; si <= hi[0]; si++) {
                float vx = si == 0 ? d1[0][0] : (si == 1 ? d1[0][1] : d1[0][2]);
                float vy = sj == 0 ? d1[1][0] : (sj == 1 ? d1[1][1] : d1[1][2]);
                float vz = sk == 0 ? d1[2][0] : (sk == 1 ? d1[2][1] : d1[2][2]);
                float d2 = dist2(vx, vy, vz);
                if ((double)d2 < sp.rsq) {
                    size_t node = (size_t)(bi + si) + (size_t)ni * ((size_t)(bj + sj) + (size_t)nj * (size_t)(bk + sk - koff));
                    long long wn, ww;
                    if (ARITH == 1) {
                        double w = kernel_weight_exact(sp, (double)d2);
                        ww = __double2ll_rn(w * kWeightScaleD);                               // 2^48
                        wn = __double2ll_rn(__dmul_rn(w, (double)value) * (double)num_scale);
                    } else {
                        float w = kernel_weight_fast<false>(sp, d2);
                        ww = __float2ll_rn(w * kWeightScaleF);
                        wn = __float2ll_rn((w * value) * num_scale);
                    }
                    int li = bi + si - ti, lj = bj + sj - tj, lk = bk + sk - tk;
                    if (tile && (unsigned)li < (unsigned)kTileEdge && (unsigned)lj < (unsigned)kTileEdge && (unsigned)lk < (unsigned)kTileEdge) {
                        int t = li + kTileEdge * (lj + kTileEdge * lk);
                        atomicAdd(tile + t, (uint32_t)(wn >> kLoBits));
                        atomicAdd(tile + kTileNodes + t, (uint32_t)wn & ((1u << kLoBits) - 1u));
                        atomicAdd(tile + 2 * kTileNodes + t, (uint32_t)(ww >> kLoBits));
                        atomicAdd(tile + 3 * kTileNodes + t, (uint32_t)ww & ((1u << kLoBits) - 1u));
                    } else {
                        atomicAdd(acc + 2 * node, (unsigned long long)wn);
                        atomicAdd(acc + 2 * node + 1, (unsigned long long)ww);
                    }
                }
            }
}

// ---- power-of-two dx: all index arithmetic in fp32, exactly --------------------------------------------
// For dx = 2^-k the offset-space coordinate q, its cell c = floor(q/dx), t = q - c*dx and dx - t are all exact
// in fp32, node positions (float)(i*dx) are exact, and the only nodes that can pass d^2 < R^2 = dx^2 are c and
// c+1 on every axis: 8 candidates per component instead of the reference's 27, same hits.
struct AxisCand { int c; float e0, e1; };      // base node, squared 1-D distances to nodes c and c+1 (+inf: no such node)

__device__ __forceinline__ AxisCand axis_cand(float q, const Grid &g, int nmin, int nmax) {
    AxisCand r;
    float fl = floorf(__fmul_rn(q, g.invdxf));
    r.c = (int)fl;
    float t = __fmaf_rn(-fl, g.dxf, q);          // q - c*dx, exact
    float u = __fsub_rn(g.dxf, t);               // (c+1)*dx - q, exact
    r.e0 = (r.c >= nmin && r.c <= nmax) ? __fmul_rn(t, t) : __int_as_float(0x7f800000);
    r.e1 = (r.c + 1 >= nmin && r.c + 1 <= nmax) ? __fmul_rn(u, u) : __int_as_float(0x7f800000);
    return r;
}

template <bool TILE>
__device__ __forceinline__ void splat_comp_pow2(const SplatParams &sp, const AxisCand &X, const AxisCand &Y, const AxisCand &Z,
                                                float value, float num_scale, int ni, int nj, int koff,
                                                unsigned long long *__restrict__ acc, uint32_t *tile, int ti, int tj, int tk) {
    const int tbase = (X.c - ti) + kTileEdge * ((Y.c - tj) + kTileEdge * (Z.c - tk));
    const long long nbase = (long long)X.c + (long long)ni * ((long long)Y.c + (long long)nj * (long long)(Z.c - koff));   // only used for valid nodes
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int a = 0; a < 2; a++) {
                float d2 = __fadd_rn(__fadd_rn(a ? X.e1 : X.e0, b ? Y.e1 : Y.e0), c ? Z.e1 : Z.e0);
                if (d2 < sp.rsqf) {
                    float w = kernel_weight_fast<true>(sp, d2);
                    long long ww = __float2ll_rn(w * kWeightScaleF);
                    long long wn = __float2ll_rn((w * value) * num_scale);
                    if (TILE) {
                        int t = tbase + a + kTileEdge * (b + kTileEdge * c);
                        atomicAdd(tile + t, (uint32_t)(wn >> kLoBits));
                        atomicAdd(tile + kTileNodes + t, (uint32_t)wn & ((1u << kLoBits) - 1u));
                        atomicAdd(tile + 2 * kTileNodes + t, (uint32_t)(ww >> kLoBits));
                        atomicAdd(tile + 3 * kTileNodes + t, (uint32_t)ww & ((1u << kLoBits) - 1u));
                    } else {
                        long long node = nbase + a + (long long)ni * (b + (long long)nj * c);
                        atomicAdd(acc + 2 * node, (unsigned long long)wn);
                        atomicAdd(acc + 2 * node + 1, (unsigned long long)ww);
                    }
                }
            }
}

template <bool TILE>
__device__ __forceinline__ void splat_pow2(const Grid &g, const SplatParams &sp, float ns, float px, float py, float pz,
                                           float vx, float vy, float vz, uint32_t *tile, int ti, int tj, int tk,
                                           unsigned long long *__restrict__ accu, unsigned long long *__restrict__ accv,
                                           unsigned long long *__restrict__ accw) {
    const int kl = g.k1 - g.k0;
    // the staggered axis of each component is unshifted (offset 0), the other two are shifted by 0.5dx
    const AxisCand ux = axis_cand(px, g, 0, g.I), uy = axis_cand(py, g, 0, g.J), uz = axis_cand(pz, g, g.k0, g.k0 + kl);
    const AxisCand sx = axis_cand(__fsub_rn(px, g.halfdxf), g, 0, g.I - 1), sy = axis_cand(__fsub_rn(py, g.halfdxf), g, 0, g.J - 1),
                   sz = axis_cand(__fsub_rn(pz, g.halfdxf), g, g.k0, g.k0 + kl - 1);
    splat_comp_pow2<TILE>(sp, ux, sy, sz, vx, ns, g.I + 1, g.J, g.k0, accu, tile, ti, tj, tk);
    splat_comp_pow2<TILE>(sp, sx, uy, sz, vy, ns, g.I, g.J + 1, g.k0, accv, tile + 4 * kTileNodes, ti, tj, tk);
    splat_comp_pow2<TILE>(sp, sx, sy, uz, vz, ns, g.I, g.J, g.k0, accw, tile + 8 * kTileNodes, ti, tj, tk);
}

// K1b (fast, brick tiles).  One CTA per brick of 8^3 cells: the brick's particles are one contiguous run of
// the sorted arrays (coalesced loads, each particle read once); their contributions land in a 10^3-node
// shared-memory tile per component through native 32-bit integer atomics (hi/lo words, see above); the tile is
// then flushed to the global 64-bit accumulators with one integer atomic per touched value.  Because every
// add is an integer add, the result is bit-identical to k_p2g_scatter's for any particle order.
template <int ARITH>
__global__ void __launch_bounds__(256) k_p2g_tile(Grid g, SplatParams sp, const int32_t *__restrict__ cell_start, uint32_t brick0,
                              const int32_t *__restrict__ index /* nullable: sorted slot -> storage slot */,
                              const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                              const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                              unsigned long long *__restrict__ accu, unsigned long long *__restrict__ accv,
                              unsigned long long *__restrict__ accw) {
    extern __shared__ uint32_t tile[];                    // [3 comps][4 words][1000 nodes] = 48 000 B
    const uint32_t b = blockIdx.x + brick0;
    const int start = cell_start[(size_t)b * kBrickCells], end = cell_start[(size_t)(b + 1) * kBrickCells];
    if (start == end) return;
    // a cell with more than 63 particles could overflow the 32-bit words (511 contributions per node): such
    // a brick accumulates straight into the 64-bit global accumulators instead
    int dense = 0;
    for (int c = threadIdx.x; c < kBrickCells; c += blockDim.x)
        dense |= (cell_start[(size_t)b * kBrickCells + c + 1] - cell_start[(size_t)b * kBrickCells + c]) > 63;
    for (int t = threadIdx.x; t < 12 * kTileNodes; t += blockDim.x) tile[t] = 0u;
    dense = __syncthreads_or(dense);
    const int bi = (int)(b % (uint32_t)g.nbi), bj = (int)((b / (uint32_t)g.nbi) % (uint32_t)g.nbj), bk = (int)(b / ((uint32_t)g.nbi * (uint32_t)g.nbj));
    const int ti = bi * kBrick - 1, tj = bj * kBrick - 1, tk = bk * kBrick - 1 + g.k0;
    const float off = (float)g.halfdx;
    const int kl = g.k1 - g.k0;
    const float ns = num_scale_f(num_exponent(sp));
    uint32_t *t0 = dense ? nullptr : tile, *t1 = dense ? nullptr : tile + 4 * kTileNodes, *t2 = dense ? nullptr : tile + 8 * kTileNodes;
    if (ARITH == 2) {
        if (dense) for (int q = start + threadIdx.x; q < end; q += blockDim.x) {
            const int r = index ? index[q] : q;
            splat_pow2<false>(g, sp, ns, x[r], y[r], z[r], vx[r], vy[r], vz[r], tile, ti, tj, tk, accu, accv, accw);
        }
        else {
            // software pipeline: the next particle is in flight while this one's ~25 shared atomics are issued
            int q = start + threadIdx.x;
            float nx = 0.f, ny = 0.f, nz = 0.f, nvx = 0.f, nvy = 0.f, nvz = 0.f;
            if (q < end) {
                const int r = index ? index[q] : q;
                nx = x[r]; ny = y[r]; nz = z[r]; nvx = vx[r]; nvy = vy[r]; nvz = vz[r];
            }
            for (; q < end; q += blockDim.x) {
                const float px = nx, py = ny, pz = nz, ux = nvx, uy = nvy, uz = nvz;
                const int qn = q + blockDim.x;
                if (qn < end) {
                    const int r = index ? index[qn] : qn;
                    nx = x[r]; ny = y[r]; nz = z[r]; nvx = vx[r]; nvy = vy[r]; nvz = vz[r];
                }
                splat_pow2<true>(g, sp, ns, px, py, pz, ux, uy, uz, tile, ti, tj, tk, accu, accv, accw);
            }
        }
    } else {
        for (int q = start + threadIdx.x; q < end; q += blockDim.x) {
            const int r = index ? index[q] : q;
            float px = x[r], py = y[r], pz = z[r];
            splat_component<0>(g, sp, 0, px, py, pz, vx[r], g.I + 1, g.J, kl, g.k0, off, ns, accu, t0, ti, tj, tk);
            splat_component<0>(g, sp, 1, px, py, pz, vy[r], g.I, g.J + 1, kl, g.k0, off, ns, accv, t1, ti, tj, tk);
            splat_component<0>(g, sp, 2, px, py, pz, vz[r], g.I, g.J, kl + 1, g.k0, off, ns, accw, t2, ti, tj, tk);
        }
    }
    if (dense) return;
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * kTileNodes; t += blockDim.x) {
        const int comp = t / kTileNodes, nloc = t - comp * kTileNodes;
        const uint32_t *tw = tile + comp * 4 * kTileNodes + nloc;
        long long wn = ((long long)(int32_t)tw[0] << kLoBits) + (long long)tw[kTileNodes];
        long long ww = ((long long)(int32_t)tw[2 * kTileNodes] << kLoBits) + (long long)tw[3 * kTileNodes];
        if (wn == 0 && ww == 0) continue;
        const int li = nloc % kTileEdge, lj = (nloc / kTileEdge) % kTileEdge, lk = nloc / (kTileEdge * kTileEdge);
        const int ni = g.I + (comp == 0), nj = g.J + (comp == 1);
        const size_t node = (size_t)(ti + li) + (size_t)ni * ((size_t)(tj + lj) + (size_t)nj * (size_t)(tk + lk - g.k0));
        unsigned long long *acc = comp == 0 ? accu : (comp == 1 ? accv : accw);
        if (wn) atomicAdd(acc + 2 * node, (unsigned long long)wn);
        if (ww) atomicAdd(acc + 2 * node + 1, (unsigned long long)ww);
    }
}

template <int ARITH>
__global__ void __launch_bounds__(256) k_p2g_scatter(Grid g, SplatParams sp, const int32_t *__restrict__ n_valid,
                              const int32_t *__restrict__ index,
                              const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                              const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                              unsigned long long *__restrict__ accu, unsigned long long *__restrict__ accv,
                              unsigned long long *__restrict__ accw) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)__ldg(n_valid)) return;
    if (index) r = index[r];
    float px = x[r], py = y[r], pz = z[r];
    float off = (float)g.halfdx;
    int kl = g.k1 - g.k0;
    float ns = num_scale_f(num_exponent(sp));
    if (ARITH == 2) { splat_pow2<false>(g, sp, ns, px, py, pz, vx[r], vy[r], vz[r], nullptr, 0, 0, 0, accu, accv, accw); return; }
    splat_component<0>(g, sp, 0, px, py, pz, vx[r], g.I + 1, g.J, kl, g.k0, off, ns, accu);
    splat_component<0>(g, sp, 1, px, py, pz, vy[r], g.I, g.J + 1, kl, g.k0, off, ns, accv);
    splat_component<0>(g, sp, 2, px, py, pz, vz[r], g.I, g.J, kl + 1, g.k0, off, ns, accw);
}

// source->containsPoint(face position)  (fluidsimulation.cpp:2489-2524)
__device__ __forceinline__ bool source_contains(const gfs_source_t &s, float fx, float fy, float fz) {
    if (s.kind == 0) {
        float vx = __fsub_rn(fx, s.p[0]), vy = __fsub_rn(fy, s.p[1]), vz = __fsub_rn(fz, s.p[2]);
        return (double)dist2(vx, vy, vz) < __dmul_rn(s.a, s.a);
    }
    return fx >= s.p[0] && fy >= s.p[1] && fz >= s.p[2] &&
           (double)fx < __dadd_rn((double)s.p[0], s.a) && (double)fy < __dadd_rn((double)s.p[1], s.b) &&
           (double)fz < __dadd_rn((double)s.p[2], s.c);
}

// normalise (ScalarField::applyWeightField, scalarfield.cpp:90-106), isValueSet = weight > 1e-9 and the
// inflow override on set faces (fluidsimulation.cpp:2571-2594).  val holds the node grid "ugrid",
// setmask its isValueSet.  Also clears the accumulators for the next substep.  Flat over the nodes of all
// three components (one 16-byte load per node); node coordinates are only recovered when sources exist.
struct FinalizeArgs {
    unsigned long long *acc[3];
    float *val[3];
    uint8_t *setmask[3];
    long long first[3];           // first node of the processed layer range, per component
    long long count[3];           // number of nodes processed, per component
};

__global__ void __launch_bounds__(256) k_p2g_finalize(Grid g, SplatParams sp, Sources src, FinalizeArgs fa) {
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c0 = fa.count[0], c01 = fa.count[0] + fa.count[1], c012 = c01 + fa.count[2];
    if (t0 >= c012) return;
    const int comp = (t0 >= c0) + (t0 >= c01);
    const size_t node = (size_t)(t0 - (comp == 0 ? 0 : (comp == 1 ? c0 : c01)) + (comp == 0 ? fa.first[0] : (comp == 1 ? fa.first[1] : fa.first[2])));
    ulonglong2 *a = reinterpret_cast<ulonglong2 *>(comp == 0 ? fa.acc[0] : (comp == 1 ? fa.acc[1] : fa.acc[2])) + node;
    ulonglong2 v = *a;
    if (v.x | v.y) *a = make_ulonglong2(0ull, 0ull);
    const long long n = (long long)v.x, w = (long long)v.y;
    float wf = (float)((double)w * (1.0 / kWeightScaleD));
    float nf = (float)((double)n * inv_num_scale_d(num_exponent(sp)));
    float value = nf;
    if (wf > 0.0f) value = nf / wf;
    bool isset = (double)wf > 1e-9;
    if (isset && src.n > 0) {
        const int ni = g.I + (comp == 0), nj = g.J + (comp == 1);
        const int i = (int)(node % (size_t)ni), j = (int)((node / (size_t)ni) % (size_t)nj), k = (int)(node / ((size_t)ni * (size_t)nj)) + g.k0;
        float fx = (float)(comp == 0 ? __dmul_rn((double)(float)i, g.dx) : __dmul_rn(__dadd_rn((double)(float)i, 0.5), g.dx));
        float fy = (float)(comp == 1 ? __dmul_rn((double)(float)j, g.dx) : __dmul_rn(__dadd_rn((double)(float)j, 0.5), g.dx));
        float fz = (float)(comp == 2 ? __dmul_rn((double)(float)k, g.dx) : __dmul_rn(__dadd_rn((double)(float)k, 0.5), g.dx));
        for (int q = 0; q < src.n; q++)
            if (source_contains(src.s[q], fx, fy, fz)) value = src.s[q].velocity[comp];
    }
    (comp == 0 ? fa.val[0] : (comp == 1 ? fa.val[1] : fa.val[2]))[node] = value;
    (comp == 0 ? fa.setmask[0] : (comp == 1 ? fa.setmask[1] : fa.setmask[2]))[node] = isset ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// K1b (exact): node-centric gather in the reference's summation order.  One thread per node of one
// component; it visits the 27 surrounding cells in ascending (k,j,i) and each cell's particles in
// stored (stable-sorted) order, i.e. exactly the order ScalarField::addPointValue would have been called
// in had the particle array been sorted by linear cell index -- float accumulation included
// (array3d.h:263-271).  Parity tool, not the fast path.
// ------------------------------------------------------------------------------------------------
// is node index `node` inside [gmin, gmax] of Grid3d::getGridIndexBounds for offset-space coordinate q?
__device__ __forceinline__ bool in_index_bounds(float q, int node, double radius, const Grid &g) {
    int c = cell_floor((double)q, g.invdx);
    float trans = __fsub_rn(q, node_pos(c, g.dx));
    int gmin = c - (int)fmax(0.0, ceil(__dmul_rn(__dsub_rn(radius, (double)trans), g.invdx)));
    int gmax = c + (int)fmax(0.0, ceil(__dmul_rn(__dadd_rn(__dsub_rn(radius, g.dx), (double)trans), g.invdx)));
    return node >= gmin && node <= gmax;
}

template <int ARITH>
__global__ void k_p2g_gather(Grid g, int comp, SplatParams sp, Sources src, const int32_t *__restrict__ cell_start,
                             const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                             const float *__restrict__ vel, float *__restrict__ val, uint8_t *__restrict__ setmask) {
    int ni = g.I + (comp == 0), nj = g.J + (comp == 1), nkl = g.k1 - g.k0 + (comp == 2);
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, kl = blockIdx.z;
    if (i >= ni || j >= nj || kl >= nkl) return;
    int k = kl + g.k0;
    float off = (float)g.halfdx;
    float gx = node_pos(i, g.dx), gy = node_pos(j, g.dx), gz = node_pos(k, g.dx);
    float field = 0.0f, weight = 0.0f;
    // A contributing particle lies within one cell (plus rounding slop) of the node in offset space, i.e.
    // within two real-space cells: scan the 5x5x5 real-space cell block, ascending (k,j,i).
    for (int ck = k - 2; ck <= k + 2; ck++) {
        if (ck < g.k0 || ck >= g.k1 || ck < 0 || ck >= g.K) continue;
        for (int cj = j - 2; cj <= j + 2; cj++) {
            if (cj < 0 || cj >= g.J) continue;
            for (int ci = i - 2; ci <= i + 2; ci++) {
                if (ci < 0 || ci >= g.I) continue;
                uint32_t key = brick_key(g, ci, cj, ck - g.k0);
                for (int r = cell_start[key], e = cell_start[key + 1]; r < e; r++) {
                    float qx = comp == 0 ? x[r] : __fsub_rn(x[r], off);
                    float qy = comp == 1 ? y[r] : __fsub_rn(y[r], off);
                    float qz = comp == 2 ? z[r] : __fsub_rn(z[r], off);
                    // the reference only visits nodes inside Grid3d::getGridIndexBounds (grid3d.h:350-371)
                    if (!in_index_bounds(qx, i, sp.radius, g) || !in_index_bounds(qy, j, sp.radius, g) ||
                        !in_index_bounds(qz, k, sp.radius, g)) continue;
                    float d2 = dist2(__fsub_rn(gx, qx), __fsub_rn(gy, qy), __fsub_rn(gz, qz));
                    if ((double)d2 < sp.rsq) {
                        double w = (ARITH == 1) ? kernel_weight_exact(sp, (double)d2) : (double)kernel_weight_fast<false>(sp, d2);
                        field = __fadd_rn(field, (float)__dmul_rn(w, (double)vel[r]));
                        weight = __fadd_rn(weight, (float)w);
                    }
                }
            }
        }
    }
    float value = field;
    if (weight > 0.0f) value = __fdiv_rn(field, weight);
    bool isset = (double)weight > 1e-9;
    if (isset) {
        for (int s = 0; s < src.n; s++) {
            float fx = (float)(comp == 0 ? __dmul_rn((double)(float)i, g.dx) : __dmul_rn(__dadd_rn((double)(float)i, 0.5), g.dx));
            float fy = (float)(comp == 1 ? __dmul_rn((double)(float)j, g.dx) : __dmul_rn(__dadd_rn((double)(float)j, 0.5), g.dx));
            float fz = (float)(comp == 2 ? __dmul_rn((double)(float)k, g.dx) : __dmul_rn(__dadd_rn((double)(float)k, 0.5), g.dx));
            if (source_contains(src.s[s], fx, fy, fz)) value = src.s[s].velocity[comp];
        }
    }
    size_t node = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)kl);
    val[node] = value;
    setmask[node] = isset ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// K1c: face assembly  (FluidSimulation::_advectVelocityField{U,V,W}, fluidsimulation.cpp:2597-2730)
// faces bordering a fluid cell take the node value if set, else the mean of the in-range 26 neighbours
// that qualify (U: |value| > 0, :2630;  V,W: isValueSet, :2675/:2720); every other face is 0.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cell_is_fluid(const Grid &g, const uint8_t *__restrict__ m, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= g.I || j >= g.J || k >= g.K) return false;   // out of range reads as solid
    int kl = k - g.k0;
    if (kl < 0 || kl >= g.k1 - g.k0) return false;
    return m[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)kl)] == GFS_FLUID;
}

struct AssembleArgs {
    const float *val[3];
    const uint8_t *setmask[3];
    float *out[3];
    int k_lo, k_hi, k_hi_w;          // cell layers [k_lo, k_hi) for the u, v faces; w face layers [k_lo, k_hi_w)
};

template <int COMP>
__device__ __forceinline__ void assemble_face(const Grid &g, const AssembleArgs &aa, int i, int j, int kl, bool borders) {
    const int ni = g.I + (COMP == 0), nj = g.J + (COMP == 1), nkl = g.k1 - g.k0 + (COMP == 2);
    const float *__restrict__ val = aa.val[COMP];
    const uint8_t *__restrict__ setmask = aa.setmask[COMP];
    const size_t node = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)kl);
    // both loads are issued before the material-dependent branch (the kernel is a chain of dependent loads otherwise)
    const uint8_t isset = setmask[node];
    const float own = val[node];
    float r = 0.0f;
    if (borders) {
        if (isset) {
            r = own;
        } else {
            double avg = 0.0, cnt = 0.0;
            for (int nk = kl - 1; nk <= kl + 1; nk++)
                for (int nj_ = j - 1; nj_ <= j + 1; nj_++)
                    for (int ni_ = i - 1; ni_ <= i + 1; ni_++) {
                        if (ni_ == i && nj_ == j && nk == kl) continue;
                        if (ni_ < 0 || nj_ < 0 || nk < 0 || ni_ >= ni || nj_ >= nj || nk >= nkl) continue;
                        const size_t nn = (size_t)ni_ + (size_t)ni * ((size_t)nj_ + (size_t)nj * (size_t)nk);
                        const float v = val[nn];
                        const bool ok = COMP == 0 ? (fabs((double)v) > 0.0) : (setmask[nn] != 0);
                        if (ok) { avg = __dadd_rn(avg, (double)v); cnt += 1.0; }
                    }
            if (cnt > 0.0) r = (float)__ddiv_rn(avg, cnt);
        }
    }
    aa.out[COMP][(size_t)i + (size_t)g.pitch[COMP] * ((size_t)j + (size_t)nj * (size_t)kl)] = r;
}

// one thread per node (i, j) of the (I+1) x (J+1) plane, blockIdx.y = layer: the u, v and w faces whose lower corner is
// that node, sharing the four material reads that decide "borders fluid"
__global__ void __launch_bounds__(256) k_assemble(Grid g, const uint8_t *__restrict__ material, AssembleArgs aa) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = (uint32_t)g.I + 1u;
    if (t >= w * ((uint32_t)g.J + 1u)) return;
    const int i = (int)(t % w), j = (int)(t / w);
    const int k = aa.k_lo + (int)blockIdx.y, kl = k - g.k0;
    // four independent material reads (no short circuit: they overlap)
    const bool f = cell_is_fluid(g, material, i, j, k), fi = cell_is_fluid(g, material, i - 1, j, k);
    const bool fj = cell_is_fluid(g, material, i, j - 1, k), fk = cell_is_fluid(g, material, i, j, k - 1);
    const bool uv = k < aa.k_hi;
    if (uv && j < g.J) assemble_face<0>(g, aa, i, j, kl, f | fi);
    if (uv && i < g.I) assemble_face<1>(g, aa, i, j, kl, f | fj);
    if (k < aa.k_hi_w && i < g.I && j < g.J) assemble_face<2>(g, aa, i, j, kl, f | fk);
}

// ------------------------------------------------------------------------------------------------
// K1b+c fused (fast arithmetic): normalisation (ScalarField::applyWeightField), isValueSet, inflow override and face
// assembly in ONE pass over the nodes, straight from the fixed-point accumulators -- the node grid "ugrid" and its
// isValueSet mask (val / setmask of k_p2g_finalize + k_assemble) are never materialised.  A face that borders fluid but
// was not set needs the 26 neighbours' values: they are re-derived from the neighbours' accumulators on the spot (rare:
// free-surface faces only).  Same arithmetic, same bits as the two-kernel form.  The accumulators are NOT cleared here
// (neighbouring threads still read them): the next P2G starts with a memset of its layers.
// ------------------------------------------------------------------------------------------------
struct FusedArgs {
    const unsigned long long *acc[3];
    float *out[3];
    int k_lo, k_hi, k_hi_w;
};

template <int COMP>
__device__ __forceinline__ float node_value(const Grid &g, const SplatParams &sp, const Sources &src, const unsigned long long *__restrict__ acc,
                                            int i, int j, int kl, double inv_ns, bool &isset) {
    const int ni = g.I + (COMP == 0), nj = g.J + (COMP == 1);
    const size_t node = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)kl);
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(acc) + node);
    const float wf = (float)((double)(long long)v.y * (1.0 / kWeightScaleD));
    const float nf = (float)((double)(long long)v.x * inv_ns);
    float value = nf;
    if (wf > 0.0f) value = nf / wf;
    isset = (double)wf > 1e-9;
    if (isset && src.n > 0) {
        const int k = kl + g.k0;
        const float fx = (float)(COMP == 0 ? __dmul_rn((double)(float)i, g.dx) : __dmul_rn(__dadd_rn((double)(float)i, 0.5), g.dx));
        const float fy = (float)(COMP == 1 ? __dmul_rn((double)(float)j, g.dx) : __dmul_rn(__dadd_rn((double)(float)j, 0.5), g.dx));
        const float fz = (float)(COMP == 2 ? __dmul_rn((double)(float)k, g.dx) : __dmul_rn(__dadd_rn((double)(float)k, 0.5), g.dx));
        for (int q = 0; q < src.n; q++)
            if (source_contains(src.s[q], fx, fy, fz)) value = src.s[q].velocity[COMP];
    }
    return value;
}

template <int COMP>
__device__ __forceinline__ void fused_face(const Grid &g, const SplatParams &sp, const Sources &src, const FusedArgs &fa, double inv_ns,
                                           int i, int j, int kl, bool borders) {
    const int ni = g.I + (COMP == 0), nj = g.J + (COMP == 1), nkl = g.k1 - g.k0 + (COMP == 2);
    float r = 0.0f;
    if (borders) {
        bool isset;
        const float own = node_value<COMP>(g, sp, src, fa.acc[COMP], i, j, kl, inv_ns, isset);
        if (isset) {
            r = own;
        } else {
            double avg = 0.0, cnt = 0.0;
            for (int nk = kl - 1; nk <= kl + 1; nk++)
                for (int nj_ = j - 1; nj_ <= j + 1; nj_++)
                    for (int ni_ = i - 1; ni_ <= i + 1; ni_++) {
                        if (ni_ == i && nj_ == j && nk == kl) continue;
                        if (ni_ < 0 || nj_ < 0 || nk < 0 || ni_ >= ni || nj_ >= nj || nk >= nkl) continue;
                        bool nset;
                        const float v = node_value<COMP>(g, sp, src, fa.acc[COMP], ni_, nj_, nk, inv_ns, nset);
                        const bool ok = COMP == 0 ? (fabs((double)v) > 0.0) : nset;
                        if (ok) { avg = __dadd_rn(avg, (double)v); cnt += 1.0; }
                    }
            if (cnt > 0.0) r = (float)__ddiv_rn(avg, cnt);
        }
    }
    fa.out[COMP][(size_t)i + (size_t)g.pitch[COMP] * ((size_t)j + (size_t)nj * (size_t)kl)] = r;
}

__global__ void __launch_bounds__(256) k_finalize_assemble(Grid g, SplatParams sp, Sources src, const uint8_t *__restrict__ material, FusedArgs fa) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = (uint32_t)g.I + 1u;
    if (t >= w * ((uint32_t)g.J + 1u)) return;
    const int i = (int)(t % w), j = (int)(t / w);
    const int k = fa.k_lo + (int)blockIdx.y, kl = k - g.k0;
    const double inv_ns = inv_num_scale_d(num_exponent(sp));
    const bool f = cell_is_fluid(g, material, i, j, k), fi = cell_is_fluid(g, material, i - 1, j, k);
    const bool fj = cell_is_fluid(g, material, i, j - 1, k), fk = cell_is_fluid(g, material, i, j, k - 1);
    const bool uv = k < fa.k_hi;
    if (uv && j < g.J) fused_face<0>(g, sp, src, fa, inv_ns, i, j, kl, f | fi);
    if (uv && i < g.I) fused_face<1>(g, sp, src, fa, inv_ns, i, j, kl, f | fj);
    if (k < fa.k_hi_w && i < g.I && j < g.J) fused_face<2>(g, sp, src, fa, inv_ns, i, j, kl, f | fk);
}

// ------------------------------------------------------------------------------------------------
// K3: MACVelocityField::extrapolateVelocityField (macvelocityfield.cpp:577-798) on a resident field.
// The reference's sweeps are sequential and in place; every one of them is order independent (a layer pass only turns
// -1 cells into L and only reads "== L-1"; a face pass writes faces that do not border layer L-1 and reads faces that
// do), so each is one flat kernel with the same arithmetic: mean in double over the neighbours in the reference's order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cell_equals(const Grid &g, const int8_t *__restrict__ a, int i, int j, int k, int value) {
    if (i < 0 || j < 0 || k < 0 || i >= g.I || j >= g.J || k >= g.K) return false;
    return a[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)k)] == value;
}

template <int COMP>
__device__ __forceinline__ bool face_borders(const Grid &g, const int8_t *__restrict__ a, int i, int j, int k, int value) {
    return cell_equals(g, a, i, j, k, value) || cell_equals(g, a, i - (COMP == 0), j - (COMP == 1), k - (COMP == 2), value);
}

struct FieldRW { float *c[3]; };

// _resetExtrapolatedFluidVelocities (:748-784) + layer 0 of _updateExtrapolationLayers (:603-613); one thread per node
__global__ void __launch_bounds__(256) k_extrapolate_reset(Grid g, const uint8_t *__restrict__ material, int8_t *__restrict__ layer, FieldRW f) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = (uint32_t)g.I + 1u;
    if (t >= w * ((uint32_t)g.J + 1u)) return;
    const int i = (int)(t % w), j = (int)(t / w), k = (int)blockIdx.y;
    const int8_t *m = reinterpret_cast<const int8_t *>(material);
    const bool fl = cell_equals(g, m, i, j, k, GFS_FLUID);
    if (i < g.I && j < g.J && k < g.K) layer[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)k)] = fl ? 0 : -1;
    if (j < g.J && k < g.K && !(fl || cell_equals(g, m, i - 1, j, k, GFS_FLUID)))
        f.c[0][(size_t)i + (size_t)g.pitch[0] * ((size_t)j + (size_t)g.J * (size_t)k)] = 0.0f;
    if (i < g.I && k < g.K && !(fl || cell_equals(g, m, i, j - 1, k, GFS_FLUID)))
        f.c[1][(size_t)i + (size_t)g.pitch[1] * ((size_t)j + (size_t)(g.J + 1) * (size_t)k)] = 0.0f;
    if (i < g.I && j < g.J && !(fl || cell_equals(g, m, i, j, k - 1, GFS_FLUID)))
        f.c[2][(size_t)i + (size_t)g.pitch[2] * ((size_t)j + (size_t)g.J * (size_t)k)] = 0.0f;
}

// _updateExtrapolationLayer (:577-601), gather form: a -1, non-solid cell with a 6-neighbour in layer L-1 becomes L
__global__ void __launch_bounds__(256) k_extrapolate_mark(Grid g, const uint8_t *__restrict__ material, int8_t *__restrict__ layer, int L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)g.I * (uint32_t)g.J) return;
    const int i = (int)(t % (uint32_t)g.I), j = (int)(t / (uint32_t)g.I), k = (int)blockIdx.y;
    const size_t c = (size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)k);
    if (layer[c] != -1 || material[c] == GFS_SOLID) return;
    // (a layer L-1 >= 1 cell is never solid; layer 0 cells are fluid)
    const bool hit = cell_equals(g, layer, i - 1, j, k, L - 1) || cell_equals(g, layer, i + 1, j, k, L - 1) ||
                     cell_equals(g, layer, i, j - 1, k, L - 1) || cell_equals(g, layer, i, j + 1, k, L - 1) ||
                     cell_equals(g, layer, i, j, k - 1, L - 1) || cell_equals(g, layer, i, j, k + 1, L - 1);
    if (hit) layer[c] = (int8_t)L;
}

template <int COMP>
__device__ __forceinline__ void extrapolate_face(const Grid &g, const uint8_t *__restrict__ material, const int8_t *__restrict__ layer,
                                                 float *__restrict__ a, int i, int j, int k, int L) {
    const int ni = g.I + (COMP == 0), nj = g.J + (COMP == 1), nk = g.K + (COMP == 2);
    if (i >= ni || j >= nj || k >= nk) return;
    if (!face_borders<COMP>(g, layer, i, j, k, L) || face_borders<COMP>(g, layer, i, j, k, L - 1) ||
        face_borders<COMP>(g, reinterpret_cast<const int8_t *>(material), i, j, k, GFS_SOLID)) return;
    const size_t pitch = (size_t)g.pitch[COMP];
    double sum = 0.0, cnt = 0.0;
    const int d6[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};       // grid3d.h:205-212
#pragma unroll
    for (int q = 0; q < 6; q++) {
        const int x = i + d6[q][0], y = j + d6[q][1], z = k + d6[q][2];
        if (x < 0 || y < 0 || z < 0 || x >= ni || y >= nj || z >= nk) continue;
        if (face_borders<COMP>(g, layer, x, y, z, L - 1)) {
            sum = __dadd_rn(sum, (double)a[(size_t)x + pitch * ((size_t)y + (size_t)nj * (size_t)z)]);
            cnt += 1.0;
        }
    }
    a[(size_t)i + pitch * ((size_t)j + (size_t)nj * (size_t)k)] = sum == 0.0 ? 0.0f : (float)__ddiv_rn(sum, cnt);
}

// _extrapolateVelocitiesForLayerIndexU/V/W (:692-744); one thread per node, its three faces
__global__ void __launch_bounds__(256) k_extrapolate_faces(Grid g, const uint8_t *__restrict__ material, const int8_t *__restrict__ layer,
                                                           FieldRW f, int L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = (uint32_t)g.I + 1u;
    if (t >= w * ((uint32_t)g.J + 1u)) return;
    const int i = (int)(t % w), j = (int)(t / w), k = (int)blockIdx.y;
    extrapolate_face<0>(g, material, layer, f.c[0], i, j, k, L);
    extrapolate_face<1>(g, material, layer, f.c[1], i, j, k, L);
    extrapolate_face<2>(g, material, layer, f.c[2], i, j, k, L);
}

// ------------------------------------------------------------------------------------------------
// A14, the resolve routine: FluidSimulation::_resolveParticleSolidCellCollision
// (fluidsimulation.cpp:3145-3179) = Collision::getLineSegmentVoxelIntersection (collision.cpp:303-400) + normalize +
// Collision::rayIntersectsAABB (:404-448, including its dir.x-for-dir.z slip) + back-off by 0.05 dx.  The G2P kernels
// only LIST the (rare) particles whose advected cell is solid -- slot and advected position -- and leave p0 in place;
// k_resolve_collisions then runs the reference's arithmetic on the list (float vec3 operations with explicit
// roundings, double voxel walk and slab test, exactly as oracle/oracle.c restates them) and bins the result.
// ------------------------------------------------------------------------------------------------
struct CollList {
    float4 *list;                  // {slot as int bits, p1.x, p1.y, p1.z}; null = keep p0 (solid test only)
    unsigned int *count;
    unsigned int cap;
    unsigned int cell_cap;         // per-cell particle cap applied while binning (0 = none), see take_ticket
};

__device__ __forceinline__ bool cell_solid_or_outside(const Grid &g, const uint8_t *__restrict__ m, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= g.I || j >= g.J || k >= g.K) return true;      // fluidmaterialgrid.cpp:25-29
    return m[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)(k - g.k0))] == GFS_SOLID;
}

__device__ inline void resolve_collision(const Grid &g, const uint8_t *__restrict__ material, const float p0[3], const float p1[3], float out[3]) {
    out[0] = p0[0]; out[1] = p0[1]; out[2] = p0[2];
    const float s = (float)g.invdx;                                  // vec3 *= (float)invdx, vmath.cpp:79-84
    float a0[3], a1[3];
    int g1[3], st[3], c[3];
    double gp[3], v[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        a0[a] = __fmul_rn(p0[a], s); a1[a] = __fmul_rn(p1[a], s);
        const int g0 = (int)floor((double)a0[a]);
        g1[a] = (int)floor((double)a1[a]);
        st[a] = g1[a] > g0 ? 1 : (g1[a] < g0 ? -1 : 0);
        c[a] = g0;
        gp[a] = (double)(g0 + (g1[a] > g0 ? 1 : 0));
        v[a] = a1[a] == a0[a] ? 1.0 : (double)__fsub_rn(a1[a], a0[a]);
    }
    const double vxvy = __dmul_rn(v[0], v[1]), vxvz = __dmul_rn(v[0], v[2]), vyvz = __dmul_rn(v[1], v[2]);
    double ex = __dmul_rn(__dsub_rn(gp[0], (double)a0[0]), vyvz), ey = __dmul_rn(__dsub_rn(gp[1], (double)a0[1]), vxvz),
           ez = __dmul_rn(__dsub_rn(gp[2], (double)a0[2]), vxvy);
    const double dex = __dmul_rn((double)st[0], vyvz), dey = __dmul_rn((double)st[1], vxvz), dez = __dmul_rn((double)st[2], vxvy);
    bool found = false;
    for (int iter = 0; iter < 1000000; iter++) {
        if (c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[0] < g.I && c[1] < g.J && c[2] < g.K &&
            material[(size_t)c[0] + (size_t)g.I * ((size_t)c[1] + (size_t)g.J * (size_t)(c[2] - g.k0))] == GFS_SOLID) { found = true; break; }
        if (c[0] == g1[0] && c[1] == g1[1] && c[2] == g1[2]) break;
        const double xr = fabs(ex), yr = fabs(ey), zr = fabs(ez);
        if (st[0] != 0 && (st[1] == 0 || xr < yr) && (st[2] == 0 || xr < zr)) { c[0] += st[0]; ex = __dadd_rn(ex, dex); }
        else if (st[1] != 0 && (st[2] == 0 || yr < zr)) { c[1] += st[1]; ey = __dadd_rn(ey, dey); }
        else if (st[2] != 0) { c[2] += st[2]; ez = __dadd_rn(ez, dez); }
    }
    if (!found) return;
    // raynorm = normalize(p1 - p0)   (vmath.h:81-92: float length, inv = (float)(1.0 / len))
    const float d0 = __fsub_rn(p1[0], p0[0]), d1 = __fsub_rn(p1[1], p0[1]), d2 = __fsub_rn(p1[2], p0[2]);
    const float lensq = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
    const float len = (float)__dsqrt_rn((double)lensq);
    const float inv = (float)__ddiv_rn(1.0, (double)len);
    const float rn[3] = {__fmul_rn(d0, inv), __fmul_rn(d1, inv), __fmul_rn(d2, inv)};
    const double eps = 1e-10;
    float dir[3] = {rn[0], rn[1], rn[2]};
    if (fabs((double)dir[0]) < eps) dir[0] = (float)(dir[0] < 0 ? -eps : eps);
    if (fabs((double)dir[1]) < eps) dir[1] = (float)(dir[1] < 0 ? -eps : eps);
    if (fabs((double)dir[0]) < eps) dir[2] = (float)(dir[2] < 0 ? -eps : eps);       // sic (collision.cpp:417)
    double tmin = 0.0, tmax = 0.0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float bmin = (float)__dmul_rn((double)(float)c[a], g.dx);               // GridIndexToPosition, grid3d.h:83-85
        const float bmax = __fadd_rn(bmin, (float)g.dx);                              // aabb.cpp:481-483
        const float dinv = (float)__ddiv_rn(1.0, (double)dir[a]);
        const double t1 = (double)__fmul_rn(__fsub_rn(bmin, p0[a]), dinv), t2 = (double)__fmul_rn(__fsub_rn(bmax, p0[a]), dinv);
        if (a == 0) { tmin = fmin(t1, t2); tmax = fmax(t1, t2); }
        else { tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2)); }
    }
    if (!(tmax > fmax(tmin, 0.0))) return;
    const float tm = (float)tmin, back = (float)__dmul_rn(0.05, g.dx);
    float r[3];
#pragma unroll
    for (int a = 0; a < 3; a++) r[a] = __fsub_rn(__fadd_rn(p0[a], __fmul_rn(dir[a], tm)), __fmul_rn(rn[a], back));
    const int i = cell_floor((double)r[0], g.invdx), j = cell_floor((double)r[1], g.invdx), k = cell_floor((double)r[2], g.invdx);
    if (cell_solid_or_outside(g, material, i, j, k)) return;
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

// ------------------------------------------------------------------------------------------------
// K2: fused G2P.  One thread per (sorted) particle:
//   vnew = sample(NEW, p0), vold = sample(SAVED, p0), both validated        (fluidsimulation.cpp:3115-3116)
//   v   <- (float)ratio*vnew + (float)(1-ratio)*((v + vnew) - vold)         (:3118-3128)
//   p1  = RK{order}(p0) through NEW; its k1 is the unvalidated sample at p0  (particleadvector.cpp:1045-1078)
//   solid test on cell(p1) (out of range reads as solid); a hit keeps p0 and is counted  (:3198-3208)
// ------------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(256) k_g2p_advect(Grid g, FieldPtrs fnew, FieldPtrs fsaved, const uint8_t *__restrict__ material,
                             int interp, int order, RkCoef rk, float ratio_pic, float ratio_flip, int64_t n,
                             const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                             const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                             float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
                             float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz,
                             unsigned long long *__restrict__ counters /* [2] = solid hits */,
                             uint32_t nkeys, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ rank_out,
                             uint32_t *__restrict__ counts, unsigned int *__restrict__ vmax_bits, CollList coll) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (r < n) {
        float px = x[r], py = y[r], pz = z[r];
        float k1x, k1y, k1z, sx, sy, sz;
        evaluate_any<ARITH>(g, fnew, interp, px, py, pz, k1x, k1y, k1z);
        evaluate_any<ARITH>(g, fsaved, interp, px, py, pz, sx, sy, sz);
        float nx = k1x, ny = k1y, nz = k1z;
        validate3(nx, ny, nz);
        validate3(sx, sy, sz);
        float ux = vx[r], uy = vy[r], uz = vz[r];
        float wx = __fadd_rn(__fmul_rn(nx, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(ux, nx), sx), ratio_flip));
        float wy = __fadd_rn(__fmul_rn(ny, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uy, ny), sy), ratio_flip));
        float wz = __fadd_rn(__fmul_rn(nz, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uz, nz), sz), ratio_flip));
        ovx[r] = wx; ovy[r] = wy; ovz[r] = wz;

        float qx, qy, qz;
        rk_advance<ARITH>(g, fnew, interp, order, rk, px, py, pz, k1x, k1y, k1z, qx, qy, qz);
        bool deferred = false;
        if (material) {
            int i = cell_floor((double)qx, g.invdx), j = cell_floor((double)qy, g.invdx), k = cell_floor((double)qz, g.invdx);
            bool solid = true;                                   // NaN -> huge negative index -> out of range -> solid
            if (i >= 0 && j >= 0 && k >= 0 && i < g.I && j < g.J && k < g.K) {
                int kl = k - g.k0;
                // a particle that leaves this slab's stored layers is not judged here: it migrates first
                solid = (kl >= 0 && kl < g.k1 - g.k0) ? material[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)kl)] == GFS_SOLID : false;
            }
            if (solid) {
                atomicAdd(&counters[2], 1ull);
                if (coll.list) {
                    const unsigned int tk = atomicAdd(coll.count, 1u);
                    if (tk < coll.cap) { coll.list[tk] = make_float4(__int_as_float((int)r), qx, qy, qz); deferred = true; }
                    else atomicAdd(&counters[3], 1ull);                  // list full: reported, see gfs_stats_t.collision_overflow
                }
                qx = px; qy = py; qz = pz;
            }
        }
        ox[r] = qx; oy[r] = qy; oz[r] = qz;
        if (keys_out) {          // bin for the next substep's counting sort while the position is in registers
            if (!deferred) {
                uint32_t key = position_key(g, nkeys, qx, qy, qz);
                rank_out[r] = take_ticket(counts, nkeys, coll.cell_cap, key);
                keys_out[r] = key;
            }
            m = fmaxf(fabsf(wx), fmaxf(fabsf(wy), fabsf(wz)));
            if (!(m < 3.0e38f)) m = 0.0f;
        }
    }
    if (keys_out) block_vmax(m, vmax_bits);
}

// ------------------------------------------------------------------------------------------------
// K2 (brick tiles, TMA).  One CTA per brick of 8^3 cells, whose particles are one contiguous run of the
// sorted arrays.  The u/v/w sub-blocks of the NEW and SAVED fields that the brick's particles can touch are
// staged into shared memory by six TMA tiled loads (cp.async.bulk.tensor.3d, completion on one mbarrier);
// TMA's zero fill of out-of-bounds box elements IS the reference's "taps outside the array read 0"
// (macvelocityfield.cpp:99-145).  NEW is staged with a margin of kMargin cells for the RK stage positions;
// a stage position that leaves the staged block falls back to global loads (same arithmetic).
// Power-of-two dx only (fp32-exact index arithmetic); other grids use k_g2p_advect.
// ------------------------------------------------------------------------------------------------
struct BrickMaps { CUtensorMap m[6]; };     // NEW u,v,w then SAVED u,v,w

#ifndef GFS_TRILINEAR_CTAS
#define GFS_TRILINEAR_CTAS 3
#endif
#ifndef GFS_TRICUBIC_CTAS
#define GFS_TRICUBIC_CTAS 2
#endif

template <int INTERP> struct BrickTile {
    // trilinear: taps c, c+1;  tricubic: taps c-1 .. c+2.  c ranges over [8b-1-M, 8b+7+M] (M = motion margin).
    // Along y and z the boxes are exactly that (z plus one unused plane, see below); along x (the contiguous axis)
    // TMA needs the box to start on a 16-byte boundary, so every box starts at node 8b-4 and is 16 nodes wide.
    //
    // Shared-memory layout (bank-conflict free for the access pattern of sorted particles).  The resident arrays keep
    // 4 zero floats in front of every row, so node 8b-4 sits at storage column 8b, and the tensor map views a row as
    // [x_hi][x_lo = 8 floats].  The 4-D box {x_lo 8, z nZ, x_hi 2, y nY} lands as
    //     word(x, y, z) = (x & 7) + 8 * (z + nZ * ((x >> 3) + 2 * y)),        x, y, z relative to the box origin
    // and with nZ ODD the bank is (x & 7) + 8 * (z +- (x >> 3)) + 16 * y  (mod 32): the lanes of a warp -- a few
    // consecutive cells along x, two candidate rows in y and two in z because of the half-cell stagger -- fall into
    // four disjoint 8-bank windows.  (The plain [z][y][x 16] box put both z candidates on the same banks: 43 % of all
    // shared-memory wavefronts of this kernel were bank conflicts, profiles/r01_k_g2p_brick.md.)
    static constexpr int kMargin = 1;
    static constexpr int kLo = (INTERP == 1) ? 1 : 0, kHi = (INTERP == 1) ? 2 : 1;
    static constexpr int kOrgX = 4, kX = 16;                                // x: nodes [8b-4, 8b+11] for every box
    static constexpr int nOrg = 1 + kMargin + kLo;                          // NEW y/z origin = 8b - nOrg
    static constexpr int nY = 9 + 2 * kMargin + kLo + kHi;
    // (trilinear keeps the dense [z][y][x 16] box: it does 1/7 of the shared loads per sample, is bound by instruction
    // issue, and the split-column address arithmetic cost it more than the conflicts did -- measured 3.53 -> 3.97 ms.)
#ifndef GFS_SKEW_TRILINEAR
#define GFS_SKEW_TRILINEAR 0
#endif
    static constexpr bool kSkew = INTERP == 1 || GFS_SKEW_TRILINEAR;
    static constexpr int nZ = kSkew ? (nY | 1) : nY;                        // odd plane count (>= nY)
    static constexpr int sOrg = 1 + kLo;                                    // SAVED (sampled at p0 only, no margin)
    static constexpr int sY = 9 + kLo + kHi;
    static constexpr int sZ = kSkew ? (sY | 1) : sY;
    static constexpr int nBox = kX * nY * nZ, sBox = kX * sY * sZ;           // floats per staged box
    static constexpr int nCount = (nBox + 31) / 32 * 32, sCount = (sBox + 31) / 32 * 32;   // 128-byte aligned slots
    static constexpr uint32_t kTxBytes = 3 * (nBox + sBox) * sizeof(float);
    static constexpr size_t kSmemBytes = 3 * (nCount + sCount) * sizeof(float) + 128 + 16;   // + alignment slack + mbarrier
    static_assert(kOrgX >= nOrg && kX - kOrgX >= 9 + kMargin + kHi, "x box must cover the y/z node range");
    static_assert(!kSkew || ((nZ & 1) && (sZ & 1)), "plane counts must be odd for the bank skew");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// box {x_lo 8, z, x_hi 2, y} of a resident field at (storage column 8 * xh, row y, plane z)
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *map, int z, int xh, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(0), "r"(z), "r"(xh), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// dense box {x 16, y, z} at (storage column x, row y, plane z)
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// word offset of box column x (0..15) in the skewed layout
template <int NZ>
__device__ __forceinline__ int tile_col(int x) { return x + (x >> 3) * (8 * NZ - 8); }      // x in [0, 16)

// one component from a staged tile: ax/ay/az are global node indices + fractions, (ox,oy,oz) the tile origin
template <int INTERP, int NZ>
__device__ __forceinline__ float tile_sample(const float *__restrict__ t, const AxisIdxF &ax, const AxisIdxF &ay, const AxisIdxF &az,
                                             int ox, int oy, int oz) {
    if (INTERP == 0 && !GFS_SKEW_TRILINEAR) {          // dense [z][y][x 16] box, NZ = rows per plane
        const float *r = t + (ax.i - ox) + 16 * ((ay.i - oy) + NZ * (az.i - oz));
        constexpr int BX = 16, BY = NZ;
        const float p000 = r[0], p100 = r[1], p010 = r[BX], p110 = r[BX + 1];
        const float p001 = r[BX * BY], p101 = r[BX * BY + 1], p011 = r[BX * BY + BX], p111 = r[BX * BY + BX + 1];
        float c00 = fmaf(ax.t, __fsub_rn(p100, p000), p000), c10 = fmaf(ax.t, __fsub_rn(p110, p010), p010);
        float c01 = fmaf(ax.t, __fsub_rn(p101, p001), p001), c11 = fmaf(ax.t, __fsub_rn(p111, p011), p011);
        float c0 = fmaf(ay.t, __fsub_rn(c10, c00), c00), c1 = fmaf(ay.t, __fsub_rn(c11, c01), c01);
        return fmaf(az.t, __fsub_rn(c1, c0), c0);
    }
    constexpr int RY = 16 * NZ, RZ = 8;                       // word strides of one row / one plane
    const int x0 = ax.i - ox;
    const float *r = t + RY * (ay.i - oy) + RZ * (az.i - oz);
    if (INTERP == 1) {
        float wx[4], wy[4], wz[4];
        cr_weights(ax.t, wx); cr_weights(ay.t, wy); cr_weights(az.t, wz);
        const int c0 = tile_col<NZ>(x0 - 1), c1 = tile_col<NZ>(x0), c2 = tile_col<NZ>(x0 + 1), c3 = tile_col<NZ>(x0 + 2);
        float acc = 0.0f;
#pragma unroll
        for (int pk = 0; pk < 4; pk++) {
            float sk = 0.0f;
#pragma unroll
            for (int pj = 0; pj < 4; pj++) {
                const float *q = r + RY * (pj - 1) + RZ * (pk - 1);
                float sj = __fmul_rn(wx[0], q[c0]);
                sj = fmaf(wx[1], q[c1], sj); sj = fmaf(wx[2], q[c2], sj); sj = fmaf(wx[3], q[c3], sj);
                sk = fmaf(wy[pj], sj, sk);
            }
            acc = fmaf(wz[pk], sk, acc);
        }
        return acc;
    }
    const float *r0 = r + tile_col<NZ>(x0), *r1 = r + tile_col<NZ>(x0 + 1);
    const float p000 = r0[0], p100 = r1[0], p010 = r0[RY], p110 = r1[RY];
    const float p001 = r0[RZ], p101 = r1[RZ], p011 = r0[RZ + RY], p111 = r1[RZ + RY];
    float c00 = fmaf(ax.t, __fsub_rn(p100, p000), p000), c10 = fmaf(ax.t, __fsub_rn(p110, p010), p010);
    float c01 = fmaf(ax.t, __fsub_rn(p101, p001), p001), c11 = fmaf(ax.t, __fsub_rn(p111, p011), p011);
    float c0 = fmaf(ay.t, __fsub_rn(c10, c00), c00), c1 = fmaf(ay.t, __fsub_rn(c11, c01), c01);
    return fmaf(az.t, __fsub_rn(c1, c0), c0);
}

// all six index/fraction pairs of a position (fp32-exact for power-of-two dx)
struct SampleIdx { AxisIdxF ux, uy, uz, sx, sy, sz; };
// Power-of-two dx: u = x / dx is exact in fp32, the staggered coordinate is u - 0.5 (also exact), and the fraction
// (x - floor(u) dx) / dx of axis_index_f equals u - floor(u) -- same bits, a third of the operations (scaling by a
// power of two commutes with every rounding involved).
template <bool MAGIC>
__device__ __forceinline__ void axis_pair(float x, const Grid &g, AxisIdxF &plain, AxisIdxF &stag) {
    const float u = __fmul_rn(x, g.invdxf), us = __fsub_rn(u, 0.5f);
    float fu, fs;
    if (MAGIC) { fu = floor_small(u, plain.i); fs = floor_small(us, stag.i); }
    else { fu = floorf(u); plain.i = (int)fu; fs = floorf(us); stag.i = (int)fs; }
    plain.t = __fsub_rn(u, fu);
    stag.t = __fsub_rn(us, fs);
}

template <bool MAGIC = false>
__device__ __forceinline__ SampleIdx sample_idx(const Grid &g, float px, float py, float pz) {
    SampleIdx s;
    axis_pair<MAGIC>(px, g, s.ux, s.sx);
    axis_pair<MAGIC>(py, g, s.uy, s.sy);
    axis_pair<MAGIC>(pz, g, s.uz, s.sz);
    return s;
}

// sample through the NEW tile if every tap of every component lies inside it, else through global memory
template <int INTERP>
__device__ __forceinline__ void evaluate_tile(const Grid &g, const FieldPtrs &f, const float *__restrict__ tile, int bx, int by, int bz,
                                              float px, float py, float pz, float &ox, float &oy, float &oz) {
    typedef BrickTile<INTERP> T;
    if (!(px >= 0.0f && py >= 0.0f && pz >= 0.0f && px < g.xmaxf && py < g.ymaxf && pz < g.zmaxf)) { ox = oy = oz = 0.0f; return; }
    const SampleIdx s = sample_idx<INTERP == 1>(g, px, py, pz);
    const int x0 = bx - T::kOrgX, y0 = by - T::nOrg, z0 = bz - T::nOrg;
    // every tap c - kLo .. c + kHi of the six index variants must lie inside the staged box
    const int lo = T::kLo;
    const unsigned spanx = (unsigned)(T::kX - 1 - T::kHi - lo), span = (unsigned)(T::nY - 1 - T::kHi - lo);
    bool in = (unsigned)(s.ux.i - x0 - lo) <= spanx && (unsigned)(s.sx.i - x0 - lo) <= spanx &&
              (unsigned)(s.uy.i - y0 - lo) <= span && (unsigned)(s.sy.i - y0 - lo) <= span &&
              (unsigned)(s.uz.i - z0 - lo) <= span && (unsigned)(s.sz.i - z0 - lo) <= span;
    if (in) {
        ox = tile_sample<INTERP, T::nZ>(tile, s.ux, s.sy, s.sz, x0, y0, z0);
        oy = tile_sample<INTERP, T::nZ>(tile + T::nCount, s.sx, s.uy, s.sz, x0, y0, z0);
        oz = tile_sample<INTERP, T::nZ>(tile + 2 * T::nCount, s.sx, s.sy, s.uz, x0, y0, z0);
    } else {
        ox = sample_component_fast<0>(g, f.c[0], INTERP, s.ux, s.sy, s.sz);
        oy = sample_component_fast<1>(g, f.c[1], INTERP, s.sx, s.uy, s.sz);
        oz = sample_component_fast<2>(g, f.c[2], INTERP, s.sx, s.sy, s.uz);
    }
}

// Fused migration (z-slab sharding over peer memory): a particle whose advected position lies below / above the layers
// this GPU owns is written by the G2P kernel itself into the neighbour GPU's arrival buffer (NVLink store), and is
// binned here into the "dead" bin nkeys + 1 -- behind every cell and the out-of-grid bin, where no kernel looks.
struct Migrate {
    int own_lo, own_hi;            // cell layers [own_lo, own_hi) stay; INT_MIN / INT_MAX on a side without neighbour
    float *out[2];                 // the neighbours' arrival buffers, 6 floats per particle (null: no migration)
    unsigned int *count;           // count[0], count[1]: leavers down / up (tickets)
    unsigned int cap;
};

// the listed colliders: reference resolve, final position, and the binning / migration the G2P kernel skipped for them
__global__ void __launch_bounds__(128) k_resolve_collisions(Grid g, const uint8_t *__restrict__ material, CollList coll,
                                                            float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
                                                            const float *__restrict__ ovx, const float *__restrict__ ovy, const float *__restrict__ ovz,
                                                            uint32_t nkeys, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ rank_out,
                                                            uint32_t *__restrict__ counts, Migrate mg, KeyRange kr) {
    const unsigned int n = min(*coll.count, coll.cap);
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const float4 e = coll.list[t];
        const int r = __float_as_int(e.x);
        const float p0[3] = {ox[r], oy[r], oz[r]}, p1[3] = {e.y, e.z, e.w};
        float q[3];
        resolve_collision(g, material, p0, p1, q);
        ox[r] = q[0]; oy[r] = q[1]; oz[r] = q[2];
        if (keys_out) {
            uint32_t key = clamp_key(position_key(g, nkeys, q[0], q[1], q[2]), nkeys, kr);
            if (key < nkeys && (mg.out[0] || mg.out[1])) {
                const int k = cell_floor((double)q[2], g.invdx);
                const int side = k < mg.own_lo ? 0 : (k >= mg.own_hi ? 1 : -1);
                if (side >= 0) {
                    const unsigned int slot = atomicAdd(mg.count + side, 1u);
                    if (slot < mg.cap && mg.out[side]) {
                        float2 *dst = reinterpret_cast<float2 *>(mg.out[side] + 6 * (size_t)slot);
                        dst[0] = make_float2(q[0], q[1]); dst[1] = make_float2(q[2], ovx[r]); dst[2] = make_float2(ovy[r], ovz[r]);
                    }
                    key = nkeys + 1;
                }
            }
            rank_out[r] = take_ticket(counts, nkeys, coll.cell_cap, key);
            keys_out[r] = key;
        }
    }
}

template <int INTERP, bool MIGRATE>
__global__ void __launch_bounds__(256, INTERP == 1 ? GFS_TRICUBIC_CTAS : GFS_TRILINEAR_CTAS) k_g2p_brick(Grid g, const __grid_constant__ BrickMaps maps, FieldPtrs fnew, FieldPtrs fsaved,
                            const uint8_t *__restrict__ material, const int32_t *__restrict__ cell_start,
                            const int32_t *__restrict__ index, const int32_t *__restrict__ tag_in, int32_t *__restrict__ tag_out,
                            int order, RkCoef rk, float ratio_pic, float ratio_flip, int64_t n,
                            const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                            const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                            float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
                            float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz,
                            unsigned long long *__restrict__ counters, uint32_t nkeys, uint32_t *__restrict__ keys_out,
                            uint32_t *__restrict__ rank_out, uint32_t *__restrict__ counts, unsigned int *__restrict__ vmax_bits,
                            Migrate mg, CollList coll, uint32_t brick0) {
    typedef BrickTile<INTERP> T;
    // dynamic shared memory: [pad to 128 B] NEW u,v,w [nCount each] | SAVED u,v,w [sCount each] | mbarrier.
    // TMA destinations must be 128-byte aligned: align by hand, static shared variables precede this block.
    extern __shared__ unsigned char smem_raw[];
    // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler must keep knowing this is
    // shared memory, or every tap becomes a generic LD instead of an LDS)
    float *tiles = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    uint64_t &bar = *reinterpret_cast<uint64_t *>(tiles + 3 * (T::nCount + T::sCount));
    const uint32_t b = blockIdx.x + brick0, nbricks = nkeys / kBrickCells;      // brick0: first brick of this launch
    // the last CTA takes the overflow bin (particles outside the grid): no tile, global path only
    const bool overflow = b == nbricks;
    const int start = cell_start[(size_t)b * kBrickCells];
    const int end = overflow ? cell_start[(size_t)nkeys + 1] : cell_start[(size_t)(b + 1) * kBrickCells];     // the dead bin follows
    if (start >= end) return;
    const int bi = (int)(b % (uint32_t)g.nbi), bj = (int)((b / (uint32_t)g.nbi) % (uint32_t)g.nbj), bk = (int)(b / ((uint32_t)g.nbi * (uint32_t)g.nbj));
    const int bx = bi * kBrick, by = bj * kBrick, bz = bk * kBrick + g.k0;       // global node index of the brick origin
    float *tnew = tiles, *tsav = tiles + 3 * T::nCount;
    if (!overflow) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(T::kTxBytes) : "memory");
#pragma unroll
            for (int c = 0; c < 3; c++) {
                // z coordinate is local to the stored layers; node 8b-4 is storage column 8 * bi
                if (T::kSkew) {
                    tma_load_box(tnew + c * T::nCount, &maps.m[c], bz - g.k0 - T::nOrg, bi, by - T::nOrg, &bar);
                    tma_load_box(tsav + c * T::sCount, &maps.m[3 + c], bz - g.k0 - T::sOrg, bi, by - T::sOrg, &bar);
                } else {
                    tma_load_3d(tnew + c * T::nCount, &maps.m[c], 8 * bi, by - T::nOrg, bz - g.k0 - T::nOrg, &bar);
                    tma_load_3d(tsav + c * T::sCount, &maps.m[3 + c], 8 * bi, by - T::sOrg, bz - g.k0 - T::sOrg, &bar);
                }
            }
        }
        __syncthreads();                                     // barrier init visible to the waiters
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    float m = 0.0f;
    // software pipeline: the next particle is loaded while this one is processed, and the ticket returned by the
    // binning atomic of this particle is only stored during the next iteration (neither latency is waited on)
    int r = start + threadIdx.x;
    float nx_ = 0.f, ny_ = 0.f, nz_ = 0.f, nvx = 0.f, nvy = 0.f, nvz = 0.f;
    int ntag = 0;
    if (r < end) {
        const int s_ = index ? index[r] : r;
        nx_ = x[s_]; ny_ = y[s_]; nz_ = z[s_]; nvx = vx[s_]; nvy = vy[s_]; nvz = vz[s_]; ntag = tag_in[s_];
    }
    int pend_r = -1;
    uint32_t pend_rank = 0;
    for (; r < end; r += blockDim.x) {
        const float px = nx_, py = ny_, pz = nz_;
        const float ux = nvx, uy = nvy, uz = nvz;
        tag_out[r] = ntag;
        {
            const int rn = r + blockDim.x;
            if (rn < end) {
                const int s_ = index ? index[rn] : rn;
                nx_ = x[s_]; ny_ = y[s_]; nz_ = z[s_]; nvx = vx[s_]; nvy = vy[s_]; nvz = vz[s_]; ntag = tag_in[s_];
            }
        }
        if (pend_r >= 0) rank_out[pend_r] = pend_rank;
        float k1x, k1y, k1z, sx, sy, sz;
        if (!overflow) {
            // p0 lies in this brick: NEW and SAVED taps are all staged, and share the index/fraction set
            const SampleIdx s = sample_idx<INTERP == 1>(g, px, py, pz);
            const int n0x = bx - T::kOrgX, n0y = by - T::nOrg, n0z = bz - T::nOrg;
            k1x = tile_sample<INTERP, T::nZ>(tnew, s.ux, s.sy, s.sz, n0x, n0y, n0z);
            k1y = tile_sample<INTERP, T::nZ>(tnew + T::nCount, s.sx, s.uy, s.sz, n0x, n0y, n0z);
            k1z = tile_sample<INTERP, T::nZ>(tnew + 2 * T::nCount, s.sx, s.sy, s.uz, n0x, n0y, n0z);
            const int s0y = by - T::sOrg, s0z = bz - T::sOrg;
            sx = tile_sample<INTERP, T::sZ>(tsav, s.ux, s.sy, s.sz, n0x, s0y, s0z);
            sy = tile_sample<INTERP, T::sZ>(tsav + T::sCount, s.sx, s.uy, s.sz, n0x, s0y, s0z);
            sz = tile_sample<INTERP, T::sZ>(tsav + 2 * T::sCount, s.sx, s.sy, s.uz, n0x, s0y, s0z);
        } else {
            evaluate_pow2(g, fnew, INTERP, px, py, pz, k1x, k1y, k1z);
            evaluate_pow2(g, fsaved, INTERP, px, py, pz, sx, sy, sz);
        }
        float nx = k1x, ny = k1y, nz = k1z;
        validate3(nx, ny, nz);
        validate3(sx, sy, sz);
        const float wx = __fadd_rn(__fmul_rn(nx, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(ux, nx), sx), ratio_flip));
        const float wy = __fadd_rn(__fmul_rn(ny, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uy, ny), sy), ratio_flip));
        const float wz = __fadd_rn(__fmul_rn(nz, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uz, nz), sz), ratio_flip));
        ovx[r] = wx; ovy[r] = wy; ovz[r] = wz;

        // RK (particleadvector.cpp:1045-1078) as ONE stage loop (one sampling site keeps the kernel inside the
        // instruction cache): stage s samples at p0 + a_s * k_{s-1}; the sum b_0 k1 + b_1 k2 + ... is built left
        // to right exactly as the reference writes it (RK4: ((k1 + 2k2) + 2k3) + k4, RK3: (2k1 + 3k2) + 4k3).
        float qx, qy, qz;
        {
            const float a1 = rk.half_dt, a2 = order == 3 ? rk.three_quarter_dt : rk.half_dt, a3 = rk.dt;
            const float b0 = order == 3 ? 2.0f : 1.0f, b1 = order == 3 ? 3.0f : 2.0f, b2 = order == 3 ? 4.0f : 2.0f;
            float kx = k1x, ky = k1y, kz = k1z;
            float sx_ = __fmul_rn(k1x, b0), sy_ = __fmul_rn(k1y, b0), sz_ = __fmul_rn(k1z, b0);     // k*1.0f == k
#pragma unroll 1
            for (int st = 1; st < order; st++) {
                const float a = st == 1 ? a1 : (st == 2 ? a2 : a3);
                const float b = st == 1 ? b1 : (st == 2 ? b2 : 1.0f);
                const float ex = axpy(px, a, kx), ey = axpy(py, a, ky), ez = axpy(pz, a, kz);
                if (overflow) evaluate_pow2(g, fnew, INTERP, ex, ey, ez, kx, ky, kz);
                else evaluate_tile<INTERP>(g, fnew, tnew, bx, by, bz, ex, ey, ez, kx, ky, kz);
                sx_ = __fadd_rn(sx_, __fmul_rn(kx, b)); sy_ = __fadd_rn(sy_, __fmul_rn(ky, b)); sz_ = __fadd_rn(sz_, __fmul_rn(kz, b));
            }
            if (order <= 2) { sx_ = kx; sy_ = ky; sz_ = kz; }          // RK1: dt*k1, RK2 (midpoint): dt*k2
            const float h = order == 4 ? rk.dt_over_6 : (order == 3 ? rk.dt_over_9 : rk.dt);
            qx = axpy(px, h, sx_); qy = axpy(py, h, sy_); qz = axpy(pz, h, sz_);
        }
        bool deferred = false;
        if (material) {
            // cell of the advected position (fp32-exact here); out of range reads as solid, NaN compares false -> solid
            bool solid = true;
            if (qx >= 0.0f && qy >= 0.0f && qz >= 0.0f && qx < g.xmaxf && qy < g.ymaxf && qz < g.zmaxf) {
                int i, j, k;
                if (INTERP == 1) { floor_small(__fmul_rn(qx, g.invdxf), i); floor_small(__fmul_rn(qy, g.invdxf), j); floor_small(__fmul_rn(qz, g.invdxf), k); }
                else { i = (int)floorf(__fmul_rn(qx, g.invdxf)); j = (int)floorf(__fmul_rn(qy, g.invdxf)); k = (int)floorf(__fmul_rn(qz, g.invdxf)); }
                const int kl = k - g.k0;
                solid = (kl >= 0 && kl < g.k1 - g.k0) ? material[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)kl)] == GFS_SOLID : false;
            }
            if (solid) {
                atomicAdd(&counters[2], 1ull);
                if (coll.list) {               // listed for k_resolve_collisions, which also bins it; p0 stays for now
                    const unsigned int tk = atomicAdd(coll.count, 1u);
                    if (tk < coll.cap) { coll.list[tk] = make_float4(__int_as_float(r), qx, qy, qz); deferred = true; }
                    else atomicAdd(&counters[3], 1ull);                  // list full: reported, see gfs_stats_t.collision_overflow
                }
                qx = px; qy = py; qz = pz;
            }
        }
        ox[r] = qx; oy[r] = qy; oz[r] = qz;
        if (keys_out && deferred) {            // its velocity still steers the fixed-point scale
            float mm = fmaxf(fabsf(wx), fmaxf(fabsf(wy), fabsf(wz)));
            if (mm < 3.0e38f) m = fmaxf(m, mm);
        }
        if (keys_out && !deferred) {
            // cell key of the advected position in fp32 (exact here, same value as position_key)
            uint32_t key = nkeys;
            if (qx >= 0.0f && qy >= 0.0f && qz >= 0.0f && qx < g.xmaxf && qy < g.ymaxf && qz < g.zmaxf) {
                int i, j, k;
                if (INTERP == 1) { floor_small(__fmul_rn(qx, g.invdxf), i); floor_small(__fmul_rn(qy, g.invdxf), j); floor_small(__fmul_rn(qz, g.invdxf), k); }
                else { i = (int)floorf(__fmul_rn(qx, g.invdxf)); j = (int)floorf(__fmul_rn(qy, g.invdxf)); k = (int)floorf(__fmul_rn(qz, g.invdxf)); }
                if (k >= g.k0 && k < g.k1) key = brick_key(g, i, j, k - g.k0);
                if (MIGRATE) {
                    const int side = k < mg.own_lo ? 0 : (k >= mg.own_hi ? 1 : -1);
                    if (side >= 0) {
                        const unsigned int slot = atomicAdd(mg.count + side, 1u);
                        if (slot < mg.cap && mg.out[side]) {
                            float2 *dst = reinterpret_cast<float2 *>(mg.out[side] + 6 * (size_t)slot);      // 24-byte records: 8-byte aligned
                            dst[0] = make_float2(qx, qy); dst[1] = make_float2(qz, wx); dst[2] = make_float2(wy, wz);
                        }
                        key = nkeys + 1;
                    }
                }
            }
            pend_rank = take_ticket(counts, nkeys, coll.cell_cap, key);
            keys_out[r] = key;
            pend_r = r;
            float mm = fmaxf(fabsf(wx), fmaxf(fabsf(wy), fabsf(wz)));
            if (mm < 3.0e38f) m = fmaxf(m, mm);
        }
    }
    if (pend_r >= 0) rank_out[pend_r] = pend_rank;
    if (keys_out) block_vmax(m, vmax_bits);
}

// ------------------------------------------------------------------------------------------------
// Host-pointer operators (packed float triples in, packed float triples out)
// ------------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void k_sample(Grid g, FieldPtrs f, int interp, int validate, int64_t n, const float *__restrict__ pos, float *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float ox, oy, oz;
    evaluate_any<ARITH>(g, f, interp, pos[3 * r], pos[3 * r + 1], pos[3 * r + 2], ox, oy, oz);
    if (validate) validate3(ox, oy, oz);
    out[3 * r] = ox; out[3 * r + 1] = oy; out[3 * r + 2] = oz;
}

template <int ARITH>
__global__ void k_advect(Grid g, FieldPtrs f, int interp, int order, RkCoef rk, int64_t n, const float *__restrict__ pos, float *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float px = pos[3 * r], py = pos[3 * r + 1], pz = pos[3 * r + 2];
    float k1x, k1y, k1z, ox, oy, oz;
    evaluate_any<ARITH>(g, f, interp, px, py, pz, k1x, k1y, k1z);
    rk_advance<ARITH>(g, f, interp, order, rk, px, py, pz, k1x, k1y, k1z, ox, oy, oz);
    out[3 * r] = ox; out[3 * r + 1] = oy; out[3 * r + 2] = oz;
}

// General ScalarField::addPointValue for arbitrary radius / offset / grid (CLScalarField::addPointValues):
// particle-centric, index bounds exactly as Grid3d::getGridIndexBounds (grid3d.h:350-371), fixed-point
// accumulation as above.  acc[2*node], acc[2*node+1].
template <int ARITH>
__global__ void k_splat_points(SplatParams sp, int vexp, double dx, float offx, float offy, float offz, int ni, int nj, int nk,
                               int64_t n, const float *__restrict__ pos, const float *__restrict__ values,
                               unsigned long long *__restrict__ acc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float q[3] = {__fsub_rn(pos[3 * r], offx), __fsub_rn(pos[3 * r + 1], offy), __fsub_rn(pos[3 * r + 2], offz)};
    float value = values ? values[r] : 1.0f;
    double inv = 1.0 / dx;
    float num_scale = num_scale_f(vexp);
    int lo[3], hi[3];
    const int size[3] = {ni, nj, nk};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int c = cell_floor((double)q[a], inv);
        float cpos = node_pos(c, dx);
        float trans = __fsub_rn(q[a], cpos);
        int gmin = c - (int)fmax(0.0, ceil(__dmul_rn(__dsub_rn(sp.radius, (double)trans), inv)));
        int gmax = c + (int)fmax(0.0, ceil(__dmul_rn(__dadd_rn(__dsub_rn(sp.radius, dx), (double)trans), inv)));
        lo[a] = max(gmin, 0);
        hi[a] = min(gmax, size[a] - 1);
    }
    for (int k = lo[2]; k <= hi[2]; k++) {
        float vz = __fsub_rn(node_pos(k, dx), q[2]);
        for (int j = lo[1]; j <= hi[1]; j++) {
            float vy = __fsub_rn(node_pos(j, dx), q[1]);
            for (int i = lo[0]; i <= hi[0]; i++) {
                float vx = __fsub_rn(node_pos(i, dx), q[0]);
                float d2 = dist2(vx, vy, vz);
                if ((double)d2 < sp.rsq) {
                    size_t node = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)k);
                    long long wn, ww;
                    if (ARITH == 1) {
                        double w = kernel_weight_exact(sp, (double)d2);
                        ww = __double2ll_rn(w * kWeightScaleD);
                        wn = __double2ll_rn(__dmul_rn(w, (double)value) * (double)num_scale);
                    } else {
                        float w = kernel_weight_fast<false>(sp, d2);
                        ww = __float2ll_rn(w * kWeightScaleF);
                        wn = __float2ll_rn((w * value) * num_scale);
                    }
                    atomicAdd(acc + 2 * node, (unsigned long long)wn);
                    atomicAdd(acc + 2 * node + 1, (unsigned long long)ww);
                }
            }
        }
    }
}

// field[n] (+)= fixed-point num, weight[n] (+)= fixed-point weight  -- epilogue of gfs_add_point_values
__global__ void k_splat_points_store(int64_t count, int vexp, const unsigned long long *__restrict__ acc,
                                     float *__restrict__ field, float *__restrict__ weight, int accumulate,
                                     int use_threshold, float threshold) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= count) return;
    if (use_threshold && field[n] > threshold) return;       // saturated before this batch: see gfs_add_points
    float nf = (float)((double)(long long)acc[2 * n] * inv_num_scale_d(vexp));
    float wf = (float)((double)(long long)acc[2 * n + 1] * (1.0 / kWeightScaleD));
    field[n] = accumulate ? __fadd_rn(field[n], nf) : nf;
    if (weight) weight[n] = accumulate ? __fadd_rn(weight[n], wf) : wf;
}

// AoS MarkerParticle_t <-> SoA
__global__ void k_aos_to_soa(int64_t n, const float *__restrict__ aos, float *x, float *y, float *z, float *vx, float *vy, float *vz, int32_t *tag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float *p = aos + 6 * r;
    x[r] = p[0]; y[r] = p[1]; z[r] = p[2]; vx[r] = p[3]; vy[r] = p[4]; vz[r] = p[5];
    tag[r] = (int32_t)r;
}
__global__ void k_soa_to_aos(int64_t n, const float *x, const float *y, const float *z, const float *vx, const float *vy, const float *vz, float *__restrict__ aos) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float *p = aos + 6 * r;
    p[0] = x[r]; p[1] = y[r]; p[2] = z[r]; p[3] = vx[r]; p[4] = vy[r]; p[5] = vz[r];
}

// Up to 16 layer-range copies (or 64-bit integer adds) in one launch: the slab exchange packs / unpacks all its
// arrays with one kernel instead of one memcpy each.  blockIdx.y selects the descriptor.
struct CopyBatch {
    int n;
    const unsigned char *src[16];
    unsigned char *dst[16];
    long long bytes[16];
    int add[16];                  // != 0: dst (uint64[]) += src (uint64[])
};

__global__ void __launch_bounds__(256) k_copy_batch(CopyBatch cb) {
    const int d = blockIdx.y;
    if (d >= cb.n) return;
    const unsigned char *__restrict__ src = cb.src[d];
    unsigned char *__restrict__ dst = cb.dst[d];
    const long long bytes = cb.bytes[d];
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cb.add[d]) {
        const unsigned long long *s8 = reinterpret_cast<const unsigned long long *>(src);
        unsigned long long *d8 = reinterpret_cast<unsigned long long *>(dst);
        for (long long t = t0; t < bytes / 8; t += stride) d8[t] += s8[t];
    } else if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes) & 15) == 0) {
        const uint4 *s16 = reinterpret_cast<const uint4 *>(src);
        uint4 *d16 = reinterpret_cast<uint4 *>(dst);
        for (long long t = t0; t < bytes / 16; t += stride) d16[t] = s16[t];
    } else {
        for (long long t = t0; t < bytes; t += stride) dst[t] = src[t];
    }
}

// ---- peer-memory exchange (CUDA IPC over NVLink): flags written by the neighbour GPU ---------------------------
// flag words live in the receiver's memory; the sender's k_signal runs after its copy kernel in stream order.
__global__ void k_signal(volatile unsigned int *peer_flag, unsigned int seq, volatile unsigned int *peer_aux, const unsigned int *aux_src, int naux) {
    for (int i = 0; i < naux; i++) peer_aux[i] = aux_src[i];
    __threadfence_system();
    *peer_flag = seq;
    __threadfence_system();
}

// spin until *flag >= seq (written over NVLink by the neighbour); gives up after ~4 s and raises *error
__device__ __forceinline__ bool wait_flag(const volatile unsigned int *flag, unsigned int seq, unsigned int *error, long long timeout) {
    const long long t0 = clock64();
    while ((int)(*flag - seq) < 0) {
        if (clock64() - t0 > timeout) { atomicExch(error, 1u); return false; }
        __nanosleep(200);
    }
    __threadfence_system();
    return true;
}

// the batched layer copy of k_copy_batch, preceded by the wait for the neighbour's data
__global__ void __launch_bounds__(256) k_copy_batch_wait(CopyBatch cb, const volatile unsigned int *flag, unsigned int seq, unsigned int *error, long long timeout) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = wait_flag(flag, seq, error, timeout) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    const int d = blockIdx.y;
    if (d >= cb.n) return;
    const unsigned char *__restrict__ src = cb.src[d];
    unsigned char *__restrict__ dst = cb.dst[d];
    const long long bytes = cb.bytes[d];
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cb.add[d]) {
        const unsigned long long *s8 = reinterpret_cast<const unsigned long long *>(src);
        unsigned long long *d8 = reinterpret_cast<unsigned long long *>(dst);
        for (long long t = t0; t < bytes / 8; t += stride) d8[t] += s8[t];
    } else if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes) & 15) == 0) {
        const uint4 *s16 = reinterpret_cast<const uint4 *>(src);
        uint4 *d16 = reinterpret_cast<uint4 *>(dst);
        for (long long t = t0; t < bytes / 16; t += stride) d16[t] = s16[t];
    } else {
        for (long long t = t0; t < bytes; t += stride) dst[t] = src[t];
    }
}

// the wait of k_copy_batch_wait alone, in ONE thread: for groups whose slabs share a GPU (tests), where thousands of
// spinning CTAs of one context would keep the other context's push kernel off the SMs for ever
__global__ void k_wait_flag(const volatile unsigned int *flag, unsigned int seq, unsigned int *error, long long timeout) {
    wait_flag(flag, seq, error, timeout);
}

// wait for both neighbours' particle flags, then publish {my kept/down/up counts, arrivals from down, arrivals from up}
// to pinned host memory: the one word set the host reads per substep
__global__ void k_gather_counts(const volatile unsigned int *flag_down, const volatile unsigned int *flag_up, unsigned int seq,
                                const unsigned int *split_counters, const volatile unsigned int *in_down, const volatile unsigned int *in_up,
                                unsigned int *host_out, unsigned int *error, long long timeout) {
    bool ok = true;
    if (flag_down) ok = wait_flag(flag_down, seq, error, timeout) && ok;
    if (flag_up) ok = wait_flag(flag_up, seq, error, timeout) && ok;
    host_out[0] = split_counters[0]; host_out[1] = split_counters[1]; host_out[2] = split_counters[2];
    host_out[3] = (flag_down && ok) ? *in_down : 0u;
    host_out[4] = (flag_up && ok) ? *in_up : 0u;
    // bit 0: this kernel's own wait timed out; bit 1: an earlier wait of the substep did (k_copy_batch_wait, k_allmax raise
    // *error in stream order before this kernel runs) -- the host reads the word in gfs_comm_migrate_finish
    host_out[5] = (ok ? 0u : 1u) | (*(volatile unsigned int *)error ? 2u : 0u);
}

// acc[first .. first+count) += src  (integer adds: the slab partial sums merge bit-exactly in any order)
__global__ void k_add_u64(long long count, unsigned long long *__restrict__ dst, const unsigned long long *__restrict__ src) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) dst[t] += src[t];
}

// Split the resident particles by the cell layer of their position: k < k_lo -> `down` (AoS, 6 floats), k >= k_hi ->
// `up`, the rest compacted into the other SoA buffer.  counters[0..2] = kept, down, up.  (A NaN position has
// no layer: it stays.)  Order inside each output is unspecified -- the next substep re-bins anyway.
__global__ void __launch_bounds__(256) k_split_by_layer(Grid g, int64_t n, int k_lo, int k_hi, int cap,
                                 const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                                 const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                                 const int32_t *__restrict__ tag,
                                 float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
                                 float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz, int32_t *__restrict__ otag,
                                 float *__restrict__ down, float *__restrict__ up, unsigned int *__restrict__ counters) {
    // one global atomic per block and destination (a per-thread ticket on a single counter serialises 10^8 atomics);
    // slots inside a block follow thread order, so the stayers keep their (nearly sorted) order
    __shared__ unsigned int s_warp[8][3], s_base[3];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int dest = -1;
    float pz = 0.0f;
    if (r < n) {
        pz = z[r];
        int k = cell_floor((double)pz, g.invdx);
        dest = (pz == pz) ? (k < k_lo ? 1 : (k >= k_hi ? 2 : 0)) : 0;
    }
    unsigned int m[3], before = 0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        m[d] = __ballot_sync(0xffffffffu, dest == d);
        if (dest == d) before = __popc(m[d] & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp][d] = __popc(m[d]);
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned int tot = 0;
        for (int w = 0; w < 8; w++) { unsigned int c = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = tot; tot += c; }
        s_base[threadIdx.x] = tot ? atomicAdd(&counters[threadIdx.x], tot) : 0u;
    }
    __syncthreads();
    if (dest < 0) return;
    const unsigned int slot = s_base[dest] + s_warp[warp][dest] + before;
    if (dest == 0) {
        ox[slot] = x[r]; oy[slot] = y[r]; oz[slot] = pz; ovx[slot] = vx[r]; ovy[slot] = vy[r]; ovz[slot] = vz[r]; otag[slot] = tag[r];
    } else if ((int)slot < cap) {
        float *o = (dest == 1 ? down : up) + 6 * (size_t)slot;
        o[0] = x[r]; o[1] = y[r]; o[2] = pz; o[3] = vx[r]; o[4] = vy[r]; o[5] = vz[r];
    }
}

__global__ void k_append_aos(int64_t n, int64_t at, const float *__restrict__ aos, float *x, float *y, float *z,
                             float *vx, float *vy, float *vz, int32_t *tag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float *p = aos + 6 * r;
    x[at + r] = p[0]; y[at + r] = p[1]; z[at + r] = p[2]; vx[at + r] = p[3]; vy[at + r] = p[4]; vz[at + r] = p[5];
    tag[at + r] = -1;
}

// arrivals of the fused migration: AoS -> SoA at the end of the arrays, binned for the next counting sort like the
// G2P epilogue bins the residents
__global__ void __launch_bounds__(256) k_append_bin(Grid g, uint32_t nkeys, int64_t n, int64_t at, const float *__restrict__ aos,
                             float *x, float *y, float *z, float *vx, float *vy, float *vz, int32_t *tag,
                             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ rank_out, uint32_t *__restrict__ counts, KeyRange kr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float2 *p = reinterpret_cast<const float2 *>(aos + 6 * r);
    const float2 a = p[0], b = p[1], c = p[2];
    x[at + r] = a.x; y[at + r] = a.y; z[at + r] = b.x; vx[at + r] = b.y; vy[at + r] = c.x; vz[at + r] = c.y;
    tag[at + r] = -1;
    const uint32_t key = clamp_key(position_key(g, nkeys, a.x, a.y, b.x), nkeys, kr);
    keys_out[at + r] = key;
    rank_out[at + r] = atomicAdd(counts + key, 1u);
}

// all-ranks maximum of one 32-bit word over peer memory: every rank stores {seq, value} into its slot of every rank's
// table (NVLink stores, one thread per peer), then waits until all slots of its own table carry seq.  Tables are double
// buffered by the parity of seq.  value is compared as an unsigned integer (bit patterns of non-negative floats order
// like the floats).  One CTA of >= world threads.
struct AllMaxPeers { unsigned long long *table[16]; };
// mode 0: post + wait; 1: post only (this rank's value is final -- right after its G2P -- long before anyone needs the
// maximum); 2: wait only (just before the splat).  Splitting takes the all-ranks round trip off the critical path.
__global__ void k_allmax(AllMaxPeers peers, int rank, int world, unsigned int seq, unsigned int *value, unsigned int *error, long long timeout, int mode) {
    __shared__ unsigned int s_max;
    const int t = threadIdx.x;
    if (t == 0) s_max = 0u;
    __syncthreads();
    const unsigned int mine = *value;
    const size_t base = (size_t)(seq & 1u) * 16;
    if (t < world && mode != 2) {
        volatile unsigned long long *slot = peers.table[t] + base + rank;
        *slot = ((unsigned long long)seq << 32) | mine;
        __threadfence_system();
    }
    if (mode == 1) return;
    if (t < world) {
        const volatile unsigned long long *in = peers.table[rank] + base + t;
        const long long t0 = clock64();
        unsigned long long v;
        bool ok = true;
        while ((unsigned int)((v = *in) >> 32) != seq) {
            if (clock64() - t0 > timeout) { atomicExch(error, 1u); ok = false; break; }
            __nanosleep(100);
        }
        if (ok) atomicMax(&s_max, (unsigned int)v);
    }
    __syncthreads();
    if (t == 0) *value = s_max;
}

// ---- order-independent 64-bit state hashes (verification hook, gfs_state_hash) ---------------------------------
// Each element contributes splitmix64(position-in-the-GLOBAL-array, value bits); contributions are summed mod 2^64, so
// the hash of a grid does not depend on which rank owns which layers, and the hash of the particle set (a sum over
// particles of a hash of their six words) does not depend on their order or distribution.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ void hash_commit(unsigned long long h, unsigned long long *out) {
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}
// rows of `ni` elements at pitch `pitch` (elements), `rows` rows starting at global row `row0`; elem_bytes 1 or 4
__global__ void __launch_bounds__(256) k_hash_grid(const void *__restrict__ base, int elem_bytes, long long ni, long long pitch, long long row0,
                                                   long long rows, unsigned long long salt, unsigned long long *out) {
    unsigned long long h = 0;
    const long long total = ni * rows;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / ni, i = t - r * ni;
        const unsigned long long v = elem_bytes == 1 ? (unsigned long long)((const uint8_t *)base)[r * pitch + i]
                                                      : (unsigned long long)((const uint32_t *)base)[r * pitch + i];
        h += mix64(mix64((unsigned long long)((row0 + r) * ni + i) ^ salt) ^ v);
    }
    hash_commit(h, out);
}
__global__ void __launch_bounds__(256) k_hash_particles(int64_t n, const uint32_t *__restrict__ x, const uint32_t *__restrict__ y, const uint32_t *__restrict__ z,
                                                        const uint32_t *__restrict__ vx, const uint32_t *__restrict__ vy, const uint32_t *__restrict__ vz,
                                                        unsigned long long *out) {
    unsigned long long h = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long a = mix64(((unsigned long long)y[r] << 32) | x[r]);
        a = mix64(a ^ (((unsigned long long)vx[r] << 32) | z[r]));
        h += mix64(a ^ (((unsigned long long)vz[r] << 32) | vy[r]));
    }
    hash_commit(h, out);
}

// after the exclusive scan of the key range [lo, hi): the end marker of its last brick and the three tail bins
// (out-of-grid, dead, end) -- what a scan over the whole table would have left there
__global__ void k_scan_tail(const uint32_t *__restrict__ counts, int32_t *__restrict__ cell_start, uint32_t lo, uint32_t hi, uint32_t nkeys) {
    const int32_t total = hi > lo ? cell_start[hi - 1] + (int32_t)counts[hi - 1] : 0;
    if (hi < nkeys) cell_start[hi] = total;
    cell_start[nkeys] = total;
    cell_start[nkeys + 1] = total + (int32_t)counts[nkeys];
    cell_start[nkeys + 2] = total + (int32_t)counts[nkeys] + (int32_t)counts[nkeys + 1];
}

__global__ void k_border_solid(Grid g, uint8_t *__restrict__ material) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, kl = blockIdx.z;
    if (i >= g.I) return;
    int k = kl + g.k0;
    bool border = i == 0 || j == 0 || k == 0 || i == g.I - 1 || j == g.J - 1 || k == g.K - 1;
    material[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)kl)] = border ? GFS_SOLID : GFS_AIR;
}

}  // namespace gfs
