// gfs_pressure.cuh -- stages 6-8 of FluidSimulation::_stepFluid on the resident grid:
//
//   k_body_force        _applyConstantBodyForces                      src/fluidsimulation.cpp:2765-2805
//   k_press_setup ...   PressureSolver::solve (MICCG(0))              src/pressuresolver.cpp:116-505
//   k_apply_pressure    _applyPressureToVelocityField                 src/fluidsimulation.cpp:2895-3061
//
// The solver is the REFERENCE'S algorithm, operation for operation, not a GPU-friendlier substitute: the modified
// incomplete Cholesky factor and its two triangular solves are sequential in the reference (cells in ascending linear
// index, each depending on its -x, -y, -z neighbours), but the result of every cell is a fixed expression of its three
// predecessors, so ANY schedule that respects that dependency produces the same bits.  Here the grid is cut into
// 16 x 8 x 4 tiles, one warp per tile: lane (j, k) walks its row along x, row (j, k) one step behind rows (j-1, k) and
// (j, k-1), whose values arrive by shuffle; tiles are handed out through a ticket in wavefront order (tx + ty + tz) and
// wait on their three predecessor tiles' completion flags (a ticket is only taken by a running warp and predecessors
// have smaller tickets, so the wait cannot deadlock).  With every product, sum, division and square root in the
// reference's order and no contraction (__dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn), the preconditioner and both
// substitutions are bit-identical to the CPU's; the only freedom left is the summation order of the two dot products per
// iteration (fixed here, so runs repeat bit for bit), which perturbs alpha and beta in the last place.  The CG
// trajectory, the iteration count and the float pressure grid therefore match the reference's
// (tests/test_gpu_pressure.py).
//
// All vectors are dense over the cells (0 outside fluid cells, where nothing is ever written): a neighbour that is not
// fluid contributes an exact zero, which is what the reference's "vidx == -1" branches do.
#pragma once
#include "gfs_kernels.cuh"

namespace gfs {

constexpr int kPressBlocks = 592;                 // grid of the flat kernels = partial sums per reduction (4 x 148)
constexpr int kPressThreads = 256;
constexpr int kTileX = 16, kTileY = 8, kTileZ = 4;
constexpr long long kPressSentinel = (long long)0xFFF8C0DEC0DEC0DEull;    // "not produced yet" (k_press_subst_df): a NaN pattern no computation yields

// flags byte per cell: bits 0-2 diag (non-solid neighbours), 3 plusi, 4 plusj, 5 plusk (MatrixCell, pressuresolver.h), 6 fluid
constexpr uint8_t kPfPlusI = 8, kPfPlusJ = 16, kPfPlusK = 32, kPfFluid = 64;

struct PressSys {
    const uint8_t *material;
    uint8_t *flags;
    double *r, *z, *s, *p, *q, *precon;       // residual, auxillary, search, pressure, forward-solve temporary, MIC(0) diagonal
    double *partial;                          // [3][kPressBlocks]: 0 = |r| max, 1 = dot(z, s), 2 = dot(z, r)
    double *sigma;                            // [2]: double buffered by iteration parity
    int *state;                               // [0] done, [1] iterations (reference's iterationNumber; -1 = rhs below tolerance), [2] spare
    double *resid;                            // [1] last max |r|
    unsigned long long *ticket;               // [3], one per sweep mode
    unsigned int *tile_done;                  // [ntiles] epoch of the last sweep that completed the tile
    const int *order;                         // tiles in wavefront order
    float *pressure;                          // float grid of _updatePressureGrid
    int I, J, K, ntx, nty, ntz, ntiles;
    long long cells;
    double scale;                             // dt / (density dx^2)
    double tol;
    long long *trace;                         // debugging (option 13): per tile {ticket, static done, halo there, steps done} in ns, or null
};

__device__ __forceinline__ int press_mat(const PressSys &S, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= S.I || j >= S.J || k >= S.K) return GFS_SOLID;
    return S.material[(size_t)i + (size_t)S.I * ((size_t)j + (size_t)S.J * (size_t)k)];
}

// The data-flow sweeps communicate through plain global memory: the producer's store and the consumer's polling load are
// STRONG (relaxed, gpu scope) operations.  A weak load -- even ld.global.cg inside asm volatile -- lets ptxas assume
// nobody else writes the location and turn "while (v == sentinel) v = load(p)" into a single load (it did).
__device__ __forceinline__ double ld_poll(const double *p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_publish(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// deterministic block reductions (fixed tree): every block that reduces the same input gets the same bits
__device__ __forceinline__ double block_sum(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kPressThreads / 32; i++) t += sh[i];
    return t;
}
__device__ __forceinline__ double block_max(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = sh[0];
    for (int i = 1; i < kPressThreads / 32; i++) t = fmax(t, sh[i]);
    return t;
}
__device__ __forceinline__ double reduce_partials_sum(const double *part, double *sh) {
    double v = 0.0;
    for (int i = threadIdx.x; i < kPressBlocks; i += kPressThreads) v += part[i];
    return block_sum(v, sh);
}
__device__ __forceinline__ double reduce_partials_max(const double *part, double *sh) {
    double v = 0.0;
    for (int i = threadIdx.x; i < kPressBlocks; i += kPressThreads) v = fmax(v, part[i]);
    return block_max(v, sh);
}

// _calculateNegativeDivergenceVector (:164-211) + _calculateMatrixCoefficients (:213-250); r = b
__global__ void __launch_bounds__(kPressThreads) k_press_setup(Grid g, FieldPtrs f, PressSys S, double dx, int sentinel) {
    __shared__ double sh[kPressThreads / 32];
    const double scale = (double)(1.0f / (float)dx);
    const float fscale = (float)scale;
    double mx = 0.0;
    for (long long c = (long long)blockIdx.x * kPressThreads + threadIdx.x; c < S.cells; c += (long long)kPressBlocks * kPressThreads) {
        const int i = (int)(c % S.I), j = (int)((c / S.I) % S.J), k = (int)(c / ((long long)S.I * S.J));
        uint8_t fl = 0;
        double b = 0.0;
        if (S.material[c] == GFS_FLUID) {
            const int mi0 = press_mat(S, i - 1, j, k), mi1 = press_mat(S, i + 1, j, k);
            const int mj0 = press_mat(S, i, j - 1, k), mj1 = press_mat(S, i, j + 1, k);
            const int mk0 = press_mat(S, i, j, k - 1), mk1 = press_mat(S, i, j, k + 1);
            const float u0 = f.c[0][(size_t)i + (size_t)g.pitch[0] * ((size_t)j + (size_t)g.J * (size_t)k)];
            const float u1 = f.c[0][(size_t)i + 1 + (size_t)g.pitch[0] * ((size_t)j + (size_t)g.J * (size_t)k)];
            const float v0 = f.c[1][(size_t)i + (size_t)g.pitch[1] * ((size_t)j + (size_t)(g.J + 1) * (size_t)k)];
            const float v1 = f.c[1][(size_t)i + (size_t)g.pitch[1] * ((size_t)j + 1 + (size_t)(g.J + 1) * (size_t)k)];
            const float w0 = f.c[2][(size_t)i + (size_t)g.pitch[2] * ((size_t)j + (size_t)g.J * (size_t)k)];
            const float w1 = f.c[2][(size_t)i + (size_t)g.pitch[2] * ((size_t)j + (size_t)g.J * ((size_t)k + 1))];
            const float sum = __fsub_rn(__fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(u1, u0), v1), v0), w1), w0);
            b = __dmul_rn(-scale, (double)sum);
            if (mi0 == GFS_SOLID) b = __dsub_rn(b, (double)__fmul_rn(fscale, u0));        // "- usolid" with usolid = 0.0f is exact
            if (mi1 == GFS_SOLID) b = __dadd_rn(b, (double)__fmul_rn(fscale, u1));
            if (mj0 == GFS_SOLID) b = __dsub_rn(b, (double)__fmul_rn(fscale, v0));
            if (mj1 == GFS_SOLID) b = __dadd_rn(b, (double)__fmul_rn(fscale, v1));
            if (mk0 == GFS_SOLID) b = __dsub_rn(b, (double)__fmul_rn(fscale, w0));
            if (mk1 == GFS_SOLID) b = __dadd_rn(b, (double)__fmul_rn(fscale, w1));
            const int n = (mi0 != GFS_SOLID) + (mi1 != GFS_SOLID) + (mj0 != GFS_SOLID) + (mj1 != GFS_SOLID) + (mk0 != GFS_SOLID) + (mk1 != GFS_SOLID);
            fl = (uint8_t)(kPfFluid | n | (mi1 == GFS_FLUID ? kPfPlusI : 0) | (mj1 == GFS_FLUID ? kPfPlusJ : 0) | (mk1 == GFS_FLUID ? kPfPlusK : 0));
            mx = fmax(mx, fabs(b));
        }
        S.flags[c] = fl;
        S.r[c] = b;
        if (sentinel && fl) { S.q[c] = __longlong_as_double(kPressSentinel); S.z[c] = __longlong_as_double(kPressSentinel); }
    }
    mx = block_max(mx, sh);
    if (threadIdx.x == 0) S.partial[blockIdx.x] = mx;
}

// one block.  it < 0: the "b.absMaxCoeff() < tolerance" early return of solve (:127-129); it >= 0: the convergence test of
// iteration `it` (:478-481).  Also re-arms the tickets of the two substitution sweeps.
__global__ void __launch_bounds__(kPressThreads) k_press_check(PressSys S, int it) {
    __shared__ double sh[kPressThreads / 32];
    if (S.state[0]) return;
    const double mx = reduce_partials_max(S.partial, sh);
    if (threadIdx.x == 0) {
        S.resid[0] = mx;
        if (mx < S.tol) { S.state[0] = 1; S.state[1] = it; }
        S.ticket[1] = 0ull; S.ticket[2] = 0ull;
    }
}

// MODE 0: _calculatePreconditionerVector (:252-310).  MODE 1 / 2: the forward / backward substitution of
// _applyPreconditioner (:312-390): q from r, then z from q.
template <int MODE>
__global__ void __launch_bounds__(kPressThreads) k_press_sweep(PressSys S, unsigned int epoch) {
    __shared__ double halo_s[kPressThreads / 32][(kTileZ + kTileY) * kTileX];
    if (S.state[0]) return;
    constexpr bool REV = MODE == 2;
    const int lane = threadIdx.x & 31, lj = lane & 7, lk = lane >> 3;
    const int a = REV ? kTileY - 1 - lj : lj, b = REV ? kTileZ - 1 - lk : lk;     // rows ahead of this one in y and z
    double *halo = halo_s[threadIdx.x >> 5];
    double *dyn = MODE == 0 ? S.precon : (MODE == 1 ? S.q : S.z);                 // the vector this sweep produces
    const double scale = S.scale, negscale = -S.scale;
    const size_t sy = (size_t)S.I, sz = (size_t)S.I * (size_t)S.J;
    const int d = REV ? 1 : -1;

    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(S.ticket + MODE, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned long long)S.ntiles) break;
        const int tile = S.order[REV ? S.ntiles - 1 - (int)t : (int)t];
        const int tx = tile % S.ntx, ty = (tile / S.ntx) % S.nty, tz = tile / (S.ntx * S.nty);
        const int i0 = tx * kTileX, j = ty * kTileY + lj, k = tz * kTileZ + lk;
        const bool row_in = j < S.J && k < S.K;
        const size_t row = (size_t)j * sy + (size_t)k * sz;

        // which cells of my row are fluid
        unsigned fl = 0;
        if (row_in)
            for (int ii = 0; ii < kTileX; ii++)
                if (i0 + ii < S.I && (S.flags[row + i0 + ii] & kPfFluid)) fl |= 1u << ii;
        const bool any = __any_sync(0xffffffffu, fl != 0);

        if (any) {
            // predecessor tiles
            if (lane < 3) {
                const int px = tx + (lane == 0 ? d : 0), py = ty + (lane == 1 ? d : 0), pz = tz + (lane == 2 ? d : 0);
                if (px >= 0 && py >= 0 && pz >= 0 && px < S.ntx && py < S.nty && pz < S.ntz) {
                    const volatile unsigned int *flag = S.tile_done + (px + S.ntx * (py + S.nty * pz));
                    while (*flag != epoch) { }
                }
            }
            __syncwarp();
            __threadfence();
            // halo: the neighbouring tiles' row j0-1 (j0+8 backward) for my 4 planes, plane k0-1 (k0+4) for my 8 rows
            for (int h = lane; h < (kTileZ + kTileY) * kTileX; h += 32) {
                const int hr = h / kTileX, ii = h % kTileX;
                int hj, hk;
                if (hr < kTileZ) { hj = ty * kTileY + (REV ? kTileY : -1); hk = tz * kTileZ + hr; }
                else { hj = ty * kTileY + (hr - kTileZ); hk = tz * kTileZ + (REV ? kTileZ : -1); }
                double v = 0.0;
                if (hj >= 0 && hk >= 0 && hj < S.J && hk < S.K && i0 + ii < S.I)
                    v = __ldcg(dyn + (size_t)(i0 + ii) + (size_t)hj * sy + (size_t)hk * sz);
                halo[h] = v;
            }
            const int ix = REV ? i0 + kTileX : i0 - 1;                         // the cell before my first one along x
            double mine = 0.0;                                                   // my value of the previous step (0: none / not fluid)
            double xh = 0.0;                                                     // the neighbouring tile's value before my first cell
            if (row_in && ix >= 0 && ix < S.I) xh = __ldcg(dyn + row + ix);
            double pc_prev = 0.0;                                                // MODE 1: precon of the previous cell along x
            uint8_t fx_prev = 0;                                                 // MODE 0: flags of the previous cell along x
            if (MODE == 1 && row_in && ix >= 0 && ix < S.I) pc_prev = S.precon[row + ix];
            if (MODE == 0 && row_in && ix >= 0 && ix < S.I) fx_prev = S.flags[row + ix];
            __syncwarp();

            for (int step = 0; step < kTileX + kTileY + kTileZ - 2; step++) {
                const int li = step - a - b;
                const bool active = (unsigned)li < (unsigned)kTileX;
                const int ii = REV ? kTileX - 1 - li : li;
                double from_y = REV ? __shfl_down_sync(0xffffffffu, mine, 1) : __shfl_up_sync(0xffffffffu, mine, 1);
                double from_z = REV ? __shfl_down_sync(0xffffffffu, mine, 8) : __shfl_up_sync(0xffffffffu, mine, 8);
                double val = 0.0;
                if (active) {
                    if (a == 0) from_y = halo[lk * kTileX + ii];
                    if (b == 0) from_z = halo[(kTileZ + lj) * kTileX + ii];
                    const double from_x = li == 0 ? xh : mine;
                    const size_t c = row + i0 + ii;
                    if ((fl >> ii) & 1u) {
                        const bool y_in = REV ? j + 1 < S.J : j > 0, z_in = REV ? k + 1 < S.K : k > 0;
                        if (MODE == 0) {
                            const uint8_t f0 = S.flags[c];
                            const uint8_t fy = y_in ? S.flags[c - sy] : 0, fz = z_in ? S.flags[c - sz] : 0;
                            const double diag = __dmul_rn((double)(f0 & 7), scale);
                            const bool ex = fx_prev & kPfFluid, ey = fy & kPfFluid, ez = fz & kPfFluid;
                            const double pi_x = ex ? __dmul_rn((fx_prev & kPfPlusI) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pi_y = ey ? __dmul_rn((fy & kPfPlusI) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pi_z = ez ? __dmul_rn((fz & kPfPlusI) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pj_x = ex ? __dmul_rn((fx_prev & kPfPlusJ) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pj_y = ey ? __dmul_rn((fy & kPfPlusJ) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pj_z = ez ? __dmul_rn((fz & kPfPlusJ) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pk_x = ex ? __dmul_rn((fx_prev & kPfPlusK) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pk_y = ey ? __dmul_rn((fy & kPfPlusK) ? 1.0 : 0.0, negscale) : 0.0;
                            const double pk_z = ez ? __dmul_rn((fz & kPfPlusK) ? 1.0 : 0.0, negscale) : 0.0;
                            const double v1 = __dmul_rn(pi_x, from_x), v2 = __dmul_rn(pj_y, from_y), v3 = __dmul_rn(pk_z, from_z);
                            const double v4 = __dmul_rn(from_x, from_x), v5 = __dmul_rn(from_y, from_y), v6 = __dmul_rn(from_z, from_z);
                            const double ta = __dmul_rn(__dmul_rn(pi_x, __dadd_rn(pj_x, pk_x)), v4);
                            const double tb = __dmul_rn(__dmul_rn(pj_y, __dadd_rn(pi_y, pk_y)), v5);
                            const double tc = __dmul_rn(__dmul_rn(pk_z, __dadd_rn(pi_z, pj_z)), v6);
                            double e = __dsub_rn(__dsub_rn(__dsub_rn(diag, __dmul_rn(v1, v1)), __dmul_rn(v2, v2)), __dmul_rn(v3, v3));
                            e = __dsub_rn(e, __dmul_rn(0.97, __dadd_rn(__dadd_rn(ta, tb), tc)));
                            if (e < __dmul_rn(0.25, diag)) e = diag;
                            if (fabs(e) > 10e-9) val = __ddiv_rn(1.0, __dsqrt_rn(e));
                            fx_prev = f0;
                        } else if (MODE == 1) {
                            const double pc = S.precon[c];
                            const double py = y_in ? S.precon[c - sy] : 0.0, pz = z_in ? S.precon[c - sz] : 0.0;
                            double tt = S.r[c];
                            tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, pc_prev), from_x));
                            tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, py), from_y));
                            tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, pz), from_z));
                            val = __dmul_rn(tt, pc);
                            pc_prev = pc;
                        } else {
                            const double pc = S.precon[c];
                            const double np = __dmul_rn(negscale, pc);
                            double tt = S.q[c];
                            tt = __dsub_rn(tt, __dmul_rn(np, from_x));
                            tt = __dsub_rn(tt, __dmul_rn(np, from_y));
                            tt = __dsub_rn(tt, __dmul_rn(np, from_z));
                            val = __dmul_rn(tt, pc);
                        }
                        dyn[c] = val;
                    } else {
                        if (MODE == 0) fx_prev = 0;                              // not fluid: flags 0, precon 0
                        if (MODE == 1) pc_prev = 0.0;
                    }
                }
                mine = val;
            }
            __threadfence();
        }
        __syncwarp();
        if (lane == 0) *(volatile unsigned int *)(S.tile_done + tile) = epoch;
    }
}

// The two substitutions of every CG iteration, same schedule and same arithmetic as k_press_sweep<1> / <2>, built for
// latency: the first version read r / precon from global memory inside the 26 dependent steps of a tile (1.8-2.2 ms per
// sweep at 256^3, ncu).  Here everything a tile needs that does NOT depend on its predecessors -- its own rows of the
// input vector and of the MIC(0) diagonal, the diagonal's halo rows, the fluid masks -- is staged into shared memory
// with coalesced loads BEFORE the warp waits for its predecessor tiles, so the wait hides the loads and the step loop
// touches only shared memory, shuffles and the one store per cell.
constexpr int kSubstWarps = 3;                      // 96-thread CTAs: 37.5 KB of static shared memory each, 5 per SM
constexpr int kRowPitch = 19;                       // doubles.  Lane (lj, lk) reads word 19 (lj + 8 lk) + step - lj - lk: the bank pair
                                                    // (2 lj + 7 lk + step) mod 16 is hit by exactly two lanes -- the minimum for 64-bit reads

struct SubstSmem {
    double in[32][kRowPitch];                       // r (forward) or q (backward), my 32 rows
    double pc[32][kRowPitch];                       // precon, my 32 rows
    double hp[kTileZ + kTileY][kTileX];             // forward only: precon of the predecessor tiles' boundary rows
    double hd[kTileZ + kTileY][kTileX];             // the produced vector on the predecessor tiles' boundary rows
};

template <bool REV>
__global__ void __launch_bounds__(kSubstWarps * 32) k_press_subst(PressSys S, unsigned int epoch) {
    __shared__ SubstSmem sm_all[kSubstWarps];
    if (S.state[0]) return;
    SubstSmem &sm = sm_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, lj = lane & 7, lk = lane >> 3;
    const int a = REV ? kTileY - 1 - lj : lj, b = REV ? kTileZ - 1 - lk : lk;
    const double *__restrict__ in = REV ? S.q : S.r;
    double *dyn = REV ? S.z : S.q;
    const double negscale = -S.scale;
    const size_t sy = (size_t)S.I, sz = (size_t)S.I * (size_t)S.J;
    const int d = REV ? 1 : -1;
    const int ly = REV ? lane + 1 : lane - 1, lz = REV ? lane + 8 : lane - 8;       // the lanes holding my y / z predecessor rows
    constexpr int MODE = REV ? 2 : 1;

    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(S.ticket + MODE, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned long long)S.ntiles) break;
        const int tile = S.order[REV ? S.ntiles - 1 - (int)t : (int)t];
        const int tx = tile % S.ntx, ty = (tile / S.ntx) % S.nty, tz = tile / (S.ntx * S.nty);
        const int i0 = tx * kTileX, j0 = ty * kTileY, k0 = tz * kTileZ;
        const int j = j0 + lj, k = k0 + lk;
        const bool row_in = j < S.J && k < S.K;
        const size_t row = (size_t)j * sy + (size_t)k * sz;

        // ---- static part, before the wait: fluid masks, my rows of the input vector and of the diagonal, the diagonal's halo
        unsigned fl = 0;
        if (row_in)
            for (int ii = 0; ii < kTileX; ii++)
                if (i0 + ii < S.I && (S.flags[row + i0 + ii] & kPfFluid)) fl |= 1u << ii;
        const bool any = __any_sync(0xffffffffu, fl != 0);
        const int ix = REV ? i0 + kTileX : i0 - 1;
        double pc_x = 0.0;
        if (any) {
            const int half = lane >> 4, ii = lane & 15;                                // two rows per pass, 16 lanes (128 B) each
            for (int rr = 0; rr < 32; rr += 2) {
                const int r2 = rr + half, rj = j0 + (r2 & 7), rk = k0 + (r2 >> 3);
                double vi = 0.0, vp = 0.0;
                if (rj < S.J && rk < S.K && i0 + ii < S.I) {
                    const size_t c = (size_t)(i0 + ii) + (size_t)rj * sy + (size_t)rk * sz;
                    vi = in[c]; vp = S.precon[c];
                }
                sm.in[r2][ii] = vi; sm.pc[r2][ii] = vp;
            }
            if (!REV) {
                for (int h = lane; h < (kTileZ + kTileY) * kTileX; h += 32) {
                    const int hr = h / kTileX, hi = h % kTileX;
                    int hj, hk;
                    if (hr < kTileZ) { hj = j0 - 1; hk = k0 + hr; } else { hj = j0 + (hr - kTileZ); hk = k0 - 1; }
                    double v = 0.0;
                    if (hj >= 0 && hk >= 0 && hj < S.J && hk < S.K && i0 + hi < S.I) v = S.precon[(size_t)(i0 + hi) + (size_t)hj * sy + (size_t)hk * sz];
                    sm.hp[hr][hi] = v;
                }
                if (row_in && ix >= 0) pc_x = S.precon[row + ix];
            }

            // ---- predecessor tiles, then the part that depends on them
            if (lane < 3) {
                const int px = tx + (lane == 0 ? d : 0), py = ty + (lane == 1 ? d : 0), pz = tz + (lane == 2 ? d : 0);
                if (px >= 0 && py >= 0 && pz >= 0 && px < S.ntx && py < S.nty && pz < S.ntz) {
                    const volatile unsigned int *flag = S.tile_done + (px + S.ntx * (py + S.nty * pz));
                    while (*flag != epoch) { }
                }
            }
            __syncwarp();
            __threadfence();
            for (int h = lane; h < (kTileZ + kTileY) * kTileX; h += 32) {
                const int hr = h / kTileX, hi = h % kTileX;
                int hj, hk;
                if (hr < kTileZ) { hj = j0 + (REV ? kTileY : -1); hk = k0 + hr; } else { hj = j0 + (hr - kTileZ); hk = k0 + (REV ? kTileZ : -1); }
                double v = 0.0;
                if (hj >= 0 && hk >= 0 && hj < S.J && hk < S.K && i0 + hi < S.I) v = __ldcg(dyn + (size_t)(i0 + hi) + (size_t)hj * sy + (size_t)hk * sz);
                sm.hd[hr][hi] = v;
            }
            double xh = 0.0;
            if (row_in && ix >= 0 && ix < S.I) xh = __ldcg(dyn + row + ix);
            __syncwarp();

            double mine = 0.0;
#pragma unroll 2
            for (int step = 0; step < kTileX + kTileY + kTileZ - 2; step++) {
                const int li = step - a - b;
                const bool active = (unsigned)li < (unsigned)kTileX;
                const int ii = (REV ? kTileX - 1 - li : li) & (kTileX - 1);
                double from_y = REV ? __shfl_down_sync(0xffffffffu, mine, 1) : __shfl_up_sync(0xffffffffu, mine, 1);
                double from_z = REV ? __shfl_down_sync(0xffffffffu, mine, 8) : __shfl_up_sync(0xffffffffu, mine, 8);
                double val = 0.0;
                if (active && ((fl >> ii) & 1u)) {
                    if (a == 0) from_y = sm.hd[lk][ii];
                    if (b == 0) from_z = sm.hd[kTileZ + lj][ii];
                    const double from_x = li == 0 ? xh : mine;
                    const double pc = sm.pc[lane][ii];
                    double tt = sm.in[lane][ii];
                    if (!REV) {
                        const double py = a == 0 ? sm.hp[lk][ii] : sm.pc[ly][ii];
                        const double pz = b == 0 ? sm.hp[kTileZ + lj][ii] : sm.pc[lz][ii];
                        tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, pc_x), from_x));
                        tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, py), from_y));
                        tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, pz), from_z));
                    } else {
                        const double np = __dmul_rn(negscale, pc);
                        tt = __dsub_rn(tt, __dmul_rn(np, from_x));
                        tt = __dsub_rn(tt, __dmul_rn(np, from_y));
                        tt = __dsub_rn(tt, __dmul_rn(np, from_z));
                    }
                    val = __dmul_rn(tt, pc);
                    dyn[row + i0 + ii] = val;
                    pc_x = pc;
                } else if (active) {
                    pc_x = 0.0;                                          // not fluid: precon 0
                }
                mine = val;
            }
            __threadfence();
        }
        __syncwarp();
        if (lane == 0) *(volatile unsigned int *)(S.tile_done + tile) = epoch;
    }
}

// Third form of the substitutions: no tile flags and no fences at all -- the DATA carries the synchronisation.  Before a
// sweep every fluid cell of the vector it produces holds a sentinel (a NaN bit pattern no computation yields; written by
// k_press_update / k_press_setup, which pass over those cells anyway), every cell is written exactly once by an aligned
// 8-byte store, and a tile simply polls the fluid cells of its halo through L2 until none of them is the sentinel.  One
// L2 round trip per tile-to-tile hand-over instead of flag + fence + halo load, and a tile starts the moment its own
// halo is there, not when the last of three whole predecessor tiles has been flagged.

//
// FINE = true goes one step further: nothing is waited for up front.  The halo is read once (whatever is there), and a
// boundary lane that meets a sentinel when it actually needs the value spins on that one address.  A tile then runs a
// few steps behind its predecessors instead of after them -- the sweep approaches the hyperplane schedule (I + J + K
// dependent steps) instead of (tiles along the diagonal) x (steps per tile).  Every step is a chain of five dependent
// fp64 operations plus a shuffle (~290 clocks measured), so the length of that chain of steps is the whole cost.
// MEASURED SLOWER (256^3: 1.8 ms per sweep against 0.68 ms): once a tile has caught up with its predecessor, every
// boundary step waits for a store to travel through L2 (~1 us), and that latency now sits on each of the 26 steps
// instead of once per tile.  Kept as option 12 = 3 for the record and for the bit-identity test.
template <bool REV, bool FINE>
__global__ void __launch_bounds__(kSubstWarps * 32) k_press_subst_df(PressSys S) {
    __shared__ SubstSmem sm_all[kSubstWarps];
    __shared__ uint8_t hf_all[kSubstWarps][(kTileZ + kTileY) * kTileX];
    if (S.state[0]) return;
    SubstSmem &sm = sm_all[threadIdx.x >> 5];
    uint8_t *hf = hf_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, lj = lane & 7, lk = lane >> 3;
    const int a = REV ? kTileY - 1 - lj : lj, b = REV ? kTileZ - 1 - lk : lk;
    const double *__restrict__ in = REV ? S.q : S.r;
    double *dyn = REV ? S.z : S.q;
    const double negscale = -S.scale;
    const size_t sy = (size_t)S.I, sz = (size_t)S.I * (size_t)S.J;
    const int ly = REV ? lane + 1 : lane - 1, lz = REV ? lane + 8 : lane - 8;
    constexpr int MODE = REV ? 2 : 1;

    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(S.ticket + MODE, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned long long)S.ntiles) break;
        const int tile = S.order[REV ? S.ntiles - 1 - (int)t : (int)t];
        long long t0 = 0, t1 = 0, t2 = 0;
        if (S.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const int tx = tile % S.ntx, ty = (tile / S.ntx) % S.nty, tz = tile / (S.ntx * S.nty);
        const int i0 = tx * kTileX, j0 = ty * kTileY, k0 = tz * kTileZ;
        const int j = j0 + lj, k = k0 + lk;
        const bool row_in = j < S.J && k < S.K;
        const size_t row = (size_t)j * sy + (size_t)k * sz;

        unsigned fl = 0;
        if (row_in)
            for (int ii = 0; ii < kTileX; ii++)
                if (i0 + ii < S.I && (S.flags[row + i0 + ii] & kPfFluid)) fl |= 1u << ii;
        if (!__any_sync(0xffffffffu, fl != 0)) continue;                          // nothing to produce, nobody waits for it

        // ---- static part: my rows of the input vector and of the diagonal, the halo's diagonal and fluid flags
        const int ix = REV ? i0 + kTileX : i0 - 1;
        const bool x_in = row_in && ix >= 0 && ix < S.I;
        double pc_x = 0.0;
        bool x_fluid = false;
        {
            const int half = lane >> 4, ii = lane & 15;
            for (int rr = 0; rr < 32; rr += 2) {
                const int r2 = rr + half, rj = j0 + (r2 & 7), rk = k0 + (r2 >> 3);
                double vi = 0.0, vp = 0.0;
                if (rj < S.J && rk < S.K && i0 + ii < S.I) {
                    const size_t c = (size_t)(i0 + ii) + (size_t)rj * sy + (size_t)rk * sz;
                    vi = in[c]; vp = S.precon[c];
                }
                sm.in[r2][ii] = vi; sm.pc[r2][ii] = vp;
            }
            for (int h = lane; h < (kTileZ + kTileY) * kTileX; h += 32) {
                const int hr = h / kTileX, hi = h % kTileX;
                int hj, hk;
                if (hr < kTileZ) { hj = j0 + (REV ? kTileY : -1); hk = k0 + hr; } else { hj = j0 + (hr - kTileZ); hk = k0 + (REV ? kTileZ : -1); }
                double v = 0.0;
                uint8_t f = 0;
                if (hj >= 0 && hk >= 0 && hj < S.J && hk < S.K && i0 + hi < S.I) {
                    const size_t c = (size_t)(i0 + hi) + (size_t)hj * sy + (size_t)hk * sz;
                    f = S.flags[c] & kPfFluid;
                    if (!REV) v = S.precon[c];
                }
                sm.hp[hr][hi] = v; hf[h] = f;
                sm.hd[hr][hi] = 0.0;
            }
            if (x_in) { x_fluid = S.flags[row + ix] & kPfFluid; if (!REV) pc_x = S.precon[row + ix]; }
        }
        __syncwarp();
        if (S.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));

        // ---- the halo of the produced vector: poll the fluid cells until their values have arrived (FINE: one look only)
        double xh = 0.0;
        const double *halo_y = dyn + (size_t)i0 + (size_t)(j0 + (REV ? kTileY : -1)) * sy + (size_t)k * sz;      // my row in the y neighbour
        const double *halo_z = dyn + (size_t)i0 + (size_t)j * sy + (size_t)(k0 + (REV ? kTileZ : -1)) * sz;      // ... in the z neighbour
        for (;;) {
            bool ok = true;
            for (int h = lane; h < (kTileZ + kTileY) * kTileX; h += 32) {
                if (!hf[h]) continue;
                const int hr = h / kTileX, hi = h % kTileX;
                int hj, hk;
                if (hr < kTileZ) { hj = j0 + (REV ? kTileY : -1); hk = k0 + hr; } else { hj = j0 + (hr - kTileZ); hk = k0 + (REV ? kTileZ : -1); }
                const double v = ld_poll(dyn + (size_t)(i0 + hi) + (size_t)hj * sy + (size_t)hk * sz);
                if (FINE) sm.hd[hr][hi] = v;
                else if (__double_as_longlong(v) == kPressSentinel) ok = false; else sm.hd[hr][hi] = v;
            }
            if (x_fluid) {
                xh = ld_poll(dyn + row + ix);
                if (!FINE && __double_as_longlong(xh) == kPressSentinel) ok = false;
            }
            if (FINE || __all_sync(0xffffffffu, ok)) break;
        }
        __syncwarp();
        if (S.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));

        double mine = 0.0;
#pragma unroll 2
        for (int step = 0; step < kTileX + kTileY + kTileZ - 2; step++) {
            const int li = step - a - b;
            const bool active = (unsigned)li < (unsigned)kTileX;
            const int ii = (REV ? kTileX - 1 - li : li) & (kTileX - 1);
            double from_y = REV ? __shfl_down_sync(0xffffffffu, mine, 1) : __shfl_up_sync(0xffffffffu, mine, 1);
            double from_z = REV ? __shfl_down_sync(0xffffffffu, mine, 8) : __shfl_up_sync(0xffffffffu, mine, 8);
            double val = 0.0;
            if (active && ((fl >> ii) & 1u)) {
                if (a == 0) {
                    from_y = sm.hd[lk][ii];
                    if (FINE) while (__double_as_longlong(from_y) == kPressSentinel) from_y = ld_poll(halo_y + ii);
                }
                if (b == 0) {
                    from_z = sm.hd[kTileZ + lj][ii];
                    if (FINE) while (__double_as_longlong(from_z) == kPressSentinel) from_z = ld_poll(halo_z + ii);
                }
                double from_x = mine;
                if (li == 0) {
                    from_x = xh;
                    if (FINE) while (__double_as_longlong(from_x) == kPressSentinel) from_x = ld_poll(dyn + row + ix);
                }
                const double pc = sm.pc[lane][ii];
                double tt = sm.in[lane][ii];
                if (!REV) {
                    const double py = a == 0 ? sm.hp[lk][ii] : sm.pc[ly][ii];
                    const double pz = b == 0 ? sm.hp[kTileZ + lj][ii] : sm.pc[lz][ii];
                    tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, pc_x), from_x));
                    tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, py), from_y));
                    tt = __dsub_rn(tt, __dmul_rn(__dmul_rn(negscale, pz), from_z));
                } else {
                    const double np = __dmul_rn(negscale, pc);
                    tt = __dsub_rn(tt, __dmul_rn(np, from_x));
                    tt = __dsub_rn(tt, __dmul_rn(np, from_y));
                    tt = __dsub_rn(tt, __dmul_rn(np, from_z));
                }
                val = __dmul_rn(tt, pc);
                st_publish(dyn + row + i0 + ii, val);
                pc_x = pc;
            } else if (active) {
                pc_x = 0.0;
            }
            mine = val;
        }
        __syncwarp();
        if (S.trace && lane == 0) {
            long long t3;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3));
            long long *tr = S.trace + 4 * ((size_t)tile + (REV ? (size_t)S.ntiles : 0));
            tr[0] = t0; tr[1] = t1; tr[2] = t2; tr[3] = t3;
        }
    }
}

// _applyMatrix (:392-433): z = A s, and the partial sums of dot(z, s)
__global__ void __launch_bounds__(kPressThreads) k_press_apply_matrix(PressSys S) {
    __shared__ double sh[kPressThreads / 32];
    if (S.state[0]) return;
    const double scale = S.scale, negscale = -S.scale;
    const long long sy = S.I, sz = (long long)S.I * S.J;
    double acc = 0.0;
    for (long long c = (long long)blockIdx.x * kPressThreads + threadIdx.x; c < S.cells; c += (long long)kPressBlocks * kPressThreads) {
        const uint8_t fl = S.flags[c];
        if (!(fl & kPfFluid)) continue;
        const int i = (int)(c % S.I), j = (int)((c / S.I) % S.J), k = (int)(c / sz);
        double val = 0.0;
        if (i > 0) val = __dadd_rn(val, S.s[c - 1]);
        if (i + 1 < S.I) val = __dadd_rn(val, S.s[c + 1]);
        if (j > 0) val = __dadd_rn(val, S.s[c - sy]);
        if (j + 1 < S.J) val = __dadd_rn(val, S.s[c + sy]);
        if (k > 0) val = __dadd_rn(val, S.s[c - sz]);
        if (k + 1 < S.K) val = __dadd_rn(val, S.s[c + sz]);
        val = __dmul_rn(val, negscale);
        const double x = S.s[c];
        val = __dadd_rn(val, __dmul_rn(__dmul_rn((double)(fl & 7), scale), x));
        S.z[c] = val;
        acc += val * x;
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) S.partial[kPressBlocks + blockIdx.x] = acc;
}

// alpha = sigma / dot(z, s); pressure += search alpha; residual += auxillary (-alpha) (:473-476); partial max |r|
__global__ void __launch_bounds__(kPressThreads) k_press_update(PressSys S, int it, int sentinel) {
    __shared__ double sh[kPressThreads / 32];
    if (S.state[0]) return;
    const double dot = reduce_partials_sum(S.partial + kPressBlocks, sh);
    const double alpha = S.sigma[it & 1] / dot, nalpha = -alpha;
    double mx = 0.0;
    for (long long c = (long long)blockIdx.x * kPressThreads + threadIdx.x; c < S.cells; c += (long long)kPressBlocks * kPressThreads) {
        if (!(S.flags[c] & kPfFluid)) continue;
        S.p[c] = __dadd_rn(S.p[c], __dmul_rn(S.s[c], alpha));
        const double r = __dadd_rn(S.r[c], __dmul_rn(S.z[c], nalpha));
        S.r[c] = r;
        mx = fmax(mx, fabs(r));
        if (sentinel) { S.q[c] = __longlong_as_double(kPressSentinel); S.z[c] = __longlong_as_double(kPressSentinel); }   // arm the data-flow sweeps
    }
    __syncthreads();
    mx = block_max(mx, sh);
    if (threadIdx.x == 0) S.partial[blockIdx.x] = mx;
}

// partial sums of dot(z, r) (:484)
__global__ void __launch_bounds__(kPressThreads) k_press_dot_zr(PressSys S) {
    __shared__ double sh[kPressThreads / 32];
    if (S.state[0]) return;
    double acc = 0.0;
    for (long long c = (long long)blockIdx.x * kPressThreads + threadIdx.x; c < S.cells; c += (long long)kPressBlocks * kPressThreads)
        if (S.flags[c] & kPfFluid) acc += S.z[c] * S.r[c];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) S.partial[2 * kPressBlocks + blockIdx.x] = acc;
}

// first = 1: search = auxillary, sigma = dot(z, r) (:463-469); else beta = sigmaNew / sigma, search = auxillary 1.0 +
// search beta (:485-487).  it = index of the iteration whose sigma is consumed; the new sigma goes to the other slot.
__global__ void __launch_bounds__(kPressThreads) k_press_search(PressSys S, int it, int first) {
    __shared__ double sh[kPressThreads / 32];
    if (S.state[0]) return;
    const double sigma_new = reduce_partials_sum(S.partial + 2 * kPressBlocks, sh);
    const double beta = first ? 0.0 : sigma_new / S.sigma[it & 1];
    for (long long c = (long long)blockIdx.x * kPressThreads + threadIdx.x; c < S.cells; c += (long long)kPressBlocks * kPressThreads) {
        if (!(S.flags[c] & kPfFluid)) continue;
        const double z = S.z[c];
        S.s[c] = first ? z : __dadd_rn(__dmul_rn(z, 1.0), __dmul_rn(S.s[c], beta));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) S.sigma[first ? 0 : (it + 1) & 1] = sigma_new;
}

// the narrowing of FluidSimulation::_updatePressureGrid (src/fluidsimulation.cpp:2884-2888)
__global__ void __launch_bounds__(kPressThreads) k_press_finish(PressSys S) {
    for (long long c = (long long)blockIdx.x * kPressThreads + threadIdx.x; c < S.cells; c += (long long)kPressBlocks * kPressThreads)
        S.pressure[c] = (S.flags[c] & kPfFluid) ? (float)S.p[c] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Stage 6: _applyConstantBodyForces (src/fluidsimulation.cpp:2765-2805).  add[c] = (float)(force[c] * dt), formed on the
// host in double as the reference does; a component whose force is exactly zero is skipped (mask).
// ------------------------------------------------------------------------------------------------
struct Float3 { float v[3]; };

__global__ void __launch_bounds__(256) k_body_force(Grid g, const uint8_t *__restrict__ material, FieldRW f, Float3 add, int mask) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = (uint32_t)g.I + 1u;
    if (t >= w * ((uint32_t)g.J + 1u)) return;
    const int i = (int)(t % w), j = (int)(t / w), k = (int)blockIdx.y;
    const int8_t *m = reinterpret_cast<const int8_t *>(material);
    const bool fl = cell_equals(g, m, i, j, k, GFS_FLUID);
    if ((mask & 1) && j < g.J && k < g.K && (fl || cell_equals(g, m, i - 1, j, k, GFS_FLUID))) {
        float *x = f.c[0] + ((size_t)i + (size_t)g.pitch[0] * ((size_t)j + (size_t)g.J * (size_t)k));
        *x = __fadd_rn(*x, add.v[0]);
    }
    if ((mask & 2) && i < g.I && k < g.K && (fl || cell_equals(g, m, i, j - 1, k, GFS_FLUID))) {
        float *x = f.c[1] + ((size_t)i + (size_t)g.pitch[1] * ((size_t)j + (size_t)(g.J + 1) * (size_t)k));
        *x = __fadd_rn(*x, add.v[1]);
    }
    if ((mask & 4) && i < g.I && j < g.J && (fl || cell_equals(g, m, i, j, k - 1, GFS_FLUID))) {
        float *x = f.c[2] + ((size_t)i + (size_t)g.pitch[2] * ((size_t)j + (size_t)g.J * (size_t)k));
        *x = __fadd_rn(*x, add.v[2]);
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 8: _applyPressureToVelocityField (src/fluidsimulation.cpp:2895-3061), out of place: a face bordering fluid
// becomes 0 when it also borders a solid, else U - scale (p1 - p0) in double with the FLOAT pressures
// (_applyPressureToFaceU: both cells are non-solid there, only its first branch is reachable); every other face is copied.
// ------------------------------------------------------------------------------------------------
template <int COMP>
__device__ __forceinline__ void pressure_face(const Grid &g, const int8_t *m, const float *__restrict__ src, float *__restrict__ dst,
                                              const float *__restrict__ pr, double scale, int i, int j, int k) {
    const size_t idx = (size_t)i + (size_t)g.pitch[COMP] * ((size_t)j + (size_t)(g.J + (COMP == 1)) * (size_t)k);
    float x = src[idx];
    if (face_borders<COMP>(g, m, i, j, k, GFS_FLUID)) {
        if (face_borders<COMP>(g, m, i, j, k, GFS_SOLID)) {
            x = 0.0f;
        } else {
            const int ci = i - (COMP == 0), cj = j - (COMP == 1), ck = k - (COMP == 2);
            const bool lo_in = ci >= 0 && cj >= 0 && ck >= 0, hi_in = i < g.I && j < g.J && k < g.K;
            const double p0 = lo_in ? (double)pr[(size_t)ci + (size_t)g.I * ((size_t)cj + (size_t)g.J * (size_t)ck)] : 0.0;
            const double p1 = hi_in ? (double)pr[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)k)] : 0.0;
            x = (float)__dsub_rn((double)x, __dmul_rn(scale, __dsub_rn(p1, p0)));
        }
    }
    dst[idx] = x;
}

__global__ void __launch_bounds__(256) k_apply_pressure(Grid g, const uint8_t *__restrict__ material, FieldPtrs src, FieldRW dst,
                                                        const float *__restrict__ pr, double scale) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = (uint32_t)g.I + 1u;
    if (t >= w * ((uint32_t)g.J + 1u)) return;
    const int i = (int)(t % w), j = (int)(t / w), k = (int)blockIdx.y;
    const int8_t *m = reinterpret_cast<const int8_t *>(material);
    if (j < g.J && k < g.K) pressure_face<0>(g, m, src.c[0], dst.c[0], pr, scale, i, j, k);
    if (i < g.I && k < g.K) pressure_face<1>(g, m, src.c[1], dst.c[1], pr, scale, i, j, k);
    if (i < g.I && j < g.J) pressure_face<2>(g, m, src.c[2], dst.c[2], pr, scale, i, j, k);
}

}  // namespace gfs
