// gfs_p2g2.cuh -- round-2 brick-tile splat (power-of-two dx): the same integer arithmetic as k_p2g_tile<2>, bit for
// bit, restructured around what ncu showed limits it (profiles/r01_k_p2g_tile.md): shared-memory atomic wavefronts
// (98 % of the LSU peak, 2.9 wavefronts per ATOMS because the lanes of a warp are the particles of ONE cell and hit
// the same nodes) and instruction issue (867 warp instructions per 32 particles, a third of them control flow around
// the per-candidate branches).
//
//   * all index / fraction / squared-distance arithmetic in cell units (dx = 2^-k: the scaling is exact, every
//     comparison and every weight has the bits of the reference-ordered float arithmetic of splat_comp_pow2);
//   * the 2^44 fixed-point scale is folded into the polynomial's constants and the per-component value, x+y partial
//     sums of the squared distances are shared between the two z candidates: 15 instructions per hit slot;
//   * one divergent region per hit slot and nothing else (ptxas cannot predicate ATOMS; see splat_slot);
//   * nodes outside the component's array are not range-checked per particle: a brick's particles can only touch the
//     10^3 tile, and the flush drops the tile nodes that do not exist (same dropped set as the reference's clamp);
//   * LANE TRANSPOSITION (TRANSPOSE = true): the brick's run is staged through shared memory in chunks of 1024
//     particles (coalesced gather in, conflict-free swizzled read out) so that the lanes of a warp work on particles of
//     32 DIFFERENT cells -- 8 consecutive cells along x in 4 rows two apart, which never share a node within one
//     candidate slot -- instead of 4 cells x 8 particles.
#pragma once
#include "gfs_kernels.cuh"

namespace gfs {

struct AxisSq { int c; float e0, e1; };        // base node (global index), squared distances to nodes c and c+1 in cell units

// unstaggered and staggered candidates of one axis (same values as axis_cand(q) / axis_cand(q - dx/2), scaled by 1/dx^2)
__device__ __forceinline__ void axis_sq(float p, float invdx, AxisSq &un, AxisSq &st) {
    const float u = __fmul_rn(p, invdx), us = __fsub_rn(u, 0.5f);
    const float fu = floorf(u), fs = floorf(us);
    un.c = (int)fu; st.c = (int)fs;
    const float t = __fsub_rn(u, fu), ts = __fsub_rn(us, fs);
    const float m = __fsub_rn(1.0f, t), ms = __fsub_rn(1.0f, ts);
    un.e0 = __fmul_rn(t, t); un.e1 = __fmul_rn(m, m);
    st.e0 = __fmul_rn(ts, ts); st.e1 = __fmul_rn(ms, ms);
}

constexpr float kC0s = 0.55555555555555556f * 17592186044416.0f;      // (5/9) 2^44, (4/9) 2^44: exact scalings of the
constexpr float kC1s = 0.44444444444444444f * 17592186044416.0f;      // float constants of kernel_weight_fast

// One hit slot.  ptxas does not predicate shared-memory atomics (a predicated red.shared becomes BSSY / @!p BRA / ATOMS /
// BSYNC around EVERY atomic -- tried with inline PTX), so the cheapest form is one divergent region per slot holding the
// weight, the two conversions and the four ATOMS: 3 control instructions per slot.
template <int C, int AB>
__device__ __forceinline__ void splat_slot(uint32_t *__restrict__ t, float xy, float ez, float vs) {
    const float d2 = __fadd_rn(xy, ez);                                         // (x^2 + y^2) + z^2, vmath.h:71-73
    if (d2 < 1.0f) {
        const float u = __fsub_rn(1.0f, d2);
        const float w44 = __fmul_rn(__fmul_rn(u, u), fmaf(u, kC1s, kC0s));      // 2^44 u^2 (5/9 + 4/9 u)
        const long long ww = __float2ll_rn(w44), wn = __float2ll_rn(__fmul_rn(w44, vs));
        constexpr int off = (AB & 1) + kTileEdge * ((AB >> 1) + kTileEdge * C);
        atomicAdd(t + off, (uint32_t)(wn >> kLoBits));
        atomicAdd(t + kTileNodes + off, (uint32_t)wn & ((1u << kLoBits) - 1u));
        atomicAdd(t + 2 * kTileNodes + off, (uint32_t)(ww >> kLoBits));
        atomicAdd(t + 3 * kTileNodes + off, (uint32_t)ww & ((1u << kLoBits) - 1u));
    }
}

// the 8 candidate nodes of one component; addr = word 0 of the component's tile at the particle's base node
__device__ __forceinline__ void splat8(uint32_t *__restrict__ addr, const AxisSq &X, const AxisSq &Y, const AxisSq &Z, float vs) {
    const float xy0 = __fadd_rn(X.e0, Y.e0), xy1 = __fadd_rn(X.e1, Y.e0), xy2 = __fadd_rn(X.e0, Y.e1), xy3 = __fadd_rn(X.e1, Y.e1);
    splat_slot<0, 0>(addr, xy0, Z.e0, vs); splat_slot<0, 1>(addr, xy1, Z.e0, vs);
    splat_slot<0, 2>(addr, xy2, Z.e0, vs); splat_slot<0, 3>(addr, xy3, Z.e0, vs);
    splat_slot<1, 0>(addr, xy0, Z.e1, vs); splat_slot<1, 1>(addr, xy1, Z.e1, vs);
    splat_slot<1, 2>(addr, xy2, Z.e1, vs); splat_slot<1, 3>(addr, xy3, Z.e1, vs);
}

__device__ __forceinline__ void splat_particle(uint32_t *__restrict__ tile, float invdx, float inv_vscale, int ti, int tj, int tk,
                                               float px, float py, float pz, float vx, float vy, float vz) {
    AxisSq ux, sx, uy, sy, uz, sz;
    axis_sq(px, invdx, ux, sx); axis_sq(py, invdx, uy, sy); axis_sq(pz, invdx, uz, sz);
    const int uxl = ux.c - ti, sxl = sx.c - ti;
    const int uyl = kTileEdge * (uy.c - tj), syl = kTileEdge * (sy.c - tj);
    const int uzl = kTileEdge * kTileEdge * (uz.c - tk), szl = kTileEdge * kTileEdge * (sz.c - tk);
    splat8(tile + (uxl + syl + szl), ux, sy, sz, __fmul_rn(vx, inv_vscale));
    splat8(tile + (4 * kTileNodes + sxl + uyl + szl), sx, uy, sz, __fmul_rn(vy, inv_vscale));
    splat8(tile + (8 * kTileNodes + sxl + syl + uzl), sx, sy, uz, __fmul_rn(vz, inv_vscale));
}

// staging slot of chunk position p: XOR swizzle of the low three bits so that both the coalesced store (consecutive p)
// and the transposed read (p = 8 lx + 128 m + const) are bank-conflict free
__device__ __forceinline__ int stage_slot(int p) { return p ^ ((((p >> 5) & 1) << 2) | ((p >> 7) & 3)); }

constexpr int kStageChunk = 1024;

template <bool TRANSPOSE>
__global__ void __launch_bounds__(TRANSPOSE ? 512 : 256, TRANSPOSE ? 2 : 4)
k_p2g_tile2(Grid g, SplatParams sp, const int32_t *__restrict__ cell_start, uint32_t brick0, const int32_t *__restrict__ index,
            const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
            const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
            unsigned long long *__restrict__ accu, unsigned long long *__restrict__ accv, unsigned long long *__restrict__ accw) {
    extern __shared__ uint32_t tile[];                    // [3 comps][4 words][1000 nodes] (+ [6][1024] staging floats)
    const uint32_t b = blockIdx.x + brick0;
    const int start = cell_start[(size_t)b * kBrickCells], end = cell_start[(size_t)(b + 1) * kBrickCells];
    if (start == end) return;
    int dense = 0;
    for (int c = threadIdx.x; c < kBrickCells; c += blockDim.x)
        dense |= (cell_start[(size_t)b * kBrickCells + c + 1] - cell_start[(size_t)b * kBrickCells + c]) > 63;
    for (int t = threadIdx.x; t < 12 * kTileNodes; t += blockDim.x) tile[t] = 0u;
    dense = __syncthreads_or(dense);
    const int bi = (int)(b % (uint32_t)g.nbi), bj = (int)((b / (uint32_t)g.nbi) % (uint32_t)g.nbj), bk = (int)(b / ((uint32_t)g.nbi * (uint32_t)g.nbj));
    const int ti = bi * kBrick - 1, tj = bj * kBrick - 1, tk = bk * kBrick - 1 + g.k0;
    const int vexp = num_exponent(sp);
    if (dense) {      // > 63 particles in a cell: 32-bit words could overflow, accumulate straight into the 64-bit grids
        const float ns = num_scale_f(vexp);
        for (int q = start + threadIdx.x; q < end; q += blockDim.x) {
            const int r = index ? index[q] : q;
            splat_pow2<false>(g, sp, ns, x[r], y[r], z[r], vx[r], vy[r], vz[r], tile, ti, tj, tk, accu, accv, accw);
        }
        return;
    }
    const float inv_vscale = __uint_as_float((unsigned)(127 - vexp) << 23);      // 2^-vexp
    const float invdx = g.invdxf;
    if (!TRANSPOSE) {
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
            splat_particle(tile, invdx, inv_vscale, ti, tj, tk, px, py, pz, ux, uy, uz);
        }
    } else {
        float *stage = reinterpret_cast<float *>(tile + 12 * kTileNodes);        // [6][kStageChunk]
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // 16 warps; warp w reads chunk positions 512 (w/8) + 64 ((w/4)&1) + 128 (lane/8) + 8 (lane&7) + 2 (w&3) + r, r = 0..1:
        // with 8 particles per cell these are 8 consecutive cells along x in rows 0,2,4,6 (or 1,3,5,7) of one z plane
        const int pbase = 512 * (warp >> 3) + 64 * ((warp >> 2) & 1) + 128 * (lane >> 3) + 8 * (lane & 7) + 2 * (warp & 3);
        // software pipeline over the chunks: the (index -> particle) gathers of chunk c+1 are in flight, in registers,
        // while chunk c is splatted out of the staging buffer
        constexpr int kPer = kStageChunk / 512;
        float rx[kPer], ry[kPer], rz[kPer], rvx[kPer], rvy[kPer], rvz[kPer];
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            const int q = start + threadIdx.x + 512 * k;
            if (q < end) {
                const int r = index ? index[q] : q;
                rx[k] = x[r]; ry[k] = y[r]; rz[k] = z[r]; rvx[k] = vx[r]; rvy[k] = vy[r]; rvz[k] = vz[r];
            }
        }
        for (int c0 = start; c0 < end; c0 += kStageChunk) {
            const int cn = min(kStageChunk, end - c0);
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int p = threadIdx.x + 512 * k;
                if (p < cn) {
                    const int s = stage_slot(p);
                    stage[s] = rx[k]; stage[kStageChunk + s] = ry[k]; stage[2 * kStageChunk + s] = rz[k];
                    stage[3 * kStageChunk + s] = rvx[k]; stage[4 * kStageChunk + s] = rvy[k]; stage[5 * kStageChunk + s] = rvz[k];
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int q = c0 + kStageChunk + threadIdx.x + 512 * k;
                if (q < end) {
                    const int r = index ? index[q] : q;
                    rx[k] = x[r]; ry[k] = y[r]; rz[k] = z[r]; rvx[k] = vx[r]; rvy[k] = vy[r]; rvz[k] = vz[r];
                }
            }
#pragma unroll 1
            for (int r = 0; r < 2; r++) {
                const int p = pbase + r;
                if (p < cn) {
                    const int s = stage_slot(p);
                    splat_particle(tile, invdx, inv_vscale, ti, tj, tk, stage[s], stage[kStageChunk + s], stage[2 * kStageChunk + s],
                                   stage[3 * kStageChunk + s], stage[4 * kStageChunk + s], stage[5 * kStageChunk + s]);
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // flush: one 64-bit integer RED per touched value; tile nodes that do not exist in the component's array (outside
    // the grid, or outside this slab's stored layers) are dropped -- the reference clamps its candidate range the same way
    const int kl = g.k1 - g.k0;
    for (int t = threadIdx.x; t < 3 * kTileNodes; t += blockDim.x) {
        const int comp = t / kTileNodes, nloc = t - comp * kTileNodes;
        const uint32_t *tw = tile + comp * 4 * kTileNodes + nloc;
        const long long wn = ((long long)(int32_t)tw[0] << kLoBits) + (long long)tw[kTileNodes];
        const long long ww = ((long long)(int32_t)tw[2 * kTileNodes] << kLoBits) + (long long)tw[3 * kTileNodes];
        if (wn == 0 && ww == 0) continue;
        const int li = nloc % kTileEdge, lj = (nloc / kTileEdge) % kTileEdge, lk = nloc / (kTileEdge * kTileEdge);
        const int ni = g.I + (comp == 0), nj = g.J + (comp == 1), nkl = kl + (comp == 2);
        const int i = ti + li, j = tj + lj, kk = tk + lk - g.k0;
        if ((unsigned)i >= (unsigned)ni || (unsigned)j >= (unsigned)nj || (unsigned)kk >= (unsigned)nkl) continue;
        const size_t node = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)kk);
        unsigned long long *acc = comp == 0 ? accu : (comp == 1 ? accv : accw);
        if (wn) atomicAdd(acc + 2 * node, (unsigned long long)wn);
        if (ww) atomicAdd(acc + 2 * node + 1, (unsigned long long)ww);
    }
}

}  // namespace gfs
