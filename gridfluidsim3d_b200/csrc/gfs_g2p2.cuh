// gfs_g2p2.cuh -- round-2 trilinear G2P brick kernel (power-of-two dx): PIC/FLIP + RK1..4 + solid test + next-step
// binning with the arithmetic of k_g2p_brick<0> bit for bit, restructured around what ncu showed limits that kernel
// (profiles/r01_k_g2p_brick.md: 1025 warp instructions per 32 particles at 68-76 % issue utilisation, IMAD address chains
// 26 % of the stall samples, F2I + FRND 8 %, 42 % of the shared-memory wavefronts bank conflicts, 78 registers):
//
//   * every index/fraction pair is 4 FADD-class instructions (no F2I, no FRND): the magic-number floor with the tile
//     origin folded into its integer bias gives the TILE-LOCAL index directly, and a component's tap address is base +
//     two multiply-adds with the eight taps at immediate offsets;
//   * no global-memory sampling code in the kernel: a particle with an RK stage position outside the staged block (only
//     possible when dt |v| exceeds the one-cell margin) is put on a list and done in full by k_g2p_slow afterwards --
//     the hot loop carries no fallback branches and fits 4 CTAs per SM;
//   * the out-of-grid bin is not this kernel's business either (k_g2p_brick handles it, launched on that bin alone);
//   * WIDE tile: a 20-word row pitch makes the staggered lookups bank-conflict free with plain address arithmetic (see
//     TriTile); it costs a quarter more shared memory (3 CTAs per SM instead of 4).  The 4-D bank-skewed box of the
//     tricubic kernel was tried here too and lost: its split-column addresses cost more issue slots than the conflicts.
#pragma once
#include "gfs_kernels.cuh"

namespace gfs {

// WIDE = false: box [z][y][x 16] (the tensor maps of k_g2p_brick<0>).  WIDE = true: box [z][y][x 20] -- four unused
// columns per row, chosen for the banks: a half-cell staggered lookup has two candidate rows in y and two in z, and with
// a 20-word row pitch and a 20*nY-word plane pitch (== 16 mod 32 for nY = 12, == 8 for nY = 10) the four (y, z) candidates
// of the cells of a warp fall into disjoint bank windows ({0,20,16,4}+x and {0,20,8,28}+x, x < 4) -- in the 16-wide box
// both shifts are multiples of 16 banks and 42 % of the kernel's shared-memory wavefronts were conflicts.
template <bool WIDE> struct TriTile {
    static constexpr int kMargin = 1;
    static constexpr int kOrgX = 4, kX = WIDE ? 20 : 16;     // x: nodes [8b-4, 8b+11] (+4 unused when WIDE)
    static constexpr int nOrg = 1 + kMargin, nY = 9 + 2 * kMargin + 1;           // NEW: y/z nodes [8b-2, 8b+9]  (12)
    static constexpr int sOrg = 1, sY = 10;                                       // SAVED: [8b-1, 8b+8]
    static constexpr int nZ = nY, sZ = sY;
    static constexpr int nBox = kX * nY * nZ, sBox = kX * sY * sZ;
    static constexpr int nCount = (nBox + 31) / 32 * 32, sCount = (sBox + 31) / 32 * 32;
    static constexpr uint32_t kTxBytes = 3 * (nBox + sBox) * sizeof(float);
    static constexpr size_t kSmemBytes = 3 * (nCount + sCount) * sizeof(float) + 128 + 16;
    static constexpr int kCtas = WIDE ? 3 : 4;
    template <int NYY> struct Str { static constexpr int sy = kX, sz = kX * NYY; };
};

struct AxL { int i; float t; };         // tile-local node index and fraction

// index and fraction of a coordinate u in cell units: i = floor(u) - org, t = u - floor(u) (exact), without the conversion
// pipe: u + 1.5*2^23 rounded toward -inf has floor(u) in its mantissa field (|u| < 2^22, see floor_small); `bias` =
// 0x4B400000 + org folds the tile origin into the one integer subtraction.  (Shifting u by the origin BEFORE the floor
// would round where the origin is negative -- the first brick along an axis starts at node -4 or -2.)
__device__ __forceinline__ AxL ax_tile(float u, int bias) {
    AxL r;
    const float kMagic = 12582912.0f;
    const float m = __fadd_rd(u, kMagic);
    r.i = __float_as_int(m) - bias;
    r.t = __fsub_rn(u, __fsub_rn(m, kMagic));
    return r;
}

struct IdxL { AxL ux, uy, uz, sx, sy, sz; };

// the six index/fraction pairs of a position (ux, uy, uz in cell units: x/dx, exact for dx = 2^-k); same values as sample_idx
__device__ __forceinline__ IdxL idx_tile(float ux, float uy, float uz, int biasx, int biasy, int biasz) {
    IdxL s;
    s.ux = ax_tile(ux, biasx); s.sx = ax_tile(__fsub_rn(ux, 0.5f), biasx);
    s.uy = ax_tile(uy, biasy); s.sy = ax_tile(__fsub_rn(uy, 0.5f), biasy);
    s.uz = ax_tile(uz, biasz); s.sz = ax_tile(__fsub_rn(uz, 0.5f), biasz);
    return s;
}

// one component from a staged tile; x, y, z: tile-local indices (y, z relative to the tile's own y/z origin)
template <int SY, int SZ>
__device__ __forceinline__ float tri_sample(const float *__restrict__ t, int x, float tx, int y, float ty, int z, float tz) {
    const float *r = t + SY * y + SZ * z + x;
    const float p000 = r[0], p100 = r[1], p010 = r[SY], p110 = r[SY + 1];
    const float p001 = r[SZ], p101 = r[SZ + 1], p011 = r[SZ + SY], p111 = r[SZ + SY + 1];
    const float c00 = fmaf(tx, __fsub_rn(p100, p000), p000), c10 = fmaf(tx, __fsub_rn(p110, p010), p010);
    const float c01 = fmaf(tx, __fsub_rn(p101, p001), p001), c11 = fmaf(tx, __fsub_rn(p111, p011), p011);
    const float c0 = fmaf(ty, __fsub_rn(c10, c00), c00), c1 = fmaf(ty, __fsub_rn(c11, c01), c01);
    return fmaf(tz, __fsub_rn(c1, c0), c0);
}

struct SlowList { int32_t *list; unsigned int *count; };      // sorted slots left to k_g2p_slow

// ADVECT = true: positions only -- RK1..4 through the NEW field (ParticleAdvector::advectParticlesRK*, the advection-only
// sub-metric): no SAVED tiles, no PIC/FLIP, the velocity arrays are neither read nor written.
template <bool WIDE, bool MIGRATE, bool ADVECT = false>
__global__ void __launch_bounds__(256, TriTile<WIDE>::kCtas)
k_g2p_tri(Grid g, const __grid_constant__ BrickMaps maps, const uint8_t *__restrict__ material, const int32_t *__restrict__ cell_start,
          uint32_t brick0, const int32_t *__restrict__ index, const int32_t *__restrict__ tag_in, int32_t *__restrict__ tag_out,
          int order, RkCoef rk, float ratio_pic, float ratio_flip,
          const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
          const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
          float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
          float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz,
          unsigned long long *__restrict__ counters, uint32_t nkeys, uint32_t *__restrict__ keys_out,
          uint32_t *__restrict__ rank_out, uint32_t *__restrict__ counts, unsigned int *__restrict__ vmax_bits,
          Migrate mg, CollList coll, SlowList slow) {
    typedef TriTile<WIDE> T;
    typedef typename T::template Str<T::nY> SN;
    typedef typename T::template Str<T::sY> SS;
    extern __shared__ unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    uint64_t &bar = *reinterpret_cast<uint64_t *>(tiles + 3 * (T::nCount + T::sCount));
    const uint32_t b = blockIdx.x + brick0;
    const int start = cell_start[(size_t)b * kBrickCells], end = cell_start[(size_t)(b + 1) * kBrickCells];
    if (start >= end) return;
    const int bi = (int)(b % (uint32_t)g.nbi), bj = (int)((b / (uint32_t)g.nbi) % (uint32_t)g.nbj), bk = (int)(b / ((uint32_t)g.nbi * (uint32_t)g.nbj));
    const int bx = bi * kBrick, by = bj * kBrick, bz = bk * kBrick + g.k0;
    float *tnew = tiles, *tsav = tiles + 3 * T::nCount;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)),
                     "r"(ADVECT ? (uint32_t)(3 * T::nBox * sizeof(float)) : T::kTxBytes) : "memory");
#pragma unroll
        for (int c = 0; c < 3; c++) {
            tma_load_3d(tnew + c * T::nCount, &maps.m[c], 8 * bi, by - T::nOrg, bz - g.k0 - T::nOrg, &bar);
            if (!ADVECT) tma_load_3d(tsav + c * T::sCount, &maps.m[3 + c], 8 * bi, by - T::sOrg, bz - g.k0 - T::sOrg, &bar);
        }
    }
    // NEW tile origin: x from node 8b-4, y/z from 8b-nOrg
    const int biasx = 0x4B400000 + (bx - T::kOrgX), biasy = 0x4B400000 + (by - T::nOrg), biasz = 0x4B400000 + (bz - T::nOrg);
    const float invdx = g.invdxf;
    int r = start + threadIdx.x;
    float nx_ = 0.f, ny_ = 0.f, nz_ = 0.f, nvx = 0.f, nvy = 0.f, nvz = 0.f;
    int ntag = 0;
    if (r < end) {
        const int s_ = index ? index[r] : r;
        nx_ = x[s_]; ny_ = y[s_]; nz_ = z[s_]; ntag = tag_in[s_];
        if (!ADVECT) { nvx = vx[s_]; nvy = vy[s_]; nvz = vz[s_]; }
    }
    __syncthreads();
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    float m = 0.0f;
    int pend_r = -1;
    uint32_t pend_rank = 0;
    const float a2 = order == 3 ? rk.three_quarter_dt : rk.half_dt;
    const float b0 = order == 3 ? 2.0f : 1.0f, b1 = order == 3 ? 3.0f : 2.0f, b2 = order == 3 ? 4.0f : 2.0f;
    const float h = order == 4 ? rk.dt_over_6 : (order == 3 ? rk.dt_over_9 : rk.dt);
    for (; r < end; r += blockDim.x) {
        const float px = nx_, py = ny_, pz = nz_;
        const float ux = nvx, uy = nvy, uz = nvz;
        const int tag = ntag;
        {
            const int rn = r + blockDim.x;
            if (rn < end) {
                const int s_ = index ? index[rn] : rn;
                nx_ = x[s_]; ny_ = y[s_]; nz_ = z[s_]; ntag = tag_in[s_];
                if (!ADVECT) { nvx = vx[s_]; nvy = vy[s_]; nvz = vz[s_]; }
            }
        }
        if (pend_r >= 0) { rank_out[pend_r] = pend_rank; pend_r = -1; }
        // ---- p0: NEW and SAVED share one index/fraction set (p0 lies in this brick: every tap is staged)
        float k1x, k1y, k1z, sx = 0.0f, sy = 0.0f, sz = 0.0f;
        {
            const IdxL s = idx_tile(__fmul_rn(px, invdx), __fmul_rn(py, invdx), __fmul_rn(pz, invdx), biasx, biasy, biasz);
            k1x = tri_sample<SN::sy, SN::sz>(tnew, s.ux.i, s.ux.t, s.sy.i, s.sy.t, s.sz.i, s.sz.t);
            k1y = tri_sample<SN::sy, SN::sz>(tnew + T::nCount, s.sx.i, s.sx.t, s.uy.i, s.uy.t, s.sz.i, s.sz.t);
            k1z = tri_sample<SN::sy, SN::sz>(tnew + 2 * T::nCount, s.sx.i, s.sx.t, s.sy.i, s.sy.t, s.uz.i, s.uz.t);
            constexpr int d = T::nOrg - T::sOrg;          // the SAVED tile starts d nodes later in y and z
            if (!ADVECT) {
            sx = tri_sample<SS::sy, SS::sz>(tsav, s.ux.i, s.ux.t, s.sy.i - d, s.sy.t, s.sz.i - d, s.sz.t);
            sy = tri_sample<SS::sy, SS::sz>(tsav + T::sCount, s.sx.i, s.sx.t, s.uy.i - d, s.uy.t, s.sz.i - d, s.sz.t);
            sz = tri_sample<SS::sy, SS::sz>(tsav + 2 * T::sCount, s.sx.i, s.sx.t, s.sy.i - d, s.sy.t, s.uz.i - d, s.uz.t);
            }
        }
        float nx = k1x, ny = k1y, nz = k1z;
        validate3(nx, ny, nz);
        validate3(sx, sy, sz);
        const float wx = __fadd_rn(__fmul_rn(nx, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(ux, nx), sx), ratio_flip));
        const float wy = __fadd_rn(__fmul_rn(ny, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uy, ny), sy), ratio_flip));
        const float wz = __fadd_rn(__fmul_rn(nz, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uz, nz), sz), ratio_flip));
        // ---- RK stages 2..order through the NEW tile (one sampling site, as k_g2p_brick)
        float kx = k1x, ky = k1y, kz = k1z;
        float sx_ = __fmul_rn(k1x, b0), sy_ = __fmul_rn(k1y, b0), sz_ = __fmul_rn(k1z, b0);
        bool leave = false;
#pragma unroll 1
        for (int st = 1; st < order; st++) {
            const float a = st == 1 ? rk.half_dt : (st == 2 ? a2 : rk.dt);
            const float bb = st == 1 ? b1 : (st == 2 ? b2 : 1.0f);
            const float ex = axpy(px, a, kx), ey = axpy(py, a, ky), ez = axpy(pz, a, kz);
            if (!(ex >= 0.0f && ey >= 0.0f && ez >= 0.0f && ex < g.xmaxf && ey < g.ymaxf && ez < g.zmaxf)) {
                kx = ky = kz = 0.0f;                      // outside the grid the field reads zero (macvelocityfield.cpp:351)
            } else {
                const IdxL s = idx_tile(__fmul_rn(ex, invdx), __fmul_rn(ey, invdx), __fmul_rn(ez, invdx), biasx, biasy, biasz);
                // taps c, c+1 of the six index variants must lie inside the staged box: 0 <= i <= n-2
                const bool in = (unsigned)s.ux.i <= 14u && (unsigned)s.sx.i <= 14u &&          // 16 loaded columns
                                (unsigned)s.uy.i <= (unsigned)(T::nY - 2) && (unsigned)s.sy.i <= (unsigned)(T::nY - 2) &&
                                (unsigned)s.uz.i <= (unsigned)(T::nY - 2) && (unsigned)s.sz.i <= (unsigned)(T::nY - 2);
                if (!in) { leave = true; break; }
                kx = tri_sample<SN::sy, SN::sz>(tnew, s.ux.i, s.ux.t, s.sy.i, s.sy.t, s.sz.i, s.sz.t);
                ky = tri_sample<SN::sy, SN::sz>(tnew + T::nCount, s.sx.i, s.sx.t, s.uy.i, s.uy.t, s.sz.i, s.sz.t);
                kz = tri_sample<SN::sy, SN::sz>(tnew + 2 * T::nCount, s.sx.i, s.sx.t, s.sy.i, s.sy.t, s.uz.i, s.uz.t);
            }
            sx_ = __fadd_rn(sx_, __fmul_rn(kx, bb)); sy_ = __fadd_rn(sy_, __fmul_rn(ky, bb)); sz_ = __fadd_rn(sz_, __fmul_rn(kz, bb));
        }
        if (leave) {          // rare: the whole particle is redone by k_g2p_slow (it still owns its slot r)
            slow.list[atomicAdd(slow.count, 1u)] = r;
            continue;
        }
        if (order <= 2) { sx_ = kx; sy_ = ky; sz_ = kz; }
        float qx = axpy(px, h, sx_), qy = axpy(py, h, sy_), qz = axpy(pz, h, sz_);
        // ---- cell of the advected position: solid test (out of range reads as solid, NaN -> solid), key, migration
        const bool ingrid = qx >= 0.0f && qy >= 0.0f && qz >= 0.0f && qx < g.xmaxf && qy < g.ymaxf && qz < g.zmaxf;
        int ci = 0, cj = 0, ck = 0;
        if (ingrid) {
            floor_small(__fmul_rn(qx, invdx), ci); floor_small(__fmul_rn(qy, invdx), cj); floor_small(__fmul_rn(qz, invdx), ck);
        }
        bool deferred = false;
        bool inside_now = ingrid;
        if (material) {
            bool solid = true;
            if (ingrid) {
                const int kl = ck - g.k0;
                solid = (kl >= 0 && kl < g.k1 - g.k0) ? material[(size_t)ci + (size_t)g.I * ((size_t)cj + (size_t)g.J * (size_t)kl)] == GFS_SOLID : false;
            }
            if (solid) {
                atomicAdd(&counters[2], 1ull);
                if (coll.list) {
                    const unsigned int tk = atomicAdd(coll.count, 1u);
                    if (tk < coll.cap) { coll.list[tk] = make_float4(__int_as_float(r), qx, qy, qz); deferred = true; }
                    else atomicAdd(&counters[3], 1ull);
                }
                qx = px; qy = py; qz = pz;
                if (!deferred && keys_out) {              // stays at p0: its key is p0's cell
                    inside_now = true;                    // p0 lies in this brick
                    floor_small(__fmul_rn(qx, invdx), ci); floor_small(__fmul_rn(qy, invdx), cj); floor_small(__fmul_rn(qz, invdx), ck);
                }
            }
        }
        ox[r] = qx; oy[r] = qy; oz[r] = qz;
        if (!ADVECT) { ovx[r] = wx; ovy[r] = wy; ovz[r] = wz; }
        tag_out[r] = tag;
        if (keys_out) {
            const float mm = ADVECT ? 0.0f : fmaxf(fabsf(wx), fmaxf(fabsf(wy), fabsf(wz)));
            if (mm < 3.0e38f) m = fmaxf(m, mm);
            if (!deferred) {
                uint32_t key = nkeys;
                if (inside_now) {
                    if (ck >= g.k0 && ck < g.k1) key = brick_key(g, ci, cj, ck - g.k0);
                    if (MIGRATE) {
                        const int side = ck < mg.own_lo ? 0 : (ck >= mg.own_hi ? 1 : -1);
                        if (side >= 0) {
                            const unsigned int slot = atomicAdd(mg.count + side, 1u);
                            if (slot < mg.cap && mg.out[side]) {
                                float2 *dst = reinterpret_cast<float2 *>(mg.out[side] + 6 * (size_t)slot);
                                dst[0] = make_float2(qx, qy); dst[1] = make_float2(qz, wx); dst[2] = make_float2(wy, wz);
                            }
                            key = nkeys + 1;
                        }
                    }
                }
                pend_rank = take_ticket(counts, nkeys, coll.cell_cap, key);
                keys_out[r] = key;
                pend_r = r;
            }
        }
    }
    if (pend_r >= 0) rank_out[pend_r] = pend_rank;
    if (keys_out && !ADVECT) block_vmax(m, vmax_bits);
}

// The particles k_g2p_tri left on its list (an RK stage position outside the staged block): the whole per-particle
// update through global memory, same arithmetic (evaluate_pow2 / rk_advance<2>), written to the particle's sorted slot.
template <bool MIGRATE, bool ADVECT = false>
__global__ void __launch_bounds__(128)
k_g2p_slow(Grid g, FieldPtrs fnew, FieldPtrs fsaved, const uint8_t *__restrict__ material, const int32_t *__restrict__ index,
           const int32_t *__restrict__ tag_in, int32_t *__restrict__ tag_out, int interp, int order, RkCoef rk, float ratio_pic, float ratio_flip,
           const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
           const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
           float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz,
           float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz,
           unsigned long long *__restrict__ counters, uint32_t nkeys, uint32_t *__restrict__ keys_out,
           uint32_t *__restrict__ rank_out, uint32_t *__restrict__ counts, unsigned int *__restrict__ vmax_bits,
           Migrate mg, CollList coll, SlowList slow) {
    const unsigned int n = *slow.count;
    float m = 0.0f;
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int r = slow.list[t];
        const int s_ = index ? index[r] : r;
        const float px = x[s_], py = y[s_], pz = z[s_];
        float k1x, k1y, k1z, sx, sy, sz;
        evaluate_pow2(g, fnew, interp, px, py, pz, k1x, k1y, k1z);
        evaluate_pow2(g, fsaved, interp, px, py, pz, sx, sy, sz);
        float nx = k1x, ny = k1y, nz = k1z;
        validate3(nx, ny, nz);
        validate3(sx, sy, sz);
        const float ux = vx[s_], uy = vy[s_], uz = vz[s_];
        const float wx = __fadd_rn(__fmul_rn(nx, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(ux, nx), sx), ratio_flip));
        const float wy = __fadd_rn(__fmul_rn(ny, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uy, ny), sy), ratio_flip));
        const float wz = __fadd_rn(__fmul_rn(nz, ratio_pic), __fmul_rn(__fsub_rn(__fadd_rn(uz, nz), sz), ratio_flip));
        float qx, qy, qz;
        rk_advance<2>(g, fnew, interp, order, rk, px, py, pz, k1x, k1y, k1z, qx, qy, qz);
        bool deferred = false;
        if (material) {
            bool solid = true;
            if (qx >= 0.0f && qy >= 0.0f && qz >= 0.0f && qx < g.xmaxf && qy < g.ymaxf && qz < g.zmaxf) {
                const int i = (int)floorf(__fmul_rn(qx, g.invdxf)), j = (int)floorf(__fmul_rn(qy, g.invdxf)), k = (int)floorf(__fmul_rn(qz, g.invdxf));
                const int kl = k - g.k0;
                solid = (kl >= 0 && kl < g.k1 - g.k0) ? material[(size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)kl)] == GFS_SOLID : false;
            }
            if (solid) {
                atomicAdd(&counters[2], 1ull);
                if (coll.list) {
                    const unsigned int tk = atomicAdd(coll.count, 1u);
                    if (tk < coll.cap) { coll.list[tk] = make_float4(__int_as_float(r), qx, qy, qz); deferred = true; }
                    else atomicAdd(&counters[3], 1ull);
                }
                qx = px; qy = py; qz = pz;
            }
        }
        ox[r] = qx; oy[r] = qy; oz[r] = qz;
        if (!ADVECT) { ovx[r] = wx; ovy[r] = wy; ovz[r] = wz; }
        tag_out[r] = tag_in[s_];
        if (keys_out) {
            const float mm = ADVECT ? 0.0f : fmaxf(fabsf(wx), fmaxf(fabsf(wy), fabsf(wz)));
            if (mm < 3.0e38f) m = fmaxf(m, mm);
            if (!deferred) {
                uint32_t key = position_key(g, nkeys, qx, qy, qz);
                if (MIGRATE && key < nkeys) {
                    const int k = cell_floor((double)qz, g.invdx);
                    const int side = k < mg.own_lo ? 0 : (k >= mg.own_hi ? 1 : -1);
                    if (side >= 0) {
                        const unsigned int slot = atomicAdd(mg.count + side, 1u);
                        if (slot < mg.cap && mg.out[side]) {
                            float2 *dst = reinterpret_cast<float2 *>(mg.out[side] + 6 * (size_t)slot);
                            dst[0] = make_float2(qx, qy); dst[1] = make_float2(qz, wx); dst[2] = make_float2(wy, wz);
                        }
                        key = nkeys + 1;
                    }
                }
                rank_out[r] = take_ticket(counts, nkeys, coll.cell_cap, key);
                keys_out[r] = key;
            }
        }
    }
    if (keys_out && !ADVECT) block_vmax(m, vmax_bits);
}

}  // namespace gfs
