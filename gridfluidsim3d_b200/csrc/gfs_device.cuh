// gfs_device.cuh -- device-side arithmetic shared by the sm_100a kernels.
//
// Index and fraction arithmetic always follows the reference's fp64 operation sequence without
// contraction (explicit __d*_rn intrinsics), so particle->cell indices are bit-exact in every mode
// (Grid3d::positionToGridIndex, /root/reference/src/grid3d.h:35-63).  The tap contraction is fp32 with
// FMAs in GFS_FAST mode and the reference's own fp64 sequence in GFS_EXACT mode
// (Interpolation::*, /root/reference/src/interpolation.cpp:26-59).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gfs {

constexpr int kBrick = 8;                 // cells per brick edge
constexpr int kBrickCells = 512;
constexpr uint32_t kKeySentinel = 0xFFFFFFFFu;

// Geometry of the (possibly z-slab-local) grid a kernel works on.
constexpr int kRowPad = 4;      // zero floats in front of every row of a resident field array (see BrickTile)

struct Grid {
    int I, J, K;          // global cell counts
    int k0, k1;           // cell layers [k0,k1) stored locally (0,K on a single GPU)
    int nbi, nbj, nbk;    // bricks over the extended node range (I+1, J+1, k1-k0+1)
    double dx, invdx;     // invdx = 1.0/dx as the reference computes it
    double xmax, ymax, zmax;   // dx*I, dx*J, dx*K (Grid3d::isPositionInGrid, grid3d.h:137-139)
    double halfdx;        // 0.5*dx
    // fp32 mirrors, used by the fast path when dx is a power of two: then x*invdx, i*dx, x-i*dx, x-0.5dx and
    // dx*I are all exact in fp32 for in-grid positions, so the fp32 index/fraction equal the reference's fp64 ones
    float dxf, invdxf, halfdxf, xmaxf, ymaxf, zmaxf;
    int pow2;             // dx is a power of two (and the extents are exactly representable in fp32)
    int pitch[3];         // row pitch (floats) of the u, v, w arrays: >= I+1, I, I; resident fields pad it to a
                          // multiple of 4 so that TMA global strides are multiples of 16 bytes
};

struct FieldPtrs {        // one MAC field: u (I+1,J,kl), v (I,J+1,kl), w (I,J,kl+1), kl = k1-k0
    const float *c[3];
};

__device__ __forceinline__ int cell_floor(double x, double invdx) {
    return (int)floor(__dmul_rn(x, invdx));
}

// brick-major key of local cell (i,j,kl)  (kl = k - k0)
__device__ __forceinline__ uint32_t brick_key(const Grid &g, int i, int j, int kl) {
    uint32_t b = ((uint32_t)(kl >> 3) * g.nbj + (uint32_t)(j >> 3)) * g.nbi + (uint32_t)(i >> 3);
    return (b << 9) | ((uint32_t)(kl & 7) << 6) | ((uint32_t)(j & 7) << 3) | (uint32_t)(i & 7);
}

// value of component `comp` at integer face index, 0 outside the component's (global) array
// (MACVelocityField::U/V/W, macvelocityfield.cpp:99-145) or outside the locally stored layers.
__device__ __forceinline__ float tap(const Grid &g, const float *__restrict__ a, int comp, int i, int j, int k) {
    int ni = g.I + (comp == 0), nj = g.J + (comp == 1);
    int kl = k - g.k0, nkl = g.k1 - g.k0 + (comp == 2);
    if ((unsigned)i >= (unsigned)ni || (unsigned)j >= (unsigned)nj || (unsigned)kl >= (unsigned)nkl) return 0.0f;
    return __ldg(a + ((size_t)i + (size_t)g.pitch[comp] * ((size_t)j + (size_t)nj * (size_t)kl)));
}

struct AxisIdx { int i; double t; };

// cell index and fraction along one axis: i = floor(x/dx), t = (x - i*dx)/dx, reference sequence
// (macvelocityfield.cpp:358-366)
__device__ __forceinline__ AxisIdx axis_index(double x, const Grid &g) {
    AxisIdx r;
    r.i = cell_floor(x, g.invdx);
    double gx = __dmul_rn((double)r.i, g.dx);
    r.t = __dmul_rn(__dsub_rn(x, gx), g.invdx);
    return r;
}

// ---- exact (fp64, uncontracted) contraction -----------------------------------------------------
__device__ __forceinline__ double cubic_exact(double p0, double p1, double p2, double p3, double x) {
    // p[1] + 0.5*x*(p[2]-p[0] + x*(2.0*p[0]-5.0*p[1]+4.0*p[2]-p[3] + x*(3.0*(p[1]-p[2])+p[3]-p[0])))
    double a = __dsub_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn(p1, p2)), p3), p0);
    double b = __dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, p0), __dmul_rn(5.0, p1)), __dmul_rn(4.0, p2)), p3);
    double c = __dadd_rn(__dsub_rn(p2, p0), __dmul_rn(x, __dadd_rn(b, __dmul_rn(x, a))));
    return __dadd_rn(p1, __dmul_rn(__dmul_rn(0.5, x), c));
}

template <int COMP>
__device__ __forceinline__ float sample_component_exact(const Grid &g, const float *__restrict__ a, int interp,
                                                        const AxisIdx &ax, const AxisIdx &ay, const AxisIdx &az) {
    if (interp == 1) {
        double col[4];
#pragma unroll 1
        for (int pk = 0; pk < 4; pk++) {
            double row[4];
#pragma unroll
            for (int pj = 0; pj < 4; pj++) {
                double p0 = tap(g, a, COMP, ax.i - 1, ay.i - 1 + pj, az.i - 1 + pk);
                double p1 = tap(g, a, COMP, ax.i,     ay.i - 1 + pj, az.i - 1 + pk);
                double p2 = tap(g, a, COMP, ax.i + 1, ay.i - 1 + pj, az.i - 1 + pk);
                double p3 = tap(g, a, COMP, ax.i + 2, ay.i - 1 + pj, az.i - 1 + pk);
                row[pj] = cubic_exact(p0, p1, p2, p3, ax.t);
            }
            col[pk] = cubic_exact(row[0], row[1], row[2], row[3], ay.t);
        }
        return (float)cubic_exact(col[0], col[1], col[2], col[3], az.t);
    }
    // trilinear, vertex order {000,100,010,001,101,011,110,111} (interpolation.cpp:51-59)
    double x = ax.t, y = ay.t, z = az.t;
    double mx = __dsub_rn(1.0, x), my = __dsub_rn(1.0, y), mz = __dsub_rn(1.0, z);
    double p0 = tap(g, a, COMP, ax.i,     ay.i,     az.i);
    double p1 = tap(g, a, COMP, ax.i + 1, ay.i,     az.i);
    double p2 = tap(g, a, COMP, ax.i,     ay.i + 1, az.i);
    double p3 = tap(g, a, COMP, ax.i,     ay.i,     az.i + 1);
    double p4 = tap(g, a, COMP, ax.i + 1, ay.i,     az.i + 1);
    double p5 = tap(g, a, COMP, ax.i,     ay.i + 1, az.i + 1);
    double p6 = tap(g, a, COMP, ax.i + 1, ay.i + 1, az.i);
    double p7 = tap(g, a, COMP, ax.i + 1, ay.i + 1, az.i + 1);
    double s = __dmul_rn(__dmul_rn(__dmul_rn(p0, mx), my), mz);
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p1, x), my), mz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p2, mx), y), mz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p3, mx), my), z));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p4, x), my), z));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p5, mx), y), z));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p6, x), y), mz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(p7, x), y), z));
    return (float)s;
}

// ---- fast (fp32) contraction ---------------------------------------------------------------------
// Catmull-Rom weights of the four taps for fraction t (same cubic as interpolation.cpp:44-46, regrouped)
__device__ __forceinline__ void cr_weights(float t, float w[4]) {
    // explicit roundings: every kernel that samples must produce the same bits, whatever it is inlined into
    const float t2 = __fmul_rn(t, t), t3 = __fmul_rn(t2, t);
    w[0] = __fmul_rn(0.5f, __fsub_rn(fmaf(2.0f, t2, -t3), t));
    w[1] = __fmul_rn(0.5f, fmaf(3.0f, t3, fmaf(-5.0f, t2, 2.0f)));
    w[2] = __fmul_rn(0.5f, fmaf(-3.0f, t3, fmaf(4.0f, t2, t)));
    w[3] = __fmul_rn(0.5f, __fsub_rn(t3, t2));
}

struct AxisIdxF { int i; float t; };

__device__ __forceinline__ AxisIdxF to_f(const AxisIdx &a) { AxisIdxF r; r.i = a.i; r.t = (float)a.t; return r; }

// fp32 index + fraction; exact (== axis_index) when dx is a power of two, see Grid::pow2
// floor of |s| < 2^22 without the conversion pipe: s + 1.5*2^23 rounded toward -inf lands on the integer grid of
// [2^23, 2^24) (ulp 1), so the sum IS floor(s) + 1.5*2^23 and its mantissa field is the integer.  Same value as
// floorf / (int)floorf for every in-range argument; F2I / FRND run at a quarter of the FADD rate.
__device__ __forceinline__ float floor_small(float s, int &i) {
    const float kMagic = 12582912.0f;                       // 1.5 * 2^23 = 0x4B400000
    const float m = __fadd_rd(s, kMagic);
    i = __float_as_int(m) - 0x4B400000;
    return __fsub_rn(m, kMagic);
}

// MAGIC: floor through floor_small (same values; pays off in the tricubic brick kernel, costs the register-starved
// trilinear one two spills -- measured both ways)
template <bool MAGIC = false>
__device__ __forceinline__ AxisIdxF axis_index_f(float x, const Grid &g) {
    AxisIdxF r;
    float s = __fmul_rn(x, g.invdxf);
    float fl;
    if (MAGIC) fl = floor_small(s, r.i);
    else { fl = floorf(s); r.i = (int)fl; }
    r.t = __fmul_rn(__fsub_rn(x, __fmul_rn(fl, g.dxf)), g.invdxf);
    return r;
}

template <int COMP>
__device__ __forceinline__ float sample_component_fast(const Grid &g, const float *__restrict__ a, int interp,
                                                       const AxisIdxF &ax, const AxisIdxF &ay, const AxisIdxF &az) {
    const int ni = g.I + (COMP == 0), nj = g.J + (COMP == 1);
    const int nkl = g.k1 - g.k0 + (COMP == 2);
    const int i = ax.i, j = ay.i, kl = az.i - g.k0;
    const float tx = ax.t, ty = ay.t, tz = az.t;
    const size_t pitch = (size_t)g.pitch[COMP];
    if (interp == 1) {
        float wx[4], wy[4], wz[4];
        cr_weights(tx, wx); cr_weights(ty, wy); cr_weights(tz, wz);
        float acc = 0.0f;
        if (i >= 1 && i + 2 < ni && j >= 1 && j + 2 < nj && kl >= 1 && kl + 2 < nkl) {   // interior: no range checks
            const float *base = a + ((size_t)(i - 1) + pitch * ((size_t)(j - 1) + (size_t)nj * (size_t)(kl - 1)));
#pragma unroll
            for (int pk = 0; pk < 4; pk++) {
                float sk = 0.0f;
#pragma unroll
                for (int pj = 0; pj < 4; pj++) {
                    const float *r = base + pitch * ((size_t)pj + (size_t)nj * (size_t)pk);
                    float sj = __fmul_rn(wx[0], __ldg(r));
                    sj = fmaf(wx[1], __ldg(r + 1), sj);
                    sj = fmaf(wx[2], __ldg(r + 2), sj);
                    sj = fmaf(wx[3], __ldg(r + 3), sj);
                    sk = fmaf(wy[pj], sj, sk);
                }
                acc = fmaf(wz[pk], sk, acc);
            }
        } else {
#pragma unroll 1
            for (int pk = 0; pk < 4; pk++) {
                float sk = 0.0f;
#pragma unroll 1
                for (int pj = 0; pj < 4; pj++) {
                    float sj = 0.0f;
#pragma unroll
                    for (int pi = 0; pi < 4; pi++)
                        sj = fmaf(wx[pi], tap(g, a, COMP, i - 1 + pi, j - 1 + pj, az.i - 1 + pk), sj);
                    sk = fmaf(wy[pj], sj, sk);
                }
                acc = fmaf(wz[pk], sk, acc);
            }
        }
        return acc;
    }
    float p000, p100, p010, p001, p101, p011, p110, p111;
    if (i >= 0 && i + 1 < ni && j >= 0 && j + 1 < nj && kl >= 0 && kl + 1 < nkl) {
        const float *r = a + ((size_t)i + pitch * ((size_t)j + (size_t)nj * (size_t)kl));
        const size_t sj = pitch, sk = pitch * (size_t)nj;
        p000 = __ldg(r);           p100 = __ldg(r + 1);
        p010 = __ldg(r + sj);      p110 = __ldg(r + sj + 1);
        p001 = __ldg(r + sk);      p101 = __ldg(r + sk + 1);
        p011 = __ldg(r + sk + sj); p111 = __ldg(r + sk + sj + 1);
    } else {
        p000 = tap(g, a, COMP, i, j, az.i);         p100 = tap(g, a, COMP, i + 1, j, az.i);
        p010 = tap(g, a, COMP, i, j + 1, az.i);     p110 = tap(g, a, COMP, i + 1, j + 1, az.i);
        p001 = tap(g, a, COMP, i, j, az.i + 1);     p101 = tap(g, a, COMP, i + 1, j, az.i + 1);
        p011 = tap(g, a, COMP, i, j + 1, az.i + 1); p111 = tap(g, a, COMP, i + 1, j + 1, az.i + 1);
    }
    float c00 = fmaf(tx, __fsub_rn(p100, p000), p000), c10 = fmaf(tx, __fsub_rn(p110, p010), p010);
    float c01 = fmaf(tx, __fsub_rn(p101, p001), p001), c11 = fmaf(tx, __fsub_rn(p111, p011), p011);
    float c0 = fmaf(ty, __fsub_rn(c10, c00), c00), c1 = fmaf(ty, __fsub_rn(c11, c01), c01);
    return fmaf(tz, __fsub_rn(c1, c0), c0);
}

// MACVelocityField::evaluateVelocityAtPosition[Linear] (macvelocityfield.cpp:545-575): zero outside the
// grid; U shifts y,z by -0.5dx, V shifts x,z, W shifts x,y (:355-356, :389-390, :423-424).
template <int ARITH>
__device__ __forceinline__ void evaluate(const Grid &g, const FieldPtrs &f, int interp, float px, float py, float pz,
                                         float &ox, float &oy, float &oz) {
    double x = px, y = py, z = pz;
    if (!(x >= 0 && y >= 0 && z >= 0 && x < g.xmax && y < g.ymax && z < g.zmax)) { ox = oy = oz = 0.0f; return; }
    AxisIdx ux = axis_index(x, g), uy = axis_index(y, g), uz = axis_index(z, g);
    AxisIdx sx = axis_index(__dsub_rn(x, g.halfdx), g), sy = axis_index(__dsub_rn(y, g.halfdx), g),
            sz = axis_index(__dsub_rn(z, g.halfdx), g);
    if (ARITH == 1) {
        ox = sample_component_exact<0>(g, f.c[0], interp, ux, sy, sz);
        oy = sample_component_exact<1>(g, f.c[1], interp, sx, uy, sz);
        oz = sample_component_exact<2>(g, f.c[2], interp, sx, sy, uz);
    } else {
        AxisIdxF fux = to_f(ux), fuy = to_f(uy), fuz = to_f(uz), fsx = to_f(sx), fsy = to_f(sy), fsz = to_f(sz);
        ox = sample_component_fast<0>(g, f.c[0], interp, fux, fsy, fsz);
        oy = sample_component_fast<1>(g, f.c[1], interp, fsx, fuy, fsz);
        oz = sample_component_fast<2>(g, f.c[2], interp, fsx, fsy, fuz);
    }
}

// Fast path with all index arithmetic in fp32 (valid when g.pow2): same results as evaluate<0>.
__device__ __forceinline__ void evaluate_pow2(const Grid &g, const FieldPtrs &f, int interp, float px, float py, float pz,
                                              float &ox, float &oy, float &oz) {
    if (!(px >= 0.0f && py >= 0.0f && pz >= 0.0f && px < g.xmaxf && py < g.ymaxf && pz < g.zmaxf)) { ox = oy = oz = 0.0f; return; }
    AxisIdxF ux = axis_index_f(px, g), uy = axis_index_f(py, g), uz = axis_index_f(pz, g);
    AxisIdxF sx = axis_index_f(__fsub_rn(px, g.halfdxf), g), sy = axis_index_f(__fsub_rn(py, g.halfdxf), g),
             sz = axis_index_f(__fsub_rn(pz, g.halfdxf), g);
    ox = sample_component_fast<0>(g, f.c[0], interp, ux, sy, sz);
    oy = sample_component_fast<1>(g, f.c[1], interp, sx, uy, sz);
    oz = sample_component_fast<2>(g, f.c[2], interp, sx, sy, uz);
}

// mode dispatch: ARITH 1 = exact fp64, 0 = fast (fp64 index), 2 = fast with fp32 index (power-of-two dx)
template <int ARITH>
__device__ __forceinline__ void evaluate_any(const Grid &g, const FieldPtrs &f, int interp, float px, float py, float pz,
                                             float &ox, float &oy, float &oz) {
    if (ARITH == 2) evaluate_pow2(g, f, interp, px, py, pz, ox, oy, oz);
    else evaluate<ARITH>(g, f, interp, px, py, pz, ox, oy, oz);
}

// ParticleAdvector::_validateOutput (particleadvector.cpp:1139-1149)
__device__ __forceinline__ void validate3(float &x, float &y, float &z) {
    if (isinf(x) || isnan(x) || isinf(y) || isnan(y) || isinf(z) || isnan(z)) { x = y = z = 0.0f; }
}

// p + s*k with the reference's float rounding (vmath.cpp:42-83), never contracted
__device__ __forceinline__ float axpy(float p, float s, float k) { return __fadd_rn(p, __fmul_rn(k, s)); }

struct RkCoef {           // (float) casts done once on the host exactly where the reference casts (particleadvector.cpp:1045-1078)
    float dt, half_dt, three_quarter_dt, dt_over_6, dt_over_9;
};

// ParticleAdvector::_RK1.._RK4.  k1 may be supplied (PIC/FLIP's vnew at p0 is RK's k1: same field, same point).
template <int ARITH>
__device__ __forceinline__ void rk_advance(const Grid &g, const FieldPtrs &f, int interp, int order, const RkCoef &c,
                                           float px, float py, float pz, float k1x, float k1y, float k1z,
                                           float &ox, float &oy, float &oz) {
    if (order == 1) { ox = axpy(px, c.dt, k1x); oy = axpy(py, c.dt, k1y); oz = axpy(pz, c.dt, k1z); return; }
    float k2x, k2y, k2z;
    evaluate_any<ARITH>(g, f, interp, axpy(px, c.half_dt, k1x), axpy(py, c.half_dt, k1y), axpy(pz, c.half_dt, k1z), k2x, k2y, k2z);
    if (order == 2) { ox = axpy(px, c.dt, k2x); oy = axpy(py, c.dt, k2y); oz = axpy(pz, c.dt, k2z); return; }
    float k3x, k3y, k3z;
    if (order == 3) {
        evaluate_any<ARITH>(g, f, interp, axpy(px, c.three_quarter_dt, k2x), axpy(py, c.three_quarter_dt, k2y),
                            axpy(pz, c.three_quarter_dt, k2z), k3x, k3y, k3z);
        float sx = __fadd_rn(__fadd_rn(__fmul_rn(k1x, 2.0f), __fmul_rn(k2x, 3.0f)), __fmul_rn(k3x, 4.0f));
        float sy = __fadd_rn(__fadd_rn(__fmul_rn(k1y, 2.0f), __fmul_rn(k2y, 3.0f)), __fmul_rn(k3y, 4.0f));
        float sz = __fadd_rn(__fadd_rn(__fmul_rn(k1z, 2.0f), __fmul_rn(k2z, 3.0f)), __fmul_rn(k3z, 4.0f));
        ox = axpy(px, c.dt_over_9, sx); oy = axpy(py, c.dt_over_9, sy); oz = axpy(pz, c.dt_over_9, sz);
        return;
    }
    evaluate_any<ARITH>(g, f, interp, axpy(px, c.half_dt, k2x), axpy(py, c.half_dt, k2y), axpy(pz, c.half_dt, k2z), k3x, k3y, k3z);
    float k4x, k4y, k4z;
    evaluate_any<ARITH>(g, f, interp, axpy(px, c.dt, k3x), axpy(py, c.dt, k3y), axpy(pz, c.dt, k3z), k4x, k4y, k4z);
    float sx = __fadd_rn(__fadd_rn(__fadd_rn(k1x, __fmul_rn(k2x, 2.0f)), __fmul_rn(k3x, 2.0f)), k4x);
    float sy = __fadd_rn(__fadd_rn(__fadd_rn(k1y, __fmul_rn(k2y, 2.0f)), __fmul_rn(k3y, 2.0f)), k4y);
    float sz = __fadd_rn(__fadd_rn(__fadd_rn(k1z, __fmul_rn(k2z, 2.0f)), __fmul_rn(k3z, 2.0f)), k4z);
    ox = axpy(px, c.dt_over_6, sx); oy = axpy(py, c.dt_over_6, sy); oz = axpy(pz, c.dt_over_6, sz);
}

}  // namespace gfs
