/*
 * oracle.c -- plain-C restatement of the reference CPU path (OpenCL disabled) for the
 * PIC/FLIP particle<->grid transfer of rlguy/GridFluidSim3D.  See oracle.h for status and rules.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Precision notes are the point of this file: the reference mixes float storage with double
 * arithmetic, and the restatement keeps each operation in the type the reference uses, in the
 * association the reference writes.  Build with -ffp-contract=off and no -march (oracle/Makefile)
 * so no operation is fused -- the reference's own flags (-O3 -std=c++11, CMakeLists.txt:34) emit
 * no FMAs on x86-64 either.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * A1. Grid3d::positionToGridIndex(vec3, dx)                           src/grid3d.h:58-63
 *     invdx is a double; float coordinate is widened by the multiply; floor; (int).
 * ---------------------------------------------------------------------------------------- */
static void cell_of(const float *p, double dx, int *i, int *j, int *k) {
    double invdx = 1.0 / dx;
    *i = (int)floor(p[0] * invdx);
    *j = (int)floor(p[1] * invdx);
    *k = (int)floor(p[2] * invdx);
}

void orc_cell_index(const float *pos, long n, double dx, int *ijk) {
    for (long p = 0; p < n; p++) cell_of(pos + 3 * p, dx, ijk + 3 * p, ijk + 3 * p + 1, ijk + 3 * p + 2);
}

/* ------------------------------------------------------------------------------------------
 * A9. Interpolation::cubic/bicubic/tricubicInterpolate                src/interpolation.cpp:26-46
 *     trilinearInterpolate                                            src/interpolation.cpp:51-59
 * ---------------------------------------------------------------------------------------- */
static double cubic(const double p[4], double x) {
    return p[1] + 0.5 * x * (p[2] - p[0] + x * (2.0 * p[0] - 5.0 * p[1] + 4.0 * p[2] - p[3] +
                                                  x * (3.0 * (p[1] - p[2]) + p[3] - p[0])));
}

static double bicubic(double p[4][4], double x, double y) {
    double arr[4];
    arr[0] = cubic(p[0], x);
    arr[1] = cubic(p[1], x);
    arr[2] = cubic(p[2], x);
    arr[3] = cubic(p[3], x);
    return cubic(arr, y);
}

static double tricubic(double p[4][4][4], double x, double y, double z) {
    double arr[4];
    arr[0] = bicubic(p[0], x, y);
    arr[1] = bicubic(p[1], x, y);
    arr[2] = bicubic(p[2], x, y);
    arr[3] = bicubic(p[3], x, y);
    return cubic(arr, z);
}

static double trilinear(const double p[8], double x, double y, double z) {
    return p[0] * (1 - x) * (1 - y) * (1 - z) +
           p[1] * x * (1 - y) * (1 - z) +
           p[2] * (1 - x) * y * (1 - z) +
           p[3] * (1 - x) * (1 - y) * z +
           p[4] * x * (1 - y) * z +
           p[5] * (1 - x) * y * z +
           p[6] * x * y * (1 - z) +
           p[7] * x * y * z;
}

/* ------------------------------------------------------------------------------------------
 * A8/A10. MACVelocityField::_interpolate{U,V,W}, _interpolateLinear{U,V,W}
 *         src/macvelocityfield.cpp:350-543.  One routine, parameterised by the component:
 *         comp 0 = U: array (I+1,J,K), y and z shifted by -0.5dx;  1 = V;  2 = W.
 *         Out-of-range taps read 0 (src/macvelocityfield.cpp:99-145, :473-480).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float *a[3];
    int I, J, K;
    double dx;
} field_t;

static int in_grid(double x, double y, double z, const field_t *f) {     /* src/grid3d.h:137-139 */
    return x >= 0 && y >= 0 && z >= 0 && x < f->dx * f->I && y < f->dx * f->J && z < f->dx * f->K;
}

static double tap(const field_t *f, int comp, int i, int j, int k) {
    int ni = f->I + (comp == 0), nj = f->J + (comp == 1), nk = f->K + (comp == 2);
    if (i < 0 || j < 0 || k < 0 || i >= ni || j >= nj || k >= nk) return 0.0;
    return (double)f->a[comp][(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)k)];
}

static double interp_component(const field_t *f, int comp, double x, double y, double z, int mode) {
    if (!in_grid(x, y, z, f)) return 0.0;

    if (comp != 0) x -= 0.5 * f->dx;
    if (comp != 1) y -= 0.5 * f->dx;
    if (comp != 2) z -= 0.5 * f->dx;

    double invdx = 1.0 / f->dx;                       /* positionToGridIndex(double...), grid3d.h:35-41 */
    int i = (int)floor(x * invdx), j = (int)floor(y * invdx), k = (int)floor(z * invdx);
    double gx = (double)i * f->dx, gy = (double)j * f->dx, gz = (double)k * f->dx;   /* grid3d.h:65-71 */

    double inv_dx = 1 / f->dx;
    double ix = (x - gx) * inv_dx, iy = (y - gy) * inv_dx, iz = (z - gz) * inv_dx;

    if (mode == ORC_TRICUBIC) {
        double pts[4][4][4];
        for (int pk = 0; pk < 4; pk++)
            for (int pj = 0; pj < 4; pj++)
                for (int pi = 0; pi < 4; pi++)
                    pts[pk][pj][pi] = tap(f, comp, pi + i - 1, pj + j - 1, pk + k - 1);
        return tricubic(pts, ix, iy, iz);
    }
    double pts[8];
    pts[0] = tap(f, comp, i,     j,     k);
    pts[1] = tap(f, comp, i + 1, j,     k);
    pts[2] = tap(f, comp, i,     j + 1, k);
    pts[3] = tap(f, comp, i,     j,     k + 1);
    pts[4] = tap(f, comp, i + 1, j,     k + 1);
    pts[5] = tap(f, comp, i,     j + 1, k + 1);
    pts[6] = tap(f, comp, i + 1, j + 1, k);
    pts[7] = tap(f, comp, i + 1, j + 1, k + 1);
    return trilinear(pts, ix, iy, iz);
}

/* MACVelocityField::evaluateVelocityAtPosition[Linear](vec3)        src/macvelocityfield.cpp:545-575
 * vec3 components are widened to double on the call; the three doubles are narrowed into a vec3. */
static void evaluate(const field_t *f, const float p[3], int mode, float out[3]) {
    double x = p[0], y = p[1], z = p[2];
    if (!in_grid(x, y, z, f)) { out[0] = out[1] = out[2] = 0.0f; return; }
    out[0] = (float)interp_component(f, 0, x, y, z, mode);
    out[1] = (float)interp_component(f, 1, x, y, z, mode);
    out[2] = (float)interp_component(f, 2, x, y, z, mode);
}

/* A11. ParticleAdvector::_tricubicInterpolateNoCL + _validateOutput   src/particleadvector.cpp:1124-1149
 * validate != 0: a NaN/Inf in any component zeroes the whole vector. */
void orc_sample(const float *pos, long n, const float *u, const float *v, const float *w,
                int I, int J, int K, double dx, int mode, int validate, float *out) {
    field_t f = {{u, v, w}, I, J, K, dx};
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n; p++) {
        float r[3];
        evaluate(&f, pos + 3 * p, mode, r);
        if (validate && (isinf(r[0]) || isnan(r[0]) || isinf(r[1]) || isnan(r[1]) || isinf(r[2]) || isnan(r[2]))) {
            r[0] = r[1] = r[2] = 0.0f;
        }
        out[3 * p] = r[0]; out[3 * p + 1] = r[1]; out[3 * p + 2] = r[2];
    }
}

/* ------------------------------------------------------------------------------------------
 * A12. ParticleAdvector::_RK4/_RK3/_RK2/_RK1                          src/particleadvector.cpp:1045-1078
 *      vec3 arithmetic is float (src/vmath.cpp:42-83: scalar*vector multiplies each component by
 *      the float scalar, sums are left to right); dt is double and each coefficient is narrowed to
 *      float exactly where the reference casts it.
 * ---------------------------------------------------------------------------------------- */
static void axpy(const float p[3], float s, const float k[3], float out[3]) {   /* p + s*k */
    out[0] = p[0] + k[0] * s; out[1] = p[1] + k[1] * s; out[2] = p[2] + k[2] * s;
}

static void rk_step(const field_t *f, const float p0[3], double dt, int order, int mode, float p1[3]) {
    float k1[3], k2[3], k3[3], k4[3], q[3], s[3];
    evaluate(f, p0, mode, k1);
    if (order == 1) { axpy(p0, (float)dt, k1, p1); return; }
    axpy(p0, (float)(0.5 * dt), k1, q);
    evaluate(f, q, mode, k2);
    if (order == 2) { axpy(p0, (float)dt, k2, p1); return; }
    if (order == 3) {
        axpy(p0, (float)(0.75 * dt), k2, q);
        evaluate(f, q, mode, k3);
        for (int c = 0; c < 3; c++) s[c] = (k1[c] * 2.0f + k2[c] * 3.0f) + k3[c] * 4.0f;
        axpy(p0, (float)(dt / 9.0f), s, p1);
        return;
    }
    axpy(p0, (float)(0.5 * dt), k2, q);
    evaluate(f, q, mode, k3);
    axpy(p0, (float)dt, k3, q);
    evaluate(f, q, mode, k4);
    for (int c = 0; c < 3; c++) s[c] = ((k1[c] + k2[c] * 2.0f) + k3[c] * 2.0f) + k4[c];
    axpy(p0, (float)(dt / 6.0f), s, p1);
}

void orc_advect(const float *pos, long n, const float *u, const float *v, const float *w,
                int I, int J, int K, double dx, double dt, int order, int mode, float *out) {
    field_t f = {{u, v, w}, I, J, K, dx};
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n; p++) rk_step(&f, pos + 3 * p, dt, order, mode, out + 3 * p);
}

/* ------------------------------------------------------------------------------------------
 * A13. FluidSimulation::_updateRangeOfMarkerParticleVelocities       src/fluidsimulation.cpp:3104-3129
 *      vnew, vold go through tricubicInterpolate (so are validated); vFLIP = (v + vnew) - vold;
 *      v = (float)ratio * vPIC + (float)(1 - ratio) * vFLIP, ratio a double (fluidsimulation.h:1161).
 * ---------------------------------------------------------------------------------------- */
static void validated(const field_t *f, const float p[3], int mode, float r[3]) {
    evaluate(f, p, mode, r);
    if (isinf(r[0]) || isnan(r[0]) || isinf(r[1]) || isnan(r[1]) || isinf(r[2]) || isnan(r[2])) {
        r[0] = r[1] = r[2] = 0.0f;
    }
}

static void picflip_one(const field_t *fn, const field_t *fs, const float p[3], const float vel[3],
                        double ratio, int mode, float out[3]) {
    float vnew[3], vold[3];
    validated(fn, p, mode, vnew);
    validated(fs, p, mode, vold);
    float rp = (float)ratio, rf = (float)(1 - ratio);
    for (int c = 0; c < 3; c++) {
        float flip = (vel[c] + vnew[c]) - vold[c];
        out[c] = vnew[c] * rp + flip * rf;
    }
}

void orc_picflip(const float *pos, const float *vel, long n,
                 const float *u, const float *v, const float *w,
                 const float *us, const float *vs, const float *ws,
                 int I, int J, int K, double dx, double ratio, int mode, float *vel_out) {
    field_t fn = {{u, v, w}, I, J, K, dx}, fs = {{us, vs, ws}, I, J, K, dx};
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n; p++) picflip_one(&fn, &fs, pos + 3 * p, vel + 3 * p, ratio, mode, vel_out + 3 * p);
}

/* ------------------------------------------------------------------------------------------
 * A3. ScalarField::addPointValue(p, value)                            src/scalarfield.cpp:167-201
 *     with setPointRadius (:40-46), _evaluateTricubicFieldFunctionForRadiusSquared (:569-571),
 *     Grid3d::getGridIndexBounds (src/grid3d.h:350-371), GridIndexToPosition(int..) (:73-75),
 *     vmath::dot in float (src/vmath.h:71-73), Array3d<float>::add (src/array3d.h:263-271).
 *     ACCUMULATES into field/weight (one ScalarField for the whole particle set -- what the
 *     reference does for N <= 5e6 and what its OpenCL path does for any N).
 *     values[p*value_stride] is the splatted scalar.
 * ---------------------------------------------------------------------------------------- */
void orc_splat(const float *pos, const float *values, long value_stride, long n, double radius,
               const float *offset, double dx, int ni, int nj, int nk, float *field, float *weight) {
    double r = radius;
    double coef1 = (4.0 / 9.0) * (1.0 / (r * r * r * r * r * r));
    double coef2 = (17.0 / 9.0) * (1.0 / (r * r * r * r));
    double coef3 = (22.0 / 9.0) * (1.0 / (r * r));
    double rsq = r * r;
    double inv = 1.0 / dx;

    for (long q = 0; q < n; q++) {
        float p[3] = {pos[3 * q] - offset[0], pos[3 * q + 1] - offset[1], pos[3 * q + 2] - offset[2]};
        double scale = (double)values[q * value_stride];

        int c[3];
        cell_of(p, dx, &c[0], &c[1], &c[2]);
        int size[3] = {ni, nj, nk}, lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            float cpos = (float)((float)c[a] * dx);          /* vec3((float)i*dx, ...) narrows */
            float trans = p[a] - cpos;
            int gmin = c[a] - (int)fmax(0, ceil((r - trans) * inv));
            int gmax = c[a] + (int)fmax(0, ceil((r - dx + trans) * inv));
            lo[a] = (int)fmax(gmin, 0);
            hi[a] = (int)fmin(gmax, size[a] - 1);
        }

        for (int k = lo[2]; k <= hi[2]; k++) {
            for (int j = lo[1]; j <= hi[1]; j++) {
                for (int i = lo[0]; i <= hi[0]; i++) {
                    float gx = (float)((float)i * dx), gy = (float)((float)j * dx), gz = (float)((float)k * dx);
                    float vx = gx - p[0], vy = gy - p[1], vz = gz - p[2];
                    float d2f = vx * vx + vy * vy + vz * vz;
                    double distsq = d2f;
                    if (distsq < rsq) {
                        double wgt = 1.0 - coef1 * distsq * distsq * distsq + coef2 * distsq * distsq - coef3 * distsq;
                        size_t idx = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)k);
                        field[idx] += (float)(wgt * scale);
                        if (weight) weight[idx] += (float)wgt;
                    }
                }
            }
        }
    }
}

/* A5. ScalarField::applyWeightField                                   src/scalarfield.cpp:90-106 */
void orc_apply_weight(float *field, const float *weight, long count) {
    for (long i = 0; i < count; i++) {
        float wgt = weight[i];
        if (wgt > 0.0) field[i] = field[i] / wgt;
    }
}

/* FluidSimulation::_initializeSolidCells                              src/fluidsimulation.cpp:1191-1213 */
void orc_border_solid(int I, int J, int K, unsigned char *m) {
    for (int k = 0; k < K; k++)
        for (int j = 0; j < J; j++)
            for (int i = 0; i < I; i++)
                if (i == 0 || j == 0 || k == 0 || i == I - 1 || j == J - 1 || k == K - 1)
                    m[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] = ORC_SOLID;
}

/* ------------------------------------------------------------------------------------------
 * A2. FluidSimulation::_updateFluidCells, marking loop                src/fluidsimulation.cpp:1998-2017
 *     interior fluid -> air, then material[cell(p)] = fluid.  The reference asserts the cell is not
 *     solid (:2015) after having removed such particles (:1933-1957); here a particle whose cell is
 *     solid or outside the grid is skipped and counted in the return value (0 for valid input).
 * ---------------------------------------------------------------------------------------- */
long orc_classify(const float *pos, long n, int I, int J, int K, double dx, unsigned char *m) {
    for (int k = 1; k < K - 1; k++)
        for (int j = 1; j < J - 1; j++)
            for (int i = 1; i < I - 1; i++) {
                size_t idx = (size_t)i + (size_t)I * ((size_t)j + (size_t)J * k);
                if (m[idx] == ORC_FLUID) m[idx] = ORC_AIR;
            }
    long bad = 0;
    for (long p = 0; p < n; p++) {
        int i, j, k;
        cell_of(pos + 3 * p, dx, &i, &j, &k);
        if (i < 0 || j < 0 || k < 0 || i >= I || j >= J || k >= K) { bad++; continue; }
        size_t idx = (size_t)i + (size_t)I * ((size_t)j + (size_t)J * k);
        if (m[idx] == ORC_SOLID) { bad++; continue; }
        m[idx] = ORC_FLUID;
    }
    return bad;
}

/* FluidMaterialGrid::isFaceBorderingMaterial{U,V,W}(.., fluid)        src/fluidmaterialgrid.cpp:119-143
 * (out-of-range cells read as solid, src/fluidmaterialgrid.cpp:25-29 -- never fluid) */
static int is_fluid(const unsigned char *m, int I, int J, int K, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= I || j >= J || k >= K) return 0;
    return m[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] == ORC_FLUID;
}

static int face_borders_fluid(const unsigned char *m, int I, int J, int K, int dir, int i, int j, int k) {
    int d[3] = {dir == 0, dir == 1, dir == 2};
    int idx[3] = {i, j, k}, size[3] = {I, J, K};
    if (idx[dir] == size[dir]) return is_fluid(m, I, J, K, i - d[0], j - d[1], k - d[2]);
    if (idx[dir] > 0) return is_fluid(m, I, J, K, i, j, k) || is_fluid(m, I, J, K, i - d[0], j - d[1], k - d[2]);
    return is_fluid(m, I, J, K, i, j, k);
}

/* source->containsPoint(p)    sphere: src/sphericalfluidsource.cpp:54-58   cuboid: src/aabb.cpp:123-126 */
static int source_contains(const orc_source_t *s, const float p[3]) {
    if (s->kind == 0) {
        float vx = p[0] - s->p[0], vy = p[1] - s->p[1], vz = p[2] - s->p[2];
        double lensq = vx * vx + vy * vy + vz * vz;          /* vmath::lengthsq is float */
        return lensq < s->a * s->a;
    }
    return p[0] >= s->p[0] && p[1] >= s->p[1] && p[2] >= s->p[2] &&
           p[0] < s->p[0] + s->a && p[1] < s->p[1] + s->b && p[2] < s->p[2] + s->c;
}

/* ------------------------------------------------------------------------------------------
 * A6 + A7. FluidSimulation::_computeVelocityScalarField               src/fluidsimulation.cpp:2526-2595
 *          FluidSimulation::_advectVelocityField{U,V,W}               src/fluidsimulation.cpp:2597-2730
 *   node grid = that component's face array; offset = face-centre offset narrowed to float;
 *   radius dx; splat; applyWeightField; isValueSet = weight > 1e-9; inflow override on set faces;
 *   faces bordering fluid take the value if set, else the mean of the 26 index-neighbours that are
 *   in range and "set" -- U tests fabs(value) > 0 (:2630), V and W test isValueSet (:2675, :2720).
 *   out must hold the face array; it is fully overwritten (cleared first, :2598).
 * ---------------------------------------------------------------------------------------- */
/* Everything of A6/A7 after the splat: `field`/`weight` hold the raw sums (they are normalised in place). */
void orc_finish_component(float *field, const float *weight, int dir, int I, int J, int K, double dx,
                          const unsigned char *material, const orc_source_t *sources, int nsources, float *out) {
    int ni = I + (dir == 0), nj = J + (dir == 1), nk = K + (dir == 2);
    size_t count = (size_t)ni * nj * nk;
    unsigned char *isset = (unsigned char *)calloc(count, 1);

    orc_apply_weight(field, weight, (long)count);

    double eps = 1e-9;
    for (size_t q = 0; q < count; q++) isset[q] = weight[q] > eps;

    for (int s = 0; s < nsources; s++) {                     /* :2489-2524 */
        float speed = (float)(double)sources[s].velocity[dir];
        for (int k = 0; k < nk; k++)
            for (int j = 0; j < nj; j++)
                for (int i = 0; i < ni; i++) {
                    size_t idx = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k);
                    if (!isset[idx]) continue;
                    float fp[3];                              /* Grid3d::FaceIndexToPosition{U,V,W}, grid3d.h:113-135 */
                    fp[0] = (float)(dir == 0 ? (float)i * dx : ((float)i + 0.5) * dx);
                    fp[1] = (float)(dir == 1 ? (float)j * dx : ((float)j + 0.5) * dx);
                    fp[2] = (float)(dir == 2 ? (float)k * dx : ((float)k + 0.5) * dx);
                    if (source_contains(&sources[s], fp)) field[idx] = speed;
                }
    }

    memset(out, 0, count * sizeof(float));
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (!face_borders_fluid(material, I, J, K, dir, i, j, k)) continue;
                size_t idx = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k);
                if (isset[idx]) { out[idx] = field[idx]; continue; }
                double avg = 0.0, wsum = 0.0;
                for (int nk_ = k - 1; nk_ <= k + 1; nk_++)       /* getNeighbourGridIndices26, grid3d.h:238-250 */
                    for (int nj_ = j - 1; nj_ <= j + 1; nj_++)
                        for (int ni_ = i - 1; ni_ <= i + 1; ni_++) {
                            if (ni_ == i && nj_ == j && nk_ == k) continue;
                            if (ni_ < 0 || nj_ < 0 || nk_ < 0 || ni_ >= ni || nj_ >= nj || nk_ >= nk) continue;
                            size_t nidx = (size_t)ni_ + (size_t)ni * ((size_t)nj_ + (size_t)nj * nk_);
                            int ok = dir == 0 ? (fabs(field[nidx]) > 0.0) : isset[nidx];
                            if (ok) { avg += field[nidx]; wsum += 1.0; }
                        }
                if (wsum > 0.0) out[idx] = (float)(avg / wsum);
            }
    free(isset);
}

void orc_p2g_component(const float *pos, const float *vel, long n, int dir, int I, int J, int K, double dx,
                       const unsigned char *material, const orc_source_t *sources, int nsources,
                       float *out) {
    int ni = I + (dir == 0), nj = J + (dir == 1), nk = K + (dir == 2);
    size_t count = (size_t)ni * nj * nk;
    float *field = (float *)calloc(count, sizeof(float));
    float *weight = (float *)calloc(count, sizeof(float));

    float offset[3];
    offset[0] = (float)(dir == 0 ? 0.0 : 0.5 * dx);
    offset[1] = (float)(dir == 1 ? 0.0 : 0.5 * dx);
    offset[2] = (float)(dir == 2 ? 0.0 : 0.5 * dx);

    orc_splat(pos, vel + dir, 3, n, dx, offset, dx, ni, nj, nk, field, weight);
    orc_finish_component(field, weight, dir, I, J, K, dx, material, sources, nsources, out);
    free(field); free(weight);
}

/* Stage 1 + stage 5 of FluidSimulation::_stepFluid for a given particle set:
 * classification (A2) then u, v, w (A6/A7).  material is updated in place. */
void orc_p2g(const float *pos, const float *vel, long n, int I, int J, int K, double dx,
             unsigned char *material, const orc_source_t *sources, int nsources,
             float *u, float *v, float *w) {
    orc_classify(pos, n, I, J, K, dx, material);
    float *out[3] = {u, v, w};
#pragma omp parallel for schedule(static, 1)
    for (int dir = 0; dir < 3; dir++)
        orc_p2g_component(pos, vel, n, dir, I, J, K, dx, material, sources, nsources, out[dir]);
}

/* ------------------------------------------------------------------------------------------
 * A14 (solid test only). FluidSimulation::_advanceRangeOfMarkerParticles   src/fluidsimulation.cpp:3198-3208
 *     g = cell(p1); out-of-range cells read as solid.  The reference then calls
 *     _resolveParticleSolidCellCollision (:3145-3179), whose every failure branch returns p0; the
 *     resolve routine itself is outside this round's scope, so a flagged
 *     particle keeps p0 and flags[p] = 1.  Returns the number of flagged particles.
 * ---------------------------------------------------------------------------------------- */
long orc_solid_test(const float *p0, float *p1, long n, int I, int J, int K, double dx,
                    const unsigned char *material, unsigned char *flags) {
    long hits = 0;
    for (long p = 0; p < n; p++) {
        int i, j, k;
        cell_of(p1 + 3 * p, dx, &i, &j, &k);
        int solid = (i < 0 || j < 0 || k < 0 || i >= I || j >= J || k >= K) ? 1
                  : material[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] == ORC_SOLID;
        /* NaN coordinates: (int)floor(NaN) is INT_MIN on x86-64 -> out of range -> solid */
        if (flags) flags[p] = (unsigned char)solid;
        if (solid) { p1[3 * p] = p0[3 * p]; p1[3 * p + 1] = p0[3 * p + 1]; p1[3 * p + 2] = p0[3 * p + 2]; hits++; }
    }
    return hits;
}

/* Stage 11 + stage 12 (without shuffle/cap): PIC/FLIP velocity update at p0, then RK advance through
 * the new field, then the solid test. */
void orc_g2p_advect(const float *pos, const float *vel, long n,
                    const float *u, const float *v, const float *w,
                    const float *us, const float *vs, const float *ws,
                    int I, int J, int K, double dx, double ratio, double dt, int order, int mode,
                    const unsigned char *material, float *pos_out, float *vel_out, unsigned char *flags) {
    orc_picflip(pos, vel, n, u, v, w, us, vs, ws, I, J, K, dx, ratio, mode, vel_out);
    orc_advect(pos, n, u, v, w, I, J, K, dx, dt, order, mode, pos_out);
    if (material) orc_solid_test(pos, pos_out, n, I, J, K, dx, material, flags);
}

/* ------------------------------------------------------------------------------------------
 * SURVEY 8(f) rank 1.  MACVelocityField::extrapolateVelocityField   src/macvelocityfield.cpp:786-798
 *   _resetExtrapolatedFluidVelocities :748-784   faces not bordering a fluid cell := 0
 *   _updateExtrapolationLayers        :603-619   layer 0 = fluid cells, layer L = non-solid cells still at -1 that are a
 *                                                6-neighbour of a non-solid layer L-1 cell (:577-601, in-place sweep)
 *   _extrapolateVelocitiesForLayerIndexU/V/W :692-744  a face that borders layer L, does not border layer L-1 and does
 *                                                not border a solid cell takes the mean (double; 0 when the sum is 0,
 *                                                :637-641) of its in-range 6 face neighbours, in the order
 *                                                (i-1,i+1,j-1,j+1,k-1,k+1) of grid3d.h:205-212, that border layer L-1
 *   "borders" on a boundary face looks at the one existing cell (macvelocityfield.h:159-173, fluidmaterialgrid.cpp:119-143)
 * ---------------------------------------------------------------------------------------- */
static int ext_cell_eq(const int *g, int I, int J, int K, int i, int j, int k, int value) {
    if (i < 0 || j < 0 || k < 0 || i >= I || j >= J || k >= K) return 0;
    return g[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] == value;
}

/* face (i,j,k) of direction dir borders a cell whose grid value equals `value` */
static int ext_face_borders(const int *g, int I, int J, int K, int dir, int i, int j, int k, int value) {
    return ext_cell_eq(g, I, J, K, i, j, k, value) ||
           ext_cell_eq(g, I, J, K, i - (dir == 0), j - (dir == 1), k - (dir == 2), value);
}

void orc_extrapolate(float *u, float *v, float *w, int I, int J, int K, const unsigned char *material, int nlayers) {
    size_t cells = (size_t)I * J * K;
    int *mat = (int *)malloc(cells * sizeof(int));
    int *layer = (int *)malloc(cells * sizeof(int));
    for (size_t c = 0; c < cells; c++) { mat[c] = material[c]; layer[c] = material[c] == ORC_FLUID ? 0 : -1; }
    float *fld[3] = {u, v, w};
    for (int dir = 0; dir < 3; dir++) {
        int ni = I + (dir == 0), nj = J + (dir == 1), nk = K + (dir == 2);
        for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++)
            if (!ext_face_borders(mat, I, J, K, dir, i, j, k, ORC_FLUID))
                fld[dir][(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)] = 0.0f;
    }
    static const int d6[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    for (int L = 1; L <= nlayers; L++)
        for (int k = 0; k < K; k++) for (int j = 0; j < J; j++) for (int i = 0; i < I; i++) {
            size_t c = (size_t)i + (size_t)I * ((size_t)j + (size_t)J * k);
            if (layer[c] != L - 1 || mat[c] == ORC_SOLID) continue;
            for (int q = 0; q < 6; q++) {
                int a = i + d6[q][0], b = j + d6[q][1], d = k + d6[q][2];
                if (a < 0 || b < 0 || d < 0 || a >= I || b >= J || d >= K) continue;
                size_t n = (size_t)a + (size_t)I * ((size_t)b + (size_t)J * d);
                if (layer[n] == -1 && mat[n] != ORC_SOLID) layer[n] = L;
            }
        }
    for (int L = 1; L <= nlayers; L++)
        for (int dir = 0; dir < 3; dir++) {
            int ni = I + (dir == 0), nj = J + (dir == 1), nk = K + (dir == 2);
            float *f = fld[dir];
            for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) {
                if (!(ext_face_borders(layer, I, J, K, dir, i, j, k, L) && !ext_face_borders(layer, I, J, K, dir, i, j, k, L - 1) &&
                      !ext_face_borders(mat, I, J, K, dir, i, j, k, ORC_SOLID))) continue;
                double sum = 0.0, weightsum = 0.0;
                for (int q = 0; q < 6; q++) {
                    int a = i + d6[q][0], b = j + d6[q][1], d = k + d6[q][2];
                    if (a < 0 || b < 0 || d < 0 || a >= ni || b >= nj || d >= nk) continue;
                    if (ext_face_borders(layer, I, J, K, dir, a, b, d, L - 1)) {
                        sum += f[(size_t)a + (size_t)ni * ((size_t)b + (size_t)nj * d)];
                        weightsum++;
                    }
                }
                double val = sum == 0.0 ? 0.0 : sum / weightsum;
                f[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)] = (float)val;
            }
        }
    free(mat); free(layer);
}

/* ------------------------------------------------------------------------------------------
 * A14, the resolve routine.
 * FluidSimulation::_resolveParticleSolidCellCollision   src/fluidsimulation.cpp:3145-3179
 *   Collision::getLineSegmentVoxelIntersection           src/collision.cpp:303-400   (voxel walk p0 -> p1)
 *   vmath::normalize                                     src/vmath.h:81-92, vmath.cpp:90-93
 *   Grid3d::GridIndexToPosition                          src/grid3d.h:83-85
 *   AABB::getMinPoint / getMaxPoint                      src/aabb.cpp:477-483
 *   Collision::rayIntersectsAABB                         src/collision.cpp:404-448   (incl. its dir.x-for-dir.z slip)
 * Every vmath::vec3 operation is a float operation (vec3 * double narrows the scalar first: vmath.cpp:71-84); the
 * voxel walk and the slab test run in double, as written.  Out-of-range cells read as solid (fluidmaterialgrid.cpp:25-29).
 * ---------------------------------------------------------------------------------------- */
static int cell_solid(const unsigned char *m, int I, int J, int K, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= I || j >= J || k >= K) return 1;
    return m[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] == ORC_SOLID;
}

void orc_resolve_collision(const float *p0, const float *p1, int I, int J, int K, double dx,
                           const unsigned char *material, float *out) {
    out[0] = p0[0]; out[1] = p0[1]; out[2] = p0[2];
    /* --- getLineSegmentVoxelIntersection: p0 *= invdx is vec3 *= float */
    double invdx = 1.0 / dx;
    float s = (float)invdx;
    float a0[3] = {p0[0] * s, p0[1] * s, p0[2] * s}, a1[3] = {p1[0] * s, p1[1] * s, p1[2] * s};
    int g0[3], g1[3], st[3], g[3];
    double gp[3], v[3];
    for (int a = 0; a < 3; a++) {
        g0[a] = (int)floor(a0[a]); g1[a] = (int)floor(a1[a]);
        st[a] = g1[a] > g0[a] ? 1 : (g1[a] < g0[a] ? -1 : 0);
        g[a] = g0[a];
        gp[a] = g0[a] + (g1[a] > g0[a] ? 1 : 0);
        v[a] = a1[a] == a0[a] ? 1 : a1[a] - a0[a];          /* float subtraction, widened */
    }
    double vxvy = v[0] * v[1], vxvz = v[0] * v[2], vyvz = v[1] * v[2];
    double err[3] = {(gp[0] - a0[0]) * vyvz, (gp[1] - a0[1]) * vxvz, (gp[2] - a0[2]) * vxvy};
    double derr[3] = {st[0] * vyvz, st[1] * vxvz, st[2] * vxvy};
    int found = 0, vox[3] = {0, 0, 0};
    for (long iter = 0; iter < 1000000; iter++) {
        if (g[0] >= 0 && g[1] >= 0 && g[2] >= 0 && g[0] < I && g[1] < J && g[2] < K &&
            material[(size_t)g[0] + (size_t)I * ((size_t)g[1] + (size_t)J * g[2])] == ORC_SOLID) {
            vox[0] = g[0]; vox[1] = g[1]; vox[2] = g[2]; found = 1; break;
        }
        if (g[0] == g1[0] && g[1] == g1[1] && g[2] == g1[2]) break;
        double xr = fabs(err[0]), yr = fabs(err[1]), zr = fabs(err[2]);
        if (st[0] != 0 && (st[1] == 0 || xr < yr) && (st[2] == 0 || xr < zr)) { g[0] += st[0]; err[0] += derr[0]; }
        else if (st[1] != 0 && (st[2] == 0 || yr < zr)) { g[1] += st[1]; err[1] += derr[1]; }
        else if (st[2] != 0) { g[2] += st[2]; err[2] += derr[2]; }
    }
    if (!found) return;
    /* --- raynorm = normalize(p1 - p0) */
    float d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    float lensq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    float len = (float)sqrt((double)lensq);
    float inv = (float)(1.0 / len);
    float rn[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
    /* --- voxel box */
    float bmin[3], bmax[3];
    for (int a = 0; a < 3; a++) { bmin[a] = (float)((float)vox[a] * dx); bmax[a] = bmin[a] + (float)dx; }
    /* --- rayIntersectsAABB */
    double eps = 1e-10;
    float dir[3] = {rn[0], rn[1], rn[2]};
    if (fabs(dir[0]) < eps) dir[0] = (float)(dir[0] < 0 ? -eps : eps);
    if (fabs(dir[1]) < eps) dir[1] = (float)(dir[1] < 0 ? -eps : eps);
    if (fabs(dir[0]) < eps) dir[2] = (float)(dir[2] < 0 ? -eps : eps);      /* sic: tests dir.x (collision.cpp:417) */
    float dirinv[3] = {(float)(1.0 / dir[0]), (float)(1.0 / dir[1]), (float)(1.0 / dir[2])};
    double t1 = (bmin[0] - p0[0]) * dirinv[0], t2 = (bmax[0] - p0[0]) * dirinv[0];
    double tmin = fmin(t1, t2), tmax = fmax(t1, t2);
    t1 = (bmin[1] - p0[1]) * dirinv[1]; t2 = (bmax[1] - p0[1]) * dirinv[1];
    tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
    t1 = (bmin[2] - p0[2]) * dirinv[2]; t2 = (bmax[2] - p0[2]) * dirinv[2];
    tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
    if (!(tmax > fmax(tmin, 0.0))) return;
    float tm = (float)tmin;
    float cp[3] = {p0[0] + dir[0] * tm, p0[1] + dir[1] * tm, p0[2] + dir[2] * tm};
    /* --- back off 0.05 dx along the ray */
    float back = (float)(0.05 * dx);
    float r[3] = {cp[0] - rn[0] * back, cp[1] - rn[1] * back, cp[2] - rn[2] * back};
    int i, j, k;
    cell_of(r, dx, &i, &j, &k);
    if (cell_solid(material, I, J, K, i, j, k)) return;
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

/* _advanceRangeOfMarkerParticles' post-pass (src/fluidsimulation.cpp:3198-3208) with the resolve: a particle whose
 * advected cell is solid goes through _resolveParticleSolidCellCollision.  flags as orc_solid_test.  Returns hits. */
long orc_collide(const float *p0, float *p1, long n, int I, int J, int K, double dx,
                 const unsigned char *material, unsigned char *flags) {
    long hits = 0;
    for (long p = 0; p < n; p++) {
        int i, j, k;
        cell_of(p1 + 3 * p, dx, &i, &j, &k);
        int solid = cell_solid(material, I, J, K, i, j, k);
        if (flags) flags[p] = (unsigned char)solid;
        if (solid) {
            float r[3];
            orc_resolve_collision(p0 + 3 * p, p1 + 3 * p, I, J, K, dx, material, r);
            p1[3 * p] = r[0]; p1[3 * p + 1] = r[1]; p1[3 * p + 2] = r[2];
            hits++;
        }
    }
    return hits;
}

/* orc_g2p_advect with the collision resolve instead of the bare solid test */
void orc_g2p_advect_resolve(const float *pos, const float *vel, long n,
                            const float *u, const float *v, const float *w,
                            const float *us, const float *vs, const float *ws,
                            int I, int J, int K, double dx, double ratio, double dt, int order, int mode,
                            const unsigned char *material, float *pos_out, float *vel_out, unsigned char *flags) {
    orc_picflip(pos, vel, n, u, v, w, us, vs, ws, I, J, K, dx, ratio, mode, vel_out);
    orc_advect(pos, n, u, v, w, I, J, K, dx, dt, order, mode, pos_out);
    if (material) orc_collide(pos, pos_out, n, I, J, K, dx, material, flags);
}
