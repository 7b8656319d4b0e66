/*
 * oracle_pressure.c -- plain-C restatement of the reference's stages 6-8: constant body forces,
 * the MICCG(0) pressure solve and the pressure update of the MAC field.  TEST INFRASTRUCTURE (see oracle.h): only
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs may call it.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).  Pinned by execution:
 * tests/test_oracle_vs_ref.py compares each function bit for bit with the unmodified reference
 * (FluidSimulation::_applyBodyForcesToVelocityField / _updatePressureGrid / _applyPressureToVelocityField through
 * oracle/ref_harness.cpp).  Same build rules as oracle.c: -ffp-contract=off, no -march.
 *
 * Layout: cell arrays are i-fastest (c = i + I (j + J k)); u has (I+1) J K, v I (J+1) K, w I J (K+1) entries
 * (src/array3d.h:394-397).  material: 0 air, 1 fluid, 2 solid (src/fluidmaterialgrid.h:29-33).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AIR 0
#define FLUID 1
#define SOLID 2

typedef struct { int I, J, K; const unsigned char *m; } mat_t;

static int mcell(const mat_t *g, int i, int j, int k) {       /* the simulator's border cells are solid: a fluid cell has all six neighbours */
    if (i < 0 || j < 0 || k < 0 || i >= g->I || j >= g->J || k >= g->K) return SOLID;
    return g->m[(size_t)i + (size_t)g->I * ((size_t)j + (size_t)g->J * (size_t)k)];
}

/* FluidMaterialGrid::isFaceBorderingMaterialU/V/W                       src/fluidmaterialgrid.cpp:119-147 */
static int face_borders(const mat_t *g, int dir, int i, int j, int k, int what) {
    const int n[3] = {g->I, g->J, g->K};
    int idx[3] = {i, j, k};
    int lo[3] = {i, j, k};
    lo[dir] -= 1;
    if (idx[dir] == n[dir]) return mcell(g, lo[0], lo[1], lo[2]) == what;
    if (idx[dir] > 0) return mcell(g, i, j, k) == what || mcell(g, lo[0], lo[1], lo[2]) == what;
    return mcell(g, i, j, k) == what;
}

static size_t face_index(int dir, int I, int J, int i, int j, int k) {
    const size_t w = (size_t)I + (dir == 0), h = (size_t)J + (dir == 1);
    return (size_t)i + w * ((size_t)j + h * (size_t)k);
}

/* ------------------------------------------------------------------------------------------
 * Stage 6.  FluidSimulation::_applyConstantBodyForces                   src/fluidsimulation.cpp:2765-2805
 *     bodyForce.x * dt is float * double = double; MACVelocityField::addU narrows it to float and adds in float
 *     (src/macvelocityfield.cpp:204-210); a component whose force is exactly zero is skipped.
 * ---------------------------------------------------------------------------------------- */
void orc_body_force(float *u, float *v, float *w, int I, int J, int K, const unsigned char *material,
                    const float force[3], double dt) {
    const mat_t g = {I, J, K, material};
    float *f[3] = {u, v, w};
    for (int dir = 0; dir < 3; dir++) {
        if (!(fabs(force[dir]) > 0.0)) continue;
        const float add = (float)(force[dir] * dt);
        const int ni = I + (dir == 0), nj = J + (dir == 1), nk = K + (dir == 2);
        for (int k = 0; k < nk; k++)
            for (int j = 0; j < nj; j++)
                for (int i = 0; i < ni; i++)
                    if (face_borders(&g, dir, i, j, k, FLUID)) f[dir][face_index(dir, I, J, i, j, k)] += add;
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 7.  PressureSolver::solve                                        src/pressuresolver.cpp:116-139
 *     unknowns are the fluid cells in the order of FluidSimulation::_fluidCellIndices, which _updateFluidCells fills in
 *     k, j, i loop order (src/fluidsimulation.cpp:2019-2039) = ascending linear cell index.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int I, J, K, n;
    const unsigned char *m;
    int *key;                   /* cell -> unknown, -1 elsewhere (GridIndexKeyMap) */
    int *cell;                  /* unknown -> i, j, k */
    char *diag, *pi, *pj, *pk;  /* MatrixCell (src/pressuresolver.h:76-84) */
    double scale;               /* dt / (density dx^2) */
} sys_t;

static int key_at(const sys_t *s, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= s->I || j >= s->J || k >= s->K) return -1;
    return s->key[(size_t)i + (size_t)s->I * ((size_t)j + (size_t)s->J * (size_t)k)];
}

/* _calculateNegativeDivergenceVector                                     src/pressuresolver.cpp:164-211
 *     the six-term sum is float arithmetic left to right, widened once; scale = (double)(1.0f / (float)dx);
 *     the solid corrections are float products subtracted from / added to the double entry. */
static void negative_divergence(const sys_t *s, const float *u, const float *v, const float *w, double dx, double *b) {
    const mat_t g = {s->I, s->J, s->K, s->m};
    const int I = s->I, J = s->J;
    const double scale = 1.0f / (float)dx;
    for (int idx = 0; idx < s->n; idx++) {
        const int i = s->cell[3 * idx], j = s->cell[3 * idx + 1], k = s->cell[3 * idx + 2];
        const float sum = u[face_index(0, I, J, i + 1, j, k)] - u[face_index(0, I, J, i, j, k)] +
                          v[face_index(1, I, J, i, j + 1, k)] - v[face_index(1, I, J, i, j, k)] +
                          w[face_index(2, I, J, i, j, k + 1)] - w[face_index(2, I, J, i, j, k)];
        b[idx] = -scale * (double)sum;
    }
    const float usolid = 0.0f, vsolid = 0.0f, wsolid = 0.0f;
    for (int idx = 0; idx < s->n; idx++) {
        const int i = s->cell[3 * idx], j = s->cell[3 * idx + 1], k = s->cell[3 * idx + 2];
        if (mcell(&g, i - 1, j, k) == SOLID) b[idx] -= (float)scale * (u[face_index(0, I, J, i, j, k)] - usolid);
        if (mcell(&g, i + 1, j, k) == SOLID) b[idx] += (float)scale * (u[face_index(0, I, J, i + 1, j, k)] - usolid);
        if (mcell(&g, i, j - 1, k) == SOLID) b[idx] -= (float)scale * (v[face_index(1, I, J, i, j, k)] - vsolid);
        if (mcell(&g, i, j + 1, k) == SOLID) b[idx] += (float)scale * (v[face_index(1, I, J, i, j + 1, k)] - vsolid);
        if (mcell(&g, i, j, k - 1) == SOLID) b[idx] -= (float)scale * (w[face_index(2, I, J, i, j, k)] - wsolid);
        if (mcell(&g, i, j, k + 1) == SOLID) b[idx] += (float)scale * (w[face_index(2, I, J, i, j, k + 1)] - wsolid);
    }
}

/* _calculateMatrixCoefficients                                           src/pressuresolver.cpp:213-250 */
static void matrix_coefficients(sys_t *s) {
    const mat_t g = {s->I, s->J, s->K, s->m};
    for (int idx = 0; idx < s->n; idx++) {
        const int i = s->cell[3 * idx], j = s->cell[3 * idx + 1], k = s->cell[3 * idx + 2];
        int n = 0;
        if (mcell(&g, i - 1, j, k) != SOLID) n++;
        if (mcell(&g, i + 1, j, k) != SOLID) n++;
        if (mcell(&g, i, j - 1, k) != SOLID) n++;
        if (mcell(&g, i, j + 1, k) != SOLID) n++;
        if (mcell(&g, i, j, k - 1) != SOLID) n++;
        if (mcell(&g, i, j, k + 1) != SOLID) n++;
        s->diag[idx] = (char)n;
        s->pi[idx] = mcell(&g, i + 1, j, k) == FLUID;
        s->pj[idx] = mcell(&g, i, j + 1, k) == FLUID;
        s->pk[idx] = mcell(&g, i, j, k + 1) == FLUID;
    }
}

/* the three "minus" neighbours of an unknown, in the order -i, -j, -k: unknown number (-1: not a fluid cell), the MIC(0)
 * diagonal there, and the scaled plusi / plusj / plusk entries of that neighbour's matrix row (0.0 where it does not exist,
 * as the reference's ternaries give) */
typedef struct { int at[3]; double precon[3]; double plus[3][3]; /* [which plus: i, j, k][neighbour] */ } lower_t;

static void lower_neighbours(const sys_t *s, int idx, const double *precon, lower_t *L) {
    const int i = s->cell[3 * idx], j = s->cell[3 * idx + 1], k = s->cell[3 * idx + 2];
    const char *flag[3] = {s->pi, s->pj, s->pk};
    const double negscale = -s->scale;
    L->at[0] = key_at(s, i - 1, j, k); L->at[1] = key_at(s, i, j - 1, k); L->at[2] = key_at(s, i, j, k - 1);
    for (int d = 0; d < 3; d++) {
        const int n = L->at[d];
        L->precon[d] = n != -1 ? precon[n] : 0.0;
        for (int a = 0; a < 3; a++) L->plus[a][d] = n != -1 ? (double)flag[a][n] * negscale : 0.0;
    }
}

/* _calculatePreconditionerVector (modified incomplete Cholesky, level 0)  src/pressuresolver.cpp:252-310
 *     e = diag - sum_d (plus_d[d] precon[d])^2 - tau * sum_d plus_d[d] (sum of the other two plus entries of d) precon[d]^2,
 *     each sum left to right over d = i, j, k; the safety rule e < sigma diag -> e = diag; precon = 1 / sqrt(e) */
static void preconditioner(const sys_t *s, double *precon) {
    const double tau = 0.97, sigma = 0.25;
    for (int idx = 0; idx < s->n; idx++) {
        lower_t L;
        lower_neighbours(s, idx, precon, &L);
        const double diag = (double)s->diag[idx] * s->scale;
        double v[3], sq[3];
        for (int d = 0; d < 3; d++) { v[d] = L.plus[d][d] * L.precon[d]; sq[d] = L.precon[d] * L.precon[d]; }
        const double cross = L.plus[0][0] * (L.plus[1][0] + L.plus[2][0]) * sq[0] +
                             L.plus[1][1] * (L.plus[0][1] + L.plus[2][1]) * sq[1] +
                             L.plus[2][2] * (L.plus[0][2] + L.plus[1][2]) * sq[2];
        double e = diag - v[0] * v[0] - v[1] * v[1] - v[2] * v[2] - tau * cross;
        if (e < sigma * diag) e = diag;
        if (fabs(e) > 10e-9) precon[idx] = 1.0 / sqrt(e);
    }
}

/* _applyPreconditioner: forward substitution over ascending unknowns (q), then backward substitution over descending
 * unknowns (vect)                                                          src/pressuresolver.cpp:312-390 */
static void apply_preconditioner(const sys_t *s, const double *precon, const double *residual, double *q, double *vect) {
    const double negscale = -s->scale;
    for (int idx = 0; idx < s->n; idx++) {
        lower_t L;
        lower_neighbours(s, idx, precon, &L);
        double qn[3];
        for (int d = 0; d < 3; d++) qn[d] = L.at[d] != -1 ? q[L.at[d]] : 0.0;
        const double t = residual[idx] - L.plus[0][0] * L.precon[0] * qn[0] -
                                         L.plus[1][1] * L.precon[1] * qn[1] -
                                         L.plus[2][2] * L.precon[2] * qn[2];
        q[idx] = t * precon[idx];
    }
    for (int idx = s->n - 1; idx >= 0; idx--) {
        const int i = s->cell[3 * idx], j = s->cell[3 * idx + 1], k = s->cell[3 * idx + 2];
        const int up[3] = {key_at(s, i + 1, j, k), key_at(s, i, j + 1, k), key_at(s, i, j, k + 1)};
        const char own[3] = {s->pi[idx], s->pj[idx], s->pk[idx]};
        const double pc = precon[idx];
        double t = q[idx];
        for (int d = 0; d < 3; d++) {
            const double ahead = up[d] != -1 ? vect[up[d]] : 0.0;
            t = t - (double)own[d] * negscale * pc * ahead;        /* ((plus negscale) precon) vect, subtracted in the order i, j, k */
        }
        vect[idx] = t * pc;
    }
}

/* _applyMatrix                                                           src/pressuresolver.cpp:392-433 */
static void apply_matrix(const sys_t *s, const double *x, double *result) {
    const double scale = s->scale, negscale = -scale;
    for (int idx = 0; idx < s->n; idx++) {
        const int i = s->cell[3 * idx], j = s->cell[3 * idx + 1], k = s->cell[3 * idx + 2];
        double val = 0.0;
        int v;
        v = key_at(s, i - 1, j, k); if (v != -1) val += x[v];
        v = key_at(s, i + 1, j, k); if (v != -1) val += x[v];
        v = key_at(s, i, j - 1, k); if (v != -1) val += x[v];
        v = key_at(s, i, j + 1, k); if (v != -1) val += x[v];
        v = key_at(s, i, j, k - 1); if (v != -1) val += x[v];
        v = key_at(s, i, j, k + 1); if (v != -1) val += x[v];
        val *= negscale;
        val += (double)s->diag[idx] * scale * x[idx];
        result[idx] = val;
    }
}

static double dot(const double *a, const double *b, int n) {      /* VectorXd::dot, src/pressuresolver.cpp:61-70 */
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += a[i] * b[i];
    return sum;
}

static double abs_max(const double *a, int n) {                   /* VectorXd::absMaxCoeff, :72-81 */
    double mx = -INFINITY;
    for (int i = 0; i < n; i++) if (fabs(a[i]) > mx) mx = fabs(a[i]);
    return mx;
}

/* solve + _solvePressureSystem                                           src/pressuresolver.cpp:116-139, 452-505
 * and FluidSimulation::_updatePressureGrid's narrowing to the float grid src/fluidsimulation.cpp:2870-2889.
 * pressure: I J K floats (0 outside fluid cells).  info[0] = CG iterations done (the reference's iterationNumber at
 * return; -1 when the right-hand side was already below the tolerance), info[1] = 1 if the iteration limit was reached.
 * Returns the last residual max-norm. */
double orc_pressure_solve(const float *u, const float *v, const float *w, int I, int J, int K, double dx,
                          const unsigned char *material, double dt, double density, double tolerance, int max_iterations,
                          float *pressure, int *info) {
    const size_t cells = (size_t)I * J * K;
    sys_t s;
    s.I = I; s.J = J; s.K = K; s.m = material;
    s.key = (int *)malloc(cells * sizeof(int));
    int n = 0;
    for (size_t c = 0; c < cells; c++) s.key[c] = material[c] == FLUID ? n++ : -1;
    s.n = n;
    s.cell = (int *)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(int));
    for (int k = 0, c = 0; k < K; k++)
        for (int j = 0; j < J; j++)
            for (int i = 0; i < I; i++, c++)
                if (s.key[c] >= 0) { s.cell[3 * s.key[c]] = i; s.cell[3 * s.key[c] + 1] = j; s.cell[3 * s.key[c] + 2] = k; }
    s.diag = (char *)calloc((size_t)n + 1, 1); s.pi = (char *)calloc((size_t)n + 1, 1);
    s.pj = (char *)calloc((size_t)n + 1, 1); s.pk = (char *)calloc((size_t)n + 1, 1);
    s.scale = dt / (density * dx * dx);

    double *b = (double *)calloc((size_t)n + 1, sizeof(double)), *x = (double *)calloc((size_t)n + 1, sizeof(double));
    double *precon = (double *)calloc((size_t)n + 1, sizeof(double)), *aux = (double *)calloc((size_t)n + 1, sizeof(double));
    double *search = (double *)calloc((size_t)n + 1, sizeof(double)), *q = (double *)calloc((size_t)n + 1, sizeof(double));
    double *residual = b;
    double err = 0.0;
    info[0] = -1; info[1] = 0;
    memset(pressure, 0, cells * sizeof(float));

    negative_divergence(&s, u, v, w, dx, b);
    err = n > 0 ? abs_max(b, n) : 0.0;
    if (n > 0 && !(err < tolerance)) {
        matrix_coefficients(&s);
        preconditioner(&s, precon);

        apply_preconditioner(&s, precon, residual, q, aux);
        memcpy(search, aux, (size_t)n * sizeof(double));
        double sigma = dot(aux, residual, n);
        int it = 0;
        info[1] = 1;
        while (it < max_iterations) {
            apply_matrix(&s, search, aux);
            const double alpha = sigma / dot(aux, search, n);
            for (int i = 0; i < n; i++) x[i] += search[i] * alpha;
            for (int i = 0; i < n; i++) residual[i] += aux[i] * (-alpha);
            err = abs_max(residual, n);
            if (err < tolerance) { info[1] = 0; break; }
            memset(q, 0, (size_t)n * sizeof(double));          /* "VectorXd q(_matSize)" is a fresh zero vector every call */
            apply_preconditioner(&s, precon, residual, q, aux);
            const double sigma_new = dot(aux, residual, n);
            const double beta = sigma_new / sigma;
            for (int i = 0; i < n; i++) search[i] = aux[i] * 1.0 + search[i] * beta;
            sigma = sigma_new;
            it++;
        }
        info[0] = it;
        for (size_t c = 0; c < cells; c++) if (s.key[c] >= 0) pressure[c] = (float)x[s.key[c]];
    }
    free(s.key); free(s.cell); free(s.diag); free(s.pi); free(s.pj); free(s.pk);
    free(b); free(x); free(precon); free(aux); free(search); free(q);
    return err;
}

/* ------------------------------------------------------------------------------------------
 * Stage 8.  FluidSimulation::_applyPressureToVelocityField               src/fluidsimulation.cpp:2895-3061
 *     temp = 0 on faces bordering a solid; faces bordering fluid and no solid get U - scale (p1 - p0) in double
 *     (_applyPressureToFaceU: both cells non-solid there, so only its first branch is reachable), narrowed by setU;
 *     _commitTemporaryVelocityFieldValues copies temp back on faces bordering fluid; other faces keep their value.
 * ---------------------------------------------------------------------------------------- */
void orc_apply_pressure(float *u, float *v, float *w, int I, int J, int K, double dx, const unsigned char *material,
                        const float *pressure, double dt, double density) {
    const mat_t g = {I, J, K, material};
    float *f[3] = {u, v, w};
    const double scale = dt / (density * dx);
    for (int dir = 0; dir < 3; dir++) {
        const int ni = I + (dir == 0), nj = J + (dir == 1), nk = K + (dir == 2);
        for (int k = 0; k < nk; k++)
            for (int j = 0; j < nj; j++)
                for (int i = 0; i < ni; i++) {
                    if (!face_borders(&g, dir, i, j, k, FLUID)) continue;
                    float *x = &f[dir][face_index(dir, I, J, i, j, k)];
                    if (face_borders(&g, dir, i, j, k, SOLID)) { *x = (float)0.0; continue; }
                    const int ci = i - (dir == 0), cj = j - (dir == 1), ck = k - (dir == 2);
                    /* (both cells exist: a fluid cell on the domain border would have tripped the reference's asserts) */
                    const double p0 = (ci < 0 || cj < 0 || ck < 0) ? 0.0 : pressure[(size_t)ci + (size_t)I * ((size_t)cj + (size_t)J * (size_t)ck)];
                    const double p1 = (i >= I || j >= J || k >= K) ? 0.0 : pressure[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * (size_t)k)];
                    const double next = *x - scale * (p1 - p0);
                    *x = (float)next;
                }
    }
}
