/*
 * oracle.h -- CPU restatement of GridFluidSim3D's PIC/FLIP particle<->grid transfer path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.  It is the checker, never the product: the
 * product path (gridfluidsim3d_b200/csrc) does not include, link or call anything in oracle/.
 *
 * Parity status: PINNED BY EXECUTION.  The reference ships no golden vectors or tests
 * (SURVEY.md §4), so every function here is checked bit-for-bit against the unmodified reference
 * sources compiled into oracle/_ref/libgfsref.so (tests/test_oracle_vs_ref.py, run wherever
 * /root/reference exists) and against fixtures generated from that library and committed under
 * tests/golden/ (oracle/make_golden.py; tests/test_oracle_golden.py, run everywhere).
 *
 * Conventions (all from the reference, paths relative to /root/reference):
 *   - positions / velocities: packed float triples, 12 B each (vmath::vec3, src/vmath.h:33-57)
 *   - grids: dense float arrays, flat = i + width*(j + height*k) (src/array3d.h:394-397);
 *     U is (I+1,J,K), V is (I,J+1,K), W is (I,J,K+1) (src/macvelocityfield.cpp:37-45)
 *   - material codes: 0 air, 1 fluid, 2 solid, one byte per cell (src/fluidmaterialgrid.h:29-33)
 *   - interpolation mode: 0 trilinear, 1 tricubic
 */
#ifndef GFS_ORACLE_H
#define GFS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_AIR   0
#define ORC_FLUID 1
#define ORC_SOLID 2

#define ORC_TRILINEAR 0
#define ORC_TRICUBIC  1

/* An active inflow source (src/fluidsimulation.cpp:2588-2594).  kind 0: sphere, centre p, radius a
 * (src/sphericalfluidsource.cpp:54-58).  kind 1: cuboid, min corner p, extents a,b,c
 * (src/cuboidfluidsource.cpp:67, src/aabb.cpp:123-126). */
typedef struct {
    int    kind;
    float  p[3];
    double a, b, c;
    float  velocity[3];
} orc_source_t;

int  orc_max_threads(void);

void orc_cell_index(const float *pos, long n, double dx, int *ijk);

void orc_sample(const float *pos, long n, const float *u, const float *v, const float *w,
                int I, int J, int K, double dx, int mode, int validate, float *out);

void orc_advect(const float *pos, long n, const float *u, const float *v, const float *w,
                int I, int J, int K, double dx, double dt, int order, int mode, float *out);

void orc_picflip(const float *pos, const float *vel, long n,
                 const float *u, const float *v, const float *w,
                 const float *us, const float *vs, const float *ws,
                 int I, int J, int K, double dx, double ratio, int mode, float *vel_out);

void orc_splat(const float *pos, const float *values, long value_stride, long n, double radius,
               const float *offset, double dx, int ni, int nj, int nk, float *field, float *weight);

void orc_apply_weight(float *field, const float *weight, long count);

void orc_border_solid(int I, int J, int K, unsigned char *material);

long orc_classify(const float *pos, long n, int I, int J, int K, double dx, unsigned char *material);

void orc_p2g_component(const float *pos, const float *vel, long n, int dir, int I, int J, int K, double dx,
                       const unsigned char *material, const orc_source_t *sources, int nsources,
                       float *out);

void orc_finish_component(float *field, const float *weight, int dir, int I, int J, int K, double dx,
                          const unsigned char *material, const orc_source_t *sources, int nsources, float *out);

void orc_p2g(const float *pos, const float *vel, long n, int I, int J, int K, double dx,
             unsigned char *material, const orc_source_t *sources, int nsources,
             float *u, float *v, float *w);

long orc_solid_test(const float *p0, float *p1, long n, int I, int J, int K, double dx,
                    const unsigned char *material, unsigned char *flags);

void orc_g2p_advect(const float *pos, const float *vel, long n,
                    const float *u, const float *v, const float *w,
                    const float *us, const float *vs, const float *ws,
                    int I, int J, int K, double dx, double ratio, double dt, int order, int mode,
                    const unsigned char *material, float *pos_out, float *vel_out, unsigned char *flags);

/* A14 with FluidSimulation::_resolveParticleSolidCellCollision (src/fluidsimulation.cpp:3145-3179) */
void orc_resolve_collision(const float *p0, const float *p1, int I, int J, int K, double dx,
                           const unsigned char *material, float *out);
long orc_collide(const float *p0, float *p1, long n, int I, int J, int K, double dx,
                 const unsigned char *material, unsigned char *flags);
void orc_g2p_advect_resolve(const float *pos, const float *vel, long n,
                            const float *u, const float *v, const float *w,
                            const float *us, const float *vs, const float *ws,
                            int I, int J, int K, double dx, double ratio, double dt, int order, int mode,
                            const unsigned char *material, float *pos_out, float *vel_out, unsigned char *flags);

/* MACVelocityField::extrapolateVelocityField (src/macvelocityfield.cpp:577-798), in place on u, v, w */
void orc_extrapolate(float *u, float *v, float *w, int I, int J, int K, const unsigned char *material, int nlayers);

/* Stages 6-8 (oracle_pressure.c): constant body forces, MICCG(0) pressure solve, pressure update.
 * src/fluidsimulation.cpp:2765-2805, 2870-2889, 2895-3061; src/pressuresolver.cpp:116-505 */
void orc_body_force(float *u, float *v, float *w, int I, int J, int K, const unsigned char *material,
                    const float force[3], double dt);
double orc_pressure_solve(const float *u, const float *v, const float *w, int I, int J, int K, double dx,
                          const unsigned char *material, double dt, double density, double tolerance, int max_iterations,
                          float *pressure, int *info);
void orc_apply_pressure(float *u, float *v, float *w, int I, int J, int K, double dx, const unsigned char *material,
                        const float *pressure, double dt, double density);

#ifdef __cplusplus
}
#endif
#endif
