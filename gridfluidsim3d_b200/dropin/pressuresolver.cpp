/* pressuresolver.cpp -- PressureSolver::solve (reference: src/pressuresolver.cpp:116-139) over the C-ABI.
 *
 * The reference builds a key map from its fluid-cell list, assembles b, A and the MIC(0) diagonal and runs CG on
 * vectors indexed by fluid cell.  Here the grids go to the device as they are (raw Array3d<float> storage + one material
 * byte per cell), the solve runs on dense per-cell vectors, and the fluid cells' entries are picked out of the returned
 * per-cell array in the caller's order.  One process-wide context serves every solver object: FluidSimulation creates
 * a PressureSolver per call (src/fluidsimulation.cpp:2881).
 */
#include "pressuresolver.h"

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>

#include "gfs_b200.h"

namespace {

struct SharedContext {
    gfs_context *ctx;
    SharedContext() : ctx(NULL) {}
    ~SharedContext() {
        if (ctx) {
            int err;
            gfs_destroy(ctx, &err);
        }
    }
    gfs_context *get() {
        if (!ctx) {
            int err;
            ctx = gfs_create(0, NULL, &err);
            if (err != GFS_SUCCESS) {
                std::cerr << "PressureSolver: " << gfs_get_error_message() << std::endl;
                std::abort();
            }
        }
        return ctx;
    }
};

SharedContext &shared() {
    static SharedContext s;
    return s;
}

}  // namespace

void VectorXd::fill(double value) {
    for (size_t i = 0; i < _vector.size(); i++) {
        _vector[i] = value;
    }
}

double VectorXd::dot(VectorXd &other) {
    FLUIDSIM_ASSERT(_vector.size() == other._vector.size());
    double sum = 0.0;
    for (size_t i = 0; i < _vector.size(); i++) {
        sum += _vector[i] * other._vector[i];
    }
    return sum;
}

double VectorXd::absMaxCoeff() {
    double m = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < _vector.size(); i++) {
        m = std::fmax(m, std::fabs(_vector[i]));
    }
    return m;
}

PressureSolver::PressureSolver() : _tolerance(1e-6), _maxIterations(200) {
}

PressureSolver::~PressureSolver() {
}

void PressureSolver::solve(PressureSolverParameters params, VectorXd &pressure) {
    GridIndexVector *cells = params.fluidCells;
    FluidMaterialGrid *mgrid = params.materialGrid;
    MACVelocityField *vfield = params.velocityField;
    FLUIDSIM_ASSERT(pressure.size() == cells->size());
    pressure.fill(0.0);
    if (cells->size() == 0) {
        return;
    }

    const int isize = mgrid->width, jsize = mgrid->height, ksize = mgrid->depth;
    std::vector<unsigned char> material((size_t)isize * jsize * ksize);
    size_t c = 0;
    for (int k = 0; k < ksize; k++) {
        for (int j = 0; j < jsize; j++) {
            for (int i = 0; i < isize; i++) {
                material[c++] = (unsigned char)(*mgrid)(i, j, k);
            }
        }
    }

    std::vector<double> dense(material.size());
    int err, iterations = 0;
    double residual = 0.0;
    gfs_pressure_solve_field(shared().get(), vfield->getRawArrayU(), vfield->getRawArrayV(), vfield->getRawArrayW(),
                             isize, jsize, ksize, params.cellwidth, &material[0], params.deltaTime, params.density,
                             _tolerance, _maxIterations, &dense[0], &iterations, &residual, &err);
    if (err != GFS_SUCCESS) {
        std::cerr << "PressureSolver::solve: " << gfs_get_error_message() << std::endl;
        std::abort();                                         /* FLUIDSIM_ASSERT semantics */
    }

    for (unsigned int idx = 0; idx < cells->size(); idx++) {
        GridIndex g = cells->at(idx);
        pressure[(int)idx] = dense[(size_t)g.i + (size_t)isize * ((size_t)g.j + (size_t)jsize * (size_t)g.k)];
    }

    if (params.logfile) {
        if (iterations >= _maxIterations) {
            params.logfile->log("Iterations limit reached.\t Estimated error : ", residual, 1);   /* src/pressuresolver.cpp:502 */
        } else if (iterations >= 0) {
            params.logfile->log("CG Iterations: ", iterations, 1);                                  /* :479 */
        }
    }
}
