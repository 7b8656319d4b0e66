/* pressuresolver.h -- drop-in replacement for the reference's src/pressuresolver.h: the same three public types
 * FluidSimulation::_updatePressureGrid uses (src/fluidsimulation.cpp:2870-2889) -- PressureSolverParameters, VectorXd,
 * PressureSolver::solve -- with the solve running on the GPU through gfs_pressure_solve_field (include/gfs_b200.h).
 *
 * Carries the reference's include guard, so `-include`-ing this file turns the reference's own header into a no-op
 * (oracle/Makefile target `dropin`).  The solver is the reference's MICCG(0), operation for operation (csrc/
 * gfs_pressure.cuh): same iteration count, pressures equal to the last place of the dot products' summation order.
 * C++11, no CUDA types.
 */
#ifndef PRESSURESOLVER_H
#define PRESSURESOLVER_H

#include <cstddef>
#include <vector>

#include "macvelocityfield.h"
#include "logfile.h"
#include "fluidmaterialgrid.h"
#include "gridindexvector.h"

/* field for field the reference's struct (src/pressuresolver.h:38-47): FluidSimulation fills it by member name */
struct PressureSolverParameters {
    double cellwidth;
    double density;
    double deltaTime;

    GridIndexVector *fluidCells;
    FluidMaterialGrid *materialGrid;
    MACVelocityField *velocityField;
    LogFile *logfile;
};

/* the solution vector type of the reference's interface (src/pressuresolver.h:53-77): one double per fluid cell */
class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(int size) : _vector((size_t)(size > 0 ? size : 0), 0.0) {}
    VectorXd(int size, double value) : _vector((size_t)(size > 0 ? size : 0), value) {}

    double operator[](int i) const { return _vector.at((size_t)i); }
    double &operator[](int i) { return _vector.at((size_t)i); }
    size_t size() { return _vector.size(); }

    void fill(double value);
    double dot(VectorXd &other);
    double absMaxCoeff();

    std::vector<double> _vector;
};

class PressureSolver {
public:
    PressureSolver();
    ~PressureSolver();

    /* pressure must have one entry per fluid cell, in the order of params.fluidCells (the reference asserts the same) */
    void solve(PressureSolverParameters params, VectorXd &pressure);

    /* src/pressuresolver.h:159-160 */
    double getTolerance() const { return _tolerance; }
    int getMaxIterations() const { return _maxIterations; }

private:
    double _tolerance;
    int _maxIterations;
};

#endif
