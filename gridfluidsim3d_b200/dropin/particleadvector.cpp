/* particleadvector.cpp -- see particleadvector.h.  C++11, no CUDA in this file: it only marshals to the C-ABI. */
#include "particleadvector.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "gfs_b200.h"

static_assert(sizeof(vmath::vec3) == 3 * sizeof(float), "vmath::vec3 must be three packed floats");

ParticleAdvector::ParticleAdvector() : _ctx(NULL), _isInitialized(false), _isOpenCLEnabled(true), _kernelWorkLoadSize(1000) {
}

ParticleAdvector::~ParticleAdvector() {
    if (_ctx) {
        int err;
        gfs_destroy(_ctx, &err);
    }
}

void ParticleAdvector::_check(int err, const char *what) {
    // FLUIDSIM_ASSERT semantics (src/fluidsimassert.h:1-17): print and abort
    if (err != GFS_SUCCESS) {
        std::cerr << "ParticleAdvector: " << what << " failed: " << gfs_get_error_message() << std::endl;
        std::abort();
    }
}

bool ParticleAdvector::initialize() {
    if (_isInitialized) {
        return true;
    }
    int err;
    _ctx = gfs_create(0, NULL, &err);
    if (err != GFS_SUCCESS) {
        std::cerr << "ParticleAdvector::initialize: " << gfs_get_error_message() << std::endl;
        return false;                                   // same contract as src/particleadvector.cpp:30-58
    }
    _isInitialized = true;
    return true;
}

void ParticleAdvector::setDevicePreference(std::string) {}
void ParticleAdvector::setDevicePreferenceGPU() {}
void ParticleAdvector::setDevicePreferenceCPU() {}

std::string ParticleAdvector::getDeviceInfo() {
    if (!_isInitialized) {
        return "";
    }
    char buf[512];
    int err;
    gfs_device_info(_ctx, buf, (int)sizeof(buf), &err);
    return err == GFS_SUCCESS ? std::string(buf) + "\n" : std::string();
}

void ParticleAdvector::printDeviceInfo() { std::cout << getDeviceInfo() << std::endl; }

std::string ParticleAdvector::getKernelInfo() {
    return "CUDA kernels (sm_100a): gfs::k_sample, gfs::k_advect (fused RK1-4), tricubic and trilinear\n";
}

void ParticleAdvector::printKernelInfo() { std::cout << getKernelInfo() << std::endl; }
gfs_context *ParticleAdvector::context() {
    if (!_isInitialized && !initialize()) {
        _check(GFS_FAIL, "initialize");
    }
    return _ctx;
}

bool ParticleAdvector::isUsingGPU() { return _isInitialized; }
bool ParticleAdvector::isUsingCPU() { return false; }
void ParticleAdvector::disableOpenCL() { _isOpenCLEnabled = false; }
void ParticleAdvector::enableOpenCL() { _isOpenCLEnabled = true; }
bool ParticleAdvector::isOpenCLEnabled() { return _isOpenCLEnabled; }
int ParticleAdvector::getKernelWorkLoadSize() { return _kernelWorkLoadSize; }
void ParticleAdvector::setKernelWorkLoadSize(int n) { _kernelWorkLoadSize = n; }

void ParticleAdvector::_advect(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt, int order,
                               std::vector<vmath::vec3> &output) {
    if (!_isInitialized && !initialize()) {
        _check(GFS_FAIL, "initialize");
    }
    int isize, jsize, ksize;
    vfield->getGridDimensions(&isize, &jsize, &ksize);
    output.clear();                                         // src/particleadvector.cpp:1084-1086
    output.resize(particles.size());
    if (particles.empty()) {
        return;
    }
    int err;
    gfs_advect(_ctx, reinterpret_cast<const float *>(&particles[0]), (int64_t)particles.size(),
               vfield->getRawArrayU(), vfield->getRawArrayV(), vfield->getRawArrayW(), isize, jsize, ksize,
               vfield->getGridCellSize(), dt, order, GFS_TRICUBIC, _isOpenCLEnabled ? GFS_FAST : GFS_EXACT,
               reinterpret_cast<float *>(&output[0]), &err);
    _check(err, "gfs_advect");
}

void ParticleAdvector::advectParticlesRK4(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                                          std::vector<vmath::vec3> &output) { _advect(particles, vfield, dt, 4, output); }
void ParticleAdvector::advectParticlesRK3(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                                          std::vector<vmath::vec3> &output) { _advect(particles, vfield, dt, 3, output); }
void ParticleAdvector::advectParticlesRK2(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                                          std::vector<vmath::vec3> &output) { _advect(particles, vfield, dt, 2, output); }
void ParticleAdvector::advectParticlesRK1(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                                          std::vector<vmath::vec3> &output) { _advect(particles, vfield, dt, 1, output); }

void ParticleAdvector::tricubicInterpolate(std::vector<vmath::vec3> &particles, MACVelocityField *vfield,
                                           std::vector<vmath::vec3> &output) {
    if (!_isInitialized && !initialize()) {
        _check(GFS_FAIL, "initialize");
    }
    // the callee sizes the output (src/particleadvector.cpp:1127-1130) and validates it (:1139-1149)
    if (output.size() < particles.size()) {
        output.resize(particles.size());
    }
    if (particles.empty()) {
        return;
    }
    int isize, jsize, ksize;
    vfield->getGridDimensions(&isize, &jsize, &ksize);
    int err;
    gfs_sample(_ctx, reinterpret_cast<const float *>(&particles[0]), (int64_t)particles.size(),
               vfield->getRawArrayU(), vfield->getRawArrayV(), vfield->getRawArrayW(), isize, jsize, ksize,
               vfield->getGridCellSize(), GFS_TRICUBIC, _isOpenCLEnabled ? GFS_FAST : GFS_EXACT, 1,
               reinterpret_cast<float *>(&output[0]), &err);
    _check(err, "gfs_sample");
}

void ParticleAdvector::tricubicInterpolate(std::vector<vmath::vec3> &particles, MACVelocityField *vfield) {
    std::vector<vmath::vec3> out(particles.size());
    tricubicInterpolate(particles, vfield, out);
    particles.swap(out);                                    // "method will overwrite particles with output data"
}
