/* clscalarfield.cpp -- see clscalarfield.h.  C++11, marshals to the C-ABI only. */
#include "clscalarfield.h"

#include <cstdlib>
#include <iostream>

#include "gfs_b200.h"

CLScalarField::CLScalarField() : _ctx(NULL), _isInitialized(false), _isOpenCLEnabled(true),
                                 _isMaxScalarFieldValueThresholdSet(false), _maxScalarFieldValueThreshold(1.0f),
                                 _kernelWorkLoadSize(1000) {
}

CLScalarField::~CLScalarField() {
    if (_ctx) {
        int err;
        gfs_destroy(_ctx, &err);
    }
}

void CLScalarField::_check(int err, const char *what) {
    if (err != GFS_SUCCESS) {
        std::cerr << "CLScalarField: " << what << " failed: " << gfs_get_error_message() << std::endl;
        std::abort();
    }
}

bool CLScalarField::initialize() {
    if (_isInitialized) {
        return true;
    }
    int err;
    _ctx = gfs_create(0, NULL, &err);
    if (err != GFS_SUCCESS) {
        std::cerr << "CLScalarField::initialize: " << gfs_get_error_message() << std::endl;
        return false;
    }
    _isInitialized = true;
    return true;
}

void CLScalarField::_splat(std::vector<vmath::vec3> &points, const float *values, double radius, vmath::vec3 offset,
                           double dx, Array3d<float> *field, Array3d<float> *weight) {
    _check(_isInitialized ? GFS_SUCCESS : GFS_FAIL, "initialize() has not been called");      // FLUIDSIM_ASSERT(_isInitialized)
    if (weight) {
        _check((field->width == weight->width && field->height == weight->height && field->depth == weight->depth)
                   ? GFS_SUCCESS : GFS_FAIL, "scalar and weight field dimensions differ");
    }
    std::vector<float> ones;
    if (!values) {
        ones.assign(points.size(), 1.0f);                     // addPoints: every point carries the value 1
        values = ones.empty() ? NULL : &ones[0];
    }
    float off[3] = {offset.x, offset.y, offset.z};
    int err;
    gfs_add_point_values(_ctx, points.empty() ? NULL : reinterpret_cast<const float *>(&points[0]), values,
                         (int64_t)points.size(), radius, off, dx, field->width, field->height, field->depth,
                         field->getRawArray(), weight ? weight->getRawArray() : NULL,
                         _isOpenCLEnabled ? 1 : 0, _isOpenCLEnabled ? GFS_FAST : GFS_EXACT, &err);
    _check(err, "gfs_add_point_values");
}

void CLScalarField::addPoints(std::vector<vmath::vec3> &points, double radius, vmath::vec3 offset, double dx,
                              Array3d<float> *field) {
    // IsotropicParticleMesher sets a threshold before every batch (src/isotropicparticlemesher.cpp:334-359); see
    // gfs_add_points for how the reference's three different skip rules map onto one order-independent rule
    _check(_isInitialized ? GFS_SUCCESS : GFS_FAIL, "initialize() has not been called");
    float off[3] = {offset.x, offset.y, offset.z};
    int err;
    gfs_add_points(_ctx, points.empty() ? NULL : reinterpret_cast<const float *>(&points[0]), (int64_t)points.size(), radius, off, dx,
                   field->width, field->height, field->depth, field->getRawArray(), _isOpenCLEnabled ? 1 : 0,
                   _isMaxScalarFieldValueThresholdSet ? 1 : 0, _maxScalarFieldValueThreshold,
                   _isOpenCLEnabled ? GFS_FAST : GFS_EXACT, &err);
    _check(err, "gfs_add_points");
}

void CLScalarField::addPoints(std::vector<vmath::vec3> &points, double radius, vmath::vec3 offset, double dx,
                              ScalarField &isfield) {
    addPoints(points, radius, offset, dx, isfield.getPointerToScalarField());
}

void CLScalarField::addPoints(std::vector<vmath::vec3> &points, ScalarField &isfield) {
    addPoints(points, isfield.getPointRadius(), isfield.getOffset(), isfield.getCellSize(), isfield.getPointerToScalarField());
}

void CLScalarField::addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, double radius,
                                   vmath::vec3 offset, double dx, Array3d<float> *field) {
    _check(points.size() == values.size() ? GFS_SUCCESS : GFS_FAIL, "points and values differ in length");
    _splat(points, values.empty() ? NULL : &values[0], radius, offset, dx, field, NULL);
}

void CLScalarField::addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, double radius,
                                   vmath::vec3 offset, double dx, Array3d<float> *scalarfield, Array3d<float> *weightfield) {
    _check(points.size() == values.size() ? GFS_SUCCESS : GFS_FAIL, "points and values differ in length");
    static const float zero = 0.0f;
    _splat(points, values.empty() ? &zero : &values[0], radius, offset, dx, scalarfield, weightfield);
}

void CLScalarField::addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, double radius,
                                   vmath::vec3 offset, double dx, ScalarField &isfield) {
    Array3d<float> *field = isfield.getPointerToScalarField();
    if (isfield.isWeightFieldEnabled()) {
        addPointValues(points, values, radius, offset, dx, field, isfield.getPointerToWeightField());
    } else {
        addPointValues(points, values, radius, offset, dx, field);
    }
}

void CLScalarField::addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, ScalarField &isfield) {
    addPointValues(points, values, isfield.getPointRadius(), isfield.getOffset(), isfield.getCellSize(), isfield);
}

void CLScalarField::setMaxScalarFieldValueThreshold(float val) {
    _isMaxScalarFieldValueThresholdSet = true;
    _maxScalarFieldValueThreshold = val;
}
void CLScalarField::setMaxScalarFieldValueThreshold() { _isMaxScalarFieldValueThresholdSet = false; }
bool CLScalarField::isMaxScalarFieldValueThresholdSet() { return _isMaxScalarFieldValueThresholdSet; }
double CLScalarField::getMaxScalarFieldValueThreshold() { return _maxScalarFieldValueThreshold; }

void CLScalarField::setDevicePreference(std::string) {}
void CLScalarField::setDevicePreferenceGPU() {}
void CLScalarField::setDevicePreferenceCPU() {}

std::string CLScalarField::getDeviceInfo() {
    if (!_isInitialized) {
        return "";
    }
    char buf[512];
    int err;
    gfs_device_info(_ctx, buf, (int)sizeof(buf), &err);
    return err == GFS_SUCCESS ? std::string(buf) + "\n" : std::string();
}
void CLScalarField::printDeviceInfo() { std::cout << getDeviceInfo() << std::endl; }
std::string CLScalarField::getKernelInfo() {
    return "CUDA kernels (sm_100a): gfs::k_splat_points (64-bit fixed-point accumulation), gfs::k_splat_points_store\n";
}
void CLScalarField::printKernelInfo() { std::cout << getKernelInfo() << std::endl; }
bool CLScalarField::isUsingGPU() { return _isInitialized; }
bool CLScalarField::isUsingCPU() { return false; }
void CLScalarField::disableOpenCL() { _isOpenCLEnabled = false; }
void CLScalarField::enableOpenCL() { _isOpenCLEnabled = true; }
bool CLScalarField::isOpenCLEnabled() { return _isOpenCLEnabled; }
int CLScalarField::getKernelWorkLoadSize() { return _kernelWorkLoadSize; }
void CLScalarField::setKernelWorkLoadSize(int n) { _kernelWorkLoadSize = n; }
