/*
 * clscalarfield.h -- drop-in replacement for the reference's OpenCL CLScalarField
 * (/root/reference/src/clscalarfield.h:64-125): same class name, public methods and include guard, forwarding to
 * gfs_add_point_values of include/gfs_b200.h.
 *
 *   enabled  (default) : the OpenCL path's semantics -- contributions are ADDED to the caller's arrays
 *                        (src/clscalarfield.cpp:1427-1455), GFS_FAST kernel weights
 *   disabled           : the CPU path's semantics -- the arrays are OVERWRITTEN with this batch's field
 *                        (src/clscalarfield.cpp:1521-1551), fp64 kernel weights
 * In both modes the accumulation is 64-bit fixed point (order-independent); it agrees with the reference's
 * particle-order fp32 accumulation to fp32 rounding, not bit for bit.
 * The max-scalar-field-value early-out (src/scalarfield.cpp:182-184) depends on particle order and is only used
 * by the surface mesher, which is outside this library's scope: setting it makes addPoints abort.
 */
#ifndef CLSCALARFIELD_H
#define CLSCALARFIELD_H

#include <string>
#include <vector>

#include "vmath.h"
#include "array3d.h"
#include "scalarfield.h"

struct gfs_context;

class CLScalarField
{
public:
    CLScalarField();
    ~CLScalarField();

    bool initialize();
    void addPoints(std::vector<vmath::vec3> &points, double radius, vmath::vec3 offset, double dx, Array3d<float> *field);
    void addPoints(std::vector<vmath::vec3> &points, double radius, vmath::vec3 offset, double dx, ScalarField &field);
    void addPoints(std::vector<vmath::vec3> &points, ScalarField &field);

    void addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, double radius, vmath::vec3 offset,
                        double dx, Array3d<float> *field);
    void addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, double radius, vmath::vec3 offset,
                        double dx, Array3d<float> *scalarfield, Array3d<float> *weightfield);
    void addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, double radius, vmath::vec3 offset,
                        double dx, ScalarField &field);
    void addPointValues(std::vector<vmath::vec3> &points, std::vector<float> &values, ScalarField &field);

    void setMaxScalarFieldValueThreshold(float val);
    void setMaxScalarFieldValueThreshold();
    bool isMaxScalarFieldValueThresholdSet();
    double getMaxScalarFieldValueThreshold();

    void setDevicePreference(std::string devtype);
    void setDevicePreferenceGPU();
    void setDevicePreferenceCPU();

    void printDeviceInfo();
    std::string getDeviceInfo();
    void printKernelInfo();
    std::string getKernelInfo();
    bool isUsingGPU();
    bool isUsingCPU();
    void disableOpenCL();
    void enableOpenCL();
    bool isOpenCLEnabled();
    int getKernelWorkLoadSize();
    void setKernelWorkLoadSize(int n);

private:
    CLScalarField(const CLScalarField &);
    CLScalarField &operator=(const CLScalarField &);

    void _splat(std::vector<vmath::vec3> &points, const float *values, double radius, vmath::vec3 offset, double dx,
                Array3d<float> *field, Array3d<float> *weight);
    void _check(int err, const char *what);

    gfs_context *_ctx;
    bool _isInitialized;
    bool _isOpenCLEnabled;
    bool _isMaxScalarFieldValueThresholdSet;
    float _maxScalarFieldValueThreshold;
    int _kernelWorkLoadSize;
};

#endif
