/*
 * particleadvector.h -- drop-in replacement for the reference's OpenCL ParticleAdvector
 * (/root/reference/src/particleadvector.h:62-110).  Same class name, same public methods, same include guard:
 * pre-include this header (g++ -include) or put it in place of the reference's file, and FluidSimulation /
 * DiffuseParticleSimulation compile against it unchanged.  Everything is forwarded to the C-ABI of
 * include/gfs_b200.h; there is no OpenCL and no CPU loop in here.
 *
 * enable/disableOpenCL keep their meaning as "which arithmetic the caller gets":
 *   enabled  (default) : GFS_FAST  -- fp32 tap contraction on the GPU (the role the OpenCL kernel had)
 *   disabled           : GFS_EXACT -- the reference CPU path's own fp64 operation sequence, evaluated on the
 *                        GPU: results are bit-identical to ParticleAdvector::_*NoCL (particleadvector.cpp:1045-1149)
 */
#ifndef PARTICLEADVECTOR_H
#define PARTICLEADVECTOR_H

#include <string>
#include <vector>

#include "vmath.h"
#include <memory>

#include "macvelocityfield.h"

struct gfs_context;

class ParticleAdvector
{
public:
    ParticleAdvector();
    ~ParticleAdvector();

    bool initialize();
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

    void advectParticlesRK4(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                            std::vector<vmath::vec3> &output);
    void advectParticlesRK3(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                            std::vector<vmath::vec3> &output);
    void advectParticlesRK2(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                            std::vector<vmath::vec3> &output);
    void advectParticlesRK1(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt,
                            std::vector<vmath::vec3> &output);

    void tricubicInterpolate(std::vector<vmath::vec3> &particles, MACVelocityField *vfield,
                             std::vector<vmath::vec3> &output);
    // method will overwrite particles with output data
    void tricubicInterpolate(std::vector<vmath::vec3> &particles, MACVelocityField *vfield);

    // ---- not in the reference's class: the device-resident path (dropin/fluidsimulation_resident.cpp) ------------
    // the CUDA context of this accelerator (initialised on demand) and a slot for the per-simulation state of the
    // resident FluidSimulation stages, which cannot add members to FluidSimulation itself
    gfs_context *context();
    std::shared_ptr<void> residentState;

private:
    ParticleAdvector(const ParticleAdvector &);              // the context is not shareable
    ParticleAdvector &operator=(const ParticleAdvector &);

    void _advect(std::vector<vmath::vec3> &particles, MACVelocityField *vfield, double dt, int order,
                 std::vector<vmath::vec3> &output);
    void _check(int err, const char *what);

    gfs_context *_ctx;
    bool _isInitialized;
    bool _isOpenCLEnabled;
    int _kernelWorkLoadSize;
};

#endif
