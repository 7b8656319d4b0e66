/*
 * Fake OpenCL C++ bindings -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference (rlguy/GridFluidSim3D) cannot be compiled without <CL/cl.hpp>,
 * and this image has no OpenCL headers or ICD.  This stub provides just enough
 * of the API surface for the UNMODIFIED reference sources to compile and for
 * ParticleAdvector::initialize() / CLScalarField::initialize() to report
 * success, so that FluidSimulation can run with OpenCL *disabled* (its CPU
 * path -- the parity oracle).  Every enqueue* entry point aborts: it is never
 * reached when disableOpenCL*() has been called.
 *
 * Written from scratch for this repository; it contains no Khronos code.
 */
#ifndef GFS_FAKE_CL_HPP
#define GFS_FAKE_CL_HPP

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

typedef int                cl_int;
typedef unsigned int       cl_uint;
typedef unsigned long long cl_ulong;
typedef unsigned long long cl_device_type;
typedef unsigned long long cl_mem_flags;
typedef long               cl_context_properties;
typedef void*              cl_platform_id;
typedef void*              cl_device_id;
typedef void*              cl_kernel;
typedef unsigned int       cl_bool;

#define CL_SUCCESS 0
#define CL_TRUE    1
#define CL_FALSE   0

#define CL_DEVICE_TYPE_DEFAULT     (1ull << 0)
#define CL_DEVICE_TYPE_CPU         (1ull << 1)
#define CL_DEVICE_TYPE_GPU         (1ull << 2)
#define CL_DEVICE_TYPE_ACCELERATOR (1ull << 3)

#define CL_CONTEXT_DEVICES  0x1081
#define CL_CONTEXT_PLATFORM 0x1084

enum {
    CL_DEVICE_TYPE = 0x1000, CL_DEVICE_MAX_WORK_GROUP_SIZE, CL_DEVICE_MAX_WORK_ITEM_SIZES,
    CL_DEVICE_MAX_CLOCK_FREQUENCY, CL_DEVICE_MAX_MEM_ALLOC_SIZE, CL_DEVICE_GLOBAL_MEM_SIZE,
    CL_DEVICE_LOCAL_MEM_SIZE, CL_DEVICE_NAME, CL_DEVICE_VENDOR, CL_DRIVER_VERSION,
    CL_DEVICE_VERSION, CL_DEVICE_OPENCL_C_VERSION,
    CL_KERNEL_FUNCTION_NAME = 0x1190, CL_KERNEL_NUM_ARGS, CL_KERNEL_ATTRIBUTES,
    CL_KERNEL_WORK_GROUP_SIZE, CL_KERNEL_LOCAL_MEM_SIZE, CL_KERNEL_PRIVATE_MEM_SIZE,
    CL_KERNEL_PREFERRED_WORK_GROUP_SIZE_MULTIPLE
};

#define CL_MEM_READ_WRITE   (1ull << 0)
#define CL_MEM_WRITE_ONLY   (1ull << 1)
#define CL_MEM_READ_ONLY    (1ull << 2)
#define CL_MEM_USE_HOST_PTR (1ull << 3)

static inline void gfs_fake_cl_fill(unsigned int name, size_t size, void *value) {
    if (!value) return;
    std::memset(value, 0, size);
    unsigned long long v = 0;
    switch (name) {
        case CL_DEVICE_TYPE:                v = CL_DEVICE_TYPE_CPU; break;
        case CL_DEVICE_MAX_WORK_GROUP_SIZE: v = 256; break;   /* power of two in [32,256] */
        case CL_DEVICE_MAX_CLOCK_FREQUENCY: v = 1000; break;
        case CL_DEVICE_MAX_MEM_ALLOC_SIZE:  v = 1ull << 30; break;
        case CL_DEVICE_GLOBAL_MEM_SIZE:     v = 1ull << 32; break;
        case CL_DEVICE_LOCAL_MEM_SIZE:      v = 48 * 1024; break;
        case CL_KERNEL_NUM_ARGS:            v = 4; break;
        case CL_KERNEL_WORK_GROUP_SIZE:     v = 256; break;
        case CL_KERNEL_PREFERRED_WORK_GROUP_SIZE_MULTIPLE: v = 32; break;
        default: break;
    }
    std::memcpy(value, &v, size < sizeof(v) ? size : sizeof(v));
}

static inline cl_int clGetDeviceInfo(cl_device_id, unsigned int name, size_t size, void *value, size_t *) {
    gfs_fake_cl_fill(name, size, value); return CL_SUCCESS;
}
static inline cl_int clGetKernelInfo(cl_kernel, unsigned int name, size_t size, void *value, size_t *) {
    gfs_fake_cl_fill(name, size, value); return CL_SUCCESS;
}
static inline cl_int clGetKernelWorkGroupInfo(cl_kernel, cl_device_id, unsigned int name, size_t size,
                                              void *value, size_t *) {
    gfs_fake_cl_fill(name, size, value); return CL_SUCCESS;
}

namespace cl {

static inline void gfs_unreachable(const char *what) {
    std::fprintf(stderr, "fake CL/cl.hpp: %s called -- OpenCL must stay disabled in the oracle build\n", what);
    std::abort();
}

class Device {
public:
    Device() {}
    cl_device_id operator()() const { return (cl_device_id)0; }
    template <size_t N> cl_int getInfo(unsigned int, char (*out)[N]) const {
        std::snprintf(*out, N, "fake-opencl (disabled)"); return CL_SUCCESS;
    }
    cl_int getInfo(unsigned int, std::vector<size_t> *out) const {
        out->clear(); out->push_back(256); out->push_back(256); out->push_back(64); return CL_SUCCESS;
    }
};

class Platform {
public:
    cl_platform_id operator()() const { return (cl_platform_id)0; }
    static cl_int get(std::vector<Platform> *out) { out->assign(1, Platform()); return CL_SUCCESS; }
};

class Context {
public:
    Context() {}
    Context(cl_device_type, cl_context_properties *, void *, void *, cl_int *err) { if (err) *err = CL_SUCCESS; }
    template <int Name> std::vector<Device> getInfo() const { return std::vector<Device>(1, Device()); }
};

class Program {
public:
    typedef std::vector<std::pair<const char *, size_t> > Sources;
    Program() {}
    Program(const Context &, const Sources &) {}
    cl_int build(const std::vector<Device> &, const char *) { return CL_SUCCESS; }
};

struct LocalSpaceArg { size_t size; };
static inline LocalSpaceArg __local(size_t size) { LocalSpaceArg a; a.size = size; return a; }

class Kernel {
public:
    Kernel() {}
    Kernel(const Program &, const char *, cl_int *err) { if (err) *err = CL_SUCCESS; }
    cl_kernel operator()() const { return (cl_kernel)0; }
    template <size_t N> cl_int getInfo(unsigned int, char (*out)[N]) const {
        std::snprintf(*out, N, "fake-kernel"); return CL_SUCCESS;
    }
    template <typename T> cl_int setArg(unsigned int, const T &) { gfs_unreachable("Kernel::setArg"); return -1; }
};

class Buffer {
public:
    Buffer() {}
    Buffer(const Context &, cl_mem_flags, size_t, void * = NULL, cl_int *err = NULL) {
        gfs_unreachable("Buffer()"); if (err) *err = -1;
    }
};

class NDRange {
public:
    NDRange() {}
    NDRange(size_t) {}
    NDRange(size_t, size_t) {}
    NDRange(size_t, size_t, size_t) {}
};
static const NDRange NullRange;

class Event {
public:
    cl_int wait() const { gfs_unreachable("Event::wait"); return -1; }
};

class CommandQueue {
public:
    CommandQueue() {}
    CommandQueue(const Context &, const Device &, cl_ulong, cl_int *err) { if (err) *err = CL_SUCCESS; }
    cl_int enqueueNDRangeKernel(const Kernel &, const NDRange &, const NDRange &, const NDRange &,
                                const std::vector<Event> * = NULL, Event * = NULL) const {
        gfs_unreachable("enqueueNDRangeKernel"); return -1;
    }
    cl_int enqueueReadBuffer(const Buffer &, cl_bool, size_t, size_t, void *,
                             const std::vector<Event> * = NULL, Event * = NULL) const {
        gfs_unreachable("enqueueReadBuffer"); return -1;
    }
    cl_int finish() const { return CL_SUCCESS; }
};

}  // namespace cl

#endif
