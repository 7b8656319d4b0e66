// TMA probe (kept as evidence for DESIGN.md section 8): tiled cp.async.bulk.tensor loads need a 16-byte aligned INNER box
// coordinate.  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 profiles/tma_alignment_probe.cu; run: ./a.out 2 <flags>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void*p){return (uint32_t)__cvta_generic_to_shared(p);}
// MODE 0: 1-D bulk copy (no tensor map); MODE 1: 2-D tensor map; MODE 2: 3-D tensor map
template<int MODE>
__global__ void k(const __grid_constant__ CUtensorMap map, const float* g, float* out, int flags){ int cx = ((flags>>8)&255)-64, cy = ((flags>>16)&255)-64;
  __shared__ __align__(128) float ts[4096];
  __shared__ __align__(8) uint64_t bars;
  extern __shared__ unsigned char raw[];
  float* td = (float*)(((uintptr_t)raw+127)&~(uintptr_t)127);
  float* t = (flags&2) ? td : ts;
  uint64_t& bar = (flags&2) ? *(uint64_t*)(td+4096) : bars;
  if(threadIdx.x==0){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar)));
    if(flags&1) asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); else asm volatile("fence.proxy.async.shared::cta;":::"memory");
  }
  if(!(flags&8)) __syncthreads();
  if(threadIdx.x==0){
    uint32_t bytes = MODE==0 ? 1024u : (MODE==1 ? 16u*16u*4u : ((flags&4)? 12u*12u*12u*4u : 16u*8u*8u*4u));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&bar)),"r"(bytes):"memory");
    if(MODE==0) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(s32(t)),"l"(g),"r"(1024),"r"(s32(&bar)):"memory");
    if(MODE==1) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"::"r"(s32(t)),"l"(&map),"r"(0),"r"(0),"r"(s32(&bar)):"memory");
    if(MODE==2) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"::"r"(s32(t)),"l"(&map),"r"(cx),"r"(cy),"r"(cy),"r"(s32(&bar)):"memory");
  }
  uint32_t done=0; while(!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }":"=r"(done):"r"(s32(&bar)):"memory");
  for(int i=threadIdx.x;i<256;i+=blockDim.x) out[i]=t[i];
}
typedef CUresult (*Fn)(CUtensorMap*,CUtensorMapDataType,cuuint32_t,void*,const cuuint64_t*,const cuuint64_t*,const cuuint32_t*,const cuuint32_t*,CUtensorMapInterleave,CUtensorMapSwizzle,CUtensorMapL2promotion,CUtensorMapFloatOOBfill);
int main(int argc,char**argv){
  int mode=atoi(argv[1]); int flags=atoi(argv[2]);
  size_t n=64*64*64; std::vector<float> h(n); for(size_t i=0;i<n;i++)h[i]=(float)i;
  float*d; cudaMalloc(&d,n*4); cudaMemcpy(d,h.data(),n*4,cudaMemcpyHostToDevice);
  void*p; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&p,cudaEnableDefault,&q); Fn fn=(Fn)p;
  CUtensorMap map; memset(&map,0,sizeof(map));
  if(mode==1){ cuuint64_t dims[2]={64,64}; cuuint64_t str[1]={64*4}; cuuint32_t box[2]={16,16}, es[2]={1,1};
    printf("encode2d %d\n",(int)fn(&map,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,2,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_NONE,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  if(mode==2){ cuuint64_t dims[3]={64,64,64}; cuuint64_t str[2]={64*4,64*64*4}; if(flags&16){dims[0]=25;dims[1]=20;dims[2]=28;str[0]=28*4;str[1]=28*20*4;} cuuint32_t box[3]={16,8,8}, es[3]={1,1,1}; if(flags&4){box[0]=12;box[1]=12;box[2]=12;}
    printf("encode3d %d\n",(int)fn(&map,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,3,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_NONE,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  float*o; cudaMalloc(&o,256*4); std::vector<float> ho(256);
  if(mode==0) k<0><<<1,128,4096*4+256>>>(map,d,o,flags); if(mode==1) k<1><<<1,128,4096*4+256>>>(map,d,o,flags); if(mode==2) k<2><<<1,128,4096*4+256>>>(map,d,o,flags);
  cudaError_t e=cudaDeviceSynchronize(); printf("mode %d: %s\n",mode,cudaGetErrorString(e));
  if(e==cudaSuccess){cudaMemcpy(ho.data(),o,ho.size()*4,cudaMemcpyDeviceToHost); printf(" t[0..3]=%g %g %g t[16]=%g t[128]=%g\n",ho[0],ho[1],ho[2],ho[16],ho[128]);}
}
