#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
struct Maps { CUtensorMap m[2]; };
__device__ __forceinline__ uint32_t s32(const void*p){return (uint32_t)__cvta_generic_to_shared(p);}
template<int MODE>
__global__ void k(const __grid_constant__ Maps maps, const CUtensorMap* gmap, float* out, int x,int y,int z, int bytes){
  extern __shared__ unsigned char raw[];
  float* t = (float*)(((uintptr_t)raw+127)&~(uintptr_t)127);
  uint64_t* bar=(uint64_t*)(t+12*12*12);
  if(threadIdx.x==0){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;":::"memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(bar)),"r"(bytes):"memory");
    const CUtensorMap* mp = MODE==0 ? &maps.m[0] : gmap;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(s32(t)),"l"(mp),"r"(x),"r"(y),"r"(z),"r"(s32(bar)):"memory");
  }
  __syncthreads();
  uint32_t done=0; while(!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }":"=r"(done):"r"(s32(bar)):"memory");
  for(int i=threadIdx.x;i<12*12*12;i+=blockDim.x) out[i]=t[i];
}
typedef CUresult (*Fn)(CUtensorMap*,CUtensorMapDataType,cuuint32_t,void*,const cuuint64_t*,const cuuint64_t*,const cuuint32_t*,const cuuint32_t*,CUtensorMapInterleave,CUtensorMapSwizzle,CUtensorMapL2promotion,CUtensorMapFloatOOBfill);
int main(int argc,char**argv){
  int mode0=atoi(argv[1]); int bxi=atoi(argv[2]); int ni=atoi(argv[3]); int cx=atoi(argv[4]); int promo=atoi(argv[5]);
  int pitch=28,nj=20,nk=28; size_t n=(size_t)pitch*nj*nk; std::vector<float> h(n); for(size_t i=0;i<n;i++)h[i]=(float)i;
  float*d; cudaMalloc(&d,n*4); cudaMemcpy(d,h.data(),n*4,cudaMemcpyHostToDevice);
  void*p; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&p,cudaEnableDefault,&q); Fn fn=(Fn)p;
  Maps maps; cuuint64_t dims[3]={(cuuint64_t)ni,(cuuint64_t)nj,(cuuint64_t)nk}; cuuint64_t str[2]={(cuuint64_t)pitch*4,(cuuint64_t)pitch*nj*4}; cuuint32_t box[3]={(cuuint32_t)bxi,12,12}, es[3]={1,1,1};
  for(int l2=0;l2<1;l2++){
  CUresult r=fn(&maps.m[0],CU_TENSOR_MAP_DATA_TYPE_FLOAT32,3,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,(CUtensorMapL2promotion)promo,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d q=%d\n",(int)r,(int)q);}
  CUtensorMap* gm; cudaMalloc(&gm,sizeof(CUtensorMap)); cudaMemcpy(gm,&maps.m[0],sizeof(CUtensorMap),cudaMemcpyHostToDevice);
  float*o; cudaMalloc(&o,12*12*12*4); std::vector<float> ho(12*12*12);
  for(int mode=mode0;mode<mode0+1;mode++){
    if(mode==0) k<0><<<1,128,16*12*12*4+256>>>(maps,gm,o,cx,cx,cx,bxi*12*12*4); else k<1><<<1,128,16*12*12*4+256>>>(maps,gm,o,cx,cx,cx,bxi*12*12*4);
    cudaError_t e=cudaDeviceSynchronize(); printf("mode %d: %s\n",mode,cudaGetErrorString(e));
    if(e==cudaSuccess){cudaMemcpy(ho.data(),o,ho.size()*4,cudaMemcpyDeviceToHost); printf(" t[0]=%g t[2+12*(2+12*2)]=%g (expect 0) t[3+12*(2+12*2)]=%g (expect 1)\n",ho[0],ho[2+12*(2+12*2)],ho[3+12*(2+12*2)]);}
    else break;
  }
}
