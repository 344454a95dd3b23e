#include <cstdint>
__global__ void k(const float *p, float *o) {
    float a,b,c,d,e,f,g,h;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a),"=f"(b),"=f"(c),"=f"(d),"=f"(e),"=f"(f),"=f"(g),"=f"(h) : "l"(p + 8 * threadIdx.x * 37));
    o[threadIdx.x] = a+b+c+d+e+f+g+h;
}
