// oracle/check_div4000.c -- TEST INFRASTRUCTURE.  Proves the kernel shortcut used in
// rplidar_ros2_driver_b200/csrc/rpl_device.cuh::dist_to_m (division by 4000 as one
// multiply + two FMAs) bit-identical to the reference expression dist_mm_q2 / 4000.0f
// (reference src/rplidar_node.cpp:588) for EVERY float a u32 can convert to.
// Build: gcc -O2 -mfma -ffp-contract=off check_div4000.c -lm   (tests/test_device_math_proofs.py)
// exhaustive check: for every float x that a u32 can convert to, is
//   q0 = x*r; e = fma(-q0, 4000, x); q = fma(e, r, q0)     (r = RN(1/4000))
// bit-identical to x / 4000.0f ?
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline float asf(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t asu(float f){uint32_t u; memcpy(&u,&f,4); return u;}
int main(){
  const float r = 1.0f/4000.0f;
  uint64_t bad=0, n=0;
  // all integers 0..2^24
  for (uint32_t i=0;i<=(1u<<24);++i){ float x=(float)i; float q0=x*r; float e=fmaf(-q0,4000.0f,x); float q=fmaf(e,r,q0); if (asu(q)!=asu(x/4000.0f)){ if(bad<5) printf("bad x=%u q=%a ref=%a\n",i,q,x/4000.0f); ++bad;} ++n; }
  // all floats in [2^24, 2^32]
  for (uint32_t u=asu(16777216.0f); u<=asu(4294967296.0f); ++u){ float x=asf(u); float q0=x*r; float e=fmaf(-q0,4000.0f,x); float q=fmaf(e,r,q0); if (asu(q)!=asu(x/4000.0f)){ if(bad<5) printf("bad x=%a\n",x); ++bad;} ++n; }
  printf("checked %llu values, mismatches %llu, r=%a\n",(unsigned long long)n,(unsigned long long)bad,r);
  return bad!=0;
}
