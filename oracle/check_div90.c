// oracle/check_div90.c -- TEST INFRASTRUCTURE.  Proves the kernel shortcut used in
// rplidar_ros2_driver_b200/csrc/rpl_device.cuh::deg_to_key (division by 90 as one multiply + two FMAs)
// bit-identical to the reference expression v * 16384.f / 90.f (reference src/sdk/src/sl_lidar_driver.cpp:107-110)
// for EVERY float v in [0, 1024] -- the ascend pass only produces angles in [0, 720].
// Build: gcc -O2 -mfma -ffp-contract=off check_div90.c -lm   (tests/test_device_math_proofs.py)
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline float asf(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t asu(float f){uint32_t u; memcpy(&u,&f,4); return u;}
int main(){
  const float r = 1.0f/90.0f;
  uint64_t bad=0, badk=0, n=0;
  for (uint32_t u=0; u<=asu(1024.0f); ++u){
    float deg=asf(u); float x=deg*16384.0f;
    float q0=x*r; float e=fmaf(-q0,90.0f,x); float q=fmaf(e,r,q0);
    float ref=x/90.0f;
    if (asu(q)!=asu(ref)){ if(bad<5) printf("bad deg=%a q=%a ref=%a\n",deg,q,ref); ++bad; }
    if ((((uint32_t)q) & 0xFFFFu) != (((uint32_t)ref) & 0xFFFFu)) ++badk;
    ++n;
  }
  printf("checked %llu values, mismatches %llu, key mismatches %llu, r=%a\n",(unsigned long long)n,(unsigned long long)bad,(unsigned long long)badk,r);
  return (bad|badk)!=0;
}
