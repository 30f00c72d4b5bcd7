// tools/check_libm_port.c -- CPU check (gcc -O2 -ffp-contract=off tools/check_libm_port.c -lm) that the fdlibm-style atan2f restated in
// cookiedough_b200/csrc/ckd_math.cuh reproduces the host libm bit for bit, and how often (float)exp / (float)pow in double differ from expf / powf.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline int32_t fw(float f){int32_t i; memcpy(&i,&f,4); return i;}
static inline float wf(int32_t i){float f; memcpy(&f,&i,4); return f;}
static const float atanhi[] = {4.6364760399e-01f,7.8539812565e-01f,9.8279368877e-01f,1.5707962513e+00f};
static const float atanlo[] = {5.0121582440e-09f,3.7748947079e-08f,3.4473217170e-08f,7.5497894159e-08f};
static const float aT[] = {3.3333334327e-01f,-2.0000000298e-01f,1.4285714924e-01f,-1.1111110449e-01f,9.0908870101e-02f,-7.6918758452e-02f,6.6610731184e-02f,-5.8335702866e-02f,4.9768779427e-02f,-3.6531571299e-02f,1.6285819933e-02f};
static float my_atanf(float x){
  float w,s1,s2,z; int32_t ix,hx,id; hx=fw(x); ix=hx&0x7fffffff;
  if(ix>=0x4c000000){ if(ix>0x7f800000) return x+x; if(hx>0) return atanhi[3]+atanlo[3]; else return -atanhi[3]-atanlo[3]; }
  if(ix<0x3ee00000){ if(ix<0x31000000){ return x; } id=-1; }
  else { x=fabsf(x);
    if(ix<0x3f980000){ if(ix<0x3f300000){id=0; x=(2.0f*x-1.0f)/(2.0f+x);} else {id=1; x=(x-1.0f)/(x+1.0f);} }
    else { if(ix<0x401c0000){id=2; x=(x-1.5f)/(1.0f+1.5f*x);} else {id=3; x=-1.0f/x;} } }
  z=x*x; w=z*z;
  s1=z*(aT[0]+w*(aT[2]+w*(aT[4]+w*(aT[6]+w*(aT[8]+w*aT[10])))));
  s2=w*(aT[1]+w*(aT[3]+w*(aT[5]+w*(aT[7]+w*aT[9]))));
  if(id<0) return x-x*(s1+s2);
  z=atanhi[id]-((x*(s1+s2)-atanlo[id])-x);
  return (hx<0)?-z:z;
}
static const float tiny=1.0e-30f, pi_o_4=7.8539818525e-01f, pi_o_2=1.5707963705e+00f, pi=3.1415927410e+00f, pi_lo=-8.7422776573e-08f;
static float my_atan2f(float y,float x){
  float z; int32_t k,m,hx,hy,ix,iy; hx=fw(x); ix=hx&0x7fffffff; hy=fw(y); iy=hy&0x7fffffff;
  if(ix>0x7f800000||iy>0x7f800000) return x+y;
  if(hx==0x3f800000) return my_atanf(y);
  m=((hy>>31)&1)|((hx>>30)&2);
  if(iy==0){ switch(m){case 0: case 1: return y; case 2: return pi+tiny; default: return -pi-tiny;} }
  if(ix==0) return (hy<0)? -pi_o_2-tiny: pi_o_2+tiny;
  if(ix==0x7f800000){ if(iy==0x7f800000){ switch(m){case 0: return pi_o_4+tiny; case 1: return -pi_o_4-tiny; case 2: return 3.0f*pi_o_4+tiny; default: return -3.0f*pi_o_4-tiny;} } else { switch(m){case 0: return 0.0f; case 1: return -0.0f; case 2: return pi+tiny; default: return -pi-tiny;} } }
  if(iy==0x7f800000) return (hy<0)? -pi_o_2-tiny: pi_o_2+tiny;
  k=(iy-ix)>>23;
  if(k>60) z=pi_o_2+0.5f*pi_lo; else if(hx<0&&k<-60) z=0.0f; else z=my_atanf(fabsf(y/x));
  switch(m){ case 0: return z; case 1: return wf(fw(z)^0x80000000); case 2: return pi-(z-pi_lo); default: return (z-pi_lo)-pi; }
}
int main(){
  uint64_t s=88172645463325252ULL; long bad=0, n=40000000; long bade=0, badp=0;
  for(long i=0;i<n;i++){
    s^=s<<13; s^=s>>7; s^=s<<17; float y=((int32_t)(s&0xffffff)-0x800000)/(float)0x200000;
    s^=s<<13; s^=s>>7; s^=s<<17; float x=((int32_t)(s&0xffffff)-0x800000)/(float)0x200000;
    if (i%7==0) { x*=1e-4f; } if (i%11==0) { y*=1e-5f; }
    float a=atan2f(y,x), b=my_atan2f(y,x);
    if(fw(a)!=fw(b)){ if(bad<5) printf("atan2f(%a,%a) libm %a mine %a\n",y,x,a,b); bad++; }
    float ex = -fabsf(y)*3.0f; float e1=expf(ex), e2=(float)exp((double)ex); if(fw(e1)!=fw(e2)) bade++;
    float px = fabsf(x)*0.25f, py = 1.0f+fabsf(y)*4.0f; float p1=powf(px,py), p2=(float)pow((double)px,(double)py); if(fw(p1)!=fw(p2)) badp++;
  }
  printf("atan2f mismatches: %ld / %ld ; expf vs (float)exp: %ld ; powf vs (float)pow: %ld\n",bad,n,bade,badp); return 0; }
