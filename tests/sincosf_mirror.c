#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
static const double SGN[4]={1.0,-1.0,-1.0,1.0};
static const double HPI_INV=0x1.45f306dc9c883p+23, HPI=0x1.921fb54442d18p+0;
static const double C0=1.0,C1=-0x1.ffffffd0c621cp-2,C2=0x1.55553e1068f19p-5,C3=-0x1.6c087e89a359dp-10,C4=0x1.99343027bf8c3p-16;
static const double S1=-0x1.555545995a603p-3,S2=0x1.1107605230bc4p-7,S3=-0x1.994eb3774cf24p-13;
static inline double sinpoly(double xs,double x2){ double s1_=fma(S3,x2,S2); double x3=x2*xs; double x5=x2*x3; double s=fma(x3,S1,xs); return fma(s1_,x5,s);}
static inline double cospoly(double x2,double sg){ double x4=x2*x2; double c1_=fma(sg*C1,x2,sg*C0); double c2_=fma(sg*C4,x2,sg*C3); double x6=x2*x4; double c=fma(x4,sg*C2,c1_); return fma(c2_,x6,c);}
static inline uint32_t top12(float y){uint32_t u; memcpy(&u,&y,4); return (u>>20)&0x7ff;}
float my_sinf(float y){ double x=y; uint32_t t=top12(y);
  if(t<0x3f4){ double x2=x*x; if(t<0x398) return y; return (float)sinpoly(x,x2);}  /* note small branch order: x3 = x*x2 */
  double r=x*HPI_INV; int n=((int32_t)r+0x800000)>>24; double xr=fma(-(double)n,HPI,x); double x2=xr*xr; double sg=(n&2)?-1.0:1.0;
  if((n&1)==0) return (float)sinpoly(xr*SGN[n&3],x2); else return (float)cospoly(x2,sg);}
float my_cosf(float y){ double x=y; uint32_t t=top12(y);
  if(t<0x3f4){ double x2=x*x; if(t<0x398) return 1.0f; return (float)cospoly(x2,1.0);}
  double r=x*HPI_INV; int n=((int32_t)r+0x800000)>>24; double xr=fma(-(double)n,HPI,x); double x2=xr*xr; double sg=(n&2)?-1.0:1.0;
  if((n&1)!=0) return (float)sinpoly(xr*SGN[n&3],x2); else return (float)cospoly(x2,sg);}
int main(){ float hi=6.2832f; uint32_t hib; memcpy(&hib,&hi,4); long bad_s=0,bad_c=0;
 #pragma omp parallel for reduction(+:bad_s,bad_c) schedule(static,1<<20)
 for(uint32_t b=0;b<=hib;b++){ float y; memcpy(&y,&b,4); float s=sinf(y),c=cosf(y); float ms=my_sinf(y),mc=my_cosf(y);
   if(memcmp(&s,&ms,4)) bad_s++; if(memcmp(&c,&mc,4)) bad_c++; }
 printf("checked %u values: sin mismatches %ld cos mismatches %ld\n",hib+1,bad_s,bad_c); return 0;}
