// Exhaustive check (all 2 139 095 041 non-negative floats) that (float)sqrt((double)x) == sqrtf(x): libmmd takes square
// roots through double (L/util/math.inl:27-45); the device uses the fp32 instruction.  gcc -O2 -ffp-contract=off ... -lm -lpthread
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <pthread.h>
typedef struct { uint32_t lo, hi; uint64_t bad; } job;
static void* run(void* a){ job* j=a; uint64_t bad=0; for(uint64_t u=j->lo; u<j->hi; ++u){ uint32_t b=(uint32_t)u; float x; memcpy(&x,&b,4); volatile double dx=(double)x; volatile double ds=sqrt(dx); float a1=(float)ds; float a2=sqrtf(x); uint32_t r1,r2; memcpy(&r1,&a1,4); memcpy(&r2,&a2,4); if(r1!=r2 && !(a1!=a1 && a2!=a2)) bad++; } j->bad=bad; return 0; }
int main(){ enum{T=8}; pthread_t th[T]; job jb[T]; uint64_t total=0x7F800001ull; /* all non-negative floats incl. +inf */
 for(int t=0;t<T;++t){ jb[t].lo=(uint32_t)(total*t/T); jb[t].hi=(uint32_t)(total*(t+1)/T); pthread_create(&th[t],0,run,&jb[t]); }
 uint64_t bad=0; for(int t=0;t<T;++t){ pthread_join(th[t],0); bad+=jb[t].bad; }
 printf("mismatches between (float)sqrt((double)x) and sqrtf(x) over all %llu non-negative floats: %llu\n",(unsigned long long)total,(unsigned long long)bad); return 0; }
