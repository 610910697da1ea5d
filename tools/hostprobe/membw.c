/* tools/hostprobe/membw.c -- DIAGNOSTIC: host copy rates. usage: membw THREADS MODE PREFETCH_BYTES; mode 0 = memcpy from
 * DRAM into a cache-resident sink (what a consumer reading a ring slot does), 1 = the same with software prefetch,
 * 2 = DRAM to DRAM with non-temporal stores (what the capture source of the object-API bench does). */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <emmintrin.h>
#include <xmmintrin.h>
static unsigned char *src, *dst; static size_t total = 196608000; static int nt, mode, pfd;
static unsigned char sinkbuf[64][49152] __attribute__((aligned(64)));
static void nt_copy(void *d, const void *s, size_t n, int pf){
  const unsigned char *sp = s; unsigned char *dp = d; for (size_t i=0;i+64<=n;i+=64){ if(pf) _mm_prefetch((const char*)sp+i+pf,_MM_HINT_T0);
   __m128i a=_mm_loadu_si128((void*)(sp+i)), b=_mm_loadu_si128((void*)(sp+i+16)), c=_mm_loadu_si128((void*)(sp+i+32)), e=_mm_loadu_si128((void*)(sp+i+48));
   _mm_stream_si128((void*)(dp+i),a); _mm_stream_si128((void*)(dp+i+16),b); _mm_stream_si128((void*)(dp+i+32),c); _mm_stream_si128((void*)(dp+i+48),e);} _mm_sfence(); }
static void pf_copy(void *d, const void *s, size_t n, int pf){
  const unsigned char *sp = s; unsigned char *dp = d; for (size_t i=0;i+64<=n;i+=64){ if(pf) _mm_prefetch((const char*)sp+i+pf,_MM_HINT_T0);
   __m128i a=_mm_loadu_si128((void*)(sp+i)), b=_mm_loadu_si128((void*)(sp+i+16)), c=_mm_loadu_si128((void*)(sp+i+32)), e=_mm_loadu_si128((void*)(sp+i+48));
   _mm_storeu_si128((void*)(dp+i),a); _mm_storeu_si128((void*)(dp+i+16),b); _mm_storeu_si128((void*)(dp+i+32),c); _mm_storeu_si128((void*)(dp+i+48),e);} }
static void *w(void *p){ long i=(long)p; size_t lo=total*i/nt, hi=total*(i+1)/nt; lo&=~63ul; hi&=~63ul;
  for (size_t o=lo;o<hi;o+=49152){ size_t n = hi-o<49152?hi-o:49152;
    if(mode==0) memcpy(sinkbuf[i],src+o,n);            /* consumer: DRAM -> cached sink, glibc */
    else if(mode==1) pf_copy(sinkbuf[i],src+o,n,pfd);  /* consumer with prefetch */
    else if(mode==2) nt_copy(dst+o,src+o,n,pfd);       /* capture: DRAM -> NT */
  }
  return 0; }
int main(int c,char**v){ nt=atoi(v[1]); mode=atoi(v[2]); pfd=atoi(v[3]); src=aligned_alloc(4096,total); dst=aligned_alloc(4096,total); memset(src,1,total); memset(dst,2,total);
  double best=1e9; for(int r=0;r<5;r++){ struct timespec a,b; pthread_t t[64]; clock_gettime(CLOCK_MONOTONIC,&a); for(long i=0;i<nt;i++)pthread_create(&t[i],0,w,(void*)i); for(int i=0;i<nt;i++)pthread_join(t[i],0); clock_gettime(CLOCK_MONOTONIC,&b); double s=(b.tv_sec-a.tv_sec)+1e-9*(b.tv_nsec-a.tv_nsec); if(s<best)best=s;}
  printf("threads %d mode %d pf %d: %.1f GB/s\n",nt,mode,pfd,total/best/1e9); }
