/* host memory bandwidth probe: T threads, each streams over its own slice (fill with non-temporal
 * stores, plain memcpy, read-sum).  Decides whether a host-side expansion of compacted G could beat PCIe. */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
typedef struct{double*a,*b;size_t n;int mode;double sum;}job;
static void*run(void*p){job*j=p;size_t n=j->n;
 if(j->mode==0){__m256d v=_mm256_set1_pd(1.5);for(size_t i=0;i<n;i+=4)_mm256_stream_pd(j->a+i,v);_mm_sfence();}
 else if(j->mode==1)memcpy(j->a,j->b,n*8);
 else{double s=0;for(size_t i=0;i<n;i++)s+=j->b[i];j->sum=s;}
 return 0;}
int main(int argc,char**argv){int T=argc>1?atoi(argv[1]):16;size_t per=(size_t)(argc>2?atol(argv[2]):64)<<20;/* doubles per thread */
 job*js=calloc(T,sizeof(job));pthread_t*th=calloc(T,sizeof(pthread_t));
 for(int t=0;t<T;t++){js[t].a=aligned_alloc(64,per*8);js[t].b=aligned_alloc(64,per*8);js[t].n=per;memset(js[t].a,0,per*8);memset(js[t].b,1,per*8);}
 const char*nm[3]={"nt-fill (write)","memcpy (read+write)","read-sum"};
 for(int mode=0;mode<3;mode++){double best=1e9;for(int r=0;r<3;r++){double t0=now();for(int t=0;t<T;t++){js[t].mode=mode;pthread_create(&th[t],0,run,&js[t]);}for(int t=0;t<T;t++)pthread_join(th[t],0);double dt=now()-t0;if(dt<best)best=dt;}
  double bytes=(double)T*per*8*(mode==1?2:1);printf("%d threads %-20s %.1f GB/s\n",T,nm[mode],bytes/best/1e9);}
 return 0;}
