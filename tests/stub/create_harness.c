/* tests/stub/create_harness.c — runs crgpu_prepare + crgpu_scene_create_prepared against the CUDA stub and prints the FNV-1a
 * checksum of every section-sized chunk of the prepared slab (the bytes that cross PCIe per frame). */
#include "crgpu.h"
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
#include <stdint.h>
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){ struct crs_scene s; if (crscene_load(&s,argv[1])) return 1; crscene_set_config(&s,1920,1080,1000,32);
 for(int i=0;i<(argc>2?atoi(argv[2]):1);i++){ crgpu_prepared *p=NULL; crgpu_scene *g=NULL; double t0=now(); int rc=crgpu_prepare(&s,&p); double t1=now();
  if(!rc) rc=crgpu_scene_create_prepared(p,0,&g); double t2=now();
  printf("create rc=%d prepare %.1f ms upload %.1f ms (%s) textures=%u\n",rc,1e3*(t1-t0),1e3*(t2-t1),rc?crgpu_last_error():"ok",s.texture_count);
  if(p){ const void *slab; size_t n; crgpu_prepared_slab(p,&slab,&n); const unsigned char *b=slab; uint64_t h=1469598103934665603ull; for(size_t k=0;k<n;k++) h=(h^b[k])*1099511628211ull;
   printf("slab %zu bytes sum %016llx\n",n,(unsigned long long)h); }
  if(g) crgpu_scene_destroy(g); crgpu_prepared_free(p);} return 0; }
