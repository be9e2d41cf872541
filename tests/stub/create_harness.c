/* tests/stub/create_harness.c — runs crgpu_scene_create against the CUDA stub and prints the running checksum of the uploads. */
#include "crgpu.h"
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
unsigned long long stub_checksum(void); void stub_reset(void);
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){ struct crs_scene s; if (crscene_load(&s,argv[1])) return 1; crscene_set_config(&s,1920,1080,1000,32);
 for(int i=0;i<(argc>2?atoi(argv[2]):1);i++){ crgpu_scene *g=NULL; stub_reset(); double t0=now(); int rc=crgpu_scene_create(&s,0,&g); double t1=now(); printf("create rc=%d %.1f ms (%s) textures=%u\n",rc,1e3*(t1-t0),rc?crgpu_last_error():"ok",s.texture_count); if(g) crgpu_scene_destroy(g);} return 0; }
