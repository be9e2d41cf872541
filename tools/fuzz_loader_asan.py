#!/usr/bin/env python3
"""Memory-safety fuzz of the scene loader (c-ray_b200/host/loader): builds it with -fsanitize=address,undefined, then loads
30 random valid scenes (leak check on) and 200 copies with one asset truncated / byte-flipped / spliced.  The loader may
reject a corrupted scene, it must never trip the sanitizers.  Round-1 result: 0 findings.  Usage: python tools/fuzz_loader_asan.py
"""
import tempfile
DRIVER = r'''#include "crloader.h"
#include <stdio.h>
#include <stdlib.h>
int main(int argc,char**argv){ int bad=0; for(int i=1;i<argc;i++){struct crs_scene s; int rc=crloader_load_json(&s,argv[i]); if(rc) {bad++; fprintf(stderr,"%s: rc=%d %s\n",argv[i],rc,crloader_last_error());} else free(s.owner);} printf("loaded %d files, %d errors\n",argc-1,bad); return 0;}
'''
import sys, os, json, random, subprocess, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'c-ray_b200'), ROOT]
import test_loader as T
base = tempfile.mkdtemp(prefix='crloader_fuzz_')
EXE = os.path.join(base, 'asan_loader')
open(os.path.join(base, 'driver.c'), 'w').write(DRIVER)
subprocess.run(['gcc', '-std=gnu11', '-O1', '-g', '-fsanitize=address,undefined', '-fno-omit-frame-pointer', '-ffp-contract=off',
                '-I' + os.path.join(ROOT, 'include'), '-o', EXE, os.path.join(base, 'driver.c')] +
               [os.path.join(ROOT, 'c-ray_b200', 'host', 'loader', f) for f in sorted(os.listdir(os.path.join(ROOT, 'c-ray_b200', 'host', 'loader'))) if f.endswith('.c')] +
               ['-lz', '-lm', '-lpthread'], check=True)
fails=0
for seed in range(30):
    d=f'{base}/{seed}'; os.makedirs(d)
    rng=random.Random(5000+seed)
    sc=T._fuzz_scene(rng,d)
    open(d+'/fuzz.json','w').write(json.dumps(sc))
    r=subprocess.run([EXE,'fuzz.json'],cwd=d,stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True,env=dict(os.environ,ASAN_OPTIONS='detect_leaks=1'))
    if 'ERROR' in r.stderr or 'runtime error' in r.stderr or r.returncode!=0:
        fails+=1; print(seed, r.stderr[-1500:])
print('clean-input fuzz: fails',fails)
# corruption fuzz
fails=0; runs=0
for seed in range(200):
    rng=random.Random(9000+seed)
    src=f'{base}/{seed%30}'; d=f'{base}/c{seed}'; shutil.copytree(src,d)
    files=[f for f in os.listdir(d) if not f.endswith('.crscene')]
    f=rng.choice(files); p=os.path.join(d,f); data=bytearray(open(p,'rb').read())
    if not data: continue
    mode=rng.randrange(4)
    if mode==0: data=data[:rng.randrange(len(data))]
    elif mode==1:
        for _ in range(rng.randrange(1,20)): data[rng.randrange(len(data))]=rng.randrange(256)
    elif mode==2:
        i=rng.randrange(len(data)); data[i:i]=bytes(rng.randrange(256) for _ in range(rng.randrange(1,64)))
    else:
        i=rng.randrange(len(data)); j=min(len(data),i+rng.randrange(1,200)); del data[i:j]
    open(p,'wb').write(bytes(data))
    r=subprocess.run([EXE,'fuzz.json'],cwd=d,stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True,errors='replace',timeout=120,env=dict(os.environ,ASAN_OPTIONS='detect_leaks=1'))
    runs+=1
    if 'AddressSanitizer' in r.stderr or 'runtime error' in r.stderr or r.returncode not in (0,):
        fails+=1; print('CORRUPT',seed,f,mode,'rc',r.returncode, r.stderr[-1200:])
    shutil.rmtree(d)
print('corruption fuzz runs',runs,'fails',fails)
shutil.rmtree(base, ignore_errors=True)
