/* shim over reference src/renderer/renderer.c for the `cray_ref_gpu` build (oracle/Makefile): the reference file is compiled in
 * place, unmodified.  INTEGRATION.md §2 changes ONE line of renderFrame (renderer.c:92: which function fills the thread-function
 * slot); to show that binding compiled and running without editing the reference source, this shim interposes on the call that
 * consumes the slot instead: renderFrame's threadStart(&r->state.threads[t]) (renderer.c:99).  With "-gpu" on the command line
 * (any -flag becomes an option tag, src/utils/args.c:207-209) or CRAY_GPU=1 in the environment, a slot that holds renderThread
 * is started with gpuRenderThread (c-ray_b200/integration/gpu_thread.c); everything else of renderFrame — tile queue, stats
 * loop, thread join — is the reference's own code. */
#include <stdlib.h>
#include <stdbool.h>
#include "utils/platform/thread.h"
#include "utils/args.h"
static int crb200_thread_start(struct crThread *t);
#define threadStart crb200_thread_start
#include "renderer/renderer.c"
#undef threadStart

void *gpuRenderThread(void *arg);

static int crb200_thread_start(struct crThread *t) {
	if (t->threadFunc == renderThread && (isSet("gpu") || getenv("CRAY_GPU"))) t->threadFunc = gpuRenderThread;
	return threadStart(t);
}
