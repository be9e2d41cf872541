/*
 * cr_threads.c — the loader's tiny fork/join helper (pthreads).  The reference parallelises scene construction only
 * across meshes (one thread per mesh BVH, src/datatypes/scene.c:51-79); here the big single mesh of the headline scene
 * is what matters, so parsing and BVH construction are parallel inside a mesh.  Thread count: CRLOADER_THREADS or the
 * number of online CPUs, capped at 64; every parallel algorithm in the loader gives the same bytes for any count.
 */
#include "cr_loader_int.h"
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

int crl_thread_count(void) {
	const char *e = getenv("CRLOADER_THREADS");
	long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
	if (n < 1) n = 1;
	if (n > 64) n = 64;
	return (int)n;
}

struct pfor { int n, next; void (*fn)(void *, int); void *arg; };

static void *pfor_worker(void *p) {
	struct pfor *f = p;
	for (;;) {
		int i = __atomic_fetch_add(&f->next, 1, __ATOMIC_RELAXED);
		if (i >= f->n) return NULL;
		f->fn(f->arg, i);
	}
}

void crl_parallel_for(int n, void (*fn)(void *, int), void *arg) {
	struct pfor f = { n, 0, fn, arg };
	int threads = crl_thread_count();
	if (threads > n) threads = n;
	pthread_t th[64];
	int started = 0;
	for (int i = 1; i < threads; ++i) if (pthread_create(&th[started], NULL, pfor_worker, &f) == 0) started++;
	pfor_worker(&f);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}
