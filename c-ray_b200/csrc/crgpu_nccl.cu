/*
 * crgpu_nccl.cu — framebuffer tile gather over NCCL (include/crgpu_nccl.h).  Host-side plumbing only:
 * tiles are packed with strided device-to-device copies, exchanged with grouped ncclSend/ncclRecv and
 * unpacked on the root; no arithmetic touches the pixels.
 */
#include "../../include/crgpu_nccl.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <vector>
#include <cstdio>

struct crgpu_comm {
	int n;
	std::vector<crgpu_scene *> scenes;
	std::vector<int> devices;
	std::vector<ncclComm_t> comms;
	std::vector<cudaStream_t> streams;
	std::vector<float *> stage;       /* per device: packed tiles (send side) */
	std::vector<float *> recv;        /* on the root device: one staging buffer per peer */
	std::vector<size_t> cap;
	int W, H;
};

#define NCHECK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { fprintf(stderr, "crgpu_nccl: %s: %s\n", #x, ncclGetErrorString(r_)); return CRGPU_ERR_CUDA; } } while (0)
#define CCHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "crgpu_nccl: %s: %s\n", #x, cudaGetErrorString(e_)); return CRGPU_ERR_CUDA; } } while (0)

extern "C" int crgpu_comm_create(crgpu_scene **scenes, int n, crgpu_comm **out) {
	if (!scenes || !out || n < 1) return CRGPU_ERR_BAD_ARGUMENT;
	crgpu_comm *c = new crgpu_comm();
	c->n = n;
	c->scenes.assign(scenes, scenes + n);
	c->devices.resize(n); c->comms.resize(n); c->streams.resize(n);
	c->stage.assign(n, nullptr); c->recv.assign(n, nullptr); c->cap.assign(n, 0);
	for (int i = 0; i < n; ++i) {
		int w = 0, h = 0;
		if (crgpu_scene_info(scenes[i], &c->devices[i], &w, &h)) { delete c; return CRGPU_ERR_BAD_ARGUMENT; }
		if (i == 0) { c->W = w; c->H = h; }
		else if (w != c->W || h != c->H) { delete c; return CRGPU_ERR_BAD_ARGUMENT; }
	}
	NCHECK(ncclCommInitAll(c->comms.data(), n, c->devices.data()));
	for (int i = 0; i < n; ++i) { CCHECK(cudaSetDevice(c->devices[i])); CCHECK(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking)); }
	*out = c;
	return CRGPU_OK;
}

static inline size_t tile_floats(const int *r) { return (size_t)(r[2] - r[0]) * (size_t)(r[3] - r[1]) * 3u; }

extern "C" int crgpu_comm_gather_tiles(crgpu_comm *c, const int *rects, const int *owner, int ntiles, int root) {
	if (!c || !rects || !owner || root < 0 || root >= c->n) return CRGPU_ERR_BAD_ARGUMENT;
	const size_t pitch = (size_t)c->W * 3u * sizeof(float);
	std::vector<size_t> total(c->n, 0);
	for (int t = 0; t < ntiles; ++t) {
		if (owner[t] < 0 || owner[t] >= c->n) return CRGPU_ERR_BAD_ARGUMENT;
		total[owner[t]] += tile_floats(rects + 4 * t);
	}
	std::vector<float *> fb(c->n);
	for (int i = 0; i < c->n; ++i) { void *p = nullptr; if (crgpu_framebuffer_device_ptr(c->scenes[i], &p, nullptr)) return CRGPU_ERR_BAD_ARGUMENT; fb[i] = (float *)p; }
	/* (re)allocate staging */
	for (int i = 0; i < c->n; ++i) {
		if (i == root || total[i] <= c->cap[i]) continue;
		CCHECK(cudaSetDevice(c->devices[i])); if (c->stage[i]) cudaFree(c->stage[i]);
		CCHECK(cudaMalloc((void **)&c->stage[i], total[i] * sizeof(float)));
		CCHECK(cudaSetDevice(c->devices[root])); if (c->recv[i]) cudaFree(c->recv[i]);
		CCHECK(cudaMalloc((void **)&c->recv[i], total[i] * sizeof(float)));
		c->cap[i] = total[i];
	}
	/* pack on every sender */
	std::vector<size_t> off(c->n, 0);
	for (int t = 0; t < ntiles; ++t) {
		const int o = owner[t];
		if (o == root) continue;
		const int *r = rects + 4 * t;
		const size_t w = (size_t)(r[2] - r[0]) * 3u * sizeof(float), rows = (size_t)(r[3] - r[1]);
		const float *src = fb[o] + ((size_t)(c->H - r[3]) * c->W + (size_t)r[0]) * 3u;        /* rows y0..y1-1 are storage rows H-y1..H-1-y0 */
		CCHECK(cudaSetDevice(c->devices[o]));
		CCHECK(cudaMemcpy2DAsync(c->stage[o] + off[o], w, src, pitch, w, rows, cudaMemcpyDeviceToDevice, c->streams[o]));
		off[o] += tile_floats(r);
	}
	/* one grouped exchange: every peer sends its packed tiles to the root */
	NCHECK(ncclGroupStart());
	for (int i = 0; i < c->n; ++i) {
		if (i == root || total[i] == 0) continue;
		NCHECK(ncclSend(c->stage[i], total[i], ncclFloat, root, c->comms[i], c->streams[i]));
		NCHECK(ncclRecv(c->recv[i], total[i], ncclFloat, i, c->comms[root], c->streams[root]));
	}
	NCHECK(ncclGroupEnd());
	/* unpack on the root */
	std::fill(off.begin(), off.end(), 0);
	CCHECK(cudaSetDevice(c->devices[root]));
	for (int t = 0; t < ntiles; ++t) {
		const int o = owner[t];
		if (o == root) continue;
		const int *r = rects + 4 * t;
		const size_t w = (size_t)(r[2] - r[0]) * 3u * sizeof(float), rows = (size_t)(r[3] - r[1]);
		float *dst = fb[root] + ((size_t)(c->H - r[3]) * c->W + (size_t)r[0]) * 3u;
		CCHECK(cudaMemcpy2DAsync(dst, pitch, c->recv[o] + off[o], w, w, rows, cudaMemcpyDeviceToDevice, c->streams[root]));
		off[o] += tile_floats(r);
	}
	for (int i = 0; i < c->n; ++i) { CCHECK(cudaSetDevice(c->devices[i])); CCHECK(cudaStreamSynchronize(c->streams[i])); }
	return CRGPU_OK;
}

extern "C" int crgpu_comm_destroy(crgpu_comm *c) {
	if (!c) return CRGPU_OK;
	for (int i = 0; i < c->n; ++i) {
		cudaSetDevice(c->devices[i]);
		if (c->stage[i]) cudaFree(c->stage[i]);
		if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
		ncclCommDestroy(c->comms[i]);
	}
	if (c->n) { cudaSetDevice(c->devices[0]); }
	for (int i = 0; i < c->n; ++i) if (c->recv[i]) cudaFree(c->recv[i]);
	delete c;
	return CRGPU_OK;
}
