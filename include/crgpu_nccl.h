/*
 * crgpu_nccl.h — the one collective of this path: gather the fp32 framebuffer tiles of every GPU into
 * one GPU's framebuffer over NVLink (libcrgpu_nccl.so, links libnccl).
 *
 * Replaces the tile return of c-ray's cluster mode, where every worker base64-encodes its finished tile
 * into a JSON message for the master (reference src/utils/protocol/worker.c:196-214 → server.c:159-174).
 * Here all GPUs of one box render tiles of the same frame and the float tiles travel device-to-device:
 * one pack kernel per sender, grouped ncclSend/ncclRecv (NCCL has no gather primitive; one message per
 * peer), one unpack kernel per peer on the root.  Two ways to form the group:
 *   - in-process: one host thread per GPU (renderFrame -j N), one communicator per device (ncclCommInitAll);
 *   - one process per GPU (torchrun / mpirun style): rank 0 makes a 128-byte id (crgpu_comm_unique_id), the
 *     launcher hands it to every rank, each rank joins with crgpu_comm_create_rank (ncclCommInitRank).
 * A communicator is meant to outlive frames: create it once, gather every frame.
 */
#pragma once
#include "crgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crgpu_comm crgpu_comm;

#define CRGPU_COMM_ID_BYTES 128

/* ---- in-process group: devices[i] is the CUDA device of member i; W x H is the frame every member renders */
int crgpu_comm_create(const int *devices, int n, crgpu_comm **out);
/* rects = 4 ints per tile (x0, y0, x1, y1; y up, end exclusive); owner[i] = member that rendered tile i; scenes[i] =
 * member i's scene (same image size everywhere).  After the call scenes[root]'s framebuffer holds every tile. */
int crgpu_comm_gather_tiles(crgpu_comm *c, crgpu_scene **scenes, const int *rects, const int *owner, int ntiles, int root);

/* ---- one process per GPU */
int crgpu_comm_unique_id(void *id_out /* CRGPU_COMM_ID_BYTES */);
int crgpu_comm_create_rank(const void *id, int rank, int world, int device, crgpu_comm **out);
/* every rank calls this with the same rects/owner (owner[i] = rank that rendered tile i) and ITS OWN scene */
int crgpu_comm_gather_tiles_rank(crgpu_comm *c, crgpu_scene *mine, const int *rects, const int *owner, int ntiles, int root);

/* rank mode: run the gather on a caller-owned CUDA stream (so a host can order it behind its render kernels and time both with
 * its own events); use_own != 0 goes back to the communicator's private stream */
int crgpu_comm_set_stream(crgpu_comm *c, void *cuda_stream, int use_own);

int crgpu_comm_destroy(crgpu_comm *c);

#ifdef __cplusplus
}
#endif
