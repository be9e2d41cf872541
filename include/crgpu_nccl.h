/*
 * crgpu_nccl.h — the one collective of this path: gather the fp32 framebuffer tiles of every GPU into
 * one GPU's framebuffer over NVLink (libcrgpu_nccl.so, links libnccl).
 *
 * Replaces the tile return of c-ray's cluster mode, where every worker base64-encodes its finished tile
 * into a JSON message for the master (reference src/utils/protocol/worker.c:196-214 → server.c:159-174).
 * Here all GPUs of one box render tiles of the same frame from the shared queue (cr_host.c) and the
 * float tiles travel device-to-device: grouped ncclSend/ncclRecv (NCCL has no gather primitive), one
 * message per peer.  Single process, one communicator per device (ncclCommInitAll).
 */
#pragma once
#include "crgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crgpu_comm crgpu_comm;

/* scenes[i] must live on distinct devices and share image dimensions. */
int crgpu_comm_create(crgpu_scene **scenes, int n, crgpu_comm **out);
/* rects = 4 ints per tile (x0, y0, x1, y1; y up, end exclusive); owner[i] = index into scenes[] of the GPU that
 * rendered tile i.  After the call scenes[root]'s framebuffer holds every tile. */
int crgpu_comm_gather_tiles(crgpu_comm *c, const int *rects, const int *owner, int ntiles, int root);
int crgpu_comm_destroy(crgpu_comm *c);

#ifdef __cplusplus
}
#endif
