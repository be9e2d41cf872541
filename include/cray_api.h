/*
 * cray_api.h — c-ray's public library API, served by the B200 path (c-ray_b200/libcrhost.so).
 *
 * A program written against the reference's `src/c-ray.h` (its own `src/main.c:14-42` is the canonical one) links
 * against libcrhost.so unchanged: the entry points below have the reference's names, argument meaning and return
 * conventions (implementation cited per function: reference src/c-ray.c, src/utils/args.c).  What differs is what is
 * behind them: scenes are built by this repository's loader (crloader.h), frames are rendered by GPU worker threads
 * through the C ABI of crgpu.h, and `-j N` counts GPUs.  Entry points of the reference that are unimplemented there
 * (ASSERT_NOT_REACHED in src/c-ray.c) and its network-worker mode are present for link compatibility and report so.
 */
#pragma once
#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* version / environment (c-ray.c:32-55) */
char *crGetVersion(void);
char *crGitHash(void);
bool  isDebug(void);
void  crInitialize(void);

/* command line (args.c:70-262): first existing file (or <arg>.json) = input; -j GPUs, -s samples, -d WxH, -t WxH;
 * every "-name" argument also becomes a tag queryable with crOptionIsSet("name") */
void  crParseArgs(int argc, char **argv);
bool  crOptionIsSet(char *key);
char *crPathArg(void);
void  crDestroyOptions(void);
char *crGetFilePath(char *fullPath);                    /* malloc'ed dirname + "/" (fileio.c:180-194) */

/* input (c-ray.c:100-106) — malloc'ed, NUL-terminated */
char *crReadFile(size_t *bytes);
char *crReadStdin(size_t *bytes);

/* renderer life cycle (c-ray.c:108-141, :247-262) */
void  crInitRenderer(void);
void  crDestroyRenderer(void);
int   crLoadSceneFromFile(char *filePath);              /* 0 ok, -1 error */
int   crLoadSceneFromBuf(char *buf);                    /* 0 ok, -1 error */
void  crStartRenderer(void);
void  crWriteImage(void);                               /* <outputFilePath><outputFileName>_<count %04d>.<png|bmp> */
void  crLog(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

/* prefs (c-ray.c:164-245) */
void     crSetThreadCount(int threadCount, bool fromSystem);
int      crGetThreadCount(void);
void     crSetSampleCount(int sampleCount);
int      crGetSampleCount(void);
void     crSetBounces(int bounces);
int      crGetBounces(void);
void     crSetTileWidth(unsigned width);
unsigned crGetTileWidth(void);
void     crSetTileHeight(unsigned height);
unsigned crGetTileHeight(void);
void     crSetImageWidth(unsigned width);
unsigned crGetImageWidth(void);
void     crSetImageHeight(unsigned height);
unsigned crGetImageHeight(void);
void     crSetOutputPath(char *filePath);
char    *crGetOutputPath(void);
void     crSetFileName(char *fileName);
char    *crGetFileName(void);
void     crSetAssetPath(void);
char    *crGetAssetPath(void);
void     crSetAntialiasing(bool on);
bool     crGetAntialiasing(void);

/* present in the reference header, not implemented by the reference either (c-ray.c:128-137,160-166,268-290) */
void crLoadMeshFromFile(char *filePath);
void crLoadMeshFromBuf(char *buf);
void crSetRenderOrder(void);
void crGetRenderOrder(void);
void crStartRenderWorker(void);
void crStartInteractive(void);
void crPauseInteractive(void);
void crGetCurrentImage(void);
void crRestartInteractive(void);

/* additions of this implementation */
const float *crGetRenderBuffer(unsigned *width, unsigned *height);   /* fp32 RGB, row 0 = image top; NULL before a render */
double       crGetRenderSeconds(void);
unsigned long long crGetRayCount(void);

#ifdef __cplusplus
}
#endif
