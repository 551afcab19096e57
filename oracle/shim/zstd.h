/* Declaration-only stand-in for <zstd.h> (the image has no zstd headers).
 * Only the opaque types that readsb's headers mention are declared so that the
 * reference's hot-path translation units parse; nothing here is ever called by
 * the demodulator path and no zstd symbol is linked. */
#ifndef B200_ORACLE_ZSTD_SHIM_H
#define B200_ORACLE_ZSTD_SHIM_H
#include <stddef.h>
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef ZSTD_CCtx ZSTD_CStream;
typedef struct ZSTD_inBuffer_s { const void *src; size_t size; size_t pos; } ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s { void *dst; size_t size; size_t pos; } ZSTD_outBuffer;
#endif
