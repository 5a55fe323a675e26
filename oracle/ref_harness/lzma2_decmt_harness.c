/* lzma2_decmt_harness.c -- memory-stream driver for the reference's multi-threaded LZMA2 decoder, linked into
 * oracle/_ref/libref_lzma.so (TEST / BENCH INFRASTRUCTURE: it only calls the unmodified reference).
 *
 * Mirrors what NCompress::NLzma2::CDecoder::Code does (CPP/7zip/Compress/Lzma2Decoder.cpp:95-186): Lzma2DecMt_Create,
 * props with numThreads and the input/output block limits of Lzma2Decoder.cpp:110-128, Lzma2DecMt_Decode over
 * ISeqInStream / ISeqOutStream wrappers -- here the two streams are plain memory buffers. */
#include <string.h>
#include "Alloc.h"
#include "Lzma2DecMt.h"

typedef struct { ISeqInStream vt; const Byte *p; size_t left; } MemIn;
typedef struct { ISeqOutStream vt; Byte *p; size_t left; size_t total; } MemOut;

static SRes mem_read(ISeqInStreamPtr pp, void *buf, size_t *size) {
    MemIn *s = Z7_CONTAINER_FROM_VTBL(pp, MemIn, vt);
    size_t n = *size < s->left ? *size : s->left;
    memcpy(buf, s->p, n); s->p += n; s->left -= n; *size = n;
    return SZ_OK;
}
static size_t mem_write(ISeqOutStreamPtr pp, const void *buf, size_t size) {
    MemOut *s = Z7_CONTAINER_FROM_VTBL(pp, MemOut, vt);
    if (size > s->left) return 0;
    memcpy(s->p, buf, size); s->p += size; s->left -= size; s->total += size;
    return size;
}

/* returns the reference's SRes; *dstSize = bytes written, *isMT = 1 if the MT path ran (independent blocks were found) */
int refh_lzma2_decode_mt(void *dst, size_t dstCap, size_t *dstSize, const void *src, size_t srcSize, unsigned prop,
                         unsigned threads, int *isMT) {
    CLzma2DecMtHandle h = Lzma2DecMt_Create(&g_AlignedAlloc, &g_MidAlloc);
    if (!h) return SZ_ERROR_MEM;
    CLzma2DecMtProps props;
    Lzma2DecMtProps_Init(&props);
    props.inBufSize_ST = 1 << 20;                    /* Lzma2Decoder.cpp:101-103 */
    props.outStep_ST = 1 << 20;
    props.numThreads = threads;
    {                                                 /* Lzma2Decoder.cpp:61-72,110-128: block limits from the dictionary size */
        const UInt32 dictSize = prop >= 40 ? 0xFFFFFFFFu : ((UInt32)(2 | (prop & 1)) << (prop / 2 + 11));
        UInt64 blockSize = (UInt64)dictSize << 2;
        if (blockSize < ((UInt64)1 << 20)) blockSize = (UInt64)1 << 20;
        if (blockSize > ((UInt64)1 << 28)) blockSize = (UInt64)1 << 28;
        if (blockSize < dictSize) blockSize = dictSize;
        blockSize = (blockSize + ((1 << 20) - 1)) & ~(UInt64)((1 << 20) - 1);
        props.outBlockMax = (size_t)blockSize;
        props.inBlockMax = (size_t)(blockSize + blockSize / 16);
    }
    MemIn in; in.vt.Read = mem_read; in.p = (const Byte *)src; in.left = srcSize;
    MemOut out; out.vt.Write = mem_write; out.p = (Byte *)dst; out.left = dstCap; out.total = 0;
    UInt64 inProcessed = 0; int mt = 0;
    SRes res = Lzma2DecMt_Decode(h, (Byte)prop, &props, &out.vt, NULL, 1, &in.vt, &inProcessed, &mt, NULL);
    Lzma2DecMt_Destroy(h);
    if (dstSize) *dstSize = out.total;
    if (isMT) *isMT = mt;
    return (int)res;
}
