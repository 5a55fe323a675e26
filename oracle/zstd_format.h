/* zstd_format.h -- Zstandard wire-format constants (RFC 8878), shared by the oracle's
 * decoder and encoder restatements.  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is
 * linked into, imported by, or executed from the product path.
 *
 * Reference locations these restate (for parity checking, /root/reference/C/zstd/):
 *   LL/ML base + extra-bit tables, default norms ... zstd_internal.h:98-164
 *   code mapping (LLcode / MLcode) ............... zstd_compress_internal.h:584-616
 *   FSE spread step .............................. fse.h:623
 */
#ifndef B2ZO_ZSTD_FORMAT_H
#define B2ZO_ZSTD_FORMAT_H
#include <stdint.h>
#include <stddef.h>

#define ZF_MAGIC            0xFD2FB528u
#define ZF_MAGIC_SKIP_MASK  0xFFFFFFF0u
#define ZF_MAGIC_SKIP       0x184D2A50u
#define ZF_BLOCK_MAX        (128u << 10)
#define ZF_MAXLL            35
#define ZF_MAXML            52
#define ZF_MAXOFF           31
#define ZF_LL_FSELOG        9
#define ZF_ML_FSELOG        9
#define ZF_OF_FSELOG        8
#define ZF_LL_DEFLOG        6
#define ZF_ML_DEFLOG        6
#define ZF_OF_DEFLOG        5
#define ZF_HUF_MAXBITS      11
#define ZF_MINMATCH         3

static const uint32_t ZF_LL_base[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,
    16,18,20,22,24,28,32,40,48,64,0x80,0x100,0x200,0x400,0x800,0x1000,0x2000,0x4000,0x8000,0x10000 };
static const uint8_t ZF_LL_bits[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
    1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
static const uint32_t ZF_ML_base[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,
    19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,
    35,37,39,41,43,47,51,59,67,83,99,0x83,0x103,0x203,0x403,0x803,0x1003,0x2003,0x4003,0x8003,0x10003 };
static const uint8_t ZF_ML_bits[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
    1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 };
static const int16_t ZF_LL_defaultNorm[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,
    2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
static const int16_t ZF_ML_defaultNorm[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
static const int16_t ZF_OF_defaultNorm[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };

static inline uint32_t zf_highbit32(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }

static inline uint32_t zf_ll_code(uint32_t ll) {
    static const uint8_t t[64] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,16,17,17,18,18,19,19,
        20,20,20,20,21,21,21,21,22,22,22,22,22,22,22,22,23,23,23,23,23,23,23,23,
        24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24 };
    return ll > 63 ? zf_highbit32(ll) + 19 : t[ll];
}
/* mlBase = matchLength - 3 */
static inline uint32_t zf_ml_code(uint32_t mlBase) {
    static const uint8_t t[128] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,
        16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,
        32,32,33,33,34,34,35,35,36,36,36,36,37,37,37,37,
        38,38,38,38,38,38,38,38,39,39,39,39,39,39,39,39,
        40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,
        41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,
        42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,
        42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42 };
    return mlBase > 127 ? zf_highbit32(mlBase) + 36 : t[mlBase];
}

#endif
