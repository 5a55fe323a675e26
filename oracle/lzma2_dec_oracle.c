/* lzma2_dec_oracle.c -- plain-C restatement of the LZMA2 decoder (7-Zip method 21: both the stock
 * LZMA2 encoder and Fast-LZMA2 emit this stream; one decoder).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Checker for the CUDA LZMA2 decoder.
 *
 * Parity pin (tests/test_oracle_lzma2.py): streams produced by the reference's Fast-LZMA2 encoder
 * (oracle/_ref/libref_lzma.so: FL2_compress, C/fast-lzma2/fl2_compress.c) and by liblzma's raw LZMA2
 * encoder (python `lzma`, an independent implementation of the same format), committed as golden fixtures.
 *
 * Reference functions restated (under /root/reference/C/):
 *   chunk control byte / sizes / props ... Lzma2Dec.c:16-36,97-165 (Lzma2Dec_UpdateState)
 *   chunk loop, dict/state/props resets ... Lzma2Dec.c:178-330 (Lzma2Dec_DecodeToDic)
 *   range decoder, bit models ............ LzmaDec.c:20-120 (NORMALIZE, GET_BIT2, TREE_DECODE)
 *   literal / match / rep decoding ....... LzmaDec.c:229-600 (LZMA_DECODE_REAL)
 *   length coder, distance slots ......... LzmaDec.c:130-227 (probability layout), :430-560
 *   match copy in the dictionary ......... LzmaDec.c:560-620
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define kNumStates 12
#define kNumPosStatesMax 16
#define kLenLow 8
#define kLenHigh 256
#define kNumLenToPosStates 4
#define kEndPosModelIndex 14
#define kNumFullDistances 128
#define kAlignSize 16
#define kTop (1u << 24)
#define PROB_INIT 1024

typedef uint16_t prob_t;

typedef struct { prob_t choice, choice2, low[kNumPosStatesMax][kLenLow], mid[kNumPosStatesMax][kLenLow], high[kLenHigh]; } len_coder;

typedef struct {
    prob_t isMatch[kNumStates][kNumPosStatesMax], isRep[kNumStates], isRepG0[kNumStates], isRepG1[kNumStates], isRepG2[kNumStates];
    prob_t isRep0Long[kNumStates][kNumPosStatesMax];
    prob_t posSlot[kNumLenToPosStates][64], specPos[kNumFullDistances - kEndPosModelIndex + 1], align[kAlignSize];
    len_coder len, repLen;
    prob_t *lit;                                   /* 0x300 << (lc+lp) */
    uint32_t lc, lp, pb, state, rep[4];
    /* range coder */
    const uint8_t *in, *inEnd; uint32_t range, code; int rcErr;
} lzma_t;

static void reset_probs(lzma_t *z) {
    prob_t *p = (prob_t *)z->isMatch; size_t n = ((uint8_t *)&z->lit - (uint8_t *)z->isMatch) / sizeof(prob_t);
    for (size_t i = 0; i < n; i++) p[i] = PROB_INIT;
    for (size_t i = 0; i < ((size_t)0x300 << (z->lc + z->lp)); i++) z->lit[i] = PROB_INIT;
    z->state = 0; z->rep[0] = z->rep[1] = z->rep[2] = z->rep[3] = 0;
}

static inline uint32_t rc_byte(lzma_t *z) { if (z->in < z->inEnd) return *z->in++; z->rcErr = 1; return 0; }
static inline uint32_t rc_bit(lzma_t *z, prob_t *p) {
    if (z->range < kTop) { z->range <<= 8; z->code = (z->code << 8) | rc_byte(z); }
    uint32_t bound = (z->range >> 11) * *p;
    if (z->code < bound) { z->range = bound; *p = (prob_t)(*p + ((2048 - *p) >> 5)); return 0; }
    z->range -= bound; z->code -= bound; *p = (prob_t)(*p - (*p >> 5)); return 1;
}
static uint32_t rc_direct(lzma_t *z, uint32_t n) {
    uint32_t r = 0;
    while (n--) {
        if (z->range < kTop) { z->range <<= 8; z->code = (z->code << 8) | rc_byte(z); }
        z->range >>= 1; z->code -= z->range;
        uint32_t t = 0u - (z->code >> 31); z->code += z->range & t;
        r = (r << 1) + (t + 1);
    }
    return r;
}
static uint32_t tree(lzma_t *z, prob_t *p, uint32_t bits) { uint32_t m = 1; for (uint32_t i = 0; i < bits; i++) m = (m << 1) | rc_bit(z, p + m); return m - (1u << bits); }
static uint32_t tree_rev(lzma_t *z, prob_t *p, uint32_t bits) { uint32_t m = 1, r = 0; for (uint32_t i = 0; i < bits; i++) { uint32_t b = rc_bit(z, p + m); m = (m << 1) | b; r |= b << i; } return r; }
static uint32_t len_decode(lzma_t *z, len_coder *l, uint32_t ps) {
    if (!rc_bit(z, &l->choice)) return 2 + tree(z, l->low[ps], 3);
    if (!rc_bit(z, &l->choice2)) return 10 + tree(z, l->mid[ps], 3);
    return 18 + tree(z, l->high, 8);
}

/* Decode one LZMA chunk: exactly `unpack` bytes into dic[*pos ...], history starts at dic[dicStart]. 0 ok, -1 corrupt. */
static int lzma_chunk(lzma_t *z, uint8_t *dic, size_t dicStart, size_t *posIO, size_t unpack, uint32_t dictSize) {
    size_t pos = *posIO, end = pos + unpack;
    if (z->inEnd - z->in < 5 || z->in[0] != 0) return -1;
    z->code = ((uint32_t)z->in[1] << 24) | ((uint32_t)z->in[2] << 16) | ((uint32_t)z->in[3] << 8) | z->in[4]; z->in += 5; z->range = 0xFFFFFFFFu; z->rcErr = 0;
    const uint32_t pbMask = (1u << z->pb) - 1, lpMask = (1u << z->lp) - 1;
    while (pos < end) {
        const uint32_t ps = (uint32_t)(pos - dicStart) & pbMask;
        if (!rc_bit(z, &z->isMatch[z->state][ps])) {
            const uint32_t prev = pos > dicStart ? dic[pos - 1] : 0;
            prob_t *p = z->lit + (size_t)0x300 * ((((uint32_t)(pos - dicStart) & lpMask) << z->lc) + (prev >> (8 - z->lc)));
            uint32_t sym = 1;
            if (z->state >= 7) {
                uint32_t mb = dic[pos - z->rep[0] - 1];
                do { uint32_t mbit = (mb >> 7) & 1; mb <<= 1; uint32_t b = rc_bit(z, p + ((1 + mbit) << 8) + sym); sym = (sym << 1) | b; if (mbit != b) break; } while (sym < 0x100);
            }
            while (sym < 0x100) sym = (sym << 1) | rc_bit(z, p + sym);
            dic[pos++] = (uint8_t)sym;
            z->state = z->state < 4 ? 0 : (z->state < 10 ? z->state - 3 : z->state - 6);
            continue;
        }
        uint32_t len;
        if (!rc_bit(z, &z->isRep[z->state])) {
            z->rep[3] = z->rep[2]; z->rep[2] = z->rep[1]; z->rep[1] = z->rep[0];
            len = len_decode(z, &z->len, ps);
            z->state = z->state < 7 ? 7 : 10;
            uint32_t slot = tree(z, z->posSlot[len - 2 < kNumLenToPosStates ? len - 2 : kNumLenToPosStates - 1], 6), dist;
            if (slot < 4) dist = slot;
            else {
                uint32_t nb = (slot >> 1) - 1; dist = (2 | (slot & 1)) << nb;
                if (slot < kEndPosModelIndex) dist += tree_rev(z, z->specPos + dist - slot - 1, nb);
                else { dist += rc_direct(z, nb - 4) << 4; dist += tree_rev(z, z->align, 4); }
            }
            z->rep[0] = dist;
            if (dist == 0xFFFFFFFFu) return -1;                   /* end marker: not allowed inside LZMA2 */
        } else {
            if (pos == dicStart) return -1;
            if (!rc_bit(z, &z->isRepG0[z->state])) {
                if (!rc_bit(z, &z->isRep0Long[z->state][ps])) {
                    z->state = z->state < 7 ? 9 : 11;
                    if (z->rep[0] >= pos - dicStart || z->rep[0] >= dictSize) return -1;
                    dic[pos] = dic[pos - z->rep[0] - 1]; pos++;
                    continue;
                }
            } else {
                uint32_t d;
                if (!rc_bit(z, &z->isRepG1[z->state])) d = z->rep[1];
                else { if (!rc_bit(z, &z->isRepG2[z->state])) d = z->rep[2]; else { d = z->rep[3]; z->rep[3] = z->rep[2]; } z->rep[2] = z->rep[1]; }
                z->rep[1] = z->rep[0]; z->rep[0] = d;
            }
            len = len_decode(z, &z->repLen, ps);
            z->state = z->state < 7 ? 8 : 11;
        }
        if (z->rep[0] >= pos - dicStart || z->rep[0] >= dictSize) return -1;
        if (len > end - pos) return -1;                           /* a match may not cross the chunk end (LzmaDec.c:1030-1033) */
        for (size_t i = 0; i < len; i++) { dic[pos] = dic[pos - z->rep[0] - 1]; pos++; }
        if (z->rcErr) return -1;
    }
    if (z->range < kTop) { z->range <<= 8; z->code = (z->code << 8) | rc_byte(z); }      /* final normalisation */
    if (z->rcErr || z->code != 0) return -1;                       /* LzmaDec.c:1020: a finished chunk leaves code == 0 */
    *posIO = pos;
    return 0;
}

/* Raw LZMA2 stream (chunks ... 0x00) -> dst. dictProp: the 1-byte coder property (Lzma2Decoder.cpp:40-48).
 * Returns decoded size, -1 corrupt, -2 dst too small.  *srcUsed (optional): bytes consumed incl. the end marker. */
int64_t b2zo_lzma2_decompress(void *dstv, size_t dstCap, const void *srcv, size_t srcSize, uint32_t dictProp, size_t *srcUsed) {
    const uint8_t *ip = (const uint8_t *)srcv, *iend = ip + srcSize;
    uint8_t *dst = (uint8_t *)dstv;
    if (dictProp > 40) return -1;
    const uint32_t dictSize = dictProp == 40 ? 0xFFFFFFFFu : ((2u | (dictProp & 1)) << (dictProp / 2 + 11));
    lzma_t *z = (lzma_t *)calloc(1, sizeof(lzma_t));
    z->lit = (prob_t *)malloc(sizeof(prob_t) * ((size_t)0x300 << 4));
    size_t pos = 0, dicStart = 0; int64_t rc = -1;
    uint32_t needInit = 0xE0;                       /* Lzma2Dec.c:108-121: lowest LZMA control byte acceptable next */
    for (;;) {
        if (ip >= iend) goto done;
        const uint32_t ctl = *ip++;
        if (ctl == 0) { rc = (int64_t)pos; break; }
        if (ctl == 1 || ctl == 2) {
            if (iend - ip < 2) goto done;
            const size_t n = (((size_t)ip[0] << 8) | ip[1]) + 1; ip += 2;
            if (ctl == 1) { dicStart = pos; needInit = 0xC0; } else if (needInit == 0xE0) goto done;
            if ((size_t)(iend - ip) < n) goto done;
            if (dstCap - pos < n) { rc = -2; goto done; }
            memcpy(dst + pos, ip, n); pos += n; ip += n;
            continue;
        }
        if (ctl < 0x80 || ctl < needInit) goto done;
        needInit = 0;
        if (iend - ip < 4) goto done;
        const size_t unpack = ((((size_t)ctl & 0x1F) << 16) | ((size_t)ip[0] << 8) | ip[1]) + 1;
        const size_t pack = (((size_t)ip[2] << 8) | ip[3]) + 1; ip += 4;
        const uint32_t mode = (ctl >> 5) & 3;
        if (mode == 3) dicStart = pos;
        if (mode >= 2) {
            if (ip >= iend) goto done;
            uint32_t d = *ip++; if (d >= 9 * 5 * 5) goto done;
            z->lc = d % 9; d /= 9; z->pb = d / 5; z->lp = d % 5;
            if (z->lc + z->lp > 4) goto done;
        }
        if (mode >= 1) reset_probs(z);
        if ((size_t)(iend - ip) < pack) goto done;
        if (dstCap - pos < unpack) { rc = -2; goto done; }
        z->in = ip; z->inEnd = ip + pack;
        if (lzma_chunk(z, dst, dicStart, &pos, unpack, dictSize)) goto done;
        if (z->in != z->inEnd) goto done;                          /* the chunk must consume its packed size exactly */
        ip += pack;
    }
done:
    if (srcUsed) *srcUsed = (size_t)(ip - (const uint8_t *)srcv);
    free(z->lit); free(z);
    return rc;
}
