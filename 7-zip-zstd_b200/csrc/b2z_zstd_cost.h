/* b2z_zstd_cost.h -- the cost model of the price-based Zstandard parse (stage Z) as arithmetic: symbol codes, integer log2
 * costs, the repcode rules.  Plain C, host + device (B2Z_HD).  Used by csrc/zstd_enc_parse.cu and, read-only, by its sequential
 * statement oracle/zstd_opt_oracle.c so that both price a sequence the same way (prices steer the choice of sequences only; what
 * is emitted is coded by stage E and checked by the reference decoder).
 *
 * Format rules followed (reference, /root/reference/C/zstd/): LL/ML codes and extra bits zstd_internal.h:98-164,
 * zstd_compress_internal.h:584-616; offset code = highbit(offBase), offBase 1..3 = repcodes with the ll == 0 shift
 * zstd_compress_internal.h:817-835 (ZSTD_updateRep).  The role of the cost model is that of zstd_opt.c:295-356 (ZSTD_getMatchPrice,
 * ZSTD_litLengthPrice, ZSTD_rawLiteralsCost); its statistics, approximations and numbers are ours. */
#ifndef B2Z_ZSTD_COST_H
#define B2Z_ZSTD_COST_H
#include "b2z_params.h"

#define ZOP_MINMATCH 3u
#define ZOP_N_OF 32u
#define ZOP_N_ML 53u
#define ZOP_N_LL 36u

/* 16 * log2(1 + (k + 0.5) / 16), k = 0..15 */
#define ZOP_FRAC_LIST 1,2,3,5,6,7,8,9,10,11,12,13,14,15,15,16
/* matchLength - 3 -> ML code for values < 128 (RFC 8878 3.1.1.3.2.1.1); extra bits per ML / LL code */
#define ZOP_MLCODE_LIST \
    0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31, \
    32,32,33,33,34,34,35,35,36,36,36,36,37,37,37,37,38,38,38,38,38,38,38,38,39,39,39,39,39,39,39,39, \
    40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41, \
    42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42
#define ZOP_MLBITS_LIST 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16
#define ZOP_LLCODE_LIST \
    0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,16,17,17,18,18,19,19,20,20,20,20,21,21,21,21, \
    22,22,22,22,22,22,22,22,23,23,23,23,23,23,23,23,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24
#define ZOP_LLBITS_LIST 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16

typedef struct {                     /* tables a coder instance reads (constant; the kernel keeps them in shared memory) */
    uint8_t frac[16], mlCode[128], mlBits[56], llCode[64], llBits[36];
} zop_tables;
#define ZOP_TABLES_INIT { { ZOP_FRAC_LIST }, { ZOP_MLCODE_LIST }, { ZOP_MLBITS_LIST }, { ZOP_LLCODE_LIST }, { ZOP_LLBITS_LIST } }

B2Z_HD uint32_t zop_highbit(uint32_t v) {
#ifdef __CUDA_ARCH__
    return 31u - (uint32_t)__clz((int)v);
#else
    return 31u - (uint32_t)__builtin_clz(v);
#endif
}
/* 16 * log2(x), x >= 1, from the leading bit and the next four */
B2Z_HD uint32_t zop_log16(const zop_tables *t, uint32_t x) {
    const uint32_t hb = zop_highbit(x);
    const uint32_t k = hb >= 4u ? ((x >> (hb - 4u)) & 15u) : ((x << (4u - hb)) & 15u);
    return 16u * hb + (k == 0u && (x & (x - 1u)) == 0u ? 0u : t->frac[k]);
}
/* cost in 1/16 bit of a symbol seen freq times out of sum (freq >= 1, sum >= freq) */
B2Z_HD uint32_t zop_cost(const zop_tables *t, uint32_t freq, uint32_t sum) {
    const uint32_t a = zop_log16(t, sum), b = zop_log16(t, freq);
    return a > b ? a - b : 1u;
}
B2Z_HD uint32_t zop_ml_code(const zop_tables *t, uint32_t mlBase) { return mlBase > 127u ? zop_highbit(mlBase) + 36u : t->mlCode[mlBase]; }
B2Z_HD uint32_t zop_ll_code(const zop_tables *t, uint32_t ll) { return ll > 63u ? zop_highbit(ll) + 19u : t->llCode[ll]; }

/* adaptive statistics of one block's sequences (stage Z updates them with the sequences it commits) */
typedef struct { uint32_t of[ZOP_N_OF], ml[ZOP_N_ML], ll[ZOP_N_LL], ofSum, mlSum, llSum; } zop_stats;

/* coder state a path leaves: repcode history (0 = not known yet in this block) and the literals since the last match */
typedef struct { uint32_t rep[3], litLen; } zop_ctx;

/* offBase of a match at distance off given the history: 1..3 = repcode (with the ll == 0 shift), else off + 3 */
B2Z_HD uint32_t zop_off_base(const zop_ctx *x, uint32_t off) {
    if (x->litLen) { if (off == x->rep[0]) return 1u; if (off == x->rep[1]) return 2u; if (off == x->rep[2]) return 3u; }
    else { if (off == x->rep[1]) return 1u; if (off == x->rep[2]) return 2u; if (x->rep[0] > 1u && off == x->rep[0] - 1u) return 3u; }
    return off + 3u;
}
/* history after a match at distance off (ZSTD_updateRep) */
B2Z_HD void zop_after_match(zop_ctx *x, uint32_t off) {
    const uint32_t ob = zop_off_base(x, off);
    if (ob > 3u) { x->rep[2] = x->rep[1]; x->rep[1] = x->rep[0]; x->rep[0] = off; }
    else {
        const uint32_t idx = ob - 1u + (x->litLen == 0u);
        if (idx != 0u) { if (idx != 1u) x->rep[2] = x->rep[1]; x->rep[1] = x->rep[0]; x->rep[0] = off; }
    }
    x->litLen = 0;
}
/* price of one sequence: literal-run code + offset code + match-length code, with their extra bits (literal BYTES are priced apart) */
B2Z_HD uint32_t zop_seq_price(const zop_tables *t, const zop_stats *s, uint32_t litLen, uint32_t offBase, uint32_t matchLen) {
    const uint32_t oc = zop_highbit(offBase), mc = zop_ml_code(t, matchLen - ZOP_MINMATCH), lc = zop_ll_code(t, litLen);
    return 16u * (oc + t->mlBits[mc] + t->llBits[lc]) + zop_cost(t, s->of[oc], s->ofSum) + zop_cost(t, s->ml[mc], s->mlSum) + zop_cost(t, s->ll[lc], s->llSum);
}
B2Z_HD void zop_count_seq(const zop_tables *t, zop_stats *s, uint32_t litLen, uint32_t offBase, uint32_t matchLen) {
    s->of[zop_highbit(offBase)]++; s->ofSum++;
    s->ml[zop_ml_code(t, matchLen - ZOP_MINMATCH)]++; s->mlSum++;
    s->ll[zop_ll_code(t, litLen)]++; s->llSum++;
}

#endif
