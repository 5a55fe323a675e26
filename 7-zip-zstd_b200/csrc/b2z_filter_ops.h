/* b2z_filter_ops.h -- the per-instruction rules of the stateless branch converters as arithmetic on one 32-bit word.
 * Plain C, host + device (B2Z_HD): used by csrc/b2z_filter.cu; the oracle (oracle/filter_oracle.c) states the same rules in its own
 * words and both are checked against the reference's converters (C/Bra.c via oracle/_ref/libref_xz.so) word by word.
 *
 * What the converters do (C/Bra.h:46-50): in CALL-type instructions the relative target becomes absolute (encode) or back
 * (decode), which makes repeated calls to one function byte-identical and so compressible.  `ia` = address of the instruction
 * (start offset property + position in the stream).  ARM64: C/Bra.c:75-123, ARM: :126-152, PPC: :155-191, SPARC: :198-252. */
#ifndef B2Z_FILTER_OPS_H
#define B2Z_FILTER_OPS_H
#include "b2z_params.h"

/* 7-Zip method ids of the filters (CPP/7zip/Compress/BranchRegister.cpp, DeltaFilter.cpp; DOC/Methods.txt) */
#define B200Z_F_DELTA 0x03u
#define B200Z_F_ARM64 0x0Au
#define B200Z_F_PPC   0x03030205u
#define B200Z_F_ARM   0x03030501u
#define B200Z_F_SPARC 0x03030805u

B2Z_HD uint32_t b2z_bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }

/* w = the instruction as the CPU reads it (little endian for ARM / ARM64, big endian for PPC / SPARC) */
B2Z_HD uint32_t b2z_conv_arm64(uint32_t w, uint32_t ia, int enc) {
    if ((w >> 26) == 0x25u) {                                       /* BL imm26: word offset */
        const uint32_t t = ia >> 2, imm = enc ? w + t : w - t;
        return 0x94000000u | (imm & 0x03FFFFFFu);
    }
    if ((w & 0x9F000000u) == 0x90000000u) {                         /* ADRP: 21-bit page offset immhi:immlo */
        const uint32_t imm = ((w >> 3) & 0x1FFFFCu) | ((w >> 29) & 3u);
        const uint32_t biased = (imm + (1u << 17)) & 0x1FFFFFu;     /* only offsets in [-2^17, 2^17) pages are converted */
        if (biased < (1u << 18)) {
            const uint32_t page = ia >> 12;
            const uint32_t b2 = (enc ? biased + page : biased - page) & 0x3FFFFu;
            const uint32_t out = (b2 - (1u << 17)) & 0x1FFFFFu;     /* back to a sign-extended 21-bit offset */
            return (w & 0x9F00001Fu) | ((out & 3u) << 29) | ((out >> 2) << 5);
        }
    }
    return w;
}
B2Z_HD uint32_t b2z_conv_arm(uint32_t w, uint32_t ia, int enc) {
    if ((w >> 24) != 0xEBu) return w;                               /* BL imm24, relative to the instruction after next */
    const uint32_t t = (ia + 8u) >> 2, imm = enc ? w + t : w - t;
    return 0xEB000000u | (imm & 0x00FFFFFFu);
}
B2Z_HD uint32_t b2z_conv_ppc(uint32_t w, uint32_t ia, int enc) {
    if ((w & 0xFC000003u) != 0x48000001u) return w;                 /* bl: AA = 0, LK = 1 */
    const uint32_t v = enc ? w + ia : w - ia;
    return 0x48000000u | (v & 0x03FFFFFFu);
}
B2Z_HD uint32_t b2z_conv_sparc(uint32_t w, uint32_t ia, int enc) {
    const uint32_t top = w >> 22;
    if (top != 0x100u && top != 0x1FFu) return w;                   /* call with a displacement in [-2^22, 2^22) words */
    const uint32_t biased = ((w & 0x3FFFFFu) + (top == 0x100u ? (1u << 22) : 0u)) << 2;      /* (disp + 2^22) * 4, < 2^25 */
    const uint32_t x = (enc ? biased + ia : biased - ia) & ((1u << 25) - 1u);
    return ((x - (1u << 24)) >> 2) | (1u << 30);
}

#endif
