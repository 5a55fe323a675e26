/* filter_oracle.c -- sequential statements of the filters csrc/b2z_filter.cu runs on the GPU.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Written from the filters' definitions (DOC/Methods.txt; C/Bra.h:46-101, C/Delta.h) in plain loops; tests/test_filters.py checks
 * every function here against the reference's own converters (C/Bra.c, C/Delta.c compiled into oracle/_ref/libref_xz.so). */
#include <string.h>
#include "oracle.h"

static uint32_t ld_le(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t ld_be(const uint8_t *p) { return (uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24); }
static void st_le(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void st_be(uint8_t *p, uint32_t v) { p[3] = (uint8_t)v; p[2] = (uint8_t)(v >> 8); p[1] = (uint8_t)(v >> 16); p[0] = (uint8_t)(v >> 24); }

/* delta: every byte minus the byte `dist` positions before it (zero before the start); decode adds them back in order */
static void delta(uint8_t *d, size_t n, uint32_t dist, int enc) {
    if (enc) { for (size_t i = n; i-- > dist;) d[i] = (uint8_t)(d[i] - d[i - dist]); }
    else for (size_t i = dist; i < n; i++) d[i] = (uint8_t)(d[i] + d[i - dist]);
}

/* x86 BCJ (C/Bra86.c:50-170, the whole buffer in one call, state 0): CALL / JMP rel32 (opcodes E8 / E9) whose operand's top byte is
 * 00 or FF become absolute (address of the next instruction added; subtracted when decoding), kept within 25 bits and sign-extended.
 * A 3-bit history of the opcode bytes just passed WITHOUT a conversion (bit 2 = one byte back, bit 1 = two, bit 0 = three) vetoes or
 * adjusts conversions that sit inside what may be another instruction's operand. */
static int is_00_ff(uint32_t b) { return (b & 0xFF) == 0x00 || (b & 0xFF) == 0xFF; }
static void x86(uint8_t *d, size_t n, uint32_t pc, int enc) {
    uint32_t hist = 0;
    size_t i = 0;
    while (i + 5 <= n) {
        if ((d[i] & 0xFE) != 0xE8) { hist >>= 1; i++; continue; }
        int convert = 0; uint32_t fix = 0;
        if (hist == 0) convert = is_00_ff(d[i + 4]);
        else if (hist == 1 || hist == 2 || hist == 4) {              /* exactly one opcode byte in the last three */
            fix = hist >> 1;                                         /* the operand byte that lines up with it: 0, 1 or 2 */
            convert = !is_00_ff(d[i + 1 + fix]) && is_00_ff(d[i + 4]);
        }
        if (!convert) { hist = (hist >> 1) | 4; i++; continue; }
        const uint32_t next = pc + (uint32_t)i + 5u;                 /* address of the instruction after this one */
        uint32_t v = ld_le(d + i + 1) + (1u << 24);                  /* biased: the 25-bit field is 0 .. 2^25 */
        v = enc ? v + next : v - next;
        if (hist != 0 && is_00_ff(v >> (8 * fix))) { v ^= (0x100u << (8 * fix)) - 1u; v = enc ? v + next : v - next; }
        st_le(d + i + 1, (v & 0x1FFFFFFu) - (1u << 24));
        hist = 0; i += 5;
    }
}

/* methodId: 7-Zip's (3 delta, 0xA ARM64, 0x3030205 PPC, 0x3030501 ARM, 0x3030805 SPARC); prop: delta distance 1..256 / start offset */
int b2zo_filter(uint32_t methodId, int enc, void *datav, size_t n, uint32_t prop) {
    uint8_t *d = (uint8_t *)datav;
    if (methodId == 3) { if (prop < 1 || prop > 256) return -1; delta(d, n, prop, enc); return 0; }
    if (methodId == 0x03030103u) { x86(d, n, prop, enc); return 0; }
    if (methodId == 0x03030701u) {                                  /* ARM Thumb: BL = halfwords F000..F7FF, F800..FFFF; 22-bit halfword offset from address + 4 */
        for (size_t i = 0; i + 4 <= (n & ~(size_t)1);) {
            const uint32_t a = d[i] | ((uint32_t)d[i + 1] << 8), b = d[i + 2] | ((uint32_t)d[i + 3] << 8);
            if ((a >> 11) != 0x1E || (b >> 11) != 0x1F) { i += 2; continue; }
            uint32_t off = ((a & 0x7FF) << 11) | (b & 0x7FF);
            const uint32_t t = ((prop + (uint32_t)i + 4u) >> 1) & 0x3FFFFF;
            off = (enc ? off + t : off + 0x400000u - t) & 0x3FFFFF;
            const uint32_t na = 0xF000u | (off >> 11), nb = 0xF800u | (off & 0x7FF);
            d[i] = (uint8_t)na; d[i + 1] = (uint8_t)(na >> 8); d[i + 2] = (uint8_t)nb; d[i + 3] = (uint8_t)(nb >> 8);
            i += 4;
        }
        return 0;
    }
    for (size_t i = 0; i + 4 <= n; i += 4) {
        const uint32_t ia = prop + (uint32_t)i;                    /* address of this instruction */
        if (methodId == 0xA) {
            uint32_t w = ld_le(d + i);
            if ((w & 0xFC000000u) == 0x94000000u) {                 /* BL: 26-bit word offset, modulo 2^26 */
                uint32_t off = w & 0x03FFFFFFu, t = (ia >> 2) & 0x03FFFFFFu;
                off = (enc ? off + t : off + 0x04000000u - t) & 0x03FFFFFFu;
                st_le(d + i, 0x94000000u | off);
            } else if ((w & 0x9F000000u) == 0x90000000u) {          /* ADRP: signed 21-bit page offset, converted when |offset| < 2^17 */
                int32_t off = (int32_t)((((w >> 5) & 0x7FFFFu) << 2) | ((w >> 29) & 3u));
                if (off & (1 << 20)) off -= (1 << 21);
                if (off >= -(1 << 17) && off < (1 << 17)) {
                    uint32_t b = (uint32_t)(off + (1 << 17)), page = (ia >> 12) & 0x3FFFFu;
                    b = (enc ? b + page : b + 0x40000u - page) & 0x3FFFFu;      /* 18-bit biased arithmetic */
                    const uint32_t o21 = (uint32_t)((int32_t)b - (1 << 17)) & 0x1FFFFFu;
                    st_le(d + i, (w & 0x9F00001Fu) | ((o21 & 3u) << 29) | ((o21 >> 2) << 5));
                }
            }
        } else if (methodId == 0x03030501u) {
            uint32_t w = ld_le(d + i);
            if ((w >> 24) == 0xEB) {                                 /* BL: 24-bit word offset from the instruction after next */
                uint32_t off = w & 0xFFFFFFu, t = ((ia + 8u) >> 2) & 0xFFFFFFu;
                off = (enc ? off + t : off + 0x1000000u - t) & 0xFFFFFFu;
                st_le(d + i, 0xEB000000u | off);
            }
        } else if (methodId == 0x03030205u) {
            uint32_t w = ld_be(d + i);
            if ((w >> 26) == 0x12 && (w & 3u) == 1u) {               /* bl: 26-bit byte offset (low bits AA = 0, LK = 1 stay) */
                uint32_t v = enc ? w + ia : w - ia;
                st_be(d + i, 0x48000000u | (v & 0x03FFFFFFu));
            }
        } else if (methodId == 0x03030805u) {
            uint32_t w = ld_be(d + i);
            const uint32_t hi = w >> 22;
            if (hi == 0x100u || hi == 0x1FFu) {                      /* call: displacement within 22 bits + sign */
                int32_t disp = (int32_t)(w & 0x3FFFFFu); if (hi == 0x1FFu) disp -= (1 << 22);
                uint32_t x = (uint32_t)(disp + (1 << 22)) << 2;      /* byte offset, biased: 0 .. 2^25 */
                x = (enc ? x + ia : x - ia) & 0x1FFFFFFu;
                st_be(d + i, ((uint32_t)(x - 0x1000000u) >> 2) | 0x40000000u);
            }
        } else return -1;
    }
    return 0;
}
