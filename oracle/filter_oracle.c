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

/* methodId: 7-Zip's (3 delta, 0xA ARM64, 0x3030205 PPC, 0x3030501 ARM, 0x3030805 SPARC); prop: delta distance 1..256 / start offset */
int b2zo_filter(uint32_t methodId, int enc, void *datav, size_t n, uint32_t prop) {
    uint8_t *d = (uint8_t *)datav;
    if (methodId == 3) { if (prop < 1 || prop > 256) return -1; delta(d, n, prop, enc); return 0; }
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
