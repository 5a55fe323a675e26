/* crc_oracle.c -- bit-at-a-time statements of the two digests csrc/b2z_crc.cu computes.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * CRC-32 (reflected 0xEDB88320, init and final xor 0xFFFFFFFF): what C/7zCrc.c:298 CrcCalc returns; check value of "123456789": 0xCBF43926.
 * CRC-64/XZ (reflected 0xC96C5795D7870F42, init and final xor all ones): C/XzCrc64.c; check value 0x995DC9BBDF1939FA. */
#include "oracle.h"

uint32_t b2zo_crc32(const void *data, size_t n) {
    const uint8_t *p = (const uint8_t *)data; uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) { c ^= p[i]; for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u))); }
    return ~c;
}
uint64_t b2zo_crc64(const void *data, size_t n) {
    const uint8_t *p = (const uint8_t *)data; uint64_t c = ~0ull;
    for (size_t i = 0; i < n; i++) { c ^= p[i]; for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xC96C5795D7870F42ull & (0ull - (c & 1ull))); }
    return ~c;
}
