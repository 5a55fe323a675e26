// zstd_dec_api.cu -- decoder half of the C ABI (placeholder until the decode kernels land).
#include <cuda_runtime.h>
#include <cstring>
#include "../../include/b200z.h"

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

extern "C" {

// Walk frame headers (and block headers) on the host: cheap, sequential, no entropy decoding.
// Follows ZSTD_getFrameHeader / ZSTD_findFrameCompressedSize (C/zstd/zstd_decompress.c:482-600,702).
int b200z_zstd_frame_info(const void* srcv, size_t srcSize, uint64_t* contentSize, uint32_t* nFrames) {
    const uint8_t* ip = (const uint8_t*)srcv; const uint8_t* iend = ip + srcSize;
    uint64_t total = 0; uint32_t frames = 0; int unknown = 0;
    while (ip < iend) {
        if (iend - ip < 4) return B200Z_E_CORRUPT;
        const uint32_t magic = rd32(ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (iend - ip < 8) return B200Z_E_CORRUPT;
            const uint32_t sz = rd32(ip + 4);
            if ((size_t)(iend - ip) < 8 + (size_t)sz) return B200Z_E_CORRUPT;
            ip += 8 + sz; continue;
        }
        if (magic != 0xFD2FB528u) return B200Z_E_CORRUPT;
        if (iend - ip < 6) return B200Z_E_CORRUPT;
        const uint32_t fhd = ip[4]; ip += 5;
        const uint32_t fcsFlag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, didFlag = fhd & 3;
        if (fhd & 8) return B200Z_E_CORRUPT;
        if (!single) ip += 1;
        static const int didBytes[4] = { 0, 1, 2, 4 };
        ip += didBytes[didFlag];
        const int fcsBytes = fcsFlag == 0 ? (int)single : (fcsFlag == 1 ? 2 : (fcsFlag == 2 ? 4 : 8));
        if (iend - ip < fcsBytes) return B200Z_E_CORRUPT;
        if (fcsBytes) { uint64_t fcs = 0; for (int i = 0; i < fcsBytes; i++) fcs |= (uint64_t)ip[i] << (8 * i); if (fcsBytes == 2) fcs += 256; total += fcs; }
        else unknown = 1;
        ip += fcsBytes;
        for (;;) {
            if (iend - ip < 3) return B200Z_E_CORRUPT;
            const uint32_t bh = ip[0] | (ip[1] << 8) | (ip[2] << 16); ip += 3;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) return B200Z_E_CORRUPT;
            const size_t adv = type == 1 ? 1 : bsize;
            if ((size_t)(iend - ip) < adv) return B200Z_E_CORRUPT;
            ip += adv;
            if (last) break;
        }
        if (checksum) { if (iend - ip < 4) return B200Z_E_CORRUPT; ip += 4; }
        frames++;
    }
    if (nFrames) *nFrames = frames;
    if (contentSize) *contentSize = total;
    return unknown ? B200Z_E_UNSUPPORTED : B200Z_OK;
}

int b200z_zstd_decompress_device(b200z_ctx*, const void*, size_t, void*, size_t, size_t*) { return B200Z_E_UNSUPPORTED; }
int b200z_zstd_decompress_host(b200z_ctx*, const void*, size_t, void*, size_t, size_t*) { return B200Z_E_UNSUPPORTED; }

}  // extern "C"
