// sevenz_api.cu -- non-solid .7z archives of many files in one GPU pass (BASELINE configs[4]).
//
// The reference's archive layer compresses a non-solid archive strictly one folder (= one file) at a time
// (CPP/7zip/Archive/7z/7zUpdate.cpp:2739-2810 calls Encode1 per folder and needs the packed size before the next,
// 7zEncode.cpp:482-487), so 100 000 files of 64 KiB are 100 000 sequential Code() calls.  Here the files of an archive are
// compressed together (b200z_zstd_compress_batch_crc_host: every file its own run of frames, CRC32s from the same bytes in HBM)
// and the container is written around the result: signature header, the packed streams back to back, and an uncompressed
// header -- the layout of 7zOut.cpp (COutArchive::WriteHeader :520-820, WriteSignature / WriteStartHeader :230-300) as
// DOC/7zFormat.txt states it:
//
//   SignatureHeader  '7z' BC AF 27 1C, version 0.4, CRC32 of the next 20 bytes, NextHeaderOffset, NextHeaderSize, NextHeaderCRC
//   packed streams   one per non-empty file
//   Header           MainStreamsInfo { PackInfo { sizes }, UnPackInfo { one folder per stream: coder 04F71101 + 5 property bytes,
//                    unpack sizes, CRCs } }, FilesInfo { EmptyStream / EmptyFile bit vectors, Names (UTF-16LE) [, MTime] }
//
// Everything in this file is host code (the writer is byte bookkeeping); the GPU work is the batch call.
#include <string>
#include <vector>
#include "b2z_ctx.h"

extern "C" int b200z_zstd_compress_batch_crc_host(b200z_ctx* ctx, const void* src, const uint64_t* sizes, uint32_t nFiles, void* dst, size_t dstCap,
                                                   uint64_t* dstOffsets, uint32_t* crcs);
extern "C" size_t b200z_zstd_compress_batch_bound(b200z_ctx* ctx, size_t totalBytes, uint32_t nFiles);

namespace {

enum : uint8_t { kEnd = 0x00, kHeader = 0x01, kMainStreamsInfo = 0x04, kFilesInfo = 0x05, kPackInfo = 0x06, kUnPackInfo = 0x07, kSize = 0x09, kCRC = 0x0A,
                 kFolder = 0x0B, kCodersUnPackSize = 0x0C, kEmptyStream = 0x0E, kEmptyFile = 0x0F, kName = 0x11, kMTime = 0x14 };

struct Out {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    // 7z NUMBER (7zOut.cpp WriteNumber): the count of leading one bits in the first byte = extra bytes, which hold the low bits first
    void num(uint64_t v) {
        uint8_t first = 0, mask = 0x80; int i;
        for (i = 0; i < 8; i++) {
            if (v < ((uint64_t)1 << (7 * (i + 1)))) { first |= (uint8_t)(v >> (8 * i)); break; }
            first |= mask; mask >>= 1;
        }
        b.push_back(first);
        for (; i > 0; i--) { b.push_back((uint8_t)v); v >>= 8; }
    }
    void bits(const std::vector<bool>& v) {                      // bit vector, MSB first (WriteBoolVector)
        uint8_t cur = 0, mask = 0x80;
        for (bool x : v) { if (x) cur |= mask; mask >>= 1; if (!mask) { b.push_back(cur); cur = 0; mask = 0x80; } }
        if (mask != 0x80) b.push_back(cur);
    }
};

uint32_t crc32_host(const uint8_t* p, size_t n) {                // C/7zCrc.c CrcCalc (reflected 0xEDB88320): header digests only
    static uint32_t tab[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t r = i; for (int k = 0; k < 8; k++) r = (r >> 1) ^ (B2Z_CRC32_POLY & (0u - (r & 1u))); tab[i] = r; } init = true; }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 255u] ^ (c >> 8);
    return ~c;
}

// UTF-8 -> UTF-16LE code units (names of an archive; invalid sequences become U+FFFD)
void utf8_to_utf16(const char* s, std::vector<uint16_t>& out) {
    const unsigned char* p = (const unsigned char*)s;
    while (*p) {
        uint32_t c = *p++; int extra = c < 0x80 ? 0 : ((c >> 5) == 6 ? 1 : ((c >> 4) == 14 ? 2 : ((c >> 3) == 30 ? 3 : -1)));
        if (extra < 0) { out.push_back(0xFFFD); continue; }
        if (extra) c &= (0x3Fu >> extra);
        bool bad = false;
        for (int k = 0; k < extra; k++) { if ((*p & 0xC0) != 0x80) { bad = true; break; } c = (c << 6) | (*p++ & 0x3F); }
        if (bad) { out.push_back(0xFFFD); continue; }
        if (c >= 0x10000) { c -= 0x10000; out.push_back((uint16_t)(0xD800 + (c >> 10))); out.push_back((uint16_t)(0xDC00 + (c & 0x3FF))); }
        else out.push_back((uint16_t)c);
    }
}

}  // namespace

extern "C" {

size_t b200z_7z_archive_bound(b200z_ctx* ctx, size_t totalBytes, uint32_t nFiles, size_t namesBytes) {
    return 32 + b200z_zstd_compress_batch_bound(ctx, totalBytes, nFiles) + (size_t)nFiles * 48 + namesBytes * 2 + 256;
}

// The container alone (no GPU): packed = the packed streams of the non-empty files back to back (file order), packSizes / unpackSizes / crcs
// per FILE (an empty file has pack size 0 and owns no stream), names = nFiles UTF-8 strings, level = the ZSTD coder's level byte
// (ZstdEncoder.h:17-32: {1, 5, level, 0, 0}).  mtimes: nullptr or nFiles Windows FILETIME values.
int b200z_7z_build_archive(const void* packed, const uint64_t* packSizes, const uint64_t* unpackSizes, const uint32_t* crcs, const char* const* names,
                           const uint64_t* mtimes, uint32_t nFiles, uint32_t level, void* dst, size_t dstCap, size_t* dstSize) {
    if (!packSizes || !unpackSizes || !crcs || !names || !dst || !dstSize) return B200Z_E_PARAM;
    uint64_t packTotal = 0; uint32_t nStreams = 0, nEmpty = 0;
    for (uint32_t i = 0; i < nFiles; i++) { packTotal += packSizes[i]; if (unpackSizes[i]) nStreams++; else { nEmpty++; if (packSizes[i]) return B200Z_E_PARAM; } }
    if (packTotal && !packed) return B200Z_E_PARAM;
    Out h;
    h.u8(kHeader);
    if (nStreams) {
        h.u8(kMainStreamsInfo);
        h.u8(kPackInfo); h.num(0); h.num(nStreams);
        h.u8(kSize); for (uint32_t i = 0; i < nFiles; i++) if (unpackSizes[i]) h.num(packSizes[i]);
        h.u8(kEnd);
        h.u8(kUnPackInfo);
        h.u8(kFolder); h.num(nStreams); h.u8(0);                 // External = 0
        for (uint32_t i = 0; i < nStreams; i++) {
            h.num(1);                                            // one coder
            h.u8(0x24);                                          // id size 4 | 0x20: has properties (7zOut.cpp WriteFolder)
            h.u8(0x04); h.u8(0xF7); h.u8(0x11); h.u8(0x01);      // method 4F71101 (ZstdRegister.cpp:13-17), big-endian id bytes
            h.num(5); h.u8(1); h.u8(5); h.u8((uint8_t)level); h.u8(0); h.u8(0);
        }
        h.u8(kCodersUnPackSize); for (uint32_t i = 0; i < nFiles; i++) if (unpackSizes[i]) h.num(unpackSizes[i]);
        h.u8(kCRC); h.u8(1);                                     // all defined
        for (uint32_t i = 0; i < nFiles; i++) if (unpackSizes[i]) h.u32(crcs[i]);
        h.u8(kEnd);
        h.u8(kEnd);                                              // (no SubStreamsInfo: one file per folder, digests are the folders')
    }
    h.u8(kFilesInfo); h.num(nFiles);
    if (nEmpty) {
        std::vector<bool> es(nFiles), ef;
        for (uint32_t i = 0; i < nFiles; i++) { es[i] = unpackSizes[i] == 0; if (es[i]) ef.push_back(true); }
        h.u8(kEmptyStream); h.num((nFiles + 7) / 8); h.bits(es);
        h.u8(kEmptyFile); h.num((nEmpty + 7) / 8); h.bits(ef);   // empty streams that are files (not directories)
    }
    {
        std::vector<uint16_t> n16;
        for (uint32_t i = 0; i < nFiles; i++) { if (!names[i]) return B200Z_E_PARAM; utf8_to_utf16(names[i], n16); n16.push_back(0); }
        h.u8(kName); h.num(n16.size() * 2 + 1); h.u8(0);
        for (uint16_t c : n16) { h.u8((uint8_t)c); h.u8((uint8_t)(c >> 8)); }
    }
    if (mtimes) {
        h.u8(kMTime); h.num((uint64_t)nFiles * 8 + 2); h.u8(1); h.u8(0);      // all defined, not external
        for (uint32_t i = 0; i < nFiles; i++) h.u64(mtimes[i]);
    }
    h.u8(kEnd);
    h.u8(kEnd);
    const size_t total = 32 + (size_t)packTotal + h.b.size();
    if (dstCap < total) return B200Z_E_DSTSIZE;
    uint8_t* d = (uint8_t*)dst;
    if (packTotal && (const uint8_t*)packed != d + 32) memmove(d + 32, packed, (size_t)packTotal);      // (the one-call writer compresses in place)
    memcpy(d + 32 + packTotal, h.b.data(), h.b.size());
    Out sh;
    sh.u64(packTotal); sh.u64(h.b.size()); sh.u32(crc32_host(h.b.data(), h.b.size()));
    const uint8_t sig[8] = { '7', 'z', 0xBC, 0xAF, 0x27, 0x1C, 0, 4 };
    memcpy(d, sig, 8);
    const uint32_t startCrc = crc32_host(sh.b.data(), 20);
    for (int i = 0; i < 4; i++) d[8 + i] = (uint8_t)(startCrc >> (8 * i));
    memcpy(d + 12, sh.b.data(), 20);
    *dstSize = total;
    return B200Z_OK;
}

// One call: files back to back in src (sizes[i] bytes each) -> a complete non-solid .7z archive in dst (method ZSTD, one folder per file).
int b200z_7z_write_archive_host(b200z_ctx* ctx, const void* src, const uint64_t* sizes, const char* const* names, const uint64_t* mtimes, uint32_t nFiles,
                                void* dst, size_t dstCap, size_t* dstSize) {
    if (!ctx || !sizes || !names || !dst || !dstSize) return B200Z_E_PARAM;
    uint64_t total = 0; size_t nameBytes = 0;
    for (uint32_t i = 0; i < nFiles; i++) { total += sizes[i]; if (!names[i]) return B200Z_E_PARAM; nameBytes += strlen(names[i]) + 1; }
    if (dstCap < b200z_7z_archive_bound(ctx, (size_t)total, nFiles, nameBytes)) return fail(ctx, B200Z_E_DSTSIZE, "dstCap < b200z_7z_archive_bound%s");
    std::vector<uint64_t> offs((size_t)nFiles + 1), pack(nFiles);
    std::vector<uint32_t> crcs(nFiles);
    int rc = b200z_zstd_compress_batch_crc_host(ctx, src, sizes, nFiles, (uint8_t*)dst + 32, dstCap - 32, offs.data(), crcs.data());
    if (rc) return rc;
    for (uint32_t i = 0; i < nFiles; i++) pack[i] = offs[i + 1] - offs[i];
    int64_t lv = 3; b200z_get_param(ctx, B200Z_P_LEVEL, &lv);
    rc = b200z_7z_build_archive((uint8_t*)dst + 32, pack.data(), sizes, crcs.data(), names, mtimes, nFiles, (uint32_t)lv, dst, dstCap, dstSize);
    if (rc) return fail(ctx, rc, "7z container: %s", rc == B200Z_E_DSTSIZE ? "destination too small" : "bad argument");
    return 0;
}

}  // extern "C"
