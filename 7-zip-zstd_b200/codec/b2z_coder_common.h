// b2z_coder_common.h -- helpers shared by the coder classes of the codec module (ZstdCoders.cpp, Lzma2Coders.cpp).
#pragma once
#include <atomic>
#include <new>
#include <vector>
#include "b2z_7zip_abi.h"
#include "../../include/b200z.h"

namespace {

HRESULT hr_from_b200z(int rc) {
    switch (rc) {
    case B200Z_OK: return S_OK;
    case B200Z_E_MEMORY: return E_OUTOFMEMORY;
    case B200Z_E_PARAM: return E_INVALIDARG;
    case B200Z_E_CORRUPT: case B200Z_E_CHECKSUM: return S_FALSE;      // data error (ZstdDecoder.cpp:115-130)
    case B200Z_E_UNSUPPORTED: return E_NOTIMPL;
    default: return E_FAIL;                                           // incl. no device: there is no CPU fallback
    }
}

HRESULT read_stream(ISequentialInStream* s, void* data, size_t* size) {   // StreamUtils.cpp:54 semantics
    size_t want = *size; *size = 0;
    while (want) {
        UInt32 cur = want < (1u << 30) ? (UInt32)want : (1u << 30), got = 0;
        HRESULT r = s->Read(data, cur, &got);
        *size += got; data = (Byte*)data + got; want -= got;
        if (r != S_OK) return r;
        if (got == 0) return S_OK;
    }
    return S_OK;
}
HRESULT write_stream(ISequentialOutStream* s, const void* data, size_t size) {   // StreamUtils.cpp:87
    while (size) {
        UInt32 cur = size < (1u << 30) ? (UInt32)size : (1u << 30), done = 0;
        HRESULT r = s->Write(data, cur, &done);
        data = (const Byte*)data + done; size -= done;
        if (r != S_OK) return r;
        if (done == 0) return E_FAIL;
    }
    return S_OK;
}

struct PinnedBuf {                                          // pinned host staging (grown on demand)
    void* p = nullptr; size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        void* q = nullptr;
        if (b200z_host_alloc_pinned(&q, n) != 0) return false;
        if (p) { memcpy(q, p, cap); b200z_host_free_pinned(p); }
        p = q; cap = n; return true;
    }
    ~PinnedBuf() { if (p) b200z_host_free_pinned(p); }
};

template <class T> struct RefCounted : T {
    std::atomic<UInt32> refs{0};
    UInt32 AddRef() override { return ++refs; }
    UInt32 Release() override { UInt32 r = --refs; if (r == 0) delete this; return r; }
    virtual ~RefCounted() {}
};

// one C++ object exposing several COM-style interfaces: a small aggregate with inner facets
struct CoderBase {
    b200z_ctx* ctx = nullptr;
    HRESULT ensure_ctx() {
        if (ctx) return S_OK;
        // device selection without a new PROPID: B200Z_DEVICE=n pins one device; B200Z_DEVICES="0-3" / "0,2,5" / "all" names
        // the devices one Code() call is dealt over.  Default: every device of the box, as the reference defaults to every core
        // (ZstdEncoder.cpp:19 nbWorkers = #CPUs).
        if (const char* e = getenv("B200Z_DEVICE")) return hr_from_b200z(b200z_create(&ctx, atoi(e)));
        std::vector<int> devs;
        const int n = b200z_device_count();
        const char* e = getenv("B200Z_DEVICES");
        if (!e || !strcmp(e, "all")) { for (int i = 0; i < n; i++) devs.push_back(i); }
        else {
            for (const char* p = e; *p;) {
                char* q; long a = strtol(p, &q, 10); if (q == p) break;
                long b = a; if (*q == '-') { p = q + 1; b = strtol(p, &q, 10); if (q == p) break; }
                for (long v = a; v <= b && v < 1024; v++) devs.push_back((int)v);
                p = q; if (*p == ',') p++;
            }
        }
        if (devs.empty()) return hr_from_b200z(B200Z_E_NODEVICE);
        return hr_from_b200z(b200z_create_multi(&ctx, devs.data(), (int)devs.size()));
    }
    ~CoderBase() { if (ctx) b200z_destroy(ctx); }
};

}  // namespace
