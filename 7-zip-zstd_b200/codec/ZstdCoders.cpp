// ZstdCoders.cpp -- NCompress::NZSTD::CEncoder / CDecoder for method 4F71101 on top of the
// b200z C ABI, plus the codec-module exports.  Host-side mirror of the reference wrappers:
//   CPP/7zip/Compress/ZstdEncoder.{h,cpp}  (props :51-229, 5-byte header :17-32/:245, Code :250-461)
//   CPP/7zip/Compress/ZstdDecoder.{h,cpp}  (SetDecoderProperties2 :32-49, CodeSpec :66-173)
//   CPP/7zip/Compress/ZstdRegister.cpp:13-17, Compress/CodecExports.cpp:153-378 (module exports)
// Same interface sets, property semantics, header bytes and HRESULT mapping; below them the
// stream is cut into batches of whole frames and handed to the GPU engine (b200z_zstd_*_host).
//
// Build: g++ -std=c++17 -O2 -fPIC -shared ZstdCoders.cpp -I../../include -L.. -lb200z -o ../libb200z_7z.so
#include <atomic>
#include <new>
#include <vector>
#include "b2z_coder_common.h"

namespace {

const uint64_t kZstdMethodId = 0x4F71101;
const uint32_t kZ7Major = 26, kZ7Minor = 1;                // module version reported to the host (C/7zVersion.h)
const Byte kZstdVerMajor = 1, kZstdVerMinor = 5;           // header bytes (ZstdEncoder.h:17-32)
const uint32_t kFastLevInc = Z7_ZSTD_FAST_LEV_INC, kUltimateLev = Z7_ZSTD_ULTIMATE_LEV;       // ICoder.h:163-166

// ------------------------------------------------------------------ encoder
class CEncoder final : public ICompressCoder, public ICompressSetCoderMt, public ICompressSetCoderProperties,
                       public ICompressSetCoderPropertiesOpt, public ICompressWriteCoderProperties, CoderBase {
    std::atomic<UInt32> refs_{0};
    Byte props_[5] = { kZstdVerMajor, kZstdVerMinor, 3, 0, 0 };
    int level_ = 3; bool max_ = false; UInt32 numThreads_ = 1;
    int windowLog_ = -1, hashLog_ = -1, chainLog_ = -1; bool long_ = false;
    PinnedBuf in_, out_;
public:
    UInt64 processedIn = 0, processedOut = 0;
    HRESULT QueryInterface(const GUID& iid, void** out) override {
        *out = nullptr;
        if (iid == kIID_IUnknown || iid == b2z_iid(4, kIID_Coder)) *out = static_cast<ICompressCoder*>(this);
        else if (iid == b2z_iid(4, kIID_SetMt)) *out = static_cast<ICompressSetCoderMt*>(this);
        else if (iid == b2z_iid(4, kIID_SetProps)) *out = static_cast<ICompressSetCoderProperties*>(this);
        else if (iid == b2z_iid(4, kIID_SetPropsOpt)) *out = static_cast<ICompressSetCoderPropertiesOpt*>(this);
        else if (iid == b2z_iid(4, kIID_WriteProps)) *out = static_cast<ICompressWriteCoderProperties*>(this);
        else return E_NOINTERFACE;
        ++refs_; return S_OK;
    }
    UInt32 AddRef() override { return ++refs_; }
    UInt32 Release() override { UInt32 r = --refs_; if (!r) delete this; return r; }

    HRESULT SetNumberOfThreads(UInt32 n) override {            // ZstdEncoder.cpp:463-474: clamp 1..256
        numThreads_ = n < 1 ? 1 : (n > 256 ? 256 : n); return S_OK;
    }
    HRESULT SetCoderProperties(const PROPID* ids, const PROPVARIANT* pv, UInt32 n) override {
        props_[0] = kZstdVerMajor; props_[1] = kZstdVerMinor; props_[2] = 3; props_[3] = props_[4] = 0;
        for (UInt32 i = 0; i < n; i++) {
            UInt32 v = pv[i].ulVal;
            switch (ids[i]) {
            case NCoderPropID::kNumThreads: SetNumberOfThreads(v); break;
            case NCoderPropID::kAdvMax:                       // ZstdEncoder.cpp:75-83: sets max, then falls into kLevel with the ultimate level
                if (!v) break;
                max_ = true; v = kUltimateLev;
                /* fall through */
            case NCoderPropID::kLevel: {                      // ZstdEncoder.cpp:84-104
                UInt32 lev = !max_ ? v : kUltimateLev;
                if (v < 1) lev = 1;
                else if (v > 22) {
                    if (v > kFastLevInc && lev != kUltimateLev) { v -= kFastLevInc; goto fast; }   // inverter: 32 + f selects fast level f
                    max_ = (lev == kUltimateLev);             // 255 from the GUI / kAdvMax = "max"
                    lev = 22;
                }
                level_ = lev == kUltimateLev ? 22 : (int)lev; // (max set earlier and a level <= 22 now: the reference keeps 255, which zstd clamps to its maximum)
                props_[2] = (Byte)(!max_ ? lev : kUltimateLev);
                break;
            }
            case NCoderPropID::kFast:
            fast:
                if (!max_) {
                    if (v < 1) v = 1; if (v > 64) v = 64;
                    level_ = -(int)v; props_[2] = (Byte)(v + kFastLevInc); break;
                }
                /* fall through (as the reference does when max is set: the value is read as kLong's) */
            case NCoderPropID::kLong: long_ = true; windowLog_ = v == 0 ? 27 : (int)(v < 10 ? 10 : (v > 31 ? 31 : v)); break;   // ZstdEncoder.cpp:128-146
            case NCoderPropID::kWindowLog: if (v < 10 || v > 31) return E_INVALIDARG; windowLog_ = (int)v; break;
            case NCoderPropID::kHashLog: if (v < 6 || v > 30) return E_INVALIDARG; hashLog_ = (int)v; break;
            case NCoderPropID::kChainLog: if (v < 6 || v > 30) return E_INVALIDARG; chainLog_ = (int)v; break;
            case NCoderPropID::kStrategy: case NCoderPropID::kSearchLog: case NCoderPropID::kMinMatch: case NCoderPropID::kTargetLen:
            case NCoderPropID::kOverlapLog: case NCoderPropID::kLdmHashLog: case NCoderPropID::kLdmSearchLength:
            case NCoderPropID::kLdmBucketSizeLog: case NCoderPropID::kLdmHashRateLog:
                break;                                        // accepted; the GPU parser has no equivalent knob yet
            default: break;
            }
        }
        return S_OK;
    }
    HRESULT SetCoderPropertiesOpt(const PROPID*, const PROPVARIANT*, UInt32) override { return S_OK; }   // kExpectedDataSize: not needed
    HRESULT WriteCoderProperties(ISequentialOutStream* out) override { return write_stream(out, props_, 5); }

    HRESULT Code(ISequentialInStream* inS, ISequentialOutStream* outS, const UInt64*, const UInt64*, ICompressProgressInfo* progress) override {
        processedIn = processedOut = 0;
        HRESULT hr = ensure_ctx(); if (hr != S_OK) return hr;
        b200z_set_param(ctx, B200Z_P_LEVEL, level_ < 1 ? 1 : level_);
        b200z_set_param(ctx, B200Z_P_FLAGS, 1);               // mcmilk MT frame convention: size hint before each frame
        if (hashLog_ >= 10 && hashLog_ <= 22) b200z_set_param(ctx, B200Z_P_HASHLOG_L, hashLog_);
        if (chainLog_ >= 10 && chainLog_ <= 22) b200z_set_param(ctx, B200Z_P_HASHLOG_S, chainLog_);
        // long=N (ZstdEncoder.cpp:322-331: long-distance matching on, window 2^N): the engine's long mode, window 2^N in frames of 8 windows
        // (21..27; a smaller N has nothing beyond stage F's reach to find, a larger one is cut to the 128 MiB the format's decoders accept by default)
        if (long_ && windowLog_ >= 21) b200z_set_param(ctx, B200Z_P_LONG, windowLog_ > 27 ? 27 : windowLog_);
        else {
            b200z_set_param(ctx, B200Z_P_LONG, 0);
            if (windowLog_ >= 17) { int64_t fl = 0; b200z_get_param(ctx, B200Z_P_FRAMELOG, &fl); b200z_set_param(ctx, B200Z_P_WINDOWLOG, windowLog_ < fl ? windowLog_ : fl); }
        }
        // batches of whole frames: 1 GiB of input per GPU pass (pinned host memory); a batch has to hold about a
        // thousand 1 MiB frames to keep the frame-parallel match finder busy
        const int nDev = b200z_device_list(ctx, nullptr, 0);
        const size_t batch = (size_t)1 << (nDev >= 4 ? 32 : (nDev >= 2 ? 31 : 30));     // the call is dealt over the devices: give them a round each
        if (!in_.reserve(batch) || !out_.reserve(b200z_zstd_compress_bound(ctx, batch))) return E_OUTOFMEMORY;
        bool wroteAny = false;
        for (;;) {
            size_t got = batch;
            hr = read_stream(inS, in_.p, &got);
            if (hr != S_OK) return hr;
            if (got == 0 && wroteAny) break;
            size_t produced = 0;
            int rc = b200z_zstd_compress_host(ctx, in_.p, got, out_.p, out_.cap, &produced);
            if (rc) return hr_from_b200z(rc);
            hr = write_stream(outS, out_.p, produced);
            if (hr != S_OK) return hr;
            wroteAny = true;
            processedIn += got; processedOut += produced;
            if (progress) { hr = progress->SetRatioInfo(&processedIn, &processedOut); if (hr != S_OK) return hr; }   // E_ABORT on user break
            if (got < batch) break;
        }
        return S_OK;
    }
};

// ------------------------------------------------------------------ decoder
class CDecoder final : public ICompressCoder, public ICompressSetDecoderProperties2, public ICompressSetCoderMt,
                       public ICompressSetOutStreamSize, public ICompressSetInStream, public ISequentialInStream, CoderBase {
    std::atomic<UInt32> refs_{0};
    // The packed stream is read piece by piece into bounded pinned staging and decoded in BATCHES of whole frames (about kOutTarget
    // decoded bytes each; one frame larger than that still goes alone): memory is bounded by the batch, not by the folder, and output
    // reaches the caller batch by batch -- the streaming the reference gets from ZSTD_decompressStream (ZstdDecoder.cpp:108-173).
    static constexpr size_t kInStep = (size_t)256 << 20;         // granularity of reads / growth of the input staging
    static constexpr uint64_t kOutTarget = (uint64_t)1 << 30;    // decoded bytes per GPU batch (the staging holds about that much packed data)
    static constexpr uint64_t kOutLimit = (uint64_t)48 << 30;    // a single frame that needs more staging than this is refused (E_OUTOFMEMORY), never attempted
    PinnedBuf in_, out_;
    size_t inFill_ = 0; bool inEof_ = false;
    size_t outSize_ = 0, outPos_ = 0;                            // current decoded batch in out_ and how much of it the caller has taken (pull mode)
    // pull mode (ZstdDecoder.cpp:182-258): SetInStream + SetOutStreamSize, then Read() until it returns 0 bytes
    ISequentialInStream* pullIn_ = nullptr; bool pullDone_ = false;
public:
    ~CDecoder() { if (pullIn_) pullIn_->Release(); }
    UInt64 processedIn = 0, processedOut = 0;
    HRESULT QueryInterface(const GUID& iid, void** out) override {
        *out = nullptr;
        if (iid == kIID_IUnknown || iid == b2z_iid(4, kIID_Coder)) *out = static_cast<ICompressCoder*>(this);
        else if (iid == b2z_iid(4, kIID_SetDecProps2)) *out = static_cast<ICompressSetDecoderProperties2*>(this);
        else if (iid == b2z_iid(4, kIID_SetMt)) *out = static_cast<ICompressSetCoderMt*>(this);
        else if (iid == b2z_iid(4, kIID_SetOutStreamSize)) *out = static_cast<ICompressSetOutStreamSize*>(this);
        else if (iid == b2z_iid(4, kIID_SetInStream)) *out = static_cast<ICompressSetInStream*>(this);
        else if (iid == b2z_iid(3, kIID_SeqIn)) *out = static_cast<ISequentialInStream*>(this);
        else return E_NOINTERFACE;
        ++refs_; return S_OK;
    }
    UInt32 AddRef() override { return ++refs_; }
    UInt32 Release() override { UInt32 r = --refs_; if (!r) delete this; return r; }
    void reset_stream() { inFill_ = 0; inEof_ = false; outSize_ = outPos_ = 0; pullDone_ = false; processedIn = processedOut = 0; }
    HRESULT SetOutStreamSize(const UInt64*) override { reset_stream(); return S_OK; }   // ZstdDecoder.cpp:57-64: (re)initialises the stream state; the
                                                                                        // size itself is not needed: frames end where their last block ends
    HRESULT SetInStream(ISequentialInStream* in) override { if (in) in->AddRef(); if (pullIn_) pullIn_->Release(); pullIn_ = in; return S_OK; }
    HRESULT ReleaseInStream() override { if (pullIn_) pullIn_->Release(); pullIn_ = nullptr; return S_OK; }
    HRESULT Read(void* data, UInt32 size, UInt32* processed) override {
        if (processed) *processed = 0;
        if (!pullIn_) return E_FAIL;
        while (outPos_ == outSize_ && !pullDone_) {                      // current batch taken: decode the next one
            bool end = false;
            HRESULT hr = next_batch(pullIn_, &end);
            if (hr != S_OK) return hr;
            if (end) pullDone_ = true;
        }
        size_t n = outSize_ - outPos_; if (n > size) n = size;
        memcpy(data, (const Byte*)out_.p + outPos_, n); outPos_ += n;
        if (processed) *processed = (UInt32)n;
        return S_OK;
    }
    HRESULT SetDecoderProperties2(const Byte*, UInt32 size) override {   // ZstdDecoder.cpp:32-49: 1/3/5 bytes, content ignored
        return (size == 1 || size == 3 || size == 5) ? S_OK : E_NOTIMPL;
    }
    HRESULT SetNumberOfThreads(UInt32) override { return S_OK; }         // no-op, as in ZstdDecoder.cpp:260-263

    // Decodes the next batch of whole frames into out_[0, outSize_).  *end: the packed stream is exhausted (outSize_ = 0).
    HRESULT next_batch(ISequentialInStream* inS, bool* end) {
        *end = false; outSize_ = outPos_ = 0;
        HRESULT hr = ensure_ctx(); if (hr != S_OK) return hr;
        for (;;) {
            // top the staging up (a short read is not the end: StreamUtils.cpp:54 semantics live in read_stream)
            if (!inEof_ && (in_.cap - inFill_ < kInStep / 2 || in_.cap == 0)) { if (!in_.reserve(inFill_ + kInStep)) return E_OUTOFMEMORY; }
            if (!inEof_ && inFill_ < in_.cap) {
                size_t got = in_.cap - inFill_;
                hr = read_stream(inS, (Byte*)in_.p + inFill_, &got);
                if (hr != S_OK) return hr;
                if (got < in_.cap - inFill_) inEof_ = true;
                inFill_ += got;
            }
            if (inFill_ == 0) { *end = true; return S_OK; }
            size_t used = 0; uint64_t bound = 0; uint32_t frames = 0;
            const int rc = b200z_zstd_frame_prefix(in_.p, inFill_, kOutTarget, &used, &bound, &frames);
            if (rc == B200Z_E_CORRUPT && frames == 0) return S_FALSE;   // (complete frames in front of the damage are still delivered first)
            if (frames == 0) {
                if (inEof_) return S_FALSE;                              // the stream ends inside a frame
                if (!in_.reserve(in_.cap + (in_.cap > kInStep ? in_.cap : kInStep))) return E_OUTOFMEMORY;   // one frame larger than the staging: grow, read on
                continue;
            }
            if (bound > kOutLimit) return E_OUTOFMEMORY;
            if (!out_.reserve((size_t)bound + 64)) return E_OUTOFMEMORY;
            size_t produced = 0;
            const int drc = b200z_zstd_decompress_host(ctx, in_.p, used, out_.p, (size_t)bound, &produced);
            if (drc == B200Z_E_DSTSIZE) return S_FALSE;                  // the blocks regenerate more than the frame declared / than blocks can hold
            if (drc) return hr_from_b200z(drc);
            memmove(in_.p, (const Byte*)in_.p + used, inFill_ - used); inFill_ -= used;
            processedIn += used; processedOut += produced;
            outSize_ = produced;
            if (produced == 0 && inFill_ == 0 && inEof_) *end = true;   // only empty / skippable frames were left
            return S_OK;
        }
    }

    HRESULT Code(ISequentialInStream* inS, ISequentialOutStream* outS, const UInt64*, const UInt64*, ICompressProgressInfo* progress) override {
        reset_stream();
        for (;;) {
            bool end = false;
            HRESULT hr = next_batch(inS, &end);
            if (hr != S_OK) return hr;
            if (outSize_) { hr = write_stream(outS, out_.p, outSize_); if (hr != S_OK) return hr; }
            if (progress) { hr = progress->SetRatioInfo(&processedIn, &processedOut); if (hr != S_OK) return hr; }
            if (end) return S_OK;
        }
    }
};

}  // namespace

// ------------------------------------------------------------------ module exports (CodecExports.cpp:153-378)
// Methods, in the reference's registration order: ZSTD (ZstdRegister.cpp:13-17), LZMA2 (Lzma2Register.cpp:16-20) and
// FLZMA2 (FastLzma2Register.cpp:13-18) -- the last two share ID 0x21: encoders are chosen by name, decoders by ID.
ICompressCoder* b2z_new_lzma2_encoder(bool fast);        // Lzma2Coders.cpp
ICompressCoder* b2z_new_lzma2_decoder();

namespace {
const uint64_t kLzma2MethodId = 0x21;
struct MethodInfo { uint64_t id; const char* name; };
const MethodInfo kMethods[3] = { { kZstdMethodId, "ZSTD" }, { kLzma2MethodId, "LZMA2" }, { kLzma2MethodId, "FLZMA2" } };
// class ids: the reference derives them from the method id alone, so LZMA2 and FLZMA2 encoders would collide; as in
// CodecExports.cpp:95-125 CreateObject resolves a clsid to the FIRST method with that id (LZMA2)
}

extern "C" {

HRESULT GetNumberOfMethods(UInt32* n) { *n = 3; return S_OK; }

HRESULT GetMethodProperty(UInt32 index, PROPID propID, PROPVARIANT* value) {
    memset(value, 0, sizeof(*value));
    if (index >= 3) return E_INVALIDARG;
    const MethodInfo& m = kMethods[index];
    switch (propID) {
    case NMethodPropID::kID: value->vt = VT_UI8; value->uhVal = m.id; break;
    case NMethodPropID::kName: value->bstrVal = b2z_alloc_bstr_ascii(m.name); if (!value->bstrVal) return E_OUTOFMEMORY; value->vt = VT_BSTR; break;
    case NMethodPropID::kDecoder: case NMethodPropID::kEncoder: {
        const GUID g = b2z_clsid(propID == NMethodPropID::kEncoder, m.id);
        value->bstrVal = b2z_alloc_bstr_bytes(&g, sizeof(g)); if (!value->bstrVal) return E_OUTOFMEMORY; value->vt = VT_BSTR; break; }
    case NMethodPropID::kDecoderIsAssigned: case NMethodPropID::kEncoderIsAssigned: value->vt = VT_BOOL; value->boolVal = -1; break;
    case NMethodPropID::kIsFilter: value->vt = VT_BOOL; value->boolVal = 0; break;
    default: break;                                            // kPackStreams: 1 stream -> left empty, as the reference does
    }
    return S_OK;
}

static HRESULT create_coder(UInt32 index, bool encoder, const GUID* iid, void** out) {
    *out = nullptr;
    if (index >= 3) return E_INVALIDARG;
    if (!(*iid == b2z_iid(4, kIID_Coder))) return E_NOINTERFACE;
    try {
        ICompressCoder* c;
        if (index == 0) c = encoder ? static_cast<ICompressCoder*>(new CEncoder()) : static_cast<ICompressCoder*>(new CDecoder());
        else c = encoder ? b2z_new_lzma2_encoder(index == 2) : b2z_new_lzma2_decoder();
        c->AddRef(); *out = c; return S_OK;
    } catch (...) { return E_OUTOFMEMORY; }
}
HRESULT CreateEncoder(UInt32 index, const GUID* iid, void** out) { return create_coder(index, true, iid, out); }
HRESULT CreateDecoder(UInt32 index, const GUID* iid, void** out) { return create_coder(index, false, iid, out); }

HRESULT CreateObject(const GUID* clsid, const GUID* iid, void** out) {
    *out = nullptr;
    for (UInt32 i = 0; i < 3; i++) {
        if (*clsid == b2z_clsid(true, kMethods[i].id)) return create_coder(i, true, iid, out);
        if (*clsid == b2z_clsid(false, kMethods[i].id)) return create_coder(i, false, iid, out);
    }
    return CLASS_E_CLASSNOTAVAILABLE;
}

HRESULT GetModuleProp(PROPID propID, PROPVARIANT* value) {
    memset(value, 0, sizeof(*value));
    if (propID == NModulePropID::kInterfaceType) { value->vt = VT_UI4; value->ulVal = 0; }          // IUnknown without virtual destructor
    else if (propID == NModulePropID::kVersion) { value->vt = VT_UI4; value->ulVal = (kZ7Major << 16) | kZ7Minor; }
    return S_OK;
}

}  // extern "C"
