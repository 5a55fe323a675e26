// coder_roundtrip.cpp -- drives libb200z_7z.so exactly as 7-Zip's codec loader does
// (CPP/7zip/UI/Common/LoadCodecs.cpp:279-303,528-563: dlsym the exports, GetModuleProp check,
// GetMethodProperty, CreateEncoder/CreateDecoder by index, then ICompressCoder::Code()).
// usage: coder_roundtrip <lib.so> <input file> <packed output file> [level] [zstd|lzma2|flzma2]      (needs a GPU)
//        coder_roundtrip <lib.so> --exports                                      (no GPU needed)
#include <dlfcn.h>
#include <cstdio>
#include <string>
#include <vector>
#include "../../7-zip-zstd_b200/codec/b2z_7zip_abi.h"

struct MemIn final : ISequentialInStream {
    const std::vector<Byte>& d; size_t pos = 0; UInt32 refs = 1;
    explicit MemIn(const std::vector<Byte>& v) : d(v) {}
    HRESULT QueryInterface(const GUID&, void** o) override { *o = nullptr; return E_NOINTERFACE; }
    UInt32 AddRef() override { return ++refs; }
    UInt32 Release() override { return --refs; }
    HRESULT Read(void* data, UInt32 size, UInt32* processed) override {
        size_t n = d.size() - pos; if (n > size) n = size; if (n > 100000) n = 100000 + (pos % 7777);   // short reads on purpose
        if (n > d.size() - pos) n = d.size() - pos;
        memcpy(data, d.data() + pos, n); pos += n; if (processed) *processed = (UInt32)n; return S_OK;
    }
};
struct MemOut final : ISequentialOutStream {
    std::vector<Byte> d; UInt32 refs = 1;
    HRESULT QueryInterface(const GUID&, void** o) override { *o = nullptr; return E_NOINTERFACE; }
    UInt32 AddRef() override { return ++refs; }
    UInt32 Release() override { return --refs; }
    HRESULT Write(const void* data, UInt32 size, UInt32* processed) override {
        d.insert(d.end(), (const Byte*)data, (const Byte*)data + size); if (processed) *processed = size; return S_OK;
    }
};
struct Progress final : ICompressProgressInfo {
    int calls = 0; UInt32 refs = 1;
    HRESULT QueryInterface(const GUID&, void** o) override { *o = nullptr; return E_NOINTERFACE; }
    UInt32 AddRef() override { return ++refs; }
    UInt32 Release() override { return --refs; }
    HRESULT SetRatioInfo(const UInt64*, const UInt64*) override { calls++; return S_OK; }
};

// pull mode, as archive handlers that read from a coder do (ICompressSetInStream + ICompressSetOutStreamSize + ISequentialInStream::Read)
static int pull_decode(ICompressCoder* dec, const std::vector<Byte>& packed, UInt64 outSize, std::vector<Byte>& out) {
    ICompressSetInStream* si = nullptr; ICompressSetOutStreamSize* so = nullptr; ISequentialInStream* rd = nullptr;
    if (dec->QueryInterface(b2z_iid(4, kIID_SetInStream), (void**)&si) != S_OK) return 1;
    if (dec->QueryInterface(b2z_iid(4, kIID_SetOutStreamSize), (void**)&so) != S_OK) return 2;
    if (dec->QueryInterface(b2z_iid(3, kIID_SeqIn), (void**)&rd) != S_OK) return 3;
    MemIn in(packed);
    if (si->SetInStream(&in) != S_OK || so->SetOutStreamSize(&outSize) != S_OK) return 4;
    out.clear();
    std::vector<Byte> buf(1 << 20);
    for (;;) {
        UInt32 got = 0;
        if (rd->Read(buf.data(), (UInt32)buf.size() - 12345u, &got) != S_OK) return 5;
        if (!got) break;
        out.insert(out.end(), buf.begin(), buf.begin() + got);
    }
    if (si->ReleaseInStream() != S_OK) return 6;
    si->Release(); so->Release(); rd->Release();
    return 0;
}

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    auto GetNumberOfMethods = (HRESULT(*)(UInt32*))dlsym(h, "GetNumberOfMethods");
    auto GetMethodProperty = (HRESULT(*)(UInt32, PROPID, PROPVARIANT*))dlsym(h, "GetMethodProperty");
    auto CreateEncoder = (HRESULT(*)(UInt32, const GUID*, void**))dlsym(h, "CreateEncoder");
    auto CreateDecoder = (HRESULT(*)(UInt32, const GUID*, void**))dlsym(h, "CreateDecoder");
    auto CreateObject = (HRESULT(*)(const GUID*, const GUID*, void**))dlsym(h, "CreateObject");
    auto GetModuleProp = (HRESULT(*)(PROPID, PROPVARIANT*))dlsym(h, "GetModuleProp");
    CHECK(GetNumberOfMethods && GetMethodProperty && CreateEncoder && CreateDecoder && CreateObject && GetModuleProp);
    UInt32 n = 0; CHECK(GetNumberOfMethods(&n) == S_OK && n == 3);
    PROPVARIANT v;
    CHECK(GetModuleProp(NModulePropID::kInterfaceType, &v) == S_OK && v.vt == VT_UI4 && v.ulVal == 0);
    CHECK(GetModuleProp(NModulePropID::kVersion, &v) == S_OK && v.ulVal == ((26u << 16) | 1u));
    CHECK(GetMethodProperty(0, NMethodPropID::kID, &v) == S_OK && v.vt == VT_UI8 && v.uhVal == 0x4F71101);
    CHECK(GetMethodProperty(0, NMethodPropID::kName, &v) == S_OK && v.vt == VT_BSTR && v.bstrVal[0] == L'Z' && v.bstrVal[3] == L'D' && v.bstrVal[4] == 0);
    CHECK(GetMethodProperty(0, NMethodPropID::kEncoder, &v) == S_OK && v.vt == VT_BSTR);
    { GUID g; memcpy(&g, v.bstrVal, 16); CHECK(g == b2z_clsid(true, 0x4F71101)); uint32_t len; memcpy(&len, (char*)v.bstrVal - 4, 4); CHECK(len == 16); }
    CHECK(GetMethodProperty(0, NMethodPropID::kIsFilter, &v) == S_OK && v.vt == VT_BOOL && v.boolVal == 0);
    // methods 1, 2: LZMA2 and FLZMA2 share ID 0x21 (Lzma2Register.cpp:16-20, FastLzma2Register.cpp:13-18)
    CHECK(GetMethodProperty(1, NMethodPropID::kID, &v) == S_OK && v.uhVal == 0x21 && GetMethodProperty(2, NMethodPropID::kID, &v) == S_OK && v.uhVal == 0x21);
    CHECK(GetMethodProperty(1, NMethodPropID::kName, &v) == S_OK && v.bstrVal[0] == L'L' && v.bstrVal[4] == L'2' && v.bstrVal[5] == 0);
    CHECK(GetMethodProperty(2, NMethodPropID::kName, &v) == S_OK && v.bstrVal[0] == L'F' && v.bstrVal[5] == L'2' && v.bstrVal[6] == 0);
    CHECK(GetMethodProperty(3, NMethodPropID::kID, &v) == E_INVALIDARG);
    const GUID iidCoder = b2z_iid(4, kIID_Coder);
    void* obj = nullptr;
    {   // LZMA2 coder objects: interface sets and property semantics (no GPU needed)
        void* o = nullptr; CHECK(CreateEncoder(2, &iidCoder, &o) == S_OK && o);
        ICompressCoder* fe = (ICompressCoder*)o; ICompressSetCoderProperties* fs = nullptr; ICompressWriteCoderProperties* fw = nullptr; void* none = nullptr;
        CHECK(fe->QueryInterface(b2z_iid(4, kIID_SetProps), (void**)&fs) == S_OK && fe->QueryInterface(b2z_iid(4, kIID_WriteProps), (void**)&fw) == S_OK);
        CHECK(fe->QueryInterface(b2z_iid(4, kIID_SetPropsOpt), &none) == E_NOINTERFACE);              // CFastEncoder has no ...PropertiesOpt
        PROPID ia[1] = { NCoderPropID::kAlgorithm }; PROPVARIANT pa[1]; memset(pa, 0, sizeof(pa)); pa[0].vt = VT_UI4; pa[0].ulVal = 4;
        CHECK(fs->SetCoderProperties(ia, pa, 1) == E_INVALIDARG);                                    // Lzma2Encoder.cpp:197-199
        PROPID ib[2] = { NCoderPropID::kDictionarySize, NCoderPropID::kLevel }; PROPVARIANT pb[2]; memset(pb, 0, sizeof(pb));
        pb[0].vt = VT_UI4; pb[0].ulVal = 1u << 22; pb[1].vt = VT_UI4; pb[1].ulVal = 5;
        CHECK(fs->SetCoderProperties(ib, pb, 2) == S_OK);
        MemOut ph; CHECK(fw->WriteCoderProperties(&ph) == S_OK && ph.d.size() == 1 && ph.d[0] == 20);   // 4 MiB dictionary -> property 20
        fs->Release(); fw->Release(); CHECK(fe->Release() == 0);
        CHECK(CreateDecoder(1, &iidCoder, &o) == S_OK && o);
        ICompressCoder* ld = (ICompressCoder*)o; ICompressSetDecoderProperties2* lp = nullptr; ICompressSetFinishMode* lf = nullptr; ICompressGetInStreamProcessedSize* lg = nullptr;
        CHECK(ld->QueryInterface(b2z_iid(4, kIID_SetDecProps2), (void**)&lp) == S_OK && ld->QueryInterface(b2z_iid(4, kIID_SetFinishMode), (void**)&lf) == S_OK);
        CHECK(ld->QueryInterface(b2z_iid(4, kIID_GetInProcessed), (void**)&lg) == S_OK);
        { ICompressSetBufSize* bs = nullptr; ICompressSetMemLimit* ml = nullptr; ICompressSetInStream* si = nullptr; ICompressSetOutStreamSize* so = nullptr; ISequentialInStream* rd = nullptr;
          CHECK(ld->QueryInterface(b2z_iid(4, kIID_SetBufSize), (void**)&bs) == S_OK && bs->SetInBufSize(0, 1 << 20) == S_OK && bs->SetOutBufSize(0, 1 << 22) == S_OK);
          CHECK(ld->QueryInterface(b2z_iid(4, kIID_SetMemLimit), (void**)&ml) == S_OK && ml->SetMemLimit((UInt64)1 << 30) == S_OK);
          CHECK(ld->QueryInterface(b2z_iid(4, kIID_SetInStream), (void**)&si) == S_OK && ld->QueryInterface(b2z_iid(4, kIID_SetOutStreamSize), (void**)&so) == S_OK);
          CHECK(ld->QueryInterface(b2z_iid(3, kIID_SeqIn), (void**)&rd) == S_OK);
          UInt32 got = 7; Byte tmp[4]; CHECK(rd->Read(tmp, 4, &got) == E_FAIL && got == 0);              // no input stream set
          bs->Release(); ml->Release(); si->Release(); so->Release(); rd->Release(); }
        const Byte ok1[1] = { 24 }, bad1[1] = { 41 };
        CHECK(lp->SetDecoderProperties2(ok1, 1) == S_OK && lp->SetDecoderProperties2(bad1, 1) == E_NOTIMPL && lp->SetDecoderProperties2(ok1, 5) == E_NOTIMPL);
        lp->Release(); lf->Release(); lg->Release(); CHECK(ld->Release() == 0);
    }
    const std::string method = argc > 5 ? argv[5] : "zstd";
    if (method != "zstd" && std::string(argv[2]) != "--exports") {
        const UInt32 idx = method == "flzma2" ? 2 : 1;
        std::vector<Byte> input;
        { FILE* f = fopen(argv[2], "rb"); CHECK(f); Byte buf[1 << 16]; size_t k; while ((k = fread(buf, 1, sizeof(buf), f)) > 0) input.insert(input.end(), buf, buf + k); fclose(f); }
        void* o = nullptr; CHECK(CreateEncoder(idx, &iidCoder, &o) == S_OK && o);
        ICompressCoder* e = (ICompressCoder*)o; ICompressSetCoderProperties* s = nullptr; ICompressWriteCoderProperties* w = nullptr;
        CHECK(e->QueryInterface(b2z_iid(4, kIID_SetProps), (void**)&s) == S_OK && e->QueryInterface(b2z_iid(4, kIID_WriteProps), (void**)&w) == S_OK);
        PROPID ids2[2] = { NCoderPropID::kLevel, NCoderPropID::kNumThreads }; PROPVARIANT pv2[2]; memset(pv2, 0, sizeof(pv2));
        pv2[0].vt = VT_UI4; pv2[0].ulVal = (UInt32)(argc > 4 ? atoi(argv[4]) : 5); pv2[1].vt = VT_UI4; pv2[1].ulVal = 8;
        CHECK(s->SetCoderProperties(ids2, pv2, 2) == S_OK);
        MemOut ph; CHECK(w->WriteCoderProperties(&ph) == S_OK && ph.d.size() == 1 && ph.d[0] == 16);
        MemIn in(input); MemOut packed; Progress prog;
        HRESULT r = e->Code(&in, &packed, nullptr, nullptr, &prog);
        if (r != S_OK) { fprintf(stderr, "lzma2 encoder Code() = 0x%08x\n", (unsigned)r); return 1; }
        CHECK(prog.calls >= 1 && !packed.d.empty() && packed.d.back() == 0);
        CHECK(CreateDecoder(idx, &iidCoder, &o) == S_OK);
        ICompressCoder* d = (ICompressCoder*)o; ICompressSetDecoderProperties2* dp = nullptr; ICompressSetFinishMode* fm = nullptr; ICompressGetInStreamProcessedSize* gp = nullptr;
        CHECK(d->QueryInterface(b2z_iid(4, kIID_SetDecProps2), (void**)&dp) == S_OK && d->QueryInterface(b2z_iid(4, kIID_SetFinishMode), (void**)&fm) == S_OK);
        CHECK(d->QueryInterface(b2z_iid(4, kIID_GetInProcessed), (void**)&gp) == S_OK);
        CHECK(dp->SetDecoderProperties2(ph.d.data(), 1) == S_OK && fm->SetFinishMode(1) == S_OK);
        MemIn pin(packed.d); MemOut back; UInt64 outSize = input.size();
        r = d->Code(&pin, &back, nullptr, &outSize, nullptr);
        if (r != S_OK) { fprintf(stderr, "lzma2 decoder Code() = 0x%08x\n", (unsigned)r); return 1; }
        CHECK(back.d == input);
        UInt64 inProc = 0; CHECK(gp->GetInStreamProcessedSize(&inProc) == S_OK && inProc == packed.d.size());
        { std::vector<Byte> pulled; int pr = pull_decode(d, packed.d, outSize, pulled); if (pr) { fprintf(stderr, "lzma2 pull mode failed at step %d\n", pr); return 1; } CHECK(pulled == input); }
        { std::vector<Byte> bad(packed.d.begin(), packed.d.begin() + packed.d.size() / 2); MemIn bi(bad); MemOut bo; CHECK(d->Code(&bi, &bo, nullptr, &outSize, nullptr) == S_FALSE); }
        { UInt64 wrong = input.size() + 1; MemIn p2(packed.d); MemOut b2; CHECK(d->Code(&p2, &b2, nullptr, &wrong, nullptr) == S_FALSE); }   // finish mode: sizes must agree
        { FILE* f = fopen(argv[3], "wb"); CHECK(f); fwrite(packed.d.data(), 1, packed.d.size(), f); fclose(f); }
        s->Release(); w->Release(); dp->Release(); fm->Release(); gp->Release();
        CHECK(e->Release() == 0 && d->Release() == 0);
        printf("roundtrip ok: %zu -> %zu bytes (%s)\n", input.size(), packed.d.size(), method.c_str());
        return 0;
    }
    CHECK(CreateEncoder(0, &iidCoder, &obj) == S_OK && obj);
    ICompressCoder* enc = (ICompressCoder*)obj;
    ICompressSetCoderProperties* sp = nullptr; ICompressWriteCoderProperties* wp = nullptr; ICompressSetCoderMt* mt = nullptr;
    CHECK(enc->QueryInterface(b2z_iid(4, kIID_SetProps), (void**)&sp) == S_OK);
    CHECK(enc->QueryInterface(b2z_iid(4, kIID_WriteProps), (void**)&wp) == S_OK);
    CHECK(enc->QueryInterface(b2z_iid(4, kIID_SetMt), (void**)&mt) == S_OK);
    void* bogus = nullptr; CHECK(enc->QueryInterface(b2z_iid(4, 0x99), &bogus) == E_NOINTERFACE);
    const int level = argc > 4 ? atoi(argv[4]) : 3;
    PROPID ids[2] = { NCoderPropID::kLevel, NCoderPropID::kNumThreads }; PROPVARIANT pv[2]; memset(pv, 0, sizeof(pv));
    pv[0].vt = VT_UI4; pv[0].ulVal = (UInt32)level; pv[1].vt = VT_UI4; pv[1].ulVal = 8;
    CHECK(sp->SetCoderProperties(ids, pv, 2) == S_OK);
    MemOut hdr; CHECK(wp->WriteCoderProperties(&hdr) == S_OK && hdr.d.size() == 5 && hdr.d[0] == 1 && hdr.d[1] == 5 && hdr.d[2] == (Byte)level);
    // fast-level inverter and max (ZstdEncoder.cpp:81-121)
    { PROPID i2[1] = { NCoderPropID::kFast }; PROPVARIANT p2[1]; memset(p2, 0, sizeof(p2)); p2[0].vt = VT_UI4; p2[0].ulVal = 5;
      void* o2 = nullptr; CHECK(CreateObject(new GUID(b2z_clsid(true, 0x4F71101)), &iidCoder, &o2) == S_OK);
      ICompressCoder* e2 = (ICompressCoder*)o2; ICompressSetCoderProperties* s2; ICompressWriteCoderProperties* w2;
      e2->QueryInterface(b2z_iid(4, kIID_SetProps), (void**)&s2); e2->QueryInterface(b2z_iid(4, kIID_WriteProps), (void**)&w2);
      CHECK(s2->SetCoderProperties(i2, p2, 1) == S_OK); MemOut h2; w2->WriteCoderProperties(&h2); CHECK(h2.d[2] == 32 + 5);
      s2->Release(); w2->Release(); CHECK(e2->Release() == 0); }
    // kLevel 255 (what MethodProps.cpp:667 / HandlerOut.cpp:196 send for "max") and kAdvMax write header byte 255 = Z7_ZSTD_ULTIMATE_LEV
    // (ICoder.h:166; 7zHandler.cpp:607,628 prints it as ZSTD:max); 32 + f through kLevel is the fast-level inverter; > 22 clamps to 22
    { struct Case { PROPID id; UInt32 val; Byte want; } cases[] = {
          { NCoderPropID::kLevel, 255, 255 }, { NCoderPropID::kAdvMax, 1, 255 }, { NCoderPropID::kLevel, 32 + 7, 32 + 7 }, { NCoderPropID::kLevel, 30, 22 },
          { NCoderPropID::kLevel, 0, 1 }, { NCoderPropID::kLevel, 22, 22 }, { NCoderPropID::kFast, 200, 32 + 64 }, { NCoderPropID::kAdvMax, 0, 3 } };
      for (const Case& c : cases) {
          void* o2 = nullptr; CHECK(CreateEncoder(0, &iidCoder, &o2) == S_OK);
          ICompressCoder* e2 = (ICompressCoder*)o2; ICompressSetCoderProperties* s2; ICompressWriteCoderProperties* w2;
          e2->QueryInterface(b2z_iid(4, kIID_SetProps), (void**)&s2); e2->QueryInterface(b2z_iid(4, kIID_WriteProps), (void**)&w2);
          PROPID i2[1] = { c.id }; PROPVARIANT p2[1]; memset(p2, 0, sizeof(p2)); p2[0].vt = VT_UI4; p2[0].ulVal = c.val;
          CHECK(s2->SetCoderProperties(i2, p2, 1) == S_OK); MemOut h2; w2->WriteCoderProperties(&h2);
          if (h2.d[2] != c.want) { fprintf(stderr, "FAIL prop %u = %u: level byte %u, want %u\n", (unsigned)c.id, (unsigned)c.val, (unsigned)h2.d[2], (unsigned)c.want); return 1; }
          s2->Release(); w2->Release(); CHECK(e2->Release() == 0); } }
    if (std::string(argv[2]) == "--exports") { printf("exports ok\n"); return 0; }

    std::vector<Byte> input;
    { FILE* f = fopen(argv[2], "rb"); CHECK(f); Byte buf[1 << 16]; size_t k; while ((k = fread(buf, 1, sizeof(buf), f)) > 0) input.insert(input.end(), buf, buf + k); fclose(f); }
    MemIn in(input); MemOut packed; Progress prog;
    HRESULT r = enc->Code(&in, &packed, nullptr, nullptr, &prog);
    if (r != S_OK) { fprintf(stderr, "encoder Code() = 0x%08x\n", (unsigned)r); return 1; }
    CHECK(prog.calls >= 1);
    CHECK(CreateDecoder(0, &iidCoder, &obj) == S_OK);
    ICompressCoder* dec = (ICompressCoder*)obj; ICompressSetDecoderProperties2* dp = nullptr;
    CHECK(dec->QueryInterface(b2z_iid(4, kIID_SetDecProps2), (void**)&dp) == S_OK);
    CHECK(dp->SetDecoderProperties2(hdr.d.data(), 5) == S_OK && dp->SetDecoderProperties2(hdr.d.data(), 2) == E_NOTIMPL);
    MemIn pin(packed.d); MemOut back; UInt64 outSize = input.size();
    r = dec->Code(&pin, &back, nullptr, &outSize, nullptr);
    if (r != S_OK) { fprintf(stderr, "decoder Code() = 0x%08x\n", (unsigned)r); return 1; }
    CHECK(back.d == input);
    { std::vector<Byte> pulled; int pr = pull_decode(dec, packed.d, outSize, pulled); if (pr) { fprintf(stderr, "zstd pull mode failed at step %d\n", pr); return 1; } CHECK(pulled == input); }
    // corrupt stream -> S_FALSE (data error)
    { std::vector<Byte> bad(packed.d.begin(), packed.d.begin() + packed.d.size() / 2); MemIn bi(bad); MemOut bo; CHECK(dec->Code(&bi, &bo, nullptr, &outSize, nullptr) == S_FALSE); }
    { FILE* f = fopen(argv[3], "wb"); CHECK(f); fwrite(packed.d.data(), 1, packed.d.size(), f); fclose(f); }
    sp->Release(); wp->Release(); mt->Release(); dp->Release();
    CHECK(enc->Release() == 0 && dec->Release() == 0);
    printf("roundtrip ok: %zu -> %zu bytes\n", input.size(), packed.d.size());
    return 0;
}
