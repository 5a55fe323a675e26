// b2z_7zip_abi.h -- the slice of 7-Zip's codec-plugin ABI that a coder module must speak,
// declared from the published ABI facts (interface IDs, method order, PROPVARIANT layout), so
// that libb200z_7z.so can be built without the 7-Zip source tree.
//
// Binary compatibility contract (Linux/Itanium C++ ABI; reference files under /root/reference/CPP):
//   - IUnknown has NO virtual destructor (7zip/ICoder.h:420-437 NModuleInterfaceType: mode 0,
//     reported through GetModuleProp(kInterfaceType)); vtable = {QueryInterface, AddRef, Release, ...}
//     (Common/MyUnknown.h / MyWindows.h:172-180).
//   - interface IDs: {23170F69-40C1-278A-0000-00gg00ss0000}, gg = group (3 stream, 4 coder),
//     ss = sub id (7zip/IDecl.h:9-24; 7zip/IStream.h:14-70; 7zip/ICoder.h:10-275).
//   - method order inside each interface = order in the reference headers (cited per interface).
//   - HRESULT values: Common/MyWindows.h:94-103.  PROPVARIANT: Common/MyWindows.h:222-250
//     (vt + 3 pad words, then an 8-byte union).  BSTR = 4-byte length prefix + UTF-32 text
//     (Common/MyWindows.cpp SysAllocStringLen), allocated with malloc.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

typedef int32_t HRESULT;
typedef uint32_t UInt32;
typedef uint64_t UInt64;
typedef uint8_t Byte;
typedef uint32_t PROPID;          // ULONG is 4 bytes in 7-Zip's Linux build (MyWindows.h:90, MyTypes)
typedef wchar_t* BSTR;

#define S_OK            ((HRESULT)0x00000000L)
#define S_FALSE         ((HRESULT)0x00000001L)
#define E_NOTIMPL       ((HRESULT)0x80004001L)
#define E_NOINTERFACE   ((HRESULT)0x80004002L)
#define E_ABORT         ((HRESULT)0x80004004L)
#define E_FAIL          ((HRESULT)0x80004005L)
#define E_OUTOFMEMORY   ((HRESULT)0x8007000EL)
#define E_INVALIDARG    ((HRESULT)0x80070057L)
#define CLASS_E_CLASSNOTAVAILABLE ((HRESULT)0x80040111L)

struct GUID { uint32_t Data1; uint16_t Data2; uint16_t Data3; uint8_t Data4[8]; };
inline bool operator==(const GUID& a, const GUID& b) { return memcmp(&a, &b, sizeof(GUID)) == 0; }

enum { VT_EMPTY = 0, VT_BSTR = 8, VT_BOOL = 11, VT_UI4 = 19, VT_UI8 = 21 };
struct PROPVARIANT {
    uint16_t vt, r1, r2, r3;
    union { uint32_t ulVal; uint64_t uhVal; int16_t boolVal; BSTR bstrVal; };
};
static_assert(sizeof(PROPVARIANT) == 16, "PROPVARIANT layout");

inline GUID b2z_iid(uint8_t group, uint8_t sub) { return GUID{ 0x23170F69, 0x40C1, 0x278A, { 0, 0, 0, group, 0, sub, 0, 0 } }; }
inline GUID b2z_clsid(bool encoder, uint64_t id) {
    GUID g{ 0x23170F69, 0x40C1, (uint16_t)(encoder ? 0x2791 : 0x2790), { 0 } };
    for (int i = 0; i < 8; i++) g.Data4[i] = (uint8_t)(id >> (8 * i));      // CodecExports.cpp:62-70 (SetUi64 little-endian)
    return g;
}

struct IUnknown {
    virtual HRESULT QueryInterface(const GUID& iid, void** out) = 0;
    virtual UInt32 AddRef() = 0;
    virtual UInt32 Release() = 0;
};
// 7zip/IStream.h:47-70
struct ISequentialInStream : IUnknown { virtual HRESULT Read(void* data, UInt32 size, UInt32* processed) = 0; };
struct ISequentialOutStream : IUnknown { virtual HRESULT Write(const void* data, UInt32 size, UInt32* processed) = 0; };
// 7zip/ICoder.h:14-31
struct ICompressProgressInfo : IUnknown { virtual HRESULT SetRatioInfo(const UInt64* inSize, const UInt64* outSize) = 0; };
struct ICompressCoder : IUnknown {
    virtual HRESULT Code(ISequentialInStream* in, ISequentialOutStream* out, const UInt64* inSize, const UInt64* outSize,
                         ICompressProgressInfo* progress) = 0;
};
// 7zip/ICoder.h:172-208
struct ICompressSetCoderPropertiesOpt : IUnknown { virtual HRESULT SetCoderPropertiesOpt(const PROPID* ids, const PROPVARIANT* props, UInt32 n) = 0; };
struct ICompressSetCoderProperties : IUnknown { virtual HRESULT SetCoderProperties(const PROPID* ids, const PROPVARIANT* props, UInt32 n) = 0; };
struct ICompressSetDecoderProperties2 : IUnknown { virtual HRESULT SetDecoderProperties2(const Byte* data, UInt32 size) = 0; };
struct ICompressWriteCoderProperties : IUnknown { virtual HRESULT WriteCoderProperties(ISequentialOutStream* out) = 0; };
struct ICompressSetCoderMt : IUnknown { virtual HRESULT SetNumberOfThreads(UInt32 n) = 0; };
struct ICompressGetInStreamProcessedSize : IUnknown { virtual HRESULT GetInStreamProcessedSize(UInt64* value) = 0; };   // ICoder.h:202-204
struct ICompressSetMemLimit : IUnknown { virtual HRESULT SetMemLimit(UInt64 memUsage) = 0; };                             // ICoder.h:221-223
struct ICompressSetInStream : IUnknown { virtual HRESULT SetInStream(ISequentialInStream* in) = 0; virtual HRESULT ReleaseInStream() = 0; };   // :257-260
struct ICompressSetOutStreamSize : IUnknown { virtual HRESULT SetOutStreamSize(const UInt64* outSize) = 0; };            // :273-275
struct ICompressSetBufSize : IUnknown { virtual HRESULT SetInBufSize(UInt32 streamIndex, UInt32 size) = 0; virtual HRESULT SetOutBufSize(UInt32 streamIndex, UInt32 size) = 0; };   // :281-285
struct ICompressSetFinishMode : IUnknown { virtual HRESULT SetFinishMode(UInt32 finishMode) = 0; };                      // ICoder.h:210-216

enum { kIID_SeqIn = 0x01, kIID_SeqOut = 0x02 };                                          // group 3
enum { kIID_Progress = 0x04, kIID_Coder = 0x05, kIID_SetPropsOpt = 0x1F, kIID_SetProps = 0x20,
       kIID_SetDecProps2 = 0x22, kIID_WriteProps = 0x23, kIID_GetInProcessed = 0x24, kIID_SetMt = 0x25, kIID_SetFinishMode = 0x26, kIID_SetMemLimit = 0x28,
       kIID_SetInStream = 0x31, kIID_SetOutStreamSize = 0x34, kIID_SetBufSize = 0x35 };             // group 4
static const GUID kIID_IUnknown = { 0, 0, 0, { 0xC0, 0, 0, 0, 0, 0, 0, 0x46 } };

// 7zip/ICoder.h:104-160 (NCoderPropID)
namespace NCoderPropID { enum {
    kDefaultProp = 0, kDictionarySize, kUsedMemorySize, kOrder, kBlockSize, kPosStateBits, kLitContextBits, kLitPosBits,
    kNumFastBytes, kMatchFinder, kMatchFinderCycles, kNumPasses, kAlgorithm, kNumThreads, kEndMarker, kLevel, kReduceSize,
    kExpectedDataSize, kBlockSize2, kCheckSize, kFilter, kMemUse, kAffinity, kBranchOffset, kHashBits, kNumThreadGroups,
    kThreadGroup, kAffinityInGroup,
    kStrategy, kFast, kLong, kWindowLog, kHashLog, kChainLog, kSearchLog, kMinMatch, kTargetLen, kOverlapLog,
    kLdmHashLog, kLdmSearchLength, kLdmBucketSizeLog, kLdmHashRateLog, kAdvMax }; }
// 7zip/ICoder.h:163-166: level bytes of the ZSTD coder's 5-byte property header
#define Z7_ZSTD_FAST_LEV_INC  32      /* 32 + f = fast level f */
#define Z7_ZSTD_ULTIMATE_LEV  255     /* "max" */
// 7zip/ICoder.h:405-419, 440-446
namespace NMethodPropID { enum { kID, kName, kDecoder, kEncoder, kPackStreams, kUnpackStreams, kDescription,
                                 kDecoderIsAssigned, kEncoderIsAssigned, kDigestSize, kIsFilter }; }
namespace NModulePropID { enum { kInterfaceType, kVersion }; }

inline BSTR b2z_alloc_bstr_bytes(const void* data, uint32_t len) {          // SysAllocStringByteLen layout
    const uint32_t size = (len + 2 * (uint32_t)sizeof(wchar_t) - 1) & ~((uint32_t)sizeof(wchar_t) - 1);
    uint8_t* p = (uint8_t*)malloc(size + 4);
    if (!p) return nullptr;
    memcpy(p, &len, 4); memset(p + 4, 0, size);
    if (data) memcpy(p + 4, data, len);
    return (BSTR)(p + 4);
}
inline BSTR b2z_alloc_bstr_ascii(const char* s) {
    const uint32_t n = (uint32_t)strlen(s), bytes = n * (uint32_t)sizeof(wchar_t);
    uint8_t* p = (uint8_t*)malloc(bytes + 4 + sizeof(wchar_t));
    if (!p) return nullptr;
    memcpy(p, &bytes, 4);
    wchar_t* w = (wchar_t*)(p + 4);
    for (uint32_t i = 0; i < n; i++) w[i] = (wchar_t)(unsigned char)s[i];
    w[n] = 0;
    return w;
}
