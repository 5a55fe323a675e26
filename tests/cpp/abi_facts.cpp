// abi_facts.cpp -- prints the ABI facts a 7-Zip codec module depends on, one per line, from EITHER declaration of the ABI:
//   g++ abi_facts.cpp                                   -> this repo's codec/b2z_7zip_abi.h (what libb200z_7z.so is compiled against)
//   g++ -DB2Z_REFERENCE_HEADERS -I/root/reference/CPP   -> the reference's own MyWindows.h / MyCom.h / ICoder.h / IStream.h
// tests/test_boundary.py compiles both and requires identical output: constants, struct layout, interface IDs, property ids and
// -- through the slot each method occupies in a probe object's vtable -- the method order of every interface the module implements.
#include <cstdio>
#include <cstddef>
#include <cstring>
#ifdef B2Z_REFERENCE_HEADERS
#include "Common/MyInitGuid.h"
#include "Common/MyWindows.h"
#include "Common/MyCom.h"
#include "7zip/ICoder.h"
#include "7zip/IStream.h"
#define IID_OF(name, group, sub) IID_##name
#define UH(v) (v).uhVal.QuadPart
static const unsigned kUltimate = Z7_ZSTD_ULTIMATE_LEV, kFastInc = Z7_ZSTD_FAST_LEV_INC;
static const GUID& IID_IUnknown_value() { return IID_IUnknown; }
#else
#include "../../7-zip-zstd_b200/codec/b2z_7zip_abi.h"
#define IID_OF(name, group, sub) b2z_iid(group, sub)
#define UH(v) (v).uhVal
static const unsigned kUltimate = Z7_ZSTD_ULTIMATE_LEV, kFastInc = Z7_ZSTD_FAST_LEV_INC;     // what codec/ZstdCoders.cpp uses
typedef uint32_t ULONG;
static const GUID& IID_IUnknown_value() { return kIID_IUnknown; }
#endif

static void guid(const char* name, const GUID& g) {
    printf("iid %s %08x-%04x-%04x-", name, (unsigned)g.Data1, (unsigned)g.Data2, (unsigned)g.Data3);
    for (int i = 0; i < 8; i++) printf("%02x", (unsigned)g.Data4[i]);
    printf("\n");
}

// vtable slot of a virtual method = byte offset stored in its member-function pointer (Itanium C++ ABI: ptr = 1 + offset)
template <class F> static long slot(F f) { long v[2] = {0, 0}; memcpy(v, &f, sizeof(f) < sizeof(v) ? sizeof(f) : sizeof(v)); return (v[0] - 1) / (long)sizeof(void*); }
#define SLOT(iface, method) printf("slot " #iface "::" #method " %ld\n", slot(&iface::method))

int main() {
    printf("sizeof HRESULT %zu PROPID %zu GUID %zu PROPVARIANT %zu wchar_t %zu\n", sizeof(HRESULT), sizeof(PROPID), sizeof(GUID), sizeof(PROPVARIANT), sizeof(wchar_t));
    printf("PROPVARIANT vt@%zu ulVal@%zu uhVal@%zu boolVal@%zu bstrVal@%zu\n", offsetof(PROPVARIANT, vt), offsetof(PROPVARIANT, ulVal), offsetof(PROPVARIANT, uhVal), offsetof(PROPVARIANT, boolVal), offsetof(PROPVARIANT, bstrVal));
    printf("VT_EMPTY %d VT_BSTR %d VT_BOOL %d VT_UI4 %d VT_UI8 %d\n", (int)VT_EMPTY, (int)VT_BSTR, (int)VT_BOOL, (int)VT_UI4, (int)VT_UI8);
    printf("S_OK %08x S_FALSE %08x E_NOTIMPL %08x E_NOINTERFACE %08x E_ABORT %08x E_FAIL %08x E_OUTOFMEMORY %08x E_INVALIDARG %08x CLASS_E_CLASSNOTAVAILABLE %08x\n",
           (unsigned)S_OK, (unsigned)S_FALSE, (unsigned)E_NOTIMPL, (unsigned)E_NOINTERFACE, (unsigned)E_ABORT, (unsigned)E_FAIL, (unsigned)E_OUTOFMEMORY, (unsigned)E_INVALIDARG, (unsigned)CLASS_E_CLASSNOTAVAILABLE);
    printf("zstd level bytes: ultimate %u fast_inc %u\n", kUltimate, kFastInc);
    guid("IUnknown", IID_IUnknown_value());
    guid("ISequentialInStream", IID_OF(ISequentialInStream, 3, 0x01)); guid("ISequentialOutStream", IID_OF(ISequentialOutStream, 3, 0x02));
    guid("ICompressProgressInfo", IID_OF(ICompressProgressInfo, 4, 0x04)); guid("ICompressCoder", IID_OF(ICompressCoder, 4, 0x05));
    guid("ICompressSetCoderPropertiesOpt", IID_OF(ICompressSetCoderPropertiesOpt, 4, 0x1F)); guid("ICompressSetCoderProperties", IID_OF(ICompressSetCoderProperties, 4, 0x20));
    guid("ICompressSetDecoderProperties2", IID_OF(ICompressSetDecoderProperties2, 4, 0x22)); guid("ICompressWriteCoderProperties", IID_OF(ICompressWriteCoderProperties, 4, 0x23));
    guid("ICompressGetInStreamProcessedSize", IID_OF(ICompressGetInStreamProcessedSize, 4, 0x24)); guid("ICompressSetCoderMt", IID_OF(ICompressSetCoderMt, 4, 0x25));
    guid("ICompressSetFinishMode", IID_OF(ICompressSetFinishMode, 4, 0x26)); guid("ICompressSetMemLimit", IID_OF(ICompressSetMemLimit, 4, 0x28));
    guid("ICompressSetInStream", IID_OF(ICompressSetInStream, 4, 0x31)); guid("ICompressSetOutStreamSize", IID_OF(ICompressSetOutStreamSize, 4, 0x34));
    guid("ICompressSetBufSize", IID_OF(ICompressSetBufSize, 4, 0x35));
    printf("NCoderPropID");
    { using namespace NCoderPropID;
      const int v[] = { kDefaultProp, kDictionarySize, kUsedMemorySize, kOrder, kBlockSize, kPosStateBits, kLitContextBits, kLitPosBits, kNumFastBytes, kMatchFinder,
                        kMatchFinderCycles, kNumPasses, kAlgorithm, kNumThreads, kEndMarker, kLevel, kReduceSize, kExpectedDataSize, kBlockSize2, kCheckSize, kFilter,
                        kMemUse, kAffinity, kBranchOffset, kHashBits, kNumThreadGroups, kThreadGroup, kAffinityInGroup, kStrategy, kFast, kLong, kWindowLog, kHashLog,
                        kChainLog, kSearchLog, kMinMatch, kTargetLen, kOverlapLog, kLdmHashLog, kLdmSearchLength, kLdmBucketSizeLog, kLdmHashRateLog, kAdvMax };
      for (int x : v) printf(" %d", x); }
    printf("\nNMethodPropID");
    { using namespace NMethodPropID; const int v[] = { kID, kName, kDecoder, kEncoder, kPackStreams, kUnpackStreams, kDescription, kDecoderIsAssigned, kEncoderIsAssigned, kDigestSize, kIsFilter };
      for (int x : v) printf(" %d", x); }
    printf("\nNModulePropID %d %d\n", (int)NModulePropID::kInterfaceType, (int)NModulePropID::kVersion);
    SLOT(IUnknown, QueryInterface); SLOT(IUnknown, AddRef); SLOT(IUnknown, Release);
    SLOT(ISequentialInStream, Read); SLOT(ISequentialOutStream, Write); SLOT(ICompressProgressInfo, SetRatioInfo); SLOT(ICompressCoder, Code);
    SLOT(ICompressSetCoderPropertiesOpt, SetCoderPropertiesOpt); SLOT(ICompressSetCoderProperties, SetCoderProperties);
    SLOT(ICompressSetDecoderProperties2, SetDecoderProperties2); SLOT(ICompressWriteCoderProperties, WriteCoderProperties);
    SLOT(ICompressGetInStreamProcessedSize, GetInStreamProcessedSize); SLOT(ICompressSetCoderMt, SetNumberOfThreads); SLOT(ICompressSetFinishMode, SetFinishMode);
    SLOT(ICompressSetMemLimit, SetMemLimit); SLOT(ICompressSetInStream, SetInStream); SLOT(ICompressSetInStream, ReleaseInStream);
    SLOT(ICompressSetOutStreamSize, SetOutStreamSize); SLOT(ICompressSetBufSize, SetInBufSize); SLOT(ICompressSetBufSize, SetOutBufSize);
    return 0;
}
