import sys, os
sys.path.insert(0, "tests")
import helpers as H
import __graft_entry__ as ge
pkg = ge.load_package()
data = pkg.corpus.entropy_class(3, 5 << 20).tobytes() + bytes(4 << 20) + pkg.corpus.g2(3 << 20).tobytes()
for fl, sl in ((22, 2), (24, 0), (23, 3)):
    c = pkg.Codec(0, frame_log=fl, window_log=fl, lzma2_slice_log=sl)
    got = c.lzma2_compress(data)
    want = H.oracle_lzma2_compress(data, frameLog=fl, windowLog=fl, flags=1 | (sl << 8))
    back = c.lzma2_decompress(got[1], got[0])
    print(fl, sl, got == want, back == data, len(got[1]), flush=True)
    c.close()
