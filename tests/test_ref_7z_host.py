"""The codec module inside the reference's own host: the unmodified 7-Zip console program + 7z.so built from /root/reference by
oracle/build_ref_7z.sh with ONLY the ZSTD / LZMA2 / FLZMA2 registrations left out, and libb200z_7z.so dropped into Codecs/.
What CPP/7zip/UI/Common/LoadCodecs.cpp:528-563 (GetModuleProp version / interface-type check), :279-303 (GetNumberOfMethods,
GetMethodProperty, CreateEncoder / CreateDecoder) and Common/CreateCoder.cpp:159-232 (lookup by name / by id) do with the module
is then the reference's code, compiled against the reference's ICoder.h / MyCom.h -- not this repo's re-declaration of the ABI.

  * CPU (`-m "not gpu"`): `7z i` accepts the module and lists 4F71101 ZSTD, 21 LZMA2, 21 FLZMA2 from it
    (the restatement of tests/main.test:17-29 for these methods).
  * GPU (`-m gpu`): archives written through the module verify and extract with the STOCK reference 7zz, and archives written by
    the stock 7zz extract through the module (tests/main.test:66-92).
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF7Z = os.path.join(ROOT, "oracle", "_ref", "7z")
HOST, STOCK = os.path.join(REF7Z, "host", "7z"), os.path.join(REF7Z, "stock", "7zz")
PKG = os.path.join(ROOT, "7-zip-zstd_b200")


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref_7z.sh")])        # no-op when built; needs /root/reference otherwise
    if not (os.path.exists(HOST) and os.path.exists(STOCK)):
        pytest.skip("oracle/_ref/7z not built (no /root/reference here)")
    codecs = os.path.join(REF7Z, "host", "Codecs")
    os.makedirs(codecs, exist_ok=True)
    shutil.copy(os.path.join(PKG, "libb200z_7z.so"), os.path.join(codecs, "b200z.so"))       # the module under test, as built now
    env = dict(os.environ, LD_LIBRARY_PATH=PKG + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))   # libb200z.so (the engine) sits beside the package

    def run(exe, *args, cwd=None, ok=True):
        p = subprocess.run([exe, *args], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        out = p.stdout.decode(errors="replace")
        if ok:
            assert p.returncode == 0, out[-3000:]
        return p.returncode, out
    return run


def test_reference_host_lists_the_modules_methods(host):
    _, out = host(HOST, "i")
    assert "Codec Load Error" not in out, out[:2000]
    libs = [l for l in out.splitlines() if "b200z.so" in l]
    assert libs and ": 26.01 :" in libs[0], out[:1500]                      # LoadCodecs.cpp:562-564: version / interface type accepted
    lib_no = libs[0].split(":")[0].strip()
    mine = [l.split() for l in out.splitlines() if l.strip().startswith(lib_no + " ") and ("ED" in l.split()[:2])]
    got = {(f[2], f[3]) for f in mine if len(f) >= 4}
    assert got == {("4F71101", "ZSTD"), ("21", "LZMA2"), ("21", "FLZMA2")}, mine
    # the host itself no longer registers them (they were left out of the object list), so lookups reach the module
    internal = [l for l in out.splitlines() if l.strip().startswith("0 ") and l.split()[-1] in ("ZSTD", "LZMA2", "FLZMA2")]
    assert not internal, internal


def _payload(pkg, tmp):
    import numpy as np
    data = pkg.corpus.g2(3 * (1 << 20) + 4567).tobytes() + bytes(70_000) + pkg.corpus.entropy_class(1, 100_000).tobytes() + pkg.corpus.entropy_class(3, 300_000).tobytes()
    path = os.path.join(tmp, "payload.bin")
    with open(path, "wb") as f:
        f.write(data)
    return path, data


@pytest.mark.gpu
@pytest.mark.parametrize("method,shown", [("-m0=zstd -mx3", "ZSTD:v1.5,l3"), ("-m0=zstd -mx19", "ZSTD:v1.5,l19"), ("-m0=zstd:max", "ZSTD:v1.5,max"),
                                            ("-m0=lzma2 -mx5", "LZMA2:"), ("-m0=flzma2 -mx5", "LZMA2:")])
def test_archives_written_through_the_module_verify_with_the_stock_reference(host, pkg, tmp_path, method, shown):
    path, data = _payload(pkg, str(tmp_path))
    arc = os.path.join(str(tmp_path), "a.7z")
    _, out = host(HOST, "a", *method.split(), arc, path, cwd=str(tmp_path))
    assert "Everything is Ok" in out, out[-1500:]
    _, out = host(STOCK, "t", arc)
    assert "Everything is Ok" in out, out[-1500:]
    _, out = host(STOCK, "l", "-slt", arc)
    assert any(l.startswith("Method = ") and shown in l for l in out.splitlines()), out[-1500:]
    outdir = os.path.join(str(tmp_path), "x"); os.makedirs(outdir)
    host(STOCK, "x", "-o" + outdir, arc)
    assert open(os.path.join(outdir, "payload.bin"), "rb").read() == data


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["-m0=zstd -mx3", "-m0=zstd -mx17", "-m0=lzma2 -mx5 -md=4m", "-m0=flzma2 -mx5", "-m0=zstd -mx1 -ms=off"])
def test_archives_of_the_stock_reference_extract_through_the_module(host, pkg, tmp_path, method):
    path, data = _payload(pkg, str(tmp_path))
    arc = os.path.join(str(tmp_path), "s.7z")
    host(STOCK, "a", *method.split(), arc, path, cwd=str(tmp_path))
    _, out = host(HOST, "t", arc)
    assert "Everything is Ok" in out, out[-1500:]
    outdir = os.path.join(str(tmp_path), "y"); os.makedirs(outdir)
    host(HOST, "x", "-o" + outdir, arc)
    assert open(os.path.join(outdir, "payload.bin"), "rb").read() == data
