"""INTEGRATION.md option 1, checked at link level: the reference's OWN CLI and block layer (bsc.cpp, libbsc.cpp) plus its host
stages (lzp, filters, adler32, platform) link against libbsc_b200.so for exactly the stage entry points the north star replaces
-- bsc_{bwt,st,coder}_{init,encode/compress,decode/decompress} -- and nothing else is missing.  Build container only (needs
/root/reference and g++); running the binary needs a GPU (without one it reports the library's GPU error code)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
LIBDIR = os.path.join(ROOT, "libbsc_b200")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "bsc.cpp")), reason="needs /root/reference")
def test_reference_cli_links_against_the_stage_abi(tmp_path):
    assert os.path.exists(os.path.join(LIBDIR, "libbsc_b200.so")), "run __graft_entry__.build() first"
    exe = str(tmp_path / "bsc_dropin")
    srcs = ["bsc.cpp", "libbsc/libbsc/libbsc.cpp", "libbsc/lzp/lzp.cpp", "libbsc/platform/platform.cpp", "libbsc/adler32/adler32.cpp",
            "libbsc/filters/detectors.cpp", "libbsc/filters/preprocessing.cpp"]
    cmd = ["/usr/bin/g++", "-O1", "-fopenmp", "-mavx2", "-w", "-DLIBBSC_OPENMP_SUPPORT", "-DLIBBSC_SORT_TRANSFORM_SUPPORT", "-DLIBBSC_ALLOW_UNALIGNED_ACCESS",
           "-I" + REF] + [os.path.join(REF, s) for s in srcs] + ["-o", exe, "-L" + LIBDIR, "-lbsc_b200", "-Wl,-rpath," + LIBDIR]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    undefined = {l.split()[-1] for l in subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout.splitlines() if " U bsc" in l}
    assert undefined == {"bsc_bwt_init", "bsc_bwt_encode", "bsc_bwt_decode", "bsc_st_init", "bsc_st_encode", "bsc_st_decode",
                         "bsc_coder_init", "bsc_coder_compress", "bsc_coder_decompress"}
    import ctypes
    lib = ctypes.CDLL(os.path.join(LIBDIR, "libbsc_b200.so"))
    for name in undefined:
        assert hasattr(lib, name), name
