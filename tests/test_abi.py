"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from strom_b200.beagle import ENGINE_LIB, SIGNATURES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("strom_beagle.h", "strom_b200.h"):
        text = "".join(l for l in open(os.path.join(ROOT, "include", h)) if not l.lstrip().startswith("#"))
        names += re.findall(r"STROM_API[^;(]*?\b(\w+)\s*\(", text)
    return names


def test_headers_declare_the_13_beagle_calls():
    names = declared_symbols()
    assert set(SIGNATURES) <= set(names)
    assert len([n for n in names if n.startswith("beagle")]) == 13


def test_engine_library_exports_every_declared_symbol():
    assert os.path.exists(ENGINE_LIB), "build the engine first: python -m strom_b200.build"
    lib = ctypes.CDLL(ENGINE_LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_oracle_exports_the_same_13(oracle):
    for name in SIGNATURES:
        assert hasattr(oracle.lib, name)


def test_drop_in_host_library_binds_only_the_13_calls():
    # libstrom_host_13call.so = the host classes compiled as the reference has them; everything it needs from the engine
    # must be one of the 13 BeagleLib calls (no engine-level entry point), and it loads without a GPU
    import subprocess
    from strom_b200.build import build_host_13call
    path = build_host_13call()
    ctypes.CDLL(path)
    out = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True).stdout
    undefined = {l.split()[-1].split("@")[0] for l in out.splitlines() if l.strip()}
    wanted = {n for n in undefined if n.startswith(("beagle", "sb200"))}
    assert wanted and wanted <= set(SIGNATURES), wanted


def test_version_string():
    lib = ctypes.CDLL(ENGINE_LIB)
    lib.sb200Version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sb200Version()


def test_product_sources_never_reference_the_oracle():
    # the product path must not import, link or execute anything under oracle/
    pkg = os.path.join(ROOT, "strom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and f != "build.py":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
