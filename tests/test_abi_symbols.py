"""CPU-side checks of the drop-in boundary: the shared library loads without a CUDA driver,
exports every symbol include/panopaea_b200.h declares, the ctypes binding covers all of them,
and with no device every compute path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "panopaea_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"PANO_API\s+[\w\s\*]+?\b(pano_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from panopaea_b200 import build as pb
    pb.build()
    from panopaea_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the .so"


def test_binding_covers_header():
    from panopaea_b200 import _lib
    assert _lib.exported_names() == declared_symbols()


def test_only_abi_symbols_are_exported():
    from panopaea_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_symbols()


def test_header_is_plain_c():
    src = '#include "panopaea_b200.h"\nint main(void){ pano_rect r = {0,0,0,0}; (void)r; return (int)sizeof(pano_step_params) == 0; }\n'
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-", "-fsyntax-only"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts_match_header():
    """ctypes mirrors vs the C compiler's view of the structs."""
    from panopaea_b200 import _lib
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "panopaea_b200.h"\n'
           'int main(void){ printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(pano_rect), sizeof(pano_pcg_info), sizeof(pano_step_params),'
           ' offsetof(pano_step_params, inflow), offsetof(pano_step_params, obstacle), offsetof(pano_pcg_info, final_residual)); return 0; }\n')
    exe = os.path.join(ROOT, "tests", "hostcheck", "layout_probe")
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    os.remove(exe)
    want = [C.sizeof(_lib.Rect), C.sizeof(_lib.PcgInfo), C.sizeof(_lib.StepParams), _lib.StepParams.inflow.offset,
            _lib.StepParams.obstacle.offset, _lib.PcgInfo.final_residual.offset]
    assert got == want


def test_no_cpu_fallback(lib):
    """Without a device the product refuses to run; it never routes through the oracle or numpy."""
    n = C.c_int(-1)
    assert lib.pano_device_count(C.byref(n)) == 0
    if n.value > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.pano_ctx_create(0, None, C.byref(h))
    assert rc == 3 and not h.value                     # PANO_ERR_CUDA
    assert b"no CPU fallback" in lib.pano_last_error()
    import panopaea_b200 as P
    with pytest.raises(P.PanoError):
        P.Context(0)
    # null handles are rejected, not dereferenced
    assert lib.pano_field_fill(None, 1.0) == 1
    assert lib.pano_fluid_step(None, None, None, None, None, None, None, None, None, None) == 1
    # the Grid3d family too
    assert lib.pano_fluid3_step(None, None, None, None, None, None, None, None, None, None) == 1
    assert lib.pano_field3_new(None, 3, 4, 4, 4, None) == 1
    assert lib.pano_trilinear(0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.25, 0.5, 0.75) == 0.25   # host arithmetic: no device needed


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "panopaea_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pano_oracle" not in text and "np_oracle" not in text and "oracle/" not in text.replace("the CPU oracle", ""), f


def _header_prototypes():
    """name -> number of parameters, from the header."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for name, args in re.findall(r"PANO_API\s+[\w\s\*]+?\b(pano_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_rust_bindings_match_header():
    """rust/panopaea-b200-sys cannot be compiled here (no Rust toolchain): at least keep its extern block in step with
    the header -- same symbols, same parameter counts, same constants."""
    src = open(os.path.join(ROOT, "rust", "panopaea-b200-sys", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"//.*", "", block)
    rust = {}
    for name, args in re.findall(r"pub fn (pano_\w+)\s*\(([^)]*)\)", block, flags=re.S):
        args = args.strip()
        rust[name] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    want = _header_prototypes()
    assert sorted(rust) == sorted(want)
    assert rust == want
    header = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    consts = dict(re.findall(r"\b(PANO_[A-Z0-9_]+)\s*=\s*(\d+)", header))
    consts.update(re.findall(r"#define\s+(PANO_[A-Z0-9_]+)\s+(\d+)", header))
    assert len(consts) >= 20
    for k, v in consts.items():
        m = re.search(r"pub const %s: \w+ = (\d+);" % k, src)
        assert m, f"{k} missing from the Rust bindings"
        assert m.group(1) == v, k
