"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/rtrb_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "raytracing_rb_b200", "csrc"), "-j4", "-s", "all"])
    from raytracing_rb_b200 import _lib
    return _lib


def header_functions():
    text = open(os.path.join(ROOT, "include", "rtrb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtrb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(built):
    L = built.lib()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert sorted(built.EXPORTS) == names


def test_struct_sizes_match_header(built, tmp_path):
    from raytracing_rb_b200 import _abi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rtrb_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(rtrb_object_desc),sizeof(rtrb_light_desc),sizeof(rtrb_texture_desc),sizeof(rtrb_scene_desc),'
                   'sizeof(rtrb_camera_desc),sizeof(rtrb_render_opts),sizeof(rtrb_stats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (_abi.ObjectDesc, _abi.LightDesc, _abi.TextureDesc, _abi.SceneDesc, _abi.CameraDesc,
                                  _abi.RenderOpts, _abi.Stats)]
    assert got == want


def test_no_cpu_fallback(built):
    """Without a CUDA device every computing entry point must fail with RTRB_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from raytracing_rb_b200 import Renderer, World, scenes
    from raytracing_rb_b200._lib import RtrbError
    w, _ = scenes.build(1)
    with pytest.raises(RtrbError) as ei:
        Renderer(World(w).to_scene_desc(), 0)
    assert ei.value.code == 2


def test_product_never_imports_oracle():
    """The product path must not import, include, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "raytracing_rb_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|(liboracle)|(oracle/)|(rtrb_oracle_)",
                     re.M)
    n = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")) or f == "Makefile":
                n += 1
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(text), os.path.join(dirpath, f)
    assert n >= 10
