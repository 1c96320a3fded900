"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/oceanwaves.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present. No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import fft_ocean_waves_b200 as fow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "oceanwaves.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ow_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_documented_surface():
    syms = header_symbols()
    assert sorted(fow.EXPORTED_SYMBOLS) == syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(fow.lib_path()), "build with: python fft-ocean-waves_b200/build.py"
    lib = C.CDLL(fow.lib_path())
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "oceanwaves.h"\nint main(void){ ow_params p; ow_outputs o; (void)p; (void)o; return OW_OK; }\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_library_contains_sm100a_code():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", fow.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = fow.load_library()
    h = C.c_void_p()
    p = fow.OceanParams().to_c()
    assert lib.ow_create(300, 1, 1, C.byref(p), 0, 0, C.byref(h)) == 1          # N not supported -> OW_ERR_INVALID
    assert b"N must be" in lib.ow_last_error(None)
    assert lib.ow_create(256, 2, 1, C.byref(p), 0, 0, C.byref(h)) == 1          # n_slots < n_cascades
    assert lib.ow_step(None, 0.0, None) == 1
    # the composition / clock entry points reject a NULL context the same way
    assert lib.ow_sample_points(None, 0, None, 1.0, 0, None, None, None) == 1
    assert lib.ow_sample_points_host(None, 0, None, 1.0, 0, None, None, None) == 1
    assert lib.ow_compose_grid(None, 0, None, 1.0, 16, 0.0, 0.0, 1.0, None, None, None) == 1
    assert lib.ow_set_time_scale(None, 1.0, 0.0) == 1 and lib.ow_step_wall_clock(None, 0.0, None) == 1
    assert lib.ow_slab_set_column_lines(None, 0) == 1


def test_no_cpu_fallback():
    """Without a CUDA device creating a simulation must raise; with one it must succeed."""
    import torch
    if torch.cuda.is_available():
        fow.FFTOceanWaves(N=256).close()
    else:
        with pytest.raises(fow.OceanWavesError):
            fow.FFTOceanWaves(N=256)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "fft-ocean-waves_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in ("ow_kernels.cuh",) and "import" not in txt, f


def test_cpp_wrapper_compiles_and_links_and_fails_loudly_without_gpu(tmp_path):
    """include/oceanwaves.hpp (the C++ host mirror of FFTOceanWaves' sim methods) builds against the library; on a
    GPU-less box create() returns false with an error text instead of computing anything."""
    import torch
    src = tmp_path / "t.cpp"
    src.write_text(
        '#include <cstdio>\n#include "oceanwaves.hpp"\n'
        "int main(){ ow::OceanSim s; ow::OceanSim::Params p; p.wind_speed = 40.f;\n"
        "  if(!s.create(256, p)){ std::printf(\"ERR %s\\n\", s.last_error().c_str()); return 3; }\n"
        "  s.generate_bit_reversed_indices(); s.generate_twiddle_factors();\n"
        "  if(s.update(0.f)) return 4;   /* ow_step before tilde_h0_k must be refused */\n"
        "  std::printf(\"OK %d\\n\", s.N()); return 0; }\n")
    exe = tmp_path / "t"
    libdir = os.path.dirname(fow.lib_path())
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-loceanwaves", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and r.stdout.startswith("OK 256"), r.stdout + r.stderr
    else:
        assert r.returncode == 3 and r.stdout.startswith("ERR "), r.stdout + r.stderr
