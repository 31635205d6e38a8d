"""The C-ABI library loads and exports every symbol include/b200gs.h declares (no compute calls
without a GPU), and the host-side operator mirror reproduces the public interface's argument errors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200gs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200gs_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from robosimgs_b200 import _cabi
    L = ctypes.CDLL(_cabi.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 7
    for s in syms:
        assert hasattr(L, s), s
    assert set(syms) == set(_cabi.EXPORTED_SYMBOLS)
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'b200gs.h')).read()
    assert L.b200gs_version() == int(re.search(r'#define B200GS_VERSION (\d+)', hdr).group(1)) >= 101


def test_buffer_sizes_are_host_only_and_monotone(built):
    from robosimgs_b200 import _cabi
    L = _cabi.lib()
    g, b, i = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    assert L.b200gs_buffer_sizes(1000, 1080, 1920, 50000, ctypes.byref(g), ctypes.byref(b), ctypes.byref(i)) == 0
    assert g.value >= 1000 * (48 + 4 * 4 + 1) and b.value >= 50000 * 24 and i.value >= 1080 * 1920 * 20
    g2 = ctypes.c_size_t()
    L.b200gs_buffer_sizes(2000, 1080, 1920, 0, ctypes.byref(g2), None, None)
    assert g2.value > g.value
    assert L.b200gs_buffer_sizes(-1, 10, 10, 0, None, None, None) != 0
    assert b"invalid" in L.b200gs_last_error()


def test_operator_surface_names_and_argument_errors(built):
    import diff_gaussian_rasterization as dgr
    import robosimgs_b200 as rb
    assert dgr.GaussianRasterizer is rb.GaussianRasterizer
    assert rb.RasterizationSettings is rb.GaussianRasterizationSettings
    fields = rb.GaussianRasterizationSettings._fields
    assert fields[:12] == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                           "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    from helpers import small_scene
    sc, cam, rs = small_scene(P=8, degree=0)
    r = rb.GaussianRasterizer(rs)
    m2 = torch.zeros_like(sc.means3D)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(sc.means3D, m2, sc.opacities, scales=sc.scales, rotations=sc.rotations)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(sc.means3D, m2, sc.opacities, shs=sc.shs, colors_precomp=sc.shs[:, 0], scales=sc.scales,
          rotations=sc.rotations)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(sc.means3D, m2, sc.opacities, shs=sc.shs)
    # no CPU fallback: CPU tensors are refused loudly
    with pytest.raises(rb.B200GSError, match="no CPU fallback"):
        r(sc.means3D, m2, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU arms touch it --
    never the product package, the alias package, the dev tools or the GPU yardstick."""
    for top in ("robosimgs_b200", "diff_gaussian_rasterization", "tools", "baseline"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "gs_oracle.h" not in txt, (top, f)


def test_header_is_plain_c_and_a_c_host_links_and_runs(built, tmp_path):
    """The drop-in boundary is a C ABI: include/b200gs.h must compile as strict C99 (no C++-isms), a C host must link
    against libb200gs.so taking the address of every declared entry point, and the host-only calls must work from C
    (version, policy rules, buffer sizes, error string) -- no device needed."""
    import shutil
    import subprocess
    from robosimgs_b200 import _cabi
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    syms = _declared_symbols()
    table = ",\n    ".join(f'{{"{s}", (void (*)(void))&{s}}}' for s in syms)
    src = tmp_path / "c_host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "b200gs.h"
struct entry { const char* name; void (*fn)(void); };
static const struct entry table[] = {
    %s
};
int main(void) {
  size_t n = sizeof(table) / sizeof(table[0]), i, g = 0, b = 0, im = 0;
  B200GSParams prm;
  B200GSAlloc a;
  memset(&prm, 0, sizeof prm);
  memset(&a, 0, sizeof a);
  for (i = 0; i < n; i++) if (!table[i].fn) return 2;
  if (b200gs_version() != B200GS_VERSION) return 3;
  if (b200gs_policy_pair_capacity(0) != 0 || b200gs_policy_pair_capacity(160000) != 160000 + 10000 + 32768) return 4;
  if (b200gs_buffer_sizes(1000, 1080, 1920, 50000, &g, &b, &im) != 0 || g == 0 || b == 0 || im == 0) return 5;
  if (b200gs_buffer_sizes(-1, 10, 10, 0, NULL, NULL, NULL) == 0 || !strstr(b200gs_last_error(), "invalid")) return 6;
  if ((B200GS_DEFER_PAIR_CHECK | B200GS_FORWARD_ONLY | B200GS_OUT_RGB8) != 7 || B200GS_BIN_SHIFT_HINT(3) != 0x400) return 7;
  printf("%%u entry points, params %%u bytes\n", (unsigned)n, (unsigned)sizeof prm);
  return 0;
}
''' % table)
    exe = tmp_path / "c_host"
    libdir = os.path.dirname(_cabi.LIB_PATH)
    cmd = [cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", str(exe), "-L", libdir, "-l:libb200gs.so", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith(f"{len(syms)} entry points") and f"{ctypes.sizeof(_cabi.B200GSParams)} bytes" in r.stdout
