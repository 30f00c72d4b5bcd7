"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports every symbol include/ckd.h declares,
the header is valid C with the same struct layouts the Python binding uses, and the product fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

from conftest import HAVE_GPU, REPO

HEADER = os.path.join(REPO, "include", "ckd.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ckd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cookiedough_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_functions()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"include/ckd.h declares symbols the library does not export: {missing}"


def test_header_is_plain_c_and_struct_layouts_match_binding():
    from cookiedough_b200 import capi
    structs = {
        "ckd_plasma_params": capi.PlasmaParams, "ckd_nautilus_params": capi.NautilusParams, "ckd_spikey_params": capi.SpikeyParams,
        "ckd_tunnel_params": capi.TunnelParams, "ckd_sinuses_params": capi.SinusesParams, "ckd_laura_params": capi.LauraParams,
        "ckd_landscape_params": capi.LandscapeParams, "ckd_tunnelscape_params": capi.TunnelscapeParams,
        "ckd_ball_params": capi.BallParams, "ckd_twister_params": capi.TwisterParams,
    }
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));\n' for n in structs)
    src = f'#include <stdio.h>\n#include "ckd.h"\nint main(void) {{ {body} printf("images %d\\n", (int)CKD_IMG_COUNT); return 0; }}\n'
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "t.c")
        exe = os.path.join(tmp, "t")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    sizes = dict(line.split() for line in out.strip().splitlines())
    for name, cls in structs.items():
        assert int(sizes[name]) == ctypes.sizeof(cls), name
    assert int(sizes["images"]) == 27
    slots = []
    for v in capi.IMAGE_SLOTS.values():
        slots += list(v) if isinstance(v, tuple) else [v]
    assert sorted(slots) == list(range(27))


def test_struct_fields_map_to_rocket_tracks():
    from cookiedough_b200 import capi
    for effect, (cls, names) in capi.TRACKS.items():
        fields = {n for n, _ in cls._fields_}
        assert set(names) <= fields, effect


@pytest.mark.skipif(HAVE_GPU, reason="this check is for GPU-less hosts")
def test_no_cpu_fallback_without_gpu():
    from cookiedough_b200 import capi
    with pytest.raises(capi.CkdError) as err:
        capi.Context(1280, 720)
    assert "no CPU fallback" in str(err.value) or "CUDA" in str(err.value)


def test_geti_matches_roundf():
    from cookiedough_b200.capi import geti
    assert [geti(v) for v in (0.4, 0.5, 1.5, 2.5, -0.5, -1.5, 511.7)] == [0, 1, 2, 3, -1, -2, 512]


def test_host_layer_exports_the_reference_entry_points():
    """include/ckd_host.h: the reference's own C++ names (demo.h, shadertoy.h, ..., util.h, rocket.h) plus the CkdHost_* /
    CkdSink_* services must be exported with C++ linkage exactly as a caller compiled against the header expects them"""
    from cookiedough_b200 import capi
    header = open(os.path.join(REPO, "include", "ckd_host.h")).read()
    header = re.sub(r"//.*", "", header)
    declared = set(re.findall(r"^\s*(?:bool|void|float|double|int|uint32_t \*|ckd_ctx \*|SyncTrack|const std::string &)\s*\*?([A-Za-z_][A-Za-z0-9_]*)\s*\(", header, flags=re.M))
    declared -= {"getf"}  # inline in the header
    assert {"Demo_Create", "Demo_Draw", "Demo_Destroy", "Nautilus_Draw", "Landscape_Create", "Polar_BlitA", "BoxBlur32", "Mix32", "TapeWarp32",
            "CkdHost_Create", "CkdSink_Open", "CkdSink_Commit", "Launch", "AddTrack", "SetLastError"} <= declared
    syms = subprocess.check_output(["nm", "-D", "--defined-only", "-C", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (?:Rocket::)?([A-Za-z_][A-Za-z0-9_]*)(?:\[abi:cxx11\])?\(", syms))
    missing = sorted(declared - exported)
    assert not missing, f"include/ckd_host.h declares functions the library does not export: {missing}"


def test_host_layer_covers_the_survey_signature_list():
    """SURVEY 8(b) 'signatures to preserve': every function and global of that list is exported by the library"""
    from cookiedough_b200 import capi
    functions = """Shadertoy_Create Shadertoy_Destroy Nautilus_Draw Sinuses_Draw Laura_Draw Plasma_Draw Tunnel_Draw Spikey_Draw
        Landscape_Create Landscape_Destroy Landscape_Draw Tunnelscape_Create Tunnelscape_Destroy Tunnelscape_Draw
        Twister_Create Twister_Destroy Twister_Draw Ball_Create Ball_Destroy Ball_Draw Ball_GetBackground Ball_HasBeams
        Polar_Create Polar_Destroy Polar_Blit Polar_BlitA Polar_Blit_2x2 BoxBlur_Create BoxBlur_Destroy BoxBlur_Horz32 BoxBlur_Vert32 BoxBlur_32
        HorizontalBoxBlur32 VerticalBoxBlur32 BoxBlur32 BoxBlurScale FxBlitter_Create FxBlitter_Destroy Fx_Blit_2x2 FxBlitter_DrawTestPattern
        Shared_Create Shared_Destroy memset32 Mix32 MixOver32 Add32 Sub32 Excl32 SoftLight32 SoftLight32A SoftLight32AA TapeWarp32 Overlay32
        Overlay32A Darken32_50 MulSrc32 MulSrc32A MixSrc32 MixSrc32S BlitSrc32 BlitSrc32A BlitAdd32 BlitAdd32A Fade32
        Demo_Create Demo_Destroy Demo_Draw""".split()
    syms = subprocess.check_output(["nm", "-D", "--defined-only", "-C", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T ([A-Za-z_][A-Za-z0-9_]*)\(", syms))
    assert not sorted(set(functions) - exported)
    data = set(re.findall(r" [BD] ([A-Za-z_][A-Za-z0-9_]*)$", syms, flags=re.M))
    assert {"g_pFxMap", "g_renderTarget", "g_gradientUnp16", "g_pNytrikTPB", "g_pXboxLogoTPB"} <= data


def test_host_hooks_for_bindings_are_exported():
    from cookiedough_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in ("ckdhost_create", "ckdhost_launch", "ckdhost_draw", "ckdhost_post", "ckdhost_demo_create", "ckdhost_demo_draw", "ckdhost_demo_destroy",
                 "ckdhost_set_pipelined", "ckdhost_flush", "ckdhost_module", "ckdhost_global", "ckdsink_open", "ckdsink_acquire", "ckdsink_commit", "ckdsink_close"):
        assert hasattr(lib, name), name
