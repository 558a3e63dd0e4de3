"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/fflins.h declares, the ctypes structs match the C layout, the C++ shim compiles against
it, and there is no CPU fallback behind it."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "fflins.h")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fflins_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fflins import api
    lib = api.load()
    names = declared_functions()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/fflins.h but not exported: {missing}"
    assert sorted(api.EXPORTS) == names, "api.EXPORTS must list exactly the header's functions"


def test_every_entry_point_cites_the_reference():
    """Each block of the header names the reference interface (file:line) it replaces."""
    src = open(HDR).read()
    for ref in ("lidar_map.cc:213-273", "lidar_map.cc:190-211", "lidar_map.cc:100-126", "lidar_map.cc:326-408",
                "lidar_factor.cc:41-118", "marginalization_info.cc:181-210", "optimizer.cc:515-548",
                "pointcloud.cc:38-51", "pointcloud.cc:61-93", "ikd_Tree.cc:360-391"):
        assert ref in src, ref


def test_struct_layouts_match_the_header(tmp_path):
    from fflins import api
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fflins.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(fflins_config),sizeof(fflins_pose),sizeof(fflins_frame_info),sizeof(fflins_assoc_stats),'
                    'offsetof(fflins_frame_info,pose),offsetof(fflins_config,max_points_per_frame));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc" if not os.path.exists("/usr/bin/gcc") else "/usr/bin/gcc", "-I", os.path.join(ROOT, "include"),
                           str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(api.Config), C.sizeof(api.Pose), C.sizeof(api.FrameInfo), C.sizeof(api.AssocStats),
                     api.FrameInfo.pose.offset, api.Config.max_points_per_frame.offset]


def test_solve_struct_layouts_match_the_header(tmp_path):
    from fflins import api
    prog = tmp_path / "sizes2.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fflins.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(fflins_solve_options),sizeof(fflins_solve_prior),sizeof(fflins_solve_report),'
                    'offsetof(fflins_solve_options,flags),offsetof(fflins_solve_report,cost_initial),'
                    'offsetof(fflins_solve_report,n_res));return 0;}\n')
    exe = tmp_path / "sizes2"
    subprocess.check_call(["gcc" if not os.path.exists("/usr/bin/gcc") else "/usr/bin/gcc", "-I", os.path.join(ROOT, "include"),
                           str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(api.SolveOptions), C.sizeof(api.SolvePrior), C.sizeof(api.SolveReport), api.SolveOptions.flags.offset,
                     api.SolveReport.cost_initial.offset, api.SolveReport.n_res.offset]
    o = api.SolveOptions()
    api.load().fflins_solve_options_default(C.byref(o))
    assert (o.max_iterations[0], o.max_iterations[1], o.chi2_threshold, o.flags) == (5, 10, 3.841, 1)   # optimizer.cc:87-88, 516


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c89.c"
    prog.write_text('#include "fflins.h"\nint main(void){fflins_config c; fflins_config_default(&c); return c.window_size == 10 ? 0 : 1;}\n')
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c99", "-pedantic", "-Wall", "-Werror",
                           "-I", os.path.join(ROOT, "include"), "-c", str(prog), "-o", str(tmp_path / "c89.o")])


def test_default_config_values():
    """fflins_config_default = config/ff_lins_robot.yaml + the hard-coded constants (lidar_map.h:87-94)."""
    from fflins import api
    c = api.default_config()
    assert (c.point_to_plane_std, c.window_size, c.point_filter_num, c.downsample_size) == (0.1, 10, 10, 0.5)
    assert (c.plane_estimation_threshold, c.minimum_keyframe_translation) == (0.1, 0.4)
    assert (c.max_search_square_distance, c.plane_scalar_threshold, c.plane_sigma_gate, c.huber_delta) == (5.0, 0.1, 4.5, 1.0)
    assert abs(c.max_keyframe_interval - 0.475) < 1e-15 and abs(c.min_keyframe_rotation - 0.17453292519943295) < 1e-15


def test_shim_compiles_against_the_abi(tmp_path):
    """The C++ mirror of lidar::LidarMap / LidarFrame / LidarFeature / LidarFactor builds and links."""
    exe = tmp_path / "shim_test"
    subprocess.check_call([GXX, "-std=c++17", "-O1", "-Wall", "-o", str(exe), os.path.join(ROOT, "tests", "shim_test.cc"),
                           "-L", os.path.join(ROOT, "ff-lins_b200"), "-lfflins_b200", "-L", os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "ff-lins_b200"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    assert exe.exists()


def test_no_cpu_fallback():
    """Without a CUDA device fflins_create refuses to run (FFLINS_E_NODEVICE); the product package never
    imports the oracle."""
    import torch
    from fflins import api
    if not torch.cuda.is_available():
        with pytest.raises(api.FflinsError) as e:
            api.Context()
        assert e.value.code == -6
    pkg = os.path.join(ROOT, "ff-lins_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "orc_" not in txt, f"{f} touches the oracle"


@pytest.mark.gpu
def test_shim_runs_on_gpu(tmp_path):
    exe = tmp_path / "shim_test"
    subprocess.check_call([GXX, "-std=c++17", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "shim_test.cc"),
                           "-L", os.path.join(ROOT, "ff-lins_b200"), "-lfflins_b200", "-L", os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "ff-lins_b200"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    out = subprocess.check_output([str(exe)], text=True)
    assert "SHIM_OK" in out, out


def test_host_library_exports_every_declared_symbol():
    """libfflins_host.so (INS mechanization, preintegration, window solver, converters) exports what include/fflins_host.h
    declares; it has no CUDA dependency."""
    from fflins import host
    src = open(os.path.join(ROOT, "include", "fflins_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(fflins_[a-z0-9_]+)\s*\(", src)) - {"fflins_lidar_eval_fn", "fflins_lidar_chi2_fn"})
    lib = host.load()
    missing = [n for n in names if not hasattr(lib, n)]
    assert len(names) >= 35 and not missing, missing
    needed = subprocess.check_output(["ldd", host.LIB_PATH], text=True)
    assert "cuda" not in needed.lower()


def test_native_session_driver_builds_and_exports():
    """tools/session_driver.cc — the native caller bench.py hands its timed loop to — links against the C-ABI only."""
    so = os.path.join(ROOT, "tools", "libfflins_driver.so")
    if not os.path.exists(so):
        subprocess.check_call([GXX, "-std=c++17", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "session_driver.cc"),
                               "-L", os.path.join(ROOT, "ff-lins_b200"), "-lfflins_b200", "-Wl,-rpath,$ORIGIN/../ff-lins_b200"])
    lib = C.CDLL(so)
    for name in ("drv_create", "drv_run", "drv_destroy", "drv_error", "drv_last_assoc", "drv_last_nres", "drv_last_blocks",
                 "drv_eval_roundtrip_us", "drv_trace_dump"):
        assert hasattr(lib, name), name
    src = open(os.path.join(ROOT, "tools", "session_driver.cc")).read()
    assert "cuda" not in src.lower() and "oracle" not in src.lower()
