"""The C-ABI library builds for sm_100a without a GPU, loads, exports every symbol that
include/orbgpu.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "orbgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orbgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_library_agree(built):
    import refactored_orb_slam2_b200._lib as L
    names = header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "liborbgpu.so does not export %s" % n
    assert sorted(L.SYMBOLS) == names, "python binding table and header drifted apart"
    assert L.lib().orbgpu_version() == 160


def test_keypoint_record_is_cv_keypoint_layout(built):
    import refactored_orb_slam2_b200 as pkg
    assert pkg.KP_DTYPE.itemsize == 28
    assert [pkg.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_no_cpu_fallback(built):
    import refactored_orb_slam2_b200 as pkg
    if pkg.lib().orbgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.OrbGpuError) as e:
        pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    assert e.value.code == -2
    with pytest.raises(pkg.OrbGpuError):
        pkg.DescriptorBank(np.zeros((4, 32), np.uint8))
    with pytest.raises(pkg.OrbGpuError):
        pkg.ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.zeros(32, np.uint8))
    # the widened rows have no CPU path either: feature sets / searches, vocabulary descent, staging, distinctive descriptors,
    # the sharded-bank exchange
    kp = np.zeros(3, pkg.KP_DTYPE)
    for call in (lambda: pkg.FrameFeatures(kp, np.zeros((3, 32), np.uint8), None, pkg.grid_geometry(0.0, 640.0, 0.0, 480.0)),
                 lambda: pkg.Vocabulary([0, 2, 2, 2], [1, 2], np.zeros((3, 32), np.uint8), [-1, 0, 1], 1),
                 lambda: pkg.staging.cvtColorGray(np.zeros((4, 4, 3), np.uint8)),
                 lambda: pkg.staging.remap(np.zeros((4, 4), np.uint8), np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)),
                 lambda: pkg.distinctive_descriptors(np.zeros((3, 32), np.uint8), [0, 3])):
        with pytest.raises(pkg.OrbGpuError) as e:
            call()
        assert e.value.code == -2
    import ctypes as C
    out, handle = C.c_void_p(), np.zeros(64, np.uint8)
    assert pkg.lib().orbgpu_exchange_create(0, 1, 16, -1, C.byref(out), handle.ctypes.data_as(C.c_void_p)) == -2


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "refactored_orb_slam2_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "oracle/").lower() or f in () or _only_comments(src), \
                    "%s refers to the oracle" % f


def _only_comments(src):
    for line in src.splitlines():
        code = line.split("//")[0].split("#")[0]
        if "oracle" in code.lower() and not code.strip().startswith(("*", '"', "'")):
            return False
    return True


def test_host_policy_matches_oracle(built):
    """ORBmatcher acceptance + rotation histogram (host code) vs the oracle's restatement."""
    import oracle
    import refactored_orb_slam2_b200 as pkg
    rng = np.random.default_rng(0)
    d1, d2 = rng.integers(0, 120, 2000), rng.integers(0, 257, 2000)
    for ratio in (0.6, 0.7, 0.75, 0.9):
        m = pkg.ORBmatcher(ratio)
        assert np.array_equal(m.accept(d1, d2), oracle.accept(d1, d2, ratio, 50))
    a1 = rng.uniform(0, 360, 500).astype(np.float32)
    a2 = np.where(rng.random(500) < 0.7, a1 - 20, rng.uniform(0, 360, 500)).astype(np.float32) % np.float32(360)
    assert np.array_equal(pkg.ORBmatcher.rotation_consistent(a1, a2), oracle.rotation_filter(a1, a2))


def test_bench_own_arm_does_not_touch_the_oracle():
    """bench.py may reach into oracle/ only from the CPU-baseline leg (cpu_reference): nowhere else may it import the oracle
    modules or put oracle/ on sys.path."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    oracle_modules = {"oracle", "ref", "gen_synth", "oracle_cv2", "slam_ref", "slam_scenarios", "voc_ref"}
    offenders = []
    for fn in ast.walk(tree):
        if isinstance(fn, (ast.FunctionDef, ast.Module)):
            for node in (fn.body if isinstance(fn, ast.Module) else ast.walk(fn)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name.split(".")[0] for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    names = [node.module.split(".")[0]]
                for n in names:
                    if n in oracle_modules and not (isinstance(fn, ast.FunctionDef) and fn.name == "cpu_reference"):
                        offenders.append((getattr(fn, "name", "<module>"), n))
    assert not offenders, offenders
    assert src.count('os.path.join(ROOT, "oracle")') == 1      # the sys.path entry is added inside cpu_reference only
