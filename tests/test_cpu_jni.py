"""The Java/JNI side of the boundary, checked without a JDK (there is none in this image):

* the generator table (finmath-lib_b200/csrc/jni/gen_jni.py) has one row per function of include/finmath_b200.h and the committed shim /
  Java class are what it generates;
* the shim compiles with -Wall -Wextra -Werror against tests/stubs/jni.h (written from the JNI specification);
* every Java_net_finmath_cuda_FinmathB200_* entry point is driven through a fake JNIEnv: without a device each computing call must
  throw RuntimeException (FMB_ENODEVICE: no CPU fallback), bad arguments IllegalArgumentException, host-only calls must work;
  on a B200 (-m gpu) the same binary runs a small end-to-end workflow through the shim and checks the numbers.
"""
import importlib.util
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI_DIR = os.path.join(ROOT, "finmath-lib_b200", "csrc", "jni")
SHIM = os.path.join(JNI_DIR, "finmath_b200_jni.c")
JAVA = os.path.join(ROOT, "finmath-lib_b200", "java", "net", "finmath", "cuda", "FinmathB200.java")
STUBS = os.path.join(ROOT, "tests", "stubs")


def _generator():
    spec = importlib.util.spec_from_file_location("gen_jni", os.path.join(JNI_DIR, "gen_jni.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shim_covers_every_function_of_the_header_and_is_up_to_date():
    g = _generator()
    header = open(os.path.join(ROOT, "include", "finmath_b200.h")).read()
    declared = set(re.findall(r"\b(fmb_[a-z0-9_]+)\s*\(", header))
    bound = {row[1] for row in g.TABLE}
    assert declared - bound == set(), "functions of the header without a JNI binding: %s" % sorted(declared - bound)
    assert bound - declared == set(), sorted(bound - declared)
    assert open(SHIM).read() == g.gen_c(), "finmath_b200_jni.c is stale: run gen_jni.py"
    assert open(JAVA).read() == g.gen_java(), "FinmathB200.java is stale: run gen_jni.py"
    # one native method per row, one Java_ function per row
    assert len(re.findall(r"public static native", open(JAVA).read())) == len(g.TABLE)
    assert len(re.findall(r"JNIEXPORT", open(SHIM).read())) == len(g.TABLE)


def _build(tmp_path):
    import __graft_entry__ as graft
    graft.build()
    lib_dir = os.path.join(ROOT, "finmath-lib_b200")
    exe = str(tmp_path / "jni_fake_env_test")
    flags = ["-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", STUBS]
    subprocess.check_call(["gcc"] + flags + ["-c", SHIM, "-o", str(tmp_path / "shim.o")])
    subprocess.check_call(["gcc"] + flags + ["-c", os.path.join(STUBS, "jni_fake_env_test.c"), "-o", str(tmp_path / "driver.o")])
    subprocess.check_call(["gcc", str(tmp_path / "driver.o"), str(tmp_path / "shim.o"), "-L", lib_dir, "-lfinmath_b200", "-lm", "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def test_shim_compiles_and_every_entry_point_runs_through_a_fake_jnienv(tmp_path):
    exe = _build(tmp_path)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")           # the no-device contract, also on a GPU box
    out = subprocess.run([exe, "nodevice"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout


@pytest.mark.gpu
def test_shim_end_to_end_on_the_device(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout
