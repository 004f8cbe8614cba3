"""Builds tests/emu/_gen/libk4emu.so and libk3emu.so: the DEVICE source of qsft_b200/csrc/k4_peel.cu / k3_gwht.cu compiled
by g++ against the CPU execution shim (cuda_emu.h) plus small host drivers (k4_emu.cpp, k3_emu.cpp).  Test infrastructure
only."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GEN = os.path.join(HERE, "_gen")
LIB = os.path.join(GEN, "libk4emu.so")


def _rewrite(body):
    body = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+(?: \w+)*?) (\w+)\[\];",
                  r"\1* \2 = reinterpret_cast<\1*>(emu_dyn_smem());", body)
    return re.sub(r"\b__shared__\b", "static", body)


def device_part(which="k4"):
    if which == "k4":
        src = open(os.path.join(ROOT, "qsft_b200", "csrc", "k4_peel.cu")).read()
        start = src.index('#include "common.cuh"') + len('#include "common.cuh"')
        end = src.index("int make_dev(")                  # host code (with <<< >>> launches) starts here
        text = _rewrite(src[start:end]) + "\n}  // namespace (closed by build_emu.py)\n"
        # the persistent on-device round loop (plain-copy variant; the TMA variant is sm_100a only)
        src = open(os.path.join(ROOT, "qsft_b200", "csrc", "k4_peel_loop.cuh")).read()
        start = src.index('#include "common.cuh"') + len('#include "common.cuh"')
        end = src.index("}  // namespace (device part; the CPU emulation cuts here)")
        return text + _rewrite(src[start:end]) + "\n}  // namespace (closed by build_emu.py)\n"
    if which == "k1":
        src = open(os.path.join(ROOT, "qsft_b200", "csrc", "k1_lattice.cu")).read()
        start = src.index('#include "common.cuh"') + len('#include "common.cuh"')
        end = src.index('extern "C" int qsft_query_lattice(')   # the anonymous namespace is already closed here
        return _rewrite(src[start:end])
    if which == "k2l":
        # operand generation of the lattice-factorised evaluation (pure integer SIMT code); the GEMM kernels are tcgen05 / TMA
        src = open(os.path.join(ROOT, "qsft_b200", "csrc", "k2_eval_lattice.cu")).read()
        start = src.index("// sum_i digit_i(a) * digit_i(b) mod 4")
        end = src.index("// ---- the GEMM ----")
        head = "namespace {\nconstexpr int LT_SCALE_BITS = %s;\n" % re.search(r"constexpr int LT_SCALE_BITS = (\d+);", src).group(1)
        return head + _rewrite(src[start:end]) + "\n}  // namespace (closed by build_emu.py)\n"
    if which == "k2":
        src = open(os.path.join(ROOT, "qsft_b200", "csrc", "k2_eval_simt.cu")).read()
        start = src.index('#include "common.cuh"') + len('#include "common.cuh"')
        end = src.index("template <int NW>\nint launch_nw(")
        return _rewrite(src[start:end]) + "\n}  // namespace (closed by build_emu.py)\n"
    src = open(os.path.join(ROOT, "qsft_b200", "csrc", "k3_gwht.cu")).read()
    start = src.index('#include "common.cuh"') + len('#include "common.cuh"')
    end = src.index("template <int Q>\nint launch_pass(")  # kernels + PassPlan / plan_pass (plain host code); launches follow
    return _rewrite(src[start:end]) + "\n}  // namespace (closed by build_emu.py)\n"


def build(force=False, which="k4"):
    os.makedirs(GEN, exist_ok=True)
    lib = os.path.join(GEN, f"lib{which}emu.so")
    inc = os.path.join(GEN, f"{which}_device.inc")
    cu = {"k4": "k4_peel.cu", "k3": "k3_gwht.cu", "k1": "k1_lattice.cu", "k2": "k2_eval_simt.cu", "k2l": "k2_eval_lattice.cu"}[which]
    text = device_part(which)
    srcs = [os.path.join(HERE, f"{which}_emu.cpp"), os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "qsft_b200", "csrc", cu)]
    if which == "k4":
        srcs += [os.path.join(ROOT, "qsft_b200", "csrc", f) for f in ("k4_peel_loop.cuh", "k4_shared.cuh")]
    fresh = os.path.exists(lib) and os.path.exists(inc) and open(inc).read() == text and \
        all(os.path.getmtime(lib) >= os.path.getmtime(p) for p in srcs)
    if fresh and not force:
        return lib
    open(inc, "w").write(text)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-w", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "qsft_b200", "csrc"),
           os.path.join(HERE, f"{which}_emu.cpp"), "-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the kernel emulation failed:\n" + res.stderr[-4000:])
    return lib


if __name__ == "__main__":
    print(build(force=True, which="k4"))
    print(build(force=True, which="k3"))
    print(build(force=True, which="k1"))
    print(build(force=True, which="k2"))
    print(build(force=True, which="k2l"))
