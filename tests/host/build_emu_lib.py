"""TEST-ONLY: builds `libst3r_emu.so`, the C ABI of the MATCH and ALIGN paths (api.cu, scan.cu, radix_sort.cu, recip.cu,
nn_simt.cu, align.cu, align_dense.cu) compiled for the HOST: every `kernel<<<grid, block, smem, stream>>>(args);` is rewritten into a call of the SIMT
emulator (tests/host/simt_emu.h through emu_cuda_shim.h), everything else - the entry points' launch sequences,
workspace carving, device-side counters - is compiled as it stands.  The tcgen05 matcher runs against a software model of
TMA / mbarriers / tensor memory / tcgen05.mma (tests/host/tcgen05_emu.h).  Used by tests/test_match_emu_host.py."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "starst3r_b200", "csrc")
HOST = os.path.join(ROOT, "tests", "host")
SOURCES = ["api.cu", "scan.cu", "radix_sort.cu", "recip.cu", "nn_simt.cu", "align.cu", "align_dense.cu",
           "gs_project.cu", "gs_bin.cu", "gs_raster.cu", "gs_backward.cu", "gs_loss.cu", "gs_adam.cu", "gs_mcmc.cu", "nn_tc.cu"]
CLUSTER_KERNELS = {"focal_weiszfeld_cluster_kernel": "WZ_CLUSTER"}     # launched cluster by cluster

STUBS = r'''
#include "common.cuh"
extern "C" int st3r_emu_launch_failed(void) { return g_emu_launch_failed ? 1 : 0; }
'''


def split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)


def rewrite(text):
    """kernel<<<g, b, s, st>>>(args);  ->  EMU_LAUNCH(kernel, g, b, s, args);"""
    out, pos, n = "", 0, 0
    while True:
        m = LAUNCH.search(text, pos)
        if not m:
            return out + text[pos:], n
        cfg = split_top_level(m.group(2))
        assert len(cfg) in (2, 3, 4), cfg
        cfg += ["0"] * (3 - len(cfg)) if len(cfg) < 3 else []
        depth, i = 1, m.end()                       # find the parenthesis that closes the argument list
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        args = text[m.end():i - 1].strip()
        name = m.group(1)
        if name in CLUSTER_KERNELS:
            out += text[pos:m.start()] + (f"emu_launch_cluster({CLUSTER_KERNELS[name]}, dim3({cfg[0]}), dim3({cfg[1]}), "
                                          f"[&]() {{ {name}({args}); }})")
            pos, n = i, n + 1
            continue
        if "<" in name:
            name = "(" + name + ")"
        out += text[pos:m.start()] + f"EMU_LAUNCH({name}, ({cfg[0]}), ({cfg[1]}), ({cfg[2]})" + (", " + args if args else "") + ")"
        pos, n = i, n + 1


def build(out_dir, defines=()):
    """`defines`: extra -D macros (e.g. the experiment switches of nn_tc.cu)."""
    os.makedirs(out_dir, exist_ok=True)
    objs, launches = [], 0
    files = [(s, open(os.path.join(CSRC, s)).read()) for s in SOURCES] + [("emu_stubs.cu", STUBS)]
    for name, text in files:
        text, n = rewrite(text)
        launches += n
        cpp = os.path.join(out_dir, name[:-3] + "_emu.cpp")
        # includes are relative to csrc/: compile a copy that lives next to nothing, with csrc/ and tests/host/ on the path
        text = text.replace('#include "../../include/starst3r_b200.h"', f'#include "{ROOT}/include/starst3r_b200.h"')
        with open(cpp, "w") as fh:
            fh.write(text)
        obj = cpp[:-4] + ".o"
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DST3R_HOST_EMU=1", "-DST3R_EMU_WHOLE=1", *[f"-D{d}" for d in defines], "-I", CSRC, "-I", HOST, "-x", "c++", "-c", cpp,
                        "-o", obj], check=True)
        objs.append(obj)
    lib = os.path.join(out_dir, "libst3r_emu.so")
    subprocess.run(["g++", "-shared", "-o", lib] + objs, check=True)
    return lib, launches


if __name__ == "__main__":
    lib, n = build(sys.argv[1] if len(sys.argv) > 1 else "/tmp/st3r_emu")
    print(lib, f"({n} kernel launches rewritten)")
