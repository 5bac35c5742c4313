"""TEST INFRASTRUCTURE: compile the package's CUDA sources for the host (see cuda_runtime.h here).

The .cu files are used as they are; three purely syntactic rewrites make them C++:

* ``kernel<<<grid, block, smem, stream>>>(args);`` -> ``p360_emul::launch(grid, block, smem, [=] { kernel(args); });``
* ``extern __shared__ T name[];``                  -> ``T *name = static_cast<T *>(p360_emul::dyn_smem());``
* inline PTX (cache-hinted loads / stores / prefetch in p360_common.cuh) -> the plain C++ access

Output: tests/emul/_build/libpano360_emul.so with the same C ABI as libpano360_b200.so,
taking host pointers.  Nothing under pano360_b200/ knows this library exists.
"""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pano360_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libpano360_emul.so")


def _matching(text, start, open_ch, close_ch):
    """Index just past the bracket that closes text[start] (which must be ``open_ch``)."""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced brackets")


def _split_top_level(text):
    parts, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        at = text.find("<<<", pos)
        if at < 0:
            out.append(text[pos:])
            return "".join(out)
        # kernel name (with optional template arguments) ends right before <<<
        name_end = at
        i = at - 1
        if text[i] == ">":                                  # template arguments: walk back to their '<'
            depth = 0
            while True:
                if text[i] == ">":
                    depth += 1
                elif text[i] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                i -= 1
            i -= 1
        while i >= 0 and (text[i].isalnum() or text[i] in "_:"):
            i -= 1
        name_start = i + 1
        close = text.index(">>>", at)
        config = _split_top_level(text[at + 3:close])
        while len(config) < 4:
            config.append("0")
        args_open = close + 3
        while text[args_open].isspace():
            args_open += 1
        assert text[args_open] == "(", text[at - 40:at + 80]
        args_close = _matching(text, args_open, "(", ")")
        stmt_end = args_close
        while text[stmt_end].isspace():
            stmt_end += 1
        assert text[stmt_end] == ";", text[at - 40:at + 120]
        name = text[name_start:name_end]
        out.append(text[pos:name_start])
        out.append("p360_emul::launch(%s, %s, %s, [=]() { %s%s; });"
                   % (config[0], config[1], config[2], name, text[args_open:args_close]))
        pos = stmt_end + 1


def rewrite_dynamic_shared(text):
    return re.sub(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;",
                  r"\1 *\2 = static_cast<\1 *>(p360_emul::dyn_smem());", text)


def rewrite_inline_ptx(text):
    def plain(match):
        body = match.group(0)
        if "ld.global" in body:
            return "r = *p;"
        if "st.global" in body:
            return "*p = v;"
        if "prefetch" in body:
            return "(void)p;"
        if "mbarrier" in body or "cp.async.bulk" in body or "fence.proxy" in body:
            return "p360_emul::no_tma();"          # (kHaveTma is false in this build: never reached)
        raise ValueError("inline PTX the host build does not know: " + body[:80])
    return re.sub(r"asm\s+volatile\s*\(.*?\)\s*;", plain, text, flags=re.S)


def translate(text):
    text = text.replace('#include "../../include/pano360_b200.h"',
                        '#include "%s"' % os.path.join(ROOT, "include", "pano360_b200.h"))
    return rewrite_launches(rewrite_dynamic_shared(rewrite_inline_ptx(text)))


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def _fingerprint():
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, f) for f in sources()] + [
            os.path.join(HERE, "cuda_runtime.h"), os.path.abspath(__file__),
            os.path.join(ROOT, "include", "pano360_b200.h")]:
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False):
    """Translate + compile when a source changed; returns the library path."""
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "fingerprint")
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB
    units = []
    for name in sources():
        with open(os.path.join(CSRC, name)) as fh:
            text = translate(fh.read())
        dst = os.path.join(OUT_DIR, name if name.endswith(".cuh") else name[:-3] + ".cpp")
        with open(dst, "w") as fh:
            fh.write(text)
        if name.endswith(".cu"):
            units.append(dst)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-march=native", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
           "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-I", HERE, "-o", LIB] + units
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("host build of the kernels failed:\n" + proc.stdout + proc.stderr)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
