"""
TEST INFRASTRUCTURE ONLY -- builds tests/emu/_build/liblqcd_b200_emu.so: the library's own .cu/.cuh sources, mechanically
translated to plain C++ (kernel launches -> emu::launch, inline PTX -> host equivalents) and compiled with g++ against the
SIMT emulator of cuda_emu.h / emu_runtime.cpp.  Purpose: pre-flight device code on the GPU-less build container (see
cuda_emu.h for what it does and does not validate).  The product never loads this library.

The translation is purely textual and fails loudly on anything it does not know (an unknown PTX string, a launch it cannot
parse), so a new construct in csrc/ cannot silently change meaning under emulation.

Usage: python tests/emu/build_emu.py [--asan] [--force]
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
CSRC = ROOT / "latticeqcd.jl_b200" / "csrc"
BUILD = HERE / "_build"

# inline PTX -> C++ (keyed by the opcode at the start of the asm string; {o0},{i0}.. = operand expressions)
PTX = {
    "st.release.sys.global.u64": "__atomic_store_n({i0}, {i1}, __ATOMIC_RELEASE);",
    "ld.relaxed.sys.global.u64": "{o0} = __atomic_load_n({i0}, __ATOMIC_ACQUIRE);",
    "ld.acquire.sys.global.u64": "{o0} = __atomic_load_n({i0}, __ATOMIC_ACQUIRE);",
    "prefetch.global.L2": "(void)({i0});",
    "ld.global.nc.L1::evict_first.v2.f64": "{o0} = ({i0})->x; {o1} = ({i0})->y;",
    "fence.mbarrier_init.release.cluster": ";",
    "mov.u64": "{o0} = (unsigned long long)__builtin_ia32_rdtsc();",
}
# helper functions whose bodies are PTX over 32-bit shared-window addresses: replaced wholesale by the mbarrier model
FUNCS = {
    "smem_u32": None,     # dropped
    "mbar_init": "__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { emu::mbar_init(bar, count); }",
    "mbar_arrive_expect_tx": "__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { emu::mbar_arrive_expect_tx(bar, bytes); }",
    "bulk_g2s": "__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { emu::bulk_g2s(dst, src, bytes, bar); }",
    "mbar_wait": "__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { emu::mbar_wait(bar, parity); }",
}


class TranslateError(RuntimeError):
    pass


def _match_close(s: str, i: int, open_c: str, close_c: str) -> int:
    """index of the bracket closing the one at s[i]"""
    depth, j, n = 0, i, len(s)
    in_str = False
    while j < n:
        c = s[j]
        if in_str:
            if c == "\\":
                j += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True
        elif c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return j
        j += 1
    raise TranslateError("unbalanced bracket")


def _split_top(s: str, sep: str) -> list[str]:
    out, depth, cur, in_str = [], 0, [], False
    i = 0
    while i < len(s):
        c = s[i]
        if in_str:
            cur.append(c)
            if c == "\\":
                cur.append(s[i + 1]); i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True; cur.append(c)
        elif c in "([{":
            depth += 1; cur.append(c)
        elif c in ")]}":
            depth -= 1; cur.append(c)
        elif c == sep and depth == 0:
            out.append("".join(cur)); cur = []
        else:
            cur.append(c)
        i += 1
    out.append("".join(cur))
    return out


def _operands(section: str) -> list[str]:
    ops = []
    for part in _split_top(section, ","):
        part = part.strip()
        if not part:
            continue
        m = re.match(r'"[^"]*"\s*\(', part)
        if not m:
            raise TranslateError(f"cannot parse asm operand {part!r}")
        ops.append(part[m.end():part.rindex(")")].strip())
    return ops


def translate_asm(src: str) -> str:
    out, i = [], 0
    pat = re.compile(r"\basm\s*(volatile)?\s*\(")
    while True:
        m = pat.search(src, i)
        if not m:
            out.append(src[i:]); break
        out.append(src[i:m.start()])
        close = _match_close(src, m.end() - 1, "(", ")")
        body = src[m.end():close]
        end = close + 1
        if src[end:end + 1] == ";":
            end += 1
        # string literal(s) first, then ':' separated operand sections
        sm = re.match(r'\s*((?:"(?:[^"\\]|\\.)*"\s*)+)', body)
        if not sm:
            raise TranslateError("asm without string literal")
        text = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', sm.group(1)))
        rest = body[sm.end():]
        secs = _split_top(rest, ":")
        outs = _operands(secs[1]) if len(secs) > 1 else []
        ins = _operands(secs[2]) if len(secs) > 2 else []
        opcode = text.strip().split()[0].rstrip(";")
        if opcode not in PTX:
            raise TranslateError(f"unknown inline PTX {opcode!r}: add it to tests/emu/build_emu.py:PTX")
        fmt = {f"o{k}": v for k, v in enumerate(outs)}
        fmt.update({f"i{k}": v for k, v in enumerate(ins)})
        out.append(PTX[opcode].format(**fmt))
        i = end
    return "".join(out)


def translate_funcs(src: str) -> str:
    for name, repl in FUNCS.items():
        m = re.search(r"__device__\s+__forceinline__\s+[\w\s\*]+?\b" + name + r"\s*\(", src)
        if not m:
            continue
        brace = src.index("{", m.end())
        close = _match_close(src, brace, "{", "}")
        src = src[:m.start()] + (repl or "") + src[close + 1:]
    return src


def translate_launches(src: str) -> str:
    out, i = [], 0
    while True:
        k = src.find("<<<", i)
        if k < 0:
            out.append(src[i:]); break
        # kernel expression: identifier, optionally followed by a template argument list, right before '<<<'
        j = k
        while src[j - 1].isspace():
            j -= 1
        if src[j - 1] == ">":
            depth, j2 = 0, j - 1
            while True:
                if src[j2] == ">":
                    depth += 1
                elif src[j2] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j2 -= 1
            j = j2
        j0 = j
        while src[j0 - 1].isalnum() or src[j0 - 1] == "_" or src[j0 - 1] == ":":
            j0 -= 1
        kernel = src[j0:k].strip()
        if not re.match(r"[A-Za-z_]", kernel):
            raise TranslateError(f"cannot find the kernel name before <<< near: {src[max(0, k - 60):k + 20]!r}")
        e = src.index(">>>", k)
        cfg = [c.strip() for c in _split_top(src[k + 3:e], ",")]
        if not 2 <= len(cfg) <= 4:
            raise TranslateError(f"bad launch configuration {cfg}")
        cfg += ["0", "nullptr"][len(cfg) - 2:]
        p = e + 3
        while src[p].isspace():
            p += 1
        if src[p] != "(":
            raise TranslateError("launch without argument list")
        close = _match_close(src, p, "(", ")")
        args = src[p + 1:close]
        out.append(src[i:j0])
        name = kernel.split("<")[0]
        out.append(f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), {cfg[2]}, {cfg[3]}, [&]() {{ {kernel}({args}); }}, \"{name}\")")
        i = close + 1
    return "".join(out)


def translate(text: str) -> str:
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_emu.h"')
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];", r"unsigned char *\1 = emu::dynamic_smem();", text)
    text = translate_funcs(text)
    text = translate_asm(text)
    text = translate_launches(text)
    if "<<<" in text or re.search(r"\basm\b", text):
        raise TranslateError("untranslated CUDA construct left")
    return text


def sources():
    sys.path.insert(0, str(ROOT / "latticeqcd.jl_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("lqcd_b200_build", ROOT / "latticeqcd.jl_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return list(mod.SOURCES)


def build(asan: bool = False, force: bool = False) -> Path:
    """(re)builds the emulated library if any input is newer; safe to call from several processes at once (pytest-xdist workers):
    one builds under an exclusive file lock, the others wait and then find it up to date"""
    import fcntl
    BUILD.mkdir(parents=True, exist_ok=True)
    with open(BUILD / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(asan, force)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(asan: bool, force: bool) -> Path:
    tag = "_asan" if asan else ""
    srcdir = BUILD / ("src" + tag)
    srcdir.mkdir(parents=True, exist_ok=True)
    out = BUILD / f"liblqcd_b200_emu{tag}.so"
    inputs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HERE / "cuda_emu.h", HERE / "emu_runtime.cpp", Path(__file__),
                                                                       ROOT / "include" / "lqcd_b200.h"]
    newest = max(p.stat().st_mtime for p in inputs)
    if out.exists() and out.stat().st_mtime > newest and not force:
        return out
    for p in list(CSRC.glob("*.cuh")):
        (srcdir / p.name).write_text(translate(p.read_text()))
    units = []
    for name in sources():
        dst = srcdir / (name + ".cpp")
        dst.write_text(translate((CSRC / name).read_text()))
        units.append(dst)
    units.append(HERE / "emu_runtime.cpp")
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-march=native", "-ffp-contract=fast", "-fno-strict-aliasing", "-w",
             f"-I{srcdir}", f"-I{HERE}", f"-I{CSRC}"]
    if asan:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    objs = []

    def cc(u: Path):
        o = srcdir / (u.name + ".o")
        r = subprocess.run(["g++", *flags, "-c", str(u), "-o", str(o)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on {u.name}:\n{r.stderr[-6000:]}")
        return o

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(cc, units))
    tmp = out.with_suffix(".so.tmp")
    link = ["g++", "-shared", "-o", str(tmp), *map(str, objs), "-lrt", "-lm"]
    if asan:
        link.append("-fsanitize=address")
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build(asan="--asan" in sys.argv, force="--force" in sys.argv))
