"""Source-to-source step of the host emulation (TEST INFRASTRUCTURE ONLY, see tests/hostemu/__init__.py).

Rewrites the two CUDA constructs g++ cannot parse, and nothing else, so that the product's .cu/.cuh
files compile UNMODIFIED IN MEANING as C++ against tests/hostemu/cuda_emu.h:

  kernel<<<grid, block, smem, stream>>>(args);   ->  SIC_EMU_LAUNCH(kernel, grid, block, smem, stream, args);
  asm volatile("ld.global[.nc|.cg].{f64,s32,f32} %0, [%1];" : "=d|r|f"(dst) : "l"(ptr));   ->  dst = *(ptr);
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(ptr));  ->  two 32-bit loads
  #include "ebe_tma.cuh"   ->  #include "fem.cuh"   (the TMA/mbarrier variant is compiled out: SIC_EBE_IMPL == 3)
  static T* g_xxx ...      ->  static thread_local T* g_xxx   (host-side singletons: one per emulated rank = host thread)
"""
import re


def _match_forward(s, i, open_ch, close_ch):
    """s[i] == open_ch; return the index just past the matching close_ch."""
    depth = 0
    k = i
    while k < len(s):
        c = s[k]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return k + 1
        k += 1
    raise ValueError("unbalanced " + open_ch)


def _kernel_start(s, j):
    """j = index of '<<<'; return the index where the kernel expression (name[<targs>]) begins."""
    k = j
    if s[k - 1] == ">":                       # template arguments
        depth = 0
        while True:
            k -= 1
            if s[k] == ">":
                depth += 1
            elif s[k] == "<":
                depth -= 1
                if depth == 0:
                    break
    while k > 0 and (s[k - 1].isalnum() or s[k - 1] in "_:"):
        k -= 1
    return k


def _split_top(s):
    """split on top-level commas"""
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{<" and not (c == "<"):
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


def translate_launches(s):
    out = ""
    pos = 0
    while True:
        j = s.find("<<<", pos)
        if j < 0:
            return out + s[pos:]
        k0 = _kernel_start(s, j)
        e = s.find(">>>", j)
        cfg = _split_top(s[j + 3:e])
        while len(cfg) < 4:
            cfg.append("0")
        a0 = e + 3
        while s[a0].isspace():
            a0 += 1
        assert s[a0] == "(", "kernel launch without argument list near: " + s[j - 40:j + 40]
        a1 = _match_forward(s, a0, "(", ")")
        args = s[a0 + 1:a1 - 1].strip()
        kern = s[k0:j]
        out += s[pos:k0] + "SIC_EMU_LAUNCH((" + kern + "), " + ", ".join(cfg) + (", " + args if args else "") + ")"
        pos = a1


_LD1 = re.compile(r'asm volatile\("ld\.global(?:\.nc|\.cg)?\.(?:f64|s32|f32|u16) %0, \[%1\];"\s*:\s*"=[drfh]"\((\w+)\)\s*:\s*"l"\((.+?)\)\);',
                  re.S)
_LD2 = re.compile(r'asm volatile\("ld\.global\.nc\.v2\.u32 \{%0, %1\}, \[%2\];"\s*:\s*"=r"\((\w+)\),\s*"=r"\((\w+)\)\s*:\s*'
                  r'"l"\((.+?)\)\);', re.S)


def strip_comments(s):
    """Remove // and /* */ comments (string literals in these sources never contain comment markers), keep newlines."""
    s = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def translate(src: str) -> str:
    s = strip_comments(src).replace('#include "ebe_tma.cuh"', '#include "fem.cuh"')
    s = _LD2.sub(lambda m: "{ const unsigned* _p2 = (const unsigned*)(%s); %s = _p2[0]; %s = _p2[1]; }"
                 % (m.group(3), m.group(1), m.group(2)), s)
    s = _LD1.sub(lambda m: "%s = *(%s);" % (m.group(1), m.group(2)), s)
    s = translate_launches(s)
    # host-side singletons of the drivers (pinned mirrors of the device scalars, timing events): one per emulated rank
    s = re.sub(r"^static (\w+\*? g_\w+)", r"static thread_local \1", s, flags=re.M)
    if "asm volatile" in s or "<<<" in s:
        raise ValueError("untranslated CUDA construct left in the source")
    return s
