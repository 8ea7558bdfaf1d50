"""Source rewriter of the CPU kernel emulator (TEST INFRASTRUCTURE ONLY, see tests/emu/README.md).

g++ cannot parse two things in the .cu/.cuh sources: the `kernel<<<grid, block, smem, stream>>>(args)` launch
syntax and inline PTX.  This script copies the sources to an output directory with

  kernel<<<g, b, s, st>>>(args)      ->  emu::launch(emu::d3(g), emu::d3(b), (size_t)(s), [&]() { kernel(args); })
  asm("rcp.approx.ftz.f64 ...")      ->  y = emu::rcp_approx(x)         (same 2^-23 seed accuracy)
  asm("rsqrt.approx.ftz.f64 ...")    ->  y = emu::rsqrt_approx(x)
  st.release / ld.acquire (sys)      ->  volatile store / load
  extern __shared__ T name[];        ->  T *name = (T *)emu::dyn_smem();
  dlopen("libnccl.so.2")             ->  dlopen(the emulated library itself), which carries tests/emu/fake_nccl.cpp

and nothing else: the kernel bodies compile unchanged against tests/emu/cuda_runtime.h.
Anything it does not recognise is an error, not a guess.
"""
import os
import re
import sys


def _match_back(text, pos, open_ch, close_ch):
    """text[pos] == close_ch; return index of the matching open_ch"""
    depth = 0
    i = pos
    while i >= 0:
        ch = text[i]
        if ch == close_ch:
            depth += 1
        elif ch == open_ch:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced %s%s before offset %d" % (open_ch, close_ch, pos))


def _match_fwd(text, pos, open_ch, close_ch):
    depth = 0
    i = pos
    while i < len(text):
        ch = text[i]
        if ch == open_ch:
            depth += 1
        elif ch == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced %s%s after offset %d" % (open_ch, close_ch, pos))


def _split_top(s):
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur).strip())
    return out


def rewrite_launches(text, fname):
    n = 0
    while True:
        p = text.find("<<<")
        if p < 0:
            break
        # kernel expression: identifier, optionally followed by <template args>
        i = p - 1
        while text[i].isspace():
            i -= 1
        end_name = i + 1
        if text[i] == ">":
            i = _match_back(text, i, "<", ">") - 1
            while text[i].isspace():
                i -= 1
        while i >= 0 and (text[i].isalnum() or text[i] in "_:"):
            i -= 1
        start_name = i + 1
        kernel = text[start_name:end_name]
        if not re.match(r"[A-Za-z_]", kernel):
            raise ValueError(f"{fname}: cannot find the kernel name before <<< at offset {p}")
        q = text.find(">>>", p)
        cfg = _split_top(text[p + 3:q])
        if not 2 <= len(cfg) <= 4:
            raise ValueError(f"{fname}: launch configuration with {len(cfg)} arguments: {cfg}")
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        j = q + 3
        while text[j].isspace():
            j += 1
        if text[j] != "(":
            raise ValueError(f"{fname}: no argument list after >>> of {kernel}")
        k = _match_fwd(text, j, "(", ")")
        args = text[j + 1:k]
        new = (f"emu::launch(\"{kernel.split('<')[0]}\", emu::d3({grid}), emu::d3({block}), (size_t)({smem}), "
               f"[&]() {{ {kernel}({args}); }})")
        text = text[:start_name] + new + text[k + 1:]
        n += 1
    return text, n


ASM_RULES = [
    (re.compile(r'asm\s*\(\s*"rcp\.approx\.ftz\.f64 %0, %1;"\s*:\s*"=d"\((\w+)\)\s*:\s*"d"\((\w+)\)\s*\)\s*;'),
     r"\1 = emu::rcp_approx(\2);"),
    (re.compile(r'asm\s*\(\s*"rsqrt\.approx\.ftz\.f64 %0, %1;"\s*:\s*"=d"\((\w+)\)\s*:\s*"d"\((\w+)\)\s*\)\s*;'),
     r"\1 = emu::rsqrt_approx(\2);"),
    (re.compile(r'asm\s+volatile\s*\(\s*"st\.release\.sys\.global\.u64 \[%0\], %1;"\s*::\s*"l"\(([^)]+)\)\s*,\s*"l"\(([^)]+)\)\s*:\s*"memory"\s*\)\s*;'),
     r"*(volatile unsigned long long *)(\1) = (\2);"),
    (re.compile(r'asm\s+volatile\s*\(\s*"ld\.acquire\.sys\.global\.u64 %0, \[%1\];"\s*:\s*"=l"\((\w+)\)\s*:\s*"l"\(([^)]+)\)\s*:\s*"memory"\s*\)\s*;'),
     r"\1 = *(const volatile unsigned long long *)(\2);"),
]

DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\]\s*;")


def rewrite(text, fname):
    text, nl = rewrite_launches(text, fname)
    for rx, rep in ASM_RULES:
        text = rx.sub(rep, text)
    text = DYN_SMEM.sub(r"\1 *\2 = (\1 *)emu::dyn_smem();", text)
    # the product opens libnccl.so.2 at run time; the emulated build opens itself, where tests/emu/fake_nccl.cpp lives
    text = text.replace('const char *names[] = {"libnccl.so.2", "libnccl.so"};',
                        'const char *names[] = {emu::self_path(), emu::self_path()};')
    # whatever inline assembly is left would reach the x86 assembler: refuse
    code = re.sub(r"//[^\n]*", "", text)
    if re.search(r"\basm\b", code):
        raise ValueError(f"{fname}: inline assembly the emulator has no rule for")
    if "extern __shared__" in code:
        raise ValueError(f"{fname}: dynamic shared memory declaration the emulator has no rule for")
    return text, nl


def main(src_dir, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    total = 0
    for name in sorted(os.listdir(src_dir)):
        if not name.endswith((".cu", ".cuh")):
            continue
        with open(os.path.join(src_dir, name)) as fh:
            text = fh.read()
        text, nl = rewrite(text, name)
        total += nl
        out = name[:-3] + ".cpp" if name.endswith(".cu") else name
        with open(os.path.join(out_dir, out), "w") as fh:
            fh.write(text)
    return total


if __name__ == "__main__":
    print(main(sys.argv[1], sys.argv[2]), "launches rewritten")
