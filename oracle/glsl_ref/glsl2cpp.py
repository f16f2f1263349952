#!/usr/bin/env python3
"""glsl2cpp.py — make the reference's own fragment shaders compilable by g++.  TEST INFRASTRUCTURE.

Reads a shader where it lies under /root/reference/src/shaders (nothing is copied into the repo; the output
goes to oracle/_ref/glsl_ref/, which is git-ignored), resolves its `#include`s the way the reference's
ShaderInclude loader does (src/core/ShaderIncludes.h: path relative to the shader directory), and applies the
few purely lexical rewrites that separate GLSL 3.30 from C++20.  No statement of the shader is reordered,
dropped or re-expressed:

  * comments and `#version` removed;
  * floating literals get an `f` suffix (GLSL literals are fp32; C++ ones would be double);
  * `uniform T x;`            -> `T x;`                 (set by the harness, as glUniform* does)
  * file-scope `in/out T x;`  -> `thread_local T x;`    (per-fragment interface variables)
  * other file-scope variables-> `thread_local ...`     (per-invocation globals: the RNG seed, ...)
  * parameter qualifiers      -> `in T x` = by value, `out/inout T x` = by reference
  * struct members get `{}` initialisers (GLSL leaves them undefined; zero is what the drivers give)
  * `main` -> `glsl_main`;
  * calls with two `rand()` arguments get the draws hoisted in source order (GLSL evaluates arguments left to right,
    g++ right to left).

`#ifdef OPT_*` blocks are left in place for the C preprocessor: the harness is compiled once per set of
defines, exactly as Renderer::InitShaders (Renderer.cpp:396-470) recompiles the shader per option set.
Types, swizzles and built-ins come from glsl_compat.h.
"""
from __future__ import annotations
import os, re, sys


def load_with_includes(path: str, root: str, seen=None) -> str:
    seen = seen or set()
    out = []
    for line in open(path, encoding="utf-8", errors="replace").read().splitlines():
        m = re.match(r"\s*#include\s+(\S+)", line)
        if m:
            inc = os.path.join(root, m.group(1).strip('"<>'))
            if inc in seen:
                continue
            seen.add(inc)
            out.append(load_with_includes(inc, root, seen))
        else:
            out.append(line)
    return "\n".join(out)


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


_FLOAT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


_MEMBER = re.compile(r"(\b\w+\s+\w+(?:\s*\[\s*\w+\s*\])?)\s*;")


def _toplevel_statement(stmt: str) -> str:
    """One file-scope statement (text up to and including its ';')."""
    s = stmt.strip()
    if not s or s == ";":
        return stmt
    lead = stmt[: len(stmt) - len(stmt.lstrip())]
    m = re.match(r"uniform\s+(.*)$", s, flags=re.S)
    if m:
        return lead + m.group(1)
    m = re.match(r"(?:in|out)\s+(\w+\s+\w+\s*;)$", s)
    if m:
        return lead + "thread_local " + m.group(1)
    if "(" in s.split("=", 1)[0]:            # function prototype
        return stmt
    return lead + "thread_local " + s


def translate(src: str) -> str:
    src = strip_comments(src)
    src = re.sub(r"^\s*#version[^\n]*", "", src, flags=re.M)
    src = re.sub(r"^\s*layout\s*\([^)]*\)\s*", "", src, flags=re.M)
    src = _FLOAT.sub(lambda m: m.group(1) + "f", src)

    out = []          # finished text
    acc = ""          # text being collected: a file-scope statement (depth 0) or a block body (depth > 0)
    depth = 0
    is_struct = False
    for line in src.split("\n"):
        if line.lstrip().startswith("#"):
            if depth == 0:
                out.append(acc); acc = ""
                out.append(line + "\n")
            else:
                acc += line + "\n"
            continue
        for ch in line:
            if ch == "{":
                if depth == 0:
                    is_struct = re.match(r"\s*struct\b", acc) is not None
                    out.append(acc + ch); acc = ""
                else:
                    acc += ch
                depth += 1
            elif ch == "}":
                depth -= 1
                if depth == 0:
                    out.append((_MEMBER.sub(r"\1{};", acc) if is_struct else acc) + ch); acc = ""
                else:
                    acc += ch
            elif ch == ";" and depth == 0:
                out.append(_toplevel_statement(acc + ch)); acc = ""
            else:
                acc += ch
        acc += "\n"
    out.append(acc)
    src = "".join(out)

    # ---- argument evaluation order ----
    # GLSL evaluates call arguments left to right (GLSL 3.30 §6.1.1); C++ leaves the order open and g++ goes right to
    # left.  It only matters where two arguments have side effects: calls that draw from the RNG twice
    # (pathtrace.glsl:405).  Hoist those draws, in order, into temporaries declared in front of the statement.
    lines = src.split("\n")
    for i, line in enumerate(lines):
        if len(re.findall(r"\brand\(\)", line)) < 2:
            continue
        prev = next((l.strip() for l in reversed(lines[:i]) if l.strip()), "")
        if prev.endswith(")") or prev.endswith("else"):
            raise SystemExit("glsl2cpp: multi-rand() statement is the body of a brace-less control statement: " + line.strip())
        k = [0]
        def hoist(m):
            k[0] += 1
            return "glsl_rand_arg%d" % (k[0] - 1)
        body = re.sub(r"\brand\(\)", hoist, line)
        indent = line[: len(line) - len(line.lstrip())]
        decl = " ".join("float glsl_rand_arg%d = rand();" % j for j in range(k[0]))
        lines[i] = indent + decl + "\n" + body
    src = "\n".join(lines)

    # ---- parameter qualifiers ----
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*(?:void)?\s*\)", "void glsl_main()", src)
    return src


def main(argv):
    if len(argv) != 4:
        print("usage: glsl2cpp.py <shader_dir> <shader.glsl> <out.inc>", file=sys.stderr)
        return 2
    root, name, out = argv[1], argv[2], argv[3]
    src = load_with_includes(os.path.join(root, name), root)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with open(out, "w") as f:
        f.write("// GENERATED from %s by oracle/glsl_ref/glsl2cpp.py — do not commit.\n" % os.path.join(root, name))
        f.write(translate(src))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
