#!/usr/bin/env python
"""Golden vectors derived from the reference's SOURCE TEXT (run in the build container; /root/reference does not
exist on the GPU box, so the vectors are committed as tests/golden/ref_*.npz next to this script).

The D reference cannot be compiled here (no D toolchain) and no reference test asserts a JPEG pixel, a converted
PixelType value or a QOIX predictor value. To pin the C oracle (oracle/*.c) to the reference rather than to a
same-author restatement, this script reads the reference's own function bodies from
/root/reference/source/gamut/**.d, transliterates their pure-expression statements to Python mechanically
(regular expressions; no arithmetic is retyped by hand), executes them with exact integer / IEEE-single
semantics (numpy int32 wrap-around, numpy float32 scalars), and stores input -> output vectors:

  ref_jpeg_idct.npz      Row!N.idct / Col!N.idct / idct / idct_4x4      jpegload.d:156-397 (sparse paths chosen
                         by s_idct_row_table / s_idct_col_table exactly as the reference does)
  ref_jpeg_upsample.npz  DCT_Upsample.P_Q!(R,C).calc / R_S!(R,C).calc, add/sub_and_store, s_max_rc dispatch,
                         idct_4x4 (transform_mcu_expand)                 jpegload.d:886-1073, 2132-2251
  ref_jpeg_colour.npz    create_look_ups + the H1V1Convert pixel expression  jpegload.d:2079-2094, 2536-2549
  ref_scanline.npz       every scanline_convert_* body of scanline.d:139-803 (46 functions)
  ref_predictors.npz     locoPredict (qoiplane10.d:84-96), locoIntraPredictionSIMD (qoi2avg.d:863-897,
                         qoi10b.d:871-903; SSE intrinsics evaluated with Intel's documented lane semantics)

Control flow that is not a pure expression (the `switch` dispatch of idct / transform_mcu_expand, the per-pixel
`for` loops) is mirrored by hand below, each place citing the lines it follows; the tables those dispatches read
are parsed from the text. If the reference text changes shape the transliterator raises instead of guessing.

tests/test_oracle_reference_text.py asserts EXACT equality of the oracle with these vectors.
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

REF = "/root/reference/source/gamut"
OUT = os.path.dirname(os.path.abspath(__file__))

i32 = np.int32


def lines_of(rel):
    return open(os.path.join(REF, rel), encoding="utf-8").read().split("\n")


def find_line(lines, needle, start=0):
    for i in range(start, len(lines)):
        if needle in lines[i]:
            return i
    raise RuntimeError("reference text changed: %r not found" % needle)


def body_after(lines, start):
    """Lines of the brace block that opens at or after `start` (exclusive of the outer braces)."""
    depth, out, i, opened = 0, [], start, False
    while i < len(lines):
        ln = lines[i]
        code = ln.split("//")[0]
        for ch in code:
            if ch == "{":
                depth += 1
                opened = True
            elif ch == "}":
                depth -= 1
        if opened:
            out.append(ln)
        if opened and depth == 0:
            break
        i += 1
    # strip the outer braces
    first = out[0]
    out[0] = first[first.index("{") + 1:]
    last = out[-1]
    out[-1] = last[:last.rindex("}")]
    return out, i


# ------------------------------------------------------------------------------------------------------------------
# D statement -> Python statement (only the shapes that occur in the transliterated bodies)
CASTS = {"int": "i32_", "ubyte": "u8_", "ushort": "u16_", "short": "i16_", "jpgd_block_t": "i16_", "float": "f32_",
         "byte": "i8_"}
TYPES = r"(?:immutable\s+|const\s+)?(?:int|float|ubyte|ushort|short|byte|Temp_Type|uint)"


def split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def d_expr(e: str) -> str:
    e = e.strip()
    e = re.sub(r"mixin\((\w+)!\(([^()]*)\)\)", r"\1(\2)", e)          # mixin(AT!(1, 0)) -> AT(1, 0)
    e = re.sub(r"mixin\((\w+)!(\d+)\)", r"\1(\2)", e)                  # mixin(ACCESS_COL!2) -> ACCESS_COL(2)
    e = re.sub(r"\bF!\((-?[\d.]+)f\)", r"F(\1)", e)                    # F!(0.415735f) -> F(0.415735)
    e = re.sub(r"\bFIX!\((-?[\d.]+)f\)", r"FIX(\1)", e)
    e = re.sub(r"(?<![\w.])(\d+\.\d+)f\b", r"f32_(\1)", e)             # 255.0f -> f32_(255.0)
    e = re.sub(r"(\w+)\.at\(([^,()]+),\s*([^()]+?)\)", r"\1[\2][\3]", e)  # P.at(r, 0) -> P[r][0]
    e = re.sub(r"\*(\w+)\+\+", r"\1.next()", e)                        # *s++ (read)
    e = re.sub(r"(\w+)\.ptr\[", r"\1[", e)                             # m_crr.ptr[cr] -> m_crr[cr]
    for t, f in CASTS.items():                                         # cast(T)(...) / cast(T)ident[...] / cast(T)ident(...)
        e = e.replace("cast(%s)(" % t, f + "(")
        e = re.sub(r"cast\(%s\)\s*(\w+(?:\[[^\]]*\]|\([^()]*\))?)" % t, f + r"(\1)", e)
    if "cast(" in e or "mixin(" in e or "!" in e.replace("!=", ""):
        raise RuntimeError("untranslated D expression: " + e)
    return e


def d_stmt(s: str):
    """One D statement (without the trailing ';') -> list of Python statements."""
    s = s.strip()
    if not s:
        return []
    m = re.match(r"^%s\s+(.*)$" % TYPES, s)
    decl_type = None
    if m:
        decl_type = re.match(r"^(?:immutable\s+|const\s+)?(\w+)", s).group(1)
        s = m.group(1)
        out = []
        for part in split_top(s):
            name, _, rhs = part.partition("=")
            rhs = d_expr(rhs)
            if "?" in rhs:
                c, _, rest = rhs.partition("?")
                a, _, b = rest.partition(":")
                rhs = "(%s if %s else %s)" % (a.strip(), c.strip(), b.strip())
            wrap = {"float": "f32_", "ubyte": "u8_", "ushort": "u16_", "short": "i16_", "byte": "i8_"}.get(decl_type)
            out.append("%s = %s" % (name.strip(), "%s(%s)" % (wrap, rhs) if wrap else rhs))
        return out
    m = re.match(r"^\*(\w+)\+\+\s*=\s*(.*)$", s)                        # *outp++ = expr
    if m:
        return ["%s.put(%s)" % (m.group(1), d_expr(m.group(2)))]
    m = re.match(r"^(\w+)\s*/=\s*(.*)$", s)                             # r /= a   (float)
    if m:
        return ["%s = f32_(%s / %s)" % (m.group(1), m.group(1), d_expr(m.group(2)))]
    m = re.match(r"^return\s+(.*)$", s)
    if m:
        return ["return " + d_expr(m.group(1))]
    m = re.match(r"^([\w\[\]\.\*\+\(\), ]+?)\s*=\s*(.*)$", s)           # lvalue = expr
    if m and not m.group(1).strip().startswith("if"):
        return ["%s = %s" % (d_expr(m.group(1)), d_expr(m.group(2)))]
    raise RuntimeError("untranslated D statement: " + s)


def d_block(lines, indent="    "):
    """C-like block (assignments, if / else if / else with or without braces, return) -> Python source."""
    text = "\n".join(ln.split("//")[0] for ln in lines)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    toks = re.findall(r"\bif\s*\([^{};]*?\)(?=\s*[\w{*])|\belse\b|\{|\}|(?!\s)(?!if\b)(?!else\b)[^{};]+;", text)
    rebuilt = "".join(toks)
    if re.sub(r"\s+", "", rebuilt) != re.sub(r"\s+", "", text):
        raise RuntimeError("block tokeniser lost text:\n" + text)
    py, depth, pending = [], 0, []        # pending: stack of "single statement" indents to close

    def emit(s):
        py.append(indent * (depth + 1) + s)

    i = 0
    single = []                            # depths opened by brace-less if/else
    while i < len(toks):
        t = toks[i].strip()
        if t.startswith("if"):
            cond = d_expr(t[t.index("(") + 1:t.rindex(")")])
            emit("if %s:" % cond)
            if i + 1 < len(toks) and toks[i + 1].strip() == "{":
                depth += 1
                single.append(False)
                i += 2
                continue
            depth += 1
            single.append(True)
        elif t == "else":
            nxt = toks[i + 1].strip()
            if nxt.startswith("if"):
                cond = d_expr(nxt[nxt.index("(") + 1:nxt.rindex(")")])
                emit("elif %s:" % cond)
                i += 1
            else:
                emit("else:")
            if toks[i + 1].strip() == "{":
                depth += 1
                single.append(False)
                i += 2
                continue
            depth += 1
            single.append(True)
        elif t == "{":
            raise RuntimeError("unexpected block")
        elif t == "}":
            depth -= 1
            single.pop()
            while single and single[-1]:
                depth -= 1
                single.pop()
        else:
            for p in d_stmt(t[:-1]):
                emit(p)
            while single and single[-1]:
                depth -= 1
                single.pop()
        i += 1
    return "\n".join(py)


# helpers visible to the transliterated code ---------------------------------------------------------------------
def f32_(x): return np.float32(x)
def i32_(x): return np.asarray(x).astype(np.int32) if isinstance(x, np.ndarray) else int(x)
def u8_(x): return (np.asarray(x).astype(np.int64) & 0xFF).astype(np.uint8) if isinstance(x, np.ndarray) else int(x) & 0xFF
def u16_(x): return int(x) & 0xFFFF
def i16_(x): return np.asarray(x).astype(np.int16) if isinstance(x, np.ndarray) else ((int(x) + 0x8000) & 0xFFFF) - 0x8000
def i8_(x): return ((int(x) + 0x80) & 0xFF) - 0x80


BASE_ENV = {"f32_": f32_, "i32_": i32_, "u8_": u8_, "u16_": u16_, "i16_": i16_, "i8_": i8_, "np": np}


# ==================================================================================================================
# JPEG: IDCT, DCT_Upsample, colour
# ==================================================================================================================
def jpeg_env():
    L = lines_of("codecs/jpegload.d")
    env = dict(BASE_ENV)
    cites = {}
    # constants: enum NAME = cast(int)VALUE;  (jpegload.d:122-135)
    for ln in L[:200]:
        m = re.match(r"^enum (CONST_BITS|PASS1_BITS|SCALEDONE|FIX_\w+) = (?:cast\(int\))?(\d+);", ln)
        if m:
            env[m.group(1)] = i32(int(m.group(2)))
    for k in ("CONST_BITS", "PASS1_BITS", "SCALEDONE", "FIX_0_298631336", "FIX_3_072711026"):
        assert k in env, k
    # DESCALE / DESCALE_ZEROSHIFT: `return <expr>;` lines are valid Python expressions (jpegload.d:137-145)
    i = find_line(L, "int DESCALE() (int x, int n)")
    env["DESCALE"] = eval("lambda x, n: " + re.search(r"return (.*);", "\n".join(L[i:i + 4])).group(1), env)
    i = find_line(L, "int DESCALE_ZEROSHIFT() (int x, int n)")
    env["DESCALE_ZEROSHIFT"] = eval("lambda x, n: " + re.search(r"return (.*);", "\n".join(L[i:i + 4])).group(1), env)
    i = find_line(L, "ubyte CLAMP() (int i)")
    clamp_txt = re.sub(r"\s+", " ", " ".join(L[i:i + 6]))
    assert "if (i < 0) i = 0; if (i > 255) i = 255; return cast(ubyte)i;" in clamp_txt, clamp_txt   # jpegload.d:147-152
    env["CLAMP"] = lambda v: np.clip(v, 0, 255)
    cites["constants"] = "jpegload.d:119-152"

    # Row!(N).idct general body (the `else` branch after NONZERO_COLS == 1)  jpegload.d:156-214
    i = find_line(L, "struct Row(int NONZERO_COLS)")
    j = find_line(L, "immutable int z2 = mixin(ACCESS_COL!2)", i)
    k = find_line(L, "pTemp[4] = DESCALE(tmp13 - btmp0", j)
    row_src = d_block(L[j:k + 1])
    # the NONZERO_COLS == 1 shortcut, jpegload.d:163-172
    j1 = find_line(L, "immutable int dcval = (pSrc[0] << PASS1_BITS);", i)
    row1_src = d_block(L[j1:j1 + 9])
    cites["Row"] = "jpegload.d:%d-%d" % (i + 1, k + 1)
    # Col!(N).idct  jpegload.d:218-292
    i = find_line(L, "struct Col (int NONZERO_ROWS)")
    j1 = find_line(L, "int dcval = DESCALE_ZEROSHIFT(pTemp[0], PASS1_BITS+3);", i)
    col1_src = d_block(L[j1:j1 + 10])
    j = find_line(L, "immutable int z2 = mixin(ACCESS_ROW!2);", i)
    k = find_line(L, "pDst_ptr[8*4] = cast(ubyte)CLAMP(i);", j)
    col_src = d_block(L[j:k + 1])
    cites["Col"] = "jpegload.d:%d-%d" % (i + 1, k + 1)

    def compile_fn(name, args, src):
        code = "def %s(%s):\n%s\n" % (name, args, src)
        exec(code, env)
        return env[name]

    compile_fn("_row_general", "pTemp, pSrc, ACCESS_COL", row_src)
    compile_fn("_row_one", "pTemp, pSrc", row1_src)
    compile_fn("_col_general", "pDst_ptr, pTemp, ACCESS_ROW", col_src)
    compile_fn("_col_one", "pDst_ptr, pTemp", col1_src)

    class V:
        """pointer-like view: V(arr (N, K), off)[i] -> arr[:, off + i] as int32 (reads), assignment writes through"""
        def __init__(self, a, off=0): self.a, self.off = a, off
        def __getitem__(self, i): return self.a[:, self.off + int(i)].astype(np.int32)
        def __setitem__(self, i, v): self.a[:, self.off + int(i)] = v
        def __add__(self, n): return V(self.a, self.off + n)
    env["V"] = V

    def Row(N, pTemp, pSrc):                      # template instantiation Row!(N)  (static if chain, :159-176)
        if N == 0:
            return
        if N == 1:
            return env["_row_one"](pTemp, pSrc)
        zero = np.zeros(pSrc.a.shape[0], np.int32)
        env["_row_general"](pTemp, pSrc, lambda x: pSrc[x] if x < N else zero)       # ACCESS_COL, :174-176

    def Col(N, pDst, pTemp):                      # Col!(N)  (:221-237)
        assert N > 0
        if N == 1:
            return env["_col_one"](pDst, pTemp)
        zero = np.zeros(pTemp.a.shape[0], np.int32)
        env["_col_general"](pDst, pTemp, lambda x: pTemp[x * 8] if x < N else zero)  # ACCESS_ROW, :239-241

    # tables, parsed from the text
    def table(name):
        i = find_line(L, name)
        txt = ""
        while "];" not in L[i]:
            txt += L[i]
            i += 1
        txt += L[i]
        return [int(x) for x in re.findall(r"\d+", txt[txt.index("=") + 1:])]
    row_table = table("static immutable ubyte[512] s_idct_row_table")
    col_table = table("static immutable ubyte[64] s_idct_col_table")
    max_rc = table("static immutable ubyte[64] s_max_rc")
    assert len(row_table) == 512 and len(col_table) == 64 and len(max_rc) == 64
    cites["tables"] = "jpegload.d:295-306, 2132-2137"

    def idct(src, max_zag):
        """idct (jpegload.d:308-376) for ONE max_zag over a batch: src (N, 64) int16 -> (N, 64) uint8."""
        n = src.shape[0]
        dst = np.zeros((n, 64), np.uint8)
        if max_zag <= 1:                            # :312-326
            i = find_line(L, "int k = ((pSrc_ptr[0] + 4) >> 3) + 128;")
            k = eval(d_expr(re.search(r"int k = (.*);", L[i]).group(1)), {"pSrc_ptr": V(src)})
            dst[:, :] = np.clip(k, 0, 255).astype(np.uint8)[:, None]
            return dst
        temp = np.zeros((n, 64), np.int32)
        for r in range(8):                          # :334-353
            Row(row_table[(max_zag - 1) * 8 + r], V(temp, r * 8), V(src, r * 8))
        nz = col_table[max_zag - 1]                 # :357
        for c in range(8):                          # :358-375
            Col(nz, V(dst, c), V(temp, c))
        return dst

    def idct_4x4(src):                              # jpegload.d:378-397
        n = src.shape[0]
        dst = np.zeros((n, 64), np.uint8)
        temp = np.zeros((n, 64), np.int32)
        for r in range(4):
            Row(4, V(temp, r * 8), V(src, r * 8))
        for c in range(8):
            Col(4, V(dst, c), V(temp, c))
        return dst

    # ---- DCT_Upsample (jpegload.d:827-1073)
    i = find_line(L, "enum FRACT_BITS = ")
    env["FRACT_BITS"] = int(re.search(r"= (\d+);", L[i]).group(1))
    assert "enum SCALE = 1 << FRACT_BITS;" in L[i + 1]
    env["SCALE"] = 1 << env["FRACT_BITS"]
    i = find_line(L, "static int D(T) (T i)")
    env["D"] = eval("lambda i: " + re.search(r"return (.*); }", L[i]).group(1), env)
    i = find_line(L, "enum F(float i) = ")
    assert "(cast(int)((i) * SCALE + 0.5f))" in L[i], L[i]
    env["F"] = lambda x: i32(int(np.float32(np.float32(x) * np.float32(env["SCALE"])) + np.float32(0.5)))   # float32 ops, cast(int) truncates
    i = find_line(L, "static struct P_Q(int NUM_ROWS, int NUM_COLS)")
    j = find_line(L, "immutable Temp_Type X000 = mixin(AT!(0, 0));", i)
    k = find_line(L, "Q.at(3, 3) = X036;", j)
    pq_src = d_block(L[j:k + 1])
    cites["P_Q"] = "jpegload.d:%d-%d" % (i + 1, k + 1)
    i = find_line(L, "static struct R_S(int NUM_ROWS, int NUM_COLS)")
    j = find_line(L, "immutable Temp_Type X100 = D(F!(0.906127f)", i)
    k = find_line(L, "S.at(3, 3) = X136;", j)
    rs_src = d_block(L[j:k + 1])
    cites["R_S"] = "jpegload.d:%d-%d" % (i + 1, k + 1)
    compile_fn("_pq", "P, Q, AT", pq_src)
    compile_fn("_rs", "R, S, AT", rs_src)
    # add_and_store / sub_and_store bodies (:886-902)
    i = find_line(L, "static void add_and_store() (jpgd_block_t* pDst, in Matrix44 a, in Matrix44 b)")
    add_src = d_block(L[i + 2:i + 6])
    i = find_line(L, "static void sub_and_store() (jpgd_block_t* pDst, in Matrix44 a, in Matrix44 b)")
    sub_src = d_block(L[i + 2:i + 6])
    compile_fn("_add_store", "pDst, a, b, r", add_src)
    compile_fn("_sub_store", "pDst, a, b, r", sub_src)
    cites["store"] = "jpegload.d:886-902"
    # the dispatch must list exactly these (rows, cols) pairs (:2164-2227)
    i = find_line(L, "switch (s_max_rc.ptr[max_zag])")
    cases = re.findall(r"case (\d)\*16\+(\d):\s*DCT_Upsample\.P_Q!\((\d), (\d)\)", "\n".join(L[i:i + 70]))
    assert len(cases) == 15 and all(a == c and b == d for a, b, c, d in cases), cases
    valid_rc = {int(a) * 16 + int(b) for a, b, _, _ in cases}

    def upsample(src, max_zag_m1):
        """Chroma part of transform_mcu_expand (jpegload.d:2150-2251) for one `max_zag` (already minus 1):
        src (N, 64) int16 -> (N, 4, 64) uint8."""
        n = src.shape[0]
        rc = max_rc[max_zag_m1]
        assert rc in valid_rc
        NR, NC = rc >> 4, rc & 15
        zero = np.zeros(n, np.int32)
        AT = lambda c, r: zero if (c >= NC or r >= NR) else src[:, c + r * 8].astype(np.int32)   # :909-911
        mk = lambda: [[None] * 4 for _ in range(4)]
        P, Q, R, S = mk(), mk(), mk(), mk()
        env["_pq"](P, Q, AT)
        env["_rs"](R, S, AT)
        add = lambda a, b: [[a[r][c] + b[r][c] for c in range(4)] for r in range(4)]
        sub = lambda a, b: [[a[r][c] - b[r][c] for c in range(4)] for r in range(4)]
        a = add(P, Q); b = sub(P, Q); c = add(R, S); d = sub(R, S)          # :2229-2234
        out = np.zeros((n, 4, 64), np.uint8)
        for t, (fn, x, y) in enumerate((("_add_store", a, c), ("_sub_store", a, c), ("_add_store", b, d), ("_sub_store", b, d))):
            temp_block = np.zeros((n, 64), np.int16)                         # jpgd_block_t[64] temp_block (zero-init), :2148
            for r in range(4):                                               # foreach (int r; 0..4)
                env[fn](V16(temp_block), x, y, r)
            out[:, t] = idct_4x4(temp_block)                                 # :2236-2250
        return out

    class V16:
        def __init__(self, a): self.a = a
        def __setitem__(self, i, v): self.a[:, int(i)] = np.asarray(v).astype(np.int16)
    env["V16"] = V16

    # ---- colour (jpegload.d:2079-2094 + the pixel expression of H1V1Convert :2536-2549)
    i = find_line(L, "enum SCALEBITS = ")
    env["SCALEBITS"] = int(re.search(r"= (\d+);", L[i]).group(1))
    assert "enum ONE_HALF = (cast(int) 1 << (SCALEBITS-1));" in L[i + 1]
    env["ONE_HALF"] = 1 << (env["SCALEBITS"] - 1)
    assert "enum FIX(float x) = (cast(int)((x) * (1L<<SCALEBITS) + 0.5f));" in L[i + 2], L[i + 2]
    env["FIX"] = lambda x: int(np.float32(np.float32(x) * np.float32(1 << env["SCALEBITS"])) + np.float32(0.5))
    i = find_line(L, "void create_look_ups ()")
    body, _ = body_after(L, i)
    j = next(q for q, ln in enumerate(body) if "int k = i - 128;" in ln)
    lut_src = d_block(body[j:j + 5])
    compile_fn("_lut_step", "i, m_crr, m_cbb, m_crg, m_cbg", lut_src)
    crr, cbb, crg, cbg = ([0] * 256 for _ in range(4))
    for q in range(256):
        env["_lut_step"](q, crr, cbb, crg, cbg)
    i = find_line(L, "void H1V1Convert ()")
    j = find_line(L, "__m128i A = _mm_setr_epi32(y + m_crr.ptr[cr],", i)
    txt = " ".join(x.strip() for x in L[j:j + 4])
    m = re.search(r"_mm_setr_epi32\((.*),\s*255\);", txt)
    exprs = [d_expr(x) for x in split_top(m.group(1))]
    assert len(exprs) == 3, exprs
    assert "_mm_packs_epi32(A, zero)" in L[j + 4] and "_mm_packus_epi16(A, zero)" in L[j + 5]
    cites["colour"] = "jpegload.d:2079-2094, %d-%d" % (j + 1, j + 6)

    def ycc_to_rgb(y, cb, cr):
        loc = {"m_crr": crr, "m_cbb": cbb, "m_crg": crg, "m_cbg": cbg, "y": int(y), "cb": int(cb), "cr": int(cr)}
        v = [eval(x, {}, loc) for x in exprs]
        v = [max(-32768, min(32767, q)) for q in v]       # _mm_packs_epi32: signed saturate to int16
        return [max(0, min(255, q)) for q in v]           # _mm_packus_epi16: unsigned saturate to u8

    return dict(idct=idct, idct_4x4=idct_4x4, upsample=upsample, ycc_to_rgb=ycc_to_rgb, tables=(crr, cbb, crg, cbg),
                cites=cites, zag=table("static immutable int[64] g_ZAG"))


def gen_jpeg():
    J = jpeg_env()
    zag = J["zag"]
    rng = np.random.default_rng(20260101)
    # ---- idct: for every max_zag 1..64, blocks whose coefficients live in zig-zag positions < max_zag
    blocks, mz, outs = [], [], []
    for max_zag in range(1, 65):
        n = 24
        b = np.zeros((n, 64), np.int16)
        for q in range(n):
            amp = (4, 24, 100, 400, 12000)[q % 5]
            vals = rng.integers(-amp, amp + 1, max_zag)
            if q % 3 == 0:
                vals[rng.random(max_zag) < 0.6] = 0                 # sparse, like real blocks
            if q == n - 1:
                vals[:] = amp                                       # saturating block
            vals[max_zag - 1] = vals[max_zag - 1] or 1              # the last coefficient defines max_zag
            for k in range(max_zag):
                b[q, zag[k]] = vals[k]
            b[q, 0] = rng.integers(-1000, 1000) if amp < 1000 else rng.integers(-8192, 8192)   # dequantised DC of 8-bit JPEG: [-1024, 1016]
        blocks.append(b)
        mz.append(np.full(n, max_zag, np.int32))
        with np.errstate(over="ignore"):
            outs.append(J["idct"](b, max_zag))
    np.savez_compressed(os.path.join(OUT, "ref_jpeg_idct.npz"), blocks=np.concatenate(blocks), max_zag=np.concatenate(mz),
                        out=np.concatenate(outs), cite=np.array(str(J["cites"])))
    # ---- upsample
    blocks, mz, outs = [], [], []
    for max_zag in range(1, 65):
        n = 16
        b = np.zeros((n, 64), np.int16)
        for q in range(n):
            amp = (4, 24, 100, 3000)[q % 4]
            vals = rng.integers(-amp, amp + 1, max_zag)
            if q % 2 == 0:
                vals[rng.random(max_zag) < 0.5] = 0
            vals[max_zag - 1] = vals[max_zag - 1] or -1
            for k in range(max_zag):
                b[q, zag[k]] = vals[k]
            b[q, 0] = rng.integers(-1000, 1000) if amp < 1000 else rng.integers(-8192, 8192)   # dequantised DC of 8-bit JPEG: [-1024, 1016]
        blocks.append(b)
        mz.append(np.full(n, max_zag, np.int32))
        with np.errstate(over="ignore"):
            outs.append(J["upsample"](b, max_zag - 1))
    np.savez_compressed(os.path.join(OUT, "ref_jpeg_upsample.npz"), blocks=np.concatenate(blocks), max_zag=np.concatenate(mz),
                        out=np.concatenate(outs), cite=np.array(str(J["cites"])))
    # ---- colour: the four tables and 4096 random + all-corner triples
    tri = [(y, cb, cr) for y in (0, 1, 127, 128, 254, 255) for cb in (0, 1, 127, 128, 129, 255) for cr in (0, 1, 127, 128, 129, 255)]
    tri += [tuple(int(v) for v in rng.integers(0, 256, 3)) for _ in range(4096)]
    tri = np.array(tri, np.int32)
    rgb = np.array([J["ycc_to_rgb"](*t) for t in tri], np.uint8)
    crr, cbb, crg, cbg = (np.array(t, np.int32) for t in J["tables"])
    np.savez_compressed(os.path.join(OUT, "ref_jpeg_colour.npz"), ycc=tri, rgb=rgb, crr=crr, cbb=cbb, crg=crg, cbg=cbg,
                        cite=np.array(J["cites"]["colour"]))
    return J["cites"]


# ==================================================================================================================
# scanline.d converters
# ==================================================================================================================
class Reader:
    def __init__(self, arr):
        self.a, self.pos = arr, 0

    def next(self):
        v = self[self.pos]
        self.pos += 1
        return v

    def __getitem__(self, i):
        v = self.a[int(i)]
        return np.float32(v) if self.a.dtype == np.float32 else int(v)      # D: ubyte/ushort promote to int


class Writer:
    def __init__(self, dtype):
        self.dtype, self.out, self.idx = dtype, [], {}

    def put(self, v):
        self.out.append(v)

    def __setitem__(self, i, v):
        self.idx[int(i)] = v

    def result(self):
        if self.idx:
            assert not self.out
            self.out = [self.idx[i] for i in range(len(self.idx))]
        if self.dtype == np.float32:
            return np.array([np.float32(v) for v in self.out], np.float32)
        return np.array([int(v) for v in self.out]).astype(self.dtype)       # stores into ubyte/ushort truncate


DT = {"ubyte": np.uint8, "ushort": np.uint16, "float": np.float32}


def gen_scanline():
    L = lines_of("scanline.d")
    rng = np.random.default_rng(20260102)
    W = 331
    fns = [i for i, ln in enumerate(L) if re.match(r"^void scanline_convert_(\w+)_to_(\w+)\(const\(ubyte\)\* inScan, ubyte\* outScan, int width", ln)]
    store = {}
    names = []
    SIZES = {"l8": (1, "ubyte"), "l16": (1, "ushort"), "lf32": (1, "float"), "la8": (2, "ubyte"), "la16": (2, "ushort"),
             "laf32": (2, "float"), "lap8": (2, "ubyte"), "lap16": (2, "ushort"), "lapf32": (2, "float"),
             "rgb8": (3, "ubyte"), "rgb16": (3, "ushort"), "rgbf32": (3, "float"), "rgba8": (4, "ubyte"),
             "rgba16": (4, "ushort"), "rgbaf32": (4, "float"), "rgbap8": (4, "ubyte"), "rgbap16": (4, "ushort"),
             "rgbapf32": (4, "float"), "bgra8": (4, "ubyte"), "bgr8": (3, "ubyte")}
    for i in fns:
        m = re.match(r"^void scanline_convert_(\w+?)_to_(\w+)\(", L[i])
        src_t, dst_t = m.group(1), m.group(2)
        if src_t not in SIZES or dst_t not in SIZES:
            raise RuntimeError("unknown pixel type in " + L[i])
        body, end = body_after(L, i)
        text = "\n".join(body)
        sch, sel = SIZES[src_t]
        dch, del_ = SIZES[dst_t]
        # input row: all small values / extremes first, then random; floats in [0, 1] (casts of
        # out-of-range floats are undefined in the reference)
        n = W * sch
        if sel == "float":
            x = rng.random(n).astype(np.float32)
            x[:256] = (np.arange(256) / np.float32(255.0)).astype(np.float32)[:min(256, n)]
            x[256:260] = [0.0, 1.0, 0.5, np.float32(1e-8)]
            if src_t in ("lapf32", "rgbapf32"):           # premultiplied input: colour <= alpha keeps results in range
                px = x.reshape(W, sch)
                px[:, :sch - 1] = px[:, :sch - 1] * px[:, sch - 1:sch]
                px[5, sch - 1] = 0.0                      # alpha == 0 branch
                px[5, :sch - 1] = 0.25
        else:
            mx = 255 if sel == "ubyte" else 65535
            x = rng.integers(0, mx + 1, n).astype(DT[sel])
            x[:min(n, 512)] = (np.arange(min(n, 512)) * (1 if sel == "ubyte" else 129)) % (mx + 1)
            if src_t in ("lap8", "rgbap8", "lap16", "rgbap16"):
                px = x.reshape(W, sch)
                px[7, sch - 1] = 0                        # alpha == 0 branch
        if "memcpy(outScan, inScan" in text:              # the copy converters (rgb8->rgb8, rgba8->rgba8, rgbaf32->rgbaf32)
            assert re.search(r"memcpy\(outScan, inScan, width \* \d( \* \w+\.sizeof)?\);", text), text
            out = x.copy()
        else:
            # prologue: pointer declarations
            env = dict(BASE_ENV)
            env["inScan"] = Reader(x.view(np.uint8)) if sel != "ubyte" else Reader(x)
            env["outScan"] = Writer(np.uint8)
            writers = ["outScan"]
            loop_at = next(q for q, ln in enumerate(body) if re.match(r"\s*for \(int x = 0; x < width; \+\+x\)", ln))
            for ln in body[:loop_at]:
                ln = ln.split("//")[0].strip()
                if not ln or ln == "version(DigitalMars) pragma(inline, false);":      # a compiler hint, no semantics
                    continue
                m1 = re.match(r"^const\((\w+)\)\*\s*(\w+) = (?:cast\(const\(\w+\)\*\)\s*)?inScan;$", ln)
                m2 = re.match(r"^(\w+)\*\s*(\w+) = (?:cast\(\w+\*\)\s*)?outScan;$", ln)
                if m1:
                    assert m1.group(1) == sel, (ln, sel)
                    env[m1.group(2)] = Reader(x)
                elif m2:
                    assert m2.group(1) == del_, (ln, del_)
                    env[m2.group(2)] = Writer(DT[del_])
                    writers.append(m2.group(2))
                else:
                    raise RuntimeError("untranslated prologue in %s: %s" % (L[i], ln))
            if sel != "ubyte" and not any(isinstance(v, Reader) and v.a is x for v in env.values()):
                raise RuntimeError("no typed input pointer in " + L[i])
            loop_body, _ = body_after(body, loop_at)
            src = d_block(loop_body)
            code = "def _px(x):\n" + src + "\n"
            exec(code, env)
            for xx in range(W):
                env["_px"](xx)
            used = [w for w in writers if env[w].out or env[w].idx]
            assert len(used) == 1, (L[i], used)
            wr = env[used[0]]
            if used[0] == "outScan":
                wr.dtype = DT[del_] if del_ == "ubyte" else wr.dtype
            out = wr.result()
            assert out.size == W * dch, (L[i], out.size)
        name = "%s_to_%s" % (src_t, dst_t)
        names.append(name)
        store["in_" + name] = x
        store["out_" + name] = out
        store["line_" + name] = np.array(i + 1)
    store["names"] = np.array(names)
    store["width"] = np.array(W)
    np.savez_compressed(os.path.join(OUT, "ref_scanline.npz"), **store)
    return names


# ==================================================================================================================
# predictors
# ==================================================================================================================
def _sat16(v):
    return np.clip(v, -32768, 32767).astype(np.int16)


SSE = {
    "_mm_setzero_si128": lambda: np.zeros(8, np.int16),
    "_mm_unpacklo_epi8": lambda a, z: a,                  # bytes already widened by _mm_loadu_si32 below (zero-extended)
    "_mm_add_epi16": lambda a, b: (a.astype(np.int32) + b).astype(np.int16),
    "_mm_sub_epi16": lambda a, b: (a.astype(np.int32) - b).astype(np.int16),
    "_mm_max_epi16": lambda a, b: np.maximum(a, b),
    "_mm_min_epi16": lambda a, b: np.minimum(a, b),
    "_mm_cmple_epi16": lambda a, b: np.where(a <= b, np.int16(-1), np.int16(0)),     # qoi2avg.d:899-902 (lt | eq)
    "_mm_cmpge_epi16": lambda a, b: np.where(a >= b, np.int16(-1), np.int16(0)),     # qoi2avg.d:904-907 (gt | eq)
    "_mm_packus_epi16": lambda a, z: np.clip(a, 0, 255).astype(np.int16),             # unsigned saturation to u8
    "_mm_set1_epi16": lambda v: np.full(8, v, np.int16),
}


def simd_fn(lines, start_marker, cite):
    i = find_line(lines, start_marker)
    body, end = body_after(lines, i)
    env = dict(SSE)
    env["np"] = np
    py = []
    for ln in body:
        ln = ln.split("//")[0].strip()
        if not ln:
            continue
        m = re.match(r"^__m128i (\w+) = _mm_loadu_si(32|64)\(&(\w)\);$", ln)
        if m:
            py.append("%s = %s" % (m.group(1), m.group(3)))
            continue
        m = re.match(r"^(?:__m128i )?(\w+) = (.*);$", ln)
        if m and "_mm_store" not in ln:
            rhs = m.group(2).replace("~", "~")
            py.append("%s = %s" % (m.group(1), rhs))
            continue
        if re.match(r"^(RGBA|qoi10_rgba_t) r;$", ln) or "_mm_storeu_si32(&r, P)" in ln or "_mm_storel_epi64(cast(__m128i*)&r, P)" in ln:
            continue
        if ln == "return r;":
            py.append("return P")
            continue
        raise RuntimeError("untranslated SIMD line (%s): %s" % (cite, ln))
    code = "def _f(a, b, c):\n" + "\n".join("    " + p for p in py) + "\n"
    exec(code, env)
    return env["_f"], "%s:%d-%d" % (cite, i + 1, end + 1)


def gen_predictors():
    rng = np.random.default_rng(20260103)
    # locoPredict, qoiplane10.d:84-96
    L = lines_of("codecs/qoiplane10.d")
    i = find_line(L, "int locoPredict(int left, int top, int topleft)")
    body, end = body_after(L, i)
    env = dict(BASE_ENV)
    exec("def locoPredict(left, top, topleft):\n" + d_block(body) + "\n", env)
    tri = rng.integers(0, 1024, (6000, 3))
    tri[:1000] = rng.integers(0, 8, (1000, 3)) * 146                 # many ties
    tri[1000:1200] = rng.integers(1000, 1024, (200, 3))
    p10 = np.array([env["locoPredict"](int(a), int(b), int(c)) for a, b, c in tri], np.int32)
    # locoIntraPredictionSIMD 8-bit RGBA, qoi2avg.d:863-897
    L2 = lines_of("codecs/qoi2avg.d")
    f8, c8 = simd_fn(L2, "static RGBA locoIntraPredictionSIMD(RGBA a, RGBA b, RGBA c)", "qoi2avg.d")
    q = rng.integers(0, 256, (6000, 3, 4))
    q[:1500] = rng.integers(0, 6, (1500, 3, 4)) * 51
    p8 = np.zeros((6000, 4), np.uint8)
    for n in range(6000):
        lanes = [np.zeros(8, np.int16) for _ in range(3)]
        for t in range(3):
            lanes[t][:4] = q[n, t]
        p8[n] = f8(*lanes)[:4].astype(np.uint8)
    # locoIntraPredictionSIMD 10-bit, qoi10b.d:871-903
    L3 = lines_of("codecs/qoi10b.d")
    f10, c10 = simd_fn(L3, "static qoi10_rgba_t locoIntraPredictionSIMD(qoi10_rgba_t a, qoi10_rgba_t b, qoi10_rgba_t c)", "qoi10b.d")
    q10 = rng.integers(0, 1024, (6000, 3, 4))
    q10[:1500] = rng.integers(0, 8, (1500, 3, 4)) * 146
    p10b = np.zeros((6000, 4), np.uint16)
    for n in range(6000):
        lanes = [np.zeros(8, np.int16) for _ in range(3)]
        for t in range(3):
            lanes[t][:4] = q10[n, t]
        p10b[n] = f10(*lanes)[:4].astype(np.uint16)
    np.savez_compressed(os.path.join(OUT, "ref_predictors.npz"), loco10_in=tri.astype(np.int32), loco10_out=p10,
                        loco8_in=q.astype(np.uint8), loco8_out=p8, loco10b_in=q10.astype(np.uint16), loco10b_out=p10b,
                        cite=np.array("qoiplane10.d:%d-%d; %s; %s" % (i + 1, end + 1, c8, c10)))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("gen_from_reference.py: %s not present (run it in the build container)" % REF)
    print("jpeg:", gen_jpeg())
    print("scanline:", len(gen_scanline()), "functions")
    gen_predictors()
    print("predictors ok")
