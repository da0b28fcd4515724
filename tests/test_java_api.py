"""The Java side of the drop-in boundary (source only: no JDK in this image).

  * java/physics/b200/IncFluidB200.java binds only symbols include/incfluid_b200.h declares, and its Panama struct
    layouts (ifl_config, ifl_step_status, ifl_body) have the field order, sizes and offsets of the ctypes mirror the
    whole test-suite drives the library with;
  * the eight program classes under java/physics/{EulerianFluids,hybridFluids}/ keep the reference's package, class and
    nested FluidSolver signatures (constructors, update, addInflow, toImage, maxTimestep, ambientT), so that the
    reference's main() programs compile against them unchanged.  The reference signatures are pinned in
    tests/golden/java_signatures.json (written by this file's `python tests/test_java_api.py --regen` from
    /root/reference/src) and re-derived from the reference itself whenever it is present.
"""
import ctypes as C
import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "physics")
REF = "/root/reference/src/physics"
GOLDEN = os.path.join(ROOT, "tests", "golden", "java_signatures.json")
PROGRAMS = {
    "EulerianFluids/ExplicitFluid1matrixless.java": 1, "EulerianFluids/ExplicitFluid2betterAdvection.java": 2,
    "EulerianFluids/ExplicitFluid3conjugateGradients.java": 3, "EulerianFluids/ExplicitFluid4solidBoundaries.java": 4,
    "EulerianFluids/ExplicitFluid5curvedBoundaries.java": 5, "EulerianFluids/ExplicitFluid6heat.java": 6,
    "EulerianFluids/ExplicitFluid7variableDensity.java": 7, "hybridFluids/ExplicitFluid8flip.java": 8,
}
METHODS = ("update", "addInflow", "toImage", "maxTimestep", "ambientT")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def solver_signatures(path):
    """package, outer class and the FluidSolver surface of one program file, whitespace-normalised"""
    text = strip_comments(open(path, encoding="utf-8", errors="replace").read())
    pkg = re.search(r"package\s+([\w.]+)\s*;", text).group(1)
    outer = re.search(r"\bclass\s+(ExplicitFluid\w+)", text).group(1)
    start = text.index("static class FluidSolver")
    nxt = re.search(r"\n    static class \w+", text[start + 10:])          # the next nested class of the same level, if any
    body = text[start: start + 10 + nxt.start()] if nxt else text[start:]
    norm = lambda s: re.sub(r"\s+", " ", s).strip()
    params = lambda s: [norm(p) for p in s.split(",")] if s.strip() else []
    sig = {"package": pkg, "class": outer, "constructors": [], "methods": {}}
    for m in re.finditer(r"\bFluidSolver\s*\(([^)]*)\)\s*\{", body):
        sig["constructors"].append([re.sub(r"\s+\w+$", "", p) for p in params(m.group(1))])         # parameter TYPES
    for m in re.finditer(r"\b(void|double)\s+(%s)\s*\(([^)]*)\)\s*(throws\s+\w+)?\s*\{" % "|".join(METHODS), body):
        sig["methods"][m.group(2)] = {"returns": m.group(1), "params": [re.sub(r"\s+\w+$", "", p) for p in params(m.group(3))],
                                      "throws": norm(m.group(4) or "")}
    return sig


def test_binding_binds_only_declared_symbols():
    hdr = strip_comments(open(os.path.join(ROOT, "include", "incfluid_b200.h")).read())
    declared = set(re.findall(r"\b(ifl_[a-z_0-9]+)\s*\(", hdr))
    src = open(os.path.join(JAVA, "b200", "IncFluidB200.java")).read()
    bound = set(re.findall(r'fn\("(ifl_[a-z_0-9]+)"', src))
    assert len(bound) >= 12 and bound <= declared, bound - declared


def java_layout(src, name):
    """[(field, size)] of `static final StructLayout NAME = MemoryLayout.structLayout(...)` incl. explicit padding"""
    body = re.search(r"StructLayout %s = MemoryLayout\.structLayout\((.*?)\);\n" % name, src, re.S).group(1)
    out = []
    for m in re.finditer(r"JAVA_(INT|DOUBLE|LONG)\.withName\(\"(\w+)\"\)|MemoryLayout\.paddingLayout\((\d+)\)|"
                         r"MemoryLayout\.sequenceLayout\((\d+), JAVA_BYTE\)\.withName\(\"(\w+)\"\)", body):
        if m.group(1):
            out.append((m.group(2), {"INT": 4, "DOUBLE": 8, "LONG": 8}[m.group(1)]))
        elif m.group(3):
            out.append((None, int(m.group(3))))
        else:
            out.append((m.group(5), int(m.group(4))))
    return out


@pytest.mark.parametrize("jname,cname", [("CONFIG", "Config"), ("STATUS", "StepStatus"), ("BODY", "Body")])
def test_struct_layouts_match_the_ctypes_mirror(ifl, jname, cname):
    src = open(os.path.join(JAVA, "b200", "IncFluidB200.java")).read()
    cls = getattr(ifl.abi, cname)
    off, got = 0, {}
    for field, size in java_layout(src, jname):
        if field is not None:
            got[field] = (off, size)
        off += size
    want = {n: (getattr(cls, n).offset, getattr(cls, n).size) for n, _ in cls._fields_}
    assert got == want
    assert off == C.sizeof(cls)


def ours():
    return {rel: solver_signatures(os.path.join(JAVA, rel)) for rel in PROGRAMS}


def test_program_classes_keep_the_reference_signatures():
    golden = json.load(open(GOLDEN))
    mine = ours()
    for rel, ref in golden.items():
        got = mine[rel]
        assert got["package"] == ref["package"] and got["class"] == ref["class"], rel
        assert got["constructors"] == ref["constructors"], (rel, got["constructors"], ref["constructors"])
        for name, m in ref["methods"].items():
            assert got["methods"].get(name) == m, (rel, name, got["methods"].get(name), m)
        # and every program delegates to the variant of its own number
        text = open(os.path.join(JAVA, rel)).read()
        assert re.search(r"IncFluidB200\.F%d_" % PROGRAMS[rel], text), rel


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_golden_signatures_are_the_reference_s():
    live = {rel: solver_signatures(os.path.join(REF, rel)) for rel in PROGRAMS}
    assert live == json.load(open(GOLDEN))


if __name__ == "__main__" and "--regen" in sys.argv:
    json.dump({rel: solver_signatures(os.path.join(REF, rel)) for rel in PROGRAMS}, open(GOLDEN, "w"), indent=1, sort_keys=True)
    print("wrote", GOLDEN)
