"""tools/gen_fortran_iface.py NAME [NAME ...] -- ISO_C_BINDING interface blocks (free-form Fortran, also valid as fixed form) for
functions declared in include/roms_b200.h: int by value, `const double*` -> real(c_double), intent(in) :: x(*), `double*` ->
intent(inout), `int*` -> integer(c_int) array, roms_b200_ctx* -> type(c_ptr), value.  Used to keep
roms_b200/fortran/roms_b200_mod.F90 in step with the header (tests/test_cpu.py checks the coverage)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hdr = open(os.path.join(ROOT, "include", "roms_b200.h")).read()
hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)


OPTIONAL = {"dndx", "dmde", "rhoA", "rhoS", "srflx", "ghats"}    # arguments that exist only under cpp options one application lacks


def wrap(prefix, items, indent="     &    "):
    lines, cur = [], prefix
    for n, it in enumerate(items):
        piece = it + (", " if n + 1 < len(items) else "")
        if len(cur) + len(piece) > 66:
            lines.append(cur.ljust(70) + "&")
            cur = indent
        cur += piece
    lines.append(cur)
    return lines


for name in sys.argv[1:]:
    m = re.search(r"\bint\s+" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S)
    assert m, name
    args = [a.strip() for a in " ".join(m.group(1).split()).split(",")]
    names, decl = [], {}
    for a in args:
        nm = re.findall(r"(\w+)\s*$", a)[0]
        names.append(nm)
        if "roms_b200_ctx" in a:
            decl.setdefault("type(c_ptr), value", []).append(nm)
        elif re.match(r"const double\s*\*", a) and name.endswith("_tile") and nm in OPTIONAL:
            decl.setdefault("real(c_double), intent(in), optional", []).append(nm + "(*)")     # absent -> null pointer (F2018 18.3.6)
        elif re.match(r"const double\s*\*", a):
            decl.setdefault("real(c_double), intent(in)", []).append(nm + "(*)")
        elif re.match(r"double\s*\*", a):
            decl.setdefault("real(c_double), intent(inout)", []).append(nm + "(*)")
        elif re.match(r"(const )?int\s*\*", a):
            decl.setdefault("integer(c_int)", []).append(nm + "(*)")
        elif re.match(r"int\b", a):
            decl.setdefault("integer(c_int), value", []).append(nm)
        else:
            raise SystemExit("unhandled argument: " + a)
    out = wrap("        integer(c_int) FUNCTION %s (" % name, names)
    out[-1] += ")"
    out[-1] = out[-1].ljust(70) + "&"
    out.append("     &                          BIND(C, name='%s')" % name)
    out.append("          IMPORT")
    for t, ns in decl.items():
        out += wrap("          %s :: " % t, ns, "     &      ")
    out.append("        END FUNCTION")
    print("\n".join(out))
