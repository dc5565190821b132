"""CPU-only static checks of the Fortran shim (shim/crm_module_b200.F90), which cannot be compiled here (no Fortran compiler in the image
or on the GPU box): its bind(C) mirror types must list exactly the members of the C structs of include/crm_b200.h in the same order and
with matching types, and every variable whose address is taken with c_loc() must be a TARGET (gfortran rejects c_loc
of a non-TARGET dummy)."""
import os
import re

import common as cm

SHIM = os.path.join(cm.ROOT, "shim", "crm_module_b200.F90")
HDR = os.path.join(cm.ROOT, "include", "crm_b200.h")


def _c_structs():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\w*\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        mem = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(?:const\s+)?(long long|int|double)\s+(.*)", decl)
            ctype, rest = m.group(1), m.group(2)
            for v in rest.split(","):
                v = v.strip()
                ptr = v.startswith("*")
                mem.append((v.lstrip("* "), "ptr" if ptr else ctype))
        out[name] = mem
    return out


def _f_types():
    src = open(SHIM).read()
    out = {}
    for name, body in re.findall(r"type,\s*bind\(C\)\s*::\s*(\w+)\s*\n(.*?)end type", src, flags=re.S | re.I):
        mem = []
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if not line:
                continue
            ftype, names = line.split("::")
            ftype = ftype.strip().lower()
            kind = {"integer(c_int)": "int", "real(c_double)": "double", "type(c_ptr)": "ptr", "integer(c_long_long)": "long long"}[ftype]
            mem += [(n.strip(), kind) for n in names.split(",")]
        out[name] = mem
    return out


def test_bind_c_types_mirror_the_header_structs():
    c, f = _c_structs(), _f_types()
    for name in ("crm_dims_t", "crm_input_t", "crm_state_t", "crm_rad_t", "crm_output_t"):
        assert name in c and name in f, name
        assert [m for m, _ in f[name]] == [m for m, _ in c[name]], name
        assert [k for _, k in f[name]] == [k for _, k in c[name]], name


def test_every_c_loc_argument_is_a_target():
    src = open(SHIM).read()
    body = src[src.index("subroutine crm("):src.index("end subroutine crm")]
    targets = set()
    for line in body.splitlines():
        code = line.split("!")[0]
        if "::" in code and re.search(r"\btarget\b", code.split("::")[0], flags=re.I):
            targets |= {re.sub(r"\(.*", "", n).strip() for n in code.split("::")[1].split(",")}
    used = {m.split("%")[0] for m in re.findall(r"c_loc\(\s*([\w%]+)\s*\)", body)}
    assert used and used <= targets, used - targets
    assert {"crm_input", "crm_state", "crm_rad", "crm_output"} <= used


def test_every_struct_member_is_assigned_before_the_run_call():
    src = open(SHIM).read()
    body = src[src.index("subroutine crm("):src.index("rc = crm_b200_run(")]
    f = _f_types()
    for var, tname in (("d", "crm_dims_t"), ("ci", "crm_input_t"), ("cs", "crm_state_t"), ("cr", "crm_rad_t"), ("co", "crm_output_t")):
        for member, _ in f[tname]:
            assert re.search(r"\b%s%%%s\s*=" % (var, member), body), (var, member)


def _c_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(crm_b200_\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        out[name] = args
    return out


def test_every_bind_c_interface_matches_a_header_prototype():
    """Each interface block of the shim names a function the header declares, with the same number of arguments, `value` on
    exactly the by-value C parameters (int lchnk / int device / void *handle) and references for the struct pointers."""
    src = open(SHIM).read()
    protos = _c_prototypes()
    blocks = re.findall(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)(.*?)end function", src, flags=re.S | re.I)
    assert len(blocks) >= 6
    for fname, fargs, cname, body in blocks:
        assert fname == cname and cname in protos, cname
        fargs = [a.strip() for a in fargs.split(",") if a.strip()]
        assert len(fargs) == len(protos[cname]), (cname, fargs, protos[cname])
        for fa, ca in zip(fargs, protos[cname]):
            decl = [ln for ln in body.splitlines() if re.search(r"::\s*%s\b" % fa, ln)]
            assert decl, (cname, fa)
            by_value = "value" in decl[0].split("::")[0].lower()
            c_by_value = ("*" not in ca) or (ca.startswith("void *") and not ca.startswith("void **"))
            assert by_value == c_by_value, (cname, fa, ca)


def test_sync_host_assigns_every_member_and_uses_targets():
    src = open(SHIM).read()
    body = src[src.index("subroutine crm_b200_sync_host("):src.index("end subroutine crm_b200_sync_host")]
    f = _f_types()
    for var, tname in (("d", "crm_dims_t"), ("cs", "crm_state_t"), ("cr", "crm_rad_t")):
        for member, _ in f[tname]:
            assert re.search(r"\b%s%%%s\s*=" % (var, member), body), (var, member)
    used = {m.split("%")[0] for m in re.findall(r"c_loc\(\s*([\w%]+)\s*\)", body)}
    assert used == {"crm_state"}
    assert re.search(r"type\(crm_state_type\),\s*target", body)
