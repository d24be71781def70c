"""julia/LandHydrologyB200.jl cannot be executed here (no Julia in the image).  What can be checked statically is checked:
its struct mirrors against include/lh_soil.h field for field (name, order, C type), and every ccall against the header's
prototypes (exported symbol, return type, argument count and pointer/scalar kinds)."""
import os
import re

import __graft_entry__ as graft

JL = os.path.join(graft.ROOT, "julia", "LandHydrologyB200.jl")
HDR = os.path.join(graft.ROOT, "include", "lh_soil.h")

C2JL = {"double": "Cdouble", "int32_t": "Int32", "int64_t": "Int64", "const double*": "Ptr{Cdouble}", "double*": "Ptr{Cdouble}",
        "lh_soil_params": "LhSoilParams", "lh_soil_face_bc": "LhSoilFaceBC"}
STRUCTS = {"lh_soil_params": "LhSoilParams", "lh_soil_face_bc": "LhSoilFaceBC", "lh_soil_config": "LhSoilConfig",
           "lh_soil_run_opts": "LhSoilRunOpts", "lh_soil_stepper": "LhSoilStepper", "lh_soil_atmos": "LhSoilAtmos"}


def header_source():
    return re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)


def c_struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header_source(), flags=re.S)
    assert m, name
    out = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        mm = re.match(r"(const double\*|double\*|[a-z0-9_]+)\s*(.*)", decl)
        ctype, rest = mm.group(1), mm.group(2)
        for item in rest.split(","):
            item = item.strip()
            ptr = item.startswith("*")
            item = item.lstrip("* ")
            arr = re.match(r"(\w+)\[(\w+)\]", item)
            t = ctype + ("*" if ptr else "")
            if arr:
                out.append((arr.group(1), f"NTuple{{{arr.group(2)}, {C2JL[t]}}}"))
            else:
                out.append((item, C2JL[t]))
    return out


def jl_struct_fields(name):
    src = open(JL).read()
    m = re.search(r"^struct %s\n(.*?)^end" % name, src, flags=re.S | re.M)
    assert m, name
    out = []
    for item in re.split(r"[;\n]", m.group(1)):
        item = item.split("#")[0].strip()
        if not item:
            continue
        fname, ftype = item.split("::")
        out.append((fname.strip(), ftype.strip()))
    return out


def test_struct_mirrors_match_the_header_field_for_field():
    for cname, jname in STRUCTS.items():
        assert jl_struct_fields(jname) == c_struct_fields(cname), (cname, jname)


def c_prototypes():
    protos = {}
    for m in re.finditer(r"\b(int32_t|int64_t|const char\*)\s+(lh_soil_\w+)\s*\(([^)]*)\)\s*;", header_source()):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        protos[name] = (ret, [] if args in ("", "void") else [a.strip() for a in args.split(",")])
    return protos


def test_every_ccall_matches_a_prototype():
    src = open(JL).read()
    protos = c_prototypes()
    calls = re.findall(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(((?:[^()]|\([^()]*\))*?)\)\s*,", src, flags=re.S)
    assert len(calls) >= 20
    ret_map = {"int32_t": "Int32", "int64_t": "Int64", "const char*": "Cstring"}
    seen = set()
    for name, ret, args in calls:
        assert name in protos, f"ccall of {name}: not declared in include/lh_soil.h"
        cret, cargs = protos[name]
        assert ret == ret_map[cret], (name, ret, cret)
        jargs = [a.strip() for a in re.split(r",(?![^{]*\})", args) if a.strip()]
        assert len(jargs) == len(cargs), (name, jargs, cargs)
        for ja, ca in zip(jargs, cargs):
            is_ptr_c = "*" in ca or "[" in ca
            is_ptr_j = ja.startswith(("Ptr{", "Ref{")) or ja == "Cstring"
            assert is_ptr_c == is_ptr_j, (name, ja, ca)
            if not is_ptr_c:
                ctype = ca.split()[0]
                assert ja == {"double": "Cdouble", "int32_t": "Int32", "int64_t": "Int64"}[ctype], (name, ja, ca)
        seen.add(name)
    # the calls a drop-in needs are all there
    for need in ("lh_soil_create", "lh_soil_destroy", "lh_soil_set_state", "lh_soil_get_state", "lh_soil_rhs", "lh_soil_get_tendency",
                 "lh_soil_step_ssprk33", "lh_soil_step", "lh_soil_run", "lh_soil_set_aux", "lh_soil_set_aux_table", "lh_soil_set_bc_values",
                 "lh_soil_set_column_params", "lh_soil_comm_unique_id", "lh_soil_comm_init", "lh_soil_budgets_allreduce"):
        assert need in seen, need


def test_no_runtime_symbol_in_ccall():
    """ccall needs a literal (name, library) pair: no `ccall((fn, LIB), ...)` with a variable."""
    src = open(JL).read()
    assert not re.search(r"ccall\(\(\s*[a-z_]\w*\s*,\s*LIB\)", src)
