"""Static checks of the Julia host file against the C header (no Julia toolchain exists in this image).

hedgehog.jl_b200/julia/HedgehogB200.jl binds libhedgehog_mc.so with `ccall`; a wrong field order or argument type there is
silent memory corruption. This test parses every `struct HH...` block and every `ccall` signature of the Julia file and
compares them with the typedefs and prototypes of include/hedgehog_mc.h: field names, order and types; entry-point names,
return types, arity and argument types; the numeric constants both files define."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "hedgehog_mc.h")).read()
JULIA = open(os.path.join(ROOT, "hedgehog.jl_b200", "julia", "HedgehogB200.jl")).read()

STRUCTS = {"hh_model": "HHModel", "hh_bk_config": "HHBkConfig", "hh_sim": "HHSim", "hh_payoff": "HHPayoff",
           "hh_result": "HHResult", "hh_tangent": "HHTangent", "hh_lsm_result": "HHLsmResult", "hh_comm": "HHComm",
           "hh_path_payoff": "HHPathPayoff"}
SCALARS = {"int32_t": {"Int32", "Cint"}, "uint32_t": {"UInt32", "Cuint"}, "int64_t": {"Int64"}, "uint64_t": {"UInt64"},
           "double": {"Float64", "Cdouble"}, "int": {"Cint", "Int32"}, "size_t": {"Csize_t"},
           "hh_allreduce_fn": {"Ptr{Cvoid}"}}
POINTEES = {"uint64_t": "UInt64", "double": "Float64", "int32_t": "Int32", "void": "Cvoid", "unsigned char": "UInt8",
            "char": "UInt8", "hh_ctx": "Cvoid", "size_t": "Csize_t"}
POINTEES.update(STRUCTS)


def strip_comments(c):
    c = re.sub(r"/\*.*?\*/", " ", c, flags=re.S)
    return re.sub(r"\\\n", " ", c)


def c_struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), strip_comments(HEADER), flags=re.S)
    assert m, f"typedef struct {name} not found in the header"
    out = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        mm = re.match(r"(const )?([A-Za-z_ 0-9]+?) ?(\*?) ?([A-Za-z_0-9]+(?: ?, ?[A-Za-z_0-9]+)*)$", decl)
        assert mm, f"cannot parse the C declaration {decl!r} in {name}"
        ctype, star, names = mm.group(2).strip(), mm.group(3), mm.group(4)
        for nm in names.split(","):
            out.append((nm.strip(), ctype + ("*" if star else "")))
    return out


def julia_struct_fields(name):
    m = re.search(r"^struct %s\b[^\n]*\n(.*?)^end" % name, JULIA, flags=re.S | re.M)
    assert m, f"struct {name} not found in the Julia file"
    out = []
    for line in m.group(1).splitlines():
        line = line.split("#")[0].strip()
        for part in filter(None, (x.strip() for x in line.split(";"))):
            nm, ty = part.split("::")
            out.append((nm.strip(), ty.strip()))
    return out


def julia_accepts(ctype, jtype):
    """Is Julia type `jtype` a correct image of C type `ctype` (as a struct field or a ccall argument)?"""
    if ctype.endswith("**"):
        return jtype in ("Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}")
    if ctype.endswith("*"):
        base = ctype[:-1].strip()
        if base == "char":
            return jtype in ("Cstring", "Ptr{UInt8}")
        want = POINTEES[base]
        jn = jtype.replace("Cdouble", "Float64").replace("Cint", "Int32")   # Julia aliases of the same bits types
        return jn in (f"Ptr{{{want}}}", f"Ref{{{want}}}")
    if ctype in STRUCTS:
        return jtype == STRUCTS[ctype]
    return jtype in SCALARS[ctype]


@pytest.mark.parametrize("cname", sorted(STRUCTS))
def test_struct_layout_matches_header(cname):
    cf, jf = c_struct_fields(cname), julia_struct_fields(STRUCTS[cname])
    assert [n for n, _ in jf] == [n for n, _ in cf], f"{STRUCTS[cname]}: field names / order differ from {cname}"
    for (n, ct), (_, jt) in zip(cf, jf):
        assert julia_accepts(ct, jt), f"{STRUCTS[cname]}.{n}: Julia type {jt} does not mirror C type {ct}"


def c_prototypes():
    protos = {}
    for m in re.finditer(r"^(const char \*|int|void) ?(hh_[a-z0-9_]+)\((.*?)\);", strip_comments(HEADER), flags=re.S | re.M):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        types = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\[[A-Z_0-9]*\]$", "*", a)                       # array parameter = pointer
                mm = re.match(r"(const )?([A-Za-z_ 0-9]+?) ?(\*{0,2}) ?([A-Za-z_0-9]+)(\*?)$", a)
                assert mm, f"cannot parse the parameter {a!r} of {name}"
                types.append(mm.group(2).strip() + mm.group(3) + mm.group(5))
        protos[name] = (ret, types)
    return protos


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(hh_[a-z0-9_]+), LIB\[\]\), ?([A-Za-z]+), ?\n?\s*\(", JULIA):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(JULIA[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), split_top(" ".join(JULIA[m.end():i - 1].split()))))
    return calls


def test_every_ccall_matches_its_prototype():
    protos, calls = c_prototypes(), julia_ccalls()
    assert len(calls) >= 12
    for name, ret, argtypes in calls:
        assert name in protos, f"ccall of {name}: no such entry point in the header"
        cret, ctypes_ = protos[name]
        assert (ret == "Cstring") if cret.startswith("const char") else (ret == {"int": "Cint", "void": "Cvoid"}[cret]), \
            f"{name}: return type {ret} vs C {cret}"
        assert len(argtypes) == len(ctypes_), f"{name}: {len(argtypes)} ccall argument types, the prototype has {len(ctypes_)}"
        for k, (ct, jt) in enumerate(zip(ctypes_, argtypes)):
            assert julia_accepts(ct, jt), f"{name}, argument {k + 1}: Julia type {jt} does not mirror C type {ct}"


def test_the_product_entry_points_are_bound():
    bound = {c[0] for c in julia_ccalls()}
    for name in ("hh_version", "hh_create", "hh_destroy", "hh_last_error", "hh_mc_european", "hh_mc_european_tangent_sums",
                 "hh_lsm_american", "hh_mc_path_dependent", "hh_peer_export", "hh_peer_connect", "hh_peer_disconnect",
                 "hh_peer_set_timeout"):
        assert name in bound, f"{name} is not bound by the Julia host file"


def test_constants_agree():
    defs = dict(re.findall(r"#define (HH_[A-Z0-9_]+) \(?(-?[0-9]+)u?\)?", HEADER))
    consts = {}
    for m in re.finditer(r"^const ([A-Z0-9_, ]+?) ?= ?(.+)$", JULIA, flags=re.M):
        names, vals = [n.strip() for n in m.group(1).split(",")], split_top(m.group(2).split("#")[0])
        if len(names) == len(vals):
            for n, v in zip(names, vals):
                mm = re.fullmatch(r"(?:Cint|UInt32|Int32)?\(?(-?[0-9]+)\)?", v.strip())
                if n.startswith("HH_") and mm:
                    consts[n] = mm.group(1)
    assert len(consts) >= 12
    for n, v in consts.items():
        assert n in defs and defs[n] == v, f"{n} = {v} in the Julia file, {defs.get(n)} in the header"


def test_reference_return_shapes_by_reading():
    """ForwardAD through Hedgehog's generic solve returns (greek = deriv,) (greeks_problem.jl:261): the host file must not
    override solve(::GreekProblem, ::ForwardAD, ...); BatchGreekProblem returns a Dict lens => greek (:559-568); the
    solution constructors take the reference's positional fields (pricing_solutions.jl:22-27, 78-84)."""
    assert not re.search(r"solve\(gprob::GreekProblem", JULIA)
    assert "occursin" not in JULIA
    assert re.search(r"Dict\(lens => partials\(price, p\)", JULIA)
    assert re.search(r"MonteCarloSolution\(prob, method, prices\[1\], ensemble\)", JULIA)
    assert re.search(r"LSMSolution\(prob, method, [^,]+, stopping_info, spot\)", JULIA)
