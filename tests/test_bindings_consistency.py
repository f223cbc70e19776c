"""Static consistency of the three descriptions of the drop-in boundary: the C header (include/kissmcmc_cuda.h),
the ctypes binding that the tests run (kissmcmc.jl_b200/_lib.py) and the Julia `ccall` module that a KissMCMC.jl
maintainer adds (kissmcmc.jl_b200/julia/CUDABackend.jl, unexecuted here: Julia is not in the image)."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "kissmcmc_cuda.h").read_text()
JULIA = (ROOT / "kissmcmc.jl_b200" / "julia" / "CUDABackend.jl").read_text()

C_TO_CTYPES = {"int64_t": C.c_int64, "uint64_t": C.c_uint64, "int32_t": C.c_int32, "double": C.c_double}
C_TO_JULIA = {"int64_t": "Int64", "uint64_t": "UInt64", "int32_t": "Int32", "double": "Float64"}


def header_struct_fields():
    body = re.search(r"typedef struct kmc_emcee_opts \{(.*?)\} kmc_emcee_opts;", HEADER, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return re.findall(r"\b(int64_t|uint64_t|int32_t|double)\s+(\w+)\s*;", body)


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return {m.group(2): m.group(3) for m in
            re.finditer(r"^(int32_t|const char \*)\s*(kmc_\w+)\(([^;]*?)\);", text, re.M | re.S)}


def test_opts_struct_is_the_same_in_c_ctypes_and_julia(km):
    fields = header_struct_fields()
    assert len(fields) == 16
    opts = km.EmceeOpts if hasattr(km, "EmceeOpts") else km._lib.EmceeOpts
    assert [(n, t) for n, t in opts._fields_] == [(name, C_TO_CTYPES[ct]) for ct, name in fields]
    assert C.sizeof(opts) == 96                                       # naturally aligned, no padding
    jl = re.search(r"struct EmceeOpts\n(.*?)\nend", JULIA, re.S).group(1)
    jl_fields = re.findall(r"^\s*(\w+)::(\w+)\s*$", jl, re.M)
    assert jl_fields == [(name, C_TO_JULIA[ct]) for ct, name in fields]


def test_every_julia_ccall_targets_a_declared_function_with_matching_arity():
    funcs = header_functions()
    calls = re.findall(r"ccall\(\(:(kmc_\w+), LIB\[\]\),\s*(\w+),\s*\(([^)]*)\)", JULIA, re.S)
    assert calls, "no ccall found in CUDABackend.jl"
    for name, ret, argt in calls:
        assert name in funcs, f"{name} is not declared in include/kissmcmc_cuda.h"
        nargs_c = 0 if funcs[name].strip() in ("", "void") else funcs[name].count(",") + 1
        jl_args = [a for a in (s.strip() for s in argt.split(",")) if a]
        assert len(jl_args) == nargs_c, f"{name}: {len(jl_args)} ccall argument types, {nargs_c} C parameters"
        assert ret == ("Cstring" if name == "kmc_last_error" else "Int32")
    # the stubs the reference-facing API needs are all present
    need = {"kmc_density_create", "kmc_density_eval", "kmc_emcee_create", "kmc_emcee_run", "kmc_emcee_copy_results",
            "kmc_emcee_nsamples", "kmc_emcee_destroy", "kmc_density_destroy", "kmc_last_error",
            # round 2: plugin options, library-owned multi-GPU, the device-side rows either side of the sampler
            "kmc_density_set_option", "kmc_density_get_info", "kmc_emcee_create_multi", "kmc_multi_run", "kmc_multi_sync",
            "kmc_multi_shape", "kmc_multi_copy_results", "kmc_multi_destroy", "kmc_emcee_squash", "kmc_make_theta0s",
            "kmc_g_pdf", "kmc_cdf_g_inv", "kmc_sample_g"}
    assert need <= {c[0] for c in calls}


def test_julia_pointer_and_scalar_argument_kinds_match_the_header():
    """Pointer parameters of the header are Ptr/Ref/Cstring in the ccall signature, scalars are scalars."""
    funcs = header_functions()
    for name, _, argt in re.findall(r"ccall\(\(:(kmc_\w+), LIB\[\]\),\s*(\w+),\s*\(([^)]*)\)", JULIA, re.S):
        c_params = [p.strip() for p in funcs[name].split(",")] if funcs[name].strip() not in ("", "void") else []
        jl_args = [a for a in (s.strip() for s in argt.split(",")) if a]
        for cp, ja in zip(c_params, jl_args):
            is_ptr_c = "*" in cp or cp.split()[0] in ("kmc_density_t", "kmc_sampler_t", "kmc_multi_t")
            is_ptr_j = ja.startswith(("Ptr{", "Ref{")) or ja == "Cstring"
            assert is_ptr_c == is_ptr_j, f"{name}: C parameter `{cp}` vs Julia `{ja}`"
            if not is_ptr_c:
                assert C_TO_JULIA[cp.split()[0]] == ja, f"{name}: C parameter `{cp}` vs Julia `{ja}`"


def test_header_cites_the_reference_for_every_entry_point_group():
    """Every block of the header names the reference lines it replaces (src/samplers.jl:NNN or the :NNN shorthand)."""
    assert HEADER.count("src/samplers.jl") >= 5
    assert len(re.findall(r"[(:, ]:\d{3}", HEADER)) + HEADER.count("src/samplers.jl:") >= 15


def test_every_header_symbol_is_bound_by_ctypes_and_exported(km):
    """include/*.h <-> the ctypes SYMBOLS list <-> the shared library's export table (no compute calls)."""
    funcs = header_functions()
    assert set(funcs) == set(km.SYMBOLS), set(funcs) ^ set(km.SYMBOLS)
    for name in funcs:
        assert hasattr(km.lib, name), name
