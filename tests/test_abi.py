"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
the ctypes binding covers them all, and the product package never routes through oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fairrec_b200.h")
PKG = os.path.join(ROOT, "recbole-fairrec_b200")
LIB = os.path.join(PKG, "libfairrec_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = declared_functions()
    for must in ("fr_focf_forward", "fr_focf_backward", "fr_focf_train_step", "fr_fullsort_topk", "fr_topk_merge",
                 "fr_item_group_stats", "fr_sort_pairs_u32"):
        assert must in names


@pytest.mark.skipif(not os.path.exists(LIB), reason="library not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/fairrec_b200.h but not exported"
    lib.fr_abi_version.restype = ctypes.c_int
    assert lib.fr_abi_version() == 1


def test_binding_covers_header():
    import recbole_fairrec_b200 as pkg
    assert sorted(pkg._lib.SIGNATURES) == declared_functions()
    # struct mirrors: same field count/order as the header
    src = open(HEADER).read()
    for cname, pystruct in (("fr_focf_step", pkg._lib.FocfStep), ("fr_fullsort", pkg._lib.FullSort),
                            ("fr_focf_shard_step", pkg._lib.FocfShardStep), ("fr_chain_layer", pkg._lib.ChainLayer),
                            ("fr_chain", pkg._lib.Chain)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"^(const\s+)?[a-z0-9_]+\s+", "", decl)
            fields += [re.sub(r"\[.*\]$", "", n.strip().lstrip("*").strip()) for n in names.split(",")]
        assert fields == [f[0] for f in pystruct._fields_], cname


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libfairrec_oracle" not in txt and "oracle/_build" not in txt, f  # never dlopen/link the checker


def test_no_cpu_fallback():
    """CPU tensors are refused instead of silently computed on the host."""
    import torch
    import recbole_fairrec_b200 as pkg
    with pytest.raises(pkg._lib.FairRecLibraryError):
        pkg._lib.ptr(torch.zeros(4))
    cfg = pkg.Config(embedding_size=8, device=torch.device("cpu"))
    from recbole_fairrec_b200.synth import SynthDataset
    model = pkg.FOCF(cfg, SynthDataset(10, 12, 5.0))
    inter = pkg.Interaction({"user_id": torch.tensor([1, 2]), "item_id": torch.tensor([1, 1]),
                             "rating": torch.tensor([3.0, 4.0]), "gender": torch.tensor([1, 2])})
    with pytest.raises(pkg._lib.FairRecLibraryError):
        model.calculate_loss(inter)


MIRRORS = (("fr_focf_step", "FocfStep"), ("fr_focf_shard_step", "FocfShardStep"), ("fr_mlp_tower", "MlpTower"),
           ("fr_nfcf_step", "NfcfStep"), ("fr_fullsort", "FullSort"), ("fr_adam_entry", "AdamEntry"),
           ("fr_spmm_plan", "SpmmPlan"), ("fr_chain_layer", "ChainLayer"), ("fr_chain", "Chain"), ("fr_chain_dp", "ChainDp"))


def test_struct_mirrors_have_the_c_compilers_size_and_field_offsets(tmp_path):
    """The header compiles as plain C (what a cgo / JNI / ctypes binding consumes) and every ctypes mirror of a `struct fr_*`
    has the size and the per-field offsets gcc gives the C struct -- field names alone would not notice a padding mismatch."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    import recbole_fairrec_b200 as pkg
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fairrec_b200.h"', 'int main(void) {']
    for cname, pyname in MIRRORS:
        lines.append(f'  printf("{cname} . %zu\\n", sizeof({cname}));')
        for field, _ in getattr(pkg._lib, pyname)._fields_:
            lines.append(f'  printf("{cname} {field} %zu\\n", offsetof({cname}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    want = {}
    for ln in out:
        if ln.strip():
            s, f, v = ln.split()
            want[(s, f)] = int(v)
    n = 0
    for cname, pyname in MIRRORS:
        cls = getattr(pkg._lib, pyname)
        assert ctypes.sizeof(cls) == want[(cname, ".")], cname
        for field, _ in cls._fields_:
            assert getattr(cls, field).offset == want[(cname, field)], (cname, field)
            n += 1
    assert n > 150


def test_ctypes_signatures_have_the_headers_parameter_counts():
    """every `fr_*` prototype of the header against the ctypes table: same number of parameters, pointer parameters bound as
    pointers, `size_t` returns bound as c_size_t"""
    import recbole_fairrec_b200 as pkg
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = re.findall(r"\b(size_t|int|void|int64_t|uint64_t|const char \*)\s*(fr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    seen = 0
    for ret, name, params in protos:
        restype, argtypes = pkg._lib.SIGNATURES[name]
        ps = [p.strip() for p in params.split(",") if p.strip() and p.strip() != "void"]
        assert len(ps) == len(argtypes), (name, len(ps), len(argtypes))
        for p, t in zip(ps, argtypes):
            is_ptr_c = "*" in p
            is_ptr_py = t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or hasattr(t, "_type_") and isinstance(
                getattr(t, "_type_", None), type)
            assert is_ptr_c == bool(is_ptr_py), (name, p, t)
        if ret == "size_t":
            assert restype is ctypes.c_size_t, name
        seen += 1
    assert seen == len(declared_functions())
