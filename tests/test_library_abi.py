"""The C-ABI library loads and exports every symbol include/aspire_b200.h
declares (no compute calls - there is no GPU here), and refuses to compute
without a device instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "aspire_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    from aspire_b200 import _lib

    names = _declared()
    assert len(names) >= 30
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table mirrors the header
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.asp_version() == 100


def test_no_cpu_fallback():
    import torch

    from aspire_b200 import AspireB200Error, _lib, ops

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.load().asp_device_check() == _lib.ASP_ERR_DEVICE
    with pytest.raises(AspireB200Error):
        ops.device_check()
    with pytest.raises(AspireB200Error):
        ops.reweight(torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                     torch.zeros(4, dtype=torch.float64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aspire_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_struct_layouts_match_the_header(tmp_path):
    """Every struct that crosses the C-ABI: sizeof and the offset of each field as a C compiler lays out
    include/aspire_b200.h (gcc, the header is plain C) against the ctypes mirrors in aspire_b200/_lib.py."""
    import ctypes as C
    import shutil
    import subprocess

    from aspire_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    pairs = {"asp_density_t": _lib.Density, "asp_beta_state_t": _lib.BetaState, "asp_peer_t": _lib.PeerDesc,
             "asp_pcn_params_t": _lib.PcnParams, "asp_flow_t": _lib.Flow, "asp_transform_t": _lib.TransformDesc,
             "asp_flow_train_t": _lib.FlowTrain}
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "aspire_b200.h")).read(), flags=re.S)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "aspire_b200.h"', "int main(void) {"]
    fields = {}
    for cname in pairs:
        tag = cname[:-2]
        body = re.search(r"typedef struct " + tag + r"\s*\{(.*?)\}\s*" + cname + ";", text, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):  # "int32_t rank, world" / "double* box[ASP_MAX_PEERS]"
                names.append(re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[[^\]]*\])?\s*$", part.strip()).group(1))
        fields[cname] = names
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        lines += [f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));' for f in names]
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True,
                   capture_output=True, text=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        mirror = [n for n, *_ in cls._fields_]
        assert mirror == fields[cname], (cname, mirror, fields[cname])
        for f in mirror:
            assert int(out[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_integration_doc_stub_matches_the_abi():
    """The ctypes stub INTEGRATION.md shows a maintainer of the reference (section 2) declares the same argument
    lists as the binding this package uses - and every function it names exists."""
    import ctypes as C

    from aspire_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    env = {"_P": C.c_void_p, "_I64": C.c_int64, "_D": C.c_double, "C": C}
    stubs = re.findall(r"_lib\.(asp_\w+)\.argtypes = \[(.*?)\]", text)
    assert len(stubs) >= 7
    for name, args in stubs:
        assert list(eval("[" + args + "]", {"__builtins__": {}}, env)) == list(_lib.SIGNATURES[name][1]), name  # noqa: S307
    for name in set(re.findall(r"\b(asp_[a-z0-9_]+)\b", text)):
        if name.rstrip("_") != name or name in ("asp_status", "asp_density_t", "asp_peer_t", "asp_transform_t",
                                                "asp_beta_state_t", "asp_pcn_params_t", "asp_flow_t", "asp_flow_train_t"):
            continue
        candidates = [s for s in _lib.SIGNATURES if s == name or s.startswith(name)]
        assert candidates, f"INTEGRATION.md names {name}, which the header does not declare"
