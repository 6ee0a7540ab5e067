"""The C-ABI shared library loads and exports every symbol include/rosnet_b200.h declares
(no compute calls: this runs on a CPU-only box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from rosnet_b200.engine import lib as L

    return L


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "rosnet_b200.h")).read()
    declared = set(re.findall(r"\b(rnb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib.EXPORTED_SYMBOLS)
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name


def test_version_and_error_string(lib):
    h = lib.load()
    assert h.rnb_abi_version() == 1
    assert isinstance(h.rnb_last_error_string(), bytes)


def test_struct_layout_matches_header(lib):
    # sizes implied by the header's field order on LP64
    assert ctypes.sizeof(lib.PermuteDesc) == 8 + 3 * 8 * lib.RNB_MAX_RANK + 16
    assert ctypes.sizeof(lib.ContractDesc) == 16 + 24 + 3 * 32 - 8 + 16 + 24 + 16 + 16
    assert ctypes.sizeof(lib.DeviceProps) == 32


def test_invalid_descriptors_fail_loudly_without_gpu(lib):
    d = lib.PermuteDesc()
    d.elem_bytes = 3
    d.rank = 1
    with pytest.raises(lib.RosnetB200Error, match="elem_bytes"):
        lib.check(lib.load().rnb_permute(ctypes.byref(d), None), "rnb_permute")
    c = lib.ContractDesc()
    c.dtype = 9
    with pytest.raises(lib.RosnetB200Error, match="dtype"):
        lib.check(lib.load().rnb_contract_grouped(ctypes.byref(c), None), "rnb_contract_grouped")


def test_product_path_has_no_cpu_fallback():
    import numpy as np
    import torch

    import rosnet_b200 as rn

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = rn.ones((4, 4), dtype=np.float64, blockshape=(2, 2))
    with pytest.raises(rn.RosnetB200Error):
        np.tensordot(a, a, [(1,), (0,)])
    with pytest.raises(rn.RosnetB200Error):
        np.transpose(a, (1, 0))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rosnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|import_module\([\"']oracle", src, re.M), os.path.join(dirpath, f)


def test_ctypes_structs_match_the_header_compiled_with_gcc(lib, tmp_path):
    """The ctypes mirrors of engine/lib.py against what a C compiler makes of include/rosnet_b200.h: sizes and the
    offsets of the fields the binding writes (a reference-side binding is checked the same way)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    checks = {
        "rnb_permute_desc": (lib.PermuteDesc, ["elem_bytes", "rank", "extent", "src_stride", "dst_stride", "src", "dst"]),
        "rnb_contract_desc": (lib.ContractDesc, ["dtype", "accumulate", "m", "a", "a_major", "b", "c", "c_nblocks", "n_out", "n_pairs", "flags", "pair_a",
                                                 "out_index", "workspace", "workspace_bytes", "c_peers", "n_peers", "blocks_per_peer"]),
        "rnb_contract_nd_desc": (lib.ContractNdDesc, ["dtype", "rank_k", "n_out", "n_pairs", "ext_m", "a_stride_m", "ext_k", "b_stride_k", "a", "c",
                                                      "c_block_stride", "pair_a", "out_index", "tab_m", "tab_kb"]),
        "rnb_operand_planes": (lib.OperandPlanes, ["mode", "hi_offset", "lo_offset"]),
        "rnb_device_props": (lib.DeviceProps, ["sm_count", "total_mem_bytes", "l2_bytes_mb"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rosnet_b200.h"', "int main(void) {"]
    for cname, (_, fields) in checks.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_probe"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, (ctype, fields) in checks.items():
        assert int(out[cname]) == ctypes.sizeof(ctype), cname
        for f in fields:
            assert int(out[f"{cname}.{f}"]) == getattr(ctype, f).offset, f"{cname}.{f}"
