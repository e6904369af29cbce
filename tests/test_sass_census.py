"""The Blackwell-native claim, checkable without a GPU: disassemble the built library (cuobjdump -sass) and require the tcgen05 / TMEM /
bulk-copy mnemonics in exactly the kernels DESIGN.md §4 says run on the tensor cores, and none in the FP32 twins.  `profiles/r02_sass_census.md`
is the committed table of the same counts (scripts/sass_census.py)."""
import collections
import re
import shutil
import subprocess

import pytest

from sttode_b200 import build as B

TC_KERNELS = ["k_dec_pair_tc", "k_dec_row_tc", "k_ode_tail_tc", "k_qkv_tc", "k_gemm_tc"]
FP32_TWINS = ["k_dec_pair", "k_dec_row", "k_ode_tail", "k_proj", "k_prologue", "k_geo_attn", "k_gemm"]


@pytest.fixture(scope="module")
def census():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    path = B.build(force=False)
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    mangled = sorted(set(re.findall(r"^\s*Function : (\S+)", out, flags=re.M)))
    plain = subprocess.run(["c++filt", "-p"], input="\n".join(mangled), capture_output=True, text=True, check=True).stdout.splitlines()
    short = {m: p.split("(")[0].split("<")[0].replace("sttode::", "") for m, p in zip(mangled, plain)}
    counts, cur, arch = {}, None, set()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = counts.setdefault(short[m.group(1)], collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", ln)
        if m:
            arch.add(m.group(1))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return counts, arch


def test_library_is_sm100a_only(census):
    _, arch = census
    assert arch == {"sm_100a"}, arch


@pytest.mark.parametrize("name", TC_KERNELS)
def test_tensor_core_kernels_issue_tcgen05(census, name):
    counts, _ = census
    c = counts[name]
    assert c["UTCHMMA"] > 0, f"{name}: no tcgen05.mma in SASS"
    assert c["LDTM"] > 0, f"{name}: accumulators are not read back from tensor memory (tcgen05.ld)"
    assert c["UTCBAR"] > 0 and c["SYNCS"] > 0, f"{name}: no tcgen05.commit / mbarrier"
    assert c["HMMA"] == 0 and c["IMMA"] == 0, f"{name}: legacy mma.sync found"


@pytest.mark.parametrize("name", ["k_dec_pair_tc", "k_dec_row_tc", "k_ode_tail_tc", "k_qkv_tc"])
def test_weight_stream_uses_bulk_copies(census, name):
    counts, _ = census
    assert counts[name]["UBLKCP"] > 0, f"{name}: weight images are not streamed with cp.async.bulk"


def test_pair_kernel_converts_hidden_layers_in_tensor_memory(census):
    # relu(D1 + rowpre) / relu(D2 + b1) are written back with tcgen05.st (STTM) and consumed as a TMEM A operand
    counts, _ = census
    assert counts["k_dec_pair_tc"]["STTM"] > 0


@pytest.mark.parametrize("name", FP32_TWINS)
def test_fp32_twins_are_ffma(census, name):
    counts, _ = census
    c = counts[name]
    assert c["UTCHMMA"] == 0 and c["HMMA"] == 0 and c["FFMA"] > 0, name


def test_committed_census_matches_the_library(census):
    """profiles/r02_sass_census.md is the table of this library, not of an older build: the UTCHMMA column must agree."""
    import os
    counts, _ = census
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_sass_census.md")
    rows = {}
    with open(path) as f:
        for ln in f:
            m = re.match(r"\| `([^`<]+)(?:<[^`]*)?` \| (\d+) \| (\d*) \|", ln)
            if m:
                rows.setdefault(m.group(1), int(m.group(3) or 0))
    for name in TC_KERNELS:
        assert rows[name] == counts[name]["UTCHMMA"], (name, rows[name], counts[name]["UTCHMMA"])
