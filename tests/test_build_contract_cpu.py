"""CPU: what the compiled library must look like for the numbers in DESIGN.md 6 to hold -- read from libfnx.so with cuobjdump, no GPU:
the blend kernels fit 8 CTAs of 128 threads on an SM (<= 64 registers, no local memory, <= 28 KB of shared memory each), stage their
record spans with bulk async copies completed on mbarriers (UBLKCP / SYNCS in SASS), the backward reduces into global memory with RED
(no atomics with return values), and nothing on this path uses tensor cores (there is no dense contraction to put there)."""
import re
import subprocess

import pytest

from fluidnexus_b200 import _lib


@pytest.fixture(scope="module")
def usage(libfnx):
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    out = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt):
        out[m.group(1)] = dict(zip(("reg", "stack", "shared", "local"), map(int, m.groups()[1:])))
    assert len(out) > 40, "cuobjdump found no kernels"
    return out


@pytest.fixture(scope="module")
def sass(libfnx):
    txt = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = f.split("\n", 1)
        out[name.strip()] = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    return out


def _kernels(d, needle):
    ks = {k: v for k, v in d.items() if needle in k}
    assert ks, needle
    return ks


def test_blend_kernels_fit_eight_ctas_per_sm(usage):
    for name, u in {**_kernels(usage, "blend_fwd_kernel"), **_kernels(usage, "blend_bwd_kernel")}.items():
        assert u["reg"] <= 64, (name, u)                   # 8 CTAs x 128 threads x 64 registers = the SM's 64 K registers
        assert u["local"] == 0 and u["stack"] <= 16, (name, u)
        assert 8 * u["shared"] <= 227 * 1024, (name, u)
    for name, u in usage.items():                          # no kernel of the library spills to local memory
        if "3fnx" in name:
            assert u["local"] == 0, (name, u)


def test_blend_kernels_stage_spans_with_bulk_copies_on_mbarriers(sass):
    for name, ins in {**_kernels(sass, "blend_fwd_kernel"), **_kernels(sass, "blend_bwd_kernel")}.items():
        assert any(i.startswith("UBLKCP") for i in ins), name                     # cp.async.bulk global -> shared
        assert any(i.startswith("SYNCS.ARRIVE.TRANS64") for i in ins) and any(i.startswith("SYNCS.PHASECHK") for i in ins), name
        assert any(i.startswith("MUFU.EX2") for i in ins), name
    for name, ins in _kernels(sass, "blend_bwd_kernel").items():
        assert any(i.startswith("REDG") for i in ins) and not any(i.startswith("ATOMG") for i in ins), name
        assert sum(i.startswith("SHFL") for i in ins) >= 12, name                 # the value-splitting butterfly


def test_no_tensor_core_or_legacy_mma_instructions_anywhere(sass):
    for name, ins in sass.items():
        bad = [i for i in ins if i.startswith(("HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG"))]
        assert not bad, (name, bad[:3])
