"""CPU side of the C ABI (include/ahsoka_b200.h): the library loads without a GPU, exports every declared
symbol, the ctypes mirrors have the C layout, and — with no CUDA device — every compute entry point fails
loudly with AHS_ERR_CUDA instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from ahsoka_b200 import api, synth

HEADER = os.path.join(api.ROOT, "include", "ahsoka_b200.h")


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ahs_[a-z_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = api.load_library()
    names = _declared_functions()
    assert {"ahs_phase_batch", "ahs_phase_batch_multi", "ahs_phase_batch_resident", "ahs_free_out", "ahs_get_limits",
            "ahs_device_count", "ahs_last_error", "ahs_chain_cost", "ahs_plan_shares", "ahs_abi_version", "ahs_pin_host", "ahs_unpin_host", "ahs_warmup"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_struct_layout_matches_the_header():
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "ahsoka_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %d\\n",' \
          'sizeof(ahs_batch_in),sizeof(ahs_batch_out),sizeof(ahs_limits),offsetof(ahs_batch_out,n_cells),offsetof(ahs_batch_out,bytes_project),AHS_ABI_VERSION);return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.dirname(HEADER), os.path.join(td, "t.c"), "-o", os.path.join(td, "t")], check=True)
        out = subprocess.run([os.path.join(td, "t")], stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    s_in, s_out, s_lim, o_cells, o_bytes, ver = (int(x) for x in out)
    assert C.sizeof(api.BatchIn) == s_in and C.sizeof(api.BatchOut) == s_out and C.sizeof(api.Limits) == s_lim
    assert api.BatchOut.n_cells.offset == o_cells and api.BatchOut.bytes_project.offset == o_bytes
    assert api.load_library().ahs_abi_version() == ver


def test_limits_and_cost_model():
    lim = api.limits()
    assert lim.max_ploidy == 6 and lim.max_clusters_position == 128 and lim.max_alleles == 15 and lim.max_positions == 32767 and lim.max_reads_cluster >= 128
    lib = api.load_library()
    assert lib.ahs_chain_cost(40, 75, 2500, 2) < lib.ahs_chain_cost(400, 750, 25000, 2)


def test_no_cpu_fallback_without_a_device():
    lib = api.load_library()
    if lib.ahs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    b = synth.generate(synth.params(2, 4, 1, 8, depth=10.0, seed=3))
    with pytest.raises(RuntimeError, match=r"failed \(2\)"):          # AHS_ERR_CUDA
        api.phase_batch(b)
    with pytest.raises(RuntimeError, match=r"failed \(2\)"):
        api.phase_batch(b, resident_iters=1)
