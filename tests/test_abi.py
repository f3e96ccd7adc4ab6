"""The drop-in boundary: the CUDA library loads, exports exactly what include/itsb.h declares, and has no CPU path."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "itsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(itsb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from itsim_b200 import _lib
    assert declared_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB)
    for name in declared_functions():
        assert hasattr(lib, name), name
    from itsim_b200 import _lib
    assert lib.itsb_abi_version() == _lib.ABI_VERSION == 5


def test_hostemu_exports_the_same_surface(built):
    lib = ctypes.CDLL(built.HOSTEMU)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_a_device(built):
    """On a box without a GPU the product fails loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from itsim_b200 import _lib, netsimple
    sim = netsimple.build()
    with pytest.raises(_lib.EngineError) as err:
        sim.run_replicas(1, seed=0, duration=1.0)
    assert err.value.code == -1


def test_missing_library_is_an_import_error(tmp_path):
    from itsim_b200 import _lib
    with pytest.raises(ImportError):
        _lib.load(str(tmp_path / "nope.so"))


def test_product_never_touches_the_oracle_or_hostemu():
    """Nothing under itsim_b200/ may import or open oracle/ or the host emulation."""
    for root, _, files in os.walk(os.path.join(ROOT, "itsim_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, name)).read()
                code = re.sub(r'""".*?"""', "", text, flags=re.S)
                code = re.sub(r"/\*.*?\*/", "", code, flags=re.S)
                code = "\n".join(line.split("#")[0] for line in code.splitlines())
                assert "hostemu" not in code, name
                assert not re.search(r"^\s*(from|import)\s+(oracle|evtrace|philox|greensim)\b", code, flags=re.M), name


def _integration_stub(lib_path):
    """Executes the ctypes stub printed in INTEGRATION.md section 1, bound to `lib_path`."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"## 1\. The ctypes stub.*?```python\n(.*?)```", text, flags=re.S).group(1)
    assert 'C.CDLL("libitsim_b200.so")' in code
    ns = {}
    exec(compile(code.replace('C.CDLL("libitsim_b200.so")', f"C.CDLL({lib_path!r})"), "INTEGRATION.md", "exec"), ns)
    return ns


@pytest.mark.parametrize("backend", [pytest.param("hostemu", id="hostemu"),
                                     pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)])
def test_the_stub_of_integration_md_runs_a_batch(built, backend):
    """The binding a maintainer of the reference would add, executed as printed: same events
    and telemetry as the package's own binding, and its structs have the header's layout."""
    import numpy as np
    from itsim_b200 import _lib, netsimple
    path = built.HOSTEMU if backend == "hostemu" else built.LIB
    stub = _integration_stub(path)
    sim = netsimple.build()
    compiled = sim.compile()
    own = compiled.desc()
    assert ctypes.sizeof(stub["ModelDesc"]) == ctypes.sizeof(own)
    desc = stub["ModelDesc"].from_buffer_copy(bytes(own))
    events, telemetry = stub["run_batch"](compiled.blobs(), desc, 3, 11, 700.0)
    sim._lib = _lib.bind(ctypes.CDLL(path)) if backend == "hostemu" else None
    eng = sim.run_replicas(3, seed=11)
    assert eng.run(700.0) == events > 1000
    assert np.array_equal(eng.telemetry(), telemetry)
    assert len(telemetry) == _lib.TM_SIZE
